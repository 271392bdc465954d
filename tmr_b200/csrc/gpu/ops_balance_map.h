/*
  ops_balance_map.h -- 2:1 balance on per-level cell bitmaps (single rank).

  Same closure as ops_balance.h (reference src/TMROctForest.cpp:2917-3089,
  balanceOctant :2763-2895, add{Face,Edge,Corner}Neighbors :2525-2745): R_l is
  the set of level-l cells that must be refined, p in R_l puts its parent and
  the parent's neighbours on p's side into R_{l-1}, leaves are the children of
  R_l members that are not in R_{l+1}.  ops_balance.h keeps every R_l as a
  sorted key array and pays a radix sort + dedup per level plus a final sort
  of all leaves (measured on the 86 M-octant cycle: 9.3 of 37 ms in 70
  launches).  Here R_l is ONE BIT per level-l cell:

    mark      every input octant sets the bit of its parent
    closure   level D-1 down to 1: every set bit ORs its (up to 8, more across
              tree edges / corners) demands into the level above -- set union
              is an atomicOr, no sort, no dedup
    rank      one scan over the words of all levels gives every refined cell
              a dense index
    count     bottom-up: leaves under every refined cell
    fill      top-down: every refined cell knows where its leaves start, so
              the leaves are written in Morton order -- no final sort

  A sibling group (the 8 children of one cell) is one byte of the map, so a
  thread works on one byte: its children's refined mask is one load.

  The maps hold nblocks * 8^l bits per level; forests whose maps exceed the
  budget (deep or very wide) use the sorted-array closure of ops_balance.h.
*/
#ifndef TMRGPU_OPS_BALANCE_MAP_H
#define TMRGPU_OPS_BALANCE_MAP_H

#include "ops_balance.h"

namespace tmrgpu {

static const int kMaxMapLevels = 20;

struct CellMaps {
  u32 *bits;              /* all levels, level l at word woff[l] */
  const u32 *wrank;       /* refined cells before each word (all levels) */
  i64 woff[kMaxMapLevels + 1];
  int D;                  /* levels 0 .. D-1 */
  int nblocks;            /* trees covered: block0 .. block0 + nblocks - 1 */
  int block0;             /* (several ranks: the trees of this rank's range) */
  TMR_HD bool has_block(i64 b) const { return b >= block0 && b < block0 + nblocks; }
  /* local cell index of global cell (block, Morton m) at level l */
  TMR_HD i64 cell_of(i64 block, u64 m, int l) const {
    return ((block - block0) << (3 * l)) | (i64)m;
  }
  TMR_HD u32 byte_of(int l, i64 group) const {
    /* refined mask of the 8 level-l cells 8 group .. 8 group + 7 */
    const i64 bit = group << 3;
    return (bits[woff[l] + (bit >> 5)] >> (int)(bit & 31)) & 0xffu;
  }
  /* dense index of level-l cell `cell` (which must be refined) */
  TMR_HD u32 rank_of(int l, i64 cell) const {
    const i64 w = woff[l] + (cell >> 5);
    return wrank[w] + (u32)popc32(bits[w] & ((1u << (int)(cell & 31)) - 1u));
  }
  TMR_HD bool test(int l, i64 cell) const {
    return (bits[woff[l] + (cell >> 5)] >> (int)(cell & 31)) & 1u;
  }
  TMR_HD i64 cells(int l) const { return (i64)nblocks << (3 * l); }
};

/* several ranks: a cell of a tree that other ranks hold too (a tree cut by a
   partition boundary, or a tree outside this rank's span) is forwarded to
   them as (level << 58 | global cell), one entry per destination */
struct MapRemote {
  const i32 *span_b0; /* per rank: first / last tree of its span (b0 > b1: none) */
  const i32 *span_b1;
  int R, me;
  int lo_shared, hi_shared; /* my first / last tree is in another rank's span */
  u64 *key;
  u32 *dest;
  unsigned long long *count;
  i64 cap;
  TMR_HD bool on() const { return R > 1; }
  TMR_HD bool needs_forward(const CellMaps &mp, i64 block) const {
    if (!mp.has_block(block)) return true;
    return (block == mp.block0 && lo_shared) ||
           (block == mp.block0 + mp.nblocks - 1 && hi_shared);
  }
  TMR_HD void forward(i64 block, u64 m, int l) const {
    const u64 k = ((u64)l << 58) | ((u64)block << (3 * l)) | m;
    for (int r = 0; r < R; r++) {
      if (r == me || block < span_b0[r] || block > span_b1[r]) continue;
      const unsigned long long s = fetch_add_u64(count, 1ULL);
      if ((i64)s < cap) {
        key[s] = k;
        dest[s] = (u32)r;
      }
    }
  }
};

/* every input octant marks its parent; level-0 octants flag their tree */
struct MapMarkParentsFn {
  const u64 *keys;
  KeyFmt fmt;
  CellMaps mp;
  int *root_flag; /* by global tree */
  MapRemote rm;
  /* parent as (tree, Morton at level L-1); false for level-0 octants */
  TMR_HD bool parent_cell(u64 k, int *L, i64 *block, u64 *m) const {
    *L = (int)(k & 31);
    if (*L == 0) return false;
    const u64 rest = k >> 5;
    *block = (i64)(rest >> (3 * fmt.D));
    *m = (rest & low_mask(3 * fmt.D)) >> (3 * (fmt.D - *L + 1));
    return true;
  }
  TMR_HD void operator()(i64 i) const {
    int L, Lp;
    i64 b, bp;
    u64 m, mq;
    if (!parent_cell(keys[i], &L, &b, &m)) {
      root_flag[(int)((keys[i] >> 5) >> (3 * fmt.D))] = 1;
      return;
    }
    if (i > 0 && parent_cell(keys[i - 1], &Lp, &bp, &mq) && Lp == L && bp == b &&
        mq == m) {
      return;
    }
    const i64 pc = mp.cell_of(b, m, L - 1);
    TMR_ATOMIC_OR_I32(&mp.bits[mp.woff[L - 1] + (pc >> 5)], 1u << (int)(pc & 31));
    if (rm.on() && rm.needs_forward(mp, b)) rm.forward(b, m, L - 1);
  }
};

/* cells received from other ranks */
struct MapOrReceivedFn {
  const u64 *key;
  CellMaps mp;
  TMR_HD void operator()(i64 i) const {
    const u64 k = key[i];
    const int l = (int)(k >> 58);
    const u64 c = k & low_mask(58);
    const i64 block = (i64)(c >> (3 * l));
    if (!mp.has_block(block)) return;
    const i64 pc = mp.cell_of(block, c & low_mask(3 * l), l);
    TMR_ATOMIC_OR_I32(&mp.bits[mp.woff[l] + (pc >> 5)], 1u << (int)(pc & 31));
  }
};

struct MapOrEmit {
  const CellMaps *mp;
  const MapRemote *rm;
  int lp; /* level of the demanded cells */
  TMR_HD void operator()(i32 block, i32 x, i32 y, i32 z) {
    const u64 m = morton3((u32)x, (u32)y, (u32)z);
    if (mp->has_block(block)) {
      const i64 c = mp->cell_of(block, m, lp);
      TMR_ATOMIC_OR_I32(&mp->bits[mp->woff[lp] + (c >> 5)], 1u << (int)(c & 31));
    }
    if (rm->on() && rm->needs_forward(*mp, block)) rm->forward(block, m, lp);
  }
};

/* which of the 27 cells around the parent (offset (ox,oy,oz) in {-1,0,1}^3,
   bit (ox+1) + 3 (oy+1) + 9 (oz+1)) a refined child with x-major digit d
   demands: the parent itself and its neighbours on the child's side across
   faces and edges (and the corner when corners are balanced) */
TMR_HD u32 child_demand_mask(int d, int corner) {
  const int s[3] = {(d & 4) ? 1 : -1, (d & 2) ? 1 : -1, (d & 1) ? 1 : -1};
  u32 need = 0;
  TMR_UNROLL
  for (int a = 0; a < 8; a++) {
    if (a == 7 && !corner) continue;
    const int ox = (a & 1) ? s[0] : 0, oy = (a & 2) ? s[1] : 0, oz = (a & 4) ? s[2] : 0;
    need |= 1u << ((ox + 1) + 3 * (oy + 1) + 9 * (oz + 1));
  }
  return need;
}

/* closure step l -> l-1, one thread per sibling group of level l.  The
   demands of the group's refined members are united first (8 refined siblings
   ask for 19 or 27 distinct cells, not 8 x 7), and a parent that does not
   touch its tree's boundary finds its neighbours by dilated +-1 on the three
   axis components of its Morton code. */
struct MapClosureFn {
  CellMaps mp;
  ConnTables t;
  int l;
  int corner;
  MapRemote rm;
  TMR_HD void operator()(i64 g) const {
    const u32 byte = mp.byte_of(l, g);
    if (!byte) return;
    u32 need = 0;
    TMR_UNROLL
    for (int d = 0; d < 8; d++) {
      if ((byte >> d) & 1u) need |= child_demand_mask(d, corner);
    }
    /* g is the cell index of the group's parent at level l-1 */
    const int shp = 3 * (l - 1);
    const u64 mq = (u64)g & low_mask(shp);
    u32 *words = mp.bits + mp.woff[l - 1];
    /* per axis: the component one step down / in place / one step up, and
       whether the parent touches the low / high tree face */
    u64 comp[3][3];
    bool lo = false, hi = false;
    TMR_UNROLL
    for (int a = 0; a < 3; a++) { /* a = 0 x, 1 y, 2 z; x is Morton bit 2 */
      const u64 am = (0x1249249249249249ULL << (2 - a)) & low_mask(shp);
      const u64 c = mq & am;
      lo = lo || c == 0;
      hi = hi || c == am;
      comp[a][0] = (c - 1) & am;
      comp[a][1] = c;
      comp[a][2] = ((c | ~am) + 1) & am;
    }
    const i32 block = (i32)(g >> shp) + mp.block0;
    if (!lo && !hi) {
      /* all demands stay inside this tree */
      const u64 base = ((u64)g >> shp) << shp; /* local tree bits */
      const bool fwd = rm.on() && rm.needs_forward(mp, block);
      while (need) {
        const int o = ctz32(need);
        need &= need - 1;
        const int oz = o / 9, oy = (o - 9 * oz) / 3, ox = o - 9 * oz - 3 * oy;
        const u64 mc = comp[0][ox] | comp[1][oy] | comp[2][oz];
        const u64 c = base | mc;
        /* neighbouring groups demand the same cells over and over: look before
           the reduction (a cached load) -- most bits are set already */
        const u32 bit = 1u << (int)(c & 31);
        if (!(words[c >> 5] & bit)) TMR_ATOMIC_OR_I32(&words[c >> 5], bit);
        if (fwd) rm.forward(block, mc, l - 1);
      }
      return;
    }
    u32 qx, qy, qz;
    unmorton3(mq, &qx, &qy, &qz);
    const i32 N = 1 << (l - 1);
    MapOrEmit emit = {&mp, &rm, l - 1};
    while (need) {
      const int o = ctz32(need);
      need &= need - 1;
      const int oz = o / 9, oy = (o - 9 * oz) / 3, ox = o - 9 * oz - 3 * oy;
      const i32 q[3] = {(i32)qx + ox - 1, (i32)qy + oy - 1, (i32)qz + oz - 1};
      tree_images(t, block, q, N, emit);
    }
  }
};

struct MapWordPopFn {
  const u32 *bits;
  TMR_HD u32 operator()(i64 w) const { return (u32)popc32(bits[w]); }
};

/* leaves under every refined cell of level l (bottom-up), one thread per
   sibling group of level l */
struct MapCountFn {
  CellMaps mp;
  int l;
  u32 *cnt; /* by dense index */
  TMR_HD void operator()(i64 g) const {
    const u32 byte = mp.byte_of(l, g);
    if (!byte) return;
    for (int d = 0; d < 8; d++) {
      if (!((byte >> d) & 1u)) continue;
      const i64 cell = (g << 3) | d;
      u32 c = 8;
      if (l + 1 < mp.D) {
        const u32 kids = mp.byte_of(l + 1, cell);
        c = 8u - (u32)popc32(kids);
        for (int e = 0; e < 8; e++) {
          if ((kids >> e) & 1u) c += cnt[mp.rank_of(l + 1, (cell << 3) | e)];
        }
      }
      cnt[mp.rank_of(l, cell)] = c;
    }
  }
};

/* leaves of every tree: refined root -> its count, unrefined root present in
   the input -> 1 */
struct MapTreeCountFn {
  CellMaps mp;
  const u32 *cnt;
  const int *root_flag;
  TMR_HD u32 operator()(i64 b) const { /* b: tree of the map (local index) */
    if (mp.test(0, b)) return cnt[mp.rank_of(0, b)];
    return root_flag[b + mp.block0] ? 1u : 0u;
  }
};

/* top-down: start of the leaves of every refined cell; children that are not
   refined are leaves and are written in place.  Several ranks: the maps span
   whole trees, the rank keeps the leaves [first, last) of that sequence (its
   own range of positions). */
struct MapFillFn {
  CellMaps mp;
  int l;
  const u32 *cnt;
  u32 *off;          /* by dense index (levels >= 1) */
  const u32 *toff;   /* per tree of the map (level 0) */
  const int *root_flag;
  KeyFmt fmt;
  u64 *out;
  u32 first, last;
  TMR_HD void put(u32 at, u64 key) const {
    if (at >= first && at < last) out[at - first] = key;
  }
  TMR_HD void operator()(i64 g) const {
    u32 byte;
    if (l == 0) {
      /* level 0: cells are trees, also write the unrefined roots */
      byte = 0;
      for (int d = 0; d < 8; d++) {
        const i64 b = (g << 3) | d;
        if (b >= mp.nblocks) break;
        if (mp.test(0, b)) {
          byte |= 1u << d;
        } else if (root_flag[b + mp.block0]) {
          put(toff[b], (u64)(b + mp.block0) << (3 * fmt.D + 5));
        }
      }
    } else {
      byte = mp.byte_of(l, g);
    }
    if (!byte) return;
    const int L = l + 1; /* level of the children */
    for (int d = 0; d < 8; d++) {
      if (!((byte >> d) & 1u)) continue;
      const i64 cell = (g << 3) | d;
      u32 at = (l == 0) ? toff[cell] : off[mp.rank_of(l, cell)];
      const u32 kids = (L < mp.D) ? mp.byte_of(L, cell) : 0u;
      const u64 block = ((u64)cell >> (3 * l)) + (u64)mp.block0;
      const u64 m = (u64)cell & low_mask(3 * l);
      for (int e = 0; e < 8; e++) {
        const i64 child = (cell << 3) | e;
        if ((kids >> e) & 1u) {
          const u32 r = mp.rank_of(L, child);
          off[r] = at;
          at += cnt[r];
        } else {
          const u64 mD = ((m << 3) | (u64)e) << (3 * (fmt.D - L));
          put(at++, (block << (3 * fmt.D + 5)) | (mD << 5) | (u64)L);
        }
      }
    }
  }
#if defined(__CUDACC__)
  /* warp form (launch_warp; ncu: 4-13 of 32 lanes active per item): the
     refined cells of the warp's 32 groups are taken four at a time, eight
     lanes per cell -- one lane per child, the children's leaf counts scanned
     within the eight lanes */
  __device__ __forceinline__ void warp(i64 base, int lane, i64 n) const {
    const i64 g = base + lane;
    if (l == 0) { /* trees: tiny, per item */
      if (g < n) (*this)(g);
      return;
    }
    const u32 byte = (g < n) ? mp.byte_of(l, g) : 0u;
    const int L = l + 1;
    const int q = lane >> 3, e = lane & 7;
    for (int d = 0; d < 8; d++) {
      unsigned mask = __ballot_sync(0xffffffffu, (byte >> d) & 1u);
      while (mask) {
        const int left = __popc(mask);
        const int nsrc = left < 4 ? left : 4;
        const bool on = q < nsrc;
        u32 size = 0, r = 0, at0 = 0;
        bool refined = false;
        i64 cell = 0;
        if (on) {
          const int src = (int)__fns(mask, 0, q + 1);
          cell = ((base + src) << 3) | d;
          at0 = off[mp.rank_of(l, cell)];
          const u32 kids = (L < mp.D) ? mp.byte_of(L, cell) : 0u;
          refined = ((kids >> e) & 1u) != 0;
          if (refined) {
            r = mp.rank_of(L, (cell << 3) | e);
            size = cnt[r];
          } else {
            size = 1;
          }
        }
        u32 incl = size;
#pragma unroll
        for (int s = 1; s < 8; s <<= 1) {
          const u32 up = __shfl_up_sync(0xffffffffu, incl, s, 8);
          if (e >= s) incl += up;
        }
        if (on) {
          const u32 at = at0 + incl - size;
          if (refined) {
            off[r] = at;
          } else {
            const u64 block = ((u64)cell >> (3 * l)) + (u64)mp.block0;
            const u64 m = (u64)cell & low_mask(3 * l);
            const u64 mD = ((m << 3) | (u64)e) << (3 * (fmt.D - L));
            put(at, (block << (3 * fmt.D + 5)) | (mD << 5) | (u64)L);
          }
        }
        for (int k = 0; k < nsrc; k++) mask &= mask - 1;
      }
    }
  }
#endif
};

/* number of leaves of the map's trees that lie before position `pos` (depth
   D, global), for pos on a leaf boundary: walks one root-to-leaf path */
struct MapLeavesBeforeFn {
  CellMaps mp;
  const u32 *cnt;
  const u32 *toff;
  u32 total;
  int D;
  const u64 *pos; /* [nq] */
  u32 *out;
  TMR_HD void operator()(i64 q) const {
    const u64 p = pos[q];
    const i64 block = (i64)(p >> (3 * D));
    if (block < mp.block0) {
      out[q] = 0;
      return;
    }
    if (block >= mp.block0 + mp.nblocks) {
      out[q] = total;
      return;
    }
    const u64 m = p & low_mask(3 * D);
    u32 n = toff[block - mp.block0];
    i64 cell = block - mp.block0;
    for (int l = 0; l < D; l++) {
      if (!mp.test(l, cell)) break; /* a leaf anchored at or before pos */
      const int e = (int)((m >> (3 * (D - l - 1))) & 7);
      const u32 kids = (l + 1 < mp.D) ? mp.byte_of(l + 1, cell) : 0u;
      for (int k = 0; k < e; k++) {
        n += ((kids >> k) & 1u) ? cnt[mp.rank_of(l + 1, (cell << 3) | k)] : 1u;
      }
      cell = (cell << 3) | e;
      if (!((kids >> e) & 1u)) break;
    }
    out[q] = n;
  }
};

/* words needed by the maps of a forest of depth D, or -1 when over budget */
inline i64 balance_map_words(int nblocks, int D, i64 budget_words, i64 *woff) {
  i64 w = 0;
  for (int l = 0; l < D; l++) {
    woff[l] = w;
    if (3 * l >= 40) return -1;
    const i64 cells = (i64)nblocks << (3 * l);
    w += ((cells + 31) >> 5) + 1;
    if (w > budget_words) return -1;
  }
  woff[D] = w;
  return w;
}

/* The shared back half of both drivers: rank the refined cells, count leaves
   bottom-up, place them top-down.  pos_lo/pos_hi: keep the leaves of that
   range of positions (several ranks), or NULL for all. */
inline int balance_map_leaves(Forest &f, CellMaps &mp, i64 words, DBuf<u32> &wrank,
                              const int *root_flag, const u64 *range /* [2] or NULL */) {
  Ctx &ctx = *f.ctx;
  const int D = mp.D;
  MapWordPopFn wp = {mp.bits};
  const i64 nref = (i64)scan_counts(ctx, words, wp, wrank.get(), "balance_map_rank");
  DBuf<u32> cnt(ctx, nref), off(ctx, nref), toff(ctx, mp.nblocks);
  for (int l = D - 1; l >= 0; l--) {
    MapCountFn cf = {mp, l, cnt.get()};
    launch(ctx, l == 0 ? ((i64)mp.nblocks + 7) / 8 : mp.cells(l - 1), cf,
           "balance_map_count");
  }
  MapTreeCountFn tc = {mp, cnt.get(), root_flag};
  const i64 total =
      (i64)scan_counts(ctx, mp.nblocks, tc, toff.get(), "balance_map_trees");
  if (total >= (1LL << 32) - 1) {
    fprintf(stderr, "TMROctForest Error: balance() would create %lld octants in the "
                    "trees of one rank\n", (long long)total);
    return 1;
  }
  u32 first = 0, last = (u32)total;
  if (range) {
    DBuf<u64> d_pos(ctx, 2);
    DBuf<u32> d_n(ctx, 2);
    copy_h2d(ctx, d_pos.get(), range, 2 * sizeof(u64));
    MapLeavesBeforeFn lb = {mp, cnt.get(), toff.get(), (u32)total, f.fmt.D,
                            d_pos.get(), d_n.get()};
    launch(ctx, 2, lb, "balance_map_range");
    u32 h_n[2] = {0, 0};
    copy_d2h(ctx, h_n, d_n.get(), sizeof(h_n));
    first = h_n[0];
    last = h_n[1];
  }
  const i64 mine = (i64)last - (i64)first;
  if (mine >= (1LL << 31)) {
    fprintf(stderr,
            "TMROctForest Error: balance() would create %lld octants on one "
            "rank (int32 index limit of the TMROctForest API)\n",
            (long long)mine);
    return 1;
  }
  DBuf<u64> out(ctx, mine);
  for (int l = 0; l < D; l++) {
    MapFillFn ff = {mp,       l,     cnt.get(), off.get(), toff.get(), root_flag,
                    f.fmt,    out.get(), first, last};
    launch_warp(ctx, l == 0 ? ((i64)mp.nblocks + 7) / 8 : mp.cells(l - 1), ff,
                "balance_map_fill");
  }
  /* a failed allocation above launched nothing: leave the forest as it was */
  if (!ctx_ok(ctx)) return check_errors(ctx, "balance");
  f.keys.swap(out);
  f.n = mine;
  f.last_out = f.n;
  return 0;
}

/* returns 0 ok, 1 error, -1 = not applicable (use the sorted-array closure) */
inline int balance_map(Forest &f, int balance_corner) {
  Ctx &ctx = *f.ctx;
  const int D = f.fmt.D;
  if (D < 1 || D > kMaxMapLevels) return -1;
  CellMaps mp;
  /* budget: 2^27 words = 512 MB of maps (+ the same for the word ranks) */
  i64 budget = (i64)1 << 27;
  if (const char *ev = getenv("TMR_B200_BALANCE_MAP_WORDS")) budget = atol(ev);
  const i64 words = balance_map_words(f.nblocks, D, budget, mp.woff);
  if (words < 0) return -1;
  trace_mark(ctx, NULL);
  f.last_mid = f.n;
  f.info.reset();
  DBuf<u32> bits(ctx, words), wrank(ctx, words);
  dev_zero(ctx, bits.get(), (size_t)words * sizeof(u32));
  mp.bits = bits.get();
  mp.wrank = wrank.get();
  mp.D = D;
  mp.nblocks = f.nblocks;
  mp.block0 = 0;
  DBuf<int> root_flag(ctx, f.nblocks);
  dev_zero(ctx, root_flag.get(), (size_t)f.nblocks * sizeof(int));
  MapRemote rm = {NULL, NULL, 1, 0, 0, 0, NULL, NULL, NULL, 0};
  MapMarkParentsFn mk = {f.keys.get(), f.fmt, mp, root_flag.get(), rm};
  launch(ctx, f.n, mk, "balance_map_mark");
  for (int l = D - 1; l >= 1; l--) {
    MapClosureFn cl = {mp, f.tables, l, balance_corner, rm};
    launch(ctx, mp.cells(l - 1), cl, "balance_map_closure");
  }
  trace_mark(ctx, "balance: closure");
  const int rc = balance_map_leaves(f, mp, words, wrank, root_flag.get(), NULL);
  if (rc) return rc;
  trace_mark(ctx, "balance: leaves");
  return check_errors(ctx, "balance");
}

}  // namespace tmrgpu

#endif
