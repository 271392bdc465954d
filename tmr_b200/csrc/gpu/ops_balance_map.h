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
  int nblocks;
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

/* every input octant marks its parent; level-0 octants flag their tree */
struct MapMarkParentsFn {
  const u64 *keys;
  KeyFmt fmt;
  CellMaps mp;
  int *root_flag;
  TMR_HD i64 parent_cell(u64 k, int *L) const {
    *L = (int)(k & 31);
    if (*L == 0) return -1;
    const u64 rest = k >> 5;
    const u64 block = rest >> (3 * fmt.D);
    const u64 m = rest & low_mask(3 * fmt.D);
    return (i64)((block << (3 * (*L - 1))) | (m >> (3 * (fmt.D - *L + 1))));
  }
  TMR_HD void operator()(i64 i) const {
    int L, Lp;
    const i64 pc = parent_cell(keys[i], &L);
    if (pc < 0) {
      root_flag[(int)((keys[i] >> 5) >> (3 * fmt.D))] = 1;
      return;
    }
    if (i > 0 && parent_cell(keys[i - 1], &Lp) == pc && Lp == L) return;
    TMR_ATOMIC_OR_I32(&mp.bits[mp.woff[L - 1] + (pc >> 5)], 1u << (int)(pc & 31));
  }
};

struct MapOrEmit {
  u32 *words; /* level l-1 */
  int sh;     /* 3 (l-1) */
  TMR_HD void operator()(i32 block, i32 x, i32 y, i32 z) {
    const u64 c = ((u64)(u32)block << sh) | morton3((u32)x, (u32)y, (u32)z);
    TMR_ATOMIC_OR_I32(&words[c >> 5], 1u << (int)(c & 31));
  }
};

/* closure step l -> l-1, one thread per sibling group of level l */
struct MapClosureFn {
  CellMaps mp;
  ConnTables t;
  int l;
  int corner;
  TMR_HD void operator()(i64 g) const {
    const u32 byte = mp.byte_of(l, g);
    if (!byte) return;
    /* g is the cell index of the group's parent at level l-1 */
    const int shp = 3 * (l - 1);
    const i32 block = (i32)(g >> shp);
    u32 qx, qy, qz;
    unmorton3((u64)g & low_mask(shp), &qx, &qy, &qz);
    const i32 N = 1 << (l - 1);
    MapOrEmit emit = {mp.bits + mp.woff[l - 1], shp};
    /* the members' own parent */
    emit(block, (i32)qx, (i32)qy, (i32)qz);
    for (int d = 0; d < 8; d++) {
      if (!((byte >> d) & 1u)) continue;
      /* x-major digit d = 4 xbit + 2 ybit + zbit: the side of the parent the
         member sits on */
      const i32 s[3] = {(d & 4) ? 1 : -1, (d & 2) ? 1 : -1, (d & 1) ? 1 : -1};
      for (int a = 1; a < 8; a++) {
        if (a == 7 && !corner) continue;
        i32 q[3] = {(i32)qx + ((a & 1) ? s[0] : 0), (i32)qy + ((a & 2) ? s[1] : 0),
                    (i32)qz + ((a & 4) ? s[2] : 0)};
        tree_images(t, block, q, N, emit);
      }
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
  TMR_HD u32 operator()(i64 b) const {
    if (mp.test(0, b)) return cnt[mp.rank_of(0, b)];
    return root_flag[b] ? 1u : 0u;
  }
};

/* top-down: start of the leaves of every refined cell; children that are not
   refined are leaves and are written in place */
struct MapFillFn {
  CellMaps mp;
  int l;
  const u32 *cnt;
  u32 *off;          /* by dense index (levels >= 1) */
  const u32 *toff;   /* per tree (level 0) */
  const int *root_flag;
  KeyFmt fmt;
  u64 *out;
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
        } else if (root_flag[b]) {
          out[toff[b]] = (u64)b << (3 * fmt.D + 5);
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
      const u64 block = (u64)cell >> (3 * l);
      const u64 m = (u64)cell & low_mask(3 * l);
      for (int e = 0; e < 8; e++) {
        const i64 child = (cell << 3) | e;
        if ((kids >> e) & 1u) {
          const u32 r = mp.rank_of(L, child);
          off[r] = at;
          at += cnt[r];
        } else {
          const u64 mD = ((m << 3) | (u64)e) << (3 * (fmt.D - L));
          out[at++] = (block << (3 * fmt.D + 5)) | (mD << 5) | (u64)L;
        }
      }
    }
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
  DBuf<int> root_flag(ctx, f.nblocks);
  dev_zero(ctx, root_flag.get(), (size_t)f.nblocks * sizeof(int));
  MapMarkParentsFn mk = {f.keys.get(), f.fmt, mp, root_flag.get()};
  launch(ctx, f.n, mk, "balance_map_mark");
  for (int l = D - 1; l >= 1; l--) {
    MapClosureFn cl = {mp, f.tables, l, balance_corner};
    launch(ctx, mp.cells(l - 1), cl, "balance_map_closure");
  }
  trace_mark(ctx, "balance: closure");
  MapWordPopFn wp = {bits.get()};
  const i64 nref = (i64)scan_counts(ctx, words, wp, wrank.get(), "balance_map_rank");
  DBuf<u32> cnt(ctx, nref), off(ctx, nref), toff(ctx, f.nblocks);
  for (int l = D - 1; l >= 0; l--) {
    MapCountFn cf = {mp, l, cnt.get()};
    launch(ctx, l == 0 ? ((i64)f.nblocks + 7) / 8 : mp.cells(l - 1), cf,
           "balance_map_count");
  }
  MapTreeCountFn tc = {mp, cnt.get(), root_flag.get()};
  const i64 total = (i64)scan_counts(ctx, f.nblocks, tc, toff.get(), "balance_map_trees");
  if (total >= (1LL << 31)) {
    fprintf(stderr,
            "TMROctForest Error: balance() would create %lld octants on one "
            "rank (int32 index limit of the TMROctForest API)\n",
            (long long)total);
    return 1;
  }
  DBuf<u64> out(ctx, total);
  for (int l = 0; l < D; l++) {
    MapFillFn ff = {mp, l, cnt.get(), off.get(), toff.get(), root_flag.get(), f.fmt,
                    out.get()};
    launch(ctx, l == 0 ? ((i64)f.nblocks + 7) / 8 : mp.cells(l - 1), ff,
           "balance_map_fill");
  }
  /* a failed allocation above launched nothing: leave the forest as it was */
  if (!ctx_ok(ctx)) return check_errors(ctx, "balance");
  f.keys.swap(out);
  f.n = total;
  f.last_out = f.n;
  trace_mark(ctx, "balance: leaves");
  return check_errors(ctx, "balance");
}

}  // namespace tmrgpu

#endif
