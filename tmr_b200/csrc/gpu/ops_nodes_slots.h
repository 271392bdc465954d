/*
  ops_nodes_slots.h -- sort-free construction of the unique node set and the
  local connectivity of an order-2 mesh (replaces the candidate sort of
  ops_nodes.h for the common case; same results, reference createLocalNodes
  src/TMROctForest.cpp:4290-4640 and createLocalConn :4660-4867).

  Every canonical node position (after transformNode and its hmax -> hmax-1
  clamp, reference :3847-4039) lies in exactly one leaf of the Morton-sorted
  element array: the LAST element whose anchor is at or before the position.
  In a 2:1 balanced, complete forest the position relative to that leaf is,
  per axis, one of three things:
      0   the leaf's own anchor coordinate,
      1   half the leaf's size above it (a hanging node on a lower face / edge
          of the leaf, created by finer neighbours),
      2   the clamped coordinate 2^30-1 (the leaf touches the upper face of
          its tree and the node sits on that face).
  So a node is named by (leaf, slot) with 27 possible slots, the unique node
  array in the reference's compareNode order (src/TMROctant.cpp:245-274) is
  "for each leaf in order, its occupied slots in slot order", and the index of
  a node is an exclusive scan of popc(slot mask) -- no sort, no candidates in
  HBM.  Slot order: the squeezed node coordinate (NodeFmt) of slot index 0 / 1
  / 2 on a leaf of 2^k cells is A<<1, A<<1 | 1<<k, A<<1 | (2<<k)-1, so node
  keys inside a leaf compare by the 6-bit code (hx hy hz lx ly lz), h = index
  >= 1, l = index == 2, x before y before z; 27 of the 64 codes are valid and
  a slot's ordinal among them indexes a 32-bit mask.

  The connectivity needs, for each of the 8 corners of each element, the leaf
  containing the corner: one predecessor search in the L2-resident radix index
  of the element keys.  A complete family of 8 siblings shares the 27 points
  of its 3x3x3 grid; 8 of them are the siblings' own anchors, the other 19
  are located once and handed to the siblings through shared memory
  (launch_block3), 2.4 searches per element instead of 7.

  Anything the slots cannot name -- a corner at a quarter position of its
  leaf (unbalanced input), a position not covered by any leaf (incomplete
  forest after a bare refine) -- raises a flag and create_nodes falls back to
  the general candidate sort.  Positions outside this rank's own Morton range
  (several ranks) go through the candidate sort as a small side list ("B"
  nodes): they sort entirely before or after the range's own nodes.
*/
#ifndef TMRGPU_OPS_NODES_SLOTS_H
#define TMRGPU_OPS_NODES_SLOTS_H

#include "ops_route.h"

namespace tmrgpu {

static const u64 kSlotValid = 0xff5533110f050301ULL; /* codes with l <= h */
static const u64 kLocB = ~0ULL;        /* outside this rank's Morton range */
static const u64 kLocFail = ~0ULL - 1; /* not nameable: general path */
static const u32 kConnB = 0xffffffffu; /* conn entry filled by the B pass */
static const u32 kPayDep = 0x80000000u; /* B payload: the corner is a dependent node */

TMR_HD int slot_ord(int c6) { return popc64(kSlotValid & ((1ULL << c6) - 1)); }
/* ordinal -> code: the 27 valid codes, one byte each, 8 per word */
TMR_HD int slot_code(int ord) {
  const int w = ord >> 3;
  const u64 k = w == 0 ? 0x1a19181210090800ULL
                       : (w == 1 ? 0x302d2c292824201bULL
                                 : (w == 2 ? 0x3c3b3a3938363432ULL : 0x3f3e3dULL));
  return (int)((k >> (8 * (ord & 7))) & 63);
}

/* The in-tree Morton code of a position is 3 D bits: a 32-bit word up to depth
   10, which halves the instruction count of the dilated arithmetic below on
   a 32-bit machine; deeper forests use 64-bit words.  M = u32 or u64. */
template <class M>
TMR_HD M axis_mask(int a) {
  return (M)(0x1249249249249249ULL << a);
}
/* add one at bit `bit` of the dilated axis-a component of a Morton code */
template <class M>
TMR_HD M morton_axis_add(M m, int a, int bit) {
  const M am = axis_mask<M>(a);
  const M nc = (((m & am) | ~am) + ((M)1 << bit)) & am;
  return (m & ~am) | nc;
}

TMR_HD void store8_u32(u32 *p, const u32 *v) {
#if defined(__CUDA_ARCH__)
  uint4 *q = reinterpret_cast<uint4 *>(p);
  q[0] = make_uint4(v[0], v[1], v[2], v[3]);
  q[1] = make_uint4(v[4], v[5], v[6], v[7]);
#else
  for (int c = 0; c < 8; c++) p[c] = v[c];
#endif
}
TMR_HD void load8_u32(const u32 *p, u32 *v) {
#if defined(__CUDA_ARCH__)
  const uint4 *q = reinterpret_cast<const uint4 *>(p);
  const uint4 a = q[0], b = q[1];
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
#else
  for (int c = 0; c < 8; c++) v[c] = p[c];
#endif
}

/* ---- rank index: predecessor search in O(1) ----------------------------------
   One entry per bucket of 64 cells of level Dg: a 64-bit map of the cells in
   which a leaf is anchored and the number of marked cells before the bucket.
   With Dg = D every leaf has its own cell, so the last leaf at or before a
   position is `start + popc(bits up to the cell) - 1`: one 16-byte load
   instead of a table read plus a binary search (measured on the 86 M-octant
   forest: 6 dependent loads and 120 warp instructions per search; the search
   loop was where the kernel waited).  When that table would outgrow its
   budget, Dg < D, a marked cell holds several leaves, `cell_first` lists the
   first leaf of every marked cell and the leaves of one cell are bisected. */
struct RankEntry {
  u32 bits[2];
  u32 start;
  u32 pad;
};

struct RankIndex {
  const RankEntry *tab;
  const u32 *cell_first; /* NULL when Dg = D; else [marked cells + 1] */
  u64 cell0;             /* first cell (level Dg, incl. the tree index) */
  int shift;             /* 3 (D - Dg) */
  TMR_HD RankEntry load(u64 b) const {
#if defined(__CUDA_ARCH__)
    const uint4 v = *reinterpret_cast<const uint4 *>(tab + b);
    RankEntry e;
    e.bits[0] = v.x;
    e.bits[1] = v.y;
    e.start = v.z;
    e.pad = v.w;
    return e;
#else
    return tab[b];
#endif
  }
  /* index of the last leaf anchored at or before position `pos` (depth D),
     -1 if none; pos must lie in the table's range */
  TMR_HD i64 pred(const u64 *keys, u64 pos) const {
    const u64 cell = (pos >> shift) - cell0;
    const RankEntry e = load(cell >> 6);
    const int bit = (int)(cell & 63);
    const u64 bits = ((u64)e.bits[1] << 32) | (u64)e.bits[0];
    const i64 ci = (i64)e.start + popc64(bits << (63 - bit)) - 1;
    if (!cell_first) return ci;
    if (ci < 0) return -1;
    if (!((bits >> bit) & 1)) return (i64)cell_first[ci + 1] - 1;
    i64 lo = cell_first[ci], hi = cell_first[ci + 1];
    const u64 q = (pos << 5) | 31ULL;
    while (lo < hi) {
      const i64 mid = lo + ((hi - lo) >> 1);
      if (keys[mid] <= q) {
        lo = mid + 1;
      } else {
        hi = mid;
      }
    }
    return lo - 1;
  }
};

/* is element i the first leaf of its level-Dg cell? */
TMR_HD bool rank_first_of_cell(const u64 *keys, i64 i, int shift) {
  return i == 0 || ((keys[i - 1] >> 5) >> shift) != ((keys[i] >> 5) >> shift);
}
struct RankMarkFn {
  const u64 *keys;
  RankEntry *tab;
  u64 cell0;
  int shift;
  TMR_HD void operator()(i64 i) const {
    if (shift && !rank_first_of_cell(keys, i, shift)) return;
    const u64 cell = ((keys[i] >> 5) >> shift) - cell0;
    const int bit = (int)(cell & 63);
    TMR_ATOMIC_OR_I32(&tab[cell >> 6].bits[bit >> 5], 1u << (bit & 31));
  }
};
struct RankBitsFn {
  const RankEntry *tab;
  TMR_HD u32 operator()(i64 b) const {
    return (u32)(popc32(tab[b].bits[0]) + popc32(tab[b].bits[1]));
  }
};
struct RankStartFn {
  RankEntry *tab;
  TMR_HD void operator()(i64 b, u32 before) const { tab[b].start = before; }
};
struct RankCellFirstFn {
  const u64 *keys;
  i64 n;
  RankIndex ix;
  u32 *cell_first;
  u32 nmarked;
  TMR_HD void operator()(i64 i) const {
    if (i == n) {
      cell_first[nmarked] = (u32)n;
      return;
    }
    if (!rank_first_of_cell(keys, i, ix.shift)) return;
    const u64 cell = ((keys[i] >> 5) >> ix.shift) - ix.cell0;
    const RankEntry e = ix.load(cell >> 6);
    const int bit = (int)(cell & 63);
    const u64 bits = ((u64)e.bits[1] << 32) | (u64)e.bits[0];
    cell_first[(i64)e.start + popc64(bits << (63 - bit)) - 1] = (u32)i;
  }
};

/* rank index over the sorted element keys [first, last] of this rank */
inline RankIndex build_rank_index(Ctx &ctx, const u64 *keys, i64 n, int D,
                                  u64 pos_first, u64 pos_end, size_t budget,
                                  DBuf<RankEntry> &tab_store, DBuf<u32> &cf_store) {
  RankIndex ix;
  int Dg = D;
  u64 c0, c1;
  while (true) {
    const int sh = 3 * (D - Dg);
    c0 = (pos_first >> sh) & ~63ULL;
    c1 = ((pos_end - 1) >> sh) + 1;
    if (((c1 - c0 + 63) >> 6) * sizeof(RankEntry) <= budget || Dg == 0) break;
    Dg--;
  }
  ix.shift = 3 * (D - Dg);
  ix.cell0 = c0;
  ix.cell_first = NULL;
  const i64 nbuckets = (i64)((c1 - c0 + 63) >> 6);
  tab_store.alloc(ctx, nbuckets);
  dev_zero(ctx, tab_store.get(), (size_t)nbuckets * sizeof(RankEntry));
  RankMarkFn mk = {keys, tab_store.get(), c0, ix.shift};
  launch(ctx, n, mk, "rank_index_mark");
  RankBitsFn bf = {tab_store.get()};
  RankStartFn sf = {tab_store.get()};
  const i64 nmarked = (i64)scan_apply(ctx, nbuckets, bf, sf, "rank_index_scan");
  ix.tab = tab_store.get();
  if (ix.shift) {
    cf_store.alloc(ctx, nmarked + 1);
    RankCellFirstFn cf = {keys, n, ix, cf_store.get(), (u32)nmarked};
    launch(ctx, n + 1, cf, "rank_index_cells");
    ix.cell_first = cf_store.get();
  }
  return ix;
}

struct SlotView {
  const u64 *keys;
  i64 E;
  KeyFmt fmt;
  ConnTables t;
  RankIndex ix;
  /* positions (depth D) of this rank's own range; with several ranks a
     position outside it belongs to another rank (B list) */
  int multi;
  u64 pos_lo, pos_hi;
  u32 *mask;  /* one rank, per leaf: occupied slots */
  /* several ranks: occupied slots (low word) and, in the high word, the slots
     that are a corner of a LOCAL element -- one 64-bit reduction per mark */
  u64 *mc;
  u32 *dmask; /* slots that are dependent (hanging) nodes */
  int *fail;

  /* leaf and slot of the canonical position (block, Morton m of the depth-D
     cell holding it, clamp bit a set where the coordinate is 2^30-1; a = 2 x,
     1 y, 0 z): (leaf << 5) | ordinal, kLocB or kLocFail */
  template <class M>
  TMR_HD u64 locate(i32 block, M m, int clamp) const {
    const int D = fmt.D;
    const u64 pos = ((u64)(u32)block << (3 * D)) | (u64)m;
    if (pos < pos_lo || pos >= pos_hi) return multi ? kLocB : kLocFail;
    const i64 j = ix.pred(keys, pos);
    if (j < 0) return kLocFail;
    const u64 kl = keys[j];
    const int k = D - (int)(kl & 31);
    const M mall = D > 0 ? (M)(((u64)1 << (3 * D)) - 1) : (M)0;
    const M ml = (M)(kl >> 5) & mall;
    if (k < 0 || (i32)(kl >> (3 * D + 5)) != block) return kLocFail;
    const M lm = k > 0 ? (M)(((u64)1 << (3 * k)) - 1) : (M)0;
    if (((ml ^ m) & ~lm) != 0) return kLocFail; /* not inside the leaf */
    const M d = m & lm;
    int c6 = 0;
    TMR_UNROLL
    for (int a = 0; a < 3; a++) {
      const M am = axis_mask<M>(a) & lm;
      const M da = d & am;
      if ((clamp >> a) & 1) {
        if (da != am) return kLocFail;
        c6 |= (8 | 1) << a;
      } else if (da != 0) {
        if (da != ((M)1 << (3 * (k - 1) + a))) return kLocFail;
        c6 |= 8 << a;
      }
    }
    return ((u64)j << 5) | (u64)slot_ord(c6);
  }
  /* canonical coordinates (after transform_node) */
  TMR_HD u64 locate_xyz(i32 block, i32 x, i32 y, i32 z) const {
    const int s = kMaxLevel - fmt.D;
    const int clamp = ((x == kHmax - 1) ? 4 : 0) | ((y == kHmax - 1) ? 2 : 0) |
                      ((z == kHmax - 1) ? 1 : 0);
    return locate<u64>(block, morton3((u32)x >> s, (u32)y >> s, (u32)z >> s), clamp);
  }
  /* leaf and slot of a node KEY (NodeFmt at Dn = D, no label bits): the low
     bit of a squeezed coordinate is set exactly when it is the clamped
     2^30-1, and halving every coordinate drops the lowest bit triple */
  TMR_HD u64 locate_key(u64 nkey) const {
    const int sh = 3 * (fmt.D + 1);
    const u64 kk = nkey & ((1ULL << sh) - 1);
    return locate<u64>((i32)(nkey >> sh), kk >> 3, (int)(kk & 7));
  }
  TMR_HD void mark(u64 v, bool corner) const {
    if (v >= kLocFail) return;
    const i64 leaf = (i64)(v >> 5);
    const u32 bit = 1u << (int)(v & 31);
    /* fire-and-forget reductions: no load to wait for */
    if (mc) {
      TMR_ATOMIC_OR_U64(&mc[leaf], (u64)bit | (corner ? ((u64)bit << 32) : 0ULL));
    } else {
      TMR_ATOMIC_OR_I32(&mask[leaf], bit);
    }
  }
};

/* corners of an order-2 element that are dependent nodes, given its child id
   and 6-bit hanging info (labelDependentNodes, reference
   src/TMROctForest.cpp:3711-3832): the corner on the far end, from the
   parent's corner, of a hanging edge (or of an edge of a hanging face) is the
   parent's edge midpoint */
TMR_HD int dep_corner_mask2(int id, int inf);
/* tabulated: 8 child ids x 64 info values (512 bytes, built on the device) */
struct DepTableFn {
  unsigned char *table;
  TMR_HD void operator()(i64 i) const {
    table[i] = (unsigned char)dep_corner_mask2((int)(i >> 6), (int)(i & 63));
  }
};

/* per element: (leaf, slot) of its 8 corners.  Runs through launch_block3:
   collect() classifies the element and queues the points that need a search,
   process() drains the queues with all lanes busy, finish() assembles the
   element's row.  M: Morton word (u32 up to depth 10, else u64). */
/* kMulti = false compiles the off-range ("B") corner path out: on one rank no
   corner can lie outside the rank's range, and the path costs the kernel six
   registers (46 instead of 40: five instead of six resident CTAs per SM) */
template <class M, bool kMulti = true>
struct NodeSlotFn {
  SlotView v;
  u32 *conn_leaf;       /* [E][8] leaf index of every corner (kConnB: B list) */
  unsigned char *slot8; /* [E][8] slot ordinal */
  const int16_t *info;  /* hanging info of the elements (with v.dmask) */
  const unsigned char *dep_table; /* [child id][info] -> dependent corners */
  /* several ranks: corners whose position is not in this rank's range */
  NodeFmt nfmt;
  u64 *b_key;
  u32 *b_pay;
  unsigned long long *b_count;
  i64 b_cap;

  struct Shared {
    u64 val[kLaunchThreads * 8];   /* [item][r or corner] */
    /* task = (item << 3) | code; one list per kind of work so that the lanes
       of a warp run the same code: family grid points (code = r), corners of
       interior elements outside complete families (code = corner), corners of
       elements on a tree face (transform_node) */
    unsigned short fam[kLaunchThreads * 4];
    unsigned short own[kLaunchThreads * 8];
    unsigned short slow[kLaunchThreads * 8];
    short fam0[kLaunchThreads]; /* first item of the item's family, or -1 */
    int nfam, nown, nslow;
    TMR_HD void reset() { nfam = nown = nslow = 0; }
  };

  TMR_HD void append_b(u64 key, u32 payload) const {
    const unsigned long long s = fetch_add_u64(b_count, 1ULL);
    if ((i64)s < b_cap) {
      b_key[s] = key;
      b_pay[s] = payload;
    }
  }

  /* point tp = gx + 3 gy + 9 gz of the 3x3x3 grid of a family: is it the
     anchor of one of the 8 siblings? */
  TMR_HD static bool family_inner(int tp) {
    return (tp % 3) < 2 && ((tp / 3) % 3) < 2 && tp < 18;
  }

  TMR_HD void collect(i64 i, int t, Shared &sh) const {
    const u64 key = v.keys[i];
    const int L = (int)(key & 31), D = v.fmt.D;
    const int s = 3 * (D - L);
    const u64 m = D > 0 ? ((key >> 5) & ((1ULL << (3 * D)) - 1)) : 0ULL;
    sh.fam0[t] = -1;
    /* does the element's cell / its parent's cell touch a tree face? */
    bool lower = false, elem_in = L > 0, par_in = L > 1;
    {
      const u64 lv = L > 0 ? ((1ULL << (3 * L)) - 1) : 0ULL;
      const u64 lvp = L > 1 ? ((1ULL << (3 * (L - 1))) - 1) : 0ULL;
      const u64 mc = m >> s, mp = mc >> 3;
      TMR_UNROLL
      for (int a = 0; a < 3; a++) {
        const u64 am = (0x1249249249249249ULL << a) & lv;
        const u64 c = mc & am;
        lower = lower || c == 0;
        elem_in = elem_in && c != 0 && c != am;
        const u64 amp = (0x1249249249249249ULL << a) & lvp;
        const u64 cp = mp & amp;
        par_in = par_in && cp != 0 && cp != amp;
      }
    }
    /* the element's own anchor is a node in place unless it lies on a lower
       tree face (then it goes through transform_node like any other corner) */
    if (L > 0 && !lower) v.mark((u64)i << 5, true);
    if (par_in) {
      const int md = (int)((m >> s) & 7);
      const i64 e0 = i - md, i0 = i - t;
      if (e0 >= i0 && e0 + 7 < i0 + kLaunchThreads && e0 + 7 < v.E) {
        const u64 k0 = v.keys[e0];
        if ((k0 & 31) == (u64)L && ((k0 >> (5 + s)) & 7) == 0 &&
            v.keys[e0 + 7] == k0 + (7ULL << (5 + s))) {
          /* 8 consecutive keys from sibling 0 to sibling 7: in a valid leaf
             set they are the 8 siblings and this element is number md */
          if (key != k0 + ((u64)md << (5 + s))) {
            *v.fail = 1;
            return;
          }
          sh.fam0[t] = (short)(e0 - i0);
          int need = 0;
          TMR_UNROLL
          for (int r = 0; r < 4; r++) {
            const int tp = md + 8 * r;
            if (tp >= 27) break;
            if (family_inner(tp)) {
              sh.val[t * 8 + r] =
                  (u64)(e0 + 4 * (tp % 3) + 2 * ((tp / 3) % 3) + tp / 9) << 5;
            } else {
              need |= 1 << r;
            }
          }
          int q = fetch_add_i32(&sh.nfam, popc32((u32)need));
          TMR_UNROLL
          for (int r = 0; r < 4; r++) {
            if (need & (1 << r)) sh.fam[q++] = (unsigned short)((t << 3) | r);
          }
          return;
        }
      }
    }
    if (elem_in) {
      sh.val[t * 8] = (u64)i << 5;
      int q = fetch_add_i32(&sh.nown, 7);
      TMR_UNROLL
      for (int c = 1; c < 8; c++) sh.own[q++] = (unsigned short)((t << 3) | c);
    } else {
      int q = fetch_add_i32(&sh.nslow, 8);
      TMR_UNROLL
      for (int c = 0; c < 8; c++) sh.slow[q++] = (unsigned short)((t << 3) | c);
    }
  }

  TMR_HD void process(i64 i0, int t, Shared &sh) const {
    const int D = v.fmt.D;
    const M mall = D > 0 ? (M)(((u64)1 << (3 * D)) - 1) : (M)0;
    /* family grid points: steps of the element size from the parent's anchor */
    for (int q = t; q < sh.nfam; q += kLaunchThreads) {
      const int task = sh.fam[q], ts = task >> 3, code = task & 7;
      const u64 key = v.keys[i0 + ts];
      const int s = 3 * (D - (int)(key & 31));
      const M m = (M)(key >> 5) & mall;
      const int tp = (int)((m >> s) & 7) + 8 * code;
      const int gz = tp / 9, gy = (tp - 9 * gz) / 3, gx = tp - 9 * gz - 3 * gy;
      M mq = (m >> (s + 3)) << (s + 3);
      if (gz == 1) mq |= (M)1 << s;
      if (gy == 1) mq |= (M)1 << (s + 1);
      if (gx == 1) mq |= (M)1 << (s + 2);
      if (gz == 2) mq = morton_axis_add<M>(mq, 0, s + 3);
      if (gy == 2) mq = morton_axis_add<M>(mq, 1, s + 4);
      if (gx == 2) mq = morton_axis_add<M>(mq, 2, s + 5);
      const u64 val = v.template locate<M>((i32)(key >> (3 * D + 5)), mq, 0);
      v.mark(val, true);
      sh.val[ts * 8 + code] = val;
    }
    /* corners of interior elements */
    for (int q = t; q < sh.nown; q += kLaunchThreads) {
      const int task = sh.own[q], ts = task >> 3, code = task & 7;
      const u64 key = v.keys[i0 + ts];
      const int s = 3 * (D - (int)(key & 31));
      M mq = (M)(key >> 5) & mall;
      if (code & 1) mq = morton_axis_add<M>(mq, 2, s + 2);
      if (code & 2) mq = morton_axis_add<M>(mq, 1, s + 1);
      if (code & 4) mq = morton_axis_add<M>(mq, 0, s);
      const u64 val = v.template locate<M>((i32)(key >> (3 * D + 5)), mq, 0);
      v.mark(val, true);
      sh.val[ts * 8 + code] = val;
    }
    /* corners of elements on a tree face: coordinates and transform_node */
    for (int q = t; q < sh.nslow; q += kLaunchThreads) {
      const int task = sh.slow[q], ts = task >> 3, c = task & 7;
      i32 b, x, y, z;
      int L;
      v.fmt.decode(v.keys[i0 + ts], &b, &x, &y, &z, &L);
      const i32 h = 1 << (kMaxLevel - L);
      x += (c & 1) * h;
      y += ((c >> 1) & 1) * h;
      z += (c >> 2) * h;
      transform_node(v.t, &b, &x, &y, &z, -1, NULL, NULL);
      const u64 val = v.locate_xyz(b, x, y, z);
      v.mark(val, true);
      sh.val[ts * 8 + c] = val;
    }
  }

  TMR_HD void finish(i64 i, int t, Shared &sh) const {
    const int f0 = sh.fam0[t];
    int gx0 = 0, gy0 = 0, gz0 = 0;
    if (f0 >= 0) {
      const u64 key = v.keys[i];
      const int md = (int)((key >> (5 + 3 * (v.fmt.D - (int)(key & 31)))) & 7);
      gz0 = md & 1;
      gy0 = (md >> 1) & 1;
      gx0 = (md >> 2) & 1;
    }
    /* dependent corners (labelDependentNodes), labelled right here in slot space */
    int dm = 0;
    if (v.dmask && info[i]) {
      const u64 key = v.keys[i];
      const int L = (int)(key & 31);
      const int md = L == 0 ? 0 : (int)((key >> (5 + 3 * (v.fmt.D - L))) & 7);
      const int id = ((md >> 2) & 1) | (md & 2) | ((md & 1) << 2);
      dm = dep_table[(id << 6) | (info[i] & 63)];
    }
    u32 leaf[8];
    u64 ords = 0;
    TMR_UNROLL
    for (int c = 0; c < 8; c++) {
      u64 val;
      if (f0 >= 0) {
        const int tp = 9 * (gz0 + (c >> 2)) + 3 * (gy0 + ((c >> 1) & 1)) + gx0 + (c & 1);
        val = sh.val[(f0 + (tp & 7)) * 8 + (tp >> 3)];
      } else {
        val = sh.val[t * 8 + c];
      }
      if (val == kLocFail) {
        *v.fail = 1;
        leaf[c] = 0;
        dm &= ~(1 << c);
      } else if (kMulti && val == kLocB) {
        i32 b, x, y, z;
        int L;
        v.fmt.decode(v.keys[i], &b, &x, &y, &z, &L);
        const i32 h = 1 << (kMaxLevel - L);
        x += (c & 1) * h;
        y += ((c >> 1) & 1) * h;
        z += (c >> 2) * h;
        transform_node(v.t, &b, &x, &y, &z, -1, NULL, NULL);
        /* payload: conn slot, bit 31 = this corner is a dependent node */
        append_b(nfmt.encode(b, x, y, z, 0),
                 (u32)(i * 8 + c) | (((dm >> c) & 1) ? kPayDep : 0u));
        leaf[c] = kConnB;
        dm &= ~(1 << c);
      } else {
        leaf[c] = (u32)(val >> 5);
        ords |= (val & 31) << (8 * c);
      }
    }
    if (dm) {
      TMR_UNROLL
      for (int c = 0; c < 8; c++) {
        if ((dm >> c) & 1) {
          TMR_ATOMIC_OR_I32(&v.dmask[leaf[c]], 1u << (int)((ords >> (8 * c)) & 31));
        }
      }
    }
    *reinterpret_cast<u64 *>(slot8 + i * 8) = ords;
    store8_u32(conn_leaf + i * 8, leaf);
  }
};

/* several ranks: the parent edge / face nodes of hanging elements whose coarse
   neighbour is remote (ParentNodeGen) must exist as nodes too */
template <class NS>
struct SlotKeyEmit {
  const SlotView *v;
  const NodeFmt *nfmt;
  const NS *ns;
  TMR_HD void operator()(i32 b, i32 x, i32 y, i32 z, int) const {
    const u64 val = v->locate_xyz(b, x, y, z);
    if (val == kLocFail) {
      *v->fail = 1;
    } else if (val == kLocB) {
      ns->append_b(nfmt->encode(b, x, y, z, 0), kConnB);
    } else {
      v->mark(val, false);
    }
  }
};


/* one rank: nodes and dependent nodes counted together, lo | hi << 31 */
struct SlotCount2Fn {
  const u32 *mask;
  const u32 *dmask;
  TMR_HD u64 operator()(i64 i) const {
    return (u64)popc32(mask[i]) | ((u64)popc32(dmask[i]) << 31);
  }
};
struct SlotInfo2Fn {
  const u32 *mask;
  const u32 *dmask;
  SlotInfo2 *slotinfo;
  TMR_HD void operator()(i64 i, u64 off) const {
#if defined(__CUDA_ARCH__)
    *reinterpret_cast<uint4 *>(slotinfo + i) =
        make_uint4((u32)(off & 0x7fffffffULL), mask[i], (u32)(off >> 31), dmask[i]);
#else
    SlotInfo2 si = {(u32)(off & 0x7fffffffULL), mask[i], (u32)(off >> 31), dmask[i]};
    slotinfo[i] = si;
#endif
  }
};
/* number of the node in slot `ord` of a leaf on ONE rank (reference
   :4113-4181): dependents -1, -2, .. and independents 0, 1, .. both in node
   order */
TMR_HD int slot_node_number(const SlotInfo2 &si, int ord) {
  const u32 below = (1u << ord) - 1u;
  const int dep_before = (int)si.dep_off + popc32(si.dmask & below);
  if ((si.dmask >> ord) & 1u) return -dep_before - 1;
  return (int)si.node_off + popc32(si.mask & below) - dep_before;
}
TMR_HD SlotInfo2 load_slotinfo2(const SlotInfo2 *p) {
#if defined(__CUDA_ARCH__)
  const uint4 v = *reinterpret_cast<const uint4 *>(p);
  SlotInfo2 si = {v.x, v.y, v.z, v.w};
  return si;
#else
  return *p;
#endif
}

/* one rank: number of the node at a canonical position, straight from the
   slots (rank index + per-leaf prefix counts) -- no node key array needed */
struct SlotLookup {
  int on;
  SlotView v;
  const SlotInfo2 *si;
  TMR_HD int number(i32 b, i32 x, i32 y, i32 z) const {
    const u64 loc = v.locate_xyz(b, x, y, z);
    if (loc >= kLocFail) return 0;
    const i64 leaf = (i64)(loc >> 5);
    const int ord = (int)(loc & 31);
    const SlotInfo2 s = load_slotinfo2(si + leaf);
    return ((s.mask >> ord) & 1u) ? slot_node_number(s, ord) : 0;
  }
};

/* key (NodeFmt at Dn = D) of slot c6 of the leaf `key` */
TMR_HD u64 slot_node_key(u64 key, int D, int c6) {
  const int k = D - (int)(key & 31);
  const u64 m = D > 0 ? ((key >> 5) & ((1ULL << (3 * D)) - 1)) : 0ULL;
  const u64 block = key >> (3 * D + 5);
  const u64 lm = (1ULL << (3 * (k + 1))) - 1;
  u64 extra = 0;
  TMR_UNROLL
  for (int a = 0; a < 3; a++) {
    if ((c6 >> a) & 1) {
      extra |= (0x1249249249249249ULL << a) & lm;
    } else if ((c6 >> (3 + a)) & 1) {
      extra |= 1ULL << (3 * k + a);
    }
  }
  return (block << (3 * (D + 1))) | (m << 3) | extra;
}

/* one rank: node keys and node numbers */
struct SlotKeys2Fn {
  const u64 *keys;
  KeyFmt fmt;
  const SlotInfo2 *slotinfo;
  u64 *node_keys;
  int *node_num;
  TMR_HD void operator()(i64 i) const {
    const SlotInfo2 si = load_slotinfo2(slotinfo + i);
    u32 mk = si.mask;
    if (!mk) return;
    const u64 key = keys[i];
    i64 o = si.node_off;
    while (mk) {
      const int ord = ctz32(mk);
      mk &= mk - 1;
      node_keys[o] = slot_node_key(key, fmt.D, slot_code(ord));
      node_num[o] = slot_node_number(si, ord);
      o++;
    }
  }
};

/* one rank: (leaf, slot) -> node NUMBER, in place: the connectivity is final
   after this pass (no local-index stage, no renumbering pass) */
struct SlotResolve2Fn {
  const SlotInfo2 *slotinfo;
  const unsigned char *slot8;
  u32 *conn;
  /* the element's 8 node numbers, also left in `leaf` for a fused consumer */
  TMR_HD void row(i64 e, u32 *leaf) const {
    load8_u32(conn + e * 8, leaf);
    const u64 ords = *reinterpret_cast<const u64 *>(slot8 + e * 8);
    TMR_UNROLL
    for (int c = 0; c < 8; c++) {
      const SlotInfo2 si = load_slotinfo2(slotinfo + leaf[c]);
      leaf[c] = (u32)slot_node_number(si, (int)((ords >> (8 * c)) & 31));
    }
    store8_u32(conn + e * 8, leaf);
  }
  TMR_HD void operator()(i64 e) const {
    u32 leaf[8];
    row(e, leaf);
  }
};


/* ---- several ranks: ownership and numbering in slot space ---------------------
   Every slot node lies in this rank's own Morton range, so this rank is its
   HOME (the reference distributes the node array by position,
   src/TMROctForest.cpp:4545): only the B nodes travel for the ownership
   decision (lowest rank creating the node from an element, :4538-4637), and
   a slot node is owned here unless a lower rank donated it (lowmask) or no
   local element creates it (mask & ~cmask).  Those few "external" slot nodes
   are listed (xref, sorted by leaf and slot) and their numbers fetched from
   the owners; every other number follows from two prefix counts per leaf. */
struct SlotInfoM { /* 16 bytes: what a node NUMBER needs */
  u32 dep_off, dmask, own_off, omask; /* omask: owned independent slots */
};
TMR_HD SlotInfoM load_slotinfom(const SlotInfoM *p) {
#if defined(__CUDA_ARCH__)
  const uint4 a = *reinterpret_cast<const uint4 *>(p);
  SlotInfoM si = {a.x, a.y, a.z, a.w};
  return si;
#else
  return *p;
#endif
}
struct SlotNumbers {
  const SlotInfoM *si;
  const u64 *xref; /* external slot nodes, (leaf << 5 | slot) ascending */
  const int *xnum; /* their numbers, from the owners */
  i64 nx;
  int dep_base; /* dependent B nodes below the rank's range */
  int own_base; /* first owned number + owned B nodes below the range */
  /* number of the node in occupied slot `ord` of `leaf` */
  TMR_HD int number(const SlotInfoM &s, i64 leaf, int ord) const {
    const u32 bit = 1u << ord, below = bit - 1u;
    if (s.dmask & bit) return -(dep_base + (int)s.dep_off + popc32(s.dmask & below)) - 1;
    if (s.omask & bit) return own_base + (int)s.own_off + popc32(s.omask & below);
    const i64 j = find_u64(xref, nx, ((u64)leaf << 5) | (u64)ord);
    return j >= 0 ? xnum[j] : 0;
  }
};
/* first scan: node positions */
struct SlotCountMFn {
  const u64 *mc;
  TMR_HD u32 operator()(i64 i) const { return (u32)popc32((u32)mc[i]); }
};
struct SlotNoStoreFn { /* scan for the total only */
  TMR_HD void operator()(i64, u32) const {}
};
struct SlotInfoMFn {
  const u64 *mc;
  u64 *slotinfo;
  TMR_HD void operator()(i64 i, u32 off) const {
    slotinfo[i] = ((u64)off << 32) | (u64)(u32)mc[i];
  }
};
/* second scan: dependents and owned independents, lo | hi << 31 */
struct SlotCount3Fn {
  const u64 *mc;
  const u32 *dmask, *lowmask;
  TMR_HD u32 om(i64 i) const { /* owned independent slots */
    const u64 m = mc[i];
    return (u32)m & (u32)(m >> 32) & ~lowmask[i] & ~dmask[i];
  }
  TMR_HD u64 operator()(i64 i) const {
    return (u64)popc32(dmask[i]) | ((u64)popc32(om(i)) << 31);
  }
};
struct SlotInfo3Fn {
  SlotCount3Fn c;
  SlotInfoM *out;
  TMR_HD void operator()(i64 i, u64 off) const {
    const u32 dep_off = (u32)(off & 0x7fffffffULL), own_off = (u32)(off >> 31);
#if defined(__CUDA_ARCH__)
    *reinterpret_cast<uint4 *>(out + i) = make_uint4(dep_off, c.dmask[i], own_off, c.om(i));
#else
    SlotInfoM si = {dep_off, c.dmask[i], own_off, c.om(i)};
    out[i] = si;
#endif
  }
};
/* home side of the ownership exchange: the received B nodes of the other
   ranks (sorted by key, run-wise minimum donor in owner_run) against my
   slots.  Replies the owner; marks my copy as external when it is not me. */
struct HomeSlotFn {
  SlotView v;
  const u64 *rkeys;
  const u32 *idx;
  const u32 *run_of;
  const int *owner_run;
  int me;
  u32 *lowmask;
  u64 *xref;
  u32 *xdest;
  unsigned long long *xcount;
  int *reply; /* by original received index */
  TMR_HD void operator()(i64 j) const {
    int o = owner_run[run_of[j]];
    const u64 loc = v.locate_key(rkeys[j]);
    if (loc < kLocFail) {
      const i64 leaf = (i64)(loc >> 5);
      const u32 bit = 1u << (int)(loc & 31);
      const u64 m = v.mc[leaf];
      if ((u32)m & bit) {
        if (((u32)(m >> 32) & bit) && me < o) o = me;
        if (o != me && (j == 0 || rkeys[j - 1] != rkeys[j])) {
          TMR_ATOMIC_OR_I32(&lowmask[leaf], bit);
          const unsigned long long q = fetch_add_u64(xcount, 1ULL);
          xref[q] = loc;
          xdest[q] = (u32)(o == 0x7fffffff ? me : o);
        }
      }
    }
    reply[idx[j]] = o;
  }
};
/* request keys of the external slot nodes (xref sorted) */
struct XrefKeyFn {
  const u64 *keys;
  int D;
  const u64 *xref;
  u64 *out;
  TMR_HD void operator()(i64 q) const {
    out[q] = slot_node_key(keys[xref[q] >> 5], D, slot_code((int)(xref[q] & 31)));
  }
};
/* B nodes: per unique node dependent / owned flags, numbers */
struct BCountFn {
  const unsigned char *udep;
  const int *owner;
  int me;
  TMR_HD u64 operator()(i64 r) const {
    return udep[r] ? 1ULL : ((owner[r] == me) ? (1ULL << 31) : 0ULL);
  }
};
struct BStoreFn {
  u64 *before;
  TMR_HD void operator()(i64 r, u64 off) const { before[r] = off; }
};
struct BExternalCountFn {
  const unsigned char *udep;
  const int *owner;
  int me;
  TMR_HD u32 operator()(i64 r) const { return (!udep[r] && owner[r] != me) ? 1u : 0u; }
};
struct BExternalFillFn {
  BExternalCountFn c;
  const u64 *ukeys;
  u64 *out_key;
  u32 *out_dest;
  u32 *out_node;
  TMR_HD void operator()(i64 r, u32 o) const {
    if (c(r)) {
      out_key[o] = ukeys[r];
      out_dest[o] = (u32)(c.owner[r] < 0 ? c.me : c.owner[r]);
      out_node[o] = (u32)r;
    }
  }
};
/* numbers of the B nodes: [dependents | owned] counted before them in node
   order = their own prefix, plus everything in the slots for the B nodes that
   follow the rank's range */
struct BNumbers {
  const u64 *before; /* dep | own << 31, exclusive */
  const unsigned char *udep;
  const int *owner;
  i64 nlow;
  int me;
  int first_owned;
  int dep_slots, own_slots; /* totals of the slot nodes */
  TMR_HD int number(i64 r) const { /* 0 for an external node */
    const int d = (int)(before[r] & 0x7fffffffULL), o = (int)(before[r] >> 31);
    if (udep[r]) return -(d + (r < nlow ? 0 : dep_slots)) - 1;
    if (owner[r] == me) return first_owned + o + (r < nlow ? 0 : own_slots);
    return 0;
  }
};
struct BNumberFn {
  BNumbers b;
  int *num;
  TMR_HD void operator()(i64 r) const { num[r] = b.number(r); }
};
/* owner side of the number exchange */
struct LookupSlotNumberFn {
  const u64 *req;
  SlotView v;
  SlotNumbers sn;
  const u64 *b_ukeys;
  i64 nbu;
  BNumbers bn;
  int *reply;
  TMR_HD void operator()(i64 i) const {
    const u64 loc = v.locate_key(req[i]);
    if (loc < kLocFail) {
      const i64 leaf = (i64)(loc >> 5);
      const int ord = (int)(loc & 31);
      const SlotInfoM s = load_slotinfom(sn.si + leaf);
      /* sn.nx = 0 here: a slot that is neither dependent nor owned answers 0 */
      reply[i] = !(((u32)v.mc[leaf] >> ord) & 1u) ? -1 : sn.number(s, leaf, ord);
    } else {
      const i64 r = find_u64(b_ukeys, nbu, req[i]);
      reply[i] = r >= 0 ? bn.number(r) : -1;
    }
  }
};
/* several ranks: number of the node at a canonical position (the dependent
   stencils' parent nodes), from the slots or, outside the rank's range, from
   the B nodes */
struct SlotLookupM {
  int on;
  SlotView v;
  SlotNumbers sn;
  NodeFmt nfmt;
  const u64 *b_ukeys;
  i64 nbu;
  const int *b_num;
  TMR_HD int number(i32 b, i32 x, i32 y, i32 z) const {
    const u64 loc = v.locate_xyz(b, x, y, z);
    if (loc < kLocFail) {
      const i64 leaf = (i64)(loc >> 5);
      const int ord = (int)(loc & 31);
      if (!(((u32)v.mc[leaf] >> ord) & 1u)) return 0;
      return sn.number(load_slotinfom(sn.si + leaf), leaf, ord);
    }
    if (loc == kLocB) {
      const i64 r = find_u64(b_ukeys, nbu, nfmt.encode(b, x, y, z, 0));
      return r >= 0 ? b_num[r] : 0;
    }
    return 0;
  }
};
/* keys and numbers of the B nodes, at their places in node order */
struct BNodeArraysFn {
  const u64 *b_ukeys;
  const int *b_num;
  i64 nlow, NA;
  u64 *node_keys;
  int *node_num;
  TMR_HD void operator()(i64 r) const {
    const i64 idx = r < nlow ? r : r + NA;
    node_keys[idx] = b_ukeys[r];
    node_num[idx] = b_num[r];
  }
};
struct StoreBExternalFn {
  const u32 *node;
  const int *number;
  int *num;
  TMR_HD void operator()(i64 k) const { num[node[k]] = number[k]; }
};
/* node keys and numbers of every occupied slot, in node order */
struct SlotKeys3Fn {
  const u64 *keys;
  KeyFmt fmt;
  const u64 *slotinfo1; /* first scan: node_off << 32 | mask */
  SlotNumbers sn;
  u64 *node_keys;
  int *node_num;
  i64 base; /* index of the first slot node */
  TMR_HD void operator()(i64 i) const {
    const u64 s1 = slotinfo1[i];
    u32 mk = (u32)s1;
    if (!mk) return;
    const SlotInfoM s = load_slotinfom(sn.si + i);
    const u64 key = keys[i];
    i64 o = base + (i64)(s1 >> 32);
    while (mk) {
      const int ord = ctz32(mk);
      mk &= mk - 1;
      node_keys[o] = slot_node_key(key, fmt.D, slot_code(ord));
      node_num[o] = sn.number(s, i, ord);
      o++;
    }
  }
};
/* (leaf, slot) -> node NUMBER, in place (B corners are placed by BPlace3Fn) */
struct SlotResolve3Fn {
  SlotNumbers sn;
  const unsigned char *slot8;
  u32 *conn;
  /* returns false when a corner is still to be placed by the B pass */
  TMR_HD bool row(i64 e, u32 *leaf) const {
    load8_u32(conn + e * 8, leaf);
    const u64 ords = *reinterpret_cast<const u64 *>(slot8 + e * 8);
    bool complete = true;
    TMR_UNROLL
    for (int c = 0; c < 8; c++) {
      if (leaf[c] == kConnB) {
        complete = false;
        continue;
      }
      const SlotInfoM s = load_slotinfom(sn.si + leaf[c]);
      leaf[c] = (u32)sn.number(s, (i64)leaf[c], (int)((ords >> (8 * c)) & 31));
    }
    store8_u32(conn + e * 8, leaf);
    return complete;
  }
  TMR_HD void operator()(i64 e) const {
    u32 leaf[8];
    row(e, leaf);
  }
};

}  // namespace tmrgpu

#endif
