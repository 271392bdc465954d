/*
  ops_balance.h -- 2:1 balance as a level-descending closure on sorted key sets.

  What the reference does (src/TMROctForest.cpp:2917-3089, balanceOctant
  :2763-2895, add{Face,Edge,Corner}Neighbors :2525-2745): a hash+queue ripple
  over "0-sibling" family representatives; a family at level L demands the
  families of its parent's face/edge(/corner) neighbours at level L-1,
  including the images of out-of-tree neighbours in every tree sharing that
  tree face/edge/corner; finally every family expands to 8 siblings and the
  finest octant at each anchor survives.

  Restated for the GPU: a family at level L is the set of children of ONE
  octant of level L-1, so the hash of families is the set R_l of octants of
  level l that must be refined.  The ripple rule becomes
      p in R_l (l >= 1)  =>  parent(p) in R_{l-1}  and  every level-(l-1)
      neighbour of parent(p) that touches p across a face/edge(/corner) is in
      R_{l-1}  (transformed into the adjacent trees when outside this one).
  Only level l-1 is ever produced from level l, so ONE sweep from the deepest
  level down closes the set; each level is: generate candidates (count ->
  scan -> fill), radix-sort them with the existing R_{l-1}, run-length dedup.
  The leaves are then the children of every R_l member that are not themselves
  in R_{l+1}.  The result is identical to the reference's (it is the closure
  of the same rule, which is independent of insertion order and rank count).
*/
#ifndef TMRGPU_OPS_BALANCE_H
#define TMRGPU_OPS_BALANCE_H

#include <stdlib.h>
#include <string.h>

#include "ops_octants.h"

namespace tmrgpu {

TMR_HD u64 low_mask(int bits) { return bits > 0 ? ((1ULL << bits) - 1) : 0ULL; }

/* (a) parent of every input octant as a combined key
       [ level l : 5 | block | Morton_l left-aligned in 3*Dp bits ]          */
struct ParentKeyFn {
  const u64 *keys;
  KeyFmt fmt;
  int Dp; /* fmt.D - 1 */
  int *root_flag;
  TMR_HD bool parent(i64 i, u64 *out) const {
    const u64 k = keys[i];
    const int L = (int)(k & 31);
    const u64 rest = k >> 5;
    const u64 block = rest >> (3 * fmt.D);
    if (L == 0) return false;
    const u64 m = rest & low_mask(3 * fmt.D);
    /* Morton code of the parent at depth L-1, then left-align to depth Dp */
    const u64 mp = (m >> (3 * (fmt.D - L + 1))) << (3 * (Dp - (L - 1)));
    *out = ((u64)(L - 1) << (fmt.bbits + 3 * Dp)) | (block << (3 * Dp)) | mp;
    return true;
  }
  TMR_HD u32 operator()(i64 i) const {
    u64 c, p;
    if (!parent(i, &c)) {
      root_flag[(int)((keys[i] >> 5) >> (3 * fmt.D))] = 1;
      return 0;
    }
    if (i > 0 && parent(i - 1, &p) && p == c) return 0;
    return 1;
  }
};

struct ParentFillFn {
  ParentKeyFn pk;
  const u32 *offset;
  u64 *out;
  TMR_HD void operator()(i64 i) const {
    u64 c, p;
    if (!pk.parent(i, &c)) return;
    if (i > 0 && pk.parent(i - 1, &p) && p == c) return;
    out[offset[i]] = c;
  }
};

/* first index of each level inside the sorted combined-key array */
struct LevelBoundsFn {
  const u64 *ckeys;
  i64 n;
  int shift; /* bbits + 3*Dp */
  i64 *bounds;
  TMR_HD void operator()(i64 l) const {
    bounds[l] = lower_bound_u64(ckeys, n, (u64)l << shift);
  }
};

/* combined key -> per-level key  pkey_l = [ block | Morton_l ] */
struct SplitLevelFn {
  const u64 *ckeys;
  int l, Dp, bbits;
  u64 *out;
  TMR_HD void operator()(i64 i) const {
    const u64 c = ckeys[i];
    const u64 block = (c >> (3 * Dp)) & low_mask(bbits);
    const u64 m = (c & low_mask(3 * Dp)) >> (3 * (Dp - l));
    out[i] = (block << (3 * l)) | m;
  }
};

/* Images of the grid cell q (integer coordinates on the N^3 grid of tree
   `block`, possibly one step outside it) in the forest: the cell itself when
   inside, otherwise its image in EVERY other tree sharing the tree face /
   edge / corner it crossed (reference addFaceNeighbors :2525-2590,
   addEdgeNeighbors :2608-2686, addCornerNeighbors :2704-2745 and their
   addAdjacent*ToQueue twins :3095-3273). */
template <class Emit>
TMR_HD void tree_images(const ConnTables &t, i32 block, const i32 q[3], i32 N,
                        Emit &emit) {
  const i32 M = N - 1;
  int out[3], nout = 0;
  for (int d = 0; d < 3; d++) {
    out[d] = (q[d] < 0 || q[d] >= N);
    nout += out[d];
  }
  if (nout == 0) {
    emit(block, q[0], q[1], q[2]);
  } else if (nout == 1) {
    const int axis = out[0] ? 0 : (out[1] ? 1 : 2);
    const int f = 2 * axis + (q[axis] < 0 ? 0 : 1);
    const int face = t.block_face_conn[6 * block + f];
    i32 a0, b0, u, v;
    face_pick(f, q[0], q[1], q[2], &a0, &b0);
    face_to_owner(t.block_face_ids[6 * block + f], M, a0, b0, &u, &v);
    for (int ip = t.face_block_ptr[face]; ip < t.face_block_ptr[face + 1]; ip++) {
      const int adj = t.face_block_conn[ip] / 6;
      if (adj == block) continue;
      const int af = t.face_block_conn[ip] % 6;
      i32 a1, b1, x, y, z;
      owner_to_face(t.block_face_ids[6 * adj + af], M, u, v, &a1, &b1);
      face_place(af, M * (af & 1), a1, b1, &x, &y, &z);
      emit(adj, x, y, z);
    }
  } else if (nout == 2) {
    int e;
    i32 u;
    if (out[1] && out[2]) {
      e = (q[1] < 0 ? 0 : 1) + (q[2] < 0 ? 0 : 2);
      u = q[0];
    } else if (out[0] && out[2]) {
      e = (q[0] < 0 ? 4 : 5) + (q[2] < 0 ? 0 : 2);
      u = q[1];
    } else {
      e = (q[0] < 0 ? 8 : 9) + (q[1] < 0 ? 0 : 2);
      u = q[2];
    }
    const int edge = t.block_edge_conn[12 * block + e];
    for (int ip = t.edge_block_ptr[edge]; ip < t.edge_block_ptr[edge + 1]; ip++) {
      const int adj = t.edge_block_conn[ip] / 12;
      if (adj == block) continue;
      const int ae = t.edge_block_conn[ip] % 12;
      const i32 uu = edge_is_reversed(t, block, e, adj, ae) ? M - u : u;
      i32 x, y, z;
      edge_place(ae, uu, M, &x, &y, &z);
      emit(adj, x, y, z);
    }
  } else {
    const int c = (q[0] < 0 ? 0 : 1) + (q[1] < 0 ? 0 : 2) + (q[2] < 0 ? 0 : 4);
    const int node = t.block_conn[8 * block + c];
    for (int ip = t.node_block_ptr[node]; ip < t.node_block_ptr[node + 1]; ip++) {
      const int adj = t.node_block_conn[ip] / 8;
      if (adj == block) continue;
      const int ac = t.node_block_conn[ip] % 8;
      emit(adj, M * (ac & 1), M * ((ac >> 1) & 1), M * (ac >> 2));
    }
  }
}

/* (c) candidates demanded by p in R_l at level l-1 */
struct BalanceGen {
  ConnTables t;
  const u64 *R;
  int l;
  int corner;

  template <class Emit>
  TMR_HD void run(i64 i, Emit &emit) const {
    const u64 pk = R[i];
    const int sh = 3 * l;
    const i32 block = (i32)(pk >> sh);
    u32 px, py, pz;
    unmorton3(pk & low_mask(sh), &px, &py, &pz);
    const i32 N = 1 << (l - 1); /* level-(l-1) grid size */
    const i32 q0[3] = {(i32)(px >> 1), (i32)(py >> 1), (i32)(pz >> 1)};
    /* outward direction per axis = the side of its parent p sits on */
    const i32 s[3] = {(px & 1) ? 1 : -1, (py & 1) ? 1 : -1, (pz & 1) ? 1 : -1};
    /* p's own parent: emitted once per sibling group */
    if (i == 0 || (R[i - 1] >> 3) != (pk >> 3)) {
      emit(block, q0[0], q0[1], q0[2]);
    }
    for (int a = 1; a < 8; a++) {
      if (a == 7 && !corner) continue;
      i32 q[3];
      for (int d = 0; d < 3; d++) q[d] = q0[d] + (((a >> d) & 1) ? s[d] : 0);
      tree_images(t, block, q, N, emit);
    }
  }
};

struct CountEmit {
  u32 n;
  TMR_HD void operator()(i32, i32, i32, i32) { n++; }
};

struct KeyEmit {
  u64 *out;
  int sh; /* 3*(l-1) */
  TMR_HD void operator()(i32 block, i32 x, i32 y, i32 z) {
    *out++ = ((u64)(u32)block << sh) | morton3((u32)x, (u32)y, (u32)z);
  }
};

struct BalanceCountFn {
  BalanceGen g;
  TMR_HD u32 operator()(i64 i) const {
    CountEmit e = {0};
    g.run(i, e);
    return e.n;
  }
};

struct BalanceFillFn {
  BalanceGen g;
  const u32 *offset;
  u64 *out;
  TMR_HD void operator()(i64 i) const {
    KeyEmit e = {out + offset[i], 3 * (g.l - 1)};
    g.run(i, e);
  }
};

/* (e) leaves = children of R_l members that are not in R_{l+1} */
struct LeafGen {
  const u64 *R;      /* R_l */
  const u64 *Rnext;  /* R_{l+1}, may be NULL */
  i64 nnext;
  int l;
  KeyFmt fmt; /* output format */
  TMR_HD u32 refined_mask(u64 pk) const {
    u32 mask = 0;
    if (nnext > 0) {
      i64 j = lower_bound_u64(Rnext, nnext, pk << 3);
      while (j < nnext && (Rnext[j] >> 3) == pk) {
        mask |= 1u << (Rnext[j] & 7);
        j++;
      }
    }
    return mask;
  }
};

struct LeafCountFn {
  LeafGen g;
  TMR_HD u32 operator()(i64 i) const {
    return 8u - (u32)popc32(g.refined_mask(g.R[i]));
  }
};

struct LeafFillFn {
  LeafGen g;
  const u32 *offset;
  u64 *out; /* already advanced to this level's base */
  TMR_HD void operator()(i64 i) const {
    const u64 pk = g.R[i];
    const u32 mask = g.refined_mask(pk);
    const int sh = 3 * g.l;
    const u64 block = pk >> sh;
    const u64 m = pk & low_mask(sh);
    const int L = g.l + 1;
    u64 *o = out + offset[i];
    for (u32 d = 0; d < 8; d++) {
      if (mask & (1u << d)) continue;
      const u64 mD = ((m << 3) | d) << (3 * (g.fmt.D - L));
      *o++ = (block << (3 * g.fmt.D + 5)) | (mD << 5) | (u64)L;
    }
  }
};

/* level-0 leaves: roots present in the input whose tree is never refined */
struct RootLeafCountFn {
  const int *root_flag;
  const u64 *R0;
  i64 n0;
  TMR_HD u32 operator()(i64 b) const {
    if (!root_flag[b]) return 0;
    return (n0 > 0 && find_u64(R0, n0, (u64)b) >= 0) ? 0u : 1u;
  }
};

struct RootLeafFillFn {
  RootLeafCountFn c;
  const u32 *offset;
  KeyFmt fmt;
  u64 *out;
  TMR_HD void operator()(i64 b) const {
    if (c(b)) out[offset[b]] = (u64)b << (3 * fmt.D + 5);
  }
};

int balance_map(Forest &f, int balance_corner); /* ops_balance_map.h */

inline int balance(Forest &f, int balance_corner) {
  Ctx &ctx = *f.ctx;
  f.last_mid = f.n;
  if (f.n == 0) return 0;
  {
    /* cell bitmaps when they fit (ops_balance_map.h); TMR_B200_BALANCE=sort
       forces the sorted-array closure below */
    const char *mode = getenv("TMR_B200_BALANCE");
    if (f.fmt.D > 0 && !(mode && strcmp(mode, "sort") == 0)) {
      const int rc = balance_map(f, balance_corner);
      if (getenv("TMR_B200_NODES_VERBOSE")) {
        fprintf(stderr, "[tmr_b200] balance: %s\n",
                rc >= 0 ? "cell bitmaps" : "sorted-array closure (maps over budget)");
      }
      if (rc >= 0) return rc;
    }
  }
  const int D = f.fmt.D;
  /* info of every surviving octant is 0 after balance (getSibling zeroes it,
     reference src/TMROctant.cpp:52) */
  f.info.reset();
  if (D == 0) { /* only level-0 octants: nothing to balance */
    f.last_out = f.n;
    return 0;
  }
  const int Dp = D - 1;
  const int bbits = f.bbits;
  const int nb = f.nblocks;

  trace_mark(ctx, NULL);
  /* (a) parents of the input octants, deduplicated against the predecessor */
  DBuf<int> root_flag(ctx, nb);
  dev_zero(ctx, root_flag.get(), (size_t)nb * sizeof(int));
  i64 npar;
  DBuf<u64> ckeys;
  {
    DBuf<u32> offset(ctx, f.n);
    ParentKeyFn pk = {f.keys.get(), f.fmt, Dp, root_flag.get()};
    npar = (i64)scan_counts(ctx, f.n, pk, offset.get(), "balance_parent_count");
    ckeys.alloc(ctx, npar);
    ParentFillFn pf = {pk, offset.get(), ckeys.get()};
    launch(ctx, f.n, pf, "balance_parent_fill");
  }
  trace_mark(ctx, "balance: parents");
  /* (b) sort by (level, block, Morton), dedup, split per level */
  std::vector<DBuf<u64> > R(D);
  std::vector<i64> nR(D, 0);
  {
    DBuf<u64> alt(ctx, npar);
    DBuf<u32> v0, v1;
    radix_sort(ctx, ckeys, alt, v0, v1, npar, 0, 5 + bbits + 3 * Dp);
    npar = unique_keep_last(ctx, ckeys, alt, v0, v1, npar, 0);
    DBuf<i64> d_bounds(ctx, D + 1);
    LevelBoundsFn lb = {ckeys.get(), npar, bbits + 3 * Dp, d_bounds.get()};
    launch(ctx, D + 1, lb, "balance_level_bounds");
    std::vector<i64> bounds(D + 1);
    copy_d2h(ctx, bounds.data(), d_bounds.get(), (size_t)(D + 1) * sizeof(i64));
    for (int l = 0; l < D; l++) {
      nR[l] = bounds[l + 1] - bounds[l];
      R[l].alloc(ctx, nR[l]);
      SplitLevelFn sp = {ckeys.get() + bounds[l], l, Dp, bbits, R[l].get()};
      launch(ctx, nR[l], sp, "balance_split_level");
    }
  }
  ckeys.reset();
  trace_mark(ctx, "balance: sort+split");

  /* (c) closure, deepest level first */
  for (int l = D - 1; l >= 1; l--) {
    if (nR[l] == 0) continue;
    DBuf<u32> offset(ctx, nR[l]);
    BalanceGen gen = {f.tables, R[l].get(), l, balance_corner};
    BalanceCountFn cnt = {gen};
    const i64 ncand =
        (i64)scan_counts(ctx, nR[l], cnt, offset.get(), "balance_cand_count");
    const i64 tot = nR[l - 1] + ncand;
    DBuf<u64> merged(ctx, tot), alt(ctx, tot);
    if (nR[l - 1] > 0) {
      copy_d2d(ctx, merged.get(), R[l - 1].get(),
               (size_t)nR[l - 1] * sizeof(u64));
    }
    BalanceFillFn fill = {gen, offset.get(), merged.get() + nR[l - 1]};
    launch(ctx, nR[l], fill, "balance_cand_fill");
    DBuf<u32> v0, v1;
    radix_sort(ctx, merged, alt, v0, v1, tot, 0, bbits + 3 * (l - 1));
    nR[l - 1] = unique_keep_last(ctx, merged, alt, v0, v1, tot, 0);
    R[l - 1].swap(merged);
  }

  trace_mark(ctx, "balance: closure");
  /* (e) leaves */
  std::vector<DBuf<u32> > loff(D);
  std::vector<i64> nleaf(D, 0);
  i64 total = 0;
  for (int l = 0; l < D; l++) {
    if (nR[l] == 0) continue;
    loff[l].alloc(ctx, nR[l]);
    LeafGen g = {R[l].get(), (l + 1 < D) ? R[l + 1].get() : NULL,
                 (l + 1 < D) ? nR[l + 1] : 0, l, f.fmt};
    LeafCountFn c = {g};
    nleaf[l] = (i64)scan_counts(ctx, nR[l], c, loff[l].get(), "balance_leaf_count");
    total += nleaf[l];
  }
  DBuf<u32> root_off(ctx, nb);
  RootLeafCountFn rc = {root_flag.get(), R[0].get(), nR[0]};
  const i64 nroot =
      (i64)scan_counts(ctx, nb, rc, root_off.get(), "balance_root_count");
  total += nroot;
  if (total >= (1LL << 31)) {
    fprintf(stderr,
            "TMROctForest Error: balance() would create %lld octants on one "
            "rank (int32 index limit of the TMROctForest API)\n",
            (long long)total);
    return 1;
  }
  DBuf<u64> out(ctx, total), out_alt(ctx, total);
  i64 base = 0;
  for (int l = 0; l < D; l++) {
    if (nR[l] == 0) continue;
    LeafGen g = {R[l].get(), (l + 1 < D) ? R[l + 1].get() : NULL,
                 (l + 1 < D) ? nR[l + 1] : 0, l, f.fmt};
    LeafFillFn fl = {g, loff[l].get(), out.get() + base};
    launch(ctx, nR[l], fl, "balance_leaf_fill");
    base += nleaf[l];
  }
  RootLeafFillFn rf = {rc, root_off.get(), f.fmt, out.get() + base};
  launch(ctx, nb, rf, "balance_root_fill");
  trace_mark(ctx, "balance: leaf fill");
  /* leaves have distinct anchors: order by (block, Morton) only */
  DBuf<u32> v0, v1;
  radix_sort(ctx, out, out_alt, v0, v1, total, 5, f.fmt.total_bits(), "leaves");
  f.keys.swap(out);
  f.n = total;
  f.last_out = f.n;
  trace_mark(ctx, "balance: final sort");
  return check_errors(ctx, "balance");
}

}  // namespace tmrgpu

#endif
