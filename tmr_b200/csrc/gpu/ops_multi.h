/*
  ops_multi.h -- the forest across several GPUs.

  The octant array is split along the Morton curve exactly like the reference:
  rank r holds a contiguous slice and `owners[r]` is its first octant; the owner
  of a position is found by the reference's scan over that table
  (getOctantMPIOwner, src/TMROctForest.cpp:2334-2343).  Every exchange of the
  reference (distributeOctants/sendOctants :2379-2509) becomes "compute the
  destination rank of each 8-byte key on the device, group by destination with
  one radix pass, all-to-all-v over NCCL".
*/
#ifndef TMRGPU_OPS_MULTI_H
#define TMRGPU_OPS_MULTI_H

#include "ops_nodes.h"
#include "ops_route.h"

namespace tmrgpu {

/* ---- owners table --------------------------------------------------------------
   owners[r] = first octant of rank r, or the sentinel (last block, hmax^3,
   tag -1) for an empty rank (reference :1805-1832, :2062-2081) */
inline int gather_owners(Forest &f, int backfill) {
  Ctx &ctx = *f.ctx;
  Comm *comm = forest_comm(f);
  if (!comm) {
    f.owners.clear();
    return 0;
  }
  Oct24 mine;
  mine.block = f.nblocks - 1;
  mine.x = mine.y = mine.z = kHmax;
  mine.tag = -1;
  mine.level = 0;
  mine.info = 0;
  if (f.n > 0) {
    u64 k;
    copy_d2h(ctx, &k, f.keys.get(), sizeof(u64));
    int level;
    f.fmt.decode(k, &mine.block, &mine.x, &mine.y, &mine.z, &level);
    mine.level = (int16_t)level;
    mine.tag = 0;
    if (f.info.get()) {
      int16_t inf;
      copy_d2h(ctx, &inf, f.info.get(), sizeof(int16_t));
      mine.info = inf;
    }
  }
  f.owners.resize(comm->size);
  comm->allgather_host(ctx, &mine, f.owners.data(), sizeof(Oct24));
  if (backfill) {
    for (int k = 1; k < comm->size; k++) {
      if (f.owners[k].tag == -1) f.owners[k] = f.owners[k - 1];
    }
  }
  return check_errors(ctx, "gather_owners");
}

/* make the key depth identical on all ranks (the deepest level anywhere) */
inline int unify_depth(Forest &f) {
  Ctx &ctx = *f.ctx;
  Comm *comm = forest_comm(f);
  if (!comm) return 0;
  const int Dg = (int)global_max(ctx, *comm, f.fmt.D);
  if (Dg != f.fmt.D) {
    if (!key_budget_ok(f, Dg)) {
      fprintf(stderr,
              "TMROctForest Error: depth %d on %d trees exceeds the 64-bit key "
              "budget of the CUDA path\n", Dg, f.nblocks);
      return 1;
    }
    RekeyFn rk = {f.keys.get(), f.fmt.D, Dg};
    launch(ctx, f.n, rk, "rekey");
    f.fmt.D = Dg;
  }
  return 0;
}

/* ---- repartition (reference :1922-2088) ------------------------------------------ */
inline int repartition(Forest &f, int max_rank) {
  Ctx &ctx = *f.ctx;
  Comm *comm = forest_comm(f);
  f.nodes.clear();
  f.interp.clear();
  if (!comm) return 0;
  const int R = comm->size, me = comm->rank;
  if (max_rank <= 0 || max_rank > R) max_rank = R;
  if (unify_depth(f)) return 1;
  std::vector<i64> ptr(R + 1, 0), nptr(R + 1, 0);
  /* element count and "carries an info array" of every rank in one gather: the
     info exchange below is a collective, so whether it happens must not depend
     on this rank alone (a rank without elements, or one whose octants were
     uploaded with all-zero info, holds no array while its peers do) */
  const i64 mine[2] = {f.n, f.info.get() ? 1 : 0};
  std::vector<i64> all(2 * (size_t)R);
  comm->allgather_host(ctx, mine, all.data(), sizeof(mine));
  bool any_info = false;
  for (int k = 0; k < R; k++) {
    ptr[k + 1] = ptr[k] + all[2 * k];
    any_info = any_info || all[2 * k + 1] != 0;
  }
  const i64 total = ptr[R];
  const i64 avg = total / max_rank, rem = total - avg * max_rank;
  for (int k = 0; k < max_rank; k++) nptr[k + 1] = nptr[k] + avg + (k < rem ? 1 : 0);
  for (int k = max_rank; k < R; k++) nptr[k + 1] = nptr[k];
  /* my slice [ptr[me], ptr[me+1]) cut by the new intervals */
  std::vector<i64> send_off(R + 1, 0), recv_off(R + 1, 0);
  for (int i = 0; i < R; i++) {
    i64 lo = nptr[i] > ptr[me] ? nptr[i] : ptr[me];
    i64 hi = nptr[i + 1] < ptr[me + 1] ? nptr[i + 1] : ptr[me + 1];
    send_off[i + 1] = send_off[i] + (hi > lo ? hi - lo : 0);
    lo = ptr[i] > nptr[me] ? ptr[i] : nptr[me];
    hi = ptr[i + 1] < nptr[me + 1] ? ptr[i + 1] : nptr[me + 1];
    recv_off[i + 1] = recv_off[i] + (hi > lo ? hi - lo : 0);
  }
  const i64 nnew = nptr[me + 1] - nptr[me];
  DBuf<u64> nk(ctx, nnew);
  comm->alltoallv(ctx, f.keys.get(), send_off.data(), nk.get(), recv_off.data(),
                  sizeof(u64));
  if (any_info) {
    if (!f.info.get() && f.n > 0) {
      f.info.alloc(ctx, f.n);
      dev_zero(ctx, f.info.get(), (size_t)f.n * sizeof(int16_t));
    }
    DBuf<int16_t> ni(ctx, nnew);
    comm->alltoallv(ctx, f.info.get(), send_off.data(), ni.get(),
                    recv_off.data(), sizeof(int16_t));
    f.info.swap(ni);
  }
  f.keys.swap(nk);
  f.n = nnew;
  if (gather_owners(f, 0)) return 1;
  return check_errors(ctx, "repartition");
}

/* ---- refine: ship representatives that fell outside this rank's range --------- */
struct PosOwnerFn {
  const u64 *keys;
  OwnerMap om;
  TMR_HD int operator()(i64 i) const { return om.owner(keys[i] >> 5); }
};

struct CountForeignFn {
  PosOwnerFn po;
  int me;
  int *flag;
  TMR_HD void operator()(i64 i) const {
    if (po(i) != me) TMR_ATOMIC_OR_I32(flag, 1);
  }
};

inline int refine_exchange(Forest &f) {
  Ctx &ctx = *f.ctx;
  Comm *comm = forest_comm(f);
  if (!comm) return 0;
  if (unify_depth(f)) return 1;
  DBuf<u64> own_store;
  OwnerMap om = make_owner_map(f, f.fmt.D, own_store);
  PosOwnerFn po = {f.keys.get(), om};
  DBuf<int> flag(ctx, 1);
  dev_zero(ctx, flag.get(), sizeof(int));
  CountForeignFn cf = {po, comm->rank, flag.get()};
  launch(ctx, f.n, cf, "refine_foreign_check");
  int h_flag = 0;
  copy_d2h(ctx, &h_flag, flag.get(), sizeof(int));
  if (!global_max(ctx, *comm, h_flag)) return 0; /* common case: all local */
  RoutePlan plan;
  make_route(ctx, *comm, f.n, po, plan);
  DBuf<u64> nk;
  route_array(ctx, *comm, plan, f.keys.get(), nk);
  const i64 has_info = global_max(ctx, *comm, f.info.get() ? 1 : 0);
  if (has_info) {
    if (!f.info.get()) {
      f.info.alloc(ctx, f.n);
      dev_zero(ctx, f.info.get(), (size_t)f.n * sizeof(int16_t));
    }
    DBuf<int16_t> ni;
    route_array(ctx, *comm, plan, f.info.get(), ni);
    f.info.swap(ni);
  }
  f.keys.swap(nk);
  f.n = plan.nrecv;
  sort_unique_elements(f);
  return check_errors(ctx, "refine_exchange");
}

/* ---- balance across ranks -----------------------------------------------------------
   Same closure as ops_balance.h, with the sets R_l distributed by the owner of
   each member's position.  Per level (deepest first): route the pending keys to
   their owners, sort+dedup, generate the candidates for the next level.  Then
   every member whose PARENT lives on another rank is copied there (so the
   parent's owner can tell which of its children are refined), leaves are
   generated by the owner of each refined octant and routed to the owner of
   their own position (the reference's final sibling exchange, :3036-3075). */
struct LevelOwnerFn {
  const u64 *pk; /* [block | Morton_l] */
  int l, D;
  OwnerMap om;
  TMR_HD int operator()(i64 i) const {
    const u64 k = pk[i];
    const u64 block = k >> (3 * l);
    const u64 m = k & low_mask(3 * l);
    return om.owner((block << (3 * D)) | (m << (3 * (D - l))));
  }
};

/* owner of the parent of a level-l member */
struct ParentOwnerFn {
  const u64 *pk;
  int l, D;
  OwnerMap om;
  TMR_HD int operator()(i64 i) const {
    const u64 k = pk[i];
    const u64 block = k >> (3 * l);
    const u64 m = (k & low_mask(3 * l)) & ~7ULL;
    return om.owner((block << (3 * D)) | (m << (3 * (D - l))));
  }
};

struct ForeignParentCountFn {
  ParentOwnerFn po;
  int me;
  TMR_HD u32 operator()(i64 i) const { return po(i) != me ? 1u : 0u; }
};

struct ForeignParentFillFn {
  ParentOwnerFn po;
  int me;
  u64 *out_keys;
  u32 *out_dest;
  TMR_HD void operator()(i64 i, u32 o) const {
    const int d = po(i);
    if (d != me) {
      out_keys[o] = po.pk[i];
      out_dest[o] = (u32)d;
    }
  }
};

struct ArrayDestFn {
  const u32 *dest;
  TMR_HD int operator()(i64 i) const { return (int)dest[i]; }
};

struct ParentOfOctantFn { /* parents of my octants as per-level keys */
  const u64 *keys;
  KeyFmt fmt;
  int l; /* wanted parent level */
  TMR_HD bool get(i64 i, u64 *out) const {
    const u64 k = keys[i];
    const int L = (int)(k & 31);
    if (L != l + 1) return false;
    const u64 rest = k >> 5;
    const u64 block = rest >> (3 * fmt.D);
    const u64 m = rest & low_mask(3 * fmt.D);
    *out = (block << (3 * l)) | (m >> (3 * (fmt.D - l)));
    return true;
  }
};

struct ParentLevelCountFn {
  ParentOfOctantFn p;
  int *root_flag;
  TMR_HD u32 operator()(i64 i) const {
    u64 c, q;
    if (p.l == 0 && (p.keys[i] & 31) == 0) {
      root_flag[(int)((p.keys[i] >> 5) >> (3 * p.fmt.D))] = 1;
    }
    if (!p.get(i, &c)) return 0;
    if (i > 0 && p.get(i - 1, &q) && q == c) return 0;
    return 1;
  }
};

struct ParentLevelFillFn {
  ParentOfOctantFn p;
  u64 *out;
  TMR_HD void operator()(i64 i, u32 o) const {
    u64 c, q;
    if (!p.get(i, &c)) return;
    if (i > 0 && p.get(i - 1, &q) && q == c) return;
    out[o] = c;
  }
};

struct LeafOwnerFn {
  const u64 *keys;
  OwnerMap om;
  TMR_HD int operator()(i64 i) const { return om.owner(keys[i] >> 5); }
};

/* ---- balance across ranks on cell bitmaps (ops_balance_map.h) ---------------------
   Every rank keeps the maps of the trees its own range of positions touches.
   A demand that lands in a tree other ranks hold too -- a tree cut by a
   partition boundary, or a tree outside this rank's span -- is also sent to
   them, level by level (the closure only ever goes from level l to l-1, so
   one exchange per level closes it exactly).  Each rank then has the final
   refinement of every cell that overlaps its range and writes its own leaves
   in Morton order: no distributed sort, no leaf routing.
   Returns -1 when not applicable (maps over budget, or octants held outside
   their owner's range, the createTrees back-fill state of reference
   :1826-1832): the sorted-array closure below takes over. */
struct MapU32DestFn {
  const u32 *dest;
  TMR_HD int operator()(i64 i) const { return (int)dest[i]; }
};

inline int balance_multi_map(Forest &f, int balance_corner) {
  Ctx &ctx = *f.ctx;
  Comm &comm = *forest_comm(f);
  const int me = comm.rank, R = comm.size;
  const int D = f.fmt.D;
  if (D < 1 || D > kMaxMapLevels) return -1;
  std::vector<u64> pos;
  owner_positions(f, D, pos);
  const u64 end = (u64)f.nblocks << (3 * D);
  std::vector<u64> lo(R), hi(R);
  std::vector<i32> span(2 * R);
  int ok = 1;
  for (int r = 0; r < R; r++) {
    lo[r] = r == 0 ? 0 : pos[r];
    hi[r] = r + 1 < R ? pos[r + 1] : end;
    if (lo[r] > end) lo[r] = end;
    if (hi[r] > end) hi[r] = end;
    if (hi[r] < lo[r]) ok = 0;
    span[r] = hi[r] > lo[r] ? (i32)(lo[r] >> (3 * D)) : 1;
    span[R + r] = hi[r] > lo[r] ? (i32)((hi[r] - 1) >> (3 * D)) : 0;
  }
  const int b0 = span[me], b1 = span[R + me];
  const int nblk = b1 >= b0 ? b1 - b0 + 1 : 0;
  CellMaps mp;
  i64 budget = (i64)1 << 27;
  if (const char *ev = getenv("TMR_B200_BALANCE_MAP_WORDS")) budget = atol(ev);
  i64 words = ok ? balance_map_words(nblk > 0 ? nblk : 1, D, budget, mp.woff) : -1;
  if (words < 0) ok = 0;
  /* my octants must lie in my owner range */
  if (ok && f.n > 0) {
    u64 kf = 0, kl = 0;
    copy_d2h(ctx, &kf, f.keys.get(), sizeof(u64));
    copy_d2h(ctx, &kl, f.keys.get() + (f.n - 1), sizeof(u64));
    if ((kf >> 5) < lo[me] || (kl >> 5) >= hi[me]) ok = 0;
  }
  {
    std::vector<i64> all(R);
    const i64 mine = ok;
    comm.allgather_host(ctx, &mine, all.data(), sizeof(i64));
    for (int r = 0; r < R; r++) {
      if (!all[r]) return -1;
    }
  }
  f.last_mid = f.n;
  f.info.reset();
  DBuf<u32> bits(ctx, words), wrank(ctx, words);
  dev_zero(ctx, bits.get(), (size_t)words * sizeof(u32));
  mp.bits = bits.get();
  mp.wrank = wrank.get();
  mp.D = D;
  mp.nblocks = nblk;
  mp.block0 = nblk > 0 ? b0 : 0;
  DBuf<int> root_flag(ctx, f.nblocks);
  dev_zero(ctx, root_flag.get(), (size_t)f.nblocks * sizeof(int));
  DBuf<i32> d_span(ctx, 2 * R);
  copy_h2d(ctx, d_span.get(), span.data(), (size_t)(2 * R) * sizeof(i32));
  int lo_shared = 0, hi_shared = 0;
  for (int r = 0; r < R; r++) {
    if (r == me || nblk == 0) continue;
    if (span[r] <= b0 && b0 <= span[R + r]) lo_shared = 1;
    if (span[r] <= b1 && b1 <= span[R + r]) hi_shared = 1;
  }
  i64 cap = 1 << 20;
  DBuf<u64> rkey(ctx, cap);
  DBuf<u32> rdest(ctx, cap);
  DBuf<unsigned long long> rcount(ctx, 1);
  dev_zero(ctx, rcount.get(), sizeof(unsigned long long));

  /* one closure stage: run `stage` (a kernel that ORs locally and appends the
     cells other ranks need), ship the appended cells, OR what arrives */
  for (int l = D; l >= 1; l--) {
    i64 nsend = 0;
    for (int attempt = 0; attempt < 2; attempt++) {
      MapRemote rm = {d_span.get(), d_span.get() + R, R,         me,
                      lo_shared,    hi_shared,        rkey.get(), rdest.get(),
                      rcount.get(), cap};
      if (l == D) {
        /* stage D: the parents of my own octants */
        MapMarkParentsFn mk = {f.keys.get(), f.fmt, mp, root_flag.get(), rm};
        launch(ctx, f.n, mk, "balance_map_mark");
      } else {
        MapClosureFn cl = {mp, f.tables, l, balance_corner, rm};
        launch(ctx, mp.cells(l - 1), cl, "balance_map_closure");
      }
      unsigned long long h = 0;
      copy_d2h(ctx, &h, rcount.get(), sizeof(h));
      dev_zero(ctx, rcount.get(), sizeof(unsigned long long));
      nsend = (i64)h;
      if (nsend <= cap) break;
      /* the appended cells did not fit: the local ORs are idempotent, rerun */
      cap = nsend + (nsend >> 2);
      rkey.alloc(ctx, cap);
      rdest.alloc(ctx, cap);
    }
    MapU32DestFn dest = {rdest.get()};
    RoutePlan plan;
    make_route(ctx, comm, nsend, dest, plan);
    DBuf<u64> got;
    route_array(ctx, comm, plan, rkey.get(), got);
    MapOrReceivedFn orf = {got.get(), mp};
    launch(ctx, plan.nrecv, orf, "balance_map_received");
  }
  const u64 range[2] = {lo[me], hi[me]};
  if (nblk == 0) {
    DBuf<u64> none;
    f.keys.swap(none);
    f.n = 0;
    f.last_out = 0;
    return check_errors(ctx, "balance_multi");
  }
  const int rc = balance_map_leaves(f, mp, words, wrank, root_flag.get(), range);
  if (rc) return rc;
  return check_errors(ctx, "balance_multi");
}

inline int balance_multi(Forest &f, int balance_corner) {
  Ctx &ctx = *f.ctx;
  Comm &comm = *forest_comm(f);
  const int me = comm.rank;
  f.last_mid = f.n;
  f.info.reset();
  if (unify_depth(f)) return 1;
  const int D = f.fmt.D;
  if (D == 0) {
    f.last_out = f.n;
    return 0;
  }
  {
    const char *mode = getenv("TMR_B200_BALANCE");
    if (!(mode && strcmp(mode, "sort") == 0)) {
      const int rc = balance_multi_map(f, balance_corner);
      if (getenv("TMR_B200_NODES_VERBOSE")) {
        fprintf(stderr, "[tmr_b200] balance (rank %d): %s\n", me,
                rc >= 0 ? "cell bitmaps" : "sorted-array closure");
      }
      if (rc >= 0) return rc;
    }
  }
  const int bbits = f.bbits, nb = f.nblocks;
  DBuf<u64> own_store;
  OwnerMap om = make_owner_map(f, D, own_store);
  DBuf<int> root_flag(ctx, nb);
  dev_zero(ctx, root_flag.get(), (size_t)nb * sizeof(int));

  std::vector<DBuf<u64> > R(D);
  std::vector<i64> nR(D, 0);
  DBuf<u64> pending; /* candidates for the level being closed */
  i64 npending = 0;
  for (int l = D - 1; l >= 0; l--) {
    /* parents (level l) of my own octants of level l+1 */
    ParentOfOctantFn pf = {f.keys.get(), f.fmt, l};
    ParentLevelCountFn pc = {pf, root_flag.get()};
    DBuf<u64> all(ctx, npending + f.n);
    if (npending) copy_d2d(ctx, all.get(), pending.get(), (size_t)npending * sizeof(u64));
    ParentLevelFillFn fl = {pf, all.get() + npending};
    const i64 npar = (i64)scan_apply(ctx, f.n, pc, fl, "balance_parents");
    const i64 nall = npending + npar;
    /* route to the owners of the positions, then sort + dedup */
    LevelOwnerFn lo = {all.get(), l, D, om};
    DBuf<u64> got;
    i64 ngot = route_keys_sparse(ctx, comm, all.get(), nall, lo, got);
    {
      DBuf<u64> alt(ctx, ngot);
      DBuf<u32> v0, v1;
      radix_sort(ctx, got, alt, v0, v1, ngot, 0, bbits + 3 * l);
      ngot = unique_keep_last(ctx, got, alt, v0, v1, ngot, 0);
    }
    R[l].swap(got);
    nR[l] = ngot;
    /* candidates for level l-1 */
    pending.reset();
    npending = 0;
    if (l >= 1 && nR[l] > 0) {
      DBuf<u32> coff(ctx, nR[l]);
      BalanceGen gen = {f.tables, R[l].get(), l, balance_corner};
      BalanceCountFn cnt = {gen};
      npending = (i64)scan_counts(ctx, nR[l], cnt, coff.get(), "balance_cand_count");
      pending.alloc(ctx, npending);
      BalanceFillFn fill = {gen, coff.get(), pending.get()};
      launch(ctx, nR[l], fill, "balance_cand_fill");
    }
  }

  /* children whose parent lives elsewhere are mirrored to the parent's owner */
  std::vector<DBuf<u64> > Rc(D);
  std::vector<i64> nRc(D, 0);
  for (int l = 1; l < D; l++) {
    ParentOwnerFn po = {R[l].get(), l, D, om};
    ForeignParentCountFn fc = {po, me};
    DBuf<u64> fk(ctx, nR[l]);
    DBuf<u32> fd(ctx, nR[l]);
    ForeignParentFillFn ff = {po, me, fk.get(), fd.get()};
    const i64 nf = (i64)scan_apply(ctx, nR[l], fc, ff, "balance_foreign_parents");
    ArrayDestFn ad = {fd.get()};
    RoutePlan plan;
    make_route(ctx, comm, nf, ad, plan);
    DBuf<u64> got;
    route_array(ctx, comm, plan, fk.get(), got);
    const i64 tot = nR[l] + plan.nrecv;
    Rc[l].alloc(ctx, tot);
    if (nR[l]) copy_d2d(ctx, Rc[l].get(), R[l].get(), (size_t)nR[l] * sizeof(u64));
    if (plan.nrecv) {
      copy_d2d(ctx, Rc[l].get() + nR[l], got.get(), (size_t)plan.nrecv * sizeof(u64));
      DBuf<u64> alt(ctx, tot);
      DBuf<u32> v0, v1;
      radix_sort(ctx, Rc[l], alt, v0, v1, tot, 0, bbits + 3 * l);
    }
    nRc[l] = tot;
  }

  /* leaves of the members I own */
  std::vector<DBuf<u32> > loff(D);
  std::vector<i64> nleaf(D, 0);
  i64 total = 0;
  for (int l = 0; l < D; l++) {
    if (nR[l] == 0) continue;
    loff[l].alloc(ctx, nR[l]);
    LeafGen g = {R[l].get(), (l + 1 < D) ? Rc[l + 1].get() : NULL,
                 (l + 1 < D) ? nRc[l + 1] : 0, l, f.fmt};
    LeafCountFn c = {g};
    nleaf[l] = (i64)scan_counts(ctx, nR[l], c, loff[l].get(), "balance_leaf_count");
    total += nleaf[l];
  }
  DBuf<u32> root_off(ctx, nb);
  RootLeafCountFn rc = {root_flag.get(), R[0].get(), nR[0]};
  const i64 nroot = (i64)scan_counts(ctx, nb, rc, root_off.get(), "balance_root_count");
  total += nroot;
  DBuf<u64> out(ctx, total);
  i64 base = 0;
  for (int l = 0; l < D; l++) {
    if (nR[l] == 0) continue;
    LeafGen g = {R[l].get(), (l + 1 < D) ? Rc[l + 1].get() : NULL,
                 (l + 1 < D) ? nRc[l + 1] : 0, l, f.fmt};
    LeafFillFn fl = {g, loff[l].get(), out.get() + base};
    launch(ctx, nR[l], fl, "balance_leaf_fill");
    base += nleaf[l];
  }
  RootLeafFillFn rf = {rc, root_off.get(), f.fmt, out.get() + base};
  launch(ctx, nb, rf, "balance_root_fill");

  /* leaves go to the owner of their own position */
  LeafOwnerFn lof = {out.get(), om};
  DBuf<u64> mine;
  const i64 nmine = route_keys_sparse(ctx, comm, out.get(), total, lof, mine);
  if (nmine >= (1LL << 31)) {
    fprintf(stderr, "TMROctForest Error: balance() leaves %lld octants on one "
                    "rank (int32 index limit)\n", (long long)nmine);
    return 1;
  }
  {
    DBuf<u64> alt(ctx, nmine);
    DBuf<u32> v0, v1;
    radix_sort(ctx, mine, alt, v0, v1, nmine, 5, f.fmt.total_bits());
  }
  f.keys.swap(mine);
  f.n = nmine;
  f.last_out = f.n;
  return check_errors(ctx, "balance_multi");
}

}  // namespace tmrgpu

#endif
