/*
  ops_route.h -- ownership lookup and all-to-all-v routing of device arrays
  (the GPU form of distributeOctants / sendOctants, reference
  src/TMROctForest.cpp:2334-2509).
*/
#ifndef TMRGPU_OPS_ROUTE_H
#define TMRGPU_OPS_ROUTE_H

#include "ops_balance.h"
#include "ops_balance_map.h"

namespace tmrgpu {

/* ---- ownership ------------------------------------------------------------- */
/* position key at Morton depth Dp: [block | Morton_Dp] */
struct OwnerMap {
  const u64 *pos; /* device, R entries, positions of owners[] at depth Dp */
  int R;
  /* this rank's own slice [my_lo, my_hi), held by value: the common answer
     "mine" costs two compares and no memory access */
  int me;
  u64 my_lo, my_hi;
  TMR_HD int owner(u64 p) const {
    if (p >= my_lo && p < my_hi) return me;
    int r = 0;
    while (r < R - 1 && pos[r + 1] <= p) r++;
    return r;
  }
};

/* owners[] as position keys at depth Dp (sentinel = one past the block) */
inline void owner_positions(const Forest &f, int Dp, std::vector<u64> &out) {
  const int R = (int)f.owners.size();
  out.resize(R);
  for (int r = 0; r < R; r++) {
    const Oct24 &o = f.owners[r];
    if (o.x >= kHmax || o.y >= kHmax || o.z >= kHmax) {
      out[r] = ((u64)(u32)(o.block + 1)) << (3 * Dp);
    } else {
      const int s = kMaxLevel - Dp;
      out[r] = ((u64)(u32)o.block << (3 * Dp)) |
               morton3((u32)o.x >> s, (u32)o.y >> s, (u32)o.z >> s);
    }
  }
}

inline OwnerMap make_owner_map(Forest &f, int Dp, DBuf<u64> &store) {
  std::vector<u64> h;
  owner_positions(f, Dp, h);
  store.alloc(*f.ctx, (i64)h.size());
  copy_h2d(*f.ctx, store.get(), h.data(), h.size() * sizeof(u64));
  const int R = (int)h.size();
  const int me = (forest_comm(f) && R > 1) ? forest_comm(f)->rank : 0;
  OwnerMap m = {store.get(), R, me, me > 0 ? h[me] : 0ULL,
                me + 1 < R ? h[me + 1] : ~0ULL};
  return m;
}

/* ---- routing ---------------------------------------------------------------- */
struct RoutePlan {
  DBuf<u32> idx;                 /* send position -> source index */
  std::vector<i64> send_off, recv_off;
  i64 nsend, nrecv;
  RoutePlan() : nsend(0), nrecv(0) {}
};

template <class DestFn>
struct DestKeyFn {
  DestFn dest;
  u64 *dk;
  u32 *idx;
  TMR_HD void operator()(i64 i) const {
    dk[i] = (u64)(u32)dest(i);
    idx[i] = (u32)i;
  }
};

/* bounds[r] = first item for rank r (r = 0..R), counts[r] = items for rank r;
   the buffer is laid out [bounds R+1 | counts R | all ranks' counts R*R] */
struct DestBoundsFn {
  const u64 *dk;
  i64 n;
  int R;
  i64 *buf;
  TMR_HD void operator()(i64 r) const {
    const i64 lo = n > 0 ? lower_bound_u64(dk, n, (u64)r) : 0;
    buf[r] = lo;
    if (r < R) {
      const i64 hi = n > 0 ? lower_bound_u64(dk, n, (u64)(r + 1)) : 0;
      buf[R + 1 + r] = hi - lo;
    }
  }
};

/* group items by destination rank and agree on the all-to-all-v layout.  The
   per-destination counts never visit the host on their way to the other
   ranks: bounds kernel -> ncclAllGather (device to device) -> one read-back
   of bounds and everybody's counts together (was: three blocking round
   trips per route) */
template <class DestFn>
void make_route(Ctx &ctx, Comm &comm, i64 n, DestFn dest, RoutePlan &plan) {
  const int R = comm.size;
  plan.idx.alloc(ctx, n);
  plan.send_off.assign(R + 1, 0);
  plan.recv_off.assign(R + 1, 0);
  DBuf<u64> dk, dk_alt;
  if (n > 0) {
    dk.alloc(ctx, n);
    dk_alt.alloc(ctx, n);
    DBuf<u32> idx_alt(ctx, n);
    DestKeyFn<DestFn> k = {dest, dk.get(), plan.idx.get()};
    launch(ctx, n, k, "route_dest");
    int bits = 1;
    while ((1 << bits) < R) bits++;
    radix_sort(ctx, dk, dk_alt, plan.idx, idx_alt, n, 0, bits);
  }
  const i64 words = (R + 1) + R + (i64)R * R;
  DBuf<i64> d_buf(ctx, words);
  DestBoundsFn b = {dk.get(), n, R, d_buf.get()};
  launch(ctx, R + 1, b, "route_bounds");
  comm.allgather_dev(ctx, d_buf.get() + (R + 1), d_buf.get() + (2 * R + 1),
                     (size_t)R * sizeof(i64));
  std::vector<i64> h(words);
  copy_d2h(ctx, h.data(), d_buf.get(), (size_t)words * sizeof(i64));
  for (int r = 0; r <= R; r++) plan.send_off[r] = h[r];
  for (int r = 0; r < R; r++) {
    plan.recv_off[r + 1] = plan.recv_off[r] + h[(2 * R + 1) + (size_t)r * R + comm.rank];
  }
  plan.nsend = n;
  plan.nrecv = plan.recv_off[R];
}

template <class T>
struct GatherFn {
  const T *src;
  const u32 *idx;
  T *dst;
  TMR_HD void operator()(i64 i) const { dst[i] = src[idx[i]]; }
};

template <class T>
struct ScatterFn {
  const T *src;
  const u32 *idx;
  T *dst;
  TMR_HD void operator()(i64 i) const { dst[idx[i]] = src[i]; }
};

/* ship src[] according to the plan; out holds plan.nrecv items grouped by
   source rank (recv_off) */
template <class T>
void route_array(Ctx &ctx, Comm &comm, const RoutePlan &plan, const T *src,
                 DBuf<T> &out) {
  DBuf<T> send(ctx, plan.nsend);
  GatherFn<T> g = {src, plan.idx.get(), send.get()};
  launch(ctx, plan.nsend, g, "route_gather");
  out.alloc(ctx, plan.nrecv);
  comm.alltoallv(ctx, send.get(), plan.send_off.data(), out.get(),
                 plan.recv_off.data(), sizeof(T));
}

/* replies travel the plan backwards: reply[] is aligned with what this rank
   RECEIVED; out[] (size nsend) is aligned with the original item order */
template <class T>
void route_back(Ctx &ctx, Comm &comm, const RoutePlan &plan, const T *reply,
                DBuf<T> &out) {
  DBuf<T> tmp(ctx, plan.nsend);
  comm.alltoallv(ctx, reply, plan.recv_off.data(), tmp.get(),
                 plan.send_off.data(), sizeof(T));
  out.alloc(ctx, plan.nsend);
  ScatterFn<T> s = {tmp.get(), plan.idx.get(), out.get()};
  launch(ctx, plan.nsend, s, "route_scatter_back");
}

/* ---- sparse routing ----------------------------------------------------------
   Almost every key already lives on its owner: keep those in place (one
   compaction) and ship only the few that do not, instead of grouping the whole
   array by destination. */
template <class DestFn>
struct ForeignCountFn {
  DestFn dest;
  int me;
  TMR_HD u32 operator()(i64 i) const { return dest(i) != me ? 1u : 0u; }
};
/* one pass splits the array: foreign keys (with their destination) are packed
   at their foreign rank `o`, the keys that stay at i - o */
template <class DestFn>
struct SplitKeysFn {
  DestFn dest;
  int me;
  const u64 *keys;
  u64 *foreign_keys;
  u32 *foreign_dest;
  u64 *local_keys;
  TMR_HD void operator()(i64 i, u32 o) const {
    const int d = dest(i);
    if (d != me) {
      foreign_keys[o] = keys[i];
      foreign_dest[o] = (u32)d;
    } else {
      local_keys[i - (i64)o] = keys[i];
    }
  }
};
struct DestArrayFn {
  const u32 *dest;
  TMR_HD int operator()(i64 i) const { return (int)dest[i]; }
};

/* out = keys that stay here (original order) followed by the keys received
   from the other ranks; returns the new count */
template <class DestFn>
i64 route_keys_sparse(Ctx &ctx, Comm &comm, const u64 *keys, i64 n, DestFn dest,
                      DBuf<u64> &out) {
  const int me = comm.rank;
  DBuf<u64> fk(ctx, n), loc(ctx, n);
  DBuf<u32> fd(ctx, n);
  ForeignCountFn<DestFn> fc = {dest, me};
  SplitKeysFn<DestFn> sp = {dest, me, keys, fk.get(), fd.get(), loc.get()};
  const i64 nf = (i64)scan_apply(ctx, n, fc, sp, "route_split");
  DestArrayFn da = {fd.get()};
  RoutePlan plan;
  make_route(ctx, comm, nf, da, plan);
  DBuf<u64> got;
  route_array(ctx, comm, plan, fk.get(), got);
  const i64 nloc = n - nf;
  if (plan.nrecv == 0) {
    out.swap(loc);
    out.set_size(nloc);
    return nloc;
  }
  out.alloc(ctx, nloc + plan.nrecv);
  copy_d2d(ctx, out.get(), loc.get(), (size_t)nloc * sizeof(u64));
  copy_d2d(ctx, out.get() + nloc, got.get(), (size_t)plan.nrecv * sizeof(u64));
  return nloc + plan.nrecv;
}

/* global max / sum / or of a small host value */
inline i64 global_max(Ctx &ctx, Comm &comm, i64 v) {
  std::vector<i64> all(comm.size);
  comm.allgather_host(ctx, &v, all.data(), sizeof(i64));
  i64 m = all[0];
  for (int r = 1; r < comm.size; r++) m = all[r] > m ? all[r] : m;
  return m;
}

}  // namespace tmrgpu

#endif
