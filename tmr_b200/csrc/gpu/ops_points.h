/*
  ops_points.h -- node locations on the device for forests whose trees are
  trilinear hexahedra (TMRTrilinearVolume), replacing the element loop of
  reference TMROctForest::evaluateNodeLocations (src/TMROctForest.cpp:5636-5671):
  every local node takes the location evaluated through the FIRST element (in
  element order) and local slot that reference it -- the reference's flags[]
  logic -- at the parametric point u + 0.5 d (1 + knot[i]) of that element.
*/
#ifndef TMRGPU_OPS_POINTS_H
#define TMRGPU_OPS_POINTS_H

#include "ops_nodes.h"

namespace tmrgpu {

TMR_HD i64 lower_bound_i32(const int *a, i64 n, int v) {
  i64 lo = 0, hi = n;
  while (lo < hi) {
    const i64 mid = lo + ((hi - lo) >> 1);
    if (a[mid] < v) {
      lo = mid + 1;
    } else {
      hi = mid;
    }
  }
  return lo;
}

/* index of a node number in the sorted list getNodeNumbers() hands out
   (reference getLocalNodeNumber :1486-1495): on one rank that list is the
   range -Nd .. owned-1 */
struct LocalIndexOf {
  const int *sorted; /* NULL on one rank */
  i64 n;
  int num_dep;
  TMR_HD i64 operator()(int number) const {
    if (!sorted) return (i64)number + num_dep;
    const i64 i = lower_bound_i32(sorted, n, number);
    return (i < n && sorted[i] == number) ? i : -1;
  }
};

struct PointFirstFn {
  const int *conn;
  LocalIndexOf index;
  u32 *first;
  TMR_HD void operator()(i64 code) const {
    const i64 i = index(conn[code]);
    if (i >= 0) TMR_ATOMIC_MIN_I32(reinterpret_cast<int *>(&first[i]), (int)code);
  }
};

struct PointEvalFn {
  const u64 *keys;
  KeyFmt fmt;
  int order;
  double knots[kMaxOrder];
  const double *corners; /* [nblocks][8][3] */
  const u32 *first;
  double *X; /* [n][3] */
  TMR_HD void operator()(i64 i) const {
    const u32 code = first[i];
    double p[3] = {0.0, 0.0, 0.0};
    if (code != 0x7fffffffu) {
      const int npe = order * order * order;
      const i64 e = (i64)(code / (u32)npe);
      const int j = (int)(code % (u32)npe);
      const int ii = j % order, jj = (j / order) % order, kk = j / (order * order);
      i32 block, x, y, z;
      int level;
      fmt.decode(keys[e], &block, &x, &y, &z, &level);
      const double d = param_coordinate(1 << (kMaxLevel - level));
      const double u = param_coordinate(x), v = param_coordinate(y), w = param_coordinate(z);
      trilinear_point(corners + 24 * (size_t)block, u + 0.5 * d * (1.0 + knots[ii]),
                      v + 0.5 * d * (1.0 + knots[jj]), w + 0.5 * d * (1.0 + knots[kk]), p);
    }
    X[3 * i] = p[0];
    X[3 * i + 1] = p[1];
    X[3 * i + 2] = p[2];
  }
};

/* h_corners: [nblocks][8][3] doubles; h_X: [num_local_nodes][3] doubles, in the
   order of the sorted node numbers (the reference's X array) */
inline int eval_trilinear_points(Forest &f, const double *h_corners, double *h_X) {
  Ctx &ctx = *f.ctx;
  NodeData &nd = f.nodes;
  if (!nd.valid) return 1;
  const i64 n = nd.num_local_nodes;
  if (n == 0) return 0;
  const i64 nc = nd.num_elements * (i64)nd.order * nd.order * nd.order;
  DBuf<double> corners(ctx, (i64)f.nblocks * 24);
  copy_h2d(ctx, corners.get(), h_corners, (size_t)f.nblocks * 24 * sizeof(double));
  DBuf<int> sorted;
  if (forest_comm(f) && build_sorted_numbers(f, sorted)) return 1;
  DBuf<u32> first(ctx, n);
  FillIntFn fill = {reinterpret_cast<int *>(first.get()), 0x7fffffff};
  launch(ctx, n, fill, "points_first_init");
  LocalIndexOf ix = {sorted.get(), n, (int)nd.num_dep_nodes};
  PointFirstFn pf = {nd.conn.get(), ix, first.get()};
  launch(ctx, nc, pf, "points_first_touch");
  DBuf<double> X(ctx, 3 * n);
  PointEvalFn pe;
  pe.keys = f.keys.get();
  pe.fmt = f.fmt;
  pe.order = nd.order;
  for (int i = 0; i < kMaxOrder; i++) pe.knots[i] = nd.knots[i];
  pe.corners = corners.get();
  pe.first = first.get();
  pe.X = X.get();
  launch(ctx, n, pe, "points_eval_trilinear");
  copy_d2h(ctx, h_X, X.get(), (size_t)(3 * n) * sizeof(double));
  return check_errors(ctx, "eval_trilinear_points");
}

}  // namespace tmrgpu

#endif
