/*
  ops_interp.h -- createInterpolation (multigrid prolongation rows in fp64 CSR),
  batched findEnclosing, and the generic TMROctantArray sort / contains.

  createInterpolation replaces reference src/TMROctForest.cpp:6611-6793:
    first_touch   atomicMin over (element, local node) of every owned fine node
                  -> the (element, slot) the reference's serial loop would use
    row scan      rows are numbered in that first-touch order (no sort needed:
                  a scan over the element-major code space)
    row_count /   per row: findEnclosing (the reference's exact binary-search +
    row_fill      forward scan, :6228-6377), tensor Lagrange weights with the
                  coarse-face collapse rule (:6501-6548), dependent-node
                  expansion, sort by column + sum duplicates (:6574-6590)
*/
#ifndef TMRGPU_OPS_INTERP_H
#define TMRGPU_OPS_INTERP_H

#include "ops_nodes.h"
#include "ops_route.h"

namespace tmrgpu {

/* the reference's comparePosition on raw coordinates
   (src/TMROctant.cpp:210-239) */
TMR_HD int compare_position(i32 b0, i32 x0, i32 y0, i32 z0, i32 b1, i32 x1,
                            i32 y1, i32 z1) {
  if (b0 != b1) return b0 - b1;
  const u32 xx = (u32)(x0 ^ x1), yx = (u32)(y0 ^ y1), zx = (u32)(z0 ^ z1);
  const u32 sor = xx | yx | zx;
  int discrim;
  if (xx > (sor ^ xx)) {
    discrim = (x0 > x1) - (x0 < x1);
  } else if (yx > (sor ^ yx)) {
    discrim = (y0 > y1) - (y0 < y1);
  } else {
    discrim = (z0 > z1) - (z0 < z1);
  }
  return discrim;
}

/* findEnclosing (reference :6228-6377) against a sorted element key array */
struct EnclosingSearch {
  const u64 *ckeys;
  i64 cn;
  KeyFmt cfmt;

  TMR_HD i64 find(i32 block, i32 x, i32 y, i32 z, int level, int info,
                  int order, const double *knots) const {
    if (cn <= 0) return -1;
    const i32 h = 1 << (kMaxLevel - level);
    const int ii = info % order;
    const int jj = (info % (order * order)) / order;
    const int kk = info / (order * order);
    i32 xi = -1, yi = -1, zi = -1;
    if (ii == 0 || ii == order - 1) {
      xi = x + (ii / (order - 1)) * h;
    } else if (order % 2 == 1 && ii == order / 2) {
      xi = x + h / 2;
    }
    if (jj == 0 || jj == order - 1) {
      yi = y + (jj / (order - 1)) * h;
    } else if (order % 2 == 1 && jj == order / 2) {
      yi = y + h / 2;
    }
    if (kk == 0 || kk == order - 1) {
      zi = z + (kk / (order - 1)) * h;
    } else if (order % 2 == 1 && kk == order / 2) {
      zi = z + h / 2;
    }
    const double xd = x + 0.5 * h * (1.0 + knots[ii]);
    const double yd = y + 0.5 * h * (1.0 + knots[jj]);
    const double zd = z + 0.5 * h * (1.0 + knots[kk]);

    i64 low = 0, high = cn - 1;
    i64 mid = low + (high - low) / 2;
    i32 mb, mx, my, mz;
    int ml;
    while (mid != low) {
      cfmt.decode(ckeys[mid], &mb, &mx, &my, &mz, &ml);
      const i32 hm = 1 << (kMaxLevel - ml);
      if (mb == block && x >= mx && x < mx + hm && y >= my && y < my + hm &&
          z >= mz && z < mz + hm) {
        break;
      }
      const int stat = compare_position(mb, mx, my, mz, block, x, y, z);
      if (stat == 0) {
        break;
      } else if (stat < 0) {
        low = mid + 1;
      } else {
        high = mid - 1;
      }
      mid = low + (high - low) / 2; /* C division truncates toward zero */
    }
    while (mid < cn) {
      cfmt.decode(ckeys[mid], &mb, &mx, &my, &mz, &ml);
      if (compare_position(mb, mx, my, mz, block, x + h, y + h, z + h) > 0) {
        break;
      }
      if (mb == block) {
        const i32 hm = 1 << (kMaxLevel - ml);
        const bool okx = (xi >= 0) ? (mx <= xi && xi <= mx + hm)
                                   : (mx <= xd && xd <= mx + hm);
        const bool oky = (yi >= 0) ? (my <= yi && yi <= my + hm)
                                   : (my <= yd && yd <= my + hm);
        const bool okz = (zi >= 0) ? (mz <= zi && zi <= mz + hm)
                                   : (mz <= zd && zd <= mz + hm);
        if (okx && oky && okz) return mid;
      }
      mid++;
    }
    return -1;
  }
};

/* the point findEnclosing uses to guess the owner of a node it did not find
   (reference :6354-6374), as a position key at depth D */
TMR_HD u64 miss_position(i32 block, i32 x, i32 y, i32 z, int level, int info,
                         int order, const double *knots, int D) {
  const i32 h = 1 << (kMaxLevel - level);
  const int ijk[3] = {info % order, (info % (order * order)) / order,
                      info / (order * order)};
  const i32 base[3] = {x, y, z};
  i32 n[3];
  for (int a = 0; a < 3; a++) {
    const int i = ijk[a];
    i32 ci = -1;
    if (i == 0 || i == order - 1) {
      ci = base[a] + (i / (order - 1)) * h;
    } else if (order % 2 == 1 && i == order / 2) {
      ci = base[a] + h / 2;
    }
    const double cd = base[a] + 0.5 * h * (1.0 + knots[i]);
    n[a] = ci < 0 ? (i32)cd : ci;
    if (n[a] == 0) {
      n[a] += 1;
    } else if (n[a] == kHmax) {
      n[a] -= 1;
    }
  }
  const int s = kMaxLevel - D;
  return ((u64)(u32)block << (3 * D)) |
         morton3((u32)n[0] >> s, (u32)n[1] >> s, (u32)n[2] >> s);
}

struct FindEnclosingFn {
  EnclosingSearch s;
  const Oct24 *nodes;
  int order;
  double knots[kMaxOrder];
  int *out;
  TMR_HD void operator()(i64 i) const {
    const Oct24 q = nodes[i];
    out[i] = (int)s.find(q.block, q.x, q.y, q.z, q.level, q.info, order, knots);
  }
};

inline int find_enclosing_batch(Forest &f, int order, const double *knots,
                                const Oct24 *h_nodes, i64 n, int *h_out) {
  Ctx &ctx = *f.ctx;
  if (n <= 0) return 0;
  if (order > 4) {
    fprintf(stderr, "TMROctForest Error: findEnclosing order %d unsupported\n",
            order);
    return 1;
  }
  DBuf<Oct24> d_nodes(ctx, n);
  DBuf<int> d_out(ctx, n);
  copy_h2d(ctx, d_nodes.get(), h_nodes, (size_t)n * sizeof(Oct24));
  FindEnclosingFn fe;
  fe.s.ckeys = f.keys.get();
  fe.s.cn = f.n;
  fe.s.cfmt = f.fmt;
  fe.nodes = d_nodes.get();
  fe.order = order;
  for (int i = 0; i < kMaxOrder; i++) fe.knots[i] = (i < order) ? knots[i] : 0.0;
  fe.out = d_out.get();
  launch(ctx, n, fe, "find_enclosing");
  copy_d2h(ctx, h_out, d_out.get(), (size_t)n * sizeof(int));
  return check_errors(ctx, "find_enclosing");
}

/* ---- createInterpolation --------------------------------------------------- */
static const u64 kNoTouch = ~0ULL;

struct FirstTouchFn {
  const int *conn;
  int lo, hi; /* owned range [lo, hi) */
  u64 *first;
  TMR_HD void operator()(i64 code) const {
    const int c = conn[code];
    if (c >= lo && c < hi) TMR_ATOMIC_MIN_U64(&first[c - lo], (u64)code);
  }
};

struct IsFirstFn {
  const int *conn;
  int lo, hi;
  const u64 *first;
  TMR_HD u32 operator()(i64 code) const {
    const int c = conn[code];
    return (c >= lo && c < hi && first[c - lo] == (u64)code) ? 1u : 0u;
  }
};

struct RowCodeFn {
  IsFirstFn is;
  const u32 *row_of;
  u64 *row_code;
  int *rows;
  TMR_HD void operator()(i64 code) const {
    if (is(code)) {
      row_code[row_of[code]] = (u64)code;
      rows[row_of[code]] = is.conn[code];
    }
  }
};


struct InterpRow {
  /* fine */
  const u64 *fkeys;
  KeyFmt ffmt;
  int forder;
  double fknots[kMaxOrder];
  /* coarse */
  EnclosingSearch s;
  int corder;
  double cknots[kMaxOrder];
  const int *cconn;
  const int *cdep_ptr;
  const int *cdep_conn;
  const double *cdep_w;
  int bernstein; /* both meshes use Bernstein points (reference :6434-6500) */
  int *overflow; /* set if a row ever exceeded its buffer (see RowCap) */

  /* per-axis weights with the collapse rule (reference :6501-6548) */
  TMR_HD void axis(int i, i32 nx, i32 h, i32 ox, i32 hc, int *start, int *end,
                   double *N) const {
    *start = 0;
    *end = corder;
    if ((i == 0 && ox == nx) || (i == forder - 1 && ox == nx + h)) {
      *start = 0;
      *end = 1;
      N[0] = 1.0;
    } else if ((i == 0 && ox + hc == nx) ||
               (i == forder - 1 && ox + hc == nx + h)) {
      *start = corder - 1;
      *end = corder;
      N[corder - 1] = 1.0;
    } else if (bernstein && forder != corder) {
      bernstein_elevation_weights(corder, i, N);
    } else {
      const double u =
          -1.0 + 2.0 * (nx + 0.5 * h * (1.0 + fknots[i]) - ox) / hc;
      if (bernstein) {
        bernstein_basis(corder, u, N);
      } else {
        lagrange_basis(corder, u, cknots, N);
      }
    }
  }

  /* enclosing coarse element of node j of the fine element with key fkey */
  TMR_HD i64 locate(u64 fkey, int j) const {
    i32 block, x, y, z;
    int level;
    ffmt.decode(fkey, &block, &x, &y, &z, &level);
    return s.find(block, x, y, z, level, j, forder, fknots);
  }

  /* builds the sorted, merged row of node j of fine element fkey inside coarse
     element t; returns its length */
  template <int kCOrder> /* sizes the coarse basis arrays */
  TMR_HD int build(u64 fkey, int j, i64 t, int *idx, double *w, int cap) const {
    i32 block, x, y, z;
    int level;
    ffmt.decode(fkey, &block, &x, &y, &z, &level);
    i32 cb, ox, oy, oz;
    int cl;
    s.cfmt.decode(s.ckeys[t], &cb, &ox, &oy, &oz, &cl);
    const i32 h = 1 << (kMaxLevel - level);
    const i32 hc = 1 << (kMaxLevel - cl);
    const int i0 = j % forder, j0 = (j % (forder * forder)) / forder,
              k0 = j / (forder * forder);
    double Nu[kCOrder], Nv[kCOrder], Nw[kCOrder];
    int is, ie, js, je, ks, ke;
    axis(i0, x, h, ox, hc, &is, &ie, Nu);
    axis(j0, y, h, oy, hc, &js, &je, Nv);
    axis(k0, z, h, oz, hc, &ks, &ke, Nw);
    const int cnpe = corder * corder * corder;
    const int *c = cconn + t * cnpe;
    int n = 0;
    for (int kk = ks; kk < ke; kk++) {
      for (int jj = js; jj < je; jj++) {
        for (int ii = is; ii < ie; ii++) {
          const int off = ii + jj * corder + kk * corder * corder;
          const double weight = Nu[ii] * Nv[jj] * Nw[kk];
          if (c[off] >= 0) {
            n = insert(idx, w, n, cap, c[off], weight);
          } else {
            const int dn = -c[off] - 1;
            for (int jp = cdep_ptr[dn]; jp < cdep_ptr[dn + 1]; jp++) {
              n = insert(idx, w, n, cap, cdep_conn[jp], weight * cdep_w[jp]);
            }
          }
        }
      }
    }
    return n;
  }

  /* sorted insert with duplicate columns summed
     (TMRIndexWeight::uniqueSort, reference src/TMRBase.h:112-135) */
  TMR_HD int insert(int *idx, double *w, int n, int cap, int col,
                    double val) const {
    int p = n;
    while (p > 0 && idx[p - 1] > col) p--;
    if (p > 0 && idx[p - 1] == col) {
      w[p - 1] += val;
      return n;
    }
    if (n >= cap) {
      TMR_ATOMIC_OR_I32(overflow, 1);
      return n;
    }
    for (int q = n; q > p; q--) {
      idx[q] = idx[q - 1];
      w[q] = w[q - 1];
    }
    idx[p] = col;
    w[p] = val;
    return n + 1;
  }
};

/* one request = one row to build: fine element key, local node, fine global
   node number (the row id), enclosing coarse element */
struct RowRequests {
  const u64 *fkey;
  const int *j;
  const i64 *t;
};

struct InterpLocateFn {
  InterpRow r;
  const u64 *row_code;
  int npe;
  u64 *fkey;
  int *j;
  i64 *t;
  TMR_HD void operator()(i64 row) const {
    const u64 code = row_code[row];
    const i64 e = (i64)(code / (u64)npe);
    const int jj = (int)(code % (u64)npe);
    fkey[row] = r.fkeys[e];
    j[row] = jj;
    t[row] = r.locate(r.fkeys[e], jj);
  }
};

struct InterpLocateRecvFn {
  InterpRow r;
  const u64 *fkey;
  const u64 *payload; /* (j << 32) | fine node number */
  int *j;
  int *num;
  i64 *t;
  int *missing;
  TMR_HD void operator()(i64 i) const {
    const int jj = (int)(payload[i] >> 32);
    j[i] = jj;
    num[i] = (int)(u32)payload[i];
    t[i] = r.locate(fkey[i], jj);
    if (t[i] < 0) TMR_ATOMIC_OR_I32(missing, 1);
  }
};

/* Capacity of the merged row buffer.  The reference allocates corder^5 (every
   coarse node dependent with a full face stencil, :6637), but the buffer here
   is kept sorted and duplicate-free at every insertion, so it never holds more
   than the DISTINCT columns a row can name: the coarse element's own nodes
   (corder^3) plus the nodes of the parent's faces and edges its dependent
   nodes hang on -- at most the 3 faces and 3 edges meeting at the child's
   corner (3 corder^2 + 3 corder).  An overflow would be a logic error; it is
   caught (the entry is dropped and the error flag raised), never written past
   the buffer. */
template <int kCOrder>
struct RowCap {
  static const int value =
      kCOrder * kCOrder * kCOrder + 3 * kCOrder * kCOrder + 3 * kCOrder + 2;
};

/* row lengths, one thread per row (a plain launch: the build is far too heavy
   to run inside the scan kernel, where each thread would do 8 of them) */
template <int kCOrder>
struct InterpCountFn {
  InterpRow r;
  RowRequests q;
  u32 *count;
  TMR_HD void operator()(i64 row) const {
    if (q.t[row] < 0) {
      count[row] = 0;
      return;
    }
    int idx[RowCap<kCOrder>::value];
    double w[RowCap<kCOrder>::value];
    count[row] = (u32)r.template build<kCOrder>(q.fkey[row], q.j[row], q.t[row],
                                                idx, w, RowCap<kCOrder>::value);
  }
};
struct StoredCountFn {
  const u32 *count;
  TMR_HD u32 operator()(i64 i) const { return count[i]; }
};

template <int kCOrder>
struct InterpFillFn {
  InterpRow r;
  RowRequests q;
  int *cols;
  double *vals;
  TMR_HD void operator()(i64 row, u32 o) const {
    if (q.t[row] < 0) return;
    int idx[RowCap<kCOrder>::value];
    double w[RowCap<kCOrder>::value];
    const int n = r.template build<kCOrder>(q.fkey[row], q.j[row], q.t[row], idx,
                                            w, RowCap<kCOrder>::value);
    for (int k = 0; k < n; k++) {
      cols[o + k] = idx[k];
      vals[o + k] = w[k];
    }
  }
};

struct RowPtrFn {
  const u32 *off;
  i64 nrows;
  u32 total;
  int *rowp;
  TMR_HD void operator()(i64 r) const {
    rowp[r] = (r < nrows) ? (int)off[r] : (int)total;
  }
};

template <int kCOrder>
struct InterpFillPlaceFn {
  InterpFillFn<kCOrder> fill;
  const u32 *off;
  TMR_HD void operator()(i64 row) const { fill(row, off[row]); }
};

/* multi-rank: rows whose coarse element is not on this rank */
struct MissCountFn {
  const i64 *t;
  TMR_HD u32 operator()(i64 i) const { return t[i] < 0 ? 1u : 0u; }
};
struct FoundCountFn {
  const i64 *t;
  TMR_HD u32 operator()(i64 i) const { return t[i] >= 0 ? 1u : 0u; }
};

struct MissFillFn {
  const u64 *fkey;
  const int *j;
  const int *num;
  const i64 *t;
  KeyFmt ffmt;
  int forder;
  double fknots[kMaxOrder];
  OwnerMap om;
  int D;
  u64 *out_key;
  u64 *out_payload;
  u32 *out_dest;
  TMR_HD void operator()(i64 i, u32 o) const {
    if (t[i] >= 0) return;
    i32 block, x, y, z;
    int level;
    ffmt.decode(fkey[i], &block, &x, &y, &z, &level);
    out_key[o] = fkey[i];
    out_payload[o] = ((u64)(u32)j[i] << 32) | (u64)(u32)num[i];
    out_dest[o] = (u32)om.owner(
        miss_position(block, x, y, z, level, j[i], forder, fknots, D));
  }
};

struct FoundFillFn {
  const u64 *fkey;
  const int *j;
  const int *num;
  const i64 *t;
  u64 *o_fkey;
  int *o_j;
  int *o_num;
  i64 *o_t;
  TMR_HD void operator()(i64 i, u32 o) const {
    if (t[i] < 0) return;
    o_fkey[o] = fkey[i];
    o_j[o] = j[i];
    o_num[o] = num[i];
    o_t[o] = t[i];
  }
};

inline int create_interp(Forest &fine, Forest &coarse) {
  Ctx &ctx = *fine.ctx;
  Comm *comm = forest_comm(fine);
  NodeData &fn = fine.nodes;
  NodeData &cn = coarse.nodes;
  InterpData &I = fine.interp;
  I.clear();
  if (!fn.valid || !cn.valid) {
    fprintf(stderr,
            "TMROctForest Error: createInterpolation needs nodes on both "
            "forests\n");
    return 1;
  }
  if (fn.interp_type != cn.interp_type) {
    fprintf(stderr,
            "TMROctForest Error: Interpolation types between meshes are not "
            "identical\n");
  }
  const int npe = fn.order * fn.order * fn.order;
  const i64 nc = fn.num_elements * npe;
  const int lo = fn.node_range_start;
  const int hi = lo + (int)fn.num_owned_nodes;
  const i64 nown = fn.num_owned_nodes;

  InterpRow r;
  r.fkeys = fine.keys.get();
  r.ffmt = fine.fmt;
  r.forder = fn.order;
  r.corder = cn.order;
  for (int i = 0; i < kMaxOrder; i++) {
    r.fknots[i] = fn.knots[i];
    r.cknots[i] = cn.knots[i];
  }
  r.s.ckeys = coarse.keys.get();
  r.s.cn = coarse.n;
  r.s.cfmt = coarse.fmt;
  r.cconn = cn.conn.get();
  r.cdep_ptr = cn.dep_ptr.get();
  r.cdep_conn = cn.dep_conn.get();
  r.cdep_w = cn.dep_weights.get();
  r.bernstein = (fn.interp_type == 2 && cn.interp_type == 2) ? 1 : 0;
  DBuf<int> row_overflow(ctx, 1);
  dev_zero(ctx, row_overflow.get(), sizeof(int));
  r.overflow = row_overflow.get();
  if (r.bernstein && fn.order - cn.order > 1) {
    fprintf(stderr,
            "TMROctForest Error: Mesh order difference across grids should be "
            "1\n");
  }

  /* rows of this rank's owned fine nodes, in first-touch order */
  i64 nrows = 0;
  DBuf<u64> q_fkey;
  DBuf<int> q_j, q_num;
  DBuf<i64> q_t;
  if (nown > 0 && nc > 0) {
    DBuf<u64> first(ctx, nown);
    dev_fill_ff(ctx, first.get(), (size_t)nown * sizeof(u64));
    FirstTouchFn ft = {fn.conn.get(), lo, hi, first.get()};
    launch(ctx, nc, ft, "interp_first_touch");
    DBuf<u32> row_of(ctx, nc);
    IsFirstFn isf = {fn.conn.get(), lo, hi, first.get()};
    nrows = (i64)scan_counts(ctx, nc, isf, row_of.get(), "interp_row_scan");
    DBuf<u64> row_code(ctx, nrows);
    q_num.alloc(ctx, nrows);
    RowCodeFn rcf = {isf, row_of.get(), row_code.get(), q_num.get()};
    launch(ctx, nc, rcf, "interp_row_codes");
    q_fkey.alloc(ctx, nrows);
    q_j.alloc(ctx, nrows);
    q_t.alloc(ctx, nrows);
    InterpLocateFn loc = {r, row_code.get(), npe, q_fkey.get(), q_j.get(), q_t.get()};
    launch(ctx, nrows, loc, "interp_locate");
  }

  int h_missing = 0;
  if (comm) {
    /* ship the nodes this rank could not place to the owner of the point
       (reference :6699-6783) and take over the ones other ranks could not */
    DBuf<u64> own_store;
    OwnerMap om = make_owner_map(coarse, coarse.fmt.D, own_store);
    DBuf<u64> mk(ctx, nrows), mp(ctx, nrows);
    DBuf<u32> md(ctx, nrows);
    MissCountFn mc = {q_t.get()};
    MissFillFn mf;
    mf.fkey = q_fkey.get();
    mf.j = q_j.get();
    mf.num = q_num.get();
    mf.t = q_t.get();
    mf.ffmt = fine.fmt;
    mf.forder = fn.order;
    for (int i = 0; i < kMaxOrder; i++) mf.fknots[i] = fn.knots[i];
    mf.om = om;
    mf.D = coarse.fmt.D;
    mf.out_key = mk.get();
    mf.out_payload = mp.get();
    mf.out_dest = md.get();
    const i64 nmiss = (i64)scan_apply(ctx, nrows, mc, mf, "interp_miss_list");
    U32DestFn mdest = {md.get()};
    RoutePlan plan;
    make_route(ctx, *comm, nmiss, mdest, plan);
    DBuf<u64> rk, rp;
    route_array(ctx, *comm, plan, mk.get(), rk);
    route_array(ctx, *comm, plan, mp.get(), rp);
    const i64 nrecv = plan.nrecv;
    /* unified request list: local found rows, then received rows */
    const i64 nfound = nrows - nmiss;
    const i64 ntot = nfound + nrecv;
    DBuf<u64> u_fkey(ctx, ntot);
    DBuf<int> u_j(ctx, ntot), u_num(ctx, ntot);
    DBuf<i64> u_t(ctx, ntot);
    FoundCountFn fc = {q_t.get()};
    FoundFillFn ff = {q_fkey.get(), q_j.get(), q_num.get(), q_t.get(),
                      u_fkey.get(), u_j.get(), u_num.get(), u_t.get()};
    scan_apply(ctx, nrows, fc, ff, "interp_found_compact");
    if (nrecv > 0) {
      copy_d2d(ctx, u_fkey.get() + nfound, rk.get(), (size_t)nrecv * sizeof(u64));
      DBuf<int> missing(ctx, 1);
      dev_zero(ctx, missing.get(), sizeof(int));
      InterpLocateRecvFn lr = {r,
                               u_fkey.get() + nfound,
                               rp.get(),
                               u_j.get() + nfound,
                               u_num.get() + nfound,
                               u_t.get() + nfound,
                               missing.get()};
      launch(ctx, nrecv, lr, "interp_locate_recv");
      copy_d2h(ctx, &h_missing, missing.get(), sizeof(int));
      if (h_missing) {
        fprintf(stderr,
                "[%d] TMROctForest Error: Destination processor does not own "
                "node\n", comm->rank);
      }
    }
    q_fkey.swap(u_fkey);
    q_j.swap(u_j);
    q_num.swap(u_num);
    q_t.swap(u_t);
    nrows = ntot;
  } else if (nrows > 0) {
    DBuf<int> missing(ctx, 1);
    dev_zero(ctx, missing.get(), sizeof(int));
    MissCountFn mc = {q_t.get()};
    DBuf<u32> tmp(ctx, nrows);
    h_missing = scan_counts(ctx, nrows, mc, tmp.get(), "interp_miss_count") > 0;
    if (h_missing) {
      fprintf(stderr,
              "TMROctForest Error: createInterpolation found fine nodes with no "
              "enclosing coarse element\n");
    }
  }

  if (nrows == 0) {
    I.valid = true;
    I.rowp.alloc(ctx, 1);
    dev_zero(ctx, I.rowp.get(), sizeof(int));
    return check_errors(ctx, "create_interp");
  }
  RowRequests q = {q_fkey.get(), q_j.get(), q_t.get()};
  DBuf<u32> off(ctx, nrows), cnt(ctx, nrows);
  /* corder is a template parameter only to size the per-thread row buffer */
#define TMR_INTERP_COUNT(C)                          \
  {                                                  \
    InterpCountFn<C> cf = {r, q, cnt.get()};         \
    launch(ctx, nrows, cf, "interp_row_count");      \
  }
  switch (r.corder) {
    case 2: TMR_INTERP_COUNT(2) break;
    case 3: TMR_INTERP_COUNT(3) break;
    case 4: TMR_INTERP_COUNT(4) break;
    case 5: TMR_INTERP_COUNT(5) break;
    case 6: TMR_INTERP_COUNT(6) break;
    case 7: TMR_INTERP_COUNT(7) break;
    case 8: TMR_INTERP_COUNT(8) break;
    case 9: TMR_INTERP_COUNT(9) break;
    case 10: TMR_INTERP_COUNT(10) break;
    case 11: TMR_INTERP_COUNT(11) break;
    case 12: TMR_INTERP_COUNT(12) break;
    case 13: TMR_INTERP_COUNT(13) break;
    case 14: TMR_INTERP_COUNT(14) break;
    case 15: TMR_INTERP_COUNT(15) break;
    default: TMR_INTERP_COUNT(16) break;
  }
#undef TMR_INTERP_COUNT
  StoredCountFn sc = {cnt.get()};
  const u64 nnz = scan_counts(ctx, nrows, sc, off.get(), "interp_row_offsets");
  cnt.reset();
  I.rowp.alloc(ctx, nrows + 1);
  RowPtrFn rp = {off.get(), nrows, (u32)nnz, I.rowp.get()};
  launch(ctx, nrows + 1, rp, "interp_row_ptr");
  I.cols.alloc(ctx, (i64)nnz);
  I.vals.alloc(ctx, (i64)nnz);
  /* offsets are known: replay the build and store */
#define TMR_INTERP_FILL(C)                                      \
  {                                                             \
    InterpFillFn<C> ffn = {r, q, I.cols.get(), I.vals.get()};   \
    InterpFillPlaceFn<C> pl = {ffn, off.get()};                 \
    launch(ctx, nrows, pl, "interp_row_fill");                  \
  }
  switch (r.corder) {
    case 2: TMR_INTERP_FILL(2) break;
    case 3: TMR_INTERP_FILL(3) break;
    case 4: TMR_INTERP_FILL(4) break;
    case 5: TMR_INTERP_FILL(5) break;
    case 6: TMR_INTERP_FILL(6) break;
    case 7: TMR_INTERP_FILL(7) break;
    case 8: TMR_INTERP_FILL(8) break;
    case 9: TMR_INTERP_FILL(9) break;
    case 10: TMR_INTERP_FILL(10) break;
    case 11: TMR_INTERP_FILL(11) break;
    case 12: TMR_INTERP_FILL(12) break;
    case 13: TMR_INTERP_FILL(13) break;
    case 14: TMR_INTERP_FILL(14) break;
    case 15: TMR_INTERP_FILL(15) break;
    default: TMR_INTERP_FILL(16) break;
  }
#undef TMR_INTERP_FILL
  {
    int h_overflow = 0;
    copy_d2h(ctx, &h_overflow, row_overflow.get(), sizeof(int));
    if (h_overflow) {
      fprintf(stderr,
              "TMROctForest Error: a prolongation row exceeded its buffer "
              "(RowCap)\n");
      ctx.last_error = "interpolation row overflow";
    }
  }
  I.rows.swap(q_num);
  I.rows.set_size(nrows);
  I.nrows = nrows;
  I.nnz = (i64)nnz;
  I.valid = true;
  return check_errors(ctx, "create_interp");
}

/* ---- generic TMROctantArray::sort / contains -------------------------------
   Arbitrary int32 coordinates (nodes may sit at 2^30-1, neighbour candidates
   may be negative), so the key is the full 144-bit
   [ block | 96-bit Morton of sign-biased coordinates | level or info ]
   sorted as three LSD words through an index payload. */
struct WideWordFn {
  const Oct24 *recs;
  const u32 *idx;
  int word;
  int node_mode;
  u64 *out;
  TMR_HD void operator()(i64 i) const {
    const Oct24 r = recs[idx[i]];
    const u32 xb = (u32)r.x ^ 0x80000000u, yb = (u32)r.y ^ 0x80000000u,
              zb = (u32)r.z ^ 0x80000000u;
    if (word == 0) {
      const u32 tail =
          (u32)(uint16_t)((node_mode ? r.info : r.level) ^ (int16_t)0x8000);
      out[i] = (morton3(xb & 0xffffu, yb & 0xffffu, zb & 0xffffu) << 16) | tail;
    } else if (word == 1) {
      out[i] = morton3(xb >> 16, yb >> 16, zb >> 16);
    } else {
      out[i] = (u64)((u32)r.block ^ 0x80000000u);
    }
  }
};

struct IotaFn {
  u32 *v;
  TMR_HD void operator()(i64 i) const { v[i] = (u32)i; }
};

TMR_HD bool same_slot(const Oct24 &a, const Oct24 &b, int node_mode) {
  if (a.block != b.block || a.x != b.x || a.y != b.y || a.z != b.z) return false;
  return node_mode ? (a.info == b.info) : true;
}

struct WideTailFn {
  const Oct24 *recs;
  const u32 *idx;
  i64 n;
  int node_mode;
  TMR_HD u32 operator()(i64 i) const {
    if (i == n - 1) return 1;
    return same_slot(recs[idx[i]], recs[idx[i + 1]], node_mode) ? 0u : 1u;
  }
};

struct WideGatherFn {
  WideTailFn tail;
  const u32 *off;
  Oct24 *out;
  TMR_HD void operator()(i64 i) const {
    if (tail(i)) out[off[i]] = tail.recs[tail.idx[i]];
  }
};

inline int array_sort(Ctx &ctx, Oct24 *h_recs, i64 n, int node_mode, i64 *nout) {
  *nout = n;
  if (n <= 1) return 0;
  DBuf<Oct24> recs(ctx, n), out(ctx, n);
  copy_h2d(ctx, recs.get(), h_recs, (size_t)n * sizeof(Oct24));
  DBuf<u64> k(ctx, n), k_alt(ctx, n);
  DBuf<u32> idx(ctx, n), idx_alt(ctx, n);
  IotaFn io = {idx.get()};
  launch(ctx, n, io, "array_iota");
  const int word_bits[3] = {64, 48, 32};
  for (int word = 0; word < 3; word++) {
    WideWordFn wf = {recs.get(), idx.get(), word, node_mode, k.get()};
    launch(ctx, n, wf, "array_wide_word");
    radix_sort(ctx, k, k_alt, idx, idx_alt, n, 0, word_bits[word]);
  }
  DBuf<u32> off(ctx, n);
  WideTailFn tf = {recs.get(), idx.get(), n, node_mode};
  const i64 m = (i64)scan_counts(ctx, n, tf, off.get(), "array_unique_scan");
  WideGatherFn g = {tf, off.get(), out.get()};
  launch(ctx, n, g, "array_gather");
  copy_d2h(ctx, h_recs, out.get(), (size_t)m * sizeof(Oct24));
  *nout = m;
  return check_errors(ctx, "array_sort");
}

struct ContainsFn {
  const Oct24 *arr;
  i64 n;
  const Oct24 *q;
  int mode; /* 0 exact element, 1 position, 2 node */
  int *out;
  TMR_HD int cmp(const Oct24 &a, const Oct24 &b) const {
    const int c = compare_position(a.block, a.x, a.y, a.z, b.block, b.x, b.y, b.z);
    if (c != 0 || mode == 1) return c;
    return mode == 0 ? (a.level - b.level) : (a.info - b.info);
  }
  TMR_HD void operator()(i64 i) const {
    const Oct24 key = q[i];
    i64 lo = 0, hi = n;
    int found = -1;
    while (lo < hi) {
      const i64 mid = lo + (hi - lo) / 2;
      const int c = cmp(key, arr[mid]);
      if (c == 0) {
        found = (int)mid;
        break;
      }
      if (c < 0) {
        hi = mid;
      } else {
        lo = mid + 1;
      }
    }
    out[i] = found;
  }
};

inline int array_contains(Ctx &ctx, const Oct24 *h_sorted, i64 n,
                          const Oct24 *h_q, i64 nq, int mode, int *h_out) {
  if (nq <= 0) return 0;
  DBuf<Oct24> arr(ctx, n), q(ctx, nq);
  DBuf<int> out(ctx, nq);
  copy_h2d(ctx, arr.get(), h_sorted, (size_t)n * sizeof(Oct24));
  copy_h2d(ctx, q.get(), h_q, (size_t)nq * sizeof(Oct24));
  ContainsFn c = {arr.get(), n, q.get(), mode, out.get()};
  launch(ctx, nq, c, "array_contains");
  copy_d2h(ctx, h_out, out.get(), (size_t)nq * sizeof(int));
  return check_errors(ctx, "array_contains");
}

}  // namespace tmrgpu

#endif
