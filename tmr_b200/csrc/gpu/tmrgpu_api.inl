/*
  tmrgpu_api.inl -- implementation of include/tmrgpu.h on top of ops_*.h.
  Included by tmrgpu_cuda.cu (the product, compiled by nvcc for sm_100a) and,
  for pre-GPU logic checks only, by tests/emu/tmrgpu_emu.cpp.
*/
#include <string.h>

#include <sstream>

#include "ops_interp.h"
#include "ops_multi.h"
#include "ops_nodes.h"
#include "ops_points.h"
#include "tmrgpu.h"

using namespace tmrgpu;

namespace tmrgpu {
int comm_unique_id(void *out, int out_bytes);
Comm *comm_create(Ctx &ctx, int rank, int size, const void *id_bytes);
void comm_destroy(Comm *c);
}  // namespace tmrgpu


struct tmrgpu_ctx {
  Ctx c;
  bool own_stream;
};

struct tmrgpu_forest {
  Forest f;
  explicit tmrgpu_forest(Ctx *c) : f(c) {}
};

/* Every mutating entry point leaves through here: a failure recorded by the
   operation (allocation, size limit, NCCL) that an early return skipped past
   is reported and CLEARED now, so it can neither be lost nor leak into the
   next operation on the context. */
static int swept(Ctx &ctx, int rc, const char *where) {
  if (rc == 0 && ctx.last_error.empty()) return 0;
  const int e = check_errors(ctx, where);
  return rc ? rc : e;
}

extern "C" {

int tmrgpu_ctx_sync(tmrgpu_ctx *ctx) { return check_errors(ctx->c, "sync"); }
void *tmrgpu_ctx_stream(tmrgpu_ctx *ctx) { return ctx->c.stream; }

int tmrgpu_profile_enable(tmrgpu_ctx *ctx, int on) {
  prof_resolve(ctx->c);
  ctx->c.profile = on;
  return 0;
}

int tmrgpu_profile_reset(tmrgpu_ctx *ctx) {
  prof_resolve(ctx->c);
  ctx->c.stats.clear();
  ctx->c.launch_count = 0;
  ctx->c.sync_count = 0;
  ctx->c.bytes_d2h = ctx->c.bytes_h2d = 0;
  return 0;
}

int tmrgpu_profile_json(tmrgpu_ctx *ctx, char *buf, int buflen) {
  prof_resolve(ctx->c);
  std::ostringstream os;
  os << "{";
  bool first = true;
  for (std::map<std::string, KernelStat>::const_iterator it =
           ctx->c.stats.begin();
       it != ctx->c.stats.end(); ++it) {
    if (!first) os << ", ";
    first = false;
    os << "\"" << it->first << "\": {\"launches\": " << it->second.launches
       << ", \"ms\": " << it->second.ms << "}";
  }
  os << "}";
  const std::string s = os.str();
  if (buf && buflen > 0) {
    const size_t nc = s.size() < (size_t)buflen - 1 ? s.size() : (size_t)buflen - 1;
    memcpy(buf, s.data(), nc);
    buf[nc] = 0;
  }
  return (int)s.size() + 1;
}

long tmrgpu_launch_count(tmrgpu_ctx *ctx) { return ctx->c.launch_count; }
long tmrgpu_sync_count(tmrgpu_ctx *ctx) { return ctx->c.sync_count; }
int64_t tmrgpu_bus_bytes(tmrgpu_ctx *ctx, int h2d) {
  return h2d ? ctx->c.bytes_h2d : ctx->c.bytes_d2h;
}

int tmrgpu_comm_unique_id(void *out, int out_bytes) {
  return comm_unique_id(out, out_bytes);
}

int tmrgpu_ctx_init_comm(tmrgpu_ctx *ctx, int rank, int size, const void *id) {
  if (ctx->c.comm) {
    comm_destroy(ctx->c.comm);
    ctx->c.comm = NULL;
  }
  if (size <= 1) return 0;
  ctx->c.comm = comm_create(ctx->c, rank, size, id);
  return ctx->c.comm ? 0 : 1;
}

int tmrgpu_ctx_rank(tmrgpu_ctx *ctx) { return ctx->c.comm ? ctx->c.comm->rank : 0; }
int tmrgpu_ctx_size(tmrgpu_ctx *ctx) { return ctx->c.comm ? ctx->c.comm->size : 1; }

int tmrgpu_forest_create(tmrgpu_ctx *ctx, tmrgpu_forest **out) {
  *out = new tmrgpu_forest(&ctx->c);
  return 0;
}

int tmrgpu_forest_set_serial(tmrgpu_forest *f, int serial) {
  f->f.serial = serial != 0;
  return 0;
}

int tmrgpu_forest_destroy(tmrgpu_forest *f) {
  delete f;
  return 0;
}

int tmrgpu_set_connectivity(
    tmrgpu_forest *F, int nblocks, int nnodes, int nedges, int nfaces,
    const int *block_conn, const int *block_edge_conn,
    const int *block_face_conn, const int *block_face_ids,
    const int *node_block_ptr, const int *node_block_conn,
    const int *edge_block_ptr, const int *edge_block_conn,
    const int *face_block_ptr, const int *face_block_conn,
    const int *node_block_owners, const int *edge_block_owners,
    const int *face_block_owners) {
  Forest &f = F->f;
  Ctx &ctx = *f.ctx;
  const int nbc = node_block_ptr[nnodes], ebc = edge_block_ptr[nedges],
            fbc = face_block_ptr[nfaces];
  /* one packed allocation, one H2D copy */
  std::vector<int> pack;
  size_t off[13];
  const int *src[13] = {block_conn,      block_edge_conn,  block_face_conn,
                        block_face_ids,  node_block_ptr,   node_block_conn,
                        edge_block_ptr,  edge_block_conn,  face_block_ptr,
                        face_block_conn, node_block_owners, edge_block_owners,
                        face_block_owners};
  const size_t len[13] = {(size_t)8 * nblocks, (size_t)12 * nblocks,
                          (size_t)6 * nblocks, (size_t)6 * nblocks,
                          (size_t)nnodes + 1,  (size_t)nbc,
                          (size_t)nedges + 1,  (size_t)ebc,
                          (size_t)nfaces + 1,  (size_t)fbc,
                          (size_t)nnodes,      (size_t)nedges,
                          (size_t)nfaces};
  for (int k = 0; k < 13; k++) {
    off[k] = pack.size();
    pack.insert(pack.end(), src[k], src[k] + len[k]);
  }
  f.table_store.reset(new DBuf<int>(ctx, (i64)pack.size()));
  int *d = f.table_store->get();
  copy_h2d(ctx, d, pack.data(), pack.size() * sizeof(int));
  ConnTables &t = f.tables;
  t.nblocks = nblocks;
  t.nnodes = nnodes;
  t.nedges = nedges;
  t.nfaces = nfaces;
  t.block_conn = d + off[0];
  t.block_edge_conn = d + off[1];
  t.block_face_conn = d + off[2];
  t.block_face_ids = d + off[3];
  t.node_block_ptr = d + off[4];
  t.node_block_conn = d + off[5];
  t.edge_block_ptr = d + off[6];
  t.edge_block_conn = d + off[7];
  t.face_block_ptr = d + off[8];
  t.face_block_conn = d + off[9];
  t.node_block_owners = d + off[10];
  t.edge_block_owners = d + off[11];
  t.face_block_owners = d + off[12];
  f.nblocks = nblocks;
  f.bbits = bits_for(nblocks);
  f.fmt.bbits = f.bbits;
  f.n = 0;
  f.keys.reset();
  f.info.reset();
  f.nodes.clear();
  return check_errors(ctx, "set_connectivity");
}

int tmrgpu_share_connectivity(tmrgpu_forest *src, tmrgpu_forest *dst) {
  dst->f.table_store = src->f.table_store;
  dst->f.tables = src->f.tables;
  dst->f.nblocks = src->f.nblocks;
  dst->f.bbits = src->f.bbits;
  dst->f.fmt.bbits = src->f.bbits;
  return 0;
}

int64_t tmrgpu_count(tmrgpu_forest *f) { return f->f.n; }

int tmrgpu_upload_octants(tmrgpu_forest *f, const tmrgpu_octant *recs,
                          int64_t n) {
  return swept(*f->f.ctx, upload_octants(f->f, reinterpret_cast<const Oct24 *>(recs), n),
               "upload_octants");
}

int tmrgpu_download_octants(tmrgpu_forest *f, tmrgpu_octant *recs) {
  return swept(*f->f.ctx, download_octants(f->f, reinterpret_cast<Oct24 *>(recs)),
               "download_octants");
}

int tmrgpu_download_info(tmrgpu_forest *F, int16_t *info) {
  Forest &f = F->f;
  if (f.n == 0) return 0;
  if (!f.info.get()) {
    memset(info, 0, (size_t)f.n * sizeof(int16_t));
    return 0;
  }
  copy_d2h(*f.ctx, info, f.info.get(), (size_t)f.n * sizeof(int16_t));
  return check_errors(*f.ctx, "download_info");
}

int tmrgpu_sort_unique(tmrgpu_forest *f) {
  sort_unique_elements(f->f);
  if (unify_depth(f->f)) return swept(*f->f.ctx, 1, "sort_unique");
  /* createRandomTrees (:1905-1916) */
  if (gather_owners(f->f, 1)) return swept(*f->f.ctx, 1, "sort_unique");
  return check_errors(*f->f.ctx, "sort_unique");
}

int tmrgpu_create_trees(tmrgpu_forest *f, int level, int block_start,
                        int block_end) {
  int rc = create_trees(f->f, level, block_start, block_end);
  if (!rc) rc = unify_depth(f->f);
  if (!rc) rc = gather_owners(f->f, 1); /* back-fill quirk of createTrees (:1826-1832) */
  return swept(*f->f.ctx, rc, "create_trees");
}

int tmrgpu_repartition(tmrgpu_forest *f, int max_rank) {
  return swept(*f->f.ctx, repartition(f->f, max_rank), "repartition");
}

int tmrgpu_get_owners(tmrgpu_forest *F, tmrgpu_octant *out) {
  Forest &f = F->f;
  for (size_t r = 0; r < f.owners.size(); r++) {
    memcpy(&out[r], &f.owners[r], sizeof(Oct24));
  }
  return 0;
}

int tmrgpu_node_range(tmrgpu_forest *F, int *out) {
  /* computed inside createNodes (reference :4165-4172): no communication here,
     so the getters built on it stay local calls like the reference's */
  const NodeData &nd = F->f.nodes;
  const int R = part_size(F->f);
  for (int r = 0; r <= R; r++) {
    out[r] = r < (int)nd.node_range.size() ? nd.node_range[r] : 0;
  }
  return nd.valid ? 0 : 1;
}

int tmrgpu_node_mirror(tmrgpu_forest *F, int which, const void **out) {
  *out = node_mirror_get(F->f, which);
  return swept(*F->f.ctx, *out ? 0 : 1, "node_mirror");
}

int tmrgpu_set_node_prefetch(tmrgpu_forest *F, int mask) {
  F->f.nodes.prefetch = mask;
  return 0;
}

int tmrgpu_refine_device(tmrgpu_forest *f, const int *d_flags, int min_level,
                         int max_level) {
  int rc = refine(f->f, d_flags, min_level, max_level);
  if (!rc) rc = refine_exchange(f->f); /* no-op on a single rank */
  return swept(*f->f.ctx, rc, "refine");
}

int tmrgpu_refine(tmrgpu_forest *F, const int *h_flags, int min_level,
                  int max_level) {
  Forest &f = F->f;
  if (!h_flags || f.n == 0) {
    int rc = refine(f, NULL, min_level, max_level);
    if (!rc) rc = refine_exchange(f);
    return swept(*f.ctx, rc, "refine");
  }
  DBuf<int> d_flags(*f.ctx, f.n);
  copy_h2d(*f.ctx, d_flags.get(), h_flags, (size_t)f.n * sizeof(int));
  int rc = refine(f, d_flags.get(), min_level, max_level);
  if (!rc) rc = refine_exchange(f);
  return swept(*f.ctx, rc, "refine");
}

int tmrgpu_exchange_counts(tmrgpu_ctx *ctx, const int *send_counts,
                           int *recv_counts) {
  Comm *comm = ctx->c.comm;
  if (!comm) {
    recv_counts[0] = send_counts[0];
    return 0;
  }
  std::vector<i64> sc(comm->size), rc(comm->size);
  for (int r = 0; r < comm->size; r++) sc[r] = send_counts[r];
  exchange_counts(ctx->c, *comm, sc.data(), rc.data());
  for (int r = 0; r < comm->size; r++) recv_counts[r] = (int)rc[r];
  return check_errors(ctx->c, "exchange_counts");
}

int tmrgpu_exchange_records(tmrgpu_ctx *ctx, const tmrgpu_octant *send,
                            const int *send_ptr, tmrgpu_octant *recv,
                            const int *recv_ptr) {
  Ctx &c = ctx->c;
  Comm *comm = c.comm;
  const int R = comm ? comm->size : 1;
  const i64 ns = send_ptr[R], nr = recv_ptr[R];
  if (!comm) {
    if (nr > 0) memcpy(recv, send + send_ptr[0], (size_t)nr * sizeof(Oct24));
    return 0;
  }
  std::vector<i64> so(R + 1), ro(R + 1);
  for (int r = 0; r <= R; r++) {
    so[r] = send_ptr[r];
    ro[r] = recv_ptr[r];
  }
  DBuf<Oct24> ds(c, ns), dr(c, nr);
  copy_h2d(c, ds.get(), send, (size_t)ns * sizeof(Oct24));
  comm->alltoallv(c, ds.get(), so.data(), dr.get(), ro.data(), sizeof(Oct24));
  copy_d2h(c, recv, dr.get(), (size_t)nr * sizeof(Oct24));
  return check_errors(c, "exchange_records");
}

int tmrgpu_balance(tmrgpu_forest *f, int balance_corner) {
  const int rc = forest_comm(f->f) ? balance_multi(f->f, balance_corner)
                                  : balance(f->f, balance_corner);
  return swept(*f->f.ctx, rc, "balance");
}

int tmrgpu_coarsen(tmrgpu_forest *src, tmrgpu_forest *dst) {
  tmrgpu_share_connectivity(src, dst);
  dst->f.serial = src->f.serial;
  int rc = coarsen_into(src->f, dst->f);
  if (!rc) rc = gather_owners(dst->f, 0); /* reference :2157-2160 */
  return swept(*dst->f.ctx, rc, "coarsen");
}

int tmrgpu_duplicate(tmrgpu_forest *src, tmrgpu_forest *dst) {
  tmrgpu_share_connectivity(src, dst);
  dst->f.nodes.prefetch = src->f.nodes.prefetch;
  dst->f.serial = src->f.serial;
  dst->f.owners = src->f.owners; /* reference :2104-2106 */
  return swept(*dst->f.ctx, duplicate_into(src->f, dst->f), "duplicate");
}

int tmrgpu_create_nodes(tmrgpu_forest *f, int order, int interp_type,
                        const double *knots) {
  return swept(*f->f.ctx, create_nodes(f->f, order, interp_type, knots), "create_nodes");
}

int tmrgpu_free_nodes(tmrgpu_forest *f) {
  f->f.nodes.clear();
  f->f.interp.clear();
  return 0;
}

int tmrgpu_node_sizes(tmrgpu_forest *F, int64_t sizes[6]) {
  const NodeData &nd = F->f.nodes;
  sizes[0] = nd.num_elements;
  sizes[1] = nd.num_local_nodes;
  sizes[2] = nd.num_dep_nodes;
  sizes[3] = nd.num_owned_nodes;
  sizes[4] = nd.dep_nnz;
  sizes[5] = nd.node_range_start;
  return nd.valid ? 0 : 1;
}

int64_t tmrgpu_node_candidates(tmrgpu_forest *f) { return f->f.nodes.num_candidates; }

int tmrgpu_download_nodes(tmrgpu_forest *F, int *conn, int *node_numbers,
                          int *dep_ptr, int *dep_conn, double *dep_weights) {
  Forest &f = F->f;
  NodeData &nd = f.nodes;
  Ctx &ctx = *f.ctx;
  if (!nd.valid) return 1;
  const i64 npe = (i64)nd.order * nd.order * nd.order;
  if (conn) {
    copy_d2h(ctx, conn, nd.conn.get(),
             (size_t)(nd.num_elements * npe) * sizeof(int));
  }
  if (node_numbers) {
    if (ensure_node_arrays(f)) return 1;
    copy_d2h(ctx, node_numbers, nd.node_num.get(),
             (size_t)nd.num_local_nodes * sizeof(int));
  }
  if (dep_ptr) {
    copy_d2h(ctx, dep_ptr, nd.dep_ptr.get(),
             (size_t)(nd.num_dep_nodes + 1) * sizeof(int));
  }
  if (dep_conn) {
    copy_d2h(ctx, dep_conn, nd.dep_conn.get(), (size_t)nd.dep_nnz * sizeof(int));
  }
  if (dep_weights) {
    copy_d2h(ctx, dep_weights, nd.dep_weights.get(),
             (size_t)nd.dep_nnz * sizeof(double));
  }
  return check_errors(ctx, "download_nodes");
}

int tmrgpu_node_device_views(tmrgpu_forest *F, const int **conn,
                             const int **node_numbers, const int **dep_ptr,
                             const int **dep_conn, const double **dep_weights,
                             const uint64_t **element_keys, int *key_depth,
                             int *block_bits) {
  Forest &f = F->f;
  NodeData &nd = f.nodes;
  if (conn) *conn = nd.conn.get();
  if (node_numbers && nd.valid && ensure_node_arrays(f)) return 1;
  if (node_numbers) *node_numbers = nd.node_num.get();
  if (dep_ptr) *dep_ptr = nd.dep_ptr.get();
  if (dep_conn) *dep_conn = nd.dep_conn.get();
  if (dep_weights) *dep_weights = nd.dep_weights.get();
  if (element_keys) *element_keys = f.keys.get();
  if (key_depth) *key_depth = f.fmt.D;
  if (block_bits) *block_bits = f.fmt.bbits;
  return nd.valid ? 0 : 1;
}

struct ElemPtrFn {
  int npe;
  int *out;
  TMR_HD void operator()(i64 i) const { out[i] = (int)(npe * i); }
};

int tmrgpu_assembler_views(tmrgpu_forest *F, tmrgpu_assembler_view *out) {
  Forest &f = F->f;
  NodeData &nd = f.nodes;
  if (!out || !nd.valid) return 1;
  const int npe = nd.order * nd.order * nd.order;
  if (ensure_node_arrays(f)) return 1;
  if (nd.elem_ptr.size() != nd.num_elements + 1) {
    nd.elem_ptr.alloc(*f.ctx, nd.num_elements + 1);
    ElemPtrFn ep = {npe, nd.elem_ptr.get()};
    launch(*f.ctx, nd.num_elements + 1, ep, "nodes_elem_ptr");
  }
  out->num_elements = nd.num_elements;
  out->num_owned_nodes = nd.num_owned_nodes;
  out->num_dep_nodes = nd.num_dep_nodes;
  out->num_local_nodes = nd.num_local_nodes;
  out->dep_nnz = nd.dep_nnz;
  out->order = nd.order;
  out->elem_ptr = nd.elem_ptr.get();
  out->conn = nd.conn.get();
  out->dep_ptr = nd.dep_ptr.get();
  out->dep_conn = nd.dep_conn.get();
  out->dep_weights = nd.dep_weights.get();
  out->node_numbers = nd.node_num.get();
  return swept(*f.ctx, 0, "assembler_views");
}

int tmrgpu_interp_device_views(tmrgpu_forest *F, const int **rows,
                               const int **rowp, const int **cols,
                               const double **vals) {
  const InterpData &I = F->f.interp;
  if (rows) *rows = I.rows.get();
  if (rowp) *rowp = I.rowp.get();
  if (cols) *cols = I.cols.get();
  if (vals) *vals = I.vals.get();
  return I.valid ? 0 : 1;
}

int tmrgpu_download_sorted_node_numbers(tmrgpu_forest *f, int *out) {
  return swept(*f->f.ctx, sorted_node_numbers(f->f, out), "sorted_node_numbers");
}

int tmrgpu_eval_trilinear_points(tmrgpu_forest *f, const double *corners, double *X) {
  return swept(*f->f.ctx, eval_trilinear_points(f->f, corners, X), "eval_trilinear_points");
}

int tmrgpu_create_interp(tmrgpu_forest *fine, tmrgpu_forest *coarse,
                         int64_t *nrows, int64_t *nnz) {
  const int rc = create_interp(fine->f, coarse->f);
  if (nrows) *nrows = fine->f.interp.nrows;
  if (nnz) *nnz = fine->f.interp.nnz;
  return swept(*fine->f.ctx, rc, "create_interp");
}

int tmrgpu_download_interp(tmrgpu_forest *F, int *rows, int *rowp, int *cols,
                           double *vals) {
  Forest &f = F->f;
  InterpData &I = f.interp;
  Ctx &ctx = *f.ctx;
  if (!I.valid) return 1;
  if (rows) copy_d2h(ctx, rows, I.rows.get(), (size_t)I.nrows * sizeof(int));
  if (rowp) copy_d2h(ctx, rowp, I.rowp.get(), (size_t)(I.nrows + 1) * sizeof(int));
  if (cols) copy_d2h(ctx, cols, I.cols.get(), (size_t)I.nnz * sizeof(int));
  if (vals) copy_d2h(ctx, vals, I.vals.get(), (size_t)I.nnz * sizeof(double));
  return check_errors(ctx, "download_interp");
}

int tmrgpu_find_enclosing(tmrgpu_forest *F, int order, const double *knots,
                          const tmrgpu_octant *nodes, int64_t n,
                          int *out_index) {
  return swept(*F->f.ctx,
               find_enclosing_batch(F->f, order, knots,
                                    reinterpret_cast<const Oct24 *>(nodes), n, out_index),
               "find_enclosing");
}

int tmrgpu_array_sort(tmrgpu_ctx *ctx, tmrgpu_octant *recs, int64_t n,
                      int use_node_index, int64_t *nout) {
  return swept(ctx->c,
               array_sort(ctx->c, reinterpret_cast<Oct24 *>(recs), n, use_node_index, nout),
               "array_sort");
}

int tmrgpu_array_contains(tmrgpu_ctx *ctx, const tmrgpu_octant *sorted,
                          int64_t n, const tmrgpu_octant *queries, int64_t nq,
                          int mode, int *out_index) {
  return swept(ctx->c,
               array_contains(ctx->c, reinterpret_cast<const Oct24 *>(sorted), n,
                              reinterpret_cast<const Oct24 *>(queries), nq, mode, out_index),
               "array_contains");
}

int tmrgpu_synth_flags(tmrgpu_forest *F, uint64_t seed, int pct, int *d_flags) {
  Forest &f = F->f;
  SynthFlagsFn s = {f.keys.get(), f.fmt, seed, pct, d_flags};
  launch(*f.ctx, f.n, s, "synth_flags");
  return 0;
}

int tmrgpu_checksum(tmrgpu_forest *F, uint64_t *out) {
  *out = checksum(F->f);
  return check_errors(*F->f.ctx, "checksum");
}

int tmrgpu_dev_alloc(tmrgpu_ctx *ctx, int64_t bytes, void **out) {
  *out = dev_alloc(ctx->c, (size_t)bytes);
  return *out ? 0 : 1;
}

int tmrgpu_dev_free(tmrgpu_ctx *ctx, void *p) {
  dev_free(ctx->c, p);
  return 0;
}

namespace {
struct CopyCountFn {
  const u32 *in;
  TMR_HD u32 operator()(i64 i) const { return in[i]; }
};
}  // namespace

int tmrgpu_test_radix_sort(tmrgpu_ctx *ctx, uint64_t *keys, uint32_t *vals,
                           int64_t n, int bit_lo, int bit_hi) {
  Ctx &c = ctx->c;
  DBuf<u64> k(c, n), ka(c, n);
  DBuf<u32> v, va;
  copy_h2d(c, k.get(), keys, (size_t)n * sizeof(u64));
  if (vals) {
    v.alloc(c, n);
    va.alloc(c, n);
    copy_h2d(c, v.get(), vals, (size_t)n * sizeof(u32));
  }
  radix_sort(c, k, ka, v, va, n, bit_lo, bit_hi);
  copy_d2h(c, keys, k.get(), (size_t)n * sizeof(u64));
  if (vals) copy_d2h(c, vals, v.get(), (size_t)n * sizeof(u32));
  return check_errors(c, "test_radix_sort");
}

int tmrgpu_test_scan(tmrgpu_ctx *ctx, const uint32_t *counts, int64_t n,
                     uint32_t *out_exclusive, uint64_t *total) {
  Ctx &c = ctx->c;
  DBuf<u32> in(c, n), out(c, n);
  copy_h2d(c, in.get(), counts, (size_t)n * sizeof(u32));
  CopyCountFn f = {in.get()};
  *total = scan_counts(c, n, f, out.get(), "test_scan");
  copy_d2h(c, out_exclusive, out.get(), (size_t)n * sizeof(u32));
  return check_errors(c, "test_scan");
}

int tmrgpu_test_fail_alloc(tmrgpu_ctx *ctx, long nth) {
  ctx->c.fail_alloc_in = nth;
  return 0;
}

int tmrgpu_host_alloc(tmrgpu_ctx *ctx, int64_t bytes, void **out) {
  *out = host_alloc(ctx->c, (size_t)bytes);
  return *out ? 0 : 1;
}

int tmrgpu_host_free(tmrgpu_ctx *ctx, void *p) {
  host_free(ctx->c, p);
  return 0;
}

int tmrgpu_copy_d2h(tmrgpu_ctx *ctx, void *dst, const void *src, int64_t bytes) {
  copy_d2h(ctx->c, dst, src, (size_t)bytes);
  return 0;
}

int tmrgpu_copy_h2d(tmrgpu_ctx *ctx, void *dst, const void *src, int64_t bytes) {
  copy_h2d(ctx->c, dst, src, (size_t)bytes);
  return 0;
}

int tmrgpu_last_counts(tmrgpu_forest *F, int64_t counts[3]) {
  counts[0] = F->f.last_in;
  counts[1] = F->f.last_mid;
  counts[2] = F->f.last_out;
  return 0;
}

}  // extern "C"
