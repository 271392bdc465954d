/*
  tmrgpu.h -- the C-ABI between the host C++ drop-in (TMROctForest /
  TMROctantArray, tmr_b200/csrc/host/) and the sm_100a CUDA layer
  (tmr_b200/csrc/gpu/).  Plain pointers and sizes only; every function returns
  0 on success and non-zero after printing "TMROctForest Error: ..." to stderr
  (the reference's error convention, e.g. src/TMROctForest.cpp:2918-2923).

  Each entry point names the reference code it replaces.  Host buffers are
  caller-owned; device state lives behind the opaque handles.  There is no CPU
  implementation behind this interface.
*/
#ifndef TMRGPU_H
#define TMRGPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tmrgpu_ctx tmrgpu_ctx;       /* device + stream + scratch    */
typedef struct tmrgpu_forest tmrgpu_forest; /* device-resident octant forest */

/* byte-identical to TMROctant (reference src/TMROctant.h:49-53) */
typedef struct {
  int32_t block, x, y, z, tag;
  int16_t level, info;
} tmrgpu_octant;

/* ---- context ------------------------------------------------------------ */
/* stream: a cudaStream_t created by the caller (e.g. torch's current stream)
   or NULL for a private stream.  Fails when no CUDA device is present. */
int tmrgpu_ctx_create(int device, void *stream, tmrgpu_ctx **out);
int tmrgpu_ctx_destroy(tmrgpu_ctx *ctx);
int tmrgpu_ctx_sync(tmrgpu_ctx *ctx);
void *tmrgpu_ctx_stream(tmrgpu_ctx *ctx);
/* per-kernel CUDA-event timing (bench roofline): enable, reset, read as JSON
   {"kernel": {"launches": n, "ms": t}, ...}; returns bytes needed */
int tmrgpu_profile_enable(tmrgpu_ctx *ctx, int on);
int tmrgpu_profile_reset(tmrgpu_ctx *ctx);
int tmrgpu_profile_json(tmrgpu_ctx *ctx, char *buf, int buflen);
long tmrgpu_launch_count(tmrgpu_ctx *ctx);
/* blocking host<->device round trips (synchronising copies, error sweeps)
   since the last tmrgpu_profile_reset: the control-plane cost of an operation */
long tmrgpu_sync_count(tmrgpu_ctx *ctx);
/* bytes copied device->host (h2d = 0) or host->device (1) since the last
   tmrgpu_profile_reset: what the end-to-end figures of bench.py count */
int64_t tmrgpu_bus_bytes(tmrgpu_ctx *ctx, int h2d);

/* ---- multi-GPU: one process per GPU, NCCL over NVLink ----------------------
   (replaces MPI_Comm + the MPI datatypes of reference src/TMRBase.cpp:42-92)
   rank 0 obtains an id (128 bytes), ships it to the other processes by any
   means (torch.distributed, MPI, a file), then every process attaches its
   context.  Collective. */
int tmrgpu_comm_unique_id(void *out, int out_bytes);
int tmrgpu_ctx_init_comm(tmrgpu_ctx *ctx, int rank, int size, const void *id);
int tmrgpu_ctx_rank(tmrgpu_ctx *ctx);
int tmrgpu_ctx_size(tmrgpu_ctx *ctx);

/* ---- forest -------------------------------------------------------------- */
int tmrgpu_forest_create(tmrgpu_ctx *ctx, tmrgpu_forest **out);
/* The forest lives on this GPU alone although the context has a communicator:
   a TMROctForest constructed on a one-rank communicator (MPI_COMM_SELF) inside
   a multi-rank job (reference src/TMROctForest.cpp:331-337 takes rank and size
   from the communicator it is given).  Inherited by duplicate / coarsen. */
int tmrgpu_forest_set_serial(tmrgpu_forest *f, int serial);
int tmrgpu_forest_destroy(tmrgpu_forest *f);

/* Super-mesh tables computed by the host class (replaces the reads of
   TMRBlockConn in reference src/TMROctForest.h:323-410).  Arrays are copied. */
int tmrgpu_set_connectivity(
    tmrgpu_forest *f, int nblocks, int nnodes, int nedges, int nfaces,
    const int *block_conn, const int *block_edge_conn,
    const int *block_face_conn, const int *block_face_ids,
    const int *node_block_ptr, const int *node_block_conn,
    const int *edge_block_ptr, const int *edge_block_conn,
    const int *face_block_ptr, const int *face_block_conn,
    const int *node_block_owners, const int *edge_block_owners,
    const int *face_block_owners);
/* share src's tables with dst (duplicate()/coarsen(): reference copyData
   src/TMROctForest.cpp:488-500) */
int tmrgpu_share_connectivity(tmrgpu_forest *src, tmrgpu_forest *dst);

/* element array <-> host records */
int64_t tmrgpu_count(tmrgpu_forest *f);
int tmrgpu_upload_octants(tmrgpu_forest *f, const tmrgpu_octant *recs,
                          int64_t n);
int tmrgpu_download_octants(tmrgpu_forest *f, tmrgpu_octant *recs);
int tmrgpu_download_info(tmrgpu_forest *f, int16_t *info);

/* sort + uniq the device element array in place (TMROctantArray::sort,
   reference src/TMROctant.cpp:357-399, as used by createRandomTrees :1884) */
int tmrgpu_sort_unique(tmrgpu_forest *f);

/* createTrees for blocks [block_start, block_end)
   (reference src/TMROctForest.cpp:1744-1833) */
int tmrgpu_create_trees(tmrgpu_forest *f, int level, int block_start,
                        int block_end);
/* refine (reference :2169-2329); flags on the host (copied inside) or already
   on the device; NULL = refine everything by one level */
int tmrgpu_refine(tmrgpu_forest *f, const int *h_flags, int min_level,
                  int max_level);
int tmrgpu_refine_device(tmrgpu_forest *f, const int *d_flags, int min_level,
                         int max_level);
/* repartition (reference :1922-2088): equal-count re-split of the global
   Morton order over the first max_rank ranks; collective */
int tmrgpu_repartition(tmrgpu_forest *f, int max_rank);
/* owners[] table (first octant of every rank), size() records */
int tmrgpu_get_owners(tmrgpu_forest *f, tmrgpu_octant *out);
/* owned-node prefix over ranks, size()+1 ints (reference node_range) */
int tmrgpu_node_range(tmrgpu_forest *f, int *out);
/* Borrowed pointer to a page-locked host copy of one node array, owned by the
   forest and valid until its node data is freed -- the storage behind the
   borrowed-pointer getters getNodeConn / getNodeNumbers / getDepNodeConn
   (reference src/TMROctForest.cpp:5686-5740).  which: 0 conn, 1 node numbers
   sorted ascending (:4246), 2 dep_ptr, 3 dep_conn, 4 dep_weights.  Only the
   array asked for is copied. */
int tmrgpu_node_mirror(tmrgpu_forest *f, int which, const void **out);
/* bit mask of arrays whose host copy createNodes starts on a second stream as
   soon as the array is final (1 conn, 2 node numbers, 4 dependent CSR), so
   the read-back overlaps the rest of createNodes; duplicates inherit it.
   Default: TMR_B200_NODE_PREFETCH or 0. */
int tmrgpu_set_node_prefetch(tmrgpu_forest *f, int mask);
/* all-to-all-v of 24-byte octant records held in HOST arrays (the public
   distributeOctants / sendOctants entry points, reference :2379-2509):
   send_ptr/recv_ptr have size()+1 entries (element offsets); the records are
   staged through the GPU and exchanged over NCCL.  exchange_counts turns
   per-destination send counts into per-source receive counts. */
int tmrgpu_exchange_counts(tmrgpu_ctx *ctx, const int *send_counts,
                           int *recv_counts);
int tmrgpu_exchange_records(tmrgpu_ctx *ctx, const tmrgpu_octant *send,
                            const int *send_ptr, tmrgpu_octant *recv,
                            const int *recv_ptr);
/* balance (reference :2917-3089) */
int tmrgpu_balance(tmrgpu_forest *f, int balance_corner);
/* coarsen / duplicate into an existing forest (reference :2097-2164) */
int tmrgpu_coarsen(tmrgpu_forest *src, tmrgpu_forest *dst);
int tmrgpu_duplicate(tmrgpu_forest *src, tmrgpu_forest *dst);

/* createNodes (reference :4064-4268).  knots: `order` interpolation knots. */
int tmrgpu_create_nodes(tmrgpu_forest *f, int order, int interp_type,
                        const double *knots);
int tmrgpu_free_nodes(tmrgpu_forest *f);
/* sizes: [0] elements [1] local nodes [2] dependent nodes [3] owned nodes
   [4] dependent nnz [5] first owned node number */
int tmrgpu_node_sizes(tmrgpu_forest *f, int64_t sizes[6]);
/* number of node-candidate keys the last createNodes sorted (roofline
   accounting of the sort passes) */
int64_t tmrgpu_node_candidates(tmrgpu_forest *f);
/* copy-out; any pointer may be NULL to skip that array */
int tmrgpu_download_nodes(tmrgpu_forest *f, int *conn, int *node_numbers,
                          int *dep_ptr, int *dep_conn, double *dep_weights);

/* all local node numbers sorted ascending (what getNodeNumbers() hands out,
   reference :4246) */
int tmrgpu_download_sorted_node_numbers(tmrgpu_forest *f, int *out);

/* Device-resident views for GPU consumers (the FE assembler and
   a bulk TACSBVecInterp ingest can read the arrays where they are instead of
   going through the host int* getters).  Pointers are device pointers owned by
   the forest, valid until the next mutating call; sizes as tmrgpu_node_sizes /
   tmrgpu_create_interp report them.  Any out-pointer may be NULL. */
int tmrgpu_node_device_views(tmrgpu_forest *f, const int **conn,
                             const int **node_numbers, const int **dep_ptr,
                             const int **dep_conn, const double **dep_weights,
                             const uint64_t **element_keys, int *key_depth,
                             int *block_bits);
/* The arrays TMROctTACSCreator::createTACS hands to TACSAssembler (reference
   src/TMR_TACSCreator.cpp:332-461: setElementConnectivity(ptr, conn),
   setDependentNodes(dep_ptr, dep_conn, dep_weights) and the three counts of
   the TACSAssembler constructor), as DEVICE pointers in exactly that shape,
   for an assembler that lives on the GPU.  elem_ptr[i] = order^3 * i is built
   on first request.  Valid until the next mutating call on the forest. */
typedef struct {
  int64_t num_elements, num_owned_nodes, num_dep_nodes, num_local_nodes, dep_nnz;
  int order;
  const int *elem_ptr;      /* [num_elements + 1] */
  const int *conn;          /* [num_elements * order^3] global node numbers */
  const int *dep_ptr;       /* [num_dep_nodes + 1] */
  const int *dep_conn;      /* [dep_nnz] */
  const double *dep_weights; /* [dep_nnz] */
  const int *node_numbers;  /* [num_local_nodes] number of each local node, node order */
} tmrgpu_assembler_view;
int tmrgpu_assembler_views(tmrgpu_forest *f, tmrgpu_assembler_view *out);
int tmrgpu_interp_device_views(tmrgpu_forest *fine, const int **rows,
                               const int **rowp, const int **cols,
                               const double **vals);

/* evaluateNodeLocations (reference :5524-5675) for forests whose trees are
   trilinear hexahedra: corners = [nblocks][8][3] (tree corner c: bit0 x, bit1
   y, bit2 z), X = [num_local_nodes][3] in the order of the sorted node numbers.
   Every node takes the location evaluated through the first element and local
   slot (in element order) that reference it, exactly the reference's flags[]
   rule, at the parametric point u + 0.5 d (1 + knot[i]) of that element. */
int tmrgpu_eval_trilinear_points(tmrgpu_forest *f, const double *corners,
                                 double *X);

/* createInterpolation (reference :6611-6793) between two forests that both
   have nodes.  Builds the CSR on the device; rows are emitted in the
   reference's call order (first touch in element order). */
int tmrgpu_create_interp(tmrgpu_forest *fine, tmrgpu_forest *coarse,
                         int64_t *nrows, int64_t *nnz);
int tmrgpu_download_interp(tmrgpu_forest *fine, int *rows, int *rowp,
                           int *cols, double *vals);
/* batched findEnclosing against this forest's elements (reference
   :6228-6377): nodes are element records whose info = local node index */
int tmrgpu_find_enclosing(tmrgpu_forest *f, int order, const double *knots,
                          const tmrgpu_octant *nodes, int64_t n,
                          int *out_index);

/* ---- TMROctantArray (reference src/TMROctant.cpp:357-424) ----------------- */
/* sort + uniq of an arbitrary host array, in place; *nout = new size.
   use_node_index: 0 element mode (keep finest per anchor), 1 node mode */
int tmrgpu_array_sort(tmrgpu_ctx *ctx, tmrgpu_octant *recs, int64_t n,
                      int use_node_index, int64_t *nout);
/* batched contains() against a SORTED host array: mode 0 exact element,
   1 position only, 2 node (position+info); out_index = match or -1 */
int tmrgpu_array_contains(tmrgpu_ctx *ctx, const tmrgpu_octant *sorted,
                          int64_t n, const tmrgpu_octant *queries, int64_t nq,
                          int mode, int *out_index);

/* ---- measurement helpers (bench.py / tests) ------------------------------- */
/* hash-driven refinement flags on the device (SURVEY.md 8(d) recipe):
   flag = record_hash(seed, octant) % 100 < pct; d_flags has count() ints */
int tmrgpu_synth_flags(tmrgpu_forest *f, uint64_t seed, int pct, int *d_flags);
/* order-independent checksum: sum of record_hash(0, octant) mod 2^64 */
int tmrgpu_checksum(tmrgpu_forest *f, uint64_t *out);
/* raw device allocations for bench-owned buffers */
int tmrgpu_dev_alloc(tmrgpu_ctx *ctx, int64_t bytes, void **out);
int tmrgpu_dev_free(tmrgpu_ctx *ctx, void *p);
/* test hook: the nth next device allocation of this context fails (0 = off).
   The operation it hits must print "TMROctForest Error", launch nothing on
   the missing buffer, return nonzero, and leave the context usable. */
int tmrgpu_test_fail_alloc(tmrgpu_ctx *ctx, long nth);
/* page-locked host buffers (cached per context) for the host mirrors the
   drop-in hands out, so that D2H runs at PCIe speed */
int tmrgpu_host_alloc(tmrgpu_ctx *ctx, int64_t bytes, void **out);
int tmrgpu_host_free(tmrgpu_ctx *ctx, void *p);
int tmrgpu_copy_d2h(tmrgpu_ctx *ctx, void *dst, const void *src, int64_t bytes);
int tmrgpu_copy_h2d(tmrgpu_ctx *ctx, void *dst, const void *src, int64_t bytes);
/* [0] octants before refine, [1] after refine, [2] after balance */
int tmrgpu_last_counts(tmrgpu_forest *f, int64_t counts[3]);

/* ---- primitive self-test hooks (tests/ only) -------------------------------
   run the radix sort / chained scan on caller data so the primitives can be
   checked against numpy in isolation */
int tmrgpu_test_radix_sort(tmrgpu_ctx *ctx, uint64_t *keys, uint32_t *vals,
                           int64_t n, int bit_lo, int bit_hi);
int tmrgpu_test_scan(tmrgpu_ctx *ctx, const uint32_t *counts, int64_t n,
                     uint32_t *out_exclusive, uint64_t *total);
const char *tmrgpu_build_kind(void);

#ifdef __cplusplus
}
#endif
#endif
