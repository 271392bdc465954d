/*
  forest_state.h -- the device-resident forest.

  Elements live in HBM as ONE sorted array of 64-bit keys (common.h: KeyFmt)
  plus an optional int16 `info` array; the reference's 24-byte TMROctant
  records (reference src/TMROctant.h:36-54) are only materialised when the
  host asks for them.  Node data (conn, numbers, dependent CSR) stays on the
  device until downloaded.
*/
#ifndef TMRGPU_FOREST_STATE_H
#define TMRGPU_FOREST_STATE_H

#include <stdlib.h>

#include <memory>
#include <thread>

#include "comm.h"
#include "prim.h"
#if defined(TMRGPU_EMU)
#include "prim_emu.h"
#else
#include "prim_cuda.cuh"
#endif

namespace tmrgpu {

/* mesh orders 2..16, as the reference (MAX_ORDER, src/TMROctForest.h:49).  Orders 2 and 3 have one node per
   corner / edge / face / block entity; from order 4 on an entity carries
   (order-2)^dim nodes that are numbered consecutively and reached through the
   edge-reversal and face-orientation permutations of reference
   createLocalConn (src/TMROctForest.cpp:4660-4867).  The bound is the size of
   the per-thread knot, basis and prolongation-row arrays, nothing structural. */
static const int kMaxOrder = 16;

/* page-locked host copy of one node array, owned by the forest: the
   reference's getters hand out borrowed pointers into forest-owned arrays
   (src/TMROctForest.cpp:5686-5740), here they point into these */
struct HostMirror {
  void *p;
  void *ev; /* pending copy (copy_d2h_async handle) */
  void *thread; /* std::thread* filling the array on the host, or NULL */
  void *aux;    /* page-locked staging the fill reads from, or NULL */
  HostMirror() : p(NULL), ev(NULL), thread(NULL), aux(NULL) {}
  void join() {
    if (thread) {
      std::thread *t = static_cast<std::thread *>(thread);
      t->join();
      delete t;
      thread = NULL;
    }
  }
};
enum { kMirrorConn = 0, kMirrorNumbers, kMirrorDepPtr, kMirrorDepConn,
       kMirrorDepWeights, kNumMirrors };

/* dep_ptr and dep_weights reach the host in compact form: one 16-bit stencil
   code per dependent node (which parent edge / face position it sits on) and
   the small table of 1-D weight rows the codes index; host threads rebuild
   the two arrays while the large copies (conn, dep_conn) are still on the
   bus.  A dependent stencil of an order-p mesh is one of 2 p rows (edge) or
   the tensor product of two of them (face) -- 12 bytes per entry shrink to 2
   bytes per NODE.  Code: 0xffff none; edge 0x8000 | bit << 4 | k; face
   b1 << 9 | i << 5 | b2 << 4 | j. */
static const unsigned short kDepCodeNone = 0xffff;
static const unsigned short kDepCodeEdge = 0x8000;
struct DepExpandJob {
  void *thread;  /* std::thread* running the expansion, joined by the getters */
  void *codes;   /* page-locked staging of the codes */
  double *table; /* page-locked staging of the weight rows */
  void *ev_codes, *ev_table;
  DepExpandJob() : thread(NULL), codes(NULL), table(NULL), ev_codes(NULL), ev_table(NULL) {}
};

/* per-leaf prefix counts of the slot construction (ops_nodes_slots.h): nodes and
   dependent nodes before the leaf, and which of its 27 slots hold either */
struct SlotInfo2 {
  u32 node_off, mask, dep_off, dmask;
};

struct NodeData {
  bool valid;
  int order;
  int interp_type;
  double knots[kMaxOrder];
  i64 num_elements;
  i64 num_local_nodes; /* unique node entries referenced on this rank */
  i64 num_dep_nodes;
  i64 num_owned_nodes;
  i64 dep_nnz;
  i64 num_candidates; /* node keys that went through the sort */
  int node_range_start; /* first global number owned by this rank */
  std::vector<int> node_range; /* owned-node prefix over ranks (size + 1) */
  NodeFmt nfmt;
  DBuf<u64> node_keys;  /* sorted unique node keys [num_local_nodes] */
  DBuf<int> node_num;   /* number of each node entry (node order) */
  DBuf<int> conn;       /* [num_elements * order^3], global numbers */
  DBuf<int> dep_ptr;    /* [num_dep_nodes + 1] */
  DBuf<int> dep_conn;   /* [dep_nnz] */
  /* one rank, slot construction: per-leaf prefix counts (16 B, SlotInfo2).  node_keys /
     node_num are then built on first request (ensure_node_arrays): createNodes itself
     and the host getters never need them */
  DBuf<SlotInfo2> slot_info;
  /* several ranks, slot construction: what node_keys / node_num are rebuilt from
     (SlotMulti, ops_nodes.h) */
  std::shared_ptr<void> slot_multi;
  DBuf<int> elem_ptr;   /* [num_elements + 1] order^3 * i, built on request (tmrgpu_assembler_views) */
  DBuf<double> dep_weights;
  DBuf<unsigned short> dep_code; /* [num_dep_nodes] stencil codes (DepExpandJob) */
  DBuf<double> dep_wtab; /* [2 kinds][2 sides][order positions][order] weight rows */
  DepExpandJob dep_job;
  /* host mirrors; prefetch = bit mask of the arrays whose copy createNodes
     starts on the copy stream as soon as the array is final (bit 0 conn, 1
     sorted node numbers, 2 the dependent CSR) */
  Ctx *mctx;
  HostMirror mirror[kNumMirrors];
  DBuf<int> sorted_numbers; /* device staging of the sorted node numbers */
  /* several ranks, slot construction: numbers of the nodes owned elsewhere
     (unsorted).  All other local numbers are two ranges (dependents, owned) */
  DBuf<int> ext_numbers;
  bool ext_numbers_valid;
  int prefetch;
  NodeData()
      : valid(false), order(2), interp_type(1), num_elements(0),
        num_local_nodes(0), num_dep_nodes(0), num_owned_nodes(0), dep_nnz(0),
        num_candidates(0), node_range_start(0), mctx(NULL), ext_numbers_valid(false),
        prefetch(0) {}
  ~NodeData() { drop_mirrors(); }
  /* the expansion job writes into the DepPtr / DepWeights mirrors: join it
     before anything is released */
  void finish_dep_job() {
    if (dep_job.thread) {
      std::thread *t = static_cast<std::thread *>(dep_job.thread);
      t->join();
      delete t;
      dep_job.thread = NULL;
    }
    if (dep_job.ev_codes) copy_wait(*mctx, dep_job.ev_codes);
    if (dep_job.ev_table) copy_wait(*mctx, dep_job.ev_table);
    dep_job.ev_codes = dep_job.ev_table = NULL;
    if (dep_job.codes) host_free(*mctx, dep_job.codes);
    if (dep_job.table) host_free(*mctx, dep_job.table);
    dep_job.codes = NULL;
    dep_job.table = NULL;
  }
  void drop_mirrors() {
    finish_dep_job();
    for (int k = 0; k < kNumMirrors; k++) {
      mirror[k].join();
      if (mirror[k].ev) copy_wait(*mctx, mirror[k].ev);
      if (mirror[k].p) host_free(*mctx, mirror[k].p);
      if (mirror[k].aux) host_free(*mctx, mirror[k].aux);
      mirror[k] = HostMirror();
    }
    sorted_numbers.reset();
  }
  void clear() {
    drop_mirrors();
    valid = false;
    node_keys.reset();
    node_num.reset();
    conn.reset();
    dep_ptr.reset();
    dep_conn.reset();
    dep_weights.reset();
    dep_code.reset();
    dep_wtab.reset();
    elem_ptr.reset();
    slot_info.reset();
    slot_multi.reset();
    ext_numbers.reset();
    ext_numbers_valid = false;
    node_range.clear();
    num_elements = num_local_nodes = num_dep_nodes = num_owned_nodes = 0;
    dep_nnz = 0;
  }
};

/* prolongation rows built by createInterpolation (fine forest owns them) */
struct InterpData {
  bool valid;
  i64 nrows, nnz;
  DBuf<int> rows; /* global fine node number of each row, in emission order */
  DBuf<int> rowp; /* [nrows + 1] */
  DBuf<int> cols;
  DBuf<double> vals;
  InterpData() : valid(false), nrows(0), nnz(0) {}
  void clear() {
    valid = false;
    nrows = nnz = 0;
    rows.reset();
    rowp.reset();
    cols.reset();
    vals.reset();
  }
};

struct Forest {
  Ctx *ctx;
  /* connectivity tables (device copies) */
  ConnTables tables;
  std::shared_ptr<DBuf<int> > table_store;
  int nblocks;
  int bbits;
  /* elements */
  DBuf<u64> keys;
  DBuf<int16_t> info; /* empty => all zero */
  i64 n;
  KeyFmt fmt;
  NodeData nodes;
  InterpData interp;
  /* SFC partition: owners[r] = first octant of rank r (reference `owners`,
     src/TMROctForest.h:270); empty for a single rank */
  std::vector<Oct24> owners;
  /* a forest built on a one-rank communicator (MPI_COMM_SELF in a multi-rank
     job) lives on this GPU alone: no exchange, whatever the context's
     communicator is */
  bool serial;
  /* counters describing the last operation (for bench/roofline reporting) */
  i64 last_in, last_mid, last_out;

  explicit Forest(Ctx *c)
      : ctx(c), nblocks(0), bbits(1), n(0), serial(false), last_in(0),
        last_mid(0), last_out(0) {
    fmt.D = 0;
    fmt.bbits = 1;
    tables = ConnTables();
    const char *ev = getenv("TMR_B200_NODE_PREFETCH");
    nodes.prefetch = ev ? atoi(ev) : 0;
  }
};

/* the communicator the forest is partitioned over (NULL = this GPU alone) */
inline Comm *forest_comm(const Forest &f) { return f.serial ? NULL : f.ctx->comm; }
inline int part_rank(const Forest &f) { return forest_comm(f) ? forest_comm(f)->rank : 0; }
inline int part_size(const Forest &f) { return forest_comm(f) ? forest_comm(f)->size : 1; }

inline int bits_for(int nblocks) {
  int b = 1;
  while ((1LL << b) < nblocks) b++;
  return b;
}

inline bool key_budget_ok(const Forest &f, int D) {
  return f.bbits + 3 * D + 5 <= 64 && D <= 19;
}

}  // namespace tmrgpu

#endif
