/*
  comm.h -- the exchange layer of the multi-GPU forest.

  The reference partitions the Morton-sorted octant array across MPI ranks and
  moves 24-byte octant records with MPI_Allgather / MPI_Alltoall + Isend/Recv
  (reference src/TMROctForest.cpp:1824,1941,2018-2059,2415,2482-2505, datatypes
  src/TMRBase.cpp:42-92).  Here one process drives one GPU; ranks exchange
  8-byte keys that are already in HBM.  `Comm` is the small surface the forest
  operations need; the product implementation is NCCL over NVLink
  (comm_nccl.cu), a test-only in-process implementation lives in tests/emu.
*/
#ifndef TMRGPU_COMM_H
#define TMRGPU_COMM_H

#include <vector>

#include "prim.h"

namespace tmrgpu {

class Comm {
 public:
  virtual ~Comm() {}
  int rank, size;
  /* every rank contributes `bytes` bytes from HOST memory; recv holds
     size*bytes (small control data: counts, partition keys) */
  virtual void allgather_host(Ctx &ctx, const void *send, void *recv,
                              size_t bytes) = 0;
  /* same with DEVICE buffers, enqueued on the context's stream, no host
     round trip: counts produced by one kernel and consumed by the next (or
     read back together with other results in one copy) */
  virtual void allgather_dev(Ctx &ctx, const void *send, void *recv,
                             size_t bytes) = 0;
  /* variable all-to-all of DEVICE buffers; offsets are in elements of
     elem_bytes bytes, arrays of size+1 entries */
  virtual void alltoallv(Ctx &ctx, const void *send, const i64 *send_off,
                         void *recv, const i64 *recv_off,
                         size_t elem_bytes) = 0;
};

/* exchange per-destination counts: send_counts[r] -> recv_counts[r] */
inline void exchange_counts(Ctx &ctx, Comm &comm, const i64 *send_counts,
                            i64 *recv_counts) {
  const int R = comm.size;
  std::vector<i64> all((size_t)R * R);
  comm.allgather_host(ctx, send_counts, all.data(), (size_t)R * sizeof(i64));
  for (int r = 0; r < R; r++) recv_counts[r] = all[(size_t)r * R + comm.rank];
}

}  // namespace tmrgpu

#endif
