/*
  comm_nccl.cu -- Comm over NCCL (NVLink 5 / NVSwitch): one communicator per
  process, all transfers enqueued on the context's stream.  all-to-all-v =
  one ncclGroup of ncclSend/ncclRecv per peer on 8-byte keys (the reference
  ships 24-byte records, src/TMROctForest.cpp:2482-2505); the self segment is
  a device-to-device copy.
*/
#include <cuda_runtime.h>
#include <nccl.h>
#include <stdio.h>
#include <string.h>

#include "comm.h"

namespace tmrgpu {

#define TMR_NCCL_OK(call)                                              \
  do {                                                                 \
    ncclResult_t r_ = (call);                                          \
    if (r_ != ncclSuccess) {                                           \
      fprintf(stderr, "TMROctForest Error: NCCL %s at %s:%d\n",        \
              ncclGetErrorString(r_), __FILE__, __LINE__);             \
      ctx.last_error = "NCCL failure";                                 \
    }                                                                  \
  } while (0)

class NcclComm : public Comm {
 public:
  ncclComm_t comm;
  void allgather_host(Ctx &ctx, const void *send, void *recv,
                      size_t bytes) override {
    if (size == 1) {
      memcpy(recv, send, bytes);
      return;
    }
    /* page-locked staging (cached by the context), so the upload does not
       need its own synchronisation: ONE blocking round trip per call */
    unsigned char *d = static_cast<unsigned char *>(
        dev_alloc(ctx, bytes * (size_t)(size + 1)));
    unsigned char *h = static_cast<unsigned char *>(
        host_alloc(ctx, bytes * (size_t)(size + 1)));
    if (!d || !h) return;
    memcpy(h, send, bytes);
    cudaStream_t st = (cudaStream_t)ctx.stream;
    cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, st);
    TMR_NCCL_OK(ncclAllGather(d, d + bytes, bytes, ncclChar, comm, st));
    copy_d2h(ctx, h + bytes, d + bytes, bytes * (size_t)size);
    memcpy(recv, h + bytes, bytes * (size_t)size);
    host_free(ctx, h);
    dev_free(ctx, d);
  }
  void allgather_dev(Ctx &ctx, const void *send, void *recv,
                     size_t bytes) override {
    if (size == 1) {
      copy_d2d(ctx, recv, send, bytes);
      return;
    }
    TMR_NCCL_OK(ncclAllGather(send, recv, bytes, ncclChar, comm,
                              (cudaStream_t)ctx.stream));
  }
  void alltoallv(Ctx &ctx, const void *send, const i64 *send_off, void *recv,
                 const i64 *recv_off, size_t elem_bytes) override {
    const unsigned char *s = static_cast<const unsigned char *>(send);
    unsigned char *r = static_cast<unsigned char *>(recv);
    cudaStream_t st = (cudaStream_t)ctx.stream;
    TMR_NCCL_OK(ncclGroupStart());
    for (int p = 0; p < size; p++) {
      const size_t sb = (size_t)(send_off[p + 1] - send_off[p]) * elem_bytes;
      const size_t rb = (size_t)(recv_off[p + 1] - recv_off[p]) * elem_bytes;
      if (p == rank) continue;
      if (sb) {
        TMR_NCCL_OK(ncclSend(s + (size_t)send_off[p] * elem_bytes, sb, ncclChar,
                             p, comm, st));
      }
      if (rb) {
        TMR_NCCL_OK(ncclRecv(r + (size_t)recv_off[p] * elem_bytes, rb, ncclChar,
                             p, comm, st));
      }
    }
    TMR_NCCL_OK(ncclGroupEnd());
    const size_t self = (size_t)(send_off[rank + 1] - send_off[rank]) * elem_bytes;
    if (self) {
      copy_d2d(ctx, r + (size_t)recv_off[rank] * elem_bytes,
               s + (size_t)send_off[rank] * elem_bytes, self);
    }
  }
};

int comm_unique_id(void *out, int out_bytes) {
  if (out_bytes < (int)sizeof(ncclUniqueId)) return 1;
  ncclUniqueId id;
  if (ncclGetUniqueId(&id) != ncclSuccess) return 1;
  memcpy(out, &id, sizeof(id));
  return 0;
}

Comm *comm_create(Ctx &ctx, int rank, int size, const void *id_bytes) {
  NcclComm *c = new NcclComm();
  c->rank = rank;
  c->size = size;
  c->comm = NULL;
  if (size > 1) {
    ncclUniqueId id;
    memcpy(&id, id_bytes, sizeof(id));
    ncclResult_t r = ncclCommInitRank(&c->comm, size, id, rank);
    if (r != ncclSuccess) {
      fprintf(stderr, "TMROctForest Error: ncclCommInitRank: %s\n",
              ncclGetErrorString(r));
      delete c;
      return NULL;
    }
  }
  (void)ctx;
  return c;
}

void comm_destroy(Comm *c) {
  NcclComm *n = static_cast<NcclComm *>(c);
  if (n && n->comm) ncclCommDestroy(n->comm);
  delete n;
}

}  // namespace tmrgpu
