/*
  tests/emu/comm_threads.cpp -- TEST INFRASTRUCTURE ONLY.
  In-process stand-in for the NCCL communicator: ranks are threads of one
  process (each with its own emulated context), collectives rendezvous on a
  barrier and copy straight out of the peers' buffers.  Lets the multi-rank
  exchange logic of ops_multi.h / ops_nodes.h be checked against the oracle's
  thread-ranks on the CPU-only build container.
*/
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <vector>

#include "comm.h"

namespace tmrgpu {

struct ThreadWorld {
  std::mutex mtx;
  bool ready;
  int size;
  pthread_barrier_t bar;
  std::vector<const void *> ptr;
  std::vector<const i64 *> off;
  /* what each rank believes the current collective is (kind, element size):
     ranks that drift apart in their collective sequence would deadlock NCCL;
     here they are caught at the rendezvous */
  std::vector<size_t> tag;
  ThreadWorld() : ready(false), size(0) {}
};

class ThreadComm : public Comm {
 public:
  ThreadWorld *w;
  void check_same_collective(size_t mine) {
    for (int r = 0; r < size; r++) {
      if (w->tag[r] != mine) {
        fprintf(stderr,
                "emulated communicator: rank %d is in collective %zx while rank %d is in "
                "%zx -- the ranks' collective sequences diverged\n",
                rank, mine, r, w->tag[r]);
        abort();
      }
    }
  }
  void allgather_host(Ctx &, const void *send, void *recv, size_t bytes) override {
    w->ptr[rank] = send;
    w->tag[rank] = ((size_t)1 << 60) | bytes;
    pthread_barrier_wait(&w->bar);
    check_same_collective(((size_t)1 << 60) | bytes);
    for (int r = 0; r < size; r++) {
      memcpy((char *)recv + (size_t)r * bytes, w->ptr[r], bytes);
    }
    pthread_barrier_wait(&w->bar);
  }
  void allgather_dev(Ctx &ctx, const void *send, void *recv, size_t bytes) override {
    allgather_host(ctx, send, recv, bytes);
  }
  void alltoallv(Ctx &, const void *send, const i64 *send_off, void *recv,
                 const i64 *recv_off, size_t eb) override {
    w->ptr[rank] = send;
    w->off[rank] = send_off;
    w->tag[rank] = ((size_t)2 << 60) | eb;
    pthread_barrier_wait(&w->bar);
    check_same_collective(((size_t)2 << 60) | eb);
    for (int p = 0; p < size; p++) {
      const i64 cnt = w->off[p][rank + 1] - w->off[p][rank];
      if (cnt > 0) {
        memcpy((char *)recv + (size_t)recv_off[p] * eb,
               (const char *)w->ptr[p] + (size_t)w->off[p][rank] * eb,
               (size_t)cnt * eb);
      }
    }
    pthread_barrier_wait(&w->bar);
  }
};

int comm_unique_id(void *out, int out_bytes) {
  if (out_bytes < (int)sizeof(void *)) return 1;
  ThreadWorld *w = new ThreadWorld();
  memset(out, 0, out_bytes);
  memcpy(out, &w, sizeof(w));
  return 0;
}

Comm *comm_create(Ctx &, int rank, int size, const void *id_bytes) {
  ThreadWorld *w;
  memcpy(&w, id_bytes, sizeof(w));
  {
    std::lock_guard<std::mutex> lk(w->mtx);
    if (!w->ready) {
      w->size = size;
      pthread_barrier_init(&w->bar, NULL, size);
      w->ptr.assign(size, (const void *)0);
      w->off.assign(size, (const i64 *)0);
      w->tag.assign(size, (size_t)0);
      w->ready = true;
    }
  }
  ThreadComm *c = new ThreadComm();
  c->rank = rank;
  c->size = size;
  c->w = w;
  return c;
}

void comm_destroy(Comm *c) { delete static_cast<ThreadComm *>(c); }

}  // namespace tmrgpu
