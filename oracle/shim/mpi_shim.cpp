/*
  oracle/shim/mpi_shim.cpp -- TEST INFRASTRUCTURE ONLY.
  Threads-as-ranks implementation of oracle/shim/mpi.h.
*/
#include "mpi.h"

#include <pthread.h>
#include <string.h>
#include <time.h>

#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

namespace {

struct Mailbox {
  std::mutex mtx;
  std::condition_variable cv;
  std::deque<std::vector<char> > msgs;
};

struct World {
  int size = 1;
  pthread_barrier_t barrier;
  bool barrier_live = false;
  std::vector<const void *> slots;
  std::vector<Mailbox *> boxes;  // [dst * size + src]
};

World g_world;
thread_local int t_rank = 0;

std::mutex g_type_mtx;
std::vector<int> g_type_extent;  // derived type handle = 1000 + index
const int kDerivedBase = 1000;

int type_size(MPI_Datatype t) {
  if (t >= kDerivedBase) {
    std::lock_guard<std::mutex> lk(g_type_mtx);
    return g_type_extent[t - kDerivedBase];
  }
  return t;
}

void world_barrier() {
  if (g_world.size > 1) {
    pthread_barrier_wait(&g_world.barrier);
  }
}

}  // namespace

extern "C" {

int MPI_Init(int *, char ***) { return MPI_SUCCESS; }
int MPI_Finalize(void) { return MPI_SUCCESS; }

double MPI_Wtime(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

int MPI_Comm_rank(MPI_Comm comm, int *rank) {
  *rank = (comm == MPI_COMM_SELF) ? 0 : t_rank;
  return MPI_SUCCESS;
}

int MPI_Comm_size(MPI_Comm comm, int *size) {
  *size = (comm == MPI_COMM_SELF) ? 1 : g_world.size;
  return MPI_SUCCESS;
}

int MPI_Barrier(MPI_Comm) {
  world_barrier();
  return MPI_SUCCESS;
}

int MPI_Type_create_struct(int count, const int lens[], const MPI_Aint disp[],
                           const MPI_Datatype types[], MPI_Datatype *newtype) {
  long extent = 0;
  int align = 4;
  for (int i = 0; i < count; i++) {
    int sz = type_size(types[i]);
    long end = disp[i] + (long)lens[i] * sz;
    if (end > extent) extent = end;
    if (sz >= 8) align = 8;
  }
  extent = ((extent + align - 1) / align) * align;
  std::lock_guard<std::mutex> lk(g_type_mtx);
  g_type_extent.push_back((int)extent);
  *newtype = kDerivedBase + (int)g_type_extent.size() - 1;
  return MPI_SUCCESS;
}

int MPI_Type_commit(MPI_Datatype *) { return MPI_SUCCESS; }
int MPI_Type_free(MPI_Datatype *) { return MPI_SUCCESS; }

int MPI_Allgather(const void *sbuf, int scount, MPI_Datatype stype, void *rbuf,
                  int, MPI_Datatype, MPI_Comm comm) {
  const size_t nbytes = (size_t)scount * type_size(stype);
  if (comm == MPI_COMM_SELF || g_world.size == 1) {
    memmove(rbuf, sbuf, nbytes);
    return MPI_SUCCESS;
  }
  g_world.slots[t_rank] = sbuf;
  world_barrier();
  for (int r = 0; r < g_world.size; r++) {
    memcpy((char *)rbuf + r * nbytes, g_world.slots[r], nbytes);
  }
  world_barrier();
  return MPI_SUCCESS;
}

int MPI_Alltoall(const void *sbuf, int scount, MPI_Datatype stype, void *rbuf,
                 int, MPI_Datatype, MPI_Comm comm) {
  const size_t nbytes = (size_t)scount * type_size(stype);
  if (comm == MPI_COMM_SELF || g_world.size == 1) {
    memmove(rbuf, sbuf, nbytes);
    return MPI_SUCCESS;
  }
  g_world.slots[t_rank] = sbuf;
  world_barrier();
  for (int r = 0; r < g_world.size; r++) {
    memcpy((char *)rbuf + r * nbytes,
           (const char *)g_world.slots[r] + t_rank * nbytes, nbytes);
  }
  world_barrier();
  return MPI_SUCCESS;
}

int MPI_Isend(const void *buf, int count, MPI_Datatype type, int dest, int,
              MPI_Comm, MPI_Request *req) {
  const size_t nbytes = (size_t)count * type_size(type);
  Mailbox *box = g_world.boxes[(size_t)dest * g_world.size + t_rank];
  {
    std::lock_guard<std::mutex> lk(box->mtx);
    box->msgs.emplace_back((const char *)buf, (const char *)buf + nbytes);
  }
  box->cv.notify_one();
  if (req) *req = 0;
  return MPI_SUCCESS;
}

int MPI_Recv(void *buf, int count, MPI_Datatype type, int src, int, MPI_Comm,
             MPI_Status *) {
  const size_t nbytes = (size_t)count * type_size(type);
  Mailbox *box = g_world.boxes[(size_t)t_rank * g_world.size + src];
  std::unique_lock<std::mutex> lk(box->mtx);
  box->cv.wait(lk, [box] { return !box->msgs.empty(); });
  std::vector<char> &m = box->msgs.front();
  memcpy(buf, m.data(), m.size() < nbytes ? m.size() : nbytes);
  box->msgs.pop_front();
  return MPI_SUCCESS;
}

int MPI_Waitall(int, MPI_Request[], MPI_Status[]) { return MPI_SUCCESS; }

void shim_barrier(void) { world_barrier(); }
int shim_rank(void) { return t_rank; }
int shim_size(void) { return g_world.size; }

void shim_world_begin(int nranks) {
  if (nranks < 1) nranks = 1;
  g_world.size = nranks;
  g_world.slots.assign(nranks, (const void *)0);
  for (size_t i = 0; i < g_world.boxes.size(); i++) delete g_world.boxes[i];
  g_world.boxes.assign((size_t)nranks * nranks, (Mailbox *)0);
  for (size_t i = 0; i < g_world.boxes.size(); i++) {
    g_world.boxes[i] = new Mailbox();
  }
  if (nranks > 1) {
    pthread_barrier_init(&g_world.barrier, NULL, nranks);
    g_world.barrier_live = true;
  }
}

void shim_attach(int rank) { t_rank = rank; }

void shim_world_end(void) {
  if (g_world.barrier_live) {
    pthread_barrier_destroy(&g_world.barrier);
    g_world.barrier_live = false;
  }
  g_world.size = 1;
  t_rank = 0;
}

void shim_run(int nranks, shim_rank_fn fn, void *arg) {
  shim_world_begin(nranks);
  std::vector<std::thread> threads;
  for (int r = 0; r < nranks; r++) {
    threads.emplace_back([r, fn, arg] {
      t_rank = r;
      fn(r, arg);
    });
  }
  for (size_t i = 0; i < threads.size(); i++) threads[i].join();
  shim_world_end();
}

}  // extern "C"
