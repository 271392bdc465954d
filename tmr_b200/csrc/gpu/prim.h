/*
  prim.h -- device-runtime plumbing shared by the forest operations: context,
  owning device buffers, copies, per-kernel timing.  The parallel primitives
  themselves (launch / scan_counts / radix sort) are declared in
  prim_cuda.cuh; the forest operations in forest_ops.h are written against
  exactly this small surface.

  A test-only host emulation of the same surface lives in tests/emu/ (it lets
  the kernel BODIES -- plain TMR_HD functors -- be exercised against the
  oracle on a machine without a GPU).  The product library never contains it.
*/
#ifndef TMRGPU_PRIM_H
#define TMRGPU_PRIM_H

#include <stddef.h>
#include <stdio.h>

#include <map>
#include <string>
#include <vector>

#include "common.h"

namespace tmrgpu {

struct KernelStat {
  long launches;
  double ms;
  KernelStat() : launches(0), ms(0.0) {}
};

class Comm;

struct Ctx {
  int device;
  Comm *comm; /* NULL = single rank */
  void *stream; /* cudaStream_t */
  void *copy_stream; /* second stream for device->host mirrors, lazily created */
  void *copy_stream2; /* lane 1: copies done by an SM kernel into page-locked memory (not queued
                         behind the copy engine's large transfers) */
  /* page-locked, device-visible words the kernels write results the host
     waits for into (scan totals, small read-backs): a PCIe store instead of a
     copy-engine job that would queue behind a 2.8 GB mirror transfer */
  unsigned long long *mailbox;
  int mailbox_next;
  int profile;  /* time every named launch with events */
  int num_sms;
  long launch_count; /* kernels launched since the last reset */
  long sync_count;   /* blocking host<->device round trips since the last reset */
  long long bytes_d2h, bytes_h2d; /* bytes copied over the bus since the last reset */
  std::map<std::string, KernelStat> stats;
  /* pending (start, stop) event pairs, resolved lazily at the next sync */
  std::vector<void *> ev_start, ev_stop;
  std::vector<std::string> ev_name;
  std::string last_error;
  long fail_alloc_in; /* fault injection for tests: the n-th next dev_alloc fails (0 = off) */
  FILE *launch_log; /* TMR_B200_LAUNCH_LOG=<file>: "<launch index> <name>" per profiled bracket */
  int trace;        /* TMR_B200_TRACE=1: print synchronised phase times */
  double trace_t0;  /* wall clock of the previous mark (s) */
  Ctx()
      : device(0), comm(NULL), stream(NULL), copy_stream(NULL), copy_stream2(NULL),
        mailbox(NULL), mailbox_next(0), profile(0),
        num_sms(148),
        launch_count(0), sync_count(0), bytes_d2h(0), bytes_h2d(0), fail_alloc_in(0), launch_log(NULL), trace(0),
        trace_t0(0.0) {}
};

/* --- runtime (prim_cuda.cu / tests/emu/prim_emu.cpp) ---------------------- */
void *dev_alloc(Ctx &ctx, size_t bytes);
void dev_free(Ctx &ctx, void *p);
void dev_cache_destroy(Ctx &ctx); /* return every cached block to the driver */
void *host_alloc(Ctx &ctx, size_t bytes); /* page-locked, cached */
void host_free(Ctx &ctx, void *p);
void copy_h2d(Ctx &ctx, void *dst, const void *src, size_t bytes);
void copy_d2h(Ctx &ctx, void *dst, const void *src, size_t bytes); /* syncs */
void copy_d2d(Ctx &ctx, void *dst, const void *src, size_t bytes);
/* device->host copy on the context's COPY stream, ordered after everything
   enqueued so far on the main stream and overlapping what follows there; dst
   must be page-locked (host_alloc).  Returns a handle for copy_wait. */
void *copy_d2h_async(Ctx &ctx, void *dst, const void *src, size_t bytes, int lane = 0);
void copy_wait(Ctx &ctx, void *handle); /* blocks until done, releases it */
/* next 8-byte mailbox slot (device-visible host memory), zeroed */
unsigned long long *mailbox_slot(Ctx &ctx);
/* blocks until done, keeps the handle; callable from a worker thread */
void copy_sync(Ctx &ctx, void *handle);
void dev_zero(Ctx &ctx, void *p, size_t bytes);
void dev_fill_ff(Ctx &ctx, void *p, size_t bytes);
void stream_sync(Ctx &ctx);
/* returns 0 when no CUDA error is pending; otherwise records it in
   ctx.last_error, prints "TMROctForest Error: ..." and returns nonzero */
int check_errors(Ctx &ctx, const char *where);
void prof_begin(Ctx &ctx, const char *name);
void prof_end(Ctx &ctx);
void prof_resolve(Ctx &ctx);
/* when tracing: synchronise and print the wall time since the last mark */
void trace_mark(Ctx &ctx, const char *label);

/* Owning, move-only device array */
template <class T>
class DBuf {
 public:
  DBuf() : ctx_(NULL), p_(NULL), n_(0) {}
  DBuf(Ctx &ctx, i64 n) : ctx_(&ctx), p_(NULL), n_(n) {
    if (n > 0) p_ = static_cast<T *>(dev_alloc(ctx, (size_t)n * sizeof(T)));
  }
  ~DBuf() { reset(); }
  DBuf(DBuf &&o) : ctx_(o.ctx_), p_(o.p_), n_(o.n_) {
    o.p_ = NULL;
    o.n_ = 0;
  }
  DBuf &operator=(DBuf &&o) {
    if (this != &o) {
      reset();
      ctx_ = o.ctx_;
      p_ = o.p_;
      n_ = o.n_;
      o.p_ = NULL;
      o.n_ = 0;
    }
    return *this;
  }
  void reset() {
    if (p_) dev_free(*ctx_, p_);
    p_ = NULL;
    n_ = 0;
  }
  void alloc(Ctx &ctx, i64 n) {
    reset();
    ctx_ = &ctx;
    n_ = n;
    if (n > 0) p_ = static_cast<T *>(dev_alloc(ctx, (size_t)n * sizeof(T)));
  }
  T *get() const { return p_; }
  i64 size() const { return n_; }
  /* shrink the logical size without reallocating */
  void set_size(i64 n) { n_ = n; }
  void swap(DBuf &o) {
    Ctx *c = ctx_;
    T *p = p_;
    i64 n = n_;
    ctx_ = o.ctx_;
    p_ = o.p_;
    n_ = o.n_;
    o.ctx_ = c;
    o.p_ = p;
    o.n_ = n;
  }

 private:
  DBuf(const DBuf &);
  DBuf &operator=(const DBuf &);
  Ctx *ctx_;
  T *p_;
  i64 n_;
};

/* --- radix sort (prim_cuda.cu) -------------------------------------------
   Stable LSD radix sort of 64-bit keys on bits [bit_lo, bit_hi), 8 or 9 bits
   per pass, ping-ponging between the two buffers.  On return `keys` (and `vals`)
   hold the sorted data; `keys_alt`/`vals_alt` are scratch of the same size.
   vals may be empty (keys only). */
void radix_sort(Ctx &ctx, DBuf<u64> &keys, DBuf<u64> &keys_alt,
                DBuf<u32> &vals, DBuf<u32> &vals_alt, i64 n, int bit_lo,
                int bit_hi, const char *tag = NULL /* profiling label */);

}  // namespace tmrgpu

#endif
