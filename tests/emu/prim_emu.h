/*
  tests/emu/prim_emu.h -- TEST INFRASTRUCTURE ONLY.

  Serial host stand-ins for the three device primitives (launch, scan_counts,
  radix_sort) so that the kernel BODIES of tmr_b200/csrc/gpu/ops_*.h -- plain
  TMR_HD functors -- and their host orchestration can be run against the
  oracle on a machine without a GPU.  Built only by tests/emu/Makefile into
  tests/emu/_build/; the product library (tmr_b200/lib) never contains or
  loads it.
*/
#ifndef TMRGPU_PRIM_EMU_H
#define TMRGPU_PRIM_EMU_H

#include "prim.h"

namespace tmrgpu {

template <class F>
void launch(Ctx &ctx, i64 n, F f, const char *) {
  for (i64 i = 0; i < n; i++) f(i);
  ctx.launch_count++;
}

/* the per-item form of a warp-cooperative body */
template <class F>
void launch_warp(Ctx &ctx, i64 n, F f, const char *) {
  for (i64 i = 0; i < n; i++) f(i);
  ctx.launch_count++;
}

struct DirectSink {
  u64 *p;
  void operator()(u64 v) { *p++ = v; }
};

template <int kMaxPer, class F>
void expand_u64(Ctx &ctx, i64 n, const u32 *off, u64, F f, u64 *out, const char *) {
  for (i64 i = 0; i < n; i++) {
    DirectSink sink = {out + off[i]};
    f(i, sink);
  }
  ctx.launch_count++;
}


static const int kLaunchThreads = 256;

inline bool ctx_ok(const Ctx &ctx) { return ctx.last_error.empty(); }

template <class F>
void launch_block3(Ctx &ctx, i64 n, F f, const char *) {
  typename F::Shared *sh = new typename F::Shared();
  for (i64 i0 = 0; i0 < n; i0 += kLaunchThreads) {
    const i64 i1 = (i0 + kLaunchThreads < n) ? i0 + kLaunchThreads : n;
    sh->reset();
    for (i64 i = i0; i < i1; i++) f.collect(i, (int)(i - i0), *sh);
    for (int t = 0; t < kLaunchThreads; t++) f.process(i0, t, *sh);
    for (i64 i = i0; i < i1; i++) f.finish(i, (int)(i - i0), *sh);
  }
  delete sh;
  ctx.launch_count++;
}

template <class F>
u64 scan_counts(Ctx &ctx, i64 n, F f, u32 *out, const char *) {
  u64 run = 0;
  for (i64 i = 0; i < n; i++) {
    const u32 c = f(i);
    out[i] = (u32)run;
    run += c;
  }
  ctx.launch_count++;
  return run;
}

template <class F, class G>
u64 scan_apply(Ctx &ctx, i64 n, F f, G g, const char *) {
  typedef decltype(f((i64)0)) T;
  u64 run = 0;
  for (i64 i = 0; i < n; i++) {
    const T c = f(i);
    g(i, (T)run);
    run += c;
  }
  ctx.launch_count++;
  return run;
}

}  // namespace tmrgpu
#endif
