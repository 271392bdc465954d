/*
  prim_cuda.cu -- CUDA runtime plumbing (caching allocator, copies, profiling
  hooks) and the hand-written LSD radix sort that replaces qsort+comparator of
  reference src/TMROctant.cpp:357-399 (design notes at the radix section).
*/
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <mutex>
#include <thread>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "prim_cuda.cuh"

namespace tmrgpu {

/* ------------------------------------------------------------------------ */
/* runtime                                                                  */
/* ------------------------------------------------------------------------ */
/* a failed runtime call is printed AND recorded in the context: the current
   operation stops launching (prim_cuda.cuh: ctx_ok) and its check_errors
   reports the failure to the caller */
#define TMR_CUDA_OK(call)                                                     \
  do {                                                                        \
    cudaError_t e_ = (call);                                                  \
    if (e_ != cudaSuccess) {                                                  \
      fprintf(stderr, "TMROctForest Error: CUDA %s at %s:%d\n",               \
              cudaGetErrorString(e_), __FILE__, __LINE__);                    \
      if (ctx.last_error.empty()) {                                           \
        ctx.last_error = std::string("CUDA ") + cudaGetErrorString(e_);       \
      }                                                                       \
    }                                                                         \
  } while (0)

/* Device memory: a size-class caching allocator in front of cudaMalloc.
   Every forest call allocates and frees multi-GB scratch arrays whose sizes
   change a little from call to call; going to the driver for those costs
   ~0.3 ms/MB (measured: an 850 ms stall inside a 137 ms step with
   cudaMallocAsync).  Blocks are rounded up to (8+k)*2^m bytes (<= 12.5 %
   slack) so a slightly different request reuses a cached block.  All work of a
   context runs on ONE stream, so a freed block can be handed out again
   immediately: stream order already serialises the old and new users. */
static size_t size_class(size_t bytes) {
  if (bytes < 512) return 512;
  int m = 0;
  while ((bytes >> m) >= 16) m++;
  /* now bytes >> m is in [8, 16) */
  size_t c = ((bytes + ((size_t)1 << m) - 1) >> m) << m;
  return c;
}

struct DevCache {
  std::multimap<size_t, void *> free_blocks; /* class -> block */
  std::map<void *, size_t> live;             /* block -> class */
  size_t cached_bytes, live_bytes, peak_bytes;
  DevCache() : cached_bytes(0), live_bytes(0), peak_bytes(0) {}
};

static std::map<Ctx *, DevCache> g_caches;
/* contexts are per thread (host/TMROctant.cpp); the maps that find a
   context's cache are shared: look-ups are serialised, the cache itself is
   only ever used by its context's thread (std::map nodes do not move) */
static std::mutex g_cache_mutex;
static DevCache &cache_of(std::map<Ctx *, DevCache> &m, Ctx &ctx) {
  std::lock_guard<std::mutex> lock(g_cache_mutex);
  return m[&ctx];
}

static void cache_release_all(DevCache &c) {
  for (std::multimap<size_t, void *>::iterator it = c.free_blocks.begin();
       it != c.free_blocks.end(); ++it) {
    cudaFree(it->second);
  }
  c.free_blocks.clear();
  c.cached_bytes = 0;
}

void dev_cache_destroy(Ctx &ctx) {
  std::lock_guard<std::mutex> lock(g_cache_mutex);
  std::map<Ctx *, DevCache>::iterator it = g_caches.find(&ctx);
  if (it == g_caches.end()) return;
  cudaStreamSynchronize((cudaStream_t)ctx.stream);
  cache_release_all(it->second);
  g_caches.erase(it);
}

size_t dev_cache_peak_bytes(Ctx &ctx) { return cache_of(g_caches, ctx).peak_bytes; }

void *dev_alloc(Ctx &ctx, size_t bytes) {
  if (ctx.fail_alloc_in > 0 && --ctx.fail_alloc_in == 0) {
    fprintf(stderr, "TMROctForest Error: device allocation of %zu bytes failed "
                    "(injected by tmrgpu_test_fail_alloc)\n", bytes);
    ctx.last_error = "device allocation failed";
    return NULL;
  }
  DevCache &c = cache_of(g_caches, ctx);
  const size_t cls = size_class(bytes);
  void *p = NULL;
  std::multimap<size_t, void *>::iterator it = c.free_blocks.find(cls);
  if (it != c.free_blocks.end()) {
    p = it->second;
    c.free_blocks.erase(it);
    c.cached_bytes -= cls;
  } else {
    cudaError_t e = cudaMalloc(&p, cls);
    if (e != cudaSuccess) {
      /* give cached blocks back to the driver and retry once */
      cudaGetLastError();
      cudaStreamSynchronize((cudaStream_t)ctx.stream);
      cache_release_all(c);
      e = cudaMalloc(&p, cls);
    }
    if (e != cudaSuccess) {
      cudaGetLastError();
      fprintf(stderr,
              "TMROctForest Error: device allocation of %zu bytes failed (%s)\n",
              cls, cudaGetErrorString(e));
      ctx.last_error = "device allocation failed";
      return NULL;
    }
  }
  c.live[p] = cls;
  c.live_bytes += cls;
  if (c.live_bytes + c.cached_bytes > c.peak_bytes) {
    c.peak_bytes = c.live_bytes + c.cached_bytes;
  }
  return p;
}

void dev_free(Ctx &ctx, void *p) {
  if (!p) return;
  DevCache &c = cache_of(g_caches, ctx);
  std::map<void *, size_t>::iterator it = c.live.find(p);
  if (it == c.live.end()) return;
  const size_t cls = it->second;
  c.live.erase(it);
  c.live_bytes -= cls;
  c.free_blocks.insert(std::make_pair(cls, p));
  c.cached_bytes += cls;
}

/* page-locked host buffers, cached the same way (cudaHostAlloc is ~1 s/GB) */
static std::map<Ctx *, DevCache> g_host_caches;

void *host_alloc(Ctx &ctx, size_t bytes) {
  DevCache &c = cache_of(g_host_caches, ctx);
  const size_t cls = size_class(bytes);
  void *p = NULL;
  std::multimap<size_t, void *>::iterator it = c.free_blocks.find(cls);
  if (it != c.free_blocks.end()) {
    p = it->second;
    c.free_blocks.erase(it);
  } else if (cudaHostAlloc(&p, cls, cudaHostAllocDefault) != cudaSuccess) {
    cudaGetLastError();
    fprintf(stderr,
            "TMROctForest Error: page-locked host allocation of %zu bytes "
            "failed\n", cls);
    return NULL;
  }
  c.live[p] = cls;
  return p;
}

void host_free(Ctx &ctx, void *p) {
  if (!p) return;
  DevCache &c = cache_of(g_host_caches, ctx);
  std::map<void *, size_t>::iterator it = c.live.find(p);
  if (it == c.live.end()) return;
  c.free_blocks.insert(std::make_pair(it->second, p));
  c.live.erase(it);
}

void copy_h2d(Ctx &ctx, void *dst, const void *src, size_t bytes) {
  if (bytes == 0 || !dst || !src || !ctx.last_error.empty()) return;
  TMR_CUDA_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice,
                              (cudaStream_t)ctx.stream));
  ctx.bytes_h2d += (long long)bytes;
  /* the source may be pageable/stack memory: make the copy complete before
     the caller can reuse it */
  TMR_CUDA_OK(cudaStreamSynchronize((cudaStream_t)ctx.stream));
  ctx.sync_count++;
}

static const int kMailboxWords = 1024; /* 8 KB */

static bool mailbox_init(Ctx &ctx) {
  if (ctx.mailbox) return true;
  void *p = NULL;
  if (cudaHostAlloc(&p, kMailboxWords * sizeof(unsigned long long),
                    cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  ctx.mailbox = static_cast<unsigned long long *>(p);
  return true;
}

unsigned long long *mailbox_slot(Ctx &ctx) {
  if (!mailbox_init(ctx)) return NULL;
  /* the upper half is the staging area of small copy_d2h calls */
  ctx.mailbox_next = (ctx.mailbox_next + 1) % (kMailboxWords / 2);
  ctx.mailbox[ctx.mailbox_next] = 0;
  return ctx.mailbox + ctx.mailbox_next;
}

__global__ void small_copy_kernel(unsigned char *dst, const unsigned char *src, int bytes) {
  for (int i = threadIdx.x; i < bytes; i += blockDim.x) dst[i] = src[i];
}

/* grid-stride 16-byte copy into device-visible host memory (lane 1) */
__global__ void sm_copy_kernel(uint4 *dst, const uint4 *src, size_t n16,
                               unsigned char *dst_tail, const unsigned char *src_tail,
                               int tail) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
    dst[i] = src[i];
  }
  if (blockIdx.x == 0 && (int)threadIdx.x < tail) dst_tail[threadIdx.x] = src_tail[threadIdx.x];
}

/* Large downloads into PAGEABLE memory (arrays the caller allocated with new /
   malloc / numpy): the driver's own staging reaches 1-2 GB/s here (measured: the
   18 GB prolongation CSR of the C3 hierarchy in 12-20 s).  Instead: DMA into
   two page-locked staging buffers in turn while host threads copy the previous
   chunk out. */
static bool dst_is_pageable(const void *p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return true;
  }
  return a.type == cudaMemoryTypeUnregistered;
}

static void parallel_memcpy(void *dst, const void *src, size_t bytes, int T) {
  std::vector<std::thread> th;
  for (int t = 1; t < T; t++) {
    const size_t a = bytes * t / T, b = bytes * (t + 1) / T;
    th.push_back(std::thread([=]() {
      memcpy(static_cast<char *>(dst) + a, static_cast<const char *>(src) + a, b - a);
    }));
  }
  memcpy(dst, src, bytes / T);
  for (size_t k = 0; k < th.size(); k++) th[k].join();
}

static void copy_d2h_staged(Ctx &ctx, void *dst, const void *src, size_t bytes) {
  const size_t chunk = (size_t)256 << 20;
  void *stage[2] = {host_alloc(ctx, chunk), host_alloc(ctx, chunk)};
  cudaEvent_t done[2];
  cudaStream_t st = (cudaStream_t)ctx.stream;
  if (!stage[0] || !stage[1]) {
    TMR_CUDA_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st));
    TMR_CUDA_OK(cudaStreamSynchronize(st));
    if (stage[0]) host_free(ctx, stage[0]);
    if (stage[1]) host_free(ctx, stage[1]);
    return;
  }
  for (int k = 0; k < 2; k++) cudaEventCreateWithFlags(&done[k], cudaEventDisableTiming);
  int T = (int)std::thread::hardware_concurrency();
  if (T > 8) T = 8;
  if (T < 1) T = 1;
  const size_t nchunks = (bytes + chunk - 1) / chunk;
  const char *s8 = static_cast<const char *>(src);
  char *d8 = static_cast<char *>(dst);
  for (size_t c = 0; c <= nchunks; c++) {
    if (c < nchunks) {
      const size_t off = c * chunk, len = bytes - off < chunk ? bytes - off : chunk;
      TMR_CUDA_OK(cudaMemcpyAsync(stage[c & 1], s8 + off, len, cudaMemcpyDeviceToHost, st));
      TMR_CUDA_OK(cudaEventRecord(done[c & 1], st));
    }
    if (c > 0) {
      const size_t p = c - 1, off = p * chunk, len = bytes - off < chunk ? bytes - off : chunk;
      TMR_CUDA_OK(cudaEventSynchronize(done[p & 1]));
      parallel_memcpy(d8 + off, stage[p & 1], len, T);
    }
  }
  for (int k = 0; k < 2; k++) {
    cudaEventDestroy(done[k]);
    host_free(ctx, stage[k]);
  }
}

void copy_d2h(Ctx &ctx, void *dst, const void *src, size_t bytes) {
  if (bytes == 0 || !dst || !src || !ctx.last_error.empty()) return;
  const size_t half = (kMailboxWords / 2) * sizeof(unsigned long long);
  if (bytes <= half && mailbox_init(ctx)) {
    /* small read-back: an SM store into page-locked memory, independent of
       what the copy engines are busy with */
    unsigned char *stage = reinterpret_cast<unsigned char *>(ctx.mailbox + kMailboxWords / 2);
    small_copy_kernel<<<1, 128, 0, (cudaStream_t)ctx.stream>>>(
        stage, static_cast<const unsigned char *>(src), (int)bytes);
    TMR_CUDA_OK(cudaStreamSynchronize((cudaStream_t)ctx.stream));
    memcpy(dst, stage, bytes);
  } else if (bytes >= ((size_t)64 << 20) && dst_is_pageable(dst)) {
    copy_d2h_staged(ctx, dst, src, bytes);
  } else {
    TMR_CUDA_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost,
                                (cudaStream_t)ctx.stream));
    TMR_CUDA_OK(cudaStreamSynchronize((cudaStream_t)ctx.stream));
  }
  ctx.sync_count++;
  ctx.bytes_d2h += (long long)bytes;
}

void *copy_d2h_async(Ctx &ctx, void *dst, const void *src, size_t bytes, int lane) {
  if (bytes == 0 || !dst || !src || !ctx.last_error.empty()) return NULL;
  void *&slot = lane ? ctx.copy_stream2 : ctx.copy_stream;
  if (!slot) {
    cudaStream_t s;
    TMR_CUDA_OK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    slot = s;
  }
  const cudaStream_t cs = (cudaStream_t)slot;
  cudaEvent_t ready, done;
  TMR_CUDA_OK(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
  TMR_CUDA_OK(cudaEventCreateWithFlags(&done, cudaEventDisableTiming));
  TMR_CUDA_OK(cudaEventRecord(ready, (cudaStream_t)ctx.stream));
  TMR_CUDA_OK(cudaStreamWaitEvent(cs, ready, 0));
  if (lane && (reinterpret_cast<size_t>(dst) % 16 == 0) &&
      (reinterpret_cast<size_t>(src) % 16 == 0)) {
    /* dst is page-locked (host_alloc) and, under unified addressing,
       device-visible at the same address */
    const size_t n16 = bytes / 16;
    const int tail = (int)(bytes % 16);
    sm_copy_kernel<<<64, 256, 0, cs>>>(
        static_cast<uint4 *>(dst), static_cast<const uint4 *>(src), n16,
        static_cast<unsigned char *>(dst) + n16 * 16,
        static_cast<const unsigned char *>(src) + n16 * 16, tail);
  } else {
    TMR_CUDA_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, cs));
  }
  ctx.bytes_d2h += (long long)bytes;
  TMR_CUDA_OK(cudaEventRecord(done, cs));
  cudaEventDestroy(ready);
  return done;
}

void copy_wait(Ctx &ctx, void *handle) {
  if (!handle) return;
  TMR_CUDA_OK(cudaEventSynchronize((cudaEvent_t)handle));
  cudaEventDestroy((cudaEvent_t)handle);
}

void copy_sync(Ctx &ctx, void *handle) {
  if (!handle) return;
  cudaSetDevice(ctx.device); /* worker threads start on device 0 */
  cudaEventSynchronize((cudaEvent_t)handle);
}

void copy_d2d(Ctx &ctx, void *dst, const void *src, size_t bytes) {
  if (bytes == 0 || !dst || !src || !ctx.last_error.empty()) return;
  TMR_CUDA_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice,
                              (cudaStream_t)ctx.stream));
}

void dev_zero(Ctx &ctx, void *p, size_t bytes) {
  if (bytes == 0 || !p || !ctx.last_error.empty()) return;
  TMR_CUDA_OK(cudaMemsetAsync(p, 0, bytes, (cudaStream_t)ctx.stream));
}

void dev_fill_ff(Ctx &ctx, void *p, size_t bytes) {
  if (bytes == 0 || !p || !ctx.last_error.empty()) return;
  TMR_CUDA_OK(cudaMemsetAsync(p, 0xff, bytes, (cudaStream_t)ctx.stream));
}

void stream_sync(Ctx &ctx) {
  TMR_CUDA_OK(cudaStreamSynchronize((cudaStream_t)ctx.stream));
  ctx.sync_count++;
}

int check_errors(Ctx &ctx, const char *where) {
  cudaError_t e = cudaStreamSynchronize((cudaStream_t)ctx.stream);
  ctx.sync_count++;
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) {
    fprintf(stderr, "TMROctForest Error: CUDA failure in %s: %s\n", where,
            cudaGetErrorString(e));
    ctx.last_error.clear();
    return 1;
  }
  if (!ctx.last_error.empty()) {
    /* recorded by the operation that just ended (allocation failure, size
       limit, NCCL): reported to its caller here, once -- the next operation
       starts clean */
    fprintf(stderr, "TMROctForest Error: %s failed: %s\n", where,
            ctx.last_error.c_str());
    ctx.last_error.clear();
    return 1;
  }
  return 0;
}

/* TMR_B200_NVTX=1: every named launch bracket is an NVTX range (shows up in
   nsys timelines and lets ncu filter with --nvtx-include) */
static int nvtx_mode() {
  static int mode = -1;
  if (mode < 0) mode = getenv("TMR_B200_NVTX") ? atoi(getenv("TMR_B200_NVTX")) : 0;
  return mode;
}

void prof_begin(Ctx &ctx, const char *name) {
  if (nvtx_mode()) nvtxRangePushA(name);
  if (!ctx.profile) return;
  if (ctx.launch_log) fprintf(ctx.launch_log, "%ld %s\n", ctx.launch_count, name);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  cudaEventRecord(a, (cudaStream_t)ctx.stream);
  ctx.ev_start.push_back(a);
  ctx.ev_stop.push_back(b);
  ctx.ev_name.push_back(name);
}

void prof_end(Ctx &ctx) {
  if (nvtx_mode()) nvtxRangePop();
  if (!ctx.profile) return;
  cudaEventRecord((cudaEvent_t)ctx.ev_stop.back(), (cudaStream_t)ctx.stream);
}

void prof_resolve(Ctx &ctx) {
  if (ctx.ev_start.empty()) return;
  cudaStreamSynchronize((cudaStream_t)ctx.stream);
  for (size_t i = 0; i < ctx.ev_start.size(); i++) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, (cudaEvent_t)ctx.ev_start[i],
                         (cudaEvent_t)ctx.ev_stop[i]);
    KernelStat &s = ctx.stats[ctx.ev_name[i]];
    s.launches++;
    s.ms += ms;
    cudaEventDestroy((cudaEvent_t)ctx.ev_start[i]);
    cudaEventDestroy((cudaEvent_t)ctx.ev_stop[i]);
  }
  ctx.ev_start.clear();
  ctx.ev_stop.clear();
  ctx.ev_name.clear();
}

void trace_mark(Ctx &ctx, const char *label) {
  if (!ctx.trace) return;
  cudaStreamSynchronize((cudaStream_t)ctx.stream);
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  const double now = ts.tv_sec + 1e-9 * ts.tv_nsec;
  if (label) {
    fprintf(stderr, "[tmr_b200 trace] %-28s %9.3f ms\n", label,
            1e3 * (now - ctx.trace_t0));
  }
  ctx.trace_t0 = now;
}

/* ------------------------------------------------------------------------ */
/* radix sort                                                               */
/* ------------------------------------------------------------------------ */
/*
  LSD radix sort, 8- or 9-bit digits, 4096-key tiles.  Per pass:
    radix_tile_hist_kernel   digit histogram of every tile (one read of the
                             keys at streaming speed, shared-memory REDs)
    radix_tile_scan_kernel   exclusive prefix of each digit column within
    radix_chunk_scan_kernel  chunks of 256 tiles, then across chunks and digits
    radix_pass_kernel        rank the tile's keys (stable), stage them in shared
                             memory in digit order, write every digit run to
                             its global position as coalesced segments
  HBM traffic per pass: 8 B (histogram) + 8 B + 8 B per key (+4+4 with a
  payload) and 4 KB of table per tile.

  Why not one-sweep with decoupled look-back (the first version): measured on
  B200 (ncu, profiles/ncu_r01_top_kernels.md) the pass kernel spent 59 % of its
  warp time waiting for predecessor tiles -- the prefix advances a few tiles
  per L2 round trip while a 4096-key tile is processed faster than that -- and
  2.77 ms per pass on 333 M keys became 1.89 + 0.52 ms with the table.

  Why not match.any for the ranking: tools/microbench_ops.cu measures ~2 clk
  per DISTINCT value per SM (61 clk for a row of random 8-bit digits, 2 for
  equal ones) against 3.2 clk for a shared-memory RED, so peers are found
  through per-warp bit masks in shared memory unless the row is dominated by a
  few digits (where the masks would serialise on one bank and match.any is
  cheap).
*/
static const int kMaxRadixBits = 9;           /* up to 512 digits per pass */
static const int kMaxRadix = 1 << kMaxRadixBits;
static const int kSortThreads = 512;
static const int kSortWarps = kSortThreads / 32;
static const int kSortItems = 8;
static const int kSortTile = kSortThreads * kSortItems; /* 4096 keys */
static const int kMaxPasses = 8;
static const int kScanChunk = 256; /* tiles per chunk of the offset table */

struct PassPlan {
  int npass;
  int shift[kMaxPasses];
  int bits[kMaxPasses];
};

template <int kBits>
__global__ void __launch_bounds__(256)
    radix_tile_hist_kernel(const u64 *__restrict__ keys, i64 n, int shift,
                           int bits, u32 *__restrict__ thist) {
  const int kRadix = 1 << kBits;
  const int kItems = kSortTile / 256;
  __shared__ u32 s_h[kRadix];
  for (int i = threadIdx.x; i < kRadix; i += 256) s_h[i] = 0;
  const i64 tile_base = (i64)blockIdx.x * kSortTile;
  const bool full = tile_base + kSortTile <= n;
  const u32 dmask = (1u << bits) - 1u;
  u64 k[kItems];
#pragma unroll
  for (int u = 0; u < kItems; u++) {
    const i64 i = tile_base + u * 256 + threadIdx.x;
    k[u] = (full || i < n) ? keys[i] : 0;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int u = 0; u < kItems; u++) {
    const u32 d = (u32)(k[u] >> shift) & dmask;
    if (full) {
      /* rows of 32 equal digits -- the rule for the high digits of
         Morton-ordered input -- would serialise 32-fold on one bank */
      int uniform;
      __match_all_sync(0xffffffffu, d, &uniform);
      if (uniform) {
        if (lane == 0) atomicAdd(&s_h[d], 32u);
      } else {
        atomicAdd(&s_h[d], 1u);
      }
    } else if (tile_base + u * 256 + threadIdx.x < n) {
      atomicAdd(&s_h[d], 1u);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kRadix; i += 256) {
    thist[(size_t)blockIdx.x * kRadix + i] = s_h[i];
  }
}

/* in place: exclusive prefix of every digit column within a chunk of
   kScanChunk tiles; the chunk totals go to ctot[chunk][R] */
template <int kBits>
__global__ void __launch_bounds__(1 << kBits)
    radix_tile_scan_kernel(u32 *__restrict__ thist, i64 tiles,
                           u32 *__restrict__ ctot) {
  const int kRadix = 1 << kBits;
  const i64 row0 = (i64)blockIdx.x * kScanChunk;
  const int rows = (int)((tiles - row0 < kScanChunk) ? (tiles - row0) : kScanChunk);
  u32 *p = thist + (size_t)row0 * kRadix + threadIdx.x;
  u32 run = 0;
  int r = 0;
  for (; r + 8 <= rows; r += 8) {
    u32 v[8];
#pragma unroll
    for (int u = 0; u < 8; u++) v[u] = p[(size_t)(r + u) * kRadix];
#pragma unroll
    for (int u = 0; u < 8; u++) {
      p[(size_t)(r + u) * kRadix] = run;
      run += v[u];
    }
  }
  for (; r < rows; r++) {
    const u32 v = p[(size_t)r * kRadix];
    p[(size_t)r * kRadix] = run;
    run += v;
  }
  ctot[(size_t)blockIdx.x * kRadix + threadIdx.x] = run;
}

/* one CTA: exclusive prefix of the chunk totals per digit (in place) and the
   exclusive scan of the digit totals -> doff[R] */
template <int kBits>
__global__ void __launch_bounds__(1 << kBits)
    radix_chunk_scan_kernel(u32 *__restrict__ ctot, int chunks,
                            u32 *__restrict__ doff) {
  const int kRadix = 1 << kBits;
  __shared__ u32 s[kRadix];
  u32 *p = ctot + threadIdx.x;
  u32 run = 0;
  int c = 0;
  for (; c + 8 <= chunks; c += 8) {
    u32 v[8];
#pragma unroll
    for (int u = 0; u < 8; u++) v[u] = p[(size_t)(c + u) * kRadix];
#pragma unroll
    for (int u = 0; u < 8; u++) {
      p[(size_t)(c + u) * kRadix] = run;
      run += v[u];
    }
  }
  for (; c < chunks; c++) {
    const u32 v = p[(size_t)c * kRadix];
    p[(size_t)c * kRadix] = run;
    run += v;
  }
  s[threadIdx.x] = run;
  __syncthreads();
  for (int d = 1; d < kRadix; d <<= 1) {
    const u32 t = (threadIdx.x >= d) ? s[threadIdx.x - d] : 0;
    __syncthreads();
    s[threadIdx.x] += t;
    __syncthreads();
  }
  doff[threadIdx.x] = s[threadIdx.x] - run;
}

/* One pass of kBits (8 or 9) bits over one tile per CTA: 512 threads x 8 keys,
   <= 64 registers, two CTAs (32 warps) per SM (three CTAs at 40 registers
   measured 4 % slower: 1.19 vs 1.12 ms per pass on 200 M keys). */
template <bool kHasVals, int kBits>
__global__ void __launch_bounds__(kSortThreads, 2)
    radix_pass_kernel(const u64 *__restrict__ kin, u64 *__restrict__ kout,
                      const u32 *__restrict__ vin, u32 *__restrict__ vout,
                      i64 n, int shift, int bits,
                      const u32 *__restrict__ digit_off,  /* [R] */
                      const u32 *__restrict__ tile_excl,  /* [tiles][R] */
                      const u32 *__restrict__ chunk_excl /* [chunks][R] */) {
  const int kRadix = 1 << kBits;
  const int kHalves = kSortThreads / kRadix; /* threads per digit: 2 or 1 */
  const int kWarpsPerHalf = kSortWarps / kHalves;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  u64 *s_keys = reinterpret_cast<u64 *>(smem_raw);            /* tile keys */
  u32 *s_vals = reinterpret_cast<u32 *>(s_keys + kSortTile);  /* tile vals */
  u32 *s_whist = s_vals + (kHasVals ? kSortTile : 0);         /* [warps][R] */
  u32 *s_wmask = s_whist + kSortWarps * kRadix;               /* [warps][R] */
  u32 *s_dbase = s_wmask + kSortWarps * kRadix;               /* [halves][R] */
  u32 *s_part = s_dbase + kHalves * kRadix;                   /* [halves][R] */
  u32 *s_goff = s_part + kHalves * kRadix;                    /* [R] */
  __shared__ u32 s_wsum[kSortWarps];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const u32 tile = blockIdx.x;
  {
    /* s_whist and s_wmask are adjacent: zero both with 16-byte stores */
    uint4 *z = reinterpret_cast<uint4 *>(s_whist);
    const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
    for (int i = tid; i < 2 * kSortWarps * kRadix / 4; i += kSortThreads) z[i] = zero;
  }
  /* the tile's global digit offsets (histogram + scan kernels): fetched now,
     used after staging */
  const int my_digit = tid & (kRadix - 1), my_half = tid >> kBits;
  u32 t_off = 0;
  if (my_half == 0) {
    t_off = digit_off[my_digit] +
            chunk_excl[(size_t)(tile / kScanChunk) * kRadix + my_digit] +
            tile_excl[(size_t)tile * kRadix + my_digit];
  }
  __syncthreads();
  const i64 tile_base = (i64)tile * kSortTile;
  const i64 rem = n - tile_base;
  const int tile_n = rem < kSortTile ? (int)rem : kSortTile;
  const u32 dmask = (1u << bits) - 1u;

  /* (1) load warp-striped */
  u64 key[kSortItems];
  u32 val[kSortItems];
  unsigned short rank[kSortItems];
  const i64 wbase = tile_base + (i64)warp * (32 * kSortItems);
#pragma unroll
  for (int j = 0; j < kSortItems; j++) {
    const i64 i = wbase + j * 32 + lane;
    key[j] = (i < n) ? kin[i] : ~0ULL;
    if (kHasVals) val[j] = (i < n) ? vin[i] : 0u;
  }

  /* (2) rank within the warp, stable.  Rows of one warp look alike, so the
     warp probes rows 0 and 4: if lane 0's digit is shared by >= 4 lanes in
     either, the tile's rows are ranked with match.any, else with the masks. */
  u32 *my_hist = s_whist + warp * kRadix;
  bool many = false;
#pragma unroll
  for (int j = 0; j < kSortItems; j += 4) {
    const u32 d = (u32)(key[j] >> shift) & dmask;
    u32 b = __ballot_sync(0xffffffffu, d == __shfl_sync(0xffffffffu, d, 0));
    b &= b - 1u;
    b &= b - 1u;
    b &= b - 1u;
    many = many || (b != 0u);
  }
  if (!many) {
    /* every lane ORs its bit into the warp's word for its digit, reads the
       word back (= the lanes holding the same digit) and clears its own bit
       again.  All peers read the warp's running count of the digit, the
       highest peer lane advances it.  Rounds of one warp are ordered by the
       __syncwarp()s, which keeps the ranking stable. */
    u32 *my_mask = s_wmask + warp * kRadix;
    const u32 lane_bit = 1u << lane, lanes_lt = lane_bit - 1u;
#pragma unroll
    for (int j = 0; j < kSortItems; j++) {
      const u32 d = (u32)(key[j] >> shift) & dmask;
      atomicOr(&my_mask[d], lane_bit);
      __syncwarp();
      const u32 peers = my_mask[d];
      const u32 old = my_hist[d];
      __syncwarp();
      atomicAnd(&my_mask[d], ~lane_bit);
      const u32 below = (u32)__popc(peers & lanes_lt);
      if ((peers >> lane) == 1u) my_hist[d] = old + below + 1u; /* last peer */
      rank[j] = (unsigned short)(old + below);
    }
  } else {
#pragma unroll
    for (int j = 0; j < kSortItems; j++) {
      const u32 d = (u32)(key[j] >> shift) & dmask;
      const u32 peers = __match_any_sync(0xffffffffu, d);
      const int leader = __ffs(peers) - 1;
      u32 old = 0;
      if (lane == leader) old = atomicAdd(&my_hist[d], (u32)__popc(peers));
      old = __shfl_sync(0xffffffffu, old, leader);
      rank[j] = (unsigned short)(old + __popc(peers & ((1u << lane) - 1u)));
    }
  }
  __syncthreads();

  /* (3) warp bases per digit: thread (half, digit) turns the counts of its
     half of the warps into exclusive prefixes, in place */
  {
    u32 run = 0;
#pragma unroll
    for (int w = 0; w < kWarpsPerHalf; w++) {
      u32 *p = s_whist + (my_half * kWarpsPerHalf + w) * kRadix + my_digit;
      const u32 c = *p;
      *p = run;
      run += c;
    }
    s_part[my_half * kRadix + my_digit] = run;
  }
  __syncthreads();
  /* tile count of the digit, exclusive scan across digits (both halves do it
     redundantly, so no thread idles into the next barrier) */
  u32 count = 0, incl = 0;
#pragma unroll
  for (int h = 0; h < kHalves; h++) count += s_part[h * kRadix + my_digit];
  incl = count;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const u32 up = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += up;
  }
  if (lane == 31) s_wsum[warp] = incl;
  __syncthreads();
  {
    /* warps of one half hold the digits in order: sum the preceding warps of
       my half */
    const int w0 = my_half * kWarpsPerHalf;
    u32 woff = 0;
#pragma unroll
    for (int w = 0; w < kWarpsPerHalf; w++) {
      if (w0 + w < warp) woff += s_wsum[w0 + w];
    }
    const u32 dbase = woff + incl - count;
    /* first position of the digit's run for the warps of my half */
    u32 before = 0;
#pragma unroll
    for (int h = 0; h < kHalves; h++) {
      if (h < my_half) before += s_part[h * kRadix + my_digit];
    }
    s_dbase[my_half * kRadix + my_digit] = dbase + before;
    /* global index of staged position q holding digit d: goff[d] + q */
    if (my_half == 0) s_goff[my_digit] = t_off - dbase;
  }
  __syncthreads();

  /* (4) stage the tile in digit order */
  const u32 *my_dbase = s_dbase + (warp / kWarpsPerHalf) * kRadix;
#pragma unroll
  for (int j = 0; j < kSortItems; j++) {
    const u32 d = (u32)(key[j] >> shift) & dmask;
    const u32 q = my_dbase[d] + my_hist[d] + rank[j];
    s_keys[q] = key[j];
    if (kHasVals) s_vals[q] = val[j];
  }
  __syncthreads();

  /* (5) write: consecutive threads write consecutive positions of a run */
#pragma unroll
  for (int j = 0; j < kSortItems; j++) {
    const int q = j * kSortThreads + tid;
    if (q < tile_n) {
      const u64 k = s_keys[q];
      const u32 d = (u32)(k >> shift) & dmask;
      const u32 g = s_goff[d] + (u32)q;
      kout[g] = k;
      if (kHasVals) vout[g] = s_vals[q];
    }
  }
}

static size_t sort_smem_bytes(bool has_vals, int radix) {
  const int halves = kSortThreads / radix;
  size_t b = (size_t)kSortTile * sizeof(u64);
  if (has_vals) b += (size_t)kSortTile * sizeof(u32);
  b += 2 * (size_t)kSortWarps * radix * sizeof(u32); /* whist, wmask */
  b += 2 * (size_t)halves * radix * sizeof(u32);     /* dbase, part */
  b += (size_t)radix * sizeof(u32);                  /* goff */
  return b;
}

/* spread `total` key bits over the fewest passes of <= 9 bits, as evenly as
   possible, never below 8 bits unless the key is shorter; the wider digits go
   LAST (most significant): Morton-ordered input makes the high digits nearly
   constant within a tile, which is where a 512-bin pass is cheap (measured:
   a 9-bit pass over the lowest node-key bits took 2.65 ms against 1.55 ms for
   an 8-bit pass) */
static PassPlan make_plan(int bit_lo, int bit_hi) {
  PassPlan pl;
  const int total = bit_hi - bit_lo;
  int npass = (total + kMaxRadixBits - 1) / kMaxRadixBits;
  if (npass < 1) npass = 1;
  pl.npass = npass;
  int at = bit_lo;
  for (int p = 0; p < npass; p++) {
    const int left = bit_hi - at;
    const int b = left / (npass - p);
    pl.shift[p] = at;
    pl.bits[p] = b;
    at += b;
  }
  return pl;
}

template <int kBits>
static void launch_tile_offsets(Ctx &ctx, const u64 *keys, i64 n, i64 tiles,
                                int shift, int bits, u32 *thist, u32 *ctot,
                                u32 *doff) {
  cudaStream_t st = (cudaStream_t)ctx.stream;
  const int chunks = (int)((tiles + kScanChunk - 1) / kScanChunk);
  radix_tile_hist_kernel<kBits><<<(unsigned)tiles, 256, 0, st>>>(keys, n, shift,
                                                               bits, thist);
  radix_tile_scan_kernel<kBits><<<chunks, 1 << kBits, 0, st>>>(thist, tiles, ctot);
  radix_chunk_scan_kernel<kBits><<<1, 1 << kBits, 0, st>>>(ctot, chunks, doff);
}

template <bool kHasVals, int kBits>
static void launch_pass(Ctx &ctx, i64 tiles, DBuf<u64> &keys, DBuf<u64> &keys_alt,
                        DBuf<u32> &vals, DBuf<u32> &vals_alt, i64 n, int shift,
                        int bits, const u32 *doff, const u32 *thist,
                        const u32 *ctot) {
  const size_t smem = sort_smem_bytes(kHasVals, 1 << kBits);
  /* the attribute is per device: set it once for every device this process
     drives (bit mask of devices done, per kernel instantiation) */
  static unsigned long long attr_devices = 0;
  const unsigned long long dev_bit = 1ULL << (ctx.device & 63);
  if (!(attr_devices & dev_bit)) {
    cudaFuncSetAttribute(radix_pass_kernel<kHasVals, kBits>,
                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_devices |= dev_bit;
  }
  radix_pass_kernel<kHasVals, kBits>
      <<<(unsigned)tiles, kSortThreads, smem, (cudaStream_t)ctx.stream>>>(
          keys.get(), keys_alt.get(), kHasVals ? vals.get() : NULL,
          kHasVals ? vals_alt.get() : NULL, n, shift, bits, doff, thist, ctot);
}

void radix_sort(Ctx &ctx, DBuf<u64> &keys, DBuf<u64> &keys_alt,
                DBuf<u32> &vals, DBuf<u32> &vals_alt, i64 n, int bit_lo,
                int bit_hi, const char *tag) {
  if (n <= 1 || bit_hi <= bit_lo || !ctx.last_error.empty()) return;
  if (!keys.get() || !keys_alt.get()) return;
  if (n >= (1LL << 32)) {
    fprintf(stderr, "TMROctForest Error: radix sort of %lld keys exceeds the "
                    "32-bit offset range\n", (long long)n);
    ctx.last_error = "radix sort too large";
    return;
  }
  const bool has_vals = vals.get() != NULL;
  const i64 tiles = (n + kSortTile - 1) / kSortTile;
  const i64 chunks = (tiles + kScanChunk - 1) / kScanChunk;
  std::string hist_name = "radix_hist", pass_name = has_vals ? "radix_pass_pairs"
                                                             : "radix_pass_keys";
  if (tag) {
    hist_name += std::string("[") + tag + "]";
    pass_name += std::string("[") + tag + "]";
  }
  /* scratch: [512 digit offsets][chunks*512 chunk prefixes][tiles*512 tile
     prefixes], all u32 */
  u32 *scratch = static_cast<u32 *>(
      dev_alloc(ctx, (size_t)(1 + chunks + tiles) * kMaxRadix * sizeof(u32)));
  if (!scratch) return;
  u32 *doff = scratch;
  u32 *ctot = doff + kMaxRadix;
  u32 *thist = ctot + (size_t)chunks * kMaxRadix;

  int lo = bit_lo;
  while (lo < bit_hi) {
    int hi = bit_hi;
    if (hi - lo > kMaxPasses * kMaxRadixBits) hi = lo + kMaxPasses * kMaxRadixBits;
    const PassPlan plan = make_plan(lo, hi);
    for (int p = 0; p < plan.npass; p++) {
      const int bits = plan.bits[p], shift = plan.shift[p];
      prof_begin(ctx, hist_name.c_str());
      if (bits > 8) {
        launch_tile_offsets<9>(ctx, keys.get(), n, tiles, shift, bits, thist, ctot,
                               doff);
      } else {
        launch_tile_offsets<8>(ctx, keys.get(), n, tiles, shift, bits, thist, ctot,
                               doff);
      }
      prof_end(ctx);
      ctx.launch_count += 3;
      prof_begin(ctx, pass_name.c_str());
      if (has_vals) {
        if (bits > 8) {
          launch_pass<true, 9>(ctx, tiles, keys, keys_alt, vals, vals_alt, n, shift,
                               bits, doff, thist, ctot);
        } else {
          launch_pass<true, 8>(ctx, tiles, keys, keys_alt, vals, vals_alt, n, shift,
                               bits, doff, thist, ctot);
        }
      } else {
        if (bits > 8) {
          launch_pass<false, 9>(ctx, tiles, keys, keys_alt, vals, vals_alt, n,
                                shift, bits, doff, thist, ctot);
        } else {
          launch_pass<false, 8>(ctx, tiles, keys, keys_alt, vals, vals_alt, n,
                                shift, bits, doff, thist, ctot);
        }
      }
      prof_end(ctx);
      ctx.launch_count++;
      keys.swap(keys_alt);
      if (has_vals) vals.swap(vals_alt);
    }
    lo = hi;
  }
  dev_free(ctx, scratch);
}

}  // namespace tmrgpu
