/*
  prim_cuda.cu -- CUDA runtime plumbing and the hand-written radix sort.

  Radix sort = "onesweep": one up-front kernel histograms every pass's digit in
  a single read of the keys; then each 8-bit pass is ONE kernel that reads each
  key once and writes it once.  Inside a pass a CTA
    (1) takes a tile ticket (so look-back only ever waits on running CTAs),
    (2) ranks its 4096 keys with warp match/ballot multi-split (stable),
    (3) publishes its 256 digit counts and resolves its global digit offsets
        by decoupled look-back over the preceding tiles' descriptors,
    (4) stages the tile in shared memory in digit order and writes each digit
        run to HBM as contiguous, coalesced segments.
  HBM traffic per pass: 8 B read + 8 B written per key (+4+4 with a payload),
  plus 2 KB of descriptors per 4096-key tile.  This replaces
  qsort+comparator in reference src/TMROctant.cpp:357-399.
*/
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <time.h>

#include "prim_cuda.cuh"

namespace tmrgpu {

/* ------------------------------------------------------------------------ */
/* runtime                                                                  */
/* ------------------------------------------------------------------------ */
#define TMR_CUDA_OK(call)                                                     \
  do {                                                                        \
    cudaError_t e_ = (call);                                                  \
    if (e_ != cudaSuccess) {                                                  \
      fprintf(stderr, "TMROctForest Error: CUDA %s at %s:%d\n",               \
              cudaGetErrorString(e_), __FILE__, __LINE__);                    \
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

static void cache_release_all(DevCache &c) {
  for (std::multimap<size_t, void *>::iterator it = c.free_blocks.begin();
       it != c.free_blocks.end(); ++it) {
    cudaFree(it->second);
  }
  c.free_blocks.clear();
  c.cached_bytes = 0;
}

void dev_cache_destroy(Ctx &ctx) {
  std::map<Ctx *, DevCache>::iterator it = g_caches.find(&ctx);
  if (it == g_caches.end()) return;
  cudaStreamSynchronize((cudaStream_t)ctx.stream);
  cache_release_all(it->second);
  g_caches.erase(it);
}

size_t dev_cache_peak_bytes(Ctx &ctx) { return g_caches[&ctx].peak_bytes; }

void *dev_alloc(Ctx &ctx, size_t bytes) {
  DevCache &c = g_caches[&ctx];
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
  DevCache &c = g_caches[&ctx];
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
  DevCache &c = g_host_caches[&ctx];
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
  DevCache &c = g_host_caches[&ctx];
  std::map<void *, size_t>::iterator it = c.live.find(p);
  if (it == c.live.end()) return;
  c.free_blocks.insert(std::make_pair(it->second, p));
  c.live.erase(it);
}

void copy_h2d(Ctx &ctx, void *dst, const void *src, size_t bytes) {
  if (bytes == 0) return;
  TMR_CUDA_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice,
                              (cudaStream_t)ctx.stream));
  /* the source may be pageable/stack memory: make the copy complete before
     the caller can reuse it */
  TMR_CUDA_OK(cudaStreamSynchronize((cudaStream_t)ctx.stream));
}

void copy_d2h(Ctx &ctx, void *dst, const void *src, size_t bytes) {
  if (bytes == 0) return;
  TMR_CUDA_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost,
                              (cudaStream_t)ctx.stream));
  TMR_CUDA_OK(cudaStreamSynchronize((cudaStream_t)ctx.stream));
}

void copy_d2d(Ctx &ctx, void *dst, const void *src, size_t bytes) {
  if (bytes == 0) return;
  TMR_CUDA_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice,
                              (cudaStream_t)ctx.stream));
}

void dev_zero(Ctx &ctx, void *p, size_t bytes) {
  if (bytes == 0) return;
  TMR_CUDA_OK(cudaMemsetAsync(p, 0, bytes, (cudaStream_t)ctx.stream));
}

void dev_fill_ff(Ctx &ctx, void *p, size_t bytes) {
  if (bytes == 0) return;
  TMR_CUDA_OK(cudaMemsetAsync(p, 0xff, bytes, (cudaStream_t)ctx.stream));
}

void stream_sync(Ctx &ctx) {
  TMR_CUDA_OK(cudaStreamSynchronize((cudaStream_t)ctx.stream));
}

int check_errors(Ctx &ctx, const char *where) {
  cudaError_t e = cudaStreamSynchronize((cudaStream_t)ctx.stream);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) {
    ctx.last_error = std::string(where) + ": " + cudaGetErrorString(e);
    fprintf(stderr, "TMROctForest Error: CUDA failure in %s: %s\n", where,
            cudaGetErrorString(e));
    return 1;
  }
  if (!ctx.last_error.empty()) return 1;
  return 0;
}

void prof_begin(Ctx &ctx, const char *name) {
  if (!ctx.profile) return;
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  cudaEventRecord(a, (cudaStream_t)ctx.stream);
  ctx.ev_start.push_back(a);
  ctx.ev_stop.push_back(b);
  ctx.ev_name.push_back(name);
}

void prof_end(Ctx &ctx) {
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
static const int kMaxRadixBits = 9;           /* up to 512 digits per pass */
static const int kMaxRadix = 1 << kMaxRadixBits;
static const int kHistThreads = 256;
static const int kSortThreads = 512;
static const int kSortWarps = kSortThreads / 32;
static const int kSortItems = 8;
static const int kSortTile = kSortThreads * kSortItems; /* 4096 keys */
static const int kMaxPasses = 8;
static const int kDefaultRadixVariant = 15;
static const int kScanChunk = 256; /* tiles per chunk of the offset table */

struct PassPlan {
  int npass;
  int shift[kMaxPasses];
  int bits[kMaxPasses];
};

/* One read of the keys -> digit histograms of every pass.  Four keys per
   thread per round are in flight; each key adds 1 to its bin of every pass
   with a shared-memory RED (3.2 clk per warp instruction on B200 when the
   lanes hit different bins, measured).  Rows whose 32 digits are equal -- the
   rule for the high digits of Morton-ordered input -- would serialise 32-fold
   on one bank, so match.all (2 clk) detects them and one lane adds 32.
   kMatch = the previous scheme (match.any aggregation: ~2 clk per distinct
   digit, 61 clk for a random row), kept for A/B measurement. */
template <bool kMatch>
__global__ void __launch_bounds__(kHistThreads)
    radix_hist_kernel(const u64 *__restrict__ keys, i64 n, PassPlan plan,
                      u32 *__restrict__ ghist) {
  __shared__ u32 s_hist[kMaxPasses * kMaxRadix];
  for (int i = threadIdx.x; i < plan.npass * kMaxRadix; i += blockDim.x) {
    s_hist[i] = 0;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  if (kMatch) {
    const i64 stride = (i64)gridDim.x * blockDim.x;
    const i64 nround = ((n + 31) / 32) * 32; /* keep warps converged */
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < nround;
         i += stride) {
      const bool valid = i < n;
      const u64 k = valid ? keys[i] : 0;
      for (int p = 0; p < plan.npass; p++) {
        const u32 d = (u32)(k >> plan.shift[p]) & ((1u << plan.bits[p]) - 1u);
        const u32 peers = __match_any_sync(0xffffffffu, valid ? d : 0xffffffffu);
        if (valid && lane == (__ffs(peers) - 1)) {
          atomicAdd(&s_hist[p * kMaxRadix + d], __popc(peers));
        }
      }
    }
  } else {
    const int kUnroll = 4;
    const i64 chunk = (i64)blockDim.x * kUnroll;
    const i64 nchunks = (n + chunk - 1) / chunk;
    for (i64 c = blockIdx.x; c < nchunks; c += gridDim.x) {
      const i64 base = c * chunk + threadIdx.x;
      u64 k[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; u++) {
        const i64 i = base + (i64)u * blockDim.x;
        k[u] = (i < n) ? keys[i] : 0;
      }
      const bool full = (c + 1) * chunk <= n; /* uniform over the CTA */
      for (int p = 0; p < plan.npass; p++) {
        const int sh = plan.shift[p];
        const u32 m = (1u << plan.bits[p]) - 1u;
        u32 *h = s_hist + p * kMaxRadix;
#pragma unroll
        for (int u = 0; u < kUnroll; u++) {
          const u32 d = (u32)(k[u] >> sh) & m;
          if (full) {
            int uniform;
            __match_all_sync(0xffffffffu, d, &uniform);
            if (uniform) {
              if (lane == 0) atomicAdd(&h[d], 32u);
            } else {
              atomicAdd(&h[d], 1u);
            }
          } else if (base + (i64)u * blockDim.x < n) {
            atomicAdd(&h[d], 1u);
          }
        }
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < plan.npass * kMaxRadix; i += blockDim.x) {
    const u32 v = s_hist[i];
    if (v) atomicAdd(&ghist[i], v);
  }
}

/* exclusive scan of each pass's bins (one CTA of 512 threads per pass) */
__global__ void radix_scan_hist_kernel(u32 *ghist) {
  __shared__ u32 s[kMaxRadix];
  u32 *h = ghist + blockIdx.x * kMaxRadix;
  const u32 v = h[threadIdx.x];
  s[threadIdx.x] = v;
  __syncthreads();
  for (int d = 1; d < kMaxRadix; d <<= 1) {
    u32 t = (threadIdx.x >= d) ? s[threadIdx.x - d] : 0;
    __syncthreads();
    s[threadIdx.x] += t;
    __syncthreads();
  }
  h[threadIdx.x] = s[threadIdx.x] - v;
}

/* ---- table mode: per-tile digit offsets without a look-back chain --------
   Measured on B200 (ncu, 86 M-octant cycle): with decoupled look-back the
   pass kernel spent 59 % of its warp time waiting for predecessor tiles -- the
   prefix can only advance a few tiles per L2 round trip, and a 4096-key tile
   is processed faster than that.  Reading the keys once more per pass (8 B per
   key at streaming speed) to histogram every tile, and scanning the [tiles][R]
   table, removes every inter-CTA dependency from the pass kernel. */
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

/* One pass of kBits (8 or 9) bits.  512 threads x 8 keys: 16 warps per CTA and
   <= 64 registers so that two CTAs (32 warps) are resident per SM and the load
   / rank / look-back / scatter phases of different CTAs overlap.  With 9-bit
   digits every thread owns one digit in the descriptor phase. */
template <bool kHasVals, int kBits, int kVar>
__global__ void __launch_bounds__(kSortThreads, 2)
    radix_pass_kernel(const u64 *__restrict__ kin, u64 *__restrict__ kout,
                      const u32 *__restrict__ vin, u32 *__restrict__ vout,
                      i64 n, int shift, int bits,
                      const u32 *__restrict__ pass_offset, /* [kRadix] */
                      u32 *ticket, u64 *lookback /* [tiles][kRadix] */,
                      const u32 *__restrict__ tile_excl,  /* [tiles][kRadix] */
                      const u32 *__restrict__ chunk_excl /* [chunks][kRadix] */) {
  const int kRadix = 1 << kBits;
  extern __shared__ unsigned char smem_raw[];
  u64 *s_keys = reinterpret_cast<u64 *>(smem_raw);            /* tile keys */
  u32 *s_vals = reinterpret_cast<u32 *>(s_keys + kSortTile);  /* tile vals */
  u32 *s_whist = s_vals + (kHasVals ? kSortTile : 0);         /* [warps][R] */
  u32 *s_dbase = s_whist + kSortWarps * kRadix;               /* [R] */
  u64 *s_goff = reinterpret_cast<u64 *>(s_dbase + kRadix);    /* [R] */
  u32 *s_wmask = reinterpret_cast<u32 *>(s_goff + kRadix);    /* [warps][R] */
  __shared__ u32 s_tile;
  __shared__ u32 s_wsum[kMaxRadix / 32];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool kTable = (kVar & 8) != 0;
  if (!kTable && tid == 0) s_tile = atomicAdd(ticket, 1u);
  for (int i = tid; i < kSortWarps * kRadix; i += kSortThreads) {
    s_whist[i] = 0;
    if (kVar & 1) s_wmask[i] = 0;
  }
  __syncthreads();
  const u32 tile = kTable ? blockIdx.x : s_tile;
  /* table mode: the tile's global digit offsets were computed by the
     histogram + scan kernels; fetch them now, use them after staging */
  u32 t_off = 0;
  if (kTable && tid < kRadix) {
    t_off = pass_offset[tid] +
            chunk_excl[(size_t)(tile / kScanChunk) * kRadix + tid] +
            tile_excl[(size_t)tile * kRadix + tid];
  }
  const i64 tile_base = (i64)tile * kSortTile;
  const i64 rem = n - tile_base;
  const int tile_n = rem < kSortTile ? (int)rem : kSortTile;
  const u32 dmask = (1u << bits) - 1u;

  /* (2) load warp-striped and rank */
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
  u32 *my_hist = s_whist + warp * kRadix;
  bool use_masks = (kVar & 1) != 0;
  if (kVar & 4) {
    /* The two peer-discovery schemes have opposite worst cases (measured on
       B200): match.any costs ~2 clk per DISTINCT digit in the row (61 clk for
       random digits, 2 for equal ones); the shared-memory masks cost one bank
       cycle per lane sharing a word (6 clk for distinct digits, 64 for equal
       ones).  Rows of one warp look alike, so the warp probes rows 0 and 4:
       if lane 0's digit is shared by >= 4 lanes in either, it uses match.any
       for the tile. */
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
    use_masks = !many;
  }
  if (use_masks) {
    /* Peer discovery through shared memory instead of match.any: every lane
       ORs its bit into the warp's word for its digit, reads the word back (=
       the lanes holding the same digit) and clears its own bit again.  All
       peers read the warp's running count of the digit, the highest peer lane
       advances it.  Rounds of one warp are ordered by the __syncwarp()s, which
       keeps the ranking stable. */
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
      /* one shared-memory atomic per distinct digit of the round; rounds of
         the same warp execute in program order, which keeps the ranking
         stable */
      if (lane == leader) old = atomicAdd(&my_hist[d], (u32)__popc(peers));
      old = __shfl_sync(0xffffffffu, old, leader);
      rank[j] = (unsigned short)(old + __popc(peers & ((1u << lane) - 1u)));
    }
  }
  __syncthreads();

  /* (3) per digit (thread t < kRadix owns digit t): warp bases, tile count,
     descriptor, look-back */
  u32 count = 0, incl = 0;
  u64 *my_desc = lookback + (size_t)tile * kRadix + tid;
  if (tid < kRadix) {
#pragma unroll
    for (int w = 0; w < kSortWarps; w++) {
      const u32 c = s_whist[w * kRadix + tid];
      s_whist[w * kRadix + tid] = count;
      count += c;
    }
    if (!kTable) {
      st_relaxed_u64(my_desc, (tile == 0 ? kStatusPrefix : kStatusAgg) | (u64)count);
    }
    incl = count;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      u32 up = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += up;
    }
    if (lane == 31) s_wsum[warp] = incl;
  }
  __syncthreads();
  u32 dbase = 0;
  if (tid < kRadix) {
    u32 woff = 0;
#pragma unroll
    for (int w = 0; w < kRadix / 32; w++) {
      if (w < warp) woff += s_wsum[w];
    }
    dbase = woff + incl - count;
    s_dbase[tid] = dbase;
  }
  __syncthreads();

  /* (4) stage the tile in digit order first: it only needs the local digit
     bases, and it gives the preceding tiles time to publish before ... */
#pragma unroll
  for (int j = 0; j < kSortItems; j++) {
    const u32 d = (u32)(key[j] >> shift) & dmask;
    const u32 q = s_dbase[d] + my_hist[d] + rank[j];
    s_keys[q] = key[j];
    if (kHasVals) s_vals[q] = val[j];
  }
  /* ... (5) the decoupled look-back resolves the global offset of each digit */
  if (tid < kRadix) {
    u64 excl = 0;
    if (kTable) {
      excl = t_off;
    } else if (tile > 0) {
      i64 p = (i64)tile - 1;
      if (kVar & 2) {
        /* kLook descriptors of consecutive predecessors are fetched at once:
           a walk of m tiles costs ~m/kLook L2 round trips instead of m.  Rows
           before tile 0 read as "inclusive prefix 0". */
        const int kLook = 8;
        bool done = false;
        while (!done) {
          u64 v[kLook];
#pragma unroll
          for (int u = 0; u < kLook; u++) {
            v[u] = (p - u >= 0)
                       ? ld_relaxed_u64(lookback + (size_t)(p - u) * kRadix + tid)
                       : kStatusPrefix;
          }
          int used = 0;
#pragma unroll
          for (int u = 0; u < kLook; u++) {
            const u64 st = v[u] & kStatusMask;
            if (!done && used == u && st != 0) {
              excl += v[u] & ~kStatusMask;
              used = u + 1;
              if (st == kStatusPrefix) done = true;
            }
          }
          p -= used;
        }
      } else {
        while (true) {
          const u64 v = ld_relaxed_u64(lookback + (size_t)p * kRadix + tid);
          const u64 st = v & kStatusMask;
          if (st == 0) continue;
          excl += v & ~kStatusMask;
          if (st == kStatusPrefix) break;
          p--;
        }
      }
      st_relaxed_u64(my_desc, kStatusPrefix | (excl + (u64)count));
    }
    /* global index of staged position q holding digit d: goff[d] + q */
    s_goff[tid] = (kTable ? 0ULL : (u64)pass_offset[tid]) + excl - (u64)dbase;
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < kSortItems; j++) {
    const int q = j * kSortThreads + tid;
    if (q < tile_n) {
      const u64 k = s_keys[q];
      const u32 d = (u32)(k >> shift) & dmask;
      const u64 g = s_goff[d] + (u64)q;
      kout[g] = k;
      if (kHasVals) vout[g] = s_vals[q];
    }
  }
}

static size_t sort_smem_bytes(bool has_vals, int radix, int var) {
  size_t b = (size_t)kSortTile * sizeof(u64);
  if (has_vals) b += (size_t)kSortTile * sizeof(u32);
  b += (size_t)kSortWarps * radix * sizeof(u32);
  b += (size_t)radix * sizeof(u32);
  b += (size_t)radix * sizeof(u64);
  if (var & 1) b += (size_t)kSortWarps * radix * sizeof(u32);
  return b;
}

/* spread `total` key bits over the fewest passes of <= 9 bits, as evenly as
   possible, never below 8 bits unless the key is shorter */
static PassPlan make_plan(int bit_lo, int bit_hi) {
  PassPlan pl;
  const int total = bit_hi - bit_lo;
  int npass = (total + kMaxRadixBits - 1) / kMaxRadixBits;
  if (npass < 1) npass = 1;
  pl.npass = npass;
  int at = bit_lo;
  for (int p = 0; p < npass; p++) {
    const int left = bit_hi - at;
    const int b = (left + (npass - p) - 1) / (npass - p);
    pl.shift[p] = at;
    pl.bits[p] = b;
    at += b;
  }
  return pl;
}

template <bool kHasVals, int kBits, int kVar>
static void launch_pass_v(Ctx &ctx, i64 tiles, DBuf<u64> &keys, DBuf<u64> &keys_alt,
                        DBuf<u32> &vals, DBuf<u32> &vals_alt, i64 n, int shift,
                        int bits, const u32 *offs, u32 *ticket, u64 *lookback,
                        const u32 *tile_excl, const u32 *chunk_excl) {
  const size_t smem = sort_smem_bytes(kHasVals, 1 << kBits, kVar);
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(radix_pass_kernel<kHasVals, kBits, kVar>,
                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_set = true;
  }
  radix_pass_kernel<kHasVals, kBits, kVar>
      <<<(unsigned)tiles, kSortThreads, smem, (cudaStream_t)ctx.stream>>>(
          keys.get(), keys_alt.get(), kHasVals ? vals.get() : NULL,
          kHasVals ? vals_alt.get() : NULL, n, shift, bits, offs, ticket,
          lookback, tile_excl, chunk_excl);
}

/* TMR_RADIX_VARIANT (measurement switch): bit 0 = peers through shared-memory
   masks instead of match.any, bit 1 = multi-descriptor look-back, bit 2 =
   per-warp choice between the masks and match.any, bit 3 = table mode (per-tile
   offsets from histogram + scan kernels, no look-back) */
static int radix_variant() {
  static int v = -1;
  if (v < 0) {
    const char *e = getenv("TMR_RADIX_VARIANT");
    v = e ? atoi(e) & 15 : kDefaultRadixVariant;
  }
  return v;
}

template <bool kHasVals, int kBits>
static void launch_pass(Ctx &ctx, i64 tiles, DBuf<u64> &keys, DBuf<u64> &keys_alt,
                        DBuf<u32> &vals, DBuf<u32> &vals_alt, i64 n, int shift,
                        int bits, const u32 *offs, u32 *ticket, u64 *lookback,
                        const u32 *tile_excl, const u32 *chunk_excl) {
  switch (radix_variant()) {
#define TMR_PASS_CASE(V)                                                       \
  case V:                                                                      \
    launch_pass_v<kHasVals, kBits, V>(ctx, tiles, keys, keys_alt, vals,        \
                                      vals_alt, n, shift, bits, offs, ticket,  \
                                      lookback, tile_excl, chunk_excl);        \
    break;
    TMR_PASS_CASE(0)
    TMR_PASS_CASE(2)
    TMR_PASS_CASE(3)
    TMR_PASS_CASE(7)
    TMR_PASS_CASE(15)
#undef TMR_PASS_CASE
    default:
      fprintf(stderr, "TMROctForest Error: unknown TMR_RADIX_VARIANT\n");
      break;
  }
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

template <bool kHasVals>
static void launch_pass_bits(Ctx &ctx, i64 tiles, DBuf<u64> &keys,
                             DBuf<u64> &keys_alt, DBuf<u32> &vals,
                             DBuf<u32> &vals_alt, i64 n, int shift, int bits,
                             const u32 *offs, u32 *ticket, u64 *lookback,
                             const u32 *tile_excl, const u32 *chunk_excl) {
  if (bits > 8) {
    launch_pass<kHasVals, 9>(ctx, tiles, keys, keys_alt, vals, vals_alt, n, shift,
                             bits, offs, ticket, lookback, tile_excl, chunk_excl);
  } else {
    launch_pass<kHasVals, 8>(ctx, tiles, keys, keys_alt, vals, vals_alt, n, shift,
                             bits, offs, ticket, lookback, tile_excl, chunk_excl);
  }
}

void radix_sort(Ctx &ctx, DBuf<u64> &keys, DBuf<u64> &keys_alt,
                DBuf<u32> &vals, DBuf<u32> &vals_alt, i64 n, int bit_lo,
                int bit_hi, const char *tag) {
  if (n <= 1 || bit_hi <= bit_lo) return;
  if (n >= (1LL << 32)) {
    fprintf(stderr, "TMROctForest Error: radix sort of %lld keys exceeds the "
                    "32-bit offset range\n", (long long)n);
    ctx.last_error = "radix sort too large";
    return;
  }
  const bool has_vals = vals.get() != NULL;
  const bool table = (radix_variant() & 8) != 0;
  cudaStream_t st = (cudaStream_t)ctx.stream;
  const i64 tiles = (n + kSortTile - 1) / kSortTile;
  const i64 chunks = (tiles + kScanChunk - 1) / kScanChunk;
  std::string hist_name = "radix_hist", pass_name = has_vals ? "radix_pass_pairs"
                                                             : "radix_pass_keys";
  if (tag) {
    hist_name += std::string("[") + tag + "]";
    pass_name += std::string("[") + tag + "]";
  }
  /* scratch, look-back mode: [kMaxPasses*512 u32 hist][tickets][tiles*512 u64
     descriptors]; table mode: [512 u32 digit offsets][chunks*512 u32]
     [tiles*512 u32] */
  const size_t hist_bytes = (size_t)kMaxPasses * kMaxRadix * sizeof(u32);
  const size_t ticket_bytes = 64;
  const size_t look_bytes =
      table ? (size_t)(tiles + chunks) * kMaxRadix * sizeof(u32)
            : (size_t)tiles * kMaxRadix * sizeof(u64);
  unsigned char *scratch = static_cast<unsigned char *>(
      dev_alloc(ctx, hist_bytes + ticket_bytes + look_bytes));
  if (!scratch) return;
  u32 *ghist = reinterpret_cast<u32 *>(scratch);
  u32 *tickets = reinterpret_cast<u32 *>(scratch + hist_bytes);
  u64 *lookback = reinterpret_cast<u64 *>(scratch + hist_bytes + ticket_bytes);
  u32 *ctot = reinterpret_cast<u32 *>(lookback);
  u32 *thist = ctot + (size_t)chunks * kMaxRadix;

  int lo = bit_lo;
  while (lo < bit_hi) {
    /* at most kMaxPasses passes are planned per sweep */
    int hi = bit_hi;
    if (hi - lo > kMaxPasses * kMaxRadixBits) hi = lo + kMaxPasses * kMaxRadixBits;
    const PassPlan plan = make_plan(lo, hi);
    if (!table) {
      dev_zero(ctx, scratch, hist_bytes + ticket_bytes);
      prof_begin(ctx, hist_name.c_str());
      if (radix_variant() & 1) {
        radix_hist_kernel<false><<<grid_for(ctx, n, kHistThreads * 4, 8),
                                   kHistThreads, 0, st>>>(keys.get(), n, plan,
                                                          ghist);
      } else {
        radix_hist_kernel<true><<<grid_for(ctx, n, kHistThreads * 4, 8),
                                  kHistThreads, 0, st>>>(keys.get(), n, plan,
                                                         ghist);
      }
      radix_scan_hist_kernel<<<plan.npass, kMaxRadix, 0, st>>>(ghist);
      prof_end(ctx);
      ctx.launch_count += 2;
    }
    for (int p = 0; p < plan.npass; p++) {
      const int bits = plan.bits[p];
      const int radix = bits > 8 ? 512 : 256;
      const u32 *offs = ghist + p * kMaxRadix;
      if (table) {
        prof_begin(ctx, hist_name.c_str());
        if (bits > 8) {
          launch_tile_offsets<9>(ctx, keys.get(), n, tiles, plan.shift[p], bits,
                                 thist, ctot, ghist);
        } else {
          launch_tile_offsets<8>(ctx, keys.get(), n, tiles, plan.shift[p], bits,
                                 thist, ctot, ghist);
        }
        prof_end(ctx);
        ctx.launch_count += 3;
        offs = ghist;
      } else {
        dev_zero(ctx, lookback, (size_t)tiles * radix * sizeof(u64));
      }
      prof_begin(ctx, pass_name.c_str());
      if (has_vals) {
        launch_pass_bits<true>(ctx, tiles, keys, keys_alt, vals, vals_alt, n,
                               plan.shift[p], bits, offs, tickets + p, lookback,
                               thist, ctot);
      } else {
        launch_pass_bits<false>(ctx, tiles, keys, keys_alt, vals, vals_alt, n,
                                plan.shift[p], bits, offs, tickets + p, lookback,
                                thist, ctot);
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
