/*
  prim_cuda.cuh -- sm_100a implementations of the two templated primitives
  every forest operation is built from:

    launch(ctx, n, f, name)          f(i) for i in [0,n): grid-stride, grid sized
                                     as a multiple of the SM count
    scan_counts(ctx, n, f, out, nm)  out[i] = sum_{j<i} f(j)   (exclusive),
                                     returns the total; ONE pass over the input
                                     (chained scan with decoupled look-back),
                                     so compaction/expansion kernels read their
                                     input once and write offsets once.

  Both take a plain functor (a TMR_HD struct), so kernel bodies stay
  host-testable.
*/
#ifndef TMRGPU_PRIM_CUDA_CUH
#define TMRGPU_PRIM_CUDA_CUH

#include <cuda_runtime.h>

#include "prim.h"

namespace tmrgpu {

static const int kLaunchThreads = 256;

/* resident CTAs per SM a kernel body asks the compiler to make room for
   (register cap = 65536 / (256 * value)); specialised next to the few bodies
   whose occupancy is worth a tighter register budget */
template <class F>
struct LaunchMinBlocks {
  static const int value = 1;
};

template <class F>
__global__ void __launch_bounds__(kLaunchThreads)
    launch_kernel(F f, i64 n) {
  const i64 stride = (i64)gridDim.x * blockDim.x;
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    f(i);
  }
}
/* the same with a register budget for kMin resident CTAs per SM (a minimum of
   1 is NOT the same as none: ptxas then spends registers freely -- measured:
   HangingFn 40 -> 54, MapClosureFn 32 -> 72 registers) */
template <class F, int kMin>
__global__ void __launch_bounds__(kLaunchThreads, kMin)
    launch_kernel_occ(F f, i64 n) {
  const i64 stride = (i64)gridDim.x * blockDim.x;
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    f(i);
  }
}

inline int grid_for(const Ctx &ctx, i64 n, int threads, int max_waves) {
  i64 blocks = (n + threads - 1) / threads;
  const i64 cap = (i64)ctx.num_sms * max_waves;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

/* false once an operation has recorded a failure (allocation, size limit,
   NCCL): nothing more is launched -- a kernel must never run on the NULL
   buffer of a failed allocation (an illegal address is a sticky error that
   takes the whole process's CUDA context with it) -- and the operation's
   closing check_errors reports it */
inline bool ctx_ok(const Ctx &ctx) { return ctx.last_error.empty(); }

template <class F>
void launch(Ctx &ctx, i64 n, F f, const char *name) {
  if (n <= 0 || !ctx_ok(ctx)) return;
  /* up to 8 resident CTAs of 256 threads per SM; 4 waves of grid-stride */
  const int grid = grid_for(ctx, n, kLaunchThreads, 8 * 4);
  prof_begin(ctx, name);
  if constexpr (LaunchMinBlocks<F>::value > 1) {
    launch_kernel_occ<F, LaunchMinBlocks<F>::value>
        <<<grid, kLaunchThreads, 0, (cudaStream_t)ctx.stream>>>(f, n);
  } else {
    launch_kernel<F><<<grid, kLaunchThreads, 0, (cudaStream_t)ctx.stream>>>(f, n);
  }
  prof_end(ctx);
  ctx.launch_count++;
}

/* ---- warp-cooperative launch ------------------------------------------------
   f.warp(base, lane, n) is called by all 32 lanes of a warp together for the
   32 consecutive items base .. base + 31 (items >= n are the lanes' own
   business): bodies whose work sits in a few of many items (sparse bitmaps)
   spread one item's inner loop over the lanes instead of leaving 28 of them
   idle.  The same functor's operator()(i) is the per-item form (used by the
   test-only emulation). */
template <class F>
__global__ void __launch_bounds__(kLaunchThreads, 4) /* up to 64 registers: no spills */
    warp_kernel(F f, i64 n) {
  const int lane = threadIdx.x & 31;
  const i64 warp0 = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const i64 nwarps = ((i64)gridDim.x * blockDim.x) >> 5;
  for (i64 base = warp0 * 32; base < n; base += nwarps * 32) f.warp(base, lane, n);
}

template <class F>
void launch_warp(Ctx &ctx, i64 n, F f, const char *name) {
  if (n <= 0 || !ctx_ok(ctx)) return;
  const int grid = grid_for(ctx, n, kLaunchThreads, 8 * 4);
  prof_begin(ctx, name);
  warp_kernel<F><<<grid, kLaunchThreads, 0, (cudaStream_t)ctx.stream>>>(f, n);
  prof_end(ctx);
  ctx.launch_count++;
}

/* ---- expand: item i appends count(i) 64-bit outputs at off[i] -------------
   off is the exclusive scan of the counts (consecutive items own consecutive
   output ranges), total = off[n].  A thread writing its few outputs straight
   to HBM touches a different 32-byte sector with every store of a warp
   instruction (measured: the node-candidate kernel ran at the L2 sector-write
   rate, 4.5 ms for 2.7 GB); here the CTA's 256 items stage their outputs in
   shared memory and the CTA writes the contiguous range with coalesced
   stores.  F is called as f(i, sink) and appends with sink(value); kMaxPer
   bounds count(i). */
struct SmemSink {
  u64 *p;
  __device__ __forceinline__ void operator()(u64 v) { *p++ = v; }
};

template <class F, int kMaxPer>
__global__ void __launch_bounds__(kLaunchThreads)
    expand_kernel(F f, i64 n, const u32 *__restrict__ off, u64 total,
                  u64 *__restrict__ out) {
  __shared__ u64 s_out[kLaunchThreads * kMaxPer];
  const i64 nblk = (n + kLaunchThreads - 1) / kLaunchThreads;
  for (i64 b = blockIdx.x; b < nblk; b += gridDim.x) {
    const i64 i0 = b * kLaunchThreads, i1 = i0 + kLaunchThreads;
    const u64 base = off[i0];
    const u64 end = (i1 < n) ? (u64)off[i1] : total;
    const i64 i = i0 + threadIdx.x;
    if (i < n) {
      SmemSink sink = {s_out + (off[i] - base)};
      f(i, sink);
    }
    __syncthreads();
    const int cnt = (int)(end - base);
    for (int q = threadIdx.x; q < cnt; q += kLaunchThreads) out[base + q] = s_out[q];
    __syncthreads();
  }
}

template <int kMaxPer, class F>
void expand_u64(Ctx &ctx, i64 n, const u32 *off, u64 total, F f, u64 *out,
                const char *name) {
  if (n <= 0 || !ctx_ok(ctx)) return;
  const int grid = grid_for(ctx, n, kLaunchThreads, 8 * 4);
  prof_begin(ctx, name);
  expand_kernel<F, kMaxPer><<<grid, kLaunchThreads, 0, (cudaStream_t)ctx.stream>>>(
      f, n, off, total, out);
  prof_end(ctx);
  ctx.launch_count++;
}


/* ---- block kernel with a shared task list ------------------------------------
   256 consecutive items per CTA, three phases separated by barriers:
     f.collect(i, t, shared)   every item classifies itself and appends the
                               work it needs to task lists in shared memory
     f.process(i0, t, shared)  all 256 threads drain the lists densely (a task
                               is not tied to the thread that queued it, so a
                               warp never idles behind one expensive item)
     f.finish(i, t, shared)    every item assembles its outputs
   F::Shared::reset() clears the list counters.  Used by the node-slot
   construction (ops_nodes_slots.h), where sibling elements share the point
   locations of their family's 27 nodes. */
template <class F>
__global__ void __launch_bounds__(kLaunchThreads, 6) /* 6 CTAs per SM: <= 40 registers */
    block3_kernel(F f, i64 n) {
  __shared__ typename F::Shared sh;
  const i64 nblk = (n + kLaunchThreads - 1) / kLaunchThreads;
  for (i64 b = blockIdx.x; b < nblk; b += gridDim.x) {
    const i64 i0 = b * kLaunchThreads;
    const i64 i = i0 + threadIdx.x;
    if (threadIdx.x == 0) sh.reset();
    __syncthreads();
    if (i < n) f.collect(i, (int)threadIdx.x, sh);
    __syncthreads();
    f.process(i0, (int)threadIdx.x, sh);
    __syncthreads();
    if (i < n) f.finish(i, (int)threadIdx.x, sh);
    __syncthreads();
  }
}

template <class F>
void launch_block3(Ctx &ctx, i64 n, F f, const char *name) {
  if (n <= 0 || !ctx_ok(ctx)) return;
  const int grid = grid_for(ctx, n, kLaunchThreads, 8 * 4);
  prof_begin(ctx, name);
  block3_kernel<F><<<grid, kLaunchThreads, 0, (cudaStream_t)ctx.stream>>>(f, n);
  prof_end(ctx);
  ctx.launch_count++;
}

/* ---- chained scan --------------------------------------------------------
   tile = 256 threads x 8 items, warp-striped: warp w owns items [256 w,
   256 w + 256) of the tile and lane l takes items l, l + 32, .. of them, so
   every load and store of f and g is a coalesced 128-byte row (a blocked
   layout -- 8 consecutive items per thread -- made each warp instruction
   touch 32 sectors: measured 1.9 ms for a 24-byte-per-item scan of 86 M items
   that moves 2 GB).  Prefixes in item order: one warp scan per row plus the
   row totals carried along.  Tile descriptors are one 64-bit word:
   [2-bit status | 62-bit value], written with a single store so status and
   value can never be observed torn. */
static const int kScanThreads = 256;
static const int kScanItems = 8;
static const int kScanTile = kScanThreads * kScanItems;
static const u64 kStatusAgg = 1ULL << 62;
static const u64 kStatusPrefix = 2ULL << 62;
static const u64 kStatusMask = 3ULL << 62;

__device__ __forceinline__ void st_relaxed_u64(u64 *p, u64 v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v)
               : "memory");
}
__device__ __forceinline__ u64 ld_relaxed_u64(const u64 *p) {
  u64 v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p)
               : "memory");
  return v;
}

/* G is called as g(i, exclusive_prefix_of_i) for every i < n.  The count type
   is whatever F returns: u32, or u64 when two counters ride in one word
   (ops_nodes_slots.h packs node and dependent-node counts as lo | hi << 31;
   a tile descriptor holds 62 value bits). */
template <class F, class G>
__global__ void __launch_bounds__(kScanThreads)
    scan_apply_kernel(F f, G g, i64 n, u64 *tile_state, u32 *ticket,
                      u64 *total) {
  typedef decltype(f((i64)0)) T;
  __shared__ u32 s_tile;
  __shared__ u64 s_warp_sum[kScanThreads / 32];
  __shared__ u64 s_tile_prefix;
  if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
  __syncthreads();
  const u32 tile = s_tile;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const i64 base = (i64)tile * kScanTile + (i64)warp * (32 * kScanItems) + lane;

  /* c[k]: count of item base + 32 k; x[k]: its exclusive prefix within the
     warp's 256 items */
  T c[kScanItems];
  u64 x[kScanItems];
  u64 carry = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; k++) {
    const i64 i = base + 32 * k;
    c[k] = (i < n) ? f(i) : (T)0;
  }
#pragma unroll
  for (int k = 0; k < kScanItems; k++) {
    T incl = c[k];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const T up = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += up;
    }
    x[k] = carry + (u64)(incl - c[k]);
    carry += (u64)__shfl_sync(0xffffffffu, incl, 31);
  }
  if (lane == 31) s_warp_sum[warp] = carry;
  __syncthreads();
  u64 warp_off = 0, tile_sum = 0;
#pragma unroll
  for (int w = 0; w < kScanThreads / 32; w++) {
    const u64 s = s_warp_sum[w];
    if (w < warp) warp_off += s;
    tile_sum += s;
  }
  /* publish the aggregate, then warp 0 looks back for the exclusive tile
     prefix, 32 predecessors per poll (lane l reads tile-1-l): a walk over m
     unfinished tiles costs ~m/32 L2 round trips */
  if (warp == 0) {
    u64 excl = 0;
    if (tile == 0) {
      if (lane == 0) st_relaxed_u64(&tile_state[0], kStatusPrefix | tile_sum);
    } else {
      if (lane == 0) st_relaxed_u64(&tile_state[tile], kStatusAgg | tile_sum);
      i64 p = (i64)tile - 1;
      while (true) {
        const i64 q = p - lane;
        /* positions before tile 0 read as "inclusive prefix 0" */
        const u64 v = (q >= 0) ? ld_relaxed_u64(&tile_state[q]) : kStatusPrefix;
        const u64 st = v & kStatusMask;
        const unsigned ready = __ballot_sync(0xffffffffu, st != 0);
        const unsigned pref = __ballot_sync(0xffffffffu, st == kStatusPrefix);
        const int nready = (~ready == 0u) ? 32 : (__ffs(~ready) - 1);
        const int fp = pref ? (__ffs(pref) - 1) : 32;
        const int take = (fp < nready) ? fp + 1 : nready;
        u64 c = (lane < take) ? (v & ~kStatusMask) : 0;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
        excl += c;
        if (fp < nready) break;
        p -= take;
      }
      if (lane == 0) {
        st_relaxed_u64(&tile_state[tile], kStatusPrefix | (excl + tile_sum));
      }
    }
    if (lane == 0) {
      s_tile_prefix = excl;
      if ((i64)(tile + 1) * kScanTile >= n) *total = excl + tile_sum;
    }
  }
  __syncthreads();
  const u64 run = s_tile_prefix + warp_off;
#pragma unroll
  for (int k = 0; k < kScanItems; k++) {
    const i64 i = base + 32 * k;
    if (i < n) g(i, (T)(run + x[k]));
  }
}

/* f(i) -> count; g(i, exclusive prefix) consumes it in the same kernel (no
   offset array round trip through HBM).  Returns the total. */
template <class F, class G>
u64 scan_apply(Ctx &ctx, i64 n, F f, G g, const char *name) {
  if (n <= 0 || !ctx_ok(ctx)) return 0;
  const i64 tiles = (n + kScanTile - 1) / kScanTile;
  const size_t bytes = (size_t)(tiles + 2) * sizeof(u64);
  u64 *scratch = static_cast<u64 *>(dev_alloc(ctx, bytes));
  if (!scratch) return 0;
  dev_zero(ctx, scratch, bytes);
  u64 *tile_state = scratch;
  u32 *ticket = reinterpret_cast<u32 *>(scratch + tiles);
  /* the last tile stores the total straight into page-locked host memory */
  unsigned long long *slot = mailbox_slot(ctx);
  u64 *total = slot ? reinterpret_cast<u64 *>(slot) : scratch + tiles + 1;
  prof_begin(ctx, name);
  scan_apply_kernel<F, G><<<(unsigned)tiles, kScanThreads, 0,
                            (cudaStream_t)ctx.stream>>>(f, g, n, tile_state,
                                                        ticket, total);
  prof_end(ctx);
  ctx.launch_count++;
  u64 h_total = 0;
  if (slot) {
    stream_sync(ctx);
    h_total = (u64)*static_cast<volatile unsigned long long *>(slot);
  } else {
    copy_d2h(ctx, &h_total, total, sizeof(u64));
  }
  dev_free(ctx, scratch);
  return h_total;
}

struct StoreOffsetFn {
  u32 *out;
  __device__ __forceinline__ void operator()(i64 i, u32 off) const {
    out[i] = off;
  }
};

template <class F>
u64 scan_counts(Ctx &ctx, i64 n, F f, u32 *out, const char *name) {
  StoreOffsetFn g = {out};
  const u64 total = scan_apply(ctx, n, f, g, name);
  if (total >> 32) {
    /* the offsets are 32-bit: a larger total has wrapped them.  Recorded, so
       that the operation's closing check_errors fails instead of handing out
       corrupt arrays */
    fprintf(stderr, "TMROctForest Error: %s: %llu outputs exceed the 32-bit offsets "
                    "of the CUDA path\n", name, (unsigned long long)total);
    ctx.last_error = "scan total exceeds 32 bits";
  }
  return total;
}

}  // namespace tmrgpu

#endif
