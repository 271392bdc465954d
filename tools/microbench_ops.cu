/* Warp-primitive throughput on sm_100a (cycles per warp instruction per SM):
   the numbers DESIGN.md section 4 uses to bound the radix ranking step.
   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/microbench_ops tools/microbench_ops.cu
   run  : tools/_build/microbench_ops  (prints one line per op) */
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

typedef unsigned int u32;

__device__ __forceinline__ u32 rng(u32 &s) {
  s ^= s << 13; s ^= s >> 17; s ^= s << 5;
  return s;
}

enum Op { MATCH_RANDOM8, MATCH_RANDOM4, MATCH_RANDOM2, MATCH_UNIFORM, BALLOT, SHFL, ATOMS_RET,
          ATOMS_NORET, ATOMS_LEADER, POPC_FFS, NOPS };

template <int OP>
__global__ void __launch_bounds__(512) k(u32 *out, int iters, u32 seed) {
  __shared__ u32 s_hist[16 * 256];
  for (int i = threadIdx.x; i < 16 * 256; i += blockDim.x) s_hist[i] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  u32 s = seed ^ (blockIdx.x * 1315423911u + threadIdx.x * 2654435761u) | 1u;
  u32 acc = 0;
  u32 v[4];
  for (int i = 0; i < 4; i++) v[i] = rng(s);
  u32 *h = s_hist + warp * 256;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 4; u++) {
      u32 x = v[u];
      if (OP == MATCH_RANDOM8) acc += __match_any_sync(0xffffffffu, x & 255u);
      if (OP == MATCH_RANDOM4) acc += __match_any_sync(0xffffffffu, x & 15u);
      if (OP == MATCH_RANDOM2) acc += __match_any_sync(0xffffffffu, x & 3u);
      if (OP == MATCH_UNIFORM) acc += __match_any_sync(0xffffffffu, (x & 0u) + it);
      if (OP == BALLOT) acc += __ballot_sync(0xffffffffu, x & 1u);
      if (OP == SHFL) acc += __shfl_sync(0xffffffffu, x, (x >> 8) & 31);
      if (OP == ATOMS_RET) acc += atomicAdd(&h[x & 255u], 1u);
      if (OP == ATOMS_NORET) atomicAdd(&h[x & 255u], 1u);
      if (OP == ATOMS_LEADER) {
        u32 old = 0;
        if (lane == (int)((x >> 8) & 31)) old = atomicAdd(&h[x & 255u], 1u);
        acc += old;
      }
      if (OP == POPC_FFS) acc += __popc(x) + __ffs(x);
      if (OP == NOPS) acc += x;
      /* cheap per-lane update so operands change every round */
      v[u] = x * 1664525u + 1013904223u + acc * (OP == NOPS ? 1u : 0u);
    }
  }
  if (acc == 0xdeadbeefu) out[0] = acc + s_hist[threadIdx.x];
  if (OP == ATOMS_NORET && threadIdx.x == 0) out[1] = s_hist[5];
}

template <int OP>
static void run(const char *name, int sms, u32 *out, double ghz) {
  const int iters = 4000, ctas_per_sm = 2;
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  k<OP><<<sms * ctas_per_sm, 512>>>(out, 100, 1);
  cudaEventRecord(a);
  k<OP><<<sms * ctas_per_sm, 512>>>(out, iters, 7);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  const double warp_inst_per_sm = (double)iters * 4 * 16 * ctas_per_sm;
  printf("%-14s %8.3f ms  %7.2f clk per warp-instr per SM (at %.3f GHz, incl. ~3 ALU ops of loop overhead)\n",
         name, ms, ms * 1e-3 * ghz * 1e9 / warp_inst_per_sm, ghz);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double ghz = khz * 1e-6;
  u32 *out; cudaMalloc(&out, 64);
  printf("%s, %d SMs, %.3f GHz nominal\n", p.name, p.multiProcessorCount, ghz);
  const int sms = p.multiProcessorCount;
  run<NOPS>("loop-only", sms, out, ghz);
  run<MATCH_RANDOM8>("match.any r256", sms, out, ghz);
  run<MATCH_RANDOM4>("match.any r16", sms, out, ghz);
  run<MATCH_RANDOM2>("match.any r4", sms, out, ghz);
  run<MATCH_UNIFORM>("match.any unif", sms, out, ghz);
  run<BALLOT>("ballot", sms, out, ghz);
  run<SHFL>("shfl.idx", sms, out, ghz);
  run<ATOMS_RET>("atoms ret r256", sms, out, ghz);
  run<ATOMS_NORET>("atoms noret", sms, out, ghz);
  run<ATOMS_LEADER>("atoms 1 lane", sms, out, ghz);
  run<POPC_FFS>("popc+ffs", sms, out, ghz);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
