/*
  tmrgpu_cuda.cu -- the product translation unit: every kernel body of
  ops_*.h instantiated through the CUDA primitives of prim_cuda.cuh for
  sm_100a, plus context creation.  There is no host fallback: context creation
  fails (and every forest call after it) when no CUDA device is usable.
*/
#include <cuda_runtime.h>

#include "tmrgpu_api.inl"

extern "C" {

int tmrgpu_ctx_create(int device, void *stream, tmrgpu_ctx **out) {
  *out = NULL;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0) {
    fprintf(stderr, "TMROctForest Error: no CUDA device available (%s)\n",
            e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    return 1;
  }
  if (device < 0 || device >= ndev) device = device % ndev;
  e = cudaSetDevice(device);
  if (e != cudaSuccess) {
    fprintf(stderr, "TMROctForest Error: cudaSetDevice(%d): %s\n", device,
            cudaGetErrorString(e));
    return 1;
  }
  tmrgpu_ctx *c = new tmrgpu_ctx();
  c->c.device = device;
  c->own_stream = false;
  if (stream) {
    c->c.stream = stream;
  } else {
    cudaStream_t s;
    e = cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
      fprintf(stderr, "TMROctForest Error: cudaStreamCreate: %s\n",
              cudaGetErrorString(e));
      delete c;
      return 1;
    }
    c->c.stream = s;
    c->own_stream = true;
  }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) {
    c->c.num_sms = prop.multiProcessorCount;
  }
  c->c.trace = getenv("TMR_B200_TRACE") ? atoi(getenv("TMR_B200_TRACE")) : 0;
  if (getenv("TMR_B200_LAUNCH_LOG")) c->c.launch_log = fopen(getenv("TMR_B200_LAUNCH_LOG"), "w");
  *out = c;
  return 0;
}

int tmrgpu_ctx_destroy(tmrgpu_ctx *ctx) {
  if (!ctx) return 0;
  prof_resolve(ctx->c);
  cudaStreamSynchronize((cudaStream_t)ctx->c.stream);
  dev_cache_destroy(ctx->c);
  if (ctx->c.launch_log) fclose(ctx->c.launch_log);
  if (ctx->c.mailbox) cudaFreeHost(ctx->c.mailbox);
  if (ctx->c.copy_stream) cudaStreamDestroy((cudaStream_t)ctx->c.copy_stream);
  if (ctx->c.copy_stream2) cudaStreamDestroy((cudaStream_t)ctx->c.copy_stream2);
  if (ctx->own_stream) cudaStreamDestroy((cudaStream_t)ctx->c.stream);
  delete ctx;
  return 0;
}

const char *tmrgpu_build_kind(void) { return "cuda-sm_100a"; }

}  // extern "C"
