// Shared declarations for libprisim_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include "../../include/prisim_b200.h"

#define PB_CHAN_CACHE 4
struct pb200_ctx {
  int device;
  int sm_count;
  char err[512];
  long long launches;
  // scratch owned by the ctx (grown on demand, freed in pb200_ctx_destroy)
  void* scratch[8];
  size_t scratch_bytes[8];
  // cached twiddle table for the delay transform
  void* twiddle;
  int twiddle_n;
  // channel grids already resident on the device (pb_channels_device): host copy + device copy
  struct {
    double* host;
    double* dev;
    int nchan, npad;
    unsigned long long stamp;
  } chan[PB_CHAN_CACHE];
  unsigned long long chan_clock;
  int skyvis_spc_env;      // PB200_SKYVIS_SPC read once at ctx creation (developer override), 0 = unset
  int dt_force_r8;         // PB200_DT_R8=1: 1024-point delay transforms through k_delay_fft_r8 instead of k_delay_fft_w32 (A/B)
};

#define PB_SPEED_OF_LIGHT 299792458.0   // scipy.constants.c
#define PB_BOLTZMANN 1.380649e-23       // scipy.constants.k
#define PB_JY 1.0e-26

inline int pb_fail(pb200_ctx* ctx, int code, const char* fmt, const char* a = "", const char* b = "") {
  if (ctx) snprintf(ctx->err, sizeof(ctx->err), fmt, a, b);
  return code;
}

#define PB_CUDA(ctx, call)                                                                    \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess) return pb_fail((ctx), PB200_ECUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
  } while (0)

#define PB_CHECK_LAUNCH(ctx, name)                                                            \
  do {                                                                                        \
    (ctx)->launches++;                                                                        \
    cudaError_t e_ = cudaGetLastError();                                                      \
    if (e_ != cudaSuccess) return pb_fail((ctx), PB200_ECUDA, "launch %s: %s", name, cudaGetErrorString(e_)); \
  } while (0)

// grow-only scratch slot
inline int pb_scratch(pb200_ctx* ctx, int slot, size_t bytes, void** out) {
  if (ctx->scratch_bytes[slot] < bytes) {
    if (ctx->scratch[slot]) cudaFree(ctx->scratch[slot]);
    ctx->scratch[slot] = nullptr;
    ctx->scratch_bytes[slot] = 0;
    size_t want = bytes + (bytes >> 2) + 256;
    cudaError_t e = cudaMalloc(&ctx->scratch[slot], want);
    if (e != cudaSuccess) return pb_fail(ctx, PB200_ENOMEM, "cudaMalloc scratch: %s", cudaGetErrorString(e));
    ctx->scratch_bytes[slot] = want;
  }
  *out = ctx->scratch[slot];
  return PB200_OK;
}

// Entry points run on ctx->device and put the caller's current device back on return (single-process
// multi-GPU callers, e.g. torch with several devices, keep their own current device).
struct PbDeviceGuard {
  int prev;
  bool changed;
  explicit PbDeviceGuard(int device) : prev(-1), changed(false) {
    if (cudaGetDevice(&prev) == cudaSuccess && prev != device) changed = (cudaSetDevice(device) == cudaSuccess);
  }
  ~PbDeviceGuard() { if (changed) cudaSetDevice(prev); }
  PbDeviceGuard(const PbDeviceGuard&) = delete;
  PbDeviceGuard& operator=(const PbDeviceGuard&) = delete;
};

// Channel frequencies on the device.  The reference passes the same `channels` array to every call
// (interferometry.py:5779-5786), so each ctx keeps the last PB_CHAN_CACHE grids resident: a call with a grid
// seen before enqueues nothing and does not synchronise; a new grid is uploaded once (that call synchronises
// the stream so the cached copy is complete before any other stream may use it).  `npad` >= nchan entries are
// stored, padding repeats the last frequency.
inline int pb_channels_device(pb200_ctx* ctx, const double* h_freqs, int nchan, int npad, cudaStream_t stream,
                              const double** d_out) {
  if (npad < nchan) npad = nchan;
  int victim = 0;
  for (int i = 0; i < PB_CHAN_CACHE; ++i) {
    auto& e = ctx->chan[i];
    if (e.dev && e.nchan == nchan && e.npad == npad && memcmp(e.host, h_freqs, sizeof(double) * nchan) == 0) {
      e.stamp = ++ctx->chan_clock;
      *d_out = e.dev;
      return PB200_OK;
    }
    if (ctx->chan[i].stamp < ctx->chan[victim].stamp) victim = i;
  }
  auto& e = ctx->chan[victim];
  if (e.dev) { cudaFree(e.dev); e.dev = nullptr; }
  if (e.host) { cudaFreeHost(e.host); e.host = nullptr; }
  cudaError_t err = cudaMallocHost((void**)&e.host, sizeof(double) * (size_t)npad);
  if (err == cudaSuccess) err = cudaMalloc((void**)&e.dev, sizeof(double) * (size_t)npad);
  if (err != cudaSuccess) {
    if (e.host) { cudaFreeHost(e.host); e.host = nullptr; }
    e.dev = nullptr;
    return pb_fail(ctx, PB200_ENOMEM, "channel grid allocation: %s", cudaGetErrorString(err));
  }
  for (int k = 0; k < npad; ++k) e.host[k] = h_freqs[k < nchan ? k : nchan - 1];
  err = cudaMemcpyAsync(e.dev, e.host, sizeof(double) * (size_t)npad, cudaMemcpyHostToDevice, stream);
  if (err == cudaSuccess) err = cudaStreamSynchronize(stream);
  if (err != cudaSuccess) {
    cudaFree(e.dev); cudaFreeHost(e.host);
    e.dev = nullptr; e.host = nullptr;
    return pb_fail(ctx, PB200_ECUDA, "channel grid upload: %s", cudaGetErrorString(err));
  }
  e.nchan = nchan; e.npad = npad; e.stamp = ++ctx->chan_clock;
  *d_out = e.dev;
  return PB200_OK;
}

static inline int pb_div_up(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---- mbarrier / TMA bulk-copy helpers (skyvis.cu, delay_transform.cu) ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// the same with an L2 cache policy (createpolicy): streams that are read once per wave are loaded evict-first
__device__ __forceinline__ void tma_bulk_g2s_hint(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t policy) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
               : "memory");
}
// TMA bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

