// Context management for libprisim_b200.so.
#include "common.cuh"
#include <cstdlib>
#include <new>

extern "C" {

int pb200_version(void) { return PB200_VERSION; }

int pb200_ctx_create(pb200_ctx** out, int device) {
  if (!out) return PB200_EINVAL;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) return PB200_ECUDA;   // no CPU fallback: fail loudly
  if (device < 0 || device >= n) return PB200_EINVAL;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return PB200_ECUDA;
  if (prop.major != 10) return PB200_EUNSUPPORTED;                           // sm_100a cubins only
  pb200_ctx* ctx = new (std::nothrow) pb200_ctx();
  if (!ctx) return PB200_ENOMEM;
  memset(ctx, 0, sizeof(*ctx));
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  // L2 set-aside for lines accessed with an evict-last policy (the fp64 running sums of the phase-sum kernel, 39 MB live per launch);
  // without it evict-last degrades to normal replacement.  Device-wide and harmless to other users of the device.
  {
    int prev = -1;
    const char* l2 = getenv("PB200_L2_PERSIST_MB");
    const size_t want = (size_t)(l2 ? atoi(l2) : 64) << 20;
    if (want > 0 && cudaGetDevice(&prev) == cudaSuccess && cudaSetDevice(device) == cudaSuccess) {
      const size_t cap = (size_t)prop.persistingL2CacheMaxSize;
      cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want < cap ? want : cap);
      cudaGetLastError();
      if (prev != device) cudaSetDevice(prev);
    }
  }
  const char* spc = getenv("PB200_SKYVIS_SPC");          // developer override of the phase-sum CTA shape (tools/variants.sh)
  ctx->skyvis_spc_env = spc ? atoi(spc) : 0;
  const char* r8 = getenv("PB200_DT_R8");
  ctx->dt_force_r8 = r8 ? atoi(r8) : 0;
  *out = ctx;
  return PB200_OK;
}

void pb200_ctx_destroy(pb200_ctx* ctx) {
  if (!ctx) return;
  PbDeviceGuard guard(ctx->device);
  for (int i = 0; i < 8; ++i)
    if (ctx->scratch[i]) cudaFree(ctx->scratch[i]);
  if (ctx->twiddle) cudaFree(ctx->twiddle);
  for (int i = 0; i < PB_CHAN_CACHE; ++i) {
    if (ctx->chan[i].dev) cudaFree(ctx->chan[i].dev);
    if (ctx->chan[i].host) cudaFreeHost(ctx->chan[i].host);
  }
  delete ctx;
}

// developer / test knobs (the environment variables PB200_SKYVIS_SPC and PB200_DT_R8 set the same fields once, at ctx creation)
int pb200_ctx_set_option(pb200_ctx* ctx, const char* name, long long value) {
  if (!ctx || !name) return PB200_EINVAL;
  if (!strcmp(name, "skyvis_spc")) {
    if (value != 0 && value != 1 && value != 2 && value != 4) return pb_fail(ctx, PB200_EINVAL, "skyvis_spc must be 0 (automatic), 1, 2 or 4");
    ctx->skyvis_spc_env = (int)value;
  } else if (!strcmp(name, "dt_force_r8")) {
    ctx->dt_force_r8 = value != 0;
  } else {
    return pb_fail(ctx, PB200_EINVAL, "pb200_ctx_set_option: unknown option %s", name);
  }
  return PB200_OK;
}

const char* pb200_last_error(const pb200_ctx* ctx) { return ctx ? ctx->err : "null ctx"; }

long long pb200_launch_count(const pb200_ctx* ctx) { return ctx ? ctx->launches : 0; }

}  // extern "C"
