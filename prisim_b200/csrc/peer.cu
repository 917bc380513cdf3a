// Peer-memory plumbing for the multi-GPU gather: the writing rank allocates the full output buffer, the other
// ranks map it over NVLink (CUDA IPC) and hand their slice to pb200_skyvis as the output pointer, so the
// kernel's epilogue (k_skyvis_finalize) stores the finished visibilities straight into the writing rank's HBM.
// This replaces the reference's exchange through _part_N.hdf5 files (scripts/run_prisim.py:1995, :2233-2276)
// without a separate collective.
#include "common.cuh"

extern "C" {

int pb200_device_alloc(pb200_ctx* ctx, size_t bytes, void** d_out) {
  if (!ctx || !d_out || bytes == 0) return ctx ? pb_fail(ctx, PB200_EINVAL, "pb200_device_alloc: bad arguments") : PB200_EINVAL;
  PbDeviceGuard guard(ctx->device);
  cudaError_t e = cudaMalloc(d_out, bytes);
  if (e != cudaSuccess) return pb_fail(ctx, PB200_ENOMEM, "cudaMalloc: %s", cudaGetErrorString(e));
  return PB200_OK;
}

int pb200_device_free(pb200_ctx* ctx, void* d_ptr) {
  if (!ctx) return PB200_EINVAL;
  PbDeviceGuard guard(ctx->device);
  PB_CUDA(ctx, cudaFree(d_ptr));
  return PB200_OK;
}

int pb200_peer_export(pb200_ctx* ctx, void* d_ptr, void* h_handle64) {
  if (!ctx || !d_ptr || !h_handle64) return ctx ? pb_fail(ctx, PB200_EINVAL, "pb200_peer_export: bad arguments") : PB200_EINVAL;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  PbDeviceGuard guard(ctx->device);
  cudaIpcMemHandle_t h;
  PB_CUDA(ctx, cudaIpcGetMemHandle(&h, d_ptr));
  memcpy(h_handle64, &h, 64);
  return PB200_OK;
}

int pb200_peer_open(pb200_ctx* ctx, const void* h_handle64, void** d_out) {
  if (!ctx || !h_handle64 || !d_out) return ctx ? pb_fail(ctx, PB200_EINVAL, "pb200_peer_open: bad arguments") : PB200_EINVAL;
  PbDeviceGuard guard(ctx->device);
  cudaIpcMemHandle_t h;
  memcpy(&h, h_handle64, 64);
  PB_CUDA(ctx, cudaIpcOpenMemHandle(d_out, h, cudaIpcMemLazyEnablePeerAccess));
  return PB200_OK;
}

int pb200_peer_close(pb200_ctx* ctx, void* d_ptr) {
  if (!ctx) return PB200_EINVAL;
  PbDeviceGuard guard(ctx->device);
  PB_CUDA(ctx, cudaIpcCloseMemHandle(d_ptr));
  return PB200_OK;
}

}  // extern "C"
