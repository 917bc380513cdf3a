// Issue-rate microbenchmarks for the pipes the skyvis kernel leans on (FP32 FMA, packed
// FFMA2, MUFU/XU, FP64, shared-memory loads) on sm_100a.  Stand-alone program: prints one
// JSON object.  MEASURED_PEAKS.json has no FP32/XU entry, so the roofline denominator quoted
// by bench.py ("of measured FFMA peak") comes from here.
//
// Each kernel runs ITER iterations of an unrolled body on every thread of a grid that fills
// all SMs with 4 warps per scheduler; throughput is reported both per wall-clock (CUDA
// events) and per SM cycle (clock64 deltas), so it is independent of the DVFS state.
#include "common.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int NCH = 16;          // independent chains per thread
constexpr int ITER = 4096;

__device__ __forceinline__ unsigned long long pk(float a, float b) {
  unsigned long long r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r;
}
__device__ __forceinline__ float lo(unsigned long long v) { float a, b; asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return a + b; }

// 1. scalar FFMA, 3 register operands: acc = acc * x + y
__global__ void k_ffma(float* out, float x, float y, long long* cyc) {
  float acc[NCH];
#pragma unroll
  for (int j = 0; j < NCH; ++j) acc[j] = threadIdx.x * 1e-3f + j;
  long long t0 = clock64();
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int j = 0; j < NCH; ++j) acc[j] = fmaf(acc[j], x, y);
  }
  long long t1 = clock64();
  float s = 0; 
#pragma unroll
  for (int j = 0; j < NCH; ++j) s += acc[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// 2. packed FFMA2: acc2 = acc2 * x2 + y2
__global__ void k_ffma2(float* out, float x, float y, long long* cyc) {
  unsigned long long acc[NCH];
  unsigned long long x2 = pk(x, x * 1.0001f), y2 = pk(y, y * 0.999f);
#pragma unroll
  for (int j = 0; j < NCH; ++j) acc[j] = pk(threadIdx.x * 1e-3f + j, j);
  long long t0 = clock64();
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int j = 0; j < NCH; ++j) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[j]) : "l"(x2), "l"(y2));
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int j = 0; j < NCH; ++j) s += lo(acc[j]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// 3. the scalar rotate+accumulate inner loop (6 FMA-pipe issues per term), 4 independent
//    (source,baseline) phasors x 4 channels each per iteration
__global__ void k_rot_scalar(float* out, float x, float y, long long* cyc) {
  float pr[4], pi[4], rr[4], ri[4], ar[16], ai[16];
#pragma unroll
  for (int j = 0; j < 4; ++j) { pr[j] = 1.f; pi[j] = 0.f; rr[j] = x + j * 1e-6f; ri[j] = y; }
#pragma unroll
  for (int j = 0; j < 16; ++j) { ar[j] = 0; ai[j] = 0; }
  float a = threadIdx.x * 1e-4f + 1.f;
  long long t0 = clock64();
  for (int it = 0; it < ITER / 4; ++it) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float nr = pr[j] * rr[j]; nr = fmaf(-pi[j], ri[j], nr);
        float ni = pr[j] * ri[j]; ni = fmaf(pi[j], rr[j], ni);
        pr[j] = nr; pi[j] = ni;
        ar[c * 4 + j] = fmaf(a, nr, ar[c * 4 + j]);
        ai[c * 4 + j] = fmaf(a, ni, ai[c * 4 + j]);
      }
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int j = 0; j < 16; ++j) s += ar[j] + ai[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// 4. the packed rotate+accumulate inner loop (3 FFMA2-class issues per term): 2 channels per
//    64-bit register, 4 independent phasor pairs
__global__ void k_rot_packed(float* out, float x, float y, long long* cyc) {
  unsigned long long PR[4], PI[4], RR[4], RI[4], NRI[4], AR[16], AI[16];
#pragma unroll
  for (int j = 0; j < 4; ++j) { PR[j] = pk(1.f, x); PI[j] = pk(0.f, y); RR[j] = pk(x + j * 1e-6f, x + j * 1e-6f); RI[j] = pk(y, y); NRI[j] = pk(-y, -y); }
#pragma unroll
  for (int j = 0; j < 16; ++j) { AR[j] = 0ull; AI[j] = 0ull; }
  unsigned long long A = pk(threadIdx.x * 1e-4f + 1.f, threadIdx.x * 1e-4f + 1.1f);
  long long t0 = clock64();
  for (int it = 0; it < ITER / 4; ++it) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        unsigned long long nr, ni;
        asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(nr) : "l"(PR[j]), "l"(RR[j]));
        asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(nr) : "l"(PI[j]), "l"(NRI[j]));
        asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(ni) : "l"(PR[j]), "l"(RI[j]));
        asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(ni) : "l"(PI[j]), "l"(RR[j]));
        PR[j] = nr; PI[j] = ni;
        asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(AR[c * 4 + j]) : "l"(A), "l"(nr));
        asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(AI[c * 4 + j]) : "l"(A), "l"(ni));
      }
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int j = 0; j < 16; ++j) s += lo(AR[j]) + lo(AI[j]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// 5. MUFU: sin.approx + cos.approx pairs (each = FMUL by 1/2pi + MUFU)
__global__ void k_mufu(float* out, float x, float y, long long* cyc) {
  float acc[NCH];
#pragma unroll
  for (int j = 0; j < NCH; ++j) acc[j] = threadIdx.x * 1e-3f + j * 0.1f;
  long long t0 = clock64();
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int j = 0; j < NCH; ++j) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(acc[j]));
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int j = 0; j < NCH; ++j) s += acc[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// 6. FP64 DFMA
__global__ void k_dfma(float* out, float x, float y, long long* cyc) {
  double acc[NCH]; double xd = x, yd = y;
#pragma unroll
  for (int j = 0; j < NCH; ++j) acc[j] = threadIdx.x * 1e-3 + j;
  long long t0 = clock64();
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int j = 0; j < NCH; ++j) acc[j] = fma(acc[j], xd, yd);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int j = 0; j < NCH; ++j) s += acc[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = (float)s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// 7. mixes: NF packed FFMA2 + NM MUFU + ND DFMA + NL LDS.128 per body; shows whether the
//    non-FMA work hides in the issue slots FFMA2 leaves free
template <int NM, int ND, int NL>
__global__ void k_mix(float* out, float x, float y, long long* cyc) {
  __shared__ float4 sm[256];
  sm[threadIdx.x] = make_float4(x, y, x, y);
  __syncthreads();
  unsigned long long acc[NCH];
  unsigned long long x2 = pk(x, x * 1.0001f), y2 = pk(y, y * 0.999f);
  float m[4]; double d[4]; float4 l = make_float4(0, 0, 0, 0);
#pragma unroll
  for (int j = 0; j < NCH; ++j) acc[j] = pk(threadIdx.x * 1e-3f + j, j);
#pragma unroll
  for (int j = 0; j < 4; ++j) { m[j] = threadIdx.x * 1e-3f + j * 0.1f; d[j] = threadIdx.x * 1e-3 + j; }
  double xd = x, yd = y;
  long long t0 = clock64();
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int j = 0; j < NCH; ++j) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[j]) : "l"(x2), "l"(y2));
#pragma unroll
    for (int j = 0; j < NM; ++j) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(m[j & 3]));
#pragma unroll
    for (int j = 0; j < ND; ++j) d[j & 3] = fma(d[j & 3], xd, yd);
#pragma unroll
    for (int j = 0; j < NL; ++j) {
      float4 v;
      asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                   : "r"((unsigned)__cvta_generic_to_shared(&sm[(it + j) & 255])));
      l.x += v.x;
    }
  }
  long long t1 = clock64();
  float s = l.x;
#pragma unroll
  for (int j = 0; j < NCH; ++j) s += lo(acc[j]);
#pragma unroll
  for (int j = 0; j < 4; ++j) s += m[j] + (float)d[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

struct Res { double ms; double cyc; };

template <typename K>
Res run(K kern, int blocks, int threads, float* out, long long* cyc) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int w = 0; w < 3; ++w) kern<<<blocks, threads>>>(out, 0.999f, 1e-3f, cyc);
  CK(cudaDeviceSynchronize());
  std::vector<double> ts;
  for (int r = 0; r < 5; ++r) {
    CK(cudaEventRecord(e0)); kern<<<blocks, threads>>>(out, 0.999f, 1e-3f, cyc); CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ts.push_back(ms);
  }
  std::sort(ts.begin(), ts.end());
  std::vector<long long> h(blocks); CK(cudaMemcpy(h.data(), cyc, blocks * sizeof(long long), cudaMemcpyDeviceToHost));
  double mx = 0; for (auto v : h) mx = std::max(mx, (double)v);
  return {ts[ts.size() / 2], mx};
}

#ifdef PB200_MICROBENCH_MAIN
int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount;
  const int threads = 256, bps = 2;           // 16 warps/SM = 4 per scheduler
  int blocks = sms * bps;
  float* out; long long* cyc;
  CK(cudaMalloc(&out, sizeof(float) * blocks * threads)); CK(cudaMalloc(&cyc, sizeof(long long) * blocks));
  double warp_per_sm = threads / 32.0 * bps;
  auto rep = [&](const char* name, Res r, double lane_ops_per_thread, const char* unit, bool last = false) {
    double total = lane_ops_per_thread * blocks * threads;
    double per_clk_sm = lane_ops_per_thread * 32.0 * warp_per_sm / r.cyc;
    printf("  \"%s\": {\"ms\": %.4f, \"cycles\": %.0f, \"%s_per_s\": %.4e, \"lane_ops_per_clk_per_sm\": %.2f, \"eff_mhz\": %.0f}%s\n",
           name, r.ms, r.cyc, unit, total / (r.ms * 1e-3), per_clk_sm, r.cyc / (r.ms * 1e-3) / 1e6, last ? "" : ",");
  };
  printf("{\n  \"device\": \"%s\", \"sms\": %d, \"clock_khz_attr\": %d,\n", p.name, sms, p.clockRate);
  rep("ffma_scalar", run(k_ffma, blocks, threads, out, cyc), (double)ITER * NCH, "fma");
  rep("ffma2_packed", run(k_ffma2, blocks, threads, out, cyc), (double)ITER * NCH * 2, "fma");
  rep("rotacc_scalar_terms", run(k_rot_scalar, blocks, threads, out, cyc), (double)(ITER / 4) * 16, "terms");
  rep("rotacc_packed_terms", run(k_rot_packed, blocks, threads, out, cyc), (double)(ITER / 4) * 16 * 2, "terms");
  rep("mufu_ex2", run(k_mufu, blocks, threads, out, cyc), (double)ITER * NCH, "mufu");
  rep("dfma", run(k_dfma, blocks, threads, out, cyc), (double)ITER * NCH, "dfma");
  rep("mix_ffma2x16_only", run(k_mix<0, 0, 0>, blocks, threads, out, cyc), (double)ITER * NCH * 2, "fma");
  rep("mix_ffma2x16_mufu2", run(k_mix<2, 0, 0>, blocks, threads, out, cyc), (double)ITER * NCH * 2, "fma");
  rep("mix_ffma2x16_mufu4", run(k_mix<4, 0, 0>, blocks, threads, out, cyc), (double)ITER * NCH * 2, "fma");
  rep("mix_ffma2x16_dfma2", run(k_mix<0, 2, 0>, blocks, threads, out, cyc), (double)ITER * NCH * 2, "fma");
  rep("mix_ffma2x16_dfma4", run(k_mix<0, 4, 0>, blocks, threads, out, cyc), (double)ITER * NCH * 2, "fma");
  rep("mix_ffma2x16_lds4", run(k_mix<0, 0, 4>, blocks, threads, out, cyc), (double)ITER * NCH * 2, "fma");
  rep("mix_ffma2x16_m2_d2_l4", run(k_mix<2, 2, 4>, blocks, threads, out, cyc), (double)ITER * NCH * 2, "fma", true);
  printf("}\n");
  return 0;
}
#endif  // PB200_MICROBENCH_MAIN

// Library entry: the three pipe rates bench.py quotes as roofline denominators.
extern "C" int pb200_microbench(pb200_ctx* ctx, double* out, int n) {
  if (!ctx || !out || n < 5) return ctx ? pb_fail(ctx, PB200_EINVAL, "pb200_microbench: need out[5]") : PB200_EINVAL;
  if (cudaSetDevice(ctx->device) != cudaSuccess) return pb_fail(ctx, PB200_ECUDA, "cudaSetDevice failed");
  const int threads = 256, bps = 2;
  int blocks = ctx->sm_count * bps;
  float* o; long long* cyc;
  if (cudaMalloc(&o, sizeof(float) * blocks * threads) != cudaSuccess) return pb_fail(ctx, PB200_ENOMEM, "microbench alloc");
  if (cudaMalloc(&cyc, sizeof(long long) * blocks) != cudaSuccess) { cudaFree(o); return pb_fail(ctx, PB200_ENOMEM, "microbench alloc"); }
  double warp_per_sm = threads / 32.0 * bps;
  Res f = run(k_ffma, blocks, threads, o, cyc);
  Res m = run(k_mufu, blocks, threads, o, cyc);
  Res d = run(k_dfma, blocks, threads, o, cyc);
  ctx->launches += 3 * 8;
  double per_thread = (double)ITER * NCH, total = per_thread * blocks * threads;
  out[0] = total / (f.ms * 1e-3);
  out[1] = per_thread * 32.0 * warp_per_sm / f.cyc;
  out[2] = total / (m.ms * 1e-3);
  out[3] = total / (d.ms * 1e-3);
  out[4] = f.cyc / (f.ms * 1e-3);
  cudaFree(o); cudaFree(cyc);
  return PB200_OK;
}
