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

// 4b. the same loop in the operand-reuse order (each instruction keeps one operand of its predecessor in the same
//     slot): t1 = PR*RR, t2 = PR*RI, are += PR*A, aim += PI*A, nr = PI*NRI + t1, ni = PI*RR + t2.  LDSA = 1: the
//     amplitudes come from shared memory (one LDS.128 broadcast per two steps per source), as in the kernel.
template <int LDSA>
__global__ void k_rot_packed_reuse(float* out, float x, float y, long long* cyc) {
  __shared__ float4 sm[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = make_float4(1.f + i * 1e-4f, 1.f, 1.01f, 0.99f);
  __syncthreads();
  unsigned long long PR[4], PI[4], RR[4], RI[4], NRI[4], AR[16], AI[16];
#pragma unroll
  for (int j = 0; j < 4; ++j) { PR[j] = pk(1.f, x); PI[j] = pk(0.f, y); RR[j] = pk(x + j * 1e-6f, x + j * 1e-6f); RI[j] = pk(y, y); NRI[j] = pk(-y, -y); }
#pragma unroll
  for (int j = 0; j < 16; ++j) { AR[j] = 0ull; AI[j] = 0ull; }
  unsigned long long A0 = pk(threadIdx.x * 1e-4f + 1.f, threadIdx.x * 1e-4f + 1.1f);
  long long t0 = clock64();
  for (int it = 0; it < ITER / 16; ++it) {
#pragma unroll
    for (int c = 0; c < 16; ++c) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        unsigned long long A = A0;
        if (LDSA) {
          const float4 v = sm[((it * 4 + j) * 8 + (c >> 1)) & 1023];
          A = (c & 1) ? pk(v.z, v.w) : pk(v.x, v.y);
        }
        unsigned long long t1, t2;
        asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(t1) : "l"(PR[j]), "l"(RR[j]));
        asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(t2) : "l"(PR[j]), "l"(RI[j]));
        asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(AR[c]) : "l"(PR[j]), "l"(A));
        asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(AI[c]) : "l"(PI[j]), "l"(A));
        asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(PR[j]) : "l"(PI[j]), "l"(NRI[j]), "l"(t1));
        asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(PI[j]) : "l"(PI[j]), "l"(RR[j]), "l"(t2));
      }
    }
  }
  long long t1c = clock64();
  float s = 0;
#pragma unroll
  for (int j = 0; j < 16; ++j) s += lo(AR[j]) + lo(AI[j]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1c - t0;
}

// 4c. the three-term packed loop of k_skyvis (Z_{j+1} = C Z_j - Z_{j-1}; accumulate A Z_j), 2 sources x 2 half blocks
//     per iteration, 64 FFMA2-class issues per source.  VAR 0: as in the kernel (C a 32-bit splat, negated addend);
//     1: C with distinct halves (full 64-bit operand); 2: no negation (Z_{j+1} = C Z_j + Z_{j-1}, diverges, timing only);
//     3: accumulate only (no recurrence); 4: recurrence only (no accumulate)
template <int VAR>
__global__ void k_three_term(float* out, float x, float y, long long* cyc) {
  float2 are[16], aim[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) { are[j] = make_float2(0.f, 0.f); aim[j] = make_float2(0.f, 0.f); }
  const float2 A0 = make_float2(threadIdx.x * 1e-4f + 1.f, 1.1f), A1 = make_float2(0.9f, threadIdx.x * 1e-4f + 1.f);
  long long t0 = clock64();
  for (int it = 0; it < ITER / 16; ++it) {
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const float c = x + (it + s) * 1e-7f;
      const float2 CC = VAR == 1 ? make_float2(c, c * 1.0000001f) : make_float2(c, c);
      const float2 RR = make_float2(c * 0.5f, c * 0.5f), RI = make_float2(y, y), NRI = make_float2(-y, -y);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float2 PRm = make_float2(1.f, c), PIm = make_float2(0.f, y + h);
        const float2 t1 = __fmul2_rn(PRm, RR), t2 = __fmul2_rn(PRm, RI);
        float2 PRc = __ffma2_rn(PIm, NRI, t1), PIc = __ffma2_rn(PIm, RR, t2);
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) {
          const int j = h * 8 + 2 * k4;
          if (k4 > 0 && VAR != 3) {
            const float sg = VAR == 2 ? 1.f : -1.f;
            const float2 PRn = __ffma2_rn(CC, PRc, make_float2(sg * PRm.x, sg * PRm.y));
            const float2 PIn = __ffma2_rn(CC, PIc, make_float2(sg * PIm.x, sg * PIm.y));
            PRm = __ffma2_rn(CC, PRn, make_float2(sg * PRc.x, sg * PRc.y));
            PIm = __ffma2_rn(CC, PIn, make_float2(sg * PIc.x, sg * PIc.y));
            const float2 tr = PRm, ti = PIm;
            PRm = PRn; PIm = PIn; PRc = tr; PIc = ti;
          }
          if (VAR != 4) {
            are[j] = __ffma2_rn(PRm, A0, are[j]);
            aim[j] = __ffma2_rn(PIm, A0, aim[j]);
            are[j + 1] = __ffma2_rn(PRc, A1, are[j + 1]);
            aim[j + 1] = __ffma2_rn(PIc, A1, aim[j + 1]);
          } else {
            are[j].x += PRm.x + PIm.x + PRc.x + PIc.x;       // keep the chain alive (scalar FADDs, timing of VAR 4 is indicative only)
          }
        }
      }
    }
  }
  long long t1c = clock64();
  float sum = 0;
#pragma unroll
  for (int j = 0; j < 16; ++j) sum += are[j].x + are[j].y + aim[j].x + aim[j].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = sum;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1c - t0;
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

// 8. FP32-pipe + FP64-pipe co-issue: NF independent FFMA2 chains interleaved with ND independent DFMA chains.
template <int NF, int ND>
__global__ void k_coissue(float* out, float x, float y, long long* cyc) {
  unsigned long long acc[NF];
  unsigned long long x2 = pk(x, x * 1.0001f), y2 = pk(y, y * 0.999f);
  double d[ND]; double xd = x, yd = y;
#pragma unroll
  for (int j = 0; j < NF; ++j) acc[j] = pk(threadIdx.x * 1e-3f + j, j);
#pragma unroll
  for (int j = 0; j < ND; ++j) d[j] = threadIdx.x * 1e-3 + j;
  constexpr int NMAX = NF > ND ? NF : ND;
  long long t0 = clock64();
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int j = 0; j < NMAX; ++j) {
      if (j * NF / NMAX != (j + 1) * NF / NMAX || NF == NMAX) {
        const int k = NF == NMAX ? j : j * NF / NMAX;
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[k]) : "l"(x2), "l"(y2));
      }
      if (j * ND / NMAX != (j + 1) * ND / NMAX || ND == NMAX) {
        const int k = ND == NMAX ? j : j * ND / NMAX;
        asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[k]) : "d"(xd), "d"(yd));
      }
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int j = 0; j < NF; ++j) s += lo(acc[j]);
#pragma unroll
  for (int j = 0; j < ND; ++j) s += (float)d[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// 9. hybrid rotate+accumulate: warps with bit 2 of the warp id clear run the packed fp32 loop over 32 channels per
//    source (96 FFMA2), the others the same recurrence in fp64 over 16 channels per source (96 DFMA), i.e. each
//    scheduler holds 2 + 2 warps at 4 warps per scheduler; NS independent sources in flight per thread.  AMP = 1:
//    amplitudes come from shared memory (LDS.128 broadcast) and the fp64 warps convert them (cvt.f64.f32).
//    FRAC64: 1 = as described, 0 = all warps fp32.
template <int AMP, int FRAC64, int NS>
__global__ void k_hybrid(float* out, float x, float y, long long* cyc) {
  __shared__ float4 sm[512];
  for (int i = threadIdx.x; i < 512; i += blockDim.x) sm[i] = make_float4(1.f + i * 1e-4f, 1.f, 1.01f, 0.99f);
  __syncthreads();
  const bool role64 = FRAC64 && ((threadIdx.x >> 7) & 1);
  float s = 0;
  long long t0 = clock64();
  if (!role64) {
    unsigned long long AR[16], AI[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) { AR[j] = 0ull; AI[j] = 0ull; }
    for (int it = 0; it < ITER / 16 / NS; ++it) {
      unsigned long long PR[NS], PI[NS], RR[NS], RI[NS], NRI[NS], A[NS];
#pragma unroll
      for (int j = 0; j < NS; ++j) {
        PR[j] = pk(1.f, x); PI[j] = pk(0.f, y); RR[j] = pk(x + (it + j) * 1e-6f, x + it * 1e-6f); RI[j] = pk(y, y); NRI[j] = pk(-y, -y);
        A[j] = pk(threadIdx.x * 1e-4f + 1.f, threadIdx.x * 1e-4f + 1.1f);
      }
#pragma unroll
      for (int c = 0; c < 16; ++c) {
#pragma unroll
        for (int j = 0; j < NS; ++j) {
          if (AMP && (c & 1) == 0) {
            float4 v = sm[((it * NS + j) * 8 + (c >> 1)) & 511];
            A[j] = pk(v.x, v.y);
            if (c & 2) A[j] = pk(v.z, v.w);
          }
          unsigned long long nr, ni;
          asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(nr) : "l"(PR[j]), "l"(RR[j]));
          asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(nr) : "l"(PI[j]), "l"(NRI[j]));
          asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(ni) : "l"(PR[j]), "l"(RI[j]));
          asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(ni) : "l"(PI[j]), "l"(RR[j]));
          PR[j] = nr; PI[j] = ni;
          asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(AR[c]) : "l"(A[j]), "l"(nr));
          asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(AI[c]) : "l"(A[j]), "l"(ni));
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) s += lo(AR[j]) + lo(AI[j]);
  } else {
    double ar[16], ai[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) { ar[j] = 0; ai[j] = 0; }
    for (int it = 0; it < ITER / 16 / NS; ++it) {
      double pr[NS], pi[NS], rr[NS], ri[NS], a[NS];
      float4 v[NS];
#pragma unroll
      for (int j = 0; j < NS; ++j) { pr[j] = 1.0; pi[j] = 0.0; rr[j] = (double)(x + (it + j) * 1e-6f); ri[j] = (double)y; a[j] = threadIdx.x * 1e-4 + 1.0; }
#pragma unroll
      for (int c = 0; c < 16; ++c) {
#pragma unroll
        for (int j = 0; j < NS; ++j) {
          if (AMP) {
            if ((c & 3) == 0) v[j] = sm[((it * NS + j) * 4 + (c >> 2)) & 511];
            a[j] = (double)((c & 3) == 0 ? v[j].x : (c & 3) == 1 ? v[j].y : (c & 3) == 2 ? v[j].z : v[j].w);
          }
          double nr, ni;
          asm volatile("mul.rn.f64 %0, %1, %2;" : "=d"(nr) : "d"(pr[j]), "d"(rr[j]));
          asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(nr) : "d"(-pi[j]), "d"(ri[j]));
          asm volatile("mul.rn.f64 %0, %1, %2;" : "=d"(ni) : "d"(pr[j]), "d"(ri[j]));
          asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(ni) : "d"(pi[j]), "d"(rr[j]));
          pr[j] = nr; pi[j] = ni;
          asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(ar[c]) : "d"(a[j]), "d"(nr));
          asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(ai[c]) : "d"(a[j]), "d"(ni));
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) s += (float)(ar[j] + ai[j]);
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// 10. dependent-issue latency of one chain per thread, one warp per scheduler
template <int OP>
__global__ void k_latency(float* out, float x, float y, long long* cyc) {
  double d = threadIdx.x * 1e-3, xd = x, yd = y; float f = threadIdx.x * 1e-3f;
  unsigned long long p = pk(f, f), x2 = pk(x, x), y2 = pk(y, y);
  long long t0 = clock64();
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      if (OP == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f) : "f"(x), "f"(y));
      if (OP == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p) : "l"(x2), "l"(y2));
      if (OP == 2) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d) : "d"(xd), "d"(yd));
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = f + lo(p) + (float)d;
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
  rep("rotacc_packed_reuse_order_terms", run(k_rot_packed_reuse<0>, blocks, threads, out, cyc), (double)(ITER / 16) * 16 * 4 * 2, "terms");
  rep("rotacc_packed_reuse_order_lds_terms", run(k_rot_packed_reuse<1>, blocks, threads, out, cyc), (double)(ITER / 16) * 16 * 4 * 2, "terms");
  rep("three_term_kernel_form_terms", run(k_three_term<0>, blocks, threads, out, cyc), (double)(ITER / 16) * 2 * 32, "terms");
  rep("three_term_c64_terms", run(k_three_term<1>, blocks, threads, out, cyc), (double)(ITER / 16) * 2 * 32, "terms");
  rep("three_term_noneg_terms", run(k_three_term<2>, blocks, threads, out, cyc), (double)(ITER / 16) * 2 * 32, "terms");
  rep("three_term_acc_only_terms", run(k_three_term<3>, blocks, threads, out, cyc), (double)(ITER / 16) * 2 * 32, "terms");
  rep("three_term_rec_only_terms", run(k_three_term<4>, blocks, threads, out, cyc), (double)(ITER / 16) * 2 * 32, "terms");
  rep("mufu_ex2", run(k_mufu, blocks, threads, out, cyc), (double)ITER * NCH, "mufu");
  rep("dfma", run(k_dfma, blocks, threads, out, cyc), (double)ITER * NCH, "dfma");
  rep("mix_ffma2x16_only", run(k_mix<0, 0, 0>, blocks, threads, out, cyc), (double)ITER * NCH * 2, "fma");
  rep("mix_ffma2x16_mufu2", run(k_mix<2, 0, 0>, blocks, threads, out, cyc), (double)ITER * NCH * 2, "fma");
  rep("mix_ffma2x16_mufu4", run(k_mix<4, 0, 0>, blocks, threads, out, cyc), (double)ITER * NCH * 2, "fma");
  rep("mix_ffma2x16_dfma2", run(k_mix<0, 2, 0>, blocks, threads, out, cyc), (double)ITER * NCH * 2, "fma");
  rep("mix_ffma2x16_dfma4", run(k_mix<0, 4, 0>, blocks, threads, out, cyc), (double)ITER * NCH * 2, "fma");
  rep("mix_ffma2x16_lds4", run(k_mix<0, 0, 4>, blocks, threads, out, cyc), (double)ITER * NCH * 2, "fma");
  rep("mix_ffma2x16_m2_d2_l4", run(k_mix<2, 2, 4>, blocks, threads, out, cyc), (double)ITER * NCH * 2, "fma");
  // co-issue: the value is FFMA-lane-equivalents, counting one DFMA as one lane op (so 118 + 58.8 = 177 if both pipes fill)
  rep("coissue_f16_d4", run(k_coissue<16, 4>, blocks, threads, out, cyc), (double)ITER * (16 * 2 + 4), "fma_plus_dfma");
  rep("coissue_f16_d8", run(k_coissue<16, 8>, blocks, threads, out, cyc), (double)ITER * (16 * 2 + 8), "fma_plus_dfma");
  rep("coissue_f16_d16", run(k_coissue<16, 16>, blocks, threads, out, cyc), (double)ITER * (16 * 2 + 16), "fma_plus_dfma");
  rep("coissue_f8_d16", run(k_coissue<8, 16>, blocks, threads, out, cyc), (double)ITER * (8 * 2 + 16), "fma_plus_dfma");
  // hybrid rotate+accumulate: terms per thread averaged over the two roles (fp32 warp 32 per source, fp64 warp 16)
  rep("hybrid_allfp32_terms", run(k_hybrid<0, 0, 4>, blocks, threads, out, cyc), (double)(ITER / 16) * 32, "terms");
  rep("hybrid_allfp32_lds_terms", run(k_hybrid<1, 0, 4>, blocks, threads, out, cyc), (double)(ITER / 16) * 32, "terms");
  rep("hybrid_2to1_ns1_terms", run(k_hybrid<0, 1, 1>, blocks, threads, out, cyc), (double)(ITER / 16) * 24, "terms");
  rep("hybrid_2to1_ns2_terms", run(k_hybrid<0, 1, 2>, blocks, threads, out, cyc), (double)(ITER / 16) * 24, "terms");
  rep("hybrid_2to1_ns4_terms", run(k_hybrid<0, 1, 4>, blocks, threads, out, cyc), (double)(ITER / 16) * 24, "terms");
  rep("hybrid_2to1_ns4_lds_cvt_terms", run(k_hybrid<1, 1, 4>, blocks, threads, out, cyc), (double)(ITER / 16) * 24, "terms");
  {
    // latency: cycles per dependent instruction (128 threads = 1 warp per scheduler, 1 CTA per SM)
    const char* nm[3] = {"latency_ffma", "latency_ffma2", "latency_dfma"};
    Res r0 = run(k_latency<0>, sms, 128, out, cyc), r1 = run(k_latency<1>, sms, 128, out, cyc), r2 = run(k_latency<2>, sms, 128, out, cyc);
    double c[3] = {r0.cyc, r1.cyc, r2.cyc};
    for (int i = 0; i < 3; ++i) printf("  \"%s\": {\"cycles_per_dependent_op\": %.2f}%s\n", nm[i], c[i] / (ITER * 16.0), i == 2 ? "" : ",");
  }
  printf("}\n");
  return 0;
}
#endif  // PB200_MICROBENCH_MAIN

// Library entry: the three pipe rates bench.py quotes as roofline denominators.
extern "C" int pb200_microbench(pb200_ctx* ctx, double* out, int n) {
  if (!ctx || !out || n < 5) return ctx ? pb_fail(ctx, PB200_EINVAL, "pb200_microbench: need out[5]") : PB200_EINVAL;
  PbDeviceGuard guard(ctx->device);
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
