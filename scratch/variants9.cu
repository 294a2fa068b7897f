// Scratch experiment 9: one word = one Box-Muller pair, "sheared" bit use (radius mantissa = low 23 bits, angle
// mantissa = top 23 bits: the 9 bits the radius does not see are the angle's leading bits), vs the masked 9-bit
// angle of experiment 8; Asian with a polynomial 2^x - 1 instead of MUFU.EX2.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#include "../optionslab_b200/csrc/mc_kernels.cuh"
using namespace b200mc;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int LAYOUT> __device__ __forceinline__ NormalPair bm_word(uint32_t w) {
  NormalPair p;
  if (LAYOUT == 0) {
    const float u = 2.0f - word_to_unit_1_2(w);
    p.rad = mufu_sqrt(-mufu_lg2(u));
    const float g = __uint_as_float((w & 511u) | 0x4b000000u);
    const float step = 6.28318530717958647692f / 512.0f;
    const float theta = fmaf(g, step, -(8388608.0f * step + 3.14159265358979f));
    p.cs = mufu_cos(theta); p.sn = mufu_sin(theta);
  } else {
    const float u = 2.0f - __uint_as_float((w & 0x7fffffu) | 0x3f800000u);
    p.rad = mufu_sqrt(-mufu_lg2(u));
    const float f = __uint_as_float((w >> 9) | 0x3f800000u);                  // [1,2) turns
    const float theta = fmaf(f, 6.28318530717958647692f, -9.42477796076937971538f);  // [-pi, pi)
    p.cs = mufu_cos(theta); p.sn = mufu_sin(theta);
  }
  return p;
}
struct Euro { float W = 0.f; __device__ __forceinline__ void operator()(const NormalPair& p) { W = fmaf(p.rad, p.cs, W); W = fmaf(p.rad, p.sn, W); }
  __device__ __forceinline__ float result() const { return W; } };
struct Asian { float l = 0.f, sum = 0.f, c, d;
  __device__ __forceinline__ void operator()(const NormalPair& p) { const float rc = p.rad * c; l = fmaf(rc, p.cs, l + d); sum += mufu_ex2(l); l = fmaf(rc, p.sn, l + d); sum += mufu_ex2(l); }
  __device__ __forceinline__ float result() const { return sum; } };
// 2^x - 1 on |x| <= 0.25: degree-5 Taylor in x*ln2 (Horner); S <- S + S*p keeps the relative rounding at 1 ulp per step
template <int DEG> __device__ __forceinline__ float exp2m1_poly(float x) {
  float p = 9.6181291076284771619e-3f;
  if (DEG >= 5) p = fmaf(1.3333558146428443423e-3f, x, 9.6181291076284771619e-3f);
  p = fmaf(p, x, 5.5504108664821579953e-2f);
  p = fmaf(p, x, 2.4022650695910071233e-1f);
  p = fmaf(p, x, 6.9314718055994530942e-1f);
  return p * x;
}
template <int DEG> struct AsianPolyT { float S = 1.f, sum = 0.f, c, d;
  __device__ __forceinline__ void operator()(const NormalPair& p) { const float rc = p.rad * c;
    S = fmaf(S, exp2m1_poly<DEG>(fmaf(rc, p.cs, d)), S); sum += S; S = fmaf(S, exp2m1_poly<DEG>(fmaf(rc, p.sn, d)), S); sum += S; }
  __device__ __forceinline__ float result() const { return sum; } };
struct Barrier { float l = 0.f, m = 0.f, c, d;
  __device__ __forceinline__ void operator()(const NormalPair& p) { const float rc = p.rad * c; l = fmaf(rc, p.cs, l + d); m = fmaxf(m, l); l = fmaf(rc, p.sn, l + d); m = fmaxf(m, l); }
  __device__ __forceinline__ float result() const { return m + l; } };
template <class C> __device__ __forceinline__ C make(float c, float d) { C a; return a; }
template <> __device__ __forceinline__ Asian make<Asian>(float c, float d) { Asian a; a.c = c; a.d = d; return a; }
using AsianPoly = AsianPolyT<5>; using AsianPoly4 = AsianPolyT<4>;
template <> __device__ __forceinline__ AsianPoly make<AsianPoly>(float c, float d) { AsianPoly a; a.c = c; a.d = d; return a; }
template <> __device__ __forceinline__ AsianPoly4 make<AsianPoly4>(float c, float d) { AsianPoly4 a; a.c = c; a.d = d; return a; }
template <> __device__ __forceinline__ Barrier make<Barrier>(float c, float d) { Barrier a; a.c = c; a.d = d; return a; }

template <class C, int UNROLL, int MINB, int LAYOUT>
__global__ void __launch_bounds__(256, MINB) k(uint32_t ppt, uint32_t n_steps, uint32_t k0, uint32_t k1, float c, float d, float* out) {
  const uint64_t base = ((uint64_t)blockIdx.x * ppt) * 256 + threadIdx.x;
  float acc = 0.f;
  for (uint32_t j = 0; j < ppt; ++j) {
    const uint64_t path = base + (uint64_t)j * 256;
    C cons = make<C>(c, d);
    const uint32_t calls = n_steps / 8;
    for (uint32_t cc = 0; cc < calls; cc += UNROLL) {
      u32x4 x[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) x[u] = draw4(path, cc + u, 0u, k0, k1);
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) { cons(bm_word<LAYOUT>(x[u].x)); cons(bm_word<LAYOUT>(x[u].y)); cons(bm_word<LAYOUT>(x[u].z)); cons(bm_word<LAYOUT>(x[u].w)); }
    }
    acc += cons.result();
  }
  out[blockIdx.x * 256 + threadIdx.x] = acc;
}
// production layout for comparison (3 calls per 16 steps)
template <int MINB>
__global__ void __launch_bounds__(256, MINB) kprod(uint32_t ppt, uint32_t n_steps, uint32_t k0, uint32_t k1, float* out) {
  const uint64_t base = ((uint64_t)blockIdx.x * ppt) * 256 + threadIdx.x;
  float acc = 0.f;
  for (uint32_t j = 0; j < ppt; ++j) acc += terminal_sum(base + (uint64_t)j * 256, n_steps, 0u, k0, k1);
  out[blockIdx.x * 256 + threadIdx.x] = acc;
}
template <class L> float time_ms(L&& launch) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  launch(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < 3; ++r) { CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); best = ms < best ? ms : best; }
  CK(cudaGetLastError());
  return best;
}
int main() {
  int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const uint32_t ppt = 32, grid = sms * 8 * 16, NS = 256;
  float* out; CK(cudaMalloc(&out, (size_t)grid * 256 * sizeof(float)));
  const double steps = (double)grid * 256 * ppt * NS;
#define REPORT(name, ms) printf("%-44s %9.3f ms  %.4e /s  (%.3f per clk per SM)\n", name, ms, steps / (ms * 1e-3), steps / (ms * 1e-3) / (sms * 1.965e9));
#define RUN(C, U, B, L) { float ms = time_ms([&] { k<C, U, B, L><<<grid, 256>>>(ppt, NS, 42u, 0u, 0.018f, 1e-4f, out); }); \
    REPORT(#C " unroll=" #U " minb=" #B " layout=" #L, ms); }
  { float ms = time_ms([&] { kprod<1><<<grid, 256>>>(ppt, NS, 42u, 0u, out); }); REPORT("production terminal_sum minb=1", ms); }
  { float ms = time_ms([&] { kprod<4><<<grid, 256>>>(ppt, NS, 42u, 0u, out); }); REPORT("production terminal_sum minb=4", ms); }
  RUN(Euro, 1, 1, 0); RUN(Euro, 2, 1, 0); RUN(Euro, 2, 4, 0);
  RUN(Euro, 1, 1, 1); RUN(Euro, 2, 1, 1); RUN(Euro, 4, 1, 1); RUN(Euro, 1, 4, 1); RUN(Euro, 2, 4, 1); RUN(Euro, 1, 6, 1); RUN(Euro, 2, 6, 1); RUN(Euro, 1, 8, 1); RUN(Euro, 2, 8, 1);
  RUN(Asian, 1, 1, 1); RUN(Asian, 2, 1, 1); RUN(Asian, 2, 4, 1); RUN(Asian, 1, 6, 1); RUN(Asian, 2, 6, 1);
  RUN(AsianPoly, 1, 1, 1); RUN(AsianPoly, 2, 1, 1); RUN(AsianPoly, 2, 4, 1); RUN(AsianPoly, 1, 6, 1); RUN(AsianPoly, 2, 6, 1); RUN(AsianPoly, 1, 8, 1);
  RUN(AsianPoly4, 1, 1, 1); RUN(AsianPoly4, 2, 1, 1); RUN(AsianPoly4, 2, 4, 1); RUN(AsianPoly4, 1, 6, 1);
  RUN(Barrier, 1, 1, 1); RUN(Barrier, 2, 1, 1); RUN(Barrier, 2, 4, 1); RUN(Barrier, 1, 6, 1); RUN(Barrier, 2, 6, 1);
  return 0;
}
