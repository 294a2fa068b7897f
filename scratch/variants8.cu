// Scratch experiment 8: one word = one Box-Muller pair (23-bit radius + 9-bit angle): 8 normals per Philox call.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#include "../optionslab_b200/csrc/mc_kernels.cuh"
using namespace b200mc;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
template <int ABITS> __device__ __forceinline__ NormalPair bm_word(uint32_t w) {
  NormalPair p;
  const float u = 2.0f - word_to_unit_1_2(w);
  p.rad = mufu_sqrt(-mufu_lg2(u));
  const float g = __uint_as_float((w & ((1u << ABITS) - 1u)) | 0x4b000000u);
  const float step = 6.28318530717958647692f / (float)(1u << ABITS);
  const float theta = fmaf(g, step, -(8388608.0f * step + 3.14159265358979f));
  p.cs = mufu_cos(theta); p.sn = mufu_sin(theta);
  return p;
}
struct Euro { float W = 0.f; __device__ __forceinline__ void operator()(const NormalPair& p) { W = fmaf(p.rad, p.cs, W); W = fmaf(p.rad, p.sn, W); }
  __device__ __forceinline__ float result() const { return W; } };
struct Asian { float l = 0.f, sum = 0.f, c, d;
  __device__ __forceinline__ void operator()(const NormalPair& p) { const float rc = p.rad * c; l = fmaf(rc, p.cs, l + d); sum += mufu_ex2(l); l = fmaf(rc, p.sn, l + d); sum += mufu_ex2(l); }
  __device__ __forceinline__ float result() const { return sum; } };
struct Barrier { float l = 0.f, m = 0.f, c, d;
  __device__ __forceinline__ void operator()(const NormalPair& p) { const float rc = p.rad * c; l = fmaf(rc, p.cs, l + d); m = fmaxf(m, l); l = fmaf(rc, p.sn, l + d); m = fmaxf(m, l); }
  __device__ __forceinline__ float result() const { return m + l; } };
template <class C> __device__ __forceinline__ C make(float c, float d) { C a; return a; }
template <> __device__ __forceinline__ Asian make<Asian>(float c, float d) { Asian a; a.c = c; a.d = d; return a; }
template <> __device__ __forceinline__ Barrier make<Barrier>(float c, float d) { Barrier a; a.c = c; a.d = d; return a; }

template <class C, int UNROLL, int MINB, int ABITS>
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
      for (int u = 0; u < UNROLL; ++u) { cons(bm_word<ABITS>(x[u].x)); cons(bm_word<ABITS>(x[u].y)); cons(bm_word<ABITS>(x[u].z)); cons(bm_word<ABITS>(x[u].w)); }
    }
    acc += cons.result();
  }
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
#define RUN(C, U, B) { float ms = time_ms([&] { k<C, U, B, 9><<<grid, 256>>>(ppt, NS, 42u, 0u, 0.018f, 1e-4f, out); }); \
    printf(#C " unroll=%d minb=%d  %9.3f ms  %.4e /s  (%.3f per clk per SM)\n", U, B, ms, steps / (ms * 1e-3), steps / (ms * 1e-3) / (sms * 1.965e9)); }
  RUN(Euro, 1, 1); RUN(Euro, 2, 1); RUN(Euro, 4, 1); RUN(Euro, 1, 6); RUN(Euro, 2, 6); RUN(Euro, 2, 4); RUN(Euro, 1, 8); RUN(Euro, 2, 5);
  RUN(Asian, 1, 1); RUN(Asian, 2, 1); RUN(Asian, 2, 4); RUN(Asian, 1, 6); RUN(Asian, 2, 6);
  RUN(Barrier, 1, 1); RUN(Barrier, 2, 1); RUN(Barrier, 2, 4); RUN(Barrier, 1, 6); RUN(Barrier, 2, 6);
  return 0;
}
