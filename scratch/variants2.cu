// Scratch experiment 2 (not product): loop structures x consumers.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../optionslab_b200/csrc/mc_kernels.cuh"
using namespace b200mc;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

struct Words12 { uint32_t w[12]; };
__device__ __forceinline__ Words12 draw12(uint64_t path, uint32_t sb, uint32_t k0, uint32_t k1) {
  const u32x4 a = draw4(path, 3 * sb, 0u, k0, k1), b = draw4(path, 3 * sb + 1, 0u, k0, k1), c = draw4(path, 3 * sb + 2, 0u, k0, k1);
  return Words12{{a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w}};
}
template <class F> __device__ __forceinline__ void consume12(const Words12& x, F&& f) {
#pragma unroll
  for (int t = 0; t < 4; ++t) { NormalPair A, B; box_muller_quad(x.w[3 * t], x.w[3 * t + 1], x.w[3 * t + 2], A, B); f(A); f(B); }
}

// consumers ---------------------------------------------------------------------------------
struct Euro { float W = 0.f; __device__ __forceinline__ void operator()(const NormalPair& p) { W = fmaf(p.rad, p.cs, W); W = fmaf(p.rad, p.sn, W); }
  __device__ __forceinline__ float result() const { return W; } };
struct Asian { float l = 0.f, sum = 0.f, c, d;
  __device__ __forceinline__ void operator()(const NormalPair& p) { const float rc = p.rad * c; l = fmaf(rc, p.cs, l + d); sum += mufu_ex2(l); l = fmaf(rc, p.sn, l + d); sum += mufu_ex2(l); }
  __device__ __forceinline__ float result() const { return sum; } };
struct Barrier { float l = 0.f, m = 0.f, c, d;
  __device__ __forceinline__ void operator()(const NormalPair& p) { const float rc = p.rad * c; l = fmaf(rc, p.cs, l + d); m = fmaxf(m, l); l = fmaf(rc, p.sn, l + d); m = fmaxf(m, l); }
  __device__ __forceinline__ float result() const { return m + l; } };
template <class C> __device__ __forceinline__ C make(float c, float d);
template <> __device__ __forceinline__ Euro make<Euro>(float, float) { return Euro{}; }
template <> __device__ __forceinline__ Asian make<Asian>(float c, float d) { Asian a; a.c = c; a.d = d; return a; }
template <> __device__ __forceinline__ Barrier make<Barrier>(float c, float d) { Barrier a; a.c = c; a.d = d; return a; }

// MODE 0: production for_each_pair; 1: draw12 then consume (simple); 2: software pipelined
template <class C, int MODE, int MINB>
__global__ void __launch_bounds__(256, MINB) k(uint32_t ppt, uint32_t n_steps, uint32_t k0, uint32_t k1, float c, float d, float* out) {
  const uint64_t base = ((uint64_t)blockIdx.x * ppt) * 256 + threadIdx.x;
  float acc = 0.f;
  for (uint32_t j = 0; j < ppt; ++j) {
    const uint64_t path = base + (uint64_t)j * 256;
    C cons = make<C>(c, d);
    const uint32_t full = n_steps >> 4;
    if (MODE == 0) {
      for_each_pair(path, n_steps, 0u, k0, k1, [&](const NormalPair& p, int) { cons(p); });
    } else if (MODE == 1) {
      for (uint32_t sb = 0; sb < full; ++sb) { const Words12 x = draw12(path, sb, k0, k1); consume12(x, cons); }
    } else {
      Words12 cur = draw12(path, 0, k0, k1);
      for (uint32_t sb = 1; sb < full; ++sb) { const Words12 nxt = draw12(path, sb, k0, k1); consume12(cur, cons); cur = nxt; }
      consume12(cur, cons);
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
  const uint32_t n_steps = 256, ppt = 32, grid = sms * 8 * 16;
  float* out; CK(cudaMalloc(&out, (size_t)grid * 256 * sizeof(float)));
  const double steps = (double)grid * 256 * ppt * n_steps;
  auto report = [&](const char* name, float ms) { printf("%-40s %9.3f ms  %.4e /s  (%.3f per clk per SM)\n", name, ms, steps / (ms * 1e-3), steps / (ms * 1e-3) / (sms * 1.965e9)); };
#define RUN(C, M, B) report(#C " mode=" #M " minb=" #B, time_ms([&] { k<C, M, B><<<grid, 256>>>(ppt, n_steps, 42u, 0u, 0.018f, 1e-4f, out); }))
  RUN(Euro, 0, 1); RUN(Euro, 1, 1); RUN(Euro, 2, 1); RUN(Euro, 2, 4); RUN(Euro, 2, 5); RUN(Euro, 1, 5);
  RUN(Asian, 0, 1); RUN(Asian, 1, 1); RUN(Asian, 2, 1); RUN(Asian, 2, 4);
  RUN(Barrier, 0, 1); RUN(Barrier, 1, 1); RUN(Barrier, 2, 1); RUN(Barrier, 2, 4);
  return 0;
}
