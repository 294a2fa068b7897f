// Scratch experiment 5: Philox multiply as IMAD.WIDE vs separate mul.hi/mul.lo; micro rates of IMAD / IMAD.HI.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../optionslab_b200/csrc/normal.cuh"
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
using namespace b200mc;
struct u4 { uint32_t x, y, z, w; };
template <int SPLIT> __device__ __forceinline__ void mhl(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
  if (SPLIT) { asm("mul.hi.u32 %0, %1, %2;" : "=r"(hi) : "r"(a), "r"(b)); asm("mul.lo.u32 %0, %1, %2;" : "=r"(lo) : "r"(a), "r"(b)); }
  else { const uint64_t p = (uint64_t)a * b; hi = (uint32_t)(p >> 32); lo = (uint32_t)p; }
}
template <int SPLIT> __device__ __forceinline__ u4 philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0, lo0, hi1, lo1;
    mhl<SPLIT>(0xD2511F53u, c0, hi0, lo0); mhl<SPLIT>(0xCD9E8D57u, c2, hi1, lo1);
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0; k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return u4{c0, c1, c2, c3};
}
template <int SPLIT> __global__ void __launch_bounds__(256, 4) k_euro(uint32_t ppt, uint32_t n_sb, uint32_t k0, uint32_t k1, float* out) {
  const uint32_t base = (blockIdx.x * ppt) * 256 + threadIdx.x;
  float acc = 0.f;
  for (uint32_t j = 0; j < ppt; ++j) {
    const uint32_t path = base + j * 256;
    float W = 0.f;
    for (uint32_t sb = 0; sb < n_sb; ++sb) {
      const u4 a = philox<SPLIT>(path, 3 * sb, 0u, 0u, k0, k1), b = philox<SPLIT>(path, 3 * sb + 1, 0u, 0u, k0, k1), c = philox<SPLIT>(path, 3 * sb + 2, 0u, 0u, k0, k1);
      const uint32_t w[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
#pragma unroll
      for (int t = 0; t < 4; ++t) { NormalPair A, B; box_muller_quad(w[3 * t], w[3 * t + 1], w[3 * t + 2], A, B);
        W = fmaf(A.rad, A.cs, W); W = fmaf(A.rad, A.sn, W); W = fmaf(B.rad, B.cs, W); W = fmaf(B.rad, B.sn, W); }
    }
    acc += W;
  }
  out[blockIdx.x * 256 + threadIdx.x] = acc;
}
template <int WHICH> __global__ void k_mul(uint32_t iters, uint32_t m, uint32_t* out) {
  uint32_t x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 2654435761u + i;
  for (uint32_t it = 0; it < iters; ++it)
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (WHICH == 0) asm volatile("mul.lo.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(m));
        if (WHICH == 1) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(m));
        if (WHICH == 2) { uint64_t p; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(x[i]), "r"(m)); x[i] = (uint32_t)p ^ (uint32_t)(p >> 32); }
      }
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s ^= x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
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
  const uint32_t ppt = 32, n_sb = 16, grid = sms * 8 * 16;
  float* out; CK(cudaMalloc(&out, (size_t)grid * 256 * 4));
  const double steps = (double)grid * 256 * ppt * n_sb * 16;
  float ms = time_ms([&] { k_euro<0><<<grid, 256>>>(ppt, n_sb, 42u, 0u, out); });
  printf("euro imad.wide   %9.3f ms %.4e /s (%.3f per clk per SM)\n", ms, steps / (ms * 1e-3), steps / (ms * 1e-3) / (sms * 1.965e9));
  ms = time_ms([&] { k_euro<1><<<grid, 256>>>(ppt, n_sb, 42u, 0u, out); });
  printf("euro hi+lo split %9.3f ms %.4e /s (%.3f per clk per SM)\n", ms, steps / (ms * 1e-3), steps / (ms * 1e-3) / (sms * 1.965e9));
  const uint32_t it = 4096; const double ops = (double)sms * 8 * 256 * it * 32;
  ms = time_ms([&] { k_mul<0><<<sms * 8, 256>>>(it, 0xD2511F53u, (uint32_t*)out); }); printf("mul.lo   %.3f per clk per SM\n", ops / (ms * 1e-3) / (sms * 1.965e9));
  ms = time_ms([&] { k_mul<1><<<sms * 8, 256>>>(it, 0xD2511F53u, (uint32_t*)out); }); printf("mul.hi   %.3f per clk per SM\n", ops / (ms * 1e-3) / (sms * 1.965e9));
  ms = time_ms([&] { k_mul<2><<<sms * 8, 256>>>(it, 0xD2511F53u, (uint32_t*)out); }); printf("mul.wide(+xor) %.3f per clk per SM\n", ops / (ms * 1e-3) / (sms * 1.965e9));
  return 0;
}
