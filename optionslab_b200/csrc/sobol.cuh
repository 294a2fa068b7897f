// Quasi-Monte Carlo European kernel: scrambled Sobol points generated on the device from a table of
// direction numbers, inverse normal CDF in registers, terminal GBM, payoff, (sum, sum^2).
//
// Replaces src/simulation/gbm_qmc.py:14-47 (Sobol(d=n_steps, scramble=True, seed).random(N) ->
// norm.ppf(clip(u, 1e-10, 1-1e-10)) -> ln S + drift*n + vol*sum(z) -> exp), the MCMethod.QMC backend of
// src/pricing_models/monte_carlo.py:94-97.  The (N, n_steps) uniform / normal arrays of the reference
// (2 x 2 GB at 1M x 252) are never materialised.
//
// Point set.  A digital net is linear over GF(2): coordinate j of point i is
//     x_j(i) = shift_j  XOR  (XOR over the set bits b of i)  c_j[b]
// with c_j[b] = v_j[b] ^ v_j[b-1] when the generator enumerates in Gray-code order with direction numbers
// v_j[b] (scipy's `Sobol._sv`, Owen-style linear matrix scramble already folded in) and digital shift
// shift_j (`Sobol._shift`).  The host passes c (natural order) and shift; the integers produced here are
// bit-identical to scipy's (tests/test_gpu_qmc.py).  u = x * 2^-bits.
//
// Work split.  Point index bits: [0,4) the 16 points a thread owns, [4,12) threadIdx, [12,...) the CTA.
// Dimensions are the OUTER loop (staged through shared memory 64 at a time), the thread's 16 running
// sums of normals live in registers:
//   per (CTA, dim)     : the CTA-bit contribution, computed once by one thread            (xcta_s)
//   per (thread, dim)  : 8 masked XORs for the threadIdx bits                             (8 LOP3)
//   per (point, dim)   : 1 XOR (the 16 points are visited in Gray order), inverse normal, 1 FADD.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "mc_kernels.cuh"

namespace b200mc {

constexpr int kSobolPoints = 16;      // points per thread
constexpr int kSobolPointBits = 4;
constexpr int kSobolTidBits = 8;      // kBlock == 256
constexpr int kSobolCtaShift = kSobolPointBits + kSobolTidBits;
constexpr int kSobolDimChunk = 64;    // dimensions staged in shared memory at a time
constexpr int kSobolWords = 32;       // table words per dimension (bit b of the point index -> word b)
static_assert(kBlock == 1 << kSobolTidBits, "thread-bit split assumes 256 threads");

struct SobolArgs {
  const b200mc_params_t* params;  // [n_opt][n_scen]
  double* partials;               // [n_opt * tiles][2 * NS]
  const uint32_t* dirnums;        // [n_steps][kSobolWords], natural (binary) order
  const uint32_t* shift;          // [n_steps]
  uint64_t point_begin;           // multiple of kBlock * kSobolPoints
  uint64_t n_points;
  uint32_t n_opt, n_scen, tiles, n_steps, bits;
  int32_t is_put;
};

// Phi^-1 in FP32.  t = min(u, 1-u) in [1e-10, 1/2], w = -ln(4t(1-t)), |x| = 1 - 2t:
//   Phi^-1(u) = sign(u - 1/2) * |x| * P(w),  P = sqrt(2)*erfinv(x)/x as a polynomial in (w - 2.5) for w < 5
//   and in (sqrt(w) - 3) for w >= 5 (the form of M. Giles' single-precision erfinv; coefficients re-fitted
//   in FP64 by tools/fit_inverse_normal.py: max abs error 2.5e-6 at |z| = 6, 6.4e-7 relative).
__device__ __forceinline__ float inverse_normal_central(float v) {  // v = w - 2.5
  float p = 3.958320986e-08f;
  p = fmaf(p, v, 4.851791015e-07f);
  p = fmaf(p, v, -4.981194608e-06f);
  p = fmaf(p, v, -6.207880342e-06f);
  p = fmaf(p, v, 3.091150937e-04f);
  p = fmaf(p, v, -1.773041744e-03f);
  p = fmaf(p, v, -5.908129905e-03f);
  p = fmaf(p, v, 3.488026846e-01f);
  p = fmaf(p, v, 2.123313560e+00f);
  return p;
}
__device__ __forceinline__ float inverse_normal_tail(float w) {
  const float v = mufu_sqrt(w) - 3.0f;
  float p = -7.605044245e-06f;
  p = fmaf(p, v, -2.237582165e-04f);
  p = fmaf(p, v, 1.646633081e-03f);
  p = fmaf(p, v, -4.775961266e-03f);
  p = fmaf(p, v, 8.189511418e-03f);
  p = fmaf(p, v, -1.092221667e-02f);
  p = fmaf(p, v, 1.334085408e-02f);
  p = fmaf(p, v, 1.416593595e+00f);
  p = fmaf(p, v, 4.006434678e+00f);
  return p;
}

// x in [0, 2^bits): the Sobol integer.  u = x * 2^-bits clipped to [1e-10, 1 - 1e-10] (gbm_qmc.py:36).
__device__ __forceinline__ float inverse_normal_from_sobol(uint32_t x, uint32_t one, float scale) {
  const uint32_t xr = one - x;
  const bool lower = x < xr;                                  // u < 1/2
  const float t = fmaxf((float)(lower ? x : xr) * scale, 1e-10f);
  const float v = fmaf(mufu_lg2(fmaf(-t, t, t)), -0.69314718055994530942f, -3.88629436111989061883f);  // w - 2.5, w = -ln(4t(1-t))
  const float ax = fmaf(-2.0f, t, 1.0f);
  float p = inverse_normal_central(v);
  if (v >= 2.5f) p = inverse_normal_tail(v + 2.5f);           // w >= 5: 0.67% of draws
  const float z = p * ax;
  return lower ? -z : z;
}

struct QmcCoef {
  float c;      // sigma*sqrt(dt) / ln2 : log2-diffusion per unit normal
  float a;      // (r - q - sigma^2/2) T / ln2 : terminal log2-drift
  float kappa;  // K / S
};

template <int NS>
__global__ void __launch_bounds__(kBlock, NS <= 2 ? 4 : 2) qmc_european_kernel(const SobolArgs a) {
  __shared__ __align__(16) uint32_t c_s[kSobolDimChunk][kSobolWords];
  __shared__ uint32_t xcta_s[kSobolDimChunk];
  __shared__ QmcCoef coef[NS];
  const uint32_t opt = blockIdx.x / a.tiles;
  const uint32_t tile = blockIdx.x - opt * a.tiles;
  if (threadIdx.x < NS) {
    const uint32_t k = threadIdx.x < a.n_scen ? threadIdx.x : a.n_scen - 1;
    const b200mc_params_t p = a.params[(size_t)opt * a.n_scen + k];
    const double inv_ln2 = 1.44269504088896340736;
    const double dt = p.T / (double)a.n_steps;
    QmcCoef q;
    q.c = (float)(p.sigma * sqrt(dt) * inv_ln2);
    q.a = (float)((p.r - p.q - 0.5 * p.sigma * p.sigma) * dt * inv_ln2 * (double)a.n_steps);
    q.kappa = (float)(p.K / p.S);
    coef[threadIdx.x] = q;
  }

  const uint32_t cta_index = (uint32_t)(a.point_begin >> kSobolCtaShift) + tile;  // point-index bits [12, bits)
  uint32_t tid_mask[kSobolTidBits];
#pragma unroll
  for (int b = 0; b < kSobolTidBits; ++b) tid_mask[b] = 0u - ((threadIdx.x >> b) & 1u);
  const uint32_t one = 1u << a.bits;
  const float scale = 1.0f / (float)one;

  float W[kSobolPoints];
#pragma unroll
  for (int k = 0; k < kSobolPoints; ++k) W[k] = 0.0f;

  for (uint32_t d0 = 0; d0 < a.n_steps; d0 += kSobolDimChunk) {
    const uint32_t nd = min((uint32_t)kSobolDimChunk, a.n_steps - d0);
    __syncthreads();  // the previous chunk is fully consumed
    for (uint32_t i = threadIdx.x; i < nd * kSobolWords; i += kBlock) (&c_s[0][0])[i] = a.dirnums[(size_t)d0 * kSobolWords + i];
    __syncthreads();
    if (threadIdx.x < nd) {
      uint32_t x = a.shift[d0 + threadIdx.x];
      for (uint32_t ci = cta_index, b = kSobolCtaShift; ci != 0; ci >>= 1, ++b)
        if (ci & 1u) x ^= c_s[threadIdx.x][b];
      xcta_s[threadIdx.x] = x;
    }
    __syncthreads();
    for (uint32_t j = 0; j < nd; ++j) {
      const uint4 lo = *reinterpret_cast<const uint4*>(&c_s[j][0]);
      const uint4 t0 = *reinterpret_cast<const uint4*>(&c_s[j][kSobolPointBits]);
      const uint4 t1 = *reinterpret_cast<const uint4*>(&c_s[j][kSobolPointBits + 4]);
      uint32_t x = xcta_s[j];
      x ^= (t0.x & tid_mask[0]) ^ (t0.y & tid_mask[1]);
      x ^= (t0.z & tid_mask[2]) ^ (t0.w & tid_mask[3]);
      x ^= (t1.x & tid_mask[4]) ^ (t1.y & tid_mask[5]);
      x ^= (t1.z & tid_mask[6]) ^ (t1.w & tid_mask[7]);
      const uint32_t flip[kSobolPointBits] = {lo.x, lo.y, lo.z, lo.w};
#pragma unroll
      for (int g = 0; g < kSobolPoints; ++g) {  // local point bits visited in Gray order: one XOR per point
        if (g > 0) x ^= flip[(g & 1) ? 0 : (g & 2) ? 1 : (g & 4) ? 2 : 3];  // lowest set bit of g
        W[g ^ (g >> 1)] += inverse_normal_from_sobol(x, one, scale);
      }
    }
  }

  float acc[2 * NS];
#pragma unroll
  for (int i = 0; i < 2 * NS; ++i) acc[i] = 0.0f;
  const bool is_put = a.is_put != 0;
  const uint64_t first = ((uint64_t)tile * kBlock + threadIdx.x) * kSobolPoints;  // local index of this thread's point 0
#pragma unroll
  for (int k = 0; k < kSobolPoints; ++k) {
    if (first + k < a.n_points) {
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        const QmcCoef q = coef[s];
        const float p = vanilla(mufu_ex2(fmaf(q.c, W[k], q.a)), q.kappa, is_put);
        acc[2 * s] += p;
        acc[2 * s + 1] = fmaf(p, p, acc[2 * s + 1]);
      }
    }
  }
  block_reduce_store<2 * NS>(acc, a.partials + (size_t)blockIdx.x * (2 * NS));
}

// Inspection: the Sobol integers of points [point_begin, point_begin + n_points) x n_dims, row-major.
__global__ void sobol_points_kernel(const uint32_t* __restrict__ dirnums, const uint32_t* __restrict__ shift, uint64_t point_begin,
                                    uint64_t n_points, uint32_t n_dims, uint32_t* __restrict__ out) {
  const uint64_t total = n_points * n_dims;
  for (uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t i = point_begin + idx / n_dims;
    const uint32_t j = (uint32_t)(idx % n_dims);
    uint32_t x = shift[j];
    for (uint32_t b = 0, v = (uint32_t)i; v != 0; v >>= 1, ++b)
      if (v & 1u) x ^= dirnums[(size_t)j * kSobolWords + b];
    out[idx] = x;
  }
}

// Inspection: the FP32 normals the QMC kernel derives from given Sobol integers.
__global__ void sobol_normals_kernel(const uint32_t* __restrict__ x, uint64_t n, uint32_t bits, float* __restrict__ out) {
  const uint32_t one = 1u << bits;
  const float scale = 1.0f / (float)one;
  for (uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (uint64_t)gridDim.x * blockDim.x)
    out[idx] = inverse_normal_from_sobol(x[idx], one, scale);
}

}  // namespace b200mc
