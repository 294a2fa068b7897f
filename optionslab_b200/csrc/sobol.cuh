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
// Work split.  Point index bits: [0,PB) the 2^PB points a thread owns, [PB,PB+8) threadIdx, above that the CTA.
// PB = 4 (16 points per thread, 4096 per CTA) amortises the per-thread work best; point sets too small to give
// every SM a CTA that way take PB = 2 or 0 (2^16 points: 16 CTAs at PB = 4, 256 at PB = 0).
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

constexpr int kSobolMaxPointBits = 4;  // up to 16 points per thread
constexpr int kSobolTidBits = 8;       // kBlock == 256
constexpr int kSobolAlignShift = kSobolMaxPointBits + kSobolTidBits;  // ranks split the sequence at multiples of 4096
constexpr int kSobolDimChunk = 64;    // dimensions staged in shared memory at a time
constexpr int kSobolWords = 32;       // table words per dimension (bit b of the point index -> word b)
static_assert(kBlock == 1 << kSobolTidBits, "thread-bit split assumes 256 threads");

struct SobolArgs {
  const b200mc_params_t* params;  // [n_opt][n_scen]
  FoldArgs fold;                  // tile partials -> out[n_opt][n_scen]
  const uint32_t* dirnums;        // [n_steps][kSobolWords], natural (binary) order
  const uint32_t* shift;          // [n_steps]
  uint64_t point_begin;           // multiple of 1 << kSobolAlignShift
  uint64_t n_points;
  uint32_t n_opt, n_scen, tiles, n_steps, bits;
  int32_t is_put;
  double* terminal_out;           // non-null: also write S_T per point ([n_points], then the mirrored [n_points] when terminal_anti)
  int32_t terminal_anti;
  // dimension split: point sets too small to fill the chip give each point tile to `dsplit` CTAs, each summing the normals of
  // its share of the dimensions; the tile's last CTA adds the partial sums (in part order) and prices the payoff
  uint32_t dsplit;
  float* wpart;                   // [n_opt * tiles][dsplit][points per tile]
  uint32_t* tile_tickets;         // [n_opt * tiles], zero between launches
};

// Phi^-1 in FP32, branch-free.  t = min(u, 1-u) in [1e-10, 1/2], y = sqrt(-2 ln t) in [1.1774, 6.7861]:
//   Phi^-1(u) = sign(u - 1/2) * P((y - c)/h),  P of degree 14 fitted in FP64 against scipy's ndtri by
//   tools/fit_inverse_normal.py (max abs error of the FP32 Horner form: 1.3e-6, at |z| = 6).  2 MUFU + 15 FFMA.
//   A piecewise central/tail form (Giles) is more accurate in relative terms near z = 0 but its tail branch
//   diverges in 19% of the warps: measured 25% slower (profiles/r01_variants13_inverse_normal.txt).
// xs = the Sobol integer scaled to 31 bits (u = xs * 2^-31; the table is staged pre-shifted, so the scaling is free).
// t = min(u, 1 - u) clipped to [1e-10, 1/2] (gbm_qmc.py:36), y = sqrt(-2 ln t) and the half u falls in.
// 31 bits, not 32, on purpose: with u on the full word 1 - u is a negation, min(u, 1 - u) an absolute value, and the compiler
// emits IABS + a SIGNED int-to-float conversion - wrong for the one value 2^31 (u = 1/2 exactly: it becomes -2^31 and the clip
// turns it into 1e-10).  (A branch-free fold with the sign applied by one LOP3 was measured 4% slower than compare-select.)
constexpr uint32_t kSobolOne = 0x80000000u;  // 1.0 on the 31-bit scale
__device__ __forceinline__ float inverse_normal_radius(uint32_t xs, bool& lower) {
  const uint32_t xr = kSobolOne - xs;
  lower = xs < xr;  // u < 1/2
  const float t = fmaxf(__uint2float_rn(lower ? xs : xr) * 4.6566128730773926e-10f, 1e-10f);
  return mufu_sqrt(mufu_lg2(t) * -1.38629436111989061883f);  // sqrt(-2 ln t)
}

#define B200MC_NDTRI_HORNER(FMA, C)                                                                   \
  p = FMA(p, v, C(3.917148571e-03f));  p = FMA(p, v, C(5.165553951e-03f));  p = FMA(p, v, C(-5.742339453e-03f)); \
  p = FMA(p, v, C(-8.094970152e-03f)); p = FMA(p, v, C(9.645064409e-03f));  p = FMA(p, v, C(-2.167277874e-03f)); \
  p = FMA(p, v, C(5.823089048e-03f));  p = FMA(p, v, C(-1.531743127e-02f)); p = FMA(p, v, C(2.503921833e-02f));  \
  p = FMA(p, v, C(-4.184841491e-02f)); p = FMA(p, v, C(7.395411314e-02f));  p = FMA(p, v, C(-1.353623019e-01f)); \
  p = FMA(p, v, C(3.068033996e+00f));  p = FMA(p, v, C(3.381260124e+00f));

__device__ __forceinline__ float inverse_normal_from_sobol(uint32_t xs) {
  bool lower;
  const float y = inverse_normal_radius(xs, lower);
  const float v = fmaf(y, 3.565869380e-01f, -1.419849035e+00f);
  float p = -2.964769098e-03f;
#define B200MC_ID(c) (c)
  B200MC_NDTRI_HORNER(fmaf, B200MC_ID)
#undef B200MC_ID
  return lower ? -p : p;
}

// Two points at once: the polynomial runs as packed FFMA2 (coefficients are FFMA2 immediates), 16 issue slots
// for two normals instead of 30; the roundings are those of the scalar form, so both give the same bits.
__device__ __forceinline__ void inverse_normal_from_sobol2(uint32_t xs0, uint32_t xs1, float& z0, float& z1) {
  bool lower0, lower1;
  const float y0 = inverse_normal_radius(xs0, lower0);
  const float y1 = inverse_normal_radius(xs1, lower1);
#define B200MC_BOTH(c) pack2((c), (c))
  const f32x2 v = fma2(pack2(y0, y1), B200MC_BOTH(3.565869380e-01f), B200MC_BOTH(-1.419849035e+00f));
  f32x2 p = B200MC_BOTH(-2.964769098e-03f);
  B200MC_NDTRI_HORNER(fma2, B200MC_BOTH)
#undef B200MC_BOTH
  float p0, p1;
  unpack2(p, p0, p1);
  z0 = lower0 ? -p0 : p0;
  z1 = lower1 ? -p1 : p1;
}

struct QmcCoef {
  float c;      // sigma*sqrt(dt) / ln2 : log2-diffusion per unit normal
  float a;      // (r - q - sigma^2/2) T / ln2 : terminal log2-drift
  float kappa;  // K / S
};

template <int NS, int PB>
__global__ void __launch_bounds__(kBlock, NS <= 2 ? 4 : 2) qmc_european_kernel(const SobolArgs a) {
  constexpr int kPoints = 1 << PB;
  constexpr int kCtaShift = PB + kSobolTidBits;
  __shared__ __align__(16) uint32_t c_s[kSobolDimChunk][kSobolWords];
  __shared__ uint32_t xcta_s[kSobolDimChunk];
  __shared__ QmcCoef coef[NS];
  __shared__ uint32_t last_part;
  const uint32_t dsplit = a.dsplit;
  const uint32_t part = blockIdx.x % dsplit;
  const uint32_t ot = blockIdx.x / dsplit;  // (option, tile)
  const uint32_t opt = ot / a.tiles;
  const uint32_t tile = ot - opt * a.tiles;
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

  const uint32_t cta_index = (uint32_t)(a.point_begin >> kCtaShift) + tile;  // point-index bits [PB + 8, bits)
  uint32_t tid_mask[kSobolTidBits];
#pragma unroll
  for (int b = 0; b < kSobolTidBits; ++b) tid_mask[b] = 0u - ((threadIdx.x >> b) & 1u);
  const uint32_t up = 31u - a.bits;  // the table is staged pre-shifted: integers x << up, u = (x << up) * 2^-31

  float W[kPoints];
#pragma unroll
  for (int k = 0; k < kPoints; ++k) W[k] = 0.0f;

  // this CTA's share of the dimensions: whole 64-dimension chunks
  const uint32_t chunks = (a.n_steps + kSobolDimChunk - 1) / kSobolDimChunk;
  const uint32_t chunks_per_part = (chunks + dsplit - 1) / dsplit;
  const uint32_t d_begin = min(part * chunks_per_part * kSobolDimChunk, a.n_steps);
  const uint32_t d_end = min(d_begin + chunks_per_part * kSobolDimChunk, a.n_steps);
  for (uint32_t d0 = d_begin; d0 < d_end; d0 += kSobolDimChunk) {
    const uint32_t nd = min((uint32_t)kSobolDimChunk, d_end - d0);
    __syncthreads();  // the previous chunk is fully consumed
    for (uint32_t i = threadIdx.x; i < nd * kSobolWords; i += kBlock) (&c_s[0][0])[i] = a.dirnums[(size_t)d0 * kSobolWords + i] << up;
    __syncthreads();
    if (threadIdx.x < nd) {
      uint32_t x = a.shift[d0 + threadIdx.x] << up;
      for (uint32_t ci = cta_index, b = kCtaShift; ci != 0; ci >>= 1, ++b)
        if (ci & 1u) x ^= c_s[threadIdx.x][b];
      xcta_s[threadIdx.x] = x;
    }
    __syncthreads();
    for (uint32_t j = 0; j < nd; ++j) {
      // words [0, PB) flip the thread's own points, words [PB, PB + 8) belong to the threadIdx bits
      uint32_t word[kSobolMaxPointBits + kSobolTidBits];
#pragma unroll
      for (int q = 0; q < (PB + kSobolTidBits + 3) / 4; ++q) {
        const uint4 v4 = *reinterpret_cast<const uint4*>(&c_s[j][4 * q]);
        word[4 * q] = v4.x, word[4 * q + 1] = v4.y, word[4 * q + 2] = v4.z, word[4 * q + 3] = v4.w;
      }
      uint32_t x = xcta_s[j];
#pragma unroll
      for (int b = 0; b < kSobolTidBits; ++b) x ^= word[PB + b] & tid_mask[b];
      if (kPoints == 1) {
        W[0] += inverse_normal_from_sobol(x);
      } else {
#pragma unroll
        for (int g = 0; g < kPoints; g += 2) {  // local point bits visited in Gray order: one XOR per point, two points per polynomial
          if (g > 0) x ^= word[(g & 2) ? 1 : (g & 4) ? 2 : 3];  // lowest set bit of the even g
          const uint32_t x0 = x;
          x ^= word[0];                                         // g + 1: lowest set bit 0
          float z0, z1;
          inverse_normal_from_sobol2(x0, x, z0, z1);
          W[g ^ (g >> 1)] += z0;
          W[(g + 1) ^ ((g + 1) >> 1)] += z1;
        }
      }
    }
  }

  if (dsplit > 1) {  // hand the partial sums to the tile's last CTA
    float* mine = a.wpart + ((size_t)ot * dsplit + part) * (kBlock * kPoints);
#pragma unroll
    for (int k = 0; k < kPoints; ++k) __stcg(mine + k * kBlock + threadIdx.x, W[k]);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
      last_part = atomicAdd(a.tile_tickets + ot, 1u) == dsplit - 1 ? 1u : 0u;
      if (last_part) a.tile_tickets[ot] = 0u;
    }
    __syncthreads();
    if (!last_part) return;
    __threadfence();
    const float* parts = a.wpart + (size_t)ot * dsplit * (kBlock * kPoints);
#pragma unroll
    for (int k = 0; k < kPoints; ++k) {
      float w = 0.0f;
      for (uint32_t pp = 0; pp < dsplit; ++pp) w += __ldcg(parts + (size_t)pp * (kBlock * kPoints) + k * kBlock + threadIdx.x);  // part order: fixed
      W[k] = w;
    }
  }
  __syncthreads();  // coef[] (written before the loop) is read below; a launch of fewer than one chunk has no barrier yet

  float acc[2 * NS];
  uint32_t paid[(NS + 3) / 4];
#pragma unroll
  for (int i = 0; i < 2 * NS; ++i) acc[i] = 0.0f;
#pragma unroll
  for (int i = 0; i < (NS + 3) / 4; ++i) paid[i] = 0u;
  const bool is_put = a.is_put != 0;
  const uint64_t first = ((uint64_t)tile * kBlock + threadIdx.x) * kPoints;  // local index of this thread's point 0
  if (a.terminal_out) {  // simulation layer (gbm_qmc.py:44-46, :70-76): the terminal prices of scenario 0, one option
    const QmcCoef q = coef[0];
    const double S = a.params[0].S;
#pragma unroll
    for (int k = 0; k < kPoints; ++k) {
      if (first + k < a.n_points) {
        a.terminal_out[first + k] = S * (double)mufu_ex2(fmaf(q.c, W[k], q.a));
        if (a.terminal_anti) a.terminal_out[a.n_points + first + k] = S * (double)mufu_ex2(fmaf(-q.c, W[k], q.a));
      }
    }
  }
#pragma unroll
  for (int k = 0; k < kPoints; ++k) {
    if (first + k < a.n_points) {
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        const QmcCoef q = coef[s];
        const float e = mufu_ex2(fmaf(q.c, W[k], q.a));
        add_sample<NS>(acc + 2 * s, paid, s, e, vanilla(e, q.kappa, is_put));
      }
    }
  }
  finish_tile<2, NS>(acc, paid, a.fold, opt, tile, a.tiles, a.n_scen,
                     [&](uint32_t k) { return vanilla_scale(a.params[(size_t)opt * a.n_scen + k], is_put); });
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
  for (uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (uint64_t)gridDim.x * blockDim.x)
    out[idx] = inverse_normal_from_sobol(x[idx] << (31u - bits));
}

}  // namespace b200mc
