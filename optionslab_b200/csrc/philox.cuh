// Philox4x32-10 counter-based generator (Salmon, Moraes, Dror, Shaw — SC'11), written from the
// published round function: two 32x32->64 multiplies by fixed constants, the high words XORed with
// the other counter words and the round key, the key bumped by Weyl constants each round.
//
// RNG stream contract of this engine (normative statement: normal.cuh; restated in
// oracle/philox_oracle.c):
//   key     = (seed & 0xffffffff, seed >> 32)
//   counter = (path & 0xffffffff, j, path >> 32, stream),   j = 0, 1, 2, ... the call index of the path
//   the 4 output words of call j are 4 Box-Muller pairs and feed the normals of steps 8j .. 8j+7 of that path
//   (word i: steps 8j + 2i and 8j + 2i + 1).  The fast-varying word j sits in c1, which round 1 only XORs.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define B200MC_HD __host__ __device__ __forceinline__
#else
#define B200MC_HD static inline
#endif

namespace b200mc {

constexpr uint32_t kPhiloxM0 = 0xD2511F53u;
constexpr uint32_t kPhiloxM1 = 0xCD9E8D57u;
constexpr uint32_t kPhiloxW0 = 0x9E3779B9u;  // golden ratio
constexpr uint32_t kPhiloxW1 = 0xBB67AE85u;  // sqrt(3) - 1

struct u32x4 {
  uint32_t x, y, z, w;
};

B200MC_HD void mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
  const uint64_t p = (uint64_t)a * (uint64_t)b;  // one IMAD.WIDE.U32 on the device
  hi = (uint32_t)(p >> 32);
  lo = (uint32_t)p;
}

template <int ROUNDS = 10>
B200MC_HD u32x4 philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < ROUNDS; ++r) {
    uint32_t hi0, lo0, hi1, lo1;
    mulhilo(kPhiloxM0, c0, hi0, lo0);
    mulhilo(kPhiloxM1, c2, hi1, lo1);
    const uint32_t n0 = hi1 ^ c1 ^ k0;
    const uint32_t n2 = hi0 ^ c3 ^ k1;
    c0 = n0;
    c1 = lo1;
    c2 = n2;
    c3 = lo0;
    k0 += kPhiloxW0;
    k1 += kPhiloxW1;
  }
  return u32x4{c0, c1, c2, c3};
}

// The ten round keys, expanded once.  Kernels take them as launch parameters (constant bank operands of the
// round XORs) so that no kernel spends loop instructions or registers re-deriving the Weyl sequence.
struct PhiloxKeys {
  uint32_t k[20];  // k[2r], k[2r+1] = key words of round r
};

B200MC_HD PhiloxKeys philox_expand_key(uint32_t k0, uint32_t k1) {
  PhiloxKeys rk;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    rk.k[2 * r] = k0 + (uint32_t)r * kPhiloxW0;
    rk.k[2 * r + 1] = k1 + (uint32_t)r * kPhiloxW1;
  }
  return rk;
}

B200MC_HD u32x4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const PhiloxKeys& rk) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0, lo0, hi1, lo1;
    mulhilo(kPhiloxM0, c0, hi0, lo0);
    mulhilo(kPhiloxM1, c2, hi1, lo1);
    const uint32_t n0 = hi1 ^ c1 ^ rk.k[2 * r];
    const uint32_t n2 = hi0 ^ c3 ^ rk.k[2 * r + 1];
    c0 = n0;
    c1 = lo1;
    c2 = n2;
    c3 = lo0;
  }
  return u32x4{c0, c1, c2, c3};
}

}  // namespace b200mc
