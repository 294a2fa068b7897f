// Host build of the DEVICE header philox.cuh (its functions are __host__ __device__): lets the CPU
// test-suite check the exact code the kernels inline against the Random123 known answers.
#include "../optionslab_b200/csrc/philox.cuh"
extern "C" void host_philox4x32_10(const uint32_t* ck, uint32_t n, uint32_t* out) {
  for (uint32_t i = 0; i < n; ++i) {
    b200mc::u32x4 x = b200mc::philox4x32<10>(ck[6 * i], ck[6 * i + 1], ck[6 * i + 2], ck[6 * i + 3], ck[6 * i + 4], ck[6 * i + 5]);
    out[4 * i] = x.x, out[4 * i + 1] = x.y, out[4 * i + 2] = x.z, out[4 * i + 3] = x.w;
  }
}
// The keyed form the simulation kernels call: round keys expanded once (philox_expand_key), then philox4x32_10.
extern "C" void host_philox4x32_10_keyed(const uint32_t* ck, uint32_t n, uint32_t* out) {
  for (uint32_t i = 0; i < n; ++i) {
    const b200mc::PhiloxKeys rk = b200mc::philox_expand_key(ck[6 * i + 4], ck[6 * i + 5]);
    b200mc::u32x4 x = b200mc::philox4x32_10(ck[6 * i], ck[6 * i + 1], ck[6 * i + 2], ck[6 * i + 3], rk);
    out[4 * i] = x.x, out[4 * i + 1] = x.y, out[4 * i + 2] = x.z, out[4 * i + 3] = x.w;
  }
}
