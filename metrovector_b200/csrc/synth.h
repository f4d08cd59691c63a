// synth.h -- counter-based synthetic data, identical on host and device (bench + tests only).
//   rk      = sm64(sm64(seed) ^ row)
//   f32/f16 : h = sm64(rk + (col>>1)); w = the 32-bit half of h selected by (col&1);
//             v = float((w & 0xffff) + (w >> 16) - 65535) * 2^-15   -> triangular on (-2, 2), exact in f32
//             f16 = round-to-nearest-even(v)
//   u8/i8   : h = sm64(rk + (col>>3)); byte (col&7) of h
// Any row can be regenerated anywhere without storing the database, which is how the 100M- and
// 1B-row configs (bigger than host RAM) are checked against the CPU oracle.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define MVF_HD __host__ __device__ __forceinline__
#else
#define MVF_HD inline
#endif

namespace mvf {
namespace synth {

MVF_HD uint64_t sm64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
MVF_HD uint64_t row_key(uint64_t seed_key, uint64_t row) { return sm64(seed_key ^ row); }
// two consecutive float columns (2p, 2p+1) from one hash
MVF_HD void float_pair(uint64_t rk, uint32_t p, float* v0, float* v1) {
    uint64_t h = sm64(rk + p);
    uint32_t lo = (uint32_t)h, hi = (uint32_t)(h >> 32);
    *v0 = (float)((int32_t)((lo & 0xffffu) + (lo >> 16)) - 65535) * 0x1p-15f;
    *v1 = (float)((int32_t)((hi & 0xffffu) + (hi >> 16)) - 65535) * 0x1p-15f;
}
// eight consecutive byte columns (8p .. 8p+7), little-endian in the returned word
MVF_HD uint64_t byte_octet(uint64_t rk, uint32_t p) { return sm64(rk + p); }

}  // namespace synth
}  // namespace mvf
