// Counter-based random numbers for the detector path: Philox4x32-10 (Salmon et al., SC'11).
//
// The reference threads ONE numpy PCG64 generator through all events (detector/simulator.py:169), so its
// stream depends on event order.  Here every draw is addressed by what it is FOR:
//   Fano normal  of (event, nucleus index, grid step k) -> counter (event_lo, event_hi, nucleus, k)
//   TB wiggle    of (event, Szudzik key)                -> counter (event_lo, event_hi, 0xFFFFFFFF, key), 16 bits
// keyed by the 64-bit seed, which makes results independent of batching and of the GPU that ran the event.
#pragma once
#include <cstdint>

namespace attpc {

struct Philox4 {
    uint32_t x, y, z, w;
};

__host__ __device__ __forceinline__ uint32_t mulhi32(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                         uint32_t k0, uint32_t k1) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = mulhi32(M0, c0), lo0 = M0 * c0;
        const uint32_t hi1 = mulhi32(M1, c2), lo1 = M1 * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0;
        c1 = lo1;
        c2 = n2;
        c3 = lo0;
        k0 += W0;
        k1 += W1;
    }
    return Philox4{c0, c1, c2, c3};
}

// 53-bit uniform in [0, 1), same bit recipe as numpy's next_double: (a >> 5) * 2^26 + (b >> 6)
__host__ __device__ __forceinline__ double uniform53(uint32_t a, uint32_t b) {
    return (double)(((uint64_t)(a >> 5) << 26) | (uint64_t)(b >> 6)) * (1.0 / 9007199254740992.0);
}

constexpr uint32_t STREAM_WIGGLE = 0xFFFFFFFFu;

// |z| never exceeds sqrt(-2 ln 2^-54) = 8.66 with the Box-Muller below.
constexpr double NORMAL_ABS_MAX = 8.66;

__device__ __forceinline__ double philox_normal(uint64_t seed, uint64_t event, uint32_t stream, uint32_t index) {
    const Philox4 r = philox4x32_10((uint32_t)event, (uint32_t)(event >> 32), stream, index, (uint32_t)seed,
                                    (uint32_t)(seed >> 32));
    // Box-Muller evaluated in single precision: the normal only enters int(mean + spread * z), whose own width
    // dwarfs a 2^-24 relative error of z.  u1 keeps its 53 random bits before the conversion, so the tail reaches
    // |z| = sqrt(-2 ln 2^-54) = 8.66 = NORMAL_ABS_MAX like the double-precision form.
    const float u1 = (float)(uniform53(r.x, r.y) + (0.5 / 9007199254740992.0));  // (0, 1]
    const float u2 = (float)(r.z >> 8) * (1.0f / 16777216.0f);
    return (double)(sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2));
}

__host__ __device__ __forceinline__ double philox_uniform(uint64_t seed, uint64_t event, uint32_t stream,
                                                          uint32_t index) {
    const Philox4 r = philox4x32_10((uint32_t)event, (uint32_t)(event >> 32), stream, index, (uint32_t)seed,
                                    (uint32_t)(seed >> 32));
    return uniform53(r.x, r.y);
}

// 16-bit uniform in [0, 1): k / 2^16.  time bucket + k / 2^16 is exact in float64 and travels as one Q16.16
// fixed-point uint32 ((bucket << 16) | k) without losing a bit.
__host__ __device__ __forceinline__ double philox_uniform16(uint64_t seed, uint64_t event, uint32_t stream,
                                                            uint32_t index) {
    const Philox4 r = philox4x32_10((uint32_t)event, (uint32_t)(event >> 32), stream, index, (uint32_t)seed,
                                    (uint32_t)(seed >> 32));
    return (double)(r.x >> 16) * (1.0 / 65536.0);
}

}  // namespace attpc
