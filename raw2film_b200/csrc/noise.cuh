// Counter-based white Gaussian noise (a7): Philox4x32 keyed by the seed, counter = (pixel quad, row,
// channel); 4 random words -> 2 Box-Muller pairs -> 4 normals for 4 consecutive pixels.  Shared by k_noise
// and the fused grain kernels, which regenerate any part of the field on the fly.
// (reference GPU path: PCG-3D hash + Box-Muller, shaders/noise.wgsl:14-62; streams differ by design)
//
// Rounds: 7.  Philox4x32-7 is the shortest variant that passes BigCrush (Salmon et al., "Parallel random numbers:
// as easy as 1, 2, 3", SC'11, table 2); the usual 10 rounds add a safety margin a film-grain field does not need,
// and the generator is 40 % of the fused grain kernel's instructions.  The moment / spectrum / independence tests
// (tests/test_gpu_full.py) run on this stream.
#pragma once
#include <cstdint>

#include <cuda_runtime.h>

namespace r2f {

constexpr int kPhiloxRounds = 7;

__device__ __forceinline__ void philox4x32(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < kPhiloxRounds; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
        c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}

// 23 random bits as the mantissa of a float in [1, 2): no integer -> float conversion (quarter-rate XU pipe)
__device__ __forceinline__ float bits_to_1_2(uint32_t x) { return __uint_as_float((x >> 9) | 0x3f800000u); }

// Two unit normals from two random words: Box-Muller on MUFU approximations (relative error ~1e-6: irrelevant for
// a noise field).  Inputs stay in the normal range, so the .ftz forms need no denormal guards.
__device__ __forceinline__ float2 box_muller(uint32_t a, uint32_t b) {
    const float u = 2.0f - bits_to_1_2(a);             // (0, 1]
    float l, r, s, c;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(u));
    const float t = l * -1.3862943611198906f;          // -2 ln(u) >= 0
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t));
    const float ang = (bits_to_1_2(b) - 1.5f) * 6.28318530717958647692f;  // [-pi, pi)
    asm("sin.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(ang));
    asm("cos.approx.ftz.f32 %0, %1;" : "=f"(c) : "f"(ang));
    return make_float2(r * c, r * s);
}

// Four unit normals for pixels x = 4*qx .. 4*qx+3 of row y, channel ch.  The stream is a pure
// function of (qx, y, ch, seed), so any kernel can regenerate any part of the field.
__device__ __forceinline__ float4 noise_quad(uint32_t qx, uint32_t y, uint32_t ch, uint32_t k0, uint32_t k1) {
    uint32_t c[4] = {qx, y, ch, 0x52324631u};
    philox4x32(c, k0, k1);
    const float2 p = box_muller(c[0], c[1]), q = box_muller(c[2], c[3]);
    return make_float4(p.x, p.y, q.x, q.y);
}

// The field at pixel (x, y, ch): lane (x + shift) & 3 of quad (x + shift) >> 2.  `shift` = (grain kernel radius) & 3
// puts the quad grid on the column grid of the fused kernels' noise tiles (tile column 0 is frame column
// 64 * bx - radius), so that a tile row is filled with aligned 128-bit stores; every kernel that touches the field
// of a render uses the same shift.
__device__ __forceinline__ float noise_at(int x, int y, int ch, uint32_t k0, uint32_t k1, int shift) {
    const uint32_t xs = (uint32_t)(x + shift);
    const float4 q = noise_quad(xs >> 2, (uint32_t)y, (uint32_t)ch, k0, k1);
    const int l = xs & 3;
    return l == 0 ? q.x : l == 1 ? q.y : l == 2 ? q.z : q.w;
}

}  // namespace r2f
