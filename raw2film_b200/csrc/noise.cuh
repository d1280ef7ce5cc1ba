// Counter-based white Gaussian noise (a7): Philox4x32-10 keyed by the seed, counter = (pixel quad, row,
// channel); 4 uniforms -> 2 Box-Muller pairs -> 4 normals for 4 consecutive pixels.  Shared by k_noise
// and the fused grain kernels, which regenerate any part of the field on the fly.
// (reference GPU path: PCG-3D hash + Box-Muller, shaders/noise.wgsl:14-62; streams differ by design)
#pragma once
#include <cstdint>

#include <cuda_runtime.h>

namespace r2f {

__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
        c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}

__device__ __forceinline__ float u01(uint32_t x) { return ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f); }

// Four unit normals for pixels x = 4*qx .. 4*qx+3 of row y, channel ch.  The stream is a pure
// function of (qx, y, ch, seed), so any kernel can regenerate any part of the field.
__device__ __forceinline__ float4 noise_quad(uint32_t qx, uint32_t y, uint32_t ch, uint32_t k0, uint32_t k1) {
    uint32_t c[4] = {qx, y, ch, 0x52324631u};
    philox4x32_10(c, k0, k1);
    // Box-Muller on MUFU approximations (relative error ~1e-6: irrelevant for a noise field)
    const float l1 = -2.0f * __logf(u01(c[0])), l2 = -2.0f * __logf(u01(c[2]));
    const float r1 = l1 * rsqrtf(l1), r2 = l2 * rsqrtf(l2);  // sqrt(l); l > 0 since u01 < 1
    float s1, c1, s2, c2;
    __sincosf(6.28318530717958647692f * (u01(c[1]) - 0.5f), &s1, &c1);  // angle in [-pi, pi)
    __sincosf(6.28318530717958647692f * (u01(c[3]) - 0.5f), &s2, &c2);
    return make_float4(r1 * c1, r1 * s1, r2 * c2, r2 * s2);
}

__device__ __forceinline__ float noise_at(int x, int y, int ch, uint32_t k0, uint32_t k1) {
    const float4 q = noise_quad((uint32_t)x >> 2, (uint32_t)y, (uint32_t)ch, k0, k1);
    const int l = x & 3;
    return l == 0 ? q.x : l == 1 ? q.y : l == 2 ? q.z : q.w;
}

}  // namespace r2f
