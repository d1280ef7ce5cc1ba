// Inner loop shared by the y-symmetric correlation kernels (r2f_conv_sym.cu, r2f_grain_sym.cu).
//
//   acc[o] += sum_{dy=0..R} sum_{j<K} w[dy][j] * (row(+dy)[o + j] + row(-dy)[o + j]),   o = 0 .. OW-1
//
// for TWO tile rows at once: `ctr0` / `ctr1` point at the window start in the centre rows of the two outputs
// rows a thread owns, the packed FFMA2 lanes are (row 0, row 1).  `wsm` holds (w, w) pairs, WROW pairs per kernel
// row, centre row halved.  PITCH (floats) must be 4 (mod 8) so that the 128-bit loads of a warp whose lanes walk
// consecutive tile rows are bank-conflict free.  Per kernel row: 4*NQ LDS.128 + K/2 weight LDS.128, 2*NQ*4 FADD,
// K*OW FFMA2 -- nothing else (cuobjdump).
#pragma once
#include <cuda_runtime.h>

namespace r2f {

// UNIFORM: `wsm` points at (w, w) pairs in the kernel's parameter space (a __grid_constant__ struct) instead of
// shared memory.  The weight index is warp-uniform, so the pairs arrive through the uniform datapath (LDCU.64 into a
// uniform register that FFMA2 takes directly as an operand): no weight LDS.128 (9 of the 41 shared-memory loads per
// kernel row at k = 17), no vector registers for weights.
template <int K, int OW, int PITCH, int WROW, bool UNIFORM = false>
__device__ __forceinline__ void sym_correlate(const float *__restrict__ ctr0, const float *__restrict__ ctr1,
                                              const float *__restrict__ wsm, float2 (&acc)[OW]) {
    constexpr int R = K / 2;
    constexpr int NQ = (OW + K - 1 + 3) / 4;
#pragma unroll 1
    for (int dy = 0; dy <= R; ++dy) {
        const float4 *a0 = reinterpret_cast<const float4 *>(ctr0 + dy * PITCH);
        const float4 *b0 = reinterpret_cast<const float4 *>(ctr0 - dy * PITCH);
        const float4 *a1 = reinterpret_cast<const float4 *>(ctr1 + dy * PITCH);
        const float4 *b1 = reinterpret_cast<const float4 *>(ctr1 - dy * PITCH);
        float2 P[NQ * 4];
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            const float4 x = a0[q], y = b0[q], z = a1[q], w = b1[q];
            P[4 * q + 0] = make_float2(x.x + y.x, z.x + w.x);
            P[4 * q + 1] = make_float2(x.y + y.y, z.y + w.y);
            P[4 * q + 2] = make_float2(x.z + y.z, z.z + w.z);
            P[4 * q + 3] = make_float2(x.w + y.w, z.w + w.w);
        }
        if (UNIFORM) {
            const float2 *wr = reinterpret_cast<const float2 *>(wsm) + dy * WROW;
#pragma unroll
            for (int j = 0; j < K; ++j) {
                const float2 wa = wr[j];
#pragma unroll
                for (int o = 0; o < OW; ++o) acc[o] = __ffma2_rn(wa, P[o + j], acc[o]);
            }
        } else {
            const float4 *wr = reinterpret_cast<const float4 *>(wsm + dy * WROW * 2);
#pragma unroll
            for (int j = 0; j < K; j += 2) {
                const float4 w4 = wr[j >> 1];
                const float2 wa = make_float2(w4.x, w4.y);
#pragma unroll
                for (int o = 0; o < OW; ++o) acc[o] = __ffma2_rn(wa, P[o + j], acc[o]);
                if (j + 1 < K) {
                    const float2 wb = make_float2(w4.z, w4.w);
#pragma unroll
                    for (int o = 0; o < OW; ++o) acc[o] = __ffma2_rn(wb, P[o + j + 1], acc[o]);
                }
            }
        }
    }
}

}  // namespace r2f
