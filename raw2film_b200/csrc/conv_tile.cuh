// Shared-memory tile staging shared by the direct-correlation kernels.
#pragma once
#include <cstdint>

#include "device_math.cuh"

namespace r2f {

// Cooperative load of a (rows x cols) tile whose top-left corner is (gy0, gx0) in a W x H plane,
// BORDER_REFLECT_101 outside the plane.  One warp per tile row, lanes along x: no div/mod, and the
// reflection is only evaluated for tiles that actually cross the frame border.
__device__ __forceinline__ void fill_tile(float *__restrict__ tile, const float *__restrict__ src, int rows, int cols,
                                          int gy0, int gx0, int H, int W, int nthreads, int pitch = 0) {
    if (pitch == 0) pitch = cols;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = nthreads >> 5;
    const bool inside_x = gx0 >= 0 && gx0 + cols <= W;
    for (int ty = warp; ty < rows; ty += nwarps) {
        const float *row = src + (size_t)reflect101(gy0 + ty, H) * W;
        float *dst = tile + ty * pitch;
        if (inside_x) {
            for (int tx = lane; tx < cols; tx += 32) dst[tx] = __ldg(row + gx0 + tx);
        } else {
            for (int tx = lane; tx < cols; tx += 32) dst[tx] = __ldg(row + reflect101(gx0 + tx, W));
        }
    }
}

// Asynchronous variant (LDGSTS / cp.async): every thread queues all of its copies before anyone waits,
// so one tile fill costs about one HBM round trip instead of one per tile row.  16-byte copies when the
// tile's left edge, the frame pitch and the tile width are 4-float aligned and the tile does not cross
// the left/right frame border; 4-byte copies with reflected addresses otherwise.
// Complete with cp_async_wait_all() + __syncthreads().
__device__ __forceinline__ void cp_async_4(float *smem_dst, const float *gsrc) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_16(float *smem_dst, const float *gsrc) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// `gpitch`: row pitch of the source plane in floats (0: W).
template <int ROWS, int COLS, int PITCH, int NT>
__device__ __forceinline__ void fill_tile_async(float *__restrict__ tile, const float *__restrict__ src, int gy0,
                                                int gx0, int H, int W, int gpitch = 0) {
    if (gpitch == 0) gpitch = W;
    const bool inside_x = gx0 >= 0 && gx0 + COLS <= W;
    if (COLS % 4 == 0 && inside_x && ((gx0 | gpitch) & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
        constexpr int QPR = COLS / 4;
        for (int idx = threadIdx.x; idx < ROWS * QPR; idx += NT) {
            const int ty = idx / QPR, q = idx - ty * QPR;
            cp_async_16(tile + ty * PITCH + 4 * q, src + (size_t)reflect101(gy0 + ty, H) * gpitch + gx0 + 4 * q);
        }
    } else {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        for (int ty = warp; ty < ROWS; ty += NT / 32) {
            const float *row = src + (size_t)reflect101(gy0 + ty, H) * gpitch;
            float *dst = tile + ty * PITCH;
            if (inside_x) {
                for (int tx = lane; tx < COLS; tx += 32) cp_async_4(dst + tx, row + gx0 + tx);
            } else {
                for (int tx = lane; tx < COLS; tx += 32) cp_async_4(dst + tx, row + reflect101(gx0 + tx, W));
            }
        }
    }
}

}  // namespace r2f
