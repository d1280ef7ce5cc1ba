// Direct 2-D correlation for kernels that are mirror-symmetric in y (K[r+dy][j] == K[r-dy][j]):
// the MTF kernels of effects.py:123-185 (|ifft2| of a radial transfer function), the halation
// kernel of effects.py:200-263 and the grain blobs all are.
//
//   out[y][x] = sum_{dy=0..r} sum_j  w[dy][j] * (in[y+dy][x+j-r] + in[y-dy][x+j-r])
//
// with the centre row stored halved (x*w == (x+x)*(w/2) exactly), so the loop has no special case.
// The row-pair sums are shared by the 8 horizontally adjacent outputs a thread owns, which cuts
// the work from k*k to ~(k+1)/2 * k multiply-adds plus (k+7)/8 adds per (dy, output).
//
// Blackwell specifics: the multiply-adds are FFMA2 (fma.rn.f32x2): one issue slot per two FMAs.
// A thread owns OW (16) adjacent outputs in each of TWO tile rows (ty and ty+32); the packed lanes are
// (row ty, row ty+32), so the
// window pair P[i] = (s_ty[i], s_ty+32[i]) is alignment-free for every tap offset j, and the
// weight operand is a pre-duplicated (w, w) pair read as a shared-memory broadcast.
// Lanes of a warp walk 32 consecutive tile rows; the row pitch is 4 (mod 8) floats, which makes
// every 128-bit window load conflict-free (8 lanes x 16 B cover all 32 banks).
//
// Reference semantics kept: correlation (not convolution), centre anchor, BORDER_REFLECT_101
// (cv.filter2D as called from effects.py:146-156).
#include <cuda_runtime.h>

#include "conv_tile.cuh"
#include "r2f_kernels.h"

namespace r2f {

namespace {

template <int K, int OW_>
struct SymCfg {
    static constexpr int R = K / 2;
    static constexpr int TW = 64, TH = 64, OW = OW_, NT = (TW / OW_) * 32;
    static constexpr int NWIN = OW + K - 1;            // window floats per thread per row
    static constexpr int NQ = (NWIN + 3) / 4;          // ... as float4 loads
    static constexpr int COLS = (TW + K - 1 + 3) / 4 * 4;  // tile width rounded up: whole 16-byte copies
    static constexpr int SHIFT = (4 - R % 4) % 4;          // tiles start at 64*bx - SHIFT, so that the tile's
                                                           // left edge (origin - R) is 4-float aligned
    static constexpr int NEED = (TW - OW) + 4 * NQ;    // right-most float a window load touches + 1
    static constexpr int P0 = ((COLS > NEED ? COLS : NEED) + 3) / 4 * 4;
    static constexpr int PITCH = (P0 % 8 == 4) ? P0 : P0 + 4;
    static constexpr int ROWS = TH + K - 1;
    static constexpr int WROW = (K + 1) / 2 * 2;       // (w,w) pairs per kernel row, even: 2 taps per LDS.128
    static constexpr int OPITCH = TW + 1;              // output staging pitch (odd: conflict-free)
    static constexpr int TILE_FLOATS = ROWS * PITCH > TH * OPITCH ? ROWS * PITCH : TH * OPITCH;
    static constexpr int SMEM_BYTES = (TILE_FLOATS + (R + 1) * WROW * 2) * 4;
};

template <int K, int OW>
__global__ void __launch_bounds__((64 / OW) * 32, OW == 8 ? 3 : 4)
k_conv2d_sym(ConvArgs a) {
    using C = SymCfg<K, OW>;
    extern __shared__ __align__(16) float smem[];
    float *tile = smem;
    float *wsm = smem + C::TILE_FLOATS;
    const int c = blockIdx.z;
    const int tx0 = blockIdx.x * C::TW - C::SHIFT, ty0 = blockIdx.y * C::TH;
    const int H = a.H, W = a.W;
    const float *__restrict__ src = a.in + (size_t)a.in_plane[c] * a.plane_stride;
    float *__restrict__ dst = a.out + (size_t)c * a.plane_stride;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (a.mode[c] != 0) {
        fill_tile_async<C::ROWS, C::COLS, C::PITCH, C::NT>(tile, src, ty0 - C::R, tx0 - C::R, H, W);
        const float *__restrict__ wg = a.ksym[c];
        for (int idx = threadIdx.x; idx < (C::R + 1) * C::WROW * 2; idx += C::NT) wsm[idx] = __ldg(wg + idx);
        cp_async_wait_all();
        __syncthreads();

        float2 acc[C::OW];
#pragma unroll
        for (int o = 0; o < C::OW; ++o) acc[o] = make_float2(0.f, 0.f);
        // rows lane and lane+32 of the tile; window starts at tile column 8*warp
        const float *ctr0 = tile + (lane + C::R) * C::PITCH + C::OW * warp;
        const float *ctr1 = ctr0 + 32 * C::PITCH;
#pragma unroll 1
        for (int dy = 0; dy <= C::R; ++dy) {
            const float4 *a0 = reinterpret_cast<const float4 *>(ctr0 + dy * C::PITCH);
            const float4 *b0 = reinterpret_cast<const float4 *>(ctr0 - dy * C::PITCH);
            const float4 *a1 = reinterpret_cast<const float4 *>(ctr1 + dy * C::PITCH);
            const float4 *b1 = reinterpret_cast<const float4 *>(ctr1 - dy * C::PITCH);
            float2 P[C::NQ * 4];
#pragma unroll
            for (int q = 0; q < C::NQ; ++q) {
                const float4 x = a0[q], y = b0[q], z = a1[q], w = b1[q];
                P[4 * q + 0] = make_float2(x.x + y.x, z.x + w.x);
                P[4 * q + 1] = make_float2(x.y + y.y, z.y + w.y);
                P[4 * q + 2] = make_float2(x.z + y.z, z.z + w.z);
                P[4 * q + 3] = make_float2(x.w + y.w, z.w + w.w);
            }
            const float4 *wr = reinterpret_cast<const float4 *>(wsm + dy * C::WROW * 2);
#pragma unroll
            for (int j = 0; j < K; j += 2) {
                const float4 w4 = wr[j >> 1];
                const float2 wa = make_float2(w4.x, w4.y);
#pragma unroll
                for (int o = 0; o < C::OW; ++o) acc[o] = __ffma2_rn(wa, P[o + j], acc[o]);
                if (j + 1 < K) {
                    const float2 wb = make_float2(w4.z, w4.w);
#pragma unroll
                    for (int o = 0; o < C::OW; ++o) acc[o] = __ffma2_rn(wb, P[o + j + 1], acc[o]);
                }
            }
        }
        __syncthreads();  // everyone is done reading the input tile: reuse it as the output stage
#pragma unroll
        for (int o = 0; o < C::OW; ++o) {
            tile[lane * C::OPITCH + C::OW * warp + o] = acc[o].x;
            tile[(lane + 32) * C::OPITCH + C::OW * warp + o] = acc[o].y;
        }
        __syncthreads();
    }

    // coalesced row-wise write-out with the fused epilogue; identity layers copy straight through
    const bool conv = a.mode[c] != 0;
    const int col = threadIdx.x & 63, rsub = threadIdx.x >> 6;
    const int gx = tx0 + col;
    if (gx >= 0 && gx < W) {
#pragma unroll 4
        for (int rr = rsub; rr < C::TH; rr += C::NT / 64) {
            const int gy = ty0 + rr;
            if (gy >= H) break;
            const size_t idx = (size_t)gy * W + gx;
            float val = conv ? tile[rr * C::OPITCH + col] : __ldg(src + idx);
            if (a.epi == EPI_DENSITY) val = density_eval(a.curve, c, val, a.eps);
            else if (a.epi == EPI_DENSITY_FAST) val = density_eval_fast(a.curve, c, val, a.eps);
            dst[idx] = val;
        }
    }
}

template <int K>
cudaError_t launch_sym(const ConvArgs &a, cudaStream_t st) {
    constexpr int OW = 16;
    using C = SymCfg<K, OW>;
    auto kfn = k_conv2d_sym<K, OW>;
    cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    dim3 grid((a.W + C::SHIFT + C::TW - 1) / C::TW, (a.H + C::TH - 1) / C::TH, 3);
    kfn<<<grid, C::NT, C::SMEM_BYTES, st>>>(a);
    return cudaGetLastError();
}

}  // namespace

int conv_sym_wrow(int k) { return (k + 1) / 2 * 2; }

bool conv_sym_supported(int k) { return k >= 3 && k <= kConvSymMaxK && (k & 1); }

cudaError_t launch_conv2d_sym(const ConvArgs &a, cudaStream_t st) {
    if (a.epi == EPI_GRAIN) return cudaErrorInvalidValue;
    switch (a.k) {
#define R2F_SYM(KK) case KK: return launch_sym<KK>(a, st);
        R2F_SYM(3) R2F_SYM(5) R2F_SYM(7) R2F_SYM(9) R2F_SYM(11) R2F_SYM(13) R2F_SYM(15) R2F_SYM(17)
        R2F_SYM(19) R2F_SYM(21) R2F_SYM(23) R2F_SYM(25) R2F_SYM(27) R2F_SYM(29) R2F_SYM(31) R2F_SYM(33)
#undef R2F_SYM
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace r2f
