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
// Tile staging: interior tiles are fetched by ONE TMA bulk-tensor copy (cp.async.bulk.tensor.3d over a
// {W, H, plane} tensor map, box = PITCH x ROWS x 1, completion on an mbarrier) issued by a single thread;
// tiles that touch the frame border need BORDER_REFLECT_101 addresses, which TMA's zero fill cannot
// express, and use per-thread cp.async copies instead.
//
// Reference semantics kept: correlation (not convolution), centre anchor, BORDER_REFLECT_101
// (cv.filter2D as called from effects.py:146-156).
#include <cstdlib>
#include <cstring>

#include <cuda.h>
#include <cuda_runtime.h>

#include "conv_tile.cuh"
#include "r2f_kernels.h"
#include "sym_conv.cuh"

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
    static constexpr int OPITCH = TW + 1;              // output staging pitch (odd: conflict-free), scalar write-out
    static constexpr int OP4 = TW + 4;                 // ... for the 128-bit write-out: 17 16-byte chunks per row (odd)
    static constexpr int TILE_FLOATS = ROWS * PITCH > TH * OP4 ? ROWS * PITCH : TH * OP4;
    static constexpr int SMEM_BYTES = TILE_FLOATS * 4;
};

// the three layers' weights as a kernel parameter (3.9 KB at k = 17, 13.5 KB at k = 33; the limit is 32 KB)
template <int K>
struct SymWeights {
    float2 w[3][(K / 2 + 1) * ((K + 1) / 2 * 2)];
};

// ---- TMA / mbarrier primitives (PTX ISA 8.x, sm_90+) ---------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned phase) {
    const unsigned addr = (unsigned)__cvta_generic_to_shared(bar);
    unsigned done = 0;
    while (!done) {
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
            : "=r"(done)
            : "r"(addr), "r"(phase)
            : "memory");
    }
}
__device__ __forceinline__ void tma_load_3d(float *smem_dst, const CUtensorMap *map, int x, int y, int z,
                                            uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(map), "r"(x), "r"(y), "r"(z),
        "r"((unsigned)__cvta_generic_to_shared(bar))
        : "memory");
}

template <int K, int OW>
__global__ void __launch_bounds__((64 / OW) * 32, OW == 8 ? 3 : 6)
k_conv2d_sym(ConvArgs a, const __grid_constant__ CUtensorMap tmap, int use_tma, int shift,
             const __grid_constant__ SymWeights<K> wts) {
    using C = SymCfg<K, OW>;
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) uint64_t tma_bar;
    float *tile = smem;
    const int c = blockIdx.z;
    const int tx0 = blockIdx.x * C::TW - shift, ty0 = ((int)blockIdx.y + a.tile_y0) * C::TH;
    const int H = a.H, W = a.W;
    const int gp = a.pitch > 0 ? a.pitch : W;   // row pitch of the planes (padded for odd widths)
    const float *__restrict__ src = a.in + (size_t)a.in_plane[c] * a.plane_stride;
    float *__restrict__ dst = a.out + (size_t)c * a.plane_stride;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    float2 acc[C::OW];
#pragma unroll
    for (int o = 0; o < C::OW; ++o) acc[o] = make_float2(0.f, 0.f);
    if (a.mode[c] != 0) {
        const int gx0 = tx0 - C::R, gy0 = ty0 - C::R;
        // TMA when every element the windows read lies inside the frame (the box may overhang by the pitch padding)
        const bool tma = use_tma && gx0 >= 0 && gy0 >= 0 && gx0 + C::COLS <= W && gy0 + C::ROWS <= H;
        if (tma) {
            if (threadIdx.x == 0) mbar_init(&tma_bar, 1);
            __syncthreads();
            if (threadIdx.x == 0) {
                mbar_expect_tx(&tma_bar, C::ROWS * C::PITCH * 4);
                tma_load_3d(tile, &tmap, gx0, gy0, a.in_plane[c], &tma_bar);
            }
        } else {
            fill_tile_async<C::ROWS, C::COLS, C::PITCH, C::NT>(tile, src, gy0, gx0, H, W, gp);
        }
        if (tma) mbar_wait(&tma_bar, 0);
        else cp_async_wait_all();
        __syncthreads();

        // rows lane and lane+32 of the tile; window starts at tile column OW*warp
        const float *ctr0 = tile + (lane + C::R) * C::PITCH + C::OW * warp;
        const float *ctr1 = ctr0 + 32 * C::PITCH;
        sym_correlate<K, C::OW, C::PITCH, C::WROW, true>(ctr0, ctr1, reinterpret_cast<const float *>(wts.w[c]), acc);
        __syncthreads();  // everyone is done reading the input tile: reuse it as the output stage
    }
    // 128-bit write-out when the tile's columns start on a 16-byte boundary of the destination rows (shift == 0: TMA
    // staging, or a kernel radius that is a multiple of 4) and the tile lies inside the frame
    // horizontally: 4 + 4 staging stores and 8 x (LDS.128, STG.128) per thread instead of 32 + 32 x (LDS.32, STG.32);
    // the scalar write-out was a fifth of the kernel's stall samples.
    const bool conv = a.mode[c] != 0;
    const bool vec_out = shift == 0 && (gp & 3) == 0 && (a.plane_stride & 3) == 0 && tx0 + C::TW <= W &&
                         (reinterpret_cast<uintptr_t>(a.out) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.in) & 15) == 0;
    auto epilogue = [&](float val) {
        if (a.epi == EPI_DENSITY) val = density_eval(a.curve, c, val, a.eps);
        else if (a.epi == EPI_DENSITY_FAST) val = density_eval_fast(a.curve, c, val, a.eps);
        return val;
    };
    if (vec_out) {
        if (conv) {
#pragma unroll
            for (int q = 0; q < C::OW / 4; ++q) {
                *reinterpret_cast<float4 *>(tile + lane * C::OP4 + C::OW * warp + 4 * q) =
                    make_float4(acc[4 * q].x, acc[4 * q + 1].x, acc[4 * q + 2].x, acc[4 * q + 3].x);
                *reinterpret_cast<float4 *>(tile + (lane + 32) * C::OP4 + C::OW * warp + 4 * q) =
                    make_float4(acc[4 * q].y, acc[4 * q + 1].y, acc[4 * q + 2].y, acc[4 * q + 3].y);
            }
            __syncthreads();
        }
#pragma unroll 4
        for (int idx = threadIdx.x; idx < C::TH * (C::TW / 4); idx += C::NT) {
            const int rr = idx / (C::TW / 4), c4 = idx % (C::TW / 4);
            const int gy = ty0 + rr;
            if (gy >= H) break;
            const size_t gi = (size_t)gy * gp + tx0 + 4 * c4;
            float4 v = conv ? *reinterpret_cast<const float4 *>(tile + rr * C::OP4 + 4 * c4)
                            : __ldg(reinterpret_cast<const float4 *>(src + gi));
            v.x = epilogue(v.x); v.y = epilogue(v.y); v.z = epilogue(v.z); v.w = epilogue(v.w);
            *reinterpret_cast<float4 *>(dst + gi) = v;
        }
        return;
    }
    if (conv) {
#pragma unroll
        for (int o = 0; o < C::OW; ++o) {
            tile[lane * C::OPITCH + C::OW * warp + o] = acc[o].x;
            tile[(lane + 32) * C::OPITCH + C::OW * warp + o] = acc[o].y;
        }
        __syncthreads();
    }

    // coalesced row-wise write-out with the fused epilogue; identity layers copy straight through
    const int col = threadIdx.x & 63, rsub = threadIdx.x >> 6;
    const int gx = tx0 + col;
    if (gx >= 0 && gx < W) {
#pragma unroll 4
        for (int rr = rsub; rr < C::TH; rr += C::NT / 64) {
            const int gy = ty0 + rr;
            if (gy >= H) break;
            const size_t idx = (size_t)gy * gp + gx;
            dst[idx] = epilogue(conv ? tile[rr * C::OPITCH + col] : __ldg(src + idx));
        }
    }
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time libcuda dependency)
using EncodeTiledFn = CUresult (*)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                   const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

// {W, H, 3 planes} float32 tensor map over the planar source with a box of pitch x rows x 1.
// Returns false when the layout does not meet TMA's alignment rules (the kernel then uses cp.async).
bool make_plane_map(const ConvArgs &a, int box_w, int box_h, CUtensorMap *map) {
    if (getenv("R2F_NO_TMA")) return false;
    EncodeTiledFn enc = encode_tiled_fn();
    const int gp = a.pitch > 0 ? a.pitch : a.W;
    if (!enc || (gp & 3) != 0 || (reinterpret_cast<uintptr_t>(a.in) & 15) != 0 || box_w > 256 || box_h > 256)
        return false;
    const cuuint64_t dims[3] = {(cuuint64_t)a.W, (cuuint64_t)a.H, 3};
    const cuuint64_t strides[2] = {(cuuint64_t)gp * 4, (cuuint64_t)a.plane_stride * 4};
    const cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(a.in), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int K>
cudaError_t launch_sym(const ConvArgs &a, cudaStream_t st) {
    constexpr int OW = 16;
    using C = SymCfg<K, OW>;
    auto kfn = k_conv2d_sym<K, OW>;
    cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    CUtensorMap map{};
    const int use_tma = make_plane_map(a, C::PITCH, C::ROWS, &map) ? 1 : 0;
    // the tiles start SHIFT columns left of a 64-column boundary so that the input tile's left edge is 16-byte aligned
    // (cp.async and TMA both need that: a TMA box starting at an unaligned column faults as an illegal instruction)
    const int shift = C::SHIFT;
    dim3 grid((a.W + shift + C::TW - 1) / C::TW, a.tile_rows > 0 ? a.tile_rows : (a.H + C::TH - 1) / C::TH, 3);
    SymWeights<K> wts;
    for (int c = 0; c < 3; ++c)
        if (a.mode[c] != 0) memcpy(wts.w[c], a.ksym_host[c], sizeof(wts.w[c]));
        else memset(wts.w[c], 0, sizeof(wts.w[c]));
    kfn<<<grid, C::NT, C::SMEM_BYTES, st>>>(a, map, use_tma, shift, wts);
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
