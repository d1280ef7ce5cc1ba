// a7 + a9 + a10 fused, for grain kernels that are mirror-symmetric in y (every grain blob kernel is):
// white noise regenerated per tile (Philox), correlated with the grain kernel by the row-pair /
// packed-FMA scheme of r2f_conv_sym.cu, scaled by the density-dependent amplitude, added, clipped,
// then tetrahedral LUT + quantise.  Reads 12 B/px of density, writes 3 B/px; the noise field and the
// grained density never touch HBM.
//
// reference: effects.py:220-236 (apply_grain), cpu_processor.py:387-397, shaders/grain.wgsl:36-92,
// utils.py:247-380 (tetrahedral LUT), cpu_processor.py:407 (quantise).
//
// Thread mapping (64x64 tile, 256 threads): warp w owns tile columns 8w..8w+7, lane l owns tile rows
// l and l+32; the FFMA2 lanes are (row l, row l+32).  The noise tile's row pitch is 4 (mod 8) floats,
// so the 128-bit window loads of a warp (one tile row per lane) are bank-conflict free.
#include <cuda_runtime.h>

#include "conv_tile.cuh"
#include "noise.cuh"
#include "r2f_kernels.h"
#include "sym_conv.cuh"

namespace r2f {

namespace {

template <int K, int OW_>
struct GrainCfg {
    static constexpr int R = K / 2;
    static constexpr int T = 64, OW = OW_, NT = (T / OW_) * 32;
    static constexpr int NWIN = OW + K - 1;
    static constexpr int NQ = (NWIN + 3) / 4;
    static constexpr int COLS = T + K - 1;
    static constexpr int NEED = (T - OW) + 4 * NQ;
    static constexpr int P0 = ((COLS > NEED ? COLS : NEED) + 3) / 4 * 4;
    static constexpr int PITCH = (P0 % 8 == 4) ? P0 : P0 + 4;
    static constexpr int ROWS = T + K - 1;
    static constexpr int WROW = (K + 1) / 2 * 2;
    static constexpr int SPITCH = 196;  // staged output row pitch in bytes: 192 + 4, so the lanes (= rows) of a
                                        // warp hit different banks when they store their packed words
    static constexpr int TILE_BYTES = ROWS * PITCH * 4 > T * SPITCH ? ROWS * PITCH * 4 : T * SPITCH;
    static constexpr int TILE_FLOATS = (TILE_BYTES + 15) / 16 * 4;
    static constexpr int DP = T + 4;                     // density tile row pitch (floats): 17 16-byte chunks, odd, so
                                                         // the 128-bit accesses of a warp (one tile row per lane) are
                                                         // bank-conflict free
    static constexpr int PRIV_FLOATS = 3 * T * DP;       // the three layers' densities, overwritten in place by the
                                                         // grained densities
    static constexpr int SMEM_BYTES = (TILE_FLOATS + PRIV_FLOATS) * 4;
};

template <int K, int OW, bool GEN, bool FASTC>
__global__ void __launch_bounds__((64 / OW) * 32, 3)
k_grain_finish_sym(const __grid_constant__ GrainFinishArgs a) {
    using C = GrainCfg<K, OW>;
    extern __shared__ __align__(16) float smem[];
    float *tile = smem;
    float *priv = smem + C::TILE_FLOATS;
    const int H = a.H, W = a.W;
    const int tx0 = blockIdx.x * C::T, ty0 = (blockIdx.y + a.tile_y0) * C::T;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t ps = a.plane_stride;

    const int nch = a.bw ? 1 : 3;
    const int xs = tx0 - C::R;                                      // global x of tile column 0
    const bool interior = xs >= 0 && xs + C::COLS <= W;             // no horizontal reflection needed
    // quads of the shifted noise grid (noise.cuh): tile column 0 starts a quad, NQT of them cover a tile row
    constexpr int NQT = (C::COLS + 3) / 4;
    static_assert(4 * NQT <= C::PITCH, "a tile row of whole quads must fit the pitch");
    const int q0 = (xs + a.noise_shift) >> 2;

    // The densities of all three layers are requested up front as asynchronous global -> shared copies (LDGSTS: no
    // register, nothing waits) and complete behind the first noise field; the grain apply overwrites a thread's own
    // values in place with the grained densities.  Full tiles are copied row-wise (a warp moves two 256-byte tile rows
    // per instruction, 4 L1 wavefronts; with the thread mapping of the correlation -- one tile row per lane -- every
    // load touched 32 lines, and the density read was 47 % of the kernel's global tag requests and 15 % of its stall
    // samples).  Frames whose rows are not 16-byte aligned copy the same rows four bytes at a time.
    // (tiles that cross the bottom or right edge still copy row-wise and zero what lies outside the frame -- with W a
    // multiple of 4 a 16-byte chunk is inside or outside as a whole: on a small frame the kernel lasts as long as its
    // slowest CTA, and the per-thread path made the last tile row three times slower)
    const int dp = a.dens_pitch > 0 ? a.dens_pitch : W;   // row pitch of the density planes (padded for odd widths)
    const bool full_tile = (dp & 3) == 0 && (ps & 3) == 0 && (reinterpret_cast<uintptr_t>(a.dens) & 15) == 0;
    if (full_tile) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float *dplane = a.dens + c * ps + (size_t)ty0 * dp + tx0;
#pragma unroll
            for (int it = 0; it < C::T * (C::T / 4) / C::NT; ++it) {
                const int idx = threadIdx.x + it * C::NT, row = idx / (C::T / 4), ch = idx % (C::T / 4);
                float *dst = priv + (c * C::T + row) * C::DP + 4 * ch;
                // a chunk lies inside the padded row as a whole; the (up to three) pad columns of an odd-width frame
                // hold unspecified values, which only feed outputs that are never written
                if (ty0 + row < H && tx0 + 4 * ch < W) cp_async_16(dst, dplane + (size_t)row * dp + 4 * ch);
                else *reinterpret_cast<float4 *>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    } else {  // rows that are not 16-byte aligned: the same row-wise copies, four bytes at a time
#pragma unroll 1
        for (int c = 0; c < 3; ++c) {
            const float *dplane = a.dens + c * ps;
            for (int idx = threadIdx.x; idx < C::T * C::T; idx += C::NT) {
                const int row = idx / C::T, col = idx % C::T;
                const int gy = ty0 + row, gx = tx0 + col;
                float *dst = priv + (c * C::T + row) * C::DP + col;
                if (gy < H && gx < W) cp_async_4(dst, dplane + (size_t)gy * dp + gx);
                else *dst = 0.0f;
            }
        }
    }

    float2 g[C::OW];
#pragma unroll 1
    for (int c = 0; c < 3; ++c) {
        if (c < nch) {
            __syncthreads();  // the previous channel's window reads are done
            if (GEN) {
                if (interior) {  // one Philox call and one aligned 128-bit store per quad of four samples
                    for (int idx = threadIdx.x; idx < C::ROWS * NQT; idx += C::NT) {
                        const int ty = idx / NQT, tq = idx - ty * NQT;  // compile-time divisor
                        const int gy = reflect101(ty0 - C::R + ty, H);
                        const float4 v = noise_quad((uint32_t)(q0 + tq), gy, c, a.seed_lo, a.seed_hi);
                        *reinterpret_cast<float4 *>(tile + ty * C::PITCH + 4 * tq) = v;
                    }
                } else {
                    for (int idx = threadIdx.x; idx < C::ROWS * C::COLS; idx += C::NT) {
                        const int ty = idx / C::COLS, tx = idx - ty * C::COLS;
                        tile[ty * C::PITCH + tx] = noise_at(reflect101(xs + tx, W), reflect101(ty0 - C::R + ty, H), c,
                                                            a.seed_lo, a.seed_hi, a.noise_shift);
                    }
                }
            } else {
                fill_tile(tile, a.noise + (size_t)c * ps, C::ROWS, C::COLS, ty0 - C::R, xs, H, W, C::NT, C::PITCH);
            }
            if (c == 0) cp_async_wait_all();  // the density copies: visible to everyone after the barrier
            __syncthreads();
#pragma unroll
            for (int o = 0; o < C::OW; ++o) g[o] = make_float2(0.f, 0.f);
            const float *ctr0 = tile + (lane + C::R) * C::PITCH + C::OW * warp;
            const float *ctr1 = ctr0 + 32 * C::PITCH;
            sym_correlate<K, C::OW, C::PITCH, C::WROW, true>(ctr0, ctr1, reinterpret_cast<const float *>(a.gkw), g);
        }
        // grain apply on channel c (black-and-white grain reuses the single field)
#pragma unroll
        for (int j = 0; j < 2 * (C::OW / 4); ++j) {  // chunk j: tile row lane + 32 * h, columns 4 * q .. 4 * q + 3 of the thread's
            constexpr int QPR = C::OW / 4;
            const int h = j / QPR, q = j % QPR;
            float4 *slot = reinterpret_cast<float4 *>(priv + (c * C::T + lane + 32 * h) * C::DP + C::OW * warp) + q;
            const float4 dq = *slot;
            const float d[4] = {dq.x, dq.y, dq.z, dq.w};
            float val[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int o = 4 * q + i;
                const float gn = h == 0 ? g[o].x : g[o].y;
                const float amp = FASTC ? fast_curve_eval(a.gfast, c, d[i]) : curve_eval(a.gcurve, c, d[i]);
                const float v = d[i] + gn * amp;
                val[i] = v > 0.0f ? v : 0.0f;
            }
            *slot = make_float4(val[0], val[1], val[2], val[3]);
        }
    }
    __syncthreads();  // all noise-tile reads done: its storage becomes the output staging area
    if (a.dens_out != nullptr) {  // grain stage only: the grained density tile leaves as planar float32 rows
#pragma unroll 1
        for (int c = 0; c < 3; ++c) {
            float *oplane = a.dens_out + c * ps;
            for (int idx = threadIdx.x; idx < C::T * (C::T / 4); idx += C::NT) {
                const int row = idx / (C::T / 4), ch = idx % (C::T / 4);
                const int gy = ty0 + row, gx = tx0 + 4 * ch;
                if (gy >= H || gx >= W) continue;
                const float4 v = *reinterpret_cast<const float4 *>(priv + (c * C::T + row) * C::DP + 4 * ch);
                float *op = oplane + (size_t)gy * W + gx;
                if ((W & 3) == 0 && (ps & 3) == 0 && (reinterpret_cast<uintptr_t>(a.dens_out) & 15) == 0) {
                    *reinterpret_cast<float4 *>(op) = v;
                } else {
                    op[0] = v.x;
                    if (gx + 1 < W) op[1] = v.y;
                    if (gx + 2 < W) op[2] = v.z;
                    if (gx + 3 < W) op[3] = v.w;
                }
            }
        }
        return;
    }
    uint8_t *stage = reinterpret_cast<uint8_t *>(tile);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        uint32_t px[C::OW];  // 0x00BBGGRR per pixel
        unsigned undecided = 0;
#pragma unroll
        for (int half = 0; half < C::OW / 4; ++half) {
            const float *slot = priv + (lane + 32 * h) * C::DP + C::OW * warp + 4 * half;
            const float4 q0d = *reinterpret_cast<const float4 *>(slot);
            const float4 q1d = *reinterpret_cast<const float4 *>(slot + C::T * C::DP);
            const float4 q2d = *reinterpret_cast<const float4 *>(slot + 2 * C::T * C::DP);
            const float d0[4] = {q0d.x, q0d.y, q0d.z, q0d.w}, d1[4] = {q1d.x, q1d.y, q1d.z, q1d.w};
            const float d2[4] = {q2d.x, q2d.y, q2d.z, q2d.w};
            // the grain stage clipped: densities are >= 0.  Branch-free, so the gathers of the pixels overlap
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (!tetra_u8_try<true>(a.ft, d0[i], d1[i], d2[i], px[4 * half + i])) undecided |= 1u << (4 * half + i);
        }
        if (undecided) {  // rare: the exact binary64 interpolation decides (unrolled: px[] stays in registers)
#pragma unroll
            for (int o = 0; o < C::OW; ++o)
                if (undecided >> o & 1u) {
                    const float *slot = priv + (lane + 32 * h) * C::DP + C::OW * warp + o;
                    px[o] = tetra_exact_u8(a.l3, slot[0], slot[C::T * C::DP], slot[2 * C::T * C::DP]);
                }
        }
        // OW pixels = 3 * OW bytes, packed words
        uint32_t *sp = reinterpret_cast<uint32_t *>(stage + (lane + 32 * h) * C::SPITCH + 3 * C::OW * warp);
#pragma unroll
        for (int q4 = 0; q4 < C::OW / 4; ++q4) {
            const uint32_t p0 = px[4 * q4], p1 = px[4 * q4 + 1], p2 = px[4 * q4 + 2], p3 = px[4 * q4 + 3];
            sp[3 * q4 + 0] = p0 | (p1 << 24);
            sp[3 * q4 + 1] = (p1 >> 8) | (p2 << 16);
            sp[3 * q4 + 2] = (p2 >> 16) | (p3 << 8);
        }
    }
    __syncthreads();
    const int tw = min(C::T, W - tx0), th = min(C::T, H - ty0);
    const int row_bytes = tw * 3;
    if (tw == C::T && ((size_t)W * 3 % 16) == 0 && (reinterpret_cast<uintptr_t>(a.out_u8) & 15) == 0) {
        for (int idx = threadIdx.x; idx < th * 12; idx += C::NT) {  // 192 bytes = 12 x 16 per row
            const int r = idx / 12, s16 = idx - r * 12;
            const uint32_t *sp = reinterpret_cast<const uint32_t *>(stage + r * C::SPITCH + s16 * 16);
            const uint4 v = make_uint4(sp[0], sp[1], sp[2], sp[3]);
            __stcs(reinterpret_cast<uint4 *>(a.out_u8 + ((size_t)(ty0 + r) * W + tx0) * 3) + s16, v);
        }
    } else {
        // rows whose byte address is not 16-byte aligned (widths that are not multiples of 16/3 pixels, edge tiles): a
        // warp writes one tile row as aligned 32-bit words funnel-shifted out of the staged bytes, plus at most three
        // leading and three trailing single bytes (byte-wise stores of the whole row cost 0.1 ms per 24 MP frame)
        const uint32_t *stage32 = reinterpret_cast<const uint32_t *>(stage);
        for (int r = warp; r < th; r += C::NT / 32) {
            uint8_t *gbase = a.out_u8 + ((size_t)(ty0 + r) * W + tx0) * 3;
            const uint8_t *sbase = stage + r * C::SPITCH;           // SPITCH is a multiple of 4
            const int mis = (int)(reinterpret_cast<uintptr_t>(gbase) & 3);
            const int lead = mis ? min(4 - mis, row_bytes) : 0;
            if (lane < lead) gbase[lane] = sbase[lane];
            const int nwords = (row_bytes - lead) >> 2;
            for (int k = lane; k < nwords; k += 32) {
                const int sb = lead + 4 * k, wi = (r * C::SPITCH + sb) >> 2;
                const uint32_t v = __funnelshift_r(stage32[wi], stage32[wi + 1], 8 * (sb & 3));   // bytes sb .. sb + 3
                *reinterpret_cast<uint32_t *>(gbase + sb) = v;
            }
            const int tail0 = lead + 4 * nwords;
            if (lane < row_bytes - tail0) gbase[tail0 + lane] = sbase[tail0 + lane];
        }
    }
}

template <int K>
cudaError_t launch_gs(const GrainFinishArgs &a, cudaStream_t st) {
    constexpr int OW = 8;  // 16-wide strips (4 warps per CTA) share more window loads but measured 0.49 vs 0.40 ms at K = 7
    using C = GrainCfg<K, OW>;
    dim3 grid((a.W + C::T - 1) / C::T, a.tile_rows > 0 ? a.tile_rows : (a.H + C::T - 1) / C::T);
    cudaError_t e;
#define R2F_GS_LAUNCH(GEN_, FC_)                                                                                   \
    do {                                                                                                          \
        auto kfn = k_grain_finish_sym<K, OW, GEN_, FC_>;                                                              \
        if ((e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES)) !=        \
            cudaSuccess)                                                                                          \
            return e;                                                                                             \
        kfn<<<grid, C::NT, C::SMEM_BYTES, st>>>(a);                                                               \
    } while (0)
    const bool fc = a.gfast.seg != nullptr;
    if (a.noise == nullptr) {
        if (fc) R2F_GS_LAUNCH(true, true);
        else R2F_GS_LAUNCH(true, false);
    } else {
        if (fc) R2F_GS_LAUNCH(false, true);
        else R2F_GS_LAUNCH(false, false);
    }
#undef R2F_GS_LAUNCH
    return cudaGetLastError();
}

}  // namespace

bool grain_finish_sym_supported(int k) { return k >= 3 && k <= kGrainSymMaxK && (k & 1); }

cudaError_t launch_grain_finish_sym(const GrainFinishArgs &a, cudaStream_t st) {
    if (a.gk_sym == nullptr || a.burn.map != nullptr) return cudaErrorInvalidValue;   // burn: dens_out + launch_finish
    switch (a.k) {
#define R2F_GS(KK) case KK: return launch_gs<KK>(a, st);
        R2F_GS(3) R2F_GS(5) R2F_GS(7) R2F_GS(9) R2F_GS(11) R2F_GS(13) R2F_GS(15) R2F_GS(17) R2F_GS(19) R2F_GS(21)
#undef R2F_GS
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace r2f
