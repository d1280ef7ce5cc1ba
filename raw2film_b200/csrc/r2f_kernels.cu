// sm_100a kernels of the raw2film render path.  Compile with -fmad=false (see device_math.cuh).
//
// Stage map (reference src/raw2film/cpu_processor.py:363-407):
//   k_pointwise   a2+a4+a5+a9+a10 fused (configs without spatial stages)
//   k_expose      a2            XYZ -> planar film exposure
//   k_conv2d      a3 / a6 / a7  direct 2-D correlation (cv2.filter2D semantics) with fused
//                               epilogues: log10+H-D curve (a4+a5) or grain apply + clip (a7)
//   k_noise       a7            white N(0,1) field (Philox4x32-10 + Box-Muller)
//   k_burn_*      a8            low-res highlight mask
//   k_finish      a8+a9+a10     burn apply, tetrahedral LUT, quantise
#include <cstdlib>

#include "conv_tile.cuh"
#include "fast_chain.cuh"
#include "noise.cuh"
#include "r2f_kernels.h"

namespace r2f {

constexpr int kThreads = 256;

__device__ __forceinline__ void store_quad_u8(uint8_t *__restrict__ out, size_t q, const uint32_t (&b)[12]) {
    uint32_t *o = reinterpret_cast<uint32_t *>(out) + 3 * q;
    __stcs(o + 0, b[0] | (b[1] << 8) | (b[2] << 16) | (b[3] << 24));
    __stcs(o + 1, b[4] | (b[5] << 8) | (b[6] << 16) | (b[7] << 24));
    __stcs(o + 2, b[8] | (b[9] << 8) | (b[10] << 16) | (b[11] << 24));
}

// stage the small tables (2-D LUT, curve rows) in shared memory
__device__ __forceinline__ void stage_tables(float *smem, Lut2D &l2, Curve1D &cv, bool want2d, bool want1d) {
    // asynchronous 16-byte copies (all of a thread's copies in flight at once) when the table is 16-byte aligned
    auto stage = [](float *dst, const float *src, int n) {
        const int nq = (reinterpret_cast<uintptr_t>(src) & 15) == 0 ? n / 4 : 0;
        for (int i = threadIdx.x; i < nq; i += blockDim.x) cp_async_16(dst + 4 * i, src + 4 * i);
        for (int i = 4 * nq + threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
    };
    float *p = smem;
    if (want2d) {  // the float4-padded copy
        const int n = l2.n * l2.n * 4;
        stage(p, reinterpret_cast<const float *>(l2.tab4), n);
        l2.tab4 = reinterpret_cast<const float4 *>(p);
        p += n;
    }
    if (want1d) {
        const int n = cv.N * 3 * 2;
        stage(p, reinterpret_cast<const float *>(cv.seg), n);
        cv.seg = reinterpret_cast<const float2 *>(p);
    }
    cp_async_wait_all();
    __syncthreads();
}

template <bool SMEM>
__device__ __forceinline__ void pixel_chain(const float (&xyz)[3], const Lut2D &l2, const Curve1D &cv, float eps,
                                            const Lut3D &l3, uint32_t &r, uint32_t &g, uint32_t &b) {
    float e0, e1, e2;
    lut2d_eval<SMEM, true>(l2, xyz[0], xyz[1], xyz[2], e0, e1, e2);
    // tables parked in shared memory are uniform by construction (launch_pointwise); the global-memory variant
    // also serves curves with a non-uniform abscissa
    const float d0 = density_eval<!SMEM>(cv, 0, e0, eps);
    const float d1 = density_eval<!SMEM>(cv, 1, e1, eps);
    const float d2 = density_eval<!SMEM>(cv, 2, e2, eps);
    tetra_quant_u8(l3, d0, d1, d2, r, g, b);
}

// ------------------------------------------------------------------------------------------
// K1: fused pointwise chain
// ------------------------------------------------------------------------------------------
// 512-thread CTAs: with the tables in shared memory (60 KB at the default sizes) three 256-thread CTAs fit an SM
// (24 warps); two 512-thread CTAs carry 32 warps in the same register file (64 registers per thread).
constexpr int kPwThreads = 512;
constexpr size_t kMaxTableSmem = 96 * 1024;  // keep >= 2 CTAs/SM resident

template <int FMT, bool SMEM_TABLES>
__global__ void __launch_bounds__(kPwThreads, 2)
k_pointwise(const void *__restrict__ in, float gain, uint8_t *__restrict__ out, size_t npix, Lut2D l2, Curve1D cv,
            float eps, Lut3D l3) {
    extern __shared__ __align__(16) float smem[];
    if (SMEM_TABLES) stage_tables(smem, l2, cv, true, true);
    const size_t nquad = npix / 4;
    const size_t stride = (size_t)gridDim.x * kPwThreads;
    for (size_t q = (size_t)blockIdx.x * kPwThreads + threadIdx.x; q < nquad; q += stride) {
        float px[4][3];
        load_quad<FMT>(in, q, gain, px);
        uint32_t b[12];
#pragma unroll
        for (int p = 0; p < 4; ++p)
            pixel_chain<SMEM_TABLES>(px[p], l2, cv, eps, l3, b[3 * p], b[3 * p + 1], b[3 * p + 2]);
        store_quad_u8(out, q, b);
    }
    if (blockIdx.x == 0 && threadIdx.x < (npix & 3)) {
        const size_t p = nquad * 4 + threadIdx.x;
        float xyz[3];
        load_px<FMT>(in, p, gain, xyz[0], xyz[1], xyz[2]);
        uint32_t r, g, b;
        pixel_chain<SMEM_TABLES>(xyz, l2, cv, eps, l3, r, g, b);
        out[p * 3] = (uint8_t)r;
        out[p * 3 + 1] = (uint8_t)g;
        out[p * 3 + 2] = (uint8_t)b;
    }
}

// ------------------------------------------------------------------------------------------
// K1 fast: the same chain through the guarded float32 fast path (fast_chain.cuh).  Pixels the fast path cannot
// prove are queued per warp and evaluated 32 at a time by the exact chain, which overwrites their bytes.
// No CTA-wide synchronisation after the table staging: ballots, the queue and the drain are warp-local.
// ------------------------------------------------------------------------------------------
constexpr int kPwQueue = 64;  // per-warp queue capacity: < 32 left over + at most 32 pushed per ballot

template <int FMT>
static __device__ __noinline__ void pw_exact_pixel(const void *__restrict__ in, float gain, uint8_t *__restrict__ out,
                                                   unsigned pix, Lut2D l2, Curve1D cv, float eps, Lut3D l3) {
    float xyz[3];
    load_px<FMT>(in, pix, gain, xyz[0], xyz[1], xyz[2]);
    float e0, e1, e2;
    lut2d_eval<true, true>(l2, xyz[0], xyz[1], xyz[2], e0, e1, e2);
    const float d0 = density_eval(cv, 0, e0, eps), d1 = density_eval(cv, 1, e1, eps), d2 = density_eval(cv, 2, e2, eps);
    float o0, o1, o2;
    tetra_eval(l3, d0, d1, d2, o0, o1, o2);
    uint8_t *o = out + (size_t)pix * 3;
    o[0] = (uint8_t)quantise_u8(o0);
    o[1] = (uint8_t)quantise_u8(o1);
    o[2] = (uint8_t)quantise_u8(o2);
}

template <int FMT, int NT, bool PAIR>
__global__ void __launch_bounds__(NT, NT >= 768 ? 1 : 2)
k_pointwise_fast(const void *__restrict__ in, float gain, uint8_t *__restrict__ out, size_t npix, Lut2D l2, Curve1D cv,
                 float eps, Lut3D l3, FastChain F, unsigned long long *__restrict__ stats) {
    extern __shared__ __align__(16) float smem[];
    // shared memory: float4-padded 2-D LUT | scaled curve segments | per-warp queues
    {
        // row pitch n + 1 vertices: without the pad, cells in neighbouring rows are n * 16 bytes = a multiple of 128
        // apart, i.e. in the same banks, and 40 % of the shared-memory wavefronts were conflicts (ncu r02_f)
        const int n = l2.n, n2f = n * (n + 1) * 4, n1f = F.N * 3 * 2;
        const float *src2 = reinterpret_cast<const float *>(l2.tab4), *src1 = reinterpret_cast<const float *>(F.fseg);
        for (int i = threadIdx.x; i < n * n; i += NT) {
            const int r = i / n, col = i - r * n;
            cp_async_16(smem + 4 * (r * (n + 1) + col), src2 + 4 * i);
        }
        for (int i = threadIdx.x; i < n1f / 4; i += NT) cp_async_16(smem + n2f + 4 * i, src1 + 4 * i);
        for (int i = (n1f / 4) * 4 + threadIdx.x; i < n1f; i += NT) smem[n2f + i] = src1[i];
        l2.tab4 = reinterpret_cast<const float4 *>(smem);
        l2.pitch4 = n + 1;
        F.fseg = reinterpret_cast<const float2 *>(smem + n2f);
        cp_async_wait_all();
        __syncthreads();
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    FastChainS S;
    S.lut2d = (unsigned)__cvta_generic_to_shared(smem);
    S.row16 = (unsigned)(l2.n + 1) * 16u;
    S.n2 = l2.n + 1;  // row pitch in vertices
    S.n2m1 = (float)(l2.n - 1);
    S.hi2 = (float)(l2.n - 2);
    S.eps = eps;
    S.cA = F.cA;
    S.cB = F.cB;
    S.pscale = F.pscale;
    S.half_m = 0.5f - F.margin;
    for (int ch = 0; ch < 3; ++ch)
        S.seg_w[ch] = S.lut2d + (unsigned)(l2.n * (l2.n + 1) * 16) + (unsigned)(ch * F.N) * 8u - (kMagicBits << 3);
    S.lut = F.lut255;
    S.n3 = F.n3;
    S.o111 = F.n3 * F.n3 + F.n3 + 1;
    S.neg_k = 0u - kMagicBits * (unsigned)S.o111;
    unsigned *wq = reinterpret_cast<unsigned *>(smem + l2.n * (l2.n + 1) * 4 + ((F.N * 3 * 2 + 3) & ~3)) + warp * kPwQueue;
    int wq_count = 0;       // warp-uniform
    unsigned deferred = 0;  // lane 0 counts for the statistics
    const size_t nquad = npix / 4;
    const size_t stride = (size_t)gridDim.x * NT;
    for (size_t qb = (size_t)blockIdx.x * NT + warp * 32; qb < nquad; qb += stride) {
        const size_t q = qb + lane;
        const bool active = q < nquad;
        unsigned bad = 0;  // bit p: pixel p of the quad is undecided
        if (active) {
            float px[4][3];
            load_quad<FMT>(in, q, gain, px);
            uint32_t b[12];
            if (PAIR) {
#pragma unroll
                for (int p = 0; p < 4; p += 2) {
                    uint32_t qa[3], qb[3];
                    bad |= chain_fast_pair(px[p], px[p + 1], S, qa, qb) << p;
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        b[3 * p + c] = qa[c];
                        b[3 * p + 3 + c] = qb[c];
                    }
                }
            } else {
#pragma unroll
                for (int p = 0; p < 4; ++p)
                    if (!chain_fast_s(px[p], S, b[3 * p], b[3 * p + 1], b[3 * p + 2])) bad |= 1u << p;
            }
            store_quad_u8(out, q, b);
        }
        // the stores above are ordered before any overwrite by another lane of this warp: __syncwarp in the drain
        if (__any_sync(0xffffffffu, bad != 0)) {
#pragma unroll 1
            for (int p = 0; p < 4; ++p) {
                const bool u = (bad >> p) & 1u;
                const unsigned bal = __ballot_sync(0xffffffffu, u);
                if (bal == 0) continue;
                if (u) wq[wq_count + __popc(bal & ((1u << lane) - 1u))] = (unsigned)(q * 4 + p);
                wq_count += __popc(bal);
                __syncwarp();
                if (wq_count >= 32) {
                    wq_count -= 32;
                    pw_exact_pixel<FMT>(in, gain, out, wq[wq_count + lane], l2, cv, eps, l3);
                    deferred += 32;
                    __syncwarp();
                }
            }
        }
    }
    __syncwarp();
    if (lane < wq_count) pw_exact_pixel<FMT>(in, gain, out, wq[lane], l2, cv, eps, l3);
    deferred += wq_count;
    if (stats != nullptr && lane == 0 && deferred) atomicAdd(stats, (unsigned long long)deferred);
    if (blockIdx.x == 0 && threadIdx.x < (npix & 3))
        pw_exact_pixel<FMT>(in, gain, out, (unsigned)(nquad * 4 + threadIdx.x), l2, cv, eps, l3);
}

size_t pointwise_fast_smem(const Lut2D &l2, const FastChain &F) {
    return ((size_t)l2.n * (l2.n + 1) * 4 + (((size_t)F.N * 3 * 2 + 3) & ~(size_t)3)) * sizeof(float) +
           (size_t)(1024 / 32) * kPwQueue * sizeof(unsigned);
}

// CTA shape.  One 1024-thread CTA per SM parks ONE copy of the tables (96 KB), which leaves ~156 KB of the SM's
// unified L1 to the 3-D LUT gathers; two 512-thread CTAs park two copies and measure 0.27 ms at 24 MP against
// 0.173 ms (profiles/r02_*).  A/B knob: R2F_PW_THREADS = 1024 (default) | 768 | 384.
static int pw_fast_threads() {
    static const int v = [] {
        const char *e = getenv("R2F_PW_THREADS");
        const int t = e ? atoi(e) : 1024;
        return (t == 384 || t == 768) ? t : 1024;
    }();
    return v;
}

// R2F_PW_PAIR=0: scalar fast chain (chain_fast_s) instead of the packed two-pixel form (chain_fast_pair)
static bool pw_fast_pair() {
    static const bool v = [] {
        const char *e = getenv("R2F_PW_PAIR");
        return !(e && e[0] == '0');
    }();
    return v;
}

cudaError_t launch_pointwise_fast(const void *in, int fmt, float gain, uint8_t *out, size_t npix, const Lut2D &l2,
                                  const Curve1D &cv, float eps, const Lut3D &l3, const FastChain &F,
                                  unsigned long long *stats, int num_sms, cudaStream_t st) {
    const size_t sm = pointwise_fast_smem(l2, F);
    if (!F.ok || sm > kMaxTableSmem + 12288 || cv.xp != nullptr || npix >= ((size_t)1 << 32))
        return cudaErrorInvalidValue;
    const int nt = pw_fast_threads(), per_sm = nt >= 768 ? 1 : 2;
    const bool pair = pw_fast_pair();
    int grid = (int)((npix / 4 + nt) / nt);
    if (grid > num_sms * per_sm) grid = num_sms * per_sm;
#define R2F_LAUNCH_PWF3(C, T, P)                                                                                \
    do {                                                                                                        \
        auto kfn = k_pointwise_fast<C, T, P>;                                                                   \
        cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);        \
        if (e != cudaSuccess) return e;                                                                         \
        kfn<<<grid, T, sm, st>>>(in, gain, out, npix, l2, cv, eps, l3, F, stats);                               \
    } while (0)
#define R2F_LAUNCH_PWF2(C, T)                                                                                   \
    do {                                                                                                        \
        if (pair) R2F_LAUNCH_PWF3(C, T, true);                                                                  \
        else R2F_LAUNCH_PWF3(C, T, false);                                                                      \
    } while (0)
#define R2F_LAUNCH_PWF(C)                                                                                       \
    do {                                                                                                        \
        if (nt == 384) R2F_LAUNCH_PWF2(C, 384);                                                                 \
        else if (nt == 768) R2F_LAUNCH_PWF2(C, 768);                                                            \
        else R2F_LAUNCH_PWF2(C, 1024);                                                                          \
    } while (0)
    switch (fmt) {
        case 0: R2F_LAUNCH_PWF(0); break;
        case 1: R2F_LAUNCH_PWF(1); break;
        case 2: R2F_LAUNCH_PWF(2); break;
        case 3: R2F_LAUNCH_PWF(3); break;
        default: return cudaErrorInvalidValue;
    }
#undef R2F_LAUNCH_PWF3
#undef R2F_LAUNCH_PWF2
#undef R2F_LAUNCH_PWF
    return cudaGetLastError();
}

static int grid_for(size_t work_items, int num_sms, int ctas_per_sm) {
    size_t want = (work_items + kThreads - 1) / kThreads;
    size_t cap = (size_t)num_sms * ctas_per_sm;
    if (want < 1) want = 1;
    return (int)(want < cap ? want : cap);
}

static size_t table_smem_bytes(const Lut2D &l2, const Curve1D &cv, bool want2d, bool want1d) {
    size_t f = 0;
    if (want2d) f += (size_t)l2.n * l2.n * 4;
    if (want1d) f += (size_t)cv.N * 3 * 2;
    return f * sizeof(float);
}


cudaError_t launch_pointwise(const void *in, int fmt, float gain, uint8_t *out, size_t npix, const Lut2D &l2,
                             const Curve1D &cv, float eps, const Lut3D &l3, int num_sms, cudaStream_t st) {
    const size_t sm = table_smem_bytes(l2, cv, true, true);
    const bool use_smem = sm <= kMaxTableSmem && cv.xp == nullptr;
    int grid = (int)((npix / 4 + kPwThreads) / kPwThreads);
    if (grid > num_sms * 2) grid = num_sms * 2;
#define R2F_LAUNCH_PW(C, S)                                                                                \
    do {                                                                                                   \
        auto kfn = k_pointwise<C, S>;                                                                      \
        if (S) {                                                                                           \
            cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm); \
            if (e != cudaSuccess) return e;                                                                \
        }                                                                                                  \
        kfn<<<grid, kPwThreads, S ? sm : 0, st>>>(in, gain, out, npix, l2, cv, eps, l3);                   \
    } while (0)
    switch (fmt * 2 + (use_smem ? 1 : 0)) {
        case 0: R2F_LAUNCH_PW(0, false); break;
        case 1: R2F_LAUNCH_PW(0, true); break;
        case 2: R2F_LAUNCH_PW(1, false); break;
        case 3: R2F_LAUNCH_PW(1, true); break;
        case 4: R2F_LAUNCH_PW(2, false); break;
        case 5: R2F_LAUNCH_PW(2, true); break;
        case 6: R2F_LAUNCH_PW(3, false); break;
        case 7: R2F_LAUNCH_PW(3, true); break;
        default: return cudaErrorInvalidValue;
    }
#undef R2F_LAUNCH_PW
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// a2: XYZ -> planar exposure
// ------------------------------------------------------------------------------------------
template <int FMT, bool SMEM_TABLES>
__global__ void __launch_bounds__(kThreads)
k_expose(const void *__restrict__ in, float gain, float *__restrict__ out, size_t plane_stride, size_t npix,
         Lut2D l2) {
    extern __shared__ __align__(16) float smem[];
    Curve1D none{};
    if (SMEM_TABLES) stage_tables(smem, l2, none, true, false);
    const size_t nquad = npix / 4;
    const size_t stride = (size_t)gridDim.x * kThreads;
    for (size_t q = (size_t)blockIdx.x * kThreads + threadIdx.x; q < nquad; q += stride) {
        float px[4][3];
        load_quad<FMT>(in, q, gain, px);
        float e[3][4];
#pragma unroll
        for (int p = 0; p < 4; ++p)
            lut2d_eval<SMEM_TABLES, true>(l2, px[p][0], px[p][1], px[p][2], e[0][p], e[1][p], e[2][p]);
#pragma unroll
        for (int c = 0; c < 3; ++c)
            reinterpret_cast<float4 *>(out + c * plane_stride)[q] = make_float4(e[c][0], e[c][1], e[c][2], e[c][3]);
    }
    if (blockIdx.x == 0 && threadIdx.x < (npix & 3)) {
        const size_t p = nquad * 4 + threadIdx.x;
        float X, Y, Z, e0, e1, e2;
        load_px<FMT>(in, p, gain, X, Y, Z);
        lut2d_eval<SMEM_TABLES, true>(l2, X, Y, Z, e0, e1, e2);
        out[p] = e0;
        out[plane_stride + p] = e1;
        out[2 * plane_stride + p] = e2;
    }
}

cudaError_t launch_expose(const void *in, int fmt, float gain, Planes out, size_t npix, const Lut2D &l2, int num_sms,
                          cudaStream_t st) {
    Curve1D none{};
    const size_t sm = table_smem_bytes(l2, none, true, false);
    const bool use_smem = sm <= kMaxTableSmem;
    const int grid = grid_for(npix / 4 + 1, num_sms, 8);
#define R2F_LAUNCH_EX(C, S)                                                                                \
    do {                                                                                                   \
        auto kfn = k_expose<C, S>;                                                                         \
        if (S) {                                                                                           \
            cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm); \
            if (e != cudaSuccess) return e;                                                                \
        }                                                                                                  \
        kfn<<<grid, kThreads, S ? sm : 0, st>>>(in, gain, out.base, out.plane_stride, npix, l2);           \
    } while (0)
    switch (fmt * 2 + (use_smem ? 1 : 0)) {
        case 0: R2F_LAUNCH_EX(0, false); break;
        case 1: R2F_LAUNCH_EX(0, true); break;
        case 2: R2F_LAUNCH_EX(1, false); break;
        case 3: R2F_LAUNCH_EX(1, true); break;
        case 4: R2F_LAUNCH_EX(2, false); break;
        case 5: R2F_LAUNCH_EX(2, true); break;
        case 6: R2F_LAUNCH_EX(3, false); break;
        case 7: R2F_LAUNCH_EX(3, true); break;
        default: return cudaErrorInvalidValue;
    }
#undef R2F_LAUNCH_EX
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// a3/a6/a7: direct 2-D correlation, BORDER_REFLECT_101, fused epilogue
//
// CTA tile TW x TH outputs of one channel.  The input tile (+ (k-1) halo) and the kernel
// weights (transposed, so the weights a thread walks are contiguous) live in shared memory.
// Each thread owns a vertical strip of 16 outputs in one column; adjacent lanes own adjacent
// columns, so every shared-memory read is conflict-free and every weight read is a broadcast.
// For kernel column j the thread slides a 16-row register window down the tile column:
// one new LDS per 16 FMAs.
// ------------------------------------------------------------------------------------------
// One kernel row: 16 FMAs on the register window, then (unless LAST) slide the window down by
// one tile row.  `u` is the compile-time slot of the row that leaves the window.
template <int LAST>
__device__ __forceinline__ void conv_step(float (&acc)[16], float (&v)[16], float w, int u, const float *&rowp,
                                          int cols) {
#pragma unroll
    for (int o = 0; o < 16; ++o) acc[o] = fmaf(w, v[(o + u) & 15], acc[o]);
    if (!LAST) {
        v[u & 15] = *rowp;
        rowp += cols;
    }
}

// The last R (< 16) kernel rows of a kernel column, fully unrolled; the final row loads nothing.
template <int R>
__device__ __forceinline__ void conv_tail(float (&acc)[16], float (&v)[16], const float4 *wp, const float *rowp,
                                          int cols) {
    float w[16];
#pragma unroll
    for (int q = 0; q < (R + 3) / 4; ++q) {
        const float4 w4 = wp[q];
        w[4 * q] = w4.x; w[4 * q + 1] = w4.y; w[4 * q + 2] = w4.z; w[4 * q + 3] = w4.w;
    }
#pragma unroll
    for (int u = 0; u < R; ++u) {
        if (u + 1 < R) conv_step<0>(acc, v, w[u], u, rowp, cols);
        else conv_step<1>(acc, v, w[u], u, rowp, cols);
    }
}

// PITCH > 0: compile-time tile row pitch (floats), so the window loads of the unrolled inner loop
// become LDS with immediate offsets (no per-step pointer arithmetic); PITCH == 0: pitch = TW + k - 1.
template <int TW, int TH, bool W_SMEM, int PITCH>
__global__ void __launch_bounds__((TW / 32) * (TH / 16) * 32)
k_conv2d(const __grid_constant__ ConvArgs a) {
    constexpr int NWX = TW / 32;
    constexpr int NT = (TW / 32) * (TH / 16) * 32;
    extern __shared__ __align__(16) float smem[];
    const int c = blockIdx.z;
    const int tx0 = blockIdx.x * TW, ty0 = blockIdx.y * TH;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int lx = (warp % NWX) * 32 + lane;
    const int ly0 = (warp / NWX) * 16;
    const int gx = tx0 + lx;
    const int H = a.H, W = a.W;
    const float *__restrict__ src = a.in + (size_t)a.in_plane[c] * a.plane_stride;
    float acc[16];

    if (a.mode[c] == 0) {
#pragma unroll
        for (int o = 0; o < 16; ++o) {
            const int gy = ty0 + ly0 + o;
            acc[o] = (gx < W && gy < H) ? src[(size_t)gy * W + gx] : 0.0f;
        }
    } else {
        const int k = a.k, kp = a.kp, rad = k / 2;
        const int cols = PITCH ? PITCH : TW + k - 1, rows = TH + k - 1;
        float *tile = smem;
        float *wsm = smem + ((rows * cols + 3) / 4) * 4;
        fill_tile(tile, src, rows, TW + k - 1, ty0 - rad, tx0 - rad, H, W, NT, cols);
        const float *__restrict__ wbase = a.kern[c];
        if (W_SMEM) {
            for (int idx = threadIdx.x; idx < k * kp; idx += NT) wsm[idx] = __ldg(wbase + idx);
            wbase = wsm;
        }
        __syncthreads();
#pragma unroll
        for (int o = 0; o < 16; ++o) acc[o] = 0.0f;
        const int nfull = k >> 4, rem = k & 15;
        for (int j = 0; j < k; ++j) {
            const float *col = tile + ly0 * cols + lx + j;
            float v[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) v[u] = col[u * cols];
            const float *rowp = col + 16 * cols;          // next tile row to enter the window
            const float4 *wp = reinterpret_cast<const float4 *>(wbase + j * kp);
            for (int b = 0; b < nfull; ++b) {             // 16 kernel rows per trip, no guards
#pragma unroll
                for (int u4 = 0; u4 < 4; ++u4) {
                    const float4 w4 = *wp++;
                    conv_step<0>(acc, v, w4.x, u4 * 4 + 0, rowp, cols);
                    conv_step<0>(acc, v, w4.y, u4 * 4 + 1, rowp, cols);
                    conv_step<0>(acc, v, w4.z, u4 * 4 + 2, rowp, cols);
                    conv_step<0>(acc, v, w4.w, u4 * 4 + 3, rowp, cols);
                }
            }
            switch (rem) {                                // k is odd: 1 <= rem <= 15
#define R2F_REM(R) case R: conv_tail<R>(acc, v, wp, rowp, cols); break;
                R2F_REM(1) R2F_REM(2) R2F_REM(3) R2F_REM(4) R2F_REM(5) R2F_REM(6) R2F_REM(7) R2F_REM(8)
                R2F_REM(9) R2F_REM(10) R2F_REM(11) R2F_REM(12) R2F_REM(13) R2F_REM(14) R2F_REM(15)
#undef R2F_REM
                default: break;
            }
        }
    }

    const size_t ps = a.plane_stride;
#pragma unroll
    for (int o = 0; o < 16; ++o) {
        const int gy = ty0 + ly0 + o;
        if (gx < W && gy < H) {
            const size_t idx = (size_t)gy * W + gx;
            float val = acc[o];
            if (a.epi == EPI_DENSITY) {
                val = density_eval<true>(a.curve, c, val, a.eps);
            } else if (a.epi == EPI_DENSITY_FAST) {
                val = density_eval_fast<true>(a.curve, c, val, a.eps);
            } else if (a.epi == EPI_GRAIN) {
                const float d = a.aux[c * ps + idx];
                const float g = val * curve_eval_any(a.curve, c, d);
                val = d + g;
                val = val > 0.0f ? val : 0.0f;
            }
            a.out[c * ps + idx] = val;
        }
    }
}

// Two alternative inner loops were built, measured at 24 MP / 17x17x3 and rejected (all 1.04-1.06 ms):
//  * horizontal 16-pixel strips with 128-bit window and weight loads (13 LDS per 272 FFMA): the vector
//    loads pin the window to consecutive registers and the FFMA register-bank conflicts cost what the
//    saved LDS gained;
//  * weights through the constant bank (kernel parameters, uniform-register FFMA operand): 62 vs 54
//    TFLOP/s in an isolated loop (scratch microbenchmark), but inside the kernel the extra guards and
//    ULDCs kept the issue slots just as busy (88 %).
// The kernel is issue-slot bound: ~66 % of its instructions are FFMA at ~84-88 % issue utilisation.

template <int TW, int TH, bool W_SMEM, int PITCH>
static cudaError_t launch_conv_cfg(const ConvArgs &a, size_t smem_bytes, cudaStream_t st) {
    auto kfn = k_conv2d<TW, TH, W_SMEM, PITCH>;
    cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (e != cudaSuccess) return e;
    dim3 grid((a.W + TW - 1) / TW, (a.H + TH - 1) / TH, 3);
    kfn<<<grid, (TW / 32) * (TH / 16) * 32, smem_bytes, st>>>(a);
    return cudaGetLastError();
}

constexpr size_t kMaxDynSmem = 227 * 1024;

static size_t conv_smem_bytes(int tw, int th, int k, int kp, bool w_smem, int pitch = 0) {
    const size_t cols = pitch ? (size_t)pitch : (size_t)(tw + k - 1);
    size_t tile = (cols * (th + k - 1) + 3) / 4 * 4;
    return (tile + (w_smem ? (size_t)k * kp : 0)) * sizeof(float);
}

cudaError_t launch_conv2d(const ConvArgs &a, cudaStream_t st) {
    const bool any_conv = a.mode[0] || a.mode[1] || a.mode[2];
    if (!any_conv) return launch_conv_cfg<64, 64, true, 0>(a, 16, st);
    size_t s;
    if (64 + a.k - 1 <= 128) {  // fixed 128-float pitch: immediate-offset window loads
        s = conv_smem_bytes(64, 64, a.k, a.kp, true, 128);
        if (s <= 100 * 1024) return launch_conv_cfg<64, 64, true, 128>(a, s, st);
    }
    s = conv_smem_bytes(64, 64, a.k, a.kp, true);
    if (s <= kMaxDynSmem) return launch_conv_cfg<64, 64, true, 0>(a, s, st);
    s = conv_smem_bytes(64, 64, a.k, a.kp, false);
    if (s <= kMaxDynSmem) return launch_conv_cfg<64, 64, false, 0>(a, s, st);
    s = conv_smem_bytes(32, 32, a.k, a.kp, false);
    if (s <= kMaxDynSmem) return launch_conv_cfg<32, 32, false, 0>(a, s, st);
    return cudaErrorInvalidValue;  // kernel wider than ~205 taps: needs the FFT path
}

// ------------------------------------------------------------------------------------------
// a8 + a9 + a10: burn apply, tetrahedral LUT, quantise (planar density -> interleaved output)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float burn_sample(const BurnArgs &b, int y, int x) {
    // scipy.ndimage.zoom(order=1, grid_mode=False) coordinate map, then edge pad / crop
    // (reference effects.py:381-385)
    const int yy = y < b.zh ? y : b.zh - 1, xx = x < b.zw ? x : b.zw - 1;
    const double cy = b.zh > 1 ? (double)yy * ((double)(b.lh - 1) / (double)(b.zh - 1)) : 0.0;
    const double cx = b.zw > 1 ? (double)xx * ((double)(b.lw - 1) / (double)(b.zw - 1)) : 0.0;
    int y0 = (int)floor(cy), x0 = (int)floor(cx);
    if (y0 > b.lh - 1) y0 = b.lh - 1;
    if (x0 > b.lw - 1) x0 = b.lw - 1;
    const double fy = cy - y0, fx = cx - x0;
    const int y1 = y0 + 1 < b.lh ? y0 + 1 : y0, x1 = x0 + 1 < b.lw ? x0 + 1 : x0;
    const double v00 = b.map[y0 * b.lw + x0], v01 = b.map[y0 * b.lw + x1];
    const double v10 = b.map[y1 * b.lw + x0], v11 = b.map[y1 * b.lw + x1];
    const double top = v00 * (1.0 - fx) + v01 * fx, bot = v10 * (1.0 - fx) + v11 * fx;
    return (float)(top * (1.0 - fy) + bot * fy);
}

template <bool F32_OUT>
__global__ void __launch_bounds__(kThreads)
k_finish(const float *__restrict__ in, size_t plane_stride, size_t npix, int W, const __grid_constant__ Lut3D l3,
         const __grid_constant__ BurnArgs burn, uint8_t *__restrict__ out_u8, float *__restrict__ out_f32,
         int f32_stage_rgb, const __grid_constant__ FastTetra ft) {
    const size_t nquad = (npix + 3) / 4;
    const size_t stride = (size_t)gridDim.x * kThreads;
    for (size_t q = (size_t)blockIdx.x * kThreads + threadIdx.x; q < nquad; q += stride) {
        float d[3][4];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float4 v = __ldcs(reinterpret_cast<const float4 *>(in + c * plane_stride) + q);
            d[c][0] = v.x; d[c][1] = v.y; d[c][2] = v.z; d[c][3] = v.w;
        }
        uint32_t b[12];
        float f[12];
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            float d0 = d[0][p], d1 = d[1][p], d2 = d[2][p];
            if (burn.map != nullptr) {
                const size_t pix = q * 4 + p;
                const int y = (int)(pix / (size_t)W), x = (int)(pix - (size_t)y * W);
                const float m = burn.strength * burn_sample(burn, y, x);
                d0 = d0 - m; d1 = d1 - m; d2 = d2 - m;
                d0 = d0 > 0.f ? d0 : 0.f; d1 = d1 > 0.f ? d1 : 0.f; d2 = d2 > 0.f ? d2 : 0.f;
            }
            if (F32_OUT && !f32_stage_rgb) {
                f[3 * p] = d0; f[3 * p + 1] = d1; f[3 * p + 2] = d2;
            } else if (F32_OUT) {
                tetra_eval(l3, d0, d1, d2, f[3 * p], f[3 * p + 1], f[3 * p + 2]);
            } else {
                // conversion-free guarded float32 tail (fast_chain.cuh) with the exact path for undecided pixels and
                // for negative / NaN densities
                const uint32_t px = tetra_u8<false>(ft, l3, d0, d1, d2);
                b[3 * p] = px & 255u; b[3 * p + 1] = (px >> 8) & 255u; b[3 * p + 2] = px >> 16;
            }
        }
        if (q * 4 + 4 <= npix) {
            if (F32_OUT) {
                float4 *o = reinterpret_cast<float4 *>(out_f32) + 3 * q;
                o[0] = make_float4(f[0], f[1], f[2], f[3]);
                o[1] = make_float4(f[4], f[5], f[6], f[7]);
                o[2] = make_float4(f[8], f[9], f[10], f[11]);
            } else {
                store_quad_u8(out_u8, q, b);
            }
        } else {
            for (size_t p = q * 4; p < npix; ++p) {
                const int pp = (int)(p - q * 4);
                for (int c = 0; c < 3; ++c) {
                    if (F32_OUT) out_f32[p * 3 + c] = f[3 * pp + c];
                    else out_u8[p * 3 + c] = (uint8_t)b[3 * pp + c];
                }
            }
        }
    }
}

cudaError_t launch_finish(Planes in, size_t npix, int H, int W, const Lut3D &l3, const BurnArgs &burn, uint8_t *out_u8,
                          float *out_f32, int f32_stage_rgb, int num_sms, cudaStream_t st, const FastTetra &ft) {
    (void)H;
    const int grid = grid_for((npix + 3) / 4, num_sms, 8);
    if (out_f32 != nullptr)
        k_finish<true><<<grid, kThreads, 0, st>>>(in.base, in.plane_stride, npix, W, l3, burn, nullptr, out_f32,
                                                 f32_stage_rgb, ft);
    else
        k_finish<false><<<grid, kThreads, 0, st>>>(in.base, in.plane_stride, npix, W, l3, burn, out_u8, nullptr, 1, ft);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// layout shuffles (taps, injected noise, stage-level entry points)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
k_planar_to_interleaved(const float *__restrict__ in, size_t plane_stride, float *__restrict__ out, size_t npix) {
    const size_t stride = (size_t)gridDim.x * kThreads;
    for (size_t p = (size_t)blockIdx.x * kThreads + threadIdx.x; p < npix; p += stride) {
#pragma unroll
        for (int c = 0; c < 3; ++c) out[p * 3 + c] = in[c * plane_stride + p];
    }
}

__global__ void __launch_bounds__(kThreads)
k_interleaved_to_planar(const float *__restrict__ in, int cin, int nch, float *__restrict__ out, size_t plane_stride,
                        size_t npix) {
    const size_t stride = (size_t)gridDim.x * kThreads;
    for (size_t p = (size_t)blockIdx.x * kThreads + threadIdx.x; p < npix; p += stride) {
        for (int c = 0; c < nch; ++c) out[c * plane_stride + p] = in[p * cin + c];
    }
}

// ------------------------------------------------------------------------------------------
// chroma noise reduction (reference effects.py:421-561; SURVEY 8f-3, runs BEFORE the render path)
//   k_cnr_rows : XYZ -> (x, y, Y) (effects.py:497-519) and the horizontal pass of the edge-clamped 1-D Gaussian
//                on x and y (effects.py:438-460), fused: a CTA turns a row segment plus its tap halo into
//                chromaticities in shared memory and blurs from there; Y goes to its plane
//   k_cnr_cols : vertical pass (effects.py:463-482) from a shared-memory tile of the row-blurred planes, then
//                xyY -> XYZ (effects.py:522-544)
// float32 products accumulated in binary64 in tap order, one rounding per pass, like the numba loops.
// (the first version was three per-pixel kernels with per-tap global loads: 0.83 ms at 24 MP, size 3.)
// ------------------------------------------------------------------------------------------
constexpr int kCnrSeg = 512;      // outputs per CTA of the row kernel (two per thread)
constexpr int kCnrMaxHalf = 24;   // chroma_nr <= 20 in the reference GUI (slider 0..10, doubled on export)
constexpr int kCnrTileW = 64, kCnrTileH = 32;
// the Gaussian taps travel as a kernel parameter: the tap index is warp-uniform, so they reach the multiplier through
// the uniform datapath instead of a shared-memory broadcast per tap
struct CnrTaps {
    float w[2 * kCnrMaxHalf + 1];
};

__global__ void __launch_bounds__(kThreads)
k_cnr_rows(const float *__restrict__ in, int cin, float *__restrict__ tmp, float *__restrict__ yplane, size_t ps, int H,
           int W, const __grid_constant__ CnrTaps taps, int half) {
    __shared__ float sx[kCnrSeg + 2 * kCnrMaxHalf], sy[kCnrSeg + 2 * kCnrMaxHalf];
    const int y = blockIdx.y, x0 = blockIdx.x * kCnrSeg;
    const int n = min(kCnrSeg, W - x0), span = n + 2 * half;
    const float *row = in + (size_t)y * W * cin;
    for (int i = threadIdx.x; i < span; i += kThreads) {
        const int gx = min(max(x0 - half + i, 0), W - 1);  // edge clamp of the blur = chromaticity of the clamped pixel
        const float X = row[(size_t)gx * cin], Y = row[(size_t)gx * cin + 1], Z = row[(size_t)gx * cin + 2];
        const float denom = (X + Y) + Z;
        const bool ok = denom > 1e-8f;
        sx[i] = ok ? __fdiv_rn(X, denom) : 0.0f;
        sy[i] = ok ? __fdiv_rn(Y, denom) : 0.0f;
        if (i >= half && i < half + n) yplane[(size_t)y * W + x0 + i - half] = Y;
    }
    __syncthreads();
    for (int o = threadIdx.x; o < n; o += kThreads) {
        double ax = 0.0, ay = 0.0;
        for (int t = 0; t <= 2 * half; ++t) {
            const float w = taps.w[t];
            ax += (double)(sx[o + t] * w);  // float32 product, binary64 sum
            ay += (double)(sy[o + t] * w);
        }
        const size_t idx = (size_t)y * W + x0 + o;
        tmp[idx] = (float)ax;
        tmp[ps + idx] = (float)ay;
    }
}

__global__ void __launch_bounds__(kThreads)
k_cnr_cols(const float *__restrict__ tmp, const float *__restrict__ yplane, float *__restrict__ out, size_t ps, int H,
           int W, const __grid_constant__ CnrTaps taps, int half) {
    extern __shared__ __align__(16) float tile[];  // [2][kCnrTileH + 2*half][kCnrTileW]
    const int x0 = blockIdx.x * kCnrTileW, y0 = blockIdx.y * kCnrTileH;
    const int rows = kCnrTileH + 2 * half;
    for (int i = threadIdx.x; i < 2 * rows * kCnrTileW; i += kThreads) {
        const int c = i / (rows * kCnrTileW), r = (i / kCnrTileW) % rows, col = i % kCnrTileW;
        const int gy = min(max(y0 - half + r, 0), H - 1), gx = min(x0 + col, W - 1);
        tile[i] = tmp[c * ps + (size_t)gy * W + gx];
    }
    __syncthreads();
    const float *tx = tile, *ty = tile + rows * kCnrTileW;
    for (int i = threadIdx.x; i < kCnrTileH * kCnrTileW; i += kThreads) {
        const int r = i / kCnrTileW, col = i % kCnrTileW;
        const int gy = y0 + r, gx = x0 + col;
        if (gy >= H || gx >= W) continue;
        double ax = 0.0, ay = 0.0;
        for (int t = 0; t <= 2 * half; ++t) {
            const float w = taps.w[t];
            ax += (double)(tx[(r + t) * kCnrTileW + col] * w);
            ay += (double)(ty[(r + t) * kCnrTileW + col] * w);
        }
        const float cx = (float)ax, cy = (float)ay;
        const size_t p = (size_t)gy * W + gx;
        const float Y = yplane[p];
        float X = 0.0f, Yo = 0.0f, Z = 0.0f;
        if (cy > 1e-8f) {
            const float inv = __fdiv_rn(Y, cy);
            X = cx * inv;
            Yo = Y;
            Z = (float)(((1.0 - (double)cx) - (double)cy) * (double)inv);  // binary64 like numba's typing
        }
        out[p * 3] = X;
        out[p * 3 + 1] = Yo;
        out[p * 3 + 2] = Z;
    }
}

cudaError_t launch_chroma_nr(const float *in, int cin, float *out, int H, int W, const float *taps_host, int ntaps,
                             float *ws /* 3 planes */, size_t ps, int num_sms, cudaStream_t st) {
    (void)num_sms;
    const int half = ntaps / 2;
    if (half > kCnrMaxHalf) return cudaErrorInvalidValue;
    CnrTaps taps_dev{};
    for (int i = 0; i < ntaps; ++i) taps_dev.w[i] = taps_host[i];
    float *tmp = ws, *yplane = ws + 2 * ps;
    k_cnr_rows<<<dim3((W + kCnrSeg - 1) / kCnrSeg, H), kThreads, 0, st>>>(in, cin, tmp, yplane, ps, H, W, taps_dev, half);
    const size_t smem = (size_t)2 * (kCnrTileH + 2 * half) * kCnrTileW * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(k_cnr_cols, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k_cnr_cols<<<dim3((W + kCnrTileW - 1) / kCnrTileW, (H + kCnrTileH - 1) / kCnrTileH), kThreads, smem, st>>>(
        tmp, yplane, out, ps, H, W, taps_dev, half);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// RGB histogram counts (reference utils.py:158-169 / shaders/histogram.wgsl pass 1):
// 3 x 256 bins over the uint8 output.  Each warp counts into its own shared-memory copy (no
// inter-warp atomic contention), copies are reduced and added to the global bins once per CTA.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
k_histogram(const uint8_t *__restrict__ img, size_t npix, unsigned int *__restrict__ counts /* [3][256] */) {
    __shared__ unsigned int sh[kThreads / 32][768];
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (kThreads / 32) * 768; i += kThreads) (&sh[0][0])[i] = 0u;
    __syncthreads();
    // 4 pixels = 12 bytes = three 32-bit words per iteration
    const size_t nquad = npix / 4, stride = (size_t)gridDim.x * kThreads;
    const uint32_t *w32 = reinterpret_cast<const uint32_t *>(img);
    unsigned int *my = sh[warp];
    for (size_t q = (size_t)blockIdx.x * kThreads + threadIdx.x; q < nquad; q += stride) {
        const uint32_t a = __ldg(w32 + 3 * q), b = __ldg(w32 + 3 * q + 1), c = __ldg(w32 + 3 * q + 2);
        const uint32_t bytes[12] = {a & 255u, (a >> 8) & 255u, (a >> 16) & 255u, a >> 24, b & 255u, (b >> 8) & 255u,
                                    (b >> 16) & 255u, b >> 24, c & 255u, (c >> 8) & 255u, (c >> 16) & 255u, c >> 24};
#pragma unroll
        for (int i = 0; i < 12; ++i) atomicAdd(&my[(i % 3) * 256 + bytes[i]], 1u);
    }
    if (blockIdx.x == 0 && threadIdx.x < (npix & 3) * 3) {
        const size_t i = nquad * 12 + threadIdx.x;
        atomicAdd(&my[(threadIdx.x % 3) * 256 + img[i]], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 768; i += kThreads) {
        unsigned int t = 0;
#pragma unroll
        for (int w = 0; w < kThreads / 32; ++w) t += sh[w][i];
        if (t) atomicAdd(counts + i, t);
    }
}

cudaError_t launch_histogram(const uint8_t *img, size_t npix, unsigned int *counts_dev, int num_sms, cudaStream_t st) {
    cudaError_t e = cudaMemsetAsync(counts_dev, 0, 768 * sizeof(unsigned int), st);
    if (e != cudaSuccess) return e;
    k_histogram<<<grid_for(npix / 4 + 1, num_sms, 4), kThreads, 0, st>>>(img, npix, counts_dev);
    return cudaGetLastError();
}

// Passes 2 and 3 of the histogram widget (reference utils.py:171-223; shaders/histogram.wgsl:63-158), one CTA:
// float32 log1p(count / peak), 3-tap moving average with replicated ends, scaling to `height`, truncation, then the
// (2,2,2,4) colour-mix lookup per widget pixel.  Same float32 operations as the NumPy code (log1p evaluated in
// binary64 and rounded once, i.e. a correctly rounded log1pf).
__global__ void __launch_bounds__(768)
k_histogram_image(const unsigned int *__restrict__ counts, int height, uint32_t mix0, uint32_t mix1, uint32_t mix2,
                  uint32_t mix3, uint32_t mix4, uint32_t mix5, uint32_t mix6, uint32_t mix7,
                  uint8_t *__restrict__ out /* height x 256 x 4 */) {
    __shared__ float f[768];
    __shared__ float red[24];
    __shared__ int hts[768];
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    auto block_max = [&](float v) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
        __syncthreads();
        if (lane == 0) red[warp] = v;
        __syncthreads();
        float m = red[0];
        for (int i = 1; i < 24; ++i) m = fmaxf(m, red[i]);
        return m;
    };
    const float c = (float)counts[t];
    float peak = fmaxf(block_max(c), 0.0f);
    if (peak == 0.0f) peak = 1.0f;
    f[t] = (float)log1p((double)__fdiv_rn(c, peak));
    __syncthreads();
    const int ch = t >> 8, bin = t & 255;
    const float left = f[ch * 256 + max(bin - 1, 0)], right = f[ch * 256 + min(bin + 1, 255)];
    const float smooth = __fdiv_rn((left + f[t]) + right, 3.0f);
    float top = fmaxf(block_max(smooth), 0.0f);
    if (top == 0.0f) top = 1.0f;
    hts[t] = (int)__fdiv_rn(smooth * (float)height, top);
    __syncthreads();
    const uint32_t mix[8] = {mix0, mix1, mix2, mix3, mix4, mix5, mix6, mix7};
    uint32_t *o32 = reinterpret_cast<uint32_t *>(out);
    for (int p = t; p < height * 256; p += 768) {
        const int y = p >> 8, x = p & 255;
        const int a0 = y >= height - hts[x], a1 = y >= height - hts[256 + x], a2 = y >= height - hts[512 + x];
        o32[p] = mix[a0 * 4 + a1 * 2 + a2];
    }
}

cudaError_t launch_histogram_image(const unsigned int *counts_dev, int height, const uint8_t *mix_host, uint8_t *out,
                                   cudaStream_t st) {
    uint32_t m[8];
    for (int i = 0; i < 8; ++i)
        m[i] = (uint32_t)mix_host[4 * i] | ((uint32_t)mix_host[4 * i + 1] << 8) | ((uint32_t)mix_host[4 * i + 2] << 16) |
               ((uint32_t)mix_host[4 * i + 3] << 24);
    k_histogram_image<<<1, 768, 0, st>>>(counts_dev, height, m[0], m[1], m[2], m[3], m[4], m[5], m[6], m[7], out);
    return cudaGetLastError();
}

// Canvas border (reference effects.py:338-357 add_canvas): fill the canvas with one colour and
// paste the rendered image at (off_y, off_x).  One thread per canvas byte triple.
__global__ void __launch_bounds__(kThreads)
k_canvas_paste(const uint8_t *__restrict__ src, int H, int W, uint8_t *__restrict__ dst, int CH, int CW, int off_y,
               int off_x, uchar3 colour) {
    const size_t total = (size_t)CH * CW;
    const size_t stride = (size_t)gridDim.x * kThreads;
    for (size_t p = (size_t)blockIdx.x * kThreads + threadIdx.x; p < total; p += stride) {
        const int y = (int)(p / CW), x = (int)(p - (size_t)y * CW);
        const int sy = y - off_y, sx = x - off_x;
        uchar3 v = colour;
        if (sy >= 0 && sy < H && sx >= 0 && sx < W) {
            const uint8_t *q = src + ((size_t)sy * W + sx) * 3;
            v = make_uchar3(q[0], q[1], q[2]);
        }
        uint8_t *d = dst + p * 3;
        d[0] = v.x; d[1] = v.y; d[2] = v.z;
    }
}

cudaError_t launch_canvas_paste(const uint8_t *src, int H, int W, uint8_t *dst, int CH, int CW, int off_y, int off_x,
                                int r, int g, int b, int num_sms, cudaStream_t st) {
    k_canvas_paste<<<grid_for((size_t)CH * CW, num_sms, 8), kThreads, 0, st>>>(
        src, H, W, dst, CH, CW, off_y, off_x, make_uchar3((unsigned char)r, (unsigned char)g, (unsigned char)b));
    return cudaGetLastError();
}

// Presentation blit (reference shaders/copy_to_int.wgsl:18-51, geometry gpu_processor.py:1416-1539): the rendered
// image scaled into a widget-sized RGBA8 buffer with letterbox / canvas areas.  Destination pixel centre ->
// normalised source uv; inside [0,1]^2: bilinear sample (texel centres at +0.5, clamp to edge), alpha 255; else
// inside the canvas rectangle: canvas colour, alpha 255; else transparent black.
__global__ void __launch_bounds__(kThreads)
k_present(const uint8_t *__restrict__ src, int H, int W, uint8_t *__restrict__ dst, int DH, int DW, PresentArgs u) {
    const size_t total = (size_t)DH * DW, stride = (size_t)gridDim.x * kThreads;
    uint32_t *o32 = reinterpret_cast<uint32_t *>(dst);
    for (size_t p = (size_t)blockIdx.x * kThreads + threadIdx.x; p < total; p += stride) {
        const int y = (int)(p / DW), x = (int)(p - (size_t)y * DW);
        const float dx = (float)x + 0.5f, dy = (float)y + 0.5f;
        const float su = (dx - u.offset_x) * u.scale_x, sv = (dy - u.offset_y) * u.scale_y;
        uint32_t px = 0u;
        if (su >= 0.0f && su <= 1.0f && sv >= 0.0f && sv <= 1.0f) {
            const float fx = su * (float)W - 0.5f, fy = sv * (float)H - 0.5f;
            const float x0f = floorf(fx), y0f = floorf(fy);
            const float ax = fx - x0f, ay = fy - y0f;
            const int x0 = min(max((int)x0f, 0), W - 1), x1 = min(max((int)x0f + 1, 0), W - 1);
            const int y0 = min(max((int)y0f, 0), H - 1), y1 = min(max((int)y0f + 1, 0), H - 1);
            const uint8_t *r0 = src + (size_t)y0 * W * 3, *r1 = src + (size_t)y1 * W * 3;
            px = 0xff000000u;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float top = (float)r0[x0 * 3 + c] * (1.0f - ax) + (float)r0[x1 * 3 + c] * ax;
                const float bot = (float)r1[x0 * 3 + c] * (1.0f - ax) + (float)r1[x1 * 3 + c] * ax;
                const int v = __float2int_rn(top * (1.0f - ay) + bot * ay);
                px |= (uint32_t)min(max(v, 0), 255) << (8 * c);
            }
        } else if (dx >= u.canvas_min_x && dx <= u.canvas_max_x && dy >= u.canvas_min_y && dy <= u.canvas_max_y) {
            px = 0xff000000u | (uint32_t)u.r | ((uint32_t)u.g << 8) | ((uint32_t)u.b << 16);
        }
        o32[p] = px;
    }
}

cudaError_t launch_present(const uint8_t *src, int H, int W, uint8_t *dst, int DH, int DW, const PresentArgs &u,
                           int num_sms, cudaStream_t st) {
    k_present<<<grid_for((size_t)DH * DW, num_sms, 8), kThreads, 0, st>>>(src, H, W, dst, DH, DW, u);
    return cudaGetLastError();
}

cudaError_t launch_planar_to_interleaved(Planes in, float *out, size_t npix, int num_sms, cudaStream_t st) {
    k_planar_to_interleaved<<<grid_for(npix, num_sms, 8), kThreads, 0, st>>>(in.base, in.plane_stride, out, npix);
    return cudaGetLastError();
}

cudaError_t launch_interleaved_to_planar(const float *in, int cin, int nch, Planes out, size_t npix, int num_sms,
                                         cudaStream_t st) {
    k_interleaved_to_planar<<<grid_for(npix, num_sms, 8), kThreads, 0, st>>>(in, cin, nch, out.base, out.plane_stride,
                                                                            npix);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// a7: white Gaussian noise.  Counter-based Philox4x32-10: counter = (quad index, channel),
// key = seed; 4 uniforms -> 2 Box-Muller pairs -> 4 normals for 4 consecutive pixels.
// (reference GPU path: PCG-3D hash + Box-Muller, shaders/noise.wgsl:14-62; streams differ by design)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
k_noise(float *__restrict__ out, size_t plane_stride, int nch, int H, int W, uint32_t k0, uint32_t k1, int shift) {
    const int qw = (W + shift + 3) >> 2;  // quads of the shifted grid that touch columns 0 .. W-1
    const size_t per_ch = (size_t)H * qw, total = per_ch * nch;
    const size_t stride = (size_t)gridDim.x * kThreads;
    for (size_t t = (size_t)blockIdx.x * kThreads + threadIdx.x; t < total; t += stride) {
        const int ch = (int)(t / per_ch);
        const size_t rem = t - (size_t)ch * per_ch;
        const int y = (int)(rem / qw), qx = (int)(rem - (size_t)y * qw);
        const float4 v = noise_quad(qx, y, ch, k0, k1);
        float *row = out + (size_t)ch * plane_stride + (size_t)y * W;
        const float vals[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int l = 0; l < 4; ++l) {
            const int x = 4 * qx + l - shift;
            if (x >= 0 && x < W) row[x] = vals[l];
        }
    }
}

cudaError_t launch_noise(Planes out, int nch, int H, int W, uint64_t seed, int shift, int num_sms, cudaStream_t st) {
    const size_t items = (size_t)H * ((W + shift + 3) / 4) * nch;
    k_noise<<<grid_for(items, num_sms, 8), kThreads, 0, st>>>(out.base, out.plane_stride, nch, H, W, (uint32_t)seed,
                                                             (uint32_t)(seed >> 32), shift);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// a7 + a8 + a9 + a10 fused: grain (noise regenerated per tile, correlated with the grain kernel,
// scaled by the density-dependent amplitude, added, clipped), burn, tetrahedral LUT, quantise.
// Replaces k_noise + k_conv2d(EPI_GRAIN) + k_finish: the noise field and the grained density never
// touch HBM (reads 12 B/px of density, writes 3 B/px).
// Same thread mapping as k_conv2d: 64x64 tile, 256 threads, a 16-row strip per thread.
// ------------------------------------------------------------------------------------------
constexpr int kGfTile = 64;

template <bool GEN>
__global__ void __launch_bounds__(256, 2)
k_grain_finish(const __grid_constant__ GrainFinishArgs a) {
    extern __shared__ __align__(16) float smem[];
    const int H = a.H, W = a.W, k = a.k, kp = a.kp, rad = k / 2;
    const int cols = kGfTile + k - 1, rows = kGfTile + k - 1;
    float *tile = smem;                                   // noise tile; later the byte staging area
    float *wsm = smem + ((rows * cols + 3) / 4) * 4;      // grain kernel
    float *priv = wsm + k * kp;                           // [3][16][256] grained densities, thread-private slots
    const int tx0 = blockIdx.x * kGfTile, ty0 = (blockIdx.y + a.tile_y0) * kGfTile;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int lx = (warp & 1) * 32 + lane, ly0 = (warp >> 1) * 16;
    const int gx = tx0 + lx;
    const size_t ps = a.plane_stride;
    for (int idx = threadIdx.x; idx < k * kp; idx += 256) wsm[idx] = __ldg(a.gk + idx);

    float g[16];
    const int nch = a.bw ? 1 : 3;
    const int xs = tx0 - rad;                                  // global x of tile column 0
    const bool interior = xs >= 0 && xs + cols <= W;           // no horizontal reflection needed
    const int q0 = (xs + a.noise_shift) >> 2, nq = (cols + 3) >> 2;  // quads of the shifted grid covering the tile row
#pragma unroll 1
    for (int c = 0; c < 3; ++c) {
        // issue this channel's 16 density loads first: their latency hides behind the noise generation
        // and the correlation below
        const float *dplane = a.dens + c * ps;
        float dval[16];
#pragma unroll
        for (int o = 0; o < 16; ++o) {
            const int gy = ty0 + ly0 + o;
            dval[o] = (gx < W && gy < H) ? __ldcs(dplane + (size_t)gy * W + gx) : 0.0f;
        }
        if (c < nch) {
            __syncthreads();  // previous channel's window reads are done
            if (GEN) {
                if (interior) {  // one Philox call per quad of four samples (quads start on tile columns 0, 4, ...)
                    for (int idx = threadIdx.x; idx < rows * nq; idx += 256) {
                        const int ty = idx / nq, tq = idx - ty * nq;
                        const int gy = reflect101(ty0 - rad + ty, H);
                        const float4 v = noise_quad((uint32_t)(q0 + tq), gy, c, a.seed_lo, a.seed_hi);
                        const float vals[4] = {v.x, v.y, v.z, v.w};
                        const int tx = 4 * tq;
#pragma unroll
                        for (int l = 0; l < 4; ++l)
                            if (tx + l < cols) tile[ty * cols + tx + l] = vals[l];
                    }
                } else {
                    for (int idx = threadIdx.x; idx < rows * cols; idx += 256) {
                        const int ty = idx / cols, tx = idx - ty * cols;
                        tile[idx] = noise_at(reflect101(tx0 - rad + tx, W), reflect101(ty0 - rad + ty, H), c, a.seed_lo,
                                             a.seed_hi, a.noise_shift);
                    }
                }
            } else {
                fill_tile(tile, a.noise + (size_t)c * ps, rows, cols, ty0 - rad, xs, H, W, 256);
            }
            __syncthreads();
#pragma unroll
            for (int o = 0; o < 16; ++o) g[o] = 0.0f;
            const int nfull = k >> 4, rem = k & 15;
            for (int j = 0; j < k; ++j) {
                const float *col = tile + ly0 * cols + lx + j;
                float v[16];
#pragma unroll
                for (int u = 0; u < 16; ++u) v[u] = col[u * cols];
                const float *rowp = col + 16 * cols;
                const float4 *wp = reinterpret_cast<const float4 *>(wsm + j * kp);
                for (int b = 0; b < nfull; ++b) {
#pragma unroll
                    for (int u4 = 0; u4 < 4; ++u4) {
                        const float4 w4 = *wp++;
                        conv_step<0>(g, v, w4.x, u4 * 4 + 0, rowp, cols);
                        conv_step<0>(g, v, w4.y, u4 * 4 + 1, rowp, cols);
                        conv_step<0>(g, v, w4.z, u4 * 4 + 2, rowp, cols);
                        conv_step<0>(g, v, w4.w, u4 * 4 + 3, rowp, cols);
                    }
                }
                switch (rem) {
#define R2F_REM(R) case R: conv_tail<R>(g, v, wp, rowp, cols); break;
                    R2F_REM(1) R2F_REM(2) R2F_REM(3) R2F_REM(4) R2F_REM(5) R2F_REM(6) R2F_REM(7) R2F_REM(8)
                    R2F_REM(9) R2F_REM(10) R2F_REM(11) R2F_REM(12) R2F_REM(13) R2F_REM(14) R2F_REM(15)
#undef R2F_REM
                    default: break;
                }
            }
        }
        // grain apply on channel c (black-and-white grain reuses the single field)
#pragma unroll
        for (int o = 0; o < 16; ++o) {
            const float d = dval[o];
            float val = d + g[o] * curve_eval_any(a.gcurve, c, d);
            val = val > 0.0f ? val : 0.0f;
            priv[(c * 16 + o) * 256 + threadIdx.x] = val;
        }
    }
    __syncthreads();  // all noise-tile reads done: its storage becomes the output staging area
    uint8_t *stage = reinterpret_cast<uint8_t *>(tile);
    // burn, tetrahedral LUT, quantise; stage the tile's bytes so rows leave as 16-byte stores
#pragma unroll 4
    for (int o = 0; o < 16; ++o) {
        const int gy = ty0 + ly0 + o;
        float d0 = priv[(0 * 16 + o) * 256 + threadIdx.x], d1 = priv[(1 * 16 + o) * 256 + threadIdx.x];
        float d2 = priv[(2 * 16 + o) * 256 + threadIdx.x];
        if (a.burn.map != nullptr && gx < W && gy < H) {
            const float m = a.burn.strength * burn_sample(a.burn, gy, gx);
            d0 = d0 - m; d1 = d1 - m; d2 = d2 - m;
            d0 = d0 > 0.f ? d0 : 0.f; d1 = d1 > 0.f ? d1 : 0.f; d2 = d2 > 0.f ? d2 : 0.f;
        }
        uint32_t q0b, q1b, q2b;
        tetra_quant_u8(a.l3, d0, d1, d2, q0b, q1b, q2b);
        uint8_t *sp = stage + ((ly0 + o) * kGfTile + lx) * 3;
        sp[0] = (uint8_t)q0b; sp[1] = (uint8_t)q1b; sp[2] = (uint8_t)q2b;
    }
    __syncthreads();
    const int tw = min(kGfTile, W - tx0), th = min(kGfTile, H - ty0);
    const int row_bytes = tw * 3;
    if (tw == kGfTile && ((size_t)W * 3 % 16) == 0 && (reinterpret_cast<uintptr_t>(a.out_u8) & 15) == 0) {
        for (int idx = threadIdx.x; idx < th * 12; idx += 256) {  // 192 bytes = 12 x 16 per row
            const int r = idx / 12, s16 = idx - r * 12;
            const uint4 v = *reinterpret_cast<const uint4 *>(stage + r * 192 + s16 * 16);
            __stcs(reinterpret_cast<uint4 *>(a.out_u8 + ((size_t)(ty0 + r) * W + tx0) * 3) + s16, v);
        }
    } else {
        for (int idx = threadIdx.x; idx < th * row_bytes; idx += 256) {
            const int r = idx / row_bytes, bcol = idx - r * row_bytes;
            a.out_u8[((size_t)(ty0 + r) * W + tx0) * 3 + bcol] = stage[r * 192 + bcol];
        }
    }
}

cudaError_t launch_grain_finish(const GrainFinishArgs &a, cudaStream_t st) {
    const int ext = kGfTile + a.k - 1;
    // noise tile (>= 12 KB so it can hold the staged output bytes) + grain kernel + private slots
    size_t tile_f = ((size_t)ext * ext + 3) / 4 * 4;
    if (tile_f < (size_t)kGfTile * kGfTile * 3 / 4) tile_f = (size_t)kGfTile * kGfTile * 3 / 4;
    const size_t smem = (tile_f + (size_t)a.k * a.kp + 3 * 16 * 256) * sizeof(float);
    if (smem > kMaxDynSmem) return cudaErrorInvalidValue;
    dim3 grid((a.W + kGfTile - 1) / kGfTile, a.tile_rows > 0 ? a.tile_rows : (a.H + kGfTile - 1) / kGfTile);
    cudaError_t e;
    if (a.noise == nullptr) {
        if ((e = cudaFuncSetAttribute(k_grain_finish<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) !=
            cudaSuccess)
            return e;
        k_grain_finish<true><<<grid, 256, smem, st>>>(a);
    } else {
        if ((e = cudaFuncSetAttribute(k_grain_finish<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem)) != cudaSuccess)
            return e;
        k_grain_finish<false><<<grid, 256, smem, st>>>(a);
    }
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// a8: highlight-burn low-res mask (reference effects.py:360-389, 404-409)
//   1. cv2.resize(INTER_AREA) of the green density plane to (lh, lw)  -- area weights restated
//      from OpenCV's computeResizeAreaTab (fractional cells, 1e-3 thresholds)
//   2. max(x - d_ref, 0)
//   3. scipy.ndimage.gaussian_filter(sigma=3, truncate=2): 13 taps, 'reflect' borders,
//      axis 0 then axis 1, float32 between the passes
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float area_weight(int s, int dcell, double scale, int ssize) {
    const double f1 = dcell * scale, f2 = f1 + scale;
    const double cell = fmin(scale, (double)ssize - f1);
    int s1 = (int)ceil(f1), s2 = (int)floor(f2);
    if (s2 > ssize - 1) s2 = ssize - 1;
    if (s1 > s2) s1 = s2;
    if (s == s1 - 1 && (double)s1 - f1 > 1e-3) return (float)(((double)s1 - f1) / cell);
    if (s >= s1 && s < s2) return (float)(1.0 / cell);
    if (s == s2 && f2 - (double)s2 > 1e-3) return (float)(fmin(fmin(f2 - (double)s2, 1.0), cell) / cell);
    return 0.0f;
}

__global__ void __launch_bounds__(kThreads)
k_burn_down(const float *__restrict__ g, int H, int W, int lh, int lw, float d_ref, float *__restrict__ out) {
    const int dy = blockIdx.y, dx = blockIdx.x;
    const double sx = (double)W / lw, sy = (double)H / lh;
    const int x_lo = max((int)floor(dx * sx) - 1, 0), x_hi = min((int)ceil((dx + 1) * sx) + 1, W);
    const int y_lo = max((int)floor(dy * sy) - 1, 0), y_hi = min((int)ceil((dy + 1) * sy) + 1, H);
    const int bw = x_hi - x_lo, bh = y_hi - y_lo;
    float part = 0.0f;
    for (int i = threadIdx.x; i < bw * bh; i += kThreads) {
        const int y = y_lo + i / bw, x = x_lo + i % bw;
        const float wgt = area_weight(x, dx, sx, W) * area_weight(y, dy, sy, H);
        if (wgt != 0.0f) part = fmaf(wgt, g[(size_t)y * W + x], part);
    }
    __shared__ float red[kThreads / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.0f;
        for (int i = 0; i < kThreads / 32; ++i) s += red[i];
        s = s - d_ref;
        out[dy * lw + dx] = s > 0.0f ? s : 0.0f;
    }
}

__device__ __forceinline__ int reflect_sym(int p, int len) {  // scipy 'reflect': d c b a | a b c d | d c b a
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = p < 0 ? -p - 1 : 2 * len - 1 - p;
    return p;
}

__global__ void k_burn_blur(const float *__restrict__ in, int lh, int lw, int axis, float *__restrict__ out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= lh * lw) return;
    const int y = idx / lw, x = idx - y * lw;
    double wsum = 0.0, w[13];
    for (int t = -6; t <= 6; ++t) {
        w[t + 6] = exp(-0.5 / 9.0 * (double)(t * t));
        wsum += w[t + 6];
    }
    double acc = 0.0;
    for (int t = -6; t <= 6; ++t) {
        const int yy = axis == 0 ? reflect_sym(y + t, lh) : y, xx = axis == 1 ? reflect_sym(x + t, lw) : x;
        acc += (w[t + 6] / wsum) * (double)in[yy * lw + xx];
    }
    out[idx] = (float)acc;
}

cudaError_t launch_burn_mask(const float *green_plane, int H, int W, int lh, int lw, float d_ref, float *tmp,
                             float *map, cudaStream_t st) {
    k_burn_down<<<dim3(lw, lh), kThreads, 0, st>>>(green_plane, H, W, lh, lw, d_ref, map);
    const int n = lh * lw;
    k_burn_blur<<<(n + 127) / 128, 128, 0, st>>>(map, lh, lw, 0, tmp);
    k_burn_blur<<<(n + 127) / 128, 128, 0, st>>>(tmp, lh, lw, 1, map);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// auto exposure (next row 8f-1; reference color_processing.py:71-99, called from raw_conversion.py:51-53):
// mean over the green samples of every second row and column of  v ** (1 / factor).
// Each term is rounded to binary32 like the reference's float32 array power; the terms are summed in
// binary64 in a fixed order (per-thread stride loop, warp shuffles, one partial per CTA, then one CTA
// over the partials), so the result is deterministic and at least as accurate as numpy's float32
// pairwise mean.
// ------------------------------------------------------------------------------------------
template <int FMT>
__global__ void __launch_bounds__(kThreads)
k_exposure_partial(const void *__restrict__ in, int H, int W, double inv_factor, double *__restrict__ partial) {
    const int hs = (H + 1) >> 1, wsub = (W + 1) >> 1;
    const size_t total = (size_t)hs * wsub;
    const size_t stride = (size_t)gridDim.x * kThreads;
    constexpr int CIN = InFmt<FMT>::cin;
    double acc = 0.0;
    for (size_t sidx = (size_t)blockIdx.x * kThreads + threadIdx.x; sidx < total; sidx += stride) {
        const int ys = (int)(sidx / wsub), xs = (int)(sidx - (size_t)ys * wsub);
        const size_t pix = (size_t)(2 * ys) * W + 2 * xs;
        float g;
        if (InFmt<FMT>::u16) g = u16_to_linear(static_cast<const uint16_t *>(in)[pix * CIN + 1], 1.0f);
        else g = static_cast<const float *>(in)[pix * CIN + 1];
        acc += (double)(float)pow((double)g, inv_factor);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    __shared__ double wsum[kThreads / 32];
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < kThreads / 32; ++i) t += wsum[i];
        partial[blockIdx.x] = t;
    }
}

__global__ void k_exposure_final(const double *__restrict__ partial, int n, double count, double *__restrict__ out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < n; ++i) t += partial[i];
        out[0] = t / count;
    }
}

cudaError_t launch_exposure_mean(const void *in, int fmt, int H, int W, double inv_factor, double *partial, int nblocks,
                                 double *out, cudaStream_t st) {
    const double count = (double)((H + 1) / 2) * (double)((W + 1) / 2);
    switch (fmt) {
        case 0: k_exposure_partial<0><<<nblocks, kThreads, 0, st>>>(in, H, W, inv_factor, partial); break;
        case 1: k_exposure_partial<1><<<nblocks, kThreads, 0, st>>>(in, H, W, inv_factor, partial); break;
        case 2: k_exposure_partial<2><<<nblocks, kThreads, 0, st>>>(in, H, W, inv_factor, partial); break;
        case 3: k_exposure_partial<3><<<nblocks, kThreads, 0, st>>>(in, H, W, inv_factor, partial); break;
        default: return cudaErrorInvalidValue;
    }
    k_exposure_final<<<1, 32, 0, st>>>(partial, nblocks, count, out);
    return cudaGetLastError();
}

}  // namespace r2f
