// Per-pixel device functions of the raw2film render path (sm_100a).
//
// Everything here is compiled with -fmad=false: each float operation is a separately
// rounded IEEE-754 operation in the order written, which is what makes the pointwise
// chain bit-exact against oracle/pointwise_oracle.c.  Where a fused multiply-add is
// wanted (the convolutions) the code calls fmaf() explicitly.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "log10_tables.h"

namespace r2f {

struct Lut2D {
    const float *tab;    // (n, n, 3) lut[x_idx][y_idx], packed (the FFT row kernels park this 48 KB copy)
    const float4 *tab4;  // same vertices padded to float4: one 128-bit read per vertex
    int n;
    int pitch4;          // vertices per row of tab4 (n in global memory; n + 1 for the shared-memory copy of the fast
                         // pointwise kernel, so that cells in neighbouring rows fall into different banks)
    __host__ __device__ Lut2D() : tab(nullptr), tab4(nullptr), n(0), pitch4(0) {}
    __host__ __device__ Lut2D(const float *t, int n_) : tab(t), tab4(nullptr), n(n_), pitch4(n_) {}
    __host__ __device__ Lut2D(const float *t, const float4 *t4, int n_) : tab(t), tab4(t4), n(n_), pitch4(n_) {}
};

struct Curve1D {
    // rows R,G,B of the (4, N) table as segments: seg[ch*N + i] = (row[i], row[i+1] - row[i]) (the float32
    // difference the interpolation would form anyway), last entry (row[N-1], 0): one 8-byte read per lookup
    const float2 *seg;
    int N;
    float x0;         // first abscissa
    float inv_range;  // float32(1 / (x_last - x_first))
    // row 0 of the table when its abscissa is NOT uniform (else nullptr): the lookup then follows np.interp
    // (binary search for the bracketing samples, binary64 slope), see curve_eval
    const float *xp;
};

struct Lut3D {
    const float4 *tab;  // (n, n, n) vertices padded to float4
    int n;
    double s;  // scale * (n - 1)
    // guarded float32 fast path for uint8 outputs (tetra_quant_u8): sf = (float)s, `margin` = proven
    // bound on |255 * (fast - exact)|; fast_exact_d != 0 when s is a power of two (fractions exact)
    float sf;
    float margin;
    int fast_ok;
};

// ---- frame input formats -------------------------------------------------------------------
// FMT 0: float32 x3 (XYZ), 1: float32 x4 (XYZ + alpha, reference GPU payload gpu_processor.py:765),
//     2: uint16 x3, 3: uint16 x4 -- what rawpy hands over (reference raw_conversion.py:38-51); the
//     device then does the reference's `astype(float32) / 65535.0` and `*= 2**calc_exposure`
//     (raw_conversion.py:51-53) itself, halving the PCIe bytes per frame.
constexpr int kFmtF32x3 = 0, kFmtF32x4 = 1, kFmtU16x3 = 2, kFmtU16x4 = 3;
template <int FMT>
struct InFmt {
    static constexpr int cin = (FMT & 1) ? 4 : 3;
    static constexpr bool u16 = FMT >= 2;
};

// float32(u) / 65535 correctly rounded (checked for all 65536 inputs: q = u*rcp refined by one FMA
// remainder step), then the float32 product with the exposure gain.
__device__ __forceinline__ float u16_to_linear(uint32_t u, float gain) {
    const float x = (float)u;
    const float rcp = 1.0f / 65535.0f;
    const float q = x * rcp;
    const float r = fmaf(-q, 65535.0f, x);
    return fmaf(r, rcp, q) * gain;
}

template <int FMT>
__device__ __forceinline__ void load_px(const void *__restrict__ in, size_t pix, float gain, float &X, float &Y,
                                        float &Z) {
    constexpr int cin = InFmt<FMT>::cin;
    if (InFmt<FMT>::u16) {
        const uint16_t *p = static_cast<const uint16_t *>(in) + pix * cin;
        X = u16_to_linear(__ldg(p), gain);
        Y = u16_to_linear(__ldg(p + 1), gain);
        Z = u16_to_linear(__ldg(p + 2), gain);
    } else {
        const float *p = static_cast<const float *>(in) + pix * cin;
        X = __ldg(p);
        Y = __ldg(p + 1);
        Z = __ldg(p + 2);
    }
}

// four consecutive pixels with vector loads (q = quad index; the frame base is 16-byte aligned)
template <int FMT>
__device__ __forceinline__ void load_quad(const void *__restrict__ in, size_t q, float gain, float (&px)[4][3]) {
    if (FMT == kFmtF32x3) {
        const float4 *p = static_cast<const float4 *>(in) + 3 * q;
        const float4 a = __ldcs(p), b = __ldcs(p + 1), c = __ldcs(p + 2);
        px[0][0] = a.x; px[0][1] = a.y; px[0][2] = a.z;
        px[1][0] = a.w; px[1][1] = b.x; px[1][2] = b.y;
        px[2][0] = b.z; px[2][1] = b.w; px[2][2] = c.x;
        px[3][0] = c.y; px[3][1] = c.z; px[3][2] = c.w;
    } else if (FMT == kFmtF32x4) {
        const float4 *p = static_cast<const float4 *>(in) + 4 * q;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float4 a = __ldcs(p + i);
            px[i][0] = a.x; px[i][1] = a.y; px[i][2] = a.z;
        }
    } else if (FMT == kFmtU16x3) {
        const uint2 *p = static_cast<const uint2 *>(in) + 3 * q;  // 12 halfwords = 24 bytes
        const uint2 a = __ldcs(p), b = __ldcs(p + 1), c = __ldcs(p + 2);
        const uint32_t w[6] = {a.x, a.y, b.x, b.y, c.x, c.y};
#pragma unroll
        for (int i = 0; i < 12; ++i) {
            const uint32_t h = (i & 1) ? (w[i >> 1] >> 16) : (w[i >> 1] & 0xffffu);
            px[i / 3][i % 3] = u16_to_linear(h, gain);
        }
    } else {
        const uint4 *p = static_cast<const uint4 *>(in) + 2 * q;  // 16 halfwords = 32 bytes
        const uint4 a = __ldcs(p), b = __ldcs(p + 1);
        const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            px[i][0] = u16_to_linear(w[2 * i] & 0xffffu, gain);
            px[i][1] = u16_to_linear(w[2 * i] >> 16, gain);
            px[i][2] = u16_to_linear(w[2 * i + 1] & 0xffffu, gain);
        }
    }
}

// ---- a2: chromaticity-indexed input LUT (reference shaders/lut_2d.wgsl:18-108) ---------
// Branch-free; the float operations and their order are exactly those of
// oracle/pointwise_oracle.c lut2d_pixel (index clamps and selects do not round).
// SMEM: L.tab is known to point into shared memory (a table parked there by the caller); the nine vertex reads
// are then explicit ld.shared instead of generic loads, which the hardware routes through the global-load queue.
__device__ __forceinline__ float lds_f32(unsigned addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}

__device__ __forceinline__ float4 lds_f32x4(unsigned addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

// VEC4: read the float4-padded copy (L.tab4; three 128-bit reads per pixel) instead of the packed one
// (nine scalar reads whose stride of 3 floats makes 56 % of the shared-memory wavefronts bank conflicts).
template <bool SMEM = false, bool VEC4 = false>
__device__ __forceinline__ void lut2d_eval(const Lut2D &L, float X, float Y, float Z, float &e0, float &e1,
                                           float &e2) {
    const float S = (X + Y) + Z;
    const bool dark = S < 1e-12f;
    const int n = L.n;
    const float inv_sum = __fdiv_rn((float)(n - 1), dark ? 1.0f : S);
    const float r = X * inv_sum, g = Y * inv_sum;
    const float rfl = floorf(r), gfl = floorf(g);
    const float hi = (float)(n - 2);
    const int ri = (int)fminf(fmaxf(rfl, 0.0f), hi);  // NaN -> 0, like the oracle's clamp
    const int gi = (int)fminf(fmaxf(gfl, 0.0f), hi);
    const float rf = r - rfl, gf = g - gfl;
    const float fs = rf + gf;
    const bool lower = fs <= 1.0f;
    const float wa = lower ? rf : 1.0f - gf;
    const float wb = lower ? gf : 1.0f - rf;
    const float wc = lower ? 1.0f - fs : fs - 1.0f;
    float a0, a1, a2, b0, b1, b2, c0, c1, c2;
    if (VEC4) {
        const int pitch = L.pitch4;
        const float4 *base = L.tab4 + (ri * pitch + gi);
        const float4 *pa = base + pitch;                     // lut[ri+1][gi]
        const float4 *pb = base + 1;                         // lut[ri][gi+1]
        const float4 *pc = base + (lower ? 0 : pitch + 1);   // lut[ri][gi] or lut[ri+1][gi+1]
        float4 va, vb, vc;
        if (SMEM) {
            va = lds_f32x4((unsigned)__cvta_generic_to_shared(pa));
            vb = lds_f32x4((unsigned)__cvta_generic_to_shared(pb));
            vc = lds_f32x4((unsigned)__cvta_generic_to_shared(pc));
        } else {
            va = __ldg(pa); vb = __ldg(pb); vc = __ldg(pc);
        }
        a0 = va.x; a1 = va.y; a2 = va.z;
        b0 = vb.x; b1 = vb.y; b2 = vb.z;
        c0 = vc.x; c1 = vc.y; c2 = vc.z;
    } else {
        const int n3 = 3 * n;
        const float *base = L.tab + (ri * n3 + gi * 3);
        const float *a = base + n3;                       // lut[ri+1][gi]
        const float *b = base + 3;                        // lut[ri][gi+1]
        const float *c = base + (lower ? 0 : n3 + 3);     // lut[ri][gi] or lut[ri+1][gi+1]
        if (SMEM) {
            const unsigned sa = (unsigned)__cvta_generic_to_shared(a), sb = (unsigned)__cvta_generic_to_shared(b);
            const unsigned sc = (unsigned)__cvta_generic_to_shared(c);
            a0 = lds_f32(sa); a1 = lds_f32(sa + 4); a2 = lds_f32(sa + 8);
            b0 = lds_f32(sb); b1 = lds_f32(sb + 4); b2 = lds_f32(sb + 8);
            c0 = lds_f32(sc); c1 = lds_f32(sc + 4); c2 = lds_f32(sc + 8);
        } else {
            a0 = a[0]; a1 = a[1]; a2 = a[2];
            b0 = b[0]; b1 = b[1]; b2 = b[2];
            c0 = c[0]; c1 = c[1]; c2 = c[2];
        }
    }
    const float v0 = ((a0 * wa + b0 * wb) + c0 * wc) * S;
    const float v1 = ((a1 * wa + b1 * wb) + c1 * wc) * S;
    const float v2 = ((a2 * wa + b2 * wb) + c2 * wc) * S;
    e0 = dark ? 0.0f : v0;
    e1 = dark ? 0.0f : v1;
    e2 = dark ? 0.0f : v2;
}

// ---- a4: log10 with lower clip (shaders/lut_1d.wgsl:23-26) -------------------------------
// The oracle defines log10 as binary64 log10 rounded once to binary32 (pointwise_oracle.c
// log10_clip1), i.e. a correctly rounded log10f.  CUDA's generic double log10 costs ~100
// instructions; log10_exact gets the same bits from a 256-entry table + degree-7 series:
//   x = 2^k * z, z in [0.699, 1.398);  r = fma(z, invc_i, -1)  (|r| <= 2^-8, error 2^-62)
//   y = (k*log10(2) + T_i) + r * P(r),  relative error <= ~3 * 2^-53.
// If y lies within 64 binary64-ulps of a binary32 rounding tie the generic routine decides
// (probability 2^-22).  Validated on the host against glibc on 3.3e8 inputs: 0 mismatches
// (tools/gen_log10_tables.py documents the tables).
static __device__ __noinline__ float log10_slow(float c) { return (float)log10((double)c); }

// Series coefficients in the constant bank: DFMA takes them as c[bank][offset] operands.  As literals the
// compiler rebuilds each 64-bit immediate with two UMOVs per use (36 instructions per pixel in k_pointwise).
static __constant__ double kLog10Poly[7] = {R2F_LOG10_A1, R2F_LOG10_A2, R2F_LOG10_A3, R2F_LOG10_A4,
                                            R2F_LOG10_A5, R2F_LOG10_A6, R2F_LOG10_A7};

__device__ __forceinline__ float log10_exact(float c) {
    const uint32_t ix = __float_as_uint(c);
    if (ix - 0x00800000u >= 0x7f000000u) return log10_slow(c);  // zero, subnormal, negative, inf, nan
    const uint32_t tmp = ix - 0x3f330000u;
    const int i = (tmp >> 15) & 255;
    const int k = (int)tmp >> 23;
    const uint32_t iz = ix - (tmp & 0xff800000u);
    // z widened to binary64 by bit construction (z is a normal float): no F2F conversion
    const double z = __hiloint2double((int)((iz >> 3) + 0x38000000u), (int)(iz << 29));
    const double2 tab = __ldg(&kLog10Tab[i]);
    const double r = fma(z, tab.x, -1.0);
    double p = kLog10Poly[6];
    p = fma(p, r, kLog10Poly[5]);
    p = fma(p, r, kLog10Poly[4]);
    p = fma(p, r, kLog10Poly[3]);
    p = fma(p, r, kLog10Poly[2]);
    p = fma(p, r, kLog10Poly[1]);
    p = fma(p, r, kLog10Poly[0]);
    const double y = fma(r, p, __ldg(&kLog10Exp[k + 160]) + tab.y);
    const uint32_t t = (uint32_t)__double2loint(y) & 0x1fffffffu;
    if ((uint32_t)(t - (0x10000000u - 64u)) < 128u) return log10_slow(c);
    return __double2float_rn(y);
}

__device__ __forceinline__ float log10_clip(float v, float eps) {
    const float c = v > eps ? v : eps;
    return log10_exact(c);
}

// Approximate variant (MUFU.LG2, abs error ~1e-6 at |log10| = 6) for stages whose input already
// carries float32 convolution noise (tolerance 1e-4 in density).
__device__ __forceinline__ float log10_clip_fast(float v, float eps) {
    const float c = v > eps ? v : eps;
    return __log2f(c) * 0.30102999566398119521f;
}

// ---- a5: per-channel curve, uniform abscissa, clamped ends (lut_1d.wgsl:43-47) -----------
// Non-uniform abscissa: np.interp semantics (what a NumPy multi_channel_interp does on the CPU): clamped ends,
// slope and offset in binary64, one rounding to binary32.  NaN -> first sample (like the uniform path).
static __device__ __noinline__ float curve_eval_interp(const Curve1D &C, int ch, float v) {
    const int N = C.N;
    const float2 *seg = C.seg + ch * N;
    if (!(v > C.xp[0])) return seg[0].x;
    if (v >= C.xp[N - 1]) return seg[N - 1].x;
    int lo = 0, hi = N - 1;  // xp[lo] <= v < xp[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (v >= C.xp[mid]) lo = mid;
        else hi = mid;
    }
    const double x0 = C.xp[lo], x1 = C.xp[lo + 1], f0 = seg[lo].x, f1 = seg[lo + 1].x;
    const double slope = (f1 - f0) / (x1 - x0);
    return (float)(slope * ((double)v - x0) + f0);
}

// Uniform abscissa only (C.xp == nullptr): the hot kernels call this one and the C ABI routes tables with a
// non-uniform abscissa to the generic kernels, which call curve_eval_any.
__device__ __forceinline__ float curve_eval(const Curve1D &C, int ch, float v) {
    float t = (v - C.x0) * C.inv_range;
    t = fminf(fmaxf(t, 0.0f), 1.0f);  // clamp; NaN -> 0 (fmaxf returns the non-NaN operand)
    const float p = t * (float)(C.N - 1);
    const int i = min((int)p, C.N - 2);
    const float f = p - (float)i;
    const float2 s = C.seg[ch * C.N + i];
    return s.x + f * s.y;
}

__device__ __forceinline__ float curve_eval_any(const Curve1D &C, int ch, float v) {
    if (C.xp != nullptr) return curve_eval_interp(C, ch, v);
    return curve_eval(C, ch, v);
}

// ANY: honour a non-uniform abscissa (generic kernels); otherwise the table is known to be uniform
template <bool ANY = false>
__device__ __forceinline__ float density_eval(const Curve1D &C, int ch, float exposure, float eps) {
    const float l = log10_clip(exposure, eps);
    return ANY ? curve_eval_any(C, ch, l) : curve_eval(C, ch, l);
}
template <bool ANY = false>
__device__ __forceinline__ float density_eval_fast(const Curve1D &C, int ch, float exposure, float eps) {
    const float l = log10_clip_fast(exposure, eps);
    return ANY ? curve_eval_any(C, ch, l) : curve_eval(C, ch, l);
}

// ---- a9: tetrahedral 3-D LUT (reference utils.py:247-380) -----------------------------------
// binary64 coordinates and accumulation, binary32 vertex differences (numba's typing of the
// reference for a float64 `scale`; bit-exact against it, tests/golden/tetra.npz).
__device__ __forceinline__ int wrap_idx(int i, int n) {
    if (i < 0) {
        i += n;
        if (i < 0) i = 0;
    }
    return i;
}

__device__ __forceinline__ void tetra_eval(const Lut3D &L, float dr_in, float dg_in, float db_in, float &o0,
                                           float &o1, float &o2) {
    const int n = L.n;
    double d[3];
    int i0[3];
    const float in[3] = {dr_in, dg_in, db_in};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double v = (double)in[k] * L.s;
        double vt = v;
        if (!(vt == vt)) vt = 0.0;
        vt = fmin(fmax(vt, -2.0e9), 2.0e9);
        int i = (int)vt;
        if (i >= n - 1) {
            i = n - 2;
            d[k] = 1.0;
        } else {
            d[k] = v - (double)i;
        }
        i0[k] = i;
    }
    const int r0 = wrap_idx(i0[0], n), g0 = wrap_idx(i0[1], n), b0 = wrap_idx(i0[2], n);
    const int r1 = wrap_idx(i0[0] + 1, n), g1 = wrap_idx(i0[1] + 1, n), b1 = wrap_idx(i0[2] + 1, n);
    const double dr = d[0], dg = d[1], db = d[2];
    // vertex 1 and 2 of the chosen tetrahedron as (r,g,b) corner selectors, and ordered fractions
    int m1r, m1g, m1b, m2r, m2g, m2b;
    double d1, d2, d3;
    if (dr >= dg) {
        if (dg >= db) { m1r = r1; m1g = g0; m1b = b0; m2r = r1; m2g = g1; m2b = b0; d1 = dr; d2 = dg; d3 = db; }
        else if (dr >= db) { m1r = r1; m1g = g0; m1b = b0; m2r = r1; m2g = g0; m2b = b1; d1 = dr; d2 = db; d3 = dg; }
        else { m1r = r0; m1g = g0; m1b = b1; m2r = r1; m2g = g0; m2b = b1; d1 = db; d2 = dr; d3 = dg; }
    } else {
        if (db >= dg) { m1r = r0; m1g = g0; m1b = b1; m2r = r0; m2g = g1; m2b = b1; d1 = db; d2 = dg; d3 = dr; }
        else if (db >= dr) { m1r = r0; m1g = g1; m1b = b0; m2r = r0; m2g = g1; m2b = b1; d1 = dg; d2 = db; d3 = dr; }
        else { m1r = r0; m1g = g1; m1b = b0; m2r = r1; m2g = g1; m2b = b0; d1 = dg; d2 = dr; d3 = db; }
    }
    const float4 c000 = __ldg(L.tab + (r0 * n + g0) * n + b0);
    const float4 cm1 = __ldg(L.tab + (m1r * n + m1g) * n + m1b);
    const float4 cm2 = __ldg(L.tab + (m2r * n + m2g) * n + m2b);
    const float4 c111 = __ldg(L.tab + (r1 * n + g1) * n + b1);
#define R2F_TETRA_CH(f)                                                                                       \
    (float)((((double)c000.f + d1 * (double)(cm1.f - c000.f)) + d2 * (double)(cm2.f - cm1.f)) +               \
            d3 * (double)(c111.f - cm2.f))
    o0 = R2F_TETRA_CH(x);
    o1 = R2F_TETRA_CH(y);
    o2 = R2F_TETRA_CH(z);
#undef R2F_TETRA_CH
}

// ---- a10: quantise (reference cpu_processor.py:407): float32 * 255, truncate ----------------
__device__ __forceinline__ uint32_t quantise_u8(float v) {
    const float q = v * 255.0f;
    if (!(q > 0.0f)) return 0u;
    if (q >= 255.0f) return 255u;
    return (uint32_t)(int)q;
}

// a9 + a10 for uint8 outputs: float32 evaluation of the same tetrahedral formula, accepted only
// when 255*value is provably on the same side of every quantisation boundary as the exact
// binary64 path (|255*(fast - exact)| <= L.margin, bound derived in r2f_set_lut3d); otherwise the
// exact path decides.  Interpolation is continuous across cells and tetrahedra, so a different
// cell choice of the float32 coordinates is covered by the same bound.
__device__ __forceinline__ bool quant_safe(float q, float m, uint32_t &out) {
    // every value in [q-m, q+m] truncates (after the [0,255] clamp) to the same integer
    const int lo = __float2int_rd(q - m), hi = __float2int_rd(q + m);
    out = (uint32_t)min(max(lo, 0), 255);
    return lo == hi;  // NaN: both conversions give 0 -> "safe" 0, same as the exact path's !(q > 0) -> 0
}

__device__ __forceinline__ void tetra_quant_u8(const Lut3D &L, float dr_in, float dg_in, float db_in, uint32_t &q0,
                                               uint32_t &q1, uint32_t &q2) {
    bool ok = false;
    if (L.fast_ok) {
        const int n = L.n;
        const float vr = dr_in * L.sf, vg = dg_in * L.sf, vb = db_in * L.sf;
        const float vmin = fminf(fminf(vr, vg), vb), vmax = fmaxf(fmaxf(vr, vg), vb);
        if (vmin >= 0.0f && vmax < 1.0e6f && vr == vr && vg == vg && vb == vb) {
            const int top = n - 2;
            const int r0 = min((int)vr, top), g0 = min((int)vg, top), b0 = min((int)vb, top);
            // fraction; at or above the last lattice plane the reference pins it to 1 (utils.py:273-289)
            const float dr = fminf(vr - (float)r0, 1.0f), dg = fminf(vg - (float)g0, 1.0f);
            const float db = fminf(vb - (float)b0, 1.0f);
            // ordered fractions d1 >= d2 >= d3 and the lattice steps of their axes (ties: any consistent
            // order gives the same interpolant; the exact path keeps the reference's branch order)
            const float d1 = fmaxf(fmaxf(dr, dg), db), d3 = fminf(fminf(dr, dg), db);
            const float d2 = fmaxf(fminf(dr, dg), fminf(fmaxf(dr, dg), db));
            const int sr = n * n, sg = n;
            const int o1 = dr == d1 ? sr : (dg == d1 ? sg : 1);
            const int o3 = db == d3 ? 1 : (dg == d3 ? sg : sr);
            const int o111 = sr + sg + 1;
            const float4 *base = L.tab + (r0 * n + g0) * n + b0;
            const float4 c000 = __ldg(base), cm1 = __ldg(base + o1), cm2 = __ldg(base + (o111 - o3));
            const float4 c111 = __ldg(base + o111);
            const float s0 = fmaf(d3, c111.x - cm2.x, fmaf(d2, cm2.x - cm1.x, fmaf(d1, cm1.x - c000.x, c000.x)));
            const float s1 = fmaf(d3, c111.y - cm2.y, fmaf(d2, cm2.y - cm1.y, fmaf(d1, cm1.y - c000.y, c000.y)));
            const float s2 = fmaf(d3, c111.z - cm2.z, fmaf(d2, cm2.z - cm1.z, fmaf(d1, cm1.z - c000.z, c000.z)));
            const float m = L.margin;
            const bool a0 = quant_safe(s0 * 255.0f, m, q0);
            const bool a1 = quant_safe(s1 * 255.0f, m, q1);
            const bool a2 = quant_safe(s2 * 255.0f, m, q2);
            ok = a0 && a1 && a2;
        }
    }
    if (!ok) {
        float o0, o1f, o2f;
        tetra_eval(L, dr_in, dg_in, db_in, o0, o1f, o2f);
        q0 = quantise_u8(o0);
        q1 = quantise_u8(o1f);
        q2 = quantise_u8(o2f);
    }
}

// ---- BORDER_REFLECT_101 index (cv2.filter2D default, reference effects.py:146-156) -----------
__device__ __forceinline__ int reflect101(int p, int len) {
    if (len == 1) return 0;
    while ((unsigned)p >= (unsigned)len) p = p < 0 ? -p : 2 * len - 2 - p;
    return p;
}

}  // namespace r2f
