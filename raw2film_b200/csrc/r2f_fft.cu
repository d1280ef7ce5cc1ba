// FFT path for the wide halation kernel (reference effects.py:200-287; a3 in SURVEY 8a).
//
// The halation kernel is (f_c*K + delta)/(f_c+1) per channel with ONE radially symmetric base
// kernel K (43x43 at 24 MP, 133x133 at 61 MP with halation_size=2): the red and green layers
// are packed into one complex image z = R + iG, and because K is real and even its spectrum is
// real, so a single complex 2-D FFT convolution filters both layers:
//     IFFT2( FFT2(z) * Khat ) = (K (*) R) + i (K (*) G).
// cv2.filter2D itself switches to a DFT for kernels larger than 11x11, so this is also the
// arithmetic the reference carries.
//
// Three kernels, all HBM-streaming, FFTs entirely in shared memory (hand-written mixed-radix
// Stockham, radices 2/3/4/5, twiddles from a double-precision-built root table):
//   k_fft_rows_fwd : 2 image rows per CTA.  Loads pixels (optionally XYZ -> 2-D LUT fused, a2),
//                    builds the BORDER_REFLECT_101 padded row, FFT length Wp, writes the row
//                    spectrum in a column-blocked layout S[Wp/NC][H][NC] (64 B segments).
//   k_fft_cols     : NC columns per CTA (one contiguous block of S).  Reflect-pads the column
//                    (vertical reflection of spectra == spectrum of the reflected rows), FFT length
//                    Hp, multiplies by the real Khat (1/(Hp*Wp) folded in), inverse FFT, writes back
//                    the H valid rows in place.  Forward, product and inverse never leave the SM.
//   k_fft_rows_inv : 2 rows per CTA, inverse FFT length Wp, crops, applies out = alpha*conv + beta*x
//                    per layer and (optionally) the fused log10 + H-D curve epilogue (a4+a5).
// Inverse transforms use IFFT(x) = swap(FFT(swap(x))).
#include "r2f_fft.h"

#include "conv_tile.cuh"

#include <cmath>
#include <cstdlib>
#include <initializer_list>
#include <vector>

namespace r2f {

// ------------------------------------------------------------------------------------------
// shared-memory Stockham FFT
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, -(a.y * b.y)), fmaf(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 mul_neg_i(float2 a) { return make_float2(a.y, -a.x); }  // a * (-i)

template <int R>
__device__ __forceinline__ void butterfly(float2 (&v)[R]);

template <>
__device__ __forceinline__ void butterfly<2>(float2 (&v)[2]) {
    const float2 a = v[0], b = v[1];
    v[0] = cadd(a, b);
    v[1] = csub(a, b);
}

template <>
__device__ __forceinline__ void butterfly<4>(float2 (&v)[4]) {
    const float2 t0 = cadd(v[0], v[2]), t1 = csub(v[0], v[2]);
    const float2 t2 = cadd(v[1], v[3]), t3 = mul_neg_i(csub(v[1], v[3]));
    v[0] = cadd(t0, t2);
    v[1] = cadd(t1, t3);
    v[2] = csub(t0, t2);
    v[3] = csub(t1, t3);
}

template <>
__device__ __forceinline__ void butterfly<3>(float2 (&v)[3]) {
    const float kHalfSqrt3 = 0.86602540378443864676f;
    const float2 s = cadd(v[1], v[2]), d = csub(v[1], v[2]);
    const float2 m = make_float2(fmaf(-0.5f, s.x, v[0].x), fmaf(-0.5f, s.y, v[0].y));
    const float2 rot = make_float2(kHalfSqrt3 * d.y, -kHalfSqrt3 * d.x);
    v[0] = cadd(v[0], s);
    v[1] = cadd(m, rot);
    v[2] = csub(m, rot);
}

template <>
__device__ __forceinline__ void butterfly<5>(float2 (&v)[5]) {
    const float c1 = 0.30901699437494742410f, c2 = -0.80901699437494742410f;
    const float s1 = 0.95105651629515357212f, s2 = 0.58778525229247312917f;
    const float2 a1 = cadd(v[1], v[4]), a2 = cadd(v[2], v[3]);
    const float2 b1 = csub(v[1], v[4]), b2 = csub(v[2], v[3]);
    const float2 r1 = make_float2(fmaf(c2, a2.x, fmaf(c1, a1.x, v[0].x)), fmaf(c2, a2.y, fmaf(c1, a1.y, v[0].y)));
    const float2 r2 = make_float2(fmaf(c1, a2.x, fmaf(c2, a1.x, v[0].x)), fmaf(c1, a2.y, fmaf(c2, a1.y, v[0].y)));
    const float2 i1 = make_float2(fmaf(s2, b2.x, s1 * b1.x), fmaf(s2, b2.y, s1 * b1.y));
    const float2 i2 = make_float2(fmaf(-s1, b2.x, s2 * b1.x), fmaf(-s1, b2.y, s2 * b1.y));
    const float2 n1 = mul_neg_i(i1), n2 = mul_neg_i(i2);
    v[0] = cadd(v[0], cadd(a1, a2));
    v[1] = cadd(r1, n1);
    v[4] = csub(r1, n1);
    v[2] = cadd(r2, n2);
    v[3] = csub(r2, n2);
}

template <>
__device__ __forceinline__ void butterfly<8>(float2 (&v)[8]) {
    const float h = 0.70710678118654752440f;
    float2 e[4] = {v[0], v[2], v[4], v[6]}, o[4] = {v[1], v[3], v[5], v[7]};
    butterfly<4>(e);
    butterfly<4>(o);
    const float2 o1 = make_float2((o[1].x + o[1].y) * h, (o[1].y - o[1].x) * h);   // * exp(-i pi/4)
    const float2 o2 = mul_neg_i(o[2]);                                             // * exp(-i pi/2)
    const float2 o3 = make_float2((o[3].y - o[3].x) * h, -(o[3].x + o[3].y) * h);  // * exp(-3i pi/4)
    v[0] = cadd(e[0], o[0]); v[4] = csub(e[0], o[0]);
    v[1] = cadd(e[1], o1);   v[5] = csub(e[1], o1);
    v[2] = cadd(e[2], o2);   v[6] = csub(e[2], o2);
    v[3] = cadd(e[3], o3);   v[7] = csub(e[3], o3);
}

// Twiddles w^1 .. w^(R-1) of one butterfly from the pass table tw[(t-1)*stride + k].  Only the powers 1, 2 and 4 are
// read; the others are products of those (one extra rounding, ~6e-8 relative).  The tables are read through L1, whose
// bandwidth -- shared with the line buffers in shared memory -- is what bounds these kernels: this drops 4 of the 7
// table reads of a radix-8 butterfly (46 -> 38 L1 wavefronts per warp and butterfly) for 4 complex multiplies on an
// FMA pipe that is a quarter busy.
template <int R>
__device__ __forceinline__ void load_twiddles(const float2 *__restrict__ tw, int stride, int k, float2 (&w)[R]) {
    w[1] = __ldg(tw + k);
    if (R > 2) w[2] = __ldg(tw + stride + k);
    if (R > 3) w[3] = cmul(w[1], w[2]);
    if (R == 5) w[4] = cmul(w[2], w[2]);
    if (R == 8) {
        w[4] = __ldg(tw + 3 * stride + k);
        w[5] = cmul(w[1], w[4]);
        w[6] = cmul(w[2], w[4]);
        w[7] = cmul(w[3], w[4]);
    }
}

// Line buffers are addressed through an XOR swizzle of the 16-byte chunk index: within every 128-byte row of a
// line (16 float2) the chunk column is XORed with the row number.  Reads of consecutive elements stay
// conflict-free (a row is only permuted), and the strided stores of the first Stockham passes (a thread
// writes R consecutive outputs, then runs of Ns = 8) spread over all banks instead of 2 or 4 bank groups.
// sw(i + m*128) == sw(i) + m*128, which the compile-time plans exploit (one swizzle per butterfly).
__device__ __forceinline__ int sw(int i) { return i ^ ((i >> 3) & 14); }

__device__ __forceinline__ float pick3(int c, float v0, float v1, float v2) { return c == 0 ? v0 : (c == 1 ? v1 : v2); }

// Streaming global loads that do not allocate in L1 (spectrum blocks and the kernel spectrum are read once;
// the L1 is kept for the twiddle tables).
__device__ __forceinline__ float4 ldg_stream4(const float4 *p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ float ldg_stream(const float *p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

// A group of threads that cooperates on one FFT line and synchronises on its own named barrier.
struct Group {
    int tid, size, bar;
};
__device__ __forceinline__ void group_sync(const Group &g) {
    asm volatile("bar.sync %0, %1;" ::"r"(g.bar), "r"(g.size) : "memory");
}

// One out-of-place Stockham pass of radix R: src[0..n) -> dst[0..n) (auto-sorting, no bit reversal).
// Butterfly j reads src[j + t*n/R] (contiguous across lanes) and writes dst[j0 + t*Ns].
// `tw` is this pass's twiddle table: tw[(t-1)*Ns + k] = exp(-2 pi i t k / (Ns R)), so lanes with
// consecutive j read consecutive twiddles.
// SCALE: the outputs are multiplied by the real table scale[] and their re/im swapped on the way out
// (the kernel-spectrum product of k_fft_cols folded into the last forward pass).
template <int R, bool SCALE = false>
__device__ __forceinline__ void fft_pass(const float2 *__restrict__ src, float2 *__restrict__ dst, int n, int Ns,
                                         const float2 *__restrict__ tw, const Group &g,
                                         const float *__restrict__ scale = nullptr) {
    const int nb = n / R;
    const bool pow2 = (Ns & (Ns - 1)) == 0;
#pragma unroll 2
    for (int j = g.tid; j < nb; j += g.size) {
        const int k = pow2 ? (j & (Ns - 1)) : (j % Ns);
        float2 v[R];
#pragma unroll
        for (int t = 0; t < R; ++t) v[t] = src[sw(j + t * nb)];
        if (Ns > 1) {
#pragma unroll
            for (int t = 1; t < R; ++t) v[t] = cmul(v[t], __ldg(tw + (t - 1) * Ns + k));
        }
        butterfly<R>(v);
        const int j0 = (j - k) * R + k;
        if (SCALE) {
#pragma unroll
            for (int t = 0; t < R; ++t) {
                const float w = ldg_stream(scale + j0 + t * Ns);
                v[t] = make_float2(v[t].y * w, v[t].x * w);
            }
        }
        if (Ns == 1 && (R & 1) == 0) {  // R consecutive outputs: 16-byte stores
#pragma unroll
            for (int t = 0; t < R / 2; ++t)
                *reinterpret_cast<float4 *>(dst + sw(j0 + 2 * t)) =
                    make_float4(v[2 * t].x, v[2 * t].y, v[2 * t + 1].x, v[2 * t + 1].y);
        } else {
#pragma unroll
            for (int t = 0; t < R; ++t) dst[sw(j0 + t * Ns)] = v[t];
        }
    }
    group_sync(g);
}

// Forward FFT of a[0..n) using b as the ping-pong partner.  The data in `a` must already be visible
// to the whole group.  Returns the buffer holding the result (a if the pass count is even).
template <bool SCALE_LAST = false>
__device__ __forceinline__ float2 *fft_forward(float2 *a, float2 *b, const FftLine &L, const Group &g,
                                               const float *__restrict__ scale = nullptr) {
    float2 *src = a, *dst = b;
    int Ns = 1;
    for (int s = 0; s < L.nrad; ++s) {
        const int R = L.rad[s];
        const float2 *tw = L.tw + L.tw_off[s];
        if (SCALE_LAST && s == L.nrad - 1) {
            switch (R) {
                case 2: fft_pass<2, true>(src, dst, L.n, Ns, tw, g, scale); break;
                case 3: fft_pass<3, true>(src, dst, L.n, Ns, tw, g, scale); break;
                case 4: fft_pass<4, true>(src, dst, L.n, Ns, tw, g, scale); break;
                case 5: fft_pass<5, true>(src, dst, L.n, Ns, tw, g, scale); break;
                default: fft_pass<8, true>(src, dst, L.n, Ns, tw, g, scale); break;
            }
        } else
        switch (R) {
            case 2: fft_pass<2>(src, dst, L.n, Ns, tw, g); break;
            case 3: fft_pass<3>(src, dst, L.n, Ns, tw, g); break;
            case 4: fft_pass<4>(src, dst, L.n, Ns, tw, g); break;
            case 5: fft_pass<5>(src, dst, L.n, Ns, tw, g); break;
            default: fft_pass<8>(src, dst, L.n, Ns, tw, g); break;
        }
        float2 *t = src;
        src = dst;
        dst = t;
        Ns *= R;
    }
    return src;
}

// ------------------------------------------------------------------------------------------
// Compile-time plans.  The generic passes above carry n, Ns, the radix and the group size as run-time
// values: two thirds of their instructions are index arithmetic (multiplies, a division for non-power-
// of-two Ns) and the five-way radix dispatch.  For the line lengths of the headline frame sizes the whole
// plan is a template: strides and twiddle offsets become immediates, the modulo a mask or a constant
// multiply, and every pass is fully unrolled.  Same arithmetic, same order, same tables.
// ------------------------------------------------------------------------------------------
template <int R, int N, int NS, int GS, bool SCALE>
__device__ __forceinline__ void fft_pass_c(const float2 *__restrict__ src, float2 *__restrict__ dst,
                                           const float2 *__restrict__ tw, const Group &g,
                                           const float *__restrict__ scale) {
    constexpr int NB = N / R;
    constexpr int ITERS = (NB + GS - 1) / GS;
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
        const int j = g.tid + it * GS;
        if ((NB % GS) != 0 && it == ITERS - 1 && j >= NB) break;
        const int k = (NS & (NS - 1)) == 0 ? (j & (NS - 1)) : (j % NS);
        float2 v[R];
#pragma unroll
        for (int t = 0; t < R; ++t) v[t] = (NB % 128 == 0) ? src[sw(j) + t * NB] : src[sw(j + t * NB)];
        if (NS > 1) {
            float2 w[R];
            load_twiddles<R>(tw, NS, k, w);
#pragma unroll
            for (int t = 1; t < R; ++t) v[t] = cmul(v[t], w[t]);
        }
        butterfly<R>(v);
        const int j0 = (j - k) * R + k;
        if (SCALE) {
#pragma unroll
            for (int t = 0; t < R; ++t) {
                const float w = ldg_stream(scale + j0 + t * NS);
                v[t] = make_float2(v[t].y * w, v[t].x * w);
            }
        }
        if (NS == 1 && (R & 1) == 0) {
#pragma unroll
            for (int t = 0; t < R / 2; ++t)
                *reinterpret_cast<float4 *>(dst + sw(j0 + 2 * t)) =
                    make_float4(v[2 * t].x, v[2 * t].y, v[2 * t + 1].x, v[2 * t + 1].y);
        } else {
#pragma unroll
            for (int t = 0; t < R; ++t) dst[(NS % 128 == 0) ? sw(j0) + t * NS : sw(j0 + t * NS)] = v[t];
        }
    }
    group_sync(g);
}

template <int N, int GS, bool SCALE_LAST, int NS, int TWOFF, int R, int... REST>
__device__ __forceinline__ float2 *fft_static(float2 *src, float2 *dst, const float2 *__restrict__ tw, const Group &g,
                                              const float *__restrict__ scale) {
    constexpr bool last = sizeof...(REST) == 0;
    fft_pass_c<R, N, NS, GS, SCALE_LAST && last>(src, dst, tw + TWOFF, g, scale);
    if constexpr (last) return dst;
    else return fft_static<N, GS, SCALE_LAST, NS * R, TWOFF + (R - 1) * NS, REST...>(dst, src, tw, g, scale);
}

// PLAN ids (kernel template parameter): 0 = run-time plan.  Keep in sync with static_plan_id() below.
//   1: n = 6144, 512-thread group, 8 8 8 4 3   (rows, 24 MP)      2: n = 4096, 512, 8 8 8 8      (columns, 24 MP)
//   3: n = 10240, 1024-thread group, 8 8 8 4 5 (rows, 61 MP)      4: n = 6912, 512, 8 8 4 3 3 3  (columns, 61 MP)
template <int PLAN, bool SCALE_LAST = false>
__device__ __forceinline__ float2 *fft_run(float2 *a, float2 *b, const FftLine &L, const Group &g,
                                           const float *__restrict__ scale = nullptr) {
    if constexpr (PLAN == 1) return fft_static<6144, 512, SCALE_LAST, 1, 0, 8, 8, 8, 4, 3>(a, b, L.tw, g, scale);
    else if constexpr (PLAN == 2) return fft_static<4096, 512, SCALE_LAST, 1, 0, 8, 8, 8, 8>(a, b, L.tw, g, scale);
    else if constexpr (PLAN == 3) return fft_static<10240, 1024, SCALE_LAST, 1, 0, 8, 8, 8, 4, 5>(a, b, L.tw, g, scale);
    else if constexpr (PLAN == 4) return fft_static<6912, 512, SCALE_LAST, 1, 0, 8, 8, 4, 3, 3, 3>(a, b, L.tw, g, scale);
    else return fft_forward<SCALE_LAST>(a, b, L, g, scale);
}

// ------------------------------------------------------------------------------------------
// In-place passes for the column kernel: forward decimation-in-frequency (natural order in, mixed-radix digit-
// reversed order out), product with the (equally permuted) kernel spectrum, inverse as a forward decimation-in-
// time transform of the re/im-swapped data (digit-reversed in, natural out).  No ping-pong partner: a column
// needs n float2 of shared memory instead of 2n, so every column of a CTA is transformed concurrently by its own
// 256-thread group (two columns per CTA and two CTAs per SM by default, see k_fft_cols_ip).  The innermost radix of the forward transform, the product and the innermost radix of
// the inverse touch the same R contiguous elements and are fused in registers (ip_mid).
// ------------------------------------------------------------------------------------------
template <int R, int N, int B, int GS, bool DIT>
__device__ __forceinline__ void ip_pass(float2 *__restrict__ x, const float2 *__restrict__ tw, const Group &g) {
    constexpr int S = B / R, NB = N / R;
    constexpr int ITERS = (NB + GS - 1) / GS;
    static_assert(S > 1, "the innermost pass is ip_mid");
#pragma unroll 1   // one butterfly's data live at a time: the 1024-thread CTA caps a thread at 64 registers
    for (int it = 0; it < ITERS; ++it) {
        const int j = g.tid + it * GS;
        if ((NB % GS) != 0 && it == ITERS - 1 && j >= NB) break;
        const int blk = j / S, k = j - blk * S;
        const int base = blk * B + k;
        float2 v[R];
#pragma unroll
        for (int t = 0; t < R; ++t) v[t] = (S % 128 == 0) ? x[sw(base) + t * S] : x[sw(base + t * S)];
        float2 w[R];
        load_twiddles<R>(tw, S, k, w);
        if (DIT) {
#pragma unroll
            for (int t = 1; t < R; ++t) v[t] = cmul(v[t], w[t]);
        }
        butterfly<R>(v);
        if (!DIT) {
#pragma unroll
            for (int t = 1; t < R; ++t) v[t] = cmul(v[t], w[t]);
        }
#pragma unroll
        for (int t = 0; t < R; ++t) x[(S % 128 == 0) ? sw(base) + t * S : sw(base + t * S)] = v[t];
    }
    group_sync(g);
}

template <int R, int N, int GS>
__device__ __forceinline__ void ip_mid(float2 *__restrict__ x, const float *__restrict__ kh, const Group &g) {
    constexpr int NB = N / R;
    constexpr int ITERS = (NB + GS - 1) / GS;
#pragma unroll 1   // one butterfly's data live at a time: the 1024-thread CTA caps a thread at 64 registers
    for (int it = 0; it < ITERS; ++it) {
        const int j = g.tid + it * GS;
        if ((NB % GS) != 0 && it == ITERS - 1 && j >= NB) break;
        const int base = j * R;
        float2 v[R];
        if ((R & 1) == 0) {
#pragma unroll
            for (int t = 0; t < R / 2; ++t) {
                const float4 q = *reinterpret_cast<const float4 *>(x + sw(base + 2 * t));
                v[2 * t] = make_float2(q.x, q.y);
                v[2 * t + 1] = make_float2(q.z, q.w);
            }
        } else {
#pragma unroll
            for (int t = 0; t < R; ++t) v[t] = x[sw(base + t)];
        }
        butterfly<R>(v);  // innermost forward radix (stride 1: no twiddles)
#pragma unroll
        for (int t = 0; t < R; ++t) {  // x Khat, swap re/im: the following forward passes compute the inverse
            const float w = ldg_stream(kh + base + t);
            v[t] = make_float2(v[t].y * w, v[t].x * w);
        }
        butterfly<R>(v);  // innermost radix of the inverse
        if ((R & 1) == 0) {
#pragma unroll
            for (int t = 0; t < R / 2; ++t)
                *reinterpret_cast<float4 *>(x + sw(base + 2 * t)) =
                    make_float4(v[2 * t].x, v[2 * t].y, v[2 * t + 1].x, v[2 * t + 1].y);
        } else {
#pragma unroll
            for (int t = 0; t < R; ++t) x[sw(base + t)] = v[t];
        }
    }
    group_sync(g);
}

template <int N, int B, int GS, int TWOFF, int R, int... REST>
__device__ __forceinline__ void ip_conv(float2 *__restrict__ x, const float2 *__restrict__ tw,
                                        const float *__restrict__ kh, const Group &g) {
    if constexpr (sizeof...(REST) == 0) {
        static_assert(B == R, "radices must multiply to N");
        ip_mid<R, N, GS>(x, kh, g);
    } else {
        ip_pass<R, N, B, GS, false>(x, tw + TWOFF, g);
        ip_conv<N, B / R, GS, TWOFF + (R - 1) * (B / R), REST...>(x, tw, kh, g);
        ip_pass<R, N, B, GS, true>(x, tw + TWOFF, g);
    }
}

// reflect-101 fill of the r-wide borders of a padded line whose interior [r, r+len) is loaded,
// and zero fill of the tail [len+2r, n).
// `off`: the padded line starts `off` elements into the buffer (row kernels, FftConvArgs::row_off); [0, off) is zeroed.
__device__ __forceinline__ void pad_line(float2 *buf, int len, int r, int n, const Group &g, int off = 0) {
    for (int p = g.tid; p < r; p += g.size) {
        buf[sw(off + p)] = buf[sw(off + r + reflect101(p - r, len))];
        buf[sw(off + r + len + p)] = buf[sw(off + r + reflect101(len + p, len))];
    }
    for (int p = off + len + 2 * r + g.tid; p < n; p += g.size) buf[sw(p)] = make_float2(0.f, 0.f);
    if (g.tid < off) buf[sw(g.tid)] = make_float2(0.f, 0.f);
}

// Row kernels: ROWS image rows per CTA, one thread group per row (ROWS == 2: two 512-thread groups
// on named barriers 1 and 2; ROWS == 1: the whole CTA).  Each row owns two line buffers.
template <int ROWS>
__device__ __forceinline__ Group row_group() {
    if (ROWS == 2) return Group{(int)(threadIdx.x & 511), 512, 1 + (int)(threadIdx.x >> 9)};
    return Group{(int)threadIdx.x, (int)blockDim.x, 0};
}

// The 2-D input LUT (48 KB at n = 64) is gathered nine times per pixel; with ~200 KB of line buffers
// the L1 has no room for it, so each group parks a copy in its idle ping-pong buffer when it fits.
// WAIT = false: the copies are only queued; the caller completes them with stage_lut2d_finish() (after it has
// issued its first frame loads, so the two round trips overlap).
template <bool WAIT = true>
__device__ __forceinline__ Lut2D stage_lut2d(const Lut2D &L, float2 *idle, int capacity_float2, const Group &g) {
    const int nfloat = L.n * L.n * 3;
    if (nfloat > 2 * capacity_float2) return L;
    float *dst = reinterpret_cast<float *>(idle);
    // asynchronous 16-byte copies that bypass L1 (.cg): the copy costs one round trip and leaves the L1
    // to the twiddle tables
    const int nq = (reinterpret_cast<uintptr_t>(L.tab) & 15) == 0 ? nfloat / 4 : 0;
    for (int i = g.tid; i < nq; i += g.size) cp_async_16(dst + 4 * i, L.tab + 4 * i);
    for (int i = 4 * nq + g.tid; i < nfloat; i += g.size) dst[i] = __ldg(L.tab + i);
    if (WAIT) {
        cp_async_wait_all();
        group_sync(g);
    }
    return Lut2D(dst, L.n);
}
__device__ __forceinline__ void stage_lut2d_finish(const Group &g) {
    cp_async_wait_all();
    group_sync(g);
}

// Interleaved frame row y -> 2-D input LUT -> (chan0 + i chan1) into `line` (offset r, swizzled) and, when asked, the
// three exposure planes.  LUT_SMEM: l2.tab points into shared memory (parked by stage_lut2d).
template <int FMT, bool LUT_SMEM>
__device__ __forceinline__ void load_row_xyz(const FftConvArgs &a, const Lut2D &l2, float2 *__restrict__ line, int y,
                                             const Group &g) {
    const int W = a.W, r = a.r + a.row_off;   // offset of pixel 0 in the line buffer
    const bool quad_aligned = (r & 3) == 0;   // a quad's four values: two aligned 16-byte stores (2 x 6 shared-memory
                                              // wavefronts per warp instead of 4 x 6 for four 8-byte stores)
    if ((W & 3) == 0) {  // row starts on a pixel-quad boundary: 128-bit frame loads
        const size_t q0 = (size_t)y * W / 4;
        constexpr int U = 2;  // quads in flight per thread (three fit a 6000-pixel row in one batch but spill)
        float px[U][4][3];
        auto request = [&](int base) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int qx = base + u * g.size;
                if (qx < W / 4) load_quad<FMT>(a.src_xyz, q0 + qx, a.gain, px[u]);
            }
        };
        int base = g.tid;
        request(base);
        if (LUT_SMEM) stage_lut2d_finish(g);  // the table copy and the first frame loads were in flight together
        while (base < W / 4) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int qx = base + u * g.size;
                if (qx < W / 4) {
                    float e[3][4];
                    float2 z[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        lut2d_eval<LUT_SMEM>(l2, px[u][i][0], px[u][i][1], px[u][i][2], e[0][i], e[1][i], e[2][i]);
                        z[i] = make_float2(pick3(a.chan[0], e[0][i], e[1][i], e[2][i]),
                                           pick3(a.chan[1], e[0][i], e[1][i], e[2][i]));
                    }
                    if (quad_aligned) {  // sw() permutes aligned pairs as a unit
                        *reinterpret_cast<float4 *>(line + sw(r + 4 * qx)) = make_float4(z[0].x, z[0].y, z[1].x, z[1].y);
                        *reinterpret_cast<float4 *>(line + sw(r + 4 * qx + 2)) = make_float4(z[2].x, z[2].y, z[3].x, z[3].y);
                    } else {
#pragma unroll
                        for (int i = 0; i < 4; ++i) line[sw(r + 4 * qx + i)] = z[i];
                    }
                    if (a.exp_planar != nullptr) {
#pragma unroll
                        for (int c = 0; c < 3; ++c)
                            __stcg(reinterpret_cast<float4 *>(a.exp_planar + c * a.plane_stride) + q0 + qx,
                                   make_float4(e[c][0], e[c][1], e[c][2], e[c][3]));
                    }
                }
            }
            base += U * g.size;
            if (base < W / 4) request(base);
        }
    } else {
        if (LUT_SMEM) stage_lut2d_finish(g);
        for (int x = g.tid; x < W; x += g.size) {
            float X, Y, Z, e0, e1, e2;
            load_px<FMT>(a.src_xyz, (size_t)y * W + x, a.gain, X, Y, Z);
            lut2d_eval<LUT_SMEM>(l2, X, Y, Z, e0, e1, e2);
            line[sw(r + x)] = make_float2(pick3(a.chan[0], e0, e1, e2), pick3(a.chan[1], e0, e1, e2));
            if (a.exp_planar != nullptr) {
                const size_t idx = (size_t)y * W + x;
                a.exp_planar[idx] = e0;
                a.exp_planar[a.plane_stride + idx] = e1;
                a.exp_planar[2 * a.plane_stride + idx] = e2;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// rows, forward
// ------------------------------------------------------------------------------------------
template <int SRC, int ROWS, int PLAN>  // SRC 0: planar planes, 1 + FMT: interleaved frame (kFmt*) through the 2-D LUT
__global__ void __launch_bounds__(1024, 1)
k_fft_rows_fwd(const __grid_constant__ FftConvArgs a) {
    extern __shared__ __align__(16) float2 fsm[];
    const int W = a.W, H = a.H, r = a.r, n = a.row.n, NC = a.nc;
    const Group g = row_group<ROWS>();
    const int half = ROWS == 2 ? (int)(threadIdx.x >> 9) : 0;
    const int y0 = a.row0 + blockIdx.x * ROWS;                              // a.row0 is a multiple of ROWS
    const int yend = a.row_count > 0 ? min(H, a.row0 + a.row_count) : H;
    const int nrows = ROWS == 1 ? 1 : min(ROWS, yend - y0);   // the grid has exactly ceil(rows / ROWS) CTAs
    float2 *bufA = fsm + (size_t)half * 2 * n, *bufB = bufA + n;
    const int y = y0 + half;
    if (half < nrows) {
        constexpr int FMT = SRC > 0 ? SRC - 1 : 0;
        if (SRC == 0) {
            for (int x = g.tid; x < W; x += g.size) {
                const size_t idx = (size_t)y * W + x;
                bufA[sw(r + a.row_off + x)] = make_float2(a.src_planar[(size_t)a.chan[0] * a.plane_stride + idx],
                                          a.src_planar[(size_t)a.chan[1] * a.plane_stride + idx]);
            }
        } else {
            const Lut2D l2 = stage_lut2d<false>(a.lut2d, bufB, n, g);   // completed inside load_row_xyz
            {   // ask L2 for the frame row that the CTA taking this one's place will read (two CTAs per SM resident)
                constexpr int BPP = FMT == kFmtF32x3 ? 12 : (FMT == kFmtF32x4 ? 16 : (FMT == kFmtU16x3 ? 6 : 8));
                const int ya = y + ROWS * a.rows_ahead;
                if (a.rows_ahead > 0 && ya < yend) {
                    const char *rowp = static_cast<const char *>(a.src_xyz) + (size_t)ya * W * BPP;
                    for (int off = g.tid * 128; off < W * BPP; off += g.size * 128)
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(rowp + off));
                }
            }
            if (l2.tab != a.lut2d.tab) load_row_xyz<FMT, true>(a, l2, bufA, y, g);   // parked in shared memory
            else load_row_xyz<FMT, false>(a, l2, bufA, y, g);
        }
        group_sync(g);
        pad_line(bufA, W, r, n, g, a.row_off);
        group_sync(g);
        fft_run<PLAN>(bufA, bufB, a.row, g);
    }
    __syncthreads();
    // blocked store: S[(b*H + y)*NC + c]; the NC columns of the CTA's rows are contiguous
    const size_t res_off = (a.row.nrad & 1) ? (size_t)n : 0;  // result buffer by pass parity
    const int nblk = n / NC;
    for (int b = threadIdx.x; b < nblk; b += blockDim.x) {  // one column block per thread: no div/mod
        for (int row = 0; row < nrows; ++row) {
            const float2 *line = fsm + (size_t)row * 2 * n + res_off;
            float2 *dp = a.S + ((size_t)b * H + (y0 + row)) * NC;
            if (NC == 4) {
                const float4 lo = *reinterpret_cast<const float4 *>(line + sw(4 * b));
                const float4 hi = *reinterpret_cast<const float4 *>(line + sw(4 * b + 2));
                *reinterpret_cast<float4 *>(dp) = lo;
                *reinterpret_cast<float4 *>(dp + 2) = hi;
            } else {
                for (int c = 0; c < NC; ++c) dp[c] = line[sw(b * NC + c)];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// columns: forward FFT, * Khat, inverse FFT, fused.  NC columns per CTA (one contiguous block of
// S), a.col_groups thread groups, each with its own ping-pong buffer.
// ------------------------------------------------------------------------------------------
template <int PLAN>
__global__ void __launch_bounds__(1024, 1)
k_fft_cols(const __grid_constant__ FftConvArgs a) {
    extern __shared__ __align__(16) float2 fsm[];
    const int H = a.H, r = a.r, n = a.col.n, NC = a.nc, NG = a.col_groups;
    const int pitch = n + 2;  // even (16-byte aligned lines), de-phases the column buffers across banks
    const int b = blockIdx.x;
    float2 *blk = a.S + (size_t)b * H * NC;
    if (NC == 4) {  // one spectrum row (4 values = 32 bytes) per thread; the loads of a batch are issued together
        constexpr int U = 4;
        for (int y0 = threadIdx.x; y0 < H; y0 += U * blockDim.x) {
            float4 lo[U], hi[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int y = y0 + u * blockDim.x;
                if (y < H) {
                    const float2 *sp = blk + (size_t)y * 4;
                    lo[u] = ldg_stream4(reinterpret_cast<const float4 *>(sp));
                    hi[u] = ldg_stream4(reinterpret_cast<const float4 *>(sp + 2));
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int y = y0 + u * blockDim.x;
                if (y < H) {
                    const int e = sw(r + y);
                    fsm[e] = make_float2(lo[u].x, lo[u].y);
                    fsm[(size_t)pitch + e] = make_float2(lo[u].z, lo[u].w);
                    fsm[(size_t)2 * pitch + e] = make_float2(hi[u].x, hi[u].y);
                    fsm[(size_t)3 * pitch + e] = make_float2(hi[u].z, hi[u].w);
                }
            }
        }
    } else {
        for (int y = threadIdx.x; y < H; y += blockDim.x) {
            const float2 *sp = blk + (size_t)y * NC;
            for (int c = 0; c < NC; ++c) fsm[(size_t)c * pitch + sw(r + y)] = sp[c];
        }
    }
    __syncthreads();
    const int gsize = (int)blockDim.x / NG;
    const int gi = (int)threadIdx.x / gsize;
    const Group g{(int)threadIdx.x % gsize, gsize, NG == 1 ? 0 : 1 + gi};
    float2 *tmp = fsm + (size_t)(NC + gi) * pitch;
    for (int c = gi; c < NC; c += NG) {
        float2 *home = fsm + (size_t)c * pitch;
        pad_line(home, H, r, n, g);
        group_sync(g);
        // forward FFT; its last pass multiplies by the real kernel spectrum and swaps re/im, so the next
        // forward FFT is the inverse
        // the spectrum of an even kernel is mirror-symmetric in the column frequency: columns v and Wp - v read the
        // same row of khat, which halves its DRAM footprint (100 -> 50 MB at 24 MP; the second reader hits L2)
        const int vcol = b * NC + c, vm = vcol <= a.row.n - vcol ? vcol : a.row.n - vcol;
        const float *kh = a.khat + (size_t)vm * n;
        float2 *spec = fft_run<PLAN, true>(home, tmp, a.col, g, kh);
        float2 *other = spec == home ? tmp : home;
        fft_run<PLAN>(spec, other, a.col, g);  // 2 * nrad passes in total: the result is back in `home`
    }
    __syncthreads();
    // keep the swapped form: k_fft_rows_inv consumes swap(x) directly
    for (int y = threadIdx.x; y < H; y += blockDim.x) {
        float2 *dp = blk + (size_t)y * NC;
        if (NC == 4) {
            const int e = sw(r + y);
            const float2 c0 = fsm[e], c1 = fsm[(size_t)pitch + e];
            const float2 c2 = fsm[(size_t)2 * pitch + e], c3 = fsm[(size_t)3 * pitch + e];
            *reinterpret_cast<float4 *>(dp) = make_float4(c0.x, c0.y, c1.x, c1.y);
            *reinterpret_cast<float4 *>(dp + 2) = make_float4(c2.x, c2.y, c3.x, c3.y);
        } else {
            for (int c = 0; c < NC; ++c) dp[c] = fsm[(size_t)c * pitch + sw(r + y)];
        }
    }
}

// In-place variant (compile-time plans only): one 256-thread group per column, n float2 per column.
// IPLAN 1: n = 4096 (8 8 8 8), IPLAN 2: n = 6912 (3 3 3 4 8 8: the line's radices reversed); a.khat rows are
// permuted by the line's perm[].
// NCOL columns per CTA: 4 = a whole 32-byte block row per thread, one 1024-thread CTA per SM; 2 = half a block
// (16 bytes per row) in a 512-thread CTA, two (n = 4096: 64 KB each) per SM.  With one CTA per SM nothing overlaps
// its load and store phases (a quarter of the kernel's stall samples at 24 MP); two independent half-block CTAs
// overlap them with each other's transforms.  The two halves of a 32-byte sector are requested by neighbouring
// CTAs within microseconds of each other, so DRAM still moves every sector once (L2 merges them).
template <int IPLAN, int NCOL>
__global__ void __launch_bounds__(256 * NCOL, 4 / NCOL)
k_fft_cols_ip(const __grid_constant__ FftConvArgs a) {
    extern __shared__ __align__(16) float2 fsm[];
    constexpr int n = IPLAN == 1 ? 4096 : 6912;
    constexpr int GS = 256, NT = 256 * NCOL;
    const int H = a.H, r = a.r;
    const int b = NCOL == 4 ? blockIdx.x : blockIdx.x >> 1;
    const int c0 = NCOL == 4 ? 0 : 2 * (blockIdx.x & 1);   // first column of the block this CTA transforms
    float2 *blk = a.S + (size_t)b * H * 4 + c0;
    if (a.cols_prefetch) {
        // L2 prefetches, one per 128-byte line: this CTA's rows of the kernel spectrum (read in the middle of the
        // transform) and the block of the CTA that will take this one's place (a.rows_ahead CTAs ahead in launch
        // order).  Pays when the kernel spectrum does not stay in L2 between frames: 61 MP 0.870 -> 0.829 ms, but
        // 24 MP (50 MB of spectrum, L2-resident) 0.210 -> 0.214 ms, so r2f_api.cu sets the flag by size.
        const int gi0 = (int)threadIdx.x / 256;
        const int vc = b * 4 + c0 + gi0, vmm = vc <= a.row.n - vc ? vc : a.row.n - vc;
        const float *khp = a.khat + (size_t)vmm * n;
        for (int i = ((int)threadIdx.x % 256) * 32; i < n; i += 256 * 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(khp + i));
        const int ahead = (int)blockIdx.x + a.rows_ahead * (NCOL == 4 ? 1 : 2);
        if (a.rows_ahead > 0 && ahead < (int)gridDim.x && (NCOL == 4 || (blockIdx.x & 1) == 0)) {
            const char *nb = reinterpret_cast<const char *>(a.S + (size_t)(NCOL == 4 ? ahead : ahead >> 1) * H * 4);
            for (int off = (int)threadIdx.x * 128; off < H * 32; off += NT * 128)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(nb + off));
        }
    }
    {
        constexpr int U = 4;
        for (int y0 = threadIdx.x; y0 < H; y0 += U * NT) {
            float4 lo[U], hi[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int y = y0 + u * NT;
                if (y < H) {
                    const float2 *sp = blk + (size_t)y * 4;
                    lo[u] = ldg_stream4(reinterpret_cast<const float4 *>(sp));
                    if (NCOL == 4) hi[u] = ldg_stream4(reinterpret_cast<const float4 *>(sp + 2));
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int y = y0 + u * NT;
                if (y < H) {
                    const int e = sw(r + y);
                    fsm[e] = make_float2(lo[u].x, lo[u].y);
                    fsm[n + e] = make_float2(lo[u].z, lo[u].w);
                    if (NCOL == 4) {
                        fsm[2 * n + e] = make_float2(hi[u].x, hi[u].y);
                        fsm[3 * n + e] = make_float2(hi[u].z, hi[u].w);
                    }
                }
            }
        }
    }
    __syncthreads();
    const int gi = (int)threadIdx.x / GS;
    const Group g{(int)threadIdx.x % GS, GS, 1 + gi};
    float2 *home = fsm + (size_t)gi * n;
    pad_line(home, H, r, n, g);
    group_sync(g);
    const int vcol = b * 4 + c0 + gi, vm = vcol <= a.row.n - vcol ? vcol : a.row.n - vcol;  // khat[v] == khat[Wp - v]
    const float *kh = a.khat + (size_t)vm * n;
    if constexpr (IPLAN == 1) ip_conv<4096, 4096, GS, 0, 8, 8, 8, 8>(home, a.col.tw_ip, kh, g);
    else ip_conv<6912, 6912, GS, 0, 3, 3, 3, 4, 8, 8>(home, a.col.tw_ip, kh, g);
    __syncthreads();
    // keep the swapped form: k_fft_rows_inv consumes swap(x) directly
    for (int y = threadIdx.x; y < H; y += NT) {
        float2 *dp = blk + (size_t)y * 4;
        const int e = sw(r + y);
        const float2 v0 = fsm[e], v1 = fsm[n + e];
        *reinterpret_cast<float4 *>(dp) = make_float4(v0.x, v0.y, v1.x, v1.y);
        if (NCOL == 4) {
            const float2 v2 = fsm[2 * n + e], v3 = fsm[3 * n + e];
            *reinterpret_cast<float4 *>(dp + 2) = make_float4(v2.x, v2.y, v3.x, v3.y);
        }
    }
}

// ------------------------------------------------------------------------------------------
// rows, inverse + epilogue
// ------------------------------------------------------------------------------------------
template <int SRC, int DENSITY, int ROWS, int PLAN>
__global__ void __launch_bounds__(1024, 1)
k_fft_rows_inv(const __grid_constant__ FftConvArgs a) {
    extern __shared__ __align__(16) float2 fsm[];
    const int W = a.W, H = a.H, r = a.r + a.row_off, n = a.row.n, NC = a.nc;   // r: offset of pixel 0 in the line
    const Group g = row_group<ROWS>();
    const int half = ROWS == 2 ? (int)(threadIdx.x >> 9) : 0;
    const int y0 = blockIdx.x * ROWS;
    const int nrows = ROWS == 1 ? 1 : min(ROWS, H - y0);   // the grid has exactly ceil(H / ROWS) CTAs
    // S holds swap(column-inverse); one more forward FFT along the row completes swap(IFFT2)
    const int nblk = n / NC;
    for (int b = threadIdx.x; b < nblk; b += blockDim.x) {
        for (int row = 0; row < nrows; ++row) {
            const float2 *sp = a.S + ((size_t)b * H + (y0 + row)) * NC;
            float2 *line = fsm + (size_t)row * 2 * n;
            if (NC == 4) {
                cp_async_16(reinterpret_cast<float *>(line + sw(4 * b)), reinterpret_cast<const float *>(sp));
                cp_async_16(reinterpret_cast<float *>(line + sw(4 * b + 2)), reinterpret_cast<const float *>(sp + 2));
            } else {
                for (int c = 0; c < NC; ++c) line[sw(b * NC + c)] = sp[c];
            }
        }
    }
    if (SRC == 0) {  // the epilogue's exposure rows: ask L2 for them now, one request per 128-byte line
        for (int row = 0; row < nrows; ++row) {
            const size_t off = (size_t)(y0 + row) * W;
            for (int i = threadIdx.x; i < 3 * ((W + 31) / 32); i += blockDim.x) {
                const int c = i / ((W + 31) / 32), seg = i - c * ((W + 31) / 32);
                const float *p = a.src_planar + (size_t)c * a.plane_stride + off + min(32 * seg, W - 1);
                asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
            }
        }
    }
    cp_async_wait_all();
    __syncthreads();
    if (half >= nrows) return;
    float2 *bufA = fsm + (size_t)half * 2 * n, *bufB = bufA + n;
    float2 *buf = fft_run<PLAN>(bufA, bufB, a.row, g);
    const size_t ps = a.plane_stride;
    const int y = y0 + half;
    constexpr int FMT = SRC > 0 ? SRC - 1 : 0;
    Lut2D l2 = a.lut2d;
    if (SRC != 0) l2 = stage_lut2d(a.lut2d, buf == bufA ? bufB : bufA, n, g);
    // out = alpha * (K (*) x) + beta * x on the two filtered layers, third layer passes through;
    // then (optionally) log10 + H-D curve
    auto finish_px = [&](const float (&src)[3], float2 zs, float (&out)[3]) {
        // compile-time register indices only (a runtime-indexed src[]/out[] would live in local memory)
        const float x0 = pick3(a.chan[0], src[0], src[1], src[2]), x1 = pick3(a.chan[1], src[0], src[1], src[2]);
        const float f0 = fmaf(a.alpha[0], zs.y, a.beta[0] * x0);   // swapped: .y = K(*)chan0
        const float f1 = fmaf(a.alpha[1], zs.x, a.beta[1] * x1);   //          .x = K(*)chan1
#pragma unroll
        for (int c = 0; c < 3; ++c) out[c] = a.chan[0] == c ? f0 : (a.chan[1] == c ? f1 : src[c]);
        if (DENSITY) {
#pragma unroll
            for (int c = 0; c < 3; ++c) out[c] = density_eval_fast(a.curve, c, out[c], a.eps);
        }
    };
    if ((W & 3) == 0 && SRC == 0) {
        // planar hand-off (the exposure planes of this row were prefetched into L2 while the transform ran)
        const size_t q0 = (size_t)y * W / 4;
        for (int qx = g.tid; qx < W / 4; qx += g.size) {
            float px[4][3], res[3][4];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float4 v = __ldcs(reinterpret_cast<const float4 *>(a.src_planar + c * ps) + q0 + qx);
                px[0][c] = v.x; px[1][c] = v.y; px[2][c] = v.z; px[3][c] = v.w;
            }
            float2 z[4];
            if ((r & 3) == 0) {  // two aligned 16-byte reads instead of four 8-byte ones (sw() keeps aligned pairs together)
                const float4 z01 = *reinterpret_cast<const float4 *>(buf + sw(r + 4 * qx));
                const float4 z23 = *reinterpret_cast<const float4 *>(buf + sw(r + 4 * qx + 2));
                z[0] = make_float2(z01.x, z01.y); z[1] = make_float2(z01.z, z01.w);
                z[2] = make_float2(z23.x, z23.y); z[3] = make_float2(z23.z, z23.w);
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) z[i] = buf[sw(r + 4 * qx + i)];
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float out[3];
                finish_px(px[i], z[i], out);
                res[0][i] = out[0]; res[1][i] = out[1]; res[2][i] = out[2];
            }
#pragma unroll
            for (int c = 0; c < 3; ++c)
                reinterpret_cast<float4 *>(a.dst_planar + c * ps)[q0 + qx] =
                    make_float4(res[c][0], res[c][1], res[c][2], res[c][3]);
        }
    } else if ((W & 3) == 0) {
        const size_t q0 = (size_t)y * W / 4;
        for (int qx = g.tid; qx < W / 4; qx += g.size) {
            float px[4][3], res[3][4];
            load_quad<FMT>(a.src_xyz, q0 + qx, a.gain, px);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float src[3], out[3];
                lut2d_eval(l2, px[i][0], px[i][1], px[i][2], src[0], src[1], src[2]);
                finish_px(src, buf[sw(r + 4 * qx + i)], out);
                res[0][i] = out[0]; res[1][i] = out[1]; res[2][i] = out[2];
            }
#pragma unroll
            for (int c = 0; c < 3; ++c)
                reinterpret_cast<float4 *>(a.dst_planar + c * ps)[q0 + qx] =
                    make_float4(res[c][0], res[c][1], res[c][2], res[c][3]);
        }
    } else {
        for (int x = g.tid; x < W; x += g.size) {
            const size_t idx = (size_t)y * W + x;
            float src[3], out[3];
            if (SRC == 0) {
#pragma unroll
                for (int c = 0; c < 3; ++c) src[c] = a.src_planar[c * ps + idx];
            } else {
                float X, Y, Z;
                load_px<FMT>(a.src_xyz, idx, a.gain, X, Y, Z);
                lut2d_eval(l2, X, Y, Z, src[0], src[1], src[2]);
            }
            finish_px(src, buf[sw(r + x)], out);
            const size_t oidx = a.dst_pitch > 0 ? (size_t)y * a.dst_pitch + x : idx;
#pragma unroll
            for (int c = 0; c < 3; ++c) a.dst_planar[c * ps + oidx] = out[c];
        }
    }
}

// ------------------------------------------------------------------------------------------
// kernel spectrum: for a kernel even in both axes,
//   Khat[u][v] = sum_di sum_dj K[c+di][c+dj] cos(2 pi u di / Hp) cos(2 pi v dj / Wp)  (real),
// evaluated in binary64 in two separable stages, stored transposed [Wp][Hp] as float32 with the
// 1/(Hp*Wp) normalisation of the inverse transform folded in.
// ------------------------------------------------------------------------------------------
__global__ void k_khat_stage1(const float *__restrict__ kern /* k x k base */, int k, int Wp,
                              const double *__restrict__ cosW, double *__restrict__ A /* k x Wp */) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (v >= Wp) return;
    const int c = k / 2;
    double acc = 0.0;
    for (int j = 0; j < k; ++j) {
        const int dj = j - c;
        const long long q = ((long long)v * (dj < 0 ? -dj : dj)) % Wp;
        acc = fma((double)kern[i * k + j], cosW[q], acc);
    }
    A[(size_t)i * Wp + v] = acc;
}

__global__ void k_khat_stage2(const double *__restrict__ A, int k, int Hp, int Wp, const double *__restrict__ cosH,
                              double norm, float *__restrict__ khat /* [Wp][Hp] */, const int *__restrict__ perm) {
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    const int v = blockIdx.y;
    if (u >= Hp) return;
    const int c = k / 2;
    double acc = 0.0;
    for (int i = 0; i < k; ++i) {
        const int di = i - c;
        const long long q = ((long long)u * (di < 0 ? -di : di)) % Hp;
        acc = fma(A[(size_t)i * Wp + v], cosH[q], acc);
    }
    khat[(size_t)v * Hp + (perm ? perm[u] : u)] = (float)(acc * norm);
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static bool factor_235(int n, std::vector<int> &rad) {
    rad.clear();
    int m = n, twos = 0;
    while (m % 2 == 0) { m /= 2; ++twos; }
    std::vector<int> odd;
    while (m % 3 == 0) { m /= 3; odd.push_back(3); }
    while (m % 5 == 0) { m /= 5; odd.push_back(5); }
    if (m != 1) return false;
    while (twos >= 3) { rad.push_back(8); twos -= 3; }
    if (twos == 2) rad.push_back(4);
    if (twos == 1) rad.push_back(2);
    rad.insert(rad.end(), odd.begin(), odd.end());
    return (int)rad.size() <= kFftMaxPasses;
}

// Padded FFT length: among the 2^a 3^b 5^c sizes in [min_n, 1.3 min_n] pick the one with the lowest
// estimated cost  n * sum_passes(weight(radix))  -- a power of two a few percent larger (radix-8
// passes) beats the smallest smooth size made of many radix-3/5 passes.
int fft_good_size(int min_n, int multiple_of) {
    int best = 0;
    double best_cost = 0.0;
    const int hi = min_n + min_n * 3 / 10 + 64;
    for (int n = min_n; n <= hi && n <= kFftMaxLen; ++n) {
        if (n % multiple_of) continue;
        std::vector<int> rad;
        if (!factor_235(n, rad)) continue;
        double w = 0.0;
        for (int R : rad) w += R == 8 ? 1.3 : R == 5 ? 1.25 : R == 2 ? 0.9 : 1.0;
        const double cost = w * (double)n;
        if (!best || cost < best_cost) {
            best = n;
            best_cost = cost;
        }
    }
    return best;
}

bool fft_make_line(int n, FftLineHost &out) {
    std::vector<int> rad;
    if (!factor_235(n, rad)) return false;
    out.n = n;
    out.rad = rad;
    out.cosines.resize(n);
    const double two_pi = 6.283185307179586476925286766559;
    for (int q = 0; q < n; ++q) out.cosines[q] = std::cos(two_pi * (double)q / (double)n);
    // per-pass twiddles, laid out [t-1][k] so that a warp reads consecutive entries
    out.roots.clear();
    out.tw_off.clear();
    int Ns = 1;
    for (int R : rad) {
        out.tw_off.push_back((int)out.roots.size());
        for (int t = 1; t < R; ++t)
            for (int k = 0; k < Ns; ++k) {
                const double ang = two_pi * (double)t * (double)k / ((double)Ns * (double)R);
                out.roots.push_back(make_float2((float)std::cos(ang), (float)(-std::sin(ang))));
            }
        Ns *= R;
    }
    // in-place passes: block B, stride S = B / R, entry (t-1)*S + k = exp(-2 pi i t k / B); the innermost pass
    // (S = 1) has no twiddles.  The in-place kernels run the radices in REVERSE order (odd radices first, radix 8
    // innermost): the early passes then have strides that are multiples of 128 (one swizzle per butterfly) and the
    // fused innermost step works on 8 contiguous elements with 16-byte accesses.
    const std::vector<int> rad_ip(rad.rbegin(), rad.rend());
    out.roots_ip.clear();
    int B = n;
    for (int R : rad_ip) {
        const int S = B / R;
        if (S > 1)
            for (int t = 1; t < R; ++t)
                for (int k = 0; k < S; ++k) {
                    const double ang = two_pi * (double)t * (double)k / (double)B;
                    out.roots_ip.push_back(make_float2((float)std::cos(ang), (float)(-std::sin(ang))));
                }
        B = S;
    }
    if (out.roots_ip.empty()) out.roots_ip.push_back(make_float2(1.f, 0.f));
    // X[k] with k = k1 + R1 k2 + R1 R2 k3 + ... ends at position k1 n/R1 + k2 n/(R1 R2) + ...
    out.perm.resize(n);
    for (int k = 0; k < n; ++k) {
        int kk = k, pos = 0, blk = n;
        for (int R : rad_ip) {
            blk /= R;
            pos += (kk % R) * blk;
            kk /= R;
        }
        out.perm[k] = pos;
    }
    return true;
}

static int inplace_plan_id(const FftLineHost &l) {
    static const char *off = getenv("R2F_FFT_INPLACE");  // tuning knob: "0" keeps the ping-pong column kernel
    if (off && off[0] == '0') return 0;
    if (l.n == 4096 && l.rad == std::vector<int>{8, 8, 8, 8}) return 1;
    if (l.n == 6912 && l.rad == std::vector<int>{8, 8, 4, 3, 3, 3}) return 2;
    return 0;
}
bool fft_cols_inplace_available(int Hp) {
    FftLineHost l;
    l.n = Hp;
    return factor_235(Hp, l.rad) && inplace_plan_id(l) != 0;
}
size_t fft_cols_inplace_smem(int Hp) { return (size_t)4 * Hp * sizeof(float2); }

constexpr size_t kFftMaxSmem = 227 * 1024;

// rows per CTA: two when both rows' ping-pong pairs fit in shared memory
static int rows_per_cta(int Wp) {
    static const char *force = getenv("R2F_FFT_ROWS");  // tuning knob (1 = one row per 512-thread CTA, 2 CTAs/SM)
    if (force && force[0] == '1') return 1;
    if (force && force[0] == '2') return (size_t)4 * Wp * sizeof(float2) <= kFftMaxSmem ? 2 : 1;
    // one row per 512-thread CTA when two such CTAs fit an SM (measured 1.6 % faster at 24 MP than two
    // rows in one 1024-thread CTA); else two rows per CTA if they fit, else one
    if ((size_t)2 * Wp * sizeof(float2) <= 100 * 1024) return 1;
    return (size_t)4 * Wp * sizeof(float2) <= kFftMaxSmem ? 2 : 1;
}
size_t fft_rows_smem(int Wp) { return (size_t)rows_per_cta(Wp) * 2 * Wp * sizeof(float2); }
size_t fft_cols_smem(int Hp, int nc, int groups) { return (size_t)(nc + groups) * (Hp + 2) * sizeof(float2); }

// Column-pass geometry: the widest column block (<= 4) dividing Wp that fits, with two thread
// groups when their ping-pong buffers fit too.
bool fft_col_geometry(int Hp, int Wp, int &nc, int &groups) {
    for (int c = 4; c >= 2; --c) {
        if (Wp % c) continue;
        for (int gcount = 2; gcount >= 1; --gcount) {
            if (gcount > c) continue;
            if (fft_cols_smem(Hp, c, gcount) <= kFftMaxSmem) {
                nc = c;
                groups = gcount;
                return true;
            }
        }
    }
    return false;
}

cudaError_t launch_khat(const float *base_kernel_dev, int k, int Hp, int Wp, const double *cosH_dev,
                        const double *cosW_dev, double *scratchA_dev, float *khat_dev, const int *perm_dev,
                        cudaStream_t st) {
    dim3 g1((Wp + 255) / 256, k);
    k_khat_stage1<<<g1, 256, 0, st>>>(base_kernel_dev, k, Wp, cosW_dev, scratchA_dev);
    dim3 g2((Hp + 255) / 256, Wp);
    k_khat_stage2<<<g2, 256, 0, st>>>(scratchA_dev, k, Hp, Wp, cosH_dev, 1.0 / ((double)Hp * (double)Wp), khat_dev,
                                      perm_dev);
    return cudaGetLastError();
}

template <typename K>
static cudaError_t set_smem(K kfn, size_t bytes) {
    return cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

// Compile-time plan matching a line (length, radix sequence) and the thread-group size, or 0.
static int static_plan_id(const FftLine &l, int group_threads) {
    auto is = [&](int n, int gs, std::initializer_list<int> rad) {
        if (l.n != n || group_threads != gs || l.nrad != (int)rad.size()) return false;
        int i = 0;
        for (int r : rad)
            if (l.rad[i++] != r) return false;
        return true;
    };
    static const char *off = getenv("R2F_FFT_STATIC");  // tuning knob: "0" forces the run-time plans
    if (off && off[0] == '0') return 0;
    if (is(6144, 512, {8, 8, 8, 4, 3})) return 1;
    if (is(4096, 512, {8, 8, 8, 8})) return 2;
    if (is(10240, 1024, {8, 8, 8, 4, 5})) return 3;
    if (is(6912, 512, {8, 8, 4, 3, 3, 3})) return 4;
    return 0;
}

template <int ROWS, int PLAN>
static cudaError_t launch_rows_fwd(const FftConvArgs &a, int src_mode, int ctas, int threads, size_t smem,
                                   cudaStream_t st) {
    cudaError_t e;
#define R2F_FWD(M)                                                                       \
    do {                                                                                 \
        if ((e = set_smem(k_fft_rows_fwd<M, ROWS, PLAN>, smem)) != cudaSuccess) return e; \
        k_fft_rows_fwd<M, ROWS, PLAN><<<ctas, threads, smem, st>>>(a);                   \
    } while (0)
    switch (src_mode) {
        case 0: R2F_FWD(0); break;
        case 1: R2F_FWD(1); break;
        case 2: R2F_FWD(2); break;
        case 3: R2F_FWD(3); break;
        default: R2F_FWD(4); break;
    }
#undef R2F_FWD
    return cudaGetLastError();
}

template <int ROWS, int PLAN>
static cudaError_t launch_rows_inv(const FftConvArgs &a, int src_mode, bool density, int ctas, int threads,
                                   size_t smem, cudaStream_t st) {
    cudaError_t e;
#define R2F_INV(M, D)                                                                       \
    do {                                                                                    \
        if ((e = set_smem(k_fft_rows_inv<M, D, ROWS, PLAN>, smem)) != cudaSuccess) return e; \
        k_fft_rows_inv<M, D, ROWS, PLAN><<<ctas, threads, smem, st>>>(a);                   \
    } while (0)
    if constexpr (PLAN != 0) {  // compile-time plans are only built for the planar source (the render path's hand-off)
        if (src_mode != 0) return cudaErrorInvalidValue;
        if (density) R2F_INV(0, 1);
        else R2F_INV(0, 0);
    } else
    switch (src_mode * 2 + (density ? 1 : 0)) {
        case 0: R2F_INV(0, 0); break;
        case 1: R2F_INV(0, 1); break;
        case 2: R2F_INV(1, 0); break;
        case 3: R2F_INV(1, 1); break;
        case 4: R2F_INV(2, 0); break;
        case 5: R2F_INV(2, 1); break;
        case 6: R2F_INV(3, 0); break;
        case 7: R2F_INV(3, 1); break;
        case 8: R2F_INV(4, 0); break;
        default: R2F_INV(4, 1); break;
    }
#undef R2F_INV
    return cudaGetLastError();
}

template <int PLAN>
static cudaError_t launch_cols(const FftConvArgs &a, int ctas, size_t smem, cudaStream_t st) {
    cudaError_t e;
    if ((e = set_smem(k_fft_cols<PLAN>, smem)) != cudaSuccess) return e;
    k_fft_cols<PLAN><<<ctas, 1024, smem, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_fft_conv(const FftConvArgs &a, int src_mode, bool density, cudaStream_t st, int stage) {
    const int rows = rows_per_cta(a.row.n);
    const size_t rs = fft_rows_smem(a.row.n), cs = fft_cols_smem(a.col.n, a.nc, a.col_groups);
    const int row_ctas = (a.H + rows - 1) / rows, col_ctas = a.row.n / a.nc;
    cudaError_t e = cudaSuccess;
    const int t1 = rs <= 100 * 1024 ? 512 : 1024;  // single-row CTAs small enough for two per SM run 512 threads
    const int row_plan = rows == 1 ? static_plan_id(a.row, t1) : 0;
    const int col_plan = static_plan_id(a.col, 1024 / a.col_groups);
    if (stage == 0 || stage == 1) {
        const int fwd_rows = a.row_count > 0 ? a.row_count : a.H;
        const int fwd_ctas = (fwd_rows + rows - 1) / rows;
        if (a.row0 % rows != 0) return cudaErrorInvalidValue;
        if (rows == 2) e = launch_rows_fwd<2, 0>(a, src_mode, fwd_ctas, 1024, rs, st);
        else if (row_plan == 1) e = launch_rows_fwd<1, 1>(a, src_mode, fwd_ctas, t1, rs, st);
        else if (row_plan == 3) e = launch_rows_fwd<1, 3>(a, src_mode, fwd_ctas, t1, rs, st);
        else e = launch_rows_fwd<1, 0>(a, src_mode, fwd_ctas, t1, rs, st);
        if (e != cudaSuccess) return e;
    }
    if ((stage == 0 || stage == 2) && a.col_inplace) {
        const size_t ips = fft_cols_inplace_smem(a.col.n);
        static const char *ncol_env = getenv("R2F_FFT_COLS_PER_CTA");  // tuning knob: "4" forces whole-block CTAs
        const bool split = 2 * (ips / 2 + 1024) <= 227 * 1024 && !(ncol_env && ncol_env[0] == '4');
        if (a.col.n == 4096 && split) {
            if ((e = set_smem(k_fft_cols_ip<1, 2>, ips / 2)) != cudaSuccess) return e;
            k_fft_cols_ip<1, 2><<<2 * col_ctas, 512, ips / 2, st>>>(a);
        } else if (a.col.n == 4096) {
            if ((e = set_smem(k_fft_cols_ip<1, 4>, ips)) != cudaSuccess) return e;
            k_fft_cols_ip<1, 4><<<col_ctas, 1024, ips, st>>>(a);
        } else if (split) {
            if ((e = set_smem(k_fft_cols_ip<2, 2>, ips / 2)) != cudaSuccess) return e;
            k_fft_cols_ip<2, 2><<<2 * col_ctas, 512, ips / 2, st>>>(a);
        } else {
            if ((e = set_smem(k_fft_cols_ip<2, 4>, ips)) != cudaSuccess) return e;
            k_fft_cols_ip<2, 4><<<col_ctas, 1024, ips, st>>>(a);
        }
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    } else if (stage == 0 || stage == 2) {
        if (col_plan == 2) e = launch_cols<2>(a, col_ctas, cs, st);
        else if (col_plan == 4) e = launch_cols<4>(a, col_ctas, cs, st);
        else e = launch_cols<0>(a, col_ctas, cs, st);
        if (e != cudaSuccess) return e;
    }
    if (stage == 0 || stage == 3) {
        FftConvArgs b = a;
        int inv_mode = src_mode;
        if (src_mode != 0 && a.exp_planar != nullptr) {  // the forward pass left the exposure planes behind
            b.src_planar = a.exp_planar;
            inv_mode = 0;
        }
        if (rows == 2) e = launch_rows_inv<2, 0>(b, inv_mode, density, row_ctas, 1024, rs, st);
        else if (row_plan == 1 && inv_mode == 0) e = launch_rows_inv<1, 1>(b, inv_mode, density, row_ctas, t1, rs, st);
        else if (row_plan == 3 && inv_mode == 0) e = launch_rows_inv<1, 3>(b, inv_mode, density, row_ctas, t1, rs, st);
        else e = launch_rows_inv<1, 0>(b, inv_mode, density, row_ctas, t1, rs, st);
    }
    return e;
}

}  // namespace r2f
