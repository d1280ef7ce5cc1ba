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

#include <cmath>
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

// complex elements a thread may hold during one pass: 16 with 512 threads (<=128 registers),
// 12 with 1024 threads (<=64 registers)
__host__ __device__ constexpr int max_elems_for(int threads) { return threads <= 512 ? 16 : 12; }

// One in-place Stockham pass of radix R over buf[0..n): every thread first pulls its butterflies
// into registers, the CTA synchronises, then results go back to their auto-sorted positions.
template <int R, int MAXE>
__device__ __forceinline__ void fft_pass(float2 *buf, int n, int Ns, const float2 *__restrict__ tw) {
    constexpr int IT = MAXE / R;
    const int nb = n / R;
    const int tws = n / (Ns * R);
    const int tid = threadIdx.x, T = blockDim.x;
    float2 v[IT][R];
#pragma unroll
    for (int it = 0; it < IT; ++it) {
        const int j = tid + it * T;
        if (j < nb) {
#pragma unroll
            for (int t = 0; t < R; ++t) v[it][t] = buf[j + t * nb];
        }
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < IT; ++it) {
        const int j = tid + it * T;
        if (j < nb) {
            const int k = j % Ns;
            if (k != 0) {
#pragma unroll
                for (int t = 1; t < R; ++t) v[it][t] = cmul(v[it][t], __ldg(tw + t * k * tws));
            }
            butterfly<R>(v[it]);
            const int j0 = (j - k) * R + k;
#pragma unroll
            for (int t = 0; t < R; ++t) buf[j0 + t * Ns] = v[it][t];
        }
    }
    __syncthreads();
}

// Forward FFT of buf[0..n) by the whole CTA (all threads must call; ends synchronised).
template <int T>
__device__ __forceinline__ void fft_forward(float2 *buf, const FftLine &L) {
    constexpr int M = max_elems_for(T);
    int Ns = 1;
    for (int s = 0; s < L.nrad; ++s) {
        const int R = L.rad[s];
        switch (R) {
            case 2: fft_pass<2, M>(buf, L.n, Ns, L.tw); break;
            case 3: fft_pass<3, M>(buf, L.n, Ns, L.tw); break;
            case 4: fft_pass<4, M>(buf, L.n, Ns, L.tw); break;
            default: fft_pass<5, M>(buf, L.n, Ns, L.tw); break;
        }
        Ns *= R;
    }
}

// reflect-101 fill of the r-wide borders of a padded line whose interior [r, r+len) is loaded,
// and zero fill of the tail [len+2r, n).  `stride` = distance between consecutive elements.
__device__ __forceinline__ void pad_line(float2 *buf, int len, int r, int n) {
    for (int p = threadIdx.x; p < r; p += blockDim.x) {
        buf[p] = buf[r + reflect101(p - r, len)];
        buf[r + len + p] = buf[r + reflect101(len + p, len)];
    }
    for (int p = len + 2 * r + threadIdx.x; p < n; p += blockDim.x) buf[p] = make_float2(0.f, 0.f);
}

// ------------------------------------------------------------------------------------------
// rows, forward
// ------------------------------------------------------------------------------------------
template <int SRC, int T>  // SRC 0: planar planes, 1: interleaved XYZ through the 2-D LUT (cin 3), 2: same, cin 4
__global__ void __launch_bounds__(T, 1)
k_fft_rows_fwd(const __grid_constant__ FftConvArgs a) {
    extern __shared__ __align__(16) float2 fsm[];
    const int W = a.W, H = a.H, r = a.r, n = a.row.n;
    const int y0 = blockIdx.x * 2;
    const int nrows = min(2, H - y0);
    for (int row = 0; row < nrows; ++row) {
        float2 *buf = fsm + (size_t)row * n;
        const int y = y0 + row;
        for (int x = threadIdx.x; x < W; x += blockDim.x) {
            float2 z;
            if (SRC == 0) {
                const size_t idx = (size_t)y * W + x;
                z.x = a.src_planar[(size_t)a.chan[0] * a.plane_stride + idx];
                z.y = a.src_planar[(size_t)a.chan[1] * a.plane_stride + idx];
            } else {
                const int cin = SRC == 1 ? 3 : 4;
                const float *px = a.src_xyz + ((size_t)y * W + x) * cin;
                float e0, e1, e2;
                lut2d_eval(a.lut2d, __ldg(px), __ldg(px + 1), __ldg(px + 2), e0, e1, e2);
                z.x = e0;
                z.y = e1;
            }
            buf[r + x] = z;
        }
    }
    __syncthreads();
    for (int row = 0; row < nrows; ++row) pad_line(fsm + (size_t)row * n, W, r, n);
    __syncthreads();
    for (int row = 0; row < nrows; ++row) fft_forward<T>(fsm + (size_t)row * n, a.row);
    // blocked store: S[(b*H + y)*NC + c], NC columns of one row are contiguous (NC*8 bytes)
    constexpr int NC = kFftColsPerBlock;
    const int per_row = n;  // n % NC == 0
    for (int idx = threadIdx.x; idx < per_row * nrows; idx += blockDim.x) {
        const int c = idx % NC;
        const int row = (idx / NC) % nrows;
        const int b = idx / (NC * nrows);
        a.S[((size_t)b * H + (y0 + row)) * NC + c] = fsm[(size_t)row * n + b * NC + c];
    }
}

// ------------------------------------------------------------------------------------------
// columns: forward FFT, * Khat, inverse FFT, fused
// ------------------------------------------------------------------------------------------
template <int T>
__global__ void __launch_bounds__(T, 1)
k_fft_cols(const __grid_constant__ FftConvArgs a) {
    extern __shared__ __align__(16) float2 fsm[];
    constexpr int NC = kFftColsPerBlock;
    const int H = a.H, r = a.r, n = a.col.n;
    const int pitch = n + 1;  // de-phase the NC column buffers across banks
    const int b = blockIdx.x;
    float2 *blk = a.S + (size_t)b * H * NC;
    for (int idx = threadIdx.x; idx < H * NC; idx += blockDim.x) {
        const int c = idx % NC, y = idx / NC;
        fsm[(size_t)c * pitch + r + y] = blk[idx];
    }
    __syncthreads();
    for (int c = 0; c < NC; ++c) pad_line(fsm + (size_t)c * pitch, H, r, n);
    __syncthreads();
    for (int c = 0; c < NC; ++c) fft_forward<T>(fsm + (size_t)c * pitch, a.col);
    // product with the real kernel spectrum; swap re/im so the next forward FFT is the inverse
    for (int idx = threadIdx.x; idx < n * NC; idx += blockDim.x) {
        const int c = idx / n, u = idx - c * n;
        const float kh = __ldg(a.khat + (size_t)(b * NC + c) * n + u);
        float2 *p = fsm + (size_t)c * pitch + u;
        const float2 z = *p;
        *p = make_float2(z.y * kh, z.x * kh);
    }
    __syncthreads();
    for (int c = 0; c < NC; ++c) fft_forward<T>(fsm + (size_t)c * pitch, a.col);
    // keep the swapped form: k_fft_rows_inv consumes swap(x) directly
    for (int idx = threadIdx.x; idx < H * NC; idx += blockDim.x) {
        const int c = idx % NC, y = idx / NC;
        blk[idx] = fsm[(size_t)c * pitch + r + y];
    }
}

// ------------------------------------------------------------------------------------------
// rows, inverse + epilogue
// ------------------------------------------------------------------------------------------
template <int SRC, int DENSITY, int T>
__global__ void __launch_bounds__(T, 1)
k_fft_rows_inv(const __grid_constant__ FftConvArgs a) {
    extern __shared__ __align__(16) float2 fsm[];
    constexpr int NC = kFftColsPerBlock;
    const int W = a.W, H = a.H, r = a.r, n = a.row.n;
    const int y0 = blockIdx.x * 2;
    const int nrows = min(2, H - y0);
    // S holds swap(column-inverse); one more forward FFT along the row completes swap(IFFT2)
    for (int idx = threadIdx.x; idx < n * nrows; idx += blockDim.x) {
        const int c = idx % NC;
        const int row = (idx / NC) % nrows;
        const int b = idx / (NC * nrows);
        fsm[(size_t)row * n + b * NC + c] = a.S[((size_t)b * H + (y0 + row)) * NC + c];
    }
    __syncthreads();
    for (int row = 0; row < nrows; ++row) fft_forward<T>(fsm + (size_t)row * n, a.row);
    const size_t ps = a.plane_stride;
    for (int row = 0; row < nrows; ++row) {
        const int y = y0 + row;
        const float2 *buf = fsm + (size_t)row * n;
        for (int x = threadIdx.x; x < W; x += blockDim.x) {
            const size_t idx = (size_t)y * W + x;
            const float2 zs = buf[r + x];  // swapped: .y = K(*)chan0, .x = K(*)chan1
            float src[3];
            if (SRC == 0) {
#pragma unroll
                for (int c = 0; c < 3; ++c) src[c] = a.src_planar[c * ps + idx];
            } else {
                const int cin = SRC == 1 ? 3 : 4;
                const float *px = a.src_xyz + idx * cin;
                lut2d_eval(a.lut2d, __ldg(px), __ldg(px + 1), __ldg(px + 2), src[0], src[1], src[2]);
            }
            float out[3] = {src[0], src[1], src[2]};
            out[a.chan[0]] = fmaf(a.alpha[0], zs.y, a.beta[0] * src[a.chan[0]]);
            out[a.chan[1]] = fmaf(a.alpha[1], zs.x, a.beta[1] * src[a.chan[1]]);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float val = out[c];
                if (DENSITY) val = density_eval(a.curve, c, val, a.eps);
                a.dst_planar[c * ps + idx] = val;
            }
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
                              double norm, float *__restrict__ khat /* [Wp][Hp] */) {
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
    khat[(size_t)v * Hp + u] = (float)(acc * norm);
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
    for (int i = 0; i + 1 < twos; i += 2) rad.push_back(4);
    if (twos & 1) rad.push_back(2);
    rad.insert(rad.end(), odd.begin(), odd.end());
    return (int)rad.size() <= kFftMaxPasses;
}

int fft_threads_for(int n) { return n <= 6656 ? 512 : 1024; }

static bool line_feasible(int n, const std::vector<int> &rad) {
    const int T = fft_threads_for(n);
    for (int R : rad) {
        const int nb = n / R;
        const int per = (nb + T - 1) / T;
        if (per * R > max_elems_for(T)) return false;
    }
    return true;
}

int fft_good_size(int min_n, int multiple_of) {
    for (int n = min_n; n <= kFftMaxLen; ++n) {
        if (n % multiple_of) continue;
        std::vector<int> rad;
        if (factor_235(n, rad) && line_feasible(n, rad)) return n;
    }
    return 0;
}

bool fft_make_line(int n, FftLineHost &out) {
    std::vector<int> rad;
    if (!factor_235(n, rad) || !line_feasible(n, rad)) return false;
    out.n = n;
    out.rad = rad;
    out.roots.resize(n);
    out.cosines.resize(n);
    const double two_pi = 6.283185307179586476925286766559;
    for (int q = 0; q < n; ++q) {
        const double ang = two_pi * (double)q / (double)n;
        out.roots[q] = make_float2((float)std::cos(ang), (float)(-std::sin(ang)));  // exp(-2 pi i q / n)
        out.cosines[q] = std::cos(ang);
    }
    return true;
}

size_t fft_rows_smem(int Wp) { return (size_t)2 * Wp * sizeof(float2); }
size_t fft_cols_smem(int Hp) { return (size_t)kFftColsPerBlock * (Hp + 1) * sizeof(float2); }

cudaError_t launch_khat(const float *base_kernel_dev, int k, int Hp, int Wp, const double *cosH_dev,
                        const double *cosW_dev, double *scratchA_dev, float *khat_dev, cudaStream_t st) {
    dim3 g1((Wp + 255) / 256, k);
    k_khat_stage1<<<g1, 256, 0, st>>>(base_kernel_dev, k, Wp, cosW_dev, scratchA_dev);
    dim3 g2((Hp + 255) / 256, Wp);
    k_khat_stage2<<<g2, 256, 0, st>>>(scratchA_dev, k, Hp, Wp, cosH_dev, 1.0 / ((double)Hp * (double)Wp), khat_dev);
    return cudaGetLastError();
}

template <typename K>
static cudaError_t set_smem(K kfn, size_t bytes) {
    return cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

template <int T>
static cudaError_t launch_rows_fwd(const FftConvArgs &a, int src_mode, int ctas, size_t smem, cudaStream_t st) {
    cudaError_t e;
#define R2F_FWD(M)                                                              \
    do {                                                                        \
        if ((e = set_smem(k_fft_rows_fwd<M, T>, smem)) != cudaSuccess) return e; \
        k_fft_rows_fwd<M, T><<<ctas, T, smem, st>>>(a);                         \
    } while (0)
    if (src_mode == 0) R2F_FWD(0); else if (src_mode == 1) R2F_FWD(1); else R2F_FWD(2);
#undef R2F_FWD
    return cudaGetLastError();
}

template <int T>
static cudaError_t launch_rows_inv(const FftConvArgs &a, int src_mode, bool density, int ctas, size_t smem,
                                   cudaStream_t st) {
    cudaError_t e;
#define R2F_INV(M, D)                                                              \
    do {                                                                           \
        if ((e = set_smem(k_fft_rows_inv<M, D, T>, smem)) != cudaSuccess) return e; \
        k_fft_rows_inv<M, D, T><<<ctas, T, smem, st>>>(a);                         \
    } while (0)
    if (density) {
        if (src_mode == 0) R2F_INV(0, 1); else if (src_mode == 1) R2F_INV(1, 1); else R2F_INV(2, 1);
    } else {
        if (src_mode == 0) R2F_INV(0, 0); else if (src_mode == 1) R2F_INV(1, 0); else R2F_INV(2, 0);
    }
#undef R2F_INV
    return cudaGetLastError();
}

template <int T>
static cudaError_t launch_cols(const FftConvArgs &a, int ctas, size_t smem, cudaStream_t st) {
    cudaError_t e;
    if ((e = set_smem(k_fft_cols<T>, smem)) != cudaSuccess) return e;
    k_fft_cols<T><<<ctas, T, smem, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_fft_conv(const FftConvArgs &a, int src_mode, bool density, cudaStream_t st) {
    const size_t rs = fft_rows_smem(a.row.n), cs = fft_cols_smem(a.col.n);
    const int tr = fft_threads_for(a.row.n), tc = fft_threads_for(a.col.n);
    const int row_ctas = (a.H + 1) / 2, col_ctas = a.row.n / kFftColsPerBlock;
    cudaError_t e = tr == 512 ? launch_rows_fwd<512>(a, src_mode, row_ctas, rs, st)
                              : launch_rows_fwd<1024>(a, src_mode, row_ctas, rs, st);
    if (e != cudaSuccess) return e;
    e = tc == 512 ? launch_cols<512>(a, col_ctas, cs, st) : launch_cols<1024>(a, col_ctas, cs, st);
    if (e != cudaSuccess) return e;
    return tr == 512 ? launch_rows_inv<512>(a, src_mode, density, row_ctas, rs, st)
                     : launch_rows_inv<1024>(a, src_mode, density, row_ctas, rs, st);
}

}  // namespace r2f
