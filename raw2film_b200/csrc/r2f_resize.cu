// Device resolution_scaling: cv2.resize as the reference calls it,
//   reference src/raw2film/utils.py:226-244  INTER_AREA when shrinking, INTER_LANCZOS4 when enlarging;
//   before the path on the float32 frame (cpu_processor.py:122-134, gpu_processor.py:748-758),
//   after it on the uint8 image (cpu_processor.py:411-412).
// The arithmetic follows OpenCV's resize.cpp in operation order (restated and pinned against cv2 in
// oracle/resize_oracle.py): every float operation is a separately rounded multiply or add (this file is compiled
// with -fmad=false like the rest of the library), integer paths are exact.
//
//   k_resize_area      resizeArea_ + computeResizeAreaTab: per destination cell a tap list (source index, weight);
//                      horizontal  buf += src * alpha  in tap order, vertical  sum = beta*buf / sum += beta*buf
//   k_resize_area_int  resizeAreaFast_: integer scale factors, block sums four terms at a time
//   k_resize_lanczos4  HResizeLanczos4 + VResizeLanczos4: 8 x 8 taps, edge-replicated indices; uint8 in 1/2048 fixed
//                      point with the (sum + 2^21) >> 22 cast, float32 as left-to-right sums
// One thread per destination pixel (three channels); the tap tables are built on the host in binary64 exactly as
// OpenCV builds them and cached per (source size, destination size).
#include <cmath>
#include <cstdint>
#include <vector>

#include <cuda_runtime.h>

#include "r2f_resize.h"

namespace r2f {

namespace {

constexpr int kRsThreads = 256;

template <typename T>
struct PixIO;
template <>
struct PixIO<float> {
    static __device__ __forceinline__ float load(const float *p) { return __ldg(p); }
    static __device__ __forceinline__ void store(float *p, float v) { *p = v; }
};
template <>
struct PixIO<uint8_t> {
    static __device__ __forceinline__ float load(const uint8_t *p) { return (float)__ldg(p); }
    // saturate_cast<uchar>(float): round half to even, clamp
    static __device__ __forceinline__ void store(uint8_t *p, float v) {
        const int r = __float2int_rn(v);
        *p = (uint8_t)min(max(r, 0), 255);
    }
};

template <typename T>
__global__ void __launch_bounds__(kRsThreads)
k_resize_area(const T *__restrict__ src, int cin, int sw, T *__restrict__ dst, int dw, int dh, AreaTabDev xt,
              AreaTabDev yt) {
    const size_t total = (size_t)dw * dh;
    for (size_t p = (size_t)blockIdx.x * kRsThreads + threadIdx.x; p < total; p += (size_t)gridDim.x * kRsThreads) {
        const int dy = (int)(p / dw), dx = (int)(p - (size_t)dy * dw);
        const int x0 = xt.start[dx], xn = xt.start[dx + 1] - x0;
        const int y0 = yt.start[dy], yn = yt.start[dy + 1] - y0;
        float sum[3] = {0.f, 0.f, 0.f};
        for (int j = 0; j < yn; ++j) {
            const float beta = yt.alpha[y0 + j];
            const T *row = src + (size_t)yt.si[y0 + j] * sw * cin;
            float buf[3] = {0.f, 0.f, 0.f};
            for (int k = 0; k < xn; ++k) {
                const float alpha = xt.alpha[x0 + k];
                const T *px = row + (size_t)xt.si[x0 + k] * cin;
#pragma unroll
                for (int c = 0; c < 3; ++c) buf[c] = buf[c] + PixIO<T>::load(px + c) * alpha;
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) sum[c] = j == 0 ? beta * buf[c] : sum[c] + beta * buf[c];
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) PixIO<T>::store(dst + p * 3 + c, sum[c]);
    }
}

__global__ void __launch_bounds__(kRsThreads)
k_resize_area_int_f32(const float *__restrict__ src, int cin, int sw, float *__restrict__ dst, int dw, int dh, int ix,
                      int iy, float scale) {
    const size_t total = (size_t)dw * dh;
    const int area = ix * iy;
    for (size_t p = (size_t)blockIdx.x * kRsThreads + threadIdx.x; p < total; p += (size_t)gridDim.x * kRsThreads) {
        const int dy = (int)(p / dw), dx = (int)(p - (size_t)dy * dw);
        const float *base = src + ((size_t)dy * iy * sw + (size_t)dx * ix) * cin;
        float sum[3] = {0.f, 0.f, 0.f};
        int k = 0;
        auto at = [&](int kk, int c) { return __ldg(base + ((size_t)(kk / ix) * sw + (kk % ix)) * cin + c); };
        for (; k + 4 <= area; k += 4)
#pragma unroll
            for (int c = 0; c < 3; ++c) sum[c] = sum[c] + (((at(k, c) + at(k + 1, c)) + at(k + 2, c)) + at(k + 3, c));
        for (; k < area; ++k)
#pragma unroll
            for (int c = 0; c < 3; ++c) sum[c] = sum[c] + at(k, c);
#pragma unroll
        for (int c = 0; c < 3; ++c) dst[p * 3 + c] = sum[c] * scale;
    }
}

__global__ void __launch_bounds__(kRsThreads)
k_resize_area_int_u8(const uint8_t *__restrict__ src, int cin, int sw, uint8_t *__restrict__ dst, int dw, int dh,
                     int ix, int iy, float scale) {
    const size_t total = (size_t)dw * dh;
    for (size_t p = (size_t)blockIdx.x * kRsThreads + threadIdx.x; p < total; p += (size_t)gridDim.x * kRsThreads) {
        const int dy = (int)(p / dw), dx = (int)(p - (size_t)dy * dw);
        const uint8_t *base = src + ((size_t)dy * iy * sw + (size_t)dx * ix) * cin;
        int sum[3] = {0, 0, 0};
        for (int j = 0; j < iy; ++j)
            for (int i = 0; i < ix; ++i)
#pragma unroll
                for (int c = 0; c < 3; ++c) sum[c] += __ldg(base + ((size_t)j * sw + i) * cin + c);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            int v;
            if (ix == 2 && iy == 2) v = (sum[c] + 2) >> 2;
            else v = min(max(__float2int_rn((float)sum[c] * scale), 0), 255);
            dst[p * 3 + c] = (uint8_t)v;
        }
    }
}

// One thread per column and strip of VS vertically adjacent destination pixels: when enlarging, neighbouring
// destination rows share most of their eight source rows, so the horizontal 8-tap sums of a source row are formed
// once and fed to every destination row of the strip that uses it (2.5x fewer loads than one pixel per thread at
// 1.5x enlargement).  Each destination pixel still accumulates its rows in tap order j = 0..7, so the float32
// result is the same as the pixel-per-thread form; the uint8 form is integer arithmetic.
constexpr int kLzStrip = 4;

template <typename T>
struct LzAcc;
template <>
struct LzAcc<uint8_t> {
    typedef int acc_t;
    static __device__ __forceinline__ int coef(const LanczosTabDev &t, int i) { return t.icoef[i]; }
    static __device__ __forceinline__ int load(const uint8_t *p) { return (int)__ldg(p); }
    static __device__ __forceinline__ uint8_t store(int v) { return (uint8_t)min(max((v + (1 << 21)) >> 22, 0), 255); }
};
template <>
struct LzAcc<float> {
    typedef float acc_t;
    static __device__ __forceinline__ float coef(const LanczosTabDev &t, int i) { return t.coef[i]; }
    static __device__ __forceinline__ float load(const float *p) { return __ldg(p); }
    static __device__ __forceinline__ float store(float v) { return v; }
};

template <typename T>
__global__ void __launch_bounds__(kRsThreads)
k_resize_lanczos4(const T *__restrict__ src, int cin, int sw, int sh, T *__restrict__ dst, int dw, int dh,
                  LanczosTabDev xt, LanczosTabDev yt) {
    typedef typename LzAcc<T>::acc_t A;
    const int nstrips = (dh + kLzStrip - 1) / kLzStrip;
    const size_t total = (size_t)dw * nstrips;
    for (size_t p = (size_t)blockIdx.x * kRsThreads + threadIdx.x; p < total; p += (size_t)gridDim.x * kRsThreads) {
        const int strip = (int)(p / dw), dx = (int)(p - (size_t)strip * dw);
        const int dy0 = strip * kLzStrip, ny = min(kLzStrip, dh - dy0);
        const int sx = xt.ofs[dx];
        int xi[8];
        A ca[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            xi[k] = min(max(sx + k - 3, 0), sw - 1) * cin;
            ca[k] = LzAcc<T>::coef(xt, dx * 8 + k);
        }
        int sy[kLzStrip];
        A acc[kLzStrip][3];
#pragma unroll
        for (int o = 0; o < kLzStrip; ++o) {
            sy[o] = yt.ofs[min(dy0 + o, dh - 1)] - 3;  // first (unclamped) source row of destination row dy0 + o
            acc[o][0] = acc[o][1] = acc[o][2] = (A)0;
        }
        const int r_first = sy[0], r_last = sy[ny - 1] + 7;  // yt.ofs is non-decreasing
        for (int rr = r_first; rr <= r_last; ++rr) {
            const T *row = src + (size_t)min(max(rr, 0), sh - 1) * sw * cin;
            A h[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) h[c] = LzAcc<T>::load(row + xi[0] + c) * ca[0];
#pragma unroll
            for (int k = 1; k < 8; ++k)
#pragma unroll
                for (int c = 0; c < 3; ++c) h[c] = h[c] + LzAcc<T>::load(row + xi[k] + c) * ca[k];
#pragma unroll
            for (int o = 0; o < kLzStrip; ++o) {
                const int j = rr - sy[o];
                if (o < ny && j >= 0 && j < 8) {
                    const A cb = LzAcc<T>::coef(yt, (dy0 + o) * 8 + j);
#pragma unroll
                    for (int c = 0; c < 3; ++c) acc[o][c] = j == 0 ? h[c] * cb : acc[o][c] + h[c] * cb;
                }
            }
        }
#pragma unroll
        for (int o = 0; o < kLzStrip; ++o)
            if (o < ny)
#pragma unroll
                for (int c = 0; c < 3; ++c) dst[((size_t)(dy0 + o) * dw + dx) * 3 + c] = LzAcc<T>::store(acc[o][c]);
    }
}

int grid_for(size_t items, int num_sms) {
    size_t want = (items + kRsThreads - 1) / kRsThreads, cap = (size_t)num_sms * 8;
    if (want < 1) want = 1;
    return (int)(want < cap ? want : cap);
}

}  // namespace

// ---- host-side tables (binary64, OpenCV's expressions) ------------------------------------------------------
bool resize_area_is_fast(int ssize, int dsize, int &iscale) {
    const double scale = 1.0 / ((double)dsize / ssize);
    iscale = (int)std::lrint(scale);  // saturate_cast<int>(double) rounds
    return std::fabs(scale - iscale) < 2.220446049250313e-16;
}

void resize_area_tab(int ssize, int dsize, AreaTabHost &t) {
    const double scale = 1.0 / ((double)dsize / ssize);
    t.start.assign(1, 0);
    t.si.clear();
    t.alpha.clear();
    for (int dx = 0; dx < dsize; ++dx) {
        const double f1 = dx * scale, f2 = f1 + scale;
        const double cell = std::fmin(scale, ssize - f1);
        int s1 = (int)std::ceil(f1), s2 = (int)std::floor(f2);
        s2 = s2 < ssize - 1 ? s2 : ssize - 1;
        s1 = s1 < s2 ? s1 : s2;
        if (s1 - f1 > 1e-3) {
            t.si.push_back(s1 - 1);
            t.alpha.push_back((float)((s1 - f1) / cell));
        }
        for (int sx = s1; sx < s2; ++sx) {
            t.si.push_back(sx);
            t.alpha.push_back((float)(1.0 / cell));
        }
        if (f2 - s2 > 1e-3) {
            t.si.push_back(s2);
            t.alpha.push_back((float)(std::fmin(std::fmin(f2 - s2, 1.0), cell) / cell));
        }
        t.start.push_back((int)t.si.size());
    }
}

static void lanczos4_coeffs(float x, float *co) {
    static const double s45 = 0.70710678118654752440084436210485;
    static const double cs[8][2] = {{1, 0}, {-s45, -s45}, {0, 1}, {s45, -s45}, {-1, 0}, {s45, s45}, {0, -1}, {-s45, s45}};
    const double y0 = -((double)x + 3) * 3.1415926535897932384626433832795 * 0.25, s0 = std::sin(y0), c0 = std::cos(y0);
    float sum = 0.f;
    for (int i = 0; i < 8; ++i) {
        const float d = (x + 3.0f) - (float)i;
        if (std::fabs(d) >= 1e-6f) {
            const double y = -(double)d * 3.1415926535897932384626433832795 * 0.25;
            co[i] = (float)((cs[i][0] * s0 + cs[i][1] * c0) / (y * y));
        } else {
            co[i] = 1e30f;
        }
        sum += co[i];
    }
    sum = 1.f / sum;
    for (int i = 0; i < 8; ++i) co[i] *= sum;
}

void resize_lanczos4_tab(int ssize, int dsize, LanczosTabHost &t) {
    const double scale = 1.0 / ((double)dsize / ssize);
    t.ofs.resize(dsize);
    t.coef.resize((size_t)dsize * 8);
    t.icoef.resize((size_t)dsize * 8);
    for (int d = 0; d < dsize; ++d) {
        float f = (float)((d + 0.5) * scale - 0.5);
        const int s = (int)std::floor(f);
        f -= (float)s;
        t.ofs[d] = s;
        lanczos4_coeffs(f, &t.coef[(size_t)d * 8]);
        for (int k = 0; k < 8; ++k) {
            long r = std::lrint((double)(t.coef[(size_t)d * 8 + k] * 2048.0f));  // saturate_cast<short>: round half even
            t.icoef[(size_t)d * 8 + k] = (int)(r < -32768 ? -32768 : (r > 32767 ? 32767 : r));
        }
    }
}

// ---- launches ----------------------------------------------------------------------------------------------
cudaError_t launch_resize_area(const void *src, bool u8, int cin, int sh, int sw, void *dst, int dh, int dw,
                               const AreaTabDev &xt, const AreaTabDev &yt, int num_sms, cudaStream_t st) {
    (void)sh;
    const int grid = grid_for((size_t)dw * dh, num_sms);
    if (u8)
        k_resize_area<uint8_t><<<grid, kRsThreads, 0, st>>>(static_cast<const uint8_t *>(src), cin, sw,
                                                             static_cast<uint8_t *>(dst), dw, dh, xt, yt);
    else
        k_resize_area<float><<<grid, kRsThreads, 0, st>>>(static_cast<const float *>(src), cin, sw,
                                                           static_cast<float *>(dst), dw, dh, xt, yt);
    return cudaGetLastError();
}

cudaError_t launch_resize_area_int(const void *src, bool u8, int cin, int sh, int sw, void *dst, int dh, int dw, int ix,
                                   int iy, int num_sms, cudaStream_t st) {
    (void)sh;
    const int grid = grid_for((size_t)dw * dh, num_sms);
    const float scale = 1.f / (float)(ix * iy);
    if (u8)
        k_resize_area_int_u8<<<grid, kRsThreads, 0, st>>>(static_cast<const uint8_t *>(src), cin, sw,
                                                          static_cast<uint8_t *>(dst), dw, dh, ix, iy, scale);
    else
        k_resize_area_int_f32<<<grid, kRsThreads, 0, st>>>(static_cast<const float *>(src), cin, sw,
                                                           static_cast<float *>(dst), dw, dh, ix, iy, scale);
    return cudaGetLastError();
}

cudaError_t launch_resize_lanczos4(const void *src, bool u8, int cin, int sh, int sw, void *dst, int dh, int dw,
                                   const LanczosTabDev &xt, const LanczosTabDev &yt, int num_sms, cudaStream_t st) {
    const int grid = grid_for((size_t)dw * ((dh + kLzStrip - 1) / kLzStrip), num_sms);
    if (u8)
        k_resize_lanczos4<uint8_t><<<grid, kRsThreads, 0, st>>>(static_cast<const uint8_t *>(src), cin, sw, sh,
                                                                 static_cast<uint8_t *>(dst), dw, dh, xt, yt);
    else
        k_resize_lanczos4<float><<<grid, kRsThreads, 0, st>>>(static_cast<const float *>(src), cin, sw, sh,
                                                               static_cast<float *>(dst), dw, dh, xt, yt);
    return cudaGetLastError();
}

}  // namespace r2f
