// FFT-based 2-D correlation for wide, even-symmetric kernels (halation).  See r2f_fft.cu.
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

#include <cuda_runtime.h>

#include "device_math.cuh"

namespace r2f {

constexpr int kFftColsPerBlock = 4;   // preferred columns per CTA in the column pass (32-byte row segments)
constexpr int kFftMaxPasses = 12;
constexpr int kFftMaxLen = 14336;

// One FFT length as the device sees it.
struct FftLine {
    int n;
    int nrad;
    int rad[kFftMaxPasses];
    int tw_off[kFftMaxPasses];  // start of each pass's twiddle table inside tw
    const float2 *tw;           // per-pass tables, pass p: tw[tw_off[p] + (t-1)*Ns + k] = exp(-2 pi i t k / (Ns R))
    // in-place decimation-in-frequency / -in-time passes (column kernel, compile-time plans): pass p works on
    // blocks of B = n / (R_1 .. R_{p-1}) with stride S = B / R_p; tables follow each other in pass order,
    // entry (t-1)*S + k = exp(-2 pi i t k / B)
    const float2 *tw_ip;
};

// Host-side description of one FFT length (factorisation + double-built root / cosine tables).
struct FftLineHost {
    int n = 0;
    std::vector<int> rad;
    std::vector<int> tw_off;
    std::vector<float2> roots;
    std::vector<float2> roots_ip;  // tables of the in-place passes (FftLine::tw_ip)
    std::vector<int> perm;         // perm[k] = position of X[k] after the in-place forward passes
    std::vector<double> cosines;
};

struct FftConvArgs {
    int H, W, r;        // frame size, kernel radius (k/2)
    FftLine row, col;   // lengths Wp (>= W + 2r, multiple of kFftColsPerBlock) and Hp (>= H + 2r)
    int nc;             // columns per CTA / per block of S (2..4, divides Wp)
    int col_groups;     // thread groups per column CTA (1 or 2)
    int row_off;        // row kernels: the padded row starts row_off elements into the line buffer (0..3), chosen so that
                        // r + row_off is a multiple of 4 and a pixel quad's four values are two aligned 16-byte accesses;
                        // a circular shift of the line shifts the filtered row by the same amount
    int dst_pitch;      // k_fft_rows_inv: row pitch of dst_planar (0: W)
    int rows_ahead;     // k_fft_rows_fwd prefetches the frame row of CTA blockIdx + rows_ahead into L2 (0: off)
    int cols_prefetch;  // column kernel: L2-prefetch its kernel-spectrum rows and the next CTA's block
    int col_inplace;    // 1: in-place column kernel (one 256-thread group per column, khat permuted by col perm)
    float2 *S;          // spectrum scratch, (Wp / nc) x H x nc complex
    const float *khat;  // [Wp][Hp] real kernel spectrum, 1/(Hp*Wp) folded in
    int chan[2];        // the two planes filtered together (R, G)
    float alpha[2], beta[2];  // out_c = alpha * (K (*) x_c) + beta * x_c
    // source: planar planes or interleaved XYZ through the 2-D input LUT
    const float *src_planar;
    const void *src_xyz;  // interleaved frame in one of the kFmt* formats
    float gain;           // exposure gain of uint16 frames
    Lut2D lut2d;
    // when set (and the source is an interleaved frame), k_fft_rows_fwd also stores the three exposure
    // planes it evaluates, and k_fft_rows_inv reads them back instead of re-evaluating the 2-D LUT
    float *exp_planar;
    size_t plane_stride;
    // destination: planar planes, optionally through log10 + H-D curve
    float *dst_planar;
    Curve1D curve;
    float eps;
    // forward row kernel only: first row and row count of this launch (0, 0 = the whole frame); lets a caller
    // start transforming the rows of a frame that is still arriving (r2f_render_banded)
    int row0, row_count;
};

int fft_good_size(int min_n, int multiple_of);
bool fft_make_line(int n, FftLineHost &out);
size_t fft_rows_smem(int Wp);
size_t fft_cols_smem(int Hp, int nc, int groups);
bool fft_col_geometry(int Hp, int Wp, int &nc, int &groups);

// Kernel spectrum of an even-symmetric k x k base kernel (device, row-major) -> khat [Wp][Hp].
// `perm_dev` (Hp ints or nullptr): khat[v][perm[u]] receives the value of row frequency u (in-place column kernel).
cudaError_t launch_khat(const float *base_kernel_dev, int k, int Hp, int Wp, const double *cosH_dev,
                        const double *cosW_dev, double *scratchA_dev, float *khat_dev, const int *perm_dev,
                        cudaStream_t st);
// True when the in-place column kernel has a compile-time plan for this line length (it always works on
// blocks of four columns).
bool fft_cols_inplace_available(int Hp);
size_t fft_cols_inplace_smem(int Hp);

// src_mode: 0 planar, 1 + kFmt* for an interleaved frame routed through the 2-D LUT.
// stage: 0 = all three kernels, 1 = rows forward, 2 = columns, 3 = rows inverse (for per-kernel timing)
cudaError_t launch_fft_conv(const FftConvArgs &a, int src_mode, bool density, cudaStream_t st, int stage = 0);

}  // namespace r2f
