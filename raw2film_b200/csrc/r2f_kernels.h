// Launch wrappers of the sm_100a kernels (implemented in r2f_kernels.cu).
// Host-side plain structs only; used by the C ABI in r2f_api.cu.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

#include "device_math.cuh"
#include "fast_chain.cuh"

namespace r2f {

// Planar float32 working image: plane c lives at base + c * plane_stride, row pitch = W -- except the density planes
// between k_fft_rows_inv, k_conv2d_sym and k_grain_finish_sym of a frame whose width is not a multiple of 4, whose rows
// are padded to a multiple of 4 floats (ConvArgs::pitch, GrainFinishArgs::dens_pitch, FftConvArgs::dst_pitch) so that
// the 16-byte paths and the TMA tensor map survive odd widths.
struct Planes {
    float *base;
    size_t plane_stride;  // floats, multiple of 64
};

inline int padded_pitch(int W) { return (W + 3) & ~3; }
inline size_t plane_stride_for(int H, int W) {
    size_t n = (size_t)H * (size_t)padded_pitch(W);
    return (n + 63) / 64 * 64;
}

enum ConvEpilogue { EPI_NONE = 0, EPI_DENSITY = 1, EPI_GRAIN = 2, EPI_DENSITY_FAST = 3 };

struct ConvArgs {
    const float *in;      // planar source
    float *out;           // planar destination
    const float *aux;     // planar density (EPI_GRAIN only)
    size_t plane_stride;
    int H, W;
    const float *kern[3];  // device, transposed + padded: kern[c][j * kp + i] = K[i][j][c]
    int k, kp;
    // y-symmetric layout for k_conv2d_sym, or nullptr: ksym[c][(dy * wrow + j) * 2 + {0,1}] = K[r+dy][j][c]
    // (duplicated pair), row dy = 0 halved, wrow = conv_sym_wrow(k)
    const float *ksym[3];
    const float *ksym_host[3];  // the same layout in host memory (kernel-parameter copy for the uniform-weight form)
    int mode[3];      // 0 = identity (exact centre delta), 1 = correlate
    int in_plane[3];  // source plane feeding output channel c
    int epi;
    Curve1D curve;  // H-D curve (EPI_DENSITY) or grain amplitude curve (EPI_GRAIN)
    float eps;
    int tile_y0, tile_rows;  // k_conv2d_sym only: rows of 64-row tiles to compute (tile_rows == 0: the whole frame)
    int pitch;               // k_conv2d_sym only: row pitch of the source and destination planes (0: W)
};

struct BurnArgs {
    const float *map;  // low-res blurred mask (lh x lw) or nullptr
    int lh, lw;        // low-res size
    int zh, zw;        // size of the zoomed map before pad/crop (scipy.ndimage.zoom output)
    float strength;
};

// K1: XYZ -> 2D LUT -> log10 -> H-D curve -> tetrahedral LUT -> u8, one pass (pointwise configs)
// `fmt` = kFmt* (device_math.cuh); `gain` is the exposure gain applied to uint16 input only.
cudaError_t launch_pointwise(const void *in, int fmt, float gain, uint8_t *out, size_t npix, const Lut2D &l2,
                             const Curve1D &cv, float eps, const Lut3D &l3, int num_sms, cudaStream_t st);
// K1 through the guarded float32 fast path (fast_chain.cuh); `stats` (device, optional) accumulates the number of
// pixels that were deferred to the exact chain.  Fails with cudaErrorInvalidValue when the tables do not qualify.
size_t pointwise_fast_smem(const Lut2D &l2, const FastChain &F);
cudaError_t launch_pointwise_fast(const void *in, int fmt, float gain, uint8_t *out, size_t npix, const Lut2D &l2,
                                  const Curve1D &cv, float eps, const Lut3D &l3, const FastChain &F,
                                  unsigned long long *stats, int num_sms, cudaStream_t st);
// XYZ (interleaved, 3 or 4 channels) -> planar exposure
cudaError_t launch_expose(const void *in, int fmt, float gain, Planes out, size_t npix, const Lut2D &l2, int num_sms,
                          cudaStream_t st);
// direct 2-D correlation, reflect-101 borders, fused epilogue
cudaError_t launch_conv2d(const ConvArgs &a, cudaStream_t st);
// same contract for kernels that are mirror-symmetric in y (r2f_conv_sym.cu): packed-FMA path
constexpr int kConvSymMaxK = 33;
int conv_sym_wrow(int k);
bool conv_sym_supported(int k);
cudaError_t launch_conv2d_sym(const ConvArgs &a, cudaStream_t st);
// planar density -> [burn] -> tetrahedral LUT -> u8 interleaved, or float32 interleaved (taps)

// `ft`: guarded float32 tetrahedral tail for the uint8 output (ft.ok == 0: the exact path for every pixel)
cudaError_t launch_finish(Planes in, size_t npix, int H, int W, const Lut3D &l3, const BurnArgs &burn, uint8_t *out_u8,
                          float *out_f32, int f32_stage_rgb, int num_sms, cudaStream_t st,
                          const FastTetra &ft = FastTetra{});
// chroma NR pre-stage (reference effects.py:421-561): needs 6 float planes of scratch
cudaError_t launch_chroma_nr(const float *in, int cin, float *out, int H, int W, const float *taps_host, int ntaps,
                             float *ws, size_t ps, int num_sms, cudaStream_t st);
// 3 x 256 histogram counts of a uint8 H x W x 3 image (reference utils.py:158-169)
cudaError_t launch_histogram(const uint8_t *img, size_t npix, unsigned int *counts_dev, int num_sms, cudaStream_t st);
// passes 2-3 of the histogram widget: counts -> (height, 256, 4) image through the 32-byte colour-mix table (host)
cudaError_t launch_histogram_image(const unsigned int *counts_dev, int height, const uint8_t *mix_host, uint8_t *out,
                                   cudaStream_t st);
// auto exposure: mean of green ** inv_factor over every second row/column (color_processing.py:71-99);
// `partial` holds nblocks doubles of scratch, `out` one double (device)
cudaError_t launch_exposure_mean(const void *in, int fmt, int H, int W, double inv_factor, double *partial, int nblocks,
                                 double *out, cudaStream_t st);
// canvas border: colour fill + paste (reference effects.py:338-357)
cudaError_t launch_canvas_paste(const uint8_t *src, int H, int W, uint8_t *dst, int CH, int CW, int off_y, int off_x,
                                int r, int g, int b, int num_sms, cudaStream_t st);
// presentation blit into a widget-sized RGBA8 buffer (reference shaders/copy_to_int.wgsl)
struct PresentArgs {
    float scale_x, scale_y, offset_x, offset_y;                       // destination pixel -> normalised source uv
    float canvas_min_x, canvas_min_y, canvas_max_x, canvas_max_y;      // canvas rectangle in destination pixels
    int r, g, b;                                                       // canvas colour
};
cudaError_t launch_present(const uint8_t *src, int H, int W, uint8_t *dst, int DH, int DW, const PresentArgs &u,
                           int num_sms, cudaStream_t st);
// layout shuffles
cudaError_t launch_planar_to_interleaved(Planes in, float *out, size_t npix, int num_sms, cudaStream_t st);
cudaError_t launch_interleaved_to_planar(const float *in, int cin, int nch, Planes out, size_t npix, int num_sms,
                                         cudaStream_t st);
// white N(0,1) noise, Philox4x32-10 + Box-Muller, planar
// `shift`: column shift of the field's quad grid, (grain kernel radius) & 3 inside a render (noise.cuh)
cudaError_t launch_noise(Planes out, int nch, int H, int W, uint64_t seed, int shift, int num_sms, cudaStream_t st);
// fused grain + burn apply + tetrahedral LUT + quantise (normal render path)
constexpr int kGrainSymMaxK = 21;
struct GrainFinishArgs {
    const float *dens;   // planar density (after MTF)
    const float *noise;  // planar injected white noise, or nullptr: regenerate from the seed per tile
    size_t plane_stride;
    int H, W;
    const float *gk;     // grain kernel, transposed + padded: gk[j * kp + i]
    const float *gk_sym; // y-symmetric packed layout (ConvArgs::ksym) or nullptr
    float2 gkw[(kGrainSymMaxK / 2 + 1) * ((kGrainSymMaxK + 1) / 2 * 2)];  // the same (w, w) pairs as a launch parameter
    int k, kp;
    int bw;              // one noise field for all three layers
    uint32_t seed_lo, seed_hi;
    Curve1D gcurve;
    FastCurve gfast;     // conversion-free form of gcurve (uniform abscissa), seg == nullptr: use gcurve
    Lut3D l3;
    FastTetra ft;        // guarded float32 tetrahedral LUT + quantise (ok == 0: exact path only)
    BurnArgs burn;
    uint8_t *out_u8;
    int dens_pitch;          // k_grain_finish_sym only: row pitch of `dens` (0: W); dens_out and noise keep W
    float *dens_out;         // k_grain_finish_sym only: not null = stop after the grain stage and write the grained
                             // density planes here (may alias `dens`; the burn mask needs the whole grained frame)
    int tile_y0, tile_rows;  // first tile row and tile-row count of this launch (0, 0 = all): banded output
    int noise_shift;         // (k / 2) & 3: column shift of the noise field's quad grid (noise.cuh)
};
cudaError_t launch_grain_finish(const GrainFinishArgs &a, cudaStream_t st);
// same contract for y-symmetric grain kernels (r2f_grain_sym.cu): row-pair sums + packed FMA; no burn apply (with
// burn, run it with dens_out and finish with launch_finish)
bool grain_finish_sym_supported(int k);
cudaError_t launch_grain_finish_sym(const GrainFinishArgs &a, cudaStream_t st);
// highlight-burn low-res mask: area down-sample of the green plane, max(x - d_ref, 0), 13-tap Gaussian (sigma 3)
cudaError_t launch_burn_mask(const float *green_plane, int H, int W, int lh, int lw, float d_ref, float *tmp,
                             float *map, cudaStream_t st);

}  // namespace r2f
