/*
 * r2f_b200.h -- C ABI of the B200-native raw2film render path (libr2f_b200.so).
 *
 * This is the drop-in boundary for the reference's per-pixel film-emulation path
 * (JanLohse/raw2film, src/raw2film/cpu_processor.py:363-407 == the body of
 * CpuProcessor.process after the LUTs are loaded, and the equivalent WebGPU
 * pipeline src/raw2film/gpu_processor.py:1695-1890).  Plain pointers and sizes only;
 * no torch / numpy types.  All image pointers passed to the *_dev entry points are
 * CUDA device pointers on the context's device; `stream` is a cudaStream_t (NULL =
 * legacy default stream).  Table setters take HOST pointers; the data is on the device when they
 * return.  Setters are copy-on-write: they never touch a buffer that a render already issued (on any
 * stream) may still be reading, so tables may be changed between the frames of a pipelined batch
 * without synchronising (reference gui_objects.py:65-115 renders mixed stocks back to back).
 *
 * Every function returns R2F_OK (0) or an error code; r2f_last_error() returns a
 * thread-local message (the reference raises plain Python exceptions; the Python shim
 * raw2film_b200/_cabi.py turns non-zero codes into RuntimeError).
 *
 * There is NO CPU fallback behind this interface.
 */
#ifndef R2F_B200_H
#define R2F_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define R2F_ABI_VERSION 2

/* status codes */
#define R2F_OK 0
#define R2F_ERR_INVALID 1 /* bad argument / table not set */
#define R2F_ERR_CUDA 2    /* CUDA runtime error (message has the cudaError string) */
#define R2F_ERR_NOMEM 3   /* workspace too small */

/* stage gates: mirror the `if` chain of cpu_processor.py:368-403 */
#define R2F_HALATION 0x01u /* cpu_processor.py:368  `if halation`                         */
#define R2F_MTF 0x02u      /* cpu_processor.py:382  `if sharpness and stock.mtf`           */
#define R2F_GRAIN 0x04u    /* cpu_processor.py:387  `if grain and stock.rms_density`       */
#define R2F_GRAIN_BW 0x08u /* cpu_processor.py:394  bw_grain = (grain == 1)               */
#define R2F_BURN 0x10u     /* cpu_processor.py:399  `if highlight_burn and (...)`          */

/* float working-space taps for parity tests (r2f_render_tap) */
#define R2F_TAP_EXPOSURE 1 /* after apply_2d_lut          cpu_processor.py:364 */
#define R2F_TAP_HALATION 2 /* after effects.halation      cpu_processor.py:369 */
#define R2F_TAP_DENSITY 3  /* after multi_channel_interp  cpu_processor.py:380 */
#define R2F_TAP_MTF 4      /* after effects.film_sharpness cpu_processor.py:383 */
#define R2F_TAP_GRAIN 5    /* after apply_grain + clip    cpu_processor.py:397 */
#define R2F_TAP_BURN 6     /* after effects.burn          cpu_processor.py:403 */
#define R2F_TAP_RGB 7      /* after apply_lut_tetrahedral cpu_processor.py:405 */

typedef struct r2f_ctx r2f_ctx;

int r2f_abi_version(void);
const char *r2f_last_error(void);

/* Replaces CpuProcessor.__init__ / GpuProcessor.__init__ device acquisition
 * (cpu_processor.py:27-29, gpu_processor.py:176-190): one context per GPU. */
int r2f_create(int device, r2f_ctx **out);
int r2f_destroy(r2f_ctx *ctx);

/* ---- table slots ----------------------------------------------------------------------------
 * A context holds R2F_MAX_SLOTS independent table sets (2-D LUT, H-D curve, 3-D LUT, halation / MTF /
 * grain kernels and curves, burn parameters, grain seed).  Setters write to, and renders read from, the
 * selected slot (slot 0 after r2f_create).  The reference processors keep ONE set and rebuild it whenever a
 * settings sub-dict changes (cpu_processor.py:157, 179, 229); a batch over mixed stocks
 * (gui.py:2472-2514) would re-upload everything on every frame -- with one slot per stock it switches by
 * index.  r2f_clear_slot drops a slot's tables (renders in flight keep theirs until they finish). */
#define R2F_MAX_SLOTS 16
int r2f_select_slot(r2f_ctx *ctx, int slot);
int r2f_clear_slot(r2f_ctx *ctx, int slot);

/* ---- table setters: the device-side half of the reference's load_* caches ------------- */

/* 2-D chromaticity input LUT, (n, n, 3) float32, index order lut[x_idx][y_idx]
 * (load_input_lut cpu_processor.py:142-164; _ensure_lut_2d gpu_processor.py:349-376). */
int r2f_set_lut2d(r2f_ctx *ctx, const float *lut, int n);

/* H-D curve, (4, N) float32: row 0 = log10-exposure abscissa, rows 1..3 = R,G,B density
 * (load_density_curve cpu_processor.py:166-188; _ensure_lut_1d gpu_processor.py:307-347).
 * A uniform abscissa (every sample within 1e-3 step of the straight line between its ends, e.g. a
 * float32 linspace) takes the normalised lookup of shaders/lut_1d.wgsl:43-47; any other strictly
 * increasing abscissa is evaluated with np.interp semantics (binary search, binary64 slope).
 * log_eps is the lower clip of log_clip (cpu_processor.py:378; shaders/lut_1d.wgsl:24). */
int r2f_set_curve1d(r2f_ctx *ctx, const float *curve, int N, float log_eps);

/* Output 3-D LUT, (n, n, n, 3) float32 indexed lut[r][g][b], sampled at density * scale
 * (load_output_lut cpu_processor.py:190-267; apply_lut_tetrahedral utils.py:247-380;
 *  call site cpu_processor.py:405 passes scale = 0.25). */
int r2f_set_lut3d(r2f_ctx *ctx, const float *lut, int n, double scale);

/* Halation kernel, (k, k, 3) float32, k odd, as built by compute_halation_kernel
 * (effects.py:239-263).  A channel whose kernel is an exact centre delta is passed through. */
int r2f_set_halation_kernel(r2f_ctx *ctx, const float *kernel, int k);

/* MTF kernel, (k, k, 3) float32, k odd, as built by mtf_kernel (effects.py:165-185). */
int r2f_set_mtf_kernel(r2f_ctx *ctx, const float *kernel, int k);

/* Grain: amplitude curve (4, N) over density (get_grain_curve, gpu_processor.py:913;
 * shaders/grain.wgsl:77-86), smoothing kernel (k, k) float32 or NULL for a 1x1 kernel
 * (grain_kernel, gpu_processor.py:927-934), and the seed of the on-device Philox stream
 * (the reference draws a fresh seed per frame, gpu_processor.py:586-592). */
int r2f_set_grain(r2f_ctx *ctx, const float *curve, int N, const float *kernel, int k, uint64_t seed);

/* Tuning / test options.  R2F_OPT_CONV_PATH selects the halation correlation path:
 * 0 = auto (FFT for wide even-symmetric kernels, direct otherwise), 1 = force direct,
 * 2 = force FFT (render fails if the kernel or frame is not eligible). */
#define R2F_OPT_CONV_PATH 1
#define R2F_OPT_CONV_SYM 2
/* R2F_OPT_FUSE_MTF: 1 (default) = in a banded call (r2f_render_banded) the MTF correlation is issued band by band
 * together with the grain + finish kernel, so the first rows of the result are final one band of MTF + grain after
 * the density planes are complete; 0 = one whole-frame MTF launch before the banded grain kernel (A/B; same bytes).
 * (A single kernel doing both was measured not to pay: both are compute-bound and want different thread shapes.)
 * R2F_OPT_FAST_CHAIN: 1 (default) = per-pixel chains evaluate a guarded float32 fast path and defer the pixels
 * whose uint8 result it cannot prove to the exact path; 0 = exact path for every pixel. */
#define R2F_OPT_FUSE_MTF 3
#define R2F_OPT_FAST_CHAIN 4
int r2f_set_option(r2f_ctx *ctx, int key, int value);

/* Re-seed the on-device noise stream only (no table upload, no synchronisation). */
int r2f_set_grain_seed(r2f_ctx *ctx, uint64_t seed);

/* Highlight burn parameters (effects.burn effects.py:392-418): d_ref of the green layer,
 * strength, and burn_scale (down-sampling divisor, effects.py:365). */
int r2f_set_burn(r2f_ctx *ctx, float d_ref, float highlight_burn, float burn_scale);

/* ---- the render call -------------------------------------------------------------------- */

/* Bytes of device scratch r2f_render needs for an H x W frame with these stage flags. */
size_t r2f_workspace_bytes(int H, int W, unsigned flags);

/* Replaces the body of CpuProcessor.process after the loaders (cpu_processor.py:363-407):
 *   in_dev   float32 H x W x in_channels (3 = XYZ, 4 = XYZ + alpha as in the GPU payload,
 *            gpu_processor.py:765), C-contiguous
 *   out_dev  uint8 H x W x 3
 *   noise_dev  optional injected white N(0,1) field, float32 H x W x noise_channels
 *            (3, or 1 with R2F_GRAIN_BW); NULL = generate on device from the seed
 *   workspace_dev / workspace_bytes  caller-owned scratch (>= r2f_workspace_bytes) */
int r2f_render(r2f_ctx *ctx, const float *in_dev, int H, int W, int in_channels, uint8_t *out_dev,
               unsigned flags, const float *noise_dev, int noise_channels, void *workspace_dev,
               size_t workspace_bytes, void *stream);

/* Input-format variant ("u16 ingest on device").  in_format R2F_IN_U16: `in_dev` is the
 * uint16 H x W x in_channels frame rawpy hands over (reference raw_conversion.py:38-48); the device
 * applies the reference's own ingest arithmetic `astype(float32) / 65535.0` then `*= gain`
 * (raw_conversion.py:51-53, gain = 2**calc_exposure) before the 2-D LUT, bit-identically.
 * R2F_IN_F32 ignores in_gain and equals r2f_render. */
#define R2F_IN_F32 0
#define R2F_IN_U16 1
int r2f_render_ex(r2f_ctx *ctx, const void *in_dev, int in_format, float in_gain, int H, int W, int in_channels,
                  uint8_t *out_dev, unsigned flags, const float *noise_dev, int noise_channels,
                  void *workspace_dev, size_t workspace_bytes, void *stream);

/* Streaming variant for callers that copy the frame in and the result out around the call (what
 * GpuProcessor.process_preloaded does, gpu_processor.py:1643-1693: upload, pipeline, read back).  The frame arrives
 * in `nbands` horizontal bands, band i = rows [r2f_band_row(H, nbands, i), r2f_band_row(H, nbands, i + 1)):
 * in_ready[i] is a cudaEvent_t the CALLER records on its copy stream once band i is on the device; out_done[i] is a
 * cudaEvent_t the LIBRARY records on `stream` once the output rows of band i are final.  The first kernel of the
 * pipeline runs per band as its rows arrive and the last one per band, so host <-> device copies overlap the
 * render inside one call.  Results are identical to r2f_render_ex. */
int r2f_band_row(int H, int nbands, int i);
int r2f_render_banded(r2f_ctx *ctx, const void *in_dev, int in_format, float in_gain, int H, int W, int in_channels,
                      uint8_t *out_dev, unsigned flags, const float *noise_dev, int noise_channels,
                      void *workspace_dev, size_t workspace_bytes, int nbands, void *const *in_ready,
                      void *const *out_done, void *stream);

/* Same pipeline, stopped after `tap_stage`; writes the float32 H x W x 3 working image. */
int r2f_render_tap(r2f_ctx *ctx, const float *in_dev, int H, int W, int in_channels, unsigned flags,
                   const float *noise_dev, int noise_channels, void *workspace_dev, size_t workspace_bytes,
                   int tap_stage, float *tap_dev, void *stream);

int r2f_render_tap_ex(r2f_ctx *ctx, const void *in_dev, int in_format, float in_gain, int H, int W,
                      int in_channels, unsigned flags, const float *noise_dev, int noise_channels,
                      void *workspace_dev, size_t workspace_bytes, int tap_stage, float *tap_dev, void *stream);

/* Host-buffer variant for non-torch callers (what a cgo/JNI/ctypes binding of
 * GpuProcessor.process_preloaded, gpu_processor.py:1643-1693, would call): copies in,
 * renders, copies out and synchronises; staging and scratch are owned by the context. */
int r2f_render_host(r2f_ctx *ctx, const float *in_host, int H, int W, int in_channels, uint8_t *out_host,
                    unsigned flags, const float *noise_host, int noise_channels);

/* ---- stage entry points (same kernels, exposed for stage-level parity tests) ------------- */

/* convolve_2d (effects.py:146-156): per-channel correlation, centre anchor, BORDER_REFLECT_101.
 * in/out: float32 H x W x 3 interleaved device images; kernel: HOST (k, k, 3). */
int r2f_convolve2d(r2f_ctx *ctx, const float *in_dev, float *out_dev, int H, int W, const float *kernel, int k,
                   void *workspace_dev, size_t workspace_bytes, void *stream);

/* White N(0,1) field from the context's Philox stream: float32 H x W x channels (1 or 3). */
int r2f_generate_noise(r2f_ctx *ctx, float *out_dev, int H, int W, int channels, uint64_t seed, void *stream);

/* chroma_nr_filter (effects.py:547-561): the pre-path chroma noise reduction.  in_dev: float32
 * H x W x in_channels XYZ; out_dev: float32 H x W x 3; taps: HOST Gaussian taps (odd count) as built by
 * gaussian_kernel_1d (effects.py:421-435); scratch >= 6 planes (r2f_workspace_bytes suffices). */
int r2f_chroma_nr(r2f_ctx *ctx, const float *in_dev, int in_channels, float *out_dev, int H, int W,
                  const float *taps, int ntaps, void *workspace_dev, size_t workspace_bytes, void *stream);

/* Counting pass of generate_histogram (utils.py:158-169; shaders/histogram.wgsl pass 1):
 * counts_dev[c * 256 + v] = number of pixels whose channel c equals v, over a uint8 H x W x 3 device image.
 * The 256-bin post-processing and rasterisation (utils.py:171-223) are host-side (hostops.histogram_image). */
int r2f_histogram(r2f_ctx *ctx, const uint8_t *img_dev, int H, int W, uint32_t *counts_dev, void *stream);

/* All three passes of generate_histogram (utils.py:145-223; shaders/histogram.wgsl:35-158) on the device: counts,
 * float32 log1p / smoothing / scaling, and the rasterised (height, 256, 4) uint8 widget image written to out_dev.
 * mix_table: HOST, the (2, 2, 2, 4) uint8 colour-mix table (32 bytes, utils.py:95-142). */
int r2f_histogram_image(r2f_ctx *ctx, const uint8_t *img_dev, int H, int W, const uint8_t *mix_table, int height,
                        uint8_t *out_dev, void *stream);

/* Reduction of calc_exposure (color_processing.py:71-99; called on every decoded frame, raw_conversion.py:51-53;
 * SURVEY 8f-1): *mean_out (host) = mean over rows 0,2,4,.. and columns 0,2,4,.. of green ** (1 / factor), green
 * taken from a device frame in either input format (uint16 is divided by 65535 first, raw_conversion.py:51).
 * Each term is rounded to binary32 like the reference's float32 array; the sum is binary64 in a fixed order.
 * The caller finishes with exp_comp = log2(ref_exposure / mean ** factor).  Synchronises the stream. */
int r2f_calc_exposure(r2f_ctx *ctx, const void *in_dev, int in_format, int H, int W, int in_channels, double factor,
                      double *mean_out, void *stream);

/* add_canvas (effects.py:338-357): fill a canvas_h x canvas_w x 3 uint8 image with (r, g, b) and paste
 * the H x W x 3 render at (off_y, off_x); geometry from get_canvas_data (effects.py:290-335). */
int r2f_canvas_paste(r2f_ctx *ctx, const uint8_t *src_dev, int H, int W, uint8_t *dst_dev, int canvas_h, int canvas_w,
                     int off_y, int off_x, int r, int g, int b, void *stream);

/* resolution_scaling (utils.py:226-244): cv2.resize on the device.  INTER_AREA is what the reference
 * uses when shrinking (the float32 frame before the path, cpu_processor.py:122-134 / gpu_processor.py:748-758),
 * INTER_LANCZOS4 when enlarging (the uint8 image after it, cpu_processor.py:411-412).  in_dev: H x W x in_channels
 * float32 (3 or 4 channels) or uint8 (3 channels); out_dev: out_h x out_w x 3 of the same type.  OpenCV's
 * arithmetic in its operation order: INTER_AREA (both types) and uint8 INTER_LANCZOS4 are bit-identical to cv2;
 * float32 INTER_LANCZOS4 agrees to a few ulp (cv2's vertical pass is host-SIMD dependent). */
#define R2F_PIX_F32 0
#define R2F_PIX_U8 2
#define R2F_INTER_AREA 0
#define R2F_INTER_LANCZOS4 1
int r2f_resize(r2f_ctx *ctx, const void *in_dev, int pix_format, int H, int W, int in_channels, void *out_dev,
               int out_h, int out_w, int interpolation, void *stream);

/* Presentation blit (shaders/copy_to_int.wgsl:18-51; geometry _bind_copy_to_dst gpu_processor.py:1416-1539): scale
 * the rendered uint8 H x W x 3 image into a dst_h x dst_w x 4 RGBA8 buffer (the preview widget), letterboxed, with
 * the canvas rectangle filled in (r, g, b) and everything else transparent.  transform: HOST, the eight floats of
 * the shader's uniform block {scale_x, scale_y, offset_x, offset_y, canvas_min_x, canvas_min_y, canvas_max_x,
 * canvas_max_y}.  Bilinear sampling with texel centres at +0.5 and clamp-to-edge, as a WebGPU linear sampler. */
int r2f_present(r2f_ctx *ctx, const uint8_t *src_dev, int H, int W, uint8_t *dst_rgba_dev, int dst_h, int dst_w,
                const float *transform, int r, int g, int b, void *stream);

/* Number of kernel launches issued by this context since creation (bench.py gpu_launches). */
uint64_t r2f_launch_count(const r2f_ctx *ctx);

/* Renders captured into a CUDA graph (cudaStreamBeginCapture on the stream handed to r2f_render) are not visible to
 * the copy-on-write table storage when the graph is replayed: call this right after every cudaGraphLaunch on the
 * stream the graph was launched on, so that tables the replay reads are not recycled under it. */
int r2f_stream_mark(r2f_ctx *ctx, void *stream);

/* Guarded fast chain statistics (R2F_OPT_FAST_CHAIN): pixels that the float32 fast path could not decide and the
 * exact chain evaluated since the last call (the counter is reset), and the selected slot's proven bound on
 * |255 * (fast - exact)| (-1 when its tables do not qualify for the fast path).  Synchronises the device. */
int r2f_fast_chain_stats(r2f_ctx *ctx, uint64_t *deferred_pixels, float *margin);

/* Per-kernel device timing (bench.py roofline): when enabled, every launch of the render
 * pipeline is bracketed by CUDA events on the launching stream.  r2f_profile_read waits for the
 * recorded events, ACCUMULATES milliseconds and launch counts per kernel id into the caller's
 * arrays (length R2F_PROF_COUNT) and clears the recorded events. */
#define R2F_PROF_POINTWISE 0 /* k_pointwise (fused a2+a4+a5+a9+a10)        */
#define R2F_PROF_EXPOSE 1    /* k_expose (a2)                              */
#define R2F_PROF_HALATION 2  /* direct k_conv2d halation + density epilogue (a3-5) */
#define R2F_PROF_DENSITY 3   /* pointwise density pass when halation is off */
#define R2F_PROF_MTF 4       /* k_conv2d MTF (a6)                           */
#define R2F_PROF_NOISE 5     /* k_noise / noise upload shuffle (a7)        */
#define R2F_PROF_GRAIN 6     /* k_conv2d grain + apply epilogue (a7)       */
#define R2F_PROF_BURN 7      /* burn mask kernels (a8)                     */
#define R2F_PROF_FINISH 8    /* k_finish (a8 apply + a9 + a10)             */
#define R2F_PROF_FFT_ROWS_FWD 9  /* k_fft_rows_fwd (a2 + row FFT)          */
#define R2F_PROF_FFT_COLS 10     /* k_fft_cols (column FFT x Khat, inverse) */
#define R2F_PROF_FFT_ROWS_INV 11 /* k_fft_rows_inv (row IFFT + a4 + a5)    */
#define R2F_PROF_COUNT 12
int r2f_profile_enable(r2f_ctx *ctx, int on);
int r2f_profile_read(r2f_ctx *ctx, double *ms_accum, uint64_t *count_accum);

#ifdef __cplusplus
}
#endif
#endif /* R2F_B200_H */
