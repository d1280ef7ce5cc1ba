/*
 * oracle/pointwise_oracle.c -- CPU restatement of the POINTWISE stages of the
 * raw2film render path.  TEST INFRASTRUCTURE ONLY: it is the checker for the
 * CUDA kernels (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline /
 * --impl reference legs).  The product (raw2film_b200/) never links or loads
 * this file.
 *
 * Build:  gcc -O2 -fPIC -shared -fopenmp -ffp-contract=off -fno-fast-math
 *         (see oracle/Makefile).  -ffp-contract=off matters: every float
 *         operation below is a separately rounded IEEE-754 binary32/binary64
 *         operation, in exactly the order written, so that the CUDA kernels
 *         (compiled -fmad=false) can be compared BIT-EXACTLY.
 *
 * Reference anchors (paths relative to /root/reference/src/raw2film):
 *   orc_apply_2d_lut      cpu_processor.py:364 -> spectral_film_lut.xy_lut.apply_2d_lut
 *                          (third-party, not in tree; restated from shaders/lut_2d.wgsl:18-108)
 *   orc_log_clip          cpu_processor.py:378 -> spectral_film_lut.utils.log_clip
 *                          (third-party; restated from shaders/lut_1d.wgsl:23-26,41)
 *   orc_curve_interp      cpu_processor.py:380 -> spectral_film_lut.utils.multi_channel_interp
 *                          (third-party; (4,N) layout from gpu_processor.py:307-333,
 *                           clamped ends from shaders/lut_1d.wgsl:43-47)
 *   orc_tetra             utils.py:247-380 apply_lut_tetrahedral (in tree; PINNED: bit-exact
 *                          against the numba original, tests/golden/tetra_*.npz)
 *   orc_quantise_u8       cpu_processor.py:407  (image * 255).astype(uint8)
 *
 * PARITY STATUS: orc_tetra / orc_quantise_u8 are pinned to the reference.  The
 * three third-party stages are "parity unpinned": spectral-film-lut (>=0.8.0,
 * pyproject.toml:28) is absent from the reference tree and cannot be installed
 * offline, so their arithmetic is the WGSL restatement the reference itself
 * ships, evaluated in float32 in the operation order documented per function.
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* torchrun exports OMP_NUM_THREADS=1; the CPU baseline asks for every host thread explicitly. */
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ---- a2: chromaticity-indexed 2D input LUT (lut_2d.wgsl:18-108) -------------------- */
static inline int clamp_floor_idx(float fl, int hi) {
    if (!(fl >= 0.0f)) return 0; /* also NaN */
    if (fl > (float)hi) return hi;
    return (int)fl;
}

static inline void lut2d_pixel(const float *px, const float *lut, int n, float *o) {
    const float X = px[0], Y = px[1], Z = px[2];
    const float S = (X + Y) + Z;                     /* lut_2d.wgsl:45 */
    if (S < 1e-12f) {                                /* lut_2d.wgsl:47 */
        o[0] = 0.0f; o[1] = 0.0f; o[2] = 0.0f;
        return;
    }
    const float scaling = (float)(n - 1);            /* lut_2d.wgsl:60 */
    const float inv_sum = scaling / S;               /* lut_2d.wgsl:63 */
    const float r = X * inv_sum, g = Y * inv_sum;    /* lut_2d.wgsl:65-66 */
    const float rfl = floorf(r), gfl = floorf(g);
    const int ri = clamp_floor_idx(rfl, n - 2);      /* lut_2d.wgsl:68-72 */
    const int gi = clamp_floor_idx(gfl, n - 2);
    const float rf = r - rfl, gf = g - gfl;          /* fract(), lut_2d.wgsl:74-75 */
    const float fs = rf + gf;
    const float *a = lut + ((size_t)(ri + 1) * n + gi) * 3;  /* lut[ri+1, gi] (lut_2d.wgsl:10-16) */
    const float *b = lut + ((size_t)ri * n + (gi + 1)) * 3;  /* lut[ri, gi+1] */
    if (fs <= 1.0f) {                                /* lower triangle, lut_2d.wgsl:81-90 */
        const float sf = 1.0f - fs;
        const float *c = lut + ((size_t)ri * n + gi) * 3;
        for (int k = 0; k < 3; ++k) o[k] = ((a[k] * rf + b[k] * gf) + c[k] * sf) * S;
    } else {                                         /* upper triangle, lut_2d.wgsl:92-104 */
        const float sf = fs - 1.0f;
        const float rf2 = 1.0f - gf, gf2 = 1.0f - rf;
        const float *c = lut + ((size_t)(ri + 1) * n + (gi + 1)) * 3;
        for (int k = 0; k < 3; ++k) o[k] = ((a[k] * rf2 + b[k] * gf2) + c[k] * sf) * S;
    }
}

void orc_apply_2d_lut(const float *xyz, int64_t npix, int cin, const float *lut, int n, float *out) {
#pragma omp parallel for schedule(static)
    for (int64_t p = 0; p < npix; ++p) lut2d_pixel(xyz + p * cin, lut, n, out + p * 3);
}

/* ---- a4: log10 with lower clip (lut_1d.wgsl:23-26) --------------------------------- */
/* log10 is evaluated in binary64 and rounded once to binary32, i.e. a correctly
 * rounded log10f up to double-rounding cases of probability ~2^-29. */
static inline float log10_clip1(float v, float eps) {
    const float c = v > eps ? v : eps;               /* max(v, eps); NaN -> eps */
    return (float)log10((double)c);
}

void orc_log_clip(float *img, int64_t count, float eps) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < count; ++i) img[i] = log10_clip1(img[i], eps);
}

/* ---- a5: per-channel 1-D curve, uniform abscissa, clamped ends ----------------------- */
/* curve is (4, N) row-major: row 0 abscissa (only its ends are used), rows 1..3 = R,G,B.
 * inv_range = float32(1 / (double(x_last) - double(x_first))) is computed by the caller
 * (gpu_processor.py:322-325 does the same on the host).  */
static inline float curve1(float v, const float *row, int N, float x0, float inv_range) {
    float t = (v - x0) * inv_range;                  /* lut_1d.wgsl:43-47 */
    t = t < 0.0f ? 0.0f : (t > 1.0f ? 1.0f : t);
    if (!(t == t)) t = 0.0f;
    const float p = t * (float)(N - 1);
    int i = (int)p;
    if (i > N - 2) i = N - 2;
    const float f = p - (float)i;
    return row[i] + f * (row[i + 1] - row[i]);
}

void orc_curve_interp(const float *in, int64_t npix, const float *curve, int N, float inv_range, float *out) {
    const float x0 = curve[0];
#pragma omp parallel for schedule(static)
    for (int64_t p = 0; p < npix; ++p)
        for (int k = 0; k < 3; ++k) out[p * 3 + k] = curve1(in[p * 3 + k], curve + (size_t)(k + 1) * N, N, x0, inv_range);
}

/* ---- a9: tetrahedral 3D LUT (utils.py:247-380) --------------------------------------- */
/* Typing follows what numba compiles for (float32[:,:,:], float32[:,:,:,:], float64):
 * coordinates/fractions are binary64, each LUT difference (cA - cB) is a binary32
 * subtraction, the products d*(cA-cB) and the running sum are binary64, and the sum is
 * rounded once to binary32 on store (utils.py:378).  Verified bit-exact, see header. */
static inline int wrap_idx(int i, int n) {           /* numba wraps negative indices */
    if (i < 0) { i += n; if (i < 0) i = 0; }
    return i;
}

static inline void tetra_pixel(const float *px, const float *lut, int n, double s, float *o) {
    double d[3];
    int i0[3];
    for (int k = 0; k < 3; ++k) {
        const double v = (double)px[k] * s;          /* utils.py:263-265 */
        double vt = v;
        if (!(vt == vt)) vt = 0.0;
        if (vt > 2.0e9) vt = 2.0e9;
        if (vt < -2.0e9) vt = -2.0e9;
        int i = (int)vt;                             /* trunc toward zero, utils.py:268-270 */
        if (i >= n - 1) { i = n - 2; d[k] = 1.0; }   /* utils.py:273-289 */
        else d[k] = v - (double)i;
        i0[k] = i;
    }
    const int r0 = wrap_idx(i0[0], n), g0 = wrap_idx(i0[1], n), b0 = wrap_idx(i0[2], n);
    const int r1 = wrap_idx(i0[0] + 1, n), g1 = wrap_idx(i0[1] + 1, n), b1 = wrap_idx(i0[2] + 1, n);
#define AT(r, g, b) (lut + (((size_t)(r) * n + (g)) * n + (b)) * 3)
    const float *c000 = AT(r0, g0, b0), *c111 = AT(r1, g1, b1);
    const float *m1, *m2; /* the two intermediate vertices of the chosen tetrahedron */
    double d1, d2, d3;
    const double dr = d[0], dg = d[1], db = d[2];
    if (dr >= dg) {
        if (dg >= db)      { m1 = AT(r1, g0, b0); m2 = AT(r1, g1, b0); d1 = dr; d2 = dg; d3 = db; } /* :300-310 */
        else if (dr >= db) { m1 = AT(r1, g0, b0); m2 = AT(r1, g0, b1); d1 = dr; d2 = db; d3 = dg; } /* :312-323 */
        else               { m1 = AT(r0, g0, b1); m2 = AT(r1, g0, b1); d1 = db; d2 = dr; d3 = dg; } /* :325-336 */
    } else {
        if (db >= dg)      { m1 = AT(r0, g0, b1); m2 = AT(r0, g1, b1); d1 = db; d2 = dg; d3 = dr; } /* :339-350 */
        else if (db >= dr) { m1 = AT(r0, g1, b0); m2 = AT(r0, g1, b1); d1 = dg; d2 = db; d3 = dr; } /* :352-363 */
        else               { m1 = AT(r0, g1, b0); m2 = AT(r1, g1, b0); d1 = dg; d2 = dr; d3 = db; } /* :365-376 */
    }
#undef AT
    for (int k = 0; k < 3; ++k) {
        const float e1 = m1[k] - c000[k];            /* binary32 differences */
        const float e2 = m2[k] - m1[k];
        const float e3 = c111[k] - m2[k];
        const double acc = (((double)c000[k] + d1 * (double)e1) + d2 * (double)e2) + d3 * (double)e3;
        o[k] = (float)acc;
    }
}

void orc_tetra(const float *in, int64_t npix, const float *lut, int n, double scale, float *out) {
    const double s = scale * (double)(n - 1);        /* utils.py:258 */
#pragma omp parallel for schedule(static)
    for (int64_t p = 0; p < npix; ++p) tetra_pixel(in + p * 3, lut, n, s, out + p * 3);
}

/* ---- a10: quantise (cpu_processor.py:407) --------------------------------------------- */
/* float32 * 255 stays float32 under NumPy 2 promotion, astype(uint8) truncates.  Values
 * outside [0,255] are clamped here (NumPy wraps them; never reached for LUTs in [0,1]). */
static inline uint8_t quant1(float v) {
    const float q = v * 255.0f;
    if (!(q > 0.0f)) return 0;
    if (q >= 255.0f) return 255;
    return (uint8_t)(int)q;
}

void orc_quantise_u8(const float *in, int64_t count, uint8_t *out) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < count; ++i) out[i] = quant1(in[i]);
}

/* ---- clip >= 0 (cpu_processor.py:397) -------------------------------------------------- */
void orc_clip_min0(float *img, int64_t count) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < count; ++i) img[i] = img[i] > 0.0f ? img[i] : 0.0f;  /* NaN -> 0 */
}

/* ---- whole pointwise chain in one pass (config C1 / C5: halation, MTF, grain off) ---- */
/* Same per-pixel functions in the order of cpu_processor.py:364,378,380,405,407; used as a
 * quick full-size checker and as the multi-threaded CPU baseline for the pointwise config. */
void orc_pointwise_chain(const float *xyz, int64_t npix, int cin,
                         const float *lut2d, int n2,
                         const float *curve, int N, float inv_range, float eps,
                         const float *lut3d, int n3, double tetra_scale,
                         uint8_t *out) {
    const float x0 = curve[0];
    const double s = tetra_scale * (double)(n3 - 1);
#pragma omp parallel for schedule(static)
    for (int64_t p = 0; p < npix; ++p) {
        float e[3], dns[3], rgb[3];
        lut2d_pixel(xyz + p * cin, lut2d, n2, e);
        for (int k = 0; k < 3; ++k)
            dns[k] = curve1(log10_clip1(e[k], eps), curve + (size_t)(k + 1) * N, N, x0, inv_range);
        tetra_pixel(dns, lut3d, n3, s, rgb);
        for (int k = 0; k < 3; ++k) out[p * 3 + k] = quant1(rgb[k]);
    }
}
