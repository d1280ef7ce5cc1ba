// C ABI of libr2f_b200.so (see include/r2f_b200.h for the contract and reference citations).
#include <cmath>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/r2f_b200.h"
#include "r2f_fft.h"
#include "r2f_kernels.h"
#include "r2f_resize.h"
#include <map>
#include <tuple>
#include <memory>

using namespace r2f;

namespace {

thread_local std::string g_err;

int fail(int code, const std::string &msg) {
    g_err = msg;
    return code;
}

int fail_cuda(cudaError_t e, const char *what) {
    g_err = std::string(what) + ": " + cudaGetErrorString(e);
    return R2F_ERR_CUDA;
}

#define CU(x)                                           \
    do {                                                \
        cudaError_t _e = (x);                           \
        if (_e != cudaSuccess) return fail_cuda(_e, #x); \
    } while (0)

struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
    cudaError_t ensure(size_t need) {
        if (need <= bytes) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
        cudaError_t e = cudaMalloc(&p, need);
        if (e == cudaSuccess) bytes = need;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
    }
};

// one spatial kernel (k x k x 3), stored per channel transposed + padded on the device
struct KernelSet {
    DevBuf buf;
    int k = 0, kp = 0;
    int mode[3] = {0, 0, 0};
    const float *chan[3] = {nullptr, nullptr, nullptr};
    bool set = false;
    // y-symmetric packed layout for k_conv2d_sym (r2f_conv_sym.cu); sym_ok when every filtered layer is
    // mirror-symmetric in y and the size is supported
    DevBuf symbuf;
    std::vector<float> sym_host;   // same contents: kernels that take the weights as a launch parameter copy from here
    size_t sym_per = 0;            // floats per layer
    int sym_channels = 0;
    const float *sym[3] = {nullptr, nullptr, nullptr};
    bool sym_ok = false;
    // FFT eligibility (r2f_fft.cu): exactly two filtered layers that share one even-symmetric base
    // kernel, K_c = alpha_c * base + beta_c * delta, third layer an exact delta.
    bool fft_ok = false;
    int fft_chan[2] = {0, 1};
    float fft_alpha[2] = {1.f, 1.f}, fft_beta[2] = {0.f, 0.f};
    // a THIRD filtered layer on the same base (black-and-white stocks filter all three layers alike,
    // effects.py:248-250): a second transform pair over (layer, layer) finishes it
    bool fft_third = false;
    int fft_chan3 = 2;
    float fft_alpha3 = 1.f, fft_beta3 = 0.f;
    DevBuf base;            // k x k base kernel, row-major, device
    uint64_t base_hash = 0;  // FNV-1a of the base kernel's bytes and size (keys the spectrum cache)
};

// device copies of the per-length FFT tables
struct FftLineDev {
    FftLineHost host;
    DevBuf roots, cosines, roots_ip, perm;
    FftLine line() const {
        FftLine l{};
        l.n = host.n;
        l.nrad = (int)host.rad.size();
        for (int i = 0; i < l.nrad; ++i) {
            l.rad[i] = host.rad[i];
            l.tw_off[i] = host.tw_off[i];
        }
        l.tw = static_cast<const float2 *>(roots.p);
        l.tw_ip = static_cast<const float2 *>(roots_ip.p);
        return l;
    }
};

}  // namespace

// Everything one film stock + settings combination needs on the device (the reference's load_* caches,
// cpu_processor.py:142-267, gpu_processor.py:792-936).  A context holds R2F_MAX_SLOTS of these so that a batch
// over mixed stocks (gui_objects.py:65-115) switches tables by index instead of re-uploading them.
struct TableSlot {
    DevBuf lut2d, lut2d4;  // packed (n, n, 3) and float4-padded copies
    int n2 = 0;
    DevBuf curve, curve_xp;  // segments; abscissa row (only when it is not uniform)
    int n1 = 0;
    float x0 = 0.f, inv_range = 0.f, eps = 1e-6f;
    DevBuf lut3d;
    int n3 = 0;
    double s3 = 0.0;
    float s3f = 0.f, margin3 = 0.f;
    int fast3 = 0;

    KernelSet hal, mtf, grain;
    DevBuf gcurve, gcurve_xp;
    int ng = 0;
    float gx0 = 0.f, ginv = 0.f;
    uint64_t seed = 0;

    bool burn_set = false;
    float d_ref = 0.f, burn_strength = 0.f, burn_scale = 50.f;

    // guarded float32 fast chain (fast_chain.cuh): host copies of what its error bound needs, derived tables
    std::vector<float> curve_host;     // the (4, N) H-D table as set
    double lut_lip[3][3] = {};         // [k][c]: largest step of output channel k between lattice neighbours along axis c
    double lut_absmax = 0.0, lut_min = 0.0, lut_max = 0.0;
    DevBuf lut255, fseg, gfseg;
    FastTetra ft{};                    // guarded float32 tetrahedral LUT of the grain / finish tails
    FastCurve gfast{};                 // conversion-free grain amplitude curve (uniform abscissa only)
    FastChain fast{};
    bool fast_valid = false;           // `fast` matches the current curve + 3-D LUT
};

struct r2f_ctx {
    int device = 0;
    int num_sms = 148;
    uint64_t launches = 0;

    std::unique_ptr<TableSlot> slots[R2F_MAX_SLOTS];
    TableSlot *t = nullptr;  // the selected slot
    int cur_slot = 0;

    // Copy-on-write table storage: a setter never overwrites a buffer a render in flight may still read.  It
    // retires the old buffer tagged with the number of renders issued so far and takes a fresh one; retired
    // buffers return to the pool once every render issued before the retirement has finished (render_done ring).
    struct Retired {
        void *p;
        size_t bytes;
        uint64_t seq;
    };
    std::vector<Retired> retired;
    std::vector<std::pair<void *, size_t>> pool;
    static constexpr int kRing = 32;
    cudaEvent_t render_done[kRing] = {};
    uint64_t issue_seq = 0, done_seq = 0;
    cudaStream_t upload_stream = nullptr;
    void *stage_host = nullptr;  // pinned staging for table uploads
    size_t stage_bytes = 0;

    DevBuf burn_buf;
    DevBuf cnr_taps;
    DevBuf expo_buf;  // r2f_calc_exposure: per-CTA partial sums + the result
    DevBuf hist_buf;  // r2f_histogram_image: counts + scalars
    DevBuf stats_buf; // fast-chain statistics (deferred pixel count)
    // r2f_resize tap tables per (kind, source size, destination size); kind 0 = area, 1 = lanczos4
    struct ResizeTab {
        DevBuf a, b, c;
        uint64_t last_use = 0;
    };
    std::map<std::tuple<int, int, int>, ResizeTab> resize_tabs;
    uint64_t resize_clock = 0;

    // r2f_render_host staging
    DevBuf h_in, h_out, h_ws, h_noise;
    cudaStream_t host_stream = nullptr;

    // FFT halation path: per-length tables, small LRU of kernel spectra keyed by kernel content + geometry
    std::map<int, std::unique_ptr<FftLineDev>> fft_lines;
    struct KhatEntry {
        uint64_t hash = 0;
        int k = 0, hp = 0, wp = 0;
        bool permuted = false;
        DevBuf buf;
        cudaEvent_t ready = nullptr;
        cudaStream_t stream = nullptr;
        uint64_t last_use = 0;
    };
    static constexpr int kKhatEntries = 4;
    KhatEntry khat[kKhatEntries];
    uint64_t khat_clock = 0;
    DevBuf khat_scratch;
    int conv_path = 0;  // R2F_OPT_CONV_PATH: 0 auto, 1 direct, 2 fft
    int conv_sym = 1;   // R2F_OPT_CONV_SYM: 1 = y-symmetric kernels take the packed-FMA kernel
    int fuse_mtf = 1;   // R2F_OPT_FUSE_MTF: 1 = banded calls issue the MTF band by band with the grain kernel
    int fast_chain = 1; // R2F_OPT_FAST_CHAIN: 1 = guarded float32 fast path of the pointwise chain

    // per-kernel profiling (r2f_profile_*)
    bool profiling = false;
    struct ProfRec {
        int id;
        cudaEvent_t a, b;
    };
    std::vector<ProfRec> prof;
};

namespace {

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// ---- copy-on-write table storage (see r2f_ctx) ---------------------------------------------------
// Renders issued so far that have finished: advance over the completed prefix of the event ring.
void sweep_done(r2f_ctx *c) {
    while (c->done_seq < c->issue_seq &&
           cudaEventQuery(c->render_done[c->done_seq % r2f_ctx::kRing]) == cudaSuccess)
        c->done_seq += 1;
    (void)cudaGetLastError();  // cudaErrorNotReady is not an error
    for (size_t i = 0; i < c->retired.size();) {
        if (c->retired[i].seq <= c->done_seq) {
            c->pool.emplace_back(c->retired[i].p, c->retired[i].bytes);
            c->retired[i] = c->retired.back();
            c->retired.pop_back();
        } else {
            ++i;
        }
    }
}

// Marks the end of one render call on its stream.
cudaError_t mark_render(r2f_ctx *c, cudaStream_t st) {
    const uint64_t seq = c->issue_seq;
    cudaEvent_t &ev = c->render_done[seq % r2f_ctx::kRing];
    if (!ev) {
        cudaError_t e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
        if (e != cudaSuccess) return e;
    } else if (c->done_seq + r2f_ctx::kRing <= seq) {
        // the ring slot still guards render seq - kRing: it has to finish before the event is re-recorded
        cudaError_t e = cudaEventSynchronize(ev);
        if (e != cudaSuccess) return e;
        sweep_done(c);
    }
    cudaError_t e = cudaEventRecord(ev, st);
    if (e == cudaSuccess) c->issue_seq = seq + 1;
    return e;
}

void retire(r2f_ctx *c, DevBuf &b) {
    if (b.p) c->retired.push_back({b.p, b.bytes, c->issue_seq});
    b.p = nullptr;
    b.bytes = 0;
}

cudaError_t acquire(r2f_ctx *c, DevBuf &b, size_t need) {
    sweep_done(c);
    size_t best = c->pool.size();
    for (size_t i = 0; i < c->pool.size(); ++i)
        if (c->pool[i].second >= need && c->pool[i].second <= 2 * need + 4096 &&
            (best == c->pool.size() || c->pool[i].second < c->pool[best].second))
            best = i;
    if (best != c->pool.size()) {
        b.p = c->pool[best].first;
        b.bytes = c->pool[best].second;
        c->pool[best] = c->pool.back();
        c->pool.pop_back();
        return cudaSuccess;
    }
    // keep the pool bounded: idle buffers are released before new memory is taken (cudaFree synchronises)
    size_t pooled = 0;
    for (auto &pb : c->pool) pooled += pb.second;
    if (pooled > ((size_t)1 << 30)) {
        for (auto &pb : c->pool) cudaFree(pb.first);
        c->pool.clear();
    }
    b.p = nullptr;
    b.bytes = 0;
    cudaError_t e = cudaMalloc(&b.p, need);
    if (e == cudaSuccess) b.bytes = need;
    return e;
}

// Host table -> a FRESH device buffer (never the one a render in flight may be reading), complete on return:
// pinned staging + asynchronous copy on the context's upload stream + synchronise.  (A plain cudaMemcpy from
// pageable memory may return before the DMA has landed, and the legacy stream does not order against the
// non-blocking streams renders run on.)
int upload(r2f_ctx *c, DevBuf &b, const void *host, size_t bytes) {
    retire(c, b);
    CU(acquire(c, b, bytes));
    if (!c->upload_stream) CU(cudaStreamCreateWithFlags(&c->upload_stream, cudaStreamNonBlocking));
    if (c->stage_bytes < bytes) {
        if (c->stage_host) cudaFreeHost(c->stage_host);
        c->stage_host = nullptr;
        c->stage_bytes = 0;
        size_t cap = bytes < ((size_t)1 << 20) ? ((size_t)1 << 20) : bytes;
        CU(cudaMallocHost(&c->stage_host, cap));
        c->stage_bytes = cap;
    }
    std::memcpy(c->stage_host, host, bytes);
    CU(cudaMemcpyAsync(b.p, c->stage_host, bytes, cudaMemcpyHostToDevice, c->upload_stream));
    CU(cudaStreamSynchronize(c->upload_stream));
    return R2F_OK;
}

// A (4, N) table's abscissa (row 0) counts as uniform when every sample lies within 1e-3 of a step of the
// straight line between its ends -- e.g. a float32 np.linspace.  Uniform tables take the reference GPU path's
// normalised lookup (shaders/lut_1d.wgsl:43-47, gpu_processor.py:322-325); anything else is evaluated with
// np.interp semantics on the stored abscissa (binary search, binary64 slope), see curve_eval.
bool abscissa_uniform(const float *xp, int N) {
    const double a = xp[0], b = xp[N - 1], step = (b - a) / (double)(N - 1);
    if (!(step > 0.0)) return step == 0.0;  // degenerate range: the normalised lookup returns row[0]
    for (int i = 0; i < N; ++i)
        if (!(std::fabs((double)xp[i] - (a + step * i)) <= 1e-3 * step)) return false;
    return true;
}

// rows 1..3 of a (4, N) curve table -> 3*N segments (value, forward difference in float32), see Curve1D;
// a non-uniform abscissa is stored next to them.
int upload_curve(r2f_ctx *c, DevBuf &b, DevBuf &xp, const float *curve, int N) {
    std::vector<float> seg((size_t)3 * N * 2);
    for (int ch = 0; ch < 3; ++ch) {
        const float *row = curve + (size_t)(ch + 1) * N;
        for (int i = 0; i < N; ++i) {
            seg[((size_t)ch * N + i) * 2] = row[i];
            seg[((size_t)ch * N + i) * 2 + 1] = i + 1 < N ? row[i + 1] - row[i] : 0.0f;
        }
    }
    int rc = upload(c, b, seg.data(), seg.size() * sizeof(float));
    if (rc != R2F_OK) return rc;
    if (abscissa_uniform(curve, N)) {
        retire(c, xp);
        return R2F_OK;
    }
    for (int i = 0; i + 1 < N; ++i)
        if (!(curve[i + 1] > curve[i])) return fail(R2F_ERR_INVALID, "curve abscissa (row 0) must be strictly increasing");
    return upload(c, xp, curve, (size_t)N * sizeof(float));
}

uint64_t fnv1a(const void *data, size_t bytes, uint64_t h = 1469598103934665603ull) {
    const unsigned char *p = static_cast<const unsigned char *>(data);
    for (size_t i = 0; i < bytes; ++i) {
        h ^= p[i];
        h *= 1099511628211ull;
    }
    return h;
}

float inv_range_of(float first, float last) {
    const double d = (double)last - (double)first;
    return d != 0.0 ? (float)(1.0 / d) : 0.0f;
}

// channels == 3: (k,k,3) interleaved; channels == 1: (k,k)
int upload_kernel(r2f_ctx *ctx, KernelSet &ks, const float *kernel, int k, int channels) {
    if (kernel == nullptr || k < 1 || (k & 1) == 0) return fail(R2F_ERR_INVALID, "kernel must be non-null with odd size");
    const int kp = (k + 3) / 4 * 4;
    const size_t per = (size_t)k * kp;
    std::vector<float> host(per * channels, 0.0f);
    int mode[3] = {1, 1, 1};
    for (int c = 0; c < channels; ++c) {
        bool delta = true;
        for (int i = 0; i < k; ++i)
            for (int j = 0; j < k; ++j) {
                const float v = kernel[((size_t)i * k + j) * channels + c];
                host[per * c + (size_t)j * kp + i] = v;
                const bool centre = (i == k / 2 && j == k / 2);
                if (centre ? (v != 1.0f) : (v != 0.0f)) delta = false;
            }
        mode[c] = delta ? 0 : 1;
    }
    int rc = upload(ctx, ks.buf, host.data(), host.size() * sizeof(float));
    if (rc != R2F_OK) return rc;
    ks.k = k;
    ks.kp = kp;
    for (int c = 0; c < 3; ++c) {
        const int src = channels == 3 ? c : 0;
        ks.mode[c] = mode[src];
        ks.chan[c] = static_cast<const float *>(ks.buf.p) + per * src;
    }
    ks.set = true;
    ks.base_hash = 0;
    ks.fft_ok = false;
    ks.fft_third = false;
    // y-symmetric layout: rows dy = 0..r, (w, w) pairs, centre row halved (exact: power-of-two scaling)
    ks.sym_ok = false;
    for (int c = 0; c < 3; ++c) ks.sym[c] = nullptr;
    if (conv_sym_supported(k)) {
        const int r = k / 2, wrow = conv_sym_wrow(k);
        const size_t sper = (size_t)(r + 1) * wrow * 2;
        std::vector<float> sh(sper * channels, 0.0f);
        bool sym = true;
        float kmax = 0.f;
        for (size_t i = 0; i < (size_t)k * k * channels; ++i) kmax = std::fmax(kmax, std::fabs(kernel[i]));
        for (int c = 0; c < channels && sym; ++c)
            for (int dy = 0; dy <= r && sym; ++dy)
                for (int j = 0; j < k; ++j) {
                    const double up = kernel[((size_t)(r - dy) * k + j) * channels + c];
                    const double dn = kernel[((size_t)(r + dy) * k + j) * channels + c];
                    if (std::fabs(up - dn) > 1e-7 * (double)kmax) {
                        sym = false;
                        break;
                    }
                    const float w = (float)(dy == 0 ? 0.25 * (up + dn) : 0.5 * (up + dn));
                    sh[sper * c + ((size_t)dy * wrow + j) * 2] = w;
                    sh[sper * c + ((size_t)dy * wrow + j) * 2 + 1] = w;
                }
        if (sym) {
            rc = upload(ctx, ks.symbuf, sh.data(), sh.size() * sizeof(float));
            if (rc != R2F_OK) return rc;
            for (int c = 0; c < 3; ++c)
                ks.sym[c] = static_cast<const float *>(ks.symbuf.p) + sper * (channels == 3 ? c : 0);
            ks.sym_host = sh;
            ks.sym_per = sper;
            ks.sym_channels = channels;
            ks.sym_ok = true;
        }
    }
    if (channels == 3 && k >= 3) {
        int conv[3], nconv = 0;
        for (int c = 0; c < 3; ++c)
            if (mode[c]) conv[nconv++] = c;
        if (nconv >= 2) {
            const int c0 = conv[0], c1 = conv[1], mid = k / 2;
            auto at = [&](int i, int j, int c) { return (double)kernel[((size_t)i * k + j) * 3 + c]; };
            bool even = true;
            for (int i = 0; i < k && even; ++i)
                for (int j = 0; j < k; ++j)
                    if (at(i, j, c0) != at(k - 1 - i, j, c0) || at(i, j, c0) != at(i, k - 1 - j, c0)) {
                        even = false;
                        break;
                    }
            // Split every filtered layer as K_c = alpha_c * B + beta_c * delta with ONE smooth base B: B is
            // layer c0 with its centre spike (the identity part of (f K + delta)/(f + 1), effects.py:255-262)
            // replaced by the largest neighbour.  Only B goes through the FFT -- its spectrum decays, so the
            // transform's rounding noise is filtered instead of passing through the delta plateau -- and
            // the identity part beta_c * x is applied exactly in float32 in the row epilogue.
            double num = 0.0, den = 0.0;
            for (int i = 0; i < k; ++i)
                for (int j = 0; j < k; ++j)
                    if (i != mid || j != mid) {
                        num += at(i, j, c1) * at(i, j, c0);
                        den += at(i, j, c0) * at(i, j, c0);
                    }
            if (even && den > 0.0) {
                const double alpha = num / den;
                double resid = 0.0;
                for (int i = 0; i < k; ++i)
                    for (int j = 0; j < k; ++j)
                        if (i != mid || j != mid) resid += std::fabs(at(i, j, c1) - alpha * at(i, j, c0));
                if (resid <= 1e-6) {
                    std::vector<float> base((size_t)k * k);
                    for (int i = 0; i < k; ++i)
                        for (int j = 0; j < k; ++j) base[(size_t)i * k + j] = kernel[((size_t)i * k + j) * 3 + c0];
                    double smooth_centre = at(mid, mid, c0);
                    if (k >= 3) {
                        const double nb = std::fmax(std::fmax(at(mid - 1, mid, c0), at(mid + 1, mid, c0)),
                                                    std::fmax(at(mid, mid - 1, c0), at(mid, mid + 1, c0)));
                        if (nb < smooth_centre) smooth_centre = nb;
                    }
                    base[(size_t)mid * k + mid] = (float)smooth_centre;
                    const double bc = (double)base[(size_t)mid * k + mid];
                    rc = upload(ctx, ks.base, base.data(), base.size() * sizeof(float));
                    if (rc != R2F_OK) return rc;
                    ks.base_hash = fnv1a(base.data(), base.size() * sizeof(float), 1469598103934665603ull ^ (uint64_t)k);
                    ks.fft_ok = true;
                    ks.fft_chan[0] = c0;
                    ks.fft_chan[1] = c1;
                    ks.fft_alpha[0] = 1.f;
                    ks.fft_beta[0] = (float)(at(mid, mid, c0) - bc);
                    ks.fft_alpha[1] = (float)alpha;
                    ks.fft_beta[1] = (float)(at(mid, mid, c1) - alpha * bc);
                    if (nconv == 3) {  // the third layer must be a multiple of the same base as well
                        const int c2 = conv[2];
                        double n3 = 0.0;
                        for (int i = 0; i < k; ++i)
                            for (int j = 0; j < k; ++j)
                                if (i != mid || j != mid) n3 += at(i, j, c2) * at(i, j, c0);
                        const double alpha3 = n3 / den;
                        double resid3 = 0.0;
                        for (int i = 0; i < k; ++i)
                            for (int j = 0; j < k; ++j)
                                if (i != mid || j != mid) resid3 += std::fabs(at(i, j, c2) - alpha3 * at(i, j, c0));
                        if (resid3 <= 1e-6) {
                            ks.fft_third = true;
                            ks.fft_chan3 = c2;
                            ks.fft_alpha3 = (float)alpha3;
                            ks.fft_beta3 = (float)(at(mid, mid, c2) - alpha3 * bc);
                        } else {
                            ks.fft_ok = false;
                        }
                    }
                }
            }
        }
    }
    return R2F_OK;
}

// ---- FFT path plumbing ------------------------------------------------------------------------
constexpr int kFftMinKernel = 9;  // below this the direct kernel is cheaper

FftLineDev *fft_line_for(r2f_ctx *c, int n);

struct FftGeometry {
    int Hp = 0, Wp = 0, nc = 0, groups = 0;
    bool inplace = false;  // in-place column kernel: one 256-thread group per column, two (or four) columns per CTA
    bool ok = false;
};

FftGeometry fft_geometry(int H, int W, int k) {
    FftGeometry g;
    const int r = k / 2;
    // multiples of 16: the line buffers are addressed through a swizzle that permutes whole 16-element rows
    g.Wp = fft_good_size(W + 2 * r, 16);
    g.Hp = fft_good_size(H + 2 * r, 16);
    if (!g.Wp || !g.Hp) return g;
    if (fft_rows_smem(g.Wp) > 227 * 1024) return g;
    if (g.Wp % 4 == 0 && fft_cols_inplace_available(g.Hp) && fft_cols_inplace_smem(g.Hp) <= 227 * 1024) {
        g.nc = 4;
        g.groups = 4;
        g.inplace = true;
    } else if (!fft_col_geometry(g.Hp, g.Wp, g.nc, g.groups)) {
        return g;
    }
    // the spectrum scratch must fit in one planar working image (3 planes of float32)
    if ((size_t)g.Wp * H * sizeof(float2) > plane_stride_for(H, W) * 3 * sizeof(float)) return g;
    g.ok = true;
    return g;
}


FftLineDev *fft_line_for(r2f_ctx *c, int n) {
    auto it = c->fft_lines.find(n);
    if (it != c->fft_lines.end()) return it->second.get();
    std::unique_ptr<FftLineDev> d(new FftLineDev());
    if (!fft_make_line(n, d->host)) return nullptr;
    if (upload(c, d->roots, d->host.roots.data(), d->host.roots.size() * sizeof(float2)) != R2F_OK) return nullptr;
    if (upload(c, d->roots_ip, d->host.roots_ip.data(), d->host.roots_ip.size() * sizeof(float2)) != R2F_OK) return nullptr;
    if (upload(c, d->perm, d->host.perm.data(), d->host.perm.size() * sizeof(int)) != R2F_OK) return nullptr;
    if (upload(c, d->cosines, d->host.cosines.data(), d->host.cosines.size() * sizeof(double)) != R2F_OK) return nullptr;
    FftLineDev *raw = d.get();
    c->fft_lines[n] = std::move(d);
    return raw;
}

// Fills the geometry-dependent parts of FftConvArgs (tables, cached kernel spectrum).
int fft_prepare(r2f_ctx *c, const KernelSet &ks, int H, int W, const FftGeometry &g, FftConvArgs &a, cudaStream_t st) {
    FftLineDev *row = fft_line_for(c, g.Wp), *col = fft_line_for(c, g.Hp);
    if (!row || !col) return fail(R2F_ERR_INVALID, "FFT plan construction failed");
    // kernel spectrum: small LRU keyed by the base kernel's content and the padded geometry, shared by all
    // table slots (stocks usually share one halation kernel)
    r2f_ctx::KhatEntry *ent = nullptr, *lru = &c->khat[0];
    for (auto &e : c->khat) {
        if (e.buf.p && e.hash == ks.base_hash && e.k == ks.k && e.hp == g.Hp && e.wp == g.Wp && e.permuted == g.inplace)
            ent = &e;
        if (!e.buf.p || (lru->buf.p && e.last_use < lru->last_use)) lru = &e;
    }
    if (!ent) {
        ent = lru;
        retire(c, ent->buf);
        CU(acquire(c, ent->buf, (size_t)g.Hp * g.Wp * sizeof(float)));
        CU(c->khat_scratch.ensure((size_t)ks.k * g.Wp * sizeof(double)));
        CU(launch_khat(static_cast<const float *>(ks.base.p), ks.k, g.Hp, g.Wp,
                       static_cast<const double *>(col->cosines.p), static_cast<const double *>(row->cosines.p),
                       static_cast<double *>(c->khat_scratch.p), static_cast<float *>(ent->buf.p),
                       g.inplace ? static_cast<const int *>(col->perm.p) : nullptr, st));
        c->launches += 2;
        if (!ent->ready) CU(cudaEventCreateWithFlags(&ent->ready, cudaEventDisableTiming));
        CU(cudaEventRecord(ent->ready, st));
        ent->stream = st;
        ent->hash = ks.base_hash;
        ent->k = ks.k;
        ent->hp = g.Hp;
        ent->wp = g.Wp;
        ent->permuted = g.inplace;
    } else if (ent->stream != st) {
        CU(cudaStreamWaitEvent(st, ent->ready, 0));  // built on another stream: order this render after it
    }
    ent->last_use = ++c->khat_clock;
    a.H = H;
    a.W = W;
    a.r = ks.k / 2;
    a.row_off = (4 - a.r % 4) % 4;
    if (W + 2 * a.r + a.row_off > g.Wp) a.row_off = 0;   // no slack in the padded length: unshifted, 8-byte accesses
    a.row = row->line();
    a.col = col->line();
    a.nc = g.nc;
    a.col_groups = g.groups;
    a.col_inplace = g.inplace ? 1 : 0;
    a.cols_prefetch = (size_t)g.Hp * g.Wp * 2 > ((size_t)64 << 20) ? 1 : 0;  // half the spectrum is stored (mirror rows)
    {
        static const char *ra = getenv("R2F_FFT_ROWS_AHEAD");  // tuning knob; default: the CTAs resident on the device
        a.rows_ahead = ra ? atoi(ra) : c->num_sms;  // measured at 24 MP: off 0.308, 148 0.302, 296 0.304 ms
    }
    a.khat = static_cast<const float *>(ent->buf.p);
    for (int i = 0; i < 2; ++i) {
        a.chan[i] = ks.fft_chan[i];
        a.alpha[i] = ks.fft_alpha[i];
        a.beta[i] = ks.fft_beta[i];
    }
    a.plane_stride = plane_stride_for(H, W);
    return R2F_OK;
}

bool want_fft(const r2f_ctx *c, const KernelSet &ks, int H, int W, FftGeometry &g) {
    if (c->conv_path == 1 || !ks.fft_ok) return false;
    if (c->conv_path == 0 && ks.k < kFftMinKernel) return false;
    g = fft_geometry(H, W, ks.k);
    return g.ok;
}

Lut2D lut2d_of(const r2f_ctx *c) {
    return Lut2D(static_cast<const float *>(c->t->lut2d.p), static_cast<const float4 *>(c->t->lut2d4.p), c->t->n2);
}
Curve1D curve_of(const r2f_ctx *c) {
    return Curve1D{static_cast<const float2 *>(c->t->curve.p), c->t->n1, c->t->x0, c->t->inv_range,
                   static_cast<const float *>(c->t->curve_xp.p)};
}
Curve1D gcurve_of(const r2f_ctx *c) {
    return Curve1D{static_cast<const float2 *>(c->t->gcurve.p), c->t->ng, c->t->gx0, c->t->ginv,
                   static_cast<const float *>(c->t->gcurve_xp.p)};
}
Lut3D lut3d_of(const r2f_ctx *c) {
    const TableSlot *t = c->t;
    return Lut3D{static_cast<const float4 *>(t->lut3d.p), t->n3, t->s3, t->s3f, t->margin3, t->fast3};
}

ConvArgs conv_args(const KernelSet &ks, const float *in, float *out, size_t ps, int H, int W) {
    ConvArgs a{};
    a.in = in;
    a.out = out;
    a.aux = nullptr;
    a.plane_stride = ps;
    a.H = H;
    a.W = W;
    a.k = ks.k;
    a.kp = ks.kp;
    for (int c = 0; c < 3; ++c) {
        a.kern[c] = ks.chan[c];
        a.ksym[c] = ks.sym_ok ? ks.sym[c] : nullptr;
        a.ksym_host[c] = ks.sym_ok ? ks.sym_host.data() + ks.sym_per * (ks.sym_channels == 3 ? c : 0) : nullptr;
        a.mode[c] = ks.mode[c];
        a.in_plane[c] = c;
    }
    a.epi = EPI_NONE;
    a.eps = 0.f;
    return a;
}

// guarded float32 tetrahedral tail of the selected tables (ok == 0 when the option or the tables rule it out)
FastTetra ft_of(const r2f_ctx *c) {
    FastTetra ft = c->t->ft;
    if (!c->fast_chain) ft.ok = 0;
    return ft;
}

// direct correlation: the y-symmetric packed-FMA kernel when the kernel set allows it, else the generic one
bool conv_takes_sym(const r2f_ctx *c, const ConvArgs &a) {
    const bool any_conv = a.mode[0] || a.mode[1] || a.mode[2];
    const bool nonuniform = a.epi != EPI_NONE && a.curve.xp != nullptr;  // only the generic kernel runs np.interp
    return c->conv_sym && any_conv && a.ksym[0] && a.ksym[1] && a.ksym[2] && a.epi != EPI_GRAIN && !nonuniform;
}
cudaError_t conv_dispatch(const r2f_ctx *c, const ConvArgs &a, cudaStream_t st) {
    if (conv_takes_sym(c, a)) return launch_conv2d_sym(a, st);
    if (a.tile_rows > 0) return cudaErrorInvalidValue;  // only the symmetric kernel renders a band of tile rows
    return launch_conv2d(a, st);
}

ConvArgs identity_args(const float *in, float *out, size_t ps, int H, int W) {
    ConvArgs a{};
    a.in = in;
    a.out = out;
    a.plane_stride = ps;
    a.H = H;
    a.W = W;
    a.k = 1;
    a.kp = 4;
    for (int c = 0; c < 3; ++c) {
        a.kern[c] = nullptr;
        a.ksym[c] = nullptr;
        a.ksym_host[c] = nullptr;
        a.mode[c] = 0;
        a.in_plane[c] = c;
    }
    return a;
}

// Brackets the launches issued while it is alive with a pair of events on `st`.
struct ProfScope {
    r2f_ctx *c;
    cudaStream_t st;
    int id;
    cudaEvent_t a = nullptr, b = nullptr;
    ProfScope(r2f_ctx *ctx, cudaStream_t s, int kid) : c(ctx), st(s), id(kid) {
        if (!c->profiling) return;
        if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) {
            a = b = nullptr;
            return;
        }
        cudaEventRecord(a, st);
    }
    ~ProfScope() {
        if (!a) return;
        cudaEventRecord(b, st);
        c->prof.push_back({id, a, b});
    }
};

struct BurnDims {
    int lh, lw, zh, zw;
};

BurnDims burn_dims(int H, int W, float burn_scale) {
    const int step = (int)std::ceil((double)std::min(H, W) / (double)burn_scale);  // effects.py:365
    BurnDims d;
    d.lw = W / step;                                                                // effects.py:372
    d.lh = H / step;
    d.zh = (int)std::lrint((double)d.lh * step);                                    // scipy zoom output shape
    d.zw = (int)std::lrint((double)d.lw * step);
    return d;
}

// Derives the fast chain's tables and its error bound from the slot's H-D curve and 3-D LUT (see fast_chain.cuh
// for the argument).  u = 2^-24; every term is a worst case over the whole table, so the bound holds for any pixel.
int fast_chain_build(r2f_ctx *c) {
    TableSlot *t = c->t;
    if (t->fast_valid) return R2F_OK;
    t->fast = FastChain{};
    t->fast_valid = true;
    const int N = t->n1, n = t->n3;
    if (t->curve_xp.p != nullptr || t->curve_host.size() != (size_t)4 * N || N < 2 || n < 2) return R2F_OK;
    const float *cv = t->curve_host.data();
    const double x0 = cv[0], xN = cv[N - 1], R = xN - x0;
    if (!(R > 0.0) || !(t->eps >= 1e-30f) || !(t->s3 > 0.0)) return R2F_OK;
    if (!(t->lut_min >= 0.0) || !(t->lut_max <= 1.0)) return R2F_OK;  // quantisation without clamps needs [0, 1]
    const double u = std::ldexp(1.0, -24), s3 = t->s3;
    double dmin = 1e300, dmax = -1e300, maxseg[3] = {0, 0, 0}, dabs[3] = {0, 0, 0};
    for (int ch = 0; ch < 3; ++ch) {
        const float *row = cv + (size_t)(ch + 1) * N;
        for (int i = 0; i < N; ++i) {
            if (!std::isfinite(row[i])) return R2F_OK;
            dmin = std::fmin(dmin, row[i]);
            dmax = std::fmax(dmax, row[i]);
            dabs[ch] = std::fmax(dabs[ch], std::fabs((double)row[i]));
            if (i + 1 < N) maxseg[ch] = std::fmax(maxseg[ch], std::fabs((double)row[i + 1] - (double)row[i]));
        }
    }
    if (!(dmin >= 0.0) || !(dmax * s3 <= (double)(n - 1) - 0.01)) return R2F_OK;  // stay inside the lattice: no clamps
    if (n > 129 || N > (1 << 20) || t->n2 > 1024) return R2F_OK;              // index arithmetic of the fast chain
    // abscissa coordinate p: exact chain vs fast chain (both against the true value)
    const double Lm = std::fmax(std::fabs(x0), std::fabs(xN)) + 1.0, log10_2 = 0.30102999566398119521;
    const double et_exact = (u * Lm + u * (R + 2.0)) / R + 3.0 * u;
    const double e_l2 = std::ldexp(1.0, -22) + 2.0 * u * (Lm / log10_2);
    const double cA = log10_2 / R, cB = -x0 / R;
    const double et_fast = e_l2 * cA + u * (Lm / R + std::fabs(x0) / R + 1.5);
    const double ep = (N - 1) * (et_exact + et_fast) + 4.0 * u * (N - 1);
    double dv[3];
    for (int ch = 0; ch < 3; ++ch)
        dv[ch] = s3 * (ep * maxseg[ch] + 2.0 * u * dabs[ch]) + 3.0 * u * (dabs[ch] * s3);
    double margin = 0.0, lipmax = 0.0;
    for (int k = 0; k < 3; ++k)
        for (int ch = 0; ch < 3; ++ch) lipmax = std::fmax(lipmax, t->lut_lip[k][ch]);
    for (int k = 0; k < 3; ++k) {
        double e = 0.0;
        for (int ch = 0; ch < 3; ++ch) e += dv[ch] * t->lut_lip[k][ch];
        e = 255.0 * (e + u * t->lut_absmax + 3.0 * u * lipmax) + 10.0 * u * 255.0 * t->lut_absmax + u * 255.0;
        margin = std::fmax(margin, e);
    }
    margin *= 1.25;  // safety factor
    if (!(margin < 0.2)) return R2F_OK;  // most pixels would be deferred: not worth it
    // scaled segments
    std::vector<float> fs((size_t)3 * N * 2);
    for (int ch = 0; ch < 3; ++ch) {
        const float *row = cv + (size_t)(ch + 1) * N;
        for (int i = 0; i < N; ++i) {
            // (midpoint of the segment in lattice units - 0.5, its forward difference): see fast_chain.cuh
            const double d = i + 1 < N ? (double)row[i + 1] - (double)row[i] : 0.0;
            fs[((size_t)ch * N + i) * 2] = (float)(((double)row[i] + 0.5 * d) * s3 - 0.5);
            fs[((size_t)ch * N + i) * 2 + 1] = (float)(d * s3);
        }
    }
    int rc = upload(c, t->fseg, fs.data(), fs.size() * sizeof(float));
    if (rc != R2F_OK) return rc;
    FastChain f{};
    f.cA = (float)cA;
    f.cB = (float)cB;
    f.pscale = std::nextafterf((float)(N - 1), 0.0f);
    f.margin = (float)margin;
    f.fseg = static_cast<const float2 *>(t->fseg.p);
    f.lut255 = static_cast<const float4 *>(t->lut255.p) + ((size_t)n * n + n + 1);  // past the guard cells
    f.N = N;
    f.n3 = n;
    f.ok = t->lut255.p != nullptr ? 1 : 0;
    t->fast = f;
    return R2F_OK;
}

int check_tables(const r2f_ctx *c, unsigned flags) {
    const TableSlot *t = c->t;
    if (!t->lut2d.p) return fail(R2F_ERR_INVALID, "2D input LUT not set (r2f_set_lut2d)");
    if (!t->curve.p) return fail(R2F_ERR_INVALID, "density curve not set (r2f_set_curve1d)");
    if (!t->lut3d.p) return fail(R2F_ERR_INVALID, "3D output LUT not set (r2f_set_lut3d)");
    if ((flags & R2F_HALATION) && !t->hal.set) return fail(R2F_ERR_INVALID, "halation kernel not set");
    if ((flags & R2F_MTF) && !t->mtf.set) return fail(R2F_ERR_INVALID, "MTF kernel not set");
    if ((flags & R2F_GRAIN) && (!t->grain.set || !t->gcurve.p)) return fail(R2F_ERR_INVALID, "grain tables not set");
    if ((flags & R2F_BURN) && !t->burn_set) return fail(R2F_ERR_INVALID, "burn parameters not set");
    return R2F_OK;
}

// Banded streaming (r2f_render_banded): the frame arrives and the result leaves in `n` horizontal bands.  The first
// kernel of the pipeline is launched per band as soon as that band's rows are on the device (in_ready[i]), and
// the last one per band with an event after each (out_done[i]) so that the caller can start copying the result
// out while later bands are still being computed.  Band i covers rows [band_row(i), band_row(i + 1)).
struct Bands {
    int n = 0;
    cudaEvent_t const *in_ready = nullptr;
    cudaEvent_t const *out_done = nullptr;
    int waited = 0, recorded = 0;  // progress, so that fallbacks can wait for / record "the rest"
};

int band_row(int H, int n, int i) {  // multiples of 64 (tile height), 0 and H at the ends
    if (i <= 0) return 0;
    if (i >= n) return H;
    const long long r = ((long long)i * H / n + 63) / 64 * 64;
    return (int)(r < H ? r : H);
}

// The whole pipeline.  tap_stage == 0: normal render to out_u8.
int render_body(r2f_ctx *c, const void *in, int in_format, float in_gain, int H, int W, int cin, uint8_t *out_u8,
                unsigned flags, const float *noise, int noise_ch, void *ws, size_t ws_bytes, int tap_stage,
                float *tap, cudaStream_t st, Bands *bands) {
    if (!c) return fail(R2F_ERR_INVALID, "null context");
    if (!in || H < 1 || W < 1 || (cin != 3 && cin != 4)) return fail(R2F_ERR_INVALID, "bad input image arguments");
    if (in_format != R2F_IN_F32 && in_format != R2F_IN_U16) return fail(R2F_ERR_INVALID, "unknown input format");
    const int fmt = (in_format == R2F_IN_U16 ? 2 : 0) + (cin == 4 ? 1 : 0);  // kFmt* of device_math.cuh
    if (tap_stage == 0 && !out_u8) return fail(R2F_ERR_INVALID, "null output");
    if (tap_stage != 0 && (!tap || tap_stage < R2F_TAP_EXPOSURE || tap_stage > R2F_TAP_RGB))
        return fail(R2F_ERR_INVALID, "bad tap arguments");
    if ((reinterpret_cast<uintptr_t>(in) & 15) != 0) return fail(R2F_ERR_INVALID, "input must be 16-byte aligned");
    if (out_u8 && (reinterpret_cast<uintptr_t>(out_u8) & 3) != 0)
        return fail(R2F_ERR_INVALID, "output must be 4-byte aligned");
    int rc = check_tables(c, flags);
    if (rc != R2F_OK) return rc;
    DeviceGuard guard(c->device);

    const int nb = bands ? bands->n : 1;
    auto wait_in = [&](int upto) -> cudaError_t {  // the rows of bands [0, upto) are needed
        for (; bands && bands->waited < upto && bands->waited < nb; ++bands->waited) {
            cudaError_t e = cudaStreamWaitEvent(st, bands->in_ready[bands->waited], 0);
            if (e != cudaSuccess) return e;
        }
        return cudaSuccess;
    };
    auto done_out = [&](int upto) -> cudaError_t {  // the output rows of bands [0, upto) are final
        for (; bands && bands->recorded < upto && bands->recorded < nb; ++bands->recorded) {
            cudaError_t e = cudaEventRecord(bands->out_done[bands->recorded], st);
            if (e != cudaSuccess) return e;
        }
        return cudaSuccess;
    };
    const size_t in_px_bytes = (size_t)cin * (in_format == R2F_IN_U16 ? 2 : 4);
    if (tap_stage != 0) CU(wait_in(nb));  // the tap paths are not banded

    const size_t npix = (size_t)H * W;
    const unsigned spatial = flags & (R2F_HALATION | R2F_MTF | R2F_GRAIN | R2F_BURN);
    const Lut2D l2 = lut2d_of(c);
    const Curve1D cv = curve_of(c);
    const Lut3D l3 = lut3d_of(c);

    if (tap_stage == 0 && spatial == 0) {  // configs C1 / C5: one fused pass
        if (c->fast_chain) {
            rc = fast_chain_build(c);
            if (rc != R2F_OK) return rc;
        }
        const FastChain &fc = c->t->fast;
        ProfScope ps_(c, st, R2F_PROF_POINTWISE);
        const bool fast = c->fast_chain && fc.ok && pointwise_fast_smem(l2, fc) <= 104 * 1024 && npix < ((size_t)1 << 32);
        if (fast && !c->stats_buf.p) {
            CU(c->stats_buf.ensure(sizeof(unsigned long long)));
            CU(cudaMemsetAsync(c->stats_buf.p, 0, sizeof(unsigned long long), st));
        }
        for (int b = 0; b < nb; ++b) {  // one launch per band (one launch in all when the call is not banded)
            const int r0 = band_row(H, nb, b), r1 = band_row(H, nb, b + 1);
            if (r1 <= r0) continue;
            CU(wait_in(b + 1));
            const void *bin = static_cast<const char *>(in) + (size_t)r0 * W * in_px_bytes;
            uint8_t *bout = out_u8 + (size_t)r0 * W * 3;
            const size_t bpix = (size_t)(r1 - r0) * W;
            if (fast)
                CU(launch_pointwise_fast(bin, fmt, in_gain, bout, bpix, l2, cv, c->t->eps, l3, fc,
                                         static_cast<unsigned long long *>(c->stats_buf.p), c->num_sms, st));
            else
                CU(launch_pointwise(bin, fmt, in_gain, bout, bpix, l2, cv, c->t->eps, l3, c->num_sms, st));
            c->launches += 1;
            CU(done_out(b + 1));
        }
        return R2F_OK;
    }

    const size_t need = r2f_workspace_bytes(H, W, flags);
    if (!ws || ws_bytes < need) return fail(R2F_ERR_NOMEM, "workspace too small (see r2f_workspace_bytes)");
    if ((reinterpret_cast<uintptr_t>(ws) & 255) != 0) return fail(R2F_ERR_INVALID, "workspace must be 256-byte aligned");
    const size_t ps = plane_stride_for(H, W);
    Planes P[3];
    for (int i = 0; i < 3; ++i) P[i] = Planes{static_cast<float *>(ws) + (size_t)i * 3 * ps, ps};
    auto export_tap = [&](Planes p) -> int {
        CU(launch_planar_to_interleaved(p, tap, npix, c->num_sms, st));
        c->launches += 1;
        return R2F_OK;
    };

    // a2 (+ a3 + a4 + a5): exposure, halation, density
    FftGeometry geo;
    const bool hal_fft = (flags & R2F_HALATION) && want_fft(c, c->t->hal, H, W, geo);
    if (c->conv_path == 2 && (flags & R2F_HALATION) && !hal_fft)
        return fail(R2F_ERR_INVALID, "FFT path forced (R2F_OPT_CONV_PATH=2) but this kernel/frame is not eligible");
    // Odd frame widths: when the density planes go straight from k_fft_rows_inv through the symmetric MTF kernel into
    // the fused symmetric grain kernel (the default full emulation), their rows are padded to a multiple of 4 floats so
    // that TMA, the 16-byte copies and the 128-bit write-out keep working (plane_stride_for reserves the room).
    int plane_pitch = W;
    if ((W & 3) != 0 && hal_fft && tap_stage == 0 && !c->t->hal.fft_third && cv.xp == nullptr &&
        (flags & R2F_GRAIN) && !(flags & R2F_BURN) && c->conv_sym && c->t->grain.sym_ok &&
        grain_finish_sym_supported(c->t->grain.k) && gcurve_of(c).xp == nullptr &&
        (size_t)(64 + c->t->grain.k - 1) * (64 + c->t->grain.k - 1) * 4 + (size_t)c->t->grain.k * c->t->grain.kp * 4 + 16384 <=
            200 * 1024) {
        bool mtf_ok = true;
        if (flags & R2F_MTF) {
            ConvArgs probe = conv_args(c->t->mtf, P[1].base, P[0].base, ps, H, W);
            mtf_ok = conv_takes_sym(c, probe);
        }
        if (mtf_ok) plane_pitch = padded_pitch(W);
    }
    if (hal_fft && tap_stage != R2F_TAP_EXPOSURE) {
        // XYZ -> 2-D LUT is fused into the row transforms; the exposure image is never materialised
        FftConvArgs fa{};
        rc = fft_prepare(c, c->t->hal, H, W, geo, fa, st);
        if (rc != R2F_OK) return rc;
        fa.S = reinterpret_cast<float2 *>(P[2].base);
        fa.src_planar = nullptr;
        fa.src_xyz = in;
        fa.gain = in_gain;
        fa.lut2d = l2;
        fa.exp_planar = P[0].base;
        fa.dst_planar = P[1].base;
        fa.curve = cv;
        fa.eps = c->t->eps;
        // the row-inverse kernel's fused log10 + curve epilogue knows uniform tables only
        const bool third = c->t->hal.fft_third;  // three filtered layers: a second pass finishes the last one
        const bool fuse_final = tap_stage != R2F_TAP_HALATION && cv.xp == nullptr;
        const bool fuse_density = fuse_final && !third;
        fa.dst_pitch = (fuse_density && plane_pitch != W) ? plane_pitch : 0;
        {
            ProfScope ps_(c, st, R2F_PROF_FFT_ROWS_FWD);
            for (int b = 0; b < nb; ++b) {  // the row transforms of a band start as soon as its rows have arrived
                const int r0 = band_row(H, nb, b), r1 = band_row(H, nb, b + 1);
                if (r1 <= r0) continue;
                CU(wait_in(b + 1));
                fa.row0 = r0;
                fa.row_count = r1 - r0;
                CU(launch_fft_conv(fa, 1 + fmt, fuse_density, st, 1));
            }
            fa.row0 = fa.row_count = 0;
        }
        for (int stage = 2; stage <= 3; ++stage) {
            ProfScope ps_(c, st, R2F_PROF_FFT_ROWS_FWD + stage - 1);
            CU(launch_fft_conv(fa, 1 + fmt, fuse_density, st, stage));
        }
        if (third) {
            // P[1] now holds the two filtered layers and the third one untouched: transform (layer3 + i layer3), mix
            // it in place, pass the finished layers through, and apply the density epilogue to all three
            FftConvArgs fb = fa;
            fb.src_xyz = nullptr;
            fb.src_planar = P[1].base;
            fb.exp_planar = nullptr;
            fb.dst_planar = P[1].base;
            fb.chan[0] = fb.chan[1] = c->t->hal.fft_chan3;
            fb.alpha[0] = fb.alpha[1] = c->t->hal.fft_alpha3;
            fb.beta[0] = fb.beta[1] = c->t->hal.fft_beta3;
            for (int stage = 1; stage <= 3; ++stage) {
                ProfScope ps_(c, st, R2F_PROF_FFT_ROWS_FWD + stage - 1);
                CU(launch_fft_conv(fb, 0, fuse_final, st, stage));
            }
            c->launches += 3;
        }
        const bool fuse_done = third ? fuse_final : fuse_density;
        c->launches += 2;  // + the one counted below
        if (tap_stage == R2F_TAP_HALATION) {
            c->launches += 1;
            return export_tap(P[1]);
        }
        if (!fuse_done) {  // separate density pass (generic kernel, np.interp lookup): P[1] -> P[0] -> P[1]
            ConvArgs a = identity_args(P[1].base, P[0].base, ps, H, W);
            a.epi = EPI_DENSITY_FAST;
            a.curve = cv;
            a.eps = c->t->eps;
            ProfScope ps_(c, st, R2F_PROF_DENSITY);
            CU(launch_conv2d(a, st));
            CU(cudaMemcpyAsync(P[1].base, P[0].base, ps * 3 * sizeof(float), cudaMemcpyDeviceToDevice, st));
            c->launches += 1;
        }
    } else {
        {
            ProfScope ps_(c, st, R2F_PROF_EXPOSE);
            for (int b = 0; b < nb; ++b) {
                const int r0 = band_row(H, nb, b), r1 = band_row(H, nb, b + 1);
                if (r1 <= r0) continue;
                CU(wait_in(b + 1));
                CU(launch_expose(static_cast<const char *>(in) + (size_t)r0 * W * in_px_bytes, fmt, in_gain,
                                 Planes{P[0].base + (size_t)r0 * W, ps}, (size_t)(r1 - r0) * W, l2, c->num_sms, st));
            }
        }
        c->launches += 1;
        if (tap_stage == R2F_TAP_EXPOSURE) return export_tap(P[0]);
        if (flags & R2F_HALATION) {
            ConvArgs a = conv_args(c->t->hal, P[0].base, P[1].base, ps, H, W);
            if (tap_stage == R2F_TAP_HALATION) {
                CU(conv_dispatch(c, a, st));
                c->launches += 1;
                return export_tap(P[1]);
            }
            a.epi = EPI_DENSITY_FAST;
            a.curve = cv;
            a.eps = c->t->eps;
            ProfScope ps_(c, st, R2F_PROF_HALATION);
            CU(conv_dispatch(c, a, st));
        } else {
            if (tap_stage == R2F_TAP_HALATION) return fail(R2F_ERR_INVALID, "halation tap requested but stage is off");
            ConvArgs a = identity_args(P[0].base, P[1].base, ps, H, W);
            a.epi = EPI_DENSITY;
            a.curve = cv;
            a.eps = c->t->eps;
            ProfScope ps_(c, st, R2F_PROF_DENSITY);
            CU(conv_dispatch(c, a, st));
        }
    }
    c->launches += 1;
    int cur = 1;
    if (tap_stage == R2F_TAP_DENSITY) return export_tap(P[cur]);

    // a6: MTF.  A banded call whose tail is MTF -> fused grain + finish runs the two kernels band by band (both are
    // tile-local; the MTF reads its halo rows from the complete density planes), so that the first rows of the result
    // can leave the device one band of MTF + grain after the density is complete, not a whole-frame MTF later.
    const bool fused_grain = tap_stage == 0 && (flags & R2F_GRAIN) && !(flags & R2F_BURN) &&
                             (size_t)(64 + c->t->grain.k - 1) * (64 + c->t->grain.k - 1) * 4 +
                                     (size_t)c->t->grain.k * c->t->grain.kp * 4 + 16384 <= 200 * 1024;
    ConvArgs mtf_band{};
    bool band_mtf = false;
    if (flags & R2F_MTF) {
        ConvArgs a = conv_args(c->t->mtf, P[cur].base, P[1 - cur].base, ps, H, W);
        a.pitch = plane_pitch != W ? plane_pitch : 0;
        if (nb > 1 && c->fuse_mtf && fused_grain && conv_takes_sym(c, a)) {
            mtf_band = a;
            band_mtf = true;
        } else {
            ProfScope ps_(c, st, R2F_PROF_MTF);
            CU(conv_dispatch(c, a, st));
        }
        c->launches += 1;
        cur = 1 - cur;
    }
    if (tap_stage == R2F_TAP_MTF) return export_tap(P[cur]);

    // a7 + a8 + a9 + a10 fused (normal render): noise regenerated per tile, nothing but the density
    // read and the uint8 write touches HBM.  Taps keep the staged kernels below.
    // Arguments of the fused grain kernels (noise regenerated per tile, or the injected field).
    auto grain_args = [&](GrainFinishArgs &ga, int nch) -> int {
        ga.dens = P[cur].base;
        ga.noise = nullptr;
        if (noise) {
            if (noise_ch != nch) return fail(R2F_ERR_INVALID, "injected noise has the wrong channel count");
            ProfScope ps_(c, st, R2F_PROF_NOISE);
            CU(launch_interleaved_to_planar(noise, nch, nch, P[2], npix, c->num_sms, st));
            c->launches += 1;
            ga.noise = P[2].base;
        }
        ga.plane_stride = ps;
        ga.H = H;
        ga.W = W;
        ga.gk = c->t->grain.chan[0];
        ga.gk_sym = c->t->grain.sym_ok ? c->t->grain.sym[0] : nullptr;
        if (ga.gk_sym && c->t->grain.sym_per * sizeof(float) <= sizeof(ga.gkw))
            memcpy(ga.gkw, c->t->grain.sym_host.data(), c->t->grain.sym_per * sizeof(float));
        ga.k = c->t->grain.k;
        ga.kp = c->t->grain.kp;
        ga.bw = nch == 1;
        ga.noise_shift = (c->t->grain.k / 2) & 3;
        ga.seed_lo = (uint32_t)c->t->seed;
        ga.seed_hi = (uint32_t)(c->t->seed >> 32);
        ga.gcurve = gcurve_of(c);
        ga.gfast = c->fast_chain ? c->t->gfast : FastCurve{};
        ga.l3 = l3;
        ga.ft = c->t->ft;
        if (!c->fast_chain) ga.ft.ok = 0;
        ga.burn = BurnArgs{};
        ga.out_u8 = out_u8;
        ga.dens_out = nullptr;
        return R2F_OK;
    };
    if (fused_grain) {
        const int nch = (flags & R2F_GRAIN_BW) ? 1 : 3;
        GrainFinishArgs ga{};
        const int grc = grain_args(ga, nch);
        if (grc != R2F_OK) return grc;
        ga.dens_pitch = plane_pitch != W ? plane_pitch : 0;
        ProfScope ps_(c, st, R2F_PROF_GRAIN);
        const bool gsym = c->conv_sym && ga.gk_sym && grain_finish_sym_supported(ga.k) && ga.gcurve.xp == nullptr;
        for (int b = 0; b < nb; ++b) {  // the result of a band can leave while the next one is computed
            const int r0 = band_row(H, nb, b), r1 = band_row(H, nb, b + 1);
            if (r1 <= r0) continue;
            ga.tile_y0 = nb > 1 ? r0 / 64 : 0;
            ga.tile_rows = nb > 1 ? (r1 - r0 + 63) / 64 : 0;
            if (band_mtf) {
                mtf_band.tile_y0 = ga.tile_y0;
                mtf_band.tile_rows = ga.tile_rows;
                CU(launch_conv2d_sym(mtf_band, st));
            }
            if (gsym) CU(launch_grain_finish_sym(ga, st));
            else CU(launch_grain_finish(ga, st));
            CU(done_out(b + 1));
        }
        c->launches += 1;
        return R2F_OK;
    }

    // a7 with the burn stage behind it (normal render): the same fused kernel stops after the grain stage and writes the
    // grained density planes in place (the burn mask is a low-resolution statistic of the whole grained frame, so the
    // tetrahedral tail cannot run in the same pass); `k_finish` applies burn + LUT afterwards.  Replaces the staged
    // noise field + generic correlation of the tap path (0.12 + 0.72 ms at 24 MP) by one 0.4 ms pass.
    bool grain_done = false;
    if (tap_stage == 0 && (flags & R2F_GRAIN) && (flags & R2F_BURN) && c->conv_sym && c->t->grain.sym_ok &&
        grain_finish_sym_supported(c->t->grain.k) && gcurve_of(c).xp == nullptr) {
        const int nch = (flags & R2F_GRAIN_BW) ? 1 : 3;
        GrainFinishArgs ga{};
        const int grc = grain_args(ga, nch);
        if (grc != R2F_OK) return grc;
        ga.dens_out = P[cur].base;
        ga.out_u8 = nullptr;
        ProfScope ps_(c, st, R2F_PROF_GRAIN);
        CU(launch_grain_finish_sym(ga, st));
        c->launches += 1;
        grain_done = true;
    }

    // a7: grain (noise -> grain-kernel correlation -> amplitude from density -> add -> clip >= 0)
    if ((flags & R2F_GRAIN) && !grain_done) {
        const int nch = (flags & R2F_GRAIN_BW) ? 1 : 3;
        if (noise) {
            if (noise_ch != nch) return fail(R2F_ERR_INVALID, "injected noise has the wrong channel count");
            ProfScope ps_(c, st, R2F_PROF_NOISE);
            CU(launch_interleaved_to_planar(noise, nch, nch, P[2], npix, c->num_sms, st));
        } else {
            ProfScope ps_(c, st, R2F_PROF_NOISE);
            CU(launch_noise(P[2], nch, H, W, c->t->seed, (c->t->grain.k / 2) & 3, c->num_sms, st));
        }
        ConvArgs a = conv_args(c->t->grain, P[2].base, P[1 - cur].base, ps, H, W);
        for (int ch = 0; ch < 3; ++ch) a.in_plane[ch] = nch == 1 ? 0 : ch;
        a.aux = P[cur].base;
        a.epi = EPI_GRAIN;
        a.curve = gcurve_of(c);
        ProfScope ps_(c, st, R2F_PROF_GRAIN);
        CU(conv_dispatch(c, a, st));
        c->launches += 2;
        cur = 1 - cur;
    }
    if (tap_stage == R2F_TAP_GRAIN) return export_tap(P[cur]);

    // a8: burn mask
    BurnArgs burn{};
    if (flags & R2F_BURN) {
        const BurnDims bd = burn_dims(H, W, c->t->burn_scale);
        if (bd.lh < 1 || bd.lw < 1) return fail(R2F_ERR_INVALID, "burn_scale too small for this frame");
        const size_t n = (size_t)bd.lh * bd.lw;
        CU(c->burn_buf.ensure(2 * n * sizeof(float)));
        float *map = static_cast<float *>(c->burn_buf.p), *tmp = map + n;
        ProfScope ps_(c, st, R2F_PROF_BURN);
        CU(launch_burn_mask(P[cur].base + ps, H, W, bd.lh, bd.lw, c->t->d_ref, tmp, map, st));
        c->launches += 3;
        burn.map = map;
        burn.lh = bd.lh;
        burn.lw = bd.lw;
        burn.zh = bd.zh;
        burn.zw = bd.zw;
        burn.strength = c->t->burn_strength;
    }
    if (tap_stage == R2F_TAP_BURN) {
        CU(launch_finish(P[cur], npix, H, W, l3, burn, nullptr, tap, 0, c->num_sms, st));
        c->launches += 1;
        return R2F_OK;
    }

    // a9 + a10
    ProfScope ps_(c, st, R2F_PROF_FINISH);
    if (tap_stage == R2F_TAP_RGB)
        CU(launch_finish(P[cur], npix, H, W, l3, burn, nullptr, tap, 1, c->num_sms, st));
    else if (burn.map != nullptr || nb == 1)
        CU(launch_finish(P[cur], npix, H, W, l3, burn, out_u8, nullptr, 1, c->num_sms, st, ft_of(c)));
    else
        for (int b = 0; b < nb; ++b) {
            const int r0 = band_row(H, nb, b), r1 = band_row(H, nb, b + 1);
            if (r1 <= r0) continue;
            CU(launch_finish(Planes{P[cur].base + (size_t)r0 * W, ps}, (size_t)(r1 - r0) * W, r1 - r0, W, l3, burn,
                             out_u8 + (size_t)r0 * W * 3, nullptr, 1, c->num_sms, st, ft_of(c)));
            CU(done_out(b + 1));
        }
    c->launches += 1;
    return R2F_OK;
}


// Every render call ends with an event on its stream: the copy-on-write table storage uses it to know when
// a replaced buffer is no longer read (mark_render / sweep_done).
int render_impl(r2f_ctx *c, const void *in, int in_format, float in_gain, int H, int W, int cin, uint8_t *out_u8,
                unsigned flags, const float *noise, int noise_ch, void *ws, size_t ws_bytes, int tap_stage,
                float *tap, cudaStream_t st, Bands *bands = nullptr) {
    int rc = render_body(c, in, in_format, in_gain, H, W, cin, out_u8, flags, noise, noise_ch, ws, ws_bytes,
                         tap_stage, tap, st, bands);
    if (c && bands && rc == R2F_OK) {  // whatever was not banded: everything is final once the stream gets here
        DeviceGuard guard(c->device);
        for (; bands->recorded < bands->n; ++bands->recorded)
            if (cudaEventRecord(bands->out_done[bands->recorded], st) != cudaSuccess) {
                rc = fail(R2F_ERR_CUDA, "cudaEventRecord (band)");
                break;
            }
    }
    if (c) {
        DeviceGuard guard(c->device);
        // a render that is being captured into a CUDA graph is marked when the graph is replayed (r2f_stream_mark)
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(st, &cap) != cudaSuccess) (void)cudaGetLastError();
        if (cap == cudaStreamCaptureStatusNone) {
            cudaError_t e = mark_render(c, st);
            if (e != cudaSuccess && rc == R2F_OK) return fail_cuda(e, "mark_render");
        }
    }
    return rc;
}

}  // namespace

extern "C" {

int r2f_abi_version(void) { return R2F_ABI_VERSION; }

const char *r2f_last_error(void) { return g_err.c_str(); }

int r2f_create(int device, r2f_ctx **out) {
    if (!out) return fail(R2F_ERR_INVALID, "null out pointer");
    int count = 0;
    CU(cudaGetDeviceCount(&count));
    if (device < 0 || device >= count) return fail(R2F_ERR_INVALID, "no such CUDA device");
    DeviceGuard guard(device);
    cudaDeviceProp prop{};
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail(R2F_ERR_INVALID, "libr2f_b200 is built for sm_100a (B200) only");
    r2f_ctx *c = new (std::nothrow) r2f_ctx();
    if (!c) return fail(R2F_ERR_NOMEM, "out of host memory");
    c->device = device;
    c->num_sms = prop.multiProcessorCount;
    c->slots[0].reset(new TableSlot());
    c->t = c->slots[0].get();
    *out = c;
    return R2F_OK;
}

namespace {
void release_kernel_set(KernelSet &k) {
    k.buf.release();
    k.symbuf.release();
    k.base.release();
}
void retire_slot(r2f_ctx *c, TableSlot &t) {
    for (DevBuf *b : {&t.lut255, &t.fseg, &t.gfseg, &t.lut2d, &t.lut2d4, &t.curve, &t.curve_xp, &t.lut3d, &t.gcurve, &t.gcurve_xp, &t.hal.buf, &t.hal.symbuf,
                      &t.hal.base, &t.mtf.buf, &t.mtf.symbuf, &t.mtf.base, &t.grain.buf, &t.grain.symbuf,
                      &t.grain.base})
        retire(c, *b);
}
}  // namespace

int r2f_destroy(r2f_ctx *c) {
    if (!c) return R2F_OK;
    DeviceGuard guard(c->device);
    cudaDeviceSynchronize();
    for (auto &sp : c->slots) {
        if (!sp) continue;
        for (DevBuf *b : {&sp->lut255, &sp->fseg, &sp->gfseg, &sp->lut2d, &sp->lut2d4, &sp->curve, &sp->curve_xp, &sp->lut3d, &sp->gcurve, &sp->gcurve_xp}) b->release();
        release_kernel_set(sp->hal);
        release_kernel_set(sp->mtf);
        release_kernel_set(sp->grain);
    }
    for (DevBuf *b : {&c->burn_buf, &c->h_in, &c->h_out, &c->h_ws, &c->h_noise, &c->khat_scratch, &c->cnr_taps,
                      &c->expo_buf, &c->hist_buf, &c->stats_buf})
        b->release();
    for (auto &e : c->khat) {
        e.buf.release();
        if (e.ready) cudaEventDestroy(e.ready);
    }
    for (auto &kv : c->resize_tabs) {
        kv.second.a.release();
        kv.second.b.release();
        kv.second.c.release();
    }
    for (auto &r : c->retired) cudaFree(r.p);
    for (auto &pb : c->pool) cudaFree(pb.first);
    for (auto &ev : c->render_done)
        if (ev) cudaEventDestroy(ev);
    for (auto &kv : c->fft_lines) {
        kv.second->roots.release();
        kv.second->roots_ip.release();
        kv.second->perm.release();
        kv.second->cosines.release();
    }
    if (c->stage_host) cudaFreeHost(c->stage_host);
    if (c->upload_stream) cudaStreamDestroy(c->upload_stream);
    if (c->host_stream) cudaStreamDestroy(c->host_stream);
    delete c;
    return R2F_OK;
}

int r2f_select_slot(r2f_ctx *c, int slot) {
    if (!c || slot < 0 || slot >= R2F_MAX_SLOTS) return fail(R2F_ERR_INVALID, "r2f_select_slot: slot out of range");
    if (!c->slots[slot]) c->slots[slot].reset(new (std::nothrow) TableSlot());
    if (!c->slots[slot]) return fail(R2F_ERR_NOMEM, "out of host memory");
    c->t = c->slots[slot].get();
    c->cur_slot = slot;
    return R2F_OK;
}

int r2f_clear_slot(r2f_ctx *c, int slot) {
    if (!c || slot < 0 || slot >= R2F_MAX_SLOTS) return fail(R2F_ERR_INVALID, "r2f_clear_slot: slot out of range");
    if (!c->slots[slot]) return R2F_OK;
    DeviceGuard guard(c->device);
    retire_slot(c, *c->slots[slot]);  // renders in flight keep their tables until they finish
    *c->slots[slot] = TableSlot();
    return R2F_OK;
}

int r2f_set_lut2d(r2f_ctx *c, const float *lut, int n) {
    if (!c || !lut || n < 2) return fail(R2F_ERR_INVALID, "r2f_set_lut2d: bad arguments");
    DeviceGuard guard(c->device);
    // vertices padded to float4: one 128-bit read per vertex (three per pixel instead of nine scalar reads)
    const size_t verts = (size_t)n * n;
    std::vector<float> padded(verts * 4);
    for (size_t v = 0; v < verts; ++v) {
        padded[4 * v + 0] = lut[3 * v + 0];
        padded[4 * v + 1] = lut[3 * v + 1];
        padded[4 * v + 2] = lut[3 * v + 2];
        padded[4 * v + 3] = 0.0f;
    }
    int rc = upload(c, c->t->lut2d4, padded.data(), padded.size() * sizeof(float));
    if (rc == R2F_OK) rc = upload(c, c->t->lut2d, lut, verts * 3 * sizeof(float));
    if (rc == R2F_OK) c->t->n2 = n;
    return rc;
}

int r2f_set_curve1d(r2f_ctx *c, const float *curve, int N, float log_eps) {
    if (!c || !curve || N < 2) return fail(R2F_ERR_INVALID, "r2f_set_curve1d: bad arguments");
    DeviceGuard guard(c->device);
    TableSlot *t = c->t;
    int rc = upload_curve(c, t->curve, t->curve_xp, curve, N);
    if (rc != R2F_OK) return rc;
    t->n1 = N;
    t->x0 = curve[0];
    t->inv_range = inv_range_of(curve[0], curve[N - 1]);
    t->eps = log_eps;
    t->curve_host.assign(curve, curve + (size_t)4 * N);
    t->fast_valid = false;
    return R2F_OK;
}

int r2f_set_lut3d(r2f_ctx *c, const float *lut, int n, double scale) {
    if (!c || !lut || n < 2) return fail(R2F_ERR_INVALID, "r2f_set_lut3d: bad arguments");
    DeviceGuard guard(c->device);
    TableSlot *t = c->t;
    const size_t verts = (size_t)n * n * n;
    std::vector<float> padded(verts * 4);
    for (size_t v = 0; v < verts; ++v) {
        padded[4 * v + 0] = lut[3 * v + 0];
        padded[4 * v + 1] = lut[3 * v + 1];
        padded[4 * v + 2] = lut[3 * v + 2];
        padded[4 * v + 3] = 0.0f;
    }
    int rc = upload(c, t->lut3d, padded.data(), padded.size() * sizeof(float));
    if (rc != R2F_OK) return rc;
    // fast chain: a copy pre-multiplied by 255 and the table's largest steps between lattice neighbours
    t->lut_min = 1e300;
    t->lut_max = -1e300;
    for (int k = 0; k < 3; ++k)
        for (int a = 0; a < 3; ++a) t->lut_lip[k][a] = 0.0;
    const size_t strides[3] = {(size_t)n * n, (size_t)n, 1};
    for (size_t v = 0; v < verts; ++v) {
        const size_t idx[3] = {v / strides[0], (v / strides[1]) % n, v % n};
        for (int k = 0; k < 3; ++k) {
            const double x = lut[3 * v + k];
            t->lut_min = std::fmin(t->lut_min, x);
            t->lut_max = std::fmax(t->lut_max, x);
            if (!(x == x)) t->lut_min = -1e300;  // NaN disqualifies the fast chain
            for (int a = 0; a < 3; ++a)
                if (idx[a] + 1 < (size_t)n)
                    t->lut_lip[k][a] = std::fmax(t->lut_lip[k][a], std::fabs((double)lut[3 * (v + strides[a]) + k] - x));
            padded[4 * v + k] = (float)(x * 255.0 - 0.5);
        }
    }
    // Guard cells in front: a lattice coordinate within rounding distance below 0 (a density of exactly 0 after the
    // fast chain's roundings) selects cell -1 with a fraction of ~1; its vertices get a weight of ~1e-7, they only
    // have to be readable.  n*n + n + 1 vertices cover index -1 on all three axes.
    const size_t nguard = (size_t)n * n + n + 1;
    {
        std::vector<float> with_guard((nguard + verts) * 4, -0.5f);
        std::memcpy(with_guard.data() + nguard * 4, padded.data(), padded.size() * sizeof(float));
        rc = upload(c, t->lut255, with_guard.data(), with_guard.size() * sizeof(float));
    }
    if (rc != R2F_OK) return rc;
    t->fast_valid = false;
    t->n3 = n;
    t->s3 = scale * (double)(n - 1);  // utils.py:258
    // Error bound of the float32 fast path (device_math.cuh tetra_quant_u8), in units of the
    // quantised output: three fused roundings of partial sums bounded by 1 + 2*range, the final
    // binary32 rounding of the exact path, the float32 product with 255, and -- when s is not a
    // power of two -- the rounding of the coordinate v = d*s propagated through the edge slopes.
    double absmax = 0.0;
    for (size_t i = 0; i < verts * 3; ++i) absmax = std::fmax(absmax, std::fabs((double)lut[i]));
    const double u = std::ldexp(1.0, -24);
    // d1 >= d2 >= d3 in [0,1] make every partial sum a convex combination of vertex values, so each of
    // the three fused roundings (and the exact path's final rounding) is at most u * absmax
    double err = 3.0 * u * absmax + u * absmax + 1e-30;
    int e2 = 0;
    t->lut_absmax = absmax;
    const bool pow2 = std::frexp(t->s3, &e2) == 0.5 && t->s3 > 0.0;
    if (!pow2) err += 3.0 * (2.0 * absmax) * (2.0 * u * (double)n);  // coordinate rounding x slopes
    double margin = 255.0 * err + 2.0 * u * 255.0 * std::fmax(1.0, absmax);
    margin *= 1.5;                                          // safety factor
    t->s3f = (float)t->s3;
    t->margin3 = (float)margin;
    t->fast3 = (margin < 0.2 && (double)t->s3f == t->s3 && std::isfinite(absmax)) ? 1 : 0;
    // FastTetra (fast_chain.cuh): the fraction is exact, so only arithmetic roundings, the clamp just below the last
    // lattice plane and -- when s is not a power of two -- the rounding of v = d * s enter the bound
    double lipmax = 0.0;
    for (int k = 0; k < 3; ++k)
        for (int a = 0; a < 3; ++a) lipmax = std::fmax(lipmax, t->lut_lip[k][a]);
    const float vtop = std::nextafterf((float)(n - 1), 0.0f);
    double mt = 255.0 * (u * absmax + 3.0 * u * lipmax) + 10.0 * u * 255.0 * absmax + u * 255.0;
    mt += 255.0 * 3.0 * lipmax * ((double)(n - 1) - (double)vtop);
    if (!pow2) mt += 255.0 * 3.0 * lipmax * (2.0 * u * (double)n);
    mt *= 1.25;
    FastTetra ft{};
    ft.lut = static_cast<const float4 *>(t->lut255.p) + nguard;
    ft.s3f = t->s3f;
    ft.vtop = vtop;
    ft.half_m = (float)(0.5 - mt);
    ft.n = n;
    ft.o111 = n * n + n + 1;
    ft.neg_k = 0u - kMagicBits * (unsigned)ft.o111;
    ft.ok = (mt < 0.2 && (double)t->s3f == t->s3 && t->lut_min >= 0.0 && t->lut_max <= 1.0 && n <= 129 &&
             std::isfinite(absmax)) ? 1 : 0;
    t->ft = ft;
    return R2F_OK;
}

int r2f_set_halation_kernel(r2f_ctx *c, const float *kernel, int k) {
    if (!c) return fail(R2F_ERR_INVALID, "null context");
    DeviceGuard guard(c->device);
    return upload_kernel(c, c->t->hal, kernel, k, 3);
}

int r2f_set_mtf_kernel(r2f_ctx *c, const float *kernel, int k) {
    if (!c) return fail(R2F_ERR_INVALID, "null context");
    DeviceGuard guard(c->device);
    return upload_kernel(c, c->t->mtf, kernel, k, 3);
}

int r2f_set_grain(r2f_ctx *c, const float *curve, int N, const float *kernel, int k, uint64_t seed) {
    if (!c || !curve || N < 2) return fail(R2F_ERR_INVALID, "r2f_set_grain: bad arguments");
    DeviceGuard guard(c->device);
    TableSlot *t = c->t;
    int rc = upload_curve(c, t->gcurve, t->gcurve_xp, curve, N);
    if (rc != R2F_OK) return rc;
    t->ng = N;
    t->gx0 = curve[0];
    t->ginv = inv_range_of(curve[0], curve[N - 1]);
    t->gfast = FastCurve{};
    if (t->gcurve_xp.p == nullptr && curve[N - 1] > curve[0] && N <= (1 << 20)) {  // uniform: conversion-free form
        std::vector<float> fs((size_t)3 * N * 2);
        for (int ch = 0; ch < 3; ++ch) {
            const float *row = curve + (size_t)(ch + 1) * N;
            for (int i = 0; i < N; ++i) {
                const double d = i + 1 < N ? (double)row[i + 1] - (double)row[i] : 0.0;
                fs[((size_t)ch * N + i) * 2] = (float)((double)row[i] + 0.5 * d);
                fs[((size_t)ch * N + i) * 2 + 1] = (float)d;
            }
        }
        rc = upload(c, t->gfseg, fs.data(), fs.size() * sizeof(float));
        if (rc != R2F_OK) return rc;
        const double R = (double)curve[N - 1] - (double)curve[0];
        t->gfast.seg = static_cast<const float2 *>(t->gfseg.p);
        t->gfast.cA = (float)(1.0 / R);
        t->gfast.cB = (float)(-(double)curve[0] / R);
        t->gfast.pscale = std::nextafterf((float)(N - 1), 0.0f);
        t->gfast.N = N;
    }
    const float one = 1.0f;  // gpu_processor.py:931-932: missing kernel -> 1x1 ones
    rc = kernel ? upload_kernel(c, t->grain, kernel, k, 1) : upload_kernel(c, t->grain, &one, 1, 1);
    if (rc != R2F_OK) return rc;
    t->seed = seed;
    return R2F_OK;
}

int r2f_set_option(r2f_ctx *c, int key, int value) {
    if (!c) return fail(R2F_ERR_INVALID, "null context");
    if (key == R2F_OPT_CONV_SYM && (value == 0 || value == 1)) {
        c->conv_sym = value;
        return R2F_OK;
    }
    if (key == R2F_OPT_CONV_PATH && value >= 0 && value <= 2) {
        c->conv_path = value;
        return R2F_OK;
    }
    if (key == R2F_OPT_FUSE_MTF && (value == 0 || value == 1)) {
        c->fuse_mtf = value;
        return R2F_OK;
    }
    if (key == R2F_OPT_FAST_CHAIN && (value == 0 || value == 1)) {
        c->fast_chain = value;
        return R2F_OK;
    }
    return fail(R2F_ERR_INVALID, "r2f_set_option: unknown key or value");
}

int r2f_set_grain_seed(r2f_ctx *c, uint64_t seed) {
    if (!c) return fail(R2F_ERR_INVALID, "null context");
    c->t->seed = seed;
    return R2F_OK;
}

int r2f_set_burn(r2f_ctx *c, float d_ref, float highlight_burn, float burn_scale) {
    if (!c || !(burn_scale > 0.f)) return fail(R2F_ERR_INVALID, "r2f_set_burn: bad arguments");
    c->t->d_ref = d_ref;
    c->t->burn_strength = highlight_burn;
    c->t->burn_scale = burn_scale;
    c->t->burn_set = true;
    return R2F_OK;
}

size_t r2f_workspace_bytes(int H, int W, unsigned flags) {
    (void)flags;  // three planar float32 working images cover every stage combination (and the taps)
    if (H < 1 || W < 1) return 0;
    return plane_stride_for(H, W) * 3 * 3 * sizeof(float);
}

int r2f_render(r2f_ctx *c, const float *in_dev, int H, int W, int in_channels, uint8_t *out_dev, unsigned flags,
               const float *noise_dev, int noise_channels, void *workspace_dev, size_t workspace_bytes, void *stream) {
    return render_impl(c, in_dev, R2F_IN_F32, 1.0f, H, W, in_channels, out_dev, flags, noise_dev, noise_channels,
                       workspace_dev, workspace_bytes, 0, nullptr, static_cast<cudaStream_t>(stream));
}

int r2f_render_ex(r2f_ctx *c, const void *in_dev, int in_format, float in_gain, int H, int W, int in_channels,
                  uint8_t *out_dev, unsigned flags, const float *noise_dev, int noise_channels, void *workspace_dev,
                  size_t workspace_bytes, void *stream) {
    return render_impl(c, in_dev, in_format, in_gain, H, W, in_channels, out_dev, flags, noise_dev, noise_channels,
                       workspace_dev, workspace_bytes, 0, nullptr, static_cast<cudaStream_t>(stream));
}

int r2f_band_row(int H, int nbands, int i) { return band_row(H, nbands < 1 ? 1 : nbands, i); }

int r2f_render_banded(r2f_ctx *c, const void *in_dev, int in_format, float in_gain, int H, int W, int in_channels,
                      uint8_t *out_dev, unsigned flags, const float *noise_dev, int noise_channels, void *workspace_dev,
                      size_t workspace_bytes, int nbands, void *const *in_ready, void *const *out_done, void *stream) {
    if (nbands < 1 || nbands > 64 || !in_ready || !out_done)
        return fail(R2F_ERR_INVALID, "r2f_render_banded: 1..64 bands with their event arrays");
    for (int i = 0; i < nbands; ++i)
        if (!in_ready[i] || !out_done[i]) return fail(R2F_ERR_INVALID, "r2f_render_banded: null event");
    Bands b;
    b.n = nbands;
    b.in_ready = reinterpret_cast<cudaEvent_t const *>(in_ready);
    b.out_done = reinterpret_cast<cudaEvent_t const *>(out_done);
    return render_impl(c, in_dev, in_format, in_gain, H, W, in_channels, out_dev, flags, noise_dev, noise_channels,
                       workspace_dev, workspace_bytes, 0, nullptr, static_cast<cudaStream_t>(stream), &b);
}

int r2f_render_tap(r2f_ctx *c, const float *in_dev, int H, int W, int in_channels, unsigned flags,
                   const float *noise_dev, int noise_channels, void *workspace_dev, size_t workspace_bytes,
                   int tap_stage, float *tap_dev, void *stream) {
    if (tap_stage == 0) return fail(R2F_ERR_INVALID, "tap_stage must be one of R2F_TAP_*");
    return render_impl(c, in_dev, R2F_IN_F32, 1.0f, H, W, in_channels, nullptr, flags, noise_dev, noise_channels,
                       workspace_dev, workspace_bytes, tap_stage, tap_dev, static_cast<cudaStream_t>(stream));
}

int r2f_render_tap_ex(r2f_ctx *c, const void *in_dev, int in_format, float in_gain, int H, int W, int in_channels,
                      unsigned flags, const float *noise_dev, int noise_channels, void *workspace_dev,
                      size_t workspace_bytes, int tap_stage, float *tap_dev, void *stream) {
    if (tap_stage == 0) return fail(R2F_ERR_INVALID, "tap_stage must be one of R2F_TAP_*");
    return render_impl(c, in_dev, in_format, in_gain, H, W, in_channels, nullptr, flags, noise_dev, noise_channels,
                       workspace_dev, workspace_bytes, tap_stage, tap_dev, static_cast<cudaStream_t>(stream));
}

int r2f_render_host(r2f_ctx *c, const float *in_host, int H, int W, int in_channels, uint8_t *out_host, unsigned flags,
                    const float *noise_host, int noise_channels) {
    if (!c || !in_host || !out_host || H < 1 || W < 1) return fail(R2F_ERR_INVALID, "r2f_render_host: bad arguments");
    if (in_channels != 3 && in_channels != 4) return fail(R2F_ERR_INVALID, "in_channels must be 3 or 4");
    DeviceGuard guard(c->device);
    if (!c->host_stream) CU(cudaStreamCreateWithFlags(&c->host_stream, cudaStreamNonBlocking));
    const size_t npix = (size_t)H * W;
    const size_t in_bytes = npix * in_channels * sizeof(float), out_bytes = npix * 3;
    CU(c->h_in.ensure(in_bytes));
    CU(c->h_out.ensure(out_bytes));
    const size_t ws_bytes = r2f_workspace_bytes(H, W, flags);
    if (flags & (R2F_HALATION | R2F_MTF | R2F_GRAIN | R2F_BURN)) CU(c->h_ws.ensure(ws_bytes));
    CU(cudaMemcpyAsync(c->h_in.p, in_host, in_bytes, cudaMemcpyHostToDevice, c->host_stream));
    const float *noise_dev = nullptr;
    if (noise_host && (flags & R2F_GRAIN)) {
        const size_t nb = npix * noise_channels * sizeof(float);
        CU(c->h_noise.ensure(nb));
        CU(cudaMemcpyAsync(c->h_noise.p, noise_host, nb, cudaMemcpyHostToDevice, c->host_stream));
        noise_dev = static_cast<const float *>(c->h_noise.p);
    }
    int rc = render_impl(c, c->h_in.p, R2F_IN_F32, 1.0f, H, W, in_channels, static_cast<uint8_t *>(c->h_out.p), flags,
                         noise_dev, noise_channels, c->h_ws.p, c->h_ws.bytes, 0, nullptr, c->host_stream);
    if (rc != R2F_OK) return rc;
    CU(cudaMemcpyAsync(out_host, c->h_out.p, out_bytes, cudaMemcpyDeviceToHost, c->host_stream));
    CU(cudaStreamSynchronize(c->host_stream));
    return R2F_OK;
}

int r2f_convolve2d(r2f_ctx *c, const float *in_dev, float *out_dev, int H, int W, const float *kernel, int k,
                   void *workspace_dev, size_t workspace_bytes, void *stream) {
    if (!c || !in_dev || !out_dev || H < 1 || W < 1) return fail(R2F_ERR_INVALID, "r2f_convolve2d: bad arguments");
    DeviceGuard guard(c->device);
    const size_t ps = plane_stride_for(H, W);
    if (!workspace_dev || workspace_bytes < ps * 6 * sizeof(float))
        return fail(R2F_ERR_NOMEM, "workspace too small (see r2f_workspace_bytes)");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    KernelSet ks;
    CU(cudaStreamSynchronize(st));
    int rc = upload_kernel(c, ks, kernel, k, 3);
    if (rc != R2F_OK) return rc;
    Planes a{static_cast<float *>(workspace_dev), ps}, b{static_cast<float *>(workspace_dev) + 3 * ps, ps};
    const size_t npix = (size_t)H * W;
    cudaError_t e = launch_interleaved_to_planar(in_dev, 3, 3, a, npix, c->num_sms, st);
    FftGeometry geo;
    // (three filtered layers need the render path's second pass: the stage entry point keeps them on the direct kernel)
    const bool use_fft = want_fft(c, ks, H, W, geo) && !ks.fft_third && workspace_bytes >= ps * 9 * sizeof(float);
    if (c->conv_path == 2 && !use_fft) {
        retire(c, ks.buf);
        retire(c, ks.base);
        retire(c, ks.symbuf);
        return fail(R2F_ERR_INVALID, "FFT path forced but this kernel/frame/workspace is not eligible");
    }
    if (e == cudaSuccess && use_fft) {
        FftConvArgs fa{};
        rc = fft_prepare(c, ks, H, W, geo, fa, st);
        if (rc == R2F_OK) {
            fa.S = reinterpret_cast<float2 *>(static_cast<float *>(workspace_dev) + 6 * ps);
            fa.src_planar = a.base;
            fa.dst_planar = b.base;
            e = launch_fft_conv(fa, 0, false, st);
        }
    } else if (e == cudaSuccess) {
        e = conv_dispatch(c, conv_args(ks, a.base, b.base, ps, H, W), st);
    }
    if (e == cudaSuccess) e = launch_planar_to_interleaved(b, out_dev, npix, c->num_sms, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e == cudaSuccess) {  // the temporary kernel's buffers are idle again: straight back to the pool
        for (DevBuf *b : {&ks.buf, &ks.base, &ks.symbuf})
            if (b->p) {
                c->pool.emplace_back(b->p, b->bytes);
                b->p = nullptr;
                b->bytes = 0;
            }
    } else {
        ks.buf.release();
        ks.base.release();
        ks.symbuf.release();
    }
    if (rc != R2F_OK) return rc;
    if (e != cudaSuccess) return fail_cuda(e, "r2f_convolve2d");
    c->launches += 3;
    return R2F_OK;
}

int r2f_generate_noise(r2f_ctx *c, float *out_dev, int H, int W, int channels, uint64_t seed, void *stream) {
    if (!c || !out_dev || H < 1 || W < 1 || (channels != 1 && channels != 3))
        return fail(R2F_ERR_INVALID, "r2f_generate_noise: bad arguments");
    DeviceGuard guard(c->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t ps = plane_stride_for(H, W), npix = (size_t)H * W;
    CU(c->h_noise.ensure(ps * 3 * sizeof(float)));
    Planes p{static_cast<float *>(c->h_noise.p), ps};
    CU(launch_noise(p, channels, H, W, seed, 0, c->num_sms, st));
    if (channels == 3) {
        CU(launch_planar_to_interleaved(p, out_dev, npix, c->num_sms, st));
    } else {
        CU(cudaMemcpyAsync(out_dev, p.base, npix * sizeof(float), cudaMemcpyDeviceToDevice, st));
    }
    c->launches += 2;
    return R2F_OK;
}

int r2f_chroma_nr(r2f_ctx *c, const float *in_dev, int in_channels, float *out_dev, int H, int W, const float *taps,
                  int ntaps, void *workspace_dev, size_t workspace_bytes, void *stream) {
    if (!c || !in_dev || !out_dev || !taps || H < 1 || W < 1 || ntaps < 1 || (ntaps & 1) == 0 ||
        (in_channels != 3 && in_channels != 4) || ntaps > 49)
        return fail(R2F_ERR_INVALID, "r2f_chroma_nr: bad arguments (odd tap count up to 49)");
    DeviceGuard guard(c->device);
    const size_t ps = plane_stride_for(H, W);
    if (!workspace_dev || workspace_bytes < ps * 6 * sizeof(float))
        return fail(R2F_ERR_NOMEM, "workspace too small (see r2f_workspace_bytes)");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CU(launch_chroma_nr(in_dev, in_channels, out_dev, H, W, taps, ntaps, static_cast<float *>(workspace_dev), ps,
                        c->num_sms, st));   // the taps travel as a kernel parameter
    c->launches += 2;
    return R2F_OK;
}

int r2f_histogram(r2f_ctx *c, const uint8_t *img_dev, int H, int W, uint32_t *counts_dev, void *stream) {
    if (!c || !img_dev || !counts_dev || H < 1 || W < 1) return fail(R2F_ERR_INVALID, "r2f_histogram: bad arguments");
    if ((reinterpret_cast<uintptr_t>(img_dev) & 3) != 0) return fail(R2F_ERR_INVALID, "image must be 4-byte aligned");
    DeviceGuard guard(c->device);
    CU(launch_histogram(img_dev, (size_t)H * W, counts_dev, c->num_sms, static_cast<cudaStream_t>(stream)));
    c->launches += 1;
    return R2F_OK;
}

int r2f_histogram_image(r2f_ctx *c, const uint8_t *img_dev, int H, int W, const uint8_t *mix_table, int height,
                        uint8_t *out_dev, void *stream) {
    if (!c || !img_dev || !mix_table || !out_dev || H < 1 || W < 1 || height < 1)
        return fail(R2F_ERR_INVALID, "r2f_histogram_image: bad arguments");
    if ((reinterpret_cast<uintptr_t>(img_dev) & 3) != 0 || (reinterpret_cast<uintptr_t>(out_dev) & 3) != 0)
        return fail(R2F_ERR_INVALID, "image and output must be 4-byte aligned");
    DeviceGuard guard(c->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CU(c->hist_buf.ensure(768 * sizeof(unsigned int)));
    unsigned int *counts = static_cast<unsigned int *>(c->hist_buf.p);
    CU(launch_histogram(img_dev, (size_t)H * W, counts, c->num_sms, st));
    CU(launch_histogram_image(counts, height, mix_table, out_dev, st));
    c->launches += 2;
    return R2F_OK;
}

int r2f_canvas_paste(r2f_ctx *c, const uint8_t *src_dev, int H, int W, uint8_t *dst_dev, int canvas_h, int canvas_w,
                     int off_y, int off_x, int r, int g, int b, void *stream) {
    if (!c || !src_dev || !dst_dev || H < 1 || W < 1 || canvas_h < 1 || canvas_w < 1)
        return fail(R2F_ERR_INVALID, "r2f_canvas_paste: bad arguments");
    DeviceGuard guard(c->device);
    CU(launch_canvas_paste(src_dev, H, W, dst_dev, canvas_h, canvas_w, off_y, off_x, r, g, b, c->num_sms,
                           static_cast<cudaStream_t>(stream)));
    c->launches += 1;
    return R2F_OK;
}

int r2f_present(r2f_ctx *c, const uint8_t *src_dev, int H, int W, uint8_t *dst_rgba_dev, int dst_h, int dst_w,
                const float *transform, int r, int g, int b, void *stream) {
    if (!c || !src_dev || !dst_rgba_dev || !transform || H < 1 || W < 1 || dst_h < 1 || dst_w < 1)
        return fail(R2F_ERR_INVALID, "r2f_present: bad arguments");
    if ((reinterpret_cast<uintptr_t>(dst_rgba_dev) & 3) != 0) return fail(R2F_ERR_INVALID, "destination must be 4-byte aligned");
    DeviceGuard guard(c->device);
    PresentArgs u{transform[0], transform[1], transform[2], transform[3], transform[4], transform[5], transform[6],
                  transform[7], r & 255, g & 255, b & 255};
    CU(launch_present(src_dev, H, W, dst_rgba_dev, dst_h, dst_w, u, c->num_sms, static_cast<cudaStream_t>(stream)));
    c->launches += 1;
    return R2F_OK;
}

int r2f_calc_exposure(r2f_ctx *c, const void *in_dev, int in_format, int H, int W, int in_channels, double factor,
                      double *mean_out, void *stream) {
    if (!c || !in_dev || !mean_out || H < 1 || W < 1 || (in_channels != 3 && in_channels != 4) || !(factor > 0.0))
        return fail(R2F_ERR_INVALID, "r2f_calc_exposure: bad arguments");
    if (in_format != R2F_IN_F32 && in_format != R2F_IN_U16) return fail(R2F_ERR_INVALID, "unknown input format");
    DeviceGuard guard(c->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int fmt = (in_format == R2F_IN_U16 ? 2 : 0) + (in_channels == 4 ? 1 : 0);
    const int nblocks = c->num_sms * 8;
    CU(c->expo_buf.ensure((size_t)(nblocks + 1) * sizeof(double)));
    double *partial = static_cast<double *>(c->expo_buf.p), *out = partial + nblocks;
    CU(launch_exposure_mean(in_dev, fmt, H, W, 1.0 / factor, partial, nblocks, out, st));
    c->launches += 2;
    CU(cudaMemcpyAsync(mean_out, out, sizeof(double), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return R2F_OK;
}

namespace {
// tap tables of one axis on the device, cached per (kind, source size, destination size)
int resize_axis_tab(r2f_ctx *c, int kind, int ssize, int dsize, r2f_ctx::ResizeTab **out) {
    auto key = std::make_tuple(kind, ssize, dsize);
    auto it = c->resize_tabs.find(key);
    if (it == c->resize_tabs.end()) {
        if (c->resize_tabs.size() >= 16) {  // drop the least recently used entry (its buffers may still be read)
            auto lru = c->resize_tabs.begin();
            for (auto jt = c->resize_tabs.begin(); jt != c->resize_tabs.end(); ++jt)
                if (jt->second.last_use < lru->second.last_use) lru = jt;
            retire(c, lru->second.a);
            retire(c, lru->second.b);
            retire(c, lru->second.c);
            c->resize_tabs.erase(lru);
        }
        r2f_ctx::ResizeTab t;
        int rc;
        if (kind == 0) {
            AreaTabHost h;
            resize_area_tab(ssize, dsize, h);
            if ((rc = upload(c, t.a, h.start.data(), h.start.size() * sizeof(int))) != R2F_OK) return rc;
            if ((rc = upload(c, t.b, h.si.data(), h.si.size() * sizeof(int))) != R2F_OK) return rc;
            if ((rc = upload(c, t.c, h.alpha.data(), h.alpha.size() * sizeof(float))) != R2F_OK) return rc;
        } else {
            LanczosTabHost h;
            resize_lanczos4_tab(ssize, dsize, h);
            if ((rc = upload(c, t.a, h.ofs.data(), h.ofs.size() * sizeof(int))) != R2F_OK) return rc;
            if ((rc = upload(c, t.b, h.coef.data(), h.coef.size() * sizeof(float))) != R2F_OK) return rc;
            if ((rc = upload(c, t.c, h.icoef.data(), h.icoef.size() * sizeof(int))) != R2F_OK) return rc;
        }
        it = c->resize_tabs.emplace(key, t).first;
    }
    it->second.last_use = ++c->resize_clock;
    *out = &it->second;
    return R2F_OK;
}
}  // namespace

int r2f_resize(r2f_ctx *c, const void *in_dev, int pix_format, int H, int W, int in_channels, void *out_dev, int out_h,
               int out_w, int interpolation, void *stream) {
    if (!c || !in_dev || !out_dev || H < 1 || W < 1 || out_h < 1 || out_w < 1)
        return fail(R2F_ERR_INVALID, "r2f_resize: bad arguments");
    if (pix_format != R2F_PIX_F32 && pix_format != R2F_PIX_U8) return fail(R2F_ERR_INVALID, "r2f_resize: unknown pixel format");
    if (in_channels != 3 && !(in_channels == 4 && pix_format == R2F_PIX_F32))
        return fail(R2F_ERR_INVALID, "r2f_resize: in_channels must be 3 (or 4 for float32 frames)");
    if (interpolation != R2F_INTER_AREA && interpolation != R2F_INTER_LANCZOS4)
        return fail(R2F_ERR_INVALID, "r2f_resize: unknown interpolation");
    DeviceGuard guard(c->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool u8 = pix_format == R2F_PIX_U8;
    if (interpolation == R2F_INTER_AREA) {
        if (out_h > H || out_w > W) return fail(R2F_ERR_INVALID, "r2f_resize: INTER_AREA is the shrinking filter");
        int ix = 1, iy = 1;
        if (resize_area_is_fast(W, out_w, ix) && resize_area_is_fast(H, out_h, iy)) {
            CU(launch_resize_area_int(in_dev, u8, in_channels, H, W, out_dev, out_h, out_w, ix, iy, c->num_sms, st));
        } else {
            r2f_ctx::ResizeTab *tx = nullptr, *ty = nullptr;
            int rc = resize_axis_tab(c, 0, W, out_w, &tx);
            if (rc == R2F_OK) rc = resize_axis_tab(c, 0, H, out_h, &ty);
            if (rc != R2F_OK) return rc;
            AreaTabDev xt{static_cast<const int *>(tx->a.p), static_cast<const int *>(tx->b.p), static_cast<const float *>(tx->c.p)};
            AreaTabDev yt{static_cast<const int *>(ty->a.p), static_cast<const int *>(ty->b.p), static_cast<const float *>(ty->c.p)};
            CU(launch_resize_area(in_dev, u8, in_channels, H, W, out_dev, out_h, out_w, xt, yt, c->num_sms, st));
        }
    } else {
        r2f_ctx::ResizeTab *tx = nullptr, *ty = nullptr;
        int rc = resize_axis_tab(c, 1, W, out_w, &tx);
        if (rc == R2F_OK) rc = resize_axis_tab(c, 1, H, out_h, &ty);
        if (rc != R2F_OK) return rc;
        LanczosTabDev xt{static_cast<const int *>(tx->a.p), static_cast<const float *>(tx->b.p), static_cast<const int *>(tx->c.p)};
        LanczosTabDev yt{static_cast<const int *>(ty->a.p), static_cast<const float *>(ty->b.p), static_cast<const int *>(ty->c.p)};
        CU(launch_resize_lanczos4(in_dev, u8, in_channels, H, W, out_dev, out_h, out_w, xt, yt, c->num_sms, st));
    }
    c->launches += 1;
    CU(mark_render(c, st));  // the tap tables are copy-on-write storage like the LUTs
    return R2F_OK;
}

uint64_t r2f_launch_count(const r2f_ctx *c) { return c ? c->launches : 0; }

int r2f_stream_mark(r2f_ctx *c, void *stream) {
    if (!c) return fail(R2F_ERR_INVALID, "null context");
    DeviceGuard guard(c->device);
    CU(mark_render(c, static_cast<cudaStream_t>(stream)));
    return R2F_OK;
}

int r2f_fast_chain_stats(r2f_ctx *c, uint64_t *deferred_pixels, float *margin) {
    if (!c) return fail(R2F_ERR_INVALID, "null context");
    DeviceGuard guard(c->device);
    unsigned long long n = 0;
    if (c->stats_buf.p) {
        CU(cudaDeviceSynchronize());
        CU(cudaMemcpy(&n, c->stats_buf.p, sizeof(n), cudaMemcpyDeviceToHost));
        CU(cudaMemset(c->stats_buf.p, 0, sizeof(n)));
    }
    if (deferred_pixels) *deferred_pixels = (uint64_t)n;
    if (margin) *margin = (c->t->fast_valid && c->t->fast.ok) ? c->t->fast.margin : -1.0f;
    return R2F_OK;
}

int r2f_profile_enable(r2f_ctx *c, int on) {
    if (!c) return fail(R2F_ERR_INVALID, "null context");
    c->profiling = on != 0;
    return R2F_OK;
}

int r2f_profile_read(r2f_ctx *c, double *ms_accum, uint64_t *count_accum) {
    if (!c || !ms_accum || !count_accum) return fail(R2F_ERR_INVALID, "r2f_profile_read: bad arguments");
    DeviceGuard guard(c->device);
    int rc = R2F_OK;
    for (auto &r : c->prof) {
        float ms = 0.f;
        cudaError_t e = cudaEventSynchronize(r.b);
        if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, r.a, r.b);
        if (e == cudaSuccess && r.id >= 0 && r.id < R2F_PROF_COUNT) {
            ms_accum[r.id] += (double)ms;
            count_accum[r.id] += 1;
        } else if (e != cudaSuccess) {
            rc = fail_cuda(e, "r2f_profile_read");
        }
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    c->prof.clear();
    return rc;
}

}  // extern "C"
