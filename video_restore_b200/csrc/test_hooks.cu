// Kernel-level test and micro-benchmark entry points of the C ABI (vr_conv3x3_test, vr_conv3x3_bench).
// They exercise exactly the kernels the product path launches; there is no alternative implementation here.
#include "../../include/vrb200.h"
#include "conv3x3_sm100.cuh"
#include "vr_common.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

using namespace vr;

extern "C" const char* vr_global_error(void) { return global_error().c_str(); }
static long long g_last_conv_cycles = 0;
extern "C" int64_t vr_last_conv_cycles(void) { return g_last_conv_cycles; }

namespace {
struct ScopedDev {
    Device dev;
    std::string err;
    bool ok = false;
    explicit ScopedDev(int ordinal) {
        dev.err = &err;
        dev.ordinal = ordinal;
        if (cudaSetDevice(ordinal) != cudaSuccess) {
            set_error(&err, "cudaSetDevice failed (no usable CUDA device; this library has no CPU path)");
            return;
        }
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, ordinal) != cudaSuccess) {
            set_error(&err, "cudaGetDeviceProperties failed");
            return;
        }
        if (p.major != 10) {
            set_error(&err, "device is not sm_100 class (found sm_" + std::to_string(p.major * 10 + p.minor) + ")");
            return;
        }
        dev.sm_count = p.multiProcessorCount;
        read_conv_env(dev);
        if (cudaStreamCreateWithFlags(&dev.stream, cudaStreamNonBlocking) != cudaSuccess) {
            set_error(&err, "cudaStreamCreate failed");
            return;
        }
        ok = true;
    }
    ~ScopedDev() {
        if (dev.bil_tab) cudaFree(dev.bil_tab);
        if (dev.stream) cudaStreamDestroy(dev.stream);
    }
};

// [px][c] fp32 -> fp16 [px][cpad], or chunk-planar [cpad/32][px][32] (FLAG_PLANAR)
std::vector<__half> to_half_padded(const float* x, size_t px, int c, int cpad, bool planar = false) {
    std::vector<__half> h(px * cpad, __float2half(0.f));
    for (size_t p = 0; p < px; ++p)
        for (int i = 0; i < c; ++i)
            h[planar ? (static_cast<size_t>(i >> 5) * px + p) * 32 + (i & 31) : p * cpad + i] = __float2half_rn(x[p * c + i]);
    return h;
}
}  // namespace

extern "C" int vr_conv3x3_test(vr_conv_test* t) {
    if (!t || !t->x || !t->weight || !t->y || t->H <= 0 || t->W <= 0) {
        set_error(nullptr, "vr_conv3x3_test: bad arguments");
        return VR_E_INVALID;
    }
    ScopedDev sd(t->device);
    if (!sd.ok) return VR_E_NODEVICE;
    Device& dev = sd.dev;
    const size_t px = static_cast<size_t>(t->H) * t->W;
    const int cin_pad = (t->cin + 31) / 32 * 32;  // multiple of 32 covers both chunk widths
    const int cout = t->cout;
    const bool rgb4 = (cout == 3);
    const bool ps4 = (cout == 48);
    const int out_c = rgb4 ? 4 : cout;

    ConvWeights w;
    int rc = pack_conv_weights(dev, t->weight, t->bias, t->act == ACT_PRELU ? t->prelu : nullptr, t->cin, cout, &w);
    if (rc) return rc;

    // FLAG_PLANAR: source, residuals and (NHWC, cout % 32 == 0) output as chunk-planar tensors, as the network uses them
    const bool planar = (t->flags & FLAG_PLANAR) != 0;
    const bool out_planar = planar && !rgb4 && !ps4 && cout % 32 == 0;
    std::vector<__half> hx = to_half_padded(t->x, px, t->cin, cin_pad, planar);
    __half *dx = nullptr, *dy = nullptr, *dr1 = nullptr, *dr2 = nullptr;
    const size_t out_elems = ps4 ? px * 16 * 4 : px * out_c;
    VR_CUDA_CHECK(cudaMalloc(&dx, hx.size() * sizeof(__half)), dev.err);
    // guard bands of 0xA5 before and after the output: the epilogue's store masks are checked, not trusted
    // (compute-sanitizer is not available on this pool)
    constexpr size_t kGuard = 8192;
    uint8_t* dy_raw = nullptr;
    VR_CUDA_CHECK(cudaMalloc(&dy_raw, out_elems * sizeof(__half) + 2 * kGuard), dev.err);
    VR_CUDA_CHECK(cudaMemset(dy_raw, 0xA5, out_elems * sizeof(__half) + 2 * kGuard), dev.err);
    dy = reinterpret_cast<__half*>(dy_raw + kGuard);
    VR_CUDA_CHECK(cudaMemset(dy, 0, out_elems * sizeof(__half)), dev.err);
    VR_CUDA_CHECK(cudaMemcpy(dx, hx.data(), hx.size() * sizeof(__half), cudaMemcpyHostToDevice), dev.err);
    if (t->res1) {
        std::vector<__half> h = to_half_padded(t->res1, px, cout, cout, out_planar);
        VR_CUDA_CHECK(cudaMalloc(&dr1, h.size() * sizeof(__half)), dev.err);
        VR_CUDA_CHECK(cudaMemcpy(dr1, h.data(), h.size() * sizeof(__half), cudaMemcpyHostToDevice), dev.err);
    }
    if (t->res2) {
        std::vector<__half> h = to_half_padded(t->res2, px, cout, cout, out_planar);
        VR_CUDA_CHECK(cudaMalloc(&dr2, h.size() * sizeof(__half)), dev.err);
        VR_CUDA_CHECK(cudaMemcpy(dr2, h.data(), h.size() * sizeof(__half), cudaMemcpyHostToDevice), dev.err);
    }

    // the uploads above ran on the legacy stream from pageable memory (the DMA may still be in flight when cudaMemcpy
    // returns) and the conv runs on a non-blocking stream: order them explicitly
    VR_CUDA_CHECK(cudaDeviceSynchronize(), dev.err);

    ConvCall c;
    c.in = dx;
    c.in_cstride = planar ? 32 : cin_pad;
    c.in_planes = planar ? cin_pad / 32 : 1;
    c.in_pstride = static_cast<long long>(px) * 32;
    c.H = t->H;
    c.W = t->W;
    c.w = &w;
    c.act = t->act;
    c.slope = t->slope;
    c.out = dy;
    c.out_cstride = out_planar ? 32 : out_c;
    c.out_pstride = static_cast<long long>(px) * 32;
    c.out_coff = 0;
    c.res1 = dr1;
    c.res1_cstride = out_planar ? 32 : cout;
    c.res1_pstride = static_cast<long long>(px) * 32;
    c.s1 = t->s1;
    c.res2 = dr2;
    c.res2_cstride = out_planar ? 32 : cout;
    c.res2_pstride = static_cast<long long>(px) * 32;
    c.s2 = t->s2;
    c.out_mode = rgb4 ? OUT_RGB4 : (ps4 ? OUT_PS4 : OUT_NHWC);
    c.base = dx;
    c.base_cstride = planar ? 32 : cin_pad;
    c.rows = t->rows;
    c.flags = t->flags;

    const int iters = t->iters > 0 ? t->iters : 1;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    rc = run_conv(dev, c);  // warm-up and the run whose result is returned when iters == 1
    if (rc == 0) {
        cudaEventRecord(e0, dev.stream);
        for (int i = 1; i < iters && rc == 0; ++i) rc = run_conv(dev, c);
        cudaEventRecord(e1, dev.stream);
    }
    cudaError_t se = cudaStreamSynchronize(dev.stream);
    if (rc == 0 && se != cudaSuccess) {
        set_error(dev.err, std::string("conv kernel failed: ") + cudaGetErrorString(se));
        rc = VR_E_CUDA;
    }
    if (rc == 0) {
        float ms = 0.f;
        if (iters > 1) {
            cudaEventElapsedTime(&ms, e0, e1);
            ms /= (iters - 1);
        }
        t->ms = ms;
        std::vector<__half> hy(out_elems);
        cudaMemcpy(hy.data(), dy, out_elems * sizeof(__half), cudaMemcpyDeviceToHost);
        std::vector<uint8_t> g0(kGuard), g1(kGuard);
        cudaMemcpy(g0.data(), dy_raw, kGuard, cudaMemcpyDeviceToHost);
        cudaMemcpy(g1.data(), dy_raw + kGuard + out_elems * sizeof(__half), kGuard, cudaMemcpyDeviceToHost);
        for (size_t i = 0; i < kGuard; ++i)
            if (g0[i] != 0xA5 || g1[i] != 0xA5) {
                set_error(dev.err, "conv kernel wrote outside its output tensor (guard band damaged)");
                rc = VR_E_CUDA;
                break;
            }
        if (ps4) {
            // [4H][4W][4] fp16 -> y[4H][4W][3] fp32
            for (size_t i = 0; i < px * 16; ++i)
                for (int ch = 0; ch < 3; ++ch) t->y[i * 3 + ch] = __half2float(hy[i * 4 + ch]);
        } else {
            for (size_t p = 0; p < px; ++p)
                for (int ch = 0; ch < cout; ++ch)
                    t->y[p * cout + ch] =
                        __half2float(hy[out_planar ? (static_cast<size_t>(ch >> 5) * px + p) * 32 + (ch & 31) : p * out_c + ch]);
        }
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(dx);
    cudaFree(dy_raw);
    if (dr1) cudaFree(dr1);
    if (dr2) cudaFree(dr2);
    free_conv_weights(&w);
    return rc;
}

static long long g_pair2_prof[64] = {0};
// wait-cycle profile of cluster 0's leader CTA in the last vr_conv_pair2_test launch (dbg_cycles[300..364), see the kernel)
extern "C" void vr_pair2_profile(int64_t* out64) {
    for (int i = 0; i < 64; ++i) out64[i] = g_pair2_prof[i];
}

// K4 hook: two consecutive dense-block layers in one launch. x [H][W][cin] (cin % 32 == 0), layer A cin -> 32, layer B
// (cin + 32) -> 32, both + bias + LeakyReLU(slope); yA / yB [H][W][32]. Tensors are the network's chunk-planar dense-block
// buffer (cin / 32 + 2 planes). iters > 1: average ms per launch in *ms. flags: VR_MAX_CTAS etc. through the environment.
extern "C" int vr_conv_pair2_test(int32_t device, int32_t H, int32_t W, int32_t cin, const float* x, const float* wa, const float* ba,
                                  const float* wb, const float* bb, float slope, float* ya, float* yb, int32_t iters, float* ms,
                                  const int32_t* gaps_x, int32_t ngx, const int32_t* gaps_y, int32_t ngy, int32_t flags) {
    if (!x || !wa || !wb || !ya || !yb || H <= 0 || W <= 0 || cin <= 0 || cin % 32 != 0 || ngx > 7 || ngy > 7) {
        set_error(nullptr, "vr_conv_pair2_test: bad arguments");
        return VR_E_INVALID;
    }
    ScopedDev sd(device);
    if (!sd.ok) return VR_E_NODEVICE;
    Device& dev = sd.dev;
    const size_t px = static_cast<size_t>(H) * W;
    const int planes = cin / 32 + 2;
    ConvWeights cwa, cwb;
    int rc = pack_conv_weights(dev, wa, ba, nullptr, cin, 32, &cwa);
    if (rc) return rc;
    rc = pack_conv_weights(dev, wb, bb, nullptr, cin + 32, 32, &cwb);
    if (rc) return rc;
    std::vector<__half> hx = to_half_padded(x, px, cin, cin, true);
    constexpr size_t kGuard = 8192;
    const size_t bytes = static_cast<size_t>(planes) * px * 32 * sizeof(__half);
    uint8_t* raw = nullptr;
    VR_CUDA_CHECK(cudaMalloc(&raw, bytes + 2 * kGuard), dev.err);
    VR_CUDA_CHECK(cudaMemset(raw, 0xA5, bytes + 2 * kGuard), dev.err);
    __half* buf = reinterpret_cast<__half*>(raw + kGuard);
    // the two output planes start as NaN patterns: every element the kernel owns must be written
    VR_CUDA_CHECK(cudaMemset(buf + static_cast<size_t>(cin / 32) * px * 32, 0xFF, 2 * px * 32 * sizeof(__half)), dev.err);
    VR_CUDA_CHECK(cudaMemcpy(buf, hx.data(), hx.size() * sizeof(__half), cudaMemcpyHostToDevice), dev.err);
    VR_CUDA_CHECK(cudaDeviceSynchronize(), dev.err);
    ConvCall c;
    c.in = buf;
    c.in_cstride = 32;
    c.in_planes = planes;
    c.in_pstride = static_cast<long long>(px) * 32;
    c.H = H;
    c.W = W;
    c.w = &cwa;
    c.w2 = &cwb;
    c.act = ACT_LRELU;
    c.slope = slope;
    c.out = buf;
    c.out_cstride = 32;
    c.out_pstride = static_cast<long long>(px) * 32;
    c.out_coff = cin;
    c.out_coff2 = cin + 32;
    c.flags = flags;
    long long* d_cyc = nullptr;
    cudaMalloc(&d_cyc, 1024 * sizeof(long long));
    cudaMemset(d_cyc, 0, 1024 * sizeof(long long));
    c.dbg_cycles = d_cyc;
    c.ngx = ngx;
    c.ngy = ngy;
    for (int i = 0; i < ngx; ++i) c.gx[i] = gaps_x[i];
    for (int i = 0; i < ngy; ++i) c.gy[i] = gaps_y[i];
    const int n = iters > 0 ? iters : 1;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    rc = run_conv(dev, c);
    if (rc == 0) {
        cudaEventRecord(e0, dev.stream);
        for (int i = 1; i < n && rc == 0; ++i) rc = run_conv(dev, c);
        cudaEventRecord(e1, dev.stream);
    }
    cudaError_t se = cudaStreamSynchronize(dev.stream);
    if (rc == 0 && se != cudaSuccess) {
        set_error(dev.err, std::string("pair2 kernel failed: ") + cudaGetErrorString(se));
        rc = VR_E_CUDA;
    }
    if (rc == 0) {
        if (ms) {
            float t = 0.f;
            if (n > 1) {
                cudaEventElapsedTime(&t, e0, e1);
                t /= (n - 1);
            }
            *ms = t;
        }
        std::vector<__half> hy(2 * px * 32);
        cudaMemcpy(hy.data(), buf + static_cast<size_t>(cin / 32) * px * 32, hy.size() * sizeof(__half), cudaMemcpyDeviceToHost);
        for (size_t p = 0; p < px; ++p)
            for (int ch = 0; ch < 32; ++ch) {
                ya[p * 32 + ch] = __half2float(hy[p * 32 + ch]);
                yb[p * 32 + ch] = __half2float(hy[(px + p) * 32 + ch]);
            }
        std::vector<uint8_t> g0(kGuard), g1(kGuard);
        cudaMemcpy(g0.data(), raw, kGuard, cudaMemcpyDeviceToHost);
        cudaMemcpy(g1.data(), raw + kGuard + bytes, kGuard, cudaMemcpyDeviceToHost);
        for (size_t i = 0; i < kGuard; ++i)
            if (g0[i] != 0xA5 || g1[i] != 0xA5) {
                set_error(dev.err, "pair2 kernel wrote outside its tensor (guard band damaged)");
                rc = VR_E_CUDA;
                break;
            }
        // the source planes must be untouched
        std::vector<__half> back(hx.size());
        cudaMemcpy(back.data(), buf, back.size() * sizeof(__half), cudaMemcpyDeviceToHost);
        if (rc == 0 && std::memcmp(back.data(), hx.data(), hx.size() * sizeof(__half)) != 0) {
            set_error(dev.err, "pair2 kernel modified its source planes");
            rc = VR_E_CUDA;
        }
    }
    if (d_cyc) {
        long long tmp[1024];
        if (cudaMemcpy(tmp, d_cyc, sizeof(tmp), cudaMemcpyDeviceToHost) == cudaSuccess)
            for (int i = 0; i < 64; ++i) g_pair2_prof[i] = tmp[300 + i];
        cudaFree(d_cyc);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(raw);
    free_conv_weights(&cwa);
    free_conv_weights(&cwb);
    return rc;
}

extern "C" int vr_conv3x3_bench(int32_t device, int32_t H, int32_t W, int32_t cin, int32_t cout, int32_t rows,
                                int32_t flags, int32_t iters, float* ms_out) {
    ScopedDev sd(device);
    if (!sd.ok) return VR_E_NODEVICE;
    Device& dev = sd.dev;
    const size_t px = static_cast<size_t>(H) * W;
    int cin_pad = (cin + 31) / 32 * 32;
    // VR_BENCH_CSTRIDE: time the layer on a wider interleaved source buffer (the network's 192-channel dense-block buffers)
    if (const char* e = std::getenv("VR_BENCH_CSTRIDE")) cin_pad = std::atoi(e) > cin_pad ? std::atoi(e) : cin_pad;
    std::vector<float> w(static_cast<size_t>(cout) * cin * 9);
    uint32_t s = 12345u;
    for (auto& v : w) {
        s = s * 1664525u + 1013904223u;
        v = (static_cast<float>(s >> 8) / 16777216.f - 0.5f) * 0.05f;
    }
    ConvWeights cw;
    int rc = pack_conv_weights(dev, w.data(), nullptr, nullptr, cin, cout, &cw);
    if (rc) return rc;
    __half *dx = nullptr, *dy = nullptr;
    int out_c = (cout == 3) ? 4 : cout;
    if (std::getenv("VR_BENCH_CSTRIDE") && cout != 3) out_c = cin_pad;
    VR_CUDA_CHECK(cudaMalloc(&dx, px * cin_pad * sizeof(__half)), dev.err);
    VR_CUDA_CHECK(cudaMalloc(&dy, px * out_c * sizeof(__half)), dev.err);
    VR_CUDA_CHECK(cudaMemset(dx, 0x11, px * cin_pad * sizeof(__half)), dev.err);  // small finite fp16 values
    ConvCall c;
    c.in = dx;
    c.in_cstride = cin_pad;
    c.H = H;
    c.W = W;
    c.w = &cw;
    c.act = ACT_LRELU;
    c.out = dy;
    c.out_cstride = out_c;
    if (std::getenv("VR_BENCH_PLANAR") && cout != 3) {
        // same bytes, chunk-planar: cin_pad / 32 source planes, destination plane(s) in a separate tensor
        c.in_cstride = 32;
        c.in_planes = cin_pad / 32;
        c.in_pstride = static_cast<long long>(px) * 32;
        c.out_cstride = 32;
        c.out_pstride = static_cast<long long>(px) * 32;
    }
    c.out_mode = (cout == 3) ? OUT_RGB4 : OUT_NHWC;
    c.rows = rows;
    c.flags = flags;
    long long* d_cyc = nullptr;
    cudaMalloc(&d_cyc, 1024 * sizeof(long long));
    cudaMemset(d_cyc, 0, 1024 * sizeof(long long));
    c.dbg_cycles = d_cyc;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int i = 0; i < 3 && rc == 0; ++i) rc = run_conv(dev, c);
    cudaEventRecord(e0, dev.stream);
    for (int i = 0; i < iters && rc == 0; ++i) rc = run_conv(dev, c);
    cudaEventRecord(e1, dev.stream);
    cudaError_t se = cudaStreamSynchronize(dev.stream);
    if (rc == 0 && se != cudaSuccess) {
        set_error(dev.err, std::string("conv bench kernel failed: ") + cudaGetErrorString(se));
        rc = VR_E_CUDA;
    }
    if (rc == 0) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        *ms_out = ms / iters;
        long long hc[1024];
        cudaMemcpy(hc, d_cyc, sizeof(hc), cudaMemcpyDeviceToHost);
        if (flags & FLAG_TRACE) {
            // per owned box of issuer w: wait start, full landed, first MMA issue, commit issued (cycles since CTA start)
            for (int i = 0; i < 32; ++i)
                for (int w = 0; w < 2; ++w) {
                    const long long* t = hc + 256 + w * 128 + i * 4;
                    std::fprintf(stderr, "[trace] box %2d warp %d: wait@%lld full+%lld issue+%lld commit+%lld\n", 16 + 2 * i + w, w,
                                 t[0], t[1] - t[0], t[2] - t[1], t[3] - t[2]);
                }
            std::fprintf(stderr, "[trace] CTA durations (cycles):");
            for (int i = 0; i < 148; ++i) std::fprintf(stderr, "%s%lld", i % 10 == 0 ? "\n   " : " ", hc[i]);
            std::fprintf(stderr, "\n");
            std::fprintf(stderr, "[trace] CTA 0: prologue done @%lld, first MMA @%lld, issuers done @%lld / %lld, CTA end @%lld\n", hc[1003],
                         hc[1000], hc[1001], hc[1002], hc[0]);
            // epilogue warps 0 / 4 (lane quarter 0 of either row parity): wait start, accumulators ready, TMEM loaded,
            // block re-initialised + released, staged in smem, stored to global
            for (int i = 0; i < 16; ++i)
                for (int w = 0; w < 2; ++w) {
                    const long long* t = hc + 512 + w * 128 + i * 8;
                    std::fprintf(stderr, "[etrace] row %2d group %d: wait@%lld ready+%lld ld+%lld release+%lld staged+%lld stored+%lld\n",
                                 8 + 2 * i + w, w, t[0], t[1] - t[0], t[2] - t[1], t[3] - t[2], t[4] - t[3], t[5] - t[4]);
                }
        }
        long long mx = 0;
        for (int i = 0; i < 256; ++i) mx = hc[i] > mx ? hc[i] : mx;
        g_last_conv_cycles = mx;  // slowest CTA of the last launch
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(dx);
    cudaFree(dy);
    cudaFree(d_cyc);
    free_conv_weights(&cw);
    return rc;
}

// ------------------------------------------------------------------------------------------------
// Device-resident timing of the HBM-bound kernels (K0/K2/K3) on synthetic frames.
// kind: 0 bilateral, 1 unsharp, 2 clahe (hist+lut+apply), 3 temporal, 4 post_crop, 5 post_blend (2x2 tiles),
//       6 pre, 7 upsample2x(64 ch)
// ------------------------------------------------------------------------------------------------
extern "C" int vr_filter_bench(int32_t device, int32_t kind, int32_t H, int32_t W, int32_t iters, float* ms_out) {
    ScopedDev sd(device);
    if (!sd.ok) return VR_E_NODEVICE;
    Device& dev = sd.dev;
    const size_t fb = static_cast<size_t>(H) * W * 3;
    const int64_t st = static_cast<int64_t>(W) * 3;
    uint8_t *a = nullptr, *b = nullptr, *c = nullptr;
    VR_CUDA_CHECK(cudaMalloc(&a, fb), dev.err);
    VR_CUDA_CHECK(cudaMalloc(&b, fb), dev.err);
    VR_CUDA_CHECK(cudaMalloc(&c, fb), dev.err);
    {
        std::vector<uint8_t> h(fb);
        uint32_t s = 99u;
        for (size_t i = 0; i < fb; ++i) {
            s = s * 1664525u + 1013904223u;
            h[i] = static_cast<uint8_t>((i / 3 % W) / 24 + (s >> 28));  // ramp + noise: non-degenerate for every filter
        }
        cudaMemcpy(a, h.data(), fb, cudaMemcpyHostToDevice);
        for (size_t i = 0; i < fb; ++i) h[i] = static_cast<uint8_t>(h[i] + ((i / 7) % 5));
        cudaMemcpy(b, h.data(), fb, cudaMemcpyHostToDevice);
    }
    int32_t* hist = nullptr;
    uint8_t* lut = nullptr;
    __half *t0 = nullptr, *t1 = nullptr;
    void* table = nullptr;
    BlendState blend_state;
    cudaMalloc(&hist, 64 * 256 * 4);
    cudaMalloc(&lut, 64 * 256);
    cudaMalloc(&table, 4096);
    const size_t tile_elems = static_cast<size_t>(H) * W * 4;
    if (kind >= 4) {
        const size_t t0_bytes = kind == 7 ? static_cast<size_t>(H) * W * 4 * 64 * 2 : kind == 5 ? tile_elems * 4 : tile_elems * 2;
        VR_CUDA_CHECK(cudaMalloc(&t0, t0_bytes), dev.err);
        VR_CUDA_CHECK(cudaMemset(t0, 0x38, kind == 5 ? t0_bytes : tile_elems * 2), dev.err);
        if (kind == 7 || kind == 6) VR_CUDA_CHECK(cudaMalloc(&t1, static_cast<size_t>(H) * W * 64 * 2), dev.err);
    }
    auto run = [&]() -> int {
        switch (kind) {
            case 0: return launch_bilateral(dev, a, st, H, W, c, st, 5, 25.f, 25.f);
            case 1: return launch_unsharp(dev, a, st, H, W, c, st, 0.5f);
            case 2: return launch_clahe(dev, a, st, H, W, c, st, 2.0f, 8, hist, lut, nullptr);
            case 3: return launch_temporal(dev, a, st, b, st, H, W, c, st, 0.2f, 12.f);
            case 4: return launch_post_crop(dev, t0, W, 0, 0, W, H, c, st, 0, 0);
            case 5: {
                // the tile grid of the `--quality max` preset at x4 (tile 2048, pad 256 output pixels; square tiles, the
                // last row / column smaller), every padded tile in its own region of t0, gather-blended
                const int T = 2048, pad = 256;
                const int ntx = (W + T - 1) / T, nty = (H + T - 1) / T;
                std::vector<BlendTile> tiles;
                size_t off = 0;
                for (int ty = 0; ty < nty; ++ty)
                    for (int tx = 0; tx < ntx; ++tx) {
                        const int x0 = std::max(tx * T - pad, 0), x1 = std::min((tx + 1) * T + pad, static_cast<int>(W));
                        const int y0 = std::max(ty * T - pad, 0), y1 = std::min((ty + 1) * T + pad, static_cast<int>(H));
                        tiles.push_back({t0 + off * 4, x0, y0, x1 - x0, y1 - y0, x1 - x0});
                        off += static_cast<size_t>(x1 - x0) * (y1 - y0);
                    }
                if (off > 2 * static_cast<size_t>(H) * W) {
                    set_error(dev.err, "vr_filter_bench: blend tiles exceed the staging buffer");
                    return VR_E_INVALID;
                }
                return launch_post_blend(dev, tiles, ntx, nty, T, pad, c, st, H, W, blend_state);
            }
            case 6: return launch_pre(dev, a, st, H, W, 0, 0, W, H, 0, t1, W, 0, 0);
            case 7: return launch_upsample2x(dev, t1, H, W, 64, reinterpret_cast<__half*>(t0));
            default: set_error(dev.err, "vr_filter_bench: unknown kind"); return VR_E_INVALID;
        }
    };
    int rc = 0;
    for (int i = 0; i < 3 && rc == 0; ++i) rc = run();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0, dev.stream);
    for (int i = 0; i < iters && rc == 0; ++i) rc = run();
    cudaEventRecord(e1, dev.stream);
    cudaError_t se = cudaStreamSynchronize(dev.stream);
    if (rc == 0 && se != cudaSuccess) {
        set_error(dev.err, std::string("filter bench kernel failed: ") + cudaGetErrorString(se));
        rc = VR_E_CUDA;
    }
    if (rc == 0) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        *ms_out = ms / iters;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    for (void* p : {static_cast<void*>(a), static_cast<void*>(b), static_cast<void*>(c), static_cast<void*>(hist),
                    static_cast<void*>(lut), static_cast<void*>(t0), static_cast<void*>(t1), table})
        if (p) cudaFree(p);
    free_blend_state(blend_state);
    return rc;
}
