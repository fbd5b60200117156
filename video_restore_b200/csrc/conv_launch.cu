// Host side of the convolution kernels K1 / K2 / K3: weight repacking into the swizzled shared-memory images, TMA
// tensor-map construction, kernel selection and launch. Kernels: conv3x3_sm100.cuh (K1), conv3x3_roll_sm100.cuh (K2),
// conv3x3_pair_sm100.cuh (K3, the default for 32 / 64 output channels).
#include "conv3x3_pair2_sm100.cuh"
#include "vr_common.h"

#include <cstdlib>
#include <cstring>
#include <mutex>

namespace vr {

void set_error(std::string* sink, const std::string& msg) {
    if (sink) *sink = msg;
    global_error() = msg;
}
std::string& global_error() {
    static thread_local std::string e;
    return e;
}

// ------------------------------------------------------------------------------------------------
// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time libcuda dependency,
// so the library also loads on the GPU-less build box for the symbol check).
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn(std::string* err) {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [&]() {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
        if (e == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    if (!fn) set_error(err, "cuTensorMapEncodeTiled entry point not available");
    return fn;
}

// NHWC fp16 activation tensor as a 4-D map (C, W, H, 1); box = 32 channels x pitch pixels x (rows + 2) lines.
// Chunk-planar tensors (cstride 32): the fourth dimension walks the planes.
static int make_act_tmap(Device& dev, const __half* ptr, int cstride, int W, int H, int rows, int kc, int planes,
                         long long pstride, CUtensorMap* out) {
    if (planes < 1) planes = 1;
    if (planes == 1) pstride = static_cast<long long>(H) * W * cstride;
    auto key = std::make_tuple(static_cast<const void*>(ptr), cstride, W, H, rows, kc, planes, pstride);
    auto it = dev.tmaps.find(key);
    if (it != dev.tmaps.end()) {
        *out = it->second;
        return 0;
    }
    EncodeTiledFn enc = get_encode_fn(dev.err);
    if (!enc) return -2;
    cuuint64_t dims[4] = {static_cast<cuuint64_t>(cstride), static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H),
                          static_cast<cuuint64_t>(planes)};
    cuuint64_t strides[3] = {static_cast<cuuint64_t>(cstride) * 2, static_cast<cuuint64_t>(W) * cstride * 2,
                             static_cast<cuuint64_t>(pstride) * 2};
    cuuint32_t box[4] = {static_cast<cuuint32_t>(kc), 130, static_cast<cuuint32_t>(rows + 2), 1};  // rows = 0: two lines (K2)
    cuuint32_t estr[4] = {1, 1, 1, 1};
    // VR_L2PROMO = 0 / 64 / 128 / 256 (default 128): L2 fetch granularity of the activation boxes (64 B per pixel and chunk)
    static const CUtensorMapL2promotion l2_promo = []() {
        const char* e = std::getenv("VR_L2PROMO");
        const int v = e ? std::atoi(e) : 128;
        return v == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE
                      : v == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                : v == 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
    }();
    CUtensorMap tm;
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<__half*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, kc == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B,
                     l2_promo,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error(dev.err, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string(static_cast<int>(r)) +
                               " (C=" + std::to_string(cstride) + " W=" + std::to_string(W) +
                               " H=" + std::to_string(H) + ")");
        return -2;
    }
    if (dev.tmaps.size() > 4096) dev.tmaps.clear();
    dev.tmaps[key] = tm;
    *out = tm;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// weights: OIHW fp32 -> [chunk][tap][n][32 ch] fp16 with the 64 B swizzle applied (16 B unit j of row n
// lands at unit j ^ ((n >> 1) & 3)), i.e. byte-for-byte what the MMA's B descriptor expects in smem.
// ------------------------------------------------------------------------------------------------
int pack_conv_weights(Device& dev, const float* w, const float* bias, const float* prelu, int cin, int cout,
                      ConvWeights* out, int kc) {
    if (cin <= 0 || cout <= 0 || cout > 64) {
        set_error(dev.err, "pack_conv_weights: unsupported channel counts");
        return -1;
    }
    if (kc == 0) {
        static const int env_kc = []() {
            const char* e = std::getenv("VR_KC");
            return e ? std::atoi(e) : 32;
        }();
        kc = env_kc;
    }
    if (kc != 16 && kc != 32) {
        set_error(dev.err, "pack_conv_weights: kc must be 16 or 32");
        return -1;
    }
    ConvWeights cw;
    cw.cin = cin;
    cw.cout = cout;
    cw.kc = kc;
    cw.npad = (cout + 15) / 16 * 16;
    cw.nchunks = (cin + kc - 1) / kc;
    const int N = cw.npad;
    const size_t elems = static_cast<size_t>(cw.nchunks) * 9 * N * kc;
    std::vector<__half> img(elems, __float2half(0.f));
    for (int c = 0; c < cw.nchunks; ++c)
        for (int t = 0; t < 9; ++t)
            for (int n = 0; n < cout; ++n)
                for (int ch = 0; ch < kc; ++ch) {
                    const int ci = c * kc + ch;
                    if (ci >= cin) continue;
                    const float v = w[(static_cast<size_t>(n) * cin + ci) * 9 + t];
                    // tap order in smem: dx-major, then dy = 2, 1, 0 (so one B descriptor spans the dy taps)
                    const int dy = t / 3, dx = t % 3;
                    const size_t row = (static_cast<size_t>(c) * 9 + dx * 3 + (2 - dy)) * N + n;
                    // swizzle of K-major rows: 64 B rows: unit ^= (row >> 1) & 3 ; 32 B rows: unit ^= (row >> 2) & 1
                    const int unit = kc == 32 ? ((ch >> 3) ^ ((n >> 1) & 3)) : ((ch >> 3) ^ ((n >> 2) & 1));
                    img[row * kc + unit * 8 + (ch & 7)] = __float2half_rn(v);
                }
    VR_CUDA_CHECK(cudaMalloc(&cw.wpack, elems * sizeof(__half)), dev.err);
    VR_CUDA_CHECK(cudaMemcpyAsync(cw.wpack, img.data(), elems * sizeof(__half), cudaMemcpyHostToDevice, dev.stream),
                  dev.err);
    if (cout == 64 && kc == 32) {
        // the same layer as two independent 32-channel halves [half][chunk][tap][32][kc] for the rolling kernel, whose
        // resident-weight budget a 64-channel layer with more than four chunks exceeds
        std::vector<__half> simg(elems, __float2half(0.f));
        const size_t half_elems = static_cast<size_t>(cw.nchunks) * 9 * 32 * kc;
        for (int n = 0; n < 64; ++n)
            for (int c = 0; c < cw.nchunks; ++c)
                for (int t = 0; t < 9; ++t) {
                    const size_t src = ((static_cast<size_t>(c) * 9 + t) * 64 + n) * kc;  // rows are already tap-ordered
                    const size_t dst = (n >> 5) * half_elems + ((static_cast<size_t>(c) * 9 + t) * 32 + (n & 31)) * kc;
                    // the swizzle depends on (n >> 1) & 3 only, identical for n and n & 31
                    std::memcpy(&simg[dst], &img[src], kc * sizeof(__half));
                }
        VR_CUDA_CHECK(cudaMalloc(&cw.wsplit, elems * sizeof(__half)), dev.err);
        VR_CUDA_CHECK(cudaMemcpyAsync(cw.wsplit, simg.data(), elems * sizeof(__half), cudaMemcpyHostToDevice, dev.stream),
                      dev.err);
        VR_CUDA_CHECK(cudaStreamSynchronize(dev.stream), dev.err);
    }
    if ((cout == 32 || cout == 64) && kc == 32) {
        // K3 (CTA pairs): per (chunk, dx) the 3*cout dy-stacked rows are split between the two CTAs of a pair:
        // [rank][chunk][dx][3*cout/2 rows][kc]; the swizzle phase of a row is unchanged (3*cout/2 is a multiple of 8)
        std::vector<__half> pimg(elems);
        const int half_rows = 3 * cout / 2;
        for (int rank = 0; rank < 2; ++rank)
            for (int c = 0; c < cw.nchunks; ++c)
                for (int dx = 0; dx < 3; ++dx)
                    std::memcpy(&pimg[((static_cast<size_t>(rank) * cw.nchunks + c) * 3 + dx) * half_rows * kc],
                                &img[((static_cast<size_t>(c) * 9 + dx * 3) * cout + static_cast<size_t>(rank) * half_rows) * kc],
                                static_cast<size_t>(half_rows) * kc * sizeof(__half));
        VR_CUDA_CHECK(cudaMalloc(&cw.wpair, elems * sizeof(__half)), dev.err);
        VR_CUDA_CHECK(cudaMemcpyAsync(cw.wpair, pimg.data(), elems * sizeof(__half), cudaMemcpyHostToDevice, dev.stream),
                      dev.err);
        VR_CUDA_CHECK(cudaStreamSynchronize(dev.stream), dev.err);
    }
    std::vector<float> b(cout, 0.f);
    if (bias) std::memcpy(b.data(), bias, cout * sizeof(float));
    VR_CUDA_CHECK(cudaMalloc(&cw.bias, cout * sizeof(float)), dev.err);
    VR_CUDA_CHECK(cudaMemcpyAsync(cw.bias, b.data(), cout * sizeof(float), cudaMemcpyHostToDevice, dev.stream),
                  dev.err);
    if (prelu) {
        VR_CUDA_CHECK(cudaMalloc(&cw.prelu, cout * sizeof(float)), dev.err);
        VR_CUDA_CHECK(cudaMemcpyAsync(cw.prelu, prelu, cout * sizeof(float), cudaMemcpyHostToDevice, dev.stream),
                      dev.err);
    }
    VR_CUDA_CHECK(cudaStreamSynchronize(dev.stream), dev.err);  // host staging vectors die here
    *out = cw;
    return 0;
}

void free_conv_weights(ConvWeights* w) {
    if (w->wpack) cudaFree(w->wpack);
    if (w->wsplit) cudaFree(w->wsplit);
    if (w->wpair) cudaFree(w->wpair);
    if (w->bias) cudaFree(w->bias);
    if (w->prelu) cudaFree(w->prelu);
    *w = ConvWeights();
}

// ------------------------------------------------------------------------------------------------
// launch
// ------------------------------------------------------------------------------------------------
template <int N, int TH, int KC, int DYS = 0, int DXS = 0>
static int launch_one(Device& dev, const CUtensorMap& tm, ConvArgs a) {
    using T = ConvTraits<N, TH, KC>;
    auto kern = conv3x3_tc_kernel<N, TH, KC, DYS, DXS>;
    static bool attr_done[64] = {};
    if (!attr_done[dev.ordinal & 63]) {
        VR_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, T::kBudget + T::kStgBytes + 1024),
                      dev.err);
        attr_done[dev.ordinal & 63] = true;
    }
    a.tiles_x = (a.W + 127) / 128;
    a.tiles_y = (a.y_end - a.y_begin + TH - 1) / TH;
    const int tiles = a.tiles_x * a.tiles_y * (DYS == 3 ? 4 : a.nlayers);  // work items of this launch (x4: all phases per tile)
    const int grid = tiles < dev.sm_count ? tiles : dev.sm_count;
    // shared-memory plan: resident weights when the whole layer fits WITHOUT reducing the pipeline depth (measured:
    // +3..5 % on 64->32, 96->32, 64->64; with only 2 activation stages left, 128->32 / 160->32 lose 4..8 %)
    const int w_bytes = a.nchunks * T::kBStage;
    const int a_stages = (T::kBudget - w_bytes) / T::kAStage;
    if (dev.weights_resident && a.nlayers == 1 && a_stages >= T::kStages && DYS != 3) {  // all-phase launch: weights change per item
        a.wres = 1;
        a.nstages = a_stages > kMaxStages ? kMaxStages : a_stages;
        a.stage_bytes = T::kAStage;
    } else {
        a.wres = 0;
        a.nstages = T::kStages;
        a.stage_bytes = T::kStageBytes;
    }
    const int smem_bytes = (a.wres ? w_bytes : 0) + a.nstages * a.stage_bytes + T::kStgBytes + 1024;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kConvThreads);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = dev.stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;  // PDL: prologue overlaps the previous kernel
    attr[0].val.programmaticStreamSerializationAllowed = dev.use_pdl ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    VR_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, tm, a), dev.err);
    dev.launches++;
    dev.conv_launches++;
    return 0;
}

// K2 work split: bands of output rows x 128-pixel strips (x channel halves). One wave of equal items is ideal; the cost
// model is waves x (band + 2 halo rows + ~2 rows of pipeline fill), minimised over the band count.
static void choose_bands(int units, int rows, int sms, int* band, int* nbands) {
    long best_cost = -1;
    for (int nb = 1; nb <= rows; ++nb) {
        const int bd = (rows + nb - 1) / nb;
        if (bd < 4 && nb > 1) break;
        const int real_nb = (rows + bd - 1) / bd;
        const long items = static_cast<long>(units) * real_nb;
        const long waves = (items + sms - 1) / sms;
        const long cost = waves * (bd + 4);
        if (best_cost < 0 || cost < best_cost) {
            best_cost = cost;
            *band = bd;
            *nbands = real_nb;
        }
    }
}

// returns 1 when the layer is not a K2 shape (caller falls back to K1), 0 on launch, < 0 on error
template <int N>
static int launch_roll(Device& dev, const CUtensorMap& tm, ConvArgs a, const ConvWeights& w) {
    using T = RollTraits<N>;
    const int w_bytes = w.nchunks * T::kBStage;
    int nslots = (T::kBudget - w_bytes) / T::kASlot;
    if (nslots < T::kMinSlots) return 1;
    if (nslots > kRollMaxSlots) nslots = kRollMaxSlots;
    auto kern = conv3x3_roll_kernel<N>;
    static bool attr_done[64] = {};
    if (!attr_done[dev.ordinal & 63]) {
        VR_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, T::kBudget + T::kStgBytes + 1024),
                      dev.err);
        attr_done[dev.ordinal & 63] = true;
    }
    a.nsplit = w.npad / N;
    a.wpack = a.nsplit == 2 ? w.wsplit : w.wpack;
    a.nstages = nslots;
    a.tiles_x = (a.W + 127) / 128;
    choose_bands(a.tiles_x * a.nsplit, a.y_end - a.y_begin, dev.sm_count, &a.band, &a.nbands);
    const int items = a.tiles_x * a.nbands * a.nsplit;
    int grid = items < dev.sm_count ? items : dev.sm_count;
    if (dev.max_ctas > 0 && grid > dev.max_ctas) grid = dev.max_ctas;
    grid -= grid % a.nsplit;  // a CTA keeps one channel half resident
    if (grid < a.nsplit) grid = a.nsplit;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kConvThreads);
    cfg.dynamicSmemBytes = w_bytes + nslots * T::kASlot + T::kStgBytes + 1024;
    cfg.stream = dev.stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = dev.use_pdl ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    VR_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, tm, a), dev.err);
    dev.launches++;
    dev.conv_launches++;
    return 0;
}

// K3: CTA pairs. returns 1 when the layer does not fit, 0 on launch, < 0 on error
template <int N>
static int launch_pair(Device& dev, const CUtensorMap& tm, ConvArgs a, const ConvWeights& w) {
    using T = PairTraits<N>;
    if (!w.wpair || w.cout != N) return 1;
    const int w_bytes = w.nchunks * T::kBHalf;
    // the direct epilogue needs 32-byte aligned 64-byte channel groups per pixel: chunk-planar tensors, or interleaved ones with
    // 32-channel multiples
    a.epi_direct = dev.epi_direct && a.omul == 1 && a.out_cstride % 32 == 0 && a.out_coff % 32 == 0 &&
                   (!a.res1 || (a.res1_cstride % 32 == 0 && a.res1_coff % 32 == 0)) &&
                   (!a.res2 || (a.res2_cstride % 32 == 0 && a.res2_coff % 32 == 0));
    // the direct epilogue has no staging buffer: its shared memory goes to activation slots
    const int stg_bytes = a.epi_direct ? 0 : T::kStgBytes;
    int nslots = (T::kBudget + T::kStgBytes - stg_bytes - w_bytes) / T::kASlot;
    if (nslots < T::kMinSlots) return 1;
    if (nslots > kPairMaxSlots) nslots = kPairMaxSlots;
    if (a.out2 && (!a.epi_direct || a.out2_cstride % 32 != 0 || a.res1 || a.res2)) return 1;
    // layers without residual operands get the instantiation that has no residual registers (N = 64: early ring release)
    const bool nores = a.epi_direct && !a.res1 && !a.res2;
    auto kern = nores ? conv3x3_pair_kernel<N, true, true>
                      : (a.epi_direct ? conv3x3_pair_kernel<N, true, false> : conv3x3_pair_kernel<N, false, false>);
    const int variant = nores ? 2 : a.epi_direct;
    static bool attr_done[3][64] = {};
    if (!attr_done[variant][dev.ordinal & 63]) {
        VR_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, T::kBudget + T::kStgBytes + 1024),
                      dev.err);
        attr_done[variant][dev.ordinal & 63] = true;
    }
    a.nsplit = 1;
    a.wpack = w.wpair;
    a.nstages = nslots;
    a.l2_hint = dev.l2_hint;
    a.l2_frac = dev.l2_frac;
    a.early64 = dev.early64;
    {
        // boxes per issuer hand-over: the next unit's operands should be landing while the current one executes
        const int env_unit = dev.pair_unit;
        // 32 channels: up to three boxes (measured: 1 -> 2 boxes -7 %, 2 -> 3 within noise). 64 channels: the ring has only
        // six logical positions, a row pair re-uses the positions of the row pair two before it, so a unit must never hold
        // both (deadlock) and should not wait on rows the other warp is still producing: two boxes, one for single-chunk layers.
        int u = N == 32 ? (nslots / 2 < 3 ? nslots / 2 : 3) : (w.nchunks >= 2 ? 2 : 1);
        if (env_unit > 0 && (N == 32 || env_unit <= w.nchunks + 1)) u = env_unit;
        if (u > nslots) u = nslots;
        a.unit = u < 1 ? 1 : u;
    }
    a.tiles_x = (a.W + 127) / 128;
    const int max_clusters = dev.sm_count / 2;
    if (dev.pair_pad) a.tiles_x = (a.tiles_x + 1) & ~1;  // A/B: pairs of adjacent strips with a padding strip
    // sub-items = strips x bands, two per cluster (any two with consecutive indices: no padding strip for odd strip counts)
    choose_bands(a.tiles_x, a.y_end - a.y_begin, 2 * max_clusters, &a.band, &a.nbands);
    const int items = (a.tiles_x * a.nbands + 1) / 2;
    int nclusters = items < max_clusters ? items : max_clusters;
    if (dev.max_ctas > 1 && nclusters > dev.max_ctas / 2) nclusters = dev.max_ctas / 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * nclusters);  // the kernel is compiled with __cluster_dims__(2, 1, 1)
    cfg.blockDim = dim3(T::kThreads);
    cfg.dynamicSmemBytes = w_bytes + nslots * T::kASlot + stg_bytes + 1024;
    cfg.stream = dev.stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = dev.use_pdl ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    VR_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, tm, a), dev.err);
    dev.launches++;
    dev.conv_launches++;
    return 0;
}

// K4: two 32-channel dense-block layers in one launch. returns 1 when the pair does not fit, 0 on launch, < 0 on error
static int launch_pair2(Device& dev, const CUtensorMap& tm, ConvArgs a, const ConvWeights& wa, const ConvWeights& wb) {
    using T = PairTraits<32>;
    if (!wa.wpair || !wb.wpair || wa.cout != 32 || wb.cout != 32 || wb.nchunks != wa.nchunks + 1) return 1;
    const int w_bytes = (wa.nchunks + wb.nchunks) * T::kBHalf;
    int nslots = (Pair2::kBudget - w_bytes - Pair2::kHand * T::kASlot) / T::kASlot;
    if (nslots < Pair2::kMinSlots) return 1;
    if (nslots > kPairMaxSlots) nslots = kPairMaxSlots;
    auto kern = conv3x3_pair2_kernel;
    static bool attr_done[64] = {};
    if (!attr_done[dev.ordinal & 63]) {
        VR_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Pair2::kBudget + 1024), dev.err);
        attr_done[dev.ordinal & 63] = true;
    }
    a.nsplit = 1;
    a.wpack = wa.wpair;
    a.wpack2 = wb.wpair;
    a.bias2 = wb.bias;
    a.nstages = nslots;
    a.unit = dev.pair_unit > 0 && dev.pair_unit <= 3 ? dev.pair_unit : 3;  // measured: 1 box per hand-over +10 %, 2 +2 %, 3 best
    a.lag = dev.k4_lag >= 2 && dev.k4_lag <= 4 ? dev.k4_lag : 0;
    a.prefetch = dev.k4_prefetch < 0 ? 0 : (dev.k4_prefetch > 16 ? 16 : dev.k4_prefetch);
    // a unit is waited for as a whole before any of its boxes is issued, so it must not hold row pairs p and p + 2 of one layer
    // (pair p + 2 re-uses ring positions pair p still occupies): with a single-chunk layer A the first steps of an item are
    // A_0, A_1, A_2 back to back
    if (wa.nchunks == 1 && a.unit > 2) a.unit = 2;
    if (a.unit > nslots) a.unit = nslots;
    a.tiles_x = (a.W + Pair2::kStrip - 1) / Pair2::kStrip;
    const int max_clusters = dev.sm_count / 2;
    choose_bands(a.tiles_x, a.y_end - a.y_begin, 2 * max_clusters, &a.band, &a.nbands);
    const int items = (a.tiles_x * a.nbands + 1) / 2;
    int nclusters = items < max_clusters ? items : max_clusters;
    if (dev.max_ctas > 1 && nclusters > dev.max_ctas / 2) nclusters = dev.max_ctas / 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * nclusters);
    cfg.blockDim = dim3(Pair2::kThreads);
    cfg.dynamicSmemBytes = w_bytes + (Pair2::kHand + nslots) * T::kASlot + 1024;
    cfg.stream = dev.stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = dev.use_pdl ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    VR_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, tm, a), dev.err);
    dev.launches++;
    dev.conv_launches++;
    return 0;
}

bool conv_supports_pair2(const Device& dev, int width) {
    if (!dev.fuse_pairs || !dev.planar || !(dev.rolling & 8)) return false;
    return dev.fuse_pairs >= 2 || (width + Pair2::kStrip - 1) / Pair2::kStrip == (width + 127) / 128;
}

bool conv_supports_out2(const Device& dev, int cout) {
    return dev.epi_direct && ((cout == 32 && (dev.rolling & 8)) || (cout == 64 && (dev.rolling & 16)));
}

void read_conv_env(Device& dev) {
    auto geti = [](const char* name, int* dst) {
        if (const char* e = std::getenv(name)) *dst = std::atoi(e);
    };
    auto getb = [](const char* name, bool* dst) {
        if (const char* e = std::getenv(name)) *dst = std::atoi(e) != 0;
    };
    getb("VR_PDL", &dev.use_pdl);
    geti("VR_ROLL", &dev.rolling);
    getb("VR_WRES", &dev.weights_resident);
    getb("VR_PAIRPAD", &dev.pair_pad);
    geti("VR_MAX_CTAS", &dev.max_ctas);
    geti("VR_EPI_DIRECT", &dev.epi_direct);
    geti("VR_K4", &dev.fuse_pairs);
    geti("VR_K4_LAG", &dev.k4_lag);
    geti("VR_K4_PREFETCH", &dev.k4_prefetch);
    geti("VR_EARLY64", &dev.early64);
    geti("VR_UNIT", &dev.pair_unit);
    geti("VR_L2HINT", &dev.l2_hint);
    if (const char* e = std::getenv("VR_L2FRAC")) dev.l2_frac = static_cast<float>(std::atof(e));
}

int run_conv(Device& dev, const ConvCall& c) {
    const ConvWeights& w = c.nlayers > 1 ? *c.lw[c.nlayers - 1] : *c.w;  // the layer with the widest channel prefix
    if (c.nlayers > 1) {
        if (c.nlayers > kMaxLayers || c.out_mode != OUT_NHWC || c.res1 || c.res2) {
            set_error(dev.err, "run_conv: unsupported multi-layer call");
            return -1;
        }
        for (int l = 0; l < c.nlayers; ++l)
            if (!c.lw[l] || c.lw[l]->npad != w.npad || c.lw[l]->kc != w.kc || c.lw[l]->cout != w.cout ||
                c.lw[l]->nchunks > w.nchunks || c.l_out_coff[l] % 8 != 0) {
                set_error(dev.err, "run_conv: multi-layer launch needs layers of one shape family");
                return -1;
            }
    }
    int rows = c.rows;
    if (rows == 0) rows = 4;
    CUtensorMap tm;
    if (c.in_cstride == 32 && (c.cin_off % 32 != 0 || (c.cin_off + w.nchunks * w.kc + 31) / 32 > c.in_planes)) {
        set_error(dev.err, "run_conv: chunk-planar source needs a 32-aligned channel prefix inside its planes");
        return -1;
    }
    int rc = make_act_tmap(dev, c.in, c.in_cstride, c.W, c.H, rows, w.kc, c.in_planes, c.in_pstride, &tm);
    if (rc) return rc;
    ConvArgs a;
    std::memset(&a, 0, sizeof(a));
    a.W = c.W;
    a.H = c.H;
    a.y_begin = c.y_begin < 0 ? 0 : c.y_begin;
    a.y_end = (c.y_end < 0 || c.y_end > c.H) ? c.H : c.y_end;
    if (a.y_end <= a.y_begin) return 0;
    a.nchunks = w.nchunks;
    a.cin_off = c.cin_off;
    a.in_cstride = c.in_cstride;
    a.out_pstride = c.out_pstride;
    a.res1_pstride = c.res1_pstride;
    a.res2_pstride = c.res2_pstride;
    a.wpack = w.wpack;
    a.bias = w.bias;
    a.prelu = w.prelu;
    a.act = c.act;
    a.slope = c.slope;
    a.out = c.out;
    a.out_cstride = c.out_cstride;
    a.out_coff = c.out_coff;
    a.out2 = c.out2;
    a.out2_cstride = c.out2_cstride;
    a.out_coff2 = c.out_coff2;
    a.cout = w.cout;
    a.res1 = c.res1;
    a.res1_cstride = c.res1_cstride;
    a.res1_coff = c.res1_coff;
    a.s1 = c.s1;
    a.res2 = c.res2;
    a.res2_cstride = c.res2_cstride;
    a.res2_coff = c.res2_coff;
    a.s2 = c.s2;
    a.out_mode = c.out_mode;
    a.base = c.base;
    a.base_cstride = c.base_cstride;
    a.flags = c.flags;
    a.dbg_cycles = c.dbg_cycles;
    a.omul = c.omul < 1 ? 1 : c.omul;
    a.opy = c.opy;
    a.opx = c.opx;
    a.nlayers = c.nlayers;
    if (c.nlayers > 1) {
        for (int l = 0; l < c.nlayers; ++l) {
            a.l_nchunks[l] = c.lw[l]->nchunks;
            a.l_wpack[l] = c.lw[l]->wpack;
            a.l_bias[l] = c.lw[l]->bias;
            a.l_out_coff[l] = c.l_out_coff[l];
        }
        if (!dev.dep_buf) {
            VR_CUDA_CHECK(cudaMalloc(&dev.dep_buf, 2 * Device::kDepRegion * sizeof(int)), dev.err);
            VR_CUDA_CHECK(cudaMemset(dev.dep_buf, 0, 2 * Device::kDepRegion * sizeof(int)), dev.err);
        }
        const int th = rows, tiles_y = ((a.y_end - a.y_begin) + th - 1) / th;
        if (c.nlayers * tiles_y > Device::kDepRegion) {
            set_error(dev.err, "run_conv: image too tall for the dependency counter region");
            return -1;
        }
        a.dep = dev.dep_buf + dev.dep_parity * Device::kDepRegion;
        a.dep_zero = dev.dep_buf + (dev.dep_parity ^ 1) * Device::kDepRegion;
        a.dep_zero_n = Device::kDepRegion;
        dev.dep_parity ^= 1;
    }
    a.ngx = c.ngx;
    a.ngy = c.ngy;
    a.gshift = c.gshift;
    for (int i = 0; i < 7; ++i) {
        a.gx[i] = c.gx[i];
        a.gy[i] = c.gy[i];
    }
    if (c.out_mode == OUT_NHWC && (w.cout % 16 != 0 || c.out_cstride % 8 != 0 || c.out_coff % 8 != 0)) {
        set_error(dev.err, "run_conv: NHWC output needs cout % 16 == 0 and 16-byte aligned channel slices");
        return -1;
    }
    if (c.out_mode == OUT_PS4 && w.npad != 48) {
        set_error(dev.err, "run_conv: pixel-shuffle output needs cout == 48");
        return -1;
    }
    if (c.w2) {
        // K4: this layer and the dense block's next one in one launch (chunk-planar tensors, full row range, LeakyReLU, no residuals)
        if (c.nlayers != 1 || c.out_mode != OUT_NHWC || c.in_cstride != 32 || c.out_cstride != 32 || c.res1 || c.res2 || c.out2 ||
            c.act != ACT_LRELU || c.slope < 0.f || c.slope > 1.f || c.dys || c.dxs || c.omul > 1 || a.y_begin != 0 || a.y_end != c.H ||
            c.out_coff % 32 != 0 || c.out_coff2 % 32 != 0 || w.kc != 32) {
            set_error(dev.err, "run_conv: this layer pair is not a K4 shape");
            return -1;
        }
        CUtensorMap tm2;
        rc = make_act_tmap(dev, c.in, c.in_cstride, c.W, c.H, 0, w.kc, c.in_planes, c.in_pstride, &tm2);  // two input rows
        if (rc) return rc;
        rc = launch_pair2(dev, tm2, a, w, *c.w2);
        if (rc == 1) {
            set_error(dev.err, "run_conv: the layer pair does not fit K4");
            return -1;
        }
        return rc;
    }
    const bool roll_shape = c.out_mode == OUT_NHWC && c.nlayers == 1 && !c.dys && !c.dxs && c.rows == 0 && w.kc == 32 &&
                            (w.cout == 32 || w.cout == 64);
    if (roll_shape && !(c.flags & FLAG_FORCE_TILE)) {
        const int mask = (c.flags & FLAG_FORCE_PAIR) ? 24 : (c.flags & FLAG_FORCE_ROLL) ? 7 : dev.rolling;
        CUtensorMap tm1;
        rc = make_act_tmap(dev, c.in, c.in_cstride, c.W, c.H, 0, w.kc, c.in_planes, c.in_pstride, &tm1);  // two input rows
        if (rc) return rc;
        rc = 1;
        if (w.cout == 32 && (mask & 8)) rc = launch_pair<32>(dev, tm1, a, w);
        if (w.cout == 64 && (mask & 16)) rc = launch_pair<64>(dev, tm1, a, w);
        if (rc <= 0) return rc;
        if (c.out2) {
            set_error(dev.err, "run_conv: a second destination needs the K3 direct epilogue (conv_supports_out2)");
            return -1;
        }
        if (w.cout == 32 && (mask & 1)) rc = launch_roll<32>(dev, tm1, a, w);
        if (w.cout == 64 && (mask & 2)) rc = launch_roll<64>(dev, tm1, a, w);                    // whole layer resident
        if (w.cout == 64 && rc == 1 && (mask & 4) && w.wsplit &&
            w.nchunks * RollTraits<64>::kBStage + RollTraits<64>::kMinSlots * RollTraits<64>::kASlot > RollTraits<64>::kBudget)
            rc = launch_roll<32>(dev, tm1, a, w);                                               // two resident halves
        if (rc <= 0) return rc;
    }
    if (c.out2) {
        set_error(dev.err, "run_conv: a second destination needs the K3 direct epilogue (conv_supports_out2)");
        return -1;
    }
    if (c.flags & (FLAG_FORCE_ROLL | FLAG_FORCE_PAIR)) {
        set_error(dev.err, "run_conv: the rolling-row kernels do not take this layer");
        return -1;
    }
    if (c.dys == 3 && c.dxs == 3) {
        // all four phases of an upsample-folded conv in one launch: lw[0..3] = the phases' pre-summed weights
        if (w.npad != 64 || rows != 4 || w.kc != 32 || c.nlayers != 1 || c.omul != 2 || c.out_mode != OUT_NHWC) {
            set_error(dev.err, "run_conv: no all-phase kernel for this layer shape");
            return -1;
        }
        for (int ph = 0; ph < 4; ++ph) {
            if (!c.lw[ph] || c.lw[ph]->npad != 64 || c.lw[ph]->kc != 32 || c.lw[ph]->nchunks != w.nchunks) {
                set_error(dev.err, "run_conv: all-phase launch needs four phase weight sets of the layer's shape");
                return -1;
            }
            a.l_wpack[ph] = c.lw[ph]->wpack;
        }
        return launch_one<64, 4, 32, 3, 3>(dev, tm, a);
    }
    if (c.dys || c.dxs) {
        // 2x2-tap phases of an upsample-folded conv: only the 64-channel / 4-row / 32-channel-chunk family is built
        if (w.npad != 64 || rows != 4 || w.kc != 32 || c.nlayers != 1 || c.dys < 1 || c.dys > 2 || c.dxs < 1 || c.dxs > 2) {
            set_error(dev.err, "run_conv: no phase kernel for this layer shape");
            return -1;
        }
        switch (c.dys * 10 + c.dxs) {
            case 11: return launch_one<64, 4, 32, 1, 1>(dev, tm, a);
            case 12: return launch_one<64, 4, 32, 1, 2>(dev, tm, a);
            case 21: return launch_one<64, 4, 32, 2, 1>(dev, tm, a);
            default: return launch_one<64, 4, 32, 2, 2>(dev, tm, a);
        }
    }
    const int key = (w.npad * 100 + rows) * 100 + w.kc;
    switch (key) {
        case (16 * 100 + 4) * 100 + 32: return launch_one<16, 4, 32>(dev, tm, a);
        case (32 * 100 + 4) * 100 + 32: return launch_one<32, 4, 32>(dev, tm, a);
        case (32 * 100 + 8) * 100 + 32: return launch_one<32, 8, 32>(dev, tm, a);
        case (48 * 100 + 4) * 100 + 32: return launch_one<48, 4, 32>(dev, tm, a);
        case (64 * 100 + 4) * 100 + 32: return launch_one<64, 4, 32>(dev, tm, a);
        case (16 * 100 + 4) * 100 + 16: return launch_one<16, 4, 16>(dev, tm, a);
        case (32 * 100 + 4) * 100 + 16: return launch_one<32, 4, 16>(dev, tm, a);
        case (32 * 100 + 8) * 100 + 16: return launch_one<32, 8, 16>(dev, tm, a);
        case (48 * 100 + 4) * 100 + 16: return launch_one<48, 4, 16>(dev, tm, a);
        case (64 * 100 + 4) * 100 + 16: return launch_one<64, 4, 16>(dev, tm, a);
        default:
            set_error(dev.err, "run_conv: no kernel instantiation for N=" + std::to_string(w.npad) +
                                   " rows=" + std::to_string(rows) + " kc=" + std::to_string(w.kc));
            return -1;
    }
}

}  // namespace vr
