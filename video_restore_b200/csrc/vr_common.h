// Host-side shared declarations for libvrb200.so (not part of the public ABI; see include/vrb200.h).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <map>
#include <string>
#include <tuple>
#include <vector>

namespace vr {

// ---- error plumbing: every internal function returns 0 / VR_E_* and records a message ----
void set_error(std::string* sink, const std::string& msg);
std::string& global_error();

#define VR_CUDA_CHECK(expr, sink)                                                                   \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess) {                                                                    \
            ::vr::set_error((sink), std::string(#expr) + ": " + cudaGetErrorString(_e));            \
            return -2;                                                                              \
        }                                                                                           \
    } while (0)

// ---- packed conv layer ----
struct ConvWeights {
    int cin = 0, cout = 0;
    int npad = 0;     // GEMM N: cout rounded up to 16
    int kc = 32;      // channels per pipeline stage (16 or 32)
    int nchunks = 0;  // ceil(cin / kc)
    __half* wpack = nullptr;  // device, [nchunks][dx][dy=2,1,0][npad][kc] swizzled
    __half* wsplit = nullptr; // device, cout == 64 only: the layer as two 32-channel halves [half][nchunks][tap][32][kc]
    __half* wpair = nullptr;  // device, cout == 32 / 64: K3's per-CTA halves [rank][nchunks][dx][3*cout/2][kc]
    float* bias = nullptr;    // device [cout]
    float* prelu = nullptr;   // device [cout] or null
};

struct ConvCall {
    const __half* in = nullptr;  // NHWC fp16
    int in_cstride = 0;          // channels per pixel of the source buffer (per plane); 32 = chunk-planar (ConvArgs)
    int in_planes = 1;           // planes of the source tensor and their distance in elements (chunk-planar only)
    long long in_pstride = 0;
    int cin_off = 0;
    int H = 0, W = 0;
    int y_begin = 0, y_end = -1;  // output row range (default: all rows)
    const ConvWeights* w = nullptr;
    int act = 0;
    float slope = 0.2f;
    __half* out = nullptr;
    int out_cstride = 0, out_coff = 0;
    __half* out2 = nullptr;  // optional second destination (K3 direct epilogue only), same channel offset / plane distance
    int out2_cstride = 0;
    // K4: the dense block's next layer fused into this launch (32 -> 32 output channels, LeakyReLU, chunk-planar tensors):
    // it reads the same source prefix plus this layer's 32 output channels and writes channel slice out_coff2
    const ConvWeights* w2 = nullptr;
    int out_coff2 = 0;
    long long out_pstride = 0, res1_pstride = 0, res2_pstride = 0;  // plane distances (cstride == 32 tensors)
    const __half* res1 = nullptr;
    int res1_cstride = 0, res1_coff = 0;
    float s1 = 1.f;
    const __half* res2 = nullptr;
    int res2_cstride = 0, res2_coff = 0;
    float s2 = 1.f;
    int out_mode = 0;  // ConvOut
    const __half* base = nullptr;
    int base_cstride = 0;
    int rows = 0;   // output rows per CTA tile (TH), 0 = default
    // phase of an upsample-folded conv: taps present (0 all, 1 = {0,1}, 2 = {1,2}) and the output pixel mapping
    int dys = 0, dxs = 0;
    int omul = 1, opy = 0, opx = 0;
    // tile-atlas gap mask (see ConvArgs): gap columns / rows in units of (1 << gshift) pixels
    int ngx = 0, ngy = 0, gshift = 0;
    int gx[7] = {0}, gy[7] = {0};
    int flags = 0;  // ConvFlags (debug ablations)
    long long* dbg_cycles = nullptr;
    // multi-layer launch: layers 0..nlayers-1 share `in` (growing channel prefix) and `out`; per layer: weights and the
    // output channel offset. nlayers == 1: the scalar fields above describe the layer.
    int nlayers = 1;
    const ConvWeights* lw[4] = {nullptr, nullptr, nullptr, nullptr};
    int l_out_coff[4] = {0, 0, 0, 0};
};

struct Device {
    int ordinal = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    std::string* err = nullptr;
    int64_t launches = 0;
    int64_t conv_launches = 0;  // of which convolution kernels (K1 / K2 / K3 / K4)
    // bilateral tables of this handle (device global memory), rebuilt only when the parameters change
    void* bil_tab = nullptr;
    int bil_d = -1;
    float bil_sc = -1.f, bil_ss = -1.f;
    std::vector<char> bil_host;
    // dependency counters of multi-layer launches: two regions used alternately (each launch zeroes the other one)
    int* dep_buf = nullptr;
    int dep_parity = 0;
    static constexpr int kDepRegion = 4 * 4096;
    // VR_MULTI=1: conv1..conv4 of a dense block in one persistent launch with tile-row dependency counters. Bit-identical;
    // measured +1 % (720p single tile) / -2.7 % (6-tile atlas: resident weights are lost, little wave tail to recover),
    // so off by default. Kept as the base of the per-RDB persistent kernel.
    bool multi_layer = false;
    bool fold_upsample = true;     // VR_FOLD_UP=0: materialise nearest x2 and run conv_up1/2 as plain 3x3 convs
    bool blend_fast = true;        // VR_BLEND_FAST=0: the general Gaussian-blend kernel on aligned frames too (A/B runs, tests)
    bool fuse_phases = true;       // VR_PHASES1=0: the four phases of a folded upsample conv as four launches (A/B runs)
    bool weights_resident = true;  // VR_WRES=0: always stream weights with the activations
    // VR_ROLL bit mask: which NHWC 3x3 layers run on the rolling-row kernels K2 / K3 instead of the tiled kernel K1:
    // 1 = 32-channel outputs, 2 = 64-channel outputs whose weights fit (cin <= 128), 4 = 64-channel outputs as two halves;
    // 8 / 16 = 32- / 64-channel outputs on the CTA-pair kernel K3 (takes precedence)
    // Default 24 + 7: K3 wherever it fits, else K2, else K1. Measured in-network (720p x4plus, interleaved A/B, sustained clocks):
    // K1 only 42.7 ms, K2 on 32-channel layers 41.1 ms, K3 on 64-channel layers + K2 39.0 ms, K3 on both 36.9 ms.
    int rolling = 31;
    bool planar = true;   // VR_PLANAR=0: interleaved [pixel][C] activation tensors instead of chunk-planar ones (A/B)
    bool pair_pad = false;  // VR_PAIRPAD=1: K3 pairs adjacent strips only (odd strip counts get a padding strip), for A/B runs
    int max_ctas = 0;     // test hook (VR_MAX_CTAS): cap K2 / K3 grids so that a CTA / CTA pair walks several work items
    bool use_pdl = true;  // VR_PDL=0 disables programmatic dependent launch of the conv kernels
    // K3 switches (A/B runs; defaults are the fast path): VR_EPI_DIRECT=0 stages the epilogue through shared memory;
    // VR_EARLY64 = 0 / 1 / 2 when a 64-channel row's ring position goes back (ConvArgs::early64); VR_UNIT boxes per issuer
    // hand-over (0 = automatic); VR_L2HINT / VR_L2FRAC cache-policy experiments (ConvArgs::l2_hint)
    int epi_direct = 1;
    // VR_K4_PREFETCH: row pairs of prefetch-only TMA requests ahead of layer A's loads. Off: measured 2 .. 14 % SLOWER, growing
    // with the distance (profiles/r2_k4_prefetch.txt) -- the TMA unit's request rate on 64-byte rows (260 per box, ~600 cycles: what
    // the box's 12 MMAs take) is itself near-critical on the 32-channel layers, so extra requests cost more than the latency they hide
    int k4_prefetch = 0;
    int k4_lag = 0;       // VR_K4_LAG: K4's row-pair lag of layer B (0 = default)
    int fuse_pairs = 1;   // VR_K4: 0 = conv1+conv2 / conv3+conv4 as separate K3 launches, 1 = K4 where it costs no extra strip, 2 = always
    int early64 = 2;
    int pair_unit = 0;
    int l2_hint = 0;
    float l2_frac = 0.5f;
    // tensor-map cache: (ptr, cstride, W, H, rows, kc, planes, pstride)
    std::map<std::tuple<const void*, int, int, int, int, int, int, long long>, CUtensorMap> tmaps;
};

int pack_conv_weights(Device& dev, const float* w_oihw, const float* bias, const float* prelu, int cin, int cout,
                      ConvWeights* out, int kc = 0);  // kc = 0: default (env VR_KC or 32)
void free_conv_weights(ConvWeights* w);
int run_conv(Device& dev, const ConvCall& c);
// K4 is enabled and its prerequisites (K3, chunk-planar tensors) hold. VR_K4 = 1 (default): only where its 126-pixel strips do not
// cost an extra strip over K3's 128 (the 6-tile atlas: 13 vs 13; 1280 wide: 11 vs 10, where the extra 10 % of MMA work eats the gain);
// 2: always; 0: never
bool conv_supports_pair2(const Device& dev, int width);
bool conv_supports_out2(const Device& dev, int cout);  // the configured kernel for a `cout`-channel NHWC layer takes ConvCall::out2
// reads every conv-related environment switch into `dev` (called when a handle / test device is created, so that one
// process can run several configurations)
void read_conv_env(Device& dev);

// ---- elementwise / filter kernels (kernels_frame.cu) ----
int launch_upsample2x(Device& dev, const __half* src, int H, int W, int C, __half* dst);
// u8 BGR frame rect -> fp16 RGB NHWC32 (zero padded channels); reflect pads past the frame edge;
// unshuffle=1 applies pixel_unshuffle(2) (12 channels, c*4 + dy*2 + dx)
// dst is an NHWC32 image of `dst_pitch` pixels per row; the tile lands at (dst_x0, dst_y0)
int launch_pre(Device& dev, const uint8_t* frame, int64_t stride, int H, int W, int x0, int y0, int w, int h,
               int unshuffle, __half* dst, int dst_pitch, int dst_x0, int dst_y0);
// RGB4 fp16 tile -> clamp, *255, rint, BGR u8 into the frame rect
int launch_post_crop(Device& dev, const __half* tile, int tile_w, int crop_x0, int crop_y0, int w, int h,
                     uint8_t* frame, int64_t stride, int dst_x0, int dst_y0);

struct BlendTile {
    const __half* data;  // RGB4 fp16, first pixel of the tile; rows are `pitch` pixels apart
    int px0, py0, pw, ph;  // padded output rect in the scaled frame
    int pitch;
};
struct BlendState {  // per-handle cache of the blend kernel's device tables (rebuilt when the tile layout changes)
    std::vector<BlendTile> last;
    void* d_table = nullptr;
    size_t table_cap = 0;
    float* d_weights = nullptr;
    size_t weights_cap = 0;
    std::vector<char> host_table;
};
int launch_post_blend(Device& dev, const std::vector<BlendTile>& tiles, int tiles_x, int tiles_y, int tile_out,
                      int pad_out, uint8_t* frame, int64_t stride, int sH, int sW, BlendState& st);
void free_blend_state(BlendState& st);
int launch_bilateral(Device& dev, const uint8_t* src, int64_t sstride, int H, int W, uint8_t* dst, int64_t dstride,
                     int d, float sigma_color, float sigma_space);
int launch_unsharp(Device& dev, const uint8_t* src, int64_t sstride, int H, int W, uint8_t* dst, int64_t dstride,
                   float amount);
// temporal stage fused into CLAHE's apply pass (vectorised path only): dst receives the UN-blended frame (the next frame's
// "previous"), `blended` (dense rows) the frame blended with `prev`. *fused tells the caller whether it happened.
struct TemporalFuse {
    const uint8_t* prev;
    uint8_t* blended;
    float alpha, tau;
};
int launch_clahe(Device& dev, const uint8_t* src, int64_t sstride, int H, int W, uint8_t* dst, int64_t dstride,
                 float clip, int grid, int32_t* d_hist, uint8_t* d_lut, uint8_t* d_luma, const TemporalFuse* tf = nullptr,
                 bool* fused = nullptr);
int launch_temporal(Device& dev, const uint8_t* cur, int64_t cstride, const uint8_t* prev, int64_t pstride, int H,
                    int W, uint8_t* dst, int64_t dstride, float alpha, float tau);
int launch_blend_weights(Device& dev, int extent, float* d_w);

}  // namespace vr
