// C ABI of libvrb200.so: handle lifetime, weight loading, the network executors (RRDBNet / SRVGGNetCompact as
// sequences of K1 launches over zero-copy NHWC buffers), RealESRGANer tiling, and the per-frame restore chain.
// Interface being replaced: reference video_upscaler.py:328-338 (constructor) and :490-505 (_process_frame).
#include "../../include/vrb200.h"
#include "conv3x3_sm100.cuh"
#include <cstdio>
#include "vr_common.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>

using namespace vr;

namespace {

struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
};

struct TileRect {
    int in_x0, in_x1, in_y0, in_y1, pad_x0, pad_x1, pad_y0, pad_y1, out_x0, out_x1, out_y0, out_y1;
};

// RealESRGANer.tile_process index arithmetic ([3P] realesrgan/utils.py), integers only.
std::vector<TileRect> make_tile_grid(int H, int W, int tile, int pad, int scale, int* tiles_x_out = nullptr,
                                     int* tiles_y_out = nullptr) {
    std::vector<TileRect> v;
    const int tiles_x = (W + tile - 1) / tile, tiles_y = (H + tile - 1) / tile;
    if (tiles_x_out) *tiles_x_out = tiles_x;
    if (tiles_y_out) *tiles_y_out = tiles_y;
    for (int y = 0; y < tiles_y; ++y)
        for (int x = 0; x < tiles_x; ++x) {
            TileRect t;
            const int ofs_x = x * tile, ofs_y = y * tile;
            t.in_x0 = ofs_x;
            t.in_x1 = std::min(ofs_x + tile, W);
            t.in_y0 = ofs_y;
            t.in_y1 = std::min(ofs_y + tile, H);
            t.pad_x0 = std::max(t.in_x0 - pad, 0);
            t.pad_x1 = std::min(t.in_x1 + pad, W);
            t.pad_y0 = std::max(t.in_y0 - pad, 0);
            t.pad_y1 = std::min(t.in_y1 + pad, H);
            t.out_x0 = (t.in_x0 - t.pad_x0) * scale;
            t.out_x1 = t.out_x0 + (t.in_x1 - t.in_x0) * scale;
            t.out_y0 = (t.in_y0 - t.pad_y0) * scale;
            t.out_y1 = t.out_y0 + (t.in_y1 - t.in_y0) * scale;
            v.push_back(t);
        }
    return v;
}

struct RawTensor {
    std::vector<float> data;
    std::vector<int64_t> shape;
};

std::string g_create_error;

}  // namespace

struct Gaps {  // tile-atlas separators in NETWORK pixels (see ConvArgs); empty for a single tile
    int ngx = 0, ngy = 0;
    int gx[7] = {0}, gy[7] = {0};
};

struct vr_handle {
    vr_config cfg;
    Device dev;
    std::string err;
    std::map<std::string, RawTensor> raw;
    std::map<std::string, ConvWeights> layers;
    bool committed = false;
    int atlas_tiles = 8;  // tiles per atlas axis (VR_ATLAS_TILES, read at vr_create)
    bool atlas_forced = false;   // VR_ATLAS_TILES given: no automatic sizing
    size_t mem_budget = 0;       // bytes the activations of ONE atlas group may take (0.8 x free memory at vr_create, or VR_MEM_BUDGET_MB)
    int auto_H = 0, auto_W = 0, auto_group = 8;  // automatic group size of the last frame geometry
    Gaps gaps;        // atlas gap mask of the frame being processed (read by conv())
    int gap_shift = 0;  // log2 of the current layer's resolution multiple
    int atlas_w = 0, atlas_h = 0;
    // network activations (NHWC fp16), grown on demand
    DevBuf in32, feat, trunk, rdb[3], up1_in, up1_out, up2_in, up2_out, hr_out, sv[2];
    std::vector<DevBuf> tile_out;
    // u8 frames
    DevBuf lr_in, lr_dn, hr[2], prev_up, stage_in, stage_out;
    bool has_prev = false;
    int prev_h = 0, prev_w = 0;
    DevBuf clahe_hist, clahe_lut;
    BlendState blend;
    // boundary frames received from the left neighbour's handle (in-process multi-GPU, pipeline.py): device buffers on THIS
    // handle's device, filled by the neighbour's thread with one cudaMemcpyPeerAsync each, recycled through a free list
    std::mutex bmu;
    std::vector<DevBuf> bfree;
    // pipelined host path (vr_submit / vr_wait): two slots
    cudaStream_t s_in = nullptr, s_out = nullptr;
    DevBuf pipe_in[2], pipe_out[2];
    cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_comp[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
    int64_t submitted = 0;
    // per-frame timing events: a ring of records so that vr_sync can average over EVERY frame enqueued since the previous
    // sync (bench.py reads the conv time of the timed steps themselves, not of one extra step)
    struct FrameEv {
        cudaEvent_t t0 = nullptr, c0 = nullptr, c1 = nullptr, t1 = nullptr;
    };
    std::vector<FrameEv> ev_ring;
    int64_t ev_next = 0, ev_pending = 0;
    FrameEv* ev_cur = nullptr;
    float last_total_ms = 0.f, last_conv_ms = 0.f;
    int32_t last_frames = 0;
    bool timing_valid = false;
};

namespace {

int ensure(vr_handle* h, DevBuf& b, size_t bytes) {
    if (b.bytes >= bytes && b.p) return 0;
    if (b.p) {
        VR_CUDA_CHECK(cudaStreamSynchronize(h->dev.stream), h->dev.err);
        cudaFree(b.p);
        b.p = nullptr;
        b.bytes = 0;
        h->dev.tmaps.clear();
    }
    bytes = (bytes + 255) / 256 * 256;
    VR_CUDA_CHECK(cudaMalloc(&b.p, bytes), h->dev.err);
    b.bytes = bytes;
    return 0;
}
void release(DevBuf& b) {
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.bytes = 0;
}

int fail(vr_handle* h, int code, const std::string& msg) {
    set_error(&h->err, msg);
    return code;
}

const ConvWeights* layer(vr_handle* h, const std::string& name) {
    auto it = h->layers.find(name);
    return it == h->layers.end() ? nullptr : &it->second;
}

#define VR_TRY(expr)            \
    do {                        \
        int _rc = (expr);       \
        if (_rc) return _rc;    \
    } while (0)

// -------------------------------------------------------------------------------------------------
// conv helper
// -------------------------------------------------------------------------------------------------
struct Act {
    __half* p = nullptr;
    int c = 0;          // channels of the tensor
    long long ps = 0;   // > 0: chunk-planar (planes of 32 channels, `ps` elements apart); 0: interleaved [pixel][c]
};
// Activation tensors with a multiple of 32 channels are chunk-planar (VR_PLANAR=0: interleaved, for A/B runs): every TMA box row
// and every output row is then one contiguous run; the interleaved form costs the 32-channel layers 16..52 %.
Act make_act(const vr_handle* h, void* ptr, int channels, size_t px) {
    Act a;
    a.p = static_cast<__half*>(ptr);
    a.c = channels;
    a.ps = (h->dev.planar && channels % 32 == 0) ? static_cast<long long>(px) * 32 : 0;
    return a;
}
struct Rows {  // output row range of one launch (row-band scheduling); default = all rows
    int y0 = 0, y1 = -1;
};
void set_io(ConvCall& c, const Act& in, const Act& out) {
    c.in = in.p;
    c.in_cstride = in.ps ? 32 : in.c;
    c.in_planes = in.ps ? in.c / 32 : 1;
    c.in_pstride = in.ps;
    c.out = out.p;
    c.out_cstride = out.ps ? 32 : out.c;
    c.out_pstride = out.ps;
}

int conv(vr_handle* h, const std::string& name, Act in, int nh, int nw, Act out, int out_coff, int act, Act res1 = Act(),
         float s1 = 1.f, Act res2 = Act(), float s2 = 1.f, int out_mode = OUT_NHWC, const __half* base = nullptr,
         int base_c = 0, Rows rows = Rows(), int phase = -1, Act out2 = Act()) {
    const ConvWeights* w = layer(h, phase == 4 ? name + ".phase0" : name);
    if (!w) return fail(h, VR_E_STATE, "missing layer " + name);
    ConvCall c;
    set_io(c, in, out);
    c.H = nh;
    c.W = nw;
    c.w = w;
    c.act = act;
    c.slope = 0.2f;
    c.out_coff = out_coff;
    c.out2 = out2.p;
    c.out2_cstride = out2.ps ? 32 : out2.c;
    c.res1 = res1.p;
    c.res1_cstride = res1.ps ? 32 : res1.c;
    c.res1_pstride = res1.ps;
    c.s1 = s1;
    c.res2 = res2.p;
    c.res2_cstride = res2.ps ? 32 : res2.c;
    c.res2_pstride = res2.ps;
    c.s2 = s2;
    c.out_mode = out_mode;
    c.base = base;
    c.base_cstride = base_c;
    c.y_begin = rows.y0;
    c.y_end = rows.y1;
    if (phase == 4) {  // all four phases in one launch (input tiles read once; weights per phase)
        c.dys = c.dxs = 3;
        c.omul = 2;
        for (int ph = 0; ph < 4; ++ph) {
            c.lw[ph] = layer(h, name + ".phase" + std::to_string(ph));
            if (!c.lw[ph]) return fail(h, VR_E_STATE, "missing layer " + name + ".phase" + std::to_string(ph));
        }
    } else if (phase >= 0) {  // output phase (py, px) of a conv folded with a preceding nearest x2 upsample
        const int py = phase >> 1, px = phase & 1;
        c.dys = py == 0 ? 1 : 2;  // py = 0 reads rows y-1, y (taps 0,1); py = 1 reads rows y, y+1 (taps 1,2)
        c.dxs = px == 0 ? 1 : 2;
        c.omul = 2;
        c.opy = py;
        c.opx = px;
    }
    c.ngx = h->gaps.ngx;
    c.ngy = h->gaps.ngy;
    c.gshift = h->gap_shift;
    for (int i = 0; i < 7; ++i) {
        c.gx[i] = h->gaps.gx[i];
        c.gy[i] = h->gaps.gy[i];
    }
    return run_conv(h->dev, c);
}

// nearest x2 of a whole activation tensor (only when VR_FOLD_UP=0): plane by plane for chunk-planar tensors
int upsample2x_act(vr_handle* h, const Act& src, int H, int W, const Act& dst) {
    if (!src.ps) return launch_upsample2x(h->dev, src.p, H, W, src.c, dst.p);
    for (int k = 0; k < src.c / 32; ++k) VR_TRY(launch_upsample2x(h->dev, src.p + k * src.ps, H, W, 32, dst.p + k * dst.ps));
    return 0;
}

// Optional row-band scheduling of the dense blocks (VR_BAND_MB = band working set in MB; default 0 = off).
// Measured on B200 (profiles/r1_band_sweep.txt): bit-identical output, but no L2 benefit -- the conv's cost per row
// is the same for a 64-row and a 720-row image (it is MMA/pipeline bound, not HBM bound) -- while every extra launch
// costs ~4-10 us, so banding by separate launches only loses. Kept for experiments.
int band_rows_for(int nh, int nw) {
    static const int target_mb = []() {
        const char* e = std::getenv("VR_BAND_MB");
        return e ? std::atoi(e) : 0;
    }();
    if (target_mb <= 0) return nh;
    long rows = static_cast<long>(target_mb) * 1000000L / (static_cast<long>(nw) * 512L);
    rows = rows / 4 * 4;
    if (rows < 16) rows = 16;
    return rows >= nh ? nh : static_cast<int>(rows);
}

// RRDBNet.forward on one (already pixel-unshuffled when scale == 2) tile of nh x nw network pixels.
// Dense concat is zero-copy: each RDB owns one 192-channel NHWC buffer; conv_k reads the channel prefix
// [0, 64 + 32(k-1)) and writes its 32 channels behind it; conv5 writes the next RDB's channels [0, 64).
int run_rrdbnet(vr_handle* h, int nh, int nw, __half* tile_out) {
    const size_t px = static_cast<size_t>(nh) * nw;
    VR_TRY(ensure(h, h->feat, px * 64 * 2));
    VR_TRY(ensure(h, h->trunk, px * 64 * 2));
    for (int i = 0; i < 3; ++i) VR_TRY(ensure(h, h->rdb[i], px * 192 * 2));
    if (!h->dev.fold_upsample) VR_TRY(ensure(h, h->up1_in, px * 4 * 64 * 2));
    VR_TRY(ensure(h, h->up1_out, px * 4 * 64 * 2));
    VR_TRY(ensure(h, h->up2_in, px * 16 * 64 * 2));
    VR_TRY(ensure(h, h->up2_out, px * 16 * 64 * 2));
    Act in32 = make_act(h, h->in32.p, 32, px);
    Act feat = make_act(h, h->feat.p, 64, px);
    Act trunk = make_act(h, h->trunk.p, 64, px);
    Act rdb[3] = {make_act(h, h->rdb[0].p, 192, px), make_act(h, h->rdb[1].p, 192, px), make_act(h, h->rdb[2].p, 192, px)};
    const int band = band_rows_for(nh, nw);
    // conv_first feeds the trunk skip (`feat`) AND the first dense block's x slot: one launch, two destinations from the
    // same epilogue registers (K3 direct epilogue); with another kernel configured (A/B switches) two launches
    if (conv_supports_out2(h->dev, 64) && feat.ps == rdb[0].ps) {
        VR_TRY(conv(h, "conv_first", in32, nh, nw, feat, 0, ACT_NONE, Act(), 1.f, Act(), 1.f, OUT_NHWC, nullptr, 0, Rows(), -1, rdb[0]));
    } else {
        VR_TRY(conv(h, "conv_first", in32, nh, nw, feat, 0, ACT_NONE));
        VR_TRY(conv(h, "conv_first", in32, nh, nw, rdb[0], 0, ACT_NONE));
    }
    for (int b = 0; b < h->cfg.num_block; ++b) {
        for (int r = 0; r < 3; ++r) {
            const std::string pre = "body." + std::to_string(b) + ".rdb" + std::to_string(r + 1) + ".conv";
            Act X = rdb[r], Y = rdb[(r + 1) % 3];
            // Row bands: conv_k of a band also produces the (5 - k) halo rows above and below that conv_{k+1..5}
            // of the SAME band need, so x1..x4 are produced and consumed while still in L2 (halo rows are
            // recomputed by the neighbouring band with identical values).
            for (int b0 = 0; b0 < nh; b0 += band) {
                const int b1 = std::min(b0 + band, nh);
                if (h->dev.multi_layer && band >= nh) {
                    // conv1..conv4 in ONE persistent launch: tiles of conv_{k+1} start as soon as the rows of conv_k
                    // they read are complete (no wave tail / pipeline refill between the four layers)
                    ConvCall mc;
                    set_io(mc, X, X);
                    mc.H = nh;
                    mc.W = nw;
                    mc.act = ACT_LRELU;
                    mc.slope = 0.2f;
                    mc.nlayers = 4;
                    for (int k = 1; k <= 4; ++k) {
                        mc.lw[k - 1] = layer(h, pre + std::to_string(k));
                        if (!mc.lw[k - 1]) return fail(h, VR_E_STATE, "missing layer " + pre + std::to_string(k));
                        mc.l_out_coff[k - 1] = 64 + 32 * (k - 1);
                    }
                    mc.ngx = h->gaps.ngx;
                    mc.ngy = h->gaps.ngy;
                    mc.gshift = h->gap_shift;
                    for (int i = 0; i < 7; ++i) {
                        mc.gx[i] = h->gaps.gx[i];
                        mc.gy[i] = h->gaps.gy[i];
                    }
                    VR_TRY(run_conv(h->dev, mc));
                } else if (conv_supports_pair2(h->dev, nw) && band >= nh) {
                    // K4: conv1 + conv2 and conv3 + conv4 as two launches; the second layer of each pair reads x .. x_{k-1} from
                    // L2 and x_k from shared memory instead of HBM (26 -> 18 plane transfers per dense block)
                    for (int k = 1; k <= 3; k += 2) {
                        const ConvWeights* wa = layer(h, pre + std::to_string(k));
                        const ConvWeights* wb = layer(h, pre + std::to_string(k + 1));
                        if (!wa || !wb) return fail(h, VR_E_STATE, "missing layer " + pre + std::to_string(k));
                        ConvCall pc;
                        set_io(pc, X, X);
                        pc.H = nh;
                        pc.W = nw;
                        pc.w = wa;
                        pc.w2 = wb;
                        pc.act = ACT_LRELU;
                        pc.slope = 0.2f;
                        pc.out_coff = 64 + 32 * (k - 1);
                        pc.out_coff2 = 64 + 32 * k;
                        pc.ngx = h->gaps.ngx;
                        pc.ngy = h->gaps.ngy;
                        pc.gshift = h->gap_shift;
                        for (int i = 0; i < 7; ++i) {
                            pc.gx[i] = h->gaps.gx[i];
                            pc.gy[i] = h->gaps.gy[i];
                        }
                        VR_TRY(run_conv(h->dev, pc));
                    }
                } else
                for (int k = 1; k <= 4; ++k) {
                    const Rows rr{std::max(b0 - (5 - k), 0), std::min(b1 + (5 - k), nh)};
                    VR_TRY(conv(h, pre + std::to_string(k), X, nh, nw, X, 64 + 32 * (k - 1), ACT_LRELU, Act(), 1.f, Act(), 1.f,
                                OUT_NHWC, nullptr, 0, rr));
                }
                const Rows r5{b0, b1};
                if (r < 2) {
                    VR_TRY(conv(h, pre + "5", X, nh, nw, Y, 0, ACT_NONE, X, 0.2f, Act(), 1.f, OUT_NHWC, nullptr, 0,
                                r5));  // x5*0.2 + x
                } else {
                    // (x5*0.2 + x)*0.2 + rrdb_in ; rrdb_in lives in rdb[0][:, 0:64] and is overwritten in place
                    VR_TRY(conv(h, pre + "5", X, nh, nw, Y, 0, ACT_NONE, X, 0.2f, Y, 0.2f, OUT_NHWC, nullptr, 0, r5));
                }
            }
        }
    }
    VR_TRY(conv(h, "conv_body", rdb[0], nh, nw, trunk, 0, ACT_NONE, feat, 1.0f));  // feat + body_feat
    Act u1i = make_act(h, h->up1_in.p, 64, px * 4), u1o = make_act(h, h->up1_out.p, 64, px * 4);
    Act u2i = make_act(h, h->up2_in.p, 64, px * 16), u2o = make_act(h, h->up2_out.p, 64, px * 16);
    if (h->dev.fold_upsample) {
        // lrelu(conv_up(nearest_x2(f))) as four 2x2-tap convs on f itself (one per output phase, pre-summed weights):
        // no upsampled tensor, 4/9 of the MACs
        // one launch per upsample conv walks the four phases of every input tile (the input is read from HBM once, not four
        // times); VR_PHASES1=0: four launches (bit-identical)
        if (h->dev.fuse_phases) {
            VR_TRY(conv(h, "conv_up1", trunk, nh, nw, u1o, 0, ACT_LRELU, Act(), 1.f, Act(), 1.f, OUT_NHWC, nullptr, 0, Rows(), 4));
        } else {
            for (int ph = 0; ph < 4; ++ph)
                VR_TRY(conv(h, "conv_up1.phase" + std::to_string(ph), trunk, nh, nw, u1o, 0, ACT_LRELU, Act(), 1.f, Act(), 1.f,
                            OUT_NHWC, nullptr, 0, Rows(), ph));
        }
        h->gap_shift = 1;  // the phases of conv_up2 run on the 2x grid: gap columns / rows are 2 pixels wide there
        if (h->dev.fuse_phases) {
            VR_TRY(conv(h, "conv_up2", u1o, 2 * nh, 2 * nw, u2o, 0, ACT_LRELU, Act(), 1.f, Act(), 1.f, OUT_NHWC, nullptr, 0, Rows(), 4));
        } else {
            for (int ph = 0; ph < 4; ++ph)
                VR_TRY(conv(h, "conv_up2.phase" + std::to_string(ph), u1o, 2 * nh, 2 * nw, u2o, 0, ACT_LRELU, Act(), 1.f, Act(), 1.f,
                            OUT_NHWC, nullptr, 0, Rows(), ph));
        }
        h->gap_shift = 2;
    } else {
        VR_TRY(upsample2x_act(h, trunk, nh, nw, u1i));
        h->gap_shift = 1;  // gap columns / rows are 2 pixels wide at 2x, 4 at 4x (nearest upsampling keeps them zero)
        VR_TRY(conv(h, "conv_up1", u1i, 2 * nh, 2 * nw, u1o, 0, ACT_LRELU));
        VR_TRY(upsample2x_act(h, u1o, 2 * nh, 2 * nw, u2i));
        h->gap_shift = 2;
        VR_TRY(conv(h, "conv_up2", u2i, 4 * nh, 4 * nw, u2o, 0, ACT_LRELU));
    }
    VR_TRY(conv(h, "conv_hr", u2o, 4 * nh, 4 * nw, u2i, 0, ACT_LRELU));  // up2_in is dead: reuse for conv_hr out
    Act to = make_act(h, tile_out, 4, 0);
    VR_TRY(conv(h, "conv_last", u2i, 4 * nh, 4 * nw, to, 0, ACT_NONE, Act(), 1.f, Act(), 1.f, OUT_RGB4));
    h->gap_shift = 0;
    return 0;
}

// SRVGGNetCompact.forward: conv+PReLU chain, last conv fused with PixelShuffle(4) + nearest-upsampled input.
int run_srvgg(vr_handle* h, int nh, int nw, __half* tile_out) {
    const size_t px = static_cast<size_t>(nh) * nw;
    VR_TRY(ensure(h, h->sv[0], px * 64 * 2));
    VR_TRY(ensure(h, h->sv[1], px * 64 * 2));
    Act in32 = make_act(h, h->in32.p, 32, px);
    Act a = make_act(h, h->sv[0].p, 64, px), b = make_act(h, h->sv[1].p, 64, px);
    VR_TRY(conv(h, "body.0", in32, nh, nw, a, 0, ACT_PRELU));
    for (int i = 0; i < h->cfg.num_conv; ++i) {
        VR_TRY(conv(h, "body." + std::to_string(2 * (i + 1)), a, nh, nw, b, 0, ACT_PRELU));
        std::swap(a, b);
    }
    Act to = make_act(h, tile_out, 4, 0);
    VR_TRY(conv(h, "body." + std::to_string(2 * (h->cfg.num_conv + 1)), a, nh, nw, to, 0, ACT_NONE, Act(), 1.f, Act(), 1.f,
                OUT_PS4, in32.p, 32));
    return 0;
}

constexpr int kEvRing = 256;
vr_handle::FrameEv* begin_frame_events(vr_handle* h) {
    if (h->ev_ring.empty()) {
        h->ev_ring.resize(kEvRing);
        for (auto& f : h->ev_ring) {
            cudaEventCreate(&f.t0);
            cudaEventCreate(&f.c0);
            cudaEventCreate(&f.c1);
            cudaEventCreate(&f.t1);
        }
    }
    vr_handle::FrameEv* f = &h->ev_ring[h->ev_next % kEvRing];
    ++h->ev_next;
    if (h->ev_pending < kEvRing) ++h->ev_pending;
    return f;
}

// The whole per-frame chain on device-resident frames; enqueues on h->dev.stream, no host sync on the way
// (except first-use allocations).
int restore_enqueue(vr_handle* h, const uint8_t* d_bgr, int H, int W, int64_t stride, uint8_t* d_out,
                    int64_t out_stride, const vr_frame_opts* opts) {
    if (!h->committed) return fail(h, VR_E_STATE, "vr_restore before vr_commit_weights");
    if (H <= 0 || W <= 0 || !d_bgr || !d_out) return fail(h, VR_E_INVALID, "vr_restore: bad frame arguments");
    vr_frame_opts o;
    std::memset(&o, 0, sizeof(o));
    if (opts) o = *opts;
    const vr_config& cfg = h->cfg;
    const int s = cfg.scale;
    const int sH = H * s, sW = W * s;
    Device& dev = h->dev;
    VR_CUDA_CHECK(cudaSetDevice(dev.ordinal), dev.err);
    vr_handle::FrameEv* fev = begin_frame_events(h);
    cudaEventRecord(fev->t0, dev.stream);

    // (1) bilateral pre-denoise on the LR frame (video_upscaler.py:495-496)
    const uint8_t* src = d_bgr;
    int64_t sstride = stride;
    if (o.denoise) {
        VR_TRY(ensure(h, h->lr_dn, static_cast<size_t>(H) * W * 3));
        VR_TRY(launch_bilateral(dev, src, sstride, H, W, static_cast<uint8_t*>(h->lr_dn.p), static_cast<int64_t>(W) * 3,
                                o.denoise_d, o.denoise_sigma_color, o.denoise_sigma_space));
        src = static_cast<const uint8_t*>(h->lr_dn.p);
        sstride = static_cast<int64_t>(W) * 3;
    }

    // (2) RealESRGANer.enhance: mod-pad (scale 2), tile loop, merge
    int Hp = H, Wp = W;
    if (s == 2) {
        Hp += H % 2;
        Wp += W % 2;
    }
    int tiles_x = 0, tiles_y = 0;
    const std::vector<TileRect> grid = make_tile_grid(Hp, Wp, cfg.tile, cfg.tile_pad, s, &tiles_x, &tiles_y);
    const bool post_chain = (o.sharpen > 0.f) || o.clahe || o.temporal;
    const size_t hr_bytes = static_cast<size_t>(sH) * sW * 3;
    const int64_t hr_stride = static_cast<int64_t>(sW) * 3;
    uint8_t* up_dst = d_out;
    int64_t up_stride = out_stride;
    if (post_chain) {
        VR_TRY(ensure(h, h->hr[0], hr_bytes));
        up_dst = static_cast<uint8_t*>(h->hr[0].p);
        up_stride = hr_stride;
    }
    const bool blend = cfg.blend == VR_BLEND_GAUSSIAN;
    // Tile atlas: all padded tiles of the frame are packed into ONE network-input image, grid column by grid column,
    // separated by 1-pixel zero gap columns / rows. A 3x3 conv cannot see across a zero gap that is re-zeroed after
    // every layer, and the gap is exactly the zero padding tile_process gives each tile, so the result is bit-identical
    // to running the tiles one by one -- with one launch per layer for the whole frame instead of one per tile.
    // More than 8 tiles along an axis (e.g. 2160p with --tile-size 256): the grid is cut into groups of at most 8 x 8 tiles, one
    // atlas each (the kernels' gap tables hold 7 separators per axis). RealESRGANer.tile_process has no such limit.
    const int net_div = s == 2 ? 2 : 1;  // x2 models run on the pixel-unshuffled (half resolution) grid
    for (size_t ti = 0; ti < grid.size(); ++ti) {
        const TileRect& t = grid[ti];
        if (s == 2 && (((t.pad_x1 - t.pad_x0) | (t.pad_y1 - t.pad_y0)) & 1))
            return fail(h, VR_E_INVALID,
                        "x2 model: padded tile extent must be even (pixel_unshuffle); use an even tile size/overlap");
    }
    // VR_ATLAS_TILES (1..8, default 8): tiles per atlas axis. Activation memory is proportional to the atlas, not to the frame: a
    // smaller group bounds it the way `--tile-size` bounds it in the reference (same result, more launches).
    int kGroup = h->atlas_tiles;
    if (!h->atlas_forced && h->mem_budget > 0) {
        // automatic: the largest group (<= 8 tiles per axis) whose activations fit the budget. Bytes per network-input pixel of an
        // RRDBNet group: in32 64 + feat 128 + trunk 128 + 3 dense-block buffers 3 x 384 + HR buffers (4 + 16 + 16) x 128 (+ 4 x 128
        // without the folded upsample) + the fp16 RGB4 output 16 x 8; the SRVGG network needs less (the figure is an upper bound)
        if (h->auto_H != H || h->auto_W != W) {
            const size_t per_px = 64 + 128 + 128 + 3 * 384 + (4 + 16 + 16) * 128 + (h->dev.fold_upsample ? 0 : 4 * 128) + 16 * 8;
            int g = 8;
            for (; g > 1; --g) {
                size_t wa = 0, ha = 0;
                for (int j = 0; j < std::min(g, tiles_x); ++j) wa += (grid[j].pad_x1 - grid[j].pad_x0) / net_div + 1;
                for (int i = 0; i < std::min(g, tiles_y); ++i)
                    ha += (grid[static_cast<size_t>(i) * tiles_x].pad_y1 - grid[static_cast<size_t>(i) * tiles_x].pad_y0) / net_div + 1;
                if (wa * ha * per_px <= h->mem_budget) break;
            }
            h->auto_H = H;
            h->auto_W = W;
            h->auto_group = g;
        }
        kGroup = std::min(kGroup, h->auto_group);
    }
    const int groups_x = (tiles_x + kGroup - 1) / kGroup, groups_y = (tiles_y + kGroup - 1) / kGroup;
    if (h->tile_out.size() < static_cast<size_t>(groups_x) * groups_y) h->tile_out.resize(static_cast<size_t>(groups_x) * groups_y);
    std::vector<BlendTile> btiles(blend ? grid.size() : 0);
    for (int gy = 0; gy < groups_y; ++gy)
        for (int gx = 0; gx < groups_x; ++gx) {
            const int tx0 = gx * kGroup, ty0 = gy * kGroup;
            const int ntx = std::min(kGroup, tiles_x - tx0), nty = std::min(kGroup, tiles_y - ty0);
            int colw[8], rowh[8], ax0[8], ay0[8];
            for (int j = 0; j < ntx; ++j) {
                const TileRect& t = grid[static_cast<size_t>(ty0) * tiles_x + tx0 + j];
                colw[j] = (t.pad_x1 - t.pad_x0) / net_div;
            }
            for (int i = 0; i < nty; ++i) {
                const TileRect& t = grid[static_cast<size_t>(ty0 + i) * tiles_x + tx0];
                rowh[i] = (t.pad_y1 - t.pad_y0) / net_div;
            }
            Gaps gaps;
            int Wa = 0, Ha = 0;
            for (int j = 0; j < ntx; ++j) {
                ax0[j] = Wa;
                Wa += colw[j];
                if (j + 1 < ntx) gaps.gx[gaps.ngx++] = Wa++;
            }
            for (int i = 0; i < nty; ++i) {
                ay0[i] = Ha;
                Ha += rowh[i];
                if (i + 1 < nty) gaps.gy[gaps.ngy++] = Ha++;
            }
            const size_t in_bytes = static_cast<size_t>(Ha) * Wa * 32 * 2;
            const bool in_realloc = h->in32.bytes < in_bytes || !h->in32.p;
            VR_TRY(ensure(h, h->in32, in_bytes));
            if (in_realloc || h->atlas_w != Wa || h->atlas_h != Ha) {
                // gap pixels of the network input are never written by pre_kernel: zero once per layout
                VR_CUDA_CHECK(cudaMemsetAsync(h->in32.p, 0, in_bytes, dev.stream), dev.err);
                h->atlas_w = Wa;
                h->atlas_h = Ha;
            }
            DevBuf& tob = h->tile_out[static_cast<size_t>(gy) * groups_x + gx];
            const int out_pitch = Wa * 4;  // every model's network output is 4x the network-input grid
            VR_TRY(ensure(h, tob, static_cast<size_t>(Ha) * 4 * out_pitch * 4 * 2));
            for (int i = 0; i < nty; ++i)
                for (int j = 0; j < ntx; ++j) {
                    const TileRect& t = grid[static_cast<size_t>(ty0 + i) * tiles_x + tx0 + j];
                    VR_TRY(launch_pre(dev, src, sstride, H, W, t.pad_x0, t.pad_y0, t.pad_x1 - t.pad_x0, t.pad_y1 - t.pad_y0,
                                      s == 2 ? 1 : 0, static_cast<__half*>(h->in32.p), Wa, ax0[j], ay0[i]));
                }
            h->gaps = gaps;
            h->gap_shift = 0;
            // conv time of the frame: first network launch .. last network launch (with several groups the few pre / merge
            // kernels between the groups' networks are inside the bracket: < 1 %)
            if (gx == 0 && gy == 0) cudaEventRecord(fev->c0, dev.stream);
            if (cfg.model_kind == VR_MODEL_RRDBNET)
                VR_TRY(run_rrdbnet(h, Ha, Wa, static_cast<__half*>(tob.p)));
            else
                VR_TRY(run_srvgg(h, Ha, Wa, static_cast<__half*>(tob.p)));
            if (gx == groups_x - 1 && gy == groups_y - 1) cudaEventRecord(fev->c1, dev.stream);
            for (int i = 0; i < nty; ++i)
                for (int j = 0; j < ntx; ++j) {
                    const size_t ti = static_cast<size_t>(ty0 + i) * tiles_x + tx0 + j;
                    const TileRect& t = grid[ti];
                    const int pw = t.pad_x1 - t.pad_x0, ph = t.pad_y1 - t.pad_y0;
                    const __half* origin =
                        static_cast<const __half*>(tob.p) + (static_cast<size_t>(ay0[i]) * 4 * out_pitch + ax0[j] * 4) * 4;
                    if (blend) {
                        btiles[ti] = {origin, t.pad_x0 * s, t.pad_y0 * s, pw * s, ph * s, out_pitch};
                    } else {
                        const int dx0 = t.in_x0 * s, dy0 = t.in_y0 * s;
                        const int w = std::min(t.in_x1 * s, sW) - dx0, hh = std::min(t.in_y1 * s, sH) - dy0;  // un-pad (post_process)
                        VR_TRY(launch_post_crop(dev, origin, out_pitch, t.out_x0, t.out_y0, w, hh, up_dst, up_stride, dx0, dy0));
                    }
                }
        }
    if (blend) {
        VR_TRY(launch_post_blend(dev, btiles, tiles_x, tiles_y, cfg.tile * s, cfg.tile_pad * s, up_dst, up_stride, sH,
                                 sW, h->blend));
    }

    // (3) enhancement stage on the HR u8 frame; the last stage writes straight into d_out
    if (post_chain) {
        int cur = 0;  // h->hr[cur] holds the current frame
        auto cur_ptr = [&]() { return static_cast<uint8_t*>(h->hr[cur].p); };
        const bool has_sharpen = o.sharpen > 0.f;
        if (has_sharpen) {
            const bool last = !o.clahe && !o.temporal;
            uint8_t* dst = d_out;
            int64_t dstride = out_stride;
            if (!last) {
                VR_TRY(ensure(h, h->hr[cur ^ 1], hr_bytes));
                dst = static_cast<uint8_t*>(h->hr[cur ^ 1].p);
                dstride = hr_stride;
            }
            VR_TRY(launch_unsharp(dev, cur_ptr(), hr_stride, sH, sW, dst, dstride, o.sharpen));
            if (!last) cur ^= 1;
        }
        const float alpha = o.temporal_alpha > 0 ? o.temporal_alpha : 0.2f;
        const float tau = o.temporal_tau > 0 ? o.temporal_tau : 12.f;
        const bool prev_ok = h->has_prev && h->prev_h == sH && h->prev_w == sW;
        bool temporal_done = false;
        if (o.clahe) {
            const bool last = !o.temporal;
            uint8_t* dst = d_out;
            int64_t dstride = out_stride;
            if (!last) {
                VR_TRY(ensure(h, h->hr[cur ^ 1], hr_bytes));
                dst = static_cast<uint8_t*>(h->hr[cur ^ 1].p);
                dstride = hr_stride;
            }
            const int g = o.clahe_grid > 0 ? o.clahe_grid : 8;
            VR_TRY(ensure(h, h->clahe_hist, static_cast<size_t>(g) * g * 256 * 4));
            VR_TRY(ensure(h, h->clahe_lut, static_cast<size_t>(g) * g * 256));
            // with a previous frame at hand the temporal blend rides on CLAHE's apply pass: the un-blended frame goes to
            // hr[cur ^ 1] (it becomes up_{t-1}), the blended one straight to d_out
            TemporalFuse tf{static_cast<const uint8_t*>(h->prev_up.p), d_out, alpha, tau};
            const bool want_fuse = o.temporal && prev_ok && out_stride == hr_stride;
            VR_TRY(launch_clahe(dev, cur_ptr(), hr_stride, sH, sW, dst, dstride, o.clahe_clip > 0 ? o.clahe_clip : 2.0f, g,
                                static_cast<int32_t*>(h->clahe_hist.p), static_cast<uint8_t*>(h->clahe_lut.p), nullptr,
                                want_fuse ? &tf : nullptr, &temporal_done));
            if (!last) cur ^= 1;
        }
        if (o.temporal && temporal_done) {
            std::swap(h->hr[cur], h->prev_up);  // up_t becomes up_{t-1}; no copy
        } else if (o.temporal) {
            if (prev_ok) {
                VR_TRY(launch_temporal(dev, cur_ptr(), hr_stride, static_cast<const uint8_t*>(h->prev_up.p), hr_stride,
                                       sH, sW, d_out, out_stride, alpha, tau));
            } else {
                VR_CUDA_CHECK(cudaMemcpy2DAsync(d_out, out_stride, cur_ptr(), hr_stride, hr_stride, sH,
                                                cudaMemcpyDeviceToDevice, dev.stream),
                              dev.err);
            }
            std::swap(h->hr[cur], h->prev_up);  // up_t becomes up_{t-1}; no copy
            h->has_prev = true;
            h->prev_h = sH;
            h->prev_w = sW;
        }
    }
    cudaEventRecord(fev->t1, dev.stream);
    h->timing_valid = false;
    return 0;
}

int finish_timing(vr_handle* h) {
    VR_CUDA_CHECK(cudaStreamSynchronize(h->dev.stream), h->dev.err);
    // average over the frames enqueued since the previous sync (at most the ring's depth; a frame whose enqueue failed half
    // way has unrecorded events and is skipped)
    double total = 0.0, conv = 0.0;
    int n = 0;
    for (int64_t i = h->ev_next - h->ev_pending; i < h->ev_next; ++i) {
        const vr_handle::FrameEv& f = h->ev_ring[i % kEvRing];
        float t = 0.f, c = 0.f;
        if (cudaEventElapsedTime(&t, f.t0, f.t1) != cudaSuccess || cudaEventElapsedTime(&c, f.c0, f.c1) != cudaSuccess) {
            cudaGetLastError();
            continue;
        }
        total += t;
        conv += c;
        ++n;
    }
    h->ev_pending = 0;
    h->last_frames = n;
    h->last_total_ms = n ? static_cast<float>(total / n) : 0.f;
    h->last_conv_ms = n ? static_cast<float>(conv / n) : 0.f;
    h->timing_valid = n > 0;
    return 0;
}

}  // namespace

// =================================================================================================
// exported ABI
// =================================================================================================
extern "C" {

int vr_create(const vr_config* cfg, vr_handle** out) {
    if (!cfg || !out) {
        g_create_error = "vr_create: null argument";
        set_error(nullptr, g_create_error);
        return VR_E_INVALID;
    }
    auto bad = [&](const std::string& m, int code) {
        g_create_error = m;
        set_error(nullptr, m);
        return code;
    };
    if (cfg->model_kind != VR_MODEL_RRDBNET && cfg->model_kind != VR_MODEL_SRVGG)
        return bad("vr_create: unknown model_kind", VR_E_INVALID);
    if (cfg->model_kind == VR_MODEL_RRDBNET && cfg->scale != 4 && cfg->scale != 2)
        return bad("vr_create: RRDBNet scale must be 4 or 2", VR_E_INVALID);
    if (cfg->model_kind == VR_MODEL_SRVGG && cfg->scale != 4) return bad("vr_create: SRVGG scale must be 4", VR_E_INVALID);
    if (cfg->num_feat != 64 || (cfg->model_kind == VR_MODEL_RRDBNET && cfg->num_grow_ch != 32))
        return bad("vr_create: only num_feat=64 / num_grow_ch=32 (the reference's architectures) are built", VR_E_INVALID);
    if (cfg->tile <= 0 || cfg->tile_pad < 0) return bad("vr_create: tile must be > 0 and tile_pad >= 0", VR_E_INVALID);
    if (cfg->pre_pad != 0) return bad("vr_create: pre_pad != 0 is not supported (reference passes 0)", VR_E_INVALID);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return bad("vr_create: no CUDA device (this library has no CPU fallback)", VR_E_NODEVICE);
    if (cfg->device < 0 || cfg->device >= ndev) return bad("vr_create: device ordinal out of range", VR_E_NODEVICE);
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, cfg->device) != cudaSuccess) return bad("cudaGetDeviceProperties failed", VR_E_CUDA);
    if (p.major != 10)
        return bad("vr_create: device is sm_" + std::to_string(p.major * 10 + p.minor) +
                       ", this library is built for sm_100a only",
                   VR_E_NODEVICE);
    std::unique_ptr<vr_handle> h(new vr_handle());
    h->cfg = *cfg;
    h->dev.ordinal = cfg->device;
    h->dev.sm_count = p.multiProcessorCount;
    h->dev.err = &h->err;
    if (const char* e = std::getenv("VR_L2PERSIST")) {
        // experiment: L2 set-aside for evict_last / persisting lines, in MB (0 = leave the driver default); see VR_L2HINT
        int max_persist = 0;
        cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, cfg->device);
        size_t want = static_cast<size_t>(std::atoi(e)) << 20;
        if (want > static_cast<size_t>(max_persist)) want = static_cast<size_t>(max_persist);
        cudaError_t pe = cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want);
        std::fprintf(stderr, "[vrb200] VR_L2PERSIST: max %d MB, set %zu MB (%s)\n", max_persist >> 20, want >> 20, cudaGetErrorString(pe));
    }
    read_conv_env(h->dev);
    if (const char* e = std::getenv("VR_ATLAS_TILES")) {
        const int v = std::atoi(e);
        h->atlas_tiles = v < 1 ? 1 : (v > 8 ? 8 : v);
        h->atlas_forced = true;
    }
    if (const char* e = std::getenv("VR_PLANAR")) h->dev.planar = std::atoi(e) != 0;
    if (const char* e = std::getenv("VR_MULTI")) h->dev.multi_layer = std::atoi(e) != 0;
    if (const char* e = std::getenv("VR_FOLD_UP")) h->dev.fold_upsample = std::atoi(e) != 0;
    if (const char* e = std::getenv("VR_PHASES1")) h->dev.fuse_phases = std::atoi(e) != 0;
    if (const char* e = std::getenv("VR_BLEND_FAST")) h->dev.blend_fast = std::atoi(e) != 0;
    if (cudaSetDevice(cfg->device) != cudaSuccess) return bad("cudaSetDevice failed", VR_E_CUDA);
    {
        // memory the activations of one atlas group may take: the tile atlas keeps a whole group's activations resident (the
        // reference's `--tile-size` bounds memory per tile), so frames whose full atlas would not fit are run as several groups
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) h->mem_budget = free_b / 10 * 8;
        if (const char* e = std::getenv("VR_MEM_BUDGET_MB")) h->mem_budget = static_cast<size_t>(std::atoll(e)) << 20;
    }
    if (cudaStreamCreateWithFlags(&h->dev.stream, cudaStreamNonBlocking) != cudaSuccess)
        return bad("cudaStreamCreate failed", VR_E_CUDA);
    *out = h.release();
    return VR_OK;
}

void vr_destroy(vr_handle* h) {
    if (!h) return;
    cudaSetDevice(h->dev.ordinal);
    if (h->dev.stream) cudaStreamSynchronize(h->dev.stream);
    for (auto& kv : h->layers) free_conv_weights(&kv.second);
    DevBuf* bufs[] = {&h->in32, &h->feat, &h->trunk, &h->rdb[0], &h->rdb[1], &h->rdb[2], &h->up1_in, &h->up1_out,
                      &h->up2_in, &h->up2_out, &h->hr_out, &h->sv[0], &h->sv[1], &h->lr_in, &h->lr_dn, &h->hr[0],
                      &h->hr[1], &h->prev_up, &h->stage_in, &h->stage_out, &h->clahe_hist, &h->clahe_lut};
    for (DevBuf* b : bufs) release(*b);
    for (auto& b : h->tile_out) release(b);
    free_blend_state(h->blend);
    for (auto& b : h->bfree) release(b);
    for (int i = 0; i < 2; ++i) {
        release(h->pipe_in[i]);
        release(h->pipe_out[i]);
        if (h->ev_in[i]) cudaEventDestroy(h->ev_in[i]);
        if (h->ev_comp[i]) cudaEventDestroy(h->ev_comp[i]);
        if (h->ev_out[i]) cudaEventDestroy(h->ev_out[i]);
    }
    if (h->s_in) cudaStreamDestroy(h->s_in);
    if (h->s_out) cudaStreamDestroy(h->s_out);
    if (h->dev.dep_buf) cudaFree(h->dev.dep_buf);
    if (h->dev.bil_tab) cudaFree(h->dev.bil_tab);
    for (auto& f : h->ev_ring) {
        cudaEventDestroy(f.t0);
        cudaEventDestroy(f.c0);
        cudaEventDestroy(f.c1);
        cudaEventDestroy(f.t1);
    }
    if (h->dev.stream) cudaStreamDestroy(h->dev.stream);
    delete h;
}

const char* vr_last_error(const vr_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int vr_load_tensor(vr_handle* h, const char* name, const float* data, const int64_t* shape, int32_t ndim) {
    if (!h) return VR_E_INVALID;
    if (!name || !data || !shape || ndim < 1 || ndim > 4) return fail(h, VR_E_INVALID, "vr_load_tensor: bad arguments");
    RawTensor t;
    size_t n = 1;
    for (int i = 0; i < ndim; ++i) {
        if (shape[i] <= 0) return fail(h, VR_E_INVALID, "vr_load_tensor: non-positive dimension");
        t.shape.push_back(shape[i]);
        n *= static_cast<size_t>(shape[i]);
    }
    t.data.assign(data, data + n);
    h->raw[name] = std::move(t);
    h->committed = false;
    return VR_OK;
}

int vr_commit_weights(vr_handle* h) {
    if (!h) return VR_E_INVALID;
    VR_CUDA_CHECK(cudaSetDevice(h->dev.ordinal), h->dev.err);
    for (auto& kv : h->layers) free_conv_weights(&kv.second);
    h->layers.clear();
    // expected layer list: (prefix, cin, cout, prelu key or "")
    struct Spec {
        std::string name;
        int cin, cout;
        std::string prelu;
    };
    std::vector<Spec> specs;
    const vr_config& c = h->cfg;
    if (c.model_kind == VR_MODEL_RRDBNET) {
        specs.push_back({"conv_first", c.scale == 2 ? 12 : 3, 64, ""});
        for (int b = 0; b < c.num_block; ++b)
            for (int r = 1; r <= 3; ++r)
                for (int k = 1; k <= 5; ++k)
                    specs.push_back({"body." + std::to_string(b) + ".rdb" + std::to_string(r) + ".conv" + std::to_string(k),
                                     64 + 32 * (k - 1), k < 5 ? 32 : 64, ""});
        for (const char* n : {"conv_body", "conv_up1", "conv_up2", "conv_hr"}) specs.push_back({n, 64, 64, ""});
        specs.push_back({"conv_last", 64, 3, ""});
    } else {
        specs.push_back({"body.0", 3, 64, "body.1.weight"});
        for (int i = 0; i < c.num_conv; ++i)
            specs.push_back({"body." + std::to_string(2 * (i + 1)), 64, 64, "body." + std::to_string(2 * (i + 1) + 1) + ".weight"});
        specs.push_back({"body." + std::to_string(2 * (c.num_conv + 1)), 64, 48, ""});
    }
    for (const Spec& s : specs) {
        auto wi = h->raw.find(s.name + ".weight");
        auto bi = h->raw.find(s.name + ".bias");
        if (wi == h->raw.end() || bi == h->raw.end()) return fail(h, VR_E_STATE, "missing tensor " + s.name + ".weight/.bias");
        const RawTensor& w = wi->second;
        if (w.shape.size() != 4 || w.shape[0] != s.cout || w.shape[1] != s.cin || w.shape[2] != 3 || w.shape[3] != 3)
            return fail(h, VR_E_INVALID, "tensor " + s.name + ".weight has the wrong shape");
        if (bi->second.data.size() != static_cast<size_t>(s.cout))
            return fail(h, VR_E_INVALID, "tensor " + s.name + ".bias has the wrong shape");
        const float* prelu = nullptr;
        if (!s.prelu.empty()) {
            auto pi = h->raw.find(s.prelu);
            if (pi == h->raw.end() || pi->second.data.size() != static_cast<size_t>(s.cout))
                return fail(h, VR_E_STATE, "missing/ill-shaped PReLU tensor " + s.prelu);
            prelu = pi->second.data.data();
        }
        ConvWeights cw;
        VR_TRY(pack_conv_weights(h->dev, w.data.data(), bi->second.data.data(), prelu, s.cin, s.cout, &cw));
        h->layers[s.name] = cw;
        if (s.name == "conv_up1" || s.name == "conv_up2") {
            // conv3x3(nearest_x2(f)) at output (2y+py, 2x+px) reads f rows {y-1, y} (py = 0) or {y, y+1} (py = 1): the
            // three original row taps collapse onto two source rows, E_py[0] = W[0], E_py[1] = W[1] + W[2] (py = 0) /
            // E_py[1] = W[0] + W[1], E_py[2] = W[2] (py = 1); same for columns. Summed in fp32, rounded to fp16 once.
            for (int ph = 0; ph < 4; ++ph) {
                const int py = ph >> 1, pxp = ph & 1;
                std::vector<float> e(w.data.size(), 0.f);
                auto tgt = [](int p, int t) { return p == 0 ? (t == 0 ? 0 : 1) : (t == 2 ? 2 : 1); };
                for (size_t oc = 0; oc < static_cast<size_t>(s.cout) * s.cin; ++oc)
                    for (int dy = 0; dy < 3; ++dy)
                        for (int dx = 0; dx < 3; ++dx)
                            e[oc * 9 + tgt(py, dy) * 3 + tgt(pxp, dx)] += w.data[oc * 9 + dy * 3 + dx];
                ConvWeights pw;
                VR_TRY(pack_conv_weights(h->dev, e.data(), bi->second.data.data(), nullptr, s.cin, s.cout, &pw));
                h->layers[s.name + ".phase" + std::to_string(ph)] = pw;
            }
        }
    }
    h->raw.clear();
    h->committed = true;
    return VR_OK;
}

int vr_restore_device_async(vr_handle* h, const uint8_t* d_bgr, int32_t H, int32_t W, int64_t stride, uint8_t* d_out,
                            int64_t out_stride, const vr_frame_opts* opts) {
    if (!h) return VR_E_INVALID;
    return restore_enqueue(h, d_bgr, H, W, stride, d_out, out_stride, opts);
}

int vr_sync(vr_handle* h) {
    if (!h) return VR_E_INVALID;
    return finish_timing(h);
}

void* vr_stream(vr_handle* h) { return h ? static_cast<void*>(h->dev.stream) : nullptr; }

int vr_restore_device(vr_handle* h, const uint8_t* d_bgr, int32_t H, int32_t W, int64_t stride, uint8_t* d_out,
                      int64_t out_stride, const vr_frame_opts* opts) {
    if (!h) return VR_E_INVALID;
    VR_TRY(restore_enqueue(h, d_bgr, H, W, stride, d_out, out_stride, opts));
    return finish_timing(h);
}

int vr_restore(vr_handle* h, const uint8_t* bgr, int32_t H, int32_t W, int64_t stride, uint8_t* out,
               int64_t out_stride, const vr_frame_opts* opts) {
    if (!h) return VR_E_INVALID;
    if (!bgr || !out || H <= 0 || W <= 0) return fail(h, VR_E_INVALID, "vr_restore: bad frame arguments");
    VR_CUDA_CHECK(cudaSetDevice(h->dev.ordinal), h->dev.err);
    const int s = h->cfg.scale;
    const size_t in_row = static_cast<size_t>(W) * 3, out_row = static_cast<size_t>(W) * s * 3;
    VR_TRY(ensure(h, h->stage_in, in_row * H));
    VR_TRY(ensure(h, h->stage_out, out_row * H * s));
    VR_CUDA_CHECK(cudaMemcpy2DAsync(h->stage_in.p, in_row, bgr, stride, in_row, H, cudaMemcpyHostToDevice, h->dev.stream),
                  h->dev.err);
    VR_TRY(restore_enqueue(h, static_cast<const uint8_t*>(h->stage_in.p), H, W, in_row,
                           static_cast<uint8_t*>(h->stage_out.p), out_row, opts));
    VR_CUDA_CHECK(cudaMemcpy2DAsync(out, out_stride, h->stage_out.p, out_row, out_row, static_cast<size_t>(H) * s,
                                    cudaMemcpyDeviceToHost, h->dev.stream),
                  h->dev.err);
    return finish_timing(h);
}

int vr_submit(vr_handle* h, const uint8_t* bgr, int32_t H, int32_t W, int64_t stride, uint8_t* out, int64_t out_stride,
              const vr_frame_opts* opts, int64_t* ticket) {
    if (!h) return VR_E_INVALID;
    if (!bgr || !out || H <= 0 || W <= 0 || !ticket) return fail(h, VR_E_INVALID, "vr_submit: bad arguments");
    VR_CUDA_CHECK(cudaSetDevice(h->dev.ordinal), h->dev.err);
    if (!h->s_in) {
        VR_CUDA_CHECK(cudaStreamCreateWithFlags(&h->s_in, cudaStreamNonBlocking), h->dev.err);
        VR_CUDA_CHECK(cudaStreamCreateWithFlags(&h->s_out, cudaStreamNonBlocking), h->dev.err);
        for (int i = 0; i < 2; ++i) {
            cudaEventCreateWithFlags(&h->ev_in[i], cudaEventDisableTiming);
            cudaEventCreateWithFlags(&h->ev_comp[i], cudaEventDisableTiming);
            cudaEventCreateWithFlags(&h->ev_out[i], cudaEventDisableTiming);
        }
    }
    const int slot = static_cast<int>(h->submitted & 1);
    if (h->submitted >= 2) VR_CUDA_CHECK(cudaEventSynchronize(h->ev_out[slot]), h->dev.err);  // slot's previous frame left
    const int s = h->cfg.scale;
    const size_t in_row = static_cast<size_t>(W) * 3, out_row = static_cast<size_t>(W) * s * 3;
    if (h->pipe_in[slot].bytes < in_row * H || h->pipe_out[slot].bytes < out_row * H * s) {
        VR_CUDA_CHECK(cudaDeviceSynchronize(), h->dev.err);  // growing a slot: nothing may still use the old one
        VR_TRY(ensure(h, h->pipe_in[slot], in_row * H));
        VR_TRY(ensure(h, h->pipe_out[slot], out_row * H * s));
    }
    VR_CUDA_CHECK(cudaMemcpy2DAsync(h->pipe_in[slot].p, in_row, bgr, stride, in_row, H, cudaMemcpyHostToDevice, h->s_in),
                  h->dev.err);
    VR_CUDA_CHECK(cudaEventRecord(h->ev_in[slot], h->s_in), h->dev.err);
    VR_CUDA_CHECK(cudaStreamWaitEvent(h->dev.stream, h->ev_in[slot], 0), h->dev.err);
    VR_TRY(restore_enqueue(h, static_cast<const uint8_t*>(h->pipe_in[slot].p), H, W, in_row,
                           static_cast<uint8_t*>(h->pipe_out[slot].p), out_row, opts));
    VR_CUDA_CHECK(cudaEventRecord(h->ev_comp[slot], h->dev.stream), h->dev.err);
    VR_CUDA_CHECK(cudaStreamWaitEvent(h->s_out, h->ev_comp[slot], 0), h->dev.err);
    VR_CUDA_CHECK(cudaMemcpy2DAsync(out, out_stride, h->pipe_out[slot].p, out_row, out_row, static_cast<size_t>(H) * s,
                                    cudaMemcpyDeviceToHost, h->s_out),
                  h->dev.err);
    VR_CUDA_CHECK(cudaEventRecord(h->ev_out[slot], h->s_out), h->dev.err);
    *ticket = h->submitted++;
    return VR_OK;
}

int vr_wait(vr_handle* h, int64_t ticket) {
    if (!h) return VR_E_INVALID;
    if (ticket < 0 || ticket >= h->submitted) return fail(h, VR_E_INVALID, "vr_wait: unknown ticket");
    if (ticket + 2 < h->submitted) return VR_OK;  // its slot has been reused: that submit already waited for it
    VR_CUDA_CHECK(cudaSetDevice(h->dev.ordinal), h->dev.err);
    VR_CUDA_CHECK(cudaEventSynchronize(h->ev_out[ticket & 1]), h->dev.err);
    return VR_OK;
}

void* vr_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    return p;
}

void vr_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

int vr_temporal_reset(vr_handle* h) {
    if (!h) return VR_E_INVALID;
    h->has_prev = false;
    return VR_OK;
}

int vr_temporal_set_prev(vr_handle* h, const uint8_t* up_prev, int32_t sH, int32_t sW, int64_t stride, int32_t is_device) {
    if (!h || !up_prev || sH <= 0 || sW <= 0) return VR_E_INVALID;
    VR_CUDA_CHECK(cudaSetDevice(h->dev.ordinal), h->dev.err);
    const size_t row = static_cast<size_t>(sW) * 3;
    VR_TRY(ensure(h, h->prev_up, row * sH));
    VR_CUDA_CHECK(cudaMemcpy2DAsync(h->prev_up.p, row, up_prev, stride, row, sH,
                                    is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, h->dev.stream),
                  h->dev.err);
    VR_CUDA_CHECK(cudaStreamSynchronize(h->dev.stream), h->dev.err);
    h->has_prev = true;
    h->prev_h = sH;
    h->prev_w = sW;
    return VR_OK;
}

int vr_temporal_get_prev(vr_handle* h, uint8_t* dst, int32_t sH, int32_t sW, int64_t stride, int32_t is_device) {
    if (!h || !dst) return VR_E_INVALID;
    if (!h->has_prev || h->prev_h != sH || h->prev_w != sW) return fail(h, VR_E_STATE, "no previous frame of that size");
    VR_CUDA_CHECK(cudaSetDevice(h->dev.ordinal), h->dev.err);
    const size_t row = static_cast<size_t>(sW) * 3;
    VR_CUDA_CHECK(cudaMemcpy2DAsync(dst, stride, h->prev_up.p, row, row, sH,
                                    is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, h->dev.stream),
                  h->dev.err);
    VR_CUDA_CHECK(cudaStreamSynchronize(h->dev.stream), h->dev.err);
    return VR_OK;
}

int vr_tile_grid(int32_t H, int32_t W, int32_t tile, int32_t tile_pad, int32_t scale, int32_t* table,
                 int32_t max_tiles) {
    if (H <= 0 || W <= 0 || tile <= 0 || tile_pad < 0 || scale <= 0) return VR_E_INVALID;
    const std::vector<TileRect> g = make_tile_grid(H, W, tile, tile_pad, scale);
    if (table) {
        const int n = std::min<int>(static_cast<int>(g.size()), max_tiles);
        for (int i = 0; i < n; ++i) std::memcpy(table + i * 12, &g[i], 12 * sizeof(int32_t));
    }
    return static_cast<int>(g.size());
}

int64_t vr_launch_count(const vr_handle* h) { return h ? h->dev.launches : 0; }
int64_t vr_conv_launch_count(const vr_handle* h) { return h ? h->dev.conv_launches : 0; }

// Test hook: one of the network's intermediate feature tensors of the LAST restored frame, as fp32 [Ha][Wa][64] over the
// whole tile atlas (a single-tile frame's atlas is the padded tile itself). RRDBNet: "feat" = conv_first output, "body" =
// output of the last RRDB (x slot of the first dense-block buffer), "trunk" = feat + conv_body(body). SRVGG: "body" = input
// of the last conv. Feature-level parity needs this: with default-init weights 345 of x4plus's 351 convs barely move the
// 8-bit frame, so a wrong layer deep in the body would hide under the two x0.2 residual scalings.
int vr_debug_activation(vr_handle* h, const char* which, float* out, int64_t capacity, int32_t* Ha, int32_t* Wa, int32_t* C) {
    if (!h || !which) return VR_E_INVALID;
    if (h->atlas_h <= 0 || h->atlas_w <= 0) return fail(h, VR_E_STATE, "vr_debug_activation: no frame has been restored yet");
    VR_CUDA_CHECK(cudaSetDevice(h->dev.ordinal), h->dev.err);
    const std::string w(which);
    const DevBuf* b = nullptr;
    if (h->cfg.model_kind == VR_MODEL_RRDBNET) {
        if (w == "feat") b = &h->feat;
        else if (w == "body") b = &h->rdb[0];
        else if (w == "trunk") b = &h->trunk;
    } else if (w == "body") {
        b = &h->sv[h->cfg.num_conv & 1];  // body.0 writes sv[0]; every body conv swaps the two buffers
    }
    if (!b || !b->p) return fail(h, VR_E_INVALID, "vr_debug_activation: unknown tensor '" + w + "'");
    const size_t px = static_cast<size_t>(h->atlas_h) * h->atlas_w;
    if (Ha) *Ha = h->atlas_h;
    if (Wa) *Wa = h->atlas_w;
    if (C) *C = 64;
    if (!out) return VR_OK;
    if (capacity < static_cast<int64_t>(px * 64)) return fail(h, VR_E_INVALID, "vr_debug_activation: output buffer too small");
    VR_CUDA_CHECK(cudaStreamSynchronize(h->dev.stream), h->dev.err);
    std::vector<__half> tmp(px * 64);
    const bool planar = h->dev.planar;
    if (planar) {  // two planes of 32 channels, px * 32 elements apart (the first two planes of a 6-plane dense-block buffer)
        VR_CUDA_CHECK(cudaMemcpy(tmp.data(), b->p, px * 64 * sizeof(__half), cudaMemcpyDeviceToHost), h->dev.err);
        for (size_t p = 0; p < px; ++p)
            for (int c = 0; c < 64; ++c) out[p * 64 + c] = __half2float(tmp[(static_cast<size_t>(c >> 5) * px + p) * 32 + (c & 31)]);
    } else {
        const int cs = (b == &h->rdb[0]) ? 192 : 64;
        std::vector<__half> t2(px * cs);
        VR_CUDA_CHECK(cudaMemcpy(t2.data(), b->p, px * cs * sizeof(__half), cudaMemcpyDeviceToHost), h->dev.err);
        for (size_t p = 0; p < px; ++p)
            for (int c = 0; c < 64; ++c) out[p * 64 + c] = __half2float(t2[p * cs + c]);
    }
    return VR_OK;
}

int vr_last_timing(const vr_handle* h, float* total_ms, float* conv_ms) {
    if (!h || !h->timing_valid) return VR_E_STATE;
    if (total_ms) *total_ms = h->last_total_ms;
    if (conv_ms) *conv_ms = h->last_conv_ms;
    return VR_OK;
}

int32_t vr_last_timing_frames(const vr_handle* h) { return (h && h->timing_valid) ? h->last_frames : 0; }

// Stand-alone temporal blend on DEVICE frames, enqueued on the handle's stream: finishes a frame-range shard's head frame
// when the left neighbour's last un-blended frame has arrived peer-to-peer (no host staging).
// ---- boundary frame between two handles of ONE process (pipeline.py: one host thread + handle per GPU) ----
// vr_boundary_send: src's temporal state (its last un-blended upscaled frame) -> a device buffer on dst's device by ONE
// cudaMemcpyPeerAsync (NVLink when the devices are peers; a plain device-to-device copy on one device), synchronised before
// returning. Called by src's thread; only dst's boundary free list is touched (under its mutex), never dst's stream.
int vr_boundary_send(vr_handle* src, vr_handle* dst, int32_t sH, int32_t sW, void** d_frame) {
    if (!src || !dst || !d_frame || sH <= 0 || sW <= 0) return VR_E_INVALID;
    if (!src->has_prev || src->prev_h != sH || src->prev_w != sW) return fail(src, VR_E_STATE, "vr_boundary_send: no previous frame of that size");
    const size_t bytes = static_cast<size_t>(sH) * sW * 3;
    DevBuf b;
    {
        std::lock_guard<std::mutex> lk(dst->bmu);
        for (size_t i = 0; i < dst->bfree.size(); ++i)
            if (dst->bfree[i].bytes >= bytes) {
                b = dst->bfree[i];
                dst->bfree.erase(dst->bfree.begin() + i);
                break;
            }
    }
    if (!b.p) {
        VR_CUDA_CHECK(cudaSetDevice(dst->dev.ordinal), src->dev.err);
        VR_CUDA_CHECK(cudaMalloc(&b.p, bytes), src->dev.err);
        b.bytes = bytes;
    }
    VR_CUDA_CHECK(cudaSetDevice(src->dev.ordinal), src->dev.err);
    if (src->dev.ordinal != dst->dev.ordinal) {
        int can = 0;
        cudaDeviceCanAccessPeer(&can, src->dev.ordinal, dst->dev.ordinal);
        if (can) {
            const cudaError_t pe = cudaDeviceEnablePeerAccess(dst->dev.ordinal, 0);
            if (pe != cudaSuccess) cudaGetLastError();  // already enabled
        }
        VR_CUDA_CHECK(cudaMemcpyPeerAsync(b.p, dst->dev.ordinal, src->prev_up.p, src->dev.ordinal, bytes, src->dev.stream), src->dev.err);
    } else {
        VR_CUDA_CHECK(cudaMemcpyAsync(b.p, src->prev_up.p, bytes, cudaMemcpyDeviceToDevice, src->dev.stream), src->dev.err);
    }
    VR_CUDA_CHECK(cudaStreamSynchronize(src->dev.stream), src->dev.err);
    *d_frame = b.p;
    return VR_OK;
}

// vr_boundary_finish: blends a chunk's head frame (host, un-blended) with the boundary frame received by vr_boundary_send and
// writes the result to `out` (host); the boundary buffer goes back to the free list. head == NULL only recycles the buffer.
int vr_boundary_finish(vr_handle* h, void* d_prev, const uint8_t* head, int64_t head_stride, uint8_t* out, int64_t out_stride,
                       int32_t sH, int32_t sW, float alpha, float tau) {
    if (!h || !d_prev) return VR_E_INVALID;
    const size_t row = static_cast<size_t>(sW) * 3, bytes = row * sH;
    int rc = VR_OK;
    if (head) {
        if (!out || sH <= 0 || sW <= 0) return fail(h, VR_E_INVALID, "vr_boundary_finish: bad arguments");
        rc = [&]() -> int {
            VR_CUDA_CHECK(cudaSetDevice(h->dev.ordinal), h->dev.err);
            VR_TRY(ensure(h, h->hr[0], bytes));
            VR_TRY(ensure(h, h->hr[1], bytes));
            VR_CUDA_CHECK(cudaMemcpy2DAsync(h->hr[0].p, row, head, head_stride, row, sH, cudaMemcpyHostToDevice, h->dev.stream), h->dev.err);
            VR_TRY(launch_temporal(h->dev, static_cast<const uint8_t*>(h->hr[0].p), row, static_cast<const uint8_t*>(d_prev), row, sH, sW,
                                   static_cast<uint8_t*>(h->hr[1].p), row, alpha > 0 ? alpha : 0.2f, tau > 0 ? tau : 12.f));
            VR_CUDA_CHECK(cudaMemcpy2DAsync(out, out_stride, h->hr[1].p, row, row, sH, cudaMemcpyDeviceToHost, h->dev.stream), h->dev.err);
            VR_CUDA_CHECK(cudaStreamSynchronize(h->dev.stream), h->dev.err);
            return VR_OK;
        }();
    }
    DevBuf b;
    b.p = d_prev;
    b.bytes = bytes;
    std::lock_guard<std::mutex> lk(h->bmu);
    h->bfree.push_back(b);
    return rc;
}

int vr_temporal_device(vr_handle* h, const uint8_t* d_cur, const uint8_t* d_prev, int32_t sH, int32_t sW, uint8_t* d_out,
                       float alpha, float tau) {
    if (!h) return VR_E_INVALID;
    if (!d_cur || !d_prev || !d_out || sH <= 0 || sW <= 0) return fail(h, VR_E_INVALID, "vr_temporal_device: bad arguments");
    VR_CUDA_CHECK(cudaSetDevice(h->dev.ordinal), h->dev.err);
    const int64_t st = static_cast<int64_t>(sW) * 3;
    return launch_temporal(h->dev, d_cur, st, d_prev, st, sH, sW, d_out, st, alpha > 0 ? alpha : 0.2f, tau > 0 ? tau : 12.f);
}

}  // extern "C"

// -------------------------------------------------------------------------------------------------
// stand-alone filter entry points (host buffers): thin staging around the same kernels
// -------------------------------------------------------------------------------------------------
namespace {
struct TmpDev {
    Device dev;
    std::string err;
    bool ok = false;
    explicit TmpDev(int ordinal) {
        dev.err = &err;
        dev.ordinal = ordinal;
        cudaDeviceProp p;
        if (cudaSetDevice(ordinal) != cudaSuccess || cudaGetDeviceProperties(&p, ordinal) != cudaSuccess) {
            set_error(&err, "no usable CUDA device (this library has no CPU fallback)");
            return;
        }
        dev.sm_count = p.multiProcessorCount;
        if (cudaStreamCreateWithFlags(&dev.stream, cudaStreamNonBlocking) != cudaSuccess) {
            set_error(&err, "cudaStreamCreate failed");
            return;
        }
        ok = true;
    }
    ~TmpDev() {
        if (dev.bil_tab) cudaFree(dev.bil_tab);
        if (dev.stream) cudaStreamDestroy(dev.stream);
    }
};
template <typename F>
int run_filter(int device, const uint8_t* src, const uint8_t* src2, int H, int W, uint8_t* dst, F&& f) {
    if (!src || !dst || H <= 0 || W <= 0) {
        set_error(nullptr, "filter: bad arguments");
        return VR_E_INVALID;
    }
    TmpDev td(device);
    if (!td.ok) return VR_E_NODEVICE;
    const size_t bytes = static_cast<size_t>(H) * W * 3;
    uint8_t *d_src = nullptr, *d_src2 = nullptr, *d_dst = nullptr;
    VR_CUDA_CHECK(cudaMalloc(&d_src, bytes), td.dev.err);
    VR_CUDA_CHECK(cudaMalloc(&d_dst, bytes), td.dev.err);
    VR_CUDA_CHECK(cudaMemcpy(d_src, src, bytes, cudaMemcpyHostToDevice), td.dev.err);
    if (src2) {
        VR_CUDA_CHECK(cudaMalloc(&d_src2, bytes), td.dev.err);
        VR_CUDA_CHECK(cudaMemcpy(d_src2, src2, bytes, cudaMemcpyHostToDevice), td.dev.err);
    }
    int rc = f(td.dev, d_src, d_src2, d_dst);
    if (rc == 0) {
        cudaError_t e = cudaStreamSynchronize(td.dev.stream);
        if (e != cudaSuccess) {
            set_error(td.dev.err, std::string("filter kernel failed: ") + cudaGetErrorString(e));
            rc = VR_E_CUDA;
        } else {
            cudaMemcpy(dst, d_dst, bytes, cudaMemcpyDeviceToHost);
        }
    }
    cudaFree(d_src);
    cudaFree(d_dst);
    if (d_src2) cudaFree(d_src2);
    return rc;
}
}  // namespace

extern "C" {

int vr_bilateral(int32_t device, const uint8_t* src, int32_t H, int32_t W, uint8_t* dst, int32_t d, float sigma_color,
                 float sigma_space) {
    const int64_t st = static_cast<int64_t>(W) * 3;
    return run_filter(device, src, nullptr, H, W, dst, [&](Device& dev, uint8_t* s, uint8_t*, uint8_t* o) {
        return launch_bilateral(dev, s, st, H, W, o, st, d, sigma_color, sigma_space);
    });
}

int vr_unsharp(int32_t device, const uint8_t* src, int32_t H, int32_t W, uint8_t* dst, float amount) {
    const int64_t st = static_cast<int64_t>(W) * 3;
    return run_filter(device, src, nullptr, H, W, dst, [&](Device& dev, uint8_t* s, uint8_t*, uint8_t* o) {
        return launch_unsharp(dev, s, st, H, W, o, st, amount);
    });
}

int vr_clahe(int32_t device, const uint8_t* src, int32_t H, int32_t W, uint8_t* dst, float clip, int32_t grid,
             int32_t* hist_out, uint8_t* lut_out) {
    const int64_t st = static_cast<int64_t>(W) * 3;
    return run_filter(device, src, nullptr, H, W, dst, [&](Device& dev, uint8_t* s, uint8_t*, uint8_t* o) {
        int32_t* d_hist = nullptr;
        uint8_t* d_lut = nullptr;
        const size_t n = static_cast<size_t>(grid) * grid * 256;
        if (grid < 1 || grid > 16) {
            set_error(dev.err, "clahe: grid must be 1..16");
            return VR_E_INVALID;
        }
        VR_CUDA_CHECK(cudaMalloc(&d_hist, n * 4), dev.err);
        VR_CUDA_CHECK(cudaMalloc(&d_lut, n), dev.err);
        int rc = launch_clahe(dev, s, st, H, W, o, st, clip, grid, d_hist, d_lut, nullptr);
        if (rc == 0 && cudaStreamSynchronize(dev.stream) == cudaSuccess) {
            if (hist_out) cudaMemcpy(hist_out, d_hist, n * 4, cudaMemcpyDeviceToHost);
            if (lut_out) cudaMemcpy(lut_out, d_lut, n, cudaMemcpyDeviceToHost);
        }
        cudaFree(d_hist);
        cudaFree(d_lut);
        return rc;
    });
}

int vr_temporal(int32_t device, const uint8_t* cur, const uint8_t* prev, int32_t H, int32_t W, uint8_t* dst,
                float alpha, float tau) {
    if (!prev) {
        set_error(nullptr, "vr_temporal: prev is null");
        return VR_E_INVALID;
    }
    const int64_t st = static_cast<int64_t>(W) * 3;
    return run_filter(device, cur, prev, H, W, dst, [&](Device& dev, uint8_t* s, uint8_t* p, uint8_t* o) {
        return launch_temporal(dev, s, st, p, st, H, W, o, st, alpha, tau);
    });
}

int vr_blend_weights(int32_t device, int32_t extent, float* w_out) {
    if (extent <= 0 || !w_out) return VR_E_INVALID;
    TmpDev td(device);
    if (!td.ok) return VR_E_NODEVICE;
    float* d = nullptr;
    VR_CUDA_CHECK(cudaMalloc(&d, extent * sizeof(float)), td.dev.err);
    int rc = launch_blend_weights(td.dev, extent, d);
    if (rc == 0) {
        if (cudaStreamSynchronize(td.dev.stream) != cudaSuccess) rc = VR_E_CUDA;
        else cudaMemcpy(w_out, d, extent * sizeof(float), cudaMemcpyDeviceToHost);
    }
    cudaFree(d);
    return rc;
}

}  // extern "C"
