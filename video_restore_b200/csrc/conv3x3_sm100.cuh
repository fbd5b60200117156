// K1: 3x3 / stride 1 / zero-pad 1 convolution as an implicit GEMM on tcgen05 tensor cores (sm_100a).
//
// Replaces every torch Conv2d(3x3) the reference dispatches to cuDNN inside RRDBNet.forward /
// SRVGGNetCompact.forward (third-party basicsr/realesrgan, constructed at reference
// video_upscaler.py:313-321 and run from RealESRGANer.enhance, video_upscaler.py:501).
//
// Formulation (per CTA tile = TH output rows x 128 output pixels, all Cout channels):
//   D[r][128 px, N] += A[(r+dy) row, px+dx][128, 16] * W[dy,dx][16, N]     for 9 taps x (Cin/16) k-steps
//   * activations NHWC fp16 in HBM; one TMA box (32 ch, 130 px, TH+2 rows) per 32-channel chunk lands the
//     haloed input tile in shared memory ONCE (64 B rows, SWIZZLE_64B); TMA out-of-bounds zero fill IS the
//     conv zero padding. The 9 taps are 9 shared-memory descriptors into that one tile (row shift dx, dy).
//   * weights pre-packed on the host into the exact swizzled smem image per chunk: one bulk copy per stage.
//   * accumulators: TH x N fp32 columns of TMEM, double buffered so the epilogue of tile i overlaps the MMAs
//     of tile i+1. One elected thread issues all MMAs. The three dy taps of one input row feed three adjacent
//     output rows, so they are ONE MMA of N = 3*Cout (B rows = weights of dy 2,1,0): A is read once per
//     (input row, dx, k) and N is 96 / 192 instead of 32 / 64 (the SS MMA is smem-read bound below N = 128).
//   * epilogue (8 warps, two groups alternating output rows): tcgen05.ld -> +bias -> LeakyReLU/PReLU ->
//     *s1 + res1 -> *s2 + res2 in fp32 -> one fp16 rounding -> 16 B stores into a channel slice of the
//     destination NHWC buffer (zero-copy concat). Residual rows are prefetched before the TMEM loads.
// Warp roles: 0..7 epilogue (warp % 4 = TMEM lane quarter), 8 = TMA producer, 9 and 10 = MMA issuers (alternating stages).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cstdint>
#include "sm100_ptx.cuh"

namespace vr {

enum ConvAct { ACT_NONE = 0, ACT_LRELU = 1, ACT_PRELU = 2 };
enum ConvOut { OUT_NHWC = 0, OUT_RGB4 = 1, OUT_PS4 = 2 };
// debug ablation flags (ConvArgs::flags): measurement only
constexpr int kMaxLayers = 4;
enum ConvFlags {
    FLAG_SKIP_TMA = 2, FLAG_SKIP_MMA = 4, FLAG_SKIP_EPI = 8, FLAG_SKIP_B = 16, FLAG_SKIP_A = 32,
    FLAG_FORCE_TILE = 64,  // kernel selection (tests): always the tiled kernel K1 ...
    FLAG_FORCE_ROLL = 128,  // ... or fail unless the rolling-row kernel K2 takes the layer
    FLAG_FORCE_PAIR = 512,  // ... or the CTA-pair rolling-row kernel K3
    FLAG_PLANAR = 1024,     // conv test hook: run the layer on chunk-planar tensors
    FLAG_TRACE = 256        // K2: CTA 0's issuers record per-box timestamps into dbg_cycles[256..512) (bench hook prints them)
};

struct ConvArgs {
    int W, H;              // conv input == output extent
    int y_begin, y_end;    // output rows [y_begin, y_end) this launch produces (row-band scheduling)
    int tiles_x, tiles_y;  // ceil(W/128), ceil((y_end - y_begin)/TH)
    int nchunks;           // Cin_padded / KC
    int cin_off;           // first input channel inside the source buffer
    int in_cstride;        // source channels per pixel (per plane); 32 = chunk-planar tensor (see chan_off)
    const __half* wpack;   // [nchunks][dx][dy=2,1,0][N][32] fp16, pre-swizzled smem image
    const float* bias;     // [cout]
    const float* prelu;    // [cout] or null
    int act;
    float slope;
    __half* out;
    int out_cstride, out_coff, cout;
    // optional second destination of the same values (K3 direct epilogue only; same channel offset and plane distance as `out`):
    // conv_first feeds both the trunk skip (`feat`) and the first dense block's x slot from one launch
    __half* out2;
    int out2_cstride;
    // K4 (conv3x3_pair2_sm100.cuh): the NEXT layer of the dense block, run in the same launch two row pairs behind this one.
    // It reads the same source planes plus this layer's output (handed over through shared memory) and writes `out_coff2`.
    const __half* wpack2;  // K3 weight halves of the second layer, nchunks + 1 chunks
    const float* bias2;
    int out_coff2;
    int lag;               // K4: row pairs layer B runs behind layer A (0 = default 2)
    int prefetch;          // K4: steps (row pairs) layer A's rows are prefetched into the L2 ahead of their TMA load (0 = off)
    // Chunk-planar tensors: a tensor with cstride == 32 stores channels [32k, 32k+32) as plane k, [H][W][32] fp16, planes
    // `pstride` elements apart (every TMA box row and every output row is then contiguous; the interleaved [H][W][C] form
    // costs the 32-channel layers 16..52 %). Any other cstride is the interleaved form (pstride unused).
    long long out_pstride, res1_pstride, res2_pstride;
    const __half* res1;
    int res1_cstride, res1_coff;
    float s1;
    const __half* res2;
    int res2_cstride, res2_coff;
    float s2;
    int out_mode;
    const __half* base;  // OUT_PS4: network input (RGB in channels 0..2), added to the 16 sub-pixels
    int base_cstride;
    int flags;  // ConvFlags
    // output pixel mapping (NHWC): output (y, x) is stored at row y*omul + opy, column x*omul + opx of an image omul times
    // as wide. omul = 2 writes one phase of a conv that was folded with a preceding nearest x2 upsample.
    int omul, opy, opx;
    // Tile atlas: several RealESRGANer tiles share one image, separated by zero gap columns / rows (the gap IS the
    // per-tile zero padding). Outputs at gap positions are forced to zero so the separation survives every layer.
    // A position x is a gap iff (x >> gshift) == gx[j] for some j (gshift = log2 of the resolution multiple).
    int ngx, ngy, gshift;
    int gx[7], gy[7];
    // shared-memory plan chosen by the host (conv_launch.cu): weights either stream with the activations (one B block
    // per stage) or, when the whole layer fits, stay resident for the CTA's lifetime and only activations stream.
    int wres;         // 1: weights resident (nchunks * kBStage bytes at the start of smem)
    int nstages;      // pipeline stages in use (<= kMaxStages)
    int stage_bytes;  // kAStage (+ kBStage when streaming weights)
    // Multi-layer launch (nlayers > 1): layers l = 0..nlayers-1 read the SAME source tensor (growing channel prefix of a
    // dense block) and differ only in the fields below; work items are (layer, tile) in layer-major order and a tile of
    // layer l waits until the tile rows of layer l-1 it reads are complete (dep counters, one int per layer and tile row).
    int nlayers;
    int l_nchunks[kMaxLayers];
    const __half* l_wpack[kMaxLayers];
    const float* l_bias[kMaxLayers];
    int l_out_coff[kMaxLayers];
    int* dep;         // [nlayers][tiles_y], zero on entry
    int* dep_zero;    // the other launch parity's region: zeroed by this launch for the next multi-layer launch
    int dep_zero_n;
    // Rolling-row kernel K2 (conv3x3_roll_sm100.cuh): work item = (band of `band` output rows, 128-pixel strip, channel half)
    int band, nbands;
    int nsplit;  // 1, or 2: the layer's 2N output channels are computed as two independent N-channel halves
    int unit;    // K3: boxes per issuer hand-over
    int early64; // K3, 64 output channels (VR_EARLY64): ring position handed back 0 = after the stores, 1 = before them,
                 // 2 (default) = additionally right after the TMEM loads in the instantiation for layers without residuals
    int epi_direct; // K3 (VR_EPI_DIRECT): each lane stores its own pixel's 32 channels with two 256-bit stores, no staging transpose
    int l2_hint; // K3 (VR_L2HINT): 1 = newest source plane and the output evict_last, older source planes evict_first;
                 // 2 / 3 = the first one / two source planes (the dense block's x) evict_last, everything else streams;
                 // 4 / 5 / 6 = the first two / first three / all source planes keep a FRACTION l2_frac of their lines
                 // (evict_last, chosen by address) and stream the rest, so that the kept part fits the L2
    float l2_frac;
    long long* dbg_cycles;  // optional: [0,256) SM cycles per CTA; [256, 496) CTA 0's per-stage issuer timestamps
};

constexpr int round_up_c(int x, int m) { return (x + m - 1) / m * m; }
constexpr int next_pow2_c(int x) { int p = 32; while (p < x) p *= 2; return p; }

constexpr int kEpiWarps = 8;
constexpr int kMaxStages = 6;
constexpr int kMmaWarps = 2;  // two issuers ping-pong pipeline stages (see the MMA section)
constexpr int kConvThreads = (kEpiWarps + 1 + kMmaWarps) * 32;

template <int N, int TH, int KC>
struct ConvTraits {
    static_assert(KC == 16 || KC == 32, "channels per pipeline stage: 16 (SWIZZLE_32B) or 32 (SWIZZLE_64B)");
    static constexpr int kRowBytes = KC * 2;  // one pixel's channel chunk
    static constexpr int kKSteps = KC / 16;   // MMAs (K = 16) per tap and stage
    static constexpr uint32_t kDescHi = KC == 32 ? ptx::kDescHiSw64 : ptx::kDescHiSw32;
    static constexpr int kInRows = TH + 2;
    static constexpr int kPitch = 130;  // 128 output pixels + 1 halo pixel each side
    static constexpr int kCopyBytes = kInRows * kPitch * kRowBytes;  // bytes one TMA box delivers
    static constexpr int kAStage = round_up_c(kCopyBytes, 1024);
    static constexpr int kBBytes = 9 * N * kRowBytes;
    static constexpr int kBStage = round_up_c(kBBytes, 1024);
    static constexpr int kStageBytes = kAStage + kBStage;
    static constexpr int kStatic = 2048;  // static __shared__: barriers, tmem slot, per-layer bias / activation tables
    // per-warp epilogue staging (NHWC outputs only): 32 pixels x N fp16 for the coalescing transpose; 16 B unit U of
    // pixel p lives at unit U ^ swz(p) (conflict-free for the per-pixel writes and the per-row reads, no padding)
    static constexpr int kStgPitch = N * 2;
    static constexpr int kStgWarp = (N % 32 == 0) ? 32 * kStgPitch : 0;
    static constexpr int kStgBytes = kEpiWarps * kStgWarp;
    static constexpr int kBudget = 227 * 1024 - 1024 - kStatic - kStgBytes;
    static constexpr int kStagesRaw = kBudget / kStageBytes;
    static constexpr int kStages = kStagesRaw > 6 ? 6 : (kStagesRaw < 1 ? 1 : kStagesRaw);
    static constexpr int kAccCols = TH * N;
    static constexpr int kTmemCols = next_pow2_c(2 * kAccCols);
    static constexpr int kSmemBytes = kStages * kStageBytes + kStgBytes + 1024;  // + slack for the manual alignment
    static_assert(kStagesRaw >= 1, "stage does not fit in shared memory");
    static_assert(kTmemCols <= 512, "accumulators do not fit in TMEM");
    static_assert(N % 16 == 0 && N >= 16 && N <= 64, "N must be 16..64 step 16");
    static_assert(TH % 2 == 0, "rows alternate between the two epilogue groups");
};

// element offset of channel `ch` of pixel `p`: chunk-planar when cs == 32, interleaved [pixel][cs] otherwise
__device__ __forceinline__ size_t chan_off(size_t p, int cs, long long ps, int ch) {
    return cs == 32 ? static_cast<size_t>(ch >> 5) * static_cast<size_t>(ps) + p * 32 + (ch & 31) : p * static_cast<size_t>(cs) + ch;
}

__device__ __forceinline__ void unpack8(const uint4& q, float* f) {
    const __half2* h = reinterpret_cast<const __half2*>(&q);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float2 t = __half22float2(h[i]);
        f[2 * i] = t.x;
        f[2 * i + 1] = t.y;
    }
}
__device__ __forceinline__ uint4 pack8(const float* f) {
    uint4 q;
    __half2* h = reinterpret_cast<__half2*>(&q);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
    return q;
}

// bias + activation over CNT consecutive channels starting at c0; the mode switch is outside the element loop
template <int CNT>
__device__ __forceinline__ void bias_act(float* v, const float* s_bias, const float* s_neg, int c0, int mode) {
    if (mode == 0) {  // identity
#pragma unroll
        for (int j = 0; j < CNT; ++j) v[j] += s_bias[c0 + j];
    } else if (mode == 1) {  // LeakyReLU with 0 <= slope <= 1: max(t, slope * t)
#pragma unroll
        for (int j = 0; j < CNT; ++j) {
            const float t = v[j] + s_bias[c0 + j];
            v[j] = fmaxf(t, t * s_neg[c0 + j]);
        }
    } else {  // general PReLU / LeakyReLU
#pragma unroll
        for (int j = 0; j < CNT; ++j) {
            const float t = v[j] + s_bias[c0 + j];
            v[j] = fmaxf(t, 0.f) + s_neg[c0 + j] * fminf(t, 0.f);
        }
    }
}

// One output row x 32 pixels (this warp's TMEM lane quarter) of an NHWC layer:
// tcgen05.ld -> +bias -> activation -> *s1 + res1 -> *s2 + res2 (fp32) -> fp16 -> per-warp swizzled staging transpose ->
// coalesced 16 B stores (consecutive lanes write consecutive units of one pixel). `t_addr` = TMEM address of the row's
// first accumulator column in this warp's lane quarter; `coff_add` shifts the output / residual channel slices (the
// second half of a layer computed as two N-channel halves); s_bias / s_neg are already offset to channel 0 of the slice.
// Residual rows are fetched before the TMEM loads: one global-load latency per row instead of one per channel group.
template <int N>
__device__ __forceinline__ void epi_row_nhwc(const ConvArgs& a, uint32_t t_addr, uint32_t stg_s, int lane, int x_base, int y,
                                             bool gap, int out_coff, int coff_add, const float* s_bias, const float* s_neg,
                                             int amode, int opy, int opx) {
    constexpr int kVec = N / 8;  // 16 B units per pixel
    constexpr int kStgPitch = N * 2;
    const int x = x_base + lane;
    const bool inb = x < a.W;
    const size_t p = static_cast<size_t>(y) * a.W + x;
    const bool has1 = a.res1 != nullptr, has2 = a.res2 != nullptr;
    uint4 q1[kVec], q2[kVec];
    if (inb && has1) {
#pragma unroll
        for (int j = 0; j < kVec; ++j)
            q1[j] = __ldg(reinterpret_cast<const uint4*>(
                a.res1 + chan_off(p, a.res1_cstride, a.res1_pstride, a.res1_coff + coff_add + j * 8)));
    }
    if (inb && has2) {
#pragma unroll
        for (int j = 0; j < kVec; ++j)  // may alias `out` (in-place RRDB skip)
            q2[j] = *reinterpret_cast<const uint4*>(
                a.res2 + chan_off(p, a.res2_cstride, a.res2_pstride, a.res2_coff + coff_add + j * 8));
    }
#pragma unroll
    for (int g = 0; g < N / 32; ++g) {
        float v[32];
        ptx::tmem_ld32(t_addr + g * 32, v);
        if (inb) {
            const int c0 = g * 32;
            bias_act<32>(v, s_bias, s_neg, c0, amode);
            if (has1) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    float f[8];
                    unpack8(q1[g * 4 + u], f);
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[u * 8 + j] = fmaf(v[u * 8 + j], a.s1, f[j]);
                }
            }
            if (has2) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    float f[8];
                    unpack8(q2[g * 4 + u], f);
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[u * 8 + j] = fmaf(v[u * 8 + j], a.s2, f[j]);
                }
            }
        }
        if (gap) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0.f;
        }
        // stage this pixel's 32 channels (64 B): unit U of pixel p lives at U ^ swz(p) (conflict-free both ways)
        const int swz_w = kVec == 8 ? (lane & 7) : ((lane >> 1) & 3);
#pragma unroll
        for (int u = 0; u < 4; ++u) ptx::sts128(stg_s + lane * kStgPitch + (((g * 4 + u) ^ swz_w) << 4), pack8(v + u * 8));
    }
    __syncwarp();
    // lane l always handles 16 B unit l % kVec of its pixels (32 % kVec == 0)
    __half* orow = a.out + chan_off(static_cast<size_t>(y * a.omul + opy) * (a.W * a.omul) + x_base * a.omul + opx, a.out_cstride,
                                    a.out_pstride, out_coff + coff_add + (lane % kVec) * 8);
#pragma unroll
    for (int i = 0; i < kVec; ++i) {
        const int idx = i * 32 + lane;
        const int px = idx / kVec, un = idx % kVec;
        if (x_base + px < a.W) {
            const int swz_r = kVec == 8 ? (px & 7) : ((px >> 1) & 3);
            const uint4 val = ptx::lds128(stg_s + px * kStgPitch + ((un ^ swz_r) << 4));
            *reinterpret_cast<uint4*>(orow + static_cast<size_t>(px) * a.omul * a.out_cstride) = val;
        }
    }
    __syncwarp();
}

// K2's variant of epi_row_nhwc. The accumulators already contain the bias (the ring block was initialised with it), all
// TMEM loads of the row are issued before one wait, and the block is re-initialised and released (`release`) BEFORE the
// arithmetic and the stores, so the MMAs of a later row can start while this row is still being written out.
// amode: 0 identity, 1 LeakyReLU with the scalar a.slope in [0, 1], 2 per-channel s_neg table (PReLU).
template <int N>
__device__ __forceinline__ void epi_row_nhwc_folded(const ConvArgs& a, uint32_t t_addr, uint32_t stg_s, int lane, int x_base, int y,
                                                    bool gap, int out_coff, int coff_add, const float* s_bias, const float* s_neg,
                                                    int amode, uint64_t* release, long long* tr = nullptr, long long t_ref = 0) {
    constexpr int kVec = N / 8;  // 16 B units per pixel
    constexpr int kStgPitch = N * 2;
    const int x = x_base + lane;
    const bool inb = x < a.W;
    const size_t p = static_cast<size_t>(y) * a.W + x;
    const bool has1 = a.res1 != nullptr, has2 = a.res2 != nullptr;
    uint4 q1[kVec], q2[kVec];
    if (inb && has1) {
#pragma unroll
        for (int j = 0; j < kVec; ++j)
            q1[j] = __ldg(reinterpret_cast<const uint4*>(
                a.res1 + chan_off(p, a.res1_cstride, a.res1_pstride, a.res1_coff + coff_add + j * 8)));
    }
    if (inb && has2) {
#pragma unroll
        for (int j = 0; j < kVec; ++j)  // may alias `out` (in-place RRDB skip)
            q2[j] = *reinterpret_cast<const uint4*>(
                a.res2 + chan_off(p, a.res2_cstride, a.res2_pstride, a.res2_coff + coff_add + j * 8));
    }
    uint32_t raw[N];
#pragma unroll
    for (int g = 0; g < N / 32; ++g) ptx::tmem_ld32_issue(t_addr + g * 32, raw + g * 32);
#pragma unroll
    for (int g = 0; g < N / 32; ++g) ptx::tmem_ld32_wait(raw + g * 32);
    if (tr) tr[2] = clock64() - t_ref;
    {
        // re-initialise the block with the bias and hand it back
#pragma unroll
        for (int g = 0; g < N / 32; ++g) {
            float bz[32];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float4 t = *reinterpret_cast<const float4*>(s_bias + g * 32 + j * 4);
                bz[j * 4] = t.x; bz[j * 4 + 1] = t.y; bz[j * 4 + 2] = t.z; bz[j * 4 + 3] = t.w;
            }
            ptx::tmem_st32(t_addr + g * 32, bz);
        }
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(release);
    }
    if (tr) tr[3] = clock64() - t_ref;
#pragma unroll
    for (int g = 0; g < N / 32; ++g) {
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[g * 32 + j]);
        if (inb) {
            if (amode == 1) {
                const float sl = a.slope;
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], v[j] * sl);
            } else if (amode == 2) {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f) + s_neg[g * 32 + j] * fminf(v[j], 0.f);
            }
            if (has1) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    float f[8];
                    unpack8(q1[g * 4 + u], f);
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[u * 8 + j] = fmaf(v[u * 8 + j], a.s1, f[j]);
                }
            }
            if (has2) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    float f[8];
                    unpack8(q2[g * 4 + u], f);
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[u * 8 + j] = fmaf(v[u * 8 + j], a.s2, f[j]);
                }
            }
        }
        if (gap) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0.f;
        }
        const int swz_w = kVec == 8 ? (lane & 7) : ((lane >> 1) & 3);
#pragma unroll
        for (int u = 0; u < 4; ++u) ptx::sts128(stg_s + lane * kStgPitch + (((g * 4 + u) ^ swz_w) << 4), pack8(v + u * 8));
    }
    __syncwarp();
    if (tr) tr[4] = clock64() - t_ref;
    // lane l always handles 16 B unit l % kVec of its pixels (32 % kVec == 0)
    __half* orow = a.out + chan_off(static_cast<size_t>(y * a.omul + a.opy) * (a.W * a.omul) + x_base * a.omul + a.opx, a.out_cstride,
                                    a.out_pstride, out_coff + coff_add + (lane % kVec) * 8);
#pragma unroll
    for (int i = 0; i < kVec; ++i) {
        const int idx = i * 32 + lane;
        const int px = idx / kVec, un = idx % kVec;
        if (x_base + px < a.W) {
            const int swz_r = kVec == 8 ? (px & 7) : ((px >> 1) & 3);
            const uint4 val = ptx::lds128(stg_s + px * kStgPitch + ((un ^ swz_r) << 4));
            *reinterpret_cast<uint4*>(orow + static_cast<size_t>(px) * a.omul * a.out_cstride) = val;
        }
    }
    __syncwarp();
}

// The MMAs of one pipeline stage (one 32- / 16-channel chunk of the haloed input tile) for the tap window dy in [kDy0, kDy1],
// dx in [kDx0, kDx1]: dy-stacked N as described in the kernel. `first_chunk`: the lowest present tap of chunk 0 is the first
// touch of an output row in this tile and must overwrite instead of accumulate.
template <int N, int TH, int KC, int kDy0, int kDy1, int kDx0, int kDx1>
__device__ __forceinline__ void conv_issue_stage(uint32_t a_lo0, uint32_t b_lo0, uint32_t d_base, bool first_chunk) {
    using T = ConvTraits<N, TH, KC>;
#pragma unroll
    for (int dx = kDx0; dx <= kDx1; ++dx) {
#pragma unroll
        for (int k = 0; k < T::kKSteps; ++k) {
#pragma unroll
            for (int rho = 0; rho < T::kInRows; ++rho) {
                const uint32_t a_lo = a_lo0 + (((rho * T::kPitch + dx) * T::kRowBytes + k * 32) >> 4);
                constexpr int kLast = TH - 1;
                // taps dy in [dy_lo, dy_hi] of this input row land in output rows rho - dy
                const int lo_r = rho - kLast > 0 ? rho - kLast : 0;
                const int hi_r = rho < 2 ? rho : 2;
                const int dy_lo = lo_r > kDy0 ? lo_r : kDy0;
                const int dy_hi = hi_r < kDy1 ? hi_r : kDy1;
                if (dy_lo > dy_hi) continue;  // this input row feeds no present tap
                const int nblk = dy_hi - dy_lo + 1;
                const int r_lo = rho - dy_hi;
                // B rows of this dx: [dy=2 | dy=1 | dy=0] x N
                const uint32_t b_lo = b_lo0 + ((((dx * 3 + (2 - dy_hi)) * N) * T::kRowBytes + k * 32) >> 4);
                const uint32_t d = d_base + r_lo * N;
                if (dx == kDx0 && k == 0 && dy_lo == kDy0 && first_chunk) {
                    // the lowest present tap is the first touch of output row rho - kDy0 in this
                    // tile: it must overwrite while the other blocks accumulate -> split the MMA
                    if (nblk > 1)
                        ptx::umma_f16<ptx::kCollNone>(d, a_lo, T::kDescHi, b_lo, T::kDescHi,
                                                      ptx::make_idesc_f16(128, (nblk > 1 ? nblk - 1 : 1) * N), 1u);
                    ptx::umma_f16<ptx::kCollNone>(d + (nblk - 1) * N, a_lo, T::kDescHi,
                                                  b_lo + (((nblk - 1) * N * T::kRowBytes) >> 4), T::kDescHi,
                                                  ptx::make_idesc_f16(128, N), 0u);
                } else {
                    ptx::umma_f16<ptx::kCollNone>(d, a_lo, T::kDescHi, b_lo, T::kDescHi, ptx::make_idesc_f16(128, nblk * N), 1u);
                }
            }
        }
    }
}

// DYS / DXS select the taps that are present: 0 = all three, 1 = {0, 1}, 2 = {1, 2} (the 2x2 sub-kernels of the four
// output phases of upsample-then-conv; the absent taps have zero weights and their MMAs are simply not issued).
template <int N, int TH, int KC, int DYS = 0, int DXS = 0>
__global__ void __launch_bounds__(kConvThreads, 1)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tmap, const ConvArgs a) {
    // DYS == DXS == 3: ALL FOUR phases of an upsample-folded conv in one launch. Work item = (tile, phase), phase fastest: the
    // CTA loads the same input tile four times back to back (three of them L2 hits) instead of four launches each streaming
    // the whole input from HBM; per phase its own pre-summed weights (a.l_wpack[phase]), tap window and output pixel offset.
    // The MMA sequence of a phase is exactly that of the single-phase instantiation: results are bit-identical.
    constexpr bool kAllPh = DYS == 3 && DXS == 3;
    using T = ConvTraits<N, TH, KC>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t s_bars[2 * kMaxStages + 5];
    __shared__ uint32_t s_tmem_slot;
    __shared__ __align__(16) float s_bias_all[kMaxLayers][N];
    __shared__ __align__(16) float s_neg[N];  // multiplier of the negative part: 1 / slope / PReLU weight
    __shared__ int s_l_nchunks[kMaxLayers], s_l_out_coff[kMaxLayers];
    __shared__ const __half* s_l_wpack[kMaxLayers];
    uint64_t* full = s_bars;
    uint64_t* empty = full + kMaxStages;
    uint64_t* tfull = empty + kMaxStages;
    uint64_t* tempty = tfull + 2;
    uint64_t* wfull = tempty + 2;  // resident weights landed
    const int nstages = a.nstages;
    uint8_t* stage0 = smem + (a.wres ? a.nchunks * T::kBStage : 0);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const long long t_start = clock64();

    if (threadIdx.x == 0) {
        for (int i = 0; i < kMaxStages; ++i) {
            ptx::mbar_init(&full[i], 1);
            ptx::mbar_init(&empty[i], 1);
        }
        ptx::mbar_init(wfull, 1);
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&tfull[i], kMmaWarps);  // every issuer commits its own MMAs of the tile
            ptx::mbar_init(&tempty[i], kEpiWarps);
        }
        ptx::fence_mbar_init();
    }
    if (warp == kEpiWarps) {
        if (lane == 0) ptx::prefetch_tmap(&tmap);
        __syncwarp();
        ptx::tmem_alloc<T::kTmemCols>(&s_tmem_slot);
    }
    if (threadIdx.x < kMaxLayers) {
        const int l = threadIdx.x;
        const bool multi = a.nlayers > 1;
        s_l_nchunks[l] = multi ? a.l_nchunks[l] : a.nchunks;
        s_l_out_coff[l] = multi ? a.l_out_coff[l] : a.out_coff;
        s_l_wpack[l] = multi ? a.l_wpack[l] : a.wpack;
    }
    for (int i = threadIdx.x; i < N * kMaxLayers; i += blockDim.x) {
        const int l = i / N, ch = i - l * N;
        const float* bp = a.nlayers > 1 ? (l < a.nlayers ? a.l_bias[l] : nullptr) : a.bias;
        s_bias_all[l][ch] = (ch < a.cout && bp) ? bp[ch] : 0.f;
    }
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        float neg = 1.f;
        if (a.act == ACT_LRELU) neg = a.slope;
        if (a.act == ACT_PRELU) neg = (i < a.cout && a.prelu) ? a.prelu[i] : 0.f;
        s_neg[i] = neg;
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = s_tmem_slot;
    const int num_tiles = a.tiles_x * a.tiles_y;
    const int num_items = kAllPh ? num_tiles * 4 : num_tiles * a.nlayers;
    if (a.wres && warp == kEpiWarps && lane == 0) {
        // weights are never written by a kernel: fetch them before waiting on the previous layer
        ptx::mbar_expect_tx(wfull, a.nchunks * T::kBBytes);
        for (int c = 0; c < a.nchunks; ++c)
            ptx::bulk_load(smem + c * T::kBStage, a.wpack + static_cast<size_t>(c) * 9 * N * KC, T::kBBytes, wfull);
    }
    // Programmatic dependent launch: everything above (barrier init, TMEM allocation, bias table; weights and bias are
    // never written by a kernel) overlaps the previous layer's tail. The next layer may start launching now; this
    // layer's activations / residuals are only touched after the previous grid has completed and flushed.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (a.dep_zero && blockIdx.x == 0 && warp == 0)
        for (int i = lane; i < a.dep_zero_n; i += 32) a.dep_zero[i] = 0;

    if (warp == kEpiWarps) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            int pre0 = 0, pre1 = 0, pre2 = 0;
            bool pre_valid = false;
            for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
                const int layer = kAllPh ? 0 : item / num_tiles, tile = kAllPh ? item >> 2 : item - layer * num_tiles;
                const int ty = tile / a.tiles_x, tx = tile - ty * a.tiles_x;
                const int x0 = tx * 128, y0 = a.y_begin + ty * TH;
                const int nch = s_l_nchunks[layer];
                const __half* wp = kAllPh ? a.l_wpack[item & 3] : s_l_wpack[layer];
                if (layer > 0) {
                    // rows y0-1 .. y0+TH of the previous layer = its tile rows ty-1 .. ty+1, all tile columns.
                    // The three counters were read one item ahead (pre0..2, below): normally no round trip here.
                    const int target = a.tiles_x * kEpiWarps;
                    const int r0 = ty > 0 ? ty - 1 : 0, r1 = ty + 1 < a.tiles_y ? ty + 1 : a.tiles_y - 1;
                    const int rm = r0 + 1 <= r1 ? r0 + 1 : r1;
                    const int* ctr = a.dep + (layer - 1) * a.tiles_y;
                    if (!(pre_valid && pre0 >= target && pre1 >= target && pre2 >= target)) {
                        const long long tw = clock64();
                        for (;;) {
                            int v0, v1, v2;
                            asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(v0) : "l"(ctr + r0) : "memory");
                            asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(v1) : "l"(ctr + rm) : "memory");
                            asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(v2) : "l"(ctr + r1) : "memory");
                            if (v0 >= target && v1 >= target && v2 >= target) break;
                            if (clock64() - tw > 4000000000LL) __trap();
                        }
                    }
                    // the rows were written with generic-proxy stores by other CTAs; TMA reads through the async proxy
                    asm volatile("fence.proxy.async;" ::: "memory");
                }
                pre_valid = false;
                {
                    // read the NEXT item's dependency counters now; their latency hides behind this item's loads
                    const int nitem = item + static_cast<int>(gridDim.x);
                    if (!kAllPh && nitem < num_items) {  // (an all-phase launch has no inter-layer dependencies: items are phases)
                        const int nl = nitem / num_tiles, nt = nitem - nl * num_tiles;
                        if (nl > 0) {
                            const int nty = nt / a.tiles_x;
                            const int r0 = nty > 0 ? nty - 1 : 0, r1 = nty + 1 < a.tiles_y ? nty + 1 : a.tiles_y - 1;
                            const int rm = r0 + 1 <= r1 ? r0 + 1 : r1;
                            const int* ctr = a.dep + (nl - 1) * a.tiles_y;
                            asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(pre0) : "l"(ctr + r0) : "memory");
                            asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(pre1) : "l"(ctr + rm) : "memory");
                            asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(pre2) : "l"(ctr + r1) : "memory");
                            pre_valid = true;
                        }
                    }
                }
                for (int c = 0; c < nch; ++c) {
                    ptx::mbar_wait(&empty[s], ph ^ 1);
                    uint8_t* st = stage0 + s * a.stage_bytes;
                    if (a.flags & FLAG_SKIP_TMA) {
                        ptx::mbar_arrive(&full[s]);
                    } else {
                        const bool ld_a = !(a.flags & FLAG_SKIP_A), ld_b = !(a.flags & FLAG_SKIP_B) && !a.wres;
                        ptx::mbar_expect_tx(&full[s], (ld_a ? T::kCopyBytes : 0) + (ld_b ? T::kBBytes : 0));
                        const int ch0 = a.cin_off + c * KC;
                        if (ld_a)
                            ptx::tma_load_4d(st, &tmap, &full[s], a.in_cstride == 32 ? 0 : ch0, x0 - 1, y0 - 1,
                                             a.in_cstride == 32 ? ch0 >> 5 : 0);
                        if (ld_b)
                            ptx::bulk_load(st + T::kAStage, wp + static_cast<size_t>(c) * 9 * N * KC, T::kBBytes, &full[s]);
                    }
                    if (++s == nstages) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp > kEpiWarps) {
        // ===================== MMA issuers (two warps, alternating pipeline stages) =====================
        // Each warp walks the whole (uniform) loop so every address stays on the uniform datapath; one elected lane
        // issues the tcgen05 instructions and the commits. Measured (tools/mma_queue_probe, mma_commit_probe): the
        // tensor pipe queues only ~8 MMAs (~400 cycles of work) and reading an mbarrier completed by TMA or
        // tcgen05.commit costs the reader ~250 cycles, so a single issuer drains the queue at every stage boundary
        // (~830 cycles lost per 36-MMA stage). With two issuers, warp w owns the stages with (global stage index) % 2
        // == w: it does its barrier waits while the other warp is issuing, then takes over through a named barrier, so
        // issue order stays deterministic (~320 cycles lost per stage; polling the next barrier mid-stage from a single
        // issuer was tried and is slower).
        const int mw = warp - (kEpiWarps + 1);  // 0 or 1
        if (a.wres) ptx::mbar_wait(wfull, 0);
        int gstage = 0;                         // global stage counter over all tiles of this CTA
        // dy-stacked N: for an input row rho the three taps dy contribute to the three output rows r = rho - dy,
        // whose accumulators are ADJACENT TMEM column blocks. With the weights of (dy = 2, 1, 0) stored as
        // consecutive B rows, one MMA of N = 3*Cout updates all three rows and reads the A tile once.
        const bool skip_mma = (a.flags & FLAG_SKIP_MMA) != 0;
        int s = 0;
        uint32_t ph = 0;
        int it = 0;
        for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
            const int buf = it & 1;
            const uint32_t aph = (it >> 1) & 1;
            const int nch = kAllPh ? a.nchunks : s_l_nchunks[item / num_tiles];
            ptx::mbar_wait(&tempty[buf], aph ^ 1);
            ptx::tc_fence_after();
            const uint32_t d_base = tmem_base + buf * T::kAccCols;
            for (int c = 0; c < nch; ++c) {
                const bool mine = (gstage & 1) == mw;
                if (mine) {
                    ptx::mbar_wait(&full[s], ph);
                    ptx::tc_fence_after();
                    // hand-over: wait until the other warp has issued the previous stage (barrier id 1 + mw).
                    // BAR.SYNC blocks lazily (at the next dependent instruction), which is all that is needed here.
                    if (gstage > 0) asm volatile("bar.sync %0, 64;" ::"r"(1 + mw) : "memory");
                }
                if (mine && ptx::elect_one()) {
                    const uint32_t a_lo0 = (ptx::smem_u32(stage0 + s * a.stage_bytes) >> 4);
                    const uint32_t b_lo0 = a.wres ? (ptx::smem_u32(smem + c * T::kBStage) >> 4) : a_lo0 + (T::kAStage >> 4);
                    if (!skip_mma) {
                        if constexpr (kAllPh) {
                            switch (item & 3) {  // phase (py, px): taps {0,1} for 0, {1,2} for 1
                                case 0: conv_issue_stage<N, TH, KC, 0, 1, 0, 1>(a_lo0, b_lo0, d_base, c == 0); break;
                                case 1: conv_issue_stage<N, TH, KC, 0, 1, 1, 2>(a_lo0, b_lo0, d_base, c == 0); break;
                                case 2: conv_issue_stage<N, TH, KC, 1, 2, 0, 1>(a_lo0, b_lo0, d_base, c == 0); break;
                                default: conv_issue_stage<N, TH, KC, 1, 2, 1, 2>(a_lo0, b_lo0, d_base, c == 0); break;
                            }
                        } else {
                            constexpr int kDy0 = DYS == 2 ? 1 : 0, kDy1 = DYS == 1 ? 1 : 2;  // present dy taps [kDy0, kDy1]
                            constexpr int kDx0 = DXS == 2 ? 1 : 0, kDx1 = DXS == 1 ? 1 : 2;
                            conv_issue_stage<N, TH, KC, kDy0, kDy1, kDx0, kDx1>(a_lo0, b_lo0, d_base, c == 0);
                        }
                    }
                    ptx::umma_commit(&empty[s]);
                }
                __syncwarp();
                if (mine) asm volatile("bar.arrive %0, 64;" ::"r"(2 - mw) : "memory");  // other warp may issue next
                ++gstage;
                if (++s == nstages) { s = 0; ph ^= 1; }
            }
            // this warp's MMAs of the tile are all issued: its commit is one of the kMmaWarps arrivals on tfull
            if (ptx::elect_one()) ptx::umma_commit(&tfull[buf]);
            __syncwarp();
        }
        // consume the last, unmatched bar.arrive so no named barrier is left half-arrived
        if (gstage > 0 && (gstage & 1) == mw) asm volatile("bar.sync %0, 64;" ::"r"(1 + mw) : "memory");
    } else {
        // ===================== epilogue warps 0..7 =====================
        const int quarter = warp & 3;  // TMEM lanes [32*quarter, 32*quarter + 32)
        const int rgrp = warp >> 2;    // rows rgrp, rgrp + 2, ...
        // activation mode for bias_act: 0 identity, 1 max-form LeakyReLU, 2 general
        const int amode = a.act == ACT_NONE ? 0 : ((a.act == ACT_LRELU && a.slope >= 0.f && a.slope <= 1.f) ? 1 : 2);
        const uint32_t stg_s = ptx::smem_u32(stage0 + nstages * a.stage_bytes + warp * T::kStgWarp);
        int it = 0;
        for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
            const int buf = it & 1;
            const uint32_t aph = (it >> 1) & 1;
            const int layer = kAllPh ? 0 : item / num_tiles, tile = kAllPh ? item >> 2 : item - layer * num_tiles;
            const int opy = kAllPh ? (item >> 1) & 1 : a.opy, opx = kAllPh ? item & 1 : a.opx;
            const int out_coff = s_l_out_coff[layer];
            const float* s_bias = s_bias_all[layer];
            const int ty = tile / a.tiles_x, tx = tile - ty * a.tiles_x;
            const int x = tx * 128 + quarter * 32 + lane;
            const int y0 = a.y_begin + ty * TH;
            const bool inb = x < a.W;
            bool xgap = false;
            for (int j = 0; j < a.ngx; ++j) xgap |= ((x >> a.gshift) == a.gx[j]);
            ptx::mbar_wait(&tfull[buf], aph);
            ptx::tc_fence_after();
            const uint32_t t_row0 = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + buf * T::kAccCols;
            const int r_end = (a.flags & FLAG_SKIP_EPI) ? 0 : TH;
#pragma unroll 1
            for (int r = rgrp; r < r_end; r += 2) {
                const int y = y0 + r;
                if (y >= a.y_end) break;  // warp-uniform
                const size_t p = static_cast<size_t>(y) * a.W + x;
                bool gap = xgap;
                for (int j = 0; j < a.ngy; ++j) gap |= ((y >> a.gshift) == a.gy[j]);
                if (a.out_mode == OUT_PS4) {
                    if constexpr (N == 48) {
                        float v[48];
                        ptx::tmem_ld16(t_row0 + r * N, v);
                        ptx::tmem_ld16(t_row0 + r * N + 16, v + 16);
                        ptx::tmem_ld16(t_row0 + r * N + 32, v + 32);
                        if (inb) {
                            float b3[3];
                            const __half* bp = a.base + p * a.base_cstride;
#pragma unroll
                            for (int ch = 0; ch < 3; ++ch) b3[ch] = __half2float(bp[ch]);
                            const int Wo = a.W * 4;
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                float px[16];
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
#pragma unroll
                                    for (int ch = 0; ch < 3; ++ch)
                                        px[j * 4 + ch] = v[ch * 16 + i * 4 + j] + s_bias[ch * 16 + i * 4 + j] + b3[ch];
                                    px[j * 4 + 3] = 0.f;
                                }
                                uint4* dst = reinterpret_cast<uint4*>(
                                    a.out + (static_cast<size_t>(y * 4 + i) * Wo + static_cast<size_t>(x) * 4) * 4);
                                dst[0] = pack8(px);
                                dst[1] = pack8(px + 8);
                            }
                        }
                    }
                } else if (a.out_mode == OUT_RGB4) {
                    float v[16];
                    ptx::tmem_ld16(t_row0 + r * N, v);
                    if (inb) {
                        bias_act<4>(v, s_bias, s_neg, 0, amode);
                        __half2 h01 = __floats2half2_rn(v[0], v[1]);
                        __half2 h23 = __floats2half2_rn(v[2], 0.f);
                        uint2 q;
                        q.x = *reinterpret_cast<uint32_t*>(&h01);
                        q.y = *reinterpret_cast<uint32_t*>(&h23);
                        *reinterpret_cast<uint2*>(a.out + p * 4) = q;
                    }
                } else {
                    if constexpr (N % 32 == 0) {
                        epi_row_nhwc<N>(a, t_row0 + r * N, stg_s, lane, tx * 128 + quarter * 32, y, gap, out_coff, 0, s_bias, s_neg,
                                        amode, opy, opx);
                        __syncwarp();
                    }
                }
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&tempty[buf]);  // TMEM buffer is free as soon as it has been read
            if (layer + 1 < a.nlayers) {
                __threadfence();  // every lane: this tile's rows are visible device-wide before the row counter moves
                __syncwarp();
                if (lane == 0) atomicAdd(a.dep + layer * a.tiles_y + ty, 1);
            }
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == kEpiWarps) {
        __syncwarp();
        ptx::tmem_dealloc<T::kTmemCols>(tmem_base);
    }
    if (a.dbg_cycles && threadIdx.x == 0) a.dbg_cycles[blockIdx.x] = clock64() - t_start;
}

}  // namespace vr
