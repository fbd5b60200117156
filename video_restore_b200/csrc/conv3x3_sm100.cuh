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
//     of tile i+1. One thread issues all MMAs; MMAs that share an A row tile (same input row, dx, k) are
//     issued back to back with collector::a fill/use/lastuse so A is read from smem once per input row.
//   * epilogue (4 warps): tcgen05.ld -> +bias -> LeakyReLU/PReLU -> *s1 + res1 -> *s2 + res2 in fp32 ->
//     one fp16 rounding -> 16 B stores into a channel slice of the destination NHWC buffer (zero-copy concat).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cstdint>
#include "sm100_ptx.cuh"

namespace vr {

enum ConvAct { ACT_NONE = 0, ACT_LRELU = 1, ACT_PRELU = 2 };
enum ConvOut { OUT_NHWC = 0, OUT_RGB4 = 1, OUT_PS4 = 2 };
// debug ablation flags (ConvArgs::flags): measurement only
enum ConvFlags { FLAG_NO_COLLECTOR = 1, FLAG_SKIP_TMA = 2, FLAG_SKIP_MMA = 4, FLAG_SKIP_EPI = 8 };

struct ConvArgs {
    int W, H;              // conv input == output extent
    int tiles_x, tiles_y;  // ceil(W/128), ceil(H/TH)
    int nchunks;           // Cin_padded / 32
    int cin_off;           // first input channel inside the source buffer
    const __half* wpack;   // [nchunks][9][N][32] fp16, pre-swizzled smem image
    const float* bias;     // [cout]
    const float* prelu;    // [cout] or null
    int act;
    float slope;
    __half* out;
    int out_cstride, out_coff, cout;
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
};

constexpr int round_up_c(int x, int m) { return (x + m - 1) / m * m; }
constexpr int next_pow2_c(int x) { int p = 32; while (p < x) p *= 2; return p; }

template <int N, int TH>
struct ConvTraits {
    static constexpr int kInRows = TH + 2;
    static constexpr int kPitch = 130;  // 128 output pixels + 1 halo pixel each side
    static constexpr int kCopyBytes = kInRows * kPitch * 64;  // bytes one TMA box delivers
    static constexpr int kCopyStride = round_up_c(kCopyBytes, 1024);
    static constexpr int kAStage = kCopyStride;
    static constexpr int kBBytes = 9 * N * 64;
    static constexpr int kBStage = round_up_c(kBBytes, 1024);
    static constexpr int kStageBytes = kAStage + kBStage;
    static constexpr int kTail = 1024;  // barriers, tmem slot, bias, prelu
    static constexpr int kBudget = 227 * 1024 - 1024 - kTail;
    static constexpr int kStagesRaw = kBudget / kStageBytes;
    static constexpr int kStages = kStagesRaw > 4 ? 4 : (kStagesRaw < 1 ? 1 : kStagesRaw);
    static constexpr int kAccCols = TH * N;
    static constexpr int kTmemCols = next_pow2_c(2 * kAccCols);
    static constexpr int kSmemBytes = kStages * kStageBytes + kTail + 1024;
    static_assert(kStagesRaw >= 1, "stage does not fit in shared memory");
    static_assert(kTmemCols <= 512, "accumulators do not fit in TMEM");
    static_assert(N % 16 == 0 && N >= 16 && N <= 64, "N must be 16..64 step 16");
};

__device__ __forceinline__ float apply_act(float v, int act, float slope, float pr) {
    if (act == ACT_LRELU) return v > 0.f ? v : v * slope;
    if (act == ACT_PRELU) return v > 0.f ? v : v * pr;
    return v;
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

template <int N, int TH, bool COLL>
__global__ void __launch_bounds__(192, 1)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tmap, const ConvArgs a) {
    using T = ConvTraits<N, TH>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* tail = smem + T::kStages * T::kStageBytes;
    uint64_t* full = reinterpret_cast<uint64_t*>(tail);
    uint64_t* empty = full + T::kStages;
    uint64_t* tfull = empty + T::kStages;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
    float* s_bias = reinterpret_cast<float*>(tail + 128);
    float* s_prelu = reinterpret_cast<float*>(tail + 128 + 256);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int i = 0; i < T::kStages; ++i) {
            ptx::mbar_init(&full[i], 1);
            ptx::mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&tfull[i], 1);
            ptx::mbar_init(&tempty[i], 4);
        }
        ptx::fence_mbar_init();
    }
    if (warp == 4) {
        if (lane == 0) ptx::prefetch_tmap(&tmap);
        __syncwarp();
        ptx::tmem_alloc<T::kTmemCols>(tmem_slot);
    }
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        s_bias[i] = (i < a.cout && a.bias) ? a.bias[i] : 0.f;
        s_prelu[i] = (i < a.cout && a.prelu) ? a.prelu[i] : 0.f;
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int num_tiles = a.tiles_x * a.tiles_y;

    if (warp == 4) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int ty = tile / a.tiles_x, tx = tile - ty * a.tiles_x;
                const int x0 = tx * 128, y0 = ty * TH;
                for (int c = 0; c < a.nchunks; ++c) {
                    ptx::mbar_wait(&empty[s], ph ^ 1);
                    uint8_t* st = smem + s * T::kStageBytes;
                    if (a.flags & FLAG_SKIP_TMA) {
                        ptx::mbar_arrive(&full[s]);
                    } else {
                        ptx::mbar_expect_tx(&full[s], T::kCopyBytes + T::kBBytes);
                        ptx::tma_load_4d(st, &tmap, &full[s], a.cin_off + c * 32, x0 - 1, y0 - 1, 0);
                        ptx::bulk_load(st + T::kAStage, a.wpack + static_cast<size_t>(c) * 9 * N * 32, T::kBBytes,
                                       &full[s]);
                    }
                    if (++s == T::kStages) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 5) {
        // ===================== MMA issuer =====================
        // The whole warp walks the (uniform) loop so every address stays on the uniform datapath; one elected
        // lane issues the tcgen05 instructions and the commits.
        constexpr uint32_t idesc = ptx::make_idesc_f16(128, N);
        const bool skip_mma = (a.flags & FLAG_SKIP_MMA) != 0;
        int s = 0;
        uint32_t ph = 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int buf = it & 1;
            const uint32_t aph = (it >> 1) & 1;
            ptx::mbar_wait(&tempty[buf], aph ^ 1);
            ptx::tc_fence_after();
            const uint32_t d_base = tmem_base + buf * T::kAccCols;
            for (int c = 0; c < a.nchunks; ++c) {
                ptx::mbar_wait(&full[s], ph);
                ptx::tc_fence_after();
                if (ptx::elect_one()) {
                    const uint32_t a_lo0 = (ptx::smem_u32(smem + s * T::kStageBytes) >> 4);
                    const uint32_t b_lo0 = a_lo0 + (T::kAStage >> 4);
                    if (!skip_mma) {
#pragma unroll
                        for (int rho = 0; rho < T::kInRows; ++rho) {
#pragma unroll
                            for (int dx = 0; dx < 3; ++dx) {
#pragma unroll
                                for (int k = 0; k < 2; ++k) {
                                    const uint32_t a_lo = a_lo0 + (((rho * T::kPitch + dx) * 64 + k * 32) >> 4);
                                    constexpr int kLast = TH - 1;
                                    const int dy_lo = rho - kLast > 0 ? rho - kLast : 0;
                                    const int dy_hi = rho < 2 ? rho : 2;
#pragma unroll
                                    for (int dy = 0; dy < 3; ++dy) {
                                        if (dy < dy_lo || dy > dy_hi) continue;
                                        const int r = rho - dy;
                                        const uint32_t b_lo = b_lo0 + ((((dy * 3 + dx) * N) * 64 + k * 32) >> 4);
                                        const uint32_t acc = (c | dy | dx | k) != 0 ? 1u : 0u;
                                        const uint32_t d = d_base + r * N;
                                        if (!COLL || dy_lo == dy_hi)
                                            ptx::umma_f16<ptx::kCollNone>(d, a_lo, ptx::kDescHiSw64, b_lo,
                                                                          ptx::kDescHiSw64, idesc, acc);
                                        else if (dy == dy_lo)
                                            ptx::umma_f16<ptx::kCollFill>(d, a_lo, ptx::kDescHiSw64, b_lo,
                                                                          ptx::kDescHiSw64, idesc, acc);
                                        else if (dy == dy_hi)
                                            ptx::umma_f16<ptx::kCollLastUse>(d, a_lo, ptx::kDescHiSw64, b_lo,
                                                                             ptx::kDescHiSw64, idesc, acc);
                                        else
                                            ptx::umma_f16<ptx::kCollUse>(d, a_lo, ptx::kDescHiSw64, b_lo,
                                                                         ptx::kDescHiSw64, idesc, acc);
                                    }
                                }
                            }
                        }
                    }
                    ptx::umma_commit(&empty[s]);
                    if (c == a.nchunks - 1) ptx::umma_commit(&tfull[buf]);
                }
                __syncwarp();
                if (++s == T::kStages) { s = 0; ph ^= 1; }
            }
        }
    } else {
        // ===================== epilogue warps 0..3 =====================
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int buf = it & 1;
            const uint32_t aph = (it >> 1) & 1;
            const int ty = tile / a.tiles_x, tx = tile - ty * a.tiles_x;
            const int x = tx * 128 + warp * 32 + lane;
            const int y0 = ty * TH;
            ptx::mbar_wait(&tfull[buf], aph);
            ptx::tc_fence_after();
            const uint32_t t_row0 = tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + buf * T::kAccCols;
            const int r_end = (a.flags & FLAG_SKIP_EPI) ? 0 : TH;
#pragma unroll 1
            for (int r = 0; r < r_end; ++r) {
                const int y = y0 + r;
                if (y >= a.H) break;  // warp-uniform
                const size_t p = static_cast<size_t>(y) * a.W + x;
                const bool inb = x < a.W;
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
                        float o[4];
#pragma unroll
                        for (int ch = 0; ch < 3; ++ch)
                            o[ch] = apply_act(v[ch] + s_bias[ch], a.act, a.slope, s_prelu[ch]);
                        o[3] = 0.f;
                        __half2 h01 = __floats2half2_rn(o[0], o[1]);
                        __half2 h23 = __floats2half2_rn(o[2], o[3]);
                        uint2 q;
                        q.x = *reinterpret_cast<uint32_t*>(&h01);
                        q.y = *reinterpret_cast<uint32_t*>(&h23);
                        *reinterpret_cast<uint2*>(a.out + p * 4) = q;
                    }
                } else {
#pragma unroll
                    for (int g = 0; g < N / 16; ++g) {
                        if (g * 16 >= a.cout) break;  // uniform
                        float v[16];
                        ptx::tmem_ld16(t_row0 + r * N + g * 16, v);
                        if (inb) {
                            const int c0 = g * 16;
#pragma unroll
                            for (int j = 0; j < 16; ++j)
                                v[j] = apply_act(v[j] + s_bias[c0 + j], a.act, a.slope, s_prelu[c0 + j]);
                            if (a.res1) {
                                const uint4* rp =
                                    reinterpret_cast<const uint4*>(a.res1 + p * a.res1_cstride + a.res1_coff + c0);
                                float f[16];
                                unpack8(rp[0], f);
                                unpack8(rp[1], f + 8);
#pragma unroll
                                for (int j = 0; j < 16; ++j) v[j] = fmaf(v[j], a.s1, f[j]);
                            }
                            if (a.res2) {
                                const uint4* rp =
                                    reinterpret_cast<const uint4*>(a.res2 + p * a.res2_cstride + a.res2_coff + c0);
                                float f[16];
                                unpack8(rp[0], f);
                                unpack8(rp[1], f + 8);
#pragma unroll
                                for (int j = 0; j < 16; ++j) v[j] = fmaf(v[j], a.s2, f[j]);
                            }
                            uint4* op = reinterpret_cast<uint4*>(a.out + p * a.out_cstride + a.out_coff + c0);
                            op[0] = pack8(v);
                            op[1] = pack8(v + 8);
                        }
                    }
                }
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&tempty[buf]);
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        __syncwarp();
        ptx::tmem_dealloc<T::kTmemCols>(tmem_base);
    }
}

}  // namespace vr
