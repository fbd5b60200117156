// K3: paired rolling-row 3x3 convolution -- K2 (conv3x3_roll_sm100.cuh) on CTA PAIRS with tcgen05 cta_group::2.
//
// A cluster of two CTAs (two SMs of one TPC) owns two 128-pixel strips x bands of rows (normally neighbours). The leader (cluster
// rank 0) issues every MMA with M = 256: each CTA supplies the activations of ITS strip (A) and HALF of the weight rows (B)
// at the same shared-memory offsets, and finds its strip's accumulators in its own TMEM (tools/mma2cta_probe). Per SM an
// MMA therefore reads 4 KB of A but only half of B:
//   * the dy-stacked N = 3 Cout MMA is tensor-bound for Cout = 32 as well (5.5 KB per 48 cycles instead of 7 KB);
//   * a layer's weights take half the shared memory per CTA, so the 192 -> 64 layer (221 KB) is resident (110 KB per CTA)
//     with full N = 192 MMAs -- no weight re-streaming per tile (K1) and no second pass over the activations (K2 halves).
// The issue sequence has no special cases at all (with cta_group::2 a partial window would need a different B split):
//   * PHANTOM rows: a band of `nrow` output rows is computed as nin2 + 2 "logical" rows (nin2 = its nrow + 2 input rows rounded
//     up to even); input row j always updates logical rows j, j+1, j+2 with one full-window MMA per (chunk, dx, k16). The two
//     logical rows before and the 2..3 after the band only collect partial sums and are discarded by the epilogue;
//   * MIRRORED ring: logical row g lives at ring position m = g mod P, P = 512/Cout - 2, and the window of an input row is
//     the three physical blocks starting at ITS first logical row's position -- never wrapping, because two spare physical
//     blocks P, P+1 extend the ring. Rows with m < 2 therefore collect their first taps in block P + m and the rest in block
//     m; the epilogue adds the two (mirror blocks restart from zero, ring blocks from the bias).
// Synchronisation: each CTA's producer loads its own strip; both count their bytes on the LEADER's full barrier. The
// leader's commits are multicast (slot release and accumulator-ready barriers exist in both CTAs); both CTAs' epilogue
// warps arrive on the leader's accumulator-free barriers (remote mbarrier arrive). Warp roles as in K1 / K2.
// Instantiations: kDirect = epilogue with direct 256-bit stores (default; false = staged through shared memory, for
// outputs that are not 32-channel groups and for A/B runs); kNoRes = layer without residual operands (no residual
// registers: 112 / 156 instead of 156 / 168, which is what lets N = 64 release its ring position before the arithmetic).
#pragma once
#include "conv3x3_roll_sm100.cuh"

namespace vr {

constexpr int kPairMaxSlots = 12;

template <int N>
struct PairTraits {
    static_assert(N == 32 || N == 64, "paired kernel: 32 or 64 output channels");
    static constexpr int KC = 32;
    static constexpr int kLineBytes = 130 * 64;                  // one input row of one channel chunk
    static constexpr int kCopyBytes = 2 * kLineBytes;            // box = two input rows
    static constexpr int kASlot = round_up_c(kCopyBytes, 512);
    static constexpr int kWin = 3 * N;                           // MMA N: the dy taps of one input row
    static constexpr int kBTap = (kWin / 2) * 64;                // this CTA's B rows of one (chunk, dx)
    static constexpr int kBHalf = 3 * kBTap;                     // ... of one chunk: 9216 (N = 32) / 18432 (N = 64) bytes
    static constexpr int kPhys = 512 / N;                        // physical ring blocks in TMEM
    static constexpr int kPeriod = kPhys - 2;                    // logical ring period (two blocks are mirrors)
    // 8 epilogue warps = two rows in flight. More were measured slower on every layer shape: 16 warps (96 registers, ~100 B of
    // spills) 160 -> 32 79.5 vs 67.6 us; 12 warps (128 registers) 64 -> 32 36.9 vs 35.3 us.
    static constexpr int kEpi = 8;                               // epilogue warps
    static constexpr int kGroups = kEpi / 4;                     // rows in flight
    static constexpr int kThreads = (kEpi + 1 + kMmaWarps) * 32;
    static constexpr int kStgBytes = kEpi * 32 * N * 2;
    static constexpr int kStatic = 3072;
    static constexpr int kBudget = 227 * 1024 - 1024 - kStatic - kStgBytes;
    static constexpr int kMinSlots = 3;
    static_assert(kBTap % 512 == 0, "B tiles must keep the swizzle phase");
};

// The twelve MMAs of a box: (dx, k16) outer, the two input rows inner; one runtime base per operand, compile-time offsets.
template <int N>
__device__ __forceinline__ void pair_issue_box(uint32_t d0, uint32_t d1, uint32_t a_lo0, uint32_t b_lo0) {
    using T = PairTraits<N>;
    constexpr uint32_t kDescHi = ptx::kDescHiSw64;
    constexpr uint32_t kIdesc = ptx::make_idesc_f16(256, T::kWin);
    constexpr uint32_t kLine = T::kLineBytes >> 4;
#pragma unroll
    for (int t = 0; t < 6; ++t) {
        const uint32_t ao = ((t >> 1) * 64 + (t & 1) * 32) >> 4, bo = ((t >> 1) * T::kBTap + (t & 1) * 32) >> 4;
        ptx::umma_f16_pair(d0, a_lo0 + ao, kDescHi, b_lo0 + bo, kDescHi, kIdesc, 1u);
        ptx::umma_f16_pair(d1, a_lo0 + kLine + ao, kDescHi, b_lo0 + bo, kDescHi, kIdesc, 1u);
    }
}

// Hand a ring position back: ring block <- bias row, mirror block <- 0, then arrive on the LEADER's barrier.
template <int N>
__device__ __forceinline__ void pair_release(uint32_t t_main, uint32_t t_mir, const float* s_bias, int lane, uint32_t release_addr) {
#pragma unroll
    for (int g = 0; g < N / 32; ++g) {
        float bz[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4 t = *reinterpret_cast<const float4*>(s_bias + g * 32 + j * 4);
            bz[j * 4] = t.x; bz[j * 4 + 1] = t.y; bz[j * 4 + 2] = t.z; bz[j * 4 + 3] = t.w;
        }
        ptx::tmem_st32(t_main + g * 32, bz);
        if (t_mir != 0xffffffffu) ptx::tmem_st32_zero(t_mir + g * 32);
    }
    ptx::tmem_st_wait();
    ptx::tc_fence_before();
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive_cluster(release_addr);
}

// activation of 32 accumulator values (amode: 0 none, 1 leaky with slope in [0, 1], 2 per-channel negative slope table)
__device__ __forceinline__ void epi_act32(float* v, int amode, float slope, const float* neg) {
    if (amode == 1) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], v[j] * slope);
    } else if (amode == 2) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f) + neg[j] * fminf(v[j], 0.f);
    }
}
// 32 channels of this lane's pixel -> fp16 -> the warp's swizzled staging buffer (16 B unit u of group g)
template <int N>
__device__ __forceinline__ void epi_stage32(const float* v, uint32_t stg_s, int lane, int g) {
    constexpr int kVec = N / 8;
    const int swz_w = kVec == 8 ? (lane & 7) : ((lane >> 1) & 3);
#pragma unroll
    for (int u = 0; u < 4; ++u) ptx::sts128(stg_s + lane * (N * 2) + (((g * 4 + u) ^ swz_w) << 4), pack8(v + u * 8));
}

// pair_release with eight registers of bias in flight instead of 32 (for the path that holds a whole 64-column row)
template <int N>
__device__ __forceinline__ void pair_release_lean(uint32_t t_main, uint32_t t_mir, const float* s_bias, int lane, uint32_t release_addr) {
#pragma unroll
    for (int q = 0; q < N / 8; ++q) {
        const float4 lo = *reinterpret_cast<const float4*>(s_bias + q * 8), hi = *reinterpret_cast<const float4*>(s_bias + q * 8 + 4);
        ptx::tmem_st8(t_main + q * 8, lo, hi);
    }
    if (t_mir != 0xffffffffu) {
#pragma unroll
        for (int g = 0; g < N / 32; ++g) ptx::tmem_st32_zero(t_mir + g * 32);
    }
    ptx::tmem_st_wait();
    ptx::tc_fence_before();
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive_cluster(release_addr);
}

// One output row x 32 pixels of this warp's lane quarter: accumulators (bias included; + mirror block when t_mir != ~0u) ->
// activation -> residuals -> fp16 -> swizzled staging -> coalesced stores. Returns true when the ring position has already
// been handed back (otherwise the caller does it).
template <int N>
__device__ __forceinline__ bool epi_row_pair(const ConvArgs& a, uint32_t t_main, uint32_t t_mir, uint32_t stg_s, int lane, int x_base,
                                             int y, bool gap, const float* s_bias, const float* s_neg, int amode,
                                             uint32_t release_addr) {
    constexpr int kVec = N / 8;
    constexpr int kStgPitch = N * 2;
    const int x = x_base + lane;
    const bool inb = x < a.W;
    const bool has1 = a.res1 != nullptr, has2 = a.res2 != nullptr;
    // The ring position is handed back as early as the registers allow -- a position held through the global stores keeps
    // the issuers waiting (with N = 64 the ring has only six logical positions). N = 32: right after the TMEM loads (the
    // whole row is 32 registers). N = 64: after the row is staged in shared memory, before the global stores (a.early64,
    // VR_EARLY64=0 for A/B runs: by the caller after the stores). Loading all 64 columns first and releasing before the
    // arithmetic was measured as well: 168 registers are not enough for it (spills), slower than this.
    const bool early = N == 32 || a.early64 != 0;
    {
        const size_t p = static_cast<size_t>(y) * a.W + x;
        uint4 q1[kVec], q2[kVec];
        if (inb && has1) {
#pragma unroll
            for (int j = 0; j < kVec; ++j)
                q1[j] = __ldg(reinterpret_cast<const uint4*>(a.res1 + chan_off(p, a.res1_cstride, a.res1_pstride, a.res1_coff + j * 8)));
        }
        if (inb && has2) {
#pragma unroll
            for (int j = 0; j < kVec; ++j)  // may alias `out` (in-place RRDB skip)
                q2[j] = *reinterpret_cast<const uint4*>(a.res2 + chan_off(p, a.res2_cstride, a.res2_pstride, a.res2_coff + j * 8));
        }
        float v0[32];
        if constexpr (N == 32) {
            if (t_mir != 0xffffffffu) {
                uint32_t r0[32], r1[32];
                ptx::tmem_ld32_issue(t_main, r0);
                ptx::tmem_ld32_issue(t_mir, r1);
                ptx::tmem_ld32_wait(r0);
                ptx::tmem_ld32_wait(r1);
#pragma unroll
                for (int j = 0; j < 32; ++j) v0[j] = __uint_as_float(r0[j]) + __uint_as_float(r1[j]);
            } else {
                ptx::tmem_ld32(t_main, v0);
            }
            pair_release<N>(t_main, t_mir, s_bias, lane, release_addr);
        }
#pragma unroll
        for (int g = 0; g < N / 32; ++g) {
            float v[32];
            if constexpr (N == 32) {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = v0[j];
            } else if (t_mir != 0xffffffffu) {
                uint32_t r0[32], r1[32];
                ptx::tmem_ld32_issue(t_main + g * 32, r0);
                ptx::tmem_ld32_issue(t_mir + g * 32, r1);
                ptx::tmem_ld32_wait(r0);
                ptx::tmem_ld32_wait(r1);
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r0[j]) + __uint_as_float(r1[j]);
            } else {
                ptx::tmem_ld32(t_main + g * 32, v);
            }
            if (inb) {
                epi_act32(v, amode, a.slope, s_neg + g * 32);
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
            epi_stage32<N>(v, stg_s, lane, g);
        }
    }
    if (N == 64 && early) pair_release<N>(t_main, t_mir, s_bias, lane, release_addr);
    __syncwarp();
    // lane l always handles 16 B unit l % kVec of its pixels (32 % kVec == 0)
    __half* orow = a.out + chan_off(static_cast<size_t>(y * a.omul + a.opy) * (a.W * a.omul) + x_base * a.omul + a.opx, a.out_cstride,
                                    a.out_pstride, a.out_coff + (lane % kVec) * 8);
#pragma unroll
    for (int i = 0; i < kVec; ++i) {
        const int idx = i * 32 + lane;
        const int px = idx / kVec, un = idx % kVec;
        if (x_base + px < a.W) {
            const int swz_r = kVec == 8 ? (px & 7) : ((px >> 1) & 3);
            const uint4 val = ptx::lds128(stg_s + px * kStgPitch + ((un ^ swz_r) << 4));
            if (a.flags & FLAG_SKIP_B) continue;  // ablation (power probe): everything but the global stores
            if (a.l2_hint == 1) ptx::stg128_hint(orow + static_cast<size_t>(px) * a.omul * a.out_cstride, val, ptx::l2_policy_evict_last());
            else *reinterpret_cast<uint4*>(orow + static_cast<size_t>(px) * a.omul * a.out_cstride) = val;
        }
    }
    __syncwarp();
    return early;  // ring position already handed back
}

// The same row without the staging transpose: a lane's 32 accumulator columns ARE 64 contiguous bytes of its pixel (one
// channel group of a chunk-planar or 32-channel-multiple tensor), so it stores them itself with two 256-bit stores -- every
// store fills whole 32-byte sectors, and the shared-memory traffic of the transpose (16 KB written + 16 KB read per row and
// CTA at N = 64) is gone. That traffic matters: the MMAs alone read 5.5 KB (N = 32) / 7 KB (N = 64) of operands per 48 / 96
// cycles from shared memory, 90 % / 57 % of its 128 B/cycle, and the TMA writes the activations on top.
template <int N, bool kNoRes>
__device__ __forceinline__ bool epi_row_pair_direct(const ConvArgs& a, uint32_t t_main, uint32_t t_mir, int lane, int x_base, int y, bool gap,
                                                    const float* s_bias, const float* s_neg, int amode, uint32_t release_addr) {
    constexpr int kG = N / 32;
    const int x = x_base + lane;
    const bool inb = x < a.W;
    const size_t p = static_cast<size_t>(y) * a.W + x;
    const bool has1 = !kNoRes && a.res1 != nullptr, has2 = !kNoRes && a.res2 != nullptr;
    if constexpr (N == 64 && kNoRes) {
        // Layers without residual operands (SRVGG body, conv_hr) have their own instantiation: without the residual registers
        // the whole 64-column row fits, both TMEM loads are issued together and the ring position goes back right after
        // them, before any arithmetic (a.early64 == 2).
        if (a.early64 == 2) {
            uint32_t acc[N];
            ptx::tmem_ld32_issue(t_main, acc);
            ptx::tmem_ld32_issue(t_main + 32, acc + 32);
            ptx::tmem_ld32_wait(acc);
            ptx::tmem_ld32_wait(acc + 32);
            if (t_mir != 0xffffffffu) {
#pragma unroll
                for (int g = 0; g < kG; ++g) {
                    float r1[32];
                    ptx::tmem_ld32(t_mir + g * 32, r1);
#pragma unroll
                    for (int j = 0; j < 32; ++j) acc[g * 32 + j] = __float_as_uint(__uint_as_float(acc[g * 32 + j]) + r1[j]);
                }
            }
            pair_release_lean<N>(t_main, t_mir, s_bias, lane, release_addr);
            const bool st = inb && !(a.flags & FLAG_SKIP_B);
#pragma unroll
            for (int g = 0; g < kG; ++g) {
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = gap ? 0.f : __uint_as_float(acc[g * 32 + j]);
                epi_act32(v, amode, a.slope, s_neg + g * 32);
                if (st) {
                    __half* o = a.out + chan_off(p, a.out_cstride, a.out_pstride, a.out_coff + g * 32);
                    const uint4 h0 = pack8(v), h1 = pack8(v + 8), h2 = pack8(v + 16), h3 = pack8(v + 24);
                    ptx::stg256(o, h0, h1);
                    ptx::stg256(o + 16, h2, h3);
                    if (a.out2) {
                        __half* o2 = a.out2 + chan_off(p, a.out2_cstride, a.out_pstride, a.out_coff + g * 32);
                        ptx::stg256(o2, h0, h1);
                        ptx::stg256(o2 + 16, h2, h3);
                    }
                }
            }
            return true;
        }
    }
    uint4 q1[kG * 4], q2[kG * 4];
    if (inb && has1) {
#pragma unroll
        for (int g = 0; g < kG; ++g) {
            const __half* r = a.res1 + chan_off(p, a.res1_cstride, a.res1_pstride, a.res1_coff + g * 32);
            ptx::ldg256_nc(r, q1[g * 4], q1[g * 4 + 1]);
            ptx::ldg256_nc(r + 16, q1[g * 4 + 2], q1[g * 4 + 3]);
        }
    }
    if (inb && has2) {
#pragma unroll
        for (int g = 0; g < kG; ++g) {  // may alias `out` (in-place RRDB skip): read and written by the same lane
            const __half* r = a.res2 + chan_off(p, a.res2_cstride, a.res2_pstride, a.res2_coff + g * 32);
            ptx::ldg256(r, q2[g * 4], q2[g * 4 + 1]);
            ptx::ldg256(r + 16, q2[g * 4 + 2], q2[g * 4 + 3]);
        }
    }
    float v0[32];
    if constexpr (N == 32) {
        if (t_mir != 0xffffffffu) {
            uint32_t r0[32], r1[32];
            ptx::tmem_ld32_issue(t_main, r0);
            ptx::tmem_ld32_issue(t_mir, r1);
            ptx::tmem_ld32_wait(r0);
            ptx::tmem_ld32_wait(r1);
#pragma unroll
            for (int j = 0; j < 32; ++j) v0[j] = __uint_as_float(r0[j]) + __uint_as_float(r1[j]);
        } else {
            ptx::tmem_ld32(t_main, v0);
        }
        pair_release<N>(t_main, t_mir, s_bias, lane, release_addr);
    }
    const bool store = inb && !(a.flags & FLAG_SKIP_B);
#pragma unroll
    for (int g = 0; g < kG; ++g) {
        float v[32];
        if constexpr (N == 32) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = v0[j];
        } else if (t_mir != 0xffffffffu) {
            uint32_t r0[32], r1[32];
            ptx::tmem_ld32_issue(t_main + g * 32, r0);
            ptx::tmem_ld32_issue(t_mir + g * 32, r1);
            ptx::tmem_ld32_wait(r0);
            ptx::tmem_ld32_wait(r1);
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r0[j]) + __uint_as_float(r1[j]);
        } else {
            ptx::tmem_ld32(t_main + g * 32, v);
        }
        epi_act32(v, amode, a.slope, s_neg + g * 32);
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
        if (gap) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0.f;
        }
        const uint4 h0 = pack8(v), h1 = pack8(v + 8), h2 = pack8(v + 16), h3 = pack8(v + 24);
        // N = 64: all TMEM reads are done after the last group; the position goes back while only the packed row is live
        if (N == 64 && g == kG - 1) pair_release<N>(t_main, t_mir, s_bias, lane, release_addr);
        if (store) {
            __half* o = a.out + chan_off(p, a.out_cstride, a.out_pstride, a.out_coff + g * 32);
            if (N == 64 && a.l2_hint >= 4) {  // a 64-channel output is the next layers' x: same keep-fraction policy as their loads
                const uint64_t pol = ptx::l2_policy_keep_fraction(a.l2_frac);
                ptx::stg256_hint(o, h0, h1, pol);
                ptx::stg256_hint(o + 16, h2, h3, pol);
            } else {
                ptx::stg256(o, h0, h1);
                ptx::stg256(o + 16, h2, h3);
            }
            if (kNoRes && a.out2) {  // second destination (conv_first): only layers without residual operands have one
                __half* o2 = a.out2 + chan_off(p, a.out2_cstride, a.out_pstride, a.out_coff + g * 32);
                ptx::stg256(o2, h0, h1);
                ptx::stg256(o2 + 16, h2, h3);
            }
        }
    }
    return true;
}

// Work decomposition: sub-items u = band * tiles_x + strip (strip fastest); cluster item i = sub-items 2i (rank 0) and 2i+1
// (rank 1). The two CTAs of a pair only have to run the same NUMBER of rows -- the leader's MMA sequence is shared, the image
// rows / strip each CTA loads and stores are its own -- so an odd strip count costs no padding strip (the tile atlas of the
// 6-tile preset is 13 strips wide). A CTA whose band is shorter than its peer's just runs extra discarded rows; a sub-item
// past the end is empty (strip beyond the image: zero-filled loads, no stores).
struct PairSub {
    int sx, y0, nrow;
};
__device__ __forceinline__ PairSub pair_sub(const ConvArgs& a, int u) {
    PairSub r;
    if (u >= a.tiles_x * a.nbands) {
        r.sx = a.tiles_x;
        r.y0 = a.y_begin;
        r.nrow = 0;
        return r;
    }
    const int b = u / a.tiles_x;
    r.sx = u - b * a.tiles_x;
    r.y0 = a.y_begin + b * a.band;
    r.nrow = (r.y0 + a.band < a.y_end ? r.y0 + a.band : a.y_end) - r.y0;
    return r;
}
__device__ __forceinline__ int pair_rows(const ConvArgs& a, int item) {  // rows the pair runs: the longer of the two bands
    const int n0 = pair_sub(a, 2 * item).nrow, n1 = pair_sub(a, 2 * item + 1).nrow;
    return n0 > n1 ? n0 : n1;
}

template <int N, bool kDirect, bool kNoRes = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(PairTraits<N>::kThreads, 1)
conv3x3_pair_kernel(const __grid_constant__ CUtensorMap tmap, const ConvArgs a) {
    using T = PairTraits<N>;
    constexpr uint32_t P = T::kPeriod;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t s_bars[2 * kPairMaxSlots + 2 * P + 1];
    __shared__ uint32_t s_tmem_slot;
    __shared__ __align__(16) float s_bias[N];
    __shared__ __align__(16) float s_neg[N];
    uint64_t* full = s_bars;                  // leader: both CTAs' boxes of a slot have landed
    uint64_t* empty = full + kPairMaxSlots;   // both: the MMAs reading the slot have retired (multicast commit)
    uint64_t* tfull = empty + kPairMaxSlots;  // both: logical ring position complete (multicast commit)
    uint64_t* tempty = tfull + P;             // leader: ring position drained and re-initialised in BOTH CTAs
    uint64_t* wfull = tempty + P;
    const int nslots = a.nstages;
    const int nch = a.nchunks;
    uint8_t* slot0 = smem + nch * T::kBHalf;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = ptx::cluster_ctarank();
    const long long t_start = clock64();

    if (threadIdx.x == 0) {
        for (int i = 0; i < kPairMaxSlots; ++i) {
            ptx::mbar_init(&full[i], 1);
            ptx::mbar_init(&empty[i], 1);
        }
        for (uint32_t i = 0; i < P; ++i) {
            ptx::mbar_init(&tfull[i], kMmaWarps);   // both issuer warps commit (multicast) their MMAs of the row
            ptx::mbar_init(&tempty[i], 8);          // four lane-quarter warps in each of the two CTAs
        }
        ptx::mbar_init(wfull, 1);
        ptx::fence_mbar_init();
    }
    if (warp == T::kEpi) {
        if (lane == 0) ptx::prefetch_tmap(&tmap);
        __syncwarp();
        ptx::tmem_alloc_pair<512>(&s_tmem_slot);
    }
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        s_bias[i] = (i < a.cout && a.bias) ? a.bias[i] : 0.f;
        float neg = 1.f;
        if (a.act == ACT_LRELU) neg = a.slope;
        if (a.act == ACT_PRELU) neg = (i < a.cout && a.prelu) ? a.prelu[i] : 0.f;
        s_neg[i] = neg;
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = s_tmem_slot;
    if (warp == T::kEpi && lane == 0) {
        // this CTA's half of the weight rows; weights are never written by a kernel: fetch before the dependency wait
        const __half* wp = a.wpack + (static_cast<size_t>(rank) * nch) * (T::kBHalf / 2);
        ptx::mbar_expect_tx(wfull, nch * T::kBHalf);
        for (int c = 0; c < nch; ++c) ptx::bulk_load(smem + c * T::kBHalf, wp + static_cast<size_t>(c) * (T::kBHalf / 2), T::kBHalf, wfull);
    }
    if (warp < T::kEpi) {
        // ring blocks start as the bias row, mirror blocks as zero; every MMA accumulates
        const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
#pragma unroll
        for (int g = 0; g < N / 32; ++g) {
            float bz[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) bz[j] = s_bias[g * 32 + j];
            for (uint32_t blk = warp >> 2; blk < static_cast<uint32_t>(T::kPhys); blk += T::kGroups) {
                if (blk < P) ptx::tmem_st32(tmem_base + lane_base + blk * N + g * 32, bz);
                else ptx::tmem_st32_zero(tmem_base + lane_base + blk * N + g * 32);
            }
        }
        ptx::tmem_st_wait();
    }
    ptx::mbar_wait(wfull, 0);
    ptx::tc_fence_before();
    ptx::cluster_sync();  // both CTAs: barriers initialised, TMEM initialised, weight halves resident
    ptx::tc_fence_after();
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");

    const int num_items = (a.tiles_x * a.nbands + 1) >> 1;
    const int cluster_id = static_cast<int>(blockIdx.x) >> 1, nclusters = static_cast<int>(gridDim.x) >> 1;

    if (warp == T::kEpi) {
        // ===================== TMA producer: this CTA's strip; bytes counted on the leader's barrier =====================
        if (lane == 0) {
            const uint32_t lead_full = ptx::map_to_rank(&full[0], 0);
            const uint64_t pol_keep = a.l2_hint >= 4 ? ptx::l2_policy_keep_fraction(a.l2_frac) : ptx::l2_policy_evict_last();
            const uint64_t pol_stream = ptx::l2_policy_evict_first();
            // chunks (source planes) below keep_end use pol_keep; l2_hint == 1 keeps the newest plane only
            const int keep_end = a.l2_hint == 1 ? nch : a.l2_hint <= 3 ? a.l2_hint - 1 : a.l2_hint == 4 ? 2 : a.l2_hint == 5 ? 3 : nch;
            const int keep_begin = a.l2_hint == 1 ? nch - 1 : 0;
            int s = 0;
            uint32_t ph = 0;
            for (int item = cluster_id; item < num_items; item += nclusters) {
                const PairSub me = pair_sub(a, 2 * item + static_cast<int>(rank));
                const int sx = me.sx, y0 = me.y0;
                const int nin2 = (pair_rows(a, item) + 3) & ~1;
                for (int j0 = 0; j0 < nin2; j0 += 2) {
                    for (int c = 0; c < nch; ++c) {
                        ptx::mbar_wait(&empty[s], ph ^ 1);
                        if (rank == 0) ptx::mbar_expect_tx(&full[s], 2 * T::kCopyBytes);
                        const int ch0 = a.cin_off + c * T::KC;
                        if (a.l2_hint)
                            ptx::tma_load_4d_pair_hint(slot0 + s * T::kASlot, &tmap, lead_full + s * 8, a.in_cstride == 32 ? 0 : ch0,
                                                       sx * 128 - 1, y0 - 1 + j0, a.in_cstride == 32 ? ch0 >> 5 : 0,
                                                       (c >= keep_begin && c < keep_end) ? pol_keep : pol_stream);
                        else
                            ptx::tma_load_4d_pair(slot0 + s * T::kASlot, &tmap, lead_full + s * 8, a.in_cstride == 32 ? 0 : ch0,
                                                  sx * 128 - 1, y0 - 1 + j0, a.in_cstride == 32 ? ch0 >> 5 : 0);
                        if (++s == nslots) { s = 0; ph ^= 1; }
                    }
                }
            }
        }
    } else if (warp > T::kEpi) {
        // ===================== MMA issuers: leader only =====================
        // Two warps alternate UNITS of `a.unit` consecutive boxes (12 MMAs each). The owner of a unit first waits for every
        // barrier the unit needs (operands landed, ring positions drained) while the other warp is still issuing, takes over
        // through a named barrier, and then ONE elected lane walks the whole unit: nothing but address arithmetic between the
        // MMAs. Both warps walk all boxes so that each can commit its own MMAs to the accumulator-ready barriers.
        if (rank == 0) {
            const int mw = warp - (T::kEpi + 1);
            const int unit = a.unit < 1 ? 1 : a.unit;
            const bool skip_mma = (a.flags & FLAG_SKIP_MMA) != 0;
            int s = 0;
            uint32_t ph = 0;
            int gunit = 0;
            uint32_t g0 = 0;  // logical rows started before the current item
            for (int item = cluster_id; item < num_items; item += nclusters) {
                const int nin2 = (pair_rows(a, item) + 3) & ~1;
                const int nb = (nin2 >> 1) * nch;  // boxes of the item: row pairs x chunks, chunk fastest
                int j0 = 0, c = 0;
                for (int n = 0; n < nb; n += unit, ++gunit) {
                    const int cnt = nb - n < unit ? nb - n : unit;
                    const bool mine = (gunit & 1) == mw;
                    if (mine) {
                        int ss = s, cc = c, jj = j0;
                        uint32_t pp = ph;
                        for (int i = 0; i < cnt; ++i) {
                            ptx::mbar_wait(&full[ss], pp);
                            if (cc == 0) {
                                // logical rows first touched by this row pair: ga + 2, ga + 3 (and ga, ga + 1 at the top of an item)
                                const uint32_t ga = g0 + jj;
                                for (uint32_t gl = (jj == 0 ? ga : ga + 2); gl < ga + 4; ++gl)
                                    ptx::mbar_wait(&tempty[gl % P], ((gl / P) & 1u) ^ 1u);
                            }
                            if (++ss == nslots) { ss = 0; pp ^= 1; }
                            if (++cc == nch) { cc = 0; jj += 2; }
                        }
                        ptx::tc_fence_after();
                        if (gunit > 0) asm volatile("bar.sync %0, 64;" ::"r"(1 + mw) : "memory");
                    }
                    if (ptx::elect_one()) {
                        int ss = s, cc = c, jj = j0;
                        for (int i = 0; i < cnt; ++i) {
                            const uint32_t ga = g0 + jj;
                            const uint32_t s0 = ga % P, s1 = (ga + 1) % P;
                            if (mine) {
                                if (!skip_mma)
                                    pair_issue_box<N>(tmem_base + s0 * N, tmem_base + s1 * N, ptx::smem_u32(slot0 + ss * T::kASlot) >> 4,
                                                      ptx::smem_u32(smem + cc * T::kBHalf) >> 4);
                                ptx::umma_commit_pair(&empty[ss]);
                            }
                            if (cc == nch - 1) {
                                // logical rows ga, ga + 1 have their last tap (after the last row pair also the trailing phantom rows):
                                // this warp's commits are one of the two arrivals on each barrier
                                ptx::umma_commit_pair(&tfull[s0]);
                                ptx::umma_commit_pair(&tfull[s1]);
                                if (jj + 2 >= nin2) {
                                    ptx::umma_commit_pair(&tfull[(ga + 2) % P]);
                                    ptx::umma_commit_pair(&tfull[(ga + 3) % P]);
                                }
                            }
                            if (++ss == nslots) ss = 0;
                            if (++cc == nch) { cc = 0; jj += 2; }
                        }
                    }
                    __syncwarp();
                    if (mine) asm volatile("bar.arrive %0, 64;" ::"r"(2 - mw) : "memory");
                    for (int i = 0; i < cnt; ++i) {
                        if (++s == nslots) { s = 0; ph ^= 1; }
                        if (++c == nch) { c = 0; j0 += 2; }
                    }
                }
                g0 += nin2 + 2;
            }
            // the last unit's arrive has no matching sync: consume it so no named barrier is left half-arrived
            if (gunit > 0 && (gunit & 1) == mw) asm volatile("bar.sync %0, 64;" ::"r"(1 + mw) : "memory");
        }
    } else {
        // ===================== epilogue warps (both CTAs): warp % 4 = TMEM lane quarter, warp / 4 = row group =====================
        const int quarter = warp & 3;
        const uint32_t rgrp = warp >> 2;
        const int amode = a.act == ACT_NONE ? 0 : ((a.act == ACT_LRELU && a.slope >= 0.f && a.slope <= 1.f) ? 1 : 2);
        const uint32_t stg_s = ptx::smem_u32(slot0 + nslots * T::kASlot + warp * (32 * N * 2));
        const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
        const uint32_t lead_tempty = ptx::map_to_rank(&tempty[0], 0);
        uint32_t g0 = 0;
        for (int item = cluster_id; item < num_items; item += nclusters) {
            const PairSub me = pair_sub(a, 2 * item + static_cast<int>(rank));
            const int sx = me.sx, y0 = me.y0, nrow = me.nrow;  // this CTA's own strip / band: real rows are l in [2, nrow + 2)
            const int nin2 = (pair_rows(a, item) + 3) & ~1;
            const int lrows = nin2 + 2;
            const int x_base = sx * 128 + quarter * 32;
            const int x = x_base + lane;
            bool xgap = false;
            for (int j = 0; j < a.ngx; ++j) xgap |= ((x >> a.gshift) == a.gx[j]);
#pragma unroll 1
            for (int l = static_cast<int>((rgrp + T::kGroups - g0 % T::kGroups) % T::kGroups); l < lrows; l += T::kGroups) {
                const uint32_t gl = g0 + l;
                const uint32_t m = gl % P;
                ptx::mbar_wait(&tfull[m], (gl / P) & 1u);
                ptx::tc_fence_after();
                const uint32_t t_main = tmem_base + lane_base + m * N;
                const uint32_t t_mir = m < 2 ? tmem_base + lane_base + (P + m) * N : 0xffffffffu;
                const bool real = l >= 2 && l < nrow + 2 && !(a.flags & FLAG_SKIP_EPI);
                bool released = false;
                if (real) {
                    const int y = y0 + l - 2;
                    bool gap = xgap;
                    for (int j = 0; j < a.ngy; ++j) gap |= ((y >> a.gshift) == a.gy[j]);
                    if constexpr (kDirect) released = epi_row_pair_direct<N, kNoRes>(a, t_main, t_mir, lane, x_base, y, gap, s_bias, s_neg, amode, lead_tempty + m * 8);
                    else released = epi_row_pair<N>(a, t_main, t_mir, stg_s, lane, x_base, y, gap, s_bias, s_neg, amode, lead_tempty + m * 8);
                }
                if (!released) pair_release<N>(t_main, t_mir, s_bias, lane, lead_tempty + m * 8);
            }
            g0 += lrows;
        }
    }

    ptx::tc_fence_before();
    ptx::cluster_sync();  // the peer may still be read (operands) or signalled (barriers) until both are done
    if (warp == T::kEpi) {
        __syncwarp();
        ptx::tmem_dealloc_pair<512>(tmem_base);
    }
    if (a.dbg_cycles && threadIdx.x == 0) a.dbg_cycles[blockIdx.x] = clock64() - t_start;
}

}  // namespace vr
