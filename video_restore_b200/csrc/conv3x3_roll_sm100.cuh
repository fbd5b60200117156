// K2: rolling-row 3x3 convolution on tcgen05 tensor cores (sm_100a) -- the NHWC body layers of RRDBNet / SRVGGNetCompact
// (same layers and reference call sites as K1, conv3x3_sm100.cuh; K1 keeps the RGB / pixel-shuffle outputs, the 2x2-tap
// upsample phases and the multi-layer launch).
//
// Why a second formulation. K1 computes TH x 128 output tiles from a (TH+2)-row haloed input tile: every input row is
// fetched 1.5x (TH = 4), the layer's weights are re-streamed for every tile unless they fit beside >= 3 stages, and the
// first / last two input rows of a tile feed only one / two output rows, so their dy-stacked MMAs run at N = Cout / 2 Cout
// where the SS MMA is shared-memory-read bound (4 KB of A per MMA regardless of N): 288 issue cycles for 192 cycles of
// tensor work on the Cout = 32 layers. K2 removes all three:
//   * a CTA owns a 128-pixel-wide strip of `band` output rows and walks DOWN it: input row rho is fetched once (one TMA
//     box of 32 ch x 130 px x 1 row per channel chunk) and feeds output rows rho-1, rho, rho+1 with ONE MMA of N = 3 Cout
//     per (chunk, dx, k16) -- full N on every row except the first / last two of the band (band ~ 50..100 rows);
//   * accumulators are a RING of 512 / Cout output rows in TMEM. Output row r is first touched (accumulate = 0) by input
//     row r-1 and complete after input row r+1; the epilogue drains it while the MMAs run 2..R-3 rows ahead;
//   * the whole layer's weights (up to 110 KB) stay resident in shared memory for the CTA's lifetime; a 192 -> 64 layer
//     (221 KB) is computed as two independent 32-channel halves (work items x2, each half resident).
// Shared memory: [weights nchunks x 9 N x 64 B][ring of <= 16 activation slots, 130 px x 64 B each][epilogue staging].
// Warp roles as in K1: 0..7 epilogue (warp % 4 = TMEM lane quarter, warp / 4 = row parity), 8 = TMA producer, 9 / 10 =
// MMA issuers alternating slots. Same operand layouts, descriptors and epilogue (epi_row_nhwc) as K1.
#pragma once
#include "conv3x3_sm100.cuh"

namespace vr {

constexpr int kRollMaxSlots = 16;

template <int N>
struct RollTraits {
    static_assert(N == 32 || N == 64, "rolling kernel: 32 or 64 output channels per pass");
    static constexpr int KC = 32;
    static constexpr int kRowBytes = KC * 2;
    static constexpr int kPitch = 130;
    static constexpr int kCopyBytes = kPitch * kRowBytes;         // one input row of one channel chunk
    static constexpr int kASlot = round_up_c(kCopyBytes, 512);    // SWIZZLE_64B atom = 8 rows x 64 B
    static constexpr int kBBytes = 9 * N * kRowBytes;             // one chunk's weights, all taps
    static constexpr int kBStage = round_up_c(kBBytes, 1024);
    static constexpr int kRing = 512 / N;                         // output rows resident in TMEM
    static constexpr int kStgBytes = kEpiWarps * 32 * N * 2;
    static constexpr int kStatic = 2048;
    static constexpr int kBudget = 227 * 1024 - 1024 - kStatic - kStgBytes;  // weights + activation slots
    static constexpr int kMinSlots = 4;
};

template <int N>
__global__ void __launch_bounds__(kConvThreads, 1)
conv3x3_roll_kernel(const __grid_constant__ CUtensorMap tmap, const ConvArgs a) {
    using T = RollTraits<N>;
    constexpr int R = T::kRing;
    constexpr uint32_t kDescHi = ptx::kDescHiSw64;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t s_bars[2 * kRollMaxSlots + 2 * R + 1];
    __shared__ uint32_t s_tmem_slot;
    __shared__ __align__(16) float s_bias[2 * N];  // both halves of a split layer
    __shared__ __align__(16) float s_neg[2 * N];
    uint64_t* full = s_bars;
    uint64_t* empty = full + kRollMaxSlots;
    uint64_t* tfull = empty + kRollMaxSlots;
    uint64_t* tempty = tfull + R;
    uint64_t* wfull = tempty + R;
    const int nslots = a.nstages;
    const int nch = a.nchunks;
    uint8_t* slot0 = smem + nch * T::kBStage;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const long long t_start = clock64();

    if (threadIdx.x == 0) {
        for (int i = 0; i < kRollMaxSlots; ++i) {
            ptx::mbar_init(&full[i], 1);
            ptx::mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < R; ++i) {
            ptx::mbar_init(&tfull[i], kMmaWarps);       // every issuer commits its own MMAs of the row
            ptx::mbar_init(&tempty[i], kEpiWarps / 2);  // the four lane-quarter warps of the row's parity group
        }
        ptx::mbar_init(wfull, 1);
        ptx::fence_mbar_init();
    }
    if (warp == kEpiWarps) {
        if (lane == 0) ptx::prefetch_tmap(&tmap);
        __syncwarp();
        ptx::tmem_alloc<512>(&s_tmem_slot);
    }
    const int ctot = N * a.nsplit;
    for (int i = threadIdx.x; i < 2 * N; i += blockDim.x) {
        s_bias[i] = (i < ctot && i < a.cout && a.bias) ? a.bias[i] : 0.f;
        float neg = 1.f;
        if (a.act == ACT_LRELU) neg = a.slope;
        if (a.act == ACT_PRELU) neg = (i < ctot && i < a.cout && a.prelu) ? a.prelu[i] : 0.f;
        s_neg[i] = neg;
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = s_tmem_slot;
    const int tiles_x = a.tiles_x;
    const int num_items = tiles_x * a.nbands * a.nsplit;
    // the host launches a grid that is a multiple of nsplit, so a CTA's items all belong to one channel half
    const int half = static_cast<int>(blockIdx.x) % a.nsplit;
    if (warp == kEpiWarps && lane == 0) {
        // weights are never written by a kernel: fetch them before waiting on the previous layer
        const __half* wp = a.wpack + static_cast<size_t>(half) * nch * 9 * N * T::KC;
        ptx::mbar_expect_tx(wfull, nch * T::kBBytes);
        for (int c = 0; c < nch; ++c)
            ptx::bulk_load(smem + c * T::kBStage, wp + static_cast<size_t>(c) * 9 * N * T::KC, T::kBBytes, wfull);
    }
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");

    if (warp == kEpiWarps) {
        // ===================== TMA producer: one box per (input row, channel chunk) =====================
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
                const int rest = item / a.nsplit;
                const int b = rest / tiles_x, sx = rest - b * tiles_x;
                const int y0 = a.y_begin + b * a.band;
                const int y1 = y0 + a.band < a.y_end ? y0 + a.band : a.y_end;
                const int nin = y1 - y0 + 2;
                for (int j = 0; j < nin; ++j) {
                    for (int c = 0; c < nch; ++c) {
                        ptx::mbar_wait(&empty[s], ph ^ 1);
                        if (a.flags & FLAG_SKIP_TMA) {
                            ptx::mbar_arrive(&full[s]);
                        } else {
                            ptx::mbar_expect_tx(&full[s], T::kCopyBytes);
                            ptx::tma_load_4d(slot0 + s * T::kASlot, &tmap, &full[s], a.cin_off + c * T::KC, sx * 128 - 1,
                                             y0 - 1 + j, 0);
                        }
                        if (++s == nslots) { s = 0; ph ^= 1; }
                    }
                }
            }
        }
    } else if (warp > kEpiWarps) {
        // ===================== MMA issuers (two warps, alternating slots; see K1 for the measurements) =====================
        const int mw = warp - (kEpiWarps + 1);
        ptx::mbar_wait(wfull, 0);
        const bool skip_mma = (a.flags & FLAG_SKIP_MMA) != 0;
        int s = 0;
        uint32_t ph = 0;
        int gstage = 0;
        int g0 = 0;  // output rows this CTA has started before the current item (TMEM ring position)
        for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
            const int b = (item / a.nsplit) / tiles_x;
            const int y0 = a.y_begin + b * a.band;
            const int nrow = (y0 + a.band < a.y_end ? y0 + a.band : a.y_end) - y0;
            for (int j = 0; j < nrow + 2; ++j) {
                // input row j of the band (image row y0 - 1 + j) feeds output rows q = j-2 (dy 2), j-1 (dy 1), j (dy 0)
                const int qa = j >= 2 ? j - 2 : 0;
                const int qb = j < nrow ? j : nrow - 1;
                const bool first_touch = j < nrow;  // output row j starts with this input row
                const int dy_hi = j - qa;
                const int nblk = qb - qa + 1;
                const int ba = (g0 + qa) % R;
                const int n1 = nblk < R - ba ? nblk : R - ba;  // blocks before the ring wraps
                const int n2 = nblk - n1;
                for (int c = 0; c < nch; ++c) {
                    const bool mine = (gstage & 1) == mw;
                    if (mine) {
                        ptx::mbar_wait(&full[s], ph);
                        if (c == 0 && first_touch) {
                            const int gt = g0 + j;  // the ring block must have been drained of output row gt - R
                            ptx::mbar_wait(&tempty[gt % R], ((gt / R) & 1) ^ 1);
                        }
                        ptx::tc_fence_after();
                        if (gstage > 0) asm volatile("bar.sync %0, 64;" ::"r"(1 + mw) : "memory");
                    }
                    if (mine && ptx::elect_one()) {
                        const uint32_t a_lo0 = ptx::smem_u32(slot0 + s * T::kASlot) >> 4;
                        const uint32_t b_lo0 = (ptx::smem_u32(smem + c * T::kBStage) >> 4) + (((2 - dy_hi) * N * T::kRowBytes) >> 4);
                        if (!skip_mma) {
#pragma unroll
                            for (int dx = 0; dx < 3; ++dx) {
#pragma unroll
                                for (int k = 0; k < 2; ++k) {
                                    const uint32_t a_lo = a_lo0 + ((dx * T::kRowBytes + k * 32) >> 4);
                                    const uint32_t b_lo = b_lo0 + ((dx * 3 * N * T::kRowBytes + k * 32) >> 4);
                                    if (dx == 0 && k == 0 && c == 0 && first_touch) {
                                        // the newest output row (last block) must overwrite: one MMA per block
                                        for (int i = 0; i < nblk; ++i)
                                            ptx::umma_f16<ptx::kCollNone>(tmem_base + ((ba + i) % R) * N, a_lo, kDescHi,
                                                                          b_lo + ((i * N * T::kRowBytes) >> 4), kDescHi,
                                                                          ptx::make_idesc_f16(128, N), i == nblk - 1 ? 0u : 1u);
                                    } else {
                                        ptx::umma_f16<ptx::kCollNone>(tmem_base + ba * N, a_lo, kDescHi, b_lo, kDescHi,
                                                                      ptx::make_idesc_f16(128, n1 * N), 1u);
                                        if (n2)
                                            ptx::umma_f16<ptx::kCollNone>(tmem_base, a_lo, kDescHi,
                                                                          b_lo + ((n1 * N * T::kRowBytes) >> 4), kDescHi,
                                                                          ptx::make_idesc_f16(128, n2 * N), 1u);
                                    }
                                }
                            }
                        }
                        ptx::umma_commit(&empty[s]);
                    }
                    __syncwarp();
                    if (mine) asm volatile("bar.arrive %0, 64;" ::"r"(2 - mw) : "memory");
                    ++gstage;
                    if (++s == nslots) { s = 0; ph ^= 1; }
                }
                // output row j-2 has received its last tap: this warp's commit is one of the two arrivals on its barrier
                if (j >= 2) {
                    if (ptx::elect_one()) ptx::umma_commit(&tfull[(g0 + j - 2) % R]);
                    __syncwarp();
                }
            }
            g0 += nrow;
        }
        if (gstage > 0 && (gstage & 1) == mw) asm volatile("bar.sync %0, 64;" ::"r"(1 + mw) : "memory");
    } else {
        // ===================== epilogue warps 0..7 =====================
        const int quarter = warp & 3;
        const int rgrp = warp >> 2;  // ring blocks (== output rows in ring order) of this parity
        const int amode = a.act == ACT_NONE ? 0 : ((a.act == ACT_LRELU && a.slope >= 0.f && a.slope <= 1.f) ? 1 : 2);
        const uint32_t stg_s = ptx::smem_u32(slot0 + nslots * T::kASlot + warp * (32 * N * 2));
        const int coff_add = half * N;
        int g0 = 0;
        for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
            const int rest = item / a.nsplit;
            const int b = rest / tiles_x, sx = rest - b * tiles_x;
            const int y0 = a.y_begin + b * a.band;
            const int nrow = (y0 + a.band < a.y_end ? y0 + a.band : a.y_end) - y0;
            const int x_base = sx * 128 + quarter * 32;
            const int x = x_base + lane;
            bool xgap = false;
            for (int j = 0; j < a.ngx; ++j) xgap |= ((x >> a.gshift) == a.gx[j]);
#pragma unroll 1
            for (int q = ((g0 & 1) == rgrp ? 0 : 1); q < nrow; q += 2) {
                const int g = g0 + q;
                const int blk = g % R;
                ptx::mbar_wait(&tfull[blk], (g / R) & 1);
                ptx::tc_fence_after();
                if (!(a.flags & FLAG_SKIP_EPI)) {
                    const int y = y0 + q;
                    bool gap = xgap;
                    for (int j = 0; j < a.ngy; ++j) gap |= ((y >> a.gshift) == a.gy[j]);
                    epi_row_nhwc<N>(a, tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + blk * N, stg_s, lane, x_base, y, gap,
                                    a.out_coff, coff_add, s_bias + coff_add, s_neg + coff_add, amode);
                }
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&tempty[blk]);
            }
            g0 += nrow;
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == kEpiWarps) {
        __syncwarp();
        ptx::tmem_dealloc<512>(tmem_base);
    }
    if (a.dbg_cycles && threadIdx.x == 0) a.dbg_cycles[blockIdx.x] = clock64() - t_start;
}

}  // namespace vr
