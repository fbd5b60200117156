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
//     box of 32 ch x 130 px x 2 rows per channel chunk) and feeds output rows rho-1, rho, rho+1 with ONE MMA of N = 3 Cout
//     per (chunk, dx, k16) -- full N on every row except the first / last two of the band (band ~ 50..100 rows);
//   * accumulators are a RING of 512 / Cout output rows in TMEM. Output row r is first touched by input row r-1 and
//     complete after input row r+1; the epilogue drains it while the MMAs run 2..R-3 rows ahead and hands the block back
//     re-initialised with the BIAS row (tcgen05.st), so every MMA accumulates and the issue sequence has no special first step;
//   * the whole layer's weights (up to 110 KB) stay resident in shared memory for the CTA's lifetime; a 192 -> 64 layer
//     (221 KB) is computed as two independent 32-channel halves (work items x2, each half resident).
// Shared memory: [weights nchunks x 9 N x 64 B][ring of <= 8 activation slots, 2 rows x 130 px x 64 B each][epilogue staging].
// Warp roles as in K1: 0..7 epilogue (warp % 4 = TMEM lane quarter, warp / 4 = row parity), 8 = TMA producer, 9 / 10 =
// MMA issuers alternating slots. Same operand layouts and descriptors as K1; epilogue epi_row_nhwc_folded (bias lives in the
// accumulators). Since K3 (conv3x3_pair_sm100.cuh) this kernel is the fall-back and the A/B baseline.
#pragma once
#include "conv3x3_sm100.cuh"

namespace vr {

constexpr int kRollMaxSlots = 8;

template <int N>
struct RollTraits {
    static_assert(N == 32 || N == 64, "rolling kernel: 32 or 64 output channels per pass");
    static constexpr int KC = 32;
    static constexpr int kRowBytes = KC * 2;
    static constexpr int kPitch = 130;
    static constexpr int kBoxRows = 2;                            // input rows per TMA box = per issuer hand-over (12 MMAs)
    static constexpr int kLineBytes = kPitch * kRowBytes;         // one input row of one channel chunk
    static constexpr int kCopyBytes = kBoxRows * kLineBytes;
    static constexpr int kASlot = round_up_c(kCopyBytes, 512);    // SWIZZLE_64B atom = 8 rows x 64 B
    static constexpr int kBBytes = 9 * N * kRowBytes;             // one chunk's weights, all taps
    static constexpr int kBStage = round_up_c(kBBytes, 1024);
    static constexpr int kRing = 512 / N;                         // output rows resident in TMEM
    static constexpr int kStgBytes = kEpiWarps * 32 * N * 2;
    static constexpr int kStatic = 2048;
    static constexpr int kBudget = 227 * 1024 - 1024 - kStatic - kStgBytes;  // weights + activation slots
    static constexpr int kMinSlots = 3;
};

// Per input row of a band: which output rows (TMEM ring blocks) its dy-stacked MMA updates.
struct RollRow {
    uint32_t d_col;     // first ring block * N is added by the caller: ring block of the lowest output row
    uint32_t n1, n2;    // blocks before / after the ring wraps
    uint32_t b_row;     // first B row block: 2 - dy_hi
    uint32_t gt;        // global output-row index of the first-touched row (valid when first_touch)
    uint32_t nblk;
    bool first_touch;   // output row j starts with this input row: its ring block must have been drained (and zeroed)
};
template <int R>
__device__ __forceinline__ RollRow roll_row(int j, int nrow, uint32_t g0) {
    RollRow r;
    const int qa = j >= 2 ? j - 2 : 0;
    const int qb = j < nrow ? j : nrow - 1;
    r.first_touch = j < nrow;
    r.nblk = static_cast<uint32_t>(qb - qa + 1);
    r.b_row = static_cast<uint32_t>(2 - (j - qa));
    r.d_col = (g0 + qa) & (R - 1);
    r.n1 = r.nblk < R - r.d_col ? r.nblk : R - r.d_col;
    r.n2 = r.nblk - r.n1;
    r.gt = g0 + j;
    return r;
}
// Generic path (band edges, ring wrap): the six MMAs (3 dx x 2 k16) of one input row, N = nblk * Cout across the row's ring
// window, split where the ring wraps. Every MMA accumulates: the epilogue hands ring blocks back zeroed.
template <int N>
__device__ __forceinline__ void roll_issue_row(const RollRow& r, uint32_t tmem_base, uint32_t a_lo0, uint32_t b_lo0) {
    constexpr uint32_t kDescHi = ptx::kDescHiSw64;
    const uint32_t b_base = b_lo0 + ((r.b_row * N * 64) >> 4);
    const uint32_t d = tmem_base + r.d_col * N;
    const uint32_t idesc1 = ptx::make_idesc_f16(128, r.n1 * N), idesc2 = ptx::make_idesc_f16(128, (r.n2 ? r.n2 : 1) * N);
#pragma unroll
    for (int t = 0; t < 6; ++t) {
        const uint32_t a_lo = a_lo0 + (((t >> 1) * 64 + (t & 1) * 32) >> 4);
        const uint32_t b_lo = b_base + (((t >> 1) * 3 * N * 64 + (t & 1) * 32) >> 4);
        ptx::umma_f16<ptx::kCollNone>(d, a_lo, kDescHi, b_lo, kDescHi, idesc1, 1u);
        if (r.n2) ptx::umma_f16<ptx::kCollNone>(tmem_base, a_lo, kDescHi, b_lo + ((r.n1 * N * 64) >> 4), kDescHi, idesc2, 1u);
    }
}
// Fast path: an interior box (both input rows feed three output rows, no ring wrap inside the four-block window). Twelve
// branch-free MMAs whose operands are one runtime base plus compile-time offsets. With dependent accumulation chains the
// tensor pipe's issue queue is shallow (tools/issue_probe): any scalar work between two MMAs is exposed, so there is none.
template <int N>
__device__ __forceinline__ void roll_issue_box_fast(uint32_t d0, uint32_t a_lo0, uint32_t b_lo0) {
    constexpr uint32_t kDescHi = ptx::kDescHiSw64;
    constexpr uint32_t kIdesc = ptx::make_idesc_f16(128, 3 * N);
    constexpr uint32_t kLine = RollTraits<N>::kLineBytes >> 4;
#pragma unroll
    for (int t = 0; t < 6; ++t) {
        const uint32_t ao = ((t >> 1) * 64 + (t & 1) * 32) >> 4, bo = ((t >> 1) * 3 * N * 64 + (t & 1) * 32) >> 4;
        ptx::umma_f16<ptx::kCollNone>(d0, a_lo0 + ao, kDescHi, b_lo0 + bo, kDescHi, kIdesc, 1u);
        ptx::umma_f16<ptx::kCollNone>(d0 + N, a_lo0 + kLine + ao, kDescHi, b_lo0 + bo, kDescHi, kIdesc, 1u);
    }
}

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
    if (warp < kEpiWarps) {
        // all MMAs accumulate, so every ring block starts as the bias row (and is re-initialised by the epilogue after each
        // drain): no accumulate = 0 special step in the issue sequence, no bias add in the epilogue
        const float* bsrc = s_bias + (static_cast<int>(blockIdx.x) % a.nsplit) * N;
        const uint32_t t0 = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16) + (warp >> 2) * 256;
#pragma unroll
        for (int g = 0; g < N / 32; ++g) {
            float bz[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) bz[j] = bsrc[g * 32 + j];
            for (int blk = 0; blk < R / 2; ++blk) ptx::tmem_st32(t0 + blk * N + g * 32, bz);
        }
        ptx::tmem_st_wait();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
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

    if (threadIdx.x == 0 && (a.flags & FLAG_TRACE) && a.dbg_cycles && blockIdx.x == 0) a.dbg_cycles[1003] = clock64() - t_start;  // prologue done
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
                for (int j = 0; j < nin; j += T::kBoxRows) {
                    for (int c = 0; c < nch; ++c) {
                        ptx::mbar_wait(&empty[s], ph ^ 1);
                        if (a.flags & FLAG_SKIP_TMA) {
                            ptx::mbar_arrive(&full[s]);
                        } else {
                            ptx::mbar_expect_tx(&full[s], T::kCopyBytes);
                            const int ch0 = a.cin_off + c * T::KC;
                            ptx::tma_load_4d(slot0 + s * T::kASlot, &tmap, &full[s], a.in_cstride == 32 ? 0 : ch0, sx * 128 - 1,
                                             y0 - 1 + j, a.in_cstride == 32 ? ch0 >> 5 : 0);
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
        const bool tracing = (a.flags & FLAG_TRACE) && a.dbg_cycles && blockIdx.x == 0;
        int s = 0;
        uint32_t ph = 0;
        int gstage = 0;
        uint32_t g0 = 0;  // output rows this CTA has started before the current item (TMEM ring position)
        for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
            const int b = (item / a.nsplit) / tiles_x;
            const int y0 = a.y_begin + b * a.band;
            const int nrow = (y0 + a.band < a.y_end ? y0 + a.band : a.y_end) - y0;
            const int nin = nrow + 2;
            for (int j0 = 0; j0 < nin; j0 += T::kBoxRows) {
                // input row j of the band (image row y0 - 1 + j) feeds output rows q = j-2 (dy 2), j-1 (dy 1), j (dy 0); a box
                // holds the rows j0 and j0 + 1 of one channel chunk
                RollRow r0 = roll_row<R>(j0, nrow, g0), r1 = roll_row<R>(j0 + 1, nrow, g0);
                const bool two = j0 + 1 < nin;
                const bool fast = two && r0.nblk == 3 && r1.nblk == 3 && r0.d_col + 4 <= static_cast<uint32_t>(R);
                for (int c = 0; c < nch; ++c) {
                    const bool mine = (gstage & 1) == mw;
                    const bool trace = tracing && mine && gstage >= 16 && gstage < 80;
                    long long* tr = a.dbg_cycles + 256 + mw * 128 + ((gstage - 16) >> 1) * 4;
                    if (trace && lane == 0) tr[0] = clock64() - t_start;
                    if (mine) {
                        ptx::mbar_wait(&full[s], ph);
                        if (trace && lane == 0) tr[1] = clock64() - t_start;
                        if (c == 0) {
                            // a ring block must have been drained of output row g - R before its first touch
                            if (r0.first_touch) ptx::mbar_wait(&tempty[r0.gt & (R - 1)], ((r0.gt / R) & 1u) ^ 1u);
                            if (two && r1.first_touch) ptx::mbar_wait(&tempty[r1.gt & (R - 1)], ((r1.gt / R) & 1u) ^ 1u);
                        }
                        ptx::tc_fence_after();
                        if (gstage > 0) asm volatile("bar.sync %0, 64;" ::"r"(1 + mw) : "memory");
                    }
                    if (mine && ptx::elect_one()) {
                        const uint32_t a_lo0 = ptx::smem_u32(slot0 + s * T::kASlot) >> 4;
                        const uint32_t b_lo0 = ptx::smem_u32(smem + c * T::kBStage) >> 4;
                        if (trace) tr[2] = clock64() - t_start;
                        if (tracing && gstage == 0) a.dbg_cycles[1000] = clock64() - t_start;  // first MMA of the CTA
                        if (!skip_mma) {
                            if (fast) {
                                roll_issue_box_fast<N>(tmem_base + r0.d_col * N, a_lo0, b_lo0);
                            } else {
                                roll_issue_row<N>(r0, tmem_base, a_lo0, b_lo0);
                                if (two) roll_issue_row<N>(r1, tmem_base, a_lo0 + (T::kLineBytes >> 4), b_lo0);
                            }
                        }
                        ptx::umma_commit(&empty[s]);
                        if (trace) tr[3] = clock64() - t_start;
                    }
                    __syncwarp();
                    if (mine) asm volatile("bar.arrive %0, 64;" ::"r"(2 - mw) : "memory");
                    ++gstage;
                    if (++s == nslots) { s = 0; ph ^= 1; }
                }
                // output rows j0-2 (and j0-1) have received their last tap: this warp's commits are one of the two arrivals each
                if (ptx::elect_one()) {
                    if (j0 >= 2) ptx::umma_commit(&tfull[(g0 + j0 - 2) & (R - 1)]);
                    if (two && j0 >= 1) ptx::umma_commit(&tfull[(g0 + j0 - 1) & (R - 1)]);
                }
                __syncwarp();
            }
            g0 += nrow;
        }
        if (gstage > 0 && (gstage & 1) == mw) asm volatile("bar.sync %0, 64;" ::"r"(1 + mw) : "memory");
        if (tracing && lane == 0) a.dbg_cycles[1001 + mw] = clock64() - t_start;  // issuer done
    } else {
        // ===================== epilogue warps 0..7 =====================
        const int quarter = warp & 3;
        const int rgrp = warp >> 2;  // ring blocks (== output rows in ring order) of this parity
        const int amode = a.act == ACT_NONE ? 0 : ((a.act == ACT_LRELU && a.slope >= 0.f && a.slope <= 1.f) ? 1 : 2);
        const uint32_t stg_s = ptx::smem_u32(slot0 + nslots * T::kASlot + warp * (32 * N * 2));
        const int coff_add = half * N;
        uint32_t g0 = 0;
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
            for (int q = ((g0 & 1u) == static_cast<uint32_t>(rgrp) ? 0 : 1); q < nrow; q += 2) {
                const uint32_t g = g0 + q;
                const uint32_t blk = g & (R - 1);
                const bool etrace = (a.flags & FLAG_TRACE) && a.dbg_cycles && blockIdx.x == 0 && quarter == 0 && lane == 0 && g >= 8 &&
                                    g < 40;
                long long* etr = a.dbg_cycles + 512 + rgrp * 128 + ((g - 8) >> 1) * 8;
                if (etrace) etr[0] = clock64() - t_start;
                ptx::mbar_wait(&tfull[blk], (g / R) & 1u);
                ptx::tc_fence_after();
                if (etrace) etr[1] = clock64() - t_start;
                const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + blk * N;
                if (!(a.flags & FLAG_SKIP_EPI)) {
                    const int y = y0 + q;
                    bool gap = xgap;
                    for (int j = 0; j < a.ngy; ++j) gap |= ((y >> a.gshift) == a.gy[j]);
                    epi_row_nhwc_folded<N>(a, t_addr, stg_s, lane, x_base, y, gap, a.out_coff, coff_add, s_bias + coff_add,
                                           s_neg + coff_add, amode, &tempty[blk], etrace ? etr : nullptr, t_start);
                } else {
                    // timing ablation: hand the block back untouched
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(&tempty[blk]);
                }
                if (etrace) etr[5] = clock64() - t_start;
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
