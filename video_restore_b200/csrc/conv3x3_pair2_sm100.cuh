// K4: TWO consecutive 32-channel layers of a dense block (conv_k and conv_k+1) in ONE launch of the CTA-pair rolling-row
// kernel K3 (conv3x3_pair_sm100.cuh) -- the dense block's HBM bytes, not its MMAs, set the sustained frame rate.
//
// Layer A = conv_k reads planes x, x1..x_{k-1} and produces x_k; layer B = conv_{k+1} reads the SAME planes plus x_k. In
// separate launches B re-reads everything from HBM (each plane is 59 MB at 720p, the L2 holds two). Here a CTA pair walks its
// (strip, band) once for both layers, B two row pairs behind A:
//   * B's activation boxes of x .. x_{k-1} are the rows A fetched a few microseconds earlier: L2 hits, no DRAM traffic;
//   * B's x_k operand never leaves the SM: A's epilogue writes each finished x_k row (fp16, after bias + LeakyReLU + gap
//     zeroing) straight into a shared-memory HAND-OFF slot in the swizzled K-major layout the MMA descriptors expect (what a
//     TMA box of that row would have produced), fences it for the async proxy and arrives on the leader's barrier; the same row
//     also goes to the x_k plane in global memory for the later layers of the block.
// No CTA ever waits for another CTA pair: everything B needs beyond its own strip / band is RECOMPUTED locally --
//   * columns: strips are 126 output pixels wide (stride 126, the MMA tile stays 128): A's 128 columns are B's 126 columns
//     plus their one-pixel halo; B's outermost two columns are discarded. Cost: 128 / 126 = 1.6 % more MMA work;
//   * rows: A computes its band plus ONE row above and below (nrow + 2 rows; B's vertical halo). Cost: 2 / band of layer A.
//   A's halo rows / columns are used through the hand-off only and never stored, so every x_k element in global memory has ONE
//   writer (K3's summation order depends on the ring position; two writers would not be bit-identical).
// TMEM: two rings with K3's N = 64 geometry (period 6 + 2 mirror blocks of 32 columns = 256 columns each). Shared memory:
// both layers' weight halves resident (46 / 83 KB per CTA for conv1+2 / conv3+4), 2 hand-off slots, 6..8 TMA slots.
// Issue order per step s: the nchA boxes of A's row pair s, then the nchA + 1 boxes of B's row pair s - 2 (hand-off box last),
// walked by the same two alternating issuer warps as K3. DRAM traffic of a dense block: 26 -> 18 plane transfers (-31 %).
#pragma once
#include "conv3x3_pair_sm100.cuh"

namespace vr {

struct Pair2 {
    using T = PairTraits<32>;
    static constexpr int N = 32;
    static constexpr uint32_t P = 6;         // logical ring period per layer
    static constexpr int kPhys = 8;          // physical blocks per layer (two mirrors)
    static constexpr int kRingCols = 256;    // TMEM columns per layer
    static constexpr int kHand = 2;          // hand-off slots (row pairs of x_k in flight between A's epilogue and B's MMAs)
    static constexpr int kLag = 2;           // B runs this many row pairs behind A
    static constexpr int kStrip = 126;       // output pixels per strip
    static constexpr int kEpi = 16;          // epilogue warps: four groups = (layer A | B) x (even | odd logical row)
    static constexpr int kThreads = (kEpi + 1 + kMmaWarps) * 32;
    static constexpr int kBudget = 227 * 1024 - 1024 - 4096;   // dynamic shared memory: weights + hand-off + TMA slots
    static constexpr int kMinSlots = 4;
};

// The static box sequence of one work item -- step s = A's row pair s (chunks 0..nchA-1), then B's row pair s - lag (chunks
// 0..nchA-1 from TMA, chunk nchA from the hand-off slot) -- is walked incrementally by the producer, the waiter and the issuer
// (three copies of the same few counters; a shared iterator struct cost registers in the issuing warps and was removed).
// sub-items as in K3, strips of 126 pixels
__device__ __forceinline__ PairSub pair2_sub(const ConvArgs& a, int u) { return pair_sub(a, u); }

// Writes this lane's pixel (32 fp16 channels = 64 B) of an x_k row into a hand-off slot: pixel index `px` of box row `r`, in the
// SWIZZLE_64B layout of a TMA box (16 B unit u of the 64 B pixel row at unit u ^ ((offset >> 7) & 3); slots are 512 B aligned).
__device__ __forceinline__ void hand_store(uint32_t slot_s, int r, int px, const uint4& h0, const uint4& h1, const uint4& h2, const uint4& h3) {
    const uint32_t o = static_cast<uint32_t>(r * 130 + px) * 64u;
    const uint32_t sw = (o >> 7) & 3u;
    ptx::sts128(slot_s + o + ((0u ^ sw) << 4), h0);
    ptx::sts128(slot_s + o + ((1u ^ sw) << 4), h1);
    ptx::sts128(slot_s + o + ((2u ^ sw) << 4), h2);
    ptx::sts128(slot_s + o + ((3u ^ sw) << 4), h3);
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Pair2::kThreads, 1)
conv3x3_pair2_kernel(const __grid_constant__ CUtensorMap tmap, const ConvArgs a) {
    using T = PairTraits<32>;
    constexpr int N = 32;
    constexpr uint32_t P = Pair2::P;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t s_bars[2 * kPairMaxSlots + 4 * P + 2 * Pair2::kHand + 1];
    __shared__ uint32_t s_tmem_slot;
    __shared__ __align__(16) float s_bias[2][N];
    uint64_t* full = s_bars;                    // leader: both CTAs' boxes of a TMA slot have landed
    uint64_t* empty = full + kPairMaxSlots;     // both: the MMAs reading the slot have retired (multicast commit)
    uint64_t* tfull = empty + kPairMaxSlots;    // both: [layer][P] logical ring position complete (multicast commit)
    uint64_t* tempty = tfull + 2 * P;           // leader: [layer][P] ring position drained and re-initialised in BOTH CTAs
    uint64_t* hfull = tempty + 2 * P;           // leader: both x_k rows of a hand-off slot written in BOTH CTAs
    uint64_t* hempty = hfull + Pair2::kHand;    // both: the MMAs reading the hand-off slot have retired (multicast commit)
    uint64_t* wfull = hempty + Pair2::kHand;
    const int nslots = a.nstages;
    const int nchA = a.nchunks, nchB = a.nchunks + 1;
    const int lag = a.lag < 1 ? Pair2::kLag : a.lag;  // B runs `lag` row pairs behind A (>= 2: the hand-off needs A's pair b + 1 drained)
    uint8_t* wB = smem + nchA * T::kBHalf;
    uint8_t* hand0 = wB + nchB * T::kBHalf;
    uint8_t* slot0 = hand0 + Pair2::kHand * T::kASlot;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = ptx::cluster_ctarank();
    const long long t_start = clock64();
    // VR profiling hook (dbg_cycles set by the test hook only): cluster 0's leader CTA accumulates the cycles its warps spend
    // in each kind of wait into dbg_cycles[300..340)
#ifdef VR_K4_PROF
    const bool prof = a.dbg_cycles != nullptr && blockIdx.x == 0;
#else
    constexpr bool prof = false;   // build with -DVR_K4_PROF for the wait-cycle counters (tools/k4_profile.py)
#endif

    if (threadIdx.x == 0) {
        for (int i = 0; i < kPairMaxSlots; ++i) {
            ptx::mbar_init(&full[i], 1);
            ptx::mbar_init(&empty[i], 1);
        }
        for (uint32_t i = 0; i < 2 * P; ++i) {
            ptx::mbar_init(&tfull[i], 1);           // the one issuing thread commits (multicast) the row's MMAs
            ptx::mbar_init(&tempty[i], 8);          // four lane-quarter warps in each of the two CTAs
        }
        for (int i = 0; i < Pair2::kHand; ++i) {
            ptx::mbar_init(&hfull[i], 16);          // two rows x four lane-quarter warps x two CTAs
            ptx::mbar_init(&hempty[i], 1);
        }
        ptx::mbar_init(wfull, 1);
        ptx::fence_mbar_init();
    }
    if (warp == Pair2::kEpi) {
        if (lane == 0) ptx::prefetch_tmap(&tmap);
        __syncwarp();
        ptx::tmem_alloc_pair<512>(&s_tmem_slot);
    }
    for (int i = threadIdx.x; i < 2 * N; i += blockDim.x) {
        const float* b = i < N ? a.bias : a.bias2;
        s_bias[i / N][i % N] = b ? b[i % N] : 0.f;
    }
    // hand-off slots: the halo pixels 0 and 129 of each row are never written by the epilogue; they only reach discarded
    // outputs, but must not hold NaN patterns from a previous kernel
    for (int i = threadIdx.x; i < Pair2::kHand * T::kASlot / 16; i += blockDim.x)
        reinterpret_cast<uint4*>(hand0)[i] = make_uint4(0u, 0u, 0u, 0u);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = s_tmem_slot;
    if (warp == Pair2::kEpi && lane == 0) {
        // this CTA's halves of both layers' weight rows; weights are never written by a kernel: fetch before the dependency wait
        const __half* wpa = a.wpack + (static_cast<size_t>(rank) * nchA) * (T::kBHalf / 2);
        const __half* wpb = a.wpack2 + (static_cast<size_t>(rank) * nchB) * (T::kBHalf / 2);
        ptx::mbar_expect_tx(wfull, (nchA + nchB) * T::kBHalf);
        for (int c = 0; c < nchA; ++c) ptx::bulk_load(smem + c * T::kBHalf, wpa + static_cast<size_t>(c) * (T::kBHalf / 2), T::kBHalf, wfull);
        for (int c = 0; c < nchB; ++c) ptx::bulk_load(wB + c * T::kBHalf, wpb + static_cast<size_t>(c) * (T::kBHalf / 2), T::kBHalf, wfull);
    }
    if (warp < Pair2::kEpi) {
        // ring blocks start as the bias row of their layer, mirror blocks as zero; every MMA accumulates
        const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
        const int L = warp >> 3;   // warps 0..7 initialise layer A's ring, 8..15 layer B's
        float bz[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) bz[j] = s_bias[L][j];
        for (uint32_t blk = (warp >> 2) & 1; blk < static_cast<uint32_t>(Pair2::kPhys); blk += 2) {
            const uint32_t t = tmem_base + lane_base + L * Pair2::kRingCols + blk * N;
            if (blk < P) ptx::tmem_st32(t, bz);
            else ptx::tmem_st32_zero(t);
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

    if (warp == Pair2::kEpi) {
        // ===================== TMA producer: this CTA's strip; bytes counted on the leader's barrier =====================
        if (lane == 0) {
            const uint32_t lead_full = ptx::map_to_rank(&full[0], 0);
            int s = 0;
            uint32_t ph = 0;
            long long pc[1] = {0};
            // plain nested loops with incremental state (the generic box iterator cost this single thread ~600 cycles per box --
            // as much as the MMAs of a box take -- so the loads of a step were issued barely faster than they were consumed)
            const bool planar_src = a.in_cstride == 32;
            auto load_box = [&](const int xc, const int yrow, const int c) {
                const long long tw0 = prof ? clock64() : 0;
                ptx::mbar_wait(&empty[s], ph ^ 1);
                if (prof) pc[0] += clock64() - tw0;
                if (a.flags & FLAG_SKIP_TMA) {  // ablation: the slot protocol without the loads
                    if (rank == 0) ptx::mbar_arrive(&full[s]);
                } else {
                    if (rank == 0) ptx::mbar_expect_tx(&full[s], 2 * T::kCopyBytes);
                    const int ch0 = a.cin_off + c * T::KC;
                    ptx::tma_load_4d_pair(slot0 + s * T::kASlot, &tmap, lead_full + s * 8, planar_src ? 0 : ch0, xc, yrow,
                                          planar_src ? ch0 >> 5 : 0);
                }
                if (++s == nslots) { s = 0; ph ^= 1; }
            };
            for (int item = cluster_id; item < num_items; item += nclusters) {
                const PairSub me = pair2_sub(a, 2 * item + static_cast<int>(rank));
                const int xc = me.sx * Pair2::kStrip - 2;  // box column 0: one pixel left of the strip's first (halo) pixel
                const int nB2 = ((pair_rows(a, item) + 3) & ~1) >> 1, nA2 = nB2 + 1;
                const int S = nB2 + lag;
                // A's input rows start one row above B's: A computes the band plus a halo row on either side
                int ya = me.y0 - 2, yb = me.y0 - 1;
                // Experiment, off by default (a.prefetch = 0): layer A's boxes come from HBM and the slot ring only reaches 0.75 .. 2
                // steps ahead, so their rows can be pulled into the L2 `pf` steps early with prefetch-only TMA requests (no shared
                // memory needed). Measured slower -- the TMA unit's request rate is the scarcer resource (see vr_common.h).
                const int pf = a.prefetch;
                if (pf > 0 && !(a.flags & FLAG_SKIP_TMA))
                    for (int st = 0; st < pf && st < nA2; ++st)
                        for (int c = 0; c < nchA; ++c) {
                            const int ch0 = a.cin_off + c * T::KC;
                            ptx::tma_prefetch_4d(&tmap, planar_src ? 0 : ch0, xc, ya + 2 * st, planar_src ? ch0 >> 5 : 0);
                        }
                for (int st = 0; st < S; ++st) {
                    if (st < nA2) {
                        if (pf > 0 && st + pf < nA2 && !(a.flags & FLAG_SKIP_TMA))
                            for (int c = 0; c < nchA; ++c) {
                                const int ch0 = a.cin_off + c * T::KC;
                                ptx::tma_prefetch_4d(&tmap, planar_src ? 0 : ch0, xc, ya + 2 * pf, planar_src ? ch0 >> 5 : 0);
                            }
                        for (int c = 0; c < nchA; ++c) load_box(xc, ya, c);
                        ya += 2;
                    }
                    if (st >= lag && st - lag < nB2) {
                        for (int c = 0; c < nchA; ++c) load_box(xc, yb, c);
                        yb += 2;
                    }
                }
            }
            if (prof) {
                a.dbg_cycles[300] = pc[0];                 // producer: cycles waiting for a free TMA slot
                a.dbg_cycles[301] = clock64() - t_start;
            }
        }
    } else if (warp > Pair2::kEpi) {
        // ===================== MMA issue: leader only; one ISSUER warp and one WAITER warp =====================
        // K3 alternates two issuing warps, each waiting for its own unit's barriers while the other issues. Measured on K4 with
        // the wait-cycle counters: that overlap does not happen -- rows receive MMAs from both warps, so both must tcgen05.commit
        // every accumulator-ready barrier, and a commit queues behind the other warp's MMAs in the (shallow) tensor-pipe FIFO:
        // the non-owner cannot run ahead to its next wait pass, and every unit pays wait pass + issue in sequence (the bare
        // barrier skeleton, no MMAs / TMA / TMEM traffic, ran at 59 of the kernel's 88 us). Here ONE elected thread issues every
        // MMA and every commit (accumulator-ready barriers need one arrival) and never touches an mbarrier; the waiter warp does
        // all the waiting one unit ahead and hands units over through two alternating pairs of named barriers.
        // Both warps walk the box sequence ONCE, with incremental state only (slot / hand-off / ring positions advance by adds
        // and wraps: no divisions, no iterator object): with ~500 cycles of tensor work per box, every ~100 scalar instructions
        // per box in either warp cost as much as the MMAs themselves. Units of `a.unit` boxes (3) are handed over; a unit per
        // (layer, row pair) with a closed-form barrier list was measured too: same on conv1+2, 10 % slower on conv3+4, whose six
        // TMA slots cannot hold the next 4-box unit while the current one is being issued.
        if (rank == 0) {
            const bool issuer = warp == Pair2::kEpi + 1;
            const int unit = a.unit < 1 ? 1 : a.unit;
            const bool skip_mma = (a.flags & FLAG_SKIP_MMA) != 0;
            const uint32_t slot0_lo = ptx::smem_u32(slot0) >> 4, hand0_lo = ptx::smem_u32(hand0) >> 4;
            const uint32_t wA_lo = ptx::smem_u32(smem) >> 4, wB_lo = ptx::smem_u32(wB) >> 4;
            constexpr uint32_t kSlotLo = T::kASlot >> 4, kBHalfLo = T::kBHalf >> 4;
            const uint32_t full_s = ptx::smem_u32(full), hfull_s = ptx::smem_u32(hfull), tempty_s = ptx::smem_u32(tempty);
            int ss = 0;              // TMA slot ring: index and phase
            uint32_t sph = 0;
            uint32_t hq = 0, hph = 0;  // hand-off slot ring: index and phase
            int gunit = 0, par = 0, ubox = 0;
            uint32_t pos[2] = {0, 0}, rev[2] = {0, 0};  // per layer: ring position / revolution count of the current pair's first row
            int k = 0;                 // waiter: barrier list entries of the current unit
            uint32_t my_bar = 0u, my_par = 0u;  // shared-memory address of this lane's barrier (0 = none)
            long long ic[4] = {0, 0, 0, 0};
            const long long t_loop = prof ? clock64() : 0;
            long long tp0 = t_loop;
            for (int item = cluster_id; item < num_items; item += nclusters) {
                const int nin2B = (pair_rows(a, item) + 3) & ~1;
                const int nB2 = nin2B >> 1, nA2 = nB2 + 1;
                const int S = nB2 + lag;
                int left = nA2 * nchA + nB2 * (nchA + 1);  // boxes of the item still to come (units do not span items)
                // one box: layer L, row pair pr, chunk c; hand = B's x_k chunk from the hand-off slot; last = last chunk of the pair
                auto box = [&](const int L, const int pr, const int c, const bool hand, const bool last, const bool last_pair) {
                    if (!issuer) {
                        if (ubox == 0) {
                            // the issuer has finished unit gunit - 2: its "ready" barrier may be signalled again
                            if (gunit >= 2) asm volatile("bar.sync %0, 64;" ::"r"(3 + par) : "memory");
                            if (prof) { const long long t = clock64(); ic[0] += t - tp0; tp0 = t; }
                            k = 0;
                            my_bar = 0u;
                        }
                        // every barrier the unit needs, ONE PER LANE (a wait on an mbarrier completed through the async proxy costs
                        // the waiting thread a few hundred cycles even when it completed long ago; a unit has up to 15 of them);
                        // selects, not branches
                        {
                            const bool me = k == lane;
                            const uint32_t bar_s = hand ? hfull_s + hq * 8u : full_s + static_cast<uint32_t>(ss) * 8u;
                            my_bar = me ? bar_s : my_bar;
                            my_par = me ? (hand ? hph : sph) : my_par;
                            ++k;
                        }
                        if (c == 0) {
                            // logical rows first touched by this row pair: +2, +3 (and +0, +1 at the top of an item)
                            const uint32_t te_s = tempty_s + static_cast<uint32_t>(L) * P * 8u;
#pragma unroll
                            for (uint32_t d = 0u; d < 4u; ++d) {
                                if (d < 2u && pr != 0) continue;
                                uint32_t p = pos[L] + d, r = rev[L];
                                if (p >= P) { p -= P; ++r; }
                                const bool me = k == lane;
                                my_bar = me ? te_s + p * 8u : my_bar;
                                my_par = me ? ((r & 1u) ^ 1u) : my_par;
                                ++k;
                            }
                        }
                    } else {
                        if (ubox == 0) {
                            asm volatile("bar.sync %0, 64;" ::"r"(1 + par) : "memory");
                            ptx::tc_fence_after();
                            if (prof) { const long long t = clock64(); ic[0] += t - tp0; tp0 = t; }
                        }
                        if (ptx::elect_one()) {
                            const uint32_t s0 = pos[L], s1 = s0 + 1 >= P ? s0 + 1 - P : s0 + 1;
                            const uint32_t tb = tmem_base + L * Pair2::kRingCols;
                            if (!skip_mma)
                                pair_issue_box<N>(tb + s0 * N, tb + s1 * N, hand ? hand0_lo + hq * kSlotLo : slot0_lo + ss * kSlotLo,
                                                  (L == 0 ? wA_lo : wB_lo) + c * kBHalfLo);
                            ptx::umma_commit_pair(hand ? &hempty[hq] : &empty[ss]);
                            if (last) {
                                uint64_t* tf = tfull + L * P;
                                ptx::umma_commit_pair(&tf[s0]);
                                ptx::umma_commit_pair(&tf[s1]);
                                if (last_pair) {
                                    const uint32_t s2 = s0 + 2 >= P ? s0 + 2 - P : s0 + 2, s3 = s0 + 3 >= P ? s0 + 3 - P : s0 + 3;
                                    ptx::umma_commit_pair(&tf[s2]);
                                    ptx::umma_commit_pair(&tf[s3]);
                                }
                            }
                        }
                    }
                    // advance the state both warps share by construction
                    if (hand) {
                        if (++hq == Pair2::kHand) { hq = 0; hph ^= 1u; }
                    } else if (++ss == nslots) {
                        ss = 0;
                        sph ^= 1u;
                    }
                    if (last) {
                        // the next pair of this layer starts two logical rows further; after an item's last pair the 2 trailing
                        // phantom rows are skipped as well (an item spans nin2 + 2 logical rows)
                        pos[L] += last_pair ? 4u : 2u;
                        if (pos[L] >= P) { pos[L] -= P; ++rev[L]; }
                    }
                    --left;
                    if (++ubox == unit || left == 0) {
                        if (!issuer) {
                            if (prof) { const long long t = clock64(); ic[2] += t - tp0; tp0 = t; }
                            if (my_bar) ptx::mbar_wait_s(my_bar, my_par);
                            __syncwarp();
                            if (prof) { const long long t = clock64(); ic[1] += t - tp0; tp0 = t; }
                            asm volatile("bar.arrive %0, 64;" ::"r"(1 + par) : "memory");
                        } else {
                            __syncwarp();
                            if (prof) { const long long t = clock64(); ic[1] += t - tp0; tp0 = t; }
                            asm volatile("bar.arrive %0, 64;" ::"r"(3 + par) : "memory");
                        }
                        ubox = 0;
                        ++gunit;
                        par ^= 1;
                    }
                };
                for (int st = 0; st < S; ++st) {
                    if (st < nA2) {
                        for (int c = 0; c < nchA; ++c) box(0, st, c, false, c == nchA - 1, st == nA2 - 1);
                    }
                    const int bq = st - lag;
                    if (bq >= 0 && bq < nB2) {
                        for (int c = 0; c < nchA; ++c) box(1, bq, c, false, false, false);
                        box(1, bq, nchA, true, true, bq == nB2 - 1);
                    }
                }
            }
            // the issuer's "done" arrivals of the last two units have no matching sync yet: consume them so that no named
            // barrier is left half-arrived
            if (!issuer)
                for (int u = gunit >= 2 ? gunit - 2 : 0; u < gunit; ++u) asm volatile("bar.sync %0, 64;" ::"r"(3 + (u & 1)) : "memory");
            if (prof && lane == 0) {
                const int o = 310 + (issuer ? 0 : 8);
                for (int i = 0; i < 3; ++i) a.dbg_cycles[o + i] = ic[i];
                a.dbg_cycles[o + 3] = clock64() - t_loop;   // main loop only (no prologue)
                a.dbg_cycles[o + 4] = gunit;
                a.dbg_cycles[o + 5] = t_loop - t_start;     // prologue incl. the wait for the previous launch
            }
        }
    } else {
        // ===================== epilogue warps (both CTAs): warp % 4 = TMEM lane quarter; warp / 4 = group: groups 0, 1 drain layer
        // A's even / odd logical rows, groups 2, 3 layer B's -- four rows in flight per CTA. The 32-channel layers are bound by the
        // epilogue's serial latency chain per row (tfull -> tcgen05.ld -> tcgen05.st re-init -> hand-off / stores: ~1.2 us per row
        // and warp, measured with the wait-cycle counters), not by its instruction count, so more rows in flight is what helps.
        const int quarter = warp & 3;
        const int myL = warp >> 3;
        const uint32_t rgrp = (warp >> 2) & 1;
        const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
        const uint32_t lead_tempty = ptx::map_to_rank(&tempty[0], 0);
        const uint32_t lead_hfull = ptx::map_to_rank(&hfull[0], 0);
        const uint32_t hand_s = ptx::smem_u32(hand0);
        const int local = quarter * 32 + lane;                 // pixel of the 128-pixel MMA tile
        const bool owned_col = local >= 1 && local <= Pair2::kStrip;
        const float slope = a.slope;
        uint32_t g0A = 0, g0B = 0, hbase = 0;  // hbase: hand-off row pairs filled before the current item
        long long ec[3] = {0, 0, 0};  // tfull waits, hempty waits, rows
        for (int item = cluster_id; item < num_items; item += nclusters) {
            const PairSub me = pair2_sub(a, 2 * item + static_cast<int>(rank));
            const int y0 = me.y0, nrow = me.nrow;
            const int nin2B = (pair_rows(a, item) + 3) & ~1, nin2A = nin2B + 2;
            const int x = me.sx * Pair2::kStrip - 1 + local;
            const bool x_in = x >= 0 && x < a.W;
            bool xgap = false;
            for (int j = 0; j < a.ngx; ++j) xgap |= (x_in && (x >> a.gshift) == a.gx[j]);
            const int nA2 = nin2A >> 1, nB2 = nin2B >> 1;
            const int S = nB2 + lag;
            // one logical row of layer L: drain, hand the ring position back, then (real rows) activation -> fp16 -> hand-off / global
            auto do_row = [&](int L, int l) {
                const uint32_t gl = (L == 0 ? g0A : g0B) + l;
                const uint32_t m = gl % P;
                const long long te0 = prof ? clock64() : 0;
                {
                    // the two barriers a hand-off row needs (accumulators complete, hand-off slot free) are waited for by two
                    // different lanes at once: each such wait costs a few hundred cycles however long ago it completed
                    const int jb = l - 2;
                    const bool ho = L == 0 && jb >= 0 && jb < nin2B && !(a.flags & FLAG_SKIP_A);
                    const uint32_t f = hbase + static_cast<uint32_t>(jb >> 1);
                    if (ho && lane == 1) ptx::mbar_wait(&hempty[f % Pair2::kHand], ((f / Pair2::kHand) & 1u) ^ 1u);
                    else if (lane == 0 || !ho) ptx::mbar_wait(&tfull[L * P + m], (gl / P) & 1u);
                    __syncwarp();
                }
                if (prof) { ec[0] += clock64() - te0; ec[2] += 1; }
                ptx::tc_fence_after();
                const uint32_t t_main = tmem_base + lane_base + L * Pair2::kRingCols + m * N;
                const uint32_t t_mir = m < 2 ? tmem_base + lane_base + L * Pair2::kRingCols + (P + m) * N : 0xffffffffu;
                // image row of this logical row: A's band starts one row above B's
                const int y = L == 0 ? y0 - 3 + l : y0 - 2 + l;
                const int jB = l - 2;  // layer A: the B input row this x_k row is
                const bool handoff = L == 0 && jB >= 0 && jB < nin2B;
                const bool real = l >= 2 && l < (L == 0 ? nrow + 4 : nrow + 2);
                if (a.flags & FLAG_SKIP_A) {  // ablation: barrier protocol only (no TMEM loads / re-initialisation, no data)
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive_cluster(lead_tempty + (L * P + m) * 8);
                    if (handoff) {
                        const uint32_t f = hbase + static_cast<uint32_t>(jB >> 1);
                        const uint32_t q = f % Pair2::kHand;
                        ptx::mbar_wait(&hempty[q], ((f / Pair2::kHand) & 1u) ^ 1u);
                        __syncwarp();
                        if (lane == 0) ptx::mbar_arrive_cluster(lead_hfull + q * 8);
                    }
                    return;
                }
                if (!handoff && !real) {  // phantom row: nothing to read, only hand the ring position back
                    pair_release<N>(t_main, t_mir, s_bias[L], lane, lead_tempty + (L * P + m) * 8);
                    return;
                }
                float v[32];
                if (t_mir != 0xffffffffu) {
                    uint32_t r0[32], r1[32];
                    ptx::tmem_ld32_issue(t_main, r0);
                    ptx::tmem_ld32_issue(t_mir, r1);
                    ptx::tmem_ld32_wait(r0);
                    ptx::tmem_ld32_wait(r1);
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r0[j]) + __uint_as_float(r1[j]);
                } else {
                    ptx::tmem_ld32(t_main, v);
                }
                pair_release<N>(t_main, t_mir, s_bias[L], lane, lead_tempty + (L * P + m) * 8);
                bool zero = !x_in || xgap || y < 0 || y >= a.H || !real;
                if (!zero)
                    for (int j = 0; j < a.ngy; ++j) zero |= ((y >> a.gshift) == a.gy[j]);
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = zero ? 0.f : fmaxf(v[j], v[j] * slope);
                const uint4 h0 = pack8(v), h1 = pack8(v + 8), h2 = pack8(v + 16), h3 = pack8(v + 24);
                if (handoff && (a.flags & FLAG_SKIP_EPI)) {  // ablation: hand-off protocol without the data
                    const uint32_t f = hbase + static_cast<uint32_t>(jB >> 1);
                    const uint32_t q = f % Pair2::kHand;
                    ptx::mbar_wait(&hempty[q], ((f / Pair2::kHand) & 1u) ^ 1u);
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive_cluster(lead_hfull + q * 8);
                } else if (handoff) {
                    const uint32_t f = hbase + static_cast<uint32_t>(jB >> 1);
                    const uint32_t q = f % Pair2::kHand;
                    // (hempty[q] was waited for by lane 1 together with the accumulator-ready barrier)
                    hand_store(hand_s + q * T::kASlot, jB & 1, 1 + local, h0, h1, h2, h3);
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive_cluster(lead_hfull + q * 8);
                }
                // global store: only rows of the band itself (A's halo rows belong to the neighbouring bands) and owned columns
                const bool in_band = L == 0 ? (l >= 3 && l < nrow + 3) : real;
                if (in_band && owned_col && x < a.W && !(a.flags & FLAG_SKIP_B)) {
                    const size_t p = static_cast<size_t>(y) * a.W + x;
                    __half* o = a.out + chan_off(p, a.out_cstride, a.out_pstride, L == 0 ? a.out_coff : a.out_coff2);
                    ptx::stg256(o, h0, h1);
                    ptx::stg256(o + 16, h2, h3);
                }
            };
            (void)S;
            if (myL == 0) {
                for (int s = 0; s < nA2; ++s) do_row(0, 2 * s + static_cast<int>(rgrp));
                do_row(0, nin2A + static_cast<int>(rgrp));
            } else {
                for (int b = 0; b < nB2; ++b) do_row(1, 2 * b + static_cast<int>(rgrp));
                do_row(1, nin2B + static_cast<int>(rgrp));
            }
            g0A += nin2A + 2;
            g0B += nin2B + 2;
            hbase += static_cast<uint32_t>(nB2);
        }
        if (prof && lane == 0 && (warp == 0 || warp == 8)) {
            for (int i = 0; i < 3; ++i) a.dbg_cycles[330 + (warp >> 3) * 4 + i] = ec[i];
            a.dbg_cycles[330 + (warp >> 3) * 4 + 3] = clock64() - t_start;
        }
    }

    ptx::tc_fence_before();
    ptx::cluster_sync();  // the peer may still be read (operands) or signalled (barriers) until both are done
    if (warp == Pair2::kEpi) {
        __syncwarp();
        ptx::tmem_dealloc_pair<512>(tmem_base);
    }
    if (a.dbg_cycles && threadIdx.x == 0) a.dbg_cycles[blockIdx.x] = clock64() - t_start;
}

}  // namespace vr
