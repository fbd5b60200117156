// Micro-probe for K2's issue path: a producer streams 8320 B boxes into a ring of smem slots (bulk copies -> mbarrier), one
// or two warps issue 6 MMAs (N = 96) per box and release the slot with tcgen05.commit. No epilogue. Which synchronisation
// structure keeps the tensor pipe at 6 x 56 = 336 cycles per box?
//   mode 0  one issuer: wait full[s] -> 6 MMAs -> commit empty[s]
//   mode 1  two issuers alternating boxes, named-barrier hand-over (K2 as first written)
//   mode 2  two INDEPENDENT streams (own slots / TMEM / producer), no hand-over
//   mode 3  one issuer, waits for box i+1 before issuing box i
//   mode 4  one issuer + watcher warp: the watcher waits on the mbarriers and publishes a sequence number in smem
//   mode 5  like 0 with mbarrier.test_wait polling
//   mode 6  one issuer, G boxes per hand-over unit: test_wait on all G full barriers first (latencies overlap), then 6G MMAs
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/issue_probe tools/issue_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../video_restore_b200/csrc/conv3x3_roll_sm100.cuh"
using namespace vr::ptx;

constexpr int kSlot = 8704, kCopy = 8320, kNSlots = 12, kBBytes = 18432;

__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}

__device__ __forceinline__ void issue_box(uint32_t tmem, uint32_t a0, uint32_t b0, int row) {
#pragma unroll
    for (int dx = 0; dx < 3; ++dx)
#pragma unroll
        for (int k = 0; k < 2; ++k)
            umma_f16<kCollNone>(tmem + (row % 6) * 32, a0 + ((dx * 64 + k * 32) >> 4), kDescHiSw64,
                                b0 + ((dx * 3 * 32 * 64 + k * 32) >> 4), kDescHiSw64, make_idesc_f16(128, 96), 1u);
}

__global__ void __launch_bounds__(192, 1) probe(const uint8_t* src, long long* out, int mode_in, int nbox, int nch, int G) {
    const bool fence = mode_in < 10;  // mode + 10: no tcgen05.fence::after_thread_sync after the full-barrier wait
    const int mode = mode_in % 10;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t full[kNSlots], empty[kNSlots], done[2];
    __shared__ uint32_t slot;
    __shared__ volatile int flag[kNSlots];
    uint8_t* bsm = smem + 8 * 16896;
    for (int i = threadIdx.x; i < (8 * 16896 + 2 * kBBytes) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x2c002c00u;
    if (threadIdx.x == 0) {
        for (int i = 0; i < kNSlots; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); flag[i] = 0; }
        mbar_init(&done[0], 1); mbar_init(&done[1], 1);
        fence_mbar_init();
    }
    if (threadIdx.x < 32) tmem_alloc<512>(&slot);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nstreams = mode == 2 ? 2 : 1;
    const int per = kNSlots / nstreams;  // slots per stream
    const uint32_t b0 = smem_u32(bsm) >> 4;
    const long long t0 = clock64();

    if (mode_in >= 70 && mode_in <= 79) {
        // fully static issue patterns with a runtime ring base (no wrap): ROWS input rows per box, N = 96 windows 32 columns apart.
        // 70: 2 rows row-major; 71: 2 rows interleaved; 72: 3 rows interleaved ((dx,k) outer); 73: 4 rows interleaved;
        // 74: 3 rows row-major; 75: 4 rows row-major; 76: 1 row (6 MMAs); 77: 2 rows, order r0 r0 r1 r1 ...
        if (warp == 2) {
            const int rows = (mode_in == 70 || mode_in == 71 || mode_in == 77) ? 2 : (mode_in == 72 || mode_in == 74) ? 3 : mode_in == 76 ? 1 : 4;
            for (int i = 0; i < nbox; ++i) {
                const uint32_t a0 = smem_u32(smem + (i % 4) * 33792) >> 4;
                const uint32_t base = tmem + (uint32_t)(((i / nch) * rows) % 9) * 32u;
                if (elect_one()) {
#define MMA_(row, t) umma_f16<kCollNone>(base + (row) * 32, a0 + ((((t) / 2) * 64 + ((t) % 2) * 32) >> 4) + (row) * 520, kDescHiSw64, \
                                         b0 + ((((t) / 2) * 3 * 32 * 64 + ((t) % 2) * 32) >> 4), kDescHiSw64, make_idesc_f16(128, 96), 1u)
                    if (mode_in == 70) {
#pragma unroll
                        for (int m = 0; m < 12; ++m) MMA_(m / 6, m % 6);
                    } else if (mode_in == 71) {
#pragma unroll
                        for (int m = 0; m < 12; ++m) MMA_(m % 2, m / 2);
                    } else if (mode_in == 72) {
#pragma unroll
                        for (int m = 0; m < 18; ++m) MMA_(m % 3, m / 3);
                    } else if (mode_in == 73) {
#pragma unroll
                        for (int m = 0; m < 24; ++m) MMA_(m % 4, m / 4);
                    } else if (mode_in == 74) {
#pragma unroll
                        for (int m = 0; m < 18; ++m) MMA_(m / 6, m % 6);
                    } else if (mode_in == 75) {
#pragma unroll
                        for (int m = 0; m < 24; ++m) MMA_(m / 6, m % 6);
                    } else if (mode_in == 76) {
#pragma unroll
                        for (int m = 0; m < 6; ++m) MMA_(0, m);
                    } else {
#pragma unroll
                        for (int m = 0; m < 12; ++m) MMA_((m / 2) % 2, (m / 4) * 2 + m % 2);
                    }
                    umma_commit(&empty[i % 4]);
                }
                __syncwarp();
            }
            if (elect_one()) umma_commit(&done[0]);
            __syncwarp();
            mbar_wait(&done[0], 0);
            if (lane == 0) out[0] = (clock64() - t0) * 2 / rows;  // normalised so the printed figure is per 6 MMAs
        }
    } else if (mode_in == 60 || mode_in == 61) {
        // K2's own issue code (roll_row / roll_issue_box) on static operands, one warp, per-box elect. 61: two warps alternating
        const int w = warp - 2;
        const bool dual = mode_in == 61;
        if (warp == 2 || (warp == 3 && dual)) {
            for (int i = 0; i < nbox; ++i) {
                const bool mine = !dual || (i & 1) == w;
                const int j0 = 2 + 2 * (i / nch), c = i % nch;
                vr::RollRow r0 = vr::roll_row<16>(j0, 100000, 0u), r1 = vr::roll_row<16>(j0 + 1, 100000, 0u);
                if (mine && dual && i > 0) asm volatile("bar.sync %0, 64;" ::"r"(1 + w) : "memory");
                if (mine && elect_one()) {
                    vr::roll_issue_box<32>(r0, r1, true, tmem, smem_u32(smem + (i % 8) * 16896) >> 4, b0, c == 0);
                    umma_commit(&empty[i % 8]);
                }
                __syncwarp();
                if (mine && dual) asm volatile("bar.arrive %0, 64;" ::"r"(2 - w) : "memory");
            }
            if (dual && nbox > 0 && (nbox & 1) == w) asm volatile("bar.sync %0, 64;" ::"r"(1 + w) : "memory");
            if (elect_one()) umma_commit(&done[w]);
            __syncwarp();
            mbar_wait(&done[w], 0);
            if (lane == 0) out[w] = clock64() - t0;
        }
    } else if (mode_in >= 49 && mode_in <= 52) {
        // static, ONE issuer warp, per-box elect: 3 (49, 51) or 4 (50, 52) input rows per box, (dx,k) outer / row inner so that
        // consecutive MMAs walk windows 32 columns apart (K1's pattern); ring advances rows*32 columns every nch boxes.
        // 51 / 52: row-major order for comparison
        const int rows = (mode_in == 49 || mode_in == 51) ? 3 : 4;
        const bool rowmajor = mode_in >= 51;
        if (warp == 2) {
            for (int i = 0; i < nbox; ++i) {
                const uint32_t a0 = smem_u32(smem + (i % 4) * 33792) >> 4;
                const uint32_t win = (uint32_t)((i / nch) * rows * 32) & 511u;
                if (elect_one()) {
                    for (int m = 0; m < 6 * rows; ++m) {
                        const int row = rowmajor ? m / 6 : m % rows, t = rowmajor ? m % 6 : m / rows;
                        const uint32_t d = (win + row * 32) & 511u;
                        const uint32_t n = d + 96 <= 512 ? 96u : 512u - d;
                        umma_f16<kCollNone>(tmem + d, a0 + (((t / 2) * 64 + (t % 2) * 32) >> 4) + row * 520, kDescHiSw64,
                                            b0 + (((t / 2) * 3 * 32 * 64 + (t % 2) * 32) >> 4), kDescHiSw64, make_idesc_f16(128, n), 1u);
                    }
                    umma_commit(&empty[i % 4]);
                }
                __syncwarp();
            }
            if (elect_one()) umma_commit(&done[0]);
            __syncwarp();
            mbar_wait(&done[0], 0);
            if (lane == 0) out[0] = (clock64() - t0) * 2 / rows;  // normalised to 12 MMAs so the printed figure is per 6 MMAs
        }
    } else if (mode_in >= 45 && mode_in <= 48) {
        // static replica of K2's issue loop: two issuer warps alternate boxes through named barriers, 12 MMAs per box.
        // 45: row-major (6 MMAs row 0 -> window w, 6 MMAs row 1 -> window w+32); 46: rows interleaved; 47: like 45, ONE warp
        // (no hand-over); 48: like 46, ONE warp. Windows advance 64 columns every nch boxes (ring of 512).
        const int w = warp - 2;
        const bool dual = mode_in <= 46, inter = (mode_in == 46 || mode_in == 48);
        if (warp == 2 || (warp == 3 && dual)) {
            for (int i = 0; i < nbox; ++i) {
                const bool mine = !dual || (i & 1) == w;
                const uint32_t a0 = smem_u32(smem + (i % 8) * 16896) >> 4;
                const uint32_t win = (uint32_t)((i / nch) * 64) & 511u;
                if (mine && dual && i > 0) asm volatile("bar.sync %0, 64;" ::"r"(1 + w) : "memory");
                if (mine && elect_one()) {
#pragma unroll
                    for (int m = 0; m < 12; ++m) {
                        const int row = inter ? (m & 1) : (m / 6), t = inter ? (m >> 1) : (m % 6);
                        const uint32_t d = (win + row * 32) & 511u;
                        const uint32_t n = d + 96 <= 512 ? 96u : 512u - d;  // crude wrap: shorter MMA at the ring end
                        umma_f16<kCollNone>(tmem + d, a0 + (((t / 2) * 64 + (t % 2) * 32) >> 4) + row * 520, kDescHiSw64,
                                            b0 + (((t / 2) * 3 * 32 * 64 + (t % 2) * 32) >> 4), kDescHiSw64, make_idesc_f16(128, n), 1u);
                    }
                    umma_commit(&empty[i % 8]);
                }
                __syncwarp();
                if (mine && dual) asm volatile("bar.arrive %0, 64;" ::"r"(2 - w) : "memory");
            }
            if (dual && nbox > 0 && (nbox & 1) == w) asm volatile("bar.sync %0, 64;" ::"r"(1 + w) : "memory");
            if (elect_one()) umma_commit(&done[w]);
            __syncwarp();
            mbar_wait(&done[w], 0);
            if (lane == 0) out[w] = clock64() - t0;
        }
    } else if (mode_in >= 40) {
        // static, per-box elect, 12 MMAs per box, D pattern by mode: 40 two windows 32 columns apart (overlapping, K2's
        // interleaved rows); 41 two disjoint windows; 42 three windows 32 apart in rotation; 43 same window for all 12;
        // 44: windows 0 / 128 apart (disjoint) but per-box bases advance by 32 like a rolling ring
        if (warp == 2) {
            for (int i = 0; i < nbox; ++i) {
                const uint32_t a0 = smem_u32(smem + (i % kNSlots) * kSlot) >> 4;
                const uint32_t ring = mode_in == 44 ? (uint32_t)(i * 2 % 8) * 32u : 0u;
                if (elect_one()) {
#pragma unroll
                    for (int m = 0; m < 12; ++m) {
                        const uint32_t d = mode_in == 40 ? (m % 2) * 32 : mode_in == 41 ? (m % 2) * 96 : mode_in == 42 ? (m % 3) * 32 : mode_in == 44 ? ((m % 2) * 128 + ring) : 0;
                        umma_f16<kCollNone>(tmem + d, a0 + ((((m / 2) / 2) * 64 + ((m / 2) % 2) * 32) >> 4) + ((m % 2) * 520), kDescHiSw64,
                                            b0 + ((((m / 2) / 2) * 3 * 32 * 64 + ((m / 2) % 2) * 32) >> 4), kDescHiSw64, make_idesc_f16(128, 96), 1u);
                    }
                }
                __syncwarp();
            }
            if (elect_one()) umma_commit(&done[0]);
            __syncwarp();
            mbar_wait(&done[0], 0);
            if (lane == 0) out[0] = clock64() - t0;
        }
    } else if (mode_in >= 30) {
        // full pipeline (producer, full-barrier waits, per-box commits), ONE issuer whose whole loop sits inside one elect.
        // 30: wait box i then issue it; 31: wait for box i+1 before issuing box i. G = rows per box (6 MMAs each).
        if (warp == 0 && lane == 0) {
            int s = 0; uint32_t ph = 0;
            for (int i = 0; i < nbox; ++i) {
                mbar_wait(&empty[s], ph ^ 1);
                mbar_expect_tx(&full[s], kCopy);
                bulk_load(smem + s * kSlot, src + (size_t)((i * 7) % 64) * kCopy, kCopy, &full[s]);
                if (++s == kNSlots) { s = 0; ph ^= 1; }
            }
        } else if (warp == 4 && mode_in == 32) {  // watcher: mbarrier waits -> plain smem sequence numbers
            if (lane == 0) {
                int s = 0; uint32_t ph = 0;
                for (int i = 0; i < nbox; ++i) {
                    mbar_wait(&full[s], ph);
                    flag[s] = i / kNSlots + 1;
                    if (++s == kNSlots) { s = 0; ph ^= 1; }
                }
            }
        } else if (warp == 2) {
            if (elect_one()) {
                int s = 0; uint32_t ph = 0;
                if (mode_in == 31) mbar_wait(&full[0], 0);
                for (int i = 0; i < nbox; ++i) {
                    if (mode_in == 30) mbar_wait(&full[s], ph);
                    else if (mode_in == 32) { while (flag[s] < i / kNSlots + 1) {} }
                    else if (mode_in == 33) {}
                    else if (i + 1 < nbox) {
                        const int ns = s + 1 == kNSlots ? 0 : s + 1;
                        mbar_wait(&full[ns], ns == 0 ? ph ^ 1 : ph);
                    }
                    const uint32_t a0 = smem_u32(smem + s * kSlot) >> 4;
                    for (int g = 0; g < G; ++g) issue_box(tmem, a0, b0, i * G + g);
                    umma_commit(&empty[s]);
                    if (++s == kNSlots) { s = 0; ph ^= 1; }
                }
                umma_commit(&done[0]);
                mbar_wait(&done[0], 0);
                out[0] = clock64() - t0;
            }
            __syncwarp();
        }
    } else if (mode_in >= 20) {
        // static operand modes: no producer, no full-barrier waits. 20: commit per box; 21: commit every 3rd box; 22: no commits;
        // 23: commit per box to ONE barrier; 24: per box commit + satisfied smem-flag poll; 25: two warps, each half the boxes,
        // commit per box, no hand-over
        if (warp == 2 && mode_in >= 26) {
            // 26: whole loop inside ONE elect; 27: per-box elect, D alternates between two windows per MMA;
            // 28: one elect + D alternating; 29: one elect, box loop unrolled x4
            if (mode_in == 26 || mode_in == 28 || mode_in == 29) {
                if (elect_one()) {
                    if (mode_in == 29) {
#pragma unroll 4
                        for (int i = 0; i < nbox; ++i) issue_box(tmem, smem_u32(smem + (i % kNSlots) * kSlot) >> 4, b0, i / nch);
                    } else {
                        for (int i = 0; i < nbox; ++i) {
                            const uint32_t a0 = smem_u32(smem + (i % kNSlots) * kSlot) >> 4;
                            if (mode_in == 26) issue_box(tmem, a0, b0, i / nch);
                            else
#pragma unroll
                                for (int m = 0; m < 6; ++m)
                                    umma_f16<kCollNone>(tmem + (m % 2) * 96, a0 + (((m / 2) * 64 + (m % 2) * 32) >> 4), kDescHiSw64,
                                                        b0 + (((m / 2) * 3 * 32 * 64 + (m % 2) * 32) >> 4), kDescHiSw64, make_idesc_f16(128, 96), 1u);
                        }
                    }
                    umma_commit(&done[0]);
                }
                __syncwarp();
            } else {
                for (int i = 0; i < nbox; ++i) {
                    const uint32_t a0 = smem_u32(smem + (i % kNSlots) * kSlot) >> 4;
                    if (elect_one()) {
#pragma unroll
                        for (int m = 0; m < 6; ++m)
                            umma_f16<kCollNone>(tmem + (m % 2) * 96, a0 + (((m / 2) * 64 + (m % 2) * 32) >> 4), kDescHiSw64,
                                                b0 + (((m / 2) * 3 * 32 * 64 + (m % 2) * 32) >> 4), kDescHiSw64, make_idesc_f16(128, 96), 1u);
                    }
                    __syncwarp();
                }
                if (elect_one()) umma_commit(&done[0]);
                __syncwarp();
            }
            mbar_wait(&done[0], 0);
            if (lane == 0) out[0] = clock64() - t0;
        } else if ((warp == 2 || (warp == 3 && mode_in == 25))) {
            const int w = warp - 2;
            const int n = mode_in == 25 ? nbox / 2 : nbox;
            for (int i = 0; i < n; ++i) {
                const int s = i % kNSlots;
                if (mode_in == 24) while (flag[s] < 0) {}
                if (elect_one()) {
                    issue_box(tmem + w * 256, smem_u32(smem + s * kSlot) >> 4, b0, i / nch);
                    if (mode_in == 20 || mode_in == 24 || mode_in == 25) umma_commit(&empty[s]);
                    if (mode_in == 21 && i % 3 == 2) umma_commit(&empty[s]);
                    if (mode_in == 23) umma_commit(&empty[0]);
                }
                __syncwarp();
            }
            if (elect_one()) { umma_commit(&done[w]); }
            __syncwarp();
            mbar_wait(&done[w], 0);
            if (lane == 0) out[w] = clock64() - t0;
        }
    } else if (warp < 2) {  // producers (warp 1 only in mode 2)
        if (warp < nstreams && lane == 0) {
            const int base = warp * per;
            int s = 0; uint32_t ph = 0;
            for (int i = 0; i < nbox; ++i) {
                mbar_wait(&empty[base + s], ph ^ 1);
                mbar_expect_tx(&full[base + s], kCopy);
                bulk_load(smem + (base + s) * kSlot, src + (size_t)((i * 7 + warp * 3) % 64) * kCopy, kCopy, &full[base + s]);
                if (++s == per) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp < 4) {  // issuers
        const int w = warp - 2;
        if (mode == 1) {
            int s = 0; uint32_t ph = 0;
            for (int i = 0; i < nbox; ++i) {
                const bool mine = (i & 1) == w;
                if (mine) {
                    mbar_wait(&full[s], ph);
                    if (fence) tc_fence_after();
                    if (i > 0) asm volatile("bar.sync %0, 64;" ::"r"(1 + w) : "memory");
                }
                if (mine && elect_one()) {
                    issue_box(tmem, smem_u32(smem + s * kSlot) >> 4, b0, i / nch);
                    umma_commit(&empty[s]);
                }
                __syncwarp();
                if (mine) asm volatile("bar.arrive %0, 64;" ::"r"(2 - w) : "memory");
                if (++s == per) { s = 0; ph ^= 1; }
            }
            if (nbox > 0 && (nbox & 1) == w) asm volatile("bar.sync %0, 64;" ::"r"(1 + w) : "memory");
            if (elect_one()) { umma_commit(&done[w]); }
            __syncwarp();
            mbar_wait(&done[w], 0);
        } else if (w < nstreams) {
            const int base = w * per;
            int s = 0; uint32_t ph = 0;
            if (mode == 3) mbar_wait(&full[base], 0);
            for (int i = 0; i < nbox; i += (mode == 6 ? G : 1)) {
                if (mode == 0 || mode == 2) {
                    mbar_wait(&full[base + s], ph);
                } else if (mode == 3) {
                    if (i + 1 < nbox) {
                        const int ns = s + 1 == per ? 0 : s + 1;
                        mbar_wait(&full[base + ns], ns == 0 ? ph ^ 1 : ph);
                    }
                } else if (mode == 4) {
                    while (flag[s] < i / per + 1) {}
                } else if (mode == 5) {
                    while (!mbar_test_wait(&full[base + s], ph)) {}
                } else if (mode == 6) {
                    for (;;) {
                        bool all = true;
                        int ss = s; uint32_t pp = ph;
                        for (int g = 0; g < G && i + g < nbox; ++g) {
                            all &= mbar_test_wait(&full[ss], pp);
                            if (++ss == per) { ss = 0; pp ^= 1; }
                        }
                        if (all) break;
                    }
                }
                if (fence) tc_fence_after();
                if (elect_one()) {
                    if (mode == 6) {
                        int ss = s;
                        for (int g = 0; g < G && i + g < nbox; ++g) {
                            issue_box(tmem + w * 256, smem_u32(smem + ss * kSlot) >> 4, b0, (i + g) / nch);
                            umma_commit(&empty[ss]);
                            if (++ss == per) ss = 0;
                        }
                    } else {
                        issue_box(tmem + w * 256, smem_u32(smem + (base + s) * kSlot) >> 4, b0, i / nch);
                        umma_commit(&empty[base + s]);
                    }
                }
                __syncwarp();
                for (int g = 0; g < (mode == 6 ? G : 1); ++g)
                    if (++s == per) { s = 0; ph ^= 1; }
            }
            if (elect_one()) { umma_commit(&done[w]); }
            __syncwarp();
            mbar_wait(&done[w], 0);
        }
        if (lane == 0) out[w] = clock64() - t0;
    } else if (warp == 4 && mode == 4) {  // watcher
        if (lane == 0) {
            int s = 0; uint32_t ph = 0;
            for (int i = 0; i < nbox; ++i) {
                mbar_wait(&full[s], ph);
                flag[s] = i / per + 1;
                if (++s == per) { s = 0; ph ^= 1; }
            }
        }
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc<512>(tmem);
}

int main() {
    long long* d; cudaMalloc(&d, 16 * 8);
    uint8_t* src; cudaMalloc(&src, 64 * kCopy); cudaMemset(src, 0x2c, 64 * kCopy);
    const int smem = 8 * 16896 + 2 * kBBytes + 4096;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int nbox = 600;
    const char* names[] = {"static: commit/box", "static: commit/3 boxes", "static: no commit", "static: commit one bar", "static: commit+flag", "static: 2 warps", "static: one elect", "static: alt D", "static: one elect+alt D", "static: one elect x4"};
    const char* names0[] = {"one issuer", "two alternating", "two streams", "one, look-ahead", "one + watcher flag", "one, test_wait", "one, G-box units"};
    for (int mode : {70, 71, 72, 73, 74, 75, 76, 77}) {
        for (int G : {1, 2, 3, 6}) {
            if (mode < 30 && G != 1) continue;
            if (mode >= 30 && mode < 40 && G > 2) continue;
            if (mode >= 40 && G != 2) continue;
            for (int nch : {2, 5}) {
            cudaMemset(d, 0, 16 * 8);
            probe<<<1, 192, smem>>>(src, d, mode, nbox, nch, G);
            probe<<<1, 192, smem>>>(src, d, mode, nbox, nch, G);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("mode %d: error %s\n", mode, cudaGetErrorString(e)); return 1; }
            long long h[16]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
            const long long t = h[0] > h[1] ? h[0] : h[1];
            const int streams = mode == 2 ? 2 : 1;
            printf("mode %d (%-18s G=%d): %8lld cycles, %6.1f per box (6 MMAs; tensor floor 336, per box of %d stream%s)\n", mode, mode >= 40 ? "static 12/box, D pattern" : mode >= 30 ? (mode == 30 ? "pipeline, one elect" : mode == 32 ? "pipeline, one elect, watcher flag" : mode == 33 ? "pipeline, NO full wait (unsafe)" : "pipeline, one elect, look-ahead") : mode >= 20 ? names[mode - 20] : names0[mode % 10], G,
                   t, (double)t / (nbox * streams * (mode >= 30 ? G : 1)), streams, streams > 1 ? "s" : "");
            printf("      (nch %d)\n", nch);
            }
        }
    }
    return 0;
}
