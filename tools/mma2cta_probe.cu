// Micro-probe: tcgen05.mma.cta_group::2 semantics on B200. A cluster of two CTAs computes D[256 x N] = A[256 x 16] * B[N x 16]^T
// with ONE instruction issued by the leader: each CTA supplies its own 128 rows of A and HALF of B's rows (rows 0..N/2-1 in
// rank 0, N/2..N-1 in rank 1) at the same shared-memory offsets; each CTA finds its 128 rows of D in its own TMEM.
// Checks the operand split / D placement K3 relies on, the multicast commit and tcgen05.alloc.cta_group::2.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma2cta_probe tools/mma2cta_probe.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include "../video_restore_b200/csrc/sm100_ptx.cuh"
using namespace vr::ptx;

constexpr int N = 96;

__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// x: [2][128][32] fp16 (A rows of rank 0 / 1, 32 channels, only k = 0..15 used), w: [N][32] fp16, out: [256][N] fp32
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) probe(const __half* x, const __half* w, float* out, int ksteps) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t done;
    __shared__ uint32_t slot;
    const uint32_t rank = cluster_rank();
    uint8_t* a_sm = smem;            // 128 rows x 64 B, SWIZZLE_64B
    uint8_t* b_sm = smem + 8192;     // N/2 rows x 64 B
    // swizzled K-major fill: element (row r, channel c) at r*64 + (((c>>3) ^ ((r>>1)&3)) << 4) + (c&7)*2
    for (int i = threadIdx.x; i < 128 * 32; i += blockDim.x) {
        const int r = i >> 5, c = i & 31;
        *reinterpret_cast<__half*>(a_sm + r * 64 + ((((c >> 3) ^ ((r >> 1) & 3))) << 4) + (c & 7) * 2) = x[(rank * 128 + r) * 32 + c];
    }
    for (int i = threadIdx.x; i < (N / 2) * 32; i += blockDim.x) {
        const int r = i >> 5, c = i & 31;
        *reinterpret_cast<__half*>(b_sm + r * 64 + ((((c >> 3) ^ ((r >> 1) & 3))) << 4) + (c & 7) * 2) = w[(rank * (N / 2) + r) * 32 + c];
    }
    if (threadIdx.x == 0) { mbar_init(&done, 1); fence_mbar_init(); }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(128u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // both CTAs' operands, barriers and TMEM are ready
    tc_fence_after();
    const uint32_t tmem = slot;
    if (rank == 0 && threadIdx.x < 32) {
        if (elect_one()) {
            const uint32_t a0 = smem_u32(a_sm) >> 4, b0 = smem_u32(b_sm) >> 4;
            const uint32_t idesc = make_idesc_f16(256, N);
            for (int k = 0; k < ksteps; ++k) {
                const uint64_t ad = (static_cast<uint64_t>(kDescHiSw64) << 32) | (a0 + k * 2);
                const uint64_t bd = (static_cast<uint64_t>(kDescHiSw64) << 32) | (b0 + k * 2);
                asm volatile(
                    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc),
                    "r"(k > 0 ? 1u : 0u)
                    : "memory");
            }
            // arrive on `done` in BOTH CTAs once the MMAs have completed
            asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                             smem_u32(&done)),
                         "h"(static_cast<uint16_t>(3))
                         : "memory");
        }
        __syncwarp();
    }
    mbar_wait(&done, 0);
    tc_fence_after();
    {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        for (int g = 0; g < N / 32; ++g) {
            float v[32];
            tmem_ld32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + g * 32, v);
            for (int j = 0; j < 32; ++j) out[(rank * 128 + warp * 32 + lane) * N + g * 32 + j] = v[j];
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u) : "memory");
}

int main() {
    std::vector<__half> hx(256 * 32), hw(N * 32);
    std::vector<float> fx(256 * 32), fw(N * 32);
    uint32_t s = 1u;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (static_cast<float>(s >> 8) / 16777216.f - 0.5f); };
    for (size_t i = 0; i < hx.size(); ++i) { hx[i] = __float2half(rnd()); fx[i] = __half2float(hx[i]); }
    for (size_t i = 0; i < hw.size(); ++i) { hw[i] = __float2half(rnd()); fw[i] = __half2float(hw[i]); }
    __half *dx, *dw; float* dout;
    cudaMalloc(&dx, hx.size() * 2); cudaMalloc(&dw, hw.size() * 2); cudaMalloc(&dout, 256 * N * 4);
    cudaMemcpy(dx, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dw, hw.data(), hw.size() * 2, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 1024);
    for (int ksteps : {1, 2}) {
        cudaMemset(dout, 0, 256 * N * 4);
        probe<<<2, 128, 32 * 1024>>>(dx, dw, dout, ksteps);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("ksteps %d: error %s\n", ksteps, cudaGetErrorString(e)); return 1; }
        std::vector<float> ho(256 * N);
        cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost);
        double max_err = 0; int bad = 0;
        for (int r = 0; r < 256; ++r)
            for (int n = 0; n < N; ++n) {
                double ref = 0;
                for (int k = 0; k < 16 * ksteps; ++k) ref += static_cast<double>(fx[r * 32 + k]) * fw[n * 32 + k];
                const double err = std::fabs(ref - ho[r * N + n]);
                if (err > max_err) max_err = err;
                if (err > 1e-3) { if (bad < 5) printf("  mismatch r=%d n=%d got %f want %f\n", r, n, ho[r * N + n], ref); ++bad; }
            }
        printf("cta_group::2 M=256 N=%d K=%d: max err %.3e, %d mismatches of %d -> %s\n", N, 16 * ksteps, max_err, bad, 256 * N,
               bad ? "FAIL" : "OK (A rows: own CTA; B rows: first half rank 0, second half rank 1; D rows in own TMEM)");
    }
    return 0;
}
