// Micro-probe: how far can the issuing thread run ahead of the tensor pipe? Timestamps after each tcgen05.mma issue
// (N = 96, M = 128, K = 16: ~55 cycles of execution each). Also: cost of an already-satisfied mbarrier try_wait
// and of tcgen05.fence::after_thread_sync on the issuing thread.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../video_restore_b200/csrc/sm100_ptx.cuh"
using namespace vr::ptx;
__global__ void __launch_bounds__(128, 1) q_kernel(long long* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar, done;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x2c002c00u;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&done, 1); fence_mbar_init(); }
    if (threadIdx.x < 32) tmem_alloc<512>(&slot);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = slot;
    if (threadIdx.x < 32 && elect_one()) {
        const uint32_t a0 = smem_u32(smem) >> 4, b0 = a0 + (51200 >> 4);
        long long ts[97];
        ts[0] = clock64();
#pragma unroll
        for (int i = 0; i < 96; ++i) {
            umma_f16<kCollNone>(tmem + (i % 2) * 96, a0 + ((i % 6) * 520), kDescHiSw64, b0 + ((i % 3) * 384), kDescHiSw64,
                                make_idesc_f16(128, 96), 1u);
            ts[i + 1] = clock64();
        }
        umma_commit(&done);
        mbar_wait(&done, 0);
        long long t_end = clock64();
        for (int i = 0; i <= 96; ++i) out[i] = ts[i] - ts[0];
        out[97] = t_end - ts[0];
        // cost of a satisfied try_wait and of the fence
        mbar_arrive(&bar);
        long long t0 = clock64();
        for (int i = 0; i < 16; ++i) mbar_try_wait(&bar, 0);
        long long t1 = clock64();
        for (int i = 0; i < 16; ++i) tc_fence_after();
        long long t2 = clock64();
        out[98] = (t1 - t0) / 16;
        out[99] = (t2 - t1) / 16;
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc<512>(tmem);
}
int main() {
    long long* d; cudaMalloc(&d, 128 * 8);
    cudaFuncSetAttribute(q_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024);
    q_kernel<<<1, 128, 210 * 1024>>>(d);
    q_kernel<<<1, 128, 210 * 1024>>>(d);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("error\n"); return 1; }
    long long h[128]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("issue-return timestamps (cycles since first issue), N=96 MMAs:\n");
    for (int i = 1; i <= 96; ++i) printf("%lld%s", h[i], i % 16 == 0 ? "\n" : " ");
    printf("all 96 complete at %lld cycles (%.1f per MMA)\n", h[97], h[97] / 96.0);
    printf("satisfied mbarrier.try_wait: %lld cycles; tcgen05.fence::after_thread_sync: %lld cycles\n", h[98], h[99]);
    return 0;
}
