// Micro-probe: latency between the completion of a batch of MMAs and the arrival of its tcgen05.commit on the mbarrier,
// as seen by the issuing thread that keeps issuing the next batches (ring of 3, like K1's smem stages).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../video_restore_b200/csrc/sm100_ptx.cuh"
using namespace vr::ptx;
__global__ void __launch_bounds__(128, 1) c_kernel(long long* out, int depth) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t ring[8], done;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x2c002c00u;
    if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(&ring[i], 1); mbar_init(&done, 1); fence_mbar_init(); }
    if (threadIdx.x < 32) tmem_alloc<512>(&slot);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = slot;
    if (threadIdx.x < 32 && elect_one()) {
        const uint32_t a0 = smem_u32(smem) >> 4, b0 = a0 + (51200 >> 4);
        long long t0 = clock64();
        for (int s = 0; s < 40; ++s) {
            long long w0 = clock64();
            if (s >= depth) mbar_wait(&ring[s % depth], ((s / depth) - 1) & 1);
            long long w1 = clock64();
            if (s < 40) { out[2 * s] = w0 - t0; out[2 * s + 1] = w1 - w0; }
#pragma unroll
            for (int i = 0; i < 36; ++i)
                umma_f16<kCollNone>(tmem + (i % 2) * 96, a0 + ((i % 6) * 520), kDescHiSw64, b0 + ((i % 3) * 384), kDescHiSw64,
                                    make_idesc_f16(128, 96), 1u);
            umma_commit(&ring[s % depth]);
        }
        umma_commit(&done);
        mbar_wait(&done, 0);
        out[100] = clock64() - t0;
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc<512>(tmem);
}
int main() {
    long long* d; cudaMalloc(&d, 128 * 8);
    cudaFuncSetAttribute(c_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024);
    for (int depth : {2, 3, 4, 6}) {
        c_kernel<<<1, 128, 210 * 1024>>>(d, depth);
        c_kernel<<<1, 128, 210 * 1024>>>(d, depth);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("error\n"); return 1; }
        long long h[128]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        printf("ring depth %d: total %lld cycles for 40 stages x 36 MMAs (%.1f per stage; pure MMA = 2016)\n  wait durations:", depth, h[100], h[100] / 40.0);
        for (int s = 0; s < 16; ++s) printf(" %lld", h[2 * s + 1]);
        printf("\n  stage start times:");
        for (int s = 0; s < 12; ++s) printf(" %lld", h[2 * s]);
        printf("\n");
    }
    return 0;
}
