// Micro-probe: the exact per-stage MMA sequence of K1 (dy-stacked N, TH = 4, KC = 32) in isolation, for several
// orders of the input-row loop. Prints cycles per stage (36 MMAs) against the smem-read model.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/mma_seq_probe tools/mma_seq_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../video_restore_b200/csrc/sm100_ptx.cuh"
using namespace vr::ptx;

// ORDER 0: dx,k outer / rho inner (kernel)   1: rho outer / dx,k inner   2: dx,k outer / rho in order 0,3,1,4,2,5
// ORDER 3: like 0 but every MMA is N = Cout on disjoint rows (old dy-separate formulation, 72 MMAs)
template <int N, int ORDER, int SYNC>
__global__ void __launch_bounds__(128, 1) seq_kernel(long long* out, int n_stages) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint64_t ring[4];
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x2c002c00u;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); for (int i = 0; i < 4; ++i) mbar_init(&ring[i], 1); fence_mbar_init(); }
    if (threadIdx.x < 32) tmem_alloc<512>(&slot);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = slot;
    constexpr int TH = 4, PITCH = 130;
    if (threadIdx.x < 32 && elect_one()) {
        const uint32_t a0 = smem_u32(smem) >> 4, b0 = a0 + (51200 >> 4);
        long long t0 = clock64();
        for (int s = 0; s < n_stages; ++s) {
            if (SYNC == 2 && s >= 3) { mbar_wait(&ring[s % 3], ((s / 3) - 1) & 1); tc_fence_after(); }
            if (ORDER == 1) {
#pragma unroll
                for (int rho = 0; rho < TH + 2; ++rho)
#pragma unroll
                    for (int dx = 0; dx < 3; ++dx)
#pragma unroll
                        for (int k = 0; k < 2; ++k) {
                            const int dy_lo = rho - 3 > 0 ? rho - 3 : 0, dy_hi = rho < 2 ? rho : 2, nblk = dy_hi - dy_lo + 1;
                            umma_f16<kCollNone>(tmem + (rho - dy_hi) * N, a0 + (((rho * PITCH + dx) * 64 + k * 32) >> 4), kDescHiSw64,
                                                b0 + ((((dx * 3 + (2 - dy_hi)) * N) * 64 + k * 32) >> 4), kDescHiSw64,
                                                make_idesc_f16(128, nblk * N), 1u);
                        }
            } else {
#pragma unroll
                for (int dx = 0; dx < 3; ++dx)
#pragma unroll
                    for (int k = 0; k < 2; ++k)
#pragma unroll
                        for (int i = 0; i < TH + 2; ++i) {
                            const int perm[6] = {0, 3, 1, 4, 2, 5};
                            const int rho = ORDER == 2 ? perm[i] : i;
                            const int dy_lo = rho - 3 > 0 ? rho - 3 : 0, dy_hi = rho < 2 ? rho : 2, nblk = dy_hi - dy_lo + 1;
                            const uint32_t a = a0 + (((rho * PITCH + dx) * 64 + k * 32) >> 4);
                            if (ORDER == 3) {
#pragma unroll
                                for (int dy = 0; dy < 3; ++dy) {
                                    if (dy < dy_lo || dy > dy_hi) continue;
                                    umma_f16<kCollNone>(tmem + (rho - dy) * N, a, kDescHiSw64,
                                                        b0 + ((((dx * 3 + (2 - dy)) * N) * 64 + k * 32) >> 4), kDescHiSw64,
                                                        make_idesc_f16(128, N), 1u);
                                }
                            } else {
                                umma_f16<kCollNone>(tmem + (rho - dy_hi) * N, a, kDescHiSw64,
                                                    b0 + ((((dx * 3 + (2 - dy_hi)) * N) * 64 + k * 32) >> 4), kDescHiSw64,
                                                    make_idesc_f16(128, nblk * N), 1u);
                            }
                        }
            }
            if (SYNC >= 1) umma_commit(&ring[s % 3]);
        }
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        out[blockIdx.x] = clock64() - t0;
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc<512>(tmem);
}
template <int N, int ORDER, int SYNC = 0>
void run(const char* label, double model) {
    long long* d; cudaMalloc(&d, 148 * 8);
    const int smem = 210 * 1024, stages = 200;
    cudaFuncSetAttribute(seq_kernel<N, ORDER, SYNC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    seq_kernel<N, ORDER, SYNC><<<148, 128, smem>>>(d, stages);
    seq_kernel<N, ORDER, SYNC><<<148, 128, smem>>>(d, stages);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("N=%d %s: %s\n", N, label, cudaGetErrorString(e)); return; }
    long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0; for (auto v : h) mx = v > mx ? v : mx;
    printf("Cout=%2d %-34s %8.1f cycles/stage  (smem-read model %6.0f, tensor-math floor %5.0f)\n", N, label,
           double(mx) / stages, model, 72.0 * 128 * N / 256);
    cudaFree(d);
}
int main() {
    run<32, 0>("dx,k outer; rho inner (kernel)", 6 * 286.0);
    run<32, 1>("rho outer; dx,k inner", 6 * 286.0);
    run<32, 2>("dx,k outer; rho 0,3,1,4,2,5", 6 * 286.0);
    run<32, 3>("72 separate N=32 MMAs", 72 * 40.0);
    run<32, 0, 1>("kernel order + commit per stage", 6 * 286.0);
    run<32, 0, 2>("kernel order + commit + ring wait", 6 * 286.0);
    run<64, 0>("dx,k outer; rho inner (kernel)", 6 * 416.0);
    run<64, 0, 1>("kernel order + commit per stage", 6 * 416.0);
    run<64, 0, 2>("kernel order + commit + ring wait", 6 * 416.0);
    run<64, 1>("rho outer; dx,k inner", 6 * 416.0);
    run<64, 2>("dx,k outer; rho 0,3,1,4,2,5", 6 * 416.0);
    run<64, 3>("72 separate N=64 MMAs", 72 * 48.0);
    return 0;
}
