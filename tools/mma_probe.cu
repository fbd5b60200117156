// Micro-probe (measurement tool, not product code): cycles per tcgen05.mma (M=128, K=16, fp16, SS operands,
// 64 B-swizzled K-major rows as K1 uses them) as a function of N and of A-collector reuse.
// Answers the design question in DESIGN.md "narrow N": is the SS MMA shared-memory-read bound at N=32/64,
// and does collector::a::fill/use/lastuse remove the repeated A reads?
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_probe tools/mma_probe.cu
//   run  : tools/mma_probe  (prints one line per configuration)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../video_restore_b200/csrc/sm100_ptx.cuh"

using namespace vr::ptx;

// mode 0: every MMA has its own A tile, no hints
// mode 1: groups of 3 MMAs share A (different B, different D), hints fill/use/lastuse
// mode 2: groups of 3 MMAs share A, no hints
// mode 3: groups of 3 MMAs share A and B is also shared (upper bound: only D changes)
template <int N, int MODE>
__global__ void __launch_bounds__(128, 1) probe_kernel(long long* cycles_out, int n_mma) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    // fill shared memory with small finite fp16 values
    for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x2c002c00u;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
    }
    if (threadIdx.x < 32) tmem_alloc<512>(&tmem_slot);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (threadIdx.x < 32 && elect_one()) {
        constexpr uint32_t idesc = make_idesc_f16(128, N);
        const uint32_t a_base = smem_u32(smem);
        const uint32_t b_base = a_base + 128 * 1024;
        constexpr int kDSlots = 512 / N > 8 ? 8 : 512 / N;
        long long t0 = clock64();
        for (int i = 0; i < n_mma; i += 12) {
#pragma unroll
            for (int j = 0; j < 12; ++j) {
                const int g = j / 3, m = j % 3;
                const uint32_t a_off = (MODE == 0) ? (j * 8320u) : (g * 8320u);
                const uint32_t b_off = (MODE == 3) ? 0u : (j % 3) * N * 64u;
                const uint32_t ad = (a_base >> 4) + (a_off >> 4);
                const uint32_t bd = (b_base >> 4) + (b_off >> 4);
                const uint32_t d = tmem + ((g + m) % kDSlots) * N;
                if (MODE == 1) {
                    if (m == 0) umma_f16<kCollFill>(d, ad, kDescHiSw64, bd, kDescHiSw64, idesc, 1u);
                    else if (m == 1) umma_f16<kCollUse>(d, ad, kDescHiSw64, bd, kDescHiSw64, idesc, 1u);
                    else umma_f16<kCollLastUse>(d, ad, kDescHiSw64, bd, kDescHiSw64, idesc, 1u);
                } else {
                    umma_f16<kCollNone>(d, ad, kDescHiSw64, bd, kDescHiSw64, idesc, 1u);
                }
            }
        }
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        long long t1 = clock64();
        cycles_out[blockIdx.x] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc<512>(tmem);
}

template <int N, int MODE>
void run(int n_ctas, const char* label) {
    long long* d;
    cudaMalloc(&d, n_ctas * sizeof(long long));
    const int smem = 210 * 1024;
    cudaFuncSetAttribute(probe_kernel<N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int n_mma = 12 * 400;
    probe_kernel<N, MODE><<<n_ctas, 128, smem>>>(d, n_mma);  // warm
    probe_kernel<N, MODE><<<n_ctas, 128, smem>>>(d, n_mma);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        printf("N=%d %s: CUDA error %s\n", N, label, cudaGetErrorString(e));
        return;
    }
    long long h[256];
    cudaMemcpy(h, d, n_ctas * sizeof(long long), cudaMemcpyDeviceToHost);
    long long mx = 0, mn = 1ll << 60;
    for (int i = 0; i < n_ctas; ++i) {
        if (h[i] > mx) mx = h[i];
        if (h[i] < mn) mn = h[i];
    }
    const double ideal = 128.0 * N / 256.0;  // cycles per MMA at tensor peak (8192 flop/clk/SM)
    printf("N=%3d %-22s ctas=%3d cyc/mma min=%7.2f max=%7.2f  tensor-peak=%6.1f  frac(max)=%.3f\n", N, label, n_ctas,
           double(mn) / n_mma, double(mx) / n_mma, ideal, ideal / (double(mx) / n_mma));
    cudaFree(d);
}

int main() {
    int sms = 148;
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    sms = p.multiProcessorCount;
    printf("device %s, %d SMs, clock %d kHz\n", p.name, sms, p.clockRate);
    for (int ctas : {1, sms}) {
        run<16, 0>(ctas, "distinct-A");
        run<16, 1>(ctas, "A x3 +hints");
        run<32, 0>(ctas, "distinct-A");
        run<32, 1>(ctas, "A x3 +hints");
        run<32, 2>(ctas, "A x3 no-hints");
        run<32, 3>(ctas, "A x3, same B");
        run<64, 0>(ctas, "distinct-A");
        run<64, 1>(ctas, "A x3 +hints");
        run<64, 2>(ctas, "A x3 no-hints");
        run<128, 0>(ctas, "distinct-A");
        run<128, 1>(ctas, "A x3 +hints");
        run<256, 0>(ctas, "distinct-A");
        run<256, 1>(ctas, "A x3 +hints");
    }
    return 0;
}
