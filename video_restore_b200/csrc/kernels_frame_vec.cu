// Vectorised variants of the HBM-bound frame kernels (K3 temporal / CLAHE / unsharp, K2 blend) for the common
// aligned case (dense rows, W % 16 == 0, 16 B aligned bases): 16 pixels = 48 B = 3 x 128-bit accesses per thread,
// fp32 staging in shared memory, per-warp histograms, float4 shared-memory reads. Arithmetic is IDENTICAL to the scalar
// kernels in kernels_frame.cu (same __fmul_rn/__fadd_rn order), so both paths are bit-exact against the oracle; the
// scalar kernels remain the fallback for ragged shapes.
#include "vr_common.h"

#include <cstring>

namespace vr {

#define VR_LAUNCH_CHECK(dev)                                   \
    do {                                                       \
        VR_CUDA_CHECK(cudaGetLastError(), (dev).err);          \
        (dev).launches++;                                      \
    } while (0)

union Px16 {  // 16 BGR pixels
    uint4 q[3];
    uint8_t b[48];
};
__device__ __forceinline__ void load_px16(const uint8_t* p, Px16& v) {
    const uint4* s = reinterpret_cast<const uint4*>(p);
    v.q[0] = __ldg(s);
    v.q[1] = __ldg(s + 1);
    v.q[2] = __ldg(s + 2);
}
__device__ __forceinline__ void store_px16(uint8_t* p, const Px16& v) {
    uint4* d = reinterpret_cast<uint4*>(p);
    d[0] = v.q[0];
    d[1] = v.q[1];
    d[2] = v.q[2];
}
__device__ __forceinline__ uint8_t sat_u8(int v) { return static_cast<uint8_t>(min(max(v, 0), 255)); }
// sat_u8(__float2int_rn(v)) in one instruction (F2IP.U8.F32): round half to even, clamp to [0, 255], NaN -> 0
__device__ __forceinline__ uint32_t f2u8(float v) {
    uint32_t r;
    asm("cvt.rni.u8.f32 %0, %1;" : "=r"(r) : "f"(v));
    return r;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// ------------------------------------------------------------------------------------------------
// temporal
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
temporal_vec_kernel(const uint8_t* __restrict__ cur, const uint8_t* __restrict__ prev, uint8_t* __restrict__ dst,
                    size_t n_chunks, float alpha, float one_minus, float tau) {
    const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (i >= n_chunks) return;
    Px16 c, p, o;
    load_px16(cur + i * 48, c);
    load_px16(prev + i * 48, p);
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const int c0 = c.b[3 * k], c1 = c.b[3 * k + 1], c2 = c.b[3 * k + 2];
        const int p0 = p.b[3 * k], p1 = p.b[3 * k + 1], p2 = p.b[3 * k + 2];
        const int d = max(max(abs(c0 - p0), abs(c1 - p1)), abs(c2 - p2));
        if (static_cast<float>(d) < tau) {
            o.b[3 * k] = f2u8((__fadd_rn(__fmul_rn(one_minus, static_cast<float>(c0)), __fmul_rn(alpha, static_cast<float>(p0)))));
            o.b[3 * k + 1] = f2u8((__fadd_rn(__fmul_rn(one_minus, static_cast<float>(c1)), __fmul_rn(alpha, static_cast<float>(p1)))));
            o.b[3 * k + 2] = f2u8((__fadd_rn(__fmul_rn(one_minus, static_cast<float>(c2)), __fmul_rn(alpha, static_cast<float>(p2)))));
        } else {
            o.b[3 * k] = c.b[3 * k];
            o.b[3 * k + 1] = c.b[3 * k + 1];
            o.b[3 * k + 2] = c.b[3 * k + 2];
        }
    }
    store_px16(dst + i * 48, o);
}
bool try_temporal_vec(Device& dev, const uint8_t* cur, int64_t cstride, const uint8_t* prev, int64_t pstride, int H,
                      int W, uint8_t* dst, int64_t dstride, float alpha, float tau, int* rc) {
    const int64_t row = static_cast<int64_t>(W) * 3;
    const size_t px = static_cast<size_t>(H) * W;
    if (cstride != row || pstride != row || dstride != row || px % 16 != 0 || !aligned16(cur) || !aligned16(prev) ||
        !aligned16(dst))
        return false;
    const size_t n = px / 16;
    temporal_vec_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, dev.stream>>>(cur, prev, dst, n, alpha,
                                                                                        1.0f - alpha, tau);
    *rc = 0;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error(dev.err, std::string("temporal_vec: ") + cudaGetErrorString(e));
        *rc = -2;
    } else {
        dev.launches++;
    }
    return true;
}

// ------------------------------------------------------------------------------------------------
// CLAHE
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int luma_of_v(int b, int g, int r) { return (r * 4899 + g * 9617 + b * 1868 + 8192) >> 14; }

// grid = (blocks per tile, tiles). Each thread takes 16-pixel chunks of the tile; one private histogram per warp.
__global__ void __launch_bounds__(256)
clahe_hist_vec_kernel(const uint8_t* __restrict__ src, int64_t sstride, int tile_w, int tile_h, int tiles_x,
                      int32_t* __restrict__ hist) {
    __shared__ unsigned int s_hist[8][256];
    for (int i = threadIdx.x; i < 8 * 256; i += 256) (&s_hist[0][0])[i] = 0;
    __syncthreads();
    const int tile = blockIdx.y;
    const int tx0 = (tile % tiles_x) * tile_w, ty0 = (tile / tiles_x) * tile_h;
    const int cpr = tile_w / 16;  // chunks per tile row
    const int n_chunks = cpr * tile_h;
    unsigned int* my = s_hist[threadIdx.x >> 5];
    for (int i = blockIdx.x * 256 + threadIdx.x; i < n_chunks; i += gridDim.x * 256) {
        const int row = i / cpr, cx = i - row * cpr;
        Px16 v;
        load_px16(src + (ty0 + row) * sstride + static_cast<int64_t>(tx0 + cx * 16) * 3, v);
#pragma unroll
        for (int k = 0; k < 16; ++k) atomicAdd(&my[luma_of_v(v.b[3 * k], v.b[3 * k + 1], v.b[3 * k + 2])], 1u);
    }
    __syncthreads();
    unsigned int s = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += s_hist[w][threadIdx.x];
    if (s) atomicAdd(&hist[tile * 256 + threadIdx.x], static_cast<int>(s));
}

// kTemporal: the temporal stage fused in -- the un-blended result (next frame's "previous") goes to dst, the frame blended with
// `prev` to `blended`: one pass instead of two (reads src + prev, writes dst + blended; the separate temporal kernel re-read dst).
template <bool kTemporal>
__global__ void __launch_bounds__(256, kTemporal ? 3 : 5)
clahe_apply_vec_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int H, int W,
                       const uint8_t* __restrict__ lut, int tiles_x, int tiles_y, float inv_tw, float inv_th,
                       const uint8_t* __restrict__ prev, uint8_t* __restrict__ blended, float alpha, float one_minus, float tau) {
    const int cpr = W / 16;
    const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (i >= static_cast<size_t>(cpr) * H) return;
    const int y = static_cast<int>(i / cpr), x0 = static_cast<int>(i % cpr) * 16;
    Px16 v, o;
    load_px16(src + i * 48, v);
    const float tyf = __fsub_rn(__fmul_rn(static_cast<float>(y), inv_th), 0.5f);
    int ty1 = static_cast<int>(floorf(tyf));
    const float ya = __fsub_rn(tyf, static_cast<float>(ty1));
    const float ya1 = __fsub_rn(1.0f, ya);
    const int ty2 = min(ty1 + 1, tiles_y - 1);
    ty1 = max(ty1, 0);
    const uint8_t* lrow1 = lut + ty1 * tiles_x * 256;
    const uint8_t* lrow2 = lut + ty2 * tiles_x * 256;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const int b = v.b[3 * k], g = v.b[3 * k + 1], r = v.b[3 * k + 2];
        const int Y = luma_of_v(b, g, r);
        const int cr = min(max(((r - Y) * 11682 + (128 << 14) + 8192) >> 14, 0), 255);
        const int cb = min(max(((b - Y) * 9241 + (128 << 14) + 8192) >> 14, 0), 255);
        const float txf = __fsub_rn(__fmul_rn(static_cast<float>(x0 + k), inv_tw), 0.5f);
        int tx1 = static_cast<int>(floorf(txf));
        const float xa = __fsub_rn(txf, static_cast<float>(tx1));
        const float xa1 = __fsub_rn(1.0f, xa);
        const int tx2 = min(tx1 + 1, tiles_x - 1);
        tx1 = max(tx1, 0);
        const float l11 = static_cast<float>(__ldg(lrow1 + tx1 * 256 + Y));
        const float l12 = static_cast<float>(__ldg(lrow1 + tx2 * 256 + Y));
        const float l21 = static_cast<float>(__ldg(lrow2 + tx1 * 256 + Y));
        const float l22 = static_cast<float>(__ldg(lrow2 + tx2 * 256 + Y));
        const float top = __fadd_rn(__fmul_rn(l11, xa1), __fmul_rn(l12, xa));
        const float bot = __fadd_rn(__fmul_rn(l21, xa1), __fmul_rn(l22, xa));
        const float res = __fadd_rn(__fmul_rn(top, ya1), __fmul_rn(bot, ya));
        const int Yn = static_cast<int>(f2u8(res));
        const int crd = cr - 128, cbd = cb - 128;
        o.b[3 * k] = sat_u8(Yn + ((cbd * 29049 + 8192) >> 14));
        o.b[3 * k + 1] = sat_u8(Yn + ((cbd * -5636 + crd * -11698 + 8192) >> 14));
        o.b[3 * k + 2] = sat_u8(Yn + ((crd * 22987 + 8192) >> 14));
    }
    store_px16(dst + i * 48, o);
    if constexpr (kTemporal) {
        Px16 p, t;
        load_px16(prev + i * 48, p);
#pragma unroll
        for (int k = 0; k < 16; ++k) {  // temporal_vec_kernel's arithmetic, on the registers just produced
            const int c0 = o.b[3 * k], c1 = o.b[3 * k + 1], c2 = o.b[3 * k + 2];
            const int p0 = p.b[3 * k], p1 = p.b[3 * k + 1], p2 = p.b[3 * k + 2];
            const int d = max(max(abs(c0 - p0), abs(c1 - p1)), abs(c2 - p2));
            if (static_cast<float>(d) < tau) {
                t.b[3 * k] = f2u8((__fadd_rn(__fmul_rn(one_minus, static_cast<float>(c0)), __fmul_rn(alpha, static_cast<float>(p0)))));
                t.b[3 * k + 1] = f2u8((__fadd_rn(__fmul_rn(one_minus, static_cast<float>(c1)), __fmul_rn(alpha, static_cast<float>(p1)))));
                t.b[3 * k + 2] = f2u8((__fadd_rn(__fmul_rn(one_minus, static_cast<float>(c2)), __fmul_rn(alpha, static_cast<float>(p2)))));
            } else {
                t.b[3 * k] = o.b[3 * k];
                t.b[3 * k + 1] = o.b[3 * k + 1];
                t.b[3 * k + 2] = o.b[3 * k + 2];
            }
        }
        store_px16(blended + i * 48, t);
    }
}

bool clahe_vec_ok(const uint8_t* src, int64_t sstride, int H, int W, const uint8_t* dst, int64_t dstride, int grid_n) {
    const int64_t row = static_cast<int64_t>(W) * 3;
    return sstride == row && dstride == row && W % grid_n == 0 && H % grid_n == 0 && (W / grid_n) % 16 == 0 &&
           aligned16(src) && aligned16(dst);
}
int launch_clahe_hist_vec(Device& dev, const uint8_t* src, int64_t sstride, int tile_w, int tile_h, int tiles_x,
                          int ntiles, int32_t* d_hist) {
    const int n_chunks = tile_w / 16 * tile_h;
    int blocks = (n_chunks + 256 * 8 - 1) / (256 * 8);
    const int cap = (dev.sm_count * 8 + ntiles - 1) / ntiles;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    clahe_hist_vec_kernel<<<dim3(blocks, ntiles), 256, 0, dev.stream>>>(src, sstride, tile_w, tile_h, tiles_x, d_hist);
    VR_LAUNCH_CHECK(dev);
    return 0;
}
int launch_clahe_apply_vec(Device& dev, const uint8_t* src, uint8_t* dst, int H, int W, const uint8_t* d_lut,
                           int tiles_x, int tiles_y, float inv_tw, float inv_th, const TemporalFuse* tf) {
    const size_t n = static_cast<size_t>(W / 16) * H;
    const unsigned grid = static_cast<unsigned>((n + 255) / 256);
    if (tf)
        clahe_apply_vec_kernel<true><<<grid, 256, 0, dev.stream>>>(src, dst, H, W, d_lut, tiles_x, tiles_y, inv_tw, inv_th, tf->prev,
                                                                   tf->blended, tf->alpha, 1.0f - tf->alpha, tf->tau);
    else
        clahe_apply_vec_kernel<false><<<grid, 256, 0, dev.stream>>>(src, dst, H, W, d_lut, tiles_x, tiles_y, inv_tw, inv_th, nullptr,
                                                                    nullptr, 0.f, 0.f, 0.f);
    VR_LAUNCH_CHECK(dev);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// unsharp: 64 x 24 output pixels per block of 192 threads. fp32 input tile (30 rows x 216 floats) and horizontal-pass tile
// (30 x 192) in shared memory; every shared-memory access is a 128-bit one, and laid out so that the eight lanes of a
// quarter warp hit eight different 16-byte bank groups (row pitches of 55 and 49 float4: odd, so lanes that differ in the
// ROW do not collide; ncu on the first version of this kernel: 30 % of its shared-memory wavefronts were bank conflicts and
// the LSU pipe was 75 % busy -- shared-memory bandwidth, not HBM and not FP32 issue, was the limiter):
//   load        aligned 128-bit global loads, all of a thread's loads in flight together -> sixteen floats -> four float4
//               stores; lane L writes its four words in the order (s + L/2) & 3 so neighbouring lanes differ in bank group
//   horizontal  one item = 16 consecutive output BYTES of one row, lanes walk down the rows: ten float4 loads (the 34 inputs
//               its taps touch), four float4 stores -- 14 wavefronts per 16 outputs (4-byte items: 32)
//   vertical    one item = 6 rows x 4 consecutive bytes: twelve float4 loads, six 32-bit global stores; 192 items = one per thread
// What remains is instruction issue: the spec (oracle/filters.py unsharp_mask) is fp32 with a separate rounding after every
// multiply and every add, i.e. 13 x 1.25 (halo rows) + 13 + 3 FP32 instructions per output byte that no FMA can merge.
// ------------------------------------------------------------------------------------------------
constexpr int kUvW = 64, kUvH = 24, kUvR = 3, kUvThreads = 192;
constexpr int kUvCols = 216;     // staged floats per row: bytes -12 .. 203 of the row relative to the block's first byte
constexpr int kUvInPitch = 220;  // 55 float4
constexpr int kUvHPitch = 196;   // 49 float4 (192 used)
constexpr int kUvTh = kUvH + 2 * kUvR;
constexpr int kUvSmem = kUvTh * (kUvInPitch + kUvHPitch) * 4;  // 49 920 B: dynamic (above the 48 KB static limit)
__constant__ float c_taps7v[7];

__global__ void __launch_bounds__(kUvThreads)
unsharp_vec_kernel(const uint8_t* __restrict__ src, int64_t sstride, int H, int W, uint8_t* __restrict__ dst,
                   int64_t dstride, float amount) {
    constexpr int th = kUvTh, tw = kUvW + 2 * kUvR;
    extern __shared__ __align__(16) float s_uv[];
    float (*s_in)[kUvInPitch] = reinterpret_cast<float (*)[kUvInPitch]>(s_uv);  // centre pixel of output byte e sits at e + 12
    float (*s_h)[kUvHPitch] = reinterpret_cast<float (*)[kUvHPitch]>(s_uv + th * kUvInPitch);
    const int tid = threadIdx.x;
    const int bx0 = blockIdx.x * kUvW, by0 = blockIdx.y * kUvH;
    const bool interior = bx0 >= 6 && bx0 + kUvW + 6 <= W && by0 >= kUvR && by0 + kUvH + kUvR <= H;
    if (interior) {
        // the row segment bytes [-16, 208) around the block = 14 x 16 bytes; floats 4 .. 219 of it are kept (column = float - 4)
        constexpr int kQuads = 14, kIters = (th * kQuads + kUvThreads - 1) / kUvThreads;
        uint4 qv[kIters];
#pragma unroll
        for (int k = 0; k < kIters; ++k) {
            const int i = tid + k * kUvThreads;
            if (i < th * kQuads) {
                const int ty = i / kQuads, q4 = i - ty * kQuads;
                qv[k] = __ldg(reinterpret_cast<const uint4*>(src + (by0 - kUvR + ty) * sstride + static_cast<int64_t>(bx0) * 3 - 16) + q4);
            }
        }
        const int rot = (tid & 31) >> 1;
#pragma unroll
        for (int k = 0; k < kIters; ++k) {
            const int i = tid + k * kUvThreads;
            if (i < th * kQuads) {
                const int ty = i / kQuads, q4 = i - ty * kQuads;
#pragma unroll
                for (int st = 0; st < 4; ++st) {
                    const int m = (st + rot) & 3;
                    const uint32_t wv = m == 0 ? qv[k].x : m == 1 ? qv[k].y : m == 2 ? qv[k].z : qv[k].w;
                    const int col = q4 * 16 + m * 4 - 4;
                    if (col >= 0 && col < kUvCols)
                        *reinterpret_cast<float4*>(&s_in[ty][col]) =
                            make_float4(static_cast<float>(wv & 0xff), static_cast<float>((wv >> 8) & 0xff),
                                        static_cast<float>((wv >> 16) & 0xff), static_cast<float>(wv >> 24));
                }
            }
        }
    } else {
#pragma unroll 4
        for (int i = tid; i < th * tw; i += kUvThreads) {
            const int ty = i / tw, tx = i - ty * tw;
            int sy = by0 - kUvR + ty, sx = bx0 - kUvR + tx;
            // REFLECT_101
            const int py = 2 * (H - 1), pxp = 2 * (W - 1);
            if (H == 1) sy = 0; else { sy %= py; if (sy < 0) sy += py; if (sy >= H) sy = py - sy; }
            if (W == 1) sx = 0; else { sx %= pxp; if (sx < 0) sx += pxp; if (sx >= W) sx = pxp - sx; }
            const uint8_t* p = src + sy * sstride + static_cast<int64_t>(sx) * 3;
            s_in[ty][3 + tx * 3 + 0] = static_cast<float>(p[0]);
            s_in[ty][3 + tx * 3 + 1] = static_cast<float>(p[1]);
            s_in[ty][3 + tx * 3 + 2] = static_cast<float>(p[2]);
        }
        // floats 0..2 and 213..215 of a row stay unwritten here: the float4 loads below fetch them, no tap uses them
    }
    __syncthreads();
    // horizontal pass: output byte e reads the same channel of pixels -3 .. +3, i.e. s_in[e + 3 + 3 t], t = 0 .. 6
    constexpr int kWide = 16, kGroups = kUvW * 3 / kWide;  // 12 items per row
    for (int i = tid; i < th * kGroups; i += kUvThreads) {
        const int g = i / th, ty = i - g * th;  // consecutive lanes: consecutive rows
        const float4* in4 = reinterpret_cast<const float4*>(&s_in[ty][g * kWide]);
        float x[kWide + 24];
#pragma unroll
        for (int v = 0; v < (kWide + 24) / 4; ++v) {
            const float4 f = in4[v];
            x[4 * v] = f.x;
            x[4 * v + 1] = f.y;
            x[4 * v + 2] = f.z;
            x[4 * v + 3] = f.w;
        }
#pragma unroll
        for (int q4 = 0; q4 < kWide / 4; ++q4) {
            float h[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                h[q] = __fmul_rn(c_taps7v[0], x[q4 * 4 + q + 3]);
#pragma unroll
                for (int t = 1; t < 7; ++t) h[q] = __fadd_rn(h[q], __fmul_rn(c_taps7v[t], x[q4 * 4 + q + 3 + 3 * t]));
            }
            *reinterpret_cast<float4*>(&s_h[ty][g * kWide + q4 * 4]) = make_float4(h[0], h[1], h[2], h[3]);
        }
    }
    __syncthreads();
    // vertical pass: one item = kUvRows rows x 4 consecutive output bytes
    constexpr int kUvRows = 6, kColGroups = kUvW * 3 / 4;  // 48
    const float one_plus = __fadd_rn(1.0f, amount);
    for (int i = tid; i < (kUvH / kUvRows) * kColGroups; i += kUvThreads) {
        const int rg = i / kColGroups, cg = i - rg * kColGroups;
        float4 hv[kUvRows + 6];
#pragma unroll
        for (int t = 0; t < kUvRows + 6; ++t) hv[t] = *reinterpret_cast<const float4*>(&s_h[rg * kUvRows + t][cg * 4]);
#pragma unroll
        for (int q = 0; q < kUvRows; ++q) {
            const int y = by0 + rg * kUvRows + q;
            const float4 xin = *reinterpret_cast<const float4*>(&s_in[rg * kUvRows + q + kUvR][cg * 4 + 12]);
            float v[4], xi[4] = {xin.x, xin.y, xin.z, xin.w};
            const float* h0 = reinterpret_cast<const float*>(&hv[q]);
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] = __fmul_rn(c_taps7v[0], h0[e]);
#pragma unroll
            for (int t = 1; t < 7; ++t) {
                const float* ht = reinterpret_cast<const float*>(&hv[q + t]);
#pragma unroll
                for (int e = 0; e < 4; ++e) v[e] = __fadd_rn(v[e], __fmul_rn(c_taps7v[t], ht[e]));
            }
            uint32_t packed = 0;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float o = __fsub_rn(__fmul_rn(one_plus, xi[e]), __fmul_rn(amount, v[e]));
                packed |= f2u8(o) << (8 * e);
            }
            if (y < H && bx0 * 3 + cg * 4 + 3 < W * 3)
                *reinterpret_cast<uint32_t*>(dst + y * dstride + static_cast<int64_t>(bx0) * 3 + cg * 4) = packed;
        }
    }
}
bool try_unsharp_vec(Device& dev, const uint8_t* src, int64_t sstride, int H, int W, uint8_t* dst, int64_t dstride,
                     float amount, int* rc) {
    if (W % 64 != 0 || sstride % 16 != 0 || dstride % 4 != 0 || (reinterpret_cast<uintptr_t>(src) & 15) ||
        (reinterpret_cast<uintptr_t>(dst) & 3) || H < 2 || W < 2)
        return false;
    static bool taps_done[64] = {};
    *rc = 0;
    if (!taps_done[dev.ordinal & 63]) {
        double k[7], sum = 0;
        for (int i = 0; i < 7; ++i) {
            k[i] = std::exp(-0.5 * (i - 3) * (i - 3));
            sum += k[i];
        }
        float kf[7];
        for (int i = 0; i < 7; ++i) kf[i] = static_cast<float>(k[i] / sum);
        if (cudaMemcpyToSymbol(c_taps7v, kf, sizeof(kf)) != cudaSuccess) {
            set_error(dev.err, "unsharp_vec: cudaMemcpyToSymbol failed");
            *rc = -2;
            return true;
        }
        if (cudaFuncSetAttribute(unsharp_vec_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kUvSmem) != cudaSuccess) {
            set_error(dev.err, "unsharp_vec: cudaFuncSetAttribute failed");
            *rc = -2;
            return true;
        }
        taps_done[dev.ordinal & 63] = true;
    }
    dim3 grid(W / kUvW, (H + kUvH - 1) / kUvH);
    unsharp_vec_kernel<<<grid, kUvThreads, kUvSmem, dev.stream>>>(src, sstride, H, W, dst, dstride, amount);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error(dev.err, std::string("unsharp_vec: ") + cudaGetErrorString(e));
        *rc = -2;
    } else {
        dev.launches++;
    }
    return true;
}

}  // namespace vr
