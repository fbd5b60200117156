// K0 / K2 / K3: the bandwidth-bound kernels around the conv stack.
//   K0  pre        u8 BGR frame rect -> fp16 RGB NHWC(32) tile   (RealESRGANer.enhance/pre_process, [3P] utils.py)
//   K0b bilateral  cv2.bilateralFilter(frame,5,25,25)            (reference video_upscaler.py:496)
//   K1u upsample2x nearest x2 NHWC                               (F.interpolate in RRDBNet.forward, [3P])
//   K2  post_crop  tile_process crop-merge + clamp/round/BGR     ([3P] utils.py tile_process/post_process/enhance)
//   K2b post_blend seamless Gaussian gather blend                (README.md:8,236; spec SURVEY 8 A7 / oracle)
//   K3  unsharp, CLAHE (hist -> clip/LUT -> apply), temporal     (README.md:9-12; spec oracle/filters.py)
// Float arithmetic that the tests require to be bit-exact uses __fmul_rn/__fadd_rn (no FMA contraction) in the
// oracle's operation order.
#include "vr_common.h"

#include <cmath>
#include <cstring>

namespace vr {

__host__ __device__ inline int reflect101(int i, int n) {
    if (n == 1) return 0;
    const int period = 2 * (n - 1);
    int m = i % period;
    if (m < 0) m += period;
    return m >= n ? period - m : m;
}

// vectorised fast paths (kernels_frame_vec.cu); each returns false when the shape/alignment is not eligible
bool try_temporal_vec(Device& dev, const uint8_t* cur, int64_t cstride, const uint8_t* prev, int64_t pstride, int H,
                      int W, uint8_t* dst, int64_t dstride, float alpha, float tau, int* rc);
bool try_unsharp_vec(Device& dev, const uint8_t* src, int64_t sstride, int H, int W, uint8_t* dst, int64_t dstride,
                     float amount, int* rc);
bool clahe_vec_ok(const uint8_t* src, int64_t sstride, int H, int W, const uint8_t* dst, int64_t dstride, int grid_n);
int launch_clahe_hist_vec(Device& dev, const uint8_t* src, int64_t sstride, int tile_w, int tile_h, int tiles_x,
                          int ntiles, int32_t* d_hist);
int launch_clahe_apply_vec(Device& dev, const uint8_t* src, uint8_t* dst, int H, int W, const uint8_t* d_lut,
                           int tiles_x, int tiles_y, float inv_tw, float inv_th, const TemporalFuse* tf);

#define VR_LAUNCH_CHECK(dev)                                   \
    do {                                                       \
        VR_CUDA_CHECK(cudaGetLastError(), (dev).err);          \
        (dev).launches++;                                      \
    } while (0)

// ------------------------------------------------------------------------------------------------
// K1u nearest x2 (NHWC fp16, 16 B per thread)
// ------------------------------------------------------------------------------------------------
__global__ void upsample2x_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int H, int W, int vec_c) {
    const size_t total = static_cast<size_t>(4) * H * W * vec_c;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int v = static_cast<int>(i % vec_c);
        const size_t p = i / vec_c;
        const int x = static_cast<int>(p % (2 * W));
        const int y = static_cast<int>(p / (2 * W));
        dst[i] = __ldg(src + (static_cast<size_t>(y >> 1) * W + (x >> 1)) * vec_c + v);
    }
}
int launch_upsample2x(Device& dev, const __half* src, int H, int W, int C, __half* dst) {
    const int vec_c = C / 8;
    const size_t total = static_cast<size_t>(4) * H * W * vec_c;
    const int block = 256;
    size_t want = (total + block - 1) / block;
    const int grid = static_cast<int>(want < static_cast<size_t>(dev.sm_count) * 16 ? want : dev.sm_count * 16);
    upsample2x_kernel<<<grid, block, 0, dev.stream>>>(reinterpret_cast<const uint4*>(src),
                                                      reinterpret_cast<uint4*>(dst), H, W, vec_c);
    VR_LAUNCH_CHECK(dev);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// K0 pre: one thread per output pixel, 64 B (32 fp16 channels) stored as 4 x uint4
// ------------------------------------------------------------------------------------------------
__global__ void pre_kernel(const uint8_t* __restrict__ frame, int64_t stride, int H, int W, int x0, int y0, int w,
                           int h, int unshuffle, __half* __restrict__ dst, int dst_pitch, int dst_x0, int dst_y0) {
    const int ow = unshuffle ? w / 2 : w;
    const int oh = unshuffle ? h / 2 : h;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= ow * oh) return;
    const int ox = idx % ow, oy = idx / ow;
    __align__(16) __half v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __float2half(0.f);
    const float inv255 = 255.0f;
    if (!unshuffle) {
        // F.pad(..., 'reflect') past the bottom/right frame edge == REFLECT_101 index map
        const int sy = reflect101(y0 + oy, H), sx = reflect101(x0 + ox, W);
        const uint8_t* p = frame + sy * stride + static_cast<int64_t>(sx) * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) v[c] = __float2half_rn(__fdiv_rn(static_cast<float>(p[2 - c]), inv255));
    } else {
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
                const int sy = reflect101(y0 + 2 * oy + dy, H), sx = reflect101(x0 + 2 * ox + dx, W);
                const uint8_t* p = frame + sy * stride + static_cast<int64_t>(sx) * 3;
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    v[c * 4 + dy * 2 + dx] = __float2half_rn(__fdiv_rn(static_cast<float>(p[2 - c]), inv255));
            }
    }
    uint4* o = reinterpret_cast<uint4*>(dst + (static_cast<size_t>(dst_y0 + oy) * dst_pitch + dst_x0 + ox) * 32);
    const uint4* s = reinterpret_cast<const uint4*>(v);
#pragma unroll
    for (int i = 0; i < 4; ++i) o[i] = s[i];
}
int launch_pre(Device& dev, const uint8_t* frame, int64_t stride, int H, int W, int x0, int y0, int w, int h,
               int unshuffle, __half* dst, int dst_pitch, int dst_x0, int dst_y0) {
    const int n = unshuffle ? (w / 2) * (h / 2) : w * h;
    pre_kernel<<<(n + 255) / 256, 256, 0, dev.stream>>>(frame, stride, H, W, x0, y0, w, h, unshuffle, dst, dst_pitch,
                                                        dst_x0, dst_y0);
    VR_LAUNCH_CHECK(dev);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// K2 post: crop-merge
// ------------------------------------------------------------------------------------------------
// numpy's (clip(x, 0, 1) * 255.0).round() (half to even) as multiply + one saturating conversion (F2IP.U8.F32): clamping
// the rounded product to [0, 255] gives the same byte as rounding the clamped input's product for every float, NaN -> 0
__device__ __forceinline__ uint8_t quant_u8(float v) {
    uint32_t r;
    asm("cvt.rni.u8.f32 %0, %1;" : "=r"(r) : "f"(__fmul_rn(v, 255.0f)));
    return static_cast<uint8_t>(r);
}
__global__ void post_crop_kernel(const __half* __restrict__ tile, int tile_w, int crop_x0, int crop_y0, int w, int h,
                                 uint8_t* __restrict__ frame, int64_t stride, int dst_x0, int dst_y0) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= w * h) return;
    const int x = idx % w, y = idx / w;
    const uint2 q = __ldg(reinterpret_cast<const uint2*>(tile + (static_cast<size_t>(crop_y0 + y) * tile_w +
                                                                 (crop_x0 + x)) * 4));
    const __half2 rg = *reinterpret_cast<const __half2*>(&q.x);
    const __half2 b_ = *reinterpret_cast<const __half2*>(&q.y);
    uint8_t* o = frame + (dst_y0 + y) * stride + static_cast<int64_t>(dst_x0 + x) * 3;
    o[0] = quant_u8(__low2float(b_));
    o[1] = quant_u8(__high2float(rg));
    o[2] = quant_u8(__low2float(rg));
}
// 4 pixels per thread: four 64-bit tile loads in flight, three 32-bit stores (w % 4 == 0, destination 4-byte aligned)
__global__ void __launch_bounds__(128)
post_crop4_kernel(const __half* __restrict__ tile, int tile_w, int crop_x0, int crop_y0, int w4, int h,
                  uint8_t* __restrict__ frame, int64_t stride, int dst_x0, int dst_y0) {
    const int xq = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (xq >= w4) return;
    const uint2* src = reinterpret_cast<const uint2*>(tile + (static_cast<size_t>(crop_y0 + y) * tile_w + crop_x0 + xq * 4) * 4);
    uint2 q[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) q[k] = __ldg(src + k);
    uint32_t b[12];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const __half2 rg = *reinterpret_cast<const __half2*>(&q[k].x);
        const __half2 b_ = *reinterpret_cast<const __half2*>(&q[k].y);
        b[3 * k + 0] = quant_u8(__low2float(b_));
        b[3 * k + 1] = quant_u8(__high2float(rg));
        b[3 * k + 2] = quant_u8(__low2float(rg));
    }
    uint32_t* o32 = reinterpret_cast<uint32_t*>(frame + (dst_y0 + y) * stride + static_cast<int64_t>(dst_x0 + xq * 4) * 3);
#pragma unroll
    for (int j = 0; j < 3; ++j) o32[j] = b[4 * j] | (b[4 * j + 1] << 8) | (b[4 * j + 2] << 16) | (b[4 * j + 3] << 24);
}
int launch_post_crop(Device& dev, const __half* tile, int tile_w, int crop_x0, int crop_y0, int w, int h,
                     uint8_t* frame, int64_t stride, int dst_x0, int dst_y0) {
    if (w <= 0 || h <= 0) return 0;
    if (dev.blend_fast && w % 4 == 0 && dst_x0 % 4 == 0 && stride % 4 == 0 && (reinterpret_cast<uintptr_t>(frame) & 3) == 0 &&
        h < 65536) {
        const int w4 = w / 4;
        post_crop4_kernel<<<dim3((w4 + 127) / 128, h), 128, 0, dev.stream>>>(tile, tile_w, crop_x0, crop_y0, w4, h, frame, stride,
                                                                             dst_x0, dst_y0);
        VR_LAUNCH_CHECK(dev);
        return 0;
    }
    const int n = w * h;
    post_crop_kernel<<<(n + 255) / 256, 256, 0, dev.stream>>>(tile, tile_w, crop_x0, crop_y0, w, h, frame, stride,
                                                              dst_x0, dst_y0);
    VR_LAUNCH_CHECK(dev);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// K2b post: Gaussian gather blend. One thread per output pixel; visits the covering tiles in row-major order.
// ------------------------------------------------------------------------------------------------
struct BlendTileDev {
    const __half* data;
    int pitch;        // pixels per tile row in memory (atlas pitch)
    const float* wx;  // g(u), u in [0, pw): precomputed once per tile layout by blend_weights_kernel
    const float* wy;  // g(v), v in [0, ph)
    int px0, py0, pw, ph;
};
__device__ __forceinline__ float blend_g(int u, int extent) {
    // oracle.realesrganer.blend_window: t = (u - (extent-1)/2) * (4/extent); g = max(exp(-0.5 t^2), 1e-3)
    const float c = __fmul_rn(static_cast<float>(extent - 1), 0.5f);
    const float inv_sigma = __fdiv_rn(4.0f, static_cast<float>(extent));
    const float t = __fmul_rn(__fsub_rn(static_cast<float>(u), c), inv_sigma);
    const float g = expf(__fmul_rn(__fmul_rn(-0.5f, t), t));
    return fmaxf(g, 1e-3f);
}
__global__ void blend_weights_kernel(int extent, float* w) {
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u < extent) w[u] = blend_g(u, extent);
}
int launch_blend_weights(Device& dev, int extent, float* d_w) {
    blend_weights_kernel<<<(extent + 255) / 256, 256, 0, dev.stream>>>(extent, d_w);
    VR_LAUNCH_CHECK(dev);
    return 0;
}
// PX output pixels per thread (4 when rows are 4-byte aligned: three 32-bit stores; else 1)
template <int PX>
__global__ void __launch_bounds__(256)
post_blend_kernel(const BlendTileDev* __restrict__ tiles, int tiles_x, int tiles_y, int tile_out, int pad_out,
                  uint8_t* __restrict__ frame, int64_t stride, int sH, int sW) {
    const int X0 = (blockIdx.x * blockDim.x + threadIdx.x) * PX;
    const int Y = blockIdx.y * blockDim.y + threadIdx.y;
    if (X0 >= sW || Y >= sH) return;
    int iy0 = (Y - pad_out) / tile_out; if (Y - pad_out < 0) iy0 = 0;
    int iy1 = (Y + pad_out) / tile_out; if (iy1 > tiles_y - 1) iy1 = tiles_y - 1;
    uint8_t outb[PX * 3];
#pragma unroll
    for (int k = 0; k < PX; ++k) {
        const int X = X0 + k;
        int ix0 = (X - pad_out) / tile_out; if (X - pad_out < 0) ix0 = 0;
        int ix1 = (X + pad_out) / tile_out; if (ix1 > tiles_x - 1) ix1 = tiles_x - 1;
        float acc_r = 0.f, acc_g = 0.f, acc_b = 0.f, wsum = 0.f;
        for (int ty = iy0; ty <= iy1; ++ty)
            for (int tx = ix0; tx <= ix1; ++tx) {
                const BlendTileDev t = tiles[ty * tiles_x + tx];
                const int u = X - t.px0, v = Y - t.py0;
                if (u < 0 || u >= t.pw || v < 0 || v >= t.ph) continue;
                const float w = __fmul_rn(__ldg(t.wy + v), __ldg(t.wx + u));
                const uint2 q = __ldg(reinterpret_cast<const uint2*>(t.data + (static_cast<size_t>(v) * t.pitch + u) * 4));
                const __half2 rg = *reinterpret_cast<const __half2*>(&q.x);
                const __half2 b_ = *reinterpret_cast<const __half2*>(&q.y);
                acc_r = __fadd_rn(acc_r, __fmul_rn(__low2float(rg), w));
                acc_g = __fadd_rn(acc_g, __fmul_rn(__high2float(rg), w));
                acc_b = __fadd_rn(acc_b, __fmul_rn(__low2float(b_), w));
                wsum = __fadd_rn(wsum, w);
            }
        outb[3 * k + 0] = quant_u8(__fdiv_rn(acc_b, wsum));
        outb[3 * k + 1] = quant_u8(__fdiv_rn(acc_g, wsum));
        outb[3 * k + 2] = quant_u8(__fdiv_rn(acc_r, wsum));
    }
    uint8_t* o = frame + Y * stride + static_cast<int64_t>(X0) * 3;
    if (PX == 4) {
        uint32_t* o32 = reinterpret_cast<uint32_t*>(o);
#pragma unroll
        for (int j = 0; j < 3; ++j)
            o32[j] = outb[4 * j] | (outb[4 * j + 1] << 8) | (outb[4 * j + 2] << 16) |
                     (static_cast<uint32_t>(outb[4 * j + 3]) << 24);
    } else {
        o[0] = outb[0];
        o[1] = outb[1];
        o[2] = outb[2];
    }
}
// The aligned path (rows 4-byte aligned, sW % 4 == 0, coordinates < 65536): 4 consecutive output pixels per thread with the
// per-tile work hoisted out of the pixel loop -- candidate tile range once per thread (division by `tile_out` as one
// multiply-high with a host-side reciprocal), one descriptor and one row weight per tile, four independent pixel loads in
// flight. Per pixel the accumulation order (tiles in row-major order) and every rounding are those of post_blend_kernel, so
// both are bit-exact against the oracle (`oracle.realesrganer` tile_process, gather form). ncu, 5120 x 2880: the general
// kernel executes 201 instructions per pixel (two integer divisions, a 48-byte descriptor and three IEEE divisions per pixel).
__device__ __forceinline__ int div_magic(int x, uint32_t magic) { return static_cast<int>(__umulhi(static_cast<uint32_t>(x), magic)); }
__global__ void __launch_bounds__(256)
post_blend4_kernel(const BlendTileDev* __restrict__ tiles, int tiles_x, int tiles_y, uint32_t tile_magic, int pad_out,
                   uint8_t* __restrict__ frame, int64_t stride, int sH, int sW) {
    const int X0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int Y = blockIdx.y * blockDim.y + threadIdx.y;
    if (X0 >= sW || Y >= sH) return;
    const int iy0 = Y >= pad_out ? div_magic(Y - pad_out, tile_magic) : 0;
    const int iy1 = min(div_magic(Y + pad_out, tile_magic), tiles_y - 1);
    const int ix0 = X0 >= pad_out ? div_magic(X0 - pad_out, tile_magic) : 0;
    const int ix1 = min(div_magic(X0 + 3 + pad_out, tile_magic), tiles_x - 1);
    float acc_r[4] = {0.f, 0.f, 0.f, 0.f}, acc_g[4] = {0.f, 0.f, 0.f, 0.f}, acc_b[4] = {0.f, 0.f, 0.f, 0.f};
    float wsum[4] = {0.f, 0.f, 0.f, 0.f};
    for (int ty = iy0; ty <= iy1; ++ty)
        for (int tx = ix0; tx <= ix1; ++tx) {
            const BlendTileDev* tp = tiles + ty * tiles_x + tx;
            const int py0 = __ldg(&tp->py0), ph = __ldg(&tp->ph);
            const int v = Y - py0;
            if (v < 0 || v >= ph) continue;
            const int px0 = __ldg(&tp->px0), pw = __ldg(&tp->pw);
            const int u0 = X0 - px0;
            if (u0 + 3 < 0 || u0 >= pw) continue;
            const float wyv = __ldg(tp->wy + v);
            const float* wxp = tp->wx;
            const __half* rowp = tp->data + (static_cast<int64_t>(v) * __ldg(&tp->pitch) + u0) * 4;
            if (u0 >= 0 && u0 + 3 < pw) {  // all four pixels inside the tile: loads first, then the arithmetic
                uint2 q[4];
                float wx[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    q[k] = __ldg(reinterpret_cast<const uint2*>(rowp + k * 4));
                    wx[k] = __ldg(wxp + u0 + k);
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float w = __fmul_rn(wyv, wx[k]);
                    const __half2 rg = *reinterpret_cast<const __half2*>(&q[k].x);
                    const __half2 b_ = *reinterpret_cast<const __half2*>(&q[k].y);
                    acc_r[k] = __fadd_rn(acc_r[k], __fmul_rn(__low2float(rg), w));
                    acc_g[k] = __fadd_rn(acc_g[k], __fmul_rn(__high2float(rg), w));
                    acc_b[k] = __fadd_rn(acc_b[k], __fmul_rn(__low2float(b_), w));
                    wsum[k] = __fadd_rn(wsum[k], w);
                }
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int u = u0 + k;
                    if (u < 0 || u >= pw) continue;
                    const float w = __fmul_rn(wyv, __ldg(wxp + u));
                    const uint2 q = __ldg(reinterpret_cast<const uint2*>(rowp + k * 4));
                    const __half2 rg = *reinterpret_cast<const __half2*>(&q.x);
                    const __half2 b_ = *reinterpret_cast<const __half2*>(&q.y);
                    acc_r[k] = __fadd_rn(acc_r[k], __fmul_rn(__low2float(rg), w));
                    acc_g[k] = __fadd_rn(acc_g[k], __fmul_rn(__high2float(rg), w));
                    acc_b[k] = __fadd_rn(acc_b[k], __fmul_rn(__low2float(b_), w));
                    wsum[k] = __fadd_rn(wsum[k], w);
                }
            }
        }
    uint32_t outb[12];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        outb[3 * k + 0] = quant_u8(__fdiv_rn(acc_b[k], wsum[k]));
        outb[3 * k + 1] = quant_u8(__fdiv_rn(acc_g[k], wsum[k]));
        outb[3 * k + 2] = quant_u8(__fdiv_rn(acc_r[k], wsum[k]));
    }
    uint32_t* o32 = reinterpret_cast<uint32_t*>(frame + Y * stride + static_cast<int64_t>(X0) * 3);
#pragma unroll
    for (int j = 0; j < 3; ++j) o32[j] = outb[4 * j] | (outb[4 * j + 1] << 8) | (outb[4 * j + 2] << 16) | (outb[4 * j + 3] << 24);
}
// floor(x / d) == umulhi(x, magic) for 0 <= x < 65536 when magic = floor(2^32 / d) + 1 and 1 <= d < 65536:
// x * (magic * d - 2^32) <= x * d < 2^32
static bool blend_magic(int tile_out, int sH, int sW, int pad_out, uint32_t* magic) {
    if (tile_out < 2 || tile_out >= 65536 || sH + pad_out >= 65536 || sW + pad_out + 3 >= 65536) return false;
    *magic = static_cast<uint32_t>((1ull << 32) / static_cast<uint32_t>(tile_out)) + 1u;
    return true;
}
int launch_post_blend(Device& dev, const std::vector<BlendTile>& tiles, int tiles_x, int tiles_y, int tile_out,
                      int pad_out, uint8_t* frame, int64_t stride, int sH, int sW, BlendState& st) {
    // device tables (tile descriptors + 1-D weight windows) are rebuilt only when the tile layout changes
    bool same = st.last.size() == tiles.size();
    for (size_t i = 0; same && i < tiles.size(); ++i)
        same = st.last[i].data == tiles[i].data && st.last[i].px0 == tiles[i].px0 && st.last[i].py0 == tiles[i].py0 &&
               st.last[i].pw == tiles[i].pw && st.last[i].ph == tiles[i].ph && st.last[i].pitch == tiles[i].pitch;
    if (!same) {
        size_t nw = 0;
        for (const BlendTile& t : tiles) nw += static_cast<size_t>(t.pw) + t.ph;
        VR_CUDA_CHECK(cudaStreamSynchronize(dev.stream), dev.err);  // previous users of the tables
        if (st.weights_cap < nw) {
            if (st.d_weights) cudaFree(st.d_weights);
            VR_CUDA_CHECK(cudaMalloc(&st.d_weights, nw * sizeof(float)), dev.err);
            st.weights_cap = nw;
        }
        const size_t tb = tiles.size() * sizeof(BlendTileDev);
        if (st.table_cap < tb) {
            if (st.d_table) cudaFree(st.d_table);
            VR_CUDA_CHECK(cudaMalloc(&st.d_table, tb), dev.err);
            st.table_cap = tb;
        }
        st.host_table.resize(tb);
        BlendTileDev* host = reinterpret_cast<BlendTileDev*>(st.host_table.data());
        size_t off = 0;
        for (size_t i = 0; i < tiles.size(); ++i) {
            float* wx = st.d_weights + off;
            float* wy = wx + tiles[i].pw;
            off += static_cast<size_t>(tiles[i].pw) + tiles[i].ph;
            host[i] = {tiles[i].data, tiles[i].pitch, wx, wy, tiles[i].px0, tiles[i].py0, tiles[i].pw, tiles[i].ph};
            int rc = launch_blend_weights(dev, tiles[i].pw, wx);
            if (rc == 0) rc = launch_blend_weights(dev, tiles[i].ph, wy);
            if (rc) return rc;
        }
        VR_CUDA_CHECK(cudaMemcpyAsync(st.d_table, st.host_table.data(), tb, cudaMemcpyHostToDevice, dev.stream),
                      dev.err);  // host_table lives in the handle: no sync needed
        st.last = tiles;
    }
    const BlendTileDev* tab = static_cast<const BlendTileDev*>(st.d_table);
    const bool vec = sW % 4 == 0 && stride % 4 == 0 && (reinterpret_cast<uintptr_t>(frame) & 3) == 0;
    uint32_t magic = 0;
    if (vec && dev.blend_fast && blend_magic(tile_out, sH, sW, pad_out, &magic)) {
        dim3 block(64, 4);
        dim3 grid((sW / 4 + 63) / 64, (sH + 3) / 4);
        post_blend4_kernel<<<grid, block, 0, dev.stream>>>(tab, tiles_x, tiles_y, magic, pad_out, frame, stride, sH, sW);
    } else if (vec) {
        dim3 block(64, 4);
        dim3 grid((sW / 4 + 63) / 64, (sH + 3) / 4);
        post_blend_kernel<4><<<grid, block, 0, dev.stream>>>(tab, tiles_x, tiles_y, tile_out, pad_out, frame, stride, sH, sW);
    } else {
        dim3 block(32, 8);
        dim3 grid((sW + 31) / 32, (sH + 7) / 8);
        post_blend_kernel<1><<<grid, block, 0, dev.stream>>>(tab, tiles_x, tiles_y, tile_out, pad_out, frame, stride, sH, sW);
    }
    VR_LAUNCH_CHECK(dev);
    return 0;
}
void free_blend_state(BlendState& st) {
    if (st.d_weights) cudaFree(st.d_weights);
    if (st.d_table) cudaFree(st.d_table);
    st = BlendState();
}

// ------------------------------------------------------------------------------------------------
// K0b bilateral (radius <= 3): 32x8 output pixels per block, haloed u8 tile + colour LUT staged in smem
// ------------------------------------------------------------------------------------------------
struct BilateralParams {
    int radius, maxk;
    int8_t dy[49], dx[49];
    float space_w[49];
};
// Tables live in GLOBAL memory owned by the Device (one per handle), not in __constant__ symbols: constants are per-device
// globals shared by every handle and stream, so two restorers on one GPU with different sigma / d would overwrite each
// other's tables under a running kernel, and refreshing them needed a host sync inside the per-frame enqueue.
struct BilateralTables {
    BilateralParams bp;
    float color_w[768];
};

constexpr int kBilBW = 32, kBilBH = 8, kBilMaxR = 3;
__global__ void __launch_bounds__(kBilBW* kBilBH)
bilateral_kernel(const uint8_t* __restrict__ src, int64_t sstride, int H, int W, uint8_t* __restrict__ dst,
                 int64_t dstride, const BilateralTables* __restrict__ tab) {
    __shared__ uint8_t tile[kBilBH + 2 * kBilMaxR][(kBilBW + 2 * kBilMaxR) * 3 + 2];
    __shared__ float s_color[768];
    __shared__ BilateralParams c_bil;
    const int tid = threadIdx.y * kBilBW + threadIdx.x;
    for (int i = tid; i < 768; i += kBilBW * kBilBH) s_color[i] = __ldg(tab->color_w + i);
    for (int i = tid; i < static_cast<int>(sizeof(BilateralParams) / 4); i += kBilBW * kBilBH)
        reinterpret_cast<uint32_t*>(&c_bil)[i] = __ldg(reinterpret_cast<const uint32_t*>(&tab->bp) + i);
    __syncthreads();
    const int R = c_bil.radius;
    const int bx0 = blockIdx.x * kBilBW - R, by0 = blockIdx.y * kBilBH - R;
    const int tw = kBilBW + 2 * R, th = kBilBH + 2 * R;
    for (int i = tid; i < tw * th; i += kBilBW * kBilBH) {
        const int ty = i / tw, tx = i % tw;
        const int sy = reflect101(by0 + ty, H), sx = reflect101(bx0 + tx, W);
        const uint8_t* p = src + sy * sstride + static_cast<int64_t>(sx) * 3;
        tile[ty][tx * 3 + 0] = p[0];
        tile[ty][tx * 3 + 1] = p[1];
        tile[ty][tx * 3 + 2] = p[2];
    }
    __syncthreads();
    const int x = blockIdx.x * kBilBW + threadIdx.x, y = blockIdx.y * kBilBH + threadIdx.y;
    if (x >= W || y >= H) return;
    const int cx = threadIdx.x + R, cy = threadIdx.y + R;
    const int b0 = tile[cy][cx * 3], g0 = tile[cy][cx * 3 + 1], r0 = tile[cy][cx * 3 + 2];
    float sb = 0.f, sg = 0.f, sr = 0.f, ws = 0.f;
    for (int k = 0; k < c_bil.maxk; ++k) {
        const int ny = cy + c_bil.dy[k], nx = cx + c_bil.dx[k];
        const int b = tile[ny][nx * 3], g = tile[ny][nx * 3 + 1], r = tile[ny][nx * 3 + 2];
        const float w = __fmul_rn(c_bil.space_w[k], s_color[abs(b - b0) + abs(g - g0) + abs(r - r0)]);
        sb = __fadd_rn(sb, __fmul_rn(static_cast<float>(b), w));
        sg = __fadd_rn(sg, __fmul_rn(static_cast<float>(g), w));
        sr = __fadd_rn(sr, __fmul_rn(static_cast<float>(r), w));
        ws = __fadd_rn(ws, w);
    }
    const float inv = __fdiv_rn(1.0f, ws);
    uint8_t* o = dst + y * dstride + static_cast<int64_t>(x) * 3;
    o[0] = static_cast<uint8_t>(min(max(__float2int_rn(__fmul_rn(sb, inv)), 0), 255));
    o[1] = static_cast<uint8_t>(min(max(__float2int_rn(__fmul_rn(sg, inv)), 0), 255));
    o[2] = static_cast<uint8_t>(min(max(__float2int_rn(__fmul_rn(sr, inv)), 0), 255));
}
int launch_bilateral(Device& dev, const uint8_t* src, int64_t sstride, int H, int W, uint8_t* dst, int64_t dstride,
                     int d, float sigma_color, float sigma_space) {
    // tables exactly as OpenCV builds them (double exp, rounded to float): bilateralFilter_8u
    if (sigma_color <= 0) sigma_color = 1;
    if (sigma_space <= 0) sigma_space = 1;
    const double gcc = -0.5 / (static_cast<double>(sigma_color) * sigma_color);
    const double gsc = -0.5 / (static_cast<double>(sigma_space) * sigma_space);
    int radius = d <= 0 ? static_cast<int>(std::lrint(sigma_space * 1.5)) : d / 2;
    if (radius < 1) radius = 1;
    if (radius > kBilMaxR) {
        set_error(dev.err, "bilateral: radius > 3 (d > 7) not supported");
        return -1;
    }
    if (!dev.bil_tab || dev.bil_d != d || dev.bil_sc != sigma_color || dev.bil_ss != sigma_space) {
    dev.bil_host.resize(sizeof(BilateralTables));
    BilateralTables& bt = *reinterpret_cast<BilateralTables*>(dev.bil_host.data());
    BilateralParams& bp = bt.bp;
    std::memset(&bp, 0, sizeof(bp));
    bp.radius = radius;
    int maxk = 0;
    for (int i = -radius; i <= radius; ++i)
        for (int j = -radius; j <= radius; ++j) {
            const double r = std::sqrt(static_cast<double>(i) * i + static_cast<double>(j) * j);
            if (r > radius) continue;
            bp.space_w[maxk] = static_cast<float>(std::exp(r * r * gsc));
            bp.dy[maxk] = static_cast<int8_t>(i);
            bp.dx[maxk] = static_cast<int8_t>(j);
            ++maxk;
        }
    bp.maxk = maxk;
    for (int i = 0; i < 768; ++i) bt.color_w[i] = static_cast<float>(std::exp(static_cast<double>(i) * i * gcc));
    // uploaded only when (d, sigma_color, sigma_space) change; stream-ordered behind the kernels still reading the old tables.
    // The staging copy lives in the Device, but a second change may follow at once: wait for the previous upload first.
    if (!dev.bil_tab) VR_CUDA_CHECK(cudaMalloc(&dev.bil_tab, sizeof(BilateralTables)), dev.err);
    VR_CUDA_CHECK(cudaMemcpyAsync(dev.bil_tab, dev.bil_host.data(), sizeof(BilateralTables), cudaMemcpyHostToDevice, dev.stream),
                  dev.err);
    VR_CUDA_CHECK(cudaStreamSynchronize(dev.stream), dev.err);  // once per parameter change, not per frame
    dev.bil_d = d;
    dev.bil_sc = sigma_color;
    dev.bil_ss = sigma_space;
    }
    dim3 block(kBilBW, kBilBH), grid((W + kBilBW - 1) / kBilBW, (H + kBilBH - 1) / kBilBH);
    bilateral_kernel<<<grid, block, 0, dev.stream>>>(src, sstride, H, W, dst, dstride,
                                                     static_cast<const BilateralTables*>(dev.bil_tab));
    VR_LAUNCH_CHECK(dev);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// K3 unsharp: separable 7-tap Gaussian (sigma 1) fused in one pass through a haloed smem tile.
// Block = 64 x 16 output pixels x 3 channels; horizontal pass into smem floats, vertical pass from smem.
// ------------------------------------------------------------------------------------------------
constexpr int kUsW = 64, kUsH = 16, kUsR = 3;
__constant__ float c_taps7[7];
__global__ void __launch_bounds__(256)
unsharp_kernel(const uint8_t* __restrict__ src, int64_t sstride, int H, int W, uint8_t* __restrict__ dst,
               int64_t dstride, float amount) {
    __shared__ uint8_t s_in[kUsH + 2 * kUsR][(kUsW + 2 * kUsR) * 3 + 2];
    __shared__ float s_h[kUsH + 2 * kUsR][kUsW * 3];
    const int tid = threadIdx.x;
    const int bx0 = blockIdx.x * kUsW, by0 = blockIdx.y * kUsH;
    constexpr int tw = kUsW + 2 * kUsR, th = kUsH + 2 * kUsR;
    for (int i = tid; i < tw * th; i += 256) {
        const int ty = i / tw, tx = i % tw;
        const int sy = reflect101(by0 - kUsR + ty, H), sx = reflect101(bx0 - kUsR + tx, W);
        const uint8_t* p = src + sy * sstride + static_cast<int64_t>(sx) * 3;
        s_in[ty][tx * 3 + 0] = p[0];
        s_in[ty][tx * 3 + 1] = p[1];
        s_in[ty][tx * 3 + 2] = p[2];
    }
    __syncthreads();
    for (int i = tid; i < th * kUsW * 3; i += 256) {
        const int ty = i / (kUsW * 3), e = i % (kUsW * 3);  // e = x*3 + c
        float h = __fmul_rn(c_taps7[0], static_cast<float>(s_in[ty][e]));
#pragma unroll
        for (int t = 1; t < 7; ++t) h = __fadd_rn(h, __fmul_rn(c_taps7[t], static_cast<float>(s_in[ty][e + t * 3])));
        s_h[ty][e] = h;
    }
    __syncthreads();
    const float one_plus = __fadd_rn(1.0f, amount);
    for (int i = tid; i < kUsH * kUsW * 3; i += 256) {
        const int ty = i / (kUsW * 3), e = i % (kUsW * 3);
        const int x = bx0 + e / 3, y = by0 + ty;
        if (x >= W || y >= H) continue;
        float v = __fmul_rn(c_taps7[0], s_h[ty][e]);
#pragma unroll
        for (int t = 1; t < 7; ++t) v = __fadd_rn(v, __fmul_rn(c_taps7[t], s_h[ty + t][e]));
        const float xin = static_cast<float>(s_in[ty + kUsR][e + kUsR * 3]);
        const float o = __fsub_rn(__fmul_rn(one_plus, xin), __fmul_rn(amount, v));
        dst[y * dstride + static_cast<int64_t>(bx0) * 3 + e] =
            static_cast<uint8_t>(min(max(__float2int_rn(o), 0), 255));
    }
}
int launch_unsharp(Device& dev, const uint8_t* src, int64_t sstride, int H, int W, uint8_t* dst, int64_t dstride,
                   float amount) {
    int vrc = 0;
    if (try_unsharp_vec(dev, src, sstride, H, W, dst, dstride, amount, &vrc)) return vrc;
    static bool taps_done[64] = {};
    if (!taps_done[dev.ordinal & 63]) {
        double k[7], sum = 0;
        for (int i = 0; i < 7; ++i) {
            k[i] = std::exp(-0.5 * (i - 3) * (i - 3));
            sum += k[i];
        }
        float kf[7];
        for (int i = 0; i < 7; ++i) kf[i] = static_cast<float>(k[i] / sum);
        VR_CUDA_CHECK(cudaMemcpyToSymbol(c_taps7, kf, sizeof(kf)), dev.err);
        taps_done[dev.ordinal & 63] = true;
    }
    dim3 grid((W + kUsW - 1) / kUsW, (H + kUsH - 1) / kUsH);
    unsharp_kernel<<<grid, 256, 0, dev.stream>>>(src, sstride, H, W, dst, dstride, amount);
    VR_LAUNCH_CHECK(dev);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// K3 CLAHE on luma (OpenCV fixed-point YCrCb, clahe.cpp semantics)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int luma_of(int b, int g, int r) { return (r * 4899 + g * 9617 + b * 1868 + 8192) >> 14; }

// pass 1: per-tile 256-bin histogram of Y over the REFLECT_101-extended image. grid = (chunks, tiles).
__global__ void __launch_bounds__(256)
clahe_hist_kernel(const uint8_t* __restrict__ src, int64_t sstride, int H, int W, int tile_w, int tile_h,
                  int tiles_x, int32_t* __restrict__ hist) {
    __shared__ unsigned int s_hist[256];
    s_hist[threadIdx.x] = 0;
    __syncthreads();
    const int tile = blockIdx.y;
    const int tx0 = (tile % tiles_x) * tile_w, ty0 = (tile / tiles_x) * tile_h;
    const int area = tile_w * tile_h;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < area; i += gridDim.x * blockDim.x) {
        const int yy = reflect101(ty0 + i / tile_w, H), xx = reflect101(tx0 + i % tile_w, W);
        const uint8_t* p = src + yy * sstride + static_cast<int64_t>(xx) * 3;
        atomicAdd(&s_hist[luma_of(p[0], p[1], p[2])], 1u);
    }
    __syncthreads();
    if (s_hist[threadIdx.x]) atomicAdd(&hist[tile * 256 + threadIdx.x], static_cast<int>(s_hist[threadIdx.x]));
}
// pass 2: clip, redistribute, prefix sum, LUT. One 256-thread block per tile. Integer work: bit-exact.
// Both reductions are warp-shuffle based: the clipped excess is summed with __shfl_down_sync (one shared-memory word per warp
// instead of up to 256 atomics on one), the 256-bin cumulative histogram is an intra-warp __shfl_up_sync scan plus the eight
// warp totals (two block barriers instead of the sixteen of a shared-memory Hillis-Steele scan).
__global__ void __launch_bounds__(256)
clahe_lut_kernel(int32_t* __restrict__ hist, uint8_t* __restrict__ lut, int clip, float lut_scale) {
    __shared__ int s_wsum[8];   // per-warp clipped excess
    __shared__ int s_wtot[8];   // per-warp histogram totals (scan)
    const int t = threadIdx.x, tile = blockIdx.x, lane = t & 31, warp = t >> 5;
    int h = hist[tile * 256 + t];
    if (clip > 0) {
        int ex = h > clip ? h - clip : 0;
        if (h > clip) h = clip;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) ex += __shfl_down_sync(0xffffffffu, ex, off);
        if (lane == 0) s_wsum[warp] = ex;
        __syncthreads();
        int clipped = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) clipped += s_wsum[w];
        const int batch = clipped / 256;
        const int residual = clipped - batch * 256;
        h += batch;
        if (residual != 0) {
            const int step = max(256 / residual, 1);
            if (t % step == 0 && t / step < residual) h += 1;
        }
    }
    hist[tile * 256 + t] = h;
    int c = h;  // inclusive scan inside the warp
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, c, off);
        if (lane >= off) c += v;
    }
    if (lane == 31) s_wtot[warp] = c;
    __syncthreads();
    for (int w = 0; w < warp; ++w) c += s_wtot[w];
    const int q = __float2int_rn(__fmul_rn(static_cast<float>(c), lut_scale));
    lut[tile * 256 + t] = static_cast<uint8_t>(min(max(q, 0), 255));
}
// pass 3: BGR -> YCrCb, bilinear LUT interpolation on Y, -> BGR
__global__ void __launch_bounds__(256)
clahe_apply_kernel(const uint8_t* __restrict__ src, int64_t sstride, int H, int W, uint8_t* __restrict__ dst,
                   int64_t dstride, const uint8_t* __restrict__ lut, int tiles_x, int tiles_y, float inv_tw,
                   float inv_th) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= W) return;
    const uint8_t* p = src + y * sstride + static_cast<int64_t>(x) * 3;
    const int b = p[0], g = p[1], r = p[2];
    const int Y = luma_of(b, g, r);
    const int cr = min(max(((r - Y) * 11682 + (128 << 14) + 8192) >> 14, 0), 255);
    const int cb = min(max(((b - Y) * 9241 + (128 << 14) + 8192) >> 14, 0), 255);
    const float txf = __fsub_rn(__fmul_rn(static_cast<float>(x), inv_tw), 0.5f);
    const float tyf = __fsub_rn(__fmul_rn(static_cast<float>(y), inv_th), 0.5f);
    int tx1 = static_cast<int>(floorf(txf)), ty1 = static_cast<int>(floorf(tyf));
    const float xa = __fsub_rn(txf, static_cast<float>(tx1)), ya = __fsub_rn(tyf, static_cast<float>(ty1));
    const float xa1 = __fsub_rn(1.0f, xa), ya1 = __fsub_rn(1.0f, ya);
    const int tx2 = min(tx1 + 1, tiles_x - 1), ty2 = min(ty1 + 1, tiles_y - 1);
    tx1 = max(tx1, 0);
    ty1 = max(ty1, 0);
    const float l11 = static_cast<float>(__ldg(lut + (ty1 * tiles_x + tx1) * 256 + Y));
    const float l12 = static_cast<float>(__ldg(lut + (ty1 * tiles_x + tx2) * 256 + Y));
    const float l21 = static_cast<float>(__ldg(lut + (ty2 * tiles_x + tx1) * 256 + Y));
    const float l22 = static_cast<float>(__ldg(lut + (ty2 * tiles_x + tx2) * 256 + Y));
    const float top = __fadd_rn(__fmul_rn(l11, xa1), __fmul_rn(l12, xa));
    const float bot = __fadd_rn(__fmul_rn(l21, xa1), __fmul_rn(l22, xa));
    const float res = __fadd_rn(__fmul_rn(top, ya1), __fmul_rn(bot, ya));
    const int Yn = min(max(__float2int_rn(res), 0), 255);
    const int crd = cr - 128, cbd = cb - 128;
    const int nb = Yn + ((cbd * 29049 + 8192) >> 14);
    const int ng = Yn + ((cbd * -5636 + crd * -11698 + 8192) >> 14);
    const int nr = Yn + ((crd * 22987 + 8192) >> 14);
    uint8_t* o = dst + y * dstride + static_cast<int64_t>(x) * 3;
    o[0] = static_cast<uint8_t>(min(max(nb, 0), 255));
    o[1] = static_cast<uint8_t>(min(max(ng, 0), 255));
    o[2] = static_cast<uint8_t>(min(max(nr, 0), 255));
}
int launch_clahe(Device& dev, const uint8_t* src, int64_t sstride, int H, int W, uint8_t* dst, int64_t dstride,
                 float clip_limit, int grid_n, int32_t* d_hist, uint8_t* d_lut, uint8_t* /*d_luma*/, const TemporalFuse* tf,
                 bool* fused) {
    if (fused) *fused = false;
    if (grid_n < 1 || grid_n > 16) {
        set_error(dev.err, "clahe: grid must be 1..16");
        return -1;
    }
    const int tiles_x = grid_n, tiles_y = grid_n;
    int We = W, He = H;
    if (W % tiles_x != 0 || H % tiles_y != 0) {  // OpenCV pads BOTH whenever either is ragged
        We = W + (tiles_x - W % tiles_x);
        He = H + (tiles_y - H % tiles_y);
    }
    const int tile_w = We / tiles_x, tile_h = He / tiles_y;
    const int area = tile_w * tile_h;
    int clip = 0;
    if (clip_limit > 0.0f) {
        clip = static_cast<int>(static_cast<double>(clip_limit) * area / 256);
        if (clip < 1) clip = 1;
    }
    const float lut_scale = 255.0f / static_cast<float>(area);
    const int ntiles = tiles_x * tiles_y;
    VR_CUDA_CHECK(cudaMemsetAsync(d_hist, 0, ntiles * 256 * sizeof(int32_t), dev.stream), dev.err);
    const bool vec = clahe_vec_ok(src, sstride, H, W, dst, dstride, grid_n);
    if (vec) {
        int rc = launch_clahe_hist_vec(dev, src, sstride, tile_w, tile_h, tiles_x, ntiles, d_hist);
        if (rc) return rc;
    } else {
        int chunks = (area + 256 * 16 - 1) / (256 * 16);
        const int max_chunks = (dev.sm_count * 8 + ntiles - 1) / ntiles;
        if (chunks > max_chunks) chunks = max_chunks;
        if (chunks < 1) chunks = 1;
        clahe_hist_kernel<<<dim3(chunks, ntiles), 256, 0, dev.stream>>>(src, sstride, H, W, tile_w, tile_h, tiles_x,
                                                                        d_hist);
        VR_LAUNCH_CHECK(dev);
    }
    clahe_lut_kernel<<<ntiles, 256, 0, dev.stream>>>(d_hist, d_lut, clip, lut_scale);
    VR_LAUNCH_CHECK(dev);
    const float inv_tw = 1.0f / static_cast<float>(tile_w), inv_th = 1.0f / static_cast<float>(tile_h);
    if (vec) {
        const bool fuse = tf && tf->prev && tf->blended && (reinterpret_cast<uintptr_t>(tf->prev) & 15) == 0 &&
                          (reinterpret_cast<uintptr_t>(tf->blended) & 15) == 0;
        if (fused) *fused = fuse;
        return launch_clahe_apply_vec(dev, src, dst, H, W, d_lut, tiles_x, tiles_y, inv_tw, inv_th, fuse ? tf : nullptr);
    }
    clahe_apply_kernel<<<dim3((W + 255) / 256, H), 256, 0, dev.stream>>>(src, sstride, H, W, dst, dstride, d_lut,
                                                                         tiles_x, tiles_y, inv_tw, inv_th);
    VR_LAUNCH_CHECK(dev);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// K3 temporal: out = gate ? rint(0.8*cur + 0.2*prev) : cur, gate = max_c |cur-prev| < tau (per pixel)
// ------------------------------------------------------------------------------------------------
__global__ void temporal_kernel(const uint8_t* __restrict__ cur, int64_t cstride, const uint8_t* __restrict__ prev,
                                int64_t pstride, int H, int W, uint8_t* __restrict__ dst, int64_t dstride,
                                float alpha, float one_minus, float tau) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= W) return;
    const uint8_t* c = cur + y * cstride + static_cast<int64_t>(x) * 3;
    const uint8_t* p = prev + y * pstride + static_cast<int64_t>(x) * 3;
    uint8_t* o = dst + y * dstride + static_cast<int64_t>(x) * 3;
    const int c0 = c[0], c1 = c[1], c2 = c[2], p0 = p[0], p1 = p[1], p2 = p[2];
    const int d = max(max(abs(c0 - p0), abs(c1 - p1)), abs(c2 - p2));
    if (static_cast<float>(d) < tau) {
        o[0] = static_cast<uint8_t>(min(max(__float2int_rn(__fadd_rn(__fmul_rn(one_minus, static_cast<float>(c0)),
                                                                     __fmul_rn(alpha, static_cast<float>(p0)))), 0), 255));
        o[1] = static_cast<uint8_t>(min(max(__float2int_rn(__fadd_rn(__fmul_rn(one_minus, static_cast<float>(c1)),
                                                                     __fmul_rn(alpha, static_cast<float>(p1)))), 0), 255));
        o[2] = static_cast<uint8_t>(min(max(__float2int_rn(__fadd_rn(__fmul_rn(one_minus, static_cast<float>(c2)),
                                                                     __fmul_rn(alpha, static_cast<float>(p2)))), 0), 255));
    } else {
        o[0] = c[0];
        o[1] = c[1];
        o[2] = c[2];
    }
}
int launch_temporal(Device& dev, const uint8_t* cur, int64_t cstride, const uint8_t* prev, int64_t pstride, int H,
                    int W, uint8_t* dst, int64_t dstride, float alpha, float tau) {
    int vrc = 0;
    if (try_temporal_vec(dev, cur, cstride, prev, pstride, H, W, dst, dstride, alpha, tau, &vrc)) return vrc;
    const float one_minus = 1.0f - alpha;
    temporal_kernel<<<dim3((W + 255) / 256, H), 256, 0, dev.stream>>>(cur, cstride, prev, pstride, H, W, dst, dstride,
                                                                      alpha, one_minus, tau);
    VR_LAUNCH_CHECK(dev);
    return 0;
}

}  // namespace vr
