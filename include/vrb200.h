/*
 * vrb200.h -- C ABI of the B200-native per-frame restoration hot path (libvrb200.so).
 *
 * Drop-in boundary for ryanjcooper/video-restore (reference, read-only at /root/reference):
 *   - vr_create + vr_load_tensor + vr_commit_weights replace the constructor call
 *       RealESRGANer(scale, model_path, model, tile, tile_pad, pre_pad=0, half, gpu_id, device)
 *     at video_upscaler.py:328-338 and the architecture choice at video_upscaler.py:313-321.
 *   - vr_restore replaces the body of OptimizedVideoUpscaler._process_frame, video_upscaler.py:490-505:
 *       cv2.bilateralFilter(frame, 5, 25, 25)            (video_upscaler.py:496)
 *       upscaler.enhance(frame, outscale=scale)          (video_upscaler.py:501)
 *     plus the README-only enhancement stage (README.md:8-12,140-141,236-240: seamless Gaussian tile blend,
 *     unsharp mask, CLAHE, temporal consistency) whose definitions are pinned by oracle/ (SURVEY.md 8 A7-A10).
 *
 * Conventions: plain pointers and sizes, no C++/torch types. Every call returns 0 on success or a negative
 * VR_E_* code; vr_last_error() gives the message. Nothing throws across the ABI, there is no global state,
 * one handle per (device, host thread). There is NO CPU fallback: without a usable sm_100 device vr_create fails.
 */
#ifndef VRB200_H_
#define VRB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VR_OK 0
#define VR_E_INVALID (-1) /* bad argument / unsupported configuration */
#define VR_E_CUDA (-2)    /* CUDA runtime or driver error (message has the CUDA string) */
#define VR_E_STATE (-3)   /* call order (e.g. restore before commit_weights), missing tensor */
#define VR_E_NODEVICE (-4)

#define VR_MODEL_RRDBNET 0 /* basicsr RRDBNet: x4plus (23 blocks), x2plus (scale 2), anime_6B (6 blocks) */
#define VR_MODEL_SRVGG 1   /* realesrgan SRVGGNetCompact: x4_v3 (32 convs, PReLU) */

#define VR_BLEND_CROP 0     /* RealESRGANer.tile_process crop-merge (what the reference code does) */
#define VR_BLEND_GAUSSIAN 1 /* seamless Gaussian-weighted gather blend (README.md:8,236) */

typedef struct vr_handle vr_handle;

typedef struct vr_config {
    int32_t model_kind;  /* VR_MODEL_* */
    int32_t scale;       /* network scale: 4, or 2 (pixel-unshuffle front end) */
    int32_t num_block;   /* RRDBNet blocks (23 / 6) */
    int32_t num_conv;    /* SRVGG body convs (32) */
    int32_t num_feat;    /* 64 */
    int32_t num_grow_ch; /* 32 */
    int32_t tile;        /* RealESRGANer tile (>0 always in the reference, video_upscaler.py:332) */
    int32_t tile_pad;    /* tile_overlap if enhanced else 10, video_upscaler.py:326 */
    int32_t pre_pad;     /* 0 in the reference, video_upscaler.py:334; only 0 is supported */
    int32_t blend;       /* VR_BLEND_* */
    int32_t device;      /* CUDA ordinal */
    int32_t reserved[5];
} vr_config;

/* Per-frame enhancement switches (all off == the plain `enhance` call). */
typedef struct vr_frame_opts {
    int32_t denoise;        /* 1: bilateral pre-denoise on the LR frame (video_upscaler.py:495-496) */
    int32_t denoise_d;      /* 5 */
    float denoise_sigma_color; /* 25 */
    float denoise_sigma_space; /* 25 */
    float sharpen;          /* unsharp amount a (0 = off), README.md:141 */
    int32_t clahe;          /* 1: CLAHE on luma, README.md:11,240 */
    float clahe_clip;       /* 2.0 */
    int32_t clahe_grid;     /* 8 */
    int32_t temporal;       /* 1: temporal consistency against the previous upscaled frame, README.md:9,237 */
    float temporal_alpha;   /* 0.2 */
    float temporal_tau;     /* 12 */
    int32_t reserved[5];
} vr_frame_opts;

/* ---- lifetime ---- */
int vr_create(const vr_config* cfg, vr_handle** out);
void vr_destroy(vr_handle* h);
const char* vr_last_error(const vr_handle* h); /* h may be NULL: error of the last failed vr_create */

/* ---- weights: one call per state_dict tensor, upstream key names
 *      (RRDBNet: conv_first, body.{i}.rdb{1..3}.conv{1..5}, conv_body, conv_up1, conv_up2, conv_hr, conv_last
 *       + ".weight" OIHW / ".bias";  SRVGG: body.{2i}.weight/.bias convs, body.{2i+1}.weight PReLU) ---- */
int vr_load_tensor(vr_handle* h, const char* name, const float* data, const int64_t* shape, int32_t ndim);
int vr_commit_weights(vr_handle* h); /* repack to the kernels' layout; errors on a missing/ill-shaped tensor */

/* ---- the hot path ---- */
/* Host buffers: bgr = uint8[H][W][3] (read-only, row stride in bytes), out = uint8[sH][sW][3]. */
int vr_restore(vr_handle* h, const uint8_t* bgr, int32_t H, int32_t W, int64_t stride, uint8_t* out,
               int64_t out_stride, const vr_frame_opts* opts);
/* Device-resident frames (same layouts, device pointers on cfg.device); work is enqueued and synchronised. */
int vr_restore_device(vr_handle* h, const uint8_t* d_bgr, int32_t H, int32_t W, int64_t stride, uint8_t* d_out,
                      int64_t out_stride, const vr_frame_opts* opts);
/* As vr_restore_device but returns after enqueueing on the handle's stream (caller uses vr_sync). */
int vr_restore_device_async(vr_handle* h, const uint8_t* d_bgr, int32_t H, int32_t W, int64_t stride,
                            uint8_t* d_out, int64_t out_stride, const vr_frame_opts* opts);
int vr_sync(vr_handle* h);
void* vr_stream(vr_handle* h); /* cudaStream_t the handle enqueues on */

/* Pipelined host-buffer path: vr_submit enqueues H2D (copy stream) -> restore (compute stream) -> D2H (copy stream)
 * for one frame and returns at once with a ticket; up to two frames are in flight, so the copies of neighbouring
 * frames overlap the compute of the current one. bgr/out must stay valid (and should be pinned, see vr_host_alloc)
 * until vr_wait(ticket) returns. Frames are processed in submission order (temporal state included). */
int vr_submit(vr_handle* h, const uint8_t* bgr, int32_t H, int32_t W, int64_t stride, uint8_t* out,
              int64_t out_stride, const vr_frame_opts* opts, int64_t* ticket);
int vr_wait(vr_handle* h, int64_t ticket);
/* Page-locked host memory for the frame buffers of vr_restore / vr_submit. */
void* vr_host_alloc(size_t bytes);
void vr_host_free(void* p);

/* Temporal-consistency state: the previous frame's un-blended upscaled result ("up_{t-1}").
 * A frame-range shard seeds it with the boundary frame received from its left neighbour. */
int vr_temporal_reset(vr_handle* h);
int vr_temporal_set_prev(vr_handle* h, const uint8_t* up_prev, int32_t sH, int32_t sW, int64_t stride,
                         int32_t is_device);
int vr_temporal_get_prev(vr_handle* h, uint8_t* dst, int32_t sH, int32_t sW, int64_t stride, int32_t is_device);

/* Temporal blend of two DEVICE frames on the handle's stream (asynchronous): a frame-range shard finishes its head frame
 * with it once the left neighbour's last un-blended frame has arrived peer-to-peer. alpha / tau <= 0: defaults 0.2 / 12. */
int vr_temporal_device(vr_handle* h, const uint8_t* d_cur, const uint8_t* d_prev, int32_t sH, int32_t sW, uint8_t* d_out,
                       float alpha, float tau);

/* Boundary frame between two handles of ONE process (one host thread + handle per GPU; replaces the shared queue of
 * video_upscaler.py:430-488 for the temporal stage): vr_boundary_send copies src's temporal state (its last un-blended upscaled
 * frame) into a device buffer on dst's device with ONE cudaMemcpyPeerAsync (NVLink between peers) and returns it in *d_frame;
 * vr_boundary_finish (dst's thread) blends a host head frame with it into a host output and recycles the buffer (head == NULL:
 * recycle only). */
int vr_boundary_send(vr_handle* src, vr_handle* dst, int32_t sH, int32_t sW, void** d_frame);
int vr_boundary_finish(vr_handle* h, void* d_prev, const uint8_t* head, int64_t head_stride, uint8_t* out, int64_t out_stride,
                       int32_t sH, int32_t sW, float alpha, float tau);

/* ---- integer tile geometry (RealESRGANer.tile_process index arithmetic), bit-exact contract ----
 * Writes up to max_tiles rows of 12 int32:
 *   {in_x0,in_x1,in_y0,in_y1, pad_x0,pad_x1,pad_y0,pad_y1, out_x0,out_x1,out_y0,out_y1}
 * (input tile, padded input tile, and the crop inside the scaled padded output tile). Returns tile count. */
int vr_tile_grid(int32_t H, int32_t W, int32_t tile, int32_t tile_pad, int32_t scale, int32_t* table,
                 int32_t max_tiles);

/* ---- enhancement kernels on their own (host buffers, uint8 BGR HWC) ---- */
int vr_bilateral(int32_t device, const uint8_t* src, int32_t H, int32_t W, uint8_t* dst, int32_t d,
                 float sigma_color, float sigma_space);
int vr_unsharp(int32_t device, const uint8_t* src, int32_t H, int32_t W, uint8_t* dst, float amount);
/* hist_out (optional): int32[grid*grid][256] clipped+redistributed histograms; lut_out (optional): uint8[grid*grid][256] */
int vr_clahe(int32_t device, const uint8_t* src, int32_t H, int32_t W, uint8_t* dst, float clip, int32_t grid,
             int32_t* hist_out, uint8_t* lut_out);
int vr_temporal(int32_t device, const uint8_t* cur, const uint8_t* prev, int32_t H, int32_t W, uint8_t* dst,
                float alpha, float tau);
/* Gaussian tile-blend weights g(u), u in [0, extent): fp32, as the blend kernel evaluates them. */
int vr_blend_weights(int32_t device, int32_t extent, float* w_out);

/* ---- kernel-level test / bench hooks ---- */
typedef struct vr_conv_test {
    int32_t H, W, cin, cout;
    const float* x;      /* [H][W][cin] */
    const float* weight; /* [cout][cin][3][3] */
    const float* bias;   /* [cout] or NULL */
    int32_t act;         /* 0 none, 1 leaky-relu(slope), 2 PReLU(prelu[cout]) */
    float slope;
    const float* prelu;
    const float* res1; /* [H][W][cout] or NULL: y = y*s1 + res1 */
    float s1;
    const float* res2; /* then y = y*s2 + res2 */
    float s2;
    float* y;          /* [H][W][cout] */
    int32_t rows;      /* output rows per CTA tile (0 = default 4) */
    int32_t flags;     /* measurement ablations: 1 no A-collector reuse, 2 skip TMA, 4 skip MMA, 8 skip epilogue */
    int32_t iters;     /* >1: repeat the launch and report the average */
    float ms;          /* out: average kernel milliseconds (CUDA events) */
    int32_t device;
} vr_conv_test;
int vr_conv3x3_test(vr_conv_test* t);
const char* vr_global_error(void);

/* K4 test / bench hook: two consecutive 32-channel dense-block layers (cin -> 32, cin + 32 -> 32, bias + LeakyReLU) in ONE launch
 * over a chunk-planar buffer, x [H][W][cin] (cin % 32 == 0), ya / yb [H][W][32]; optional tile-atlas gap columns / rows;
 * flags: measurement ablations -- 4 skip MMA, 8 hand-off without data, 16 no global stores. */
int vr_conv_pair2_test(int32_t device, int32_t H, int32_t W, int32_t cin, const float* x, const float* wa, const float* ba,
                       const float* wb, const float* bb, float slope, float* ya, float* yb, int32_t iters, float* ms,
                       const int32_t* gaps_x, int32_t ngx, const int32_t* gaps_y, int32_t ngy, int32_t flags);

/* Wait-cycle profile of the last vr_conv_pair2_test launch (cluster 0's leader CTA): 64 counters, see conv3x3_pair2_sm100.cuh. */
void vr_pair2_profile(int64_t* out64);

/* Device-resident conv benchmark on zero-copy synthetic data: returns average ms per launch. */
int vr_conv3x3_bench(int32_t device, int32_t H, int32_t W, int32_t cin, int32_t cout, int32_t rows,
                     int32_t flags, int32_t iters, float* ms_out);

/* SM cycles (clock64) the slowest CTA spent inside the last vr_conv3x3_bench launch: cycles / time = real SM clock. */
int64_t vr_last_conv_cycles(void);

/* Device-resident timing of the HBM-bound kernels on a synthetic HxW uint8 frame (average ms per call).
 * kind: 0 bilateral, 1 unsharp, 2 CLAHE (hist+LUT+apply), 3 temporal, 4 post crop-merge, 5 post Gaussian blend (2x2
 * tiles), 6 pre (u8 -> fp16 NHWC32), 7 nearest x2 upsample (64 channels). */
int vr_filter_bench(int32_t device, int32_t kind, int32_t H, int32_t W, int32_t iters, float* ms_out);

/* Feature-level parity hook: an intermediate tensor of the LAST restored frame as fp32 [Ha][Wa][64] over the tile atlas
 * (single tile: the padded tile). RRDBNet: "feat" (conv_first), "body" (last RRDB's output), "trunk" (feat + conv_body(body));
 * SRVGG: "body" (input of the last conv). out may be NULL to query the extents. capacity in floats. */
int vr_debug_activation(vr_handle* h, const char* which, float* out, int64_t capacity, int32_t* Ha, int32_t* Wa,
                        int32_t* C);

/* Counters since handle creation: kernels launched by this library, for bench.py's gpu_launches. */
int64_t vr_launch_count(const vr_handle* h);
int64_t vr_conv_launch_count(const vr_handle* h); /* of which convolution kernels */
/* Device ms per frame of the whole chain / of the network (conv kernels), averaged over the frames enqueued since the
 * previous vr_sync (CUDA events on the handle's stream, one record per frame, the last 256 at most);
 * vr_last_timing_frames = how many frames that average covers. */
int vr_last_timing(const vr_handle* h, float* total_ms, float* conv_ms);
int32_t vr_last_timing_frames(const vr_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* VRB200_H_ */
