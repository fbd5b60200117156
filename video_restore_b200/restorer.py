"""Host-side mirror of the reference's per-frame interface, over the C ABI of libvrb200.so.

`RealESRGANer` keeps the constructor and `enhance(img, outscale) -> (uint8 BGR, 'RGB')` contract the reference
uses (video_upscaler.py:328-338, :501); `FrameRestorer` is `_process_frame` (video_upscaler.py:490-505) plus the
README-only enhancement stage. All arithmetic happens in hand-written sm_100a kernels; this file only moves
pointers. There is no CPU fallback: construction raises VrError without a B200 and the built library.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import VrConfig, VrError, VrFrameOpts
from .models import MODEL_ZOO


@dataclass
class FrameOpts:
    """Per-frame enhancement switches; defaults == plain `enhance`."""
    denoise: bool = False          # bilateral 5/25/25 on the LR frame, video_upscaler.py:495-496
    denoise_d: int = 5
    denoise_sigma_color: float = 25.0
    denoise_sigma_space: float = 25.0
    sharpen: float = 0.0           # unsharp amount, README.md:141
    clahe: bool = False            # README.md:11,240
    clahe_clip: float = 2.0
    clahe_grid: int = 8
    temporal: bool = False         # README.md:9,237
    temporal_alpha: float = 0.2
    temporal_tau: float = 12.0

    def to_c(self) -> VrFrameOpts:
        return VrFrameOpts(denoise=int(self.denoise), denoise_d=self.denoise_d,
                           denoise_sigma_color=self.denoise_sigma_color, denoise_sigma_space=self.denoise_sigma_space,
                           sharpen=self.sharpen, clahe=int(self.clahe), clahe_clip=self.clahe_clip,
                           clahe_grid=self.clahe_grid, temporal=int(self.temporal),
                           temporal_alpha=self.temporal_alpha, temporal_tau=self.temporal_tau)


def _as_spec(model):
    if isinstance(model, str):
        if model not in MODEL_ZOO:
            raise ValueError(f"Unsupported model: {model}")  # same message as video_upscaler.py:323
        return dict(MODEL_ZOO[model])
    if isinstance(model, dict):
        return dict(model)
    raise TypeError("model must be a model-zoo name or a spec dict(kind, scale, num_block, num_conv)")


class FrameRestorer:
    """One instance per (GPU, host thread), like `self.models[gpu_id]` in the reference (video_upscaler.py:340)."""

    def __init__(self, model="RealESRGAN_x4plus", state_dict=None, tile=512, tile_pad=10, pre_pad=0,
                 blend="crop", gpu_id=0):
        self._h = C.c_void_p()
        self._lib = _lib.load()
        spec = _as_spec(model)
        self.spec = spec
        self.scale = spec["scale"]
        cfg = VrConfig(model_kind=_lib.VR_MODEL_RRDBNET if spec["kind"] == "rrdb" else _lib.VR_MODEL_SRVGG,
                       scale=spec["scale"], num_block=spec.get("num_block", 0), num_conv=spec.get("num_conv", 0),
                       num_feat=64, num_grow_ch=32, tile=int(tile), tile_pad=int(tile_pad), pre_pad=int(pre_pad),
                       blend=_lib.VR_BLEND_GAUSSIAN if blend == "gaussian" else _lib.VR_BLEND_CROP,
                       device=int(gpu_id))
        rc = self._lib.vr_create(C.byref(cfg), C.byref(self._h))
        if rc != 0:
            self._h = C.c_void_p()
            raise VrError(f"vr_create failed ({rc}): {(self._lib.vr_last_error(None) or b'').decode()}")
        if state_dict is not None:
            self.load_state_dict(state_dict)

    # -- weights ---------------------------------------------------------------------------------
    def load_state_dict(self, state_dict) -> None:
        """`state_dict`: mapping upstream key -> array-like float (torch tensors or numpy), as in a .pth's
        params_ema / params."""
        for name, t in state_dict.items():
            a = np.ascontiguousarray(t.detach().cpu().numpy() if hasattr(t, "detach") else t, dtype=np.float32)
            shape = (C.c_int64 * a.ndim)(*a.shape)
            _lib.check(self._lib.vr_load_tensor(self._h, name.encode(), a.ctypes.data_as(C.c_void_p), shape, a.ndim),
                       self._h)
        _lib.check(self._lib.vr_commit_weights(self._h), self._h)

    # -- the hot path ----------------------------------------------------------------------------
    def process_frame(self, frame: np.ndarray, opts: FrameOpts | None = None, out: np.ndarray | None = None):
        """uint8 [H,W,3] BGR (read-only is fine) -> freshly allocated uint8 [sH,sW,3] BGR."""
        if frame.dtype != np.uint8 or frame.ndim != 3 or frame.shape[2] != 3:
            raise ValueError("frame must be uint8 HxWx3 BGR")
        if frame.strides[2] != 1 or frame.strides[1] != 3:
            frame = np.ascontiguousarray(frame)
        H, W, _ = frame.shape
        s = self.scale
        if out is None:
            out = np.empty((H * s, W * s, 3), np.uint8)
        elif (not isinstance(out, np.ndarray) or out.dtype != np.uint8 or out.shape != (H * s, W * s, 3)
              or out.strides[1:] != (3, 1) or out.strides[0] < W * s * 3 or not out.flags.writeable):
            raise ValueError(f"out must be a writable uint8 array of shape {(H * s, W * s, 3)} with dense pixels")
        o = (opts or FrameOpts()).to_c()
        rc = self._lib.vr_restore(self._h, frame.ctypes.data_as(C.c_void_p), H, W, frame.strides[0],
                                  out.ctypes.data_as(C.c_void_p), out.strides[0], C.byref(o))
        _lib.check(rc, self._h)
        return out

    zero_copy_stream = True  # process_stream accepts out_pool (pipeline.py hands frames to the encoder without a copy)

    def process_stream(self, frames, opts: FrameOpts | None = None, out_pool=None):
        """Generator over restored frames, in order, with H2D / compute / D2H of neighbouring frames overlapped
        (vr_submit / vr_wait, two frames in flight). `frames` is any iterable of uint8 HxWx3 BGR arrays. Each yielded
        array is a view of a pinned buffer that is overwritten three frames later: copy it to keep it longer -- or
        pass `out_pool` (an object with get(shape) -> page-locked uint8 array, may block): every frame is then written
        into a buffer of the caller's, who owns it from the moment it is yielded."""
        o = (opts or FrameOpts()).to_c()
        n_slots = 3
        ring_in, ring_out = getattr(self, "_rings", ([], []))  # pinned rings are kept across calls (slow to allocate)
        i = 0
        pending = []  # (ticket, out_array)
        for frame in frames:
            if frame.dtype != np.uint8 or frame.ndim != 3 or frame.shape[2] != 3:
                raise ValueError("frame must be uint8 HxWx3 BGR")
            H, W, _ = frame.shape
            s = self.scale
            if not ring_in or ring_in[0].shape != frame.shape or (out_pool is None and not ring_out):
                while pending:
                    t, out = pending.pop(0)
                    _lib.check(self._lib.vr_wait(self._h, t), self._h)
                    yield out
                ring_in = [_lib.pinned_array((H, W, 3)) for _ in range(n_slots)]
                ring_out = [] if out_pool is not None else [_lib.pinned_array((H * s, W * s, 3)) for _ in range(n_slots)]
                self._rings = (ring_in, ring_out)
            slot = i % n_slots
            if len(pending) >= 2:  # keep two in flight: the slot about to be reused has been handed out already
                t, out = pending.pop(0)
                _lib.check(self._lib.vr_wait(self._h, t), self._h)
                yield out
            np.copyto(ring_in[slot], frame)
            dst = out_pool.get((H * s, W * s, 3)) if out_pool is not None else ring_out[slot]
            ticket = C.c_int64(0)
            _lib.check(self._lib.vr_submit(self._h, ring_in[slot].ctypes.data_as(C.c_void_p), H, W,
                                           ring_in[slot].strides[0], dst.ctypes.data_as(C.c_void_p),
                                           dst.strides[0], C.byref(o), C.byref(ticket)), self._h)
            pending.append((int(ticket.value), dst))
            i += 1
        while pending:
            t, out = pending.pop(0)
            _lib.check(self._lib.vr_wait(self._h, t), self._h)
            yield out

    def process_frame_device(self, d_in: int, H: int, W: int, d_out: int, opts: FrameOpts | None = None,
                             sync: bool = True) -> None:
        """Device-resident frames (raw pointers on this restorer's GPU), dense rows."""
        o = (opts or FrameOpts()).to_c()
        fn = self._lib.vr_restore_device if sync else self._lib.vr_restore_device_async
        _lib.check(fn(self._h, C.c_void_p(d_in), H, W, W * 3, C.c_void_p(d_out), W * self.scale * 3, C.byref(o)),
                   self._h)

    def sync(self) -> None:
        _lib.check(self._lib.vr_sync(self._h), self._h)

    @staticmethod
    def alloc_host(shape) -> np.ndarray:
        """uint8 array over page-locked memory (fast H2D / D2H staging, e.g. boundary frames in pipeline.py)."""
        return _lib.pinned_array(tuple(shape))

    @property
    def stream(self) -> int:
        return int(self._lib.vr_stream(self._h) or 0)

    # -- temporal state (frame-range sharding hands the boundary frame across) ----------------------
    def temporal_reset(self) -> None:
        _lib.check(self._lib.vr_temporal_reset(self._h), self._h)

    def temporal_set_prev(self, up_prev, device_ptr: bool = False, shape=None) -> None:
        if device_ptr:
            sH, sW = shape
            _lib.check(self._lib.vr_temporal_set_prev(self._h, C.c_void_p(int(up_prev)), sH, sW, sW * 3, 1), self._h)
        else:
            a = np.ascontiguousarray(up_prev, dtype=np.uint8)
            _lib.check(self._lib.vr_temporal_set_prev(self._h, a.ctypes.data_as(C.c_void_p), a.shape[0], a.shape[1],
                                                      a.strides[0], 0), self._h)

    def temporal_get_prev(self, sH: int, sW: int, device_ptr: int | None = None, out: np.ndarray | None = None):
        if device_ptr is not None:
            _lib.check(self._lib.vr_temporal_get_prev(self._h, C.c_void_p(int(device_ptr)), sH, sW, sW * 3, 1), self._h)
            return None
        a = out if out is not None else np.empty((sH, sW, 3), np.uint8)
        if a.dtype != np.uint8 or a.shape != (sH, sW, 3) or a.strides[1:] != (3, 1) or not a.flags.writeable:
            raise ValueError(f"out must be a writable uint8 array of shape {(sH, sW, 3)} with dense pixels")
        _lib.check(self._lib.vr_temporal_get_prev(self._h, a.ctypes.data_as(C.c_void_p), sH, sW, a.strides[0], 0),
                   self._h)
        return a

    def temporal_blend_device(self, d_cur: int, d_prev: int, sH: int, sW: int, d_out: int, alpha: float = 0.2,
                              tau: float = 12.0) -> None:
        """Temporal blend of two device frames on this restorer's stream (asynchronous; `sync()` to wait)."""
        _lib.check(self._lib.vr_temporal_device(self._h, C.c_void_p(int(d_cur)), C.c_void_p(int(d_prev)), sH, sW,
                                                C.c_void_p(int(d_out)), alpha, tau), self._h)

    # -- boundary frame between two restorers of one process (pipeline.py) ---------------------------
    peer_boundary = True

    def boundary_send(self, dst: "FrameRestorer", sH: int, sW: int) -> int:
        """This restorer's last un-blended frame -> a device buffer on `dst`'s GPU (one cudaMemcpyPeerAsync); returns the
        device pointer, to be handed to dst.boundary_finish by dst's thread."""
        p = C.c_void_p()
        _lib.check(self._lib.vr_boundary_send(self._h, dst._h, sH, sW, C.byref(p)), self._h)
        return int(p.value)

    def boundary_finish(self, d_prev: int, head: np.ndarray | None, out: np.ndarray | None, alpha: float = 0.2,
                        tau: float = 12.0) -> None:
        """Blend the host head frame with the received boundary frame into `out` (host); recycles the device buffer."""
        if head is None:
            _lib.check(self._lib.vr_boundary_finish(self._h, C.c_void_p(d_prev), None, 0, None, 0, 0, 0, alpha, tau), self._h)
            return
        if head.dtype != np.uint8 or out.dtype != np.uint8 or head.shape != out.shape or head.strides[1:] != (3, 1) \
                or out.strides[1:] != (3, 1):
            raise ValueError("boundary_finish: head / out must be uint8 HxWx3 with dense rows")
        _lib.check(self._lib.vr_boundary_finish(self._h, C.c_void_p(d_prev), head.ctypes.data_as(C.c_void_p), head.strides[0],
                                                out.ctypes.data_as(C.c_void_p), out.strides[0], head.shape[0], head.shape[1],
                                                alpha, tau), self._h)

    # -- introspection ---------------------------------------------------------------------------
    @property
    def launch_count(self) -> int:
        return int(self._lib.vr_launch_count(self._h))

    def debug_activation(self, which: str) -> np.ndarray:
        """Test hook: fp32 [Ha, Wa, 64] intermediate tensor ("feat" / "body" / "trunk") of the last restored frame."""
        ha, wa, c = C.c_int32(0), C.c_int32(0), C.c_int32(0)
        _lib.check(self._lib.vr_debug_activation(self._h, which.encode(), None, 0, C.byref(ha), C.byref(wa), C.byref(c)),
                   self._h)
        out = np.empty((ha.value, wa.value, c.value), np.float32)
        _lib.check(self._lib.vr_debug_activation(self._h, which.encode(), out.ctypes.data_as(C.c_void_p), out.size,
                                                 C.byref(ha), C.byref(wa), C.byref(c)), self._h)
        return out

    @property
    def conv_launch_count(self) -> int:
        return int(self._lib.vr_conv_launch_count(self._h))

    def last_timing(self):
        t, c = C.c_float(0), C.c_float(0)
        _lib.check(self._lib.vr_last_timing(self._h, C.byref(t), C.byref(c)), self._h)
        return float(t.value), float(c.value)

    def last_timing_frames(self) -> int:
        """Number of frames `last_timing()` averages over (frames enqueued since the previous sync)."""
        return int(self._lib.vr_last_timing_frames(self._h))

    def close(self) -> None:
        if getattr(self, "_h", None) and self._h.value:
            self._lib.vr_destroy(self._h)
            self._h = C.c_void_p()
        self._rings = ([], [])  # the pinned frame rings go back now, not whenever the collector finds them

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001 - interpreter shutdown
            pass


class RealESRGANer:
    """Signature-compatible stand-in for realesrgan.RealESRGANer as the reference constructs it
    (video_upscaler.py:328-338). `model` is a model-zoo name or spec instead of an nn.Module; `model_path` may be a
    .pth (params_ema / params) or None together with `state_dict`."""

    def __init__(self, scale, model_path=None, dni_weight=None, model="RealESRGAN_x4plus", tile=0, tile_pad=10,
                 pre_pad=10, half=True, device=None, gpu_id=None, state_dict=None, blend="crop"):
        spec = _as_spec(model)
        if scale != spec["scale"]:
            raise ValueError(f"scale={scale} does not match the model's network scale {spec['scale']}")
        if tile <= 0:
            # upstream runs the whole frame as one tile when tile == 0; same arithmetic as one tile covering it
            tile = 1 << 20
        if not half:
            raise VrError("the B200 path computes in fp16 with fp32 accumulation; half=False has no CUDA-free fallback")
        if gpu_id is None:
            gpu_id = 0
            if device is not None:
                idx = getattr(device, "index", None)
                if idx is None and isinstance(device, str) and ":" in device:
                    idx = int(device.split(":")[1])
                gpu_id = idx or 0
        if state_dict is None:
            if model_path is None:
                raise ValueError("either model_path or state_dict is required")
            import torch
            loadnet = torch.load(model_path, map_location="cpu")
            state_dict = loadnet["params_ema"] if "params_ema" in loadnet else loadnet.get("params", loadnet)
        self.scale = scale
        self.tile_size = tile
        self.tile_pad = tile_pad
        self.pre_pad = pre_pad
        self.half = half
        if pre_pad < 0:
            raise ValueError("pre_pad must be >= 0")
        # pre_pad (upstream default 10; the reference passes 0, video_upscaler.py:334) is a reflect pad of the bottom / right
        # edge before everything else and a crop afterwards: done here on the host frame, the C ABI always sees pre_pad = 0
        self._r = FrameRestorer(model=spec, state_dict=state_dict, tile=tile, tile_pad=tile_pad, pre_pad=0,
                                blend=blend, gpu_id=gpu_id)

    def enhance(self, img, outscale=None, alpha_upsampler="realesrgan"):
        """uint8 HxWx3 BGR -> (uint8 BGR, 'RGB') like upstream. The reference always passes outscale == scale
        (video_upscaler.py:501,718); any other outscale is upstream's final step, a Lanczos resize of the network-scale
        result on the host with the very same OpenCV call."""
        h0, w0 = img.shape[:2]
        if self.pre_pad:
            if self.pre_pad >= min(h0, w0):
                raise ValueError("pre_pad must be smaller than the frame (reflect padding)")
            img = np.pad(img, ((0, self.pre_pad), (0, self.pre_pad), (0, 0)), mode="reflect")
        out = self._r.process_frame(img)
        if self.pre_pad:
            out = np.ascontiguousarray(out[:h0 * self.scale, :w0 * self.scale])
        if outscale is not None and float(outscale) != float(self.scale):
            import cv2

            out = cv2.resize(out, (int(w0 * outscale), int(h0 * outscale)), interpolation=cv2.INTER_LANCZOS4)
        return out, "RGB"


# -- stand-alone filters (host arrays) ------------------------------------------------------------
def _u8(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    if a.ndim != 3 or a.shape[2] != 3:
        raise ValueError("expected uint8 HxWx3")
    return a


def bilateral_filter(img, d=5, sigma_color=25.0, sigma_space=25.0, device=0):
    a = _u8(img); out = np.empty_like(a)
    _lib.check(_lib.load().vr_bilateral(device, a.ctypes.data_as(C.c_void_p), a.shape[0], a.shape[1],
                                        out.ctypes.data_as(C.c_void_p), d, sigma_color, sigma_space))
    return out


def unsharp_mask(img, amount, device=0):
    a = _u8(img); out = np.empty_like(a)
    _lib.check(_lib.load().vr_unsharp(device, a.ctypes.data_as(C.c_void_p), a.shape[0], a.shape[1],
                                      out.ctypes.data_as(C.c_void_p), amount))
    return out


def clahe_bgr(img, clip=2.0, grid=8, device=0, return_tables=False):
    a = _u8(img); out = np.empty_like(a)
    hist = np.zeros((grid * grid, 256), np.int32)
    lut = np.zeros((grid * grid, 256), np.uint8)
    _lib.check(_lib.load().vr_clahe(device, a.ctypes.data_as(C.c_void_p), a.shape[0], a.shape[1],
                                    out.ctypes.data_as(C.c_void_p), clip, grid, hist.ctypes.data_as(C.c_void_p),
                                    lut.ctypes.data_as(C.c_void_p)))
    return (out, hist, lut) if return_tables else out


def temporal_blend(cur, prev, alpha=0.2, tau=12.0, device=0):
    a = _u8(cur); p = _u8(prev); out = np.empty_like(a)
    _lib.check(_lib.load().vr_temporal(device, a.ctypes.data_as(C.c_void_p), p.ctypes.data_as(C.c_void_p), a.shape[0],
                                       a.shape[1], out.ctypes.data_as(C.c_void_p), alpha, tau))
    return out


def blend_weights(extent, device=0):
    w = np.zeros(extent, np.float32)
    _lib.check(_lib.load().vr_blend_weights(device, extent, w.ctypes.data_as(C.c_void_p)))
    return w


def tile_grid(H, W, tile, tile_pad, scale):
    lib = _lib.load()
    n = lib.vr_tile_grid(H, W, tile, tile_pad, scale, None, 0)
    if n < 0:
        raise VrError(f"vr_tile_grid: invalid arguments ({n})")
    tab = np.zeros((n, 12), np.int32)
    lib.vr_tile_grid(H, W, tile, tile_pad, scale, tab.ctypes.data_as(C.c_void_p), n)
    return tab
