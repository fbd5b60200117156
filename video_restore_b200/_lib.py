"""ctypes binding of libvrb200.so (include/vrb200.h). Loading fails loudly: there is no CPU fallback."""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

_PKG = Path(__file__).resolve().parent
# VR_LIB: another build of the same library (A/B runs of kernel changes on one box); default the in-tree build
LIB_PATH = Path(os.environ["VR_LIB"]) if os.environ.get("VR_LIB") else _PKG / "libvrb200.so"

VR_MODEL_RRDBNET, VR_MODEL_SRVGG = 0, 1
VR_BLEND_CROP, VR_BLEND_GAUSSIAN = 0, 1


class VrError(RuntimeError):
    """Raised for any non-zero status from the C ABI (the reference raises Python exceptions per frame,
    video_upscaler.py:476-481)."""


class VrConfig(C.Structure):
    _fields_ = [
        ("model_kind", C.c_int32), ("scale", C.c_int32), ("num_block", C.c_int32), ("num_conv", C.c_int32),
        ("num_feat", C.c_int32), ("num_grow_ch", C.c_int32), ("tile", C.c_int32), ("tile_pad", C.c_int32),
        ("pre_pad", C.c_int32), ("blend", C.c_int32), ("device", C.c_int32), ("reserved", C.c_int32 * 5),
    ]


class VrFrameOpts(C.Structure):
    _fields_ = [
        ("denoise", C.c_int32), ("denoise_d", C.c_int32), ("denoise_sigma_color", C.c_float),
        ("denoise_sigma_space", C.c_float), ("sharpen", C.c_float), ("clahe", C.c_int32),
        ("clahe_clip", C.c_float), ("clahe_grid", C.c_int32), ("temporal", C.c_int32),
        ("temporal_alpha", C.c_float), ("temporal_tau", C.c_float), ("reserved", C.c_int32 * 5),
    ]


class VrConvTest(C.Structure):
    _fields_ = [
        ("H", C.c_int32), ("W", C.c_int32), ("cin", C.c_int32), ("cout", C.c_int32),
        ("x", C.c_void_p), ("weight", C.c_void_p), ("bias", C.c_void_p),
        ("act", C.c_int32), ("slope", C.c_float), ("prelu", C.c_void_p),
        ("res1", C.c_void_p), ("s1", C.c_float), ("res2", C.c_void_p), ("s2", C.c_float),
        ("y", C.c_void_p), ("rows", C.c_int32), ("flags", C.c_int32),
        ("iters", C.c_int32), ("ms", C.c_float), ("device", C.c_int32),
    ]


# name -> (restype, argtypes); also the list the CPU test checks against the header's declarations
SIGNATURES = {
    "vr_create": (C.c_int, [C.POINTER(VrConfig), C.POINTER(C.c_void_p)]),
    "vr_destroy": (None, [C.c_void_p]),
    "vr_last_error": (C.c_char_p, [C.c_void_p]),
    "vr_load_tensor": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.POINTER(C.c_int64), C.c_int32]),
    "vr_commit_weights": (C.c_int, [C.c_void_p]),
    "vr_restore": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_void_p, C.c_int64,
                             C.POINTER(VrFrameOpts)]),
    "vr_restore_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_void_p, C.c_int64,
                                    C.POINTER(VrFrameOpts)]),
    "vr_restore_device_async": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_void_p,
                                          C.c_int64, C.POINTER(VrFrameOpts)]),
    "vr_submit": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_void_p, C.c_int64,
                            C.POINTER(VrFrameOpts), C.POINTER(C.c_int64)]),
    "vr_wait": (C.c_int, [C.c_void_p, C.c_int64]),
    "vr_host_alloc": (C.c_void_p, [C.c_size_t]),
    "vr_host_free": (None, [C.c_void_p]),
    "vr_sync": (C.c_int, [C.c_void_p]),
    "vr_stream": (C.c_void_p, [C.c_void_p]),
    "vr_temporal_reset": (C.c_int, [C.c_void_p]),
    "vr_temporal_set_prev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_int32]),
    "vr_temporal_get_prev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_int32]),
    "vr_tile_grid": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32]),
    "vr_bilateral": (C.c_int, [C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_float,
                               C.c_float]),
    "vr_unsharp": (C.c_int, [C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_float]),
    "vr_clahe": (C.c_int, [C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_float, C.c_int32,
                           C.c_void_p, C.c_void_p]),
    "vr_temporal": (C.c_int, [C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_float,
                              C.c_float]),
    "vr_blend_weights": (C.c_int, [C.c_int32, C.c_int32, C.c_void_p]),
    "vr_conv3x3_test": (C.c_int, [C.POINTER(VrConvTest)]),
    "vr_conv_pair2_test": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(C.c_float), C.c_void_p,
                                     C.c_int32, C.c_void_p, C.c_int32, C.c_int32]),
    "vr_pair2_profile": (None, [C.c_void_p]),
    "vr_global_error": (C.c_char_p, []),
    "vr_conv3x3_bench": (C.c_int, [C.c_int32] * 8 + [C.POINTER(C.c_float)]),
    "vr_last_conv_cycles": (C.c_int64, []),
    "vr_filter_bench": (C.c_int, [C.c_int32] * 5 + [C.POINTER(C.c_float)]),
    "vr_debug_activation": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64, C.POINTER(C.c_int32),
                                      C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "vr_launch_count": (C.c_int64, [C.c_void_p]),
    "vr_conv_launch_count": (C.c_int64, [C.c_void_p]),
    "vr_last_timing": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "vr_last_timing_frames": (C.c_int32, [C.c_void_p]),
    "vr_boundary_send": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]),
    "vr_boundary_finish": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int32, C.c_int32,
                                     C.c_float, C.c_float]),
    "vr_temporal_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_float,
                                     C.c_float]),
}

_lib = None


def load() -> C.CDLL:
    """Load the CUDA library; raise (never fall back) if it is missing or incomplete."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise VrError(f"{LIB_PATH} is missing: run `python -m video_restore_b200.build` "
                      "(the restoration path is CUDA-only; there is no CPU fallback)")
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is absent
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def global_error() -> str:
    return (load().vr_global_error() or b"").decode()


def check(rc: int, handle=None) -> None:
    if rc == 0:
        return
    lib = load()
    msg = lib.vr_last_error(handle) if handle else lib.vr_global_error()
    raise VrError(f"libvrb200 status {rc}: {(msg or b'').decode()}")


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def conv3x3(x, weight, bias=None, act=0, slope=0.2, prelu=None, res1=None, s1=1.0, res2=None, s2=1.0,
            rows=0, flags=0, iters=1, device=0):
    """Kernel-level hook: x [H,W,Cin] f32, weight [Cout,Cin,3,3] -> y [H,W,Cout] f32 (cout==48: [4H,4W,3])."""
    lib = load()
    x = _f32(x); weight = _f32(weight); bias = _f32(bias); prelu = _f32(prelu); res1 = _f32(res1); res2 = _f32(res2)
    H, W, cin = x.shape
    cout = weight.shape[0]
    y = np.zeros((4 * H, 4 * W, 3) if cout == 48 else (H, W, cout), np.float32)
    ptr = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    t = VrConvTest(H=H, W=W, cin=cin, cout=cout, x=ptr(x), weight=ptr(weight), bias=ptr(bias), act=act,
                   slope=slope, prelu=ptr(prelu), res1=ptr(res1), s1=s1, res2=ptr(res2), s2=s2, y=ptr(y),
                   rows=rows, flags=flags, iters=iters, ms=0.0, device=device)
    check(lib.vr_conv3x3_test(C.byref(t)))
    return y, float(t.ms)


def conv_pair2(x, wa, ba, wb, bb, slope=0.2, iters=1, gaps_x=(), gaps_y=(), device=0, flags=0):
    """K4 hook: x [H,W,Cin] (Cin % 32 == 0); layer A Cin -> 32, layer B Cin + 32 -> 32 (input = concat(x, yA)), both with bias and
    LeakyReLU. Returns (yA, yB, ms)."""
    lib = load()
    x = _f32(x); wa = _f32(wa); ba = _f32(ba); wb = _f32(wb); bb = _f32(bb)
    H, W, cin = x.shape
    ya = np.zeros((H, W, 32), np.float32)
    yb = np.zeros((H, W, 32), np.float32)
    gx = np.asarray(gaps_x, np.int32)
    gy = np.asarray(gaps_y, np.int32)
    ms = C.c_float(0)
    ptr = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    check(lib.vr_conv_pair2_test(device, H, W, cin, ptr(x), ptr(wa), ptr(ba), ptr(wb), ptr(bb), slope, ptr(ya), ptr(yb), iters,
                                 C.byref(ms), ptr(gx) if gx.size else None, gx.size, ptr(gy) if gy.size else None, gy.size, flags))
    return ya, yb, float(ms.value)


def pair2_profile():
    """{name: cycles} of cluster 0's leader CTA in the last conv_pair2 launch."""
    a = np.zeros(64, np.int64)
    load().vr_pair2_profile(a.ctypes.data_as(C.c_void_p))
    names = {0: "producer.wait_empty", 1: "producer.total"}
    for i, n in enumerate(("wait_for_waiter", "issue", "-", "loop_total", "units", "prologue")):
        names[10 + i] = f"issuer.{n}"
    for i, n in enumerate(("wait_for_issuer", "barrier_waits", "walk", "loop_total", "units", "prologue")):
        names[18 + i] = f"waiter.{n}"
    for g in (0, 1):
        for i, n in enumerate(("wait_tfull", "wait_hempty", "rows", "total")):
            names[30 + g * 4 + i] = f"epilogue{g}.{n}"
    return {n: int(a[i]) for i, n in names.items()}


def conv3x3_bench(H, W, cin, cout, rows=0, flags=0, iters=20, device=0) -> float:
    lib = load()
    ms = C.c_float(0)
    check(lib.vr_conv3x3_bench(device, H, W, cin, cout, rows, flags, iters, C.byref(ms)))
    return float(ms.value)


def last_conv_cycles() -> int:
    return int(load().vr_last_conv_cycles())


FILTER_KINDS = {"bilateral": 0, "unsharp": 1, "clahe": 2, "temporal": 3, "post_crop": 4, "post_blend": 5, "pre": 6,
                "upsample2x": 7}
# algorithmic bytes per pixel of the HxW frame each kind is timed on
FILTER_BYTES_PER_PX = {"bilateral": 6, "unsharp": 6, "clahe": 9, "temporal": 9, "post_crop": 11, "post_blend": 11,
                       "pre": 67, "upsample2x": 128 + 512}


def filter_bench(kind: str, H: int, W: int, iters: int = 20, device: int = 0):
    """Returns (ms per call, achieved GB/s on the algorithmic bytes)."""
    ms = C.c_float(0)
    check(load().vr_filter_bench(device, FILTER_KINDS[kind], H, W, iters, C.byref(ms)))
    return float(ms.value), FILTER_BYTES_PER_PX[kind] * H * W / (ms.value * 1e-3) / 1e9


def pinned_array(shape, dtype=np.uint8):
    """numpy array over page-locked host memory (freed when the array's base buffer is garbage collected)."""
    lib = load()
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = lib.vr_host_alloc(n)
    if not p:
        raise VrError("vr_host_alloc failed")
    buf = (C.c_uint8 * n).from_address(p)
    arr = np.frombuffer(buf, dtype=dtype).reshape(shape)
    import weakref
    weakref.finalize(buf, lib.vr_host_free, C.c_void_p(p))
    return arr
