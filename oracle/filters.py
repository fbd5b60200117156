"""CPU restatements (numpy) of the enhancement-stage filters.

TEST INFRASTRUCTURE -- see oracle/__init__.py.

  bilateral_filter   restates cv2.bilateralFilter(frame, 5, 25, 25), the reference's only own arithmetic call on the
                     hot path (video_upscaler.py:496). PINNED bit-exactly against OpenCV 4.13 in tests.
  bgr_to_ycrcb / ycrcb_to_bgr / clahe_u8 / clahe_bgr
                     README-only "CLAHE colour enhancement" (README.md:11,240). Spec = OpenCV's own
                     cvtColor(BGR2YCrCb) + createCLAHE(2.0,(8,8)).apply + cvtColor(YCrCb2BGR); PINNED bit-exactly.
  unsharp_mask       README-only (README.md:12,141); float definition of SURVEY 8 A8 is the spec.
  temporal_blend     README-only (README.md:9,237); non-recursive definition of SURVEY 8 A10 is the spec.

All float arithmetic below is fp32 with one rounding per operation (numpy never fuses multiply-add); the CUDA
kernels use __fmul_rn/__fadd_rn in the same order where bit-exactness is claimed.
"""
from __future__ import annotations

import numpy as np

_F = np.float32


def _reflect101(idx: np.ndarray, n: int) -> np.ndarray:
    """cv2.BORDER_REFLECT_101 index map (gfedcb|abcdefgh|gfedcba)."""
    if n == 1:
        return np.zeros_like(idx)
    period = 2 * (n - 1)
    m = np.mod(idx, period)
    return np.where(m >= n, period - m, m)


# ------------------------------------------------------------------------------------------------
# bilateral (OpenCV bilateralFilter_8u, 3 channels)
# ------------------------------------------------------------------------------------------------
def bilateral_tables(d: int, sigma_color: float, sigma_space: float):
    """(radius, [(dy,dx)...], space_weight f32[maxk], color_weight f32[256*3]) exactly as OpenCV builds them."""
    if sigma_color <= 0:
        sigma_color = 1.0
    if sigma_space <= 0:
        sigma_space = 1.0
    gauss_color_coeff = -0.5 / (float(sigma_color) * float(sigma_color))
    gauss_space_coeff = -0.5 / (float(sigma_space) * float(sigma_space))
    radius = int(round(sigma_space * 1.5)) if d <= 0 else d // 2
    radius = max(radius, 1)
    color_weight = np.exp(np.arange(256 * 3, dtype=np.float64) ** 2 * gauss_color_coeff).astype(np.float32)
    offs, sw = [], []
    for i in range(-radius, radius + 1):
        for j in range(-radius, radius + 1):
            r = np.sqrt(float(i) * i + float(j) * j)
            if r > radius:
                continue
            sw.append(np.float32(np.exp(r * r * gauss_space_coeff)))
            offs.append((i, j))
    return radius, offs, np.asarray(sw, np.float32), color_weight


def bilateral_filter(src: np.ndarray, d: int = 5, sigma_color: float = 25.0, sigma_space: float = 25.0) -> np.ndarray:
    """uint8 [H,W,3] -> uint8 [H,W,3]; cv2.bilateralFilter with BORDER_DEFAULT (REFLECT_101)."""
    assert src.dtype == np.uint8 and src.ndim == 3 and src.shape[2] == 3
    H, W, _ = src.shape
    radius, offs, space_w, color_w = bilateral_tables(d, sigma_color, sigma_space)
    ys = _reflect101(np.arange(-radius, H + radius), H)
    xs = _reflect101(np.arange(-radius, W + radius), W)
    pad = src[ys][:, xs].astype(np.int32)
    c0 = pad[radius:radius + H, radius:radius + W]
    wsum = np.zeros((H, W), _F)
    acc = np.zeros((H, W, 3), _F)
    for k, (i, j) in enumerate(offs):
        nb = pad[radius + i:radius + i + H, radius + j:radius + j + W]
        dist = np.abs(nb - c0).sum(axis=2)
        w = (space_w[k] * color_w[dist]).astype(_F)
        acc += nb.astype(_F) * w[:, :, None]
        wsum += w
    inv = (_F(1.0) / wsum).astype(_F)
    out = np.rint(acc * inv[:, :, None])
    return np.clip(out, 0, 255).astype(np.uint8)


# ------------------------------------------------------------------------------------------------
# colour conversion, OpenCV 8-bit fixed point (yuv_shift = 14)
# ------------------------------------------------------------------------------------------------
def bgr_to_ycrcb(bgr: np.ndarray) -> np.ndarray:
    b = bgr[:, :, 0].astype(np.int32)
    g = bgr[:, :, 1].astype(np.int32)
    r = bgr[:, :, 2].astype(np.int32)
    half = 1 << 13
    delta = 128 << 14
    y = (r * 4899 + g * 9617 + b * 1868 + half) >> 14
    cr = ((r - y) * 11682 + delta + half) >> 14
    cb = ((b - y) * 9241 + delta + half) >> 14
    out = np.stack([y, cr, cb], axis=2)
    return np.clip(out, 0, 255).astype(np.uint8)


def ycrcb_to_bgr(ycc: np.ndarray) -> np.ndarray:
    y = ycc[:, :, 0].astype(np.int32)
    cr = ycc[:, :, 1].astype(np.int32) - 128
    cb = ycc[:, :, 2].astype(np.int32) - 128
    half = 1 << 13
    b = y + ((cb * 29049 + half) >> 14)
    g = y + ((cb * -5636 + cr * -11698 + half) >> 14)
    r = y + ((cr * 22987 + half) >> 14)
    out = np.stack([b, g, r], axis=2)
    return np.clip(out, 0, 255).astype(np.uint8)


# ------------------------------------------------------------------------------------------------
# CLAHE (OpenCV clahe.cpp, 8-bit)
# ------------------------------------------------------------------------------------------------
def clahe_tables(src: np.ndarray, clip_limit: float = 2.0, grid: int = 8):
    """Returns (hist int32[grid*grid,256] after clip+redistribute, lut uint8[grid*grid,256], tile_h, tile_w)."""
    assert src.dtype == np.uint8 and src.ndim == 2
    H, W = src.shape
    tiles_x = tiles_y = grid
    if W % tiles_x == 0 and H % tiles_y == 0:
        ext = src
    else:
        pad_b = tiles_y - (H % tiles_y)
        pad_r = tiles_x - (W % tiles_x)
        ys = _reflect101(np.arange(0, H + pad_b), H)
        xs = _reflect101(np.arange(0, W + pad_r), W)
        ext = src[ys][:, xs]
    tile_h, tile_w = ext.shape[0] // tiles_y, ext.shape[1] // tiles_x
    area = tile_h * tile_w
    lut_scale = _F(255) / _F(area)
    clip = 0
    if clip_limit > 0.0:
        clip = max(int(clip_limit * area / 256), 1)
    hists = np.zeros((tiles_y * tiles_x, 256), np.int32)
    luts = np.zeros((tiles_y * tiles_x, 256), np.uint8)
    for ty in range(tiles_y):
        for tx in range(tiles_x):
            t = ext[ty * tile_h:(ty + 1) * tile_h, tx * tile_w:(tx + 1) * tile_w]
            hist = np.bincount(t.ravel(), minlength=256).astype(np.int32)
            if clip > 0:
                over = hist > clip
                clipped = int((hist[over] - clip).sum())
                hist[over] = clip
                batch = clipped // 256
                residual = clipped - batch * 256
                hist += batch
                if residual != 0:
                    step = max(256 // residual, 1)
                    i = 0
                    while i < 256 and residual > 0:
                        hist[i] += 1
                        i += step
                        residual -= 1
            cs = np.cumsum(hist).astype(_F)
            lut = np.rint(cs * lut_scale)
            hists[ty * tiles_x + tx] = hist
            luts[ty * tiles_x + tx] = np.clip(lut, 0, 255).astype(np.uint8)
    return hists, luts, tile_h, tile_w


def clahe_u8(src: np.ndarray, clip_limit: float = 2.0, grid: int = 8) -> np.ndarray:
    """cv2.createCLAHE(clipLimit, (grid, grid)).apply(src) for uint8 single-channel input."""
    H, W = src.shape
    _, luts, tile_h, tile_w = clahe_tables(src, clip_limit, grid)
    inv_tw = _F(1.0) / _F(tile_w)
    inv_th = _F(1.0) / _F(tile_h)
    txf = (np.arange(W, dtype=_F) * inv_tw - _F(0.5)).astype(_F)
    tyf = (np.arange(H, dtype=_F) * inv_th - _F(0.5)).astype(_F)
    tx1 = np.floor(txf).astype(np.int32)
    ty1 = np.floor(tyf).astype(np.int32)
    xa = (txf - tx1.astype(_F)).astype(_F)
    ya = (tyf - ty1.astype(_F)).astype(_F)
    xa1 = (_F(1.0) - xa).astype(_F)
    ya1 = (_F(1.0) - ya).astype(_F)
    tx2 = np.minimum(tx1 + 1, grid - 1)
    ty2 = np.minimum(ty1 + 1, grid - 1)
    tx1 = np.maximum(tx1, 0)
    ty1 = np.maximum(ty1, 0)
    v = src.astype(np.int64)
    l11 = luts[(ty1[:, None] * grid + tx1[None, :]), v].astype(_F)
    l12 = luts[(ty1[:, None] * grid + tx2[None, :]), v].astype(_F)
    l21 = luts[(ty2[:, None] * grid + tx1[None, :]), v].astype(_F)
    l22 = luts[(ty2[:, None] * grid + tx2[None, :]), v].astype(_F)
    top = (l11 * xa1[None, :] + l12 * xa[None, :]).astype(_F)
    bot = (l21 * xa1[None, :] + l22 * xa[None, :]).astype(_F)
    res = (top * ya1[:, None] + bot * ya[:, None]).astype(_F)
    return np.clip(np.rint(res), 0, 255).astype(np.uint8)


def clahe_bgr(bgr: np.ndarray, clip_limit: float = 2.0, grid: int = 8) -> np.ndarray:
    """BGR -> YCrCb, CLAHE on Y, -> BGR."""
    ycc = bgr_to_ycrcb(bgr)
    ycc = ycc.copy()
    ycc[:, :, 0] = clahe_u8(np.ascontiguousarray(ycc[:, :, 0]), clip_limit, grid)
    return ycrcb_to_bgr(ycc)


# ------------------------------------------------------------------------------------------------
# unsharp mask (spec ours)
# ------------------------------------------------------------------------------------------------
def gaussian_taps7() -> np.ndarray:
    """7-tap sigma=1.0 kernel, normalised in float64 then rounded to fp32 (the CUDA kernel embeds these values)."""
    i = np.arange(-3, 4, dtype=np.float64)
    k = np.exp(-0.5 * i * i)
    return (k / k.sum()).astype(np.float32)


def unsharp_mask(src: np.ndarray, amount: float) -> np.ndarray:
    """out = sat_u8(rint((1+a)*x - a*blur)), blur = separable 7-tap Gaussian (rows then columns), REFLECT_101,
    fp32, taps accumulated left-to-right / top-to-bottom."""
    assert src.dtype == np.uint8 and src.ndim == 3
    H, W, _ = src.shape
    k = gaussian_taps7()
    x = src.astype(_F)
    xs = _reflect101(np.arange(-3, W + 3), W)
    ys = _reflect101(np.arange(-3, H + 3), H)
    xp = x[:, xs]
    hb = np.zeros_like(x)
    for t in range(7):
        hb += k[t] * xp[:, t:t + W]
    hp = hb[ys]
    vb = np.zeros_like(x)
    for t in range(7):
        vb += k[t] * hp[t:t + H]
    a = _F(amount)
    out = (_F(1.0) + a) * x - a * vb
    return np.clip(np.rint(out), 0, 255).astype(np.uint8)


# ------------------------------------------------------------------------------------------------
# temporal consistency (spec ours, non-recursive)
# ------------------------------------------------------------------------------------------------
def temporal_blend(cur: np.ndarray, prev: np.ndarray | None, alpha: float = 0.2, tau: float = 12.0) -> np.ndarray:
    """d = max_c |cur - prev|; w = alpha if d < tau else 0; out = sat_u8(rint((1-w)*cur + w*prev)).
    `prev` is the previous frame's UN-blended result; None (first frame) passes through."""
    if prev is None:
        return cur.copy()
    assert cur.shape == prev.shape and cur.dtype == np.uint8
    ci = cur.astype(np.int32)
    pi = prev.astype(np.int32)
    d = np.abs(ci - pi).max(axis=2)
    gate = d.astype(_F) < _F(tau)
    a = _F(alpha)
    one_minus = _F(1.0) - a
    mixed = (one_minus * cur.astype(_F)) + (a * prev.astype(_F))
    out = np.where(gate[:, :, None], np.rint(mixed), cur.astype(_F))
    return np.clip(out, 0, 255).astype(np.uint8)
