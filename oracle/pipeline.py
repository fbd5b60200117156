"""CPU restatement of the per-frame driver (reference video_upscaler.py:490-505) plus the enhancement stage.

TEST INFRASTRUCTURE -- see oracle/__init__.py.

Order of operations for one frame t (the first two lines are the reference's code, the rest is the
README-only stage whose definitions are ours):
    f  = bilateral(frame, 5, 25, 25)            if opts.denoise          video_upscaler.py:495-496
    up = RealESRGANer.enhance(f, outscale=s)    (crop or Gaussian blend) video_upscaler.py:501
    up = unsharp(up, a)                         if opts.sharpen > 0      README.md:12,141
    up = clahe_bgr(up, 2.0, 8)                  if opts.clahe            README.md:11,240
    out_t = temporal(up_t, up_{t-1})            if opts.temporal         README.md:9,237
`up_{t-1}` is the previous frame's result BEFORE its own temporal blend (non-recursive), so frame-range shards
need exactly one boundary frame from their left neighbour.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import filters
from .archs import MODEL_ZOO, build_model
from .realesrganer import RealESRGANer


@dataclass
class FrameOpts:
    denoise: bool = False
    denoise_d: int = 5
    denoise_sigma_color: float = 25.0
    denoise_sigma_space: float = 25.0
    sharpen: float = 0.0
    clahe: bool = False
    clahe_clip: float = 2.0
    clahe_grid: int = 8
    temporal: bool = False
    temporal_alpha: float = 0.2
    temporal_tau: float = 12.0


class OracleRestorer:
    """Mirror of the product's FrameRestorer, on the CPU in fp32."""

    def __init__(self, model_name="RealESRGAN_x4plus", tile=512, tile_pad=10, blend="crop", model=None, seed=0):
        spec = MODEL_ZOO[model_name]
        self.scale = spec["scale"]
        self.model = model if model is not None else build_model(model_name, seed)
        self.upsampler = RealESRGANer(scale=self.scale, model=self.model, tile=tile, tile_pad=tile_pad, pre_pad=0,
                                      half=False, blend=blend)
        self.prev_up = None

    def reset(self):
        self.prev_up = None

    # same temporal-state surface as the product's FrameRestorer (used by the sharder tests)
    def temporal_reset(self):
        self.prev_up = None

    def temporal_set_prev(self, up_prev):
        self.prev_up = np.ascontiguousarray(up_prev, dtype=np.uint8).copy()

    def temporal_get_prev(self, sH=None, sW=None):
        return self.prev_up.copy()

    def upscale_only(self, frame: np.ndarray, opts: FrameOpts) -> np.ndarray:
        """Everything except the temporal blend: returns up_t."""
        f = frame
        if opts.denoise:
            f = filters.bilateral_filter(f, opts.denoise_d, opts.denoise_sigma_color, opts.denoise_sigma_space)
        up, _ = self.upsampler.enhance(f, outscale=self.scale)
        if opts.sharpen > 0:
            up = filters.unsharp_mask(up, opts.sharpen)
        if opts.clahe:
            up = filters.clahe_bgr(up, opts.clahe_clip, opts.clahe_grid)
        return up

    def process_frame(self, frame: np.ndarray, opts: FrameOpts | None = None) -> np.ndarray:
        opts = opts or FrameOpts()
        up = self.upscale_only(frame, opts)
        out = up
        if opts.temporal:
            out = filters.temporal_blend(up, self.prev_up, opts.temporal_alpha, opts.temporal_tau)
            self.prev_up = up
        return out
