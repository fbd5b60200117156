"""CPU restatement of realesrgan.RealESRGANer (pre_process / tile_process / post_process / enhance).

TEST INFRASTRUCTURE -- see oracle/__init__.py. Parity status: **unpinned** (realesrgan 0.3.0 `realesrgan/utils.py`
is not on disk; restated from the published algorithm, anchored on the reference call sites
video_upscaler.py:326 (tile_pad rule), :328-338 (constructor), :501 (enhance(frame, outscale=scale))).

Differences from upstream, all deliberate and behaviour-preserving for the reference's usage:
  * device is always CPU, half=False (fp32) -- this is "the reference's PyTorch CPU path" of the north star;
  * only the 3-channel uint8 BGR branch of `enhance` is restated (the reference feeds bgr24 frames,
    video_upscaler.py:232,246); gray / RGBA / 16-bit branches are out of scope;
  * `outscale != scale` is upstream's last step, a cv2 INTER_LANCZOS4 resize of the uint8 result, restated at the end of
    `enhance` (the reference itself always passes outscale == scale, :501,:718);
  * an extra, opt-in `blend="gaussian"` mode implements the README-only seamless tile blending
    (README.md:8,236; definition: SURVEY.md 8 A7) -- not part of upstream.
"""
from __future__ import annotations

import math

import numpy as np
import torch
from torch.nn import functional as F


def tile_grid(height: int, width: int, tile: int, tile_pad: int, scale: int) -> np.ndarray:
    """Integer tile tables of tile_process, row-major over (y, x). int32 [n_tiles, 12]:
    in_x0,in_x1,in_y0,in_y1, pad_x0,pad_x1,pad_y0,pad_y1, out_x0,out_x1,out_y0,out_y1
    where out_* is the crop inside the scaled padded output tile."""
    tiles_x = math.ceil(width / tile)
    tiles_y = math.ceil(height / tile)
    rows = []
    for y in range(tiles_y):
        for x in range(tiles_x):
            ofs_x, ofs_y = x * tile, y * tile
            in_x0, in_x1 = ofs_x, min(ofs_x + tile, width)
            in_y0, in_y1 = ofs_y, min(ofs_y + tile, height)
            pad_x0, pad_x1 = max(in_x0 - tile_pad, 0), min(in_x1 + tile_pad, width)
            pad_y0, pad_y1 = max(in_y0 - tile_pad, 0), min(in_y1 + tile_pad, height)
            out_x0 = (in_x0 - pad_x0) * scale
            out_x1 = out_x0 + (in_x1 - in_x0) * scale
            out_y0 = (in_y0 - pad_y0) * scale
            out_y1 = out_y0 + (in_y1 - in_y0) * scale
            rows.append([in_x0, in_x1, in_y0, in_y1, pad_x0, pad_x1, pad_y0, pad_y1, out_x0, out_x1, out_y0, out_y1])
    return np.asarray(rows, dtype=np.int32).reshape(-1, 12)


def blend_window(extent: int) -> np.ndarray:
    """1-D Gaussian tile weight over a padded output extent: sigma = extent/4, centre (extent-1)/2,
    floored at 1e-3; evaluated in fp32 exactly as written here."""
    u = np.arange(extent, dtype=np.float32)
    c = np.float32((extent - 1) * 0.5)
    inv_sigma = np.float32(4.0) / np.float32(extent)
    t = (u - c) * inv_sigma
    g = np.exp(np.float32(-0.5) * t * t).astype(np.float32)
    return np.maximum(g, np.float32(1e-3)).astype(np.float32)


class RealESRGANer:
    def __init__(self, scale, model, tile=0, tile_pad=10, pre_pad=10, half=False, blend="crop"):
        self.scale = scale
        self.tile_size = tile
        self.tile_pad = tile_pad
        self.pre_pad = pre_pad
        self.mod_scale = None
        self.half = False  # CPU oracle is fp32
        self.blend = blend
        self.model = model.eval()

    def pre_process(self, img: np.ndarray) -> None:
        img_t = torch.from_numpy(np.transpose(img, (2, 0, 1))).float()
        self.img = img_t.unsqueeze(0)
        if self.pre_pad != 0:
            self.img = F.pad(self.img, (0, self.pre_pad, 0, self.pre_pad), "reflect")
        if self.scale == 2:
            self.mod_scale = 2
        elif self.scale == 1:
            self.mod_scale = 4
        if self.mod_scale is not None:
            self.mod_pad_h, self.mod_pad_w = 0, 0
            _, _, h, w = self.img.size()
            if h % self.mod_scale != 0:
                self.mod_pad_h = self.mod_scale - h % self.mod_scale
            if w % self.mod_scale != 0:
                self.mod_pad_w = self.mod_scale - w % self.mod_scale
            self.img = F.pad(self.img, (0, self.mod_pad_w, 0, self.mod_pad_h), "reflect")

    def process(self) -> None:
        self.output = self.model(self.img)

    def tile_process(self) -> None:
        batch, channel, height, width = self.img.shape
        s = self.scale
        self.output = self.img.new_zeros((batch, channel, height * s, width * s))
        grid = tile_grid(height, width, self.tile_size, self.tile_pad, s)
        if self.blend == "gaussian":
            acc = torch.zeros_like(self.output)
            wsum = torch.zeros((1, 1, height * s, width * s), dtype=torch.float32)
        for (ix0, ix1, iy0, iy1, px0, px1, py0, py1, ox0, ox1, oy0, oy1) in grid.tolist():
            input_tile = self.img[:, :, py0:py1, px0:px1]
            output_tile = self.model(input_tile)
            if self.blend == "gaussian":
                # gather form, tiles visited in row-major (y then x) order; fp32 accumulate
                gy = torch.from_numpy(blend_window((py1 - py0) * s))
                gx = torch.from_numpy(blend_window((px1 - px0) * s))
                w2 = gy[:, None] * gx[None, :]
                acc[:, :, py0 * s:py1 * s, px0 * s:px1 * s] += output_tile * w2
                wsum[:, :, py0 * s:py1 * s, px0 * s:px1 * s] += w2
            else:
                self.output[:, :, iy0 * s:iy1 * s, ix0 * s:ix1 * s] = output_tile[:, :, oy0:oy1, ox0:ox1]
        if self.blend == "gaussian":
            self.output = acc / wsum

    def post_process(self) -> torch.Tensor:
        if self.mod_scale is not None:
            _, _, h, w = self.output.size()
            self.output = self.output[:, :, 0:h - self.mod_pad_h * self.scale, 0:w - self.mod_pad_w * self.scale]
        if self.pre_pad != 0:
            _, _, h, w = self.output.size()
            self.output = self.output[:, :, 0:h - self.pre_pad * self.scale, 0:w - self.pre_pad * self.scale]
        return self.output

    @torch.no_grad()
    def enhance(self, img: np.ndarray, outscale=None):
        if img.ndim != 3 or img.shape[2] != 3 or img.dtype != np.uint8:
            raise ValueError("oracle restates only the uint8 HxWx3 BGR branch")
        img = img.astype(np.float32)
        max_range = 65535 if np.max(img) > 256 else 255  # never true for uint8 input
        img = img / max_range
        img_mode = "RGB"
        img = np.ascontiguousarray(img[:, :, ::-1])  # cv2.cvtColor(img, cv2.COLOR_BGR2RGB) on float32: a channel swap
        self.pre_process(img)
        if self.tile_size > 0:
            self.tile_process()
        else:
            self.process()
        output_img = self.post_process()
        output_img = output_img.data.squeeze(0).float().cpu().clamp_(0, 1).numpy()
        output_img = np.transpose(output_img[[2, 1, 0], :, :], (1, 2, 0))
        output = (output_img * 255.0).round().astype(np.uint8)
        if outscale is not None and outscale != float(self.scale):
            import cv2  # upstream's own last step: cv2.resize(..., interpolation=cv2.INTER_LANCZOS4) on the uint8 result

            h_in, w_in = img.shape[:2]
            output = cv2.resize(output, (int(w_in * outscale), int(h_in * outscale)), interpolation=cv2.INTER_LANCZOS4)
        return output, img_mode
