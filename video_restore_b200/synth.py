"""Synthetic frames and weights (there are no datasets or checkpoints in this environment).

Frame recipe: blurred uniform noise stretched to full range + horizontal ramp + N(0,4) noise,
translated 1 px per frame with fresh noise -- exercises the bilateral filter, CLAHE clipping and the temporal gate.
Pure numpy (no cv2) so it runs identically here and on the GPU box.
"""
from __future__ import annotations

import numpy as np


def _box_blur(a: np.ndarray, r: int) -> np.ndarray:
    """Separable running-mean blur with wrap-around (cheap, deterministic)."""
    out = a
    for axis in (0, 1):
        acc = np.zeros_like(out)
        for d in range(-r, r + 1):
            acc += np.roll(out, d, axis=axis)
        out = acc / (2 * r + 1)
    return out


def synth_frame(height: int, width: int, seed: int = 0, index: int = 0) -> np.ndarray:
    """uint8 [H,W,3] BGR. `seed` fixes the scene, `index` is the frame number (1 px/frame pan + fresh noise)."""
    rng = np.random.default_rng(seed)
    base = rng.random((height, width + 64, 3), dtype=np.float32)
    base = _box_blur(_box_blur(base, 2), 2)
    lo, hi = base.min(), base.max()
    base = (base - lo) / max(hi - lo, 1e-6)
    shift = index % 64
    scene = base[:, shift:shift + width] * 178.0
    ramp = np.linspace(0.0, 76.0, width, dtype=np.float32)[None, :, None]
    noise = np.random.default_rng(seed * 1000003 + index + 1).normal(0.0, 4.0, (height, width, 3)).astype(np.float32)
    return np.clip(scene + ramp + noise, 0, 255).astype(np.uint8)


def random_state_dict(model_name: str, seed: int = 0, centre_output: bool = True):
    """Random-init weights with the upstream key names, as float32 numpy arrays.

    Uses torch's own default initialisers through the oracle-independent layer list below (kaiming-normal*0.1 for
    RDB convs, torch Conv2d default elsewhere, PReLU 0.25) -- SURVEY 8 A6. `centre_output` sets the last conv's
    bias to 0.5 so random-init outputs land mid-range instead of clamping at 0 (a random net's output is ~0).
    """
    import torch
    from torch import nn

    from .models import MODEL_ZOO, conv_layers

    spec = MODEL_ZOO[model_name]
    torch.manual_seed(seed)
    sd = {}
    for name, cin, cout, kind in conv_layers(spec):
        conv = nn.Conv2d(cin, cout, 3, 1, 1)
        if kind == "rdb":
            nn.init.kaiming_normal_(conv.weight)
            conv.weight.data *= 0.1
            conv.bias.data.fill_(0)
        sd[name + ".weight"] = conv.weight.detach().numpy().copy()
        sd[name + ".bias"] = conv.bias.detach().numpy().copy()
    if spec["kind"] == "srvgg":
        for i in range(spec["num_conv"] + 1):
            sd[f"body.{2 * i + 1}.weight"] = np.full((64,), 0.25, np.float32)
    if centre_output and spec["kind"] == "rrdb":  # SRVGG adds the input image itself, already mid-range
        sd["conv_last.bias"] = sd["conv_last.bias"] + np.float32(0.5)
    return sd
