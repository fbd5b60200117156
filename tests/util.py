"""Shared helpers for the parity tests: same weights into the oracle (CPU fp32) and the CUDA path."""
from __future__ import annotations

import numpy as np
import torch

from oracle.archs import build_model
from video_restore_b200.synth import random_state_dict, synth_frame  # noqa: F401


def oracle_model_from_sd(model_name: str, sd: dict):
    m = build_model(model_name, seed=None)
    m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=True)
    return m.eval()


def psnr_u8(a: np.ndarray, b: np.ndarray) -> float:
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    return float("inf") if mse == 0 else 10.0 * np.log10(255.0 ** 2 / mse)


def max_lsb(a: np.ndarray, b: np.ndarray) -> int:
    return int(np.abs(a.astype(np.int32) - b.astype(np.int32)).max())


def inrange_state_dict(model_name: str, seed: int = 0, gain: float = 1.0):
    """Random-init weights re-balanced so that the 8-bit frame is NOT dominated by clamping and the body matters.

    With the default init every RRDB returns 1.2 x its input plus a small term, so after 23 blocks the trunk is
    ~66 x conv_first's output, the x4plus / x2plus frames come out 42 % / 48 % saturated, and the 345 dense-block convs
    contribute almost nothing. Here conv_body is scaled by 1.2^-num_block (body and skip contribute comparably, outputs stay in
    range) and, with `gain` > 1, every dense-block conv by `gain` (gain 10 = plain kaiming-normal: each conv then moves the
    features by O(1)). Same tensors go into the oracle and the CUDA path."""
    from video_restore_b200.models import MODEL_ZOO

    sd = random_state_dict(model_name, seed)
    spec = MODEL_ZOO[model_name]
    if spec["kind"] != "rrdb":
        return sd
    k = np.float32(1.2 ** -spec["num_block"])
    sd["conv_body.weight"] = sd["conv_body.weight"] * k
    sd["conv_body.bias"] = sd["conv_body.bias"] * k
    if gain != 1.0:
        for name in sd:
            if ".rdb" in name and name.endswith(".weight"):
                sd[name] = sd[name] * np.float32(gain)
    return sd


def rel_l2(a: np.ndarray, b: np.ndarray) -> float:
    return float(np.linalg.norm((a - b).ravel().astype(np.float64)) / max(np.linalg.norm(b.ravel().astype(np.float64)), 1e-30))


def psnr_unsaturated(a: np.ndarray, ref: np.ndarray) -> tuple[float, float]:
    """(PSNR over the pixels the reference does not clamp, fraction of such pixels): clamped pixels agree trivially."""
    m = (ref > 0) & (ref < 255)
    if not m.any():
        return float("inf"), 0.0
    return psnr_u8(a[m], ref[m]), float(m.mean())
