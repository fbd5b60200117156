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
