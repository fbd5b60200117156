"""RealESRGANer.tile_process index arithmetic: golden tables + covering properties; and the C ABI's integer twin
(vr_tile_grid needs no GPU) must be bit-identical."""
from pathlib import Path

import numpy as np
import pytest

from oracle.realesrganer import tile_grid

G = Path(__file__).resolve().parent / "golden"


@pytest.fixture(scope="module")
def gold():
    return np.load(G / "tile_grids.npz")


def test_golden_tables(gold):
    for i, c in enumerate(gold["cases"].tolist()):
        assert np.array_equal(tile_grid(*c), gold[f"grid_{i}"]), c


def test_cabi_tile_grid_bit_exact(gold):
    from video_restore_b200.restorer import tile_grid as c_tile_grid

    for i, c in enumerate(gold["cases"].tolist()):
        assert np.array_equal(c_tile_grid(*c), gold[f"grid_{i}"]), c


def test_known_answers():
    g = tile_grid(1080, 1920, 512, 32, 2)      # BASELINE config 3: 12 tiles, thin last row 1080 % 512 = 56
    assert g.shape == (12, 12)
    assert g[-1].tolist() == [1536, 1920, 1024, 1080, 1504, 1920, 992, 1080, 64, 832, 64, 176]
    g = tile_grid(720, 1280, 512, 64, 4)       # config 4 --enhanced: 6 tiles
    assert g.shape[0] == 6 and g[0].tolist() == [0, 512, 0, 512, 0, 576, 0, 576, 0, 2048, 0, 2048]
    assert tile_grid(720, 1280, 1536, 10, 4).shape[0] == 1


@pytest.mark.parametrize("case", [(256, 256, 128, 16, 4), (37, 53, 16, 3, 4), (1080, 1920, 512, 32, 2),
                                  (100, 100, 30, 40, 4), (129, 129, 128, 1, 4)])
def test_tiles_partition_the_output(case):
    H, W, tile, pad, s = case
    cover = np.zeros((H * s, W * s), np.int32)
    for (ix0, ix1, iy0, iy1, px0, px1, py0, py1, ox0, ox1, oy0, oy1) in tile_grid(*case).tolist():
        cover[iy0 * s:iy1 * s, ix0 * s:ix1 * s] += 1
        assert 0 <= px0 <= ix0 < ix1 <= px1 <= W and 0 <= py0 <= iy0 < iy1 <= py1 <= H
        assert ox1 - ox0 == (ix1 - ix0) * s and oy1 - oy0 == (iy1 - iy0) * s
        assert ox1 <= (px1 - px0) * s and oy1 <= (py1 - py0) * s       # crop lies inside the padded output tile
        assert ix0 - px0 <= pad and px1 - ix1 <= pad
    assert (cover == 1).all()                                           # every output pixel written exactly once
