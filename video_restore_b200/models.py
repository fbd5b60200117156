"""Model zoo of the drop-in path: `--model` name -> architecture hyper-parameters and layer list.

Mirrors the constructor arguments at reference video_upscaler.py:313-321 (RealESRGAN_x4plus, RealESRGAN_x4_v3,
RealESRGAN_x4plus_anime_6B) and adds RealESRGAN_x2plus, which the reference documents (README.md:158,281) but whose
argparse choices omit (video_upscaler.py:653). State-dict key names are upstream's (basicsr RRDBNet /
realesrgan SRVGGNetCompact) so real checkpoints load unchanged.
"""
from __future__ import annotations

MODEL_ZOO = {
    "RealESRGAN_x4plus": dict(kind="rrdb", scale=4, num_block=23, num_conv=0),
    "RealESRGAN_x2plus": dict(kind="rrdb", scale=2, num_block=23, num_conv=0),
    "RealESRGAN_x4plus_anime_6B": dict(kind="rrdb", scale=4, num_block=6, num_conv=0),
    "RealESRGAN_x4_v3": dict(kind="srvgg", scale=4, num_block=0, num_conv=32),
}

NUM_FEAT = 64
NUM_GROW = 32


def conv_layers(spec):
    """[(state_dict prefix, cin, cout, kind)] in execution order; kind 'rdb' marks kaiming*0.1 initialised convs."""
    out = []
    if spec["kind"] == "rrdb":
        cin0 = 3 * (4 if spec["scale"] == 2 else 1)
        out.append(("conv_first", cin0, NUM_FEAT, "plain"))
        for b in range(spec["num_block"]):
            for r in (1, 2, 3):
                for c in range(1, 6):
                    cin = NUM_FEAT + (c - 1) * NUM_GROW
                    cout = NUM_GROW if c < 5 else NUM_FEAT
                    out.append((f"body.{b}.rdb{r}.conv{c}", cin, cout, "rdb"))
        for name in ("conv_body", "conv_up1", "conv_up2", "conv_hr"):
            out.append((name, NUM_FEAT, NUM_FEAT, "plain"))
        out.append(("conv_last", NUM_FEAT, 3, "plain"))
    else:
        out.append(("body.0", 3, NUM_FEAT, "plain"))
        for i in range(spec["num_conv"]):
            out.append((f"body.{2 * (i + 1)}", NUM_FEAT, NUM_FEAT, "plain"))
        out.append((f"body.{2 * (spec['num_conv'] + 1)}", NUM_FEAT, 3 * 16, "plain"))
    return out


def flops_per_input_pixel(spec) -> int:
    """Algorithmic conv FLOPs (2*MAC) per network-input pixel: x4plus 35 853 696, x4_v3 2 418 048,
    x2plus 8 966 016 (per frame pixel), anime_6B 11 412 864."""
    total = 0.0
    for name, cin, cout, _ in conv_layers(spec):
        mult = 1.0
        if spec["kind"] == "rrdb":
            if name in ("conv_up1",):
                mult = 4.0
            elif name in ("conv_up2", "conv_hr", "conv_last"):
                mult = 16.0
            if spec["scale"] == 2:
                mult /= 4.0  # the network runs on the pixel-unshuffled (half-resolution) grid
        total += 2.0 * 9 * cin * cout * mult
    return int(round(total))
