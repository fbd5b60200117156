"""CPU restatement (PyTorch fp32) of the two network architectures the reference instantiates.

TEST INFRASTRUCTURE -- see oracle/__init__.py. Parity status for this file: **unpinned** (the upstream packages
are absent; structure cross-checked by parameter counts and state_dict key sets in tests/test_oracle_archs.py).

Restated from the published upstream algorithm (third-party, not vendored in /root/reference):
  * basicsr 1.4.2  basicsr/archs/rrdbnet_arch.py  (RRDBNet, RRDB, ResidualDenseBlock)
  * basicsr 1.4.2  basicsr/archs/arch_util.py     (pixel_unshuffle, default_init_weights)
  * realesrgan 0.3.0  realesrgan/archs/srvgg_arch.py  (SRVGGNetCompact)
Reference call sites: video_upscaler.py:314-315 (x4plus), :317-318 (x4_v3), :320-321 (anime_6B).
"""
from __future__ import annotations

import torch
from torch import nn
from torch.nn import functional as F


def pixel_unshuffle(x: torch.Tensor, scale: int) -> torch.Tensor:
    """[b,c,hh,hw] -> [b, c*scale^2, hh/scale, hw/scale]; out channel = c*scale^2 + dy*scale + dx."""
    b, c, hh, hw = x.size()
    if hh % scale != 0 or hw % scale != 0:
        raise RuntimeError(f"pixel_unshuffle: extent {hh}x{hw} not divisible by {scale}")
    h, w = hh // scale, hw // scale
    xv = x.view(b, c, h, scale, w, scale)
    return xv.permute(0, 1, 3, 5, 2, 4).reshape(b, c * scale * scale, h, w)


@torch.no_grad()
def default_init_weights(modules, scale: float = 1.0) -> None:
    for m in modules:
        nn.init.kaiming_normal_(m.weight)
        m.weight.data *= scale
        if m.bias is not None:
            m.bias.data.fill_(0)


class ResidualDenseBlock(nn.Module):
    def __init__(self, num_feat: int = 64, num_grow_ch: int = 32):
        super().__init__()
        self.conv1 = nn.Conv2d(num_feat, num_grow_ch, 3, 1, 1)
        self.conv2 = nn.Conv2d(num_feat + num_grow_ch, num_grow_ch, 3, 1, 1)
        self.conv3 = nn.Conv2d(num_feat + 2 * num_grow_ch, num_grow_ch, 3, 1, 1)
        self.conv4 = nn.Conv2d(num_feat + 3 * num_grow_ch, num_grow_ch, 3, 1, 1)
        self.conv5 = nn.Conv2d(num_feat + 4 * num_grow_ch, num_feat, 3, 1, 1)
        self.lrelu = nn.LeakyReLU(negative_slope=0.2, inplace=True)
        default_init_weights([self.conv1, self.conv2, self.conv3, self.conv4, self.conv5], 0.1)

    def forward(self, x):
        x1 = self.lrelu(self.conv1(x))
        x2 = self.lrelu(self.conv2(torch.cat((x, x1), 1)))
        x3 = self.lrelu(self.conv3(torch.cat((x, x1, x2), 1)))
        x4 = self.lrelu(self.conv4(torch.cat((x, x1, x2, x3), 1)))
        x5 = self.conv5(torch.cat((x, x1, x2, x3, x4), 1))
        return x5 * 0.2 + x


class RRDB(nn.Module):
    def __init__(self, num_feat: int, num_grow_ch: int = 32):
        super().__init__()
        self.rdb1 = ResidualDenseBlock(num_feat, num_grow_ch)
        self.rdb2 = ResidualDenseBlock(num_feat, num_grow_ch)
        self.rdb3 = ResidualDenseBlock(num_feat, num_grow_ch)

    def forward(self, x):
        out = self.rdb1(x)
        out = self.rdb2(out)
        out = self.rdb3(out)
        return out * 0.2 + x


class RRDBNet(nn.Module):
    def __init__(self, num_in_ch=3, num_out_ch=3, scale=4, num_feat=64, num_block=23, num_grow_ch=32):
        super().__init__()
        self.scale = scale
        if scale == 2:
            num_in_ch = num_in_ch * 4
        elif scale == 1:
            num_in_ch = num_in_ch * 16
        self.conv_first = nn.Conv2d(num_in_ch, num_feat, 3, 1, 1)
        self.body = nn.Sequential(*[RRDB(num_feat, num_grow_ch) for _ in range(num_block)])
        self.conv_body = nn.Conv2d(num_feat, num_feat, 3, 1, 1)
        self.conv_up1 = nn.Conv2d(num_feat, num_feat, 3, 1, 1)
        self.conv_up2 = nn.Conv2d(num_feat, num_feat, 3, 1, 1)
        self.conv_hr = nn.Conv2d(num_feat, num_feat, 3, 1, 1)
        self.conv_last = nn.Conv2d(num_feat, num_out_ch, 3, 1, 1)
        self.lrelu = nn.LeakyReLU(negative_slope=0.2, inplace=True)

    def forward(self, x):
        if self.scale == 2:
            feat = pixel_unshuffle(x, scale=2)
        elif self.scale == 1:
            feat = pixel_unshuffle(x, scale=4)
        else:
            feat = x
        feat = self.conv_first(feat)
        body_feat = self.conv_body(self.body(feat))
        feat = feat + body_feat
        feat = self.lrelu(self.conv_up1(F.interpolate(feat, scale_factor=2, mode="nearest")))
        feat = self.lrelu(self.conv_up2(F.interpolate(feat, scale_factor=2, mode="nearest")))
        return self.conv_last(self.lrelu(self.conv_hr(feat)))


class SRVGGNetCompact(nn.Module):
    def __init__(self, num_in_ch=3, num_out_ch=3, num_feat=64, num_conv=16, upscale=4, act_type="prelu"):
        super().__init__()
        if act_type != "prelu":
            raise ValueError("only act_type='prelu' is used by the reference (video_upscaler.py:318)")
        self.num_in_ch, self.num_out_ch, self.num_feat = num_in_ch, num_out_ch, num_feat
        self.num_conv, self.upscale = num_conv, upscale
        self.body = nn.ModuleList()
        self.body.append(nn.Conv2d(num_in_ch, num_feat, 3, 1, 1))
        self.body.append(nn.PReLU(num_parameters=num_feat))
        for _ in range(num_conv):
            self.body.append(nn.Conv2d(num_feat, num_feat, 3, 1, 1))
            self.body.append(nn.PReLU(num_parameters=num_feat))
        self.body.append(nn.Conv2d(num_feat, num_out_ch * upscale * upscale, 3, 1, 1))
        self.upsampler = nn.PixelShuffle(upscale)

    def forward(self, x):
        out = x
        for layer in self.body:
            out = layer(out)
        out = self.upsampler(out)
        base = F.interpolate(x, scale_factor=self.upscale, mode="nearest")
        return out + base


# --- model zoo: `--model` name -> constructor args (video_upscaler.py:313-321; x2plus from README.md:158,281) ---
MODEL_ZOO = {
    "RealESRGAN_x4plus": dict(kind="rrdb", scale=4, num_block=23),
    "RealESRGAN_x2plus": dict(kind="rrdb", scale=2, num_block=23),
    "RealESRGAN_x4plus_anime_6B": dict(kind="rrdb", scale=4, num_block=6),
    "RealESRGAN_x4_v3": dict(kind="srvgg", scale=4, num_conv=32),
}


def build_model(name: str, seed: int | None = 0) -> nn.Module:
    """Random-init network of the named architecture (default torch init + RDB kaiming*0.1)."""
    spec = MODEL_ZOO[name]
    if seed is not None:
        torch.manual_seed(seed)
    if spec["kind"] == "rrdb":
        m = RRDBNet(3, 3, scale=spec["scale"], num_feat=64, num_block=spec["num_block"], num_grow_ch=32)
    else:
        m = SRVGGNetCompact(3, 3, num_feat=64, num_conv=spec["num_conv"], upscale=4, act_type="prelu")
    return m.eval()
