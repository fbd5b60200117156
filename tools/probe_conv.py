"""GPU bring-up probe for K1 (run with gpurun): conv3x3 tcgen05 kernel vs a numpy fp32 reference.

usage: python tools/probe_conv.py <group>   groups: basic halo dx3 shapes epi special bench
Each group is its own process so one trapped kernel does not poison the others.
"""
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from video_restore_b200 import _lib  # noqa: E402


def ref_conv(x, w, b=None):
    """x [H,W,Cin] (already fp16-rounded), w [Cout,Cin,3,3] (fp16-rounded) -> [H,W,Cout] fp32, zero pad 1."""
    H, W, cin = x.shape
    xp = np.zeros((H + 2, W + 2, cin), np.float32)
    xp[1:-1, 1:-1] = x
    y = np.zeros((H, W, w.shape[0]), np.float32)
    for dy in range(3):
        for dx in range(3):
            y += (xp[dy:dy + H, dx:dx + W].reshape(-1, cin) @ w[:, :, dy, dx].T).reshape(H, W, -1)
    if b is not None:
        y += b
    return y


def h16(a):
    return a.astype(np.float16).astype(np.float32)


def case(name, H, W, cin, cout, seed=0, act=0, prelu=False, res=0, **kw):
    rng = np.random.default_rng(seed)
    x = h16(rng.standard_normal((H, W, cin)).astype(np.float32))
    w = h16((rng.standard_normal((cout, cin, 3, 3)) * (1.0 / np.sqrt(9 * cin))).astype(np.float32))
    b = rng.standard_normal(cout).astype(np.float32) * 0.1
    pr = (rng.random(cout).astype(np.float32) * 0.5) if prelu else None
    r1 = h16(rng.standard_normal((H, W, cout)).astype(np.float32)) if res >= 1 else None
    r2 = h16(rng.standard_normal((H, W, cout)).astype(np.float32)) if res >= 2 else None
    ref = ref_conv(x, w, b)
    if act == 1:
        ref = np.where(ref > 0, ref, ref * 0.2)
    if prelu:
        ref = np.where(ref > 0, ref, ref * pr)
    if r1 is not None:
        ref = ref * 0.2 + r1
    if r2 is not None:
        ref = ref * 0.2 + r2
    if cout == 48:  # pixel shuffle 4 + nearest base (first 3 input channels)
        ps = ref.reshape(H, W, 3, 4, 4).transpose(0, 3, 1, 4, 2).reshape(4 * H, 4 * W, 3)
        base = np.repeat(np.repeat(x[:, :, :3], 4, axis=0), 4, axis=1)
        ref = ps + base
    t0 = time.time()
    try:
        y, ms = _lib.conv3x3(x, w, b, act=2 if prelu else act, prelu=pr, res1=r1, s1=0.2, res2=r2, s2=0.2, **kw)
    except Exception as e:  # noqa: BLE001
        print(f"[{name}] ERROR {e}", flush=True)
        return False
    err = np.abs(y - ref)
    tol = 2e-2 + 4e-3 * np.abs(ref)
    ok = bool((err <= tol).all())
    print(f"[{name}] H={H} W={W} cin={cin} cout={cout} kw={kw} max_err={err.max():.4e} "
          f"mean_err={err.mean():.3e} ref_rms={np.sqrt((ref**2).mean()):.3f} ok={ok} ({time.time()-t0:.1f}s)",
          flush=True)
    if not ok:
        bad = np.argwhere(err > tol)
        print(f"    mismatches: {len(bad)} of {err.size}; first: {bad[:6].tolist()}", flush=True)
        ys, xs = np.unique(bad[:, 0]), np.unique(bad[:, 1])
        print(f"    bad rows {ys[:12].tolist()}.. bad cols {xs[:12].tolist()}..", flush=True)
    return ok


def main():
    group = sys.argv[1] if len(sys.argv) > 1 else "basic"
    # only the conv hooks exist in early bring-up builds
    if "--partial" in sys.argv:
        keep = ("vr_conv3x3_test", "vr_global_error", "vr_conv3x3_bench")
        _lib.SIGNATURES = {k: v for k, v in _lib.SIGNATURES.items() if k in keep}
    if group == "basic":
        case("halo-coll", 8, 128, 32, 32)
        case("halo-2chunk", 8, 128, 64, 32)
    elif group == "shapes":
        case("multi-tile", 37, 300, 64, 32)
        case("cin96", 16, 256, 96, 32)
        case("cin160", 16, 256, 160, 32)
        case("cin192-64", 21, 200, 192, 64)
        case("rows8", 40, 256, 128, 32, rows=8)
        case("small", 5, 17, 64, 64)
        case("wide", 9, 1280, 64, 64)
    elif group == "epi":
        case("lrelu", 12, 140, 64, 32, act=1)
        case("prelu", 12, 140, 64, 64, prelu=True)
        case("res1", 12, 140, 192, 64, res=1)
        case("res2", 12, 140, 192, 64, res=2)
    elif group == "special":
        case("cin3", 12, 140, 3, 64)
        case("cin12", 12, 140, 12, 64)
        case("rgb", 12, 140, 64, 3)
        case("ps4", 12, 140, 64, 48)
    elif group == "roll":
        R = 128  # FLAG_FORCE_ROLL
        case("r-basic", 8, 128, 32, 32, flags=R)
        case("r-2chunk", 8, 128, 64, 32, flags=R)
        case("r-tall", 75, 128, 64, 32, flags=R)          # ring wraps (16 blocks) several times
        case("r-tall64", 75, 200, 64, 64, flags=R)        # ring of 8
        case("r-multi", 37, 300, 64, 32, flags=R)
        case("r-cin160", 40, 256, 160, 32, flags=R)
        case("r-split", 41, 200, 192, 64, flags=R)        # two resident 32-channel halves
        case("r-split-res2", 23, 140, 192, 64, res=2, flags=R)
        case("r-lrelu", 12, 140, 64, 32, act=1, flags=R)
        case("r-prelu", 12, 140, 64, 64, prelu=True, flags=R)
        case("r-small", 5, 17, 64, 64, flags=R)
        case("r-1row", 1, 33, 64, 32, flags=R)
        case("r-2row", 2, 130, 96, 32, flags=R)
        case("r-cin3", 12, 140, 3, 64, flags=R)
        case("r-big", 720, 1280, 64, 32, flags=R)         # 140 items, bands of 52 rows
        case("r-big-split", 300, 1280, 192, 64, res=2, flags=R)
    elif group == "pair":
        Pf = 512  # FLAG_FORCE_PAIR
        case("p-basic", 8, 256, 32, 32, flags=Pf)
        case("p-2chunk", 8, 128, 64, 32, flags=Pf)           # odd strip count: the pair's second strip is empty
        case("p-tall", 75, 256, 64, 32, flags=Pf)            # mirrored ring (period 14) several times round
        case("p-tall64", 75, 200, 64, 64, flags=Pf)          # period 6
        case("p-multi", 37, 300, 64, 32, flags=Pf)
        case("p-cin160", 40, 256, 160, 32, flags=Pf)
        case("p-conv5", 41, 200, 192, 64, flags=Pf)          # 192 -> 64 resident as two 110 KB halves, N = 192
        case("p-conv5-res2", 23, 140, 192, 64, res=2, flags=Pf)
        case("p-lrelu", 12, 140, 64, 32, act=1, flags=Pf)
        case("p-prelu", 12, 140, 64, 64, prelu=True, flags=Pf)
        case("p-small", 5, 17, 64, 64, flags=Pf)
        case("p-1row", 1, 33, 64, 32, flags=Pf)
        case("p-2row", 2, 130, 96, 32, flags=Pf)
        case("p-cin3", 12, 140, 3, 64, flags=Pf)
        case("p-big", 720, 1280, 64, 32, flags=Pf)
        case("p-big-conv5", 300, 1280, 192, 64, res=2, flags=Pf)
    elif group == "planar":
        for base, nm in ((64, "K1"), (128, "K2"), (512, "K3")):
            fl = base + 1024
            case(f"{nm}-pl-64-32", 37, 300, 64, 32, flags=fl)
            case(f"{nm}-pl-160-32", 40, 256, 160, 32, act=1, flags=fl)
            case(f"{nm}-pl-conv5", 41, 200, 192, 64, res=2, flags=fl)
            case(f"{nm}-pl-64-64", 23, 140, 64, 64, prelu=True, res=1, flags=fl)
            case(f"{nm}-pl-cin3", 12, 140, 3, 64, flags=fl)
        case("K1-pl-rgb", 12, 140, 64, 3, flags=64 + 1024)
        case("K1-pl-ps4", 12, 140, 64, 48, flags=64 + 1024)
    elif group == "rbench":
        H, W = 720, 1280
        for cin, cout in [(64, 32), (96, 32), (128, 32), (160, 32), (192, 64), (64, 64)]:
            for label, rows, fl in (("K1", 4, 0), ("K2", 0, 128), ("K2-skip-epi", 0, 128 + 8), ("K3", 0, 512),
                                    ("K3-skip-epi", 0, 512 + 8), ("K3-skip-mma", 0, 512 + 4)):
                try:
                    ms = _lib.conv3x3_bench(H, W, cin, cout, rows=rows, flags=fl, iters=20)
                    cyc = _lib.last_conv_cycles()
                    print(f"[rbench] {cin}->{cout} {label:>12}: {ms*1e3:8.1f} us  {2.0*H*W*cin*cout*9/ms/1e9:7.1f} TFLOP/s  "
                          f"{cyc} cyc/CTA -> {cyc / (ms * 1e3):.0f} MHz", flush=True)
                except Exception as e:  # noqa: BLE001
                    print(f"[rbench] {cin}->{cout} {label}: ERROR {e}", flush=True)
                    return
    elif group == "e64":
        # 64-channel K3 layers under the current VR_EARLY64 setting (conv3x3_bench layers have no residual operands)
        for H, W in ((720, 1280), (480, 854)):
            for cin, cout in [(64, 64), (128, 64), (192, 64)]:
                ms = _lib.conv3x3_bench(H, W, cin, cout, rows=0, flags=512, iters=30)
                print(f"[e64] {H}x{W} {cin}->{cout} K3: {ms*1e3:8.1f} us  {2.0*H*W*cin*cout*9/ms/1e9:7.1f} TFLOP/s", flush=True)
    elif group == "one":
        cin, cout = int(sys.argv[2]), int(sys.argv[3])
        fl = int(sys.argv[4]) if len(sys.argv) > 4 and sys.argv[4].isdigit() else 0
        ms = _lib.conv3x3_bench(720, 1280, cin, cout, rows=4, flags=fl, iters=3)
        print(f"[one] {cin}->{cout} flags={fl}: {ms:.4f} ms  {2.0*720*1280*cin*cout*9/ms/1e9:.1f} TFLOP/s", flush=True)
    elif group == "l2":
        # same layer on a full 720p image (working set >> L2) and on a 128-row band (fits L2): per-pixel cost
        for cin, cout in [(64, 32), (160, 32), (192, 64), (64, 64)]:
            for H in (720, 256, 128, 64):
                ms = _lib.conv3x3_bench(H, 1280, cin, cout, rows=4, flags=0, iters=20)
                mb = H * 1280 * (cin + cout) * 2 / 1e6
                print(f"[l2] {cin}->{cout} H={H:4d} ({mb:6.1f} MB in+out): {ms*1e3:8.1f} us  "
                      f"{ms*1e6/(H*1280):.4f} ns/px  {2.0*H*1280*cin*cout*9/ms/1e9:.0f} TFLOP/s", flush=True)
    elif group == "bench":
        H, W = 720, 1280
        names = {0: "full", 16: "no-weights-ld", 32: "no-act-ld", 2: "skip-tma", 4: "skip-mma", 8: "skip-epi", 10: "mma-only",
                 12: "tma-only", 6: "epi-only"}
        for cin, cout in [(64, 32), (160, 32), (192, 64), (64, 64)]:
            for rows in (4, 8):
                if rows == 8 and cout != 32:
                    continue
                for fl in (0, 16, 32, 2, 4, 8, 10, 12, 6):
                    try:
                        ms = _lib.conv3x3_bench(H, W, cin, cout, rows=rows, flags=fl, iters=10)
                        tf = 2.0 * H * W * cin * cout * 9 / ms / 1e9
                        cyc = _lib.last_conv_cycles()
                        print(f"[bench] {cin}->{cout} rows={rows} {names[fl]:>12}: {ms:.4f} ms  {tf:.1f} TFLOP/s  "
                              f"{cyc} cyc/CTA -> {cyc / (ms * 1e3):.0f} MHz", flush=True)
                    except Exception as e:  # noqa: BLE001
                        print(f"[bench] {cin}->{cout} rows={rows} flags={fl}: ERROR {e}", flush=True)
                        return

if __name__ == "__main__":
    main()
