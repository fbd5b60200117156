"""Steady-state (power-capped) behaviour of K1 ablations: ~2 s per configuration, NVML clock/power sampled meanwhile."""
import sys, threading, time, statistics
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import pynvml
from video_restore_b200 import _lib
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
names = {0: "full", 2: "skip-tma", 4: "skip-mma", 8: "skip-epi", 10: "mma-only", 12: "tma-only", 6: "epi-only",
         512: "K3 full", 516: "K3 skip-mma", 520: "K3 skip-epi", 528: "K3 no-stores", 64: "K1 full"}
def run(cin, cout, fl, iters, rows=4, H=720):
    samples = []; stop = threading.Event()
    def samp():
        while not stop.is_set():
            samples.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0))
            stop.wait(0.05)
    t = threading.Thread(target=samp); t.start()
    ms = _lib.conv3x3_bench(H, 1280, cin, cout, rows=rows, flags=fl, iters=iters)
    stop.set(); t.join()
    tail = samples[len(samples) // 2:] or samples
    clk = statistics.median(s[0] for s in tail); pw = statistics.median(s[1] for s in tail)
    tf = 2.0 * H * 1280 * cin * cout * 9 / ms / 1e9
    print(f"[power] H={H} {ms*1e6/H:6.1f} ns/row {cin}->{cout} rows={rows} {names[fl]:>9}: {ms*1e3:7.1f} us/iter  {tf:7.1f} TFLOP/s  SM {clk:.0f} MHz  {pw:.0f} W  ({len(samples)} samples)", flush=True)
mode = sys.argv[1] if len(sys.argv) > 1 else "ablate"
if mode == "layers":
    for cin, cout in [(64, 32), (96, 32), (128, 32), (160, 32), (192, 64), (64, 64)]:
        base = _lib.conv3x3_bench(720, 1280, cin, cout, rows=4, flags=0, iters=5)
        run(cin, cout, 0, max(50, int(2000.0 / base)))
elif mode == "k3":
    # sustained K3 (CTA pairs) per layer shape, with the MMAs or the epilogue removed, next to K1
    for cin, cout in [(64, 32), (160, 32), (192, 64), (64, 64)]:
        for fl, rows in ((64, 4), (512, 0), (516, 0), (520, 0)):
            base = _lib.conv3x3_bench(720, 1280, cin, cout, rows=rows, flags=fl, iters=5)
            run(cin, cout, fl, max(50, int(2000.0 / base)), rows)
elif mode == "epi":
    # what in the epilogue costs power: full / without the global stores / without the whole epilogue
    for cin, cout in [(192, 64), (160, 32), (64, 64)]:
        for fl in (512, 528, 520):
            base = _lib.conv3x3_bench(720, 1280, cin, cout, rows=0, flags=fl, iters=5)
            run(cin, cout, fl, max(50, int(3000.0 / base)), 0)
elif mode == "l2":
    # same layer on an image whose source + destination fit the 126 MB L2 (re-read from L2 every iteration) and on 720 rows
    for cin, cout in [(128, 32), (64, 32)]:
        for H in (208, 720):
            base = _lib.conv3x3_bench(H, 1280, cin, cout, rows=0, flags=512, iters=5)
            run(cin, cout, 512, max(50, int(2500.0 / base)), 0, H)
elif mode == "rows":
    for cin in (64, 96, 128, 160):
        for rows in (4, 8):
            base = _lib.conv3x3_bench(720, 1280, cin, 32, rows=rows, flags=0, iters=5)
            run(cin, 32, 0, max(50, int(2500.0 / base)), rows)
else:
    for cin, cout in [(160, 32), (192, 64)]:
        for fl in (0, 10, 2, 8, 12, 6):
            base = _lib.conv3x3_bench(720, 1280, cin, cout, rows=4, flags=fl, iters=5)
            run(cin, cout, fl, max(50, int(2500.0 / base)))
