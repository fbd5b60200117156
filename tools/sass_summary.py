"""Per-kernel SASS evidence of the Blackwell paths in libvrb200.so: counts of UTCHMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld /
.st), UTMALDG (TMA tensor load), UBLKCP (bulk copy), SYNCS (mbarrier), STG.E.ENL2.256 / LDG.E.ENL2.256 (256-bit global accesses),
SHFL (warp shuffle) per kernel, from `cuobjdump -sass`. Writes profiles/sass_summary.txt ("committed SASS evidence").

    python tools/sass_summary.py [path/to/libvrb200.so]
"""
import re
import subprocess
import sys
from collections import OrderedDict
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
lib = Path(sys.argv[1]) if len(sys.argv) > 1 else ROOT / "video_restore_b200" / "libvrb200.so"
sass = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True, check=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
pats = OrderedDict([("UTCHMMA", r"\bUTCHMMA"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"), ("UTMALDG", r"\bUTMALDG"),
                    ("UBLKCP", r"\bUBLKCP"), ("SYNCS", r"\bSYNCS"), ("STG.256", r"\bSTG\.E\.ENL2\.256"),
                    ("LDG.256", r"\bLDG\.E\.ENL2\.256"), ("SHFL", r"\bSHFL"), ("HMMA", r"\bHMMA")])
rows = []
for i, blk in enumerate(re.split(r"\n\s*Function : ", sass)[1:]):
    rows.append((names[i], [len(re.findall(p, blk)) for p in pats.values()], blk.count("\n")))
short = lambda n: re.sub(r"\(.*", "", n).replace("vr::", "").replace("(anonymous namespace)::", "")[:58]
out = [f"# cuobjdump -sass {lib.name} (sm_100a), instruction counts per kernel; built by tools/sass_summary.py",
       f"# UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / .st (TMEM), UTMALDG = cp.async.bulk.tensor (TMA), UBLKCP = cp.async.bulk,",
       f"# SYNCS = mbarrier ops, STG.256 / LDG.256 = 256-bit global accesses, SHFL = warp shuffles, HMMA = legacy mma.sync (none expected)",
       f"{'kernel':60s}" + "".join(f"{k:>9s}" for k in pats) + f"{'lines':>8s}"]
for n, c, ln in sorted(rows, key=lambda r: (-r[1][0], r[0])):
    out.append(f"{short(n):60s}" + "".join(f"{v:9d}" for v in c) + f"{ln:8d}")
tot = [sum(r[1][i] for r in rows) for i in range(len(pats))]
out.append(f"{'TOTAL (' + str(len(rows)) + ' kernels)':60s}" + "".join(f"{v:9d}" for v in tot))
text = "\n".join(out) + "\n"
(ROOT / "profiles" / "sass_summary.txt").write_text(text)
print(text)
