"""One line per kernel launch from an `ncu --set full` report: duration, DRAM bytes, SM throughput, issue-slot use, registers.
usage: python tools/ncu_kernel_table.py gpurun_out/X.ncu-rep > profiles/X.txt   (needs ncu on PATH; runs on the CPU box)"""
import csv
import io
import subprocess
import sys

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, rows = rows[0], rows[2:]


def g(r, n, default="0"):
    return r[hdr.index(n)] if n in hdr else default


for r in rows:
    rd, wr, dur = float(g(r, "dram__bytes_read.sum")), float(g(r, "dram__bytes_write.sum")), float(g(r, "gpu__time_duration.sum"))
    print(f"{g(r, 'Kernel Name')[:34]:34s} grid={g(r, 'launch__grid_size'):>7s} dur={dur:8.2f}us dram_rd={rd:8.2f}MB dram_wr={wr:7.2f}MB "
          f"({(rd + wr) / max(dur, 1e-9) * 1e-0 / 1e0:6.2f} MB/us) sm_tput={float(g(r, 'sm__throughput.avg.pct_of_peak_sustained_elapsed')):5.1f}% "
          f"issue={float(g(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active')):5.1f}% "
          f"tensor={float(g(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', '0') or 0):5.1f}% regs={g(r, 'launch__registers_per_thread')}")
