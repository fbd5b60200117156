"""bench.py's JSON-line contract, checked on the CPU for the arm that needs no GPU (`--impl reference`, tiny crop)."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def _run(*args, env=None):
    import os

    e = dict(os.environ)
    e.update(env or {})
    done = subprocess.run([sys.executable, str(ROOT / "bench.py"), *args], capture_output=True, text=True, timeout=600, env=e, cwd=ROOT)
    assert done.returncode == 0, done.stderr[-2000:]
    lines = [l for l in done.stdout.splitlines() if l.strip()]
    return lines


def test_reference_arm_prints_one_contract_line():
    lines = _run("--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-crop", "32")
    assert len(lines) == 1, "exactly ONE JSON line on stdout"
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0
    assert d["unit"] == "frames/s" and d["higher_is_better"] is True and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert d["value"] > 0 and abs(d["ms_per_step"]) > 0
    assert d["config"]["workload"] == "c4_x4plus_720p_qmax_enhanced" and d["config"]["model"] == "RealESRGAN_x4plus"
    cb = d["cpu_baseline"]
    assert cb["value"] == d["value"] and cb["cores"] >= 1 and cb["kind"].startswith("port") and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    # under torchrun only rank 0 runs the CPU arm; the other ranks print nothing and exit 0
    lines = _run("--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", "--cpu-crop", "32",
                 env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2", "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": "29999"})
    assert lines == []
