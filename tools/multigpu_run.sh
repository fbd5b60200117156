#!/bin/bash
# Multi-GPU records of round 2 (run on the GPU box through `gpurun --gpus N -- bash tools/multigpu_run.sh N`).
# N = 2: bit-identity tests on two distinct devices, BASELINE config 3 and the default workload on 2 GPUs, CLI thread / process paths.
# N = 8: default workload weak scaling, BASELINE config 5 (3000 frames, strong scaling, boundary exchange inside), CLI --procs.
N=${1:-2}
OUT=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
GPUS=$(seq -s ' ' 0 $((N-1)))
if [ "$N" = "2" ]; then
  python -m pytest tests/test_multigpu.py -m gpu -q -s -p no:cacheprovider > $OUT/r2_multigpu_tests_2gpu.log 2>&1; tail -6 $OUT/r2_multigpu_tests_2gpu.log
  $TR bench.py --gpus 2 > $OUT/r2_bench_2gpu_c4_enhanced.json 2> $OUT/r2_bench_2gpu_c4.err
  $TR bench.py --gpus 2 --workload c3_x2plus_1080p_seamless > $OUT/r2_bench_2gpu_c3.json 2>> $OUT/r2_bench_2gpu_c4.err
  python bench.py --gpus 1 --workload c3_x2plus_1080p_seamless --no-cpu-baseline > $OUT/r2_bench_1gpu_c3.json 2>/dev/null
  python video_upscaler.py in out --synthetic 192 --quality max --enhanced --gpus 0 > $OUT/r2_cli_1gpu.txt 2>&1
  python video_upscaler.py in out --synthetic 384 --quality max --enhanced --gpus 0 1 > $OUT/r2_cli_2gpu_threads.txt 2>&1
  python video_upscaler.py in out --synthetic 384 --quality max --enhanced --gpus 0 1 --procs > $OUT/r2_cli_2gpu_procs.txt 2>&1
  tail -n 2 $OUT/r2_cli_1gpu.txt $OUT/r2_cli_2gpu_threads.txt $OUT/r2_cli_2gpu_procs.txt
else
  $TR bench.py --gpus $N > $OUT/r2_bench_${N}gpu_c4_enhanced.json 2> $OUT/r2_bench_${N}gpu.err
  $TR bench.py --gpus $N --workload c5_x4plus_1080p_temporal --frames 3000 > $OUT/r2_bench_${N}gpu_c5_3000frames.json 2>> $OUT/r2_bench_${N}gpu.err
  python video_upscaler.py in out --synthetic $((128*N)) --quality max --enhanced --gpus $GPUS --procs > $OUT/r2_cli_${N}gpu_procs.txt 2>&1
  python video_upscaler.py in out --synthetic 192 --quality max --enhanced --gpus 0 > $OUT/r2_cli_1gpu_b.txt 2>&1
  tail -n 2 $OUT/r2_cli_${N}gpu_procs.txt $OUT/r2_cli_1gpu_b.txt
fi
python - <<PY
import json, glob
for f in sorted(glob.glob("$OUT/r2_bench_*gpu*.json")):
    try:
        d = json.load(open(f))
        print(f.split("/")[-1], "n", d["n_gpus"], round(d["value"], 2), "fps", "e2e", round(d.get("e2e", {}).get("value", 0), 2), "exchange ms", round(d["boundary_exchange_ms"], 2), d["scaling"], "clk", d["clocks"]["sm_mhz"])
    except Exception as e:
        print(f, "unreadable:", e)
PY
