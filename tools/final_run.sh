#!/bin/bash
# Round-end measurement pass on one B200 (gpurun -- bash tools/final_run.sh): GPU tests, smoke, every bench workload, the
# reference arm, long runs, ncu launch list + one dense block with --set full, frame-kernel GB/s. Outputs under gpurun_out/.
O=gpurun_out
python -m pytest tests -m gpu -q -s -p no:cacheprovider > $O/r2_gputests_final.log 2>&1; tail -4 $O/r2_gputests_final.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_smoke.txt 2>&1; tail -1 $O/r2_smoke.txt
python bench.py > $O/r2_bench_c4_enhanced.json 2> $O/r2_bench.err; tail -c 300 $O/r2_bench.err
for w in c4_x4plus_720p_qmax_plain c1_x4plus_256_tile128 c2_x4v3_480p_fast c3_x2plus_1080p_seamless c5_x4plus_1080p; do
  python bench.py --workload $w --no-cpu-baseline > $O/r2_bench_$w.json 2>> $O/r2_bench.err
done
python bench.py --steps 200 --no-cpu-baseline > $O/r2_bench_c4_enhanced_long.json 2>> $O/r2_bench.err
python bench.py --steps 200 --no-cpu-baseline --workload c4_x4plus_720p_qmax_plain > $O/r2_bench_c4_plain_long.json 2>> $O/r2_bench.err
[ -n "$VR_SKIP_REF" ] || python bench.py --impl reference > $O/r2_bench_reference_arm.json 2>> $O/r2_bench.err
python tools/bench_filters.py > $O/r2_filters_gbs.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/r2_ncu_launches_c4_enhanced.csv python tools/run_frames.py --workload c4_x4plus_720p_qmax_enhanced --frames 1 > /dev/null 2>&1
[ -n "$VR_SKIP_REF" ] || ncu --set full --clock-control none --import-source on -k regex:"conv3x3_pair" -s 13 -c 3 -o $O/r2_k4_rdb_enhanced -f python tools/run_frames.py --workload c4_x4plus_720p_qmax_enhanced --frames 1 > $O/r2_ncu_rdb_enh.log 2>&1
ncu --set full --clock-control none -k regex:"bilateral|pre_kernel|post_blend|unsharp|clahe|temporal" -s 8 -c 16 -o $O/r2_filters_innet -f python tools/run_frames.py --workload c4_x4plus_720p_qmax_enhanced --frames 2 > $O/r2_ncu_filters.log 2>&1
python - <<PY
import json, glob
for f in sorted(glob.glob("$O/r2_bench_c*.json") + ["$O/r2_bench_reference_arm.json"]):
    try:
        d = json.load(open(f))
        print(f.split("/")[-1], round(d["value"], 3), "fps", "e2e", round(d.get("e2e", {}).get("value", 0), 3), "frac", round(d.get("roofline", {}).get("frac", 0), 3), "clk", d.get("clocks", {}).get("sm_mhz"), "tol", d.get("tolerance"))
    except Exception as e:
        print(f, "unreadable", e)
PY
