set -x
run() { VR_L2PERSIST=$3 VR_L2HINT=$1 VR_L2FRAC=$2 python tools/env_frames.py c4_x4plus_720p_qmax_plain 40; }
run 0 0.5 0
run 4 0.5 64
run 4 0.375 48
run 4 0.5 96
for v in "4 0.5 64"; do set -- $v; VR_L2PERSIST=$3 VR_L2HINT=$1 VR_L2FRAC=$2 ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:pair -s 206 -c 15 --csv --log-file gpurun_out/l2exp_$1_$2_$3.csv python tools/run_frames.py > /dev/null 2>&1; python - <<PY
import csv
rows=[l for l in open("gpurun_out/l2exp_$1_$2_$3.csv") if l.startswith('"')]
r=list(csv.DictReader(rows))
agg={}
for d in r:
    agg.setdefault(d["ID"],{})[d["Metric Name"]]=float(d["Metric Value"].replace(",",""))
tot_r=sum(v["dram__bytes_read.sum"] for v in agg.values()); tot_w=sum(v["dram__bytes_write.sum"] for v in agg.values()); t=sum(v["gpu__time_duration.sum"] for v in agg.values())
print("L2EXP hint=$1 frac=$2 persist=$3: 15 launches read", tot_r, "write", tot_w, "time", t, [ (round(v["dram__bytes_read.sum"]/1e6),round(v["gpu__time_duration.sum"]/1e3)) for v in agg.values()])
PY
done
