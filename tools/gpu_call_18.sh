#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_codec.py tests/test_gpu_codec_stream.py tests/test_gpu_e2e.py -x -q > gpurun_out/t_codec.log 2>&1; tail -4 gpurun_out/t_codec.log
timeout 300 python -m pytest tests/test_gpu_ar.py -x -q -k "tiny or two_frames" > gpurun_out/t_ar_tiny.log 2>&1; tail -2 gpurun_out/t_ar_tiny.log
: > gpurun_out/r02_l2_residency.txt
for kf in 0.0 0.45 0.61 0.8 1.0; do
  Q3_KEEP_FRACTION=$kf timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum --clock-control none -k regex:q3_step_kernel -s 2 -c 1 --csv --log-file gpurun_out/l2_$kf.csv python tools/ncu_targets.py > /dev/null 2>&1
  python - <<PY >> gpurun_out/r02_l2_residency.txt
import csv
rows = [r for r in csv.DictReader(l for l in open("gpurun_out/l2_$kf.csv") if not l.startswith("=="))]
m = {r["Metric Name"]: (float(r["Metric Value"].replace(",", "")), r["Metric Unit"]) for r in rows}
print("keep_fraction $kf:", {k: v for k, v in m.items()})
PY
done
cat gpurun_out/r02_l2_residency.txt
timeout 200 python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline --no-parity-check > gpurun_out/bench_d.json 2> gpurun_out/bench_d.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_d.json"))
    print("value", round(d["value"], 1), "ms/frame-step", round(d["roofline"]["ms_per_frame_step"], 3), "frac", round(d["roofline"]["frac"], 4), d["breakdown_ms_per_step"], "first_packet", round(d["first_packet_ms"] or 0, 1))
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/bench_d.err").read()[-1500:])
PY
