#!/usr/bin/env bash
# first GPU call of round 2: does the new frame-step kernel work, and how fast is it
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ar.py -x -q -k "tiny or large_batch or free_running or sampler or long_context" > gpurun_out/t_ar_tiny.log 2>&1; tail -5 gpurun_out/t_ar_tiny.log
timeout 200 python -m pytest tests/test_gpu_ar.py -x -q -k "two_frames" > gpurun_out/t_ar_full2.log 2>&1; tail -3 gpurun_out/t_ar_full2.log
for fl in 0 1 2; do
  Q3_FLAGS=$fl timeout 200 python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline --no-parity-check > gpurun_out/bench_f$fl.json 2> gpurun_out/bench_f$fl.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_f$fl.json"))
    print("flags=$fl value", round(d["value"], 1), "ms/frame-step", round(d["roofline"]["ms_per_frame_step"], 3), "frac", round(d["roofline"]["frac"], 4), d["breakdown_ms_per_step"], "first_packet", d["first_packet_ms"])
except Exception as e:
    print("flags=$fl bench failed", e)
    print(open("gpurun_out/bench_f$fl.err").read()[-1500:])
PY
done
timeout 200 python tools/critical_path.py --batch 8 > gpurun_out/critical_b8.txt 2>&1; tail -20 gpurun_out/critical_b8.txt
