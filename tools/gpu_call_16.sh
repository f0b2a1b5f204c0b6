#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ar.py -x -q -k "tiny or large_batch or free_running or sampler or long_context or two_frames" > gpurun_out/t_ar_tiny.log 2>&1; tail -3 gpurun_out/t_ar_tiny.log
timeout 300 python -m pytest tests/test_gpu_ar.py -x -q -k "headline and 1.7b-32" -s > gpurun_out/t_b32.log 2>&1; grep -E "parity\]|passed|failed|Error|assert" gpurun_out/t_b32.log | tail -5
for b in 32 8; do
timeout 200 python bench.py --steps 2 --warmup 1 --batch $b --no-extras --no-cpu-baseline --no-parity-check > gpurun_out/bench_b$b.json 2> gpurun_out/bench_b$b.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_b$b.json"))
    print("B=$b value", round(d["value"], 1), "ms/frame-step", round(d["roofline"]["ms_per_frame_step"], 3), "frac", round(d["roofline"]["frac"], 4), d["breakdown_ms_per_step"], "first_packet", round(d["first_packet_ms"] or 0, 1))
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/bench_b$b.err").read()[-1500:])
PY
done
