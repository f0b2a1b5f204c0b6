#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ar.py -x -q -k "tiny or large_batch or free_running or sampler or long_context or two_frames" > gpurun_out/t_ar_tiny.log 2>&1; tail -3 gpurun_out/t_ar_tiny.log
run_bench() {
  tag=$1; shift
  env "$@" timeout 200 python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline --no-parity-check > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_$tag.json"))
    print("$tag: value", round(d["value"], 1), "ms/frame-step", round(d["roofline"]["ms_per_frame_step"], 3), "frac", round(d["roofline"]["frac"], 4), "first_packet", round(d["first_packet_ms"] or 0, 1))
except Exception as e:
    print("$tag bench failed", e); print(open("gpurun_out/bench_$tag.err").read()[-800:])
PY
}
run_bench default Q3_FLAGS=0
run_bench smemfull Q3_SMEM_FULL=1
timeout 200 python tools/phase_ablation.py 8 > gpurun_out/ablation_b8.txt 2>&1; tail -18 gpurun_out/ablation_b8.txt
