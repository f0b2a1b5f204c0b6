#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
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
run_bench sb2 Q3_SLOT_BLOCKS=2
run_bench nomma Q3_FLAGS=16
run_bench sb2nomma Q3_SLOT_BLOCKS=2 Q3_FLAGS=16
run_bench greedy_b1 Q3_FLAGS=0 X=1
