#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
run_bench() {  # $1 tag, rest: env assignments
  tag=$1; shift
  env "$@" timeout 200 python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline --no-parity-check > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_$tag.json"))
    print("$tag: value", round(d["value"], 1), "ms/frame-step", round(d["roofline"]["ms_per_frame_step"], 3), "frac", round(d["roofline"]["frac"], 4), "first_packet", round(d["first_packet_ms"] or 0, 1))
except Exception as e:
    print("$tag bench failed", e); print(open("gpurun_out/bench_$tag.err").read()[-1200:])
PY
}
run_bench default Q3_FLAGS=0
run_bench nopolicy Q3_FLAGS=4
run_bench keep0 Q3_FLAGS=0 Q3_KEEP_FRACTION=0.0
run_bench keep1 Q3_FLAGS=0 Q3_KEEP_FRACTION=1.0
run_bench ldgstage Q3_FLAGS=2
timeout 200 python tools/critical_path.py --batch 8 > gpurun_out/critical_b8.txt 2>&1; tail -16 gpurun_out/critical_b8.txt
timeout 200 python tools/critical_path.py --batch 1 > gpurun_out/critical_b1.txt 2>&1; tail -16 gpurun_out/critical_b1.txt
