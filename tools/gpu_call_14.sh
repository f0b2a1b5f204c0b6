#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_ar.py::test_full_shape_headline_variants > gpurun_out/t_all_fast.log 2>&1; tail -12 gpurun_out/t_all_fast.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 python bench.py --steps 3 --warmup 2 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench_full.json"))
    print({k: d[k] for k in ("value", "ms_per_step", "first_packet_ms")}, d["roofline"]["frac"], d["breakdown_ms_per_step"])
    print("e2e", d["e2e"]["value"], "parity", d["parity_check"], "cpu", {k: d["cpu_baseline"][k] for k in ("value", "kind", "cores")})
    print("extras", json.dumps(d.get("extras"))[:1500])
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/bench_full.err").read()[-2000:])
PY
