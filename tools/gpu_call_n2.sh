#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_n2.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "n_gpus", "ms_per_step", "unit")}, "e2e", d["e2e"]["value"], "strong", d["strong_scaling"])
except Exception as e:
    print("bench N=2 failed", e); print(open("gpurun_out/bench_n2.err").read()[-2500:])
PY
grep -c "NCCL INFO" gpurun_out/bench_n2.err; grep -m3 "comm\|nranks" gpurun_out/bench_n2.err | cut -c1-200
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err; cut -c1-300 gpurun_out/bench_ref_n2.json
