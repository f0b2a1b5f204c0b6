#!/usr/bin/env bash
# One GPU call that re-validates a round: full GPU suite, smoke, the driver's bench line, the launch list and the
# ncu captures that profiles/ summarises.   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash tools/gpu_validate.sh'
set -u
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.json
timeout 1500 python -m pytest tests -m gpu -q -x -s > gpurun_out/t_all.log 2>&1; grep -E "parity\]|passed|failed|Error|assert" gpurun_out/t_all.log | tail -12
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench_full.json"))
    print({k: d[k] for k in ("value", "ms_per_step", "first_packet_ms")}, "frac", d["roofline"]["frac"], d["breakdown_ms_per_step"])
    print("e2e", d["e2e"]["value"], "parity ok", d["parity_check"]["ok"], "cpu", {k: d["cpu_baseline"][k] for k in ("value", "kind", "cores")})
    print("extras", json.dumps(d.get("extras"))[:1800])
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/bench_full.err").read()[-2000:])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 1 --warmup 0 --no-extras --no-cpu-baseline --no-parity-check > gpurun_out/ncu_bench.log 2>&1
python - <<'PY'
import csv, collections
lines = [l for l in open('gpurun_out/r02_launches.csv') if not l.startswith('==')]
rows = [r for r in csv.DictReader(lines) if r.get('Metric Name') == 'gpu__time_duration.sum']
agg = collections.OrderedDict()
for r in rows:
    k = r['Kernel Name'].split('(')[0][-40:]
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += float(r['Metric Value'].replace(',', '')) / 1e3
tot = sum(a[1] for a in agg.values())
with open('gpurun_out/r02_launchlist_summary.csv', 'w') as f:
    f.write("kernel,launches,total_us,share\n")
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        f.write(f"{k},{a[0]},{a[1]:.1f},{a[1] / tot:.4f}\n")
print(open('gpurun_out/r02_launchlist_summary.csv').read()[:600])
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:q3_step_kernel -s 2 -c 1 -o gpurun_out/r02_decode -f python tools/ncu_targets.py > gpurun_out/ncu_decode.log 2>&1; tail -1 gpurun_out/ncu_decode.log
timeout 600 ncu --set full --clock-control none -k regex:tap_gemm_kernel -s 166 -c 16 -o gpurun_out/r02_codec_gemm -f python tools/ncu_targets.py > gpurun_out/ncu_gemm.log 2>&1; tail -1 gpurun_out/ncu_gemm.log
ls -la gpurun_out/*.ncu-rep | tail -3
