#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
# 1. launch list of one bench step (cold-cache, serialised: compare SHARES)
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
print(open('gpurun_out/r02_launchlist_summary.csv').read()[:1500])
PY
# 2. the frame-step kernel, full set
timeout 900 ncu --set full --clock-control none --import-source on -k regex:q3_step_kernel -s 2 -c 1 -o gpurun_out/r02_decode python tools/ncu_targets.py > gpurun_out/ncu_decode.log 2>&1; tail -2 gpurun_out/ncu_decode.log
# 3. the codec GEMMs, full set (a few of the big residual-stack launches)
timeout 600 ncu --set full --clock-control none -k regex:tap_gemm_kernel -s 260 -c 8 -o gpurun_out/r02_codec_gemm python tools/ncu_targets.py > gpurun_out/ncu_gemm.log 2>&1; tail -2 gpurun_out/ncu_gemm.log
ls -la gpurun_out/*.ncu-rep
