set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_codec.py tests/test_gpu_codec_stream.py tests/test_gpu_e2e.py tests/test_gpu_ar.py -k "not full_shape" -x -q > gpurun_out/t_codec.log 2>&1; tail -3 gpurun_out/t_codec.log
timeout 200 python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline --no-parity-check > gpurun_out/bench_d.json 2> gpurun_out/bench_d.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_d.json"))
    print("value", round(d["value"], 1), "ms/frame-step", round(d["roofline"]["ms_per_frame_step"], 3), "frac", round(d["roofline"]["frac"], 4), d["breakdown_ms_per_step"], "first_packet", round(d["first_packet_ms"] or 0, 1))
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/bench_d.err").read()[-1500:])
PY
Q3_GEMM_TRACE=1 timeout 200 python tools/codec_breakdown.py --trace 2> gpurun_out/codec_trace.txt
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/codec_launches.csv python tools/codec_breakdown.py > /dev/null 2>&1
python tools/codec_breakdown.py --join gpurun_out/codec_trace.txt gpurun_out/codec_launches.csv > gpurun_out/codec_breakdown.txt 2>&1; tail -5 gpurun_out/codec_breakdown.txt
