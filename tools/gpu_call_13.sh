#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_codec_stream.py tests/test_gpu_scheduler.py tests/test_gpu_e2e.py tests/test_gpu_codec.py -x -q > gpurun_out/t_new.log 2>&1; tail -25 gpurun_out/t_new.log
timeout 300 python -m pytest tests/test_gpu_ar.py -x -q -k "headline and 1.7b-1-8" -s > gpurun_out/t_b1.log 2>&1; grep -E "parity\]|passed|failed|Error|assert" gpurun_out/t_b1.log | tail -5
