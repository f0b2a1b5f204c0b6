#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 200 python tools/critical_path.py --batch 8 > gpurun_out/critical_b8.txt 2>&1; tail -14 gpurun_out/critical_b8.txt
timeout 200 python tools/critical_path.py --batch 8 --flags 8 > gpurun_out/critical_b8_f8.txt 2>&1; tail -14 gpurun_out/critical_b8_f8.txt
timeout 600 python -m pytest tests/test_gpu_ar.py -x -q -k "headline and (1.7b-32 or 0.6b-1-8)" -s > gpurun_out/t_ar_full.log 2>&1; grep -E "parity|passed|failed|Error|assert" gpurun_out/t_ar_full.log | tail -12
