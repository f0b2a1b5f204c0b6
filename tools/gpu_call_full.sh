#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.json
timeout 1500 python -m pytest tests -m gpu -q -x -s > gpurun_out/t_all.log 2>&1; grep -E "parity\]|passed|failed|Error|assert" gpurun_out/t_all.log | tail -20
