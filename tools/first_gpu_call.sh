#!/usr/bin/env bash
# One GPU call that answers everything the next round needs first (≈ 4 minutes on a B200):
#   /usr/local/graft/bin/gpurun --timeout 600 -- 'bash tools/first_gpu_call.sh'
# 1. micro-benchmarks that size the cluster design (ROUND2_PLAN.md step 0)
# 2. the full GPU test suite
# 3. the headline bench line (+ first packet, encoder probe)
# 4. encoder / speaker-encoder timings
set -u
mkdir -p gpurun_out
nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/cluster_probe tools/cluster_probe.cu \
  && timeout 60 tools/cluster_probe > gpurun_out/cluster_probe.txt 2>&1
cat gpurun_out/cluster_probe.txt
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/t_all.log 2>&1; tail -3 gpurun_out/t_all.log
timeout 300 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; cut -c1-400 gpurun_out/bench.json
timeout 100 python tools/encoder_time.py > gpurun_out/enc_time.txt 2>&1; tail -3 gpurun_out/enc_time.txt
timeout 100 python - > gpurun_out/spk_time.txt 2>&1 <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import torch
import qwen3_tts_b200  # noqa: F401
from qwen3_tts_b200 import synthetic
from qwen3_tts_b200.config import SpeakerEncoderConfig
from qwen3_tts_b200.speaker_encoder import SpeakerEncoder
cfg = SpeakerEncoderConfig()
enc = SpeakerEncoder(cfg, synthetic.random_speaker_encoder_weights(cfg, 1), device="cuda:0")
for B, T in ((1, 72000), (8, 72000), (1, 240000)):
    wav = (torch.randn(B, T, device="cuda:0") * 0.1).clamp(-1, 1)
    for _ in range(3):
        enc.embed_waveform(wav)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(5):
        enc.embed_waveform(wav)
    e1.record(); torch.cuda.synchronize()
    print(f"speaker embedding B={B} T={T}: {e0.elapsed_time(e1) / 5:.3f} ms, launches {enc.last_launches()}")
PY
tail -3 gpurun_out/spk_time.txt
