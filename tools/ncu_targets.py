"""Workload for the ncu captures: 1.7B shapes, B=8 — prefill, 2 x 16 frame-steps (the second launch is the one to
capture: -k regex:q3_step_kernel -s 2 -c 1), then a codec decode of 8 x 125 frames (tap_gemm_kernel launches)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import qwen3_tts_b200 as q
from qwen3_tts_b200 import synthetic
from qwen3_tts_b200.pipeline import TTSEngine
dev = "cuda:0"
cfg = synthetic.cfg_1p7b(); ccfg = q.CodecConfig()
eng = TTSEngine(cfg, synthetic.random_tts_weights(cfg, device="cpu", seed=0), ccfg, synthetic.random_codec_weights(ccfg, device="cpu", seed=0),
                device=dev, max_batch=8, max_ctx=320, codec_max_frames=136)
H = cfg.talker.hidden_size
B = 8
lens = [16 + 8 * (i % 8) + 11 + (12 if i % 2 else 0) for i in range(B)]
g = torch.Generator().manual_seed(1)
embs = [(torch.randn(L, H, generator=g) * 0.5).bfloat16().to(dev) for L in lens]
trail = [torch.zeros(0, H, dtype=torch.bfloat16, device=dev)] * B
pad = (torch.randn(H, generator=g) * 0.1).bfloat16().to(dev)
sp = q.SamplingParams(max_new_tokens=200, suppress_eos=True, seed=1234)
eng.ar.prefill(embs, trail, pad, sp)
codes = torch.zeros(B, 160, 16, dtype=torch.int32, device=dev)
eng.ar.decode(60, codes)      # context ~ mid-utterance
eng.ar.decode(16, codes)      # <- capture this launch
torch.cuda.synchronize()
wav = eng.codec.chunked_decode(codes[:, :125].transpose(1, 2))
torch.cuda.synchronize()
print("ok", eng.ar.progress()[0], tuple(wav.shape))
