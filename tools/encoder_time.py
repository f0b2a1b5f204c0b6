import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import qwen3_tts_b200 as q
from qwen3_tts_b200 import synthetic
from qwen3_tts_b200.config import EncoderConfig
from qwen3_tts_b200.codec_encoder import CodecEncoder
cfg = EncoderConfig(); enc = CodecEncoder(cfg, synthetic.random_encoder_weights(cfg, 2), device="cuda:0")
cases = ((1, 72000),) if len(sys.argv) > 1 else ((1, 72000), (8, 72000), (1, 240000))
for B, T in cases:
    wav = (torch.randn(B, T, device="cuda:0") * 0.1).clamp(-1, 1)
    n = 1 if len(sys.argv) > 1 else 5
    for _ in range(0 if len(sys.argv) > 1 else 3): enc.forward(wav)
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): enc.forward(wav)
    e1.record(); torch.cuda.synchronize()
    print(f"encode B={B} T={T}: {e0.elapsed_time(e1)/n:.3f} ms, launches {enc.last_launches()}")
