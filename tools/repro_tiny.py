import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import qwen3_tts_b200 as q
from qwen3_tts_b200 import synthetic
from qwen3_tts_b200.engine import AREngine
dev = "cuda:0"
cfg = synthetic.cfg_tiny()
W = synthetic.random_tts_weights(cfg, device=dev, seed=0)
eng = AREngine(cfg, W, device=dev, max_batch=4, max_ctx=128)
H = cfg.talker.hidden_size
g = torch.Generator().manual_seed(0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
embs = [(torch.randn(7 + i, H, generator=g) * 0.5).bfloat16() for i in range(B)]
pad = (torch.randn(H, generator=g) * 0.1).bfloat16()
sp = q.SamplingParams(do_sample=False, subtalker_dosample=False, max_new_tokens=4, suppress_eos=True)
out = eng.generate(embs, [torch.zeros(0, H)] * B, pad, sp)
print([o.shape for o in out], out[0][0].tolist())
