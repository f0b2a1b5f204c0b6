"""Ablation of the GEMV phase: same phase repeated 400x (whole grid, steady state) with parts disabled."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import qwen3_tts_b200 as q
from qwen3_tts_b200 import synthetic, _lib
from qwen3_tts_b200.engine import AREngine
dev = "cuda:0"; cfg = synthetic.cfg_1p7b(); W = synthetic.random_tts_weights(cfg, device=dev, seed=0)
eng = AREngine(cfg, W, device=dev, max_batch=8, max_ctx=256); H = cfg.talker.hidden_size
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
embs = [(torch.randn(40, H) * 0.5).bfloat16() for _ in range(B)]; pad = (torch.randn(H) * 0.1).bfloat16()
eng.prefill(embs, [torch.zeros(0, H)] * B, pad, q.SamplingParams(max_new_tokens=8, suppress_eos=True))
codes = torch.zeros(B, 8, 16, dtype=torch.int32, device=dev); eng.decode(2, codes); torch.cuda.synchronize()
def t(first, mask, count=400):
    eng.lib.q3_debug_set_skip(eng.h, mask); ms = C.c_float()
    _lib.check(eng.lib.q3_debug_time_phases(eng.h, first, 1, count, C.byref(ms), None)); eng.lib.q3_debug_set_skip(eng.h, 0)
    return ms.value * 1e3 / count
tk = 15 * 28
import os
print("grid", os.environ.get("Q3_GRID", "148"))
for name, idx in (("cp qkv (normed, K=1024)", 1), ("cp o (plain, K=2048)", 3), ("cp gate_up", 4), ("talker qkv", tk), ("talker down (K=6144)", tk + 4)):
    print(f"{name:26s} full {t(idx,0):5.2f} | empty body {t(idx,16):5.2f} | no stage {t(idx,1):5.2f} | no loop {t(idx,2):5.2f} | no epi {t(idx,4):5.2f} | "
          f"no preload {t(idx,8):5.2f} | stage only {t(idx,2|4):5.2f} | loop only {t(idx,1|4):5.2f} | epi only {t(idx,1|2):5.2f}")

print("floor ablation (cp qkv, body skipped): full floor", f"{t(1,16):.2f}", "| no L2 prefetch", f"{t(1,16|32):.2f}", "| no nw prefetch", f"{t(1,16|64):.2f}", "| neither", f"{t(1,16|32|64):.2f}")
print("full phase without L2 prefetch:", f"{t(1,32):.2f}", " without nw prefetch:", f"{t(1,64):.2f}")
