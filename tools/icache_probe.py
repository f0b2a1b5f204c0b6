"""Is the fused kernel instruction-fetch bound?  Time the SAME phase repeated vs phases of different kinds alternating."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import qwen3_tts_b200 as q
from qwen3_tts_b200 import synthetic, _lib
from qwen3_tts_b200.engine import AREngine
dev = "cuda:0"
cfg = synthetic.cfg_1p7b()
W = synthetic.random_tts_weights(cfg, device=dev, seed=0)
eng = AREngine(cfg, W, device=dev, max_batch=8, max_ctx=256)
H = cfg.talker.hidden_size
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
embs = [(torch.randn(40, H) * 0.5).bfloat16() for _ in range(B)]
pad = (torch.randn(H) * 0.1).bfloat16()
eng.prefill(embs, [torch.zeros(0, H)] * B, pad, q.SamplingParams(max_new_tokens=8, suppress_eos=True))
codes = torch.zeros(B, 8, 16, dtype=torch.int32, device=dev)
eng.decode(2, codes); torch.cuda.synchronize()
n = -eng.lib.q3_describe_frame_program(eng.h, None, 0)
kinds = (C.c_int32 * n)(); eng.lib.q3_describe_frame_program(eng.h, kinds, n)
kinds = list(kinds)
def t(first, span, count):
    ms = C.c_float()
    _lib.check(eng.lib.q3_debug_time_phases(eng.h, first, span, count, C.byref(ms), None))
    return ms.value * 1e3 / (count * span)
# frame program layout (1.7B, joint pass 0): [proj, (qkv, attn, o, gate_up, down) x5, head, sample] x15, then talker
print("kinds[0:8] =", kinds[:8])
print(f"cp proj only              : {t(0,1,400):6.2f} us/phase")
print(f"cp qkv only               : {t(1,1,400):6.2f} us/phase")
print(f"cp attn only              : {t(2,1,400):6.2f} us/phase")
print(f"cp o only                 : {t(3,1,400):6.2f} us/phase")
print(f"cp gate_up only           : {t(4,1,400):6.2f} us/phase")
print(f"cp qkv+attn alternating   : {t(1,2,200):6.2f} us/phase")
print(f"cp layer (5 phases)       : {t(1,5,80):6.2f} us/phase")
print(f"cp pass (27 phases)       : {t(0,27,15):6.2f} us/phase")
tk = 15 * 28
print(f"talker qkv only           : {t(tk,1,200):6.2f} us/phase")
print(f"talker attn only          : {t(tk+1,1,200):6.2f} us/phase")
print(f"talker gate_up only       : {t(tk+3,1,100):6.2f} us/phase")
print(f"talker layer (5 phases)   : {t(tk,5,40):6.2f} us/phase")
