"""Event-timed cost of single phase kinds of the frame program (the same phase repeated 400x over the whole grid, steady
state, no device timestamps), with parts of the GEMV phase disabled through the kernel's experiment flags.
Answers: what does a phase cost when its code is hot, and which segment of it."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import qwen3_tts_b200 as q
from qwen3_tts_b200 import synthetic, _lib
from qwen3_tts_b200.engine import AREngine
dev = "cuda:0"; cfg = synthetic.cfg_1p7b(); W = synthetic.random_tts_weights(cfg, device="cpu", seed=0)
eng = AREngine(cfg, W, device=dev, max_batch=8, max_ctx=256); H = cfg.talker.hidden_size
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
embs = [(torch.randn(40, H) * 0.5).bfloat16() for _ in range(B)]; pad = (torch.randn(H) * 0.1).bfloat16()
eng.prefill(embs, [torch.zeros(0, H)] * B, pad, q.SamplingParams(max_new_tokens=8, suppress_eos=True))
codes = torch.zeros(B, 8, 16, dtype=torch.int32, device=dev); eng.decode(2, codes); torch.cuda.synchronize()
n = -eng.lib.q3_describe_frame_program(eng.h, None, 0)
kinds = (C.c_int32 * n)(); eng.lib.q3_describe_frame_program(eng.h, kinds, n); kinds = list(kinds)
def t(first, span, count, mask=0):
    eng.lib.q3_debug_set_skip(eng.h, mask); ms = C.c_float()
    _lib.check(eng.lib.q3_debug_time_phases(eng.h, first, span, count, C.byref(ms), None)); eng.lib.q3_debug_set_skip(eng.h, 0)
    return ms.value * 1e3 / (count * span)
def find(kind, nth=0):
    idx = [i for i, k in enumerate(kinds) if k == kind]
    return idx[nth]
names = [("cp qkv", 10, 2), ("cp o", 12, 2), ("cp gate_up", 13, 1), ("cp down", 12, 3), ("cp head", 14, 1), ("talker qkv", 0, 0), ("talker o", 2, 0),
         ("talker gate_up", 3, 0), ("talker down", 2, 1), ("talker head", 4, 0)]
print(f"B={B}; us per phase, same phase x400")
print(f"{'phase':16s} {'full':>6s} {'noMMA':>6s} {'noStage':>7s} {'noLoop':>6s} {'noEpi':>6s} {'floor':>6s}  (floor = ring consumed, nothing else)")
for name, kind, nth in names:
    i = find(kind, nth)
    print(f"{name:16s} {t(i,1,400):6.2f} {t(i,1,400,16):6.2f} {t(i,1,400,32):7.2f} {t(i,1,400,64):6.2f} {t(i,1,400,128):6.2f} {t(i,1,400,32|64|128):6.2f}")
for name, kind in (("cp attn", 110), ("talker attn", 100)):
    i = find(kind, 1)
    print(f"{name:16s} {t(i,1,400):6.2f}")
i = find(10, 2)
print(f"cp layer (5 phases x80, alternating kinds): {t(i,5,80):6.2f} us/phase;  cp pass (27 phases x15): {t(find(10,5) ,27,15):6.2f}")
i = find(0, 0)
print(f"talker layer (5 phases x80): {t(i,5,80):6.2f} us/phase")
