"""Per-phase device-timestamp profile of one fused frame-step (CTA 0's view): where the time of the decode
kernel goes, by phase kind.  Usage: python tools/profile_frame.py [--batch 8] [--ctx 120] [--model 1.7b]"""
import argparse
import collections
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--ctx", type=int, default=60)
ap.add_argument("--model", default="1.7b")
ap.add_argument("--frames", type=int, default=40)
ap.add_argument("--out", default=None)
a = ap.parse_args()
import qwen3_tts_b200 as q  # noqa: E402
from qwen3_tts_b200 import synthetic  # noqa: E402
from qwen3_tts_b200.engine import AREngine  # noqa: E402

dev = "cuda:0"
cfg = {"1.7b": synthetic.cfg_1p7b, "0.6b": synthetic.cfg_0p6b, "tiny": synthetic.cfg_tiny}[a.model]()
W = synthetic.random_tts_weights(cfg, device=dev, seed=0)
eng = AREngine(cfg, W, device=dev, max_batch=32, max_ctx=a.ctx + a.frames + 16)
H = cfg.talker.hidden_size
g = torch.Generator().manual_seed(0)
embs = [(torch.randn(a.ctx, H, generator=g) * 0.5).bfloat16() for _ in range(a.batch)]
pad = (torch.randn(H, generator=g) * 0.1).bfloat16()
sp = q.SamplingParams(max_new_tokens=a.frames + 1, suppress_eos=True)
eng.prefill(embs, [torch.zeros(0, H)] * a.batch, pad, sp)
codes = torch.zeros(a.batch, a.frames, 16, dtype=torch.int32, device=dev)
eng.decode(8, codes)  # warm
torch.cuda.synchronize()
kinds, t_end, t_bar, marks = eng.profile_frame(8, codes)
n = len(kinds)
start = np.concatenate([[t_bar[0] - (t_bar[0] - t_end[0])], t_bar[:-1]])  # phase i starts when barrier i-1 ended
start[0] = t_end[0]  # unknown start of phase 0: count only its barrier
body = (t_end - start) / 1e3
wait = (t_bar - t_end) / 1e3
names = {0: "gemv", 1: "attn", 2: "sample"}
epi = {0: "store(qkv)", 1: "bias(proj)", 2: "resid(o/down)", 3: "swiglu(gate_up)", 4: "logits(head)"}
agg = collections.OrderedDict()
sub = collections.OrderedDict()
for i, k in enumerate(kinds):
    ty, stack, e = k // 100, (k // 10) % 10, k % 10
    key = f"{'talker' if stack == 0 else 'cp':6s} {names[ty]:6s} {epi[e] if ty == 0 else ''}"
    d = agg.setdefault(key, [0, 0.0, 0.0])
    d[0] += 1; d[1] += body[i]; d[2] += wait[i]
    if i > 0 and marks[i, 0] > 0:
        pts = [start[i]] + [m for m in marks[i] if m > 0] + [t_end[i]]
        segs = np.diff(np.array(pts, dtype=np.float64)) / 1e3
        sd = sub.setdefault(key, [0, np.zeros(6)])
        sd[0] += 1; sd[1][:len(segs)] += segs
total = (t_bar[-1] - t_end[0]) / 1e3
print(f"B={a.batch} ctx={a.ctx} model={a.model}: {n} phases, frame-step {total:.1f} us (CTA 0 view)")
print(f"{'kind':40s} {'count':>5s} {'body_us':>9s} {'avg':>7s} {'barrier_us':>10s} {'avg':>7s}")
for k, (c, b, w) in agg.items():
    print(f"{k:40s} {c:5d} {b:9.1f} {b / c:7.2f} {w:10.1f} {w / c:7.2f}")
print(f"{'TOTAL':40s} {n:5d} {body.sum():9.1f} {body.mean():7.2f} {wait.sum():10.1f} {wait.mean():7.2f}")
print('inner segments (avg us): entry->m2, m2->m3, m3->m4, m4->m5, ->end')
for k, (c, v) in sub.items():
    print(f"{k:40s} " + ' '.join(f'{x / c:7.2f}' for x in v))
if a.out:
    json.dump({"batch": a.batch, "ctx": a.ctx, "model": a.model, "total_us": total,
               "by_kind": {k: {"count": c, "body_us": b, "barrier_us": w} for k, (c, b, w) in agg.items()}},
              open(a.out, "w"), indent=1)
