"""All-CTA timestamp analysis of one fused frame-step: per phase kind, who arrives last at the barrier and why."""
import argparse, collections, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import qwen3_tts_b200 as q
from qwen3_tts_b200 import synthetic
from qwen3_tts_b200.engine import AREngine
ap = argparse.ArgumentParser(); ap.add_argument("--batch", type=int, default=1); ap.add_argument("--ctx", type=int, default=60); ap.add_argument("--flags", type=int, default=0)
a = ap.parse_args()
dev = "cuda:0"; cfg = synthetic.cfg_1p7b(); W = synthetic.random_tts_weights(cfg, device=dev, seed=0)
eng = AREngine(cfg, W, device=dev, max_batch=32, max_ctx=a.ctx + 64); H = cfg.talker.hidden_size
eng.lib.q3_debug_set_skip(eng.h, a.flags)
embs = [(torch.randn(a.ctx, H) * 0.5).bfloat16() for _ in range(a.batch)]; pad = (torch.randn(H) * 0.1).bfloat16()
eng.prefill(embs, [torch.zeros(0, H)] * a.batch, pad, q.SamplingParams(max_new_tokens=40, suppress_eos=True))
codes = torch.zeros(a.batch, 40, 16, dtype=torch.int32, device=dev); eng.decode(8, codes); torch.cuda.synchronize()
kinds, *_ = eng.profile_frame(8, codes); T = eng.last_profile_all.astype(np.float64) / 1e3  # us
names = {0: "gemv", 1: "attn", 2: "sample"}; epi = {0: "qkv", 1: "proj", 2: "o/down", 3: "gate_up", 4: "head"}
agg = collections.OrderedDict()
for i, k in enumerate(kinds):
    if i == 0: continue
    ty, stack, e = k // 100, (k // 10) % 10, k % 10
    key = f"{'talker' if stack == 0 else 'cp':6s} {names[ty]:6s} {epi[e] if ty == 0 else ''}"
    start, end, passed = T[i, :, 6], T[i, :, 0], T[i, :, 1]
    rel_prev = T[i - 1, :, 1]                      # when each CTA left the previous barrier
    t0 = rel_prev.max()                            # last CTA released from previous barrier
    last = int(np.argmax(end))
    d = agg.setdefault(key, [])
    m = T[i, last, 2:6]
    cyc = lambda j: T[i, last, j] * 1e3 / 1965.0
    d.append(dict(wait=cyc(5), comp=cyc(9), prod=cyc(10), pre=(T[i, last, 8] - m[0]) if (T[i, last, 8] > 0 and m[0] > 0) else 0.0, w0done=(T[i, last, 7] - m[0]) if (T[i, last, 7] > 0 and m[0] > 0) else 0.0,
                  phase=end.max() - rel_prev.min(), skew_release=rel_prev.max() - rel_prev.min(),
                  body_last=end[last] - rel_prev[last], body_med=np.median(end - rel_prev), body_min=(end - rel_prev).min(),
                  arrive_to_release=passed.min() - end.max(), release_spread=passed.max() - passed.min(),
                  seg=[(m[0] - rel_prev[last]) if m[0] > 0 else 0, (m[1] - m[0]) if m[1] > 0 else 0, (m[2] - m[1]) if m[2] > 0 else 0,
                       0.0, (end[last] - max(m[:3].max(), rel_prev[last]))]))
print(f"B={a.batch} ctx={a.ctx}: all-CTA view; times in us (mean over phases of the kind)")
print(f"{'kind':28s} {'n':>4s} {'phase':>6s} {'body_last':>9s} {'body_med':>8s} {'body_min':>8s} {'arr->rel':>8s} {'rel_spread':>10s} | last CTA segments: entry+stage loop epi (unused) exit")
tot = 0
for k, v in agg.items():
    f = lambda n: np.mean([x[n] for x in v])
    seg = np.mean([x["seg"] for x in v], axis=0)
    tot += f("phase") * len(v)
    print(f"[w0: pre-loop {f('pre'):5.2f} wait {f('wait'):5.2f} compute {f('comp'):5.2f} produce {f('prod'):5.2f} run done +{f('w0done'):5.2f}] ", end="")
    print(f"{k:28s} {len(v):4d} {f('phase'):6.2f} {f('body_last'):9.2f} {f('body_med'):8.2f} {f('body_min'):8.2f} {f('arrive_to_release'):8.2f} {f('release_spread'):10.2f} | " + " ".join(f"{x:5.2f}" for x in seg))
print(f"sum of phase times {tot:.0f} us")
