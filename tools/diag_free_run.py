"""Diagnostic: frame-0/1 logits of a free-running run vs a teacher-forced run vs the oracle (1.7B shapes)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import talker as OT
from tests import helpers as Hh
from qwen3_tts_b200.engine import AREngine
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
DEV = "cuda:0"
cfg = OT.cfg_1p7b(); cfg.text_vocab_size = 1000
W = OT.random_weights(cfg, seed=20 + B, with_text=False)
Wb, Wf = Hh.bf16_weights(W); del W
lens = [200 + (37 * i) % 61 if B <= 8 else 40 + (53 * i) % 190 for i in range(B)]
if B > 8: lens[0] = 230
embs, trail, pad = Hh.make_inputs(cfg, lens, [(3 * i) % 4 for i in range(B)], seed=21 + B)
N = 2
sp = OT.SamplingCfg(do_sample=False, subtalker_dosample=False, max_new_tokens=N + 1, suppress_eos=True)
ref = OT.generate(Wf, cfg, [e.float() for e in embs], [t.float() for t in trail], pad.float(), sp, record_logits=True)
forced = torch.stack(ref.codes).numpy()
eng = AREngine(Hh.to_pkg_cfg(cfg), Wb, device=DEV, max_batch=32, max_ctx=max(lens) + N + 40)
V = cfg.talker.vocab_size; G = cfg.num_code_groups
def run(force):
    tl = torch.zeros(N + 1, B, V, dtype=torch.float32, device=DEV)
    cl = torch.zeros(N, G - 1, B, cfg.cp.vocab_size, dtype=torch.float32, device=DEV)
    f = torch.from_numpy(forced.astype(np.int32)).to(DEV).contiguous() if force else None
    eng.set_debug(f, N if force else 0, tl, cl)
    eng.prefill(embs, trail, pad, Hh.to_pkg_sampling(sp))
    codes = torch.zeros(B, N, G, dtype=torch.int32, device=DEV)
    eng.decode(N, codes); torch.cuda.synchronize()
    eng.set_debug(None, 0, None, None)
    return codes.cpu().numpy(), tl.cpu().numpy(), cl.cpu().numpy()
for name, force in (("forced", True), ("free", False), ("free2", False), ("forced2", True)):
    codes, tl, cl = run(force)
    r0 = ref.record["talker_logits"][0]
    d0 = np.abs(tl[0] - r0).max(-1) / r0.std()
    am = np.argmax(tl[0], -1); om = np.argmax(r0, -1)
    print(name, "frame0: rows with argmax != oracle:", np.nonzero(am != om)[0].tolist(), "max err/std per row (top 5):", np.sort(d0)[-5:].round(3).tolist(),
          "codes[:,0,0]==oracle:", int((codes[:, 0, 0] == forced[:, 0, 0]).sum()), "/", B, " full match frac", float((codes == forced).mean()))
    bad = np.nonzero(codes[:, 0, 0] != forced[:, 0, 0])[0]
    for b in bad[:6]:
        print("   row", b, "len", lens[b], "engine c0", codes[b, 0, 0], "argmax(engine logits)", am[b], "oracle", forced[b, 0, 0], "oracle margin", float(np.sort(r0[b])[-1] - np.sort(r0[b])[-2]))
