"""GPU parity of the fused AR frame-step kernel against the CPU oracle (through the C ABI)."""
import numpy as np
import pytest
import torch

from oracle import talker as OT, sampler as OS, philox
from tests import helpers as Hh

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _tiny(identity_proj=False):
    cfg = OT.cfg_tiny()
    cfg.talker.rope_theta = 1e6
    cfg.cp.rope_theta = 1e4
    if identity_proj:
        cfg.cp = OT.StackCfg(cfg.talker.hidden_size, 2, 4, 2, 128, 256, 2048, rope_theta=1e4)
    return cfg


def _engine(cfg, Wb, max_ctx=256):
    from qwen3_tts_b200.engine import AREngine
    return AREngine(Hh.to_pkg_cfg(cfg), Wb, device=DEV, max_batch=32, max_ctx=max_ctx)


def _check_logits(eng_l, ora_l, what):
    d = np.abs(eng_l - ora_l)
    assert d.max() < Hh.LOGIT_TOL, f"{what}: max |diff| {d.max():.4f}"
    assert d.mean() < Hh.LOGIT_TOL / 5, f"{what}: mean |diff| {d.mean():.4f}"
    # argmax must agree wherever the oracle's top-2 margin exceeds the tolerance
    srt = np.sort(ora_l, -1)
    margin = srt[..., -1] - srt[..., -2]
    agree = np.argmax(eng_l, -1) == np.argmax(ora_l, -1)
    assert agree[margin > 2 * Hh.LOGIT_TOL].all(), f"{what}: argmax disagrees outside the tolerance band"


@pytest.mark.parametrize("identity_proj", [False, True])
@pytest.mark.parametrize("lens", [[7], [5, 9, 7], [3, 40, 17, 33, 8]])
def test_teacher_forced_logits_tiny(lens, identity_proj):
    cfg = _tiny(identity_proj)
    Wb, Wf = Hh.bf16_weights(OT.random_weights(cfg, seed=1))
    B = len(lens)
    embs, trail, pad = Hh.make_inputs(cfg, lens, [(2 * i + 1) % 5 for i in range(B)], seed=3)
    N = 5
    sp = OT.SamplingCfg(do_sample=False, subtalker_dosample=False, max_new_tokens=N + 1, suppress_eos=True)
    ref = OT.generate(Wf, cfg, [e.float() for e in embs], [t.float() for t in trail], pad.float(), sp,
                      record_logits=True)
    forced = torch.stack(ref.codes).numpy()
    eng = _engine(cfg, Wb)
    codes, tl, cl, prog = Hh.run_engine_forced(eng, embs, trail, pad, Hh.to_pkg_sampling(sp), forced, DEV)
    assert prog[0] == N
    assert (codes == forced).all()
    for f in range(N + 1):
        _check_logits(tl[f], ref.record["talker_logits"][f], f"talker frame {f}")
    G = cfg.num_code_groups
    for f in range(N):
        for j in range(G - 1):
            _check_logits(cl[f, j], ref.record["cp_logits"][f * (G - 1) + j], f"cp frame {f} head {j}")
    eng.close()


def test_large_batch_split_pass0_and_eos():
    """B=20 (> 16: code-predictor pass 0 runs as two passes), EOS forced on some rows: finished rows keep
    stepping with pad=eos, output trimmed at first EOS (modeling_qwen3_tts.py:2283-2290)."""
    cfg = _tiny()
    Wb, Wf = Hh.bf16_weights(OT.random_weights(cfg, seed=2))
    B, N = 20, 6
    lens = [4 + (3 * i) % 11 for i in range(B)]
    embs, trail, pad = Hh.make_inputs(cfg, lens, [i % 4 for i in range(B)], seed=5)
    sp = OT.SamplingCfg(do_sample=False, subtalker_dosample=False, max_new_tokens=N + 1)
    g = np.random.default_rng(0)
    forced = g.integers(0, 2000, size=(B, N, cfg.num_code_groups))
    forced[3, 2, 0] = cfg.codec_eos_token_id
    forced[7, 4, 0] = cfg.codec_eos_token_id
    forced[3, 3:, 0] = cfg.codec_eos_token_id  # HF feeds pad(=eos) to finished rows
    forced[7, 5:, 0] = cfg.codec_eos_token_id
    ref = OT.generate(Wf, cfg, [e.float() for e in embs], [t.float() for t in trail], pad.float(), sp,
                      record_logits=True, forced_codes=forced)
    eng = _engine(cfg, Wb)
    codes, tl, cl, prog = Hh.run_engine_forced(eng, embs, trail, pad, Hh.to_pkg_sampling(sp), forced, DEV)
    fd, n_valid, fin = prog
    assert fd == N
    assert n_valid[3] == 2 and n_valid[7] == 4 and fin[3] == 1 and fin[7] == 1
    assert [len(c) for c in ref.codes] == [min(n, N) for n in n_valid]
    for f in range(N + 1):
        _check_logits(tl[f], ref.record["talker_logits"][f], f"talker frame {f}")
    eng.close()


def test_free_running_greedy_matches_oracle():
    cfg = _tiny()
    Wb, Wf = Hh.bf16_weights(OT.random_weights(cfg, seed=4))
    lens = [6, 11]
    embs, trail, pad = Hh.make_inputs(cfg, lens, [3, 0], seed=7)
    sp = OT.SamplingCfg(do_sample=False, subtalker_dosample=False, max_new_tokens=9, suppress_eos=True)
    ref = OT.generate(Wf, cfg, [e.float() for e in embs], [t.float() for t in trail], pad.float(), sp,
                      record_logits=True)
    eng = _engine(cfg, Wb)
    out = eng.generate(embs, trail, pad, Hh.to_pkg_sampling(sp))
    G = cfg.num_code_groups
    for b in range(len(lens)):
        o, r = out[b].cpu().numpy(), ref.codes[b].numpy()
        assert o.shape == r.shape == (8, G)
        if (o == r).all():
            continue
        # first divergence must sit on a near-tie of the oracle (bf16 reduction order)
        f, gidx = np.argwhere(o != r)[0]
        lg = ref.record["talker_logits"][f][b] if gidx == 0 else ref.record["cp_logits"][f * (G - 1) + gidx - 1][b]
        s = np.sort(lg)
        assert s[-1] - s[-2] < 2 * Hh.LOGIT_TOL, (b, f, gidx, s[-1] - s[-2])
    eng.close()


def test_sampler_against_oracle_on_engine_logits():
    """Sampling (rep-penalty, min-new-tokens, suppress, temperature, top-k, top-p, inverse CDF with Philox
    uniforms): feed the ENGINE's own captured logits to the oracle's sampler and require the same tokens."""
    cfg = _tiny()
    Wb, _ = Hh.bf16_weights(OT.random_weights(cfg, seed=6))
    lens = [5, 8, 13]
    B, N, G = len(lens), 12, cfg.num_code_groups
    embs, trail, pad = Hh.make_inputs(cfg, lens, [1, 2, 0], seed=9)
    for top_p, seed in ((1.0, 11), (0.8, 12)):
        sp = OT.SamplingCfg(do_sample=True, subtalker_dosample=True, top_k=50, top_p=top_p, temperature=0.9,
                            subtalker_top_k=50, subtalker_top_p=top_p, subtalker_temperature=0.9,
                            repetition_penalty=1.05, max_new_tokens=N + 1, seed=seed)
        eng = _engine(cfg, Wb)
        V, Vc = cfg.talker.vocab_size, cfg.cp.vocab_size
        tl = torch.zeros(N + 1, B, V, dtype=torch.float32, device=DEV)
        cl = torch.zeros(N, G - 1, B, Vc, dtype=torch.float32, device=DEV)
        eng.set_debug(None, 0, tl, cl)
        eng.prefill(embs, trail, pad, Hh.to_pkg_sampling(sp))
        codes = torch.zeros(B, N + 1, G, dtype=torch.int32, device=DEV)
        eng.decode(N, codes)
        torch.cuda.synchronize()
        fd, n_valid, fin = eng.progress()
        codes, tl, cl = codes.cpu().numpy(), tl.cpu().numpy(), cl.cpu().numpy()
        lp = OT.talker_logits_processors(cfg, sp)
        mism = 0
        for b in range(B):
            gen = []
            for f in range(min(fd, N)):
                s = OS.process_logits(tl[f, b], generated_ids=gen, **lp)
                u = philox.uniform(sp.seed, b, f, 0)
                tok, cdf = OS.sample_from_scores(s, do_sample=True, u=u)
                got = int(codes[b, f, 0])
                if tok != got:
                    # only tolerated on a CDF boundary (fp32 scan order vs fp64 cumsum)
                    lo = cdf[got - 1] if got > 0 else 0.0
                    assert min(abs(u - lo), abs(u - cdf[got])) < 1e-4, (b, f, tok, got, u)
                    mism += 1
                gen.append(got)
                if got == cfg.codec_eos_token_id:
                    break
                for j in range(G - 1):
                    s = OS.process_logits(cl[f, j, b], do_sample=True, temperature=sp.subtalker_temperature,
                                          top_k=sp.subtalker_top_k, top_p=sp.subtalker_top_p)
                    u = philox.uniform(sp.seed, b, f, j + 1)
                    tok, cdf = OS.sample_from_scores(s, do_sample=True, u=u)
                    got = int(codes[b, f, j + 1])
                    if tok != got:
                        lo = cdf[got - 1] if got > 0 else 0.0
                        assert min(abs(u - lo), abs(u - cdf[got])) < 1e-4, (b, f, j, tok, got, u)
                        mism += 1
        assert mism <= 3
        eng.close()


def test_full_shape_1p7b_two_frames():
    """Expected shipped 1.7B shapes, seeded random weights, B=2, teacher-forced."""
    cfg = OT.cfg_1p7b()
    cfg.text_vocab_size = 1000
    W = OT.random_weights(cfg, seed=0, with_text=False)
    Wb, Wf = Hh.bf16_weights(W)
    del W
    lens = [9, 14]
    embs, trail, pad = Hh.make_inputs(cfg, lens, [1, 0], seed=1)
    N = 2
    sp = OT.SamplingCfg(do_sample=False, subtalker_dosample=False, max_new_tokens=N + 1, suppress_eos=True)
    ref = OT.generate(Wf, cfg, [e.float() for e in embs], [t.float() for t in trail], pad.float(), sp,
                      record_logits=True)
    forced = torch.stack(ref.codes).numpy()
    eng = _engine(cfg, Wb, max_ctx=512)
    codes, tl, cl, prog = Hh.run_engine_forced(eng, embs, trail, pad, Hh.to_pkg_sampling(sp), forced, DEV)
    assert prog[0] == N
    scale = float(np.std(ref.record["talker_logits"][0]))
    for f in range(N + 1):
        d = np.abs(tl[f] - ref.record["talker_logits"][f])
        # calibration: the oracle's own bf16-vs-fp32 gap at this shape is max 0.11*std, mean 0.024*std
        assert d.max() < 0.2 * scale and d.mean() < 0.05 * scale, (f, d.max(), d.mean(), scale)
    G = cfg.num_code_groups
    for f in range(N):
        for j in range(G - 1):
            r = ref.record["cp_logits"][f * (G - 1) + j]
            d = np.abs(cl[f, j] - r)
            assert d.max() < 0.2 * float(np.std(r)) and d.mean() < 0.05 * float(np.std(r)), (f, j, d.max(), d.mean())
    eng.close()


MAX_TOL, MEAN_TOL = 0.4, 0.075  # x logit std: 1.5x PyTorch bf16's own gap to fp32 at these shapes (see below)


def _full_shape_case(model, lens, N, seed):
    """Teacher-forced logits of the engine vs the fp32 oracle at the expected shipped shapes.
    Returns (engine, inputs, oracle result, forced codes) so callers can add free-running checks."""
    cfg = OT.cfg_1p7b() if model == "1.7b" else OT.cfg_0p6b()
    cfg.text_vocab_size = 1000
    W = OT.random_weights(cfg, seed=seed, with_text=False)
    Wb, Wf = Hh.bf16_weights(W)
    del W
    B = len(lens)
    embs, trail, pad = Hh.make_inputs(cfg, lens, [(3 * i) % 4 for i in range(B)], seed=seed + 1)
    sp = OT.SamplingCfg(do_sample=False, subtalker_dosample=False, max_new_tokens=N + 1, suppress_eos=True)
    ref = OT.generate(Wf, cfg, [e.float() for e in embs], [t.float() for t in trail], pad.float(), sp,
                      record_logits=True)
    forced = torch.stack(ref.codes).numpy()
    eng = _engine(cfg, Wb, max_ctx=max(lens) + N + 40)
    codes, tl, cl, prog = Hh.run_engine_forced(eng, embs, trail, pad, Hh.to_pkg_sampling(sp), forced, DEV)
    assert prog[0] == N
    assert (codes == forced).all()
    G = cfg.num_code_groups
    worst = 0.0
    for f in range(N + 1):
        r = ref.record["talker_logits"][f]
        scale = float(np.std(r))
        d = np.abs(tl[f] - r)
        # calibration (tools/calibrate_tolerance.py -> profiles/r02_tolerance_calibration.txt): PyTorch's OWN bf16 run of
        # these shapes (ctx 200-230, 8 frames, two seeds) deviates from the fp32 oracle by up to max 0.18*std / mean
        # 0.037*std on the talker logits and max 0.27*std / mean 0.051*std on the code-predictor logits; the engine is
        # held to 1.5x those figures (MAX_TOL, MEAN_TOL)
        assert d.max() < MAX_TOL * scale and d.mean() < MEAN_TOL * scale, (model, B, "talker", f, d.max(), d.mean(), scale)
        worst = max(worst, d.max() / scale)
        srt = np.sort(r, -1)
        margin = srt[..., -1] - srt[..., -2]
        agree = np.argmax(tl[f], -1) == np.argmax(r, -1)
        assert agree[margin > 0.4 * scale].all(), (model, B, "talker argmax outside the tolerance band", f)
    for f in range(N):
        for j in range(G - 1):
            r = ref.record["cp_logits"][f * (G - 1) + j]
            scale = float(np.std(r))
            d = np.abs(cl[f, j] - r)
            assert d.max() < MAX_TOL * scale and d.mean() < MEAN_TOL * scale, (model, B, "cp", f, j, d.max(), d.mean())
            worst = max(worst, d.max() / scale)
    # free-running greedy from the same prompts: report the exact-match rate against the oracle's own greedy codes
    out = eng.generate(embs, trail, pad, Hh.to_pkg_sampling(sp))
    tot = same = rows_equal = 0
    for b in range(B):
        o, r = out[b].cpu().numpy(), ref.codes[b].numpy()
        assert o.shape == r.shape == (N, G)
        tot += o.size
        same += int((o == r).sum())
        rows_equal += int((o == r).all())
        if not (o == r).all():  # first divergence must sit on a near-tie of the oracle's PROCESSED scores
            f, gidx = np.argwhere(o != r)[0]
            lg = ref.record["talker_logits"][f][b] if gidx == 0 else ref.record["cp_logits"][f * (G - 1) + gidx - 1][b]
            sc = lg
            if gidx == 0:  # codebook 0 goes through the talker's processors (suppress range, min-new-tokens, rep. penalty)
                sc = OS.process_logits(lg, generated_ids=[int(x) for x in r[:f, 0]], **OT.talker_logits_processors(cfg, sp))
                sc[cfg.codec_eos_token_id] = -np.inf  # suppress_eos
            s = np.sort(sc[np.isfinite(sc)])
            assert s[-1] - s[-2] < 0.4 * float(np.std(lg)), (model, B, b, f, gidx, s[-1] - s[-2])
    Hh.report_parity(f"full_shape_{model}_B{B}", {"model": model, "batch": B, "frames": N, "ctx_max": max(lens) + N,
                                                  "worst_logit_err_over_std": worst, "free_running_code_match": same / tot,
                                                  "free_running_rows_identical": rows_equal / B})
    eng.close()


# The exact kernel instantiations the bench and the BASELINE configs run: B=8 -> 16 columns in pass 0 (NT=2),
# B=32 -> NT=4 with the split pass 0, 0.6B -> Identity projection at the real hidden sizes; ctx >= 200, >= 8 frames.
@pytest.mark.parametrize("model,B,N", [("1.7b", 8, 8), ("1.7b", 32, 8), ("0.6b", 1, 8), ("0.6b", 8, 8), ("1.7b", 1, 8)])
def test_full_shape_headline_variants(model, B, N):
    lens = [200 + (37 * i) % 61 if B <= 8 else 40 + (53 * i) % 190 for i in range(B)]
    if B > 8:
        lens[0] = 230
    _full_shape_case(model, lens, N, seed=20 + B)


def test_long_context_cross_cta_split_attention():
    """ctx > 128 with few rows => the talker attention is split across CTAs (flash-decoding style) and combined by
    the last arriver; checked against the oracle at ctx ~ 300 and ~ 600 (3 and 5 splits)."""
    cfg = _tiny()
    Wb, Wf = Hh.bf16_weights(OT.random_weights(cfg, seed=8))
    for lens in ([300], [600, 150]):
        B = len(lens)
        embs, trail, pad = Hh.make_inputs(cfg, lens, [0] * B, seed=13)
        N = 3
        sp = OT.SamplingCfg(do_sample=False, subtalker_dosample=False, max_new_tokens=N + 1, suppress_eos=True)
        ref = OT.generate(Wf, cfg, [e.float() for e in embs], [t.float() for t in trail], pad.float(), sp, record_logits=True)
        forced = torch.stack(ref.codes).numpy()
        eng = _engine(cfg, Wb, max_ctx=1024)
        codes, tl, cl, prog = Hh.run_engine_forced(eng, embs, trail, pad, Hh.to_pkg_sampling(sp), forced, DEV)
        assert prog[0] == N
        for f in range(N + 1):
            _check_logits(tl[f], ref.record["talker_logits"][f], f"ctx{lens} talker frame {f}")
        eng.close()


def test_per_step_hidden_states_second_return_value():
    """generate()'s second return (modeling_qwen3_tts.py:2281,2290): the final-norm hidden state of the newest position
    of every step.  Check: codec_head(hidden[s]) reproduces the step's own raw logits, for every row and step."""
    cfg = _tiny()
    Wb, _ = Hh.bf16_weights(OT.random_weights(cfg, seed=12))
    lens = [7, 12, 5]
    B, N = len(lens), 6
    embs, trail, pad = Hh.make_inputs(cfg, lens, [2, 0, 1], seed=14)
    sp = OT.SamplingCfg(do_sample=True, subtalker_dosample=True, max_new_tokens=N + 1, suppress_eos=True, seed=3)
    eng = _engine(cfg, Wb)
    V = cfg.talker.vocab_size
    tl = torch.zeros(N + 1, B, V, dtype=torch.float32, device=DEV)
    eng.set_debug(None, 0, tl, None)
    codes, hid = eng.generate(embs, trail, pad, Hh.to_pkg_sampling(sp), return_hidden=True)
    eng.set_debug(None, 0, None, None)
    head = Wb["talker.codec_head.weight"].float().to(DEV)
    for b in range(B):
        assert hid[b].shape == (codes[b].shape[0], cfg.talker.hidden_size) and codes[b].shape[0] == N
        lg = hid[b].float() @ head.t()                                     # (N, V): steps 0..N-1
        ref = tl[:N, b]
        assert (lg - ref).abs().max().item() < 0.05, (b, (lg - ref).abs().max().item())
    eng.close()
