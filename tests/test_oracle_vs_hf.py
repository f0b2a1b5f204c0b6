"""Oracle pieces pinned against the INSTALLED third-party code the reference calls (transformers; the reference pins
4.57.3, this image has 5.5.0) — no /root/reference needed, so these run on the GPU box too: the logits-processor list HF
itself builds for the talker's generate() kwargs, and the control flow of GenerationMixin._sample."""
import numpy as np
import pytest
import torch


def _hf_processor_list(**gen_kwargs):
    """The LogitsProcessorList HF itself builds for these generate() kwargs (GenerationMixin._get_logits_processor of
    the installed transformers; the reference pins 4.57.3) — asked of a throw-away 1-layer GPT-2."""
    from transformers import GenerationConfig, GPT2Config, GPT2LMHeadModel, LogitsProcessorList
    m = GPT2LMHeadModel(GPT2Config(n_layer=1, n_head=1, n_embd=8, vocab_size=3072, n_positions=16))
    gc = GenerationConfig(**gen_kwargs)
    m._prepare_special_tokens(gc, kwargs_has_attention_mask=True, device="cpu")  # sets _eos_token_tensor (min_new_tokens needs it)
    return m._get_logits_processor(generation_config=gc, input_ids_seq_length=0, encoder_input_ids=None,
                                   prefix_allowed_tokens_fn=None, logits_processor=LogitsProcessorList(), device="cpu")


def test_sampler_matches_hf_logits_processors():
    """oracle/sampler.py::process_logits against the third-party processors the reference configures at
    modeling_qwen3_tts.py:2044-2066 / :2272-2278 — both the ORDER HF applies them in and the formulas, on the talker's
    settings (repetition penalty over generated ids, min_new_tokens=2, suppress [V-1024, V) \\ {eos}, T, top-k, top-p)."""
    from oracle import sampler as OSm
    V, eos = 3072, 2150
    rng = np.random.default_rng(0)
    suppress = [i for i in range(V - 1024, V) if i != eos]
    for trial, (top_p, n_gen) in enumerate([(1.0, 0), (1.0, 1), (0.8, 5), (0.95, 40)]):
        kw = dict(do_sample=True, top_k=50, top_p=top_p, temperature=0.9, repetition_penalty=1.05, min_new_tokens=2,
                  eos_token_id=eos, suppress_tokens=suppress, max_new_tokens=100, pad_token_id=eos)
        procs = _hf_processor_list(**kw)
        names = [type(p).__name__ for p in procs]
        want = ["RepetitionPenaltyLogitsProcessor", "MinNewTokensLengthLogitsProcessor", "SuppressTokensLogitsProcessor",
                "TemperatureLogitsWarper", "TopKLogitsWarper"] + (["TopPLogitsWarper"] if top_p < 1.0 else [])
        assert names == want, names
        logits = (rng.standard_normal(V) * 3).astype(np.float32)
        gen = rng.integers(0, 2048, size=n_gen)
        ids = torch.from_numpy(gen.astype(np.int64))[None]          # HF's input_ids = generated ids only (inputs_embeds prompt)
        hf = procs(ids, torch.from_numpy(logits)[None].clone())[0].numpy()
        mine = OSm.process_logits(logits, generated_ids=list(gen), repetition_penalty=1.05, min_new_tokens=2,
                                  eos_token_id=eos, suppress_lo=V - 1024, suppress_hi=V, do_sample=True, temperature=0.9,
                                  top_k=50, top_p=top_p)
        assert np.array_equal(np.isfinite(hf), np.isfinite(mine)), f"trial {trial}: kept sets differ"
        keep = np.isfinite(hf)
        assert np.abs(hf[keep] - mine[keep]).max() < 1e-6
    # greedy: no warpers, same three processors
    names = [type(p).__name__ for p in _hf_processor_list(do_sample=False, repetition_penalty=1.05, min_new_tokens=2,
                                                           eos_token_id=eos, suppress_tokens=suppress, max_new_tokens=9,
                                                           pad_token_id=eos)]
    assert names == ["RepetitionPenaltyLogitsProcessor", "MinNewTokensLengthLogitsProcessor", "SuppressTokensLogitsProcessor"]


def test_generation_loop_glue_matches_hf_sample():
    """oracle/talker.py::hf_sample_loop (the control flow the oracle's generate() runs on) against the real
    `GenerationMixin._sample` of the installed transformers, driven the way the reference drives it: inputs_embeds
    prompt (so the processors only ever see GENERATED ids), left-padded batch, eos == pad, min_new_tokens=2,
    repetition penalty, suppress list, greedy.  The model is a throw-away 2-layer GPT-2: the glue is model-agnostic.
    Rows are made to finish at different steps by choosing as EOS a token that row 0 emits mid-sequence."""
    from transformers import GPT2Config, GPT2LMHeadModel
    from oracle import talker as OT
    torch.manual_seed(0)
    V, N = 300, 12
    m = GPT2LMHeadModel(GPT2Config(n_layer=2, n_head=2, n_embd=32, vocab_size=V, n_positions=64)).eval()
    with torch.no_grad():
        for p_ in m.parameters():
            p_.mul_(3.0)                      # sharper logits: greedy paths that wander instead of repeating
    lens = [5, 3, 7]
    B, L = len(lens), max(lens)
    E = torch.randn(B, L, 32)
    mask = torch.zeros(B, L, dtype=torch.long)
    for b, l in enumerate(lens):
        mask[b, L - l:] = 1                   # left padding, like modeling_qwen3_tts.py:2239-2254
    suppress = list(range(V - 40, V))

    E_all, mask_all = E, mask

    def hf(eos, max_new, rows=(0, 1, 2)):
        E, mask = E_all[list(rows)], mask_all[list(rows)]
        sup = [t for t in suppress if t != eos]
        with torch.no_grad():
            return m.generate(inputs_embeds=E, attention_mask=mask, do_sample=False, max_new_tokens=max_new, min_new_tokens=2,
                              repetition_penalty=1.05, suppress_tokens=sup, eos_token_id=eos, pad_token_id=eos)

    def mine(eos, max_new, rows=(0, 1, 2)):
        E, mask = E_all[list(rows)], mask_all[list(rows)]
        B = len(rows)
        wte = m.get_input_embeddings()
        hist = []

        def logits_after(tokens):             # full re-forward: prompt embeddings + embeddings of the tokens so far
            x, am = E, mask
            if tokens:
                t = torch.stack(tokens, 1)
                x = torch.cat([E, wte(t)], 1)
                am = torch.cat([mask, torch.ones(B, t.shape[1], dtype=torch.long)], 1)
            pos = (am.cumsum(-1) - 1).clamp(min=0)
            with torch.no_grad():
                return m(inputs_embeds=x, attention_mask=am, position_ids=pos).logits[:, -1].float()

        def step_fn(tok, step):
            hist.append(tok)
            return logits_after(hist)

        lp = dict(repetition_penalty=1.05, min_new_tokens=2, eos_token_id=eos, suppress_lo=V - 40, suppress_hi=V,
                  do_sample=False, temperature=1.0, top_k=0, top_p=1.0)
        toks, steps = OT.hf_sample_loop(logits_after([]), step_fn, B, lp, False, 0, max_new, eos)
        return torch.stack(toks, 1), steps

    free = hf(eos=V - 1, max_new=N)           # EOS suppressed-range token that never fires: the free-running paths
    assert free.shape == (B, N)
    a, steps = mine(V - 1, N)
    assert torch.equal(a, free) and steps == N - 1
    # an EOS that row 0 emits at step >= 2 (min_new_tokens) and that the other rows emit later or never
    cand = [int(t) for t in free[0, 2:8] if all(int(t) not in free[r, :3].tolist() for r in range(1, B))]
    assert cand, "test construction: no usable EOS candidate"
    eos = cand[0]
    ref = hf(eos, N)
    out, steps = mine(eos, N)
    n = ref.shape[1]
    assert out.shape[1] == n == steps + 1, (out.shape, ref.shape, steps)
    assert torch.equal(out, ref)
    first = [(ref[b] == eos).nonzero()[0].item() if (ref[b] == eos).any() else n for b in range(B)]
    assert min(first) < n - 1 or n < N        # at least one row really finished early and was padded / the loop stopped
    # every row finished -> the loop stops early (EosTokenCriteria): row 0 alone ends right after its EOS
    solo_ref, (solo, solo_steps) = hf(eos, N, rows=(0,)), mine(eos, N, rows=(0,))
    assert solo_ref.shape[1] == first[0] + 1 < N and torch.equal(solo, solo_ref) and solo_steps == first[0]
    # max_new_tokens cap: exactly that many tokens, i.e. max_new_tokens - 1 complete frames in the reference (:2283-2290)
    assert hf(V - 1, 4).shape == (B, 4) and mine(V - 1, 4)[0].shape == (B, 4)
