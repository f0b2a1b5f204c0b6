"""CPU oracle: talker + code predictor forward passes and the nested generation loop.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Plain torch ops in the dtype of the weights handed in
(fp32 for strict checks, bf16 to mirror the reference's GPU rounding points).  Every function cites the
reference lines it restates (paths relative to /root/reference/qwen_tts/core/models/).

Weights are a flat dict keyed by the reference's own state_dict names with the `talker.` prefix, e.g.
`talker.model.layers.0.self_attn.q_proj.weight`, `talker.code_predictor.lm_head.3.weight`.
"""
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np
import torch
import torch.nn.functional as F

from . import sampler as S
from . import philox


@dataclass
class StackCfg:
    """Shape of one decoder stack (talker or code predictor).  configuration_qwen3_tts.py:187-212,370-404."""
    hidden_size: int
    num_layers: int
    num_heads: int
    num_kv_heads: int
    head_dim: int
    intermediate_size: int
    vocab_size: int
    rms_eps: float = 1e-6
    rope_theta: float = 1e6


@dataclass
class TTSCfg:
    talker: StackCfg
    cp: StackCfg
    num_code_groups: int = 16
    text_hidden_size: int = 2048
    text_vocab_size: int = 151936
    codec_eos_token_id: int = 2150
    codec_pad_id: int = 2148
    codec_bos_id: int = 2149
    codec_think_id: int = 2154
    codec_nothink_id: int = 2155
    codec_think_bos_id: int = 2156
    codec_think_eos_id: int = 2157
    tts_bos_token_id: int = 151672
    tts_eos_token_id: int = 151673
    tts_pad_token_id: int = 151671


@dataclass
class SamplingCfg:
    """Generation kwargs.  inference/qwen3_tts_model.py:287-352 defaults; modeling_qwen3_tts.py:2044-2066."""
    do_sample: bool = True
    top_k: int = 50
    top_p: float = 1.0
    temperature: float = 0.9
    repetition_penalty: float = 1.05
    subtalker_dosample: bool = True
    subtalker_top_k: int = 50
    subtalker_top_p: float = 1.0
    subtalker_temperature: float = 0.9
    min_new_tokens: int = 2
    max_new_tokens: int = 2048
    suppress_eos: bool = False  # benchmark-only switch: fixed horizon with random weights
    seed: int = 0


# ----------------------------------------------------------------------------------------------
# leaf ops
# ----------------------------------------------------------------------------------------------
def rms_norm(x, w, eps):
    """modeling_qwen3_tts.py:605-610 — fp32 mean-square, rsqrt, cast to input dtype, THEN times weight."""
    dt = x.dtype
    xf = x.to(torch.float32)
    var = xf.pow(2).mean(-1, keepdim=True)
    xf = xf * torch.rsqrt(var + eps)
    return w * xf.to(dt)


def rotate_half(x):
    """modeling_qwen3_tts.py:615-619."""
    h = x.shape[-1] // 2
    return torch.cat((-x[..., h:], x[..., :h]), dim=-1)


def rope_cos_sin(positions, head_dim, theta, dtype):
    """modeling_qwen3_tts.py:546-559 / :581-592 — fp32 tables, cast to activation dtype.
    The talker's 3 M-RoPE position rows are identical for TTS (:1500-1502,1711), so the interleaved
    M-RoPE (:692-712) degenerates to plain RoPE.  positions: (B,S) int64."""
    inv = 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.int64).to(torch.float32) / head_dim))
    freqs = positions.to(torch.float32)[..., None] * inv  # (B,S,d/2)
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos().to(dtype), emb.sin().to(dtype)


def apply_rope(q, k, cos, sin):
    """modeling_qwen3_tts.py:878-882 (and :722-723).  q,k: (B,h,S,d); cos,sin: (B,S,d)."""
    cos = cos.unsqueeze(1)
    sin = sin.unsqueeze(1)
    return (q * cos) + (rotate_half(q) * sin), (k * cos) + (rotate_half(k) * sin)


def eager_attention(q, k, v, mask, scaling, n_rep):
    """modeling_qwen3_tts.py:622-657 — repeat_kv, QK^T*scale + mask, fp32 softmax, cast, PV."""
    B, nkv, S, d = k.shape
    k = k[:, :, None].expand(B, nkv, n_rep, S, d).reshape(B, nkv * n_rep, S, d)
    v = v[:, :, None].expand(B, nkv, n_rep, S, d).reshape(B, nkv * n_rep, S, d)
    w = torch.matmul(q, k.transpose(2, 3)) * scaling
    if mask is not None:
        w = w + mask
    w = F.softmax(w, dim=-1, dtype=torch.float32).to(q.dtype)
    o = torch.matmul(w, v)
    return o.transpose(1, 2).contiguous()


def decoder_layer(W, pfx, cfg: StackCfg, x, cos, sin, mask, cache, li):
    """modeling_qwen3_tts.py:1393-1424 (talker) / :985-1012 (code predictor); attention :761-805 / :916-958;
    MLP :853-855.  cache: list of [k,v] per layer, (B,nkv,S,d), concatenated like DynamicCache.update (:785)."""
    B, Sq, _ = x.shape
    p = f"{pfx}.layers.{li}"
    res = x
    h = rms_norm(x, W[f"{p}.input_layernorm.weight"], cfg.rms_eps)
    q = F.linear(h, W[f"{p}.self_attn.q_proj.weight"]).view(B, Sq, -1, cfg.head_dim)
    k = F.linear(h, W[f"{p}.self_attn.k_proj.weight"]).view(B, Sq, -1, cfg.head_dim)
    v = F.linear(h, W[f"{p}.self_attn.v_proj.weight"]).view(B, Sq, -1, cfg.head_dim)
    q = rms_norm(q, W[f"{p}.self_attn.q_norm.weight"], cfg.rms_eps).transpose(1, 2)
    k = rms_norm(k, W[f"{p}.self_attn.k_norm.weight"], cfg.rms_eps).transpose(1, 2)
    v = v.transpose(1, 2)
    q, k = apply_rope(q, k, cos, sin)
    if cache[li] is None:
        cache[li] = [k, v]
    else:
        cache[li] = [torch.cat((cache[li][0], k), dim=2), torch.cat((cache[li][1], v), dim=2)]
    kk, vv = cache[li]
    a = eager_attention(q, kk, vv, mask, cfg.head_dim ** -0.5, cfg.num_heads // cfg.num_kv_heads)
    a = F.linear(a.reshape(B, Sq, -1), W[f"{p}.self_attn.o_proj.weight"])
    x = res + a
    res = x
    h = rms_norm(x, W[f"{p}.post_attention_layernorm.weight"], cfg.rms_eps)
    g = F.linear(h, W[f"{p}.mlp.gate_proj.weight"])
    u = F.linear(h, W[f"{p}.mlp.up_proj.weight"])
    h = F.linear(F.silu(g) * u, W[f"{p}.mlp.down_proj.weight"])
    return res + h


def stack_forward(W, pfx, cfg: StackCfg, x, positions, mask, cache):
    """modeling_qwen3_tts.py:1520-1561 / :1112-1153 — layers then final RMSNorm."""
    cos, sin = rope_cos_sin(positions, cfg.head_dim, cfg.rope_theta, x.dtype)
    for li in range(cfg.num_layers):
        x = decoder_layer(W, pfx, cfg, x, cos, sin, mask, cache, li)
    return rms_norm(x, W[f"{pfx}.norm.weight"], cfg.rms_eps)


def causal_mask(valid_kv, q_pos_in_cache, dtype):
    """Additive mask equivalent to HF create_causal_mask with a 2-D padding mask
    (call sites modeling_qwen3_tts.py:1510-1518, :1105-1110).
    valid_kv: (B,ctx) bool — key slot holds a real token; q_pos_in_cache: (Sq,) cache index of each query."""
    B, ctx = valid_kv.shape
    kidx = torch.arange(ctx)
    allowed = (kidx[None, None, :] <= q_pos_in_cache[None, :, None]) & valid_kv[:, None, :]
    m = torch.zeros(B, 1, q_pos_in_cache.numel(), ctx, dtype=dtype)
    m.masked_fill_(~allowed[:, None], torch.finfo(dtype).min)
    return m


# ----------------------------------------------------------------------------------------------
# code predictor (one frame: 15 passes)
# ----------------------------------------------------------------------------------------------
def code_predictor_frame(W, cfg: TTSCfg, past_hidden, c0, sp: SamplingCfg, frame_idx, rows=None,
                         forced: Optional[np.ndarray] = None, record=None):
    """modeling_qwen3_tts.py:1669-1681 (caller) and :1250-1312 (forward): 2-token prefill -> lm_head[0],
    then 14 single-token passes; pass j embeds with codec_embedding[j-1] and reads lm_head[j]; positions
    0..16, fresh cache per frame.  past_hidden: (B,1,H); c0: (B,) int64.  Returns (B,15) int64."""
    B = c0.shape[0]
    G = cfg.num_code_groups
    pfx = "talker.code_predictor.model"
    cache = [None] * cfg.cp.num_layers
    e0 = F.embedding(c0[:, None], W["talker.model.codec_embedding.weight"])  # :1670
    x = torch.cat((past_hidden, e0), dim=1)  # (B,2,H)
    out = torch.zeros(B, G - 1, dtype=torch.int64)
    rows = list(range(B)) if rows is None else rows
    for j in range(G - 1):
        if j > 0:
            x = F.embedding(out[:, j - 1:j], W[f"{pfx}.codec_embedding.{j - 1}.weight"])  # :1281
        if "talker.code_predictor.small_to_mtp_projection.weight" in W:  # :1171-1174,1282
            x = F.linear(x, W["talker.code_predictor.small_to_mtp_projection.weight"],
                         W["talker.code_predictor.small_to_mtp_projection.bias"])
        Sq = x.shape[1]
        start = 0 if j == 0 else j + 1
        pos = torch.arange(start, start + Sq)[None].expand(B, -1)
        ctx = start + Sq
        mask = causal_mask(torch.ones(B, ctx, dtype=torch.bool), torch.arange(start, start + Sq), x.dtype)
        h = stack_forward(W, pfx, cfg.cp, x, pos, mask, cache)
        logits = F.linear(h[:, -1], W[f"talker.code_predictor.lm_head.{j}.weight"]).to(torch.float32)  # :1299
        if record is not None:
            record.setdefault("cp_logits", []).append(logits.numpy().copy())
        for b in range(B):
            s = S.process_logits(logits[b].numpy(), do_sample=sp.subtalker_dosample,
                                 temperature=sp.subtalker_temperature, top_k=sp.subtalker_top_k,
                                 top_p=sp.subtalker_top_p)
            u = philox.uniform(sp.seed, rows[b], frame_idx, j + 1) if sp.subtalker_dosample else None
            tok, _ = S.sample_from_scores(s, do_sample=sp.subtalker_dosample, u=u)
            out[b, j] = tok
        if forced is not None:
            out[:, j] = torch.from_numpy(forced[:, j])
    return out


def embed_frame(W, cfg: TTSCfg, codes16):
    """modeling_qwen3_tts.py:1682-1687 — sum of the 16 per-codebook embeddings (E_0 = talker codec_embedding,
    E_g = code_predictor.codec_embedding[g-1]).  codes16: (B,16) -> (B,1,H)."""
    parts = [F.embedding(codes16[:, :1], W["talker.model.codec_embedding.weight"])]
    for g in range(1, cfg.num_code_groups):
        parts.append(F.embedding(codes16[:, g:g + 1], W[f"talker.code_predictor.model.codec_embedding.{g - 1}.weight"]))
    return torch.cat(parts, dim=1).sum(1, keepdim=True)


# ----------------------------------------------------------------------------------------------
# talker generation loop (restates HF _sample + the forward's generate branch)
# ----------------------------------------------------------------------------------------------
@dataclass
class GenResult:
    codes: List[torch.Tensor]                 # per row (N_i,16) int64, trimmed at first EOS (:2283-2290)
    steps: int = 0
    record: dict = field(default_factory=dict)


def talker_logits_processors(cfg: TTSCfg, sp: SamplingCfg):
    V = cfg.talker.vocab_size
    return dict(repetition_penalty=sp.repetition_penalty, min_new_tokens=sp.min_new_tokens,
                eos_token_id=cfg.codec_eos_token_id, suppress_lo=V - 1024, suppress_hi=V,
                do_sample=sp.do_sample, temperature=sp.temperature, top_k=sp.top_k, top_p=sp.top_p)


def hf_sample_loop(first_logits, step_fn, B, lp, do_sample, seed, max_new_tokens, eos, forced_tok=None, on_logits=None):
    """The control flow of HF `GenerationMixin._sample` as the reference drives it (talker.generate at
    modeling_qwen3_tts.py:2272-2278; third-party transformers==4.57.3), separated from the model so that it can be
    pinned against the real `_sample` with any model (tests/test_oracle_vs_reference.py):
      scores = processors(generated ids of the row, logits); token = argmax | sample;
      rows that already emitted EOS keep stepping but receive pad_token_id (= eos);
      a row finishes on EOS; the loop ends when every row has finished or max_new_tokens tokens exist.
    first_logits: (B, V) fp32 from the prefill; step_fn(tokens (B,), step) -> next (B, V) logits.
    Returns (list of (B,) token tensors, number of step_fn calls)."""
    generated = [[] for _ in range(B)]
    finished = [False] * B

    def pick(logits, idx):
        tok = torch.zeros(B, dtype=torch.int64)
        for b in range(B):
            s = S.process_logits(logits[b].numpy(), generated_ids=generated[b], **lp)
            u = philox.uniform(seed, b, idx, 0) if do_sample else None
            t, _ = S.sample_from_scores(s, do_sample=do_sample, u=u)
            tok[b] = eos if finished[b] else t  # HF pads finished rows with pad_token_id (= eos)
        f = forced_tok(idx) if forced_tok is not None else None
        return tok if f is None else f

    def update(tok):
        for b in range(B):
            if not finished[b]:
                generated[b].append(int(tok[b]))
                finished[b] = int(tok[b]) == eos

    if on_logits is not None:
        on_logits(first_logits)
    tok = pick(first_logits, 0)
    update(tok)
    tokens = [tok]
    step = 0  # == generation_step of the reference after prefill (:1666,1741)
    while not all(finished) and len(tokens) < max_new_tokens:
        logits = step_fn(tok, step)
        if on_logits is not None:
            on_logits(logits)
        step += 1
        tok = pick(logits, step)
        update(tok)
        tokens.append(tok)
    return tokens, step


def generate(W, cfg: TTSCfg, inputs_embeds: List[torch.Tensor], trailing_text: List[torch.Tensor],
             tts_pad_embed: torch.Tensor, sp: SamplingCfg, record_logits=False,
             forced_codes: Optional[np.ndarray] = None) -> GenResult:
    """Restates `talker.generate(...)` as called at modeling_qwen3_tts.py:2272-2278 plus the post-trim
    :2280-2290, for a batch given as per-row *unpadded* prefill embeddings (the left-pad + mask of
    :2239-2254 is rebuilt here exactly).

    inputs_embeds[i]: (L_i,H); trailing_text[i]: (Tt_i,H); tts_pad_embed: (H,) or (1,1,H).
    forced_codes: optional (B, n_frames, 16) teacher forcing (tokens recorded, then overwritten).
    """
    B = len(inputs_embeds)
    H = cfg.talker.hidden_size
    dt = inputs_embeds[0].dtype
    lens = torch.tensor([e.shape[0] for e in inputs_embeds])
    Lmax = int(lens.max())
    x = torch.zeros(B, Lmax, H, dtype=dt)
    valid = torch.zeros(B, Lmax, dtype=torch.bool)
    for i, e in enumerate(inputs_embeds):  # left padding (:2239-2254)
        x[i, Lmax - e.shape[0]:] = e
        valid[i, Lmax - e.shape[0]:] = True
    tts_pad = tts_pad_embed.reshape(1, 1, H).to(dt)
    Tt = max(t.shape[0] for t in trailing_text)
    trail = tts_pad.expand(B, Tt, H).clone()  # right-pad with tts_pad (:2255-2269)
    for i, t in enumerate(trailing_text):
        trail[i, :t.shape[0]] = t
    # prefill positions (:1794-1796): cumsum(mask)-1, masked -> 1
    pos = valid.long().cumsum(-1) - 1
    pos = pos.masked_fill(~valid, 1)
    cache = [None] * cfg.talker.num_layers
    mask = causal_mask(valid, torch.arange(Lmax), dt)
    h = stack_forward(W, "talker.model", cfg.talker, x, pos, mask, cache)
    past_hidden = h[:, -1:, :]
    logits = F.linear(past_hidden[:, 0], W["talker.codec_head.weight"]).to(torch.float32)

    rec = {}
    lp = talker_logits_processors(cfg, sp)
    if sp.suppress_eos:
        lp = dict(lp, min_new_tokens=1 << 30)
    frames = []  # list of (B,16)
    eos = cfg.codec_eos_token_id
    state = dict(past_hidden=past_hidden, valid_kv=valid.clone())

    def on_logits(lg):
        if record_logits:
            rec.setdefault("talker_logits", []).append(lg.numpy().copy())

    def forced_c0(idx):
        if forced_codes is not None and idx < forced_codes.shape[1]:
            return torch.from_numpy(forced_codes[:, idx, 0].astype(np.int64))
        return None

    def step_fn(c0, step):
        """One decode step: emits the 16 codes of frame `step` (:1669-1692) and returns the next codebook-0 logits."""
        forced_rest = None
        if forced_codes is not None and step < forced_codes.shape[1]:
            forced_rest = forced_codes[:, step, 1:].astype(np.int64)
        rest = code_predictor_frame(W, cfg, state["past_hidden"], c0, sp, step, forced=forced_rest,
                                    record=rec if record_logits else None)
        codes16 = torch.cat((c0[:, None], rest), dim=1)
        frames.append(codes16)
        xe = embed_frame(W, cfg, codes16)
        if step < Tt:
            xe = xe + trail[:, step].unsqueeze(1)
        else:
            xe = xe + tts_pad
        # positions (:1699-1711): cache_position + rope_deltas = len_i + step
        posd = (lens + step)[:, None]
        state["valid_kv"] = torch.cat((state["valid_kv"], torch.ones(B, 1, dtype=torch.bool)), dim=1)
        ctx = state["valid_kv"].shape[1]
        m = causal_mask(state["valid_kv"], torch.tensor([ctx - 1]), dt)
        hh = stack_forward(W, "talker.model", cfg.talker, xe, posd, m, cache)
        state["past_hidden"] = hh[:, -1:, :]
        return F.linear(state["past_hidden"][:, 0], W["talker.codec_head.weight"]).to(torch.float32)

    _, step = hf_sample_loop(logits, step_fn, B, lp, sp.do_sample, sp.seed, sp.max_new_tokens, eos, forced_c0, on_logits)
    # post-trim (:2280-2290)
    out = []
    if frames:
        allc = torch.stack(frames, dim=1)  # (B,N,16)
        for b in range(B):
            is_stop = (allc[b, :, 0] == eos)
            n = int(torch.argmax(is_stop.int())) if bool(is_stop.any()) else allc.shape[1]
            out.append(allc[b, :n])
    else:
        out = [torch.zeros(0, cfg.num_code_groups, dtype=torch.int64) for _ in range(B)]
    return GenResult(codes=out, steps=step, record=rec)


# ----------------------------------------------------------------------------------------------
# random weights of a given shape (no checkpoints offline: SURVEY §0 F5)
# ----------------------------------------------------------------------------------------------
def random_weights(cfg: TTSCfg, seed=0, dtype=torch.float32, std=0.02, with_text=True, text_vocab=None):
    g = torch.Generator().manual_seed(seed)
    W = {}

    def lin(name, out_f, in_f, bias=False, s=None):
        W[name + ".weight"] = (torch.randn(out_f, in_f, generator=g) * (s or std)).to(dtype)
        if bias:
            W[name + ".bias"] = (torch.randn(out_f, generator=g) * 0.01).to(dtype)

    def norm(name, n):
        W[name + ".weight"] = (1.0 + 0.1 * torch.randn(n, generator=g)).to(dtype)

    def stack(pfx, c: StackCfg):
        for i in range(c.num_layers):
            p = f"{pfx}.layers.{i}"
            lin(f"{p}.self_attn.q_proj", c.num_heads * c.head_dim, c.hidden_size)
            lin(f"{p}.self_attn.k_proj", c.num_kv_heads * c.head_dim, c.hidden_size)
            lin(f"{p}.self_attn.v_proj", c.num_kv_heads * c.head_dim, c.hidden_size)
            lin(f"{p}.self_attn.o_proj", c.hidden_size, c.num_heads * c.head_dim)
            norm(f"{p}.self_attn.q_norm", c.head_dim)
            norm(f"{p}.self_attn.k_norm", c.head_dim)
            lin(f"{p}.mlp.gate_proj", c.intermediate_size, c.hidden_size)
            lin(f"{p}.mlp.up_proj", c.intermediate_size, c.hidden_size)
            lin(f"{p}.mlp.down_proj", c.hidden_size, c.intermediate_size)
            norm(f"{p}.input_layernorm", c.hidden_size)
            norm(f"{p}.post_attention_layernorm", c.hidden_size)
        norm(f"{pfx}.norm", c.hidden_size)

    t, c = cfg.talker, cfg.cp
    stack("talker.model", t)
    W["talker.model.codec_embedding.weight"] = (torch.randn(t.vocab_size, t.hidden_size, generator=g) * std).to(dtype)
    lin("talker.codec_head", t.vocab_size, t.hidden_size, s=0.05)
    stack("talker.code_predictor.model", c)
    for j in range(cfg.num_code_groups - 1):
        W[f"talker.code_predictor.model.codec_embedding.{j}.weight"] = \
            (torch.randn(c.vocab_size, t.hidden_size, generator=g) * std).to(dtype)
        lin(f"talker.code_predictor.lm_head.{j}", c.vocab_size, c.hidden_size, s=0.05)
    if c.hidden_size != t.hidden_size:
        lin("talker.code_predictor.small_to_mtp_projection", c.hidden_size, t.hidden_size, bias=True)
    if with_text:
        tv = text_vocab or cfg.text_vocab_size
        W["talker.model.text_embedding.weight"] = (torch.randn(tv, cfg.text_hidden_size, generator=g) * std).to(dtype)
        lin("talker.text_projection.linear_fc1", cfg.text_hidden_size, cfg.text_hidden_size, bias=True)
        lin("talker.text_projection.linear_fc2", t.hidden_size, cfg.text_hidden_size, bias=True)
    return W


def cfg_1p7b() -> TTSCfg:
    """Expected shipped 12Hz-1.7B shapes (SURVEY §8, App. B.2 — unverified offline)."""
    return TTSCfg(talker=StackCfg(2048, 28, 16, 8, 128, 6144, 3072),
                  cp=StackCfg(1024, 5, 16, 8, 128, 3072, 2048))


def cfg_0p6b() -> TTSCfg:
    return TTSCfg(talker=StackCfg(1024, 28, 16, 8, 128, 3072, 3072),
                  cp=StackCfg(1024, 5, 16, 8, 128, 3072, 2048), text_hidden_size=2048)


def cfg_tiny(vocab=3072) -> TTSCfg:
    """Small shapes for fast unit parity (same head_dim/groups so every code path is exercised)."""
    return TTSCfg(talker=StackCfg(256, 3, 4, 2, 128, 512, vocab),
                  cp=StackCfg(128, 2, 4, 2, 128, 256, 2048),
                  text_hidden_size=256, text_vocab_size=1000,
                  tts_bos_token_id=997, tts_eos_token_id=998, tts_pad_token_id=996)
