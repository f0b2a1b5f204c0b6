"""CPU oracle: Qwen3-TTS-Tokenizer-12Hz codec DECODER (codes -> 24 kHz waveform).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Restates
qwen_tts/core/tokenizer_12hz/modeling_qwen3_tts_tokenizer_v2.py; weights are a flat dict keyed by the
reference decoder's own state_dict names (e.g. `decoder.1.block.2.conv1.conv.weight`).
"""
import math
from dataclasses import dataclass
from typing import Tuple

import torch
import torch.nn.functional as F

from .talker import rms_norm, apply_rope, rope_cos_sin


@dataclass
class CodecCfg:
    """configuration_qwen3_tts_tokenizer_v2.py:72-93 defaults (+ codebook_dim, which has no default in code,
    modeling…v2.py:831-836; 512 expected — SURVEY App. B.2)."""
    codebook_size: int = 2048
    codebook_dim: int = 512
    hidden_size: int = 1024
    latent_dim: int = 1024
    rope_theta: float = 10000.0
    num_heads: int = 16
    num_kv_heads: int = 16
    head_dim: int = 64
    sliding_window: int = 72
    intermediate_size: int = 3072
    rms_eps: float = 1e-5
    num_layers: int = 8
    num_quantizers: int = 16
    upsample_rates: Tuple[int, ...] = (8, 5, 4, 3)
    upsampling_ratios: Tuple[int, ...] = (2, 2)
    decoder_dim: int = 1536

    @property
    def total_upsample(self):
        return int(math.prod(self.upsample_rates) * math.prod(self.upsampling_ratios))

    def to_reference_kwargs(self):
        return dict(codebook_size=self.codebook_size, codebook_dim=self.codebook_dim, hidden_size=self.hidden_size,
                    latent_dim=self.latent_dim, rope_theta=self.rope_theta, num_attention_heads=self.num_heads,
                    num_key_value_heads=self.num_kv_heads, head_dim=self.head_dim,
                    sliding_window=self.sliding_window, intermediate_size=self.intermediate_size,
                    rms_norm_eps=self.rms_eps, num_hidden_layers=self.num_layers,
                    num_quantizers=self.num_quantizers, upsample_rates=tuple(self.upsample_rates),
                    upsampling_ratios=tuple(self.upsampling_ratios), decoder_dim=self.decoder_dim)


def cfg_tiny_codec() -> CodecCfg:
    return CodecCfg(codebook_size=64, codebook_dim=32, hidden_size=64, latent_dim=64, num_heads=4, num_kv_heads=4,
                    head_dim=16, sliding_window=6, intermediate_size=96, num_layers=2, num_quantizers=16,
                    upsample_rates=(8, 5, 4, 3), upsampling_ratios=(2, 2), decoder_dim=96)


# ----------------------------------------------------------------------------------------------
def causal_conv1d(x, w, b, dilation=1, stride=1, groups=1):
    """…v2.py:159-192 — left pad (k-1)*d+1-stride, right 'extra' pad so the length is stride-aligned."""
    k_eff = (w.shape[-1] - 1) * dilation + 1
    padding = k_eff - stride
    length = x.shape[-1]
    n_frames = (length - k_eff + padding) / stride + 1
    ideal = (math.ceil(n_frames) - 1) * stride + (k_eff - padding)
    extra = ideal - length
    x = F.pad(x, (padding, extra))
    return F.conv1d(x, w, b, stride=stride, dilation=dilation, groups=groups)


def causal_conv_transpose1d(x, w, b, stride):
    """…v2.py:195-208 — ConvTranspose1d then drop the last (k - stride) samples."""
    y = F.conv_transpose1d(x, w, b, stride=stride)
    rp = w.shape[-1] - stride
    return y[..., : y.shape[-1] - rp] if rp > 0 else y


def snake_beta(x, alpha, beta):
    """…v2.py:602-616 — x + 1/(exp(beta)+1e-9) * sin^2(x*exp(alpha)), per channel."""
    a = torch.exp(alpha)[None, :, None]
    b = torch.exp(beta)[None, :, None]
    return x + (1.0 / (b + 1e-9)) * torch.pow(torch.sin(x * a), 2)


def convnext_block(W, p, x):
    """…v2.py:227-243."""
    inp = x
    C = x.shape[1]
    h = causal_conv1d(x, W[f"{p}.dwconv.conv.weight"], W[f"{p}.dwconv.conv.bias"], groups=C)
    h = h.permute(0, 2, 1)
    h = F.layer_norm(h, (C,), W[f"{p}.norm.weight"], W[f"{p}.norm.bias"], eps=1e-6)
    h = F.linear(h, W[f"{p}.pwconv1.weight"], W[f"{p}.pwconv1.bias"])
    h = F.gelu(h)
    h = F.linear(h, W[f"{p}.pwconv2.weight"], W[f"{p}.pwconv2.bias"])
    h = W[f"{p}.gamma"] * h
    return inp + h.permute(0, 2, 1)


def rvq_decode(W, cfg: CodecCfg, codes):
    """…v2.py:815-821 -> :773-777 -> :721-727 -> :707-711 -> :676-679.  codes: (B,K,T) -> (B,codebook_dim,T)."""
    def one(pfx, cc):  # cc: (B,k,T)
        q = None
        for i in range(cc.shape[1]):
            es = W[f"{pfx}.vq.layers.{i}._codebook.embedding_sum"]
            cu = W[f"{pfx}.vq.layers.{i}._codebook.cluster_usage"]
            emb = es / cu.clamp(min=1e-5)[:, None]
            e = F.embedding(cc[:, i], emb).transpose(1, 2)  # (B,dim,T)
            q = e if q is None else q + e
        return F.conv1d(q, W[f"{pfx}.output_proj.weight"])
    out = one("quantizer.rvq_first", codes[:, :1])
    if codes.shape[1] > 1:
        out = out + one("quantizer.rvq_rest", codes[:, 1:])
    return out


def pre_transformer(W, cfg: CodecCfg, x):
    """…v2.py:501-575 (model), :450-472 (layer), :321-354 (attention, no q/k norm), sliding causal mask
    (HF create_sliding_window_causal_mask: key k visible to query q iff 0 <= q-k < window).  x: (B,T,latent)."""
    B, T, _ = x.shape
    p = "pre_transformer"
    x = F.linear(x, W[f"{p}.input_proj.weight"], W[f"{p}.input_proj.bias"])
    pos = torch.arange(T)[None].expand(B, -1)
    cos, sin = rope_cos_sin(pos, cfg.head_dim, cfg.rope_theta, x.dtype)
    qi = torch.arange(T)[:, None]
    ki = torch.arange(T)[None, :]
    allowed = (ki <= qi) & (qi - ki < cfg.sliding_window)
    mask = torch.zeros(T, T, dtype=x.dtype).masked_fill(~allowed, torch.finfo(x.dtype).min)[None, None]
    nrep = cfg.num_heads // cfg.num_kv_heads
    for i in range(cfg.num_layers):
        lp = f"{p}.layers.{i}"
        res = x
        h = rms_norm(x, W[f"{lp}.input_layernorm.weight"], cfg.rms_eps)
        q = F.linear(h, W[f"{lp}.self_attn.q_proj.weight"]).view(B, T, -1, cfg.head_dim).transpose(1, 2)
        k = F.linear(h, W[f"{lp}.self_attn.k_proj.weight"]).view(B, T, -1, cfg.head_dim).transpose(1, 2)
        v = F.linear(h, W[f"{lp}.self_attn.v_proj.weight"]).view(B, T, -1, cfg.head_dim).transpose(1, 2)
        q, k = apply_rope(q, k, cos, sin)
        if nrep > 1:
            k = k.repeat_interleave(nrep, dim=1)
            v = v.repeat_interleave(nrep, dim=1)
        w = torch.matmul(q, k.transpose(2, 3)) * (cfg.head_dim ** -0.5) + mask
        w = F.softmax(w, dim=-1, dtype=torch.float32).to(q.dtype)
        a = torch.matmul(w, v).transpose(1, 2).reshape(B, T, -1)
        a = F.linear(a, W[f"{lp}.self_attn.o_proj.weight"])
        x = res + W[f"{lp}.self_attn_layer_scale.scale"] * a
        res = x
        h = rms_norm(x, W[f"{lp}.post_attention_layernorm.weight"], cfg.rms_eps)
        h = F.linear(F.silu(F.linear(h, W[f"{lp}.mlp.gate_proj.weight"])) * F.linear(h, W[f"{lp}.mlp.up_proj.weight"]),
                     W[f"{lp}.mlp.down_proj.weight"])
        x = res + W[f"{lp}.mlp_layer_scale.scale"] * h
    x = rms_norm(x, W[f"{p}.norm.weight"], cfg.rms_eps)
    return F.linear(x, W[f"{p}.output_proj.weight"], W[f"{p}.output_proj.bias"])


def decoder_forward(W, cfg: CodecCfg, codes, taps=None):
    """…v2.py:869-884.  codes: (B,K,T) int64 -> wav (B,1,T*total_upsample) in [-1,1]."""
    if codes.shape[1] != cfg.num_quantizers:
        raise ValueError(f"Expected {cfg.num_quantizers} layer of codes, got {codes.shape[1]}")
    h = rvq_decode(W, cfg, codes)
    if taps is not None:
        taps["rvq"] = h
    h = causal_conv1d(h, W["pre_conv.conv.weight"], W["pre_conv.conv.bias"]).transpose(1, 2)
    if taps is not None:
        taps["pre_conv"] = h
    h = pre_transformer(W, cfg, h).permute(0, 2, 1)
    if taps is not None:
        taps["pre_transformer"] = h
    for i, f in enumerate(cfg.upsampling_ratios):
        h = causal_conv_transpose1d(h, W[f"upsample.{i}.0.conv.weight"], W[f"upsample.{i}.0.conv.bias"], f)
        h = convnext_block(W, f"upsample.{i}.1", h)
    if taps is not None:
        taps["upsample"] = h
    w = causal_conv1d(h, W["decoder.0.conv.weight"], W["decoder.0.conv.bias"])
    if taps is not None:
        taps["decoder0"] = w
    for i, r in enumerate(cfg.upsample_rates):
        p = f"decoder.{i + 1}.block"
        w = snake_beta(w, W[f"{p}.0.alpha"], W[f"{p}.0.beta"])
        w = causal_conv_transpose1d(w, W[f"{p}.1.conv.weight"], W[f"{p}.1.conv.bias"], r)
        for u, dil in enumerate((1, 3, 9)):
            q = f"{p}.{u + 2}"
            res = w
            w = snake_beta(w, W[f"{q}.act1.alpha"], W[f"{q}.act1.beta"])
            w = causal_conv1d(w, W[f"{q}.conv1.conv.weight"], W[f"{q}.conv1.conv.bias"], dilation=dil)
            w = snake_beta(w, W[f"{q}.act2.alpha"], W[f"{q}.act2.beta"])
            w = causal_conv1d(w, W[f"{q}.conv2.conv.weight"], W[f"{q}.conv2.conv.bias"])
            w = w + res
        if taps is not None:
            taps[f"block{i}"] = w
    n = len(cfg.upsample_rates)
    w = snake_beta(w, W[f"decoder.{n + 1}.alpha"], W[f"decoder.{n + 1}.beta"])
    w = causal_conv1d(w, W[f"decoder.{n + 2}.conv.weight"], W[f"decoder.{n + 2}.conv.bias"])
    return w.clamp(min=-1, max=1)


def chunked_decode(W, cfg: CodecCfg, codes, chunk_size=300, left_context_size=25):
    """…v2.py:886-896 — NOT equal to a full forward after the first chunk; replicated as is."""
    wavs = []
    start = 0
    while start < codes.shape[-1]:
        end = min(start + chunk_size, codes.shape[-1])
        ctx = left_context_size if start - left_context_size > 0 else start
        wav = decoder_forward(W, cfg, codes[..., start - ctx:end])
        wavs.append(wav[..., ctx * cfg.total_upsample:])
        start = end
    return torch.cat(wavs, dim=-1)


def decode(W, cfg: CodecCfg, audio_codes):
    """Qwen3TTSTokenizerV2Model.decode, …v2.py:993-1024.  audio_codes: (B,T,K) padded with -1."""
    lengths = (audio_codes[..., 0] > -1).sum(1) * cfg.total_upsample
    codes = torch.clamp(audio_codes, min=0)
    wav = chunked_decode(W, cfg, codes.transpose(1, 2)).squeeze(1)
    return [a[:l] for a, l in zip(wav, lengths)]


# ----------------------------------------------------------------------------------------------
def random_weights(cfg: CodecCfg, seed=0, dtype=torch.float32):
    """Seeded weights with the reference decoder's state_dict names/shapes; scales chosen so that the
    activations stay O(1) through the ~1920x upsampling stack (no checkpoints offline)."""
    g = torch.Generator().manual_seed(seed)
    W = {}

    def rn(*shape, s=1.0):
        return (torch.randn(*shape, generator=g) * s).to(dtype)

    def conv(name, cout, cin, k, groups=1, transposed=False):
        fan = (cin // groups) * k
        shape = (cin, cout, k) if transposed else (cout, cin // groups, k)
        if transposed:
            fan = cin * max(1, k // 2)
        W[f"{name}.weight"] = rn(*shape, s=1.0 / math.sqrt(fan))
        W[f"{name}.bias"] = rn(cout, s=0.02)

    half = cfg.codebook_dim // 2
    for pfx, n in (("quantizer.rvq_first", 1), ("quantizer.rvq_rest", cfg.num_quantizers - 1)):
        for i in range(n):
            W[f"{pfx}.vq.layers.{i}._codebook.embedding_sum"] = rn(cfg.codebook_size, half, s=0.5)
            W[f"{pfx}.vq.layers.{i}._codebook.cluster_usage"] = (torch.rand(cfg.codebook_size, generator=g) + 0.5).to(dtype)
        W[f"{pfx}.input_proj.weight"] = rn(half, cfg.codebook_dim, 1, s=1 / math.sqrt(cfg.codebook_dim))
        W[f"{pfx}.output_proj.weight"] = rn(cfg.codebook_dim, half, 1, s=1 / math.sqrt(half))
    conv("pre_conv.conv", cfg.latent_dim, cfg.codebook_dim, 3)
    p = "pre_transformer"
    Hh = cfg.hidden_size

    def lin(name, o, i, bias=False):
        W[f"{name}.weight"] = rn(o, i, s=1 / math.sqrt(i))
        if bias:
            W[f"{name}.bias"] = rn(o, s=0.02)

    lin(f"{p}.input_proj", Hh, cfg.latent_dim, True)
    lin(f"{p}.output_proj", cfg.latent_dim, Hh, True)
    for i in range(cfg.num_layers):
        lp = f"{p}.layers.{i}"
        lin(f"{lp}.self_attn.q_proj", cfg.num_heads * cfg.head_dim, Hh)
        lin(f"{lp}.self_attn.k_proj", cfg.num_kv_heads * cfg.head_dim, Hh)
        lin(f"{lp}.self_attn.v_proj", cfg.num_kv_heads * cfg.head_dim, Hh)
        lin(f"{lp}.self_attn.o_proj", Hh, cfg.num_heads * cfg.head_dim)
        lin(f"{lp}.mlp.gate_proj", cfg.intermediate_size, Hh)
        lin(f"{lp}.mlp.up_proj", cfg.intermediate_size, Hh)
        lin(f"{lp}.mlp.down_proj", Hh, cfg.intermediate_size)
        W[f"{lp}.input_layernorm.weight"] = (1 + 0.1 * torch.randn(Hh, generator=g)).to(dtype)
        W[f"{lp}.post_attention_layernorm.weight"] = (1 + 0.1 * torch.randn(Hh, generator=g)).to(dtype)
        W[f"{lp}.self_attn_layer_scale.scale"] = (0.3 + 0.05 * torch.randn(Hh, generator=g)).to(dtype)
        W[f"{lp}.mlp_layer_scale.scale"] = (0.3 + 0.05 * torch.randn(Hh, generator=g)).to(dtype)
    W[f"{p}.norm.weight"] = (1 + 0.1 * torch.randn(Hh, generator=g)).to(dtype)
    C = cfg.latent_dim
    for i, f in enumerate(cfg.upsampling_ratios):
        conv(f"upsample.{i}.0.conv", C, C, f, transposed=True)
        q = f"upsample.{i}.1"
        conv(f"{q}.dwconv.conv", C, C, 7, groups=C)
        W[f"{q}.norm.weight"] = (1 + 0.1 * torch.randn(C, generator=g)).to(dtype)
        W[f"{q}.norm.bias"] = rn(C, s=0.02)
        lin(f"{q}.pwconv1", 4 * C, C, True)
        lin(f"{q}.pwconv2", C, 4 * C, True)
        W[f"{q}.gamma"] = (0.3 + 0.05 * torch.randn(C, generator=g)).to(dtype)
    conv("decoder.0.conv", cfg.decoder_dim, C, 7)
    for i, r in enumerate(cfg.upsample_rates):
        cin = cfg.decoder_dim // 2 ** i
        cout = cfg.decoder_dim // 2 ** (i + 1)
        bp = f"decoder.{i + 1}.block"
        W[f"{bp}.0.alpha"] = rn(cin, s=0.3)
        W[f"{bp}.0.beta"] = rn(cin, s=0.3)
        conv(f"{bp}.1.conv", cout, cin, 2 * r, transposed=True)
        for u in range(3):
            q = f"{bp}.{u + 2}"
            W[f"{q}.act1.alpha"] = rn(cout, s=0.3)
            W[f"{q}.act1.beta"] = rn(cout, s=0.3)
            conv(f"{q}.conv1.conv", cout, cout, 7)
            W[f"{q}.act2.alpha"] = rn(cout, s=0.3)
            W[f"{q}.act2.beta"] = rn(cout, s=0.3)
            conv(f"{q}.conv2.conv", cout, cout, 1)
            W[f"{q}.conv1.conv.weight"] *= 0.5
            W[f"{q}.conv2.conv.weight"] *= 0.5
    n = len(cfg.upsample_rates)
    cl = cfg.decoder_dim // 2 ** n
    W[f"decoder.{n + 1}.alpha"] = rn(cl, s=0.3)
    W[f"decoder.{n + 1}.beta"] = rn(cl, s=0.3)
    conv(f"decoder.{n + 2}.conv", 1, cl, 7)
    W[f"decoder.{n + 2}.conv.weight"] *= 0.3
    return W


# ----------------------------------------------------------------------------------------------
# Stateful streaming decoder (a NEW surface — the reference only has chunked_decode, which is not
# equal to the full forward after the first chunk, F9).  Every layer of the decoder is causal, so a packet can be
# decoded exactly given a bounded amount of per-layer history: this class states WHICH history (the spec the CUDA
# engine's stateful path has to implement) and tests/test_oracle_golden.py proves packets == decoder_forward(full).
#   conv k (stride 1, dilation d)      : the last (k-1)*d input samples          (zeros before the first packet = causal pad)
#   ConvTranspose k=2r, stride r       : the last input sample (overlap-add of two half kernels)
#   ConvTranspose k=stride             : nothing
#   sliding-window transformer         : K/V of the last window-1 frames per layer + the absolute position (RoPE)
# ----------------------------------------------------------------------------------------------
class StreamingDecoder:
    def __init__(self, W, cfg: CodecCfg, batch: int = 1):
        self.W, self.cfg, self.B = W, cfg, batch
        self.state = {}
        self.pos = 0  # frames consumed so far

    def _conv(self, name, x, w, b, dilation=1, groups=1):
        ctx = (w.shape[-1] - 1) * dilation
        prev = self.state.get(name)
        if prev is None:
            prev = torch.zeros(x.shape[0], x.shape[1], ctx, dtype=x.dtype)
        xx = torch.cat([prev, x], dim=-1)
        if ctx:
            self.state[name] = xx[..., xx.shape[-1] - ctx:]
        return F.conv1d(xx, w, b, dilation=dilation, groups=groups)

    def _convT(self, name, x, w, b, stride):
        k = w.shape[-1]
        if k == stride:                      # no overlap between output blocks
            return F.conv_transpose1d(x, w, b, stride=stride)
        assert k == 2 * stride, "decoder blocks use k = 2*stride (…v2.py:638-658)"
        prev = self.state.get(name)
        if prev is None:
            prev = torch.zeros(x.shape[0], x.shape[1], 1, dtype=x.dtype)
        xx = torch.cat([prev, x], dim=-1)
        self.state[name] = x[..., -1:]
        y = F.conv_transpose1d(xx, w, b, stride=stride)
        return y[..., stride: stride + x.shape[-1] * stride]

    def _transformer(self, x):
        W, cfg = self.W, self.cfg
        B, T, _ = x.shape
        p = "pre_transformer"
        x = F.linear(x, W[f"{p}.input_proj.weight"], W[f"{p}.input_proj.bias"])
        pos = (self.pos + torch.arange(T))[None].expand(B, -1)
        cos, sin = rope_cos_sin(pos, cfg.head_dim, cfg.rope_theta, x.dtype)
        nrep = cfg.num_heads // cfg.num_kv_heads
        for i in range(cfg.num_layers):
            lp = f"{p}.layers.{i}"
            res = x
            h = rms_norm(x, W[f"{lp}.input_layernorm.weight"], cfg.rms_eps)
            q = F.linear(h, W[f"{lp}.self_attn.q_proj.weight"]).view(B, T, -1, cfg.head_dim).transpose(1, 2)
            k = F.linear(h, W[f"{lp}.self_attn.k_proj.weight"]).view(B, T, -1, cfg.head_dim).transpose(1, 2)
            v = F.linear(h, W[f"{lp}.self_attn.v_proj.weight"]).view(B, T, -1, cfg.head_dim).transpose(1, 2)
            q, k = apply_rope(q, k, cos, sin)
            pk, pv = self.state.get(f"k{i}"), self.state.get(f"v{i}")
            if pk is not None:
                k, v = torch.cat([pk, k], 2), torch.cat([pv, v], 2)
            keep = cfg.sliding_window - 1
            self.state[f"k{i}"], self.state[f"v{i}"] = k[:, :, -keep:] if keep else k[:, :, :0], v[:, :, -keep:] if keep else v[:, :, :0]
            S = k.shape[2]
            kpos = self.pos + T - S + torch.arange(S)          # absolute positions of the cached + new keys
            qpos = self.pos + torch.arange(T)
            allowed = (kpos[None, :] <= qpos[:, None]) & (qpos[:, None] - kpos[None, :] < cfg.sliding_window)
            mask = torch.zeros(T, S, dtype=x.dtype).masked_fill(~allowed, torch.finfo(x.dtype).min)[None, None]
            kk, vv = (k.repeat_interleave(nrep, dim=1), v.repeat_interleave(nrep, dim=1)) if nrep > 1 else (k, v)
            w = torch.matmul(q, kk.transpose(2, 3)) * (cfg.head_dim ** -0.5) + mask
            w = F.softmax(w, dim=-1, dtype=torch.float32).to(q.dtype)
            a = torch.matmul(w, vv).transpose(1, 2).reshape(B, T, -1)
            a = F.linear(a, W[f"{lp}.self_attn.o_proj.weight"])
            x = res + W[f"{lp}.self_attn_layer_scale.scale"] * a
            res = x
            h = rms_norm(x, W[f"{lp}.post_attention_layernorm.weight"], cfg.rms_eps)
            h = F.linear(F.silu(F.linear(h, W[f"{lp}.mlp.gate_proj.weight"])) * F.linear(h, W[f"{lp}.mlp.up_proj.weight"]),
                         W[f"{lp}.mlp.down_proj.weight"])
            x = res + W[f"{lp}.mlp_layer_scale.scale"] * h
        x = rms_norm(x, W[f"{p}.norm.weight"], cfg.rms_eps)
        return F.linear(x, W[f"{p}.output_proj.weight"], W[f"{p}.output_proj.bias"])

    @torch.no_grad()
    def push(self, codes):
        """codes: (B, K, n) the next n frames -> wav (B, 1, n * total_upsample), equal to the same slice of
        decoder_forward over everything pushed so far."""
        W, cfg = self.W, self.cfg
        h = rvq_decode(W, cfg, codes)
        h = self._conv("pre_conv", h, W["pre_conv.conv.weight"], W["pre_conv.conv.bias"]).transpose(1, 2)
        h = self._transformer(h).permute(0, 2, 1)
        for i, f in enumerate(cfg.upsampling_ratios):
            h = self._convT(f"up{i}", h, W[f"upsample.{i}.0.conv.weight"], W[f"upsample.{i}.0.conv.bias"], f)
            q = f"upsample.{i}.1"
            C = h.shape[1]
            inp = h
            g = self._conv(f"dw{i}", h, W[f"{q}.dwconv.conv.weight"], W[f"{q}.dwconv.conv.bias"], groups=C).permute(0, 2, 1)
            g = F.layer_norm(g, (C,), W[f"{q}.norm.weight"], W[f"{q}.norm.bias"], eps=1e-6)
            g = F.linear(F.gelu(F.linear(g, W[f"{q}.pwconv1.weight"], W[f"{q}.pwconv1.bias"])), W[f"{q}.pwconv2.weight"],
                         W[f"{q}.pwconv2.bias"])
            h = inp + (W[f"{q}.gamma"] * g).permute(0, 2, 1)
        w = self._conv("dec0", h, W["decoder.0.conv.weight"], W["decoder.0.conv.bias"])
        for i, r in enumerate(cfg.upsample_rates):
            p = f"decoder.{i + 1}.block"
            w = snake_beta(w, W[f"{p}.0.alpha"], W[f"{p}.0.beta"])
            w = self._convT(f"ct{i}", w, W[f"{p}.1.conv.weight"], W[f"{p}.1.conv.bias"], r)
            for u, dil in enumerate((1, 3, 9)):
                q = f"{p}.{u + 2}"
                res = w
                w = snake_beta(w, W[f"{q}.act1.alpha"], W[f"{q}.act1.beta"])
                w = self._conv(f"c1_{i}_{u}", w, W[f"{q}.conv1.conv.weight"], W[f"{q}.conv1.conv.bias"], dilation=dil)
                w = snake_beta(w, W[f"{q}.act2.alpha"], W[f"{q}.act2.beta"])
                w = F.conv1d(w, W[f"{q}.conv2.conv.weight"], W[f"{q}.conv2.conv.bias"]) + res
        n = len(cfg.upsample_rates)
        w = snake_beta(w, W[f"decoder.{n + 1}.alpha"], W[f"decoder.{n + 1}.beta"])
        w = self._conv("out", w, W[f"decoder.{n + 2}.conv.weight"], W[f"decoder.{n + 2}.conv.bias"])
        self.pos += codes.shape[-1]
        return w.clamp(min=-1, max=1)

    def state_bytes(self, dtype_bytes=2):
        """History the engine has to carry per row (bf16 by default)."""
        return sum(int(v.numel()) // max(self.B, 1) * dtype_bytes for v in self.state.values())
