"""CPU oracle: Qwen3-TTS-Tokenizer-12Hz codec ENCODER (24 kHz waveform -> 16 x 12.5 Hz codes).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference's encoder IS the third-party `transformers` MimiModel with its decoder half removed
(qwen_tts/core/tokenizer_12hz/modeling_qwen3_tts_tokenizer_v2.py:899-908), called at :981-983:
all 32 quantisers are computed and the first `encoder_valid_num_quantizers` = 16 kept, then each row is
trimmed to ceil(valid_samples / 1920) frames.  The arithmetic therefore lives in
transformers/models/mimi/modeling_mimi.py (reference pins transformers==4.57.3, pyproject.toml:23; the
copy installed here is 5.5.0 — this restatement is pinned against THAT copy in
tests/test_oracle_vs_reference.py::test_mimi_encoder_matches_hf, and cites its line numbers):

  MimiConv1d            :214-351   causal left pad k_eff - stride (+ right pad to a stride multiple); pad_mode
  MimiResnetBlock       :412-451   x + conv_k1(ELU(conv_k3(ELU(x))))           (compress 2, identity shortcut)
  MimiEncoder (SEANet)  :454-496   conv k7, 4 x [resblock, ELU, strided conv k=2r s=r] (r = 4,5,6,8), ELU, conv k3
  MimiTransformerLayer  :926-993   LayerNorm -> MHA (RoPE, causal, sliding window 250) -> LayerScale -> +res ->
                                   LayerNorm -> fc1/GELU(erf)/fc2 -> LayerScale -> +res     (no final norm)
  downsample            :1420-1430 conv k4 s2, no bias, pad_mode "replicate" (25 Hz -> 12.5 Hz)
  Split RVQ encode      :1176-1338 semantic (1 layer) and acoustic (n-1 layers) chains, each with its own 1x1
                                   input_proj; nearest centroid by Euclidean distance, residual update

Weights: a flat dict keyed by MimiModel's own state_dict names (fp32).
"""
import math
from dataclasses import dataclass, field
from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F


@dataclass
class MimiEncCfg:
    """transformers MimiConfig defaults (the reference builds MimiConfig(**encoder_config) with an empty dict by
    default, configuration_qwen3_tts_tokenizer_v2.py:155-162)."""
    num_filters: int = 64
    kernel_size: int = 7
    last_kernel_size: int = 3
    residual_kernel_size: int = 3
    compress: int = 2
    ratios: Tuple[int, ...] = (4, 5, 6, 8)        # reversed(upsampling_ratios = [8, 6, 5, 4])
    hidden_size: int = 512
    num_layers: int = 8
    num_heads: int = 8
    head_dim: int = 64
    intermediate_size: int = 2048
    sliding_window: int = 250
    rope_theta: float = 10000.0
    norm_eps: float = 1e-5
    codebook_size: int = 2048
    codebook_dim: int = 256
    num_quantizers: int = 32
    num_semantic_quantizers: int = 1
    valid_num_quantizers: int = 16                 # encoder_valid_num_quantizers (:147)
    downsample_stride: int = 2                     # encodec_frame_rate 25 / frame_rate 12.5
    encode_downsample_rate: int = 1920

    @property
    def hop(self):
        return int(math.prod(self.ratios) * self.downsample_stride)

    def to_hf_kwargs(self):
        return dict(num_filters=self.num_filters, kernel_size=self.kernel_size, last_kernel_size=self.last_kernel_size,
                    residual_kernel_size=self.residual_kernel_size, compress=self.compress,
                    upsampling_ratios=list(reversed(self.ratios)), hidden_size=self.hidden_size,
                    num_hidden_layers=self.num_layers, num_attention_heads=self.num_heads,
                    num_key_value_heads=self.num_heads, head_dim=self.head_dim,
                    intermediate_size=self.intermediate_size, sliding_window=self.sliding_window,
                    norm_eps=self.norm_eps, codebook_size=self.codebook_size, codebook_dim=self.codebook_dim,
                    vector_quantization_hidden_dimension=self.codebook_dim, num_quantizers=self.num_quantizers,
                    num_semantic_quantizers=self.num_semantic_quantizers, upsample_groups=self.hidden_size)


def cfg_tiny_encoder() -> MimiEncCfg:
    return MimiEncCfg(num_filters=8, hidden_size=64, num_layers=2, num_heads=4, head_dim=16, intermediate_size=96,
                      sliding_window=6, codebook_size=64, codebook_dim=32, num_quantizers=32, valid_num_quantizers=16)


# ----------------------------------------------------------------------------------------------
def mimi_conv1d(x, w, b, stride=1, dilation=1, pad_mode="constant"):
    """MimiConv1d.forward (:331-351), causal branch: left pad k_eff - stride, right pad up to a stride multiple."""
    k_eff = (w.shape[-1] - 1) * dilation + 1
    pad_total = k_eff - stride
    length = x.shape[-1]
    n_frames = (length - k_eff + pad_total) / stride + 1
    ideal = (math.ceil(n_frames) - 1) * stride + (k_eff - pad_total)
    extra = ideal - length
    x = F.pad(x, (pad_total, extra), mode=pad_mode)
    return F.conv1d(x, w, b, stride=stride, dilation=dilation)


def seanet_encoder(W: Dict[str, torch.Tensor], cfg: MimiEncCfg, x, stages=None):
    """MimiEncoder.forward (:454-496).  x: (B, 1, T) -> (B, hidden, T / prod(ratios))."""
    def conv(name, x, **kw):
        return mimi_conv1d(x, W[f"encoder.layers.{name}.conv.weight"], W.get(f"encoder.layers.{name}.conv.bias"), **kw)

    idx = 0
    h = conv(f"{idx}", x)
    if stages is not None:
        stages.append(("conv0", h))
    idx += 1
    for si, r in enumerate(cfg.ratios):
        # residual block (num_residual_layers = 1 -> dilation 1), identity shortcut (:412-451)
        y = conv(f"{idx}.block.1", F.elu(h))
        y = conv(f"{idx}.block.3", F.elu(y))
        h = h + y
        if stages is not None:
            stages.append((f"res{si}", h))
        idx += 1
        idx += 1  # the nn.ELU() module occupies an index
        h = conv(f"{idx}", F.elu(h), stride=r)
        if stages is not None:
            stages.append((f"down{si}", h))
        idx += 1
    idx += 1  # ELU
    h = conv(f"{idx}", F.elu(h))
    if stages is not None:
        stages.append(("conv_last", h))
    return h


def rope_tables(cfg: MimiEncCfg, T: int):
    """MimiRotaryEmbedding (:515-578): fp32 inv_freq x positions, emb = cat(freqs, freqs)."""
    inv = 1.0 / (cfg.rope_theta ** (torch.arange(0, cfg.head_dim, 2, dtype=torch.int64).float() / cfg.head_dim))
    fr = torch.arange(T).float()[:, None] * inv[None, :]
    return fr.cos(), fr.sin()  # (T, head_dim/2)


def encoder_transformer(W, cfg: MimiEncCfg, x, stages=None):
    """MimiTransformerModel.forward (:1015-1140) without cache.  x: (B, T, hidden)."""
    B, T, C = x.shape
    nh, hd = cfg.num_heads, cfg.head_dim
    cos, sin = rope_tables(cfg, T)
    cos = torch.cat([cos, cos], -1)[None, None]
    sin = torch.cat([sin, sin], -1)[None, None]
    i = torch.arange(T)[:, None]
    j = torch.arange(T)[None, :]
    allowed = (j <= i) & (i - j < cfg.sliding_window)  # create_sliding_window_causal_mask
    bias = torch.zeros(T, T).masked_fill(~allowed, float("-inf"))

    def rot(t):
        return torch.cat([-t[..., hd // 2:], t[..., : hd // 2]], -1)

    for l in range(cfg.num_layers):
        p = f"encoder_transformer.layers.{l}."
        h = F.layer_norm(x, (C,), W[p + "input_layernorm.weight"], W[p + "input_layernorm.bias"], cfg.norm_eps)
        q = (h @ W[p + "self_attn.q_proj.weight"].T).view(B, T, nh, hd).transpose(1, 2)
        k = (h @ W[p + "self_attn.k_proj.weight"].T).view(B, T, nh, hd).transpose(1, 2)
        v = (h @ W[p + "self_attn.v_proj.weight"].T).view(B, T, nh, hd).transpose(1, 2)
        q = q * cos + rot(q) * sin
        k = k * cos + rot(k) * sin
        a = torch.softmax((q @ k.transpose(2, 3)) / math.sqrt(hd) + bias, dim=-1, dtype=torch.float32)
        o = (a @ v).transpose(1, 2).reshape(B, T, nh * hd) @ W[p + "self_attn.o_proj.weight"].T
        x = x + W[p + "self_attn_layer_scale.scale"] * o
        h = F.layer_norm(x, (C,), W[p + "post_attention_layernorm.weight"], W[p + "post_attention_layernorm.bias"],
                         cfg.norm_eps)
        h = F.gelu(h @ W[p + "mlp.fc1.weight"].T) @ W[p + "mlp.fc2.weight"].T
        x = x + W[p + "mlp_layer_scale.scale"] * h
        if stages is not None:
            stages.append((f"tr{l}", x.transpose(1, 2)))
    return x


def codebook_embed(W, prefix):
    """MimiEuclideanCodebook.embed (:1192-1195)."""
    return W[prefix + "embed_sum"] / W[prefix + "cluster_usage"].clamp(min=1e-5)[:, None]


def rvq_encode(W, cfg: MimiEncCfg, emb, n_q: int, margins=None):
    """MimiSplitResidualVectorQuantizer.encode (:1311-1338).  emb: (B, hidden, T) -> codes (B, n_q, T).
    `margins` (optional list) receives, per quantiser, the gap between the best and second-best squared distance."""
    out = []
    for which, n in (("semantic", cfg.num_semantic_quantizers), ("acoustic", n_q - cfg.num_semantic_quantizers)):
        if n <= 0:
            continue
        p = f"quantizer.{which}_residual_vector_quantizer."
        r = F.conv1d(emb, W[p + "input_proj.weight"])          # (B, D, T)
        for qi in range(n):
            E = codebook_embed(W, p + f"layers.{qi}.codebook.")  # (K, D)
            x = r.transpose(1, 2).reshape(-1, E.shape[1])       # (B*T, D)
            d = torch.cdist(x[None].float(), E[None].float(), p=2)[0]
            ind = d.argmin(-1)
            if margins is not None:
                top2 = torch.topk(d * d, 2, dim=-1, largest=False).values
                margins.append((top2[:, 1] - top2[:, 0]).view(emb.shape[0], -1))
            out.append(ind.view(emb.shape[0], -1))
            r = r - E[ind].view(emb.shape[0], -1, E.shape[1]).transpose(1, 2)
    return torch.stack(out, 1)


@torch.no_grad()
def encode(W, cfg: MimiEncCfg, wav: torch.Tensor, n_q: int = None, stages: List = None, margins: List = None):
    """MimiModel._encode_frame (:1455-1488) on (B, T) fp32 waveforms -> codes (B, n_q, ceil(T / hop))."""
    n_q = cfg.valid_num_quantizers if n_q is None else n_q
    x = wav[:, None, :].float()
    h = seanet_encoder(W, cfg, x, stages)
    h = encoder_transformer(W, cfg, h.transpose(1, 2), stages).transpose(1, 2)
    h = mimi_conv1d(h, W["downsample.conv.weight"], None, stride=cfg.downsample_stride, pad_mode="replicate")
    if stages is not None:
        stages.append(("downsample", h))
    return rvq_encode(W, cfg, h, n_q, margins)


@torch.no_grad()
def tokenizer_encode(W, cfg: MimiEncCfg, wavs: List[torch.Tensor], n_q: int = None):
    """Qwen3TTSTokenizerV2Model.encode (…v2.py:961-991): right-pad to the longest, encode, keep the first 16
    quantisers, trim row i to ceil(len_i / encode_downsample_rate) frames; returns a list of (T_i, n_q) int64."""
    L = max(int(w.shape[0]) for w in wavs)
    x = torch.zeros(len(wavs), L)
    for i, w in enumerate(wavs):
        x[i, : w.shape[0]] = w
    codes = encode(W, cfg, x, n_q)
    return [codes[i, :, : -(-int(w.shape[0]) // cfg.encode_downsample_rate)].transpose(0, 1).contiguous()
            for i, w in enumerate(wavs)]


def random_weights(cfg: MimiEncCfg, seed=0) -> Dict[str, torch.Tensor]:
    """Seeded weights with MimiModel's state_dict names/shapes (encoder half only)."""
    g = torch.Generator().manual_seed(seed)
    W = {}

    def rn(*shape, s=1.0):
        return torch.randn(*shape, generator=g) * s

    def conv(name, cin, cout, k, bias=True):
        W[f"{name}.conv.weight"] = rn(cout, cin, k, s=1.0 / math.sqrt(cin * k))
        if bias:
            W[f"{name}.conv.bias"] = rn(cout, s=0.05)

    idx = 0
    conv(f"encoder.layers.{idx}", 1, cfg.num_filters, cfg.kernel_size)
    idx += 1
    dim = cfg.num_filters
    for r in cfg.ratios:
        conv(f"encoder.layers.{idx}.block.1", dim, dim // cfg.compress, cfg.residual_kernel_size)
        conv(f"encoder.layers.{idx}.block.3", dim // cfg.compress, dim, 1)
        idx += 2
        conv(f"encoder.layers.{idx}", dim, dim * 2, 2 * r)
        idx += 1
        dim *= 2
    idx += 1
    conv(f"encoder.layers.{idx}", dim, cfg.hidden_size, cfg.last_kernel_size)
    C, I = cfg.hidden_size, cfg.intermediate_size
    for l in range(cfg.num_layers):
        p = f"encoder_transformer.layers.{l}."
        for n in ("q_proj", "k_proj", "v_proj", "o_proj"):
            W[p + f"self_attn.{n}.weight"] = rn(C, C, s=1.0 / math.sqrt(C))
        W[p + "mlp.fc1.weight"] = rn(I, C, s=1.0 / math.sqrt(C))
        W[p + "mlp.fc2.weight"] = rn(C, I, s=1.0 / math.sqrt(I))
        for n in ("input_layernorm", "post_attention_layernorm"):
            W[p + n + ".weight"] = 1.0 + rn(C, s=0.1)
            W[p + n + ".bias"] = rn(C, s=0.05)
        W[p + "self_attn_layer_scale.scale"] = 0.3 + rn(C, s=0.05)
        W[p + "mlp_layer_scale.scale"] = 0.3 + rn(C, s=0.05)
    W["downsample.conv.weight"] = rn(C, C, 2 * cfg.downsample_stride, s=1.0 / math.sqrt(C * 4))
    D, K = cfg.codebook_dim, cfg.codebook_size
    for which, n in (("semantic", cfg.num_semantic_quantizers),
                     ("acoustic", cfg.num_quantizers - cfg.num_semantic_quantizers)):
        p = f"quantizer.{which}_residual_vector_quantizer."
        W[p + "input_proj.weight"] = rn(D, C, 1, s=1.0 / math.sqrt(C))
        for qi in range(n):
            usage = torch.rand(K, generator=g) * 3 + 0.5
            W[p + f"layers.{qi}.codebook.cluster_usage"] = usage
            # centroid scale shrinks with depth like a trained RVQ, so that deeper levels stay informative
            W[p + f"layers.{qi}.codebook.embed_sum"] = rn(K, D, s=0.8 * (0.75 ** qi)) * usage[:, None]
    return W
