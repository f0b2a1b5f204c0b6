"""Build and drive the REFERENCE modules (through oracle/ref_shims.py) so the oracle can be pinned
against them in the build container.  TEST INFRASTRUCTURE ONLY; never runs on the GPU box.

The reference's HF `generate()` loops do not run under transformers 5.5.0, so the reference
modules are driven by hand with a `DynamicCache`, exactly as the survey's probes did.
"""
import torch

from . import ref_shims
from .talker import TTSCfg


def build_reference_talker(cfg: TTSCfg, text_vocab=None, dtype=torch.float32):
    ref_shims.install()
    from qwen_tts.core.models.configuration_qwen3_tts import (Qwen3TTSTalkerConfig,
                                                              Qwen3TTSTalkerCodePredictorConfig)
    from qwen_tts.core.models.modeling_qwen3_tts import Qwen3TTSTalkerForConditionalGeneration
    t, c = cfg.talker, cfg.cp
    cp_cfg = dict(vocab_size=c.vocab_size, hidden_size=c.hidden_size, intermediate_size=c.intermediate_size,
                  num_hidden_layers=c.num_layers, num_attention_heads=c.num_heads,
                  num_key_value_heads=c.num_kv_heads, head_dim=c.head_dim, rms_norm_eps=c.rms_eps,
                  rope_theta=c.rope_theta, num_code_groups=cfg.num_code_groups, pad_token_id=None)
    tcfg = Qwen3TTSTalkerConfig(
        code_predictor_config=cp_cfg, vocab_size=t.vocab_size, hidden_size=t.hidden_size,
        intermediate_size=t.intermediate_size, num_hidden_layers=t.num_layers,
        num_attention_heads=t.num_heads, num_key_value_heads=t.num_kv_heads, head_dim=t.head_dim,
        rms_norm_eps=t.rms_eps, rope_theta=t.rope_theta,
        rope_scaling={"rope_type": "default", "mrope_section": [t.head_dim // 2 - 2 * (t.head_dim // 6),
                                                                  t.head_dim // 6, t.head_dim // 6],
                      "interleaved": True},
        num_code_groups=cfg.num_code_groups, text_hidden_size=cfg.text_hidden_size,
        text_vocab_size=text_vocab or cfg.text_vocab_size,
        codec_eos_token_id=cfg.codec_eos_token_id, codec_pad_id=cfg.codec_pad_id, codec_bos_id=cfg.codec_bos_id,
        codec_think_id=cfg.codec_think_id, codec_nothink_id=cfg.codec_nothink_id,
        codec_think_bos_id=cfg.codec_think_bos_id, codec_think_eos_id=cfg.codec_think_eos_id,
        pad_token_id=None)
    tcfg._attn_implementation = "eager"
    tcfg.code_predictor_config._attn_implementation = "eager"
    m = Qwen3TTSTalkerForConditionalGeneration(tcfg).eval().to(dtype)
    return m


def load_weights_into_reference(module, W):
    """W uses `talker.`-prefixed names (reference top-level state_dict names)."""
    sd = {k[len("talker."):]: v for k, v in W.items() if k.startswith("talker.")}
    missing, unexpected = module.load_state_dict(sd, strict=False)
    missing = [k for k in missing if "rotary_emb" not in k and "inv_freq" not in k
               and not k.startswith(("model.text_embedding", "text_projection"))]  # text path: host code, optional here
    assert not unexpected, unexpected
    assert not missing, missing


def build_reference_codec_decoder(ccfg, dtype=torch.float32):
    ref_shims.install()
    from qwen_tts.core.tokenizer_12hz.configuration_qwen3_tts_tokenizer_v2 import Qwen3TTSTokenizerV2DecoderConfig
    from qwen_tts.core.tokenizer_12hz.modeling_qwen3_tts_tokenizer_v2 import Qwen3TTSTokenizerV2Decoder
    cfg = Qwen3TTSTokenizerV2DecoderConfig(**ccfg.to_reference_kwargs())
    cfg._attn_implementation = "eager"
    m = Qwen3TTSTokenizerV2Decoder(cfg).eval().to(dtype)
    return m
