"""Checkpoint reading — the `from_pretrained` half of the drop-in boundary.

The reference loads through HF `AutoModel.from_pretrained` (inference/qwen3_tts_model.py:82-121,
core/models/modeling_qwen3_tts.py:1843-1941, inference/qwen3_tts_tokenizer.py:63-99).  The on-disk format is plain
HF: `config.json` + `model.safetensors` (or a sharded `model.safetensors.index.json`), a `speech_tokenizer/`
sub-directory with the codec's own `config.json` + safetensors, and `generation_config.json`.  This module reads
that format WITHOUT importing the reference: config dictionaries are completed with the reference's constructor
defaults (cited per table) and turned into the engine's config records; tensors stay keyed by the reference's
state_dict names, which is what the engines' weight converters consume.
"""
import glob
import json
import os
from types import SimpleNamespace
from typing import Dict, Optional, Tuple

import torch

from .config import CodecConfig, EncoderConfig, SpeakerEncoderConfig, TTSConfig

# core/models/configuration_qwen3_tts.py:187-212 (Qwen3TTSTalkerCodePredictorConfig.__init__ defaults)
CODE_PREDICTOR_DEFAULTS = dict(vocab_size=2048, hidden_size=1024, intermediate_size=3072, num_hidden_layers=5,
                               num_attention_heads=16, num_key_value_heads=8, head_dim=128, rms_norm_eps=1e-6,
                               rope_theta=10000, rope_scaling=None, num_code_groups=32)
# core/models/configuration_qwen3_tts.py:370-404 (Qwen3TTSTalkerConfig.__init__ defaults; no head_dim default — the
# modeling code falls back to hidden_size // num_attention_heads, modeling_qwen3_tts.py:737)
TALKER_DEFAULTS = dict(vocab_size=3072, hidden_size=1024, intermediate_size=2048, num_hidden_layers=20,
                       num_attention_heads=16, num_key_value_heads=2, rms_norm_eps=1e-6, rope_theta=10000,
                       rope_scaling=None, num_code_groups=32, text_hidden_size=2048, codec_eos_token_id=4198,
                       codec_think_id=4202, codec_nothink_id=4203, codec_think_bos_id=4204, codec_think_eos_id=4205,
                       codec_pad_id=4196, codec_bos_id=4197, spk_id=None, spk_is_dialect=None, codec_language_id=None)
# core/models/configuration_qwen3_tts.py:465-477 (Qwen3TTSConfig.__init__ defaults)
TOP_DEFAULTS = dict(tokenizer_type=None, tts_model_size=None, tts_model_type=None, im_start_token_id=151644,
                    im_end_token_id=151645, tts_pad_token_id=151671, tts_bos_token_id=151672, tts_eos_token_id=151673)
# core/tokenizer_12hz/configuration_qwen3_tts_tokenizer_v2.py:72-93 (Qwen3TTSTokenizerV2DecoderConfig defaults;
# codebook_dim has none — modeling…v2.py:831-836 reads it from the checkpoint's config)
DECODER_DEFAULTS = dict(codebook_size=2048, hidden_size=1024, latent_dim=1024, max_position_embeddings=8000,
                        rope_theta=10000, num_attention_heads=16, num_key_value_heads=16, sliding_window=72,
                        intermediate_size=3072, rms_norm_eps=1e-5, num_hidden_layers=8, num_quantizers=16,
                        upsample_rates=(8, 5, 4, 3), upsampling_ratios=(2, 2), decoder_dim=1536)
# core/tokenizer_12hz/configuration_qwen3_tts_tokenizer_v2.py:143-152 (Qwen3TTSTokenizerV2Config defaults)
TOKENIZER_DEFAULTS = dict(encoder_valid_num_quantizers=16, input_sample_rate=24000, output_sample_rate=24000,
                          decode_upsample_rate=1920, encode_downsample_rate=1920)


def _ns(d: Optional[dict], defaults: dict) -> SimpleNamespace:
    out = dict(defaults)
    out.update(d or {})
    return SimpleNamespace(**out)


def read_json(path: str) -> dict:
    with open(path, "r", encoding="utf-8") as f:
        return json.load(f)


def resolve_dir(name_or_path: str) -> str:
    """Local directory, or a hub snapshot if `huggingface_hub` can provide one (the build/bench boxes are offline)."""
    if os.path.isdir(name_or_path):
        return name_or_path
    try:
        from huggingface_hub import snapshot_download
        return snapshot_download(name_or_path)
    except Exception as e:  # pragma: no cover - needs network
        raise FileNotFoundError(f"{name_or_path!r} is not a local directory and could not be fetched: {e}") from e


def read_state_dict(directory: str, device="cpu", dtype: Optional[torch.dtype] = None,
                    prefixes: Optional[Tuple[str, ...]] = None) -> Dict[str, torch.Tensor]:
    """All tensors of a HF safetensors checkpoint (single file, sharded index, or any *.safetensors), optionally only
    those whose name starts with one of `prefixes`; floating tensors are cast to `dtype` when given."""
    from safetensors import safe_open
    index = os.path.join(directory, "model.safetensors.index.json")
    if os.path.exists(index):
        files = sorted({os.path.join(directory, f) for f in read_json(index)["weight_map"].values()})
    else:
        single = os.path.join(directory, "model.safetensors")
        files = [single] if os.path.exists(single) else sorted(glob.glob(os.path.join(directory, "*.safetensors")))
    if not files:
        raise FileNotFoundError(f"no safetensors weights under {directory}")
    out: Dict[str, torch.Tensor] = {}
    for fn in files:
        with safe_open(fn, framework="pt", device=str(device)) as f:
            for k in f.keys():
                if prefixes is not None and not k.startswith(prefixes):
                    continue
                t = f.get_tensor(k)
                if dtype is not None and t.is_floating_point():
                    t = t.to(dtype)
                out[k] = t
    return out


def tts_config_from_dict(cfg: dict):
    """config.json of a Qwen3-TTS checkpoint -> (TTSConfig, meta).  meta carries what the host-side generate() needs:
    speaker / dialect / language tables (lower-cased by the model class), model type and size."""
    talker_d = dict(cfg.get("talker_config") or {})
    cp = _ns(talker_d.pop("code_predictor_config", None), CODE_PREDICTOR_DEFAULTS)
    talker = _ns(talker_d, TALKER_DEFAULTS)
    talker.code_predictor_config = cp
    top = _ns({k: v for k, v in cfg.items() if k in TOP_DEFAULTS}, TOP_DEFAULTS)
    top.talker_config = talker
    tcfg = TTSConfig.from_hf(top)
    meta = dict(spk_id=talker.spk_id or {}, spk_is_dialect=talker.spk_is_dialect or {},
                codec_language_id=talker.codec_language_id or {}, tts_model_type=top.tts_model_type,
                tts_model_size=top.tts_model_size, tokenizer_type=top.tokenizer_type,
                speaker_encoder_config=SpeakerEncoderConfig.from_dict(cfg.get("speaker_encoder_config")))
    return tcfg, meta


def load_tts_checkpoint(directory: str, device="cuda:0", dtype=torch.bfloat16):
    """-> (TTSConfig, weights keyed by the reference state_dict names, meta, generate_config | None).
    Base checkpoints also carry `speaker_encoder.*` (modeling_qwen3_tts.py:1822-1825): returned fp32, prefix stripped,
    in meta["speaker_encoder_weights"] (empty for the other model types)."""
    tcfg, meta = tts_config_from_dict(read_json(os.path.join(directory, "config.json")))
    W = read_state_dict(directory, device=device, dtype=dtype, prefixes=("talker.",))
    spk = {}
    if meta["tts_model_type"] == "base":
        spk = {k[len("speaker_encoder."):]: v for k, v in
               read_state_dict(directory, device=device, dtype=torch.float32, prefixes=("speaker_encoder.",)).items()}
    meta["speaker_encoder_weights"] = spk
    gen = None
    gpath = os.path.join(directory, "generation_config.json")
    if os.path.exists(gpath):
        gen = read_json(gpath)
    return tcfg, W, meta, gen


def tokenizer_configs_from_dict(cfg: dict):
    """speech_tokenizer/config.json -> (CodecConfig, EncoderConfig, rates)."""
    from transformers import MimiConfig
    top = _ns({k: v for k, v in cfg.items() if k in TOKENIZER_DEFAULTS}, TOKENIZER_DEFAULTS)
    dec_d = dict(cfg.get("decoder_config") or {})
    if "codebook_dim" not in dec_d:
        raise ValueError("decoder_config.codebook_dim is required (it has no default in the reference either)")
    dec = _ns(dec_d, DECODER_DEFAULTS)
    ccfg = CodecConfig.from_hf(dec)
    enc_d = {k: v for k, v in (cfg.get("encoder_config") or {}).items()
             if k not in ("model_type", "transformers_version", "architectures", "torch_dtype", "dtype")}
    ecfg = EncoderConfig.from_hf(MimiConfig(**enc_d), valid_num_quantizers=top.encoder_valid_num_quantizers,
                                 encode_downsample_rate=top.encode_downsample_rate)
    if ccfg.total_upsample != top.decode_upsample_rate:
        raise ValueError(f"decode_upsample_rate {top.decode_upsample_rate} != product of the decoder strides {ccfg.total_upsample}")
    rates = dict(input_sample_rate=top.input_sample_rate, output_sample_rate=top.output_sample_rate,
                 decode_upsample_rate=top.decode_upsample_rate, encode_downsample_rate=top.encode_downsample_rate)
    return ccfg, ecfg, rates


def load_speech_tokenizer_checkpoint(directory: str, device="cuda:0"):
    """-> (CodecConfig, decoder weights, EncoderConfig, encoder weights, rates).  The checkpoint is the state_dict of
    Qwen3TTSTokenizerV2Model (…v2.py:932-959): `decoder.<Qwen3TTSTokenizerV2Decoder names>` and
    `encoder.<MimiModel names>`; the leading component is stripped so that each engine sees its module's own names."""
    ccfg, ecfg, rates = tokenizer_configs_from_dict(read_json(os.path.join(directory, "config.json")))
    sd = read_state_dict(directory, device=device, dtype=torch.float32, prefixes=("decoder.", "encoder."))
    dec = {k[len("decoder."):]: v for k, v in sd.items() if k.startswith("decoder.")}
    enc = {k[len("encoder."):]: v for k, v in sd.items() if k.startswith("encoder.")}
    return ccfg, dec, ecfg, enc, rates
