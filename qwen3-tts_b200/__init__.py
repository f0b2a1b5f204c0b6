"""qwen3-tts_b200 — B200-native (sm_100a) hot paths of Qwen3-TTS behind the reference's Python API.

Importable as `qwen3_tts_b200` (see the shim module at the repo root; the directory name carries a hyphen).
"""
from .config import CodecConfig, SamplingParams, StackConfig, TTSConfig  # noqa: F401

__all__ = ["TTSConfig", "StackConfig", "SamplingParams", "CodecConfig"]
