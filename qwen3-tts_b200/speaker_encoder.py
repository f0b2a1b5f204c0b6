"""Host side of the speaker x-vector path (Qwen3TTSForConditionalGeneration.extract_speaker_embedding,
core/models/modeling_qwen3_tts.py:1941-1954) on libqwen3tts_b200.so.

`weights`: the reference's `speaker_encoder.*` state_dict entries with the prefix stripped, any float dtype.  The mel
front-end tables are computed here on the CPU: Hann window (torch.hann_window, periodic), the exact DFT twiddles, and the
Slaney-scale, area-normalised triangular filterbank that `librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax)` produces
with its defaults (htk=False, norm="slaney") — librosa itself is not a dependency of this package.
"""
import ctypes as C
import math
from typing import Dict

import numpy as np
import torch

from . import _lib
from .config import SpeakerEncoderConfig


def _hz_to_mel(f):
    f = np.asanyarray(f, dtype=np.float64)
    f_sp, min_log_hz, logstep = 200.0 / 3, 1000.0, np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_hz / f_sp + np.log(np.maximum(f, 1e-30) / min_log_hz) / logstep, f / f_sp)


def _mel_to_hz(m):
    m = np.asanyarray(m, dtype=np.float64)
    f_sp, min_log_hz, logstep = 200.0 / 3, 1000.0, np.log(6.4) / 27.0
    min_log_mel = min_log_hz / f_sp
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_filterbank(sr, n_fft, n_mels, fmin, fmax) -> np.ndarray:
    """(n_mels, 1 + n_fft // 2) float32 — Slaney mel scale, triangles scaled by 2 / (upper edge - lower edge)."""
    fmax = sr / 2.0 if fmax is None else fmax
    fftfreqs = np.linspace(0.0, sr / 2.0, 1 + n_fft // 2)
    mel_f = _mel_to_hz(np.linspace(_hz_to_mel(fmin), _hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    w = np.zeros((n_mels, 1 + n_fft // 2))
    for i in range(n_mels):
        w[i] = np.maximum(0.0, np.minimum(-ramps[i] / fdiff[i], ramps[i + 2] / fdiff[i + 1]))
    w *= (2.0 / (mel_f[2: n_mels + 2] - mel_f[:n_mels]))[:, None]
    return w.astype(np.float32)


class SpeakerEncoder:
    def __init__(self, cfg: SpeakerEncoderConfig, weights: Dict[str, torch.Tensor], device="cuda:0"):
        self.lib = _lib.load()
        self.cfg = cfg
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("SpeakerEncoder needs a CUDA device (no CPU fallback)")
        sc = _lib.SpkCfg()
        sc.mel_dim, sc.enc_dim, sc.n_blocks = int(cfg.mel_dim), int(cfg.enc_dim), len(cfg.enc_channels)
        for i, (c, k, d) in enumerate(zip(cfg.enc_channels, cfg.enc_kernel_sizes, cfg.enc_dilations)):
            sc.channels[i], sc.kernel_sizes[i], sc.dilations[i] = int(c), int(k), int(d)
        sc.attention_channels, sc.res2net_scale, sc.se_channels = (int(cfg.enc_attention_channels),
                                                                   int(cfg.enc_res2net_scale), int(cfg.enc_se_channels))
        sc.n_fft, sc.hop, sc.win, sc.device = int(cfg.n_fft), int(cfg.hop_size), int(cfg.win_size), self.device.index or 0
        h = C.c_void_p()
        _lib.check(self.lib.q3_spk_create(C.byref(sc), C.byref(h)))
        self.h = h
        for name, t in weights.items():
            self._put(name, t)
        n = cfg.n_fft
        ang = 2.0 * math.pi * torch.arange(n, dtype=torch.float64) / n
        self._put("mel.window", torch.hann_window(cfg.win_size))
        self._put("mel.cos", torch.cos(ang).float())
        self._put("mel.sin", torch.sin(ang).float())
        fb = mel_filterbank(cfg.sample_rate, cfg.n_fft, cfg.mel_dim, cfg.fmin, cfg.fmax)
        self._put("mel.fbT", torch.from_numpy(np.ascontiguousarray(fb.T)))
        _lib.check(self.lib.q3_spk_finalize(self.h))

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.q3_spk_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def _put(self, name, x):
        x = x.detach().to(self.device, torch.float32).contiguous()
        shape = (C.c_int64 * x.dim())(*x.shape)
        _lib.check(self.lib.q3_spk_load_tensor(self.h, name.encode(), x.data_ptr(), shape, x.dim()))

    def frames(self, n_samples: int) -> int:
        return self.lib.q3_spk_frames(self.h, int(n_samples))

    def mel(self, wav: torch.Tensor) -> torch.Tensor:
        """(B, T) waveform -> (B, mel_dim, frames) log-mel == mel_spectrogram(...) of the reference (:396-448)."""
        wav = wav.to(self.device, torch.float32).contiguous()
        B, T = wav.shape
        out = torch.empty(B, self.cfg.mel_dim, self.frames(T), dtype=torch.float32, device=self.device)
        st = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self.lib.q3_spk_mel(self.h, wav.data_ptr(), B, T, out.data_ptr(), C.c_void_p(st)))
        return out

    def forward(self, mels: torch.Tensor) -> torch.Tensor:
        """Qwen3TTSSpeakerEncoder.forward: mels (B, frames, mel_dim) -> (B, enc_dim)."""
        m = mels.to(self.device, torch.float32).transpose(1, 2).contiguous()
        B, _, L = m.shape
        emb = torch.empty(B, self.cfg.enc_dim, dtype=torch.float32, device=self.device)
        st = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self.lib.q3_spk_embed(self.h, None, B, 0, m.data_ptr(), L, emb.data_ptr(), C.c_void_p(st)))
        return emb

    def embed_waveform(self, wav: torch.Tensor) -> torch.Tensor:
        """(B, T) 24 kHz waveform -> (B, enc_dim): mel front end + encoder in one call."""
        wav = wav.to(self.device, torch.float32).contiguous()
        B, T = wav.shape
        emb = torch.empty(B, self.cfg.enc_dim, dtype=torch.float32, device=self.device)
        st = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self.lib.q3_spk_embed(self.h, wav.data_ptr(), B, T, None, 0, emb.data_ptr(), C.c_void_p(st)))
        return emb

    def last_launches(self) -> int:
        return self.lib.q3_spk_last_launch_count(self.h)
