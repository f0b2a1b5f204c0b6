"""CPU oracle: speaker x-vector path used by voice cloning (not yet on the GPU).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Restates
  * `mel_spectrogram`            qwen_tts/core/models/modeling_qwen3_tts.py:396-448 (called with n_fft 1024, 128 mels,
                                 hop 256, win 1024, fmin 0, fmax 12000 at :1941-1952),
  * `Qwen3TTSSpeakerEncoder`     :300-393 (ECAPA-TDNN) with its blocks: TimeDelayNetBlock :229-250, Res2NetBlock :95-126,
                                 SqueezeExcitationBlock :129-157, SqueezeExcitationRes2NetBlock :253-297,
                                 AttentiveStatisticsPooling :160-226.
The mel FILTERBANK comes from a third-party dependency that is absent here (`librosa.filters.mel`, librosa is
unpinned in pyproject.toml): `slaney_mel_filterbank` restates its published algorithm (Slaney mel scale,
area normalisation) — that one function is "parity unpinned"; everything else is pinned against the reference's
own module in tests/test_oracle_vs_reference.py.

Weights: a flat dict keyed by the reference's `speaker_encoder.` state_dict names (prefix stripped).
"""
import math
from dataclasses import dataclass
from typing import Dict, Tuple

import numpy as np
import torch
import torch.nn.functional as F


@dataclass
class SpkEncCfg:
    """Qwen3TTSSpeakerEncoderConfig defaults (core/models/configuration_qwen3_tts.py:47-57)."""
    mel_dim: int = 128
    enc_dim: int = 1024
    enc_channels: Tuple[int, ...] = (512, 512, 512, 512, 1536)
    enc_kernel_sizes: Tuple[int, ...] = (5, 3, 3, 3, 1)
    enc_dilations: Tuple[int, ...] = (1, 2, 3, 4, 1)
    enc_attention_channels: int = 128
    enc_res2net_scale: int = 8
    enc_se_channels: int = 128
    sample_rate: int = 24000


def cfg_tiny_spk() -> SpkEncCfg:
    return SpkEncCfg(mel_dim=16, enc_dim=24, enc_channels=(32, 32, 32, 64), enc_kernel_sizes=(5, 3, 3, 1),
                     enc_dilations=(1, 2, 3, 1), enc_attention_channels=8, enc_res2net_scale=4, enc_se_channels=8)


# ---------------------------------------------------------------------------------------------- mel front end
def _hz_to_mel(f):
    f = np.asanyarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz, logstep = 1000.0, np.log(6.4) / 27.0
    min_log_mel = min_log_hz / f_sp
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-30) / min_log_hz) / logstep, mels)


def _mel_to_hz(m):
    m = np.asanyarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    min_log_hz, logstep = 1000.0, np.log(6.4) / 27.0
    min_log_mel = min_log_hz / f_sp
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def slaney_mel_filterbank(sr, n_fft, n_mels, fmin, fmax) -> np.ndarray:
    """librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax) with its defaults htk=False, norm="slaney": triangular
    filters on the Slaney mel scale, each scaled by 2 / (upper edge - lower edge).  (n_mels, 1 + n_fft // 2) float32."""
    fmax = sr / 2.0 if fmax is None else fmax
    fftfreqs = np.linspace(0.0, sr / 2.0, 1 + n_fft // 2)
    mel_f = _mel_to_hz(np.linspace(_hz_to_mel(fmin), _hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    w = np.zeros((n_mels, 1 + n_fft // 2))
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        w[i] = np.maximum(0.0, np.minimum(lower, upper))
    w *= (2.0 / (mel_f[2: n_mels + 2] - mel_f[:n_mels]))[:, None]
    return w.astype(np.float32)


def mel_spectrogram(y: torch.Tensor, n_fft=1024, num_mels=128, sampling_rate=24000, hop_size=256, win_size=1024, fmin=0,
                    fmax=12000, mel_basis: torch.Tensor = None) -> torch.Tensor:
    """:396-448 — reflect-pad (n_fft - hop)/2 both sides, STFT (Hann, center=False), magnitude sqrt(re^2+im^2+1e-9),
    mel projection, log(clamp(., 1e-5)).  y: (B, T) -> (B, num_mels, frames)."""
    if mel_basis is None:
        mel_basis = torch.from_numpy(slaney_mel_filterbank(sampling_rate, n_fft, num_mels, fmin, fmax))
    pad = (n_fft - hop_size) // 2
    y = F.pad(y[:, None, :], (pad, pad), mode="reflect")[:, 0]
    spec = torch.stft(y, n_fft, hop_length=hop_size, win_length=win_size, window=torch.hann_window(win_size),
                      center=False, pad_mode="reflect", normalized=False, onesided=True, return_complex=True)
    mag = torch.sqrt(torch.view_as_real(spec).pow(2).sum(-1) + 1e-9)
    return torch.log(torch.clamp(torch.matmul(mel_basis, mag), min=1e-5))


# ---------------------------------------------------------------------------------------------- ECAPA-TDNN
def _conv_same_reflect(x, w, b, dilation=1):
    """nn.Conv1d(padding="same", padding_mode="reflect"): total pad d*(k-1), left = total // 2, right = the rest."""
    total = dilation * (w.shape[-1] - 1)
    left = total // 2
    if total > 0:
        x = F.pad(x, (left, total - left), mode="reflect")
    return F.conv1d(x, w, b, dilation=dilation)


def _tdnn(W, p, x, dilation=1):
    """TimeDelayNetBlock (:229-250): ReLU(conv)."""
    return F.relu(_conv_same_reflect(x, W[p + ".conv.weight"], W[p + ".conv.bias"], dilation))


def _res2net(W, p, x, scale, dilation):
    """Res2NetBlock.forward (:114-126)."""
    outs, prev = [], None
    for i, part in enumerate(torch.chunk(x, scale, dim=1)):
        if i == 0:
            out = part
        elif i == 1:
            out = _tdnn(W, f"{p}.blocks.{i - 1}", part, dilation)
        else:
            out = _tdnn(W, f"{p}.blocks.{i - 1}", part + prev, dilation)
        prev = out
        outs.append(out)
    return torch.cat(outs, dim=1)


def _se(W, p, x):
    """SqueezeExcitationBlock.forward (:150-157): channel gate from the time mean."""
    m = x.mean(dim=2, keepdim=True)
    m = F.relu(F.conv1d(m, W[p + ".conv1.weight"], W[p + ".conv1.bias"]))
    m = torch.sigmoid(F.conv1d(m, W[p + ".conv2.weight"], W[p + ".conv2.bias"]))
    return x * m


def _asp(W, p, x, eps=1e-12):
    """AttentiveStatisticsPooling.forward (:203-226) with the all-ones length mask the reference builds."""
    L = x.shape[-1]

    def stats(x, m):
        mean = (m * x).sum(2)
        std = torch.sqrt((m * (x - mean.unsqueeze(2)).pow(2)).sum(2).clamp(eps))
        return mean, std

    mean, std = stats(x, torch.full((x.shape[0], 1, L), 1.0 / L))
    att = torch.cat([x, mean.unsqueeze(2).repeat(1, 1, L), std.unsqueeze(2).repeat(1, 1, L)], dim=1)
    att = F.conv1d(torch.tanh(_tdnn(W, p + ".tdnn", att)), W[p + ".conv.weight"], W[p + ".conv.bias"])
    att = F.softmax(att, dim=2)
    mean, std = stats(x, att)
    return torch.cat((mean, std), dim=1).unsqueeze(2)


@torch.no_grad()
def speaker_encoder(W: Dict[str, torch.Tensor], cfg: SpkEncCfg, mels: torch.Tensor) -> torch.Tensor:
    """Qwen3TTSSpeakerEncoder.forward (:371-393).  mels: (B, frames, mel_dim) -> (B, enc_dim)."""
    x = mels.transpose(1, 2)
    feats = []
    x = _tdnn(W, "blocks.0", x, cfg.enc_dilations[0])
    feats.append(x)
    for i in range(1, len(cfg.enc_channels) - 1):
        p = f"blocks.{i}"
        r = x
        h = _tdnn(W, p + ".tdnn1", x)
        h = _res2net(W, p + ".res2net_block", h, cfg.enc_res2net_scale, cfg.enc_dilations[i])
        h = _tdnn(W, p + ".tdnn2", h)
        x = _se(W, p + ".se_block", h) + r
        feats.append(x)
    x = _tdnn(W, "mfa", torch.cat(feats[1:], dim=1), cfg.enc_dilations[-1])
    x = _asp(W, "asp", x)
    return F.conv1d(x, W["fc.weight"], W["fc.bias"]).squeeze(-1)


@torch.no_grad()
def extract_speaker_embedding(W, cfg: SpkEncCfg, audio: torch.Tensor) -> torch.Tensor:
    """Qwen3TTSForConditionalGeneration.extract_speaker_embedding (:1941-1954): 24 kHz mono (T,) -> (enc_dim,)."""
    mels = mel_spectrogram(audio[None].float(), num_mels=cfg.mel_dim).transpose(1, 2)
    return speaker_encoder(W, cfg, mels)[0]


def random_weights(cfg: SpkEncCfg, seed=0) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    W = {}

    def conv(p, cin, cout, k):
        W[p + ".weight"] = torch.randn(cout, cin, k, generator=g) / math.sqrt(cin * k)
        W[p + ".bias"] = torch.randn(cout, generator=g) * 0.05

    ch, ks = cfg.enc_channels, cfg.enc_kernel_sizes
    conv("blocks.0.conv", cfg.mel_dim, ch[0], ks[0])
    s = cfg.enc_res2net_scale
    for i in range(1, len(ch) - 1):
        p = f"blocks.{i}"
        conv(p + ".tdnn1.conv", ch[i - 1], ch[i], 1)
        for j in range(s - 1):
            conv(f"{p}.res2net_block.blocks.{j}.conv", ch[i] // s, ch[i] // s, ks[i])
        conv(p + ".tdnn2.conv", ch[i], ch[i], 1)
        conv(p + ".se_block.conv1", ch[i], cfg.enc_se_channels, 1)
        conv(p + ".se_block.conv2", cfg.enc_se_channels, ch[i], 1)
    conv("mfa.conv", ch[-1], ch[-1], ks[-1])
    conv("asp.tdnn.conv", ch[-1] * 3, cfg.enc_attention_channels, 1)
    conv("asp.conv", cfg.enc_attention_channels, ch[-1], 1)
    conv("fc", ch[-1] * 2, cfg.enc_dim, 1)
    return W
