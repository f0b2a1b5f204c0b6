"""GPU parity of the fp32 codec ENCODER (csrc/codec_encoder.cu, through the C ABI) against the CPU oracle.

The codes are discrete, so the bar is: every stage activation within a small fp32 tolerance of the oracle's, and the
codes IDENTICAL except where the oracle's own best/second-best centroid gap is below that tolerance (a flip at level
q of a frame legitimately changes every deeper level of that frame, so only the FIRST differing level is judged)."""
import os

import numpy as np
import pytest
import torch

from oracle import mimi_encoder as M
from oracle.make_golden import micro_encoder_cfg

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
STAGE_RTOL = 2e-4      # max |gpu - oracle| / max |oracle| per stage (fp32, different summation order)
MARGIN_TOL = 5e-3      # squared-distance gap below which a nearest-centroid decision may legitimately flip


def _pkg_cfg(c):
    import qwen3_tts_b200 as q
    from qwen3_tts_b200.config import EncoderConfig
    kw = {k: getattr(c, k) for k in EncoderConfig.__dataclass_fields__ if hasattr(c, k)}
    return EncoderConfig(**kw)


def _check(cfg, W, wav, check_stages=True):
    from qwen3_tts_b200.codec_encoder import CodecEncoder
    enc = CodecEncoder(_pkg_cfg(cfg), W, device=DEV, max_frames=512)
    stages, margins = [], []
    ref = M.encode(W, cfg, wav, stages=stages, margins=margins)
    codes, got = enc.forward_with_stages(wav.to(DEV))
    codes = codes.cpu()
    assert codes.shape == ref.shape
    if check_stages:
        assert [n for n, _ in stages] == enc.stage_names
        for name, t in stages:
            g = got[name].cpu()
            assert g.shape == t.shape, (name, g.shape, t.shape)
            err = float((g - t).abs().max() / t.abs().max().clamp(min=1e-20))
            assert err < STAGE_RTOL, f"stage {name}: relative error {err:.2e}"
    # codes: first differing level per (row, frame) must be a near-tie in the oracle
    diff = codes != ref
    n_flip = 0
    for b, t in zip(*np.nonzero(diff.any(1).numpy())):
        q = int(np.argmax(diff[b, :, t].numpy()))
        gap = float(margins[q][b, t])
        assert gap < MARGIN_TOL, f"row {b} frame {t} level {q}: codes differ with oracle gap {gap:.3e}"
        n_flip += 1
    assert n_flip <= max(1, int(0.05 * ref.shape[0] * ref.shape[2])), f"{n_flip} frames differ"
    assert enc.last_launches() > 0
    return codes, ref


@pytest.mark.parametrize("B,T", [(1, 1), (2, 959), (1, 1920), (3, 5000), (2, 23000)])
def test_tiny_encoder_matches_oracle(B, T):
    cfg = M.cfg_tiny_encoder()
    W = M.random_weights(cfg, seed=3)
    g = torch.Generator().manual_seed(T)
    wav = (torch.randn(B, T, generator=g) * 0.1).clamp(-1, 1)
    _check(cfg, W, wav)


def test_default_encoder_3s_matches_oracle():
    """BASELINE config 1 (tokenizer round trip, 3 s @ 24 kHz): 72 000 samples -> (16, 38) codes."""
    cfg = M.MimiEncCfg()
    W = M.random_weights(cfg, seed=5)
    g = torch.Generator().manual_seed(0)
    wav = (torch.randn(1, 72000, generator=g) * 0.1).clamp(-1, 1)
    codes, ref = _check(cfg, W, wav)
    assert codes.shape == (1, 16, 38)


def test_golden_codes_and_ragged_batch():
    from qwen3_tts_b200.codec_encoder import CodecEncoder
    z = np.load(os.path.join(GOLD, "encoder_micro.npz"))
    cfg = micro_encoder_cfg()
    W = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("W::")}
    wav = torch.from_numpy(z["wav"])
    codes, ref = _check(cfg, W, wav, check_stages=False)
    assert (ref.numpy() == z["codes"]).all()
    # list API = Qwen3TTSTokenizerV2Model.encode: pad, encode, trim to ceil(len / 1920) frames
    enc = CodecEncoder(_pkg_cfg(cfg), W, device=DEV, max_frames=512)
    a, b = wav[0, :9000], wav[1, :4000]
    rows = enc.encode([a, b])
    want = M.tokenizer_encode(W, cfg, [a, b])
    assert [tuple(r.shape) for r in rows] == [tuple(w.shape) for w in want] == [(5, 16), (3, 16)]
    for r, w in zip(rows, want):
        assert (r.cpu() == w).float().mean() > 0.9


def test_tokenizer_round_trip_3s():
    """BASELINE config 1 through the product wrapper: 3 s @ 24 kHz -> encode -> (38, 16) codes -> decode -> 72 960
    samples (seeded random weights; shapes, ranges and determinism, not audio quality)."""
    import qwen3_tts_b200 as q
    from qwen3_tts_b200 import synthetic
    from qwen3_tts_b200.config import EncoderConfig
    from qwen3_tts_b200.model import Qwen3TTSTokenizer
    ecfg = EncoderConfig()
    ccfg = q.CodecConfig()
    tok = Qwen3TTSTokenizer(ccfg, synthetic.random_codec_weights(ccfg, device=DEV, seed=1), device=DEV, max_frames=64,
                            encoder_cfg=ecfg, encoder_weights=synthetic.random_encoder_weights(ecfg, seed=2))
    g = np.random.default_rng(0)
    wav = np.clip(g.standard_normal(72000).astype(np.float32) * 0.1, -1, 1)
    enc = tok.encode(wav, sr=24000)
    assert len(enc.audio_codes) == 1 and tuple(enc.audio_codes[0].shape) == (38, 16)
    c = enc.audio_codes[0]
    assert c.dtype == torch.long and int(c.min()) >= 0 and int(c.max()) < ecfg.codebook_size
    again = tok.encode([wav, wav[:30000]], sr=24000).audio_codes
    assert (again[0] == c).all() and tuple(again[1].shape) == (16, 16) and (again[1] == c[:16]).float().mean() > 0.9
    wavs, sr = tok.decode(enc)
    assert sr == 24000 and wavs[0].shape == (72960,) and np.isfinite(wavs[0]).all() and np.abs(wavs[0]).max() <= 1.0
