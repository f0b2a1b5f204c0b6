"""Codec ENCODER oracle (oracle/mimi_encoder.py) pinned against the third-party implementation the reference wraps
(transformers MimiModel — installed copy 5.5.0, the reference pins 4.57.3), against the committed golden codes, and
through the size-independent properties the path offers (causality, batch independence, frame count)."""
import os

import numpy as np
import pytest
import torch

from oracle import mimi_encoder as M
from oracle.make_golden import micro_encoder_cfg

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _hf(cfg, W):
    from transformers import MimiConfig, MimiModel
    hf = MimiModel(MimiConfig(**cfg.to_hf_kwargs())).eval()
    missing, unexpected = hf.load_state_dict(W, strict=False)
    assert not unexpected
    assert all(not k.startswith(("encoder.", "encoder_transformer.", "downsample.")) for k in missing), missing
    return hf


@pytest.mark.parametrize("T", [1, 959, 1920, 5000, 23000])
def test_oracle_equals_hf_mimi(T):
    """Bit-exact codes incl. edge lengths: 1 sample, one short of a 25 Hz frame, exactly one code frame, ragged, and
    23000 samples = 24 transformer frames > the tiny sliding window (6), so the window mask is exercised."""
    cfg = M.cfg_tiny_encoder()
    W = M.random_weights(cfg, seed=3)
    hf = _hf(cfg, W)
    g = torch.Generator().manual_seed(T)
    wav = (torch.randn(2, T, generator=g) * 0.1).clamp(-1, 1)
    with torch.no_grad():
        ref = hf.encode(wav[:, None, :], return_dict=True).audio_codes[:, : cfg.valid_num_quantizers]
    out = M.encode(W, cfg, wav)
    assert out.shape == ref.shape == (2, 16, -(-T // cfg.hop))
    assert (out == ref).all()


def test_oracle_equals_hf_mimi_default_config():
    cfg = M.MimiEncCfg()
    W = M.random_weights(cfg, seed=5)
    hf = _hf(cfg, W)
    g = torch.Generator().manual_seed(0)
    wav = (torch.randn(1, 12000, generator=g) * 0.1).clamp(-1, 1)
    with torch.no_grad():
        ref = hf.encode(wav[:, None, :], return_dict=True).audio_codes[:, :16]
    assert (M.encode(W, cfg, wav) == ref).all()


def test_golden_codes():
    z = np.load(os.path.join(GOLD, "encoder_micro.npz"))
    cfg = micro_encoder_cfg()
    W = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("W::")}
    out = M.encode(W, cfg, torch.from_numpy(z["wav"]))
    assert (out.numpy() == z["codes"]).all()


def test_causality_and_ragged_batch():
    """Every layer is causal, so (i) codes of a frame-aligned prefix equal the prefix of the codes and (ii) a row's
    codes do not depend on how far the batch is right-padded — the reason the reference may pad and trim (…v2.py:983)."""
    cfg = M.cfg_tiny_encoder()
    W = M.random_weights(cfg, seed=7)
    g = torch.Generator().manual_seed(1)
    a = (torch.randn(4 * cfg.hop + 300, generator=g) * 0.1)
    b = (torch.randn(2 * cfg.hop, generator=g) * 0.1)
    full = M.encode(W, cfg, a[None])
    pre = M.encode(W, cfg, a[None, : 3 * cfg.hop])
    assert (full[:, :, :3] == pre).all()
    rows = M.tokenizer_encode(W, cfg, [a, b])
    assert rows[0].shape == (5, 16) and rows[1].shape == (2, 16)
    assert (rows[0] == full[0].T).all()
    assert (rows[1] == M.encode(W, cfg, b[None])[0].T).all()
