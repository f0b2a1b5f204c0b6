"""Pin the oracle against the reference's OWN modules (run through oracle/ref_shims.py).

Runs only where /root/reference exists (the build container).  The reference ships no tests or golden
vectors, so this is the strongest pin available: same seeded weights in both, reference
modules driven by hand with a DynamicCache (HF generate() cannot run under transformers 5.5.0).
"""
import numpy as np
import pytest
import torch

from oracle import talker as T

pytestmark = pytest.mark.reference


def _setup(seed=1):
    from oracle import ref_driver as R
    cfg = T.cfg_tiny()
    cfg.talker.rope_theta = 1e6
    cfg.cp.rope_theta = 1e4
    W = T.random_weights(cfg, seed=seed)
    m = R.build_reference_talker(cfg)
    R.load_weights_into_reference(m, W)
    return cfg, W, m


def test_talker_and_code_predictor_teacher_forced():
    from transformers.cache_utils import DynamicCache
    cfg, W, m = _setup()
    torch.manual_seed(0)
    B, lens, H = 3, [5, 9, 7], cfg.talker.hidden_size
    embs = [torch.randn(l, H) * 0.5 for l in lens]
    trail = [torch.randn(n, H) * 0.1 for n in (2, 1, 4)]
    pad = torch.randn(H) * 0.1
    sp = T.SamplingCfg(do_sample=False, subtalker_dosample=False, max_new_tokens=6, suppress_eos=True)
    r = T.generate(W, cfg, embs, trail, pad, sp, record_logits=True)
    n_frames = r.codes[0].shape[0]
    assert n_frames == 5
    codes = torch.stack(r.codes)  # (B,N,16)

    # ---- reference: prefill
    Lmax = max(lens)
    x = torch.zeros(B, Lmax, H)
    mask = torch.zeros(B, Lmax, dtype=torch.long)
    for i, e in enumerate(embs):
        x[i, Lmax - len(e):] = e
        mask[i, Lmax - len(e):] = 1
    cache = DynamicCache()
    m.rope_deltas = None
    with torch.no_grad():
        out = m(inputs_embeds=x, attention_mask=mask, past_key_values=cache, use_cache=True,
                cache_position=torch.arange(Lmax))
    tl = r.record["talker_logits"]
    assert np.abs(tl[0] - out.logits[:, -1].numpy()).max() < 2e-5
    past_hidden = out.past_hidden
    Tt = max(t.shape[0] for t in trail)
    cp_i = 0
    for step in range(n_frames):
        c0 = codes[:, step, 0]
        # ---- reference code predictor, driven by hand (:1250-1312)
        cpc = DynamicCache()
        e0 = m.get_input_embeddings()(c0[:, None])
        with torch.no_grad():
            o = m.code_predictor(inputs_embeds=torch.cat((past_hidden, e0), dim=1), past_key_values=cpc,
                                 use_cache=True)
        ref_logits = [o.logits[:, -1]]
        gs = o.generation_steps
        for j in range(1, cfg.num_code_groups - 1):
            with torch.no_grad():
                o = m.code_predictor(input_ids=codes[:, step, j:j + 1], past_key_values=cpc, use_cache=True,
                                     generation_steps=gs)
            gs = o.generation_steps
            ref_logits.append(o.logits[:, -1])
        for j, rl in enumerate(ref_logits):
            ol = r.record["cp_logits"][cp_i + j]
            assert np.abs(ol - rl.numpy()).max() < 2e-5, (step, j)
            assert (np.argmax(ol, -1) == codes[:, step, j + 1].numpy()).all()
        cp_i += cfg.num_code_groups - 1
        # ---- reference talker decode step via the inner model (:1682-1727)
        hid = [e0] + [m.code_predictor.get_input_embeddings()[i](codes[:, step, i + 1:i + 2])
                      for i in range(cfg.num_code_groups - 1)]
        xe = torch.cat(hid, dim=1).sum(1, keepdim=True)
        padv = pad.view(1, 1, H)
        tr = torch.stack([t[step] if step < t.shape[0] else pad for t in trail])[:, None]
        xe = xe + (tr if step < Tt else padv)
        mask = torch.cat((mask, torch.ones(B, 1, dtype=torch.long)), dim=1)
        cp = torch.tensor([Lmax + step])
        pos = (cp[0] + m.rope_deltas).view(1, B, 1).expand(3, -1, -1)
        with torch.no_grad():
            mo = m.model(inputs_embeds=xe, attention_mask=mask, position_ids=pos, past_key_values=cache,
                         use_cache=True, cache_position=cp)
            logits = m.codec_head(mo.last_hidden_state)
        past_hidden = mo.last_hidden_state[:, -1:]
        assert np.abs(tl[step + 1] - logits[:, -1].numpy()).max() < 3e-5, step


def test_leaf_ops_match_reference():
    from oracle import ref_shims
    ref_shims.install()
    from qwen_tts.core.models import modeling_qwen3_tts as M
    torch.manual_seed(0)
    x = torch.randn(2, 5, 64)
    n = M.Qwen3TTSRMSNorm(64, eps=1e-6)
    n.weight.data = torch.randn(64)
    assert torch.equal(n(x), T.rms_norm(x, n.weight.data, 1e-6))
    xb = x.bfloat16()
    nb = n.to(torch.bfloat16)
    assert torch.equal(nb(xb), T.rms_norm(xb, nb.weight.data, 1e-6))
    assert torch.equal(M.rotate_half(x), T.rotate_half(x))


def _codec_pair(cfg, seed=3):
    from oracle import codec as C, ref_driver as R
    W = C.random_weights(cfg, seed=seed)
    m = R.build_reference_codec_decoder(cfg)
    sd = m.state_dict()
    missing = [k for k in sd if k not in W and "rotary_emb" not in k]
    extra = [k for k in W if k not in sd]
    assert not missing and not extra, (missing[:5], extra[:5])
    for k in sd:
        if k in W:
            assert sd[k].shape == W[k].shape, (k, sd[k].shape, W[k].shape)
    m.load_state_dict(W, strict=False)
    return W, m


def test_codec_decoder_tiny_matches_reference():
    from oracle import codec as C
    cfg = C.cfg_tiny_codec()
    W, m = _codec_pair(cfg)
    g = torch.Generator().manual_seed(5)
    codes = torch.randint(0, cfg.codebook_size, (2, 16, 13), generator=g)
    with torch.no_grad():
        ref = m(codes)
    out = C.decoder_forward(W, cfg, codes)
    assert out.shape == ref.shape == (2, 1, 13 * 1920)
    assert (out - ref).abs().max() < 2e-5
    assert out.abs().max() > 0.05  # not a degenerate all-zero / all-clamped signal
    # chunked decode with several chunks + wrapper semantics (pad -1, trim)
    codes_long = torch.randint(0, cfg.codebook_size, (2, 16, 40), generator=g)
    with torch.no_grad():
        ref_c = m.chunked_decode(codes_long, chunk_size=16, left_context_size=5)
    out_c = C.chunked_decode(W, cfg, codes_long, chunk_size=16, left_context_size=5)
    assert (out_c - ref_c).abs().max() < 2e-5


def test_codec_decoder_default_shapes_names():
    """Full default config: parameter names/shapes line up with the reference (195.08 M params)."""
    from oracle import codec as C
    cfg = C.CodecCfg()
    W, m = _codec_pair(cfg)
    n = sum(v.numel() for k, v in W.items() if "input_proj.weight" not in k or "pre_transformer" in k)
    assert abs(sum(p.numel() for p in m.parameters()) - sum(v.numel() for v in W.values())) == 0
    codes = torch.randint(0, cfg.codebook_size, (1, 16, 3))
    with torch.no_grad():
        ref = m(codes)
    out = C.decoder_forward(W, cfg, codes)
    assert (out - ref).abs().max() < 5e-5


@pytest.mark.parametrize("which", ["tiny", "default"])
def test_speaker_encoder_and_mel_match_reference(which):
    """ECAPA-TDNN x-vector (modeling_qwen3_tts.py:300-393) and the log-mel front end (:396-448).  The mel FILTERBANK is
    librosa's (absent): the reference function is run with oracle.speaker_encoder's restatement patched in, so STFT,
    magnitude, projection and log are pinned; the filterbank itself stays 'parity unpinned'."""
    from oracle import ref_shims, speaker_encoder as S
    ref_shims.install()
    from qwen_tts.core.models import modeling_qwen3_tts as RM
    from qwen_tts.core.models.configuration_qwen3_tts import Qwen3TTSSpeakerEncoderConfig
    cfg = S.cfg_tiny_spk() if which == "tiny" else S.SpkEncCfg()
    rc = Qwen3TTSSpeakerEncoderConfig(mel_dim=cfg.mel_dim, enc_dim=cfg.enc_dim, enc_channels=list(cfg.enc_channels),
                                      enc_kernel_sizes=list(cfg.enc_kernel_sizes), enc_dilations=list(cfg.enc_dilations),
                                      enc_attention_channels=cfg.enc_attention_channels,
                                      enc_res2net_scale=cfg.enc_res2net_scale, enc_se_channels=cfg.enc_se_channels)
    m = RM.Qwen3TTSSpeakerEncoder(rc).eval()
    W = S.random_weights(cfg, seed=1)
    m.load_state_dict(W)  # strict: every reference parameter has a counterpart
    torch.manual_seed(0)
    mels = torch.randn(2, 57, cfg.mel_dim)
    with torch.no_grad():
        ref = m(mels)
    assert (ref - S.speaker_encoder(W, cfg, mels)).abs().max() < 1e-5
    RM.librosa_mel_fn = lambda sr, n_fft, n_mels, fmin, fmax: S.slaney_mel_filterbank(sr, n_fft, n_mels, fmin, fmax)
    y = (torch.randn(2, 9000) * 0.1).clamp(-1, 1)
    refm = RM.mel_spectrogram(y, n_fft=1024, num_mels=cfg.mel_dim, sampling_rate=24000, hop_size=256, win_size=1024,
                              fmin=0, fmax=12000)
    assert (refm - S.mel_spectrogram(y, num_mels=cfg.mel_dim)).abs().max() < 1e-5
    fb = S.slaney_mel_filterbank(24000, 1024, 128, 0, 12000)
    assert fb.shape == (128, 513) and (fb >= 0).all() and (fb.sum(1) > 0).all()
