"""The call shapes of the reference's own example scripts (its only executable checks) —
examples/test_tokenizer_12hz.py and examples/test_model_12hz_base.py / _custom_voice.py / _voice_design.py — run against
this package's public API with the four engines replaced by CPU recorders: every documented way of calling
encode / decode / generate_* / create_voice_clone_prompt must be accepted and return the documented shapes."""
import numpy as np
import pytest
import torch

from tests.helpers import write_tiny_checkpoint


@pytest.fixture()
def stubbed(monkeypatch):
    from qwen3_tts_b200 import model as M, codec_encoder as CE, speaker_encoder as SE

    class _AR:
        def __init__(self, cfg, w, device=None, max_batch=None, max_ctx=None):
            self.max_batch, self.G = max_batch, cfg.num_code_groups

        def generate(self, embeds, trail, pad, sp, return_hidden=False):
            assert all(e.dim() == 2 for e in embeds) and len(embeds) == len(trail)
            codes = [torch.full((3 + i, self.G), i, dtype=torch.long) for i in range(len(embeds))]
            return (codes, [torch.zeros(c.shape[0], 8) for c in codes]) if return_hidden else codes

    class _Dec:
        total_upsample = 1920

        def __init__(self, cfg, w, device=None, max_frames=None):
            pass

        def decode(self, ac):
            return [torch.zeros(int((row[:, 0] > -1).sum()) * 1920) for row in ac]

    class _Enc:
        def __init__(self, cfg, w, device=None):
            pass

        def encode(self, wavs):
            return [torch.ones(-(-int(x.shape[0]) // 1920), 16, dtype=torch.long) for x in wavs]

    class _Spk:
        def __init__(self, cfg, w, device=None):
            self.cfg = cfg

        def embed_waveform(self, wav):
            return torch.zeros(wav.shape[0], self.cfg.enc_dim)

    monkeypatch.setattr(M, "AREngine", _AR)
    monkeypatch.setattr(M, "CodecDecoder", _Dec)
    monkeypatch.setattr(CE, "CodecEncoder", _Enc)
    monkeypatch.setattr(SE, "SpeakerEncoder", _Spk)
    return M


def _proc(text=None, **kw):
    body = [(ord(c) * 7) % 900 for c in text if c not in "<|>"][:12]
    return {"input_ids": torch.tensor([[1, 2, 3] + body + [4, 5, 6, 7, 8]])}


def _wav(path, n, sr=24000):
    from scipy.io import wavfile
    wavfile.write(path, sr, (np.sin(np.arange(n) * 0.05) * 8000).astype(np.int16))
    return str(path)


def test_tokenizer_example_call_shapes(stubbed, tmp_path):
    """examples/test_tokenizer_12hz.py:26-67."""
    write_tiny_checkpoint(str(tmp_path))
    tok = stubbed.Qwen3TTSTokenizer.from_pretrained(str(tmp_path / "speech_tokenizer"), device_map="cpu")
    a1, a2 = _wav(tmp_path / "a1.wav", 30000), _wav(tmp_path / "a2.wav", 50000, sr=16000)
    enc1 = tok.encode(a1)                                           # single path
    wavs1, sr1 = tok.decode(enc1)
    assert sr1 == 24000 and len(wavs1) == 1 and wavs1[0].shape == (16 * 1920,) and wavs1[0].dtype == np.float32
    enc2 = tok.encode([a1, a2])                                     # batch of paths (second one is resampled 16k -> 24k)
    assert [tuple(c.shape) for c in enc2.audio_codes] == [(16, 16), (40, 16)]
    wavs2, _ = tok.decode(enc2)
    assert [w.shape[0] for w in wavs2] == [16 * 1920, 40 * 1920]
    wd1, _ = tok.decode({"audio_codes": enc2.audio_codes[0]})       # dict
    wd2, _ = tok.decode([{"audio_codes": c} for c in enc2.audio_codes])                  # list[dict]
    wd3, _ = tok.decode([{"audio_codes": c.cpu().numpy()} for c in enc2.audio_codes])    # list[dict] with numpy
    assert len(wd1) == 1 and [w.shape for w in wd2] == [w.shape for w in wd3] == [w.shape for w in wavs2]
    y = np.zeros(24000 * 2, np.float32)
    enc3 = tok.encode(y, sr=24000)                                  # numpy + sr
    assert tuple(enc3.audio_codes[0].shape) == (25, 16)
    with pytest.raises(ValueError):
        tok.encode(y)
    assert (tok.get_model_type(), tok.get_input_sample_rate(), tok.get_output_sample_rate(),
            tok.get_encode_downsample_rate(), tok.get_decode_upsample_rate()) == ("qwen3_tts_tokenizer_12hz", 24000, 24000, 1920, 1920)


def test_voice_clone_example_call_shapes(stubbed, tmp_path):
    """examples/test_model_12hz_base.py:91-188: {single, batch} prompts x {single, batch} texts x {direct, prompt-then-
    generate} x {ICL, x-vector only}, with the reference's kwargs (dtype / attn_implementation included)."""
    write_tiny_checkpoint(str(tmp_path), model_type="base", spk_enc_dim=256)
    tts = stubbed.Qwen3TTSModel.from_pretrained(str(tmp_path), device_map="cpu", dtype=torch.bfloat16,
                                                attn_implementation="flash_attention_2", processor=_proc)
    ref1, ref2 = _wav(tmp_path / "r1.wav", 26000), _wav(tmp_path / "r2.wav", 30000)
    t1 = "Okay. Yeah. I resent you."
    tb = [t1, "a second reference transcript"]
    syn1, synb, langb = "Good one. Okay, fine.", ["Good one. Okay, fine.", "second sentence"], ["Chinese", "English"]
    kw = dict(max_new_tokens=2048, do_sample=True, top_k=50, top_p=1.0, temperature=0.9, repetition_penalty=1.05,
              subtalker_dosample=True, subtalker_top_k=50, subtalker_top_p=1.0, subtalker_temperature=0.9)
    for xv in (False, True):
        w, sr = tts.generate_voice_clone(text=syn1, language="Auto", ref_audio=ref1, ref_text=t1, x_vector_only_mode=xv, **kw)
        assert sr == 24000 and len(w) == 1 and w[0].ndim == 1
        items = tts.create_voice_clone_prompt(ref_audio=ref1, ref_text=t1, x_vector_only_mode=xv)
        assert len(items) == 1 and items[0].icl_mode == (not xv) and (items[0].ref_code is None) == xv
        w, _ = tts.generate_voice_clone(text=syn1, language="Auto", voice_clone_prompt=items, **kw)
        assert len(w) == 1
        w, _ = tts.generate_voice_clone(text=synb, language=langb, ref_audio=ref1, ref_text=t1, x_vector_only_mode=xv, **kw)
        assert len(w) == 2                                          # one prompt reused for a batch of texts
        w, _ = tts.generate_voice_clone(text=synb, language=langb, voice_clone_prompt=items, **kw)
        assert len(w) == 2
        w, _ = tts.generate_voice_clone(text=synb, language=langb, ref_audio=[ref1, ref2], ref_text=tb,
                                        x_vector_only_mode=[xv, xv], **kw)
        assert len(w) == 2
        items2 = tts.create_voice_clone_prompt(ref_audio=[ref1, ref2], ref_text=tb, x_vector_only_mode=[xv, xv])
        w, _ = tts.generate_voice_clone(text=synb, language=langb, voice_clone_prompt=items2, **kw)
        assert len(w) == 2 and all(x.dtype == np.float32 for x in w)
        # ICL output is cut proportionally: generated frames only (stub: 3 and 4 frames)
        assert [x.shape[0] for x in w] == [3 * 1920, 4 * 1920]


def test_custom_voice_and_voice_design_example_call_shapes(stubbed, tmp_path):
    """examples/test_model_12hz_custom_voice.py / _voice_design.py: single and batch, instruct optional."""
    write_tiny_checkpoint(str(tmp_path))
    tts = stubbed.Qwen3TTSModel.from_pretrained(str(tmp_path), device_map="cpu", processor=_proc)
    w, sr = tts.generate_custom_voice(text="hello", language="English", speaker="Alice", instruct="very happy")
    assert sr == 24000 and len(w) == 1
    w, _ = tts.generate_custom_voice(text=["a", "b"], language=["Chinese", "English"], speaker=["alice", "bob"],
                                     instruct=["", "slow"], max_new_tokens=64)
    assert len(w) == 2
    with pytest.raises(ValueError):
        tts.generate_voice_design(text="x", instruct="y")           # wrong model type
    assert tts.get_supported_speakers() == ["alice", "bob"] and "english" in tts.get_supported_languages()
    vd_dir = tmp_path / "vd"
    write_tiny_checkpoint(str(vd_dir), model_type="voice_design")
    vd = stubbed.Qwen3TTSModel.from_pretrained(str(vd_dir), device_map="cpu", processor=_proc)
    w, sr = vd.generate_voice_design(text="hello there", language="English", instruct="a calm, low voice")
    assert sr == 24000 and len(w) == 1
    w, _ = vd.generate_voice_design(text=["a", "b"], language=["Chinese", "English"], instruct=["warm", "bright"])
    assert len(w) == 2
    with pytest.raises(ValueError):
        vd.generate_custom_voice(text="x", speaker="alice")         # wrong model type the other way round
