"""Mint golden vectors from the REFERENCE's own modules (through oracle/ref_shims.py) and commit them under
tests/golden/.  Run in the build container only:  python -m oracle.make_golden

The reference ships no fixtures; these pin the oracle wherever /root/reference is absent (GPU box).
Shapes are deliberately micro (weights travel inside the .npz)."""
import os

import numpy as np
import torch

from . import codec as OC
from . import mimi_encoder as OM
from . import speaker_encoder as OS
from . import ref_driver as R
from . import talker as OT

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def micro_tts_cfg():
    V = 1200
    return OT.TTSCfg(talker=OT.StackCfg(64, 2, 2, 1, 32, 128, V, rope_theta=1e6),
                     cp=OT.StackCfg(32, 2, 2, 1, 32, 64, 64, rope_theta=1e4), text_hidden_size=64, text_vocab_size=100,
                     codec_eos_token_id=V - 10, codec_pad_id=V - 12, codec_bos_id=V - 11, tts_bos_token_id=97,
                     tts_eos_token_id=98, tts_pad_token_id=96)


def micro_codec_cfg():
    return OC.CodecCfg(codebook_size=32, codebook_dim=16, hidden_size=32, latent_dim=32, num_heads=2, num_kv_heads=2,
                       head_dim=16, sliding_window=4, intermediate_size=48, num_layers=2, num_quantizers=16,
                       upsample_rates=(8, 5, 4, 3), upsampling_ratios=(2, 2), decoder_dim=48)


def make_talker():
    from transformers.cache_utils import DynamicCache
    cfg = micro_tts_cfg()
    W = OT.random_weights(cfg, seed=11, with_text=False)
    m = R.build_reference_talker(cfg, text_vocab=100)
    R.load_weights_into_reference(m, {k: v for k, v in W.items()})
    g = torch.Generator().manual_seed(5)
    H, B, lens, N, G = 64, 2, [4, 7], 4, 16
    embs = [torch.randn(l, H, generator=g) * 0.5 for l in lens]
    trail = [torch.randn(n, H, generator=g) * 0.1 for n in (2, 0)]
    pad = torch.randn(H, generator=g) * 0.1
    codes = torch.randint(0, 64, (B, N, G), generator=g)
    codes[:, :, 0] = torch.randint(0, 150, (B, N), generator=g)
    Lmax = max(lens)
    x = torch.zeros(B, Lmax, H)
    mask = torch.zeros(B, Lmax, dtype=torch.long)
    for i, e in enumerate(embs):
        x[i, Lmax - len(e):] = e
        mask[i, Lmax - len(e):] = 1
    cache = DynamicCache()
    m.rope_deltas = None
    tl, cl = [], []
    with torch.no_grad():
        out = m(inputs_embeds=x, attention_mask=mask, past_key_values=cache, use_cache=True, cache_position=torch.arange(Lmax))
        tl.append(out.logits[:, -1].numpy().copy())
        past_hidden = out.past_hidden
        for step in range(N):
            c0 = codes[:, step, 0]
            cpc = DynamicCache()
            e0 = m.get_input_embeddings()(c0[:, None])
            o = m.code_predictor(inputs_embeds=torch.cat((past_hidden, e0), dim=1), past_key_values=cpc, use_cache=True)
            cl.append(o.logits[:, -1].numpy().copy())
            gs = o.generation_steps
            for j in range(1, G - 1):
                o = m.code_predictor(input_ids=codes[:, step, j:j + 1], past_key_values=cpc, use_cache=True, generation_steps=gs)
                gs = o.generation_steps
                cl.append(o.logits[:, -1].numpy().copy())
            hid = [e0] + [m.code_predictor.get_input_embeddings()[i](codes[:, step, i + 1:i + 2]) for i in range(G - 1)]
            xe = torch.cat(hid, dim=1).sum(1, keepdim=True)
            tr = torch.stack([t[step] if step < t.shape[0] else pad for t in trail])[:, None]
            xe = xe + tr
            mask = torch.cat((mask, torch.ones(B, 1, dtype=torch.long)), dim=1)
            cp = torch.tensor([Lmax + step])
            pos = (cp[0] + m.rope_deltas).view(1, B, 1).expand(3, -1, -1)
            mo = m.model(inputs_embeds=xe, attention_mask=mask, position_ids=pos, past_key_values=cache, use_cache=True, cache_position=cp)
            past_hidden = mo.last_hidden_state[:, -1:]
            tl.append(m.codec_head(mo.last_hidden_state)[:, -1].numpy().copy())
    blob = {f"W::{k}": v.numpy() for k, v in W.items()}
    blob.update({f"emb{i}": e.numpy() for i, e in enumerate(embs)})
    blob.update({f"trail{i}": t.numpy() for i, t in enumerate(trail)})
    blob.update(pad=pad.numpy(), codes=codes.numpy(), talker_logits=np.stack(tl), cp_logits=np.stack(cl))
    np.savez_compressed(os.path.join(OUT, "talker_micro.npz"), **blob)
    print("talker_micro.npz", np.stack(tl).shape, np.stack(cl).shape)


def make_codec():
    cfg = micro_codec_cfg()
    W = OC.random_weights(cfg, seed=13)
    m = R.build_reference_codec_decoder(cfg)
    m.load_state_dict(W, strict=False)
    g = torch.Generator().manual_seed(2)
    codes = torch.randint(0, cfg.codebook_size, (2, 16, 9), generator=g)
    with torch.no_grad():
        wav = m(codes)
        wav_c = m.chunked_decode(codes, chunk_size=4, left_context_size=2)
    blob = {f"W::{k}": v.numpy() for k, v in W.items()}
    blob.update(codes=codes.numpy(), wav=wav.numpy(), wav_chunked=wav_c.numpy())
    np.savez_compressed(os.path.join(OUT, "codec_micro.npz"), **blob)
    print("codec_micro.npz", wav.shape, float(wav.abs().max()))


def micro_encoder_cfg():
    return OM.MimiEncCfg(num_filters=4, hidden_size=32, num_layers=2, num_heads=2, head_dim=16, intermediate_size=48,
                         sliding_window=5, codebook_size=32, codebook_dim=16, num_quantizers=32, valid_num_quantizers=16)


def make_encoder():
    """Golden codes from the third-party encoder the reference wraps: transformers MimiModel (installed 5.5.0; the
    reference pins 4.57.3) driven exactly like Qwen3TTSTokenizerV2Model.encode (…v2.py:977-983)."""
    from transformers import MimiConfig, MimiModel
    cfg = micro_encoder_cfg()
    W = OM.random_weights(cfg, seed=17)
    hf = MimiModel(MimiConfig(**cfg.to_hf_kwargs())).eval()
    missing, unexpected = hf.load_state_dict(W, strict=False)
    assert not unexpected, unexpected
    g = torch.Generator().manual_seed(4)
    wav = (torch.randn(2, 15000, generator=g) * 0.1).clamp(-1, 1)  # 16 transformer frames > window 5, 8 code frames
    with torch.no_grad():
        codes = hf.encode(wav[:, None, :], return_dict=True).audio_codes[:, : cfg.valid_num_quantizers]
    blob = {f"W::{k}": v.numpy() for k, v in W.items()}
    blob.update(wav=wav.numpy(), codes=codes.numpy())
    np.savez_compressed(os.path.join(OUT, "encoder_micro.npz"), **blob)
    print("encoder_micro.npz", tuple(codes.shape))


def make_speaker():
    """Golden x-vectors from the reference's own Qwen3TTSSpeakerEncoder (modeling_qwen3_tts.py:300-393) and log-mels from
    its mel_spectrogram (:396-448; the absent librosa filterbank is supplied by oracle.speaker_encoder)."""
    from . import ref_shims
    ref_shims.install()
    from qwen_tts.core.models import modeling_qwen3_tts as RM
    from qwen_tts.core.models.configuration_qwen3_tts import Qwen3TTSSpeakerEncoderConfig
    cfg = OS.cfg_tiny_spk()
    rc = Qwen3TTSSpeakerEncoderConfig(mel_dim=cfg.mel_dim, enc_dim=cfg.enc_dim, enc_channels=list(cfg.enc_channels),
                                      enc_kernel_sizes=list(cfg.enc_kernel_sizes), enc_dilations=list(cfg.enc_dilations),
                                      enc_attention_channels=cfg.enc_attention_channels,
                                      enc_res2net_scale=cfg.enc_res2net_scale, enc_se_channels=cfg.enc_se_channels)
    m = RM.Qwen3TTSSpeakerEncoder(rc).eval()
    W = OS.random_weights(cfg, seed=19)
    m.load_state_dict(W)
    g = torch.Generator().manual_seed(6)
    wav = (torch.randn(2, 6000, generator=g) * 0.1).clamp(-1, 1)
    RM.librosa_mel_fn = lambda sr, n_fft, n_mels, fmin, fmax: OS.slaney_mel_filterbank(sr, n_fft, n_mels, fmin, fmax)
    mel = RM.mel_spectrogram(wav, n_fft=1024, num_mels=cfg.mel_dim, sampling_rate=24000, hop_size=256, win_size=1024,
                             fmin=0, fmax=12000)
    with torch.no_grad():
        emb = m(mel.transpose(1, 2))
    blob = {f"W::{k}": v.numpy() for k, v in W.items()}
    blob.update(wav=wav.numpy(), mel=mel.numpy(), emb=emb.numpy())
    np.savez_compressed(os.path.join(OUT, "speaker_micro.npz"), **blob)
    print("speaker_micro.npz", tuple(mel.shape), tuple(emb.shape))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    import sys
    which = sys.argv[1:] or ["talker", "codec", "encoder", "speaker"]
    if "talker" in which:
        make_talker()
    if "codec" in which:
        make_codec()
    if "encoder" in which:
        make_encoder()
    if "speaker" in which:
        make_speaker()
