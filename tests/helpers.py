"""Shared helpers for the parity tests (test infrastructure; may import oracle/)."""
import numpy as np
import torch

from oracle import talker as OT

LOGIT_TOL = 0.06      # |engine - fp32 oracle| on raw logits (logit std ~0.8 with the seeded weights); the
                      # bf16-vs-fp32 gap of the *oracle itself* is ~0.03


def to_pkg_cfg(ocfg: OT.TTSCfg):
    import qwen3_tts_b200 as q

    def st(s):
        return q.StackConfig(s.hidden_size, s.num_layers, s.num_heads, s.num_kv_heads, s.head_dim,
                             s.intermediate_size, s.vocab_size, s.rms_eps, s.rope_theta)
    return q.TTSConfig(talker=st(ocfg.talker), cp=st(ocfg.cp), num_code_groups=ocfg.num_code_groups,
                       text_hidden_size=ocfg.text_hidden_size, text_vocab_size=ocfg.text_vocab_size,
                       codec_eos_token_id=ocfg.codec_eos_token_id, codec_pad_id=ocfg.codec_pad_id,
                       codec_bos_id=ocfg.codec_bos_id, tts_bos_token_id=ocfg.tts_bos_token_id,
                       tts_eos_token_id=ocfg.tts_eos_token_id, tts_pad_token_id=ocfg.tts_pad_token_id)


def to_pkg_sampling(sp: OT.SamplingCfg):
    import qwen3_tts_b200 as q
    return q.SamplingParams(**{k: getattr(sp, k) for k in q.SamplingParams.__dataclass_fields__})


def bf16_weights(W):
    """bf16-rounded weights: (bf16 dict for the engine, fp32 upcast of the SAME values for the oracle)."""
    Wb = {k: v.to(torch.bfloat16) for k, v in W.items()}
    Wf = {k: v.to(torch.float32) for k, v in Wb.items()}
    return Wb, Wf


def make_inputs(cfg, lens, trail_lens, seed=0):
    g = torch.Generator().manual_seed(seed)
    H = cfg.talker.hidden_size
    embs = [(torch.randn(l, H, generator=g) * 0.5).bfloat16() for l in lens]
    trail = [(torch.randn(n, H, generator=g) * 0.1).bfloat16() for n in trail_lens]
    pad = (torch.randn(H, generator=g) * 0.1).bfloat16()
    return embs, trail, pad


def run_engine_forced(eng, embs, trail, pad, sp_pkg, forced: np.ndarray, dev):
    """Teacher-forced run with raw-logit capture.  forced: (B, N, G) int."""
    B, N, G = forced.shape
    V, Vc = eng.cfg.talker.vocab_size, eng.cfg.cp.vocab_size
    f = torch.from_numpy(forced.astype(np.int32)).to(dev).contiguous()
    tl = torch.zeros(N + 1, B, V, dtype=torch.float32, device=dev)
    cl = torch.zeros(N, G - 1, B, Vc, dtype=torch.float32, device=dev)
    eng.set_debug(f, N, tl, cl)
    eng.prefill(embs, trail, pad, sp_pkg)
    codes = torch.zeros(B, N, G, dtype=torch.int32, device=dev)
    eng.decode(N, codes)
    torch.cuda.synchronize()
    prog = eng.progress()
    eng.set_debug(None, 0, None, None)
    return codes.cpu().numpy(), tl.cpu().numpy(), cl.cpu().numpy(), prog
