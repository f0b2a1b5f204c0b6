"""End-to-end hot path on one GPU: prefill -> fused frame-step decode -> codec decode.

`TTSEngine.synthesize` is the call a user of seam B + seam C makes: prefill embeddings in (host or device),
waveforms out (host numpy, like Qwen3TTSTokenizer.decode -> `.to(float32).cpu().numpy()`,
inference/qwen3_tts_tokenizer.py:364)."""
from typing import List, Sequence

import numpy as np
import torch

from .codec import CodecDecoder
from .config import CodecConfig, SamplingParams, TTSConfig
from .engine import AREngine


class TTSEngine:
    def __init__(self, cfg: TTSConfig, weights, codec_cfg: CodecConfig, codec_weights, device="cuda:0", max_batch=32,
                 max_ctx=4096, codec_max_frames=1024):
        self.device = torch.device(device)
        self.ar = AREngine(cfg, weights, device=device, max_batch=max_batch, max_ctx=max_ctx)
        self.codec = CodecDecoder(codec_cfg, codec_weights, device=device, max_frames=codec_max_frames, max_batch=max_batch)
        self.cfg, self.codec_cfg = cfg, codec_cfg
        self.timing = {}

    @torch.no_grad()
    def generate_codes(self, inputs_embeds, trailing_text, tts_pad_embed, sp: SamplingParams) -> List[torch.Tensor]:
        return self.ar.generate(inputs_embeds, trailing_text, tts_pad_embed, sp)

    @torch.no_grad()
    def decode_codes(self, codes: Sequence[torch.Tensor]) -> List[torch.Tensor]:
        """List[(T_i,K)] -> List[(T_i*1920,)] fp32 on device; == Qwen3TTSTokenizer.decode's padding with -1 +
        Qwen3TTSTokenizerV2Model.decode (inference/qwen3_tts_tokenizer.py:329, …v2.py:993-1024)."""
        B = len(codes)
        Tm = max(max(int(c.shape[0]) for c in codes), 1)
        K = self.codec_cfg.num_quantizers
        ac = torch.full((B, Tm, K), -1, dtype=torch.int64, device=self.device)
        for i, c in enumerate(codes):
            ac[i, :c.shape[0]] = c.to(self.device)
        return self.codec.decode(ac)

    @torch.no_grad()
    def synthesize(self, inputs_embeds, trailing_text, tts_pad_embed, sp: SamplingParams, to_host=True):
        """inputs_embeds / trailing_text / tts_pad_embed may live in (pinned) host memory: they are copied to the
        device here, and the waveforms are copied back — both inside the caller's timed region."""
        dev = self.device
        emb = [e.to(dev, non_blocking=True) for e in inputs_embeds]
        tr = [t.to(dev, non_blocking=True) for t in trailing_text]
        pad = tts_pad_embed.to(dev, non_blocking=True)
        codes = self.ar.generate(emb, tr, pad, sp)
        wavs = self.decode_codes(codes)
        if not to_host:
            return wavs, codes
        out = [w.to(torch.float32).cpu().numpy() for w in wavs]
        return out, codes

    @torch.no_grad()
    def stream_synthesize(self, inputs_embeds, trailing_text, tts_pad_embed, sp: SamplingParams, packet_frames: int = 4,
                          left_context="stateful"):
        """Streaming OUTPUT (new surface: the reference returns audio whole).  Yields, per packet of
        `packet_frames` frames (4 frames = 320 ms, Qwen3-TTS report §3.4), a list with one fp32 numpy waveform chunk per
        row.  Default: the STATEFUL codec stream (q3_codec_stream_*) — each packet costs its own frames only and the
        concatenated chunks equal the one-shot causal decode.  `left_context=25` decodes every packet together with 25
        already-emitted frames, exactly like the reference's chunked_decode does between chunks (…v2.py:886-896);
        `left_context=None` re-decodes the whole prefix."""
        dev = self.device
        emb = [e.to(dev, non_blocking=True) for e in inputs_embeds]
        tr = [t.to(dev, non_blocking=True) for t in trailing_text]
        pad = tts_pad_embed.to(dev, non_blocking=True)
        K = self.codec_cfg.num_quantizers
        up = self.codec.total_upsample
        B = len(emb)
        if left_context == "stateful":
            key = (B, int(packet_frames))
            if getattr(self, "_cstream_key", None) != key:
                if getattr(self, "_cstream", None) is not None:
                    self._cstream.close()
                self._cstream = self.codec.open_stream(B, max_packet_frames=int(packet_frames))
                self._cstream_key = key
            cs = self._cstream
            cs.reset()
            for pkt in self.ar.stream(emb, tr, pad, sp, packet_frames=packet_frames):
                n_new = [int(p.shape[0]) for p in pkt]
                n = max(n_new) if n_new else 0
                if n == 0:
                    yield [np.zeros(0, dtype=np.float32) for _ in pkt]
                    continue
                # rows that already finished (or got fewer frames) are padded with code 0: their state no longer matters
                codes = torch.zeros(B, n, K, dtype=torch.int64, device=dev)
                for b, p in enumerate(pkt):
                    if p.shape[0]:
                        codes[b, :p.shape[0]] = p
                wav = cs.push(codes.transpose(1, 2).contiguous())[:, 0].to(torch.float32).cpu().numpy()
                yield [wav[b, :n_new[b] * up] for b in range(B)]
            return
        hist = [torch.zeros(0, K, dtype=torch.int64, device=dev) for _ in emb]
        for pkt in self.ar.stream(emb, tr, pad, sp, packet_frames=packet_frames):
            out = [np.zeros(0, dtype=np.float32) for _ in pkt]
            # rows whose window has the same shape (the normal case: every live row got `packet_frames` new frames on top
            # of the same history) go through ONE batched codec call — rows are independent in the decoder, so this is
            # bit-identical to decoding them one by one, at 1/B of the launches
            groups = {}
            for b, new in enumerate(pkt):
                if new.shape[0] == 0:
                    continue
                ctx = hist[b].shape[0] if left_context is None else min(left_context, hist[b].shape[0])
                groups.setdefault((ctx, int(new.shape[0])), []).append(b)
            for (ctx, n_new), rows in groups.items():
                windows = torch.stack([torch.cat([hist[b][hist[b].shape[0] - ctx:], pkt[b]], 0) for b in rows], 0)  # (n, T, K)
                wav = self.codec.forward(windows.transpose(1, 2).contiguous())[:, 0, ctx * up:]
                wav = wav.to(torch.float32).cpu().numpy()
                for i, b in enumerate(rows):
                    out[b] = wav[i]
            for b, new in enumerate(pkt):
                if new.shape[0]:
                    hist[b] = torch.cat([hist[b], new], 0)
            yield out

    def close(self):
        self.ar.close()
        self.codec.close()
