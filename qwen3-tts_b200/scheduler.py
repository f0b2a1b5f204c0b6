"""Continuous batching over the AR engine.

The reference serves one static, left-padded batch per `generate()` call (modeling_qwen3_tts.py:2239-2254) behind a
Gradio queue (cli/demo.py:629): a request that arrives while a batch is running waits for the whole batch, and a
batch runs until its LONGEST utterance ends.  Here a session of `n_slots` rows runs continuously; a request is
prefilled into a free slot between two decode chunks (`packet_frames` frame-steps) while the other rows keep their
K/V, positions, sampling history and Philox streams, and leaves as soon as it samples EOS.  Every per-row quantity of
the fused kernel is keyed by the row's own frame counter and by the request's Philox key, so a request generates
exactly what it would generate alone (tests/test_gpu_scheduler.py).
"""
from collections import deque
from dataclasses import dataclass, field
from typing import Deque, Dict, List, Optional

import torch

from .config import SamplingParams
from .engine import AREngine


@dataclass
class _Request:
    rid: int
    embeds: torch.Tensor
    trailing: torch.Tensor
    max_frames: int
    slot: int = -1
    emitted: int = 0
    chunks: List[torch.Tensor] = field(default_factory=list)


class ContinuousBatcher:
    def __init__(self, engine: AREngine, tts_pad_embed: torch.Tensor, sp: SamplingParams, n_slots: int = 8, packet_frames: int = 4,
                 max_trailing: int = 256):
        self.eng, self.sp = engine, sp
        self.n_slots, self.packet = int(n_slots), int(packet_frames)
        self.G = engine.cfg.num_code_groups
        self.max_frames = max(int(sp.max_new_tokens) - 1, 1)
        self.codes = torch.zeros(self.n_slots, self.max_frames, self.G, dtype=torch.int32, device=engine.device)
        self.pending: Deque[_Request] = deque()
        self.running: Dict[int, _Request] = {}     # slot -> request
        self.done: Dict[int, torch.Tensor] = {}
        self._next_id = 0
        self.max_trailing = int(max_trailing)
        engine.session_begin(self.n_slots, tts_pad_embed, sp, max_trailing=self.max_trailing)
        self.frames_total = 0

    # ------------------------------------------------------------------ API
    def submit(self, inputs_embeds: torch.Tensor, trailing_text: Optional[torch.Tensor] = None, key: Optional[int] = None,
               max_frames: Optional[int] = None) -> int:
        """Queue one request; returns its id (also its Philox row key unless `key` is given)."""
        H = self.eng.cfg.talker.hidden_size
        rid = self._next_id if key is None else int(key)
        self._next_id = max(self._next_id, rid) + 1
        tr = trailing_text if trailing_text is not None else torch.zeros(0, H)
        if tr.reshape(-1, H).shape[0] > self.max_trailing:
            raise ValueError(f"trailing text of {tr.shape[0]} positions exceeds max_trailing={self.max_trailing}")
        L = int(inputs_embeds.reshape(-1, H).shape[0])
        room = self.eng.max_ctx - L
        if room < 1:
            raise ValueError(f"prompt of {L} positions does not fit max_ctx={self.eng.max_ctx}")
        self.pending.append(_Request(rid, inputs_embeds, tr, min(max_frames or self.max_frames, self.max_frames, room)))
        return rid

    def step(self) -> List[int]:
        """Admit what fits, decode one packet, collect what finished.  Returns the ids that completed in this step."""
        free = [s for s in range(self.n_slots) if s not in self.running]
        batch = []
        while free and self.pending:
            r = self.pending.popleft()
            r.slot = free.pop(0)
            batch.append(r)
        if batch:
            self.eng.admit([r.slot for r in batch], [r.rid for r in batch], [r.embeds for r in batch], [r.trailing for r in batch])
            for r in batch:
                self.running[r.slot] = r
        if not self.running:
            return []
        n = min([self.packet] + [r.max_frames - r.emitted for r in self.running.values()])
        self.eng.decode(max(n, 1), self.codes)
        torch.cuda.current_stream(self.eng.device).synchronize()
        fd, n_valid, fin = self.eng.progress()
        self.frames_total = fd
        finished, give_up = [], []
        for slot, r in list(self.running.items()):
            have = min(n_valid[slot], r.max_frames)
            r.emitted = have
            if fin[slot] or have >= r.max_frames:
                self.done[r.rid] = self.codes[slot, :have].to(torch.int64).clone()
                finished.append(r.rid)
                if not fin[slot]:
                    give_up.append(slot)  # horizon reached without EOS: the row is abandoned, its slot is free again
                del self.running[slot]
        if give_up:
            self.eng.release_slots(give_up)
        return finished

    def run(self) -> Dict[int, torch.Tensor]:
        """Drain the queue; returns {request id: (N_i, G) codes}."""
        while self.pending or self.running:
            self.step()
        return self.done
