"""Logits processors + sampling: restatement of the HF semantics the reference relies on.

Third-party dependency (absent from /root/reference): transformers==4.57.3 (pyproject.toml:23) —
`RepetitionPenaltyLogitsProcessor`, `MinNewTokensLengthLogitsProcessor`, `SuppressTokensLogitsProcessor`,
`TemperatureLogitsWarper`, `TopKLogitsWarper`, `TopPLogitsWarper`, `GenerationMixin._sample`.
Reference call sites: modeling_qwen3_tts.py:2044-2066 (talker kwargs), :1671-1680 (code predictor kwargs),
:2272-2278 (talker.generate).  Order and formulas: SURVEY.md §A.4.

Sampling itself is inverse-CDF with an externally supplied uniform (oracle/philox.py) — documented as
differing from torch.multinomial's RNG stream.
"""
import numpy as np

NEG_INF = -np.inf


def process_logits(logits, *, generated_ids=(), repetition_penalty=1.0, min_new_tokens=0, eos_token_id=None,
                   suppress_lo=None, suppress_hi=None, do_sample=False, temperature=1.0, top_k=0, top_p=1.0):
    """logits: 1-D float array (one row).  Returns processed fp32 scores (with -inf for removed tokens)."""
    s = np.asarray(logits, dtype=np.float32).copy()
    V = s.shape[0]
    # 1. repetition penalty over *generated* tokens only (prompt is inputs_embeds => HF input_ids starts empty)
    if repetition_penalty != 1.0 and len(generated_ids):
        ids = np.unique(np.asarray(generated_ids, dtype=np.int64))
        v = s[ids]
        s[ids] = np.where(v < 0, v * np.float32(repetition_penalty), v / np.float32(repetition_penalty))
    # 2. min new tokens
    if eos_token_id is not None and len(generated_ids) < min_new_tokens:
        s[eos_token_id] = NEG_INF
    # 3. suppress tokens [lo, hi) \ {eos}
    if suppress_lo is not None:
        keep = s[eos_token_id] if (eos_token_id is not None and suppress_lo <= eos_token_id < suppress_hi) else None
        s[suppress_lo:suppress_hi] = NEG_INF
        if keep is not None:
            s[eos_token_id] = keep
    if do_sample:
        # 4a. temperature
        if temperature != 1.0:
            s = s / np.float32(temperature)
        # 4b. top-k (ties at the threshold are kept: HF removes scores < kth largest)
        if top_k and top_k > 0:
            k = min(int(top_k), V)
            kth = np.partition(s, V - k)[V - k]
            s = np.where(s < kth, np.float32(NEG_INF), s)
        # 4c. top-p (skipped at >= 1.0): ascending sort, drop tokens whose cumulative prob <= 1-p, keep >= 1
        if top_p < 1.0:
            order = np.argsort(s, kind="stable")
            ss = s[order]
            m = ss[np.isfinite(ss)].max()
            p = np.exp(ss - m, dtype=np.float32)
            p = p / p.sum(dtype=np.float32)
            cum = np.cumsum(p, dtype=np.float32)
            remove = cum <= np.float32(1.0 - top_p)
            remove[-1] = False
            s[order[remove]] = NEG_INF
    return s.astype(np.float32)


def sample_from_scores(scores, *, do_sample, u=None):
    """argmax (first max index, like torch.argmax) or inverse-CDF over softmax(scores) in token-id order.
    Returns the token id; for do_sample also returns the cdf for tolerance-aware checks."""
    s = np.asarray(scores, dtype=np.float32)
    if not do_sample:
        return int(np.argmax(s)), None
    m = s.max()
    p = np.exp(s - m, dtype=np.float32)  # exp(-inf)=0
    cdf = np.cumsum(p.astype(np.float64))
    target = float(u) * cdf[-1]
    idx = int(np.searchsorted(cdf, target, side="right"))
    idx = min(idx, len(s) - 1)
    # never select a zero-probability token
    while p[idx] == 0 and idx > 0:
        idx -= 1
    return idx, cdf / cdf[-1]
