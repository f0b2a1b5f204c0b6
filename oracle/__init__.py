"""CPU oracle for the Qwen3-TTS 12 Hz hot paths.

TEST INFRASTRUCTURE ONLY.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline /
`--impl reference` legs may import this package; the product path (`qwen3-tts_b200/`) never does and fails
loudly if its CUDA library is missing.

The oracle is a plain-PyTorch restatement (no HF `generate`, no HF modules) of the reference's arithmetic for
  * the talker + code-predictor autoregressive loop   (qwen_tts/core/models/modeling_qwen3_tts.py)
  * the 12 Hz codec decoder                            (qwen_tts/core/tokenizer_12hz/modeling_qwen3_tts_tokenizer_v2.py)
  * the HF logits processors / sampling used by both   (transformers==4.57.3, third-party, restated)

Parity pinning: the reference ships NO tests, golden vectors or fixtures, so the
oracle is pinned against outputs of the reference's own modules run in the build container through
`oracle/ref_shims.py` (`tests/test_oracle_vs_reference.py`, skipped when /root/reference is absent) and
through committed fixtures minted by `oracle/make_golden.py` (`tests/golden/*.npz`).  The HF
`GenerationMixin` loop itself cannot run under the installed transformers 5.5.0, so the *loop* (not the
module arithmetic) is "parity unpinned" beyond code reading; see DESIGN.md §Oracle.
"""
