"""Property-based checks (hypothesis) of the host-side invariants the multi-GPU split and the checkpoint reader rely on."""
import os

import pytest
import torch
from hypothesis import given, settings, strategies as st


@settings(max_examples=200, deadline=None)
@given(n=st.integers(0, 500), world=st.integers(1, 16))
def test_shard_bounds_partition(n, world):
    from qwen3_tts_b200 import parallel as P
    b = [P.shard_bounds(n, world, r) for r in range(world)]
    assert b[0][0] == 0 and b[-1][1] == n
    assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))           # contiguous, no gaps, no overlap
    sizes = [h - l for l, h in b]
    assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)


@settings(max_examples=200, deadline=None)
@given(lengths=st.lists(st.integers(1, 400), min_size=0, max_size=40), world=st.integers(1, 8))
def test_length_balanced_order_is_a_partition_with_lpt_bound(lengths, world):
    from qwen3_tts_b200 import parallel as P
    bins = P.length_balanced_order(lengths, world)
    assert len(bins) == world and sorted(i for b in bins for i in b) == list(range(len(lengths)))
    assert all(b == sorted(b) for b in bins)
    if lengths:
        loads = [sum(lengths[i] for i in b) for b in bins]
        # longest-processing-time-first: no rank exceeds the mean by more than one (largest) item
        assert max(loads) <= sum(lengths) / world + max(lengths)


@settings(max_examples=25, deadline=None)
@given(cut=st.lists(st.integers(0, 9), min_size=1, max_size=3, unique=True), seed=st.integers(0, 5))
def test_sharded_safetensors_read_back_identically(tmp_path_factory, cut, seed):
    """Any split of a state_dict over safetensors shards + index file reads back to the same tensors and dtypes."""
    import json
    from safetensors.torch import save_file
    from qwen3_tts_b200 import checkpoint
    d = tmp_path_factory.mktemp("ckpt")
    g = torch.Generator().manual_seed(seed)
    W = {f"talker.t{i}.weight": (torch.randn(3 + i, 4, generator=g)).to(torch.bfloat16 if i % 2 else torch.float32) for i in range(10)}
    W["talker.ids"] = torch.arange(7)
    keys = sorted(W)
    bounds = [0] + sorted(cut) + [len(keys)]
    wm = {}
    for j in range(len(bounds) - 1):
        part = keys[bounds[j]:bounds[j + 1]]
        if not part:
            continue
        fn = f"model-{j:05d}.safetensors"
        save_file({k: W[k].contiguous() for k in part}, os.path.join(d, fn))
        wm.update({k: fn for k in part})
    json.dump({"metadata": {}, "weight_map": wm}, open(os.path.join(d, "model.safetensors.index.json"), "w"))
    got = checkpoint.read_state_dict(str(d), device="cpu", prefixes=("talker.",))
    assert set(got) == set(W) and all(got[k].dtype == W[k].dtype and torch.equal(got[k], W[k]) for k in W)
    cast = checkpoint.read_state_dict(str(d), device="cpu", dtype=torch.bfloat16)
    assert all(v.dtype == (torch.bfloat16 if W[k].is_floating_point() else W[k].dtype) for k, v in cast.items())
