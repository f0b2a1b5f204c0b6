"""N > 1 host logic on CPU: world_size-2 gloo, batch split + variable-length gather."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from qwen3_tts_b200 import parallel as P
    reqs = [dict(id=i, n=3 + (i * 7) % 5) for i in range(7)]
    seen = []

    def fn(rs):
        seen.extend(r["id"] for r in rs)
        return [np.full(r["n"], r["id"], dtype=np.float32) for r in rs]  # variable-length "waveforms"

    out = P.run_data_parallel(fn, reqs)
    ok = all(o.shape == (r["n"],) and (o == r["id"]).all() for o, r in zip(out, reqs))
    out2 = P.run_data_parallel(fn, reqs, lengths=[r["n"] for r in reqs])
    ok = ok and all((o == r["id"]).all() for o, r in zip(out2, reqs))
    t = P.max_over_ranks(float(rank + 1), device="cpu")
    q.put((rank, ok, sorted(seen[:len(seen) // 2]) if False else len(seen), t))
    dist.destroy_process_group()


def test_world2_split_and_gather():
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    os.environ["PYTHONPATH"] = root + os.pathsep + os.environ.get("PYTHONPATH", "")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
    assert all(ok for _, ok, _, _ in res)
    assert sum(n for _, _, n, _ in res) == 14  # each request ran exactly once per call (2 calls x 7)
    assert all(t == 2.0 for *_, t in res)


def test_shard_bounds_and_balance():
    from qwen3_tts_b200 import parallel as P
    for n in (0, 1, 7, 8, 32):
        for w in (1, 2, 4, 8):
            b = [P.shard_bounds(n, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1
    bins = P.length_balanced_order([100, 1, 1, 1, 50, 50], 2)
    assert sorted(sum(bins, [])) == list(range(6))
    loads = [sum([100, 1, 1, 1, 50, 50][i] for i in b) for b in bins]
    assert abs(loads[0] - loads[1]) <= 3
