"""bench.py contract checks that do not need a GPU: the reference arm prints exactly ONE JSON line on stdout with the
contract's keys, and both arms share one metric / unit string (the driver divides the two lines)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--model", "tiny", "--steps", "2",
                        "--warmup", "1", "--batch", "2", "--frames", "6"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert src.count('"unit": UNIT') >= 4 and 'UNIT = "' in src          # every unit field of both arms is the one constant
    unit = src.split('UNIT = "', 1)[1].split('"', 1)[0]
    assert d["impl"] == "reference" and d["unit"] == unit and d["metric"] == "speech_tokens_per_s"
    for k in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
              "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["e2e"]["unit"] == unit and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["frames_sampled"] >= 2
    assert d["value"] > 0 and abs(d["value"] - cb["value"]) < 1e-9
