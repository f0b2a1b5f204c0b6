"""Per-launch view of one codec decode (8 x 125 frames by default).

Run twice on the GPU box:
  Q3_GEMM_TRACE=1 python tools/codec_breakdown.py --trace 2> gpurun_out/codec_trace.txt
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/codec_launches.csv python tools/codec_breakdown.py
then `python tools/codec_breakdown.py --join gpurun_out/codec_trace.txt gpurun_out/codec_launches.csv` (CPU) prints, per
tap-GEMM launch, shape, duration, TFLOP/s and the share of the whole decode.
"""
import argparse, csv, os, re, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def run(trace):
    import torch
    from qwen3_tts_b200 import synthetic
    from qwen3_tts_b200.codec import CodecDecoder
    dev = "cuda:0"
    from qwen3_tts_b200.config import CodecConfig
    cfg = CodecConfig()
    W = synthetic.random_codec_weights(cfg, device="cpu", seed=0)
    dec = CodecDecoder(cfg, W, device=dev, max_batch=8, max_frames=128)
    codes = torch.randint(0, 2048, (8, 16, 125), dtype=torch.int32, device=dev)
    for _ in range(2):
        dec.forward(codes)
    torch.cuda.synchronize()
    if trace:
        sys.stderr.write("[trace-begin]\n"); sys.stderr.flush()
    torch.cuda.cudart().cudaProfilerStart()
    dec.forward(codes)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    if trace:
        sys.stderr.write("[trace-end]\n"); sys.stderr.flush()


def join(trace_path, csv_path):
    txt = open(trace_path).read()
    txt = txt[txt.rindex("[trace-begin]"):]
    shapes = [dict((k, int(v)) for k, v in re.findall(r"(\w+)=(-?\d+)", l)) for l in txt.splitlines() if l.startswith("[tap_gemm]")]
    rows = []
    with open(csv_path) as f:
        lines = [l for l in f if not l.startswith("==")]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            v = float(r["Metric Value"].replace(",", ""))
            unit = r["Metric Unit"]
            us = v / 1e3 if unit in ("ns", "nsecond") else v if unit in ("us", "usecond") else v * 1e3
            rows.append((r["Kernel Name"], us))
    total = sum(us for _, us in rows)
    gi = 0
    print(f"{len(rows)} launches, {total / 1e3:.2f} ms (serialised, cold-cache ncu times); {len(shapes)} tap-GEMM launches traced")
    agg = {}
    for name, us in rows:
        short = name.split("(")[0]
        if "tap_gemm" in name and gi < len(shapes):
            s = shapes[gi]; gi += 1
            fl = 2.0 * s["B"] * s["T"] * s["N"] * s["Kp"] * s["taps"]
            print(f"  gemm B={s['B']} T={s['T']:6d} N={s['N']:5d} K={s['Kp']:5d}x{s['taps']} bn={s['bn']:3d} tiles={s['tiles']:5d} act={s['act']} "
                  f"{us:8.1f} us {fl / us / 1e6:7.1f} TFLOP/s {100 * us / total:5.1f}%")
        a = agg.setdefault(short, [0, 0.0]); a[0] += 1; a[1] += us
    print("by kernel:")
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"  {k:40s} {n:4d} launches {us / 1e3:8.3f} ms {100 * us / total:5.1f}%")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--trace", action="store_true")
    ap.add_argument("--join", nargs=2)
    a = ap.parse_args()
    if a.join:
        join(*a.join)
    else:
        run(a.trace)
