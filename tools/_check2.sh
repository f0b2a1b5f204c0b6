set -u
mkdir -p gpurun_out
python - <<'PY' 2>&1 | tail -12
import sys, json, torch
sys.argv = ["bench.py"]
import bench, qwen3_tts_b200 as q
from qwen3_tts_b200 import synthetic
from qwen3_tts_b200.pipeline import TTSEngine
from qwen3_tts_b200.config import CodecConfig
class A: frames = 125
dev = "cuda:0"; torch.cuda.set_device(0)
cfg = bench.model_cfg("1.7b"); ccfg = CodecConfig()
W = synthetic.random_tts_weights(cfg, device="cpu", seed=0); CW = synthetic.random_codec_weights(ccfg, device="cpu", seed=0)
eng = TTSEngine(cfg, W, ccfg, CW, device=dev, max_batch=32, max_ctx=260, codec_max_frames=125 + 8 + bench.REF_FRAMES)
spk = dict(do_sample=True, top_k=50, top_p=1.0, temperature=0.9, repetition_penalty=1.05, subtalker_dosample=True, subtalker_top_k=50, subtalker_top_p=1.0, subtalker_temperature=0.9)
import inspect
try:
    spk = bench.SPK if hasattr(bench, "SPK") else spk
except Exception: pass
for B in (4, 32):
    try:
        r = bench.voice_clone_probe(eng, q, cfg, W, A, spk, B, 125, dev)
        print(B, json.dumps({k: r[k] for k in ("ms_total_wall", "ms", "frames_per_s", "rtf")}))
    except Exception as e:
        import traceback; traceback.print_exc()
PY
for f in 0 4; do Q3_FLAGS=$f timeout 200 python bench.py --batch 1 --steps 2 --warmup 1 --no-extras --no-cpu-baseline --no-parity-check 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('B=1 flags=$f', round(d['roofline']['ms_per_frame_step'],3), round(d['value'],1))"; done
