"""CPU study behind the codec encoder's precision choice: how many RVQ codes flip, relative to the fp32
oracle, if the encoder's GEMM-shaped work ran with tf32 or bf16 operands (fp32 accumulate)?  Random weights, default
Mimi shapes, 4 x 2 s of noise = 100 code frames.  Usage: python tools/encoder_precision_study.py"""
import inspect
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from oracle import mimi_encoder as M  # noqa: E402


def tf32(x):  # round-to-nearest-even to 10 explicit mantissa bits
    i = x.contiguous().view(torch.int32)
    return ((i + (((i >> 13) & 1) + 0x0FFF)) & ~0x1FFF).view(torch.float32)


def bf16(x):
    return x.to(torch.bfloat16).float()


def rounded_transformer(rf):
    src = inspect.getsource(M.encoder_transformer)
    src = re.sub(r"\(h @ W\[(.*?)\]\.T\)", r"(RF(h) @ RF(W[\1]).T)", src)
    src = src.replace('o = (a @ v).transpose(1, 2).reshape(B, T, nh * hd) @ W[p + "self_attn.o_proj.weight"].T',
                      'o = RF((a @ v).transpose(1, 2).reshape(B, T, nh * hd)) @ RF(W[p + "self_attn.o_proj.weight"]).T')
    src = src.replace('h = F.gelu(h @ W[p + "mlp.fc1.weight"].T) @ W[p + "mlp.fc2.weight"].T',
                      'h = RF(F.gelu(RF(h) @ RF(W[p + "mlp.fc1.weight"]).T)) @ RF(W[p + "mlp.fc2.weight"]).T')
    assert src.count("RF(") == 10, src.count("RF(")
    ns = dict(M.__dict__)
    ns["RF"] = rf
    exec(src, ns)
    return ns["encoder_transformer"]


def main():
    cfg = M.MimiEncCfg()
    W = M.random_weights(cfg, seed=5)
    g = torch.Generator().manual_seed(0)
    wav = (torch.randn(4, 48000, generator=g) * 0.1).clamp(-1, 1)
    ref = M.encode(W, cfg, wav)
    conv0, tr0 = M.mimi_conv1d, M.encoder_transformer
    print(f"{'operands':28s} frames-with-a-flip  codebook-0 flips  all codes")
    for name, rf, with_tr in (("tf32: convs", tf32, False), ("tf32: convs + linears", tf32, True),
                              ("bf16: convs", bf16, False), ("bf16: convs + linears", bf16, True)):
        M.mimi_conv1d = (lambda rf: (lambda x, w, b, stride=1, dilation=1, pad_mode="constant":
                                     conv0(rf(x), rf(w), b, stride=stride, dilation=dilation, pad_mode=pad_mode)))(rf)
        M.encoder_transformer = rounded_transformer(rf) if with_tr else tr0
        d = M.encode(W, cfg, wav) != ref
        M.mimi_conv1d, M.encoder_transformer = conv0, tr0
        print(f"{name:28s} {float(d.any(1).float().mean()):17.3f} {float(d[:, 0].float().mean()):17.3f} {float(d.float().mean()):10.4f}")


if __name__ == "__main__":
    main()
