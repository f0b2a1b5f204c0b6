"""CPU model check of the frame-step kernel's weight-ring geometry (csrc/ar_ring.cuh is shared by host and device):
producer and consumer enumerate identical pieces for every CTA / warp, and the pieces tile every weight matrix
exactly once.  Shapes: the frame programs of the 1.7B / 0.6B / tiny configurations on 148 SMs and on odd grids."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def model_bin(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("ring") / "ring_model")
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", out, os.path.join(ROOT, "tests", "cpp", "ring_model.cpp")], check=True)
    return out


def frame_program(H, I, Hc, Ic, L, Lc, V, Vc, G=16, nh=16, nkv=8, hd=128, proj=False):
    """(n_tiles, kb) per phase of the frame program (0,0 = attention / sample phase), as build_programs lays it out."""
    def layers(hid, inter, n):
        out = []
        for _ in range(n):
            out += [((nh + 2 * nkv) * hd // 16, hid // 32), (0, 0), (hid // 16, nh * hd // 32), (2 * inter // 16, hid // 32),
                    (hid // 16, inter // 32)]
        return out
    ph = []
    for j in range(G - 1):
        if j == 0 and proj:
            ph.append((Hc // 16, H // 32))
        ph += layers(Hc, Ic, Lc) + [(Vc // 16, Hc // 32), (0, 0)]
    ph += layers(H, I, L) + [(V // 16, H // 32), (0, 0)]
    return ph


CASES = {
    "1.7b": frame_program(2048, 6144, 1024, 3072, 28, 5, 3072, 2048, proj=True),
    "0.6b": frame_program(1024, 3072, 1024, 3072, 28, 5, 3072, 2048),
    "tiny": frame_program(256, 512, 128, 256, 2, 2, 3072, 2048, G=4, nh=4, nkv=2, proj=True),
    "odd": [(7, 3), (0, 0), (1, 1), (333, 5), (20, 64), (9, 192)],
}


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("grid,sb,r,niter", [(148, 4, 3, 2), (148, 2, 2, 1), (132, 4, 4, 2), (16, 1, 5, 3)])
def test_ring_producer_matches_consumer(model_bin, name, grid, sb, r, niter):
    ph = CASES[name]
    args = [model_bin, str(grid), str(niter), str(sb), str(r), str(len(ph))]
    for nt, kb in ph:
        args += [str(nt), str(kb)]
    p = subprocess.run(args, capture_output=True, text=True)
    assert p.returncode == 0 and p.stdout.startswith("OK"), p.stdout + p.stderr
