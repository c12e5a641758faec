"""Generates tests/golden/ref_cases_{in,out}.bin: seeded inputs pushed through the UNMODIFIED reference's public API
by oracle/_ref/ref_harness (the reference's own CUDA code, so this must run on the GPU box):

    gpurun -- 'python tests/golden/make_golden.py && cp tests/golden/ref_cases_*.bin gpurun_out/'

The committed fixtures pin the CPU oracle (tests/test_golden.py, CPU) and the CUDA kernels (GPU)."""
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle as orc, refio  # noqa: E402


def build_inputs():
    rng = np.random.default_rng(20261018)

    def rand_fr(n):
        out = rng.integers(0, 1 << 32, size=(n, 8), dtype=np.uint64).astype(np.uint32)
        out[:, 7] %= 1944954707
        return out

    def small_fr(n, lim):
        return orc.fr_from_ints([int(v) for v in rng.integers(-lim, lim, size=n)], mont=True)

    def points(n):
        ks = orc.to_limbs([int.from_bytes(rng.bytes(31), "little") for _ in range(n)])
        return orc.g1_mul(orc.g1_generator(), ks, fast=True)

    box = {}
    a, b = rand_fr(64), rand_fr(64)
    a[:4] = orc.to_limbs([0, 1, orc.FR_P - 1, orc.FR_R]); b[:4] = orc.to_limbs([orc.FR_P - 1, orc.FR_P - 1, orc.FR_P - 1, 0])
    box["ops.a"], box["ops.b"], box["ops.x"] = a, b, rand_fr(1)
    box["fold.a"], box["fold.u"], box["fold.w"] = rand_fr(100), rand_fr(7), np.array([4, 3], np.uint32)
    box["sc.a"], box["sc.b"], box["sc.u"], box["sc.v"] = rand_fr(37), rand_fr(37), rand_fr(6), rand_fr(6)
    P, Q = points(16), points(16)
    Q[0] = P[0]; Q[1] = orc.g1_neg(P[1:2])[0]; Q[2, 24:] = 0
    x = rand_fr(16); x[0] = 0; x[1] = orc.to_limbs([1])[0]
    box["g1.p"], box["g1.q"], box["g1.x"], box["g1.u"] = P, Q, x, rand_fr(4)
    ng, m = 64, 64
    box["com.g"] = points(ng)
    t = small_fr(ng * m, 1 << 13); t[:ng] = rand_fr(ng)            # row 0 full-width, the rest quantised-weight sized
    box["com.t"] = t
    box["com.s"], box["com.u"] = rand_fr(ng), rand_fr(6)
    box["com.uo"] = rand_fr(12)
    B, I, O = 3, 13, 7
    w = (rng.uniform(-1, 1, size=(I, O)) / np.sqrt(I)).astype(np.float32)
    xin = rng.standard_normal((B, I)).astype(np.float32)
    w[0, :4] = [0.0, -0.0, 2.5 / 65536, -2.5 / 65536]
    box["fc.w"], box["fc.x"], box["fc.dims"] = w, xin, np.array([B, I, O, 16], np.uint32)
    return box


def main():
    box = build_inputs()
    inp, outp = os.path.join(HERE, "ref_cases_in.bin"), os.path.join(HERE, "ref_cases_out.bin")
    refio.write_box(inp, box)
    import torch
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(os.path.dirname(torch.__file__), "lib") + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    subprocess.check_call([os.path.join(ROOT, "oracle", "_ref", "ref_harness"), "run", inp, outp], env=env)
    out = refio.read_box(outp)
    print("reference produced", len(out), "arrays; cuda_status", out["cuda_status"])


if __name__ == "__main__":
    main()
