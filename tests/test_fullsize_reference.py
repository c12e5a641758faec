"""Full-size differential parity against the UNMODIFIED reference (BASELINE.json configs 2-5), live on the GPU box.

`oracle/_ref/ref_harness full <case> ...` (oracle/ref_harness.cu linked against the reference's own objects) regenerates
seeded inputs, runs the reference's public API (partial_me, inner_product_sumcheck, Fr_me, G1_me, Commitment::me_open,
zkReLU::operator(), binary/hadamard sumchecks; /root/reference/zkfc.cu:128-145, zkrelu.cu:79-100, commitment.cu:43-92) and
dumps every proof element.  Here the same inputs are regenerated (numpy's legacy MT19937 seeding == std::mt19937(seed)),
pushed through the C ABI (the routes bench.py times: integer weight folds, packed zkReLU, batched-MSM opening) and
compared: Fr bit for bit, G1 as points.  The same harness source built against the drop-in headers
(zkdl_b200/host/zk_harness) must produce the same container, which covers the C++ shim at full size as well."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "ref_harness")
TWIN = os.path.join(ROOT, "zkdl_b200", "host", "zk_harness")

from oracle import oracle as orc
from oracle import refio

G1_KEYS = ("open.com_eval", "open.proof", "open.com_rows", "msm.full", "msm.small")
FR_P_LIMBS = np.array([(orc.FR_P >> (32 * i)) & 0xffffffff for i in range(8)], dtype=np.uint32)


@pytest.fixture(scope="module")
def zk():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from zkdl_b200 import capi
    capi.lib()
    return capi


def eq(a, b):
    return np.array_equal(np.asarray(a, dtype=np.uint32).reshape(-1), np.asarray(b, dtype=np.uint32).reshape(-1))


def mt_raw(seed, n):
    bg = np.random.MT19937()
    bg._legacy_seeding(seed)                      # init_genrand(seed) == std::mt19937(seed)
    return bg.random_raw(n).astype(np.uint32)


def signed_small(n, bits, seed):
    """ref_harness.cu rand_signed: v = (mt() & (2^bits - 1)) - 2^(bits-1)."""
    return (mt_raw(seed, n) & np.uint32((1 << bits) - 1)).astype(np.int64) - (1 << (bits - 1))


def fr_plain_from_signed(v):
    """numpy int64 (|v| < 2^32) -> plain (non-Montgomery) Fr limbs, negatives as p - |v|."""
    out = np.zeros((len(v), 8), dtype=np.uint32)
    a = np.abs(v).astype(np.uint64)
    neg = v < 0
    out[~neg, 0] = a[~neg].astype(np.uint32)
    an = a[neg]
    lo = (np.uint64(FR_P_LIMBS[0]) + (np.uint64(1) << np.uint64(32)) - an) & np.uint64(0xffffffff)     # p0 = 1
    borrow = (an > np.uint64(FR_P_LIMBS[0])).astype(np.uint32)
    rows = np.tile(FR_P_LIMBS, (len(an), 1))
    rows[:, 0] = lo.astype(np.uint32)
    rows[:, 1] = FR_P_LIMBS[1] - borrow
    out[neg] = rows
    return out


def dev_signed(zk, n, bits, seed, mont=True):
    t = zk.to_device(fr_plain_from_signed(signed_small(n, bits, seed)))
    return zk.fr_elementwise(zk.OP_MONT, t, out=t) if mont else t


def seeded_points(zk, n, seed):
    from zkdl_b200 import mlp
    return zk.g1_mul(zk.to_device(mlp._generator()), zk.to_device(zk.random_vec(seed, n)))


def run_harness(binary, tmp_path, name, *args):
    import torch
    if not os.path.exists(binary):
        pytest.skip(f"{binary} not present")
    out = str(tmp_path / f"{name}.bin")
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(os.path.dirname(torch.__file__), "lib") + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    r = subprocess.run([binary, "full", *[str(a) for a in args], out], capture_output=True, text=True, timeout=1500, env=env)
    assert r.returncode == 0, (r.stdout[-300:], r.stderr[-300:])
    box = refio.read_box(out)
    assert box["cuda_status"][0] == 0
    return box


def compare_boxes(ref, got):
    assert set(ref) == set(got)
    for k in ref:
        if k in G1_KEYS:
            assert orc.g1_eq(ref[k].reshape(-1, 36), got[k].reshape(-1, 36)).all(), k
        else:
            assert eq(ref[k], got[k]), k


def check_fc(zk, ref, I, O, B, wbits, with_open):
    import torch
    kb, ki, ko = (orc.ceil_log2(v) for v in (B, I, O))
    dW, dX = dev_signed(zk, I * O, wbits, 1), dev_signed(zk, B * I, 17, 2)
    mm = zk.MatmulWeights(dW, I, O)
    Z = zk.fr_matmul_prepared(dX, mm, B)
    u_bs, u_in, u_out = zk.random_vec(101, kb), zk.random_vec(102, ki), zk.random_vec(103, ko)
    Xr, Wr = zk.fr_partial_me(dX, u_bs, I), zk.fr_partial_me(dW, u_out, 1)
    assert eq(zk.to_host(Xr), ref["fc.xr"]) and eq(zk.to_host(Wr), ref["fc.wr"])
    assert eq(zk.to_host(zk.ip_sumcheck(Xr, Wr, u_in)), ref["fc.ip"])
    assert eq(zk.to_host(zk.fr_me(Z, np.concatenate([u_out, u_bs]))), ref["fc.zu"])
    assert eq(zk.to_host(zk.fr_sum(Z)), ref["fc.z_sum"])
    if not with_open:
        mm.close()
        return
    ng = 1 << ((orc.ceil_log2(I * O) + 1) // 2)
    ncom = I * O // ng
    G, com = seeded_points(zk, ng, 7), seeded_points(zk, ncom, 9)
    gens, com_tab = zk.G1Table(G, full=True), zk.G1Table(com, full=True)
    u = np.concatenate([u_out, u_in])
    k = orc.ceil_log2(ncom)
    u_hi = u[len(u) - k:]
    assert eq(zk.to_host(zk.fr_partial_me(dW, u_hi, ng)), ref["open.tf"])
    assert orc.g1_eq(zk.to_host(zk.commit(gens, dW[: 2 * ng])), ref["open.com_rows"].reshape(-1, 36)).all()
    nip = 3 * ki + 2
    for w_int in (None, mm):                          # Fr-table folds and the integer weight folds bench.py uses
        pfr, pg1 = zk.zkfc_prove(dX, dW, Z, B, I, O, gens, com_tab, u_bs, u_in, u_out, w_int=w_int)
        pfr, pg1 = zk.to_host(pfr), zk.to_host(pg1)
        assert eq(pfr[:nip], ref["fc.ip"]) and eq(pfr[nip], ref["fc.zu"]) and eq(pfr[nip + 1], ref["open.ret"])
        assert orc.g1_eq(pg1[:1], ref["open.com_eval"].reshape(-1, 36)).all()
        assert orc.g1_eq(pg1[1:], ref["open.proof"].reshape(-1, 36)).all()
    # the pieces the multi-GPU plan proves separately give the same segments
    pfr_s, _ = zk.zkfc_prove(dX, dW, Z, B, I, O, gens, com_tab, u_bs, u_in, u_out, parts=zk.FC_SUMCHECK, w_int=mm)
    pfr_o, pg1_o = zk.zkfc_prove(dX, dW, Z, B, I, O, gens, com_tab, u_bs, u_in, u_out, parts=zk.FC_OPENING, w_int=mm)
    assert eq(zk.to_host(pfr_s)[: nip + 1], pfr[: nip + 1]) and eq(zk.to_host(pfr_o)[nip + 1], pfr[nip + 1])
    assert orc.g1_eq(zk.to_host(pg1_o), pg1).all()
    torch.cuda.synchronize()
    mm.close(); gens.close(); com_tab.close()


def check_relu(zk, ref, I, O, B, wbits, generic):
    dW, dX = dev_signed(zk, I * O, wbits, 1), dev_signed(zk, B * I, 17, 2)
    mm = zk.MatmulWeights(dW, I, O)
    Z = zk.fr_matmul_prepared(dX, mm, B)
    mm.close()
    del dW, dX
    L = orc.ceil_log2(B * O)
    A, sign, magp, remp, bad = zk.relu_packed(Z)
    assert int(bad.item()) == 0
    assert eq(zk.to_host(zk.fr_me(A, zk.random_vec(201, L))), ref["relu.a_me"])
    assert eq(zk.to_host(zk.fr_me(sign, zk.random_vec(202, L))), ref["relu.sign_me"])
    mag, rem = zk.relu_expand(magp, remp)             # the reference's 0/1 Fr tables (zkrelu.cu:30-38)
    assert eq(zk.to_host(zk.fr_me(mag, zk.random_vec(203, L + 5))), ref["relu.mag_me"])
    assert eq(zk.to_host(zk.fr_me(rem, zk.random_vec(204, L + 4))), ref["relu.rem_me"])
    ch = [zk.random_vec(211, L + 5), zk.random_vec(212, L + 5), zk.random_vec(213, L + 4), zk.random_vec(214, L + 4),
          zk.random_vec(215, L), zk.random_vec(216, L), zk.random_vec(217, L)]
    exp = np.concatenate([ref[k] for k in ("relu.mag_sc", "relu.mag_rec", "relu.rem_sc", "relu.rem_rec", "relu.hp")])
    assert eq(zk.to_host(zk.zkrelu_prove_packed(Z, sign, magp, remp, *ch)), exp)          # the route bench.py times
    if generic:                                       # the generic sumchecks on the Fr tables (zkdl_zkrelu_prove)
        assert eq(zk.to_host(zk.zkrelu_prove(Z, sign, mag, rem, *ch)), exp)


@pytest.mark.parametrize("k,B,generic", [(11, 256, True), (10, 4096, False)], ids=["demo-hidden-layer-2048x2048-B256", "deep-narrow-1024x1024-B4096"])
def test_layer_against_reference(zk, tmp_path, k, B, generic):
    """One whole layer at BASELINE size: config 4's hidden layer (|G| = 2048, n = 2^19) and config 5's (B = 4096, n = 2^22)."""
    ref = run_harness(REF, tmp_path, "ref", "layer", k, B)
    check_fc(zk, ref, 1 << k, 1 << k, B, 13, True)
    check_relu(zk, ref, 1 << k, 1 << k, B, 13, generic)
    compare_boxes(ref, run_harness(TWIN, tmp_path, "twin", "layer", k, B))


def test_fc4096_sumcheck_set_against_reference(zk, tmp_path):
    """Config 2: 4096x4096 16-bit weights, batch 256: X.partial_me, W.partial_me, inner-product sumcheck, Z(u)."""
    ref = run_harness(REF, tmp_path, "ref", "fc", 12, 256)
    check_fc(zk, ref, 4096, 4096, 256, 16, False)
    compare_boxes(ref, run_harness(TWIN, tmp_path, "twin", "fc", 12, 256))


@pytest.mark.parametrize("k", [18, 20])
def test_msm_against_reference(zk, tmp_path, k):
    """Config 3: (G * s).sum() with 255-bit scalars and with 16-bit signed scalars, plain and fixed-base Pippenger."""
    ref = run_harness(REF, tmp_path, "ref", "msm", k, 0)
    n = 1 << k
    G = seeded_points(zk, n, 7)
    s = zk.to_device(zk.random_vec(8, n))
    w = dev_signed(zk, n, 16, 3, mont=False)
    tab = zk.G1Table(G, full=False)
    assert orc.g1_eq(zk.to_host(zk.msm(tab, s, 1, False)), ref["msm.full"].reshape(-1, 36)).all()
    assert orc.g1_eq(zk.to_host(zk.msm(tab, w, 1, False)), ref["msm.small"].reshape(-1, 36)).all()
    tab.close()
    if k <= 18:
        tab = zk.G1Table(G, full=True)
        assert orc.g1_eq(zk.to_host(zk.msm(tab, s, 1, False)), ref["msm.full"].reshape(-1, 36)).all()
        tab.close()
    compare_boxes(ref, run_harness(TWIN, tmp_path, "twin", "msm", k, 0))


def test_split_against_reference(zk, tmp_path):
    """FrTensor::split (fr-tensor.cu:376-397) of the drop-in shim (two strided 2-D copies + ragged tail) on a ragged table,
    windows 64 / 1 / 7 / n-1, against the reference's kernel."""
    compare_boxes(run_harness(REF, tmp_path, "ref", "split", 12, 64), run_harness(TWIN, tmp_path, "twin", "split", 12, 64))


def test_random_generator_against_reference(zk, tmp_path):
    """FrTensor::random / random_int (fr-tensor.cu:302-347): the reference's kernels launched with a fixed seed against
    zkdl_fr_random / zkdl_fr_random_int (curand XORWOW, curand_init(seed, index, 0)), 5000 elements (ragged last CTA)."""
    compare_boxes(run_harness(REF, tmp_path, "ref", "random", 5000, 123456789), run_harness(TWIN, tmp_path, "twin", "random", 5000, 123456789))
