"""Full-size, size-independent parity properties (the CPU oracle only finishes small cases in seconds): the whole
18.2 M-parameter batch-256 demo proof is checked by the verifier identities of SURVEY.md §4 (zkdl_b200/verify.py) —
sumcheck round consistency, matmul claim vs Z(u), the opening recursion against com(u_hi) and the folded generator —
plus linearity / additive-split properties of the MSM engine at sizes the oracle cannot reach."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from zkdl_b200 import capi as zk, mlp, verify
    zk.lib()
    ws, x = mlp.synthetic_mlp(mlp.demo_layer_dims(), 256, seed=0)
    P = mlp.MLPProver(ws, gen_seed=3)
    P.forward(x)
    return zk, mlp, verify, P


def test_demo_proof_verifies_at_full_size(env):
    zk, mlp, verify, P = env
    proof = P.prove(seed=77)
    assert len(proof) == 15
    for part, (kind, i, ch, _mask) in zip(proof, P.last_tasks):
        L = P.layers[i]
        if kind == "fc":
            info = verify.verify_zkfc(part[2], part[3], L.G, P.B, L.I, L.O, *ch, gens_table=None)
            assert 0 <= info["z_eval"] < verify.P
        else:
            u_z, v_z, u_r, v_r, u_rec, u_hp, v_hp = ch
            assert verify.verify_zkrelu(part[2], P.B * L.O, u_z, v_z, u_r, v_r, u_hp, v_hp)


def test_tampered_proofs_are_rejected(env):
    zk, mlp, verify, P = env
    proof = P.prove(seed=78, fc_layers=[2], relu_layers=[2])
    (kr, ir, chr_, _m1), (kf, if_, chf, _m2) = P.last_tasks
    relu_part = next(p for p in proof if p[0] == "relu"); fc_part = next(p for p in proof if p[0] == "fc")
    L = P.layers[2]
    verify.verify_zkfc(fc_part[2], fc_part[3], L.G, P.B, L.I, L.O, *chf)
    for idx in (0, 7, 3 * 11 + 1, 3 * 11 + 2, 3 * 11 + 3):              # a round coefficient, a final value, Z(u), open_ret
        bad = fc_part[2].clone(); bad[idx, 0] ^= 1
        with pytest.raises(verify.VerifyError):
            verify.verify_zkfc(bad, fc_part[3], L.G, P.B, L.I, L.O, *chf)
    for idx in (0, 1, 5, 34):                                           # com(u_hi), T of round 0, T1 of round 1, G_final
        bad = fc_part[3].clone(); bad[idx] = fc_part[3][(idx + 2) % 35]
        with pytest.raises(verify.VerifyError):
            verify.verify_zkfc(fc_part[2], bad, L.G, P.B, L.I, L.O, *chf)
    u_z, v_z, u_r, v_r, u_rec, u_hp, v_hp = chr_
    n = P.B * L.O
    verify.verify_zkrelu(relu_part[2], n, u_z, v_z, u_r, v_r, u_hp, v_hp)
    for idx in (1, 30, 3 * 24 + 1 + 32 + 4, relu_part[2].shape[0] - 1):
        bad = relu_part[2].clone(); bad[idx, 1] ^= 4
        with pytest.raises(verify.VerifyError):
            verify.verify_zkrelu(bad, n, u_z, v_z, u_r, v_r, u_hp, v_hp)


def test_msm_linearity_and_additive_split_large(env):
    """MSM(s + t) = MSM(s) + MSM(t); MSM over [0,n) = MSM over [0,n/2) + MSM over [n/2,n)  (n = 2^16, 255-bit scalars)."""
    zk, mlp, verify, P = env
    n = 1 << 16
    G = zk.g1_mul(zk.to_device(mlp._generator()), zk.to_device(zk.random_vec(31, n)))
    s, t = zk.to_device(zk.random_vec(32, n)), zk.to_device(zk.random_vec(33, n))
    tab = zk.G1Table(G, full=False)
    ms, mt = zk.msm(tab, s, 1, False), zk.msm(tab, t, 1, False)
    mst = zk.msm(tab, zk.fr_elementwise(zk.OP_ADD, s, t), 1, False)
    assert verify._same_point(mst, zk.g1_elementwise(zk.G1_ADD, ms, mt))
    lo, hi = zk.G1Table(G[: n // 2].contiguous(), full=False), zk.G1Table(G[n // 2:].contiguous(), full=False)
    parts = zk.g1_elementwise(zk.G1_ADD, zk.msm(lo, s[: n // 2].contiguous(), 1, False), zk.msm(hi, s[n // 2:].contiguous(), 1, False))
    assert verify._same_point(ms, parts)
    # fixed-base (window tables) and plain Pippenger agree
    full = zk.G1Table(G[:4096].contiguous(), full=True); plain = zk.G1Table(G[:4096].contiguous(), full=False)
    assert verify._same_point(zk.msm(full, s[:4096].contiguous(), 1, False), zk.msm(plain, s[:4096].contiguous(), 1, False))
    for tb in (tab, lo, hi, full, plain):
        tb.close()
