"""Fiat-Shamir mode (zkdl_b200/fiat_shamir.py, csrc/fs_kernels.cu): the transcript-driven proofs
(a) verify with challenges the VERIFIER recomputes, (b) are bit-identical to what the injected-challenge kernels (the ones
pinned against the oracle / reference) produce for the same challenges - which also pins the device SHA-256 against
hashlib - and (c) stop verifying when any proof element or any public value is tampered with."""
import copy

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import oracle as orc


@pytest.fixture(scope="module")
def setup():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from zkdl_b200 import capi as zk, mlp, fiat_shamir as fs
    zk.lib()
    ws, x = mlp.synthetic_mlp([(20, 32), (32, 64), (64, 30), (30, 16)], 8, seed=4)
    P = mlp.MLPProver(ws, gen_seed=6)
    P.forward(x)
    public, proofs = fs.prove(P)
    return zk, fs, P, public, proofs


def eq(a, b):
    return np.array_equal(np.asarray(a, dtype=np.uint32).reshape(-1), np.asarray(b, dtype=np.uint32).reshape(-1))


def test_fs_proof_verifies(setup):
    zk, fs, P, public, proofs = setup
    assert [(p[0], p[1]) for p in proofs] == [("fc", 3), ("relu", 2), ("fc", 2), ("relu", 1), ("fc", 1), ("relu", 0), ("fc", 0)]
    assert fs.verify_all(public, P.B, proofs)


def test_fs_proofs_equal_injected_mode_for_the_transcript_challenges(setup):
    zk, fs, P, public, proofs = setup
    root = fs.public_root(public, P.B)
    for p in proofs:
        L = P.layers[p[1]]
        if p[0] == "fc":
            u_bs, u_in, u_out = fs.challenges_fc(root, p[1], P.B, L.I, L.O, zk.to_host(p[2]))
            X = P.A[p[1] - 1] if p[1] > 0 else P.X
            pfr, pg1 = zk.zkfc_prove(X, L.W, P.Z[p[1]], P.B, L.I, L.O, L.gens, L.com_table, u_bs, u_in, u_out)
            assert eq(zk.to_host(pfr), zk.to_host(p[2]))
            assert orc.g1_eq(zk.to_host(pg1), zk.to_host(p[3])).all()
        else:
            ch = fs.challenges_relu(root, p[1], P.B * L.O, zk.to_host(p[2]))
            sign, magp, remp = P.aux[p[1]]
            ref = zk.zkrelu_prove_packed(P.Z[p[1]], sign, magp, remp, *ch)
            assert eq(zk.to_host(ref), zk.to_host(p[2]))


def test_fs_tampering_is_rejected(setup):
    zk, fs, P, public, proofs = setup
    from zkdl_b200 import verify

    def rejected(pub, prs):
        with pytest.raises(verify.VerifyError):
            fs.verify_all(pub, P.B, prs)

    def with_fr(idx, row, limb=0):
        prs = list(proofs)
        t = prs[idx][2].clone(); t[row, limb] ^= 1
        prs[idx] = prs[idx][:2] + (t,) + prs[idx][3:]
        return prs

    rejected(public, with_fr(0, 1))                      # a coefficient of the matmul sumcheck: every later challenge changes
    nip = 3 * 5 + 2                                      # layer 3: I = 32
    rejected(public, with_fr(0, nip))                    # the claimed Z(u)
    rejected(public, with_fr(1, 4))                      # a binary-sumcheck coefficient
    rejected(public, with_fr(1, proofs[1][2].shape[0] - 1))     # final value of the Hadamard sumcheck
    pub = copy.deepcopy(public); pub[2]["commitment"][0] = public[2]["commitment"][1]
    rejected(pub, proofs)                                # the root, hence every challenge, depends on the public commitments
    rejected(public, proofs[:-1])


def test_fs_proof_file_roundtrip(setup, tmp_path):
    """prove --fiat-shamir -> file -> verify from the file alone: the file holds no challenges (version 2), the verifier
    re-derives them; a flipped proof limb or a swapped public commitment is rejected."""
    zk, fs, P, public, proofs = setup
    from zkdl_b200 import proof_file, serialize, verify
    path = tmp_path / "fs.zkp"
    n = proof_file.export_fs(P, public, proofs, str(path))
    assert n == path.stat().st_size
    s = proof_file.verify_file(str(path))
    assert [(k, i) for k, i, _ in s] == [(p[0], p[1]) for p in proofs]
    pub, tasks = serialize.loads(path.read_bytes())
    assert pub["fiat_shamir"] and all(t["challenges"] == [] for t in tasks)
    bad = tmp_path / "bad.zkp"
    t = copy.deepcopy(tasks); t[0]["fr"][2, 0] ^= 1
    bad.write_bytes(serialize.dumps(pub, t, fiat_shamir=True))
    with pytest.raises(verify.VerifyError):
        proof_file.verify_file(str(bad))
    p2 = copy.deepcopy(pub); p2["layers"][1]["commitment"][0] = pub["layers"][1]["commitment"][1]
    bad.write_bytes(serialize.dumps(p2, tasks, fiat_shamir=True))
    with pytest.raises(verify.VerifyError):
        proof_file.verify_file(str(bad))
