"""prove -> file -> verify (zkdl_b200/proof_file.py): a proof written in the wire format verifies from the file alone;
corrupted files are rejected either by the parser or by the verifier identities."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def proof_path(tmp_path_factory):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from zkdl_b200 import capi as zk, mlp, proof_file
    zk.lib()
    ws, x = mlp.synthetic_mlp([(20, 32), (32, 64), (64, 30), (30, 16)], 8, seed=4)
    P = mlp.MLPProver(ws, gen_seed=6)
    P.forward(x)
    proof = P.prove(seed=21)
    path = tmp_path_factory.mktemp("zkp") / "proof.zkp"
    n = proof_file.export(P, proof, str(path))
    assert n == path.stat().st_size
    return path


def test_file_verifies(proof_path):
    from zkdl_b200 import proof_file
    s = proof_file.verify_file(str(proof_path))
    assert [k for k, _, _ in s] == ["fc", "relu", "fc", "relu", "fc", "relu", "fc"]
    assert [i for _, i, _ in s] == [3, 2, 2, 1, 1, 0, 0]


def test_corrupted_files_are_rejected(proof_path, tmp_path):
    from zkdl_b200 import proof_file, serialize, verify
    blob = proof_path.read_bytes()
    public, tasks = serialize.loads(blob)
    bad = tmp_path / "bad.zkp"

    def rejected(pub, tks):
        bad.write_bytes(serialize.dumps(pub, tks))
        with pytest.raises((verify.VerifyError, ValueError)):
            proof_file.verify_file(str(bad))

    import copy
    t = copy.deepcopy(tasks); t[0]["fr"][2, 0] ^= 1; rejected(public, t)                     # a sumcheck coefficient
    t = copy.deepcopy(tasks); t[0]["g1"][3] = tasks[0]["g1"][4]; rejected(public, t)         # an opening point
    t = copy.deepcopy(tasks); t[2]["challenges"][1][0, 0] ^= 1; rejected(public, t)          # a challenge
    t = copy.deepcopy(tasks); t[1]["fr"][-1, 0] ^= 1; rejected(public, t)                    # final value of the Hadamard sumcheck
    p = copy.deepcopy(public); p["layers"][3]["commitment"][0] = public["layers"][3]["commitment"][1]; rejected(p, tasks)
    p = copy.deepcopy(public); p["layers"][2]["generators"][5] = public["layers"][2]["generators"][6]; rejected(p, tasks)
    # structure: a file with a layer proof missing, duplicated or reordered, or with challenge vectors whose
    # lengths do not follow from the public shapes, must not pass
    rejected(public, tasks[:-1])
    rejected(public, tasks[1:])
    rejected(public, tasks + [tasks[-1]])
    rejected(public, [tasks[1], tasks[0]] + tasks[2:])
    rejected(public, [])
    t = copy.deepcopy(tasks); t[0]["challenges"][1] = t[0]["challenges"][1][:-1]; rejected(public, t)
    t = copy.deepcopy(tasks); t[1]["challenges"][4] = np.concatenate([t[1]["challenges"][4], t[1]["challenges"][4][:1]]); rejected(public, t)
    t = copy.deepcopy(tasks); t[1]["fr"] = t[1]["fr"][:-1]; rejected(public, t)
    bad.write_bytes(blob[: len(blob) // 2])
    with pytest.raises(ValueError):
        proof_file.verify_file(str(bad))


def test_subgroup_check_rejects_cofactor_points(proof_path):
    """on-curve is not enough, BLS12-381 G1 has a cofactor.  The first abscissa with a square right-hand side gives
    a curve point outside the prime-order subgroup (probability 1 - 1/h); real generators pass."""
    from zkdl_b200 import capi as zk, serialize, verify
    public, _ = serialize.loads(proof_path.read_bytes())
    verify.verify_subgroup(zk.to_device(public["layers"][0]["generators"]))
    x = 1
    while True:
        b = bytearray(x.to_bytes(48, "big")); b[0] |= 0x80
        try:
            pt = serialize.g1_decompress(bytes(b)); break
        except ValueError:
            x += 1
    with pytest.raises(verify.VerifyError):
        verify.verify_subgroup(zk.to_device(pt))
