"""Linked mode on the GPU (zkdl_b200/linked.py over the C ABI): the chain verifies, tampering is rejected, the file round
trip verifies, and the whole proof equals what the CPU oracle produces from the same tables and the same transcript
(Fr bit for bit, G1 as points)."""
import copy

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def chain():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from zkdl_b200 import capi as zk, mlp, linked
    zk.lib()
    ws, x = mlp.synthetic_mlp([(12, 16), (16, 8), (8, 5)], 4, seed=9)
    P = mlp.MLPProver(ws, gen_seed=3)
    P.forward(x)
    P.check_range()
    public, proof = linked.prove(P, check=True)
    return P, public, proof


def test_chain_verifies_and_round_trips(chain, tmp_path):
    from zkdl_b200 import linked, proof_file
    P, public, proof = chain
    assert linked.verify_linked(public, proof)
    path = str(tmp_path / "chain.zkp")
    assert linked.export(public, proof, path) > 0
    assert linked.verify_file(path)
    assert proof_file.main(["verify", path]) == 0


def test_tampered_chain_is_rejected(chain):
    from zkdl_b200 import linked, verify
    P, public, proof = chain

    def tampered(edit):
        bad = copy.deepcopy(proof)
        edit(bad)
        with pytest.raises(verify.VerifyError):
            linked.verify_linked(public, bad)

    def bump(arr, row=0):
        arr[row, 0] ^= 1

    tampered(lambda p: bump(p["output"], 2))
    tampered(lambda p: bump(p["input"], 5))
    tampered(lambda p: bump(p["steps"][2]["ip"], 4))
    tampered(lambda p: bump(p["steps"][1]["r_mag"], 0))
    tampered(lambda p: bump(p["steps"][3]["hp"], -2))                      # M~(q)
    tampered(lambda p: bump(p["steps"][3]["bin_rem"], -1))
    tampered(lambda p: bump(p["steps"][1]["opens"][1]["ret"]))
    tampered(lambda p: p["steps"][3]["opens"].reverse())
    tampered(lambda p: p["aux_com"][0].reverse())                           # sign / rem commitments swapped


def test_chain_equals_the_oracle(chain, monkeypatch):
    """The same host logic on the CPU oracle (tests/zk_cpu_mock.py) with the device's tables: identical roots, hence identical
    challenges, hence every proof element must agree."""
    import zk_cpu_mock as mock
    from oracle import oracle as orc
    from zkdl_b200 import capi as zk, fiat_shamir, linked, verify
    P, public, proof = chain
    H = mock.HostCopy(P, zk)
    for mod in (fiat_shamir, linked, verify):
        monkeypatch.setattr(mod, "zk", mock)
    pub_o, proof_o = linked.prove(H, check=True)
    def canon(pts):                                                        # infinity has many limb images (z = 0)
        a = np.array(pts, dtype=np.uint32).reshape(-1, 36)
        a[(a[:, 24:] == 0).all(axis=1)] = 0
        return a

    for a, b in zip(public, pub_o):
        assert np.array_equal(canon(a["generators"]), canon(b["generators"])) and np.array_equal(canon(a["commitment"]), canon(b["commitment"]))
    for ca, cb in zip(proof["aux_com"], proof_o["aux_com"]):
        for x, y in zip(ca, cb):
            assert np.array_equal(x, y), "auxiliary commitments differ from the oracle's"
    assert len(proof["steps"]) == len(proof_o["steps"])
    for s, t in zip(proof["steps"], proof_o["steps"]):
        opens = [s["open_w"]] if s["kind"] == "fc" else s["opens"]
        opens_o = [t["open_w"]] if t["kind"] == "fc" else t["opens"]
        for k in [k for k in s if k not in ("kind", "layer", "open_w", "opens")]:
            assert np.array_equal(s[k], t[k]), f"{s['kind']} {s['layer']}: {k} differs from the oracle's"
        for o, oo in zip(opens, opens_o):
            assert np.array_equal(o["ret"], oo["ret"])
            assert orc.g1_eq(o["g1"], oo["g1"]).all(), f"{s['kind']} {s['layer']}: opening points differ from the oracle's"
