"""Linked mode (zkdl_b200/linked.py) on toy shapes with the CPU oracle standing in for the C-ABI library (tests/zk_cpu_mock.py):
the chain verifies, every link is the table's own evaluation, and any tampered element is rejected."""
import copy

import numpy as np
import pytest

import zk_cpu_mock as mock
from zkdl_b200 import fiat_shamir, linked, verify


@pytest.fixture()
def cpu_zk(monkeypatch):
    for mod in (fiat_shamir, linked, verify):
        monkeypatch.setattr(mod, "zk", mock)
    return mock


@pytest.fixture(scope="module")
def toy():
    return mock.ToyProver([(6, 8), (8, 4), (4, 3)], batch=3, seed=5)


def test_linked_chain_verifies_and_rejects_tampering(cpu_zk, toy, tmp_path):
    public, proof = linked.prove(toy, check=True)
    assert linked.verify_linked(public, proof)
    path = str(tmp_path / "chain.zkp")
    size = linked.export(public, proof, path)                              # wire format version 3 and back
    assert size > 0 and linked.verify_file(path)
    blob = bytearray(open(path, "rb").read())
    blob[-40] ^= 1                                                         # inside the last opening's points
    open(path, "wb").write(bytes(blob))
    with pytest.raises((verify.VerifyError, ValueError)):
        linked.verify_file(path)
    kinds = [(s["kind"], s["layer"]) for s in proof["steps"]]
    assert kinds == [("fc", 2), ("relu", 1), ("fc", 1), ("relu", 0), ("fc", 0)]

    def tampered(edit):
        bad = copy.deepcopy(proof)
        edit(bad)
        with pytest.raises(verify.VerifyError):
            linked.verify_linked(public, bad)

    def bump(arr, row=0):
        arr[row, 0] ^= 1

    tampered(lambda p: bump(p["output"]))                                  # another output
    tampered(lambda p: bump(p["input"], 1))                                # another input
    tampered(lambda p: bump(p["steps"][0]["ip"], 1))                       # a sumcheck coefficient
    tampered(lambda p: bump(p["steps"][1]["r_mag"], 3))                    # a recover row
    tampered(lambda p: bump(p["steps"][1]["r_rem"], 15))
    tampered(lambda p: bump(p["steps"][1]["hp"], -1))                      # sign~(q)
    tampered(lambda p: bump(p["steps"][3]["bin_sign"], -1))
    tampered(lambda p: bump(p["steps"][3]["opens"][4]["ret"]))             # an opened value
    tampered(lambda p: p["steps"].pop())                                   # a missing link
    tampered(lambda p: p["aux_com"][0].reverse())                          # sign / rem commitments swapped
    # the public model is part of the transcript root: another generator or weight commitment changes every challenge
    for key, row in (("generators", 1), ("commitment", 0)):
        pub = copy.deepcopy(public)
        pub[1][key][row] = pub[1][key][row + 1]
        with pytest.raises(verify.VerifyError):
            linked.verify_linked(pub, proof)


def test_linked_rejects_a_wrong_activation(cpu_zk):
    """A prover whose ReLU output is inconsistent with its own decomposition (A != M o sign) cannot close the chain."""
    P = mock.ToyProver([(6, 8), (8, 4)], batch=2, seed=7)
    a = mock.to_host(P.A[0]).copy()
    a[1, 0] ^= 2
    P.A[0] = mock.to_device(a)
    z = mock.orc.fr_matmul(a, mock.to_host(P.layers[1].W), P.B, P.layers[1].I, P.layers[1].O)
    P.Z[1] = mock.to_device(z)                                             # a consistent last layer on top of the wrong A
    public, proof = linked.prove(P)
    with pytest.raises(verify.VerifyError):
        linked.verify_linked(public, proof)


def test_linked_batch_one_pads_short_tables(cpu_zk):
    """Batch 1: no batch variables, and the sign table (4 cells) is shorter than the generator set (8): committed zero-extended,
    opened at the zero-extended point."""
    P = mock.ToyProver([(16, 4), (4, 3)], batch=1, seed=3)
    assert P.B == 1 and P.layers[0].ngens == 8 and mock.to_host(P.aux[0][0]).shape[0] == 4
    public, proof = linked.prove(P, check=True)
    assert [c.shape[0] for c in proof["aux_com"][0]] == [1, 16, 8]
    assert linked.verify_linked(public, proof)


def test_linked_magnitude_range(cpu_zk):
    """Tiny pre-activations: -2^15 <= Z < 0 rounds the magnitude UP to exactly 2^31 (bit 31 set, low bits clear), which the
    range sumcheck b31 o low31 = 0 accepts; the other decomposition of a positive Z, (sign 0, M + 2^31), which would zero
    an activation, is rejected by it."""
    P = mock.ToyProver([(6, 8), (8, 4)], batch=2, seed=11, weight_scale=1e-4, input_scale=1e-3)
    magp = P.aux[0][1].numpy().view(np.uint32)
    assert (magp == 1 << 31).any(), "the toy case should contain a rounded-up tiny negative"
    public, proof = linked.prove(P, check=True)
    assert linked.verify_linked(public, proof)
    # the cheat: entry e is positive with a small magnitude; claim it negative with M + 2^31 and output 0
    P = mock.ToyProver([(6, 8), (8, 4)], batch=2, seed=12)
    sign, magp, remp = P.aux[0]
    sg = mock.to_host(sign).copy()
    m = magp.numpy().view(np.uint32).copy()
    e = int(np.flatnonzero(sg.any(axis=1) & (m > 0) & (m < (1 << 31)))[0])
    sg[e] = 0
    m[e] += 1 << 31
    a = mock.to_host(P.A[0]).copy()
    a[e] = 0
    import torch
    P.aux[0] = (mock.to_device(sg), torch.from_numpy(m.view(np.int32).copy()), remp)
    P.A[0] = mock.to_device(a)
    P.Z[1] = mock.to_device(mock.orc.fr_matmul(a, mock.to_host(P.layers[1].W), P.B, P.layers[1].I, P.layers[1].O))
    public, proof = linked.prove(P, check=True)                           # every link holds: the decomposition is consistent
    with pytest.raises(verify.VerifyError, match="sumcheck: round 0"):
        linked.verify_linked(public, proof)
