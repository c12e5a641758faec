"""Host side of the Fiat-Shamir transcript (zkdl_b200/fiat_shamir.py): the hashing rules the device kernel mirrors."""
import hashlib

import numpy as np

from zkdl_b200 import fiat_shamir as fs


def test_round_is_sha256_of_state_and_the_three_limb_images():
    T = fs.Transcript(b"\x00" * 32, "fc", 3)
    s0 = T.s
    c = np.arange(24, dtype=np.uint32).reshape(3, 8) * np.uint32(0x01020304)
    x = T.round(c[0], c[1], c[2])
    d = hashlib.sha256(s0 + c.astype("<u4").tobytes()).digest()
    assert T.s == d
    want = np.frombuffer(d, dtype="<u4").copy(); want[7] %= 1944954707
    assert np.array_equal(x, want) and x[7] < 1944954707


def test_vectors_are_deterministic_domain_separated_and_below_p():
    a, b = fs.Transcript(b"r" * 32, "fc", 1), fs.Transcript(b"r" * 32, "fc", 1)
    va, vb = a.vector(5), b.vector(5)
    assert np.array_equal(va, vb) and a.s == b.s and va.shape == (5, 8) and (va[:, 7] < 1944954707).all()
    assert not np.array_equal(va, fs.Transcript(b"r" * 32, "relu", 1).vector(5))
    assert not np.array_equal(va, fs.Transcript(b"r" * 32, "fc", 2).vector(5))
    assert not np.array_equal(va[0], va[1])
    assert fs.Transcript(b"r" * 32, "fc", 1).vector(0).shape == (0, 8)
    P = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
    for row in va:
        assert sum(int(x) << (32 * i) for i, x in enumerate(row)) < P


def test_replay_matches_round_by_round():
    proof = (np.arange(7 * 8, dtype=np.uint32).reshape(7, 8) + 5)
    a, b = fs.Transcript(b"q" * 32, "relu", 0), fs.Transcript(b"q" * 32, "relu", 0)
    xs = a.rounds(proof, 2)
    assert np.array_equal(xs[0], b.round(proof[0], proof[1], proof[2])) and np.array_equal(xs[1], b.round(proof[3], proof[4], proof[5]))
