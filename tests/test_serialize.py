"""Wire format (zkdl_b200/serialize.py): Fr / G1 encodings against known vectors and the oracle's group law, file
round trips and malformed-input rejection.  CPU only; the prove -> file -> verify loop is in test_proof_file_gpu.py."""
import numpy as np
import pytest

from oracle import oracle as orc
from zkdl_b200 import serialize as ser

rng = np.random.default_rng(11)
GEN_COMPRESSED = bytes.fromhex("97f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb")


def normalise(points):
    """[n,36] Jacobian limbs (any z) -> z = 1 representatives, with Python integers (what zkdl_g1_normalize does on the GPU)."""
    P = ser.FQ_P
    out = np.zeros_like(points)
    for i, row in enumerate(points):
        x, y, z = (ser._ints(row[12 * k: 12 * k + 12].reshape(1, 12))[0] * ser.FQ_RINV % P for k in range(3))
        if z == 0:
            continue
        zi = pow(z, -1, P)
        xa, ya = x * zi * zi % P, y * zi * zi * zi % P
        out[i] = np.concatenate([ser._limbs([xa * ser.FQ_R % P], 12)[0], ser._limbs([ya * ser.FQ_R % P], 12)[0], ser._limbs([ser.FQ_R], 12)[0]])
    return out


def some_points(n):
    ks = orc.to_limbs([int.from_bytes(rng.bytes(31), "little") for _ in range(n)])
    return orc.g1_mul(orc.g1_generator(), ks, fast=True)


def test_generator_known_vector():
    g = normalise(orc.g1_generator())
    assert ser.g1_compress(g) == GEN_COMPRESSED
    back = ser.g1_decompress(GEN_COMPRESSED)
    assert np.array_equal(back, g)
    neg = normalise(orc.g1_neg(orc.g1_generator()))
    c = ser.g1_compress(neg)
    assert c[0] == GEN_COMPRESSED[0] ^ 0x20 and c[1:] == GEN_COMPRESSED[1:]          # same x, the other root
    assert np.array_equal(ser.g1_decompress(c), neg)


def test_g1_round_trips_and_infinity():
    pts = normalise(some_points(40))
    pts[7] = 0                                                                        # infinity (z = 0)
    for enc, dec, width in ((ser.g1_compress, ser.g1_decompress, 48), (ser.g1_uncompressed, ser.g1_from_uncompressed, 96)):
        blob = enc(pts)
        assert len(blob) == width * len(pts)
        back = dec(blob)
        assert np.array_equal(back, pts)
        assert orc.g1_on_curve(np.delete(back, 7, axis=0)).all()
    with pytest.raises(ValueError):
        ser.g1_compress(orc.g1_double(some_points(1)))                               # z != 1: must be normalised first
    bad = bytearray(ser.g1_compress(pts[:1])); bad[0] &= 0x7F
    with pytest.raises(ValueError):
        ser.g1_decompress(bytes(bad))
    with pytest.raises(ValueError):                                                   # x = 1 is not on y^2 = x^3 + 4 ... (5 is no square)
        ser.g1_decompress(bytes([0x80]) + bytes(46) + bytes([1]))
    with pytest.raises(ValueError):
        ser.g1_from_uncompressed(bytes(47) + bytes([1]) + bytes(47) + bytes([1]))


def test_fr_round_trip_and_canonical_check():
    vals = [0, 1, ser.FR_P - 1, ser.FR_R] + [int.from_bytes(rng.bytes(40), "little") % ser.FR_P for _ in range(50)]
    limbs = orc.to_limbs(vals)
    blob = ser.fr_to_bytes(limbs)
    assert np.array_equal(ser.fr_from_bytes(blob), limbs)
    plain = [v * ser.FR_RINV % ser.FR_P for v in vals]
    assert blob[:32] == plain[0].to_bytes(32, "little") and blob[64:96] == plain[2].to_bytes(32, "little")
    with pytest.raises(ValueError):
        ser.fr_from_bytes(ser.FR_P.to_bytes(32, "little"))
    with pytest.raises(ValueError):
        ser.fr_from_bytes(bytes(31))


def fake_proof():
    pts = normalise(some_points(12))
    layers = [{"in_dim": 3, "out_dim": 4, "I": 4, "O": 4, "generators": pts[:4], "commitment": pts[4:8]},
              {"in_dim": 4, "out_dim": 2, "I": 4, "O": 2, "generators": pts[8:10], "commitment": pts[10:12]}]
    fr = lambda n: orc.to_limbs([int.from_bytes(rng.bytes(40), "little") % ser.FR_P for _ in range(n)])
    tasks = [{"kind": "fc", "layer": 1, "challenges": [fr(0), fr(2), fr(1)], "fr": fr(10), "g1": pts[2:7]},
             {"kind": "relu", "layer": 0, "challenges": [fr(k) for k in (7, 7, 6, 6, 2, 2, 2)], "fr": fr(90), "g1": None}]
    return {"batch": 2, "layers": layers}, tasks


def test_file_round_trip_and_rejections():
    public, tasks = fake_proof()
    blob = ser.dumps(public, tasks)
    pub2, tasks2 = ser.loads(blob)
    assert pub2["batch"] == 2 and len(pub2["layers"]) == 2
    for a, b in zip(public["layers"], pub2["layers"]):
        assert all(a[k] == b[k] for k in ("in_dim", "out_dim", "I", "O"))
        assert np.array_equal(a["generators"], b["generators"]) and np.array_equal(a["commitment"], b["commitment"])
    for a, b in zip(tasks, tasks2):
        assert a["kind"] == b["kind"] and a["layer"] == b["layer"] and np.array_equal(a["fr"], b["fr"])
        assert all(np.array_equal(x, y) for x, y in zip(a["challenges"], b["challenges"]))
        assert (a["g1"] is None and b["g1"] is None) or np.array_equal(a["g1"], b["g1"])
    assert ser.dumps(pub2, tasks2) == blob
    for bad in (blob[:-1], blob + b"\0", b"XKDLPRF1" + blob[8:], blob[:8] + (2).to_bytes(4, "little") + blob[12:]):
        with pytest.raises(ValueError):
            ser.loads(bad)


def test_uncompressed_infinity_must_be_canonical():
    """the infinity flag with non-zero payload, or compression / sort flags on an uncompressed point, are rejected."""
    import pytest
    from zkdl_b200 import serialize as sz
    assert (sz.g1_from_uncompressed(bytes([0x40]) + bytes(95))[0, 24:] == 0).all()
    for bad in (bytes([0x40]) + bytes(94) + b"\x01", bytes([0x41]) + bytes(95), bytes([0xC0]) + bytes(95), bytes([0x20]) + bytes(95)):
        with pytest.raises(ValueError):
            sz.g1_from_uncompressed(bad)
