"""Wire format for the proofs of the FC-layer path.  The reference computes its proofs and drops
them (`zkFC::prove` / `zkReLU::prove` return void); this module writes what a verifier needs to a file and reads it back.

Encodings (the ones the BLS12-381 ecosystem uses, so other tooling can parse the elements):
  * Fr element   32 bytes, little-endian canonical integer (NOT Montgomery);
  * G1 point     48 bytes compressed, big-endian x with the three flag bits of the zcash/IETF serialisation in the top
                 byte: 0x80 compressed, 0x40 infinity, 0x20 y is the lexicographically larger root;
  * public G1    96 bytes uncompressed (big-endian x || y, 0x40 flag for infinity): the generators, so that loading
                 them costs no square roots.
The loader rebuilds exactly the limb arrays the C ABI works on (Montgomery limbs are canonical, so the map is a
bijection): a loaded proof verifies with zkdl_b200.verify unchanged.

File layout ("ZKDLPRF1"): header, the public part per layer (shapes, generators, weight commitment), then one record per
proved task in proving order (kind, layer, challenge vectors, Fr rows, G1 rows).  All integers little-endian u32.

Host-side big-integer code only (a few hundred elements per proof); nothing here touches the GPU."""
import struct

import numpy as np

FR_P = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
FQ_P = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
FR_R, FQ_R = (1 << 256) % FR_P, (1 << 384) % FQ_P
FR_RINV, FQ_RINV = pow(FR_R, -1, FR_P), pow(FQ_R, -1, FQ_P)
MAGIC = b"ZKDLPRF1"


def _ints(rows):
    rows = np.ascontiguousarray(rows, dtype="<u4")
    n = 4 * rows.shape[-1]
    b = rows.tobytes()                                                  # little-endian limbs, low limb first = the integer's bytes
    return [int.from_bytes(b[i:i + n], "little") for i in range(0, len(b), n)]


def _limbs(vals, w):
    if not len(vals):
        return np.zeros((0, w), dtype=np.uint32)
    return np.frombuffer(b"".join(int(v).to_bytes(4 * w, "little") for v in vals), dtype="<u4").reshape(-1, w).astype(np.uint32)


# ------------------------------------------------------------------------------------------------ Fr
def fr_to_bytes(limbs):
    """[n, 8] Montgomery limbs -> n * 32 bytes (little-endian canonical integers)."""
    return b"".join((v * FR_RINV % FR_P).to_bytes(32, "little") for v in _ints(limbs))


def fr_from_bytes(buf):
    if len(buf) % 32:
        raise ValueError("Fr block is not a multiple of 32 bytes")
    vals = [int.from_bytes(buf[i:i + 32], "little") for i in range(0, len(buf), 32)]
    if any(v >= FR_P for v in vals):
        raise ValueError("non-canonical Fr element")
    return _limbs([v * FR_R % FR_P for v in vals], 8)


# ------------------------------------------------------------------------------------------------ G1
def _affine_ints(points):
    """[n, 36] Jacobian limbs with z = 1 (zkdl_g1_normalize) or z = 0 -> [(x, y) plain integers or None]."""
    pts = np.ascontiguousarray(points, dtype=np.uint32).reshape(-1, 36)
    xs, ys, zs = _ints(pts[:, :12]), _ints(pts[:, 12:24]), _ints(pts[:, 24:])
    out = []
    for x, y, z in zip(xs, ys, zs):
        if z == 0:
            out.append(None)
            continue
        if z != FQ_R:
            raise ValueError("point is not normalised (z != 1): pass it through zkdl_g1_normalize first")
        out.append((x * FQ_RINV % FQ_P, y * FQ_RINV % FQ_P))
    return out


def _jacobian(aff):
    rows = np.zeros((len(aff), 36), dtype=np.uint32)
    for i, p in enumerate(aff):
        if p is None:
            continue                                                   # infinity: z = 0
        rows[i] = np.concatenate([_limbs([p[0] * FQ_R % FQ_P], 12)[0], _limbs([p[1] * FQ_R % FQ_P], 12)[0], _limbs([FQ_R], 12)[0]])
    return rows


def g1_compress(points):
    out = []
    for p in _affine_ints(points):
        if p is None:
            out.append(bytes([0xC0]) + bytes(47))
            continue
        x, y = p
        b = bytearray(x.to_bytes(48, "big"))
        b[0] |= 0x80 | (0x20 if y > FQ_P - y else 0)
        out.append(bytes(b))
    return b"".join(out)


def g1_decompress(buf):
    if len(buf) % 48:
        raise ValueError("G1 block is not a multiple of 48 bytes")
    aff = []
    for i in range(0, len(buf), 48):
        b = buf[i:i + 48]
        if not b[0] & 0x80:
            raise ValueError("compression flag missing")
        if b[0] & 0x40:
            if (b[0] & 0x3F) or any(b[1:]):
                raise ValueError("malformed point at infinity")
            aff.append(None)
            continue
        x = int.from_bytes(bytes([b[0] & 0x1F]) + b[1:], "big")
        if x >= FQ_P:
            raise ValueError("x coordinate out of range")
        y2 = (x * x * x + 4) % FQ_P
        y = pow(y2, (FQ_P + 1) // 4, FQ_P)                              # p = 3 mod 4
        if y * y % FQ_P != y2:
            raise ValueError("x is not the abscissa of a curve point")
        if bool(b[0] & 0x20) != (y > FQ_P - y):
            y = FQ_P - y
        aff.append((x, y))
    return _jacobian(aff)


def g1_uncompressed(points):
    out = []
    for p in _affine_ints(points):
        out.append(bytes([0x40]) + bytes(95) if p is None else p[0].to_bytes(48, "big") + p[1].to_bytes(48, "big"))
    return b"".join(out)


def g1_from_uncompressed(buf, check=True):
    if len(buf) % 96:
        raise ValueError("G1 block is not a multiple of 96 bytes")
    aff = []
    for i in range(0, len(buf), 96):
        b = buf[i:i + 96]
        if b[0] & 0xA0:
            raise ValueError("compression / sort flags set in an uncompressed point")
        if b[0] & 0x40:
            if (b[0] & 0x3F) or any(b[1:]):
                raise ValueError("malformed point at infinity")
            aff.append(None)
            continue
        x, y = int.from_bytes(b[:48], "big"), int.from_bytes(b[48:], "big")
        if x >= FQ_P or y >= FQ_P or (check and (y * y - x * x * x - 4) % FQ_P):
            raise ValueError("not a curve point")
        aff.append((x, y))
    return _jacobian(aff)


# ------------------------------------------------------------------------------------------------ file
def _u32(*v):
    return struct.pack("<%dI" % len(v), *v)


def dumps(public, tasks, fiat_shamir=False, linked=None):
    """public: [{in_dim, out_dim, I, O, generators [n,36] normalised, commitment [m,36] normalised}] per layer;
    tasks: [{kind: "fc"|"relu", layer, challenges: [[k,8] limbs ...], fr: [r,8] limbs, g1: [s,36] normalised or None}]
    in proving order; batch is stored in the header.  Version 1 = injected challenges (stored per task); version 2 =
    Fiat-Shamir mode (zkdl_b200/fiat_shamir.py): same layout, every task's challenge list is empty - the verifier derives them.
    Version 3 = linked mode (zkdl_b200/linked.py): version 2 plus, after the public part, the input and output tables and the
    three auxiliary row commitments (sign, mag_bin, rem_bin) of every zkReLU layer (uncompressed, like the generators);
    linked = {"input", "output", "aux_com"}."""
    out = [MAGIC, _u32(3 if linked is not None else 2 if fiat_shamir else 1, len(public["layers"]), public["batch"])]
    for L in public["layers"]:
        out.append(_u32(L["in_dim"], L["out_dim"], L["I"], L["O"], len(L["generators"]), len(L["commitment"])))
        out.append(g1_uncompressed(L["generators"]))
        out.append(g1_compress(L["commitment"]))
    if linked is not None:
        for key in ("input", "output"):
            out.append(_u32(len(linked[key]))); out.append(fr_to_bytes(linked[key]))
        out.append(_u32(len(linked["aux_com"])))
        for coms in linked["aux_com"]:
            for c in coms:
                out.append(_u32(len(c))); out.append(g1_uncompressed(c))    # ~10^5 points for the demo model: no square roots on load
    out.append(_u32(len(tasks)))
    for t in tasks:
        out.append(_u32(0 if t["kind"] == "fc" else 1, t["layer"], len(t["challenges"])))
        for c in t["challenges"]:
            c = np.asarray(c, dtype=np.uint32).reshape(-1, 8)
            out.append(_u32(len(c))); out.append(fr_to_bytes(c))
        out.append(_u32(len(t["fr"]))); out.append(fr_to_bytes(t["fr"]))
        g1 = t.get("g1")
        out.append(_u32(0 if g1 is None else len(g1)))
        if g1 is not None:
            out.append(g1_compress(g1))
    return b"".join(out)


class _Reader:
    def __init__(self, buf):
        self.b, self.o = buf, 0

    def take(self, n):
        if self.o + n > len(self.b):
            raise ValueError("truncated proof file")
        v = self.b[self.o:self.o + n]; self.o += n
        return v

    def u32(self, n=1):
        v = struct.unpack("<%dI" % n, self.take(4 * n))
        return v[0] if n == 1 else v


def loads(buf):
    r = _Reader(buf)
    if r.take(8) != MAGIC:
        raise ValueError("not a zkdl_b200 proof file")
    version, nl, batch = r.u32(3)
    if version not in (1, 2, 3):
        raise ValueError("unsupported proof file version %d" % version)
    layers = []
    for _ in range(nl):
        in_dim, out_dim, I, O, ng, nc = r.u32(6)
        layers.append({"in_dim": in_dim, "out_dim": out_dim, "I": I, "O": O,
                       "generators": g1_from_uncompressed(r.take(96 * ng)), "commitment": g1_decompress(r.take(48 * nc))})
    linked = None
    if version == 3:
        linked = {"input": fr_from_bytes(r.take(32 * r.u32())), "output": fr_from_bytes(r.take(32 * r.u32()))}
        linked["aux_com"] = [[g1_from_uncompressed(r.take(96 * r.u32())) for _ in range(3)] for _ in range(r.u32())]
    tasks = []
    for _ in range(r.u32()):
        kind, layer, nch = r.u32(3)
        if kind > 1 or layer >= nl:
            raise ValueError("malformed task record")
        ch = [fr_from_bytes(r.take(32 * r.u32())) for _ in range(nch)]
        fr = fr_from_bytes(r.take(32 * r.u32()))
        ng1 = r.u32()
        tasks.append({"kind": "fc" if kind == 0 else "relu", "layer": layer, "challenges": ch, "fr": fr,
                      "g1": g1_decompress(r.take(48 * ng1)) if ng1 else None})
    if r.o != len(buf):
        raise ValueError("trailing bytes after the last task")
    if version >= 2 and any(t["challenges"] for t in tasks):
        raise ValueError("a Fiat-Shamir proof file must not carry challenges")
    return {"batch": batch, "layers": layers, "fiat_shamir": version == 2, "linked": linked}, tasks
