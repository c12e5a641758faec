"""Fiat-Shamir mode of the layer proofs: every challenge is derived from a SHA-256 transcript
instead of the reference's std::random_device streams (random_vec, /root/reference/proof.cu:3-11; zkfc.cu:135-137,
zkrelu.cu:85-98), so a proof binds to the public model, to the claimed evaluations and to its own earlier rounds, and the
verifier recomputes the challenges instead of trusting the prover's.

Transcript (one per layer proof, seeded from a digest of the public part so that layer proofs stay independent and can
still be spread over streams / GPUs):
    S_0            = SHA-256("zkdl_b200/fs/v1" || root || kind || layer)
    absorb(bytes)  : S <- SHA-256(S || bytes)
    vector draw    : x_i = limbs(SHA-256(S || 0x01 || i_le32)), top limb % 0x73eda753;  then S <- SHA-256(S || 0x02)
    sumcheck round : S <- SHA-256(S || c0 || c1 || c2);  x_j = limbs(S), top limb % 0x73eda753      (ON THE DEVICE,
                     zkdl_sumcheck_fs / csrc/fs_kernels.cu: no host round trip between rounds)
Field elements are hashed as their 32-byte little-endian limb images (Montgomery form, as they appear in the proof); a
challenge is "a value < p read as Montgomery form", the reference's own convention for random_vec.

Proof layout per layer = the reference's; the challenges are not part of the proof.
What is bound: zkFC - (u_bs, u_out) to the public root; Z(u) is absorbed before the matmul sumcheck, whose fold challenges
u_in are the evaluation point of the weight opening.  zkReLU - the eq points to the root, every fold challenge to the
previous rounds, u_recover to both binary sumchecks.  In this mode the per-layer claims are NOT chained to each other, as in the
reference; zkdl_b200/linked.py builds the chained proof (§8f-3) on the same transcript and device-side hashing."""
import hashlib

import numpy as np

from . import capi as zk
from . import verify

DOMAIN = b"zkdl_b200/fs/v1"
TOP = 1944954707


def _fr_bytes(limbs):
    return np.ascontiguousarray(np.asarray(limbs, dtype=np.uint32)).astype("<u4").tobytes()


def _challenge(digest):
    x = np.frombuffer(digest, dtype="<u4").astype(np.uint32).copy()
    x[7] %= TOP
    return x


class Transcript:
    def __init__(self, root, kind, layer):
        self.s = hashlib.sha256(DOMAIN + root + kind.encode() + int(layer).to_bytes(4, "little")).digest()

    def absorb(self, data):
        self.s = hashlib.sha256(self.s + data).digest()

    def absorb_fr(self, limbs):
        self.absorb(_fr_bytes(limbs))

    def vector(self, k):
        out = np.zeros((k, 8), dtype=np.uint32)
        for i in range(k):
            out[i] = _challenge(hashlib.sha256(self.s + b"\x01" + i.to_bytes(4, "little")).digest())
        self.s = hashlib.sha256(self.s + b"\x02").digest()
        return out

    def round(self, c0, c1, c2):
        """What the device does after summing a round's coefficients (fs_kernels.cu transcript_round)."""
        self.s = hashlib.sha256(self.s + _fr_bytes(c0) + _fr_bytes(c1) + _fr_bytes(c2)).digest()
        return _challenge(self.s)

    def rounds(self, proof, k):
        """Replays k sumcheck rounds over proof[0 : 3k]; returns the k fold challenges."""
        out = np.zeros((k, 8), dtype=np.uint32)
        for j in range(k):
            out[j] = self.round(proof[3 * j], proof[3 * j + 1], proof[3 * j + 2])
        return out


def public_root(layers, batch):
    """Digest of the public part: batch, padded shapes, generators and weight commitments (normalised Jacobian limbs)."""
    def canon(points):                                       # infinity (z = 0) has many limb images: hash it as all zeros
        a = np.array(points, dtype=np.uint32).reshape(-1, 36)
        a[(a[:, 24:] == 0).all(axis=1)] = 0
        return a.astype("<u4").tobytes()
    h = hashlib.sha256(DOMAIN + b"/root" + int(batch).to_bytes(4, "little"))
    for L in layers:
        h.update(np.array([L["in_dim"], L["out_dim"], L["I"], L["O"]], dtype="<u4").tobytes())
        h.update(canon(L["generators"]))
        h.update(canon(L["commitment"]))
    return h.digest()


def public_part(P):
    """MLPProver -> the public description a verifier holds (same fields as proof_file.export)."""
    return [{"in_dim": L.in_dim, "out_dim": L.out_dim, "I": L.I, "O": L.O,
             "generators": zk.to_host(zk.g1_normalize(L.G)), "commitment": zk.to_host(zk.g1_normalize(L.com))} for L in P.layers]


def _clog(v):
    return 0 if v <= 1 else (int(v) - 1).bit_length()


# ------------------------------------------------------------------------------------------------ prover
def prove_fc(P, i, root):
    """zkFC::prove (zkfc.cu:128-145) with transcript challenges.  Returns (proof_fr, proof_g1) in the reference's layout."""
    L = P.layers[i]
    B, kb, ki, ko = P.B, _clog(P.B), _clog(L.I), _clog(L.O)
    X = P.A[i - 1] if i > 0 else P.X
    T = Transcript(root, "fc", i)
    u_bs, u_out = T.vector(kb), T.vector(ko)
    z = zk.fr_me(P.Z[i], np.concatenate([u_out, u_bs]))                       # the claim the sumcheck reduces
    T.absorb_fr(zk.to_host(z))
    Xr = zk.fr_partial_me(X, u_bs, L.I) if kb else X
    Wr = zk.fr_partial_me(L.W, u_out, 1)
    ip, u_in, state = zk.sumcheck_fs(zk.FS_IP, Xr, Wr, None, ki, T.s)         # u_in[j] is hashed out of round j on the device
    u_in = zk.to_host(u_in)
    pfr, pg1 = zk.zkfc_prove(X, L.W, P.Z[i], B, L.I, L.O, L.gens, L.com_table, u_bs, u_in, u_out, parts=zk.FC_OPENING, w_int=L.mm)
    nip = 3 * ki + 2
    pfr[:nip] = ip
    pfr[nip] = z[0]
    return pfr, pg1


def prove_relu(P, i, root):
    """zkReLU::prove (zkrelu.cu:79-100) with transcript challenges, on the reference's 0/1 Fr tables."""
    L = P.layers[i]
    n = P.B * L.O
    Lg = _clog(n)
    sign, magp, remp = P.aux[i]
    mag, rem = zk.relu_expand(magp, remp)
    T = Transcript(root, "relu", i)
    u_z = T.vector(Lg + 5)
    p_mag, _, T.s = zk.sumcheck_fs(zk.FS_BIN, mag, None, u_z, Lg + 5, T.s)
    u_r = T.vector(Lg + 4)
    p_rem, _, T.s = zk.sumcheck_fs(zk.FS_BIN, rem, None, u_r, Lg + 4, T.s)
    u_rec = T.vector(Lg)                                                      # after both binary sumchecks
    r_mag, r_rem = zk.fr_partial_me(mag, u_rec, 32), zk.fr_partial_me(rem, u_rec, 16)
    T.absorb_fr(zk.to_host(r_mag)); T.absorb_fr(zk.to_host(r_rem))
    u_hp = T.vector(Lg)
    p_hp, _, T.s = zk.sumcheck_fs(zk.FS_HP, P.Z[i], sign, u_hp, Lg, T.s)
    import torch
    return torch.cat([p_mag, r_mag, p_rem, r_rem, p_hp])


def prove(P):
    """All layer proofs in the reference's order (demo.cu:124-138).  Returns (public, [(kind, layer, proof_fr[, proof_g1])])."""
    public = public_part(P)
    root = public_root(public, P.B)
    nl = len(P.layers)
    out = [("fc", nl - 1) + prove_fc(P, nl - 1, root)]
    for i in range(nl - 2, -1, -1):
        out.append(("relu", i, prove_relu(P, i, root)))
        out.append(("fc", i) + prove_fc(P, i, root))
    return public, out


# ------------------------------------------------------------------------------------------------ verifier
def challenges_fc(root, i, B, I, O, proof_fr):
    """Re-derives (u_bs, u_in, u_out) of layer i's zkFC proof from the transcript."""
    kb, ki, ko = _clog(B), _clog(I), _clog(O)
    fr = np.asarray(proof_fr, dtype=np.uint32).reshape(-1, 8)
    T = Transcript(root, "fc", i)
    u_bs, u_out = T.vector(kb), T.vector(ko)
    T.absorb_fr(fr[3 * ki + 2])
    return u_bs, T.rounds(fr, ki), u_out


def challenges_relu(root, i, n, proof_fr):
    Lg = _clog(n)
    fr = np.asarray(proof_fr, dtype=np.uint32).reshape(-1, 8)
    T = Transcript(root, "relu", i)
    o = 0
    u_z = T.vector(Lg + 5); v_z = T.rounds(fr[o:], Lg + 5); o += 3 * (Lg + 5) + 1
    r_mag = fr[o: o + 32]; o += 32
    u_r = T.vector(Lg + 4); v_r = T.rounds(fr[o:], Lg + 4); o += 3 * (Lg + 4) + 1
    r_rem = fr[o: o + 16]; o += 16
    u_rec = T.vector(Lg)
    T.absorb_fr(r_mag); T.absorb_fr(r_rem)
    u_hp = T.vector(Lg); v_hp = T.rounds(fr[o:], Lg)
    return u_z, v_z, u_r, v_r, u_rec, u_hp, v_hp


def verify_all(public, batch, proofs):
    """Verifier for prove(): recomputes every challenge from the transcript, then runs the checks of zkdl_b200/verify.py
    (sumcheck rounds, opening recursion, com(u_hi) against the public commitments, W~(u) against the sumcheck's final value).
    Raises verify.VerifyError."""
    root = public_root(public, batch)
    nl = len(public)
    expected = [("fc", nl - 1)] + [(k, i) for i in range(nl - 2, -1, -1) for k in ("relu", "fc")]
    if [(p[0], p[1]) for p in proofs] != expected:
        raise verify.VerifyError("not exactly the layer proofs of the public model")
    for p in proofs:
        L = public[p[1]]
        if p[0] == "fc":
            fr = zk.to_host(p[2])
            u_bs, u_in, u_out = challenges_fc(root, p[1], batch, L["I"], L["O"], fr)
            G, com = zk.to_device(L["generators"]), zk.to_device(L["commitment"])
            verify.verify_zkfc(p[2], p[3], G, batch, L["I"], L["O"], u_bs, u_in, u_out)
            u = np.concatenate([u_out.reshape(-1, 8), u_in.reshape(-1, 8)])
            klo = (G.shape[0] - 1).bit_length()
            verify.verify_commitment_eval(com, p[3][:1], u[klo:])
        else:
            n = batch * L["O"]
            u_z, v_z, u_r, v_r, u_rec, u_hp, v_hp = challenges_relu(root, p[1], n, zk.to_host(p[2]))
            verify.verify_zkrelu(p[2], n, u_z, v_z, u_r, v_r, u_hp, v_hp)
    return True
