"""TEST INFRASTRUCTURE: the subset of zkdl_b200.capi that the host-side protocol logic (fiat_shamir.py, linked.py, verify.py)
calls, answered by the CPU oracle on torch CPU tensors.  It lets the -m "not gpu" suite run the transcript / chaining /
verifier logic on toy shapes without a GPU; the product never imports it (the C-ABI library has no CPU path)."""
import hashlib

import numpy as np
import torch

from oracle import oracle as orc

OP_ADD, OP_SUB, OP_MUL, OP_NEG, OP_MONT, OP_UNMONT = range(6)
FS_IP, FS_HP, FS_BIN = 0, 1, 2
TOP = 1944954707
FQ_ONE = [196605, 1980301312, 3289120770, 3958636555, 1405573306, 1598593111, 1884444485, 2010011731, 2723605613, 1543969431, 4202751123, 368467651]


def to_host(t):
    return t.detach().cpu().numpy().view(np.uint32)


def to_device(arr):
    a = np.ascontiguousarray(np.asarray(arr, dtype=np.uint32))
    return torch.from_numpy(a.view(np.int32).copy())


def _u(rows):
    if rows is None:
        return np.zeros((0, 8), np.uint32)
    return np.asarray(rows, dtype=np.uint32).reshape(-1, 8)


def fr_elementwise(op, a, b=None, out=None):
    assert op in (OP_MONT, OP_UNMONT) and b is None
    res = to_device((orc.fr_mont if op == OP_MONT else orc.fr_unmont)(to_host(a)))
    if out is not None:
        out.copy_(res)
        return out
    return res


def fr_me(a, u):
    return to_device(orc.fr_me(to_host(a), _u(u)).reshape(1, 8))


def fr_partial_me(a, u, window):
    return to_device(orc.fr_partial_me(to_host(a), _u(u), window))


def relu_expand(magp, remp):
    m = magp.numpy().view(np.uint32).astype(np.uint64)
    r = remp.numpy().view(np.uint16).astype(np.uint64)
    one = orc.fr_mont(orc.to_limbs([1]))[0]
    mag = np.zeros((len(m) * 32, 8), np.uint32)
    rem = np.zeros((len(r) * 16, 8), np.uint32)
    mb = ((m[:, None] >> np.arange(32, dtype=np.uint64)) & 1).reshape(-1).astype(bool)
    rb = ((r[:, None] >> np.arange(16, dtype=np.uint64)) & 1).reshape(-1).astype(bool)
    mag[mb] = one
    rem[rb] = one
    return to_device(mag), to_device(rem)


def _challenge(digest):
    x = np.frombuffer(digest, dtype="<u4").astype(np.uint32).copy()
    x[7] %= TOP
    return x


def sumcheck_fs(kind, a, b, u_eq, k, state):
    """The device transcript of csrc/fs_kernels.cu restated round by round: S <- SHA-256(S || c0 || c1 || c2), x_j = limbs(S)."""
    a = to_host(a).copy()
    b = to_host(b).copy() if b is not None else None
    u = _u(u_eq)
    s = bytes(state)
    rows, xs = [], np.zeros((max(k, 1), 8), np.uint32)
    zeros = np.zeros((k, 8), np.uint32)
    for j in range(k):
        if kind == FS_IP:
            tri = orc.ip_sumcheck(a, b, zeros[j:])[:3]
        elif kind == FS_HP:
            tri = orc.hp_sumcheck(a, b, u[j:], zeros[j:])[:3]
        else:
            tri = orc.bin_sumcheck(a, u[j:], zeros[j:])[:3]
        rows.append(tri)
        s = hashlib.sha256(s + tri.astype("<u4").tobytes()).digest()
        x = _challenge(s)
        xs[j] = x
        a = orc.fr_partial_me(a, x.reshape(1, 8), 1)
        if b is not None:
            b = orc.fr_partial_me(b, x.reshape(1, 8), 1)
    fin = [a[:1]] + ([b[:1]] if kind != FS_BIN else [])
    proof = np.concatenate(rows + fin) if rows else np.concatenate(fin)
    return to_device(proof), to_device(xs[:k]), s


# ------------------------------------------------------------------ G1
def g1_normalize(a):
    aff, inf = orc.g1_to_affine(to_host(a))
    out = np.zeros((len(aff), 36), np.uint32)
    out[:, :24] = aff
    out[:, 24:] = np.array(FQ_ONE, np.uint32)
    out[inf] = 0
    return to_device(out)


def g1_mul(P, x):
    return to_device(orc.g1_mul(to_host(P), to_host(x), fast=True))


def g1_sum(a):
    return to_device(orc.g1_sum(to_host(a)))


class G1Table:
    def __init__(self, points, full=True):
        self.points = to_host(points).copy()
        self.n = self.points.shape[0]

    def close(self):
        pass


def msm(table, scalars, m, scalars_mont):
    s = to_host(scalars)[: m * table.n]
    if scalars_mont:
        s = orc.fr_unmont(s)
    prod = orc.g1_mul(table.points, s, fast=True)
    return to_device(np.concatenate([orc.g1_sum(prod[r * table.n:(r + 1) * table.n]) for r in range(m)]))


def commit(table, t):
    return to_device(orc.commit(table.points, to_host(t), fast=True))


def open_(gens, com_table, t, u):
    g, proof, ret = orc.open_(to_host(t), gens.points, com_table.points, _u(u), fast=True)
    return to_device(g), to_device(proof), to_device(ret.reshape(1, 8))


# ------------------------------------------------------------------ a toy prover state (what mlp.MLPProver holds after forward())
class _Layer:
    pass


class ToyProver:
    def __init__(self, dims, batch, seed=0, weight_scale=1.0, input_scale=1.0):
        rng = np.random.default_rng(seed)
        gen = orc.g1_generator()
        B = 1 << max(0, (batch - 1).bit_length())
        self.B, self.layers = B, []
        for (i_dim, o_dim) in dims:
            L = _Layer()
            L.in_dim, L.out_dim = i_dim, o_dim
            L.I, L.O = 1 << (i_dim - 1).bit_length(), 1 << (o_dim - 1).bit_length()
            L.ngens = 1 << (((i_dim * o_dim - 1).bit_length() + 1) // 2)
            sc = orc.to_limbs([int(v) for v in rng.integers(1, 1 << 62, L.ngens)])
            L.G = to_device(orc.g1_mul(gen, sc, fast=True))
            L.gens = G1Table(L.G)
            w = ((rng.random((i_dim, o_dim)) * 2 - 1) / np.sqrt(i_dim) * weight_scale).astype(np.float32)
            L.W = to_device(orc.fr_mont(orc.float_to_fr(w, L.I, L.O)))
            L.com = commit(L.gens, L.W)
            L.com_table = G1Table(L.com)
            self.layers.append(L)
        x = (rng.standard_normal((batch, dims[0][0])) * input_scale).astype(np.float32)
        self.X = to_device(orc.fr_mont(orc.float_to_fr(x, B, self.layers[0].I)))
        self.forward_from(0, to_host(self.X))

    def forward_from(self, first, cur):
        """(Re)computes layers first.. from the activation table `cur` (numpy limbs)."""
        if first == 0:
            self.Z, self.A, self.aux = [], [], []
        del self.Z[first:], self.A[first:], self.aux[first:]
        B = self.B
        for i, L in enumerate(self.layers):
            if i < first:
                continue
            z = orc.fr_matmul(cur, to_host(L.W), B, L.I, L.O)
            self.Z.append(to_device(z))
            if i + 1 < len(self.layers):
                a, sign, mag, rem, bad = orc.relu(z)
                assert bad == 0
                one = orc.fr_mont(orc.to_limbs([1]))[0]
                mb = (mag == one).all(axis=1).reshape(-1, 32).astype(np.uint64)
                rb = (rem == one).all(axis=1).reshape(-1, 16).astype(np.uint64)
                magp = (mb << np.arange(32, dtype=np.uint64)).sum(axis=1).astype(np.uint32).view(np.int32)
                remp = (rb << np.arange(16, dtype=np.uint64)).sum(axis=1).astype(np.uint16).view(np.int16)
                self.A.append(to_device(a))
                self.aux.append((to_device(sign), torch.from_numpy(magp.copy()), torch.from_numpy(remp.copy())))
                cur = a


class HostCopy:
    """A device prover's state (mlp.MLPProver after forward()) copied to CPU tensors, for running the same host logic on the oracle."""

    def __init__(self, P, dev_zk):
        self.B, self.layers = P.B, []
        for L in P.layers:
            C = _Layer()
            for k in ("in_dim", "out_dim", "I", "O", "ngens"):
                setattr(C, k, getattr(L, k))
            C.G, C.W, C.com = (to_device(dev_zk.to_host(t)) for t in (L.G, L.W, L.com))
            C.gens, C.com_table = G1Table(C.G), G1Table(C.com)
            self.layers.append(C)
        self.X = to_device(dev_zk.to_host(P.X))
        self.Z = [to_device(dev_zk.to_host(z)) for z in P.Z]
        self.A = [to_device(dev_zk.to_host(a)) for a in P.A]
        self.aux = [(to_device(dev_zk.to_host(s)), m.cpu().clone(), r.cpu().clone()) for s, m, r in P.aux]
