"""ctypes binding of the CPU oracle (oracle/zkdl_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs as the checker.  The product (zkdl_b200/) never imports this module.

Array conventions (numpy, C-contiguous, dtype uint32, little-endian limbs as in the reference PODs):
  Fr  -> [n, 8]    Fq -> [n, 12]    G1 affine -> [n, 24]    G1 Jacobian -> [n, 36]
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

FR_P = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
FQ_P = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
FR_R = (1 << 256) % FR_P
FQ_R = (1 << 384) % FQ_P


def build():
    """Compile libzkdl_oracle.so (gcc, seconds)."""
    subprocess.check_call(["make", "-s", "-C", _HERE, "libzkdl_oracle.so"])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libzkdl_oracle.so")
        src = os.path.join(_HERE, "zkdl_oracle.c")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
            build()
        _LIB = C.CDLL(path)
        _LIB.orc_fr_partial_me.restype = C.c_size_t
        _LIB.orc_relu.restype = C.c_size_t
        _LIB.orc_ceil_log2.restype = C.c_uint32
        _LIB.orc_g1_eq.restype = C.c_int
        _LIB.orc_g1_on_curve.restype = C.c_int
        _LIB.orc_num_threads.restype = C.c_int
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _c(a, w):
    a = np.ascontiguousarray(a, dtype=np.uint32)
    assert a.ndim == 2 and a.shape[1] == w, (a.shape, w)
    return a


def _sz(n):
    return C.c_size_t(int(n))


# ---------------------------------------------------------------- int <-> limb helpers
def to_limbs(vals, nlimbs=8):
    vals = list(vals)
    out = np.zeros((len(vals), nlimbs), dtype=np.uint32)
    for i, v in enumerate(vals):
        v = int(v)
        for j in range(nlimbs):
            out[i, j] = (v >> (32 * j)) & 0xFFFFFFFF
    return out


def from_limbs(arr):
    arr = np.asarray(arr, dtype=np.uint32)
    if arr.ndim == 1:
        arr = arr[None, :]
    out = []
    for row in arr:
        v = 0
        for j, x in enumerate(row):
            v |= int(x) << (32 * j)
        out.append(v)
    return out


def fr_from_ints(vals, mont=True):
    """Signed python ints -> Fr limbs (Montgomery form if mont)."""
    return to_limbs([((int(v) % FR_P) * (FR_R if mont else 1)) % FR_P for v in vals], 8)


def fr_to_ints(arr, mont=True):
    rinv = pow(FR_R, -1, FR_P)
    return [(v * rinv) % FR_P if mont else v for v in from_limbs(arr)]


# ---------------------------------------------------------------- Fr
def _bin(fn, a, b):
    a, b = _c(a, 8), _c(b, 8)
    out = np.empty_like(a)
    getattr(lib(), fn)(_p(a), _p(b), _p(out), _sz(len(a)))
    return out


def _un(fn, a):
    a = _c(a, 8)
    out = np.empty_like(a)
    getattr(lib(), fn)(_p(a), _p(out), _sz(len(a)))
    return out


def fr_add(a, b): return _bin("orc_fr_add", a, b)
def fr_sub(a, b): return _bin("orc_fr_sub", a, b)
def fr_mul(a, b): return _bin("orc_fr_mul", a, b)
def fr_mont(a): return _un("orc_fr_mont", a)
def fr_unmont(a): return _un("orc_fr_unmont", a)
def fr_neg(a): return _un("orc_fr_neg", a)


def fr_bcast(a, x, op):
    a, x = _c(a, 8), _c(np.asarray(x).reshape(1, 8), 8)
    out = np.empty_like(a)
    lib().orc_fr_bcast(_p(a), _p(x), C.c_int({"add": 0, "sub": 1, "mul": 2}[op]), _p(out), _sz(len(a)))
    return out


def fr_sum(a):
    a = _c(a, 8)
    out = np.zeros((1, 8), np.uint32)
    lib().orc_fr_sum(_p(a), _sz(len(a)), _p(out))
    return out[0]


def fr_me(a, u):
    a, u = _c(a, 8), _c(np.asarray(u).reshape(-1, 8), 8)
    out = np.zeros((1, 8), np.uint32)
    lib().orc_fr_me(_p(a), _sz(len(a)), _p(u), _sz(len(u)), _p(out))
    return out[0]


def fr_partial_me(a, u, window):
    a, u = _c(a, 8), _c(np.asarray(u).reshape(-1, 8), 8)
    out = np.zeros((len(a) + 2 * int(window) + 1, 8), np.uint32)
    n = lib().orc_fr_partial_me(_p(a), _sz(len(a)), _p(u), _sz(len(u)), _sz(window), _p(out))
    return out[:n].copy()


def ip_sumcheck(a, b, u):
    a, b, u = _c(a, 8), _c(b, 8), _c(np.asarray(u).reshape(-1, 8), 8)
    proof = np.zeros((3 * len(u) + 2, 8), np.uint32)
    lib().orc_ip_sumcheck(_p(a), _p(b), _sz(len(a)), _p(u), _sz(len(u)), _p(proof))
    return proof


def hp_sumcheck(a, b, u, v):
    a, b = _c(a, 8), _c(b, 8)
    u, v = _c(np.asarray(u).reshape(-1, 8), 8), _c(np.asarray(v).reshape(-1, 8), 8)
    proof = np.zeros((3 * len(u) + 2, 8), np.uint32)
    lib().orc_hp_sumcheck(_p(a), _p(b), _sz(len(a)), _p(u), _p(v), _sz(len(u)), _p(proof))
    return proof


def bin_sumcheck(a, u, v):
    a = _c(a, 8)
    u, v = _c(np.asarray(u).reshape(-1, 8), 8), _c(np.asarray(v).reshape(-1, 8), 8)
    proof = np.zeros((3 * len(u) + 1, 8), np.uint32)
    lib().orc_bin_sumcheck(_p(a), _sz(len(a)), _p(u), _p(v), _sz(len(u)), _p(proof))
    return proof


def random_vec(seed, n):
    out = np.zeros((n, 8), np.uint32)
    lib().orc_random_vec(C.c_uint32(seed), _sz(n), _p(out))
    return out


def ceil_log2(n):
    return int(lib().orc_ceil_log2(C.c_uint32(n)))


def float_to_fr(fs, rows_out, cols_out):
    fs = np.ascontiguousarray(fs, dtype=np.float32)
    assert fs.ndim == 2
    out = np.zeros((rows_out * cols_out, 8), np.uint32)
    lib().orc_float_to_fr(_p(fs), _p(out), C.c_uint32(fs.shape[0]), C.c_uint32(rows_out), C.c_uint32(fs.shape[1]), C.c_uint32(cols_out))
    return out


def fr_matmul(A, B, rows_a, cols_a, cols_b):
    A, B = _c(A, 8), _c(B, 8)
    assert len(A) == rows_a * cols_a and len(B) == cols_a * cols_b
    out = np.zeros((rows_a * cols_b, 8), np.uint32)
    lib().orc_fr_matmul(_p(A), _p(B), _p(out), _sz(rows_a), _sz(cols_a), _sz(cols_b))
    return out


def relu(X):
    X = _c(X, 8)
    n = len(X)
    Z, sign = np.zeros((n, 8), np.uint32), np.zeros((n, 8), np.uint32)
    mag, rem = np.zeros((32 * n, 8), np.uint32), np.zeros((16 * n, 8), np.uint32)
    bad = lib().orc_relu(_p(X), _p(Z), _p(sign), _p(mag), _p(rem), _sz(n))
    return Z, sign, mag, rem, int(bad)


# ---------------------------------------------------------------- G1
def fq_mul(a, b):
    a, b = _c(a, 12), _c(b, 12)
    out = np.empty_like(a)
    lib().orc_fq_mul(_p(a), _p(b), _p(out), _sz(len(a)))
    return out


def g1_generator():
    out = np.zeros((1, 36), np.uint32)
    lib().orc_g1_generator(_p(out))
    return out


def g1_double(a):
    a = _c(a, 36); out = np.empty_like(a)
    lib().orc_g1_double(_p(a), _p(out), _sz(len(a))); return out


def g1_add(a, b):
    a, b = _c(a, 36), _c(b, 36); out = np.empty_like(a)
    lib().orc_g1_add(_p(a), _p(b), _p(out), _sz(len(a))); return out


def g1_add_mixed(a, b):
    a, b = _c(a, 36), _c(b, 24); out = np.empty_like(a)
    lib().orc_g1_add_mixed(_p(a), _p(b), _p(out), _sz(len(a))); return out


def g1_neg(a):
    a = _c(a, 36); out = np.empty_like(a)
    lib().orc_g1_neg(_p(a), _p(out), _sz(len(a))); return out


def g1_mul(P, x, fast=False):
    """out[i] = [x[i]] P[i mod len(P)] with the raw limbs of x as the integer scalar."""
    P, x = _c(P, 36), _c(x, 8)
    out = np.zeros((len(x), 36), np.uint32)
    (lib().orc_g1_mul_fast if fast else lib().orc_g1_mul)(_p(P), _sz(len(P)), _p(x), _sz(len(x)), _p(out))
    return out


def g1_sum(a):
    a = _c(a, 36); out = np.zeros((1, 36), np.uint32)
    lib().orc_g1_sum(_p(a), _sz(len(a)), _p(out)); return out


def g1_me(a, u):
    a, u = _c(a, 36), _c(np.asarray(u).reshape(-1, 8), 8)
    out = np.zeros((1, 36), np.uint32)
    lib().orc_g1_me(_p(a), _sz(len(a)), _p(u), _sz(len(u)), _p(out)); return out


def g1_to_affine(a):
    a = _c(a, 36)
    out = np.zeros((len(a), 24), np.uint32); inf = np.zeros(len(a), np.uint8)
    lib().orc_g1_to_affine(_p(a), _p(out), _p(inf), _sz(len(a)))
    return out, inf.astype(bool)


def g1_eq(a, b):
    """Elementwise same-point test (projective)."""
    a, b = _c(a, 36), _c(b, 36)
    assert a.shape == b.shape
    return np.array([bool(lib().orc_g1_eq(_p(a[i:i + 1]), _p(b[i:i + 1]))) for i in range(len(a))])


def g1_on_curve(a):
    a = _c(a, 36)
    return np.array([bool(lib().orc_g1_on_curve(_p(a[i:i + 1]))) for i in range(len(a))])


def commit(G, t, fast=False):
    G, t = _c(G, 36), _c(t, 8)
    assert len(t) % len(G) == 0
    com = np.zeros((len(t) // len(G), 36), np.uint32)
    lib().orc_commit(_p(G), _sz(len(G)), _p(t), _sz(len(t)), _p(com), C.c_int(int(fast)))
    return com


def commit_as_written(G, t):
    G, t = _c(G, 36), _c(t, 8)
    com = np.zeros((len(t) // len(G), 36), np.uint32)
    lib().orc_commit_as_written(_p(G), _sz(len(G)), _p(t), _sz(len(t)), _p(com))
    return com


def me_open(t, G, u, fast=False):
    t, G, u = _c(t, 8), _c(G, 36), _c(np.asarray(u).reshape(-1, 8), 8)
    assert len(t) == len(G)
    proof = np.zeros((3 * len(u) + 1, 36), np.uint32); ret = np.zeros((1, 8), np.uint32)
    lib().orc_me_open(_p(t), _p(G), _sz(len(t)), _p(u), _sz(len(u)), _p(proof), _p(ret), C.c_int(int(fast)))
    return proof, ret[0]


def open_(t, G, com, u, fast=False):
    """Commitment::open (commitment.cu:83-92): returns (com(u_hi), me_open proof, final scalar)."""
    G, com = _c(G, 36), _c(com, 36)
    u = _c(np.asarray(u).reshape(-1, 8), 8)
    k = ceil_log2(len(com))
    u_hi, u_lo = u[len(u) - k:], u[:len(u) - k]
    if len(G) != (1 << len(u_lo)):
        raise ValueError("Incompatible dimensions")
    g_temp = g1_me(com, u_hi)
    tf = fr_partial_me(t, u_hi, 1 << len(u_lo))
    proof, ret = me_open(tf, G, u_lo, fast=fast)
    return g_temp, proof, ret


def msm_pippenger(bases_affine, scalars_plain, threads=0):
    b, s = _c(bases_affine, 24), _c(scalars_plain, 8)
    out = np.zeros((1, 36), np.uint32)
    lib().orc_msm_pippenger(_p(b), _p(s), _sz(len(b)), _p(out), C.c_int(threads))
    return out


def num_threads():
    return int(lib().orc_num_threads())


# ---------------------------------------------------------------- compositions (zkfc.cu:128-145, zkrelu.cu:79-100)
def zkfc_prove(X, W, Z, G, com, B, I, O, u_bs, u_in, u_out, fast=True):
    """Element order of SURVEY App. A.12.  X,W,Z Montgomery Fr tables; returns dict of proof parts."""
    Xr = fr_partial_me(X, u_bs, I) if len(u_bs) else np.array(X, copy=True)
    Wr = fr_partial_me(W, u_out, 1)
    ip = ip_sumcheck(Xr, Wr, u_in)
    zu = fr_me(Z, np.concatenate([u_out, u_bs]))
    g_temp, opening, ret = open_(W, G, com, np.concatenate([u_out, u_in]), fast=fast)
    return {"ip": ip, "z_eval": zu, "com_eval": g_temp, "opening": opening, "open_ret": ret}


def zkrelu_prove(Xpre, sign, mag_bin, rem_bin, u_z, v_z, u_r, v_r, u_rec, u_hp, v_hp):
    out = {}
    out["mag_sc"] = bin_sumcheck(mag_bin, u_z, v_z)
    out["mag_rec"] = fr_partial_me(mag_bin, u_rec, 32)
    out["rem_sc"] = bin_sumcheck(rem_bin, u_r, v_r)
    out["rem_rec"] = fr_partial_me(rem_bin, u_rec, 16)
    out["hp"] = hp_sumcheck(Xpre, sign, u_hp, v_hp)
    return out
