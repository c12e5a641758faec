"""ctypes binding of libzkdl_b200.so (the C ABI in include/zkdl_b200.h).

PyTorch is used only for device memory and streams.  Tensors are torch.int32 CUDA tensors viewed as limb arrays:
Fr -> [n, 8], G1 affine -> [n, 24], G1 Jacobian -> [n, 36].  There is no CPU fallback: importing works anywhere
(so the CPU test-suite can check the exported symbols), every compute call needs a CUDA device and raises otherwise.
"""
import ctypes as C
import os
import re
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
LIB_PATH = os.environ.get("ZKDL_LIB", os.path.join(_HERE, "libzkdl_b200.so"))      # ZKDL_LIB: A/B a differently tuned build of the same library
HEADER_PATH = os.path.join(_ROOT, "include", "zkdl_b200.h")
_lib = None

ERR = {1: "Incompatible dimensions", 2: "CUDA error", 3: "bad argument", 4: "NCCL error"}
OP_ADD, OP_SUB, OP_MUL, OP_NEG, OP_MONT, OP_UNMONT = range(6)
G1_ADD, G1_SUB, G1_NEG, G1_MADD, G1_MSUB = range(5)


class ZkdlError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"zkdl_b200 error {code} ({ERR.get(code, '?')}): {msg}")
        self.code = code


class DimensionError(ZkdlError):
    """Raised where the reference throws std::runtime_error("Incompatible dimensions")."""


def build(verbose=False):
    """Compile libzkdl_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-C", os.path.join(_HERE, "csrc"), "-j4"]
    subprocess.check_call(cmd, stdout=None if verbose else subprocess.DEVNULL)


def declared_symbols():
    """Every function name declared in include/zkdl_b200.h."""
    src = open(HEADER_PATH).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(zkdl_[a-z0-9_]+)\s*\(", src)))


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(the CUDA extension is the product; there is no fallback)")
        _lib = C.CDLL(LIB_PATH)
        _lib.zkdl_last_error.restype = C.c_char_p
        _lib.zkdl_launch_count.restype = C.c_uint64
        _lib.zkdl_partial_me_size.restype = C.c_size_t
        _lib.zkdl_partial_me_size.argtypes = [C.c_size_t, C.c_size_t, C.c_size_t]
        _lib.zkdl_zkrelu_proof_size.restype = C.c_size_t
        _lib.zkdl_zkrelu_proof_size.argtypes = [C.c_size_t]
        _lib.zkdl_g1_table_size.restype = C.c_size_t
        _lib.zkdl_g1_table_bytes.restype = C.c_size_t
        _lib.zkdl_ceil_log2.restype = C.c_uint32
    return _lib


def _check(rc):
    if rc != 0:
        msg = lib().zkdl_last_error().decode()
        raise (DimensionError if rc == 1 else ZkdlError)(rc, msg)


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("zkdl_b200 needs a CUDA device (no CPU fallback)")
    return torch


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream():
    return C.c_void_p(_torch().cuda.current_stream().cuda_stream)


def _sz(n):
    return C.c_size_t(int(n))


def _host_fr(a):
    """numpy [k,8] uint32 (or None) -> ctypes pointer kept alive by the returned array."""
    import numpy as np
    if a is None:
        return None, C.c_void_p(0)
    a = np.ascontiguousarray(np.asarray(a, dtype=np.uint32).reshape(-1, 8))
    return a, a.ctypes.data_as(C.c_void_p)


def empty(n, width):
    torch = _torch()
    return torch.empty((int(n), width), dtype=torch.int32, device="cuda")


def to_device(arr):
    """numpy uint32 limb array -> CUDA int32 tensor."""
    import numpy as np
    torch = _torch()
    a = np.ascontiguousarray(np.asarray(arr, dtype=np.uint32))
    return torch.from_numpy(a.view(np.int32)).cuda()


def to_host(t):
    import numpy as np
    return t.detach().cpu().numpy().view(np.uint32)


def launch_count():
    return int(lib().zkdl_launch_count())


# ------------------------------------------------------------------ Fr
def fr_elementwise(op, a, b=None, out=None):
    out = empty(a.shape[0], 8) if out is None else out
    if b is not None and b.shape[0] != a.shape[0]:
        raise DimensionError(1, "Incompatible dimensions")          # fr-tensor.cu:124
    _check(lib().zkdl_fr_elementwise(op, _ptr(a), _ptr(b), _ptr(out), _sz(a.shape[0]), _stream()))
    return out


def fr_broadcast(op, a, x_host, out=None):
    out = empty(a.shape[0], 8) if out is None else out
    keep, xp = _host_fr(x_host)
    _check(lib().zkdl_fr_broadcast(op, _ptr(a), xp, _ptr(out), _sz(a.shape[0]), _stream()))
    return out


def fr_sum(a):
    out = empty(1, 8)
    _check(lib().zkdl_fr_sum(_ptr(a), _sz(a.shape[0]), _ptr(out), _stream()))
    return out


def fr_fold(a, x_host):
    n = a.shape[0]
    out = empty((n + 1) // 2, 8)
    keep, xp = _host_fr(x_host)
    _check(lib().zkdl_fr_fold(_ptr(a), _ptr(out), xp, _sz(n), _stream()))
    return out


def fr_partial_fold(a, x_host, window):
    n = a.shape[0]
    out = empty(window * ((n + 2 * window - 1) // (2 * window)), 8)
    keep, xp = _host_fr(x_host)
    _check(lib().zkdl_fr_partial_fold(_ptr(a), _ptr(out), xp, _sz(n), _sz(window), _stream()))
    return out


def fr_me(a, u_host):
    keep, up = _host_fr(u_host)
    k = 0 if keep is None else len(keep)
    out = empty(1, 8)
    _check(lib().zkdl_fr_me(_ptr(a), _sz(a.shape[0]), up, _sz(k), _ptr(out), _stream()))
    return out


def fr_partial_me(a, u_host, window):
    keep, up = _host_fr(u_host)
    k = 0 if keep is None else len(keep)
    n = a.shape[0]
    out = empty(max(1, lib().zkdl_partial_me_size(n, k, window)), 8)
    _check(lib().zkdl_fr_partial_me(_ptr(a), _sz(n), up, _sz(k), _sz(window), _ptr(out), _stream()))
    return out[: lib().zkdl_partial_me_size(n, k, window)]


def ip_sumcheck(a, b, u_host):
    keep, up = _host_fr(u_host)
    k = len(keep)
    if a.shape[0] != b.shape[0]:
        raise DimensionError(1, "Incompatible dimensions")
    proof = empty(3 * k + 2, 8)
    _check(lib().zkdl_ip_sumcheck(_ptr(a), _ptr(b), _sz(a.shape[0]), up, _sz(k), _ptr(proof), _stream()))
    return proof


def hp_sumcheck(a, b, u_host, v_host):
    ku, up = _host_fr(u_host)
    kv, vp = _host_fr(v_host)
    if len(ku) != len(kv) or a.shape[0] != b.shape[0]:
        raise DimensionError(1, "Incompatible dimensions")
    k = len(ku)
    proof = empty(3 * k + 2, 8)
    _check(lib().zkdl_hp_sumcheck(_ptr(a), _ptr(b), _sz(a.shape[0]), up, vp, _sz(k), _ptr(proof), _stream()))
    return proof


def bin_sumcheck(a, u_host, v_host):
    ku, up = _host_fr(u_host)
    kv, vp = _host_fr(v_host)
    if len(ku) != len(kv):
        raise DimensionError(1, "Incompatible dimensions")
    k = len(ku)
    proof = empty(3 * k + 1, 8)
    _check(lib().zkdl_bin_sumcheck(_ptr(a), _sz(a.shape[0]), up, vp, _sz(k), _ptr(proof), _stream()))
    return proof


def fr_random(n, seed):
    """FrTensor::random (fr-tensor.cu:337-368) with an explicit 64-bit seed: the reference's curand XORWOW stream."""
    out = empty(n, 8)
    _check(lib().zkdl_fr_random(_ptr(out), _sz(n), C.c_uint64(int(seed)), _stream()))
    return out


def fr_random_int(n, num_bits, seed):
    """FrTensor::random_int (fr-tensor.cu:302-335) with an explicit seed."""
    out = empty(n, 8)
    _check(lib().zkdl_fr_random_int(_ptr(out), C.c_uint32(num_bits), _sz(n), C.c_uint64(int(seed)), _stream()))
    return out


FS_IP, FS_HP, FS_BIN = 0, 1, 2


def sumcheck_fs(kind, a, b, u_eq, k, state):
    """zkdl_sumcheck_fs: a sumcheck whose fold challenges come from a SHA-256 transcript ON THE DEVICE.  state: 32 bytes.
    Returns (proof [3k + 1|2, 8], challenges [k, 8], state_out bytes) - reading the last two synchronises the stream."""
    import numpy as np
    torch = _torch()
    nfin = 1 if kind == FS_BIN else 2
    proof, xs = empty(3 * k + nfin, 8), empty(max(k, 1), 8)
    st_out = torch.empty(8, dtype=torch.int32, device="cuda")
    keep, up = _host_fr(u_eq if (u_eq is not None and len(u_eq)) else None)
    sb = (C.c_uint8 * 32).from_buffer_copy(bytes(state))
    _check(lib().zkdl_sumcheck_fs(C.c_int(kind), _ptr(a), _ptr(b), _sz(a.shape[0]), up, _sz(k), sb, _ptr(proof), _ptr(xs), _ptr(st_out), _stream()))
    words = st_out.cpu().numpy().view(np.uint32)                   # big-endian digest words
    return proof, xs[:k], b"".join(int(w).to_bytes(4, "big") for w in words)


def float_to_fr(fs, rows_out, cols_out):
    """fs: float32 CUDA tensor [rows, cols] -> Fr [rows_out*cols_out, 8] (not Montgomery)."""
    fs = fs.contiguous()
    out = empty(rows_out * cols_out, 8)
    _check(lib().zkdl_float_to_fr(_ptr(fs), _ptr(out), C.c_uint32(fs.shape[0]), C.c_uint32(rows_out), C.c_uint32(fs.shape[1]),
                                  C.c_uint32(cols_out), _stream()))
    return out


def fr_matmul(A, B, rows_a, cols_a, cols_b):
    if A.shape[0] != rows_a * cols_a or B.shape[0] != cols_a * cols_b:
        raise DimensionError(1, "Incompatible dimensions")
    out = empty(rows_a * cols_b, 8)
    _check(lib().zkdl_fr_matmul(_ptr(A), _ptr(B), _ptr(out), _sz(rows_a), _sz(cols_a), _sz(cols_b), _stream()))
    return out


class MatmulWeights:
    """Quantised integer copy of a weight table for fr_matmul (zkdl_mm_weights)."""

    def __init__(self, W, rows, cols):
        if W.shape[0] != rows * cols:
            raise DimensionError(1, "Incompatible dimensions")
        self.W, self.rows, self.cols = W, rows, cols
        self.handle = C.c_void_p()
        _check(lib().zkdl_mm_weights_create(_ptr(W), _sz(rows), _sz(cols), C.byref(self.handle), _stream()))

    def close(self):
        if self.handle:
            lib().zkdl_mm_weights_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def fr_matmul_prepared(A, weights, rows_a):
    if A.shape[0] != rows_a * weights.rows:
        raise DimensionError(1, "Incompatible dimensions")
    out = empty(rows_a * weights.cols, 8)
    _check(lib().zkdl_fr_matmul_prepared(_ptr(A), _ptr(weights.W), weights.handle, _ptr(out), _sz(rows_a), _stream()))
    return out


def fr_matmul_prepared_relu(A, weights, rows_a):
    """zkFC::operator() + zkReLU::operator() of one hidden layer: (Z, act, sign, packed mag, packed rem, out-of-range counter)."""
    torch = _torch()
    if A.shape[0] != rows_a * weights.rows:
        raise DimensionError(1, "Incompatible dimensions")
    n = rows_a * weights.cols
    Z, act, sign = empty(n, 8), empty(n, 8), empty(n, 8)
    mag = torch.empty(n, dtype=torch.int32, device="cuda"); rem = torch.empty(n, dtype=torch.int16, device="cuda")
    bad = torch.empty(1, dtype=torch.int32, device="cuda")                     # zeroed by the library
    _check(lib().zkdl_fr_matmul_prepared_relu(_ptr(A), _ptr(weights.W), weights.handle, _ptr(Z), _sz(rows_a), _ptr(act), _ptr(sign),
                                              _ptr(mag), _ptr(rem), _ptr(bad), _stream()))
    return Z, act, sign, mag, rem, bad


def relu(X):
    torch = _torch()
    n = X.shape[0]
    Z, sign, mag, rem = empty(n, 8), empty(n, 8), empty(32 * n, 8), empty(16 * n, 8)
    bad = torch.zeros(1, dtype=torch.int32, device="cuda")
    _check(lib().zkdl_relu(_ptr(X), _ptr(Z), _ptr(sign), _ptr(mag), _ptr(rem), _sz(n), _ptr(bad), _stream()))
    return Z, sign, mag, rem, bad


def relu_packed(X):
    """Z, sign, packed mag (int32 [n]), packed rem (int16 [n]), out-of-range counter."""
    torch = _torch()
    n = X.shape[0]
    Z, sign = empty(n, 8), empty(n, 8)
    mag = torch.empty(n, dtype=torch.int32, device="cuda"); rem = torch.empty(n, dtype=torch.int16, device="cuda")
    bad = torch.zeros(1, dtype=torch.int32, device="cuda")
    _check(lib().zkdl_relu_packed(_ptr(X), _ptr(Z), _ptr(sign), _ptr(mag), _ptr(rem), _sz(n), _ptr(bad), _stream()))
    return Z, sign, mag, rem, bad


def relu_expand(mag_packed, rem_packed):
    n = mag_packed.shape[0]
    mag, rem = empty(32 * n, 8), empty(16 * n, 8)
    _check(lib().zkdl_relu_expand(_ptr(mag_packed), _ptr(rem_packed), _ptr(mag), _ptr(rem), _sz(n), _stream()))
    return mag, rem


# ------------------------------------------------------------------ G1
def g1_elementwise(op, a, b=None):
    out = empty(a.shape[0], 36)
    nb = 0 if b is None else b.shape[0]
    _check(lib().zkdl_g1_elementwise(op, _ptr(a), _ptr(b), _sz(nb), _ptr(out), _sz(a.shape[0]), _stream()))
    return out


def g1_affine_to_jacobian(a):
    out = empty(a.shape[0], 36)
    _check(lib().zkdl_g1_affine_to_jacobian(_ptr(a), _ptr(out), _sz(a.shape[0]), _stream()))
    return out


def g1_mul(P, x):
    out = empty(x.shape[0], 36)
    _check(lib().zkdl_g1_mul(_ptr(P), _sz(P.shape[0]), _ptr(x), _sz(x.shape[0]), _ptr(out), _stream()))
    return out


def g1_sum(a):
    out = empty(1, 36)
    _check(lib().zkdl_g1_sum(_ptr(a), _sz(a.shape[0]), _ptr(out), _stream()))
    return out


def g1_me(a, u_host):
    keep, up = _host_fr(u_host)
    k = 0 if keep is None else len(keep)
    out = empty(1, 36)
    _check(lib().zkdl_g1_me(_ptr(a), _sz(a.shape[0]), up, _sz(k), _ptr(out), _stream()))
    return out


def g1_normalize(a):
    out = empty(a.shape[0], 36)
    _check(lib().zkdl_g1_normalize(_ptr(a), _ptr(out), _sz(a.shape[0]), _stream()))
    return out


class G1Table:
    """Fixed-base window tables for a generator set (zkdl_g1_table)."""

    def __init__(self, points, full=True):
        self.handle = C.c_void_p(0)
        self.n = points.shape[0]
        _check(lib().zkdl_g1_table_create(_ptr(points), _sz(self.n), 0, int(bool(full)), C.byref(self.handle), _stream()))
        _torch().cuda.current_stream().synchronize()

    @property
    def nbytes(self):
        return int(lib().zkdl_g1_table_bytes(self.handle))

    def close(self):
        if self.handle:
            _torch().cuda.synchronize()
            lib().zkdl_g1_table_destroy(self.handle)
            self.handle = C.c_void_p(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def msm(table, scalars, m, scalars_mont):
    if scalars.shape[0] < m * table.n:
        raise DimensionError(1, "Incompatible dimensions")
    out = empty(m, 36)
    _check(lib().zkdl_msm(table.handle, _ptr(scalars), _sz(m), int(bool(scalars_mont)), _ptr(out), _stream()))
    return out


def commit(table, t):
    if t.shape[0] % table.n != 0:
        raise DimensionError(1, "Incompatible dimensions")
    out = empty(t.shape[0] // table.n, 36)
    _check(lib().zkdl_commit(table.handle, _ptr(t), _sz(t.shape[0]), _ptr(out), _stream()))
    return out


def me_open(table, t, u_host):
    keep, up = _host_fr(u_host)
    k = 0 if keep is None else len(keep)
    proof, ret = empty(3 * k + 1, 36), empty(1, 8)
    _check(lib().zkdl_me_open(table.handle, _ptr(t), _sz(t.shape[0]), up, _sz(k), _ptr(proof), _ptr(ret), _stream()))
    return proof, ret


def open_(gens, com_table, t, u_host):
    keep, up = _host_fr(u_host)
    ku = len(keep)
    khi = int(lib().zkdl_ceil_log2(C.c_uint32(com_table.n)))
    klo = max(ku - khi, 0)
    com_eval, proof, ret = empty(1, 36), empty(3 * klo + 1, 36), empty(1, 8)
    _check(lib().zkdl_open(gens.handle, com_table.handle, _ptr(t), _sz(t.shape[0]), up, _sz(ku), _ptr(com_eval), _ptr(proof), _ptr(ret), _stream()))
    return com_eval, proof, ret


FC_SUMCHECK, FC_OPENING = 1, 2                   # ZKDL_FC_* / ZKDL_RELU_* part masks (include/zkdl_b200.h)
RELU_MAG, RELU_REM, RELU_HP = 1, 2, 4


def zkfc_prove(X, W, Z, B, I, O, gens, com_table, u_bs, u_in, u_out, parts=FC_SUMCHECK | FC_OPENING, w_int=None):
    """zkFC::prove; `parts` restricts it to the sumcheck and/or the opening (only those proof segments are written)."""
    nfr, ng1 = C.c_size_t(0), C.c_size_t(0)
    lib().zkdl_zkfc_proof_sizes(_sz(B), _sz(I), _sz(O), _sz(gens.n), C.byref(nfr), C.byref(ng1))
    pfr, pg1 = empty(nfr.value, 8), empty(ng1.value, 36)
    k1, p1 = _host_fr(u_bs if len(u_bs) else None)
    k2, p2 = _host_fr(u_in)
    k3, p3 = _host_fr(u_out)
    wi = w_int.handle if w_int is not None else C.c_void_p(0)
    _check(lib().zkdl_zkfc_prove_parts(_ptr(X), _ptr(W), wi, _ptr(Z), _sz(B), _sz(I), _sz(O), gens.handle, com_table.handle, p1, p2, p3,
                                       _ptr(pfr), _ptr(pg1), C.c_uint(parts), _stream()))
    return pfr, pg1


def zkfc_segments(I, parts):
    """Row ranges of proof_fr written by `parts` (proof_g1 belongs to the opening as a whole)."""
    nip = 3 * ((I - 1).bit_length()) + 2
    out = []
    if parts & FC_SUMCHECK:
        out.append((0, nip + 1))
    if parts & FC_OPENING:
        out.append((nip + 1, nip + 2))
    return out


def zkrelu_segments(n, parts):
    """Row ranges of the zkReLU proof written by `parts`."""
    L = (n - 1).bit_length()
    a = 3 * (L + 5) + 1 + 32
    b = a + 3 * (L + 4) + 1 + 16
    out = []
    if parts & RELU_MAG:
        out.append((0, a))
    if parts & RELU_REM:
        out.append((a, b))
    if parts & RELU_HP:
        out.append((b, b + 3 * L + 2))
    return out


def zkrelu_prove(X, sign, mag_bin, rem_bin, u_z, v_z, u_r, v_r, u_rec, u_hp, v_hp):
    n = X.shape[0]
    pfr = empty(lib().zkdl_zkrelu_proof_size(n), 8)
    hs = [_host_fr(a) for a in (u_z, v_z, u_r, v_r, u_rec, u_hp, v_hp)]
    _check(lib().zkdl_zkrelu_prove(_ptr(X), _ptr(sign), _ptr(mag_bin), _ptr(rem_bin), _sz(n), *[h[1] for h in hs], _ptr(pfr), _stream()))
    return pfr


def zkrelu_prove_packed(X, sign, mag_packed, rem_packed, u_z, v_z, u_r, v_r, u_rec, u_hp, v_hp, parts=RELU_MAG | RELU_REM | RELU_HP):
    n = X.shape[0]
    pfr = empty(lib().zkdl_zkrelu_proof_size(n), 8)
    hs = [_host_fr(a) for a in (u_z, v_z, u_r, v_r, u_rec, u_hp, v_hp)]
    _check(lib().zkdl_zkrelu_prove_packed_parts(_ptr(X), _ptr(sign), _ptr(mag_packed), _ptr(rem_packed), _sz(n), *[h[1] for h in hs],
                                                _ptr(pfr), C.c_uint(parts), _stream()))
    return pfr


def scratch_reserve(nbytes):
    """Pre-sizes the scratch arenas of the current stream and of the calling thread's side streams (zkdl_scratch_reserve),
    so that no later call on them has to grow an arena (cudaMalloc + synchronisation) inside a timed region."""
    _check(lib().zkdl_scratch_reserve(_sz(nbytes), _stream()))


_scratch_gen = [0]


def scratch_release_all():
    """Frees the library's idle scratch arenas on the current device (zkdl_scratch_release_all; synchronises the device).
    CUDA graphs captured over library calls hold scratch addresses: scratch_generation() tells their owners to re-capture."""
    _scratch_gen[0] += 1
    _check(lib().zkdl_scratch_release_all())


def scratch_generation():
    return _scratch_gen[0]


def prof_enable(on=True):
    """Per-kernel profiler of the library (zkdl_prof_enable): clears the records and switches event bracketing on/off."""
    _check(lib().zkdl_prof_enable(int(bool(on))))


def prof_dump():
    """{kernel: {"launches", "ms", "bytes", "fr_mul", "fq_mul"}} since the last prof_enable (synchronises the device)."""
    lib().zkdl_prof_dump.restype = C.c_size_t
    need = lib().zkdl_prof_dump(None, _sz(0))
    buf = C.create_string_buffer(int(need) + 16)
    lib().zkdl_prof_dump(buf, _sz(len(buf)))
    out = {}
    for line in buf.value.decode().splitlines():
        f = line.rsplit(" ", 5)
        out[f[0]] = {"launches": int(float(f[1])), "ms": float(f[2]), "bytes": float(f[3]), "fr_mul": float(f[4]), "fq_mul": float(f[5])}
    return out


def random_vec(seed, n):
    import numpy as np
    out = np.zeros((n, 8), np.uint32)
    lib().zkdl_random_vec_host(C.c_uint32(seed), _sz(n), out.ctypes.data_as(C.c_void_p))
    return out
