"""Verifier-side checks of the proof elements produced by zkFC::prove / zkReLU::prove (SURVEY.md §4 "protocol
identities", §8f rank 1).  The reference computes these proofs and throws them away; it has no verifier.  This module
checks, for proofs of ANY size and without re-running the prover:

  * every sumcheck round polynomial against the running claim and the final claim against the returned evaluations;
  * the matmul claim: round 0 of the inner-product sumcheck against Z(u_out || u_bs);
  * the opening: T_0 = [R] com(u_hi), the folding recursion T' = [x(1-x)] T + [(1-x)^2] T0 + [x^2] T1 for every round,
    the final [s_final] G_final = T_k, G_final = <w, G> for the public fold weights w, and s_final = W~(u) = the
    sumcheck's final weight evaluation.

Scalar arithmetic on the ~100 field elements of a proof is plain Python integers; every G1 operation (scalar
multiplication, addition, MSM, normalisation for equality) goes through the C ABI (zkdl_b200.capi) on the GPU.
It doubles as the full-size, size-independent parity property used by tests/test_fullsize_properties.py."""
import numpy as np

from . import capi as zk

P = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
R = (1 << 256) % P
RINV = pow(R, -1, P)


def limbs_to_int(row):
    v = 0
    for j, x in enumerate(np.asarray(row, dtype=np.uint32).reshape(-1)):
        v |= int(x) << (32 * j)
    return v


def int_to_limbs(vals):
    out = np.zeros((len(vals), 8), dtype=np.uint32)
    for i, v in enumerate(vals):
        for j in range(8):
            out[i, j] = (int(v) >> (32 * j)) & 0xFFFFFFFF
    return out


def plain(row):
    """Montgomery limbs -> the field element they represent."""
    return limbs_to_int(row) * RINV % P


class VerifyError(AssertionError):
    pass


def _req(cond, msg):
    if not cond:
        raise VerifyError(msg)


# ------------------------------------------------------------------------------------------------ sumchecks
def verify_ip(proof, u, claim):
    """inner_product_sumcheck (proof.cu:72-108): g_j(0) + g_j(1) = claim_j, claim_{j+1} = g_j(u_j), last = a(0) b(0)."""
    p = [plain(r) for r in proof]
    k = len(u)
    _req(len(p) == 3 * k + 2, "ip: wrong proof length")
    for j in range(k):
        c0, c1, c2 = p[3 * j: 3 * j + 3]
        _req((2 * c0 + c1 + c2) % P == claim % P, f"ip: round {j} does not match the running claim")
        x = plain(u[j])
        claim = (c0 + c1 * x + c2 * x * x) % P
    _req(claim == p[-2] * p[-1] % P, "ip: final claim != a(0) * b(0)")
    return p[-2], p[-1]


def verify_weighted(proof, u, v, claim, nfinal, final_check):
    """hadamard / binary sumcheck (proof.cu:110-200): (1-u_j) g_j(0) + u_j g_j(1) = claim_j, claim_{j+1} = g_j(v_j)."""
    p = [plain(r) for r in proof]
    k = len(u)
    _req(len(p) == 3 * k + nfinal, "sumcheck: wrong proof length")
    for j in range(k):
        c0, c1, c2 = p[3 * j: 3 * j + 3]
        uj, vj = plain(u[j]), plain(v[j])
        lhs = ((1 - uj) * c0 + uj * (c0 + c1 + c2)) % P
        if claim is None:
            claim = lhs                              # the reference leaves the initial Hadamard claim implicit
        _req(lhs == claim % P, f"sumcheck: round {j} does not match the running claim")
        claim = (c0 + c1 * vj + c2 * vj * vj) % P
    _req(final_check(claim, p[3 * k:]), "sumcheck: final claim does not match the returned evaluations")
    return p[3 * k:]


def verify_bin(proof, u, v):
    return verify_weighted(proof, u, v, 0, 1, lambda c, f: c == f[0] * (f[0] - 1) % P)


def verify_hp(proof, u, v):
    return verify_weighted(proof, u, v, None, 2, lambda c, f: c == f[0] * f[1] % P)


# ------------------------------------------------------------------------------------------------ G1 helpers (GPU)
def _same_point(a, b):
    na, nb = zk.to_host(zk.g1_normalize(a)), zk.to_host(zk.g1_normalize(b))
    return bool(np.array_equal(na, nb))


def _lincomb(points, scalars):
    """sum_i [scalars_i] points_i for a handful of Jacobian points (device [m,36]) and python ints."""
    sc = zk.to_device(int_to_limbs([s % P for s in scalars]))
    return zk.g1_sum(zk.g1_mul(points, sc))


def verify_opening(G, com_eval, open_proof, open_ret, u_lo, gens_table=None):
    """Commitment::me_open transcript (commitment.cu:43-81) against com(u_hi).  G: device [n,36] generators."""
    k = len(u_lo)
    n = G.shape[0]
    _req(n == 1 << k and open_proof.shape[0] == 3 * k + 1, "opening: wrong sizes")
    T_cur = _lincomb(com_eval, [R])                                    # scalars are Montgomery limbs: x(s) = s * R
    for j in range(k):
        T, T0, T1 = open_proof[3 * j: 3 * j + 1], open_proof[3 * j + 1: 3 * j + 2], open_proof[3 * j + 2: 3 * j + 3]
        _req(_same_point(T, T_cur), f"opening: T of round {j} does not match the running commitment")
        x = plain(u_lo[j])
        T_cur = _lincomb(open_proof[3 * j: 3 * j + 3], [x * (1 - x), (1 - x) * (1 - x), x * x])
    G_final = open_proof[3 * k: 3 * k + 1]
    s_final = limbs_to_int(zk.to_host(open_ret)[0])                    # raw Montgomery limbs as the scalar (defect B4)
    _req(_same_point(_lincomb(G_final, [s_final]), T_cur), "opening: [s_final] G_final != folded commitment")
    # G_final = <w, G>, w_b = prod_j (bit_j(b) ? 1 - x_j : x_j)   (G' = G1 + [x](G0 - G1))
    w = [1]
    for j in range(k):
        x = plain(u_lo[j])
        w = [wb * x % P for wb in w] + [wb * (1 - x) % P for wb in w]
    tab = gens_table or zk.G1Table(G, full=False)
    exp = zk.msm(tab, zk.to_device(int_to_limbs(w)), 1, False)
    if gens_table is None:
        tab.close()
    _req(_same_point(G_final, exp), "opening: G_final is not the folded generator")
    return s_final * RINV % P                                          # the opened evaluation W~(u)


def verify_subgroup(points, what="point"):
    """BLS12-381 G1 has cofactor 0x396c8c005555e1568c00aaab0000aaab: an on-curve point need not lie in the prime-order subgroup
    the protocol lives in.  [r] P must be the point at infinity for every generator, commitment and proof element."""
    if points is None or points.shape[0] == 0:
        return
    r = zk.to_device(np.repeat(int_to_limbs([P]), points.shape[0], axis=0))
    z = zk.to_host(zk.g1_mul(points, r))[:, 24:]
    _req(not z.any(), f"{what}: not in the prime-order subgroup")


def verify_commitment_eval(com, com_eval, u_hi):
    """com(u_hi) (proof_g1[0]) against the PUBLIC row commitments: sum_r eq(u_hi, r) com[r]  (g1-tensor.cu:463-491)."""
    w = [1]
    for row in np.asarray(u_hi).reshape(-1, 8):
        x = plain(row)
        w = [wb * (1 - x) % P for wb in w] + [wb * x % P for wb in w]
    _req(com.shape[0] <= len(w), "commitment: more rows than the challenge addresses")
    tab = zk.G1Table(com, full=False)
    exp = zk.msm(tab, zk.to_device(int_to_limbs(w[: com.shape[0]])), 1, False)
    tab.close()
    _req(_same_point(com_eval, exp), "commitment: com(u_hi) is not the evaluation of the public commitment")


def verify_zkfc(proof_fr, proof_g1, G, B, I, O, u_bs, u_in, u_out, gens_table=None):
    """zkFC::prove (zkfc.cu:128-145): proof_fr = [ip][Z(u)][open_ret], proof_g1 = [com(u_hi)][me_open ...]."""
    fr = zk.to_host(proof_fr)
    ki = len(u_in)
    nip = 3 * ki + 2
    z_eval = plain(fr[nip])
    a0, b0 = verify_ip(fr[:nip], u_in, z_eval)                         # sum_i Xr_i Wr_i = Z(u_out || u_bs)
    u = np.concatenate([np.asarray(u_out).reshape(-1, 8), np.asarray(u_in).reshape(-1, 8)])
    klo = (G.shape[0] - 1).bit_length()
    w_eval = verify_opening(G, proof_g1[:1], proof_g1[1:], proof_fr[nip + 1: nip + 2], u[:klo], gens_table)
    _req(w_eval == b0, "zkFC: opened W~(u_out || u_in) != the sumcheck's final weight evaluation")
    return {"z_eval": z_eval, "x_eval": a0, "w_eval": b0}


def verify_zkrelu(proof_fr, n, u_z, v_z, u_r, v_r, u_hp, v_hp):
    """zkReLU::prove (zkrelu.cu:79-100): two binary sumchecks (claim 0) and the Hadamard sumcheck chain.  The 32 + 16
    partial_me(u_recover, .) rows are evaluations at a point the binary sumchecks never visit, and the fragments carry neither
    X~(u_recover) nor sign~(u_recover): the recover relation cannot be checked from the reference's proof elements, and the
    Hadamard sumcheck's initial claim is implicit (accepted as the first round states it)."""
    fr = zk.to_host(proof_fr)
    L = (n - 1).bit_length()
    o = 0
    verify_bin(fr[o: o + 3 * (L + 5) + 1], u_z, v_z); o += 3 * (L + 5) + 1 + 32
    verify_bin(fr[o: o + 3 * (L + 4) + 1], u_r, v_r); o += 3 * (L + 4) + 1 + 16
    verify_hp(fr[o: o + 3 * L + 2], u_hp, v_hp)
    return True
