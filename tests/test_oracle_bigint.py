"""Pins the C oracle (oracle/zkdl_oracle.c) against a Python big-int model and the protocol identities of
SURVEY.md §4.  CPU only.  The reference ships no tests; reference-generated fixtures are covered in
tests/test_golden.py."""
import numpy as np
import pytest
from oracle import oracle as orc
import bigint_model as bm

rng = np.random.default_rng(1234)


def rand_fr(n, mont=True):
    return orc.to_limbs([int.from_bytes(rng.bytes(40), "little") % bm.FR_P for _ in range(n)], 8)


def test_constants():
    g = orc.g1_generator()[0]
    assert bm.jac_limbs_to_affine(g) == (bm.GX, bm.GY)
    assert (bm.GY * bm.GY - bm.GX ** 3 - 4) % bm.FQ_P == 0
    one = orc.fr_mont(orc.to_limbs([1]))
    assert orc.from_limbs(one)[0] == bm.FR_R
    # reference ONE constant bls12-381.cu:3
    assert list(one[0]) == [4294967294, 1, 215042, 1485092858, 3971764213, 2576109551, 2898593135, 405057881]
    assert orc.ceil_log2(0) == 0 and orc.ceil_log2(1) == 0 and orc.ceil_log2(2) == 1 and orc.ceil_log2(1000) == 10 and orc.ceil_log2(1024) == 10


def test_fr_arith():
    a, b = rand_fr(200), rand_fr(200)
    ai, bi = orc.from_limbs(a), orc.from_limbs(b)
    assert orc.from_limbs(orc.fr_add(a, b)) == [(x + y) % bm.FR_P for x, y in zip(ai, bi)]
    assert orc.from_limbs(orc.fr_sub(a, b)) == [(x - y) % bm.FR_P for x, y in zip(ai, bi)]
    assert orc.from_limbs(orc.fr_mul(a, b)) == [(x * y * bm.FR_RINV) % bm.FR_P for x, y in zip(ai, bi)]
    assert orc.from_limbs(orc.fr_mont(a)) == [(x * bm.FR_R) % bm.FR_P for x in ai]
    assert orc.from_limbs(orc.fr_unmont(a)) == [(x * bm.FR_RINV) % bm.FR_P for x in ai]
    assert orc.from_limbs(orc.fr_neg(a)) == [(-x) % bm.FR_P for x in ai]
    assert orc.from_limbs(orc.fr_sum(a))[0] == sum(ai) % bm.FR_P
    # edge values
    e = orc.to_limbs([0, 1, bm.FR_P - 1, bm.FR_P - 2, bm.FR_R])
    ei = orc.from_limbs(e)
    for x in range(5):
        bb = np.repeat(e[x:x + 1], 5, 0)
        assert orc.from_limbs(orc.fr_mul(e, bb)) == [(v * ei[x] * bm.FR_RINV) % bm.FR_P for v in ei]
        assert orc.from_limbs(orc.fr_add(e, bb)) == [(v + ei[x]) % bm.FR_P for v in ei]
        assert orc.from_limbs(orc.fr_sub(e, bb)) == [(v - ei[x]) % bm.FR_P for v in ei]


def test_fq_mul():
    a = orc.to_limbs([int.from_bytes(rng.bytes(56), "little") % bm.FQ_P for _ in range(100)], 12)
    b = orc.to_limbs([int.from_bytes(rng.bytes(56), "little") % bm.FQ_P for _ in range(100)], 12)
    assert orc.from_limbs(orc.fq_mul(a, b)) == [(x * y * bm.FQ_RINV) % bm.FQ_P for x, y in zip(orc.from_limbs(a), orc.from_limbs(b))]


def model_me(vals, us):
    """vals, us: plain field ints (not Montgomery)."""
    vals = list(vals)
    for x in us:
        nxt = []
        for g in range((len(vals) + 1) // 2):
            a0 = vals[2 * g]; a1 = vals[2 * g + 1] if 2 * g + 1 < len(vals) else 0
            nxt.append((a0 + x * (a1 - a0)) % bm.FR_P)
        vals = nxt
    return vals


@pytest.mark.parametrize("n", [1, 2, 5, 8, 13, 64])
def test_fr_me_and_partial(n):
    k = max(1, orc.ceil_log2(n))
    a, u = rand_fr(n), rand_fr(k)
    ai = orc.fr_to_ints(a); ui = orc.fr_to_ints(u)
    got = orc.fr_to_ints(orc.fr_me(a, u)[None, :])[0]
    assert got == model_me(ai, ui)[0]
    # partial_me with window 1 == plain folds
    pm = orc.fr_partial_me(a, u[:1], 1)
    assert orc.fr_to_ints(pm) == model_me(ai, ui[:1])


def test_partial_me_rows():
    B, I = 8, 6
    a, u = rand_fr(B * I), rand_fr(3)
    ai = orc.fr_to_ints(a); ui = orc.fr_to_ints(u)
    got = orc.fr_to_ints(orc.fr_partial_me(a, u, I))
    for i in range(I):
        assert got[i] == model_me([ai[b * I + i] for b in range(B)], ui)[0]


def eq_weights(us):
    w = [1]
    for j, x in enumerate(us):   # u[0] binds LSB
        w = [w[i & ((1 << j) - 1)] * ((x if (i >> j) & 1 else (1 - x))) % bm.FR_P for i in range(1 << (j + 1))]
    return w


@pytest.mark.parametrize("n", [2, 7, 16, 33])
def test_ip_sumcheck_identities(n):
    k = orc.ceil_log2(n)
    a, b, u = rand_fr(n), rand_fr(n), rand_fr(k)
    pr = orc.fr_to_ints(orc.ip_sumcheck(a, b, u))
    ai, bi, ui = orc.fr_to_ints(a), orc.fr_to_ints(b), orc.fr_to_ints(u)
    claim = sum(x * y for x, y in zip(ai, bi)) % bm.FR_P
    for j in range(k):
        c0, c1, c2 = pr[3 * j: 3 * j + 3]
        assert (2 * c0 + c1 + c2) % bm.FR_P == claim
        claim = (c0 + c1 * ui[j] + c2 * ui[j] * ui[j]) % bm.FR_P
    assert claim == pr[-2] * pr[-1] % bm.FR_P
    assert pr[-2] == model_me(ai, ui)[0] and pr[-1] == model_me(bi, ui)[0]


@pytest.mark.parametrize("n", [2, 6, 16, 21])
def test_hp_and_bin_sumcheck_identities(n):
    k = orc.ceil_log2(n)
    a, b, u, v = rand_fr(n), rand_fr(n), rand_fr(k), rand_fr(k)
    ai, bi, ui, vi = orc.fr_to_ints(a), orc.fr_to_ints(b), orc.fr_to_ints(u), orc.fr_to_ints(v)
    pad = (1 << k) - n
    w = eq_weights(ui)
    pr = orc.fr_to_ints(orc.hp_sumcheck(a, b, u, v))
    claim = sum(w[i] * x * y for i, (x, y) in enumerate(zip(ai + [0] * pad, bi + [0] * pad))) % bm.FR_P
    for j in range(k):
        c0, c1, c2 = pr[3 * j: 3 * j + 3]
        g0, g1 = c0, (c0 + c1 + c2) % bm.FR_P
        assert ((1 - ui[j]) * g0 + ui[j] * g1) % bm.FR_P == claim
        gv = (c0 + c1 * vi[j] + c2 * vi[j] ** 2) % bm.FR_P
        claim = gv  # next claim is the eq-weighted sum at v_j without the (1-u)/(u) factor for later rounds
        # claim_j+1 = sum_g eq(u[j+1:],g) a'(g) b'(g): equals g_j(v_j) by construction
    assert claim == pr[-2] * pr[-1] % bm.FR_P
    # binary sumcheck: claim is sum eq * a(a-1)
    prb = orc.fr_to_ints(orc.bin_sumcheck(a, u, v))
    claim = sum(w[i] * x * (x - 1) for i, x in enumerate(ai + [0] * pad)) % bm.FR_P
    for j in range(k):
        c0, c1, c2 = prb[3 * j: 3 * j + 3]
        assert ((1 - ui[j]) * c0 + ui[j] * (c0 + c1 + c2)) % bm.FR_P == claim
        claim = (c0 + c1 * vi[j] + c2 * vi[j] ** 2) % bm.FR_P
    assert claim == prb[-1] * (prb[-1] - 1) % bm.FR_P


def test_random_vec_mt19937():
    # std::mt19937 default-seeded known answer: 10000th draw of mt19937(5489) is 4123659995
    import random
    v = orc.random_vec(5489, 1250)          # 1250*8 = 10000 draws
    assert int(v[-1, 7]) == 4123659995 % 1944954707
    assert int(orc.random_vec(5489, 1)[0, 0]) == 3499211612
    assert all(x < bm.FR_P for x in orc.from_limbs(orc.random_vec(7, 64)))


def test_g1_group_law_vs_affine_model():
    G = orc.g1_generator()
    Gp = (bm.GX, bm.GY)
    ks = [1, 2, 3, 5, 0xdeadbeef, bm.FR_P - 1, bm.FR_P, 0]
    P = orc.g1_mul(G, orc.to_limbs(ks))
    for row, k in zip(P, ks):
        assert bm.jac_limbs_to_affine(row) == bm.aff_mul(Gp, k)
    assert orc.g1_on_curve(P).all()
    Pf = orc.g1_mul(G, orc.to_limbs(ks), fast=True)
    assert orc.g1_eq(P, Pf).all()
    # add / double / mixed / neg incl. edge cases P+P, P+(-P), inf
    A, B = P[[0, 1, 2, 3, 4, 5, 7, 0]], P[[0, 3, 4, 5, 4, 0, 2, 7]]
    S = orc.g1_add(A, B)
    for a, b, s in zip(A, B, S):
        assert bm.jac_limbs_to_affine(s) == bm.aff_add(bm.jac_limbs_to_affine(a), bm.jac_limbs_to_affine(b))
    S = orc.g1_add(A, orc.g1_neg(A))
    assert all(bm.jac_limbs_to_affine(s) is None for s in S)
    aff, inf = orc.g1_to_affine(P)
    for row, a, i in zip(P, aff, inf):
        m = bm.jac_limbs_to_affine(row)
        if m is None: assert i
        else:
            assert orc.from_limbs(a[0:12])[0] * bm.FQ_RINV % bm.FQ_P == m[0]
            assert orc.from_limbs(a[12:24])[0] * bm.FQ_RINV % bm.FQ_P == m[1]
    M = orc.g1_add_mixed(P[[1, 7, 2]], aff[[2, 2, 2]])
    assert bm.jac_limbs_to_affine(M[0]) == bm.aff_mul(Gp, 5)
    assert bm.jac_limbs_to_affine(M[1]) == bm.aff_mul(Gp, 3)
    assert bm.jac_limbs_to_affine(M[2]) == bm.aff_mul(Gp, 6)


def test_g1_sum_me_commit_open_identities():
    n = 8
    G0 = orc.g1_generator()
    ks = [int.from_bytes(rng.bytes(31), "little") for _ in range(n)]
    G = orc.g1_mul(G0, orc.to_limbs(ks), fast=True)
    Gp = (bm.GX, bm.GY)
    s = orc.g1_sum(G)
    assert bm.jac_limbs_to_affine(s[0]) == bm.aff_mul(Gp, sum(ks) % bm.FR_P)
    # G1_me: P' = P0 + [unmont(x)](P1 - P0)
    u = rand_fr(3); ui = orc.fr_to_ints(u)
    w = eq_weights(ui)
    me = orc.g1_me(G, u)
    assert bm.jac_limbs_to_affine(me[0]) == bm.aff_mul(Gp, sum(wi * k for wi, k in zip(w, ks)) % bm.FR_P)
    # commit (intended): com[r] = sum_c t[r,c] G[c]
    t_int = [int(x) for x in rng.integers(-2000, 2000, size=2 * n)]
    t = orc.fr_from_ints(t_int, mont=True)
    com = orc.commit(G, t)
    for r in range(2):
        exp = sum(t_int[r * n + c] * ks[c] for c in range(n)) % bm.FR_P
        assert bm.jac_limbs_to_affine(com[r]) == bm.aff_mul(Gp, exp)
    assert orc.g1_eq(com, orc.commit(G, t, fast=True)).all()
    # me_open identities: scalars are the Montgomery limbs read as integers
    tt = rand_fr(n); uu = rand_fr(3)
    proof, ret = orc.me_open(tt, G, uu)
    proof_f, ret_f = orc.me_open(tt, G, uu, fast=True)
    assert orc.g1_eq(proof, proof_f).all() and (ret == ret_f).all()
    s_int = orc.from_limbs(tt)            # raw limbs (== s*R mod r as integers)
    x = orc.fr_to_ints(uu)
    T = bm.jac_limbs_to_affine(proof[0])
    assert T == bm.aff_mul(Gp, sum(si * k for si, k in zip(s_int, ks)) % bm.FR_P)
    cur_s, cur_k = s_int, ks
    for j in range(3):
        T, T0, T1 = [bm.jac_limbs_to_affine(p) for p in proof[3 * j:3 * j + 3]]
        ns = len(cur_s) // 2
        assert T == bm.aff_mul(Gp, sum(a * b for a, b in zip(cur_s, cur_k)) % bm.FR_P)
        assert T0 == bm.aff_mul(Gp, sum(cur_s[2 * g] * cur_k[2 * g + 1] for g in range(ns)) % bm.FR_P)
        assert T1 == bm.aff_mul(Gp, sum(cur_s[2 * g + 1] * cur_k[2 * g] for g in range(ns)) % bm.FR_P)
        # fold: s' = s0 + u (s1 - s0) on Montgomery values => raw limbs fold linearly too
        cur_s = [(cur_s[2 * g] + x[j] * (cur_s[2 * g + 1] - cur_s[2 * g])) % bm.FR_P for g in range(ns)]
        cur_k = [(cur_k[2 * g + 1] + x[j] * (cur_k[2 * g] - cur_k[2 * g + 1])) % bm.FR_P for g in range(ns)]
    assert bm.jac_limbs_to_affine(proof[9]) == bm.aff_mul(Gp, cur_k[0])
    assert orc.from_limbs(ret)[0] == cur_s[0]


def test_quantise_relu_matmul():
    fs = np.array([[0.5, -0.25, 1e-6, -1e-6, 3.0000076, -0.0], [1.5, 2.5 / 65536, -2.5 / 65536, 0.49999 / 65536, 100.0, -100.0]], np.float32)
    q = orc.float_to_fr(fs, 4, 8)
    qi = orc.from_limbs(q)
    def exp(x):
        x = np.float32(x) * np.float32(65536.0)
        v = int(np.floor(abs(float(x)) + 0.5))
        return (-v) % bm.FR_P if np.signbit(x) else v
    for r in range(4):
        for c in range(8):
            e = exp(fs[r, c]) if (r < 2 and c < 6) else 0
            assert qi[r * 8 + c] == e
    # relu: model
    xs = [0, 1, 65535, 65536, 98304, 32768, 32767, (1 << 47) - 1, -1, -65536, -(1 << 47), -32768, -32769, 12345678901]
    X = orc.fr_from_ints(xs, mont=True)
    Z, sign, mag, rem, bad = orc.relu(X)
    assert bad == 0
    zi = orc.fr_to_ints(Z); sg = orc.fr_to_ints(sign); mb = orc.fr_to_ints(mag); rb = orc.fr_to_ints(rem)
    for i, x in enumerate(xs):
        m = x if x >= 0 else (1 << 47) + x
        rs = (m >> 15) & 1; rm = m & 32767
        r = rm - 32768 if rs else rm
        qv = ((m - r) >> 16) & 0xFFFFFFFF
        assert sg[i] == (1 if x >= 0 else 0)
        assert zi[i] == (qv if x >= 0 else 0)
        assert [mb[32 * i + k] for k in range(32)] == [(qv >> k) & 1 for k in range(32)]
        assert [rb[16 * i + k] for k in range(15)] == [(rm >> k) & 1 for k in range(15)] and rb[16 * i + 15] == rs
        if x >= 0: assert qv == (x + 32768) >> 16       # round-half-up rescale
    _, _, _, _, bad = orc.relu(orc.fr_from_ints([1 << 47, -(1 << 47) - 1], mont=True))
    assert bad == 2
    A = [int(v) for v in rng.integers(-50, 50, size=6)]; Bm = [int(v) for v in rng.integers(-50, 50, size=12)]
    Cm = orc.fr_to_ints(orc.fr_matmul(orc.fr_from_ints(A), orc.fr_from_ints(Bm), 2, 3, 4))
    for r in range(2):
        for c in range(4):
            assert Cm[r * 4 + c] == sum(A[r * 3 + k] * Bm[k * 4 + c] for k in range(3)) % bm.FR_P


def test_cpu_pippenger_matches_ladder():
    n = 40
    G0 = orc.g1_generator()
    ks = [int.from_bytes(rng.bytes(31), "little") for _ in range(n)]
    G = orc.g1_mul(G0, orc.to_limbs(ks), fast=True)
    aff, inf = orc.g1_to_affine(G)
    sc = orc.to_limbs([int.from_bytes(rng.bytes(40), "little") % bm.FR_P for _ in range(n)])
    got = orc.msm_pippenger(aff, sc)
    exp = orc.g1_sum(orc.g1_mul(G, sc, fast=True))
    assert orc.g1_eq(got, exp).all()
