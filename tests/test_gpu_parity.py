"""GPU parity: every C-ABI entry point of libzkdl_b200.so against the CPU oracle on the same seeded inputs.
Fr results are compared bit-for-bit; G1 results as points (projective equality), SURVEY.md §8c."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import oracle as orc


@pytest.fixture(scope="module")
def zk():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from zkdl_b200 import capi
    capi.lib()
    return capi


rng = np.random.default_rng(2024)


def rand_fr(n):
    out = rng.integers(0, 1 << 32, size=(n, 8), dtype=np.uint64).astype(np.uint32)
    out[:, 7] %= 1944954707
    return out


def small_fr(n, lo=-(1 << 15), hi=1 << 15):
    return orc.fr_from_ints([int(v) for v in rng.integers(lo, hi, size=n)], mont=True)


def eq(a, b):
    return np.array_equal(np.asarray(a, dtype=np.uint32).reshape(-1), np.asarray(b, dtype=np.uint32).reshape(-1))


def test_fr_elementwise_and_sum(zk):
    n = 5000
    a, b = rand_fr(n), rand_fr(n)
    a[:4] = orc.to_limbs([0, 1, orc.FR_P - 1, orc.FR_R]); b[:4] = orc.to_limbs([orc.FR_P - 1, orc.FR_P - 1, orc.FR_P - 1, 0])
    da, db = zk.to_device(a), zk.to_device(b)
    assert eq(zk.to_host(zk.fr_elementwise(zk.OP_ADD, da, db)), orc.fr_add(a, b))
    assert eq(zk.to_host(zk.fr_elementwise(zk.OP_SUB, da, db)), orc.fr_sub(a, b))
    assert eq(zk.to_host(zk.fr_elementwise(zk.OP_MUL, da, db)), orc.fr_mul(a, b))
    assert eq(zk.to_host(zk.fr_elementwise(zk.OP_NEG, da)), orc.fr_neg(a))
    assert eq(zk.to_host(zk.fr_elementwise(zk.OP_MONT, da)), orc.fr_mont(a))
    assert eq(zk.to_host(zk.fr_elementwise(zk.OP_UNMONT, da)), orc.fr_unmont(a))
    x = rand_fr(1)
    for op, name in ((zk.OP_ADD, "add"), (zk.OP_SUB, "sub"), (zk.OP_MUL, "mul")):
        assert eq(zk.to_host(zk.fr_broadcast(op, da, x)), orc.fr_bcast(a, x, name))
    assert eq(zk.to_host(zk.fr_sum(da)), orc.fr_sum(a))
    assert eq(zk.to_host(zk.fr_sum(da[:1])), a[0])


@pytest.mark.parametrize("n", [1, 2, 3, 17, 256, 1000, 4096, 5001, 70000])
def test_fold_me_partial(zk, n):
    a = rand_fr(n); da = zk.to_device(a)
    k = max(orc.ceil_log2(n), 0)
    u = rand_fr(max(k, 1))
    assert eq(zk.to_host(zk.fr_fold(da, u[:1])), orc.fr_partial_me(a, u[:1], 1))
    if n > 1:
        assert eq(zk.to_host(zk.fr_me(da, u[:k])), orc.fr_me(a, u[:k]))
    for w in (1, 3, 16):
        for kk in (1, 2, 3, 5):
            if n > w * (1 << (kk - 1)):
                got = zk.to_host(zk.fr_partial_me(da, u[:kk] if kk <= len(u) else rand_fr(kk), w)) if kk <= len(u) else None
                if got is not None:
                    assert eq(got, orc.fr_partial_me(a, u[:kk], w)), (n, w, kk)
        assert eq(zk.to_host(zk.fr_partial_fold(da, u[:1], w)), orc.fr_partial_me(a, u[:1], w))


def test_me_dimension_errors(zk):
    da = zk.to_device(rand_fr(8))
    with pytest.raises(zk.DimensionError):
        zk.fr_me(da, rand_fr(2))
    with pytest.raises(zk.DimensionError):
        zk.fr_me(da, rand_fr(4))
    with pytest.raises(zk.DimensionError):
        zk.ip_sumcheck(da, da, rand_fr(2))
    with pytest.raises(zk.DimensionError):
        zk.fr_partial_me(da, rand_fr(2), 4)


@pytest.mark.parametrize("n", [2, 3, 8, 100, 2048, 2049, 4096, 6000, 40000])
def test_sumchecks(zk, n):
    k = orc.ceil_log2(n)
    a, b, u, v = rand_fr(n), rand_fr(n), rand_fr(k), rand_fr(k)
    da, db = zk.to_device(a), zk.to_device(b)
    assert eq(zk.to_host(zk.ip_sumcheck(da, db, u)), orc.ip_sumcheck(a, b, u))
    assert eq(zk.to_host(zk.hp_sumcheck(da, db, u, v)), orc.hp_sumcheck(a, b, u, v))
    assert eq(zk.to_host(zk.bin_sumcheck(da, u, v)), orc.bin_sumcheck(a, u, v))


def test_sumcheck_binary_table(zk):
    n = 1 << 13
    bits = rng.integers(0, 2, size=n)
    a = orc.fr_from_ints([int(x) for x in bits], mont=True)
    k = 13
    u, v = rand_fr(k), rand_fr(k)
    got = zk.to_host(zk.bin_sumcheck(zk.to_device(a), u, v))
    assert eq(got, orc.bin_sumcheck(a, u, v))
    assert not got[0].any()          # C0 of a 0/1 table is identically zero


def test_quantise_matmul_relu(zk):
    import torch
    fs = rng.standard_normal((37, 53)).astype(np.float32)
    fs[0, :6] = [0.0, -0.0, 1e-9, -1e-9, 2.5 / 65536, -2.5 / 65536]
    fs[1, :4] = [1e9, -1e9, np.inf, np.nan]
    got = zk.to_host(zk.float_to_fr(torch.from_numpy(fs).cuda(), 64, 64))
    assert eq(got, orc.float_to_fr(fs, 64, 64))
    A, B = small_fr(24 * 40), small_fr(40 * 33)
    got = zk.to_host(zk.fr_matmul(zk.to_device(A), zk.to_device(B), 24, 40, 33))
    assert eq(got, orc.fr_matmul(A, B, 24, 40, 33))
    # operands outside the small-integer range take the generic Fr path; mixed and boundary cases
    A2, B2 = rand_fr(5 * 7), rand_fr(7 * 3)
    assert eq(zk.to_host(zk.fr_matmul(zk.to_device(A2), zk.to_device(B2), 5, 7, 3)), orc.fr_matmul(A2, B2, 5, 7, 3))
    lim = [(1 << 31) - 1, -(1 << 31), 1 << 31, -(1 << 31) - 1, 0, 1, -1]
    A3 = orc.fr_from_ints([lim[i % 7] for i in range(70 * 65)]); B3 = orc.fr_from_ints([lim[(3 * i + 1) % 5] for i in range(65 * 66)])
    assert eq(zk.to_host(zk.fr_matmul(zk.to_device(A3), zk.to_device(B3), 70, 65, 66)), orc.fr_matmul(A3, B3, 70, 65, 66))
    A4 = orc.fr_from_ints([lim[i % 2] for i in range(130 * 100)]); B4 = orc.fr_from_ints([lim[(i + 1) % 2] for i in range(100 * 67)])
    assert eq(zk.to_host(zk.fr_matmul(zk.to_device(A4), zk.to_device(B4), 130, 100, 67)), orc.fr_matmul(A4, B4, 130, 100, 67))
    xs = [int(v) for v in rng.integers(-(1 << 46), 1 << 46, size=3000)] + [0, 1, -1, 32767, 32768, -32768, -32769, (1 << 47) - 1, -(1 << 47), 1 << 47, -(1 << 47) - 1]
    X = orc.fr_from_ints(xs, mont=True)
    Z, sign, mag, rem, bad = zk.relu(zk.to_device(X))
    oZ, osign, omag, orem, obad = orc.relu(X)
    assert int(bad.item()) == obad == 2
    assert eq(zk.to_host(Z), oZ) and eq(zk.to_host(sign), osign) and eq(zk.to_host(mag), omag) and eq(zk.to_host(rem), orem)


@pytest.mark.parametrize("case", ["tc_small", "tc_boundary", "a_too_big", "w_too_big", "generic_fr", "ragged_shape", "tc_long_k",
                                  "umma_boundary", "umma_a_too_big", "umma_tiles"])
def test_matmul_prepared_routes(zk, case):
    """zkdl_fr_matmul_prepared: the int8 tensor-core route, the int32 SIMT route and the generic Fr route agree with the
    oracle's Fr matmul bit for bit; the route is picked on the device from the operands' magnitudes."""
    M, K, N = 64, 128, 64
    if case == "ragged_shape":
        M, K, N = 24, 40, 33
    if case == "tc_long_k":
        M, K, N = 128, 2048, 128
    if case.startswith("umma"):                       # shapes the tcgen05 kernel takes (M % 128 == 0, N % 64 == 0, K % 128 == 0)
        M, K, N = (256, 384, 192) if case == "umma_tiles" else (128, 256, 192)
    amax, wmax = (1 << 23) - 1, (1 << 15) - 1
    if case in ("tc_small", "ragged_shape", "tc_long_k", "umma_tiles"):
        a = [int(v) for v in rng.integers(-(1 << 20), 1 << 20, size=M * K)]
        w = [int(v) for v in rng.integers(-(1 << 12), 1 << 12, size=K * N)]
    elif case == "generic_fr":
        a = w = None
    else:
        a = [(amax, -amax, 0, 1, -1, 255, -256, 65535, -65536)[i % 9] for i in range(M * K)]
        w = [(wmax, -wmax, 0, 1, -1, 255, -256, 128, -129)[(5 * i + 2) % 9] for i in range(K * N)]
        if case in ("a_too_big", "umma_a_too_big"):
            a[7] = 1 << 23
        if case == "w_too_big":
            w[11] = -(1 << 15)
    A = rand_fr(M * K) if a is None else orc.fr_from_ints(a, mont=True)
    W = rand_fr(K * N) if w is None else orc.fr_from_ints(w, mont=True)
    dW = zk.to_device(W)
    prep = zk.MatmulWeights(dW, K, N)
    got = zk.to_host(zk.fr_matmul_prepared(zk.to_device(A), prep, M))
    assert eq(got, orc.fr_matmul(A, W, M, K, N))
    assert eq(zk.to_host(zk.fr_matmul(zk.to_device(A), dW, M, K, N)), got)
    prep.close()


def make_points(n, seed=5):
    r = np.random.default_rng(seed)
    ks = orc.to_limbs([int.from_bytes(r.bytes(31), "little") for _ in range(n)])
    return orc.g1_mul(orc.g1_generator(), ks, fast=True)


def test_g1_elementwise_mul_sum(zk):
    n = 70
    P, Q = make_points(n, 1), make_points(n, 2)
    Q[0] = P[0]                               # P + P
    Q[1] = orc.g1_neg(P[1:2])[0]              # P + (-P)
    Q[2, 24:] = 0                             # Q = infinity
    P[3, 24:] = 0                             # P = infinity
    dP, dQ = zk.to_device(P), zk.to_device(Q)
    assert orc.g1_eq(zk.to_host(zk.g1_elementwise(zk.G1_ADD, dP, dQ)), orc.g1_add(P, Q)).all()
    assert orc.g1_eq(zk.to_host(zk.g1_elementwise(zk.G1_SUB, dP, dQ)), orc.g1_add(P, orc.g1_neg(Q))).all()
    assert orc.g1_eq(zk.to_host(zk.g1_elementwise(zk.G1_NEG, dP)), orc.g1_neg(P)).all()
    assert orc.g1_eq(zk.to_host(zk.g1_elementwise(zk.G1_ADD, dP, dQ[5:6])), orc.g1_add(P, np.repeat(Q[5:6], n, 0))).all()
    aff, inf = orc.g1_to_affine(Q)
    ok = ~inf
    got = zk.to_host(zk.g1_elementwise(zk.G1_MADD, zk.to_device(P[ok]), zk.to_device(aff[ok])))
    assert orc.g1_eq(got, orc.g1_add(P[ok], Q[ok])).all()
    x = rand_fr(n); x[0] = 0; x[1] = orc.to_limbs([1])[0]; x[2] = orc.to_limbs([orc.FR_P - 1])[0]; x[4] = 0xFFFFFFFF
    got = zk.to_host(zk.g1_mul(dP, zk.to_device(x)))
    assert orc.g1_eq(got, orc.g1_mul(P, x, fast=True)).all()
    x2 = rand_fr(2 * n)
    assert orc.g1_eq(zk.to_host(zk.g1_mul(dP, zk.to_device(x2))), orc.g1_mul(P, x2, fast=True)).all()   # broadcast of the G1 side
    for m in (1, 2, 63, 64, 70):
        assert orc.g1_eq(zk.to_host(zk.g1_sum(dP[:m])), orc.g1_sum(P[:m])).all()
    nz = zk.to_host(zk.g1_normalize(dP))
    assert orc.g1_eq(nz, P).all()
    one = orc.g1_generator()[0, 24:]
    assert all((row[24:] == one).all() or not row[24:].any() for row in nz)


@pytest.mark.parametrize("n,full", [(8, True), (64, True), (100, True), (64, False), (1000, False)])
def test_msm_against_ladder(zk, n, full):
    G = make_points(n, 7)
    if n >= 64:
        G[5] = G[4]                          # repeated base
        G[6, 24:] = 0                        # infinity base
    tab = zk.G1Table(zk.to_device(G), full=full)
    m = 3
    s = rand_fr(m * n)
    s[0] = 0; s[1] = orc.to_limbs([1])[0]; s[2] = orc.to_limbs([orc.FR_P - 1])[0]; s[3] = orc.to_limbs([(orc.FR_P - 1) // 2])[0]
    s[4] = orc.to_limbs([(orc.FR_P + 1) // 2])[0]; s[5] = s[4]
    got = zk.to_host(zk.msm(tab, zk.to_device(s), m, False))
    exp = np.concatenate([orc.g1_sum(orc.g1_mul(G, s[r * n:(r + 1) * n], fast=True)) for r in range(m)])
    assert orc.g1_eq(got, exp).all()
    # Montgomery scalars == commit semantics (intended row-wise Pedersen commitment)
    t = small_fr(m * n)
    got = zk.to_host(zk.commit(tab, zk.to_device(t)))
    assert orc.g1_eq(got, orc.commit(G, t, fast=True)).all()
    t2 = rand_fr(n)
    assert orc.g1_eq(zk.to_host(zk.commit(tab, zk.to_device(t2))), orc.commit(G, t2, fast=True)).all()
    tab.close()


def test_msm_all_equal_bases_and_scalars(zk):
    n = 256
    G = np.repeat(orc.g1_generator(), n, 0)           # Commitment(n, G1Jacobian_generator) before randomisation (demo.cu:81)
    tab = zk.G1Table(zk.to_device(G), full=True)
    s = np.repeat(orc.to_limbs([12345]), n, 0)
    got = zk.to_host(zk.msm(tab, zk.to_device(s), 1, False))
    exp = orc.g1_mul(orc.g1_generator(), orc.to_limbs([12345 * n]), fast=True)
    assert orc.g1_eq(got, exp).all()
    tab.close()


@pytest.mark.parametrize("n", [1, 2, 8, 64])
def test_me_open_and_g1_me(zk, n):
    k = orc.ceil_log2(n)
    G = make_points(n, 11)
    t, u = rand_fr(n), rand_fr(max(k, 1))[:k]
    tab = zk.G1Table(zk.to_device(G), full=True)
    proof, ret = zk.me_open(tab, zk.to_device(t), u if k else None)
    oproof, oret = orc.me_open(t, G, u, fast=True)
    assert orc.g1_eq(zk.to_host(proof), oproof).all()
    assert eq(zk.to_host(ret), oret)
    if n > 1:
        got = zk.to_host(zk.g1_me(zk.to_device(G), u))
        assert orc.g1_eq(got, orc.g1_me(G, u)).all()
    tab.close()


def test_g1_me_non_pow2(zk):
    G = make_points(11, 3)
    u = rand_fr(4)
    assert orc.g1_eq(zk.to_host(zk.g1_me(zk.to_device(G), u)), orc.g1_me(G, u)).all()


@pytest.mark.parametrize("B,I,O", [(4, 8, 16), (1, 16, 8), (8, 32, 32)])
def test_zkfc_prove_and_open(zk, B, I, O):
    kb, ki, ko = (orc.ceil_log2(v) for v in (B, I, O))
    ng = 1 << ((orc.ceil_log2(I * O) + 1) // 2)           # demo.cu:81 rule on padded sizes
    G = make_points(ng, 21)
    X, W = small_fr(B * I), small_fr(I * O)
    Z = orc.fr_matmul(X, W, B, I, O)
    gens = zk.G1Table(zk.to_device(G), full=True)
    dW = zk.to_device(W)
    com = zk.commit(gens, dW)
    ocom = orc.commit(G, W, fast=True)
    assert orc.g1_eq(zk.to_host(com), ocom).all()
    com_tab = zk.G1Table(com, full=True)
    u_bs, u_in, u_out = rand_fr(max(kb, 1))[:kb], rand_fr(ki), rand_fr(ko)
    pfr, pg1 = zk.zkfc_prove(zk.to_device(X), dW, zk.to_device(Z), B, I, O, gens, com_tab, u_bs, u_in, u_out)
    ref = orc.zkfc_prove(X, W, Z, G, ocom, B, I, O, u_bs, u_in, u_out, fast=True)
    pfr, pg1 = zk.to_host(pfr), zk.to_host(pg1)
    nip = 3 * ki + 2
    assert eq(pfr[:nip], ref["ip"])
    assert eq(pfr[nip], ref["z_eval"])
    assert eq(pfr[nip + 1], ref["open_ret"])
    assert orc.g1_eq(pg1[:1], ref["com_eval"]).all()
    assert orc.g1_eq(pg1[1:], ref["opening"]).all()
    # sumcheck self-consistency: 2 c0 + c1 + c2 of round 0 equals Z(u_out || u_bs)  
    c = orc.fr_to_ints(pfr[:3])
    assert (2 * c[0] + c[1] + c[2]) % orc.FR_P == orc.fr_to_ints(pfr[nip:nip + 1])[0]
    # with the integer copy of the weights both weights.partial_me passes fold integers against eq tables: same proof
    prep = zk.MatmulWeights(dW, I, O)
    pfr2, pg12 = zk.zkfc_prove(zk.to_device(X), dW, zk.to_device(Z), B, I, O, gens, com_tab, u_bs, u_in, u_out, w_int=prep)
    assert eq(zk.to_host(pfr2), pfr) and orc.g1_eq(zk.to_host(pg12), pg1).all()
    prep.close()
    gens.close(); com_tab.close()


@pytest.mark.parametrize("n", [2, 64, 4096, 1 << 15])
def test_zkrelu_prove(zk, n):
    L = orc.ceil_log2(n)
    xs = [int(v) for v in rng.integers(-(1 << 40), 1 << 40, size=n)]
    X = orc.fr_from_ints(xs, mont=True)
    dX = zk.to_device(X)
    Z, sign, mag, rem, bad = zk.relu(dX)
    oZ, osign, omag, orem, _ = orc.relu(X)
    ch = [rand_fr(L + 5), rand_fr(L + 5), rand_fr(L + 4), rand_fr(L + 4), rand_fr(L), rand_fr(L), rand_fr(L)]
    got = zk.to_host(zk.zkrelu_prove(dX, sign, mag, rem, *ch))
    ref = orc.zkrelu_prove(X, osign, omag, orem, *ch)
    exp = np.concatenate([ref["mag_sc"], ref["mag_rec"], ref["rem_sc"], ref["rem_rec"], ref["hp"]])
    assert eq(got, exp)
    # packed auxiliary input: same proof from 48 bits per activation
    Z2, sign2, magp, remp, bad2 = zk.relu_packed(dX)
    assert eq(zk.to_host(Z2), oZ) and eq(zk.to_host(sign2), osign)
    mag2, rem2 = zk.relu_expand(magp, remp)
    assert eq(zk.to_host(mag2), omag) and eq(zk.to_host(rem2), orem)
    got2 = zk.to_host(zk.zkrelu_prove_packed(dX, sign2, magp, remp, *ch))
    assert eq(got2, exp)


def test_random_vec_matches_oracle(zk):
    assert eq(zk.random_vec(12345, 33), orc.random_vec(12345, 33))


@pytest.mark.parametrize("world", [2, 3, 8])
def test_subtask_partition_reproduces_the_whole_proof(zk, world):
    """The sub-layer partition (parallel.partition_subtasks): every rank's pieces, proved separately through the
    *_parts entry points and re-assembled, are bit-identical to the proof of the undivided prover."""
    import torch
    from zkdl_b200 import mlp, parallel
    dims = [(20, 32), (32, 64), (64, 30), (30, 16)]
    ws, x = mlp.synthetic_mlp(dims, 8, seed=3)
    P = mlp.MLPProver(ws, gen_seed=2)
    P.forward(x)
    whole = P.prove(seed=11, streams=1)
    shapes = [(L.I, L.O) for L in P.layers]
    meta = {(k, i): (L.I, L.ngens, P.B * L.O) for i, L in enumerate(P.layers) for k in ("fc", "relu")}
    order = [(p[0], p[1]) for p in whole]
    plans = parallel.partition_subtasks(shapes, P.B, world)
    flats = []
    for r in range(world):
        res = P.prove(seed=11, parts=plans[r], streams=4 if r % 2 else 1)
        flats.append(parallel.pack_owned(res, plans[r], meta) if res else torch.zeros(0, dtype=torch.int32, device="cuda"))
    got = parallel.assemble(flats, plans, meta, order)
    for p in whole:
        key = (p[0], p[1])
        assert eq(zk.to_host(got[key][0]), zk.to_host(p[2]))
        if key[0] == "fc":
            assert orc.g1_eq(zk.to_host(got[key][1]), zk.to_host(p[3])).all()
