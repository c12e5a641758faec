"""Linked mode (/root/reference/README.md:44-46 lists it as future work): ONE Fiat-Shamir proof in
which every claim a layer proof ends in is the claim the next one starts from, instead of the reference's 15 unrelated
fragments (demo.cu:124-138), and in which the ReLU auxiliary tables are committed so that the claims about them are
opened instead of taken on trust.  This is protocol design on top of the path's own building blocks, not a restatement:
the sumchecks, commitments and openings are the reference's (proof.cu:72-200, commitment.cu:29-92) through the C ABI.

Statement: for the public input table X and output table Z_last and the committed weights, Z_last = MLP_W(X) in the
reference's quantised arithmetic (zkfc.cu:6-47, zkrelu.cu:11-42).

Chain, from the output backwards (r = evaluation point over (out, batch) index bits, low bit first as in FrTensor::operator()):
    claim  Z_i~(r) = z                         z for the last layer: the verifier evaluates the public output itself
    zkFC   sum_k X_i~(u_bs, k) W_i~(k, u_out) = z   inner-product sumcheck (fold challenges u_in from the transcript)
           -> W_i~(u_out, u_in) opened against the public weight commitment;  X_i~(u_in, u_bs) = a  is the next claim
    i = 0  the verifier evaluates the public input at (u_in, u_bs) itself
    zkReLU X_i = A_{i-1} = M o sign  (zkrelu.cu:41):  sum_x eq(p, x) M(x) sign(x) = a   Hadamard sumcheck with the eq
           point p = (u_in, u_bs)  ->  M~(q), sign~(q) at the fold point q
           recover rows at q (the reference takes them at an unrelated point): mag_bin~(k, q), rem_bin~(k, q);
           M~(q) = sum_k 2^k mag_bin~(k, q), rem~(q) = sum_{k<15} 2^k rem_bin~(k, q) - 2^15 rem_bin~(15, q)
           Z_{i-1}~(q) = 2^16 M~(q) + rem~(q) - 2^47 (1 - sign~(q))     (relu_kernel's decomposition)  -> next claim, r = q
           binary sumchecks on mag_bin, rem_bin AND sign (the reference has none for sign)
           range of the magnitude: relu_kernel's M lies in [0, 2^31] (2^31 itself for -2^15 <= Z < 0, rounded up), but 32 bits
           hold twice that, and (sign, M) / (1 - sign, M + 2^31) would decompose the same Z with different outputs.  So
           M <= 2^31 is proved as  b31(x) * low31(x) = 0  for all x:  sum_x eq(t, x) b31(x) low31(x) = 0, a Hadamard sumcheck
           at a transcript point t, ending in b31~(q2), low31~(q2), which a second set of recover rows at q2 gives.  With it
           the decomposition is unique except Z in [-2^15, 2^15), where both forms give the output 0.
           openings against the auxiliary commitments: mag_bin at (tau_m, q), at (tau_2, q2) and at its binary sumcheck's fold
           point, rem_bin at (tau_r, q) and at its fold point, sign at q and at its fold point (tau_*: transcript points
           batching the 32 / 16 recover rows into one evaluation)
The transcript is seeded with the public model, the input, the output and the auxiliary commitments, so every challenge
depends on all of them.  Not covered: the reference's opening folds with the coordinates of the evaluation point
(commitment.cu:43-81) - it is kept as is."""
import hashlib

import numpy as np

from . import capi as zk
from . import fiat_shamir as fs
from . import verify
from .verify import P as MOD, plain, _req

Q_BITS, R_BITS = 32, 16                               # zkrelu.cu:73-76
N_OPENS = 7                                           # openings per zkReLU step


def _clog(v):
    return 0 if v <= 1 else (int(v) - 1).bit_length()


def _canon_g1(points):
    a = np.array(zk.to_host(zk.g1_normalize(points)), dtype=np.uint32).reshape(-1, 36)
    a[(a[:, 24:] == 0).all(axis=1)] = 0
    return a


def _cat(*vs):
    vs = [np.asarray(v, dtype=np.uint32).reshape(-1, 8) for v in vs]
    return np.concatenate(vs) if vs else np.zeros((0, 8), np.uint32)


def _pad_table(t, n):
    """Zero-extends a table to n cells: t'~(u, 0...0) = t~(u), so a table shorter than the generator set is opened at the
    zero-extended point."""
    if t.shape[0] >= n:
        return t
    import torch
    return torch.cat([t, torch.zeros((n - t.shape[0], 8), dtype=t.dtype, device=t.device)])


def _pad_point(u, k):
    u = _cat(u)
    return u if len(u) >= k else _cat(u, np.zeros((k - len(u), 8), np.uint32))


def _eq_weights(tau):
    w = [1]
    for row in _cat(tau):
        x = plain(row)
        w = [wb * (1 - x) % MOD for wb in w] + [wb * x % MOD for wb in w]
    return w


def linked_root(public, batch, x_host, z_host, aux_com):
    h = hashlib.sha256(fs.DOMAIN + b"/linked" + fs.public_root(public, batch))
    h.update(np.ascontiguousarray(x_host, dtype=np.uint32).astype("<u4").tobytes())
    h.update(np.ascontiguousarray(z_host, dtype=np.uint32).astype("<u4").tobytes())
    for coms in aux_com:
        for c in coms:
            h.update(np.ascontiguousarray(c, dtype=np.uint32).astype("<u4").tobytes())
    return h.digest()


# ------------------------------------------------------------------------------------------------ prover
class _Aux:
    """The auxiliary tables of one zkReLU layer as Montgomery Fr tables, and their row commitments."""

    def __init__(self, P, j):
        import torch
        L = P.layers[j]
        sign, magp, remp = P.aux[j]
        self.sign = sign
        self.mag, self.rem = zk.relu_expand(magp, remp)                     # 0/1 tables, index = 32 x + k / 16 x + k
        m = torch.zeros((magp.shape[0], 8), dtype=torch.int32, device=magp.device)
        m[:, 0] = magp                                                      # the packed magnitude IS mag_rescaled (zkrelu.cu:31)
        self.M = zk.fr_elementwise(zk.OP_MONT, m, out=m)
        lo, top = torch.zeros_like(m), torch.zeros_like(m)
        lo[:, 0] = magp & 0x7FFFFFFF                                        # low31 and b31 of the same integers
        top[:, 0] = (magp >> 31) & 1
        self.low, self.top = zk.fr_elementwise(zk.OP_MONT, lo, out=lo), zk.fr_elementwise(zk.OP_MONT, top, out=top)
        self.tables = [_pad_table(t, L.ngens) for t in (self.sign, self.mag, self.rem)]
        self.com = [zk.commit(L.gens, t) for t in self.tables]             # unmont(0/1) = 0/1: one-digit scalars
        self.com_host = [_canon_g1(c) for c in self.com]


def _open(L, table, com, u):
    """Commitment::open of `table` (committed row-wise under layer L's generators) at u."""
    u = _pad_point(u, _clog(table.shape[0]))
    tab = zk.G1Table(com, full=False)
    com_eval, proof, ret = zk.open_(L.gens, tab, table, u)
    out = {"ret": zk.to_host(ret).copy(), "g1": _canon_g1(_torch_cat([com_eval, proof]))}
    tab.close()
    return out


def _torch_cat(ts):
    import torch
    return torch.cat(list(ts))


def prove(P, check=False):
    """P: mlp.MLPProver after forward().  Returns (public, proof).  check=True asserts every chained claim against the
    prover's own tables as it goes (debugging aid: a broken link is reported where it happens)."""
    public = fs.public_part(P)
    nl = len(P.layers)
    B, kb = P.B, _clog(P.B)
    aux = [_Aux(P, j) for j in range(nl - 1)]
    x_host, z_host = zk.to_host(P.X).copy(), zk.to_host(P.Z[-1]).copy()
    aux_com = [a.com_host for a in aux]
    T = fs.Transcript(linked_root(public, B, x_host, z_host, aux_com), "linked", 0)
    Ll = P.layers[-1]
    r = T.vector(_clog(Ll.O) + kb)
    steps = []
    for i in range(nl - 1, -1, -1):
        L = P.layers[i]
        ki, ko = _clog(L.I), _clog(L.O)
        u_out, u_bs = r[:ko], r[ko:]
        X = P.A[i - 1] if i > 0 else P.X
        if check:
            claim = plain(zk.to_host(zk.fr_me(P.Z[i], r))[0])
            if i < nl - 1:
                assert claim == z_next, f"layer {i}: chained Z claim does not match the table"
        Xr = zk.fr_partial_me(X, u_bs, L.I) if kb else X
        Wr = zk.fr_partial_me(L.W, u_out, 1)
        ip, u_in, T.s = zk.sumcheck_fs(zk.FS_IP, Xr, Wr, None, ki, T.s)
        ip, u_in = zk.to_host(ip).copy(), zk.to_host(u_in).copy()
        T.absorb_fr(ip[3 * ki:])
        opening = _open(L, L.W, L.com, _cat(u_out, u_in))
        steps.append({"kind": "fc", "layer": i, "ip": ip, "open_w": opening})
        if i == 0:
            break
        # ---- zkReLU of layer i - 1: claim A~(p) = ip's final a(0)
        j, A = i - 1, aux[i - 1]
        Lj = P.layers[j]
        Lg = _clog(B * Lj.O)
        p = _cat(u_in, u_bs)
        hp, q, T.s = zk.sumcheck_fs(zk.FS_HP, A.M, A.sign, p, Lg, T.s)
        hp, q = zk.to_host(hp).copy(), zk.to_host(q).copy()
        T.absorb_fr(hp[3 * Lg:])
        r_mag = zk.to_host(zk.fr_partial_me(A.mag, q, Q_BITS)).copy()
        r_rem = zk.to_host(zk.fr_partial_me(A.rem, q, R_BITS)).copy()
        T.absorb_fr(r_mag); T.absorb_fr(r_rem)
        u_z = T.vector(Lg + 5)
        p_mag, v_z, T.s = zk.sumcheck_fs(zk.FS_BIN, A.mag, None, u_z, Lg + 5, T.s)
        p_mag, v_z = zk.to_host(p_mag).copy(), zk.to_host(v_z).copy()
        T.absorb_fr(p_mag[-1:])
        u_r = T.vector(Lg + 4)
        p_rem, v_r, T.s = zk.sumcheck_fs(zk.FS_BIN, A.rem, None, u_r, Lg + 4, T.s)
        p_rem, v_r = zk.to_host(p_rem).copy(), zk.to_host(v_r).copy()
        T.absorb_fr(p_rem[-1:])
        u_s = T.vector(Lg)
        p_sign, v_s, T.s = zk.sumcheck_fs(zk.FS_BIN, A.sign, None, u_s, Lg, T.s)
        p_sign, v_s = zk.to_host(p_sign).copy(), zk.to_host(v_s).copy()
        T.absorb_fr(p_sign[-1:])
        t2 = T.vector(Lg)                                                    # M <= 2^31:  b31 o low31 = 0
        top, q2, T.s = zk.sumcheck_fs(zk.FS_HP, A.top, A.low, t2, Lg, T.s)
        top, q2 = zk.to_host(top).copy(), zk.to_host(q2).copy()
        T.absorb_fr(top[3 * Lg:])
        r_top = zk.to_host(zk.fr_partial_me(A.mag, q2, Q_BITS)).copy()
        T.absorb_fr(r_top)
        tau_m, tau_r, tau_2 = T.vector(5), T.vector(4), T.vector(5)
        t_sign, t_mag, t_rem = A.tables
        c_sign, c_mag, c_rem = A.com
        opens = [_open(Lj, t_mag, c_mag, _cat(tau_m, q)), _open(Lj, t_mag, c_mag, v_z),
                 _open(Lj, t_rem, c_rem, _cat(tau_r, q)), _open(Lj, t_rem, c_rem, v_r),
                 _open(Lj, t_sign, c_sign, q), _open(Lj, t_sign, c_sign, v_s), _open(Lj, t_mag, c_mag, _cat(tau_2, q2))]
        steps.append({"kind": "relu", "layer": j, "hp": hp, "r_mag": r_mag, "r_rem": r_rem, "bin_mag": p_mag, "bin_rem": p_rem,
                      "bin_sign": p_sign, "top": top, "r_top": r_top, "opens": opens})
        if check:
            Mq, sq = plain(hp[-2]), plain(hp[-1])
            remq = (sum(plain(r_rem[k]) << k for k in range(15)) - (plain(r_rem[15]) << 15)) % MOD
            z_next = ((Mq << 16) + remq - (1 << 47) * (1 - sq)) % MOD
        r = q
    return public, {"batch": B, "input": x_host, "output": z_host, "aux_com": aux_com, "steps": steps}


# ------------------------------------------------------------------------------------------------ verifier
def _check_open(G, gens_table, com_dev, rec, u, expect, what):
    g = _clog(G.shape[0])
    rows = com_dev.shape[0]
    u = _pad_point(u, g + _clog(rows))
    _req(len(u) == g + _clog(rows), f"{what}: evaluation point does not address the committed table")
    g1 = zk.to_device(rec["g1"])
    _req(g1.shape[0] == 3 * g + 2, f"{what}: wrong opening length")
    verify.verify_subgroup(g1, what)
    verify.verify_commitment_eval(com_dev, g1[:1], u[g:])
    val = verify.verify_opening(G, g1[:1], g1[1:], zk.to_device(rec["ret"]), u[:g], gens_table)
    _req(val == expect % MOD, f"{what}: opened value differs from the claimed evaluation")


def verify_linked(public, proof):
    """Checks the whole chain from the public output down to the public input.  Raises verify.VerifyError."""
    nl = len(public)
    B = int(proof["batch"])
    kb = _clog(B)
    steps = proof["steps"]
    expected = [("fc", nl - 1)] + [(k, i) for i in range(nl - 2, -1, -1) for k in ("relu", "fc")]
    _req([(s["kind"], s["layer"]) for s in steps] == expected, "not exactly the chain of the public model")
    _req(len(proof["aux_com"]) == nl - 1, "auxiliary commitments do not match the model")
    x_host, z_host = np.asarray(proof["input"], np.uint32), np.asarray(proof["output"], np.uint32)
    _req(x_host.shape == (B * public[0]["I"], 8) and z_host.shape == (B * public[-1]["O"], 8), "input / output shapes")
    _req(B == 1 << kb, "batch is not a power of two")
    gens, tabs, aux_dev = [], [], []
    for j, L in enumerate(public):
        G = zk.to_device(L["generators"])
        ng = 1 << ((_clog(L["in_dim"] * L["out_dim"]) + 1) // 2)                                       # demo.cu:81
        _req(L["I"] == 1 << _clog(L["in_dim"]) and L["O"] == 1 << _clog(L["out_dim"]) and G.shape[0] == ng
             and len(L["commitment"]) * ng == L["I"] * L["O"] and (j == 0 or public[j - 1]["O"] == L["I"]),
             f"layer {j}: public shapes are inconsistent")
        verify.verify_subgroup(G, f"layer {j} generators")
        verify.verify_subgroup(zk.to_device(L["commitment"]), f"layer {j} commitment")
        gens.append(G); tabs.append(zk.G1Table(G, full=False))
        if j < nl - 1:
            n = B * L["O"]
            rows = [max(n * w // G.shape[0], 1) for w in (1, Q_BITS, R_BITS)]
            _req([c.shape[0] for c in proof["aux_com"][j]] == rows, f"relu {j}: auxiliary commitment sizes")
            aux_dev.append([zk.to_device(c) for c in proof["aux_com"][j]])
            for c in aux_dev[-1]:
                verify.verify_subgroup(c, f"relu {j} auxiliary commitment")
    T = fs.Transcript(linked_root(public, B, x_host, z_host, proof["aux_com"]), "linked", 0)
    r = T.vector(_clog(public[-1]["O"]) + kb)
    z = plain(zk.to_host(zk.fr_me(zk.to_device(z_host), r))[0])                       # the public output at r
    it = iter(steps)
    try:
        for i in range(nl - 1, -1, -1):
            L = public[i]
            ki, ko = _clog(L["I"]), _clog(L["O"])
            u_out, u_bs = r[:ko], r[ko:]
            s = next(it)
            ip = np.asarray(s["ip"], np.uint32)
            _req(ip.shape == (3 * ki + 2, 8), f"fc {i}: wrong sumcheck length")
            u_in = T.rounds(ip, ki)
            a0, b0 = verify.verify_ip(ip, u_in, z)
            T.absorb_fr(ip[3 * ki:])
            _check_open(gens[i], tabs[i], zk.to_device(L["commitment"]), s["open_w"], _cat(u_out, u_in), b0, f"fc {i} weight opening")
            if i == 0:
                x_eval = plain(zk.to_host(zk.fr_me(zk.to_device(x_host), _cat(u_in, u_bs)))[0])
                _req(x_eval == a0, "fc 0: the chain does not end in the public input")
                break
            j = i - 1
            Lj = public[j]
            Lg = _clog(B * Lj["O"])
            s = next(it)
            hp, r_mag, r_rem = (np.asarray(s[k], np.uint32) for k in ("hp", "r_mag", "r_rem"))
            p_mag, p_rem, p_sign = (np.asarray(s[k], np.uint32) for k in ("bin_mag", "bin_rem", "bin_sign"))
            top, r_top = np.asarray(s["top"], np.uint32), np.asarray(s["r_top"], np.uint32)
            _req(hp.shape == (3 * Lg + 2, 8) and r_mag.shape == (Q_BITS, 8) and r_rem.shape == (R_BITS, 8)
                 and top.shape == (3 * Lg + 2, 8) and r_top.shape == (Q_BITS, 8), f"relu {j}: wrong lengths")
            _req(len(s["opens"]) == N_OPENS, f"relu {j}: {N_OPENS} openings expected")
            p = _cat(u_in, u_bs)
            q = T.rounds(hp, Lg)
            Mq, sq = verify.verify_weighted(hp, p, q, a0, 2, lambda c, f: c == f[0] * f[1] % MOD)
            T.absorb_fr(hp[3 * Lg:])
            T.absorb_fr(r_mag); T.absorb_fr(r_rem)
            rm, rr = [plain(x) for x in r_mag], [plain(x) for x in r_rem]
            _req(sum(v << k for k, v in enumerate(rm)) % MOD == Mq, f"relu {j}: recover rows do not add up to the magnitude")
            remq = (sum(rr[k] << k for k in range(R_BITS - 1)) - (rr[R_BITS - 1] << (R_BITS - 1))) % MOD
            u_z = T.vector(Lg + 5); v_z = T.rounds(p_mag, Lg + 5)
            f_mag = verify.verify_bin(p_mag, u_z, v_z)[0]
            T.absorb_fr(p_mag[-1:])
            u_r = T.vector(Lg + 4); v_r = T.rounds(p_rem, Lg + 4)
            f_rem = verify.verify_bin(p_rem, u_r, v_r)[0]
            T.absorb_fr(p_rem[-1:])
            u_s = T.vector(Lg); v_s = T.rounds(p_sign, Lg)
            f_sign = verify.verify_bin(p_sign, u_s, v_s)[0]
            T.absorb_fr(p_sign[-1:])
            t2 = T.vector(Lg); q2 = T.rounds(top, Lg)
            b31, low = verify.verify_weighted(top, t2, q2, 0, 2, lambda c, f: c == f[0] * f[1] % MOD)
            T.absorb_fr(top[3 * Lg:])
            T.absorb_fr(r_top)
            rt = [plain(x) for x in r_top]
            _req(rt[Q_BITS - 1] == b31 and sum(v << k for k, v in enumerate(rt[:Q_BITS - 1])) % MOD == low,
                 f"relu {j}: second recover rows do not match the magnitude-range sumcheck")
            tau_m, tau_r, tau_2 = T.vector(5), T.vector(4), T.vector(5)
            e_mag = sum(w * v for w, v in zip(_eq_weights(tau_m), rm)) % MOD
            e_rem = sum(w * v for w, v in zip(_eq_weights(tau_r), rr)) % MOD
            e_top = sum(w * v for w, v in zip(_eq_weights(tau_2), rt)) % MOD
            c_sign, c_mag, c_rem = aux_dev[j]
            o = s["opens"]
            G, gt = gens[j], tabs[j]
            _check_open(G, gt, c_mag, o[0], _cat(tau_m, q), e_mag, f"relu {j} mag_bin at the recover point")
            _check_open(G, gt, c_mag, o[1], v_z, f_mag, f"relu {j} mag_bin at its sumcheck point")
            _check_open(G, gt, c_rem, o[2], _cat(tau_r, q), e_rem, f"relu {j} rem_bin at the recover point")
            _check_open(G, gt, c_rem, o[3], v_r, f_rem, f"relu {j} rem_bin at its sumcheck point")
            _check_open(G, gt, c_sign, o[4], q, sq, f"relu {j} sign at the recover point")
            _check_open(G, gt, c_sign, o[5], v_s, f_sign, f"relu {j} sign at its sumcheck point")
            _check_open(G, gt, c_mag, o[6], _cat(tau_2, q2), e_top, f"relu {j} mag_bin at the range point")
            z = ((Mq << 16) + remq - (1 << 47) * (1 - sq)) % MOD                    # Z_j~(q): the next layer's claim
            r = q
    finally:
        for t in tabs:
            t.close()
    return True




# ------------------------------------------------------------------------------------------------ wire format (serialize.py, version 3)
def to_tasks(proof):
    """The chain as serialize.py task records: fc = [ip | open ret] + opening points, relu = [hp | recover rows | three binary
    sumchecks | range sumcheck | second recover rows | open rets] + the openings' points."""
    tasks = []
    for s in proof["steps"]:
        if s["kind"] == "fc":
            fr, g1 = _cat(s["ip"], s["open_w"]["ret"]), s["open_w"]["g1"]
        else:
            fr = _cat(s["hp"], s["r_mag"], s["r_rem"], s["bin_mag"], s["bin_rem"], s["bin_sign"], s["top"], s["r_top"],
                      *[o["ret"] for o in s["opens"]])
            g1 = np.concatenate([o["g1"] for o in s["opens"]])
        tasks.append({"kind": s["kind"], "layer": s["layer"], "challenges": [], "fr": fr, "g1": g1})
    return tasks


def from_tasks(public, batch, extra, tasks):
    """Inverse of to_tasks for a loaded file; raises verify.VerifyError when a record does not have the protocol's lengths."""
    steps = []
    for t in tasks:
        _req(t["layer"] < len(public) and t["g1"] is not None, "malformed record")
        L = public[t["layer"]]
        g = _clog(len(L["generators"]))
        fr, g1 = t["fr"], t["g1"]
        if t["kind"] == "fc":
            ki = _clog(L["I"])
            _req(len(fr) == 3 * ki + 3 and len(g1) == 3 * g + 2, f"fc {t['layer']}: wrong number of proof elements")
            steps.append({"kind": "fc", "layer": t["layer"], "ip": fr[:3 * ki + 2], "open_w": {"ret": fr[3 * ki + 2:], "g1": g1}})
        else:
            Lg = _clog(batch * L["O"])
            cuts = np.cumsum([3 * Lg + 2, Q_BITS, R_BITS, 3 * (Lg + 5) + 1, 3 * (Lg + 4) + 1, 3 * Lg + 1, 3 * Lg + 2, Q_BITS])
            _req(len(fr) == cuts[-1] + N_OPENS and len(g1) == N_OPENS * (3 * g + 2), f"relu {t['layer']}: wrong number of proof elements")
            hp, r_mag, r_rem, b_mag, b_rem, b_sign, top, r_top, rets = np.split(fr, cuts)
            opens = [{"ret": rets[k:k + 1], "g1": g1[k * (3 * g + 2):(k + 1) * (3 * g + 2)]} for k in range(N_OPENS)]
            steps.append({"kind": "relu", "layer": t["layer"], "hp": hp, "r_mag": r_mag, "r_rem": r_rem, "bin_mag": b_mag,
                          "bin_rem": b_rem, "bin_sign": b_sign, "top": top, "r_top": r_top, "opens": opens})
    return {"batch": batch, "input": extra["input"], "output": extra["output"], "aux_com": extra["aux_com"], "steps": steps}


def export(public, proof, path):
    from . import serialize
    blob = serialize.dumps({"batch": proof["batch"], "layers": public}, to_tasks(proof),
                           linked={k: proof[k] for k in ("input", "output", "aux_com")})
    with open(path, "wb") as f:
        f.write(blob)
    return len(blob)


def verify_file(path):
    from . import serialize
    with open(path, "rb") as f:
        public, tasks = serialize.loads(f.read())
    _req(public.get("linked") is not None, "not a linked-mode proof file")
    return verify_linked(public["layers"], from_tasks(public["layers"], public["batch"], public["linked"], tasks))
