"""BASELINE.json configs 2, 3 and 5 (the bench line itself is config 0/3/4): prints one JSON line per measurement.
   python tools/run_configs.py [fc4096] [msm [max_log]] [deep]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from zkdl_b200 import capi as zk, mlp

def timed(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

def small_fr(n, bits, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    v = torch.randint(-(1 << (bits - 1)), 1 << (bits - 1), (n, 1), generator=g, device="cuda", dtype=torch.int32).float() / 65536.0
    q = zk.float_to_fr(v.reshape(-1, 1), n, 1)
    return zk.fr_elementwise(zk.OP_MONT, q, out=q)

args = sys.argv[1:] or ["fc4096", "msm", "deep"]
if "fc4096" in args:                                  # config 2: standalone zkFC matmul sumcheck, 4096x4096, batch 256
    I = O = 4096; B = 256
    W, X = small_fr(I * O, 16, 1), small_fr(B * I, 16, 2)
    Z = zk.fr_matmul(X, W, B, I, O)
    u_bs, u_in, u_out = zk.random_vec(3, 8), zk.random_vec(4, 12), zk.random_vec(5, 12)
    def sc():
        Xr = zk.fr_partial_me(X, u_bs, I); Wr = zk.fr_partial_me(W, u_out, 1)
        zk.ip_sumcheck(Xr, Wr, u_in); zk.fr_me(Z, np.concatenate([u_out, u_bs]))
    ms = timed(sc)
    alg = 96 * (I * O) * (1 - 2.0 ** -12) + 96 * (B * I) * (1 - 2.0 ** -8) + 96 * (B * O) * (1 - 2.0 ** -20)
    print(json.dumps({"config": "fc4096 sumcheck set (X.partial_me, W.partial_me, ip sumcheck, Z(u))", "ms": ms, "algorithmic_GB": alg / 1e9,
                      "algorithmic_GBps": alg / ms / 1e6, "reference_ms": "28-145 (profiles/r1_reference_times_b200.jsonl, ref_fc_sumcheck)"}))
    ms = timed(lambda: zk.fr_matmul(X, W, B, I, O))
    print(json.dumps({"config": "fc4096 forward matmul 256x4096x4096", "ms": ms}))
    ng = 4096
    G = zk.g1_mul(zk.to_device(mlp._generator()), zk.to_device(zk.random_vec(6, ng)))
    t0 = time.time(); gens = zk.G1Table(G, full=True); torch.cuda.synchronize(); t_tab = time.time() - t0
    ms_c = timed(lambda: zk.commit(gens, W), reps=3, warm=1)
    com = zk.commit(gens, W); com_tab = zk.G1Table(com, full=True)
    ms_p = timed(lambda: zk.zkfc_prove(X, W, Z, B, I, O, gens, com_tab, u_bs, u_in, u_out))
    print(json.dumps({"config": "fc4096 commit (4096 rows x 4096 generators, 16-bit weights)", "ms": ms_c, "Mpts_per_s": I * O / ms_c / 1e3, "table_build_s": t_tab}))
    print(json.dumps({"config": "fc4096 zkFC::prove (sumcheck + opening)", "ms": ms_p}))
    del W, X, Z, gens, com_tab
if "msm" in args:                                     # config 3: G1 Pedersen commitment MSM sweep, m = 1
    mx = int(args[args.index("msm") + 1]) if len(args) > args.index("msm") + 1 and args[args.index("msm") + 1].isdigit() else 22
    gen = zk.to_device(mlp._generator())
    for lg in range(16, mx + 1, 2):
        N = 1 << lg
        ks = zk.to_device(zk.random_vec(7, N))
        t0 = time.time(); G = zk.g1_mul(gen, ks); torch.cuda.synchronize(); t_gen = time.time() - t0
        tab = zk.G1Table(G, full=False)
        full = zk.to_device(zk.random_vec(8, N))
        ms_full = timed(lambda: zk.msm(tab, full, 1, False), reps=3, warm=1)
        s16 = small_fr(N, 16, 9)
        ms_16 = timed(lambda: zk.commit(tab, s16), reps=3, warm=1)
        print(json.dumps({"config": f"msm 2^{lg} plain Pippenger", "ms_255bit": ms_full, "Mpts_per_s_255bit": N / ms_full / 1e3,
                          "ms_16bit_signed": ms_16, "Mpts_per_s_16bit": N / ms_16 / 1e3, "bases_gen_s": t_gen}))
        tab.close(); del G, ks, full, s16
if "deep" in args:                                    # config 5: depth 32, width 1024, batch 1 vs 4096
    dims = [(1024, 1024)] * 32
    for batch in (1, 4096):
        ws, x = mlp.synthetic_mlp(dims, batch, seed=3)
        t0 = time.time(); P = mlp.MLPProver(ws); torch.cuda.synchronize(); t_setup = time.time() - t0
        ms_f = timed(lambda: P.forward(x), reps=2, warm=1)
        ms_p = timed(lambda: P.prove(seed=1), reps=3, warm=2)
        print(json.dumps({"config": f"deep-narrow 32x(1024->1024), batch {batch}", "prove_ms": ms_p, "forward_ms": ms_f, "setup_s": t_setup,
                          "mem_GB": torch.cuda.max_memory_allocated() / 1e9}))
        del P, ws, x
if "fs" in args:                                      # Fiat-Shamir mode of the demo MLP (config 4 with transcript challenges)
    from zkdl_b200 import fiat_shamir as fs
    ws, x = mlp.synthetic_mlp(mlp.demo_layer_dims(), 256, seed=0)
    P = mlp.MLPProver(ws, gen_seed=1)
    P.forward(x)
    fs.prove(P)
    torch.cuda.synchronize(); t0 = time.time()
    public, proofs = fs.prove(P)
    torch.cuda.synchronize(); t_prove = time.time() - t0
    t0 = time.time(); ok = fs.verify_all(public, P.B, proofs); t_verify = time.time() - t0
    print(json.dumps({"config": "demo MLP 18.2M params, batch 256, Fiat-Shamir mode (device-side transcript, reference Fr tables, layers sequential)",
                      "prove_ms": t_prove * 1e3, "verify_s": t_verify, "verified": bool(ok)}))
if "linked" in args:                                  # linked mode of the demo MLP: one chained proof, auxiliary tables committed and opened
    from zkdl_b200 import linked
    ws, x = mlp.synthetic_mlp(mlp.demo_layer_dims(), 256, seed=0)
    P = mlp.MLPProver(ws, gen_seed=1)
    P.forward(x); P.check_range()
    linked.prove(P)
    torch.cuda.synchronize(); t0 = time.time()
    public, proof = linked.prove(P)
    torch.cuda.synchronize(); t_prove = time.time() - t0
    t0 = time.time(); aux = [linked._Aux(P, j) for j in range(len(P.layers) - 1)]; torch.cuda.synchronize(); t_aux = time.time() - t0
    del aux
    t0 = time.time(); ok = linked.verify_linked(public, proof); t_verify = time.time() - t0
    path = os.path.join(os.environ.get("TMPDIR", "/tmp"), "linked.zkp")
    t0 = time.time(); nbytes = linked.export(public, proof, path); t_export = time.time() - t0
    t0 = time.time(); ok2 = linked.verify_file(path); t_file = time.time() - t0
    print(json.dumps({"config": "demo MLP 18.2M params, batch 256, linked mode (one chained Fiat-Shamir proof; sign / mag_bin / rem_bin committed, 7 openings per zkReLU, magnitude range sumcheck)",
                      "prove_ms": t_prove * 1e3, "of_which_aux_expand_commit_ms": t_aux * 1e3, "verify_s": t_verify, "verified": bool(ok),
                      "file_bytes": nbytes, "export_s": t_export, "load_and_verify_s": t_file, "file_verified": bool(ok2),
                      "mem_GB": torch.cuda.max_memory_allocated() / 1e9}))
