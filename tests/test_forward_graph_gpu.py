"""MLPProver.forward(graph=True): the forward pass replayed from CUDA graphs gives exactly the tables of the eager pass, for
the captured input and for later inputs, and survives zk.scratch_release_all() (which frees the arenas the graphs point into)."""
import pytest

pytestmark = pytest.mark.gpu


def test_forward_graph_matches_eager():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from zkdl_b200 import capi as zk, mlp
    zk.lib()
    dims = [(784, 256), (256, 384), (384, 128), (128, 10)]
    ws, x0 = mlp.synthetic_mlp(dims, 128, seed=2)
    _, x1 = mlp.synthetic_mlp(dims, 128, seed=3)
    E, G = mlp.MLPProver(ws, gen_seed=1), mlp.MLPProver(ws, gen_seed=1)

    def same(x):
        E.forward(x)
        G.forward(x, graph=True)
        torch.cuda.synchronize()
        assert len(G.ready) == len(dims) and torch.equal(E.X, G.X)
        for a, b in zip(E.Z + E.A, G.Z + G.A):
            assert torch.equal(a, b)
        for ea, ga in zip(E.aux, G.aux):
            assert all(torch.equal(a, b) for a, b in zip(ea, ga))
        G.check_range()

    same(x0)
    same(x1)                                   # replay with another input
    same(x0)
    zk.scratch_release_all()                   # arenas freed: the next call must re-capture, not replay into freed memory
    same(x1)
    # a proof from graph-produced tables verifies like any other
    from zkdl_b200 import verify
    proof = G.prove(seed=5, overlap_forward=True)
    torch.cuda.synchronize()
    for part, (kind, i, ch, mask) in zip(proof, G.last_tasks):
        L = G.layers[i]
        if kind == "fc":
            verify.verify_zkfc(part[2], part[3], L.G, G.B, L.I, L.O, *ch)
        else:
            verify.verify_zkrelu(part[2], G.B * L.O, ch[0], ch[1], ch[2], ch[3], ch[5], ch[6])


@pytest.mark.parametrize("M,K,N,amax,wmax", [
    (256, 1024, 2048, 1 << 20, 1 << 14),       # tcgen05 route: zkReLU applied in the epilogue
    (128, 2048, 64, 1 << 23, 1 << 14),         # same route, |Z| reaches 2^48: out-of-range entries are counted, not decomposed
    (8, 32, 64, 1 << 20, 1 << 14),             # shape does not tile: mma.sync / integer kernels + the separate relu pass
    (128, 128, 64, 1 << 30, 1 << 14),          # activations too large for the byte planes: integer route + relu pass
])
def test_fused_layer_equals_product_then_relu(M, K, N, amax, wmax):
    """zkdl_fr_matmul_prepared_relu == zkdl_fr_matmul_prepared followed by zkdl_relu_packed, bit for bit, on every route."""
    import numpy as np
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from zkdl_b200 import capi as zk
    zk.lib()
    g = torch.Generator(device="cpu").manual_seed(M + K + N)

    def fr_from(v):
        limbs = torch.zeros((v.numel(), 8), dtype=torch.int64)
        a = v.abs()
        limbs[:, 0] = a & 0xFFFFFFFF
        limbs[:, 1] = a >> 32
        t = zk.to_device(limbs.numpy().astype(np.uint32))
        t = zk.fr_elementwise(zk.OP_MONT, t)
        neg = zk.fr_elementwise(zk.OP_NEG, t)
        m = (v < 0).cuda()
        t[m] = neg[m]
        return t

    def ints(n, top):
        v = torch.randint(-top + 1, top, (n,), generator=g, dtype=torch.int64)
        v[:4] = torch.tensor([top - 1, -(top - 1), 0, 1])
        return v

    va, vw = ints(M * K, amax), ints(K * N, wmax)
    if amax == 1 << 23:                          # rows 0 / 1 times column 0: +-K (2^23 - 1)(2^14 - 1) ~ +-2^48, outside +-2^47
        va[:K] = amax - 1
        va[K: 2 * K] = -(amax - 1)
        vw[0::N] = wmax - 1
    A, W = fr_from(va), fr_from(vw)
    mm = zk.MatmulWeights(W, K, N)
    z0 = zk.fr_matmul_prepared(A, mm, M)
    a0, s0, q0, r0, b0 = zk.relu_packed(z0)
    z1, a1, s1, q1, r1, b1 = zk.fr_matmul_prepared_relu(A, mm, M)
    torch.cuda.synchronize()
    assert torch.equal(z0, z1) and torch.equal(a0, a1) and torch.equal(s0, s1) and torch.equal(q0, q1) and torch.equal(r0, r1)
    assert int(b0.item()) == int(b1.item())
    if amax == 1 << 23:
        assert int(b1.item()) >= 2, "this case is meant to contain out-of-range pre-activations"
