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
