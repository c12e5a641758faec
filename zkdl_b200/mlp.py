"""Host-side driver of the demo path (mirrors /root/reference/demo.cu:23-143 over the C ABI): load float weights,
build generators, commit, quantised forward pass, then the backward proving loop zkFC::prove / zkReLU::prove.

Python mirror of zkdl_b200/host/demo.cpp, used by bench.py and the tests.  Randomness is injected: generators and
challenges come from seeded streams (the reference uses std::random_device)."""
import numpy as np

from . import capi as zk


def ceil_log2(n):
    return 0 if n <= 1 else (int(n) - 1).bit_length()


def pad2(n):
    return 1 << ceil_log2(n)


def demo_layer_dims():
    """model.py:14-30 — the 18.2 M-parameter ModulusLab benchmark MLP."""
    d = [784, 1000, 1773, 1773, 1773, 1773, 1773, 1124, 1000]
    return list(zip(d[:-1], d[1:]))


class Layer:
    pass


class MLPProver:
    """weights: list of float32 CUDA tensors shaped [in, out] (== nn.Linear.weight.t(), demo.cu:72)."""

    def __init__(self, weights, gen_seed=1, world=1, rank=0, all_gather=None):
        """world > 1: Commitment::commit is sharded by row over the ranks (commitment.cu:29-41): rank r
        commits rows [r m/P, (r+1) m/P) of every weight table and the row commitments are all-gathered
        (parallel.commit_sharded); generators and window tables are replicated (|G| <= 4096 points)."""
        import torch
        from . import parallel
        self.layers = []
        rng = np.random.default_rng(gen_seed)
        gen = zk.to_device(_generator())
        for w in weights:
            L = Layer()
            L.in_dim, L.out_dim = int(w.shape[0]), int(w.shape[1])
            L.I, L.O = pad2(L.in_dim), pad2(L.out_dim)
            L.ngens = 1 << ((ceil_log2(L.in_dim * L.out_dim) + 1) // 2)                      # demo.cu:81
            # generators *= FrTensor::random(n) (demo.cu:82): the reference's curand stream, seeded per layer
            L.G = zk.g1_mul(gen, zk.fr_random(L.ngens, int(rng.integers(0, 1 << 63))))
            L.gens = zk.G1Table(L.G, full=True)
            q = zk.float_to_fr(w.contiguous(), L.I, L.O)                                     # zkfc.cu:90-100
            L.W = zk.fr_elementwise(zk.OP_MONT, q, out=q)
            L.com = parallel.commit_sharded(zk.commit, L.gens, L.W, L.ngens, world, rank, all_gather)   # zkfc.cu:102
            L.com_table = zk.G1Table(L.com, full=True)
            L.mm = zk.MatmulWeights(L.W, L.I, L.O)                                           # integer copy for the forward product
            self.layers.append(L)
        self.n_params = sum(L.in_dim * L.out_dim for L in self.layers)
        torch.cuda.synchronize()

    def forward(self, x, graph=False):
        """x: float32 CUDA [batch, in].  Keeps Z_i, A_i and the ReLU aux tables (demo.cu:23-38).
        graph=True replays the pass from CUDA graphs captured on the first call with this batch shape (one graph per layer,
        so the per-layer events of prove(overlap_forward=True) stay): the ~65 small launches of a batch-256 pass are
        launch-bound from Python (0.78 ms), the kernels themselves take a fraction of that.  The tables then live in the
        graphs' memory pool and are OVERWRITTEN by the next forward(graph=True) call."""
        import torch
        if graph:
            return self._forward_graph(x)
        self.B = pad2(x.shape[0])
        self.Z, self.A, self.aux, self.bad, self.ready = [], [], [], [], []
        cur = None
        for i in range(len(self.layers)):
            cur = self._forward_layer(i, x if i == 0 else cur)
            ev = torch.cuda.Event(); ev.record()                                            # layer i's tables are final here
            self.ready.append(ev)
        return self.Z[-1]

    def _forward_layer(self, i, cur):
        L = self.layers[i]
        if i == 0:
            X = zk.float_to_fr(cur.contiguous(), self.B, L.I)                                # zkfc.cu:106-115
            self.X = cur = zk.fr_elementwise(zk.OP_MONT, X, out=X)                           # demo.cu:119
        if i + 1 < len(self.layers):                                                         # product + zkReLU in one call, aux kept bit-packed
            z, a, sign, mag, rem, bad = zk.fr_matmul_prepared_relu(cur, L.mm, self.B)
            self.Z.append(z); self.A.append(a); self.aux.append((sign, mag, rem)); self.bad.append(bad)
            return a
        z = zk.fr_matmul_prepared(cur, L.mm, self.B)
        self.Z.append(z)
        return z

    def _forward_graph(self, x):
        import torch
        key = (tuple(x.shape), zk.scratch_generation())
        main = torch.cuda.current_stream()
        if getattr(self, "_fg_key", None) != key:
            # Capture.  The library's scratch arenas are keyed by stream and grow with cudaMalloc, which a capture forbids:
            # two eager passes on the capture stream size its arena first.  Scratch addresses are baked into the graphs, so
            # the stream is reserved for them and zk.scratch_release_all() (which frees arenas) invalidates the capture.
            self._fg_key, self._fg_graphs = None, []
            st = self._fg_stream = getattr(self, "_fg_stream", None) or torch.cuda.Stream()
            self._fg_x = torch.empty_like(x)
            self._fg_x.copy_(x)
            st.wait_stream(main)
            with torch.cuda.stream(st):
                for _ in range(2):
                    self.forward(self._fg_x)
            st.synchronize()
            self.B = pad2(x.shape[0])
            self.Z, self.A, self.aux, self.bad = [], [], [], []
            pool = torch.cuda.graph_pool_handle()
            cur = None
            for i in range(len(self.layers)):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, pool=pool, stream=st, capture_error_mode="thread_local"):   # NCCL watchdog / pool threads may touch CUDA meanwhile
                    cur = self._forward_layer(i, self._fg_x if i == 0 else cur)
                self._fg_graphs.append(g)
            self._fg_key = key
            self._fg_tables = (self.X, list(self.Z), list(self.A), list(self.aux), list(self.bad))
        # the tables the graphs write (an eager forward() in between rebinds self.Z ... to its own tensors)
        self.B = pad2(x.shape[0])
        self.X, self.Z, self.A, self.aux, self.bad = (self._fg_tables[0], *(list(t) for t in self._fg_tables[1:]))
        self._fg_x.copy_(x, non_blocking=True)
        self.ready = []
        for g in self._fg_graphs:
            g.replay()
            ev = torch.cuda.Event(); ev.record()
            self.ready.append(ev)
        return self.Z[-1]

    def check_range(self):
        """Activations outside +-2^47 are undefined in the reference (relu_kernel, zkrelu.cu:16-28); here they are
        counted on the device.  Synchronises; raises if the last forward pass saw any."""
        import torch
        if self.bad and int(torch.stack(self.bad).sum().item()) != 0:
            raise ValueError("zkReLU input outside +-2^47: the decomposition (and the reference's) is undefined for it")

    def prove(self, seed=0, fc_layers=None, relu_layers=None, streams=8, threads=None, parts=None, overlap_forward=False):
        """Backward proving loop (demo.cu:124-138).  Returns the proof parts in the reference's order.
        fc_layers / relu_layers restrict the work to a subset (layer-parallel multi-GPU); `parts` = {("fc"|"relu", layer):
        part mask} restricts it further to independent parts of a layer's proof (parallel.partition_subtasks): the
        returned buffers are full-size with only the owned segments written.  Challenges are drawn for every layer
        regardless, so a proof element does not depend on which rank produced it.
        Every layer's proof is independent of the others (all randomness is fresh), so the per-layer
        proofs are issued on `streams` CUDA streams: the latency-bound bucket reductions of one layer's opening overlap
        the bandwidth-bound sumcheck passes of another.  With threads=True each stream is fed by its own host thread
        (ctypes releases the GIL, the library is thread-safe), so the ~1200 kernel launches of a proof are issued in
        parallel instead of from one core.
        overlap_forward: the pieces are issued in ascending layer order and each waits only for ITS layer's event of the
        last forward() call, so proving the first layers overlaps the rest of the forward pass (the end-to-end path)."""
        import os
        import torch
        if threads is None:                              # ZKDL_PROVE_THREADS=0: issue from the calling thread (NVTX-scoped profiling)
            threads = os.environ.get("ZKDL_PROVE_THREADS", "1") != "0"
        ctr = [seed]

        def rv(k):
            ctr[0] += 1
            return zk.random_vec(ctr[0], k)

        B, kb = self.B, ceil_log2(self.B)
        nl = len(self.layers)
        tasks = []                                   # (kind, layer, challenges) in the reference's order

        def owned(kind, i, every):
            if parts is not None:
                return parts.get((kind, i), 0)
            sel = fc_layers if kind == "fc" else relu_layers
            return every if sel is None or i in sel else 0

        def fc(i):
            L = self.layers[i]
            mask = owned("fc", i, zk.FC_SUMCHECK | zk.FC_OPENING)
            if not mask:
                ctr[0] += 3                                                                  # another rank draws these
                return
            ch = (rv(kb), rv(ceil_log2(L.I)), rv(ceil_log2(L.O)))                             # zkfc.cu:135-137
            tasks.append(("fc", i, ch, mask))

        def relu(i):
            mask = owned("relu", i, zk.RELU_MAG | zk.RELU_REM | zk.RELU_HP)
            if not mask:
                ctr[0] += 7
                return
            Lg = ceil_log2(B * self.layers[i].O)
            ch = [rv(Lg + 5), rv(Lg + 5), rv(Lg + 4), rv(Lg + 4), rv(Lg)]                    # zkrelu.cu:85-89
            ch += [rv(Lg), rv(Lg)]                                                           # zkrelu.cu:97-98
            tasks.append(("relu", i, ch, mask))

        fc(nl - 1)
        for i in range(nl - 2, -1, -1):
            relu(i)
            fc(i)

        self.last_tasks = tasks                       # (kind, layer, challenges, parts): what a verifier needs besides the proof

        def run(task):
            kind, i, ch, mask = task
            L = self.layers[i]
            if kind == "fc":
                Xin = self.A[i - 1] if i > 0 else self.X
                return ("fc", i) + zk.zkfc_prove(Xin, L.W, self.Z[i], B, L.I, L.O, L.gens, L.com_table, *ch, parts=mask, w_int=L.mm)
            sign, mag, rem = self.aux[i]
            return ("relu", i, zk.zkrelu_prove_packed(self.Z[i], sign, mag, rem, *ch, parts=mask))

        main = torch.cuda.current_stream()
        if streams <= 1 or len(tasks) <= 1:
            return [run(t) for t in tasks]
        streams = min(streams, len(tasks))
        dev = torch.cuda.current_device()
        if len(getattr(self, "_streams", [])) < streams:
            # One CUDA stream and ONE dedicated host thread per slot: the library's side streams and scratch arenas are
            # per host thread / per stream, so a stable slot <-> thread pairing keeps every arena at its final size after
            # the first call (a shared pool that shuffles slots over threads grows arenas sporadically for many calls).
            from concurrent.futures import ThreadPoolExecutor
            old = len(getattr(self, "_streams", []))
            self._streams = getattr(self, "_streams", []) + [torch.cuda.Stream() for _ in range(streams - old)]
            self._pools = getattr(self, "_pools", []) + [ThreadPoolExecutor(max_workers=1) for _ in range(streams - old)]
            # measured peak of a slot's main arena: ~100 MB for a 2^19-activation zkReLU proof (eq / folded tables, look-up tables),
            # ~30 MB for an opening; side-stream arenas get half.  Arenas still grow on demand if this is ever too small.
            need = 96 * max(max(L.I * L.O, 4 * B * L.O) for L in self.layers)

            def reserve(slot):
                torch.cuda.set_device(dev)
                with torch.cuda.stream(self._streams[slot]):
                    zk.scratch_reserve(need)

            for f in [self._pools[s_].submit(reserve, s_) for s_ in range(old, streams)]:
                f.result()
            torch.cuda.synchronize()
        start = torch.cuda.Event()
        order = list(range(len(tasks)))
        if overlap_forward:
            order.sort(key=lambda j: (tasks[j][1], tasks[j][0] != "fc"))      # fc i needs Z_i only; relu i also its aux tables
        else:
            start.record(main)

        def worker(slot):
            torch.cuda.set_device(dev)
            st = self._streams[slot]
            if not overlap_forward:
                st.wait_event(start)
            res = []
            with torch.cuda.stream(st):
                for j in order[slot::streams]:
                    if overlap_forward:
                        st.wait_event(self.ready[tasks[j][1]])
                    res.append((j, run(tasks[j])))
            ev = torch.cuda.Event(); ev.record(st)
            return res, ev

        if threads:
            outs = [f.result() for f in [self._pools[s_].submit(worker, s_) for s_ in range(streams)]]
        else:
            outs = [worker(s_) for s_ in range(streams)]
        out = [None] * len(tasks)
        for res, ev in outs:
            main.wait_event(ev)
            for j, part in res:
                out[j] = part
                for t in part[2:]:
                    t.record_stream(main)            # proof tensors were allocated on side streams
        return out


def _generator():
    """G1Jacobian_generator (g1-tensor.cuh:28-63)."""
    gx = [4250078230, 1555269520, 2574712821, 2014837863, 339452353, 357537223, 4090554183, 4037962445, 568063040, 3989728972, 2651585397, 302085953]
    gy = [216474225, 3131872213, 2031680910, 2351063834, 1460086222, 3713621779, 1346392468, 1370249257, 2902481344, 236751935, 1342743146, 196886268]
    one = [196605, 1980301312, 3289120770, 3958636555, 1405573306, 1598593111, 1884444485, 2010011731, 2723605613, 1543969431, 4202751123, 368467651]
    return np.array([gx + gy + one], dtype=np.uint32)


def synthetic_mlp(dims, batch, seed=0, device="cuda"):
    """PyTorch-default U(-1/sqrt(in), 1/sqrt(in)) weights and N(0,1) inputs of the named shapes (model.py)."""
    import torch
    g = torch.Generator(device="cpu").manual_seed(seed)
    ws = [((torch.rand(i, o, generator=g) * 2 - 1) / (i ** 0.5)).float().to(device) for i, o in dims]
    x = torch.randn(batch, dims[0][0], generator=g).float().to(device)
    return ws, x
