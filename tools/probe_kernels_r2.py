"""A handful of launches of the hot kernels at bench sizes, for `ncu --set full` (keeps the capture small):
one commitment opening at |G| = 2048 (k_msm_accumulate / k_msm_combine<8> / k_msm_reduce_scan), the tcgen05 forward product
256 x 2048 x 2048 (k_umma_matmul), one packed zkReLU proof at n = 2^19 (k_bin_packed3 / k_bin_r34 / k_sc_round / k_sc_tail)
and the three-round fold of the FC4096 weight table (k_fr_fold_multi<3>, the `roofline` kernel of bench.py)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from zkdl_b200 import capi as zk, mlp

ng = 2048
G = zk.g1_mul(zk.to_device(mlp._generator()), zk.fr_random(ng, 5))
tab = zk.G1Table(G, full=True)
t = zk.to_device(zk.random_vec(6, ng))
for _ in range(2):
    zk.me_open(tab, t, zk.random_vec(7, 11))
I = O = 2048; B = 256
g = torch.Generator(device="cuda").manual_seed(3)
def small_fr(cnt, bits):
    v = torch.randint(-(1 << (bits - 1)), 1 << (bits - 1), (cnt, 1), generator=g, device="cuda", dtype=torch.int32).float() / 65536.0
    q = zk.float_to_fr(v, cnt, 1)
    return zk.fr_elementwise(zk.OP_MONT, q, out=q)
W, X = small_fr(I * O, 13), small_fr(B * I, 17)
mm = zk.MatmulWeights(W, I, O)
for _ in range(2):
    Z = zk.fr_matmul_prepared(X, mm, B)
A, sign, magp, remp, bad = zk.relu_packed(Z)
L = 19
ch = [zk.random_vec(11, L + 5), zk.random_vec(12, L + 5), zk.random_vec(13, L + 4), zk.random_vec(14, L + 4), zk.random_vec(15, L), zk.random_vec(16, L), zk.random_vec(17, L)]
for _ in range(2):
    zk.zkrelu_prove_packed(Z, sign, magp, remp, *ch)
torch.cuda.synchronize()
del W, X, Z, A, sign
n = 1 << 24
Wb = torch.randint(-(2 ** 31), 2 ** 31 - 1, (n, 8), dtype=torch.int32, device="cuda"); Wb[:, 7] &= 0x3FFFFFFF
for _ in range(2):
    zk.fr_partial_me(Wb, zk.random_vec(1, 3), 1)
torch.cuda.synchronize()
print("ok")
