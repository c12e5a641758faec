"""A handful of launches of the roofline kernels at bench sizes, for `ncu --set full` (keep the capture small)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from zkdl_b200 import capi as zk, mlp
n = 1 << 24
W = torch.randint(-(2 ** 31), 2 ** 31 - 1, (n, 8), dtype=torch.int32, device="cuda"); W[:, 7] &= 0x3FFFFFFF
for _ in range(2):
    zk.fr_partial_me(W, zk.random_vec(1, 3), 1)                 # k_fr_fold_multi<3> on 512 MiB
del W
a = torch.randint(-(2 ** 31), 2 ** 31 - 1, (1 << 22, 8), dtype=torch.int32, device="cuda"); a[:, 7] &= 0x3FFFFFFF
zk.bin_sumcheck(a, zk.random_vec(2, 22), zk.random_vec(3, 22))  # k_sc_round<2> from 2^22 entries
ng = 2048
G = zk.g1_mul(zk.to_device(mlp._generator()), zk.to_device(zk.random_vec(5, ng)))
tab = zk.G1Table(G, full=True)
t = zk.to_device(zk.random_vec(6, ng))
zk.me_open(tab, t, zk.random_vec(7, 11))                        # k_msm_accumulate of one opening (34 rows x 2048)
torch.cuda.synchronize()
print("ok")
