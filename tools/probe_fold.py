import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from zkdl_b200 import capi as zk
n = 1 << 24
W = torch.randint(-(2 ** 31), 2 ** 31 - 1, (n, 8), dtype=torch.int32, device="cuda"); W[:, 7] &= 0x3FFFFFFF
u3 = zk.random_vec(77, 3)
for _ in range(3): zk.fr_partial_me(W, u3, 1)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): zk.fr_partial_me(W, u3, 1)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"block={os.environ.get('ZKDL_FOLD_BLOCK','256')} cap={os.environ.get('ZKDL_FOLD_CAP','8')}: {ms*1e3:.1f} us, algorithmic {84*n/ms/1e6:.0f} GB/s = {84*n/ms/1e6/6530.3:.3f} of HBM peak")
