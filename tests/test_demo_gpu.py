"""End-to-end drop-in check (SURVEY.md §4 'End-to-end'): ./demo traced_model.pt sample_input.pt of the new build must
write a demo.out byte-identical to the reference's own demo (oracle/_ref/demo, built from /root/reference for sm_100)
on the same TorchScript files, print the same three stdout lines, and its dumped proof must verify the sumcheck /
opening identities."""
import os
import re
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

GEN = r'''
import sys, torch, torch.nn as nn
torch.manual_seed(int(sys.argv[1]))
dims = [int(v) for v in sys.argv[3].split(",")]
def save_tensor(t, fn):
    m = nn.Module(); m.register_parameter("0", nn.Parameter(t)); torch.jit.script(m).save(fn)
layers = []
for i in range(len(dims) - 1):
    layers.append(nn.Linear(dims[i], dims[i + 1], bias=False))
    if i < len(dims) - 2: layers.append(nn.ReLU())
model = nn.Sequential(*layers).to("cuda").eval()
x = torch.randn(int(sys.argv[2]), dims[0]).to("cuda")
save_tensor(x, "sample_input.pt")
torch.jit.trace(model, x[:1]).save("traced_model.pt")
'''


FULL = "784,1000,1773,1773,1773,1773,1773,1124,1000"      # model.py:14-30, the 18.2 M-parameter ModulusLab MLP (configs 1 and 4)


@pytest.mark.parametrize("batch,dims", [(5, "20,33,17,10"), (1, "16,64,8"), (256, FULL), (1, FULL)],
                         ids=["toy-b5", "toy-b1", "model.py-b256", "model.py-b1"])
def test_demo_out_identical_to_reference(tmp_path, batch, dims):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    ours = os.path.join(ROOT, "zkdl_b200", "host", "demo")
    ref = os.path.join(ROOT, "oracle", "_ref", "demo")
    if not os.path.exists(ours):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "zkdl_b200", "host"), "-j3"])
    if not os.path.exists(ref):
        pytest.skip("reference build (oracle/_ref/demo) not present")
    subprocess.check_call([sys.executable, "-c", GEN, "7", str(batch), dims], cwd=tmp_path)
    d_ref, d_our = tmp_path / "ref", tmp_path / "our"
    d_ref.mkdir(); d_our.mkdir()
    r = subprocess.run([ref, "../traced_model.pt", "../sample_input.pt"], cwd=d_ref, capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stderr[-500:]
    env = dict(os.environ, ZKDL_SEED="1234", ZKDL_DUMP_PROOF="proof.txt")
    o = subprocess.run([ours, "../traced_model.pt", "../sample_input.pt"], cwd=d_our, capture_output=True, text=True, timeout=600, env=env)
    assert o.returncode == 0, o.stderr[-500:]
    assert (d_ref / "demo.out").read_bytes() == (d_our / "demo.out").read_bytes()
    for pat in (r"Total number of parameters: (\d+)", r"Proof time: [0-9.e+-]+ seconds per data point\.", r"Current CUDA status: 0"):
        assert re.search(pat, r.stdout) and re.search(pat, o.stdout), (pat, r.stdout, o.stdout)
    assert re.search(r"parameters: (\d+)", r.stdout).group(1) == re.search(r"parameters: (\d+)", o.stdout).group(1)
    proof = (d_our / "proof.txt").read_text().split("\n")
    assert sum(1 for l in proof if l.startswith("fc ")) == dims.count(",") and any(l.startswith("relu ") for l in proof)
