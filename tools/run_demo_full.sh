#!/bin/bash
# The drop-in ./demo (zkdl_b200/host/demo) on the full 18.2M-param model, batch 256, for 0 (blocking loop) / 4 / 8 / 15 streams (cold first pass and
# warm last pass of ZKDL_DEMO_REPS=3), and with "ref" the reference ./demo (oracle/_ref/demo) on the same files + both demo.out hashes.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
W=$(mktemp -d); cd "$W"
python - <<'PY'
import torch, torch.nn as nn
torch.manual_seed(0)
def save_tensor(t, fn):
    m = nn.Module(); m.register_parameter("0", nn.Parameter(t)); torch.jit.script(m).save(fn)
d = [784, 1000, 1773, 1773, 1773, 1773, 1773, 1124, 1000]
layers = []
for i in range(8):
    layers.append(nn.Linear(d[i], d[i + 1], bias=False))
    if i < 7: layers.append(nn.ReLU())
model = nn.Sequential(*layers).to("cuda").eval()
x = torch.randn(256, 784).to("cuda")
save_tensor(x, "sample_input.pt")
torch.jit.trace(model, x[:1]).save("traced_model.pt")
PY
for t in 0 4 8 15; do
  SECONDS=0; echo -n "{\"streams\": $t, \"out\": \""; ZKDL_DEMO_STREAMS=$t ZKDL_DEMO_REPS=3 $ROOT/zkdl_b200/host/demo traced_model.pt sample_input.pt 2>&1 | tr "\n" " "; echo "\", \"wall_s\": $SECONDS}"
done
echo "{\"demo_out_sha256_ours\": \"$(sha256sum demo.out | cut -c1-64)\"}"
if [ "$1" == "ref" ]; then
  LD_LIBRARY_PATH=$(python -c 'import torch,os;print(os.path.dirname(torch.__file__))')/lib:$LD_LIBRARY_PATH $ROOT/oracle/_ref/demo traced_model.pt sample_input.pt 2>&1 | tr '\n' ' '; echo
  echo "{\"demo_out_sha256_reference\": \"$(sha256sum demo.out | cut -c1-64)\"}"
fi
