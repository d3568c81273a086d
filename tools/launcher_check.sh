#!/bin/bash
# train(n_proc=2) from a plain `python script.py` (the reference's usage, estorch.py:272-308): needs 2 GPUs.
out=gpurun_out; mkdir -p $out; tmp=$(mktemp -d)
cat > $tmp/user_script.py <<PY
import os, sys, numpy as np, torch
sys.path.insert(0, "$PWD"); sys.path.insert(0, "$PWD/tests")
import estorch_b200 as E
from test_api_cpu import MLP
g = torch.Generator().manual_seed(1)
obs, tgt = torch.randn(256, 128, generator=g), torch.randn(256, 288, generator=g)
class Q(E.ES):
    def log(self):
        print(f"rank {self.rank} log step {self.step} episode {self.episode_reward:.5f}", flush=True)
es = Q(MLP, E.DeviceAgent, torch.optim.Adam, population_size=256, sigma=0.02, policy_kwargs={"dims": [128, 512, 288]},
       agent_kwargs=dict(obs=obs, target=tgt), optimizer_kwargs={"lr": 0.01}, noise_table_size=1 << 22)
print("constructed", os.environ.get("RANK"), flush=True)
es.train(n_steps=6, n_proc=2)
torch.cuda.synchronize()
np.save("$tmp/theta_rank%d.npy" % es.rank, es._slots[0].theta.cpu().numpy())
print("done rank", es.rank, flush=True)
PY
: > $out/r02_launcher_check.log
for graph in 1 0; do
  echo "== ESTORCH_B200_GRAPH=$graph" >> $out/r02_launcher_check.log
  s=$(date +%s)
  ESTORCH_B200_GRAPH=$graph timeout 90 python $tmp/user_script.py >> $out/r02_launcher_check.log 2>&1
  echo "exit $? after $(( $(date +%s) - s )) s" >> $out/r02_launcher_check.log
  python - <<PY >> $out/r02_launcher_check.log 2>&1
import numpy as np
a, b = np.load("$tmp/theta_rank0.npy"), np.load("$tmp/theta_rank1.npy")
print("ranks agree:", bool(np.array_equal(a, b)))
PY
done
grep -v "^W\|^\[W\|OMP_NUM\|^\*\*\*" $out/r02_launcher_check.log | tail -30
