#!/usr/bin/env python
"""Small driver for ncu: a few launches of each hot kernel at the north-star shape."""
import argparse, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from estorch_b200.backend import CudaBackend, new_state, adam_desc

ap = argparse.ArgumentParser()
ap.add_argument("--pairs", type=int, default=2048)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--what", default="both")
a = ap.parse_args()
be = CudaBackend(torch.device("cuda", 0))
dims = [128, 512, 512, 512, 512, 288]
n = sum(dims[i] * dims[i + 1] + dims[i + 1] for i in range(len(dims) - 1))
P, pairs, B = 2 * a.pairs, a.pairs, 256
table = be.alloc(1 << 28); be.fill_noise_table(table, 42)
offs = be.alloc(pairs, dtype=torch.int64); order = be.alloc(pairs, dtype=torch.int32)
be.make_offsets(42, None, 0, 0, pairs, table.numel(), n, offs, order)
theta, m, v = torch.randn(n, device=be.device) * 0.05, be.zeros(n), be.zeros(n)
obs, tgt = torch.randn(B, dims[0], device=be.device), torch.randn(B, dims[-1], device=be.device)
ret = be.zeros(P); st = new_state(be.device); ranks = be.zeros(P, dtype=torch.int32)
tb16 = be.alloc(table.numel(), dtype=torch.float16); assert be.shadow_f16(table, tb16) == 0
for _ in range(a.iters):
    if a.what in ("both", "eval"):
        be.eval_mlp(dims, theta, table, offs, order, pairs, 0.02, obs, tgt, ret[:pairs], ret[pairs:], precision="f16",
                    table16=tb16)
    if a.what in ("both", "grad"):
        be.rank_grad_adam(ret, None, 1.0, 0.0, P, tb16, offs, order, theta, m, v, st, adam_desc(lr=0.01), ranks, None, None)
torch.cuda.synchronize()
print("done")
