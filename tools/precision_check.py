"""Accuracy of the tensor-core evaluate modes vs the exact fp32 kernel at the north-star size."""
import os, sys, json, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from estorch_b200.backend import CudaBackend
be = CudaBackend(torch.device("cuda", 0))
dims = [128, 512, 512, 512, 512, 288]
n = sum(dims[i] * dims[i + 1] + dims[i + 1] for i in range(len(dims) - 1))
P = 4096; pairs = P // 2
torch.manual_seed(0)
mods = []
for i in range(len(dims) - 1):
    l = torch.nn.Linear(dims[i], dims[i + 1]); mods += [l.weight.detach().reshape(-1), l.bias.detach()]
theta = torch.cat(mods).to(be.device).contiguous()
g = torch.Generator().manual_seed(1234)
obs, tgt = torch.randn(256, 128, generator=g).to(be.device), torch.randn(256, 288, generator=g).to(be.device)
table = be.alloc(1 << 28); be.fill_noise_table(table, 42)
offs = be.alloc(pairs, dtype=torch.int64); order = be.alloc(pairs, dtype=torch.int32)
be.make_offsets(42, None, 0, 0, pairs, table.numel(), n, offs, order)
th16 = be.alloc(n, dtype=torch.bfloat16); tb16 = be.alloc(table.numel(), dtype=torch.bfloat16)
be.shadow_bf16(theta, th16); be.shadow_bf16(table, tb16)
res = {}
for mode in ("fp32", "bf16", "bf16s"):
    r = be.zeros(P)
    be.eval_mlp(dims, theta, table, offs, order, pairs, 0.02, obs, tgt, r[:pairs], r[pairs:], precision=mode,
                theta16=th16, table16=tb16)
    res[mode] = r.double().cpu().numpy()
ref = res["fp32"]
def ranks(x):
    o = np.argsort(x, kind="stable"); rk = np.empty_like(o); rk[o] = np.arange(len(x)); return rk
for mode in ("bf16", "bf16s"):
    x = res[mode]
    rr = np.corrcoef(ranks(ref), ranks(x))[0, 1]
    print(json.dumps({"mode": mode, "max_rel_err_returns": float(np.max(np.abs(x - ref)) / np.max(np.abs(ref))),
                      "spread_of_returns": float(ref.std() / abs(ref.mean())),
                      "err_over_spread": float(np.abs(x - ref).max() / ref.std()),
                      "spearman_rank_corr": float(rr),
                      "centred_rank_weight_rel_l2": float(np.linalg.norm(ranks(ref) - ranks(x)) / np.linalg.norm(ranks(ref) - (P - 1) / 2))}))
