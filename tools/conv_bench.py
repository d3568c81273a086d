"""One timing of the conv + VirtualBatchNorm evaluate at BASELINE config-5 scale
(population 1024, reference batch 128, observation batch B)."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from estorch_b200.backend import CudaBackend
be = CudaBackend(torch.device("cuda", 0))
A, R, B, P = 4, 128, int(sys.argv[1]) if len(sys.argv) > 1 else 32, 1024
n = 677268
pairs = P // 2
table = be.alloc(1 << 26); be.fill_noise_table(table, 42)
offs = be.alloc(pairs, dtype=torch.int64); order = be.alloc(pairs, dtype=torch.int32)
be.make_offsets(42, None, 0, 0, pairs, table.numel(), n, offs, order)
theta = torch.randn(n, device=be.device) * 0.02
xref, obs, tgt = torch.rand(R, 4, 84, 84, device=be.device), torch.rand(B, 4, 84, 84, device=be.device), torch.randn(B, A, device=be.device)
scratch = torch.empty(be.conv_scratch_bytes(R, B), dtype=torch.uint8, device=be.device)
ret = be.zeros(P)
def run():
    be.eval_conv_vbn(A, theta, table, offs, order, pairs, 0.02, xref, obs, tgt, ret[:pairs], ret[pairs:], scratch)
run(); torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); run(); b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b)
macs = P * ((R + B) * (6400 * 256 + 2592 * 256) + B * (2592 * 256 + 256 * A))
print(json.dumps({"kernel": "eval_conv_vbn", "P": P, "ref_batch": R, "B": B, "ms": ms, "TFLOPs_fp32": 2 * macs / ms / 1e9}))
