"""Time the rank + gradient kernel alone (CUDA events, us per launch) in the geometry every GPU count gives it:
world W -> rank-major returns, pairs/W local pairs.  n_small isolates phase A + launch + grid.sync (phase B ~ 0)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from estorch_b200.backend import CudaBackend, adam_desc
be = CudaBackend(torch.device("cuda", 0))
P, pairs = 4096, 2048
table = be.alloc(1 << 28); be.fill_noise_table(table, 42)
tb16 = be.alloc(table.numel(), dtype=torch.float16); assert be.shadow_f16(table, tb16) == 0
torch.manual_seed(0)
ret = torch.randn(P, device=be.device)
nov = torch.randn(P, device=be.device)
tag = sys.argv[1] if len(sys.argv) > 1 else "lib"


def timed(run, iters=50):
    for _ in range(5):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        run()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


for n in (1001760, 4096):
    theta, m, v = (torch.zeros(n, device=be.device) for _ in range(3))
    grad = be.alloc(n); ranks = be.alloc(P, dtype=torch.int32); ranks2 = be.alloc(P, dtype=torch.int32)
    state = be.zeros(32, dtype=torch.uint8)
    ad = adam_desc(lr=0.01, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, clamp=1.0)
    out = []
    for W in (1, 2, 4, 8):
        pl = pairs // W
        offs = be.alloc(pl, dtype=torch.int64); order = be.alloc(pl, dtype=torch.int32)
        be.make_offsets(42, None, 0, 0, pl, table.numel(), n, offs, order)
        if W == 1:
            us = timed(lambda: be.rank_grad_adam(ret, None, 1.0, 0.0, P, tb16, offs, order, theta, m, v, state, ad, ranks, None, grad))
            out.append(f"W=1 rank+grad+Adam {us:.1f}")
            us = timed(lambda: be.rank_grad_adam(ret, nov, 0.5, 0.5, P, tb16, offs, order, theta, m, v, state, ad, ranks, ranks2, grad))
            out.append(f"(+novelty column {us:.1f})")
        us = timed(lambda: be.rank_grad(ret, None, 1.0, 0.0, P, tb16, offs, order, 0, pl, n, grad, ranks, None, world=W))
        out.append(f"W={W} rank+partial {us:.1f}")
    us = timed(lambda: be.clamp_adam(grad, P, theta, m, v, state, ad, None))
    out.append(f"clamp+Adam {us:.1f}")
    print(f"{tag}: n={n}: " + "  ".join(out) + "  (us per launch)", flush=True)
