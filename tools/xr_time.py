"""Under torchrun: us per launch of the rank + gradient kernel with the cross-GPU sum inside (NVLink peer memory)
against the same kernel + NCCL all-reduce + clamp/Adam, same geometry, back-to-back launches on every rank."""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from estorch_b200.backend import CudaBackend, adam_desc
rank, W = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
be = CudaBackend(torch.device("cuda", torch.cuda.current_device()))
P, pairs = 4096, 2048
pl = pairs // W
table = be.alloc(1 << 28); be.fill_noise_table(table, 42)
tb16 = be.alloc(table.numel(), dtype=torch.float16); assert be.shadow_f16(table, tb16) == 0
torch.manual_seed(0)
ret = torch.randn(P, device=be.device)


def timed(run, iters=100):
    for _ in range(5):
        run()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        run()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters * 1e3], device=be.device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


for n in (1001760, 4096):
    theta, m, v = (torch.zeros(n, device=be.device) for _ in range(3))
    grad = be.alloc(n); ranks = be.alloc(P, dtype=torch.int32)
    state = be.zeros(32, dtype=torch.uint8)
    ad = adam_desc(lr=0.01, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, clamp=1.0)
    offs = be.alloc(pl, dtype=torch.int64); order = be.alloc(pl, dtype=torch.int32)
    be.make_offsets(42, None, 0, rank * pl, pl, table.numel(), n, offs, order)
    mine = be.peer_alloc(be.xr_workspace_bytes(n))
    handles = [None] * W
    dist.all_gather_object(handles, mine[1])
    peers = [mine[0] if r == rank else be.peer_open(handles[r]) for r in range(W)]
    t_xr = timed(lambda: be.rank_grad_xr_adam(ret, None, 1.0, 0.0, P, W, rank, tb16, offs, order, rank * pl, pl, peers,
                                              theta, m, v, state, ad, ranks, None, grad))
    t_rg = timed(lambda: be.rank_grad(ret, None, 1.0, 0.0, P, tb16, offs, order, rank * pl, pl, n, grad, ranks, None, world=W))

    def nccl_path():
        be.rank_grad(ret, None, 1.0, 0.0, P, tb16, offs, order, rank * pl, pl, n, grad, ranks, None, world=W)
        dist.all_reduce(grad)
        be.clamp_adam(grad, P, theta, m, v, state, ad, None)
    t_nc = timed(nccl_path)
    t_ar = timed(lambda: dist.all_reduce(grad))
    if rank == 0:
        print(f"W={W} n={n}: rank+gradient+NVLink sum+Adam {t_xr:.1f} | rank+partial gradient alone {t_rg:.1f} | "
              f"rank+partial, NCCL all-reduce, clamp+Adam {t_nc:.1f} | NCCL all-reduce alone {t_ar:.1f}  (us, max over ranks)",
              flush=True)
dist.destroy_process_group()
