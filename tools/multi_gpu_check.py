"""Launched with torchrun on W GPUs: fused ES and fused NSRA-ES stay bit-identical across
ranks on real NCCL, and match a single-GPU run of the same problem on rank 0's device."""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import estorch_b200 as E
from test_api_cpu import MLP

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
g = torch.Generator().manual_seed(3)
for name, cls, dims, kw, akw in (("ES", E.ES, [128, 512, 288], {}, {}),
                                 ("NSRA_ES", E.NSRA_ES, [24, 64, 64, 4], {"weight_t": 2}, {"bc_obs": 64, "bc_dim": 256})):
    obs, tgt = torch.randn(256, dims[0], generator=g), torch.randn(256, dims[-1], generator=g)
    torch.manual_seed(5); np.random.seed(11)

    class Q(cls):
        def log(self):
            pass
    es = Q(MLP, E.DeviceAgent, torch.optim.Adam, population_size=256, sigma=0.02, policy_kwargs={"dims": dims},
           agent_kwargs=dict(obs=obs, target=tgt, **akw), optimizer_kwargs={"lr": 0.01}, noise_table_size=1 << 22,
           log_interval=int(os.environ.get("LOG_INTERVAL", "1")), **kw)   # > 1: the post-update rollout is folded
    assert es._fused and es.n_workers == world
    es.train(n_steps=3)
    theta = torch.stack([s.theta for s in es._slots])
    ret = torch.from_numpy(es.population_returns).to(theta.device)
    for t in (theta, ret):
        ref = t.clone(); dist.broadcast(ref, src=0)
        assert torch.equal(ref, t), f"{name}: rank {rank} diverged from rank 0"
    assert torch.isfinite(theta).all() and torch.isfinite(ret).all()
    if rank == 0:
        print(f"{name}: {world} ranks bit-identical after 3 generations; precision={es._precision}; "
              f"episode {es.episode_reward:.5f}", flush=True)
dist.destroy_process_group()
