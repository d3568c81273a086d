"""Generations/s through the public API for the other BASELINE configs (1 GPU):
config 4 NSRA-ES (BipedalWalker-shape MLP 24-64-64-4, population 2048, k=10, M=3, 256-D BC)
config 5 conv + VirtualBatchNorm policy (84x84x4), population 1024, 128 reference frames."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import estorch_b200 as E
from test_api_cpu import MLP
from test_api_gpu import AtariPolicy

def timed_train(es, steps, warm):
    es.train(warm); torch.cuda.synchronize()
    t0 = time.perf_counter(); es.train(steps); torch.cuda.synchronize()
    return steps / (time.perf_counter() - t0)

g = torch.Generator().manual_seed(0)
which = sys.argv[1] if len(sys.argv) > 1 else "both"
if which in ("both", "nsra"):
    obs, tgt = torch.randn(256, 24, generator=g), torch.randn(256, 4, generator=g)
    class Q(E.NSRA_ES):
        def log(self): pass
    es = Q(MLP, E.DeviceAgent, torch.optim.Adam, population_size=2048, sigma=0.02, weight_t=10,
           policy_kwargs={"dims": [24, 64, 64, 4]}, agent_kwargs=dict(obs=obs, target=tgt, bc_obs=64, bc_dim=256),
           optimizer_kwargs={"lr": 0.01}, noise_table_size=1 << 26)
    r = timed_train(es, 100, 5)
    print(json.dumps({"config": "NSRA-ES bipedal-shape MLP pop=2048 (BASELINE config 4), 1 GPU", "fused": es._fused,
                      "generations_per_s": r, "archive": len(es._archive), "weight": es.weight}), flush=True)
if which in ("both", "conv"):
    xref = torch.rand(128, 4, 84, 84, generator=g)
    obs, tgt = torch.rand(32, 4, 84, 84, generator=g), torch.randn(32, 4, generator=g)
    class C(E.ES):
        def log(self): pass
    es = C(AtariPolicy, E.DeviceAgent, torch.optim.Adam, population_size=1024, sigma=0.02,
           policy_kwargs=dict(n_actions=4, xref=xref), agent_kwargs=dict(obs=obs, target=tgt),
           optimizer_kwargs={"lr": 0.01}, noise_table_size=1 << 26)
    r = timed_train(es, 5, 1)
    print(json.dumps({"config": "conv+VirtualBatchNorm policy pop=1024, xref 128 frames, B=32 (BASELINE config 5), 1 GPU",
                      "fused": es._fused and es._is_conv, "generations_per_s": r}), flush=True)
