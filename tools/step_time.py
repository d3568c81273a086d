"""ms per generation of the fused ES path under different host modes (graph replay on/off, deferred rollout on/off)."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import MLP, synthetic_batch, WORKLOADS
from estorch_b200 import ES, DeviceAgent
wl = WORKLOADS["north_star"]
obs, tgt = synthetic_batch(wl["dims"], wl["batch"])
class Q(ES):
    def log(self):
        pass
for graph in ("1", "0"):
    for li in (10 ** 9, 1):
        os.environ["ESTORCH_B200_GRAPH"] = graph
        torch.manual_seed(0)
        es = Q(MLP, DeviceAgent, torch.optim.Adam, population_size=wl["population_size"], sigma=wl["sigma"],
               policy_kwargs={"dims": wl["dims"]}, agent_kwargs=dict(obs=obs, target=tgt), optimizer_kwargs={"lr": 0.01},
               noise_table_size=1 << 26, log_interval=li)
        es.train(5)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); t0 = time.perf_counter()
        es.train(40)
        e1.record(); torch.cuda.synchronize()
        print(f"graph={graph} log_interval={li}: {e0.elapsed_time(e1) / 40:.3f} ms/generation (host {1e3 * (time.perf_counter() - t0) / 40:.3f}), "
              f"graphs cached {len(es.__dict__.get('_graphs', {}))}, launches {es._be.launches}", flush=True)
        del es
