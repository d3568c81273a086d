"""ms per generation of the fused ES path under different host modes (graph replay on/off, deferred rollout
on/off), in blocks of 25 generations so that drift (clock throttling) shows; SM clock sampled per block."""
import os, subprocess, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import MLP, synthetic_batch, WORKLOADS
from estorch_b200 import ES, DeviceAgent
wl = WORKLOADS["north_star"]
obs, tgt = synthetic_batch(wl["dims"], wl["batch"])
log2 = int(sys.argv[1]) if len(sys.argv) > 1 else 28
blocks = int(sys.argv[2]) if len(sys.argv) > 2 else 6
def clock():
    try:
        return subprocess.run(["nvidia-smi", "--id=0", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits"],
                              capture_output=True, text=True, timeout=5).stdout.strip()
    except Exception:
        return "?"
class Q(ES):
    def log(self):
        pass
for graph in ("1", "0"):
    for li in (10 ** 9, 1):
        os.environ["ESTORCH_B200_GRAPH"] = graph
        torch.manual_seed(0)
        es = Q(MLP, DeviceAgent, torch.optim.Adam, population_size=wl["population_size"], sigma=wl["sigma"],
               policy_kwargs={"dims": wl["dims"]}, agent_kwargs=dict(obs=obs, target=tgt), optimizer_kwargs={"lr": 0.01},
               noise_table_size=1 << log2, log_interval=li)
        es.train(5)
        torch.cuda.synchronize()
        out = []
        for b in range(blocks):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            es.train(25)
            e1.record(); torch.cuda.synchronize()
            out.append(f"{e0.elapsed_time(e1) / 25:.3f}@{clock().replace(', ', '/')}")
        print(f"graph={graph} log_interval={li} table=2^{log2}: ms/generation per block of 25 = {' '.join(out)}; "
              f"graphs cached {sum(isinstance(v, tuple) for v in es.__dict__.get('_graphs', {}).values())}", flush=True)
        del es
        torch.cuda.empty_cache()
