#!/usr/bin/env python
"""bench.py -- generations/sec of the ES hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one ES generation (estorch.py:214-248): evaluate the P = 4096
mirrored members of the 1M-parameter MLP over a synthetic observation batch,
centred-rank the returns, reduce sum_j w_j * noise_j, negate/clamp, Adam,
post-update rollout.  Prints ONE JSON line on rank 0.

  value  device-resident throughput: inputs already in HBM, no host
         synchronisation inside the timed region (CUDA events, max over ranks)
  e2e    the same through the public API with HOST buffers: every generation
         uploads the observation/target batch from pinned host memory and reads
         population_returns + episode_reward back (host clock around K steps)
  roofline / kernels   per-kernel achieved rate vs MEASURED_PEAKS.json
  cpu_baseline         the reference's CPU algorithm (oracle/reference_port.py)
                       on this box's host cores, bounded sample, rank 0 / N=1

`--impl reference` runs only the CPU port (rank 0), same metric/config.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json metric: pop=4096, 1M-param MLP (obs 128, 4x512 hidden, 288 out)
    "north_star": dict(dims=[128, 512, 512, 512, 512, 288], population_size=4096, sigma=0.02, batch=256),
    # configs[1]: CartPole-shape 2x64 MLP, pop 4096
    "cartpole": dict(dims=[4, 64, 64, 2], population_size=4096, sigma=0.1, batch=256),
}


class MLP(torch.nn.Module):
    def __init__(self, dims):
        super().__init__()
        layers = []
        for i in range(len(dims) - 1):
            layers.append(torch.nn.Linear(dims[i], dims[i + 1]))
            if i + 2 < len(dims):
                layers.append(torch.nn.ReLU())
        self.net = torch.nn.Sequential(*layers)

    def forward(self, x):
        return self.net(x)


def n_params(dims):
    return sum(dims[i] * dims[i + 1] + dims[i + 1] for i in range(len(dims) - 1))


def synthetic_batch(dims, batch):
    g = torch.Generator().manual_seed(1234)
    return torch.randn(batch, dims[0], generator=g), torch.randn(batch, dims[-1], generator=g)


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p["hbm_gbs"], tensor_burst=p["bf16_tflops"], tensor=p["bf16_tflops_sustained"],
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tensor_burst=1590.0, tensor=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def window(self, t0, t1):
        rows = [r for t, r in self.rows if t0 <= t <= t1] or [r for _, r in self.rows[-3:]]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": float(rows[0][1]) if rows[0][1].replace(".", "").isdigit() else None,
                "power_w_max": max((float(r[2]) for r in rows if r[2].replace(".", "").isdigit()), default=None),
                "samples": len(rows), "reasons": reasons}

    def stop(self):
        if self.proc:
            self.proc.terminate()


# ----------------------------------------------------------------------------- CPU reference arm
def cpu_reference(wl, steps, warmup, budget_s=20.0):
    """Time the reference's CPU algorithm on a bounded population sample and scale
    linearly in P (sample, cat, rollouts, mm are all O(P*n))."""
    from oracle.reference_port import time_reference
    from estorch_b200.agents import DeviceAgent
    dims, P, sigma = wl["dims"], wl["population_size"], wl["sigma"]
    obs, tgt = synthetic_batch(dims, wl["batch"])
    agent = DeviceAgent(obs, tgt)
    torch.manual_seed(0)
    probe_P = 16
    dt, _ = time_reference(MLP, {"dims": dims}, agent, probe_P, sigma, steps=1)
    per_member = dt / probe_P
    sample_P = int(min(P, max(16, budget_s / max(1, steps + warmup) / per_member)))
    sample_P -= sample_P % 2
    dt, phases = time_reference(MLP, {"dims": dims}, agent, sample_P, sigma, steps=steps, warmup=warmup)
    scale = P / sample_P
    return dict(seconds_per_generation_full=dt * scale, value=1.0 / (dt * scale), sample_P=sample_P,
                seconds_per_generation_sample=dt, phases_sample_s=phases, cores=torch.get_num_threads())


def run_reference_arm(args, wl, rank, world):
    if rank != 0:
        return
    r = cpu_reference(wl, args.steps, min(args.warmup, 1), budget_s=60.0)
    cb = {"value": r["value"], "unit": "generations/s", "cores": r["cores"], "kind": "port",
          "sample": f"population_size={r['sample_P']} of {wl['population_size']} per step, same policy/batch; "
                    f"time scaled linearly in P (all phases are O(P*n)); 1 process, {r['cores']} torch threads "
                    f"of {os.cpu_count()} cpus; phases(s/sample-step)="
                    f"{ {k: round(v, 4) for k, v in r['phases_sample_s'].items()} }"}
    line = {"impl": "reference", "metric": "generations/sec at pop=4096, 1M-param MLP", "value": r["value"],
            "unit": "generations/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * r["seconds_per_generation_full"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, wl, world), "cpu_baseline": cb,
            "e2e": {"value": r["value"], "unit": "generations/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0}}
    emit(line)


def workload_config(args, wl, world):
    return {"workload": f"{args.workload}: ES generation, population_size={wl['population_size']} "
                        f"({wl['population_size'] // 2} antithetic pairs), MLP {wl['dims']} "
                        f"(n={n_params(wl['dims'])}), synthetic obs batch B={wl['batch']}, sigma={wl['sigma']}, "
                        f"Adam lr=0.01, noise table 2^{args.table_log2} fp32",
            "population_size": wl["population_size"], "n_parameters": n_params(wl["dims"]),
            "batch": wl["batch"], "parallelism": f"pairs sharded over {world} GPU(s)",
            "l2": "noise stream per step (>= 8 GB at north_star) exceeds L2; no flush needed"}


# ----------------------------------------------------------------------------- our arm
def run_ours(args, wl, rank, world, local_rank):
    import torch.distributed as dist
    from estorch_b200 import ES, DeviceAgent
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
    dims, P, sigma, B = wl["dims"], wl["population_size"], wl["sigma"], wl["batch"]
    n, pairs = n_params(dims), P // 2
    obs, tgt = synthetic_batch(dims, B)

    class Streaming(DeviceAgent):
        """e2e: each generation's batch comes from pinned host memory."""
        stream_from_host = False

        def __init__(self, obs, target):
            super().__init__(obs, target)
            self.h_obs, self.h_tgt = self.obs.clone().pin_memory(), self.target.clone().pin_memory()

        def next_batch(self, step):
            return (self.h_obs, self.h_tgt) if self.stream_from_host else None

    class Bench(ES):
        read_back = False

        def log(self):
            if self.read_back:   # the step's result, on the host (D2H of P returns + 32-byte state)
                self.last = (self.population_returns, self.episode_reward)

    torch.manual_seed(0)
    es = Bench(MLP, Streaming, torch.optim.Adam, population_size=P, sigma=sigma,
               policy_kwargs={"dims": dims}, agent_kwargs=dict(obs=obs, target=tgt),
               optimizer_kwargs={"lr": 0.01}, noise_table_size=1 << args.table_log2, noise_seed=42,
               log_interval=10 ** 9, eval_precision=args.eval_precision)
    assert es._fused, "bench: fused device path is not active"
    be = es._be

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- value: device-resident, no host sync in the loop
    sampler = ClockSampler(local_rank) if rank == 0 else None   # nvidia-smi needs ~1 s to start streaming
    es.train(args.warmup)
    barrier()
    launches0 = be.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ev0.record()
    es.train(args.steps)
    ev1.record()
    barrier()
    t1 = time.perf_counter()
    ms = max_over_ranks(ev0.elapsed_time(ev1))
    launches = be.launches - launches0
    clocks = sampler.window(t0, t1) if sampler else None

    # ---- e2e: host buffers in, host results out, every generation
    es.agent.stream_from_host, es.read_back, es._log_interval = True, True, 1
    es.train(min(args.warmup, 3))
    barrier()
    t0 = time.perf_counter()
    es.train(args.steps)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    if sampler:
        sampler.stop()

    if rank != 0:
        return
    peaks = load_peaks()
    # ---- per-kernel device time (CUDA events on the launching stream), N=1 geometry of this rank
    pl = es._pairs_local
    slot = es._slots[0]

    def timed(fn, iters=5):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
        for a, b in evs:
            a.record(); fn(); b.record()
        torch.cuda.synchronize()
        return sorted(a.elapsed_time(b) for a, b in evs)[iters // 2]

    from estorch_b200.backend import adam_desc
    R = es._returns if world == 1 else es._rm_buffers()[0]
    gt = es._grad_table()                       # the exact fp16 copy of the table when there is one
    tbytes = gt.element_size()
    rp, rm_ = (R[es._pair_begin: es._pair_begin + pl], R[pairs + es._pair_begin: pairs + es._pair_begin + pl]) \
        if world == 1 else (R[rank, 0], R[rank, 1])
    t_eval = timed(lambda: be.eval_mlp(es._spec.dims, slot.theta, es._table, es._offsets, es._order, pl, sigma,
                                       es._obs, es._tgt, rp, rm_, **es._eval_kw(slot)), iters=3)
    scratch = [t.clone() for t in (slot.theta, slot.m, slot.v)]
    ad = adam_desc(lr=0.01)
    if world == 1:
        st_scratch = slot.state.clone()
        t_grad = timed(lambda: be.rank_grad_adam(R, None, 1.0, 0.0, P, gt, es._offsets, es._order,
                                                 scratch[0], scratch[1], scratch[2], st_scratch, ad,
                                                 es._ranks, None, None))
    else:
        rmaj = gt.dtype == torch.float16
        t_grad = timed(lambda: be.rank_grad(R.view(-1) if rmaj else es._returns, None, 1.0, 0.0, P, gt, es._offsets,
                                            es._order, es._pair_begin, pl, n, es._grad, es._ranks, None,
                                            world=world if rmaj else 1))
    # algorithmic bytes per launch on this rank (SURVEY 8d with the table's element size: the engine's
    # table entries are fp16-representable and both kernels stream the exact 16-bit copy)
    bytes_grad = tbytes * n * pl + 28 * n + 8 * P
    bytes_eval = tbytes * n * pl + 4 * n + 4 * B * (dims[0] + dims[-1]) + 4 * P
    flops_eval = 2.0 * n * B * 2 * pl
    k_grad = {"kernel": "rank_grad_adam" if world == 1 else "rank_grad", "bound": "hbm",
              "achieved": bytes_grad / t_grad / 1e6, "peak": peaks["hbm"], "unit": "GB/s",
              "frac": bytes_grad / t_grad / 1e6 / peaks["hbm"], "ms": t_grad, "traffic": None,
              "algorithmic_bytes": bytes_grad,
              "note": "frac > 1 is L2 reuse: the rows of one generation overlap in the 1 GiB table and every CTA "
                      "walks the pairs in offset-sorted order, so the table is fetched from HBM once (ncu: 1.08 GB "
                      "DRAM read per launch) and the other 7/8 of the algorithmic bytes are L2 hits"}
    k_eval = {"kernel": "eval_mlp_" + es._precision, "bound": "tensor", "achieved": flops_eval / t_eval / 1e9,
              "peak": peaks["tensor"], "unit": "TFLOP/s", "frac": flops_eval / t_eval / 1e9 / peaks["tensor"],
              "ms": t_eval, "traffic": None, "algorithmic_bytes": bytes_eval, "flops": flops_eval,
              "hbm_frac_of_noise_stream": bytes_eval / t_eval / 1e6 / peaks["hbm"]}
    # DRAM traffic per launch (dram__bytes_read.sum + dram__bytes_write.sum) from the committed
    # `ncu --set full` capture of exactly these kernels / shapes on one GPU
    # (profiles/ncu_traffic.json <- profiles/r01s2_ncu_full_summary.txt); null elsewhere
    if args.workload == "north_star" and world == 1:
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
            k_grad["traffic"] = tr.get(k_grad["kernel"])
            k_eval["traffic"] = tr.get(k_eval["kernel"])
        except (OSError, ValueError):
            pass
    dominant = k_eval if t_eval >= t_grad else k_grad
    roofline = {k: dominant[k] for k in ("bound", "achieved", "peak", "unit", "frac", "traffic")}
    roofline["kernel"] = dominant["kernel"]
    roofline["peak_source"] = peaks["source"]

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_reference(wl, steps=1, warmup=0, budget_s=15.0)
        cpu = {"value": r["value"], "unit": "generations/s", "cores": r["cores"], "kind": "port",
               "sample": f"1 generation at population_size={r['sample_P']} of {P}, scaled linearly in P; "
                         f"{r['cores']} torch threads of {os.cpu_count()} cpus; "
                         f"{r['seconds_per_generation_sample']:.2f} s/sample-generation"}
    line = {"metric": "generations/sec at pop=4096, 1M-param MLP", "value": args.steps / (ms / 1e3),
            "unit": "generations/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None,
            "dtype": {"f16": "f32 (fp32-equivalent: evaluate GEMMs on tcgen05 with fp16 operands = 11-bit significand, "
                             "TF32 class, each weight formed in fp32 from fp32 theta + the exactly-16-bit noise value and "
                             "rounded once, f32 accumulate, f32 bias/loss; f32 ranks, reduction, Adam)",
                      "bf16": "bf16 operands / f32 accumulate (evaluate GEMMs); f32 noise, ranks, reduction, Adam",
                      "bf16s": "bf16 operands formed from bf16 shadows of theta/noise, f32 accumulate (evaluate "
                               "GEMMs); f32 noise table, ranks, reduction, Adam"}.get(es._precision, "f32"),
            "data": "synthetic", "config": workload_config(args, wl, world),
            "e2e": {"value": args.steps / e2e_s, "unit": "generations/s",
                    "h2d_bytes_per_step": int(obs.numel() * 4 + tgt.numel() * 4),
                    "d2h_bytes_per_step": int(4 * P + 32)},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "kernels": [k_grad, k_eval],
            "cpu_baseline": cpu}
    emit(line)


_JSON_FD = None


def emit(line):
    """The ONE JSON line, on the process's real stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="north_star", choices=sorted(WORKLOADS))
    ap.add_argument("--table-log2", type=int, default=28)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--eval-precision", default="auto", choices=["auto", "fp32", "f16", "bf16", "bf16s"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    wl = WORKLOADS[args.workload]
    # stdout carries exactly one JSON line: anything a library prints there (NCCL's version
    # banner at communicator creation, ...) is sent to stderr instead
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference_arm(args, wl, rank, world)
        return
    if world != args.gpus:
        if args.gpus > 1:
            raise SystemExit(f"--gpus {args.gpus} needs torchrun with {args.gpus} processes (WORLD_SIZE={world})")
    run_ours(args, wl, rank, world, local_rank)
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
