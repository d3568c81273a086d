#!/usr/bin/env python
"""Kernel micro-benchmarks (CUDA events on the launching stream, warm-up, L2 flush
by construction: every launch streams >> 126 MB).  Writes JSON lines to stdout."""
import argparse
import json
import sys
import os
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from estorch_b200.backend import CudaBackend, new_state, adam_desc  # noqa: E402


def timeit(fn, iters=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in evs:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2], ts[0]


def mlp_n(dims):
    return sum(dims[i] * dims[i + 1] + dims[i + 1] for i in range(len(dims) - 1))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--table-log2", type=int, default=28)
    ap.add_argument("--skip-big-eval", action="store_true")
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    be = CudaBackend(torch.device("cuda", 0))
    peak = 6569.6
    table_len = 1 << args.table_log2
    table = be.alloc(table_len)
    be.fill_noise_table(table, 42)
    torch.cuda.synchronize()

    for name, dims, P, B in (("north_star_1M", [128, 512, 512, 512, 512, 288], 4096, 256),
                             ("cartpole", [4, 64, 64, 2], 4096, 256),
                             ("bipedal", [24, 64, 64, 4], 2048, 256)):
        if args.only and name != args.only:
            continue
        n, pairs = mlp_n(dims), P // 2
        offs = be.alloc(pairs, dtype=torch.int64)
        order = be.alloc(pairs, dtype=torch.int32)
        be.make_offsets(42, None, 0, 0, pairs, table_len, n, offs, order)
        ret = torch.randn(P, device=be.device)
        theta, m, v = (torch.randn(n, device=be.device) * 0.05), be.zeros(n), be.zeros(n)
        st = new_state(be.device)
        ad = adam_desc(lr=0.01)
        ranks = be.zeros(P, dtype=torch.int32)
        bytes_grad = 4 * n * pairs + 28 * n + 8 * P
        for label, od in (("sorted", order), ("unsorted", None)):
            med, best = timeit(lambda: be.rank_grad_adam(ret, None, 1.0, 0.0, P, table, offs, od, theta, m, v,
                                                          st, ad, ranks, None, None))
            print(json.dumps({"kernel": "rank_grad_adam", "config": name, "order": label, "n": n, "P": P,
                              "ms_median": med, "ms_best": best, "GBps": bytes_grad / med / 1e6,
                              "frac_hbm": bytes_grad / med / 1e6 / peak}), flush=True)
        obs = torch.randn(B, dims[0], device=be.device)
        tgt = torch.randn(B, dims[-1], device=be.device)
        epairs = pairs
        if n > 100000:
            epairs = 64
        rets = be.zeros(2 * epairs)
        for label, od in (("sorted", order if epairs == pairs else None),):
            if n > 100000 and args.skip_big_eval:
                continue
            med, best = timeit(lambda: be.eval_mlp(dims, theta, table, offs[:epairs].contiguous(), od, epairs, 0.02,
                                                    obs, tgt, rets[:epairs], rets[epairs:]), iters=5, warmup=2)
            flops = 2.0 * n * B * 2 * epairs
            print(json.dumps({"kernel": "eval_mlp_fp32", "config": name, "pairs": epairs, "B": B,
                              "ms_median": med, "ms_best": best, "TFLOPs": flops / med / 1e9,
                              "ms_extrapolated_full": med * pairs / epairs}), flush=True)
        if be.eval_supports_bf16(dims, B):
            rets2 = be.zeros(P)
            med, best = timeit(lambda: be.eval_mlp(dims, theta, table, offs, order, pairs, 0.02, obs, tgt,
                                                    rets2[:pairs], rets2[pairs:], precision="bf16"), iters=5, warmup=2)
            flops = 2.0 * n * B * 2 * pairs
            print(json.dumps({"kernel": "eval_mlp_bf16", "config": name, "pairs": pairs, "B": B, "ms_median": med,
                              "ms_best": best, "TFLOPs": flops / med / 1e9, "frac_tensor": flops / med / 1e9 / 1431.4}),
                  flush=True)
            th16 = be.alloc(n, dtype=torch.bfloat16); tb16 = be.alloc(table_len, dtype=torch.bfloat16)
            be.shadow_bf16(theta, th16); be.shadow_bf16(table, tb16)
            med, best = timeit(lambda: be.eval_mlp(dims, theta, table, offs, order, pairs, 0.02, obs, tgt,
                                                    rets2[:pairs], rets2[pairs:], precision="bf16s",
                                                    theta16=th16, table16=tb16), iters=5, warmup=2)
            print(json.dumps({"kernel": "eval_mlp_bf16s", "config": name, "pairs": pairs, "B": B, "ms_median": med,
                              "ms_best": best, "TFLOPs": flops / med / 1e9, "frac_tensor": flops / med / 1e9 / 1431.4}),
                  flush=True)
            one = be.zeros(1)
            med, best = timeit(lambda: be.eval_mlp_center(dims, theta, obs, tgt, one, precision="bf16"), iters=5, warmup=2)
            print(json.dumps({"kernel": "eval_mlp_center_bf16", "config": name, "ms_median": med}), flush=True)
        one = be.zeros(1)
        med, best = timeit(lambda: be.eval_mlp_center(dims, theta, obs, tgt, one), iters=5, warmup=2)
        print(json.dumps({"kernel": "eval_mlp_center", "config": name, "ms_median": med}), flush=True)


if __name__ == "__main__":
    main()
