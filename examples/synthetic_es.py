#!/usr/bin/env python
"""Usage example, written like the reference's examples/cartpole_es.py:43-61 but with a
device agent: classic ES on a synthetic batch-regression task.

    python examples/synthetic_es.py                 # one GPU
    python examples/synthetic_es.py --n-proc 2      # re-launches itself under torchrun, one process per GPU
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from estorch_b200 import ES, DeviceAgent  # noqa: E402


class Policy(torch.nn.Module):
    def __init__(self, n_input, n_output):
        super().__init__()
        self.linear_1 = torch.nn.Linear(n_input, 64)
        self.activation_1 = torch.nn.ReLU()
        self.linear_2 = torch.nn.Linear(64, 64)
        self.activation_2 = torch.nn.ReLU()
        self.linear_3 = torch.nn.Linear(64, n_output)

    def forward(self, x):
        return self.linear_3(self.activation_2(self.linear_2(self.activation_1(self.linear_1(x)))))


class Trainer(ES):
    def log(self):                                   # called on rank 0 after every generation
        if self.step % 10 == 0:
            print(f"step {self.step:4d}  episode {self.episode_reward:.5f}  best {self.best_reward:.5f}  "
                  f"max population {self.population_returns.max():.5f}", flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-proc", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    args = ap.parse_args()
    g = torch.Generator().manual_seed(0)
    obs = torch.randn(256, 4, generator=g)
    teacher = Policy(4, 2)
    with torch.no_grad():
        target = teacher(obs)                        # a realisable target: the return can approach 0
    torch.manual_seed(1)
    es = Trainer(Policy, DeviceAgent, torch.optim.Adam, population_size=512, sigma=0.02,
                 policy_kwargs={"n_input": 4, "n_output": 2}, agent_kwargs={"obs": obs, "target": target},
                 optimizer_kwargs={"lr": 0.01}, noise_table_size=1 << 24)
    es.train(n_steps=args.steps, n_proc=args.n_proc)
    if es.rank == 0:
        first = es.agent.rollout(teacher.__class__(4, 2))
        print(f"done: best reward {es.best_reward:.5f} (a random policy scores about {first:.3f}); "
              f"world size {es.n_workers}", flush=True)
        assert es.best_reward > -0.05
