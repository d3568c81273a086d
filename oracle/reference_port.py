"""CPU port of the reference's whole generation, for TIMING the reference's own
way of doing the work on the host cores (bench.py `cpu_baseline` and
`--impl reference`).  TEST/BENCH INFRASTRUCTURE ONLY.

Unlike ``es_oracle`` (which restates the arithmetic at the noise-injection
boundary), this port keeps the reference's *cost structure*: fresh
``Normal(0, sigma).sample`` per generation, the two ``[P, n]`` concatenations,
one ``vector_to_parameters`` + ``agent.rollout`` per member, centred ranks via
numpy argsort, ``torch.mm`` over the full ``[P, n]`` epsilon, per-parameter
grad scatter + clamp and ``torch.optim.Adam.step`` -- estorch/estorch.py:187-246.
The reference itself (pure Python, needs mpi4py which is not installed) cannot
travel to the GPU box; its arithmetic is pinned separately by tests/golden.
"""
from __future__ import annotations

import time

import numpy as np
import torch

from .es_oracle import rank_transformation


class PhaseTimer:
    def __init__(self):
        self.t = {}

    def add(self, name, dt):
        self.t[name] = self.t.get(name, 0.0) + dt


def reference_generation(policy, target, agent, optimizer, population_size, sigma, timer=None):
    """One iteration of estorch.py:214-248 on CPU tensors.  Returns
    (population_returns [P,1] float32, episode_reward)."""
    timer = timer or PhaseTimer()
    with torch.no_grad():
        t0 = time.perf_counter()
        flat = torch.nn.utils.parameters_to_vector(policy.parameters())          # :188
        noise = torch.distributions.normal.Normal(0, sigma).sample(               # :189-190
            [population_size // 2, flat.shape[0]])
        flat = flat.detach().cpu()
        members = torch.cat((flat + noise, flat - noise))                         # :192
        signed_noise = torch.cat((noise, -noise))                                 # :193
        t1 = time.perf_counter()
        timer.add("sample", t1 - t0)

        rewards = []
        for row in members:                                                       # :197-201
            torch.nn.utils.vector_to_parameters(row, target.parameters())
            rewards.append(agent.rollout(target))
        returns = np.array(rewards, dtype=np.float32)[:, np.newaxis]             # :202
        t2 = time.perf_counter()
        timer.add("returns", t2 - t1)

        centred = torch.from_numpy(rank_transformation(returns.squeeze())).unsqueeze(0).float()   # :175-176
        grad = (torch.mm(centred, signed_noise) / (population_size * sigma)).squeeze()           # :177-178
        t3 = time.perf_counter()
        timer.add("grad", t3 - t2)

        cursor = 0
        for param in policy.parameters():                                         # :237-244
            count = int(np.prod(param.shape))
            param.grad = -grad[cursor:cursor + count].view(param.shape)
            param.grad.data.clamp_(-1.0, 1.0)
            cursor += count
        optimizer.step()                                                          # :245
        episode_reward = agent.rollout(policy)                                    # :182
        t4 = time.perf_counter()
        timer.add("update", t4 - t3)
    return returns, episode_reward


def time_reference(policy_cls, policy_kwargs, agent, population_size, sigma, steps, warmup=0, lr=0.01):
    """Build policy/target/Adam the way the reference constructor does
    (estorch.py:136-142) and time ``steps`` generations.  Returns
    (seconds_per_generation, phase seconds dict)."""
    policy = policy_cls(**policy_kwargs)
    target = policy_cls(**policy_kwargs)
    optimizer = torch.optim.Adam(policy.parameters(), lr=lr)
    for _ in range(warmup):
        reference_generation(policy, target, agent, optimizer, population_size, sigma)
    timer = PhaseTimer()
    t0 = time.perf_counter()
    for _ in range(steps):
        reference_generation(policy, target, agent, optimizer, population_size, sigma, timer)
    dt = (time.perf_counter() - t0) / max(1, steps)
    return dt, {k: v / max(1, steps) for k, v in timer.t.items()}
