"""Agent protocol.

Reference protocol (estorch.py:182,200 / :405,428,439): an object with
``rollout(policy) -> float`` (ES) or ``-> (float, bc: 1-D array)`` (NS family).
Arbitrary Python rollouts cannot run inside a kernel, so the device path needs
an agent that *declares* its computation:

``DeviceAgent`` -- rollout is "run the policy over a fixed observation batch
and reduce": ``return = -mean((policy(obs) - target)**2)`` and (NS family)
``bc = policy(obs[:bc_obs]).flatten()[:bc_dim]`` (the 256-float behaviour
characteristic shape of examples/nsra_es.py:45-49).  It still implements the
reference protocol (``rollout``) with plain torch, so the very same object
runs under the CPU reference -- that is how the goldens and the CPU baseline
are produced.

Any other agent (a gym loop, ...) is a *host agent*: the engine materialises
each member's parameter row on the device and calls ``rollout`` on the host,
exactly like estorch.py:195-202.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch


class DeviceAgent:
    """Batch-regression agent evaluated by the fused CUDA kernel."""

    def __init__(self, obs: torch.Tensor, target: torch.Tensor, bc_obs: int = 0, bc_dim: int = 0):
        if obs.dim() < 2 or target.dim() != 2 or obs.shape[0] != target.shape[0]:
            raise ValueError("obs must be [B, in] (or [B, C, H, W] for conv policies) and target [B, out]")
        self.obs = obs.detach().to(torch.float32).contiguous()
        self.target = target.detach().to(torch.float32).contiguous()
        self.bc_obs = int(bc_obs)
        self.bc_dim = int(bc_dim)
        if bool(self.bc_obs) != bool(self.bc_dim):
            raise ValueError("bc_obs and bc_dim must both be zero or both be positive")

    # -- reference protocol (host / torch) ---------------------------------
    def rollout(self, policy):
        with torch.no_grad():
            p = next(policy.parameters())
            out = policy(self.obs.to(p.device))
            reward = float(-((out - self.target.to(p.device)) ** 2).mean())
            if self.bc_dim:
                bc = out[:self.bc_obs].flatten()[:self.bc_dim].detach().cpu().numpy().copy()
                return reward, bc
        return reward

    # -- device protocol ------------------------------------------------------
    def next_batch(self, step: int) -> Optional[Tuple[torch.Tensor, torch.Tensor]]:
        """Called once per generation before evaluation.  Return ``None`` to keep
        the current batch, or ``(obs, target)`` host tensors (ideally pinned) of
        the same shapes to upload for this generation."""
        return None


SyntheticAgent = DeviceAgent
