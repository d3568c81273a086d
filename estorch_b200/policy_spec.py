"""Recognise policies the device kernels can evaluate.

The reference accepts any ``nn.Module`` class as ``policy`` and runs it on the
host (estorch.py:136,142,195-202).  The fused evaluate kernel needs the
architecture, so the module is inspected once: a chain
``Linear -> ReLU -> ... -> Linear`` whose parameters are registered in forward
order (examples/cartpole_es.py:6-20, examples/nsra_es.py:52-67) becomes an
``MLPSpec``.  Anything else returns ``None`` and the engine uses the
materialising path (rows built on the device, rollout on the host).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import torch
from torch import nn


@dataclass(frozen=True)
class MLPSpec:
    dims: tuple  # (in, h1, ..., out)

    @property
    def n_parameters(self) -> int:
        d = self.dims
        return sum(d[i] * d[i + 1] + d[i + 1] for i in range(len(d) - 1))


def _leaf_modules(module: nn.Module) -> List[nn.Module]:
    return [m for m in module.modules() if len(list(m.children())) == 0]


def mlp_spec_from_module(module: nn.Module, probe: bool = True) -> Optional[MLPSpec]:
    """Return the MLPSpec of ``module`` or None.

    Structural test: the leaf modules are only Linear / ReLU; every Linear has a
    bias; widths chain; parameters are registered in layer order.
    Behavioural test (``probe``): a random batch through the module equals the
    chain evaluated from its flat parameters -- this rejects modules whose
    ``forward`` does something else with the same layers.
    """
    leaves = _leaf_modules(module)
    linears = [m for m in leaves if isinstance(m, nn.Linear)]
    others = [m for m in leaves if not isinstance(m, (nn.Linear, nn.ReLU))]
    if not linears or others or len(linears) > 8:
        return None
    if any(l.bias is None for l in linears):
        return None
    dims = [linears[0].in_features]
    for l in linears:
        if l.in_features != dims[-1]:
            return None
        dims.append(l.out_features)
    params = list(module.parameters())
    expect = [p for l in linears for p in (l.weight, l.bias)]
    if len(params) != len(expect) or any(a is not b for a, b in zip(params, expect)):
        return None
    spec = MLPSpec(tuple(dims))
    if probe:
        with torch.no_grad():
            dev = params[0].device
            x = torch.randn(3, dims[0], device=dev, dtype=params[0].dtype)
            try:
                y = module(x)
            except Exception:
                return None
            h = x
            for i, l in enumerate(linears):
                h = torch.nn.functional.linear(h, l.weight, l.bias)
                if i + 1 < len(linears):
                    h = torch.relu(h)
            if y.shape != h.shape or not torch.allclose(y, h, rtol=1e-4, atol=1e-5):
                return None
    return spec


@dataclass(frozen=True)
class ConvVBNSpec:
    """The conv + VirtualBatchNorm policy of the reference's Atari example
    (examples/atari.py:14-37): conv1 4->16 k8 s4, VBN(16), conv2 16->32 k4 s2,
    VBN(32), fc1 2592->256, fc2 256->n_actions, with the reference batch ``xref``."""
    n_actions: int
    ref_batch: int

    @property
    def n_parameters(self) -> int:
        return 4096 + 16 + 16 + 16 + 8192 + 32 + 32 + 32 + 256 * 2592 + 256 + 256 * self.n_actions + self.n_actions


def conv_vbn_spec_from_module(module: nn.Module) -> Optional[ConvVBNSpec]:
    """Recognise the Atari-example architecture: leaf modules, in registration order,
    Conv2d(4,16,8,4) / VirtualBatchNorm(16) / Conv2d(16,32,4,2) / VirtualBatchNorm(32) /
    Linear(2592,256) / Linear(256,A), plus an ``xref`` tensor ``[R,4,84,84]``; the forward
    is probed against the same chain evaluated from the module's own parameters."""
    from .vbn import VirtualBatchNorm
    leaves = _leaf_modules(module)
    if len(leaves) != 6:
        return None
    c1, b1, c2, b2, f1, f2 = leaves
    ok = (isinstance(c1, nn.Conv2d) and (c1.in_channels, c1.out_channels, c1.kernel_size, c1.stride, c1.padding)
          == (4, 16, (8, 8), (4, 4), (0, 0)) and c1.bias is not None
          and isinstance(b1, VirtualBatchNorm) and b1.num_features == 16 and abs(b1.eps - 1e-5) < 1e-12
          and isinstance(c2, nn.Conv2d) and (c2.in_channels, c2.out_channels, c2.kernel_size, c2.stride, c2.padding)
          == (16, 32, (4, 4), (2, 2), (0, 0)) and c2.bias is not None
          and isinstance(b2, VirtualBatchNorm) and b2.num_features == 32 and abs(b2.eps - 1e-5) < 1e-12
          and isinstance(f1, nn.Linear) and (f1.in_features, f1.out_features) == (2592, 256) and f1.bias is not None
          and isinstance(f2, nn.Linear) and f2.in_features == 256 and f2.bias is not None)
    xref = getattr(module, "xref", None)
    if not ok or not torch.is_tensor(xref) or xref.dim() != 4 or tuple(xref.shape[1:]) != (4, 84, 84) or xref.shape[0] < 2:
        return None
    expect = [c1.weight, c1.bias, b1.weight, b1.bias, c2.weight, c2.bias, b2.weight, b2.bias,
              f1.weight, f1.bias, f2.weight, f2.bias]
    params = list(module.parameters())
    if len(params) != len(expect) or any(a is not b for a, b in zip(params, expect)):
        return None
    with torch.no_grad():
        F = torch.nn.functional
        dev = params[0].device
        x = torch.rand(2, 4, 84, 84, device=dev)
        xr = xref.to(dev)
        try:
            y = module(x)
        except Exception:
            return None

        def vbn(t, ref, m):
            mean, var = ref.mean(0, keepdim=True), ref.var(0, keepdim=True)
            return (t - mean) / torch.sqrt(var + m.eps) * m.weight.view(1, -1, 1, 1) + m.bias.view(1, -1, 1, 1)
        r1 = F.conv2d(xr, c1.weight, c1.bias, stride=4)
        h1 = torch.relu(vbn(F.conv2d(x, c1.weight, c1.bias, stride=4), r1, b1))
        r1n = torch.relu(vbn(r1, r1, b1))
        r2 = F.conv2d(r1n, c2.weight, c2.bias, stride=2)
        h2 = torch.relu(vbn(F.conv2d(h1, c2.weight, c2.bias, stride=2), r2, b2))
        want = F.linear(torch.relu(F.linear(h2.reshape(-1, 2592), f1.weight, f1.bias)), f2.weight, f2.bias)
        if y.shape != want.shape or not torch.allclose(y, want, rtol=1e-3, atol=1e-4):
            return None
    return ConvVBNSpec(int(f2.out_features), int(xref.shape[0]))
