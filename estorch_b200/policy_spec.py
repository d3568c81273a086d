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
