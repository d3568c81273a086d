"""VirtualBatchNorm (Salimans et al. 2016) with the interface of the reference's
``estorch.VirtualBatchNorm`` (estorch/modules.py:6-58).

Contract kept from the reference: constructor ``(num_features, eps=1e-5)``;
learnable per-channel ``weight`` (ones) and ``bias`` (zeros); a TWO-CALL protocol --
the first ``forward`` after a reset receives the *reference batch*, records the
per-(C,H,W) mean and unbiased variance over the batch dimension and normalises
that batch; the next ``forward`` normalises the real batch with the recorded
statistics and forgets them (modules.py:48-58).  The statistics are not part of
``state_dict`` (plain attributes in the reference too).  ``mean`` / ``var`` remain
readable attributes (``None`` when no statistics are held).

The device evaluate kernel for conv policies (estk_eval_conv_vbn) implements the
same arithmetic; this module is what user policies are built from and what the
hooks path runs.
"""
from typing import Optional, Tuple

import torch
from torch import nn


class VirtualBatchNorm(nn.Module):
    def __init__(self, num_features: int, eps: float = 1e-5):
        super().__init__()
        self.num_features, self.eps = num_features, eps
        self.weight = nn.Parameter(torch.empty(num_features))
        self.bias = nn.Parameter(torch.empty(num_features))
        self._ref_stats: Optional[Tuple[torch.Tensor, torch.Tensor]] = None   # (variance, mean), each [1,C,H,W]
        self.reset_parameters()

    def reset_parameters(self) -> None:
        with torch.no_grad():
            self.weight.fill_(1.0)
            self.bias.zero_()

    # -- statistics, exposed under the reference's attribute names ----------------
    @property
    def var(self):
        return None if self._ref_stats is None else self._ref_stats[0]

    @property
    def mean(self):
        return None if self._ref_stats is None else self._ref_stats[1]

    def _affine(self, x: torch.Tensor) -> torch.Tensor:
        variance, mean = self._ref_stats
        per_channel = (1, self.num_features, 1, 1)
        centred = (x - mean) / torch.sqrt(variance + self.eps)
        return centred * self.weight.view(per_channel) + self.bias.view(per_channel)

    def normalize(self, x: torch.Tensor) -> torch.Tensor:
        """Normalise ``x`` with the statistics currently held."""
        return self._affine(x)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self._ref_stats is None:
            # call 1: x is the reference batch (torch.var_mean: unbiased over dim 0)
            self._ref_stats = torch.var_mean(x, dim=0, keepdim=True)
            return self._affine(x)
        # call 2: the real batch; statistics are single-use
        out = self._affine(x)
        self._ref_stats = None
        return out
