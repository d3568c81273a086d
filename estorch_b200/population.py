"""Lazy stand-ins for the reference's ``population_parameters`` / ``epsilon``.

The reference materialises both as dense ``[P, n]`` CPU tensors every
generation (estorch.py:187-193; 16-33 GB each at the north-star sizes).  Here
a row is a pure function of (theta snapshot, noise table, offset, sigma), so
both are handles that build rows on the device on demand
(``estk_perturb_rows``).  ``population_parameters[idx]`` -- the one access the
reference's examples make (examples/early_stopping.py:22) -- works unchanged.
"""
from __future__ import annotations

import torch


class _NoiseRows:
    def __init__(self, backend, theta, table, offsets, sigma, population_size):
        self._be = backend
        self._theta = theta            # snapshot of the centre this population was drawn around
        self._table = table
        self._offsets = offsets        # [pairs] int64, ALL pairs of the population
        self.sigma = float(sigma)
        self.population_size = int(population_size)
        self.n = theta.numel()

    @property
    def shape(self):
        return (self.population_size, self.n)

    def size(self, dim=None):
        return self.shape if dim is None else self.shape[dim]

    def __len__(self):
        return self.population_size

    def _build(self, begin, count, want_eps):
        out = torch.empty(count, self.n, dtype=torch.float32, device=self._theta.device)
        if want_eps:
            self._be.perturb_rows(self._theta, self._table, self._offsets, self.population_size // 2,
                                  self.sigma, begin, count, None, out)
        else:
            self._be.perturb_rows(self._theta, self._table, self._offsets, self.population_size // 2,
                                  self.sigma, begin, count, out, None)
        return out

    def _get(self, idx, want_eps):
        if isinstance(idx, slice):
            start, stop, step = idx.indices(self.population_size)
            if step != 1:
                return torch.stack([self._get(i, want_eps) for i in range(start, stop, step)])
            if stop <= start:
                return torch.empty(0, self.n, device=self._theta.device)
            return self._build(start, stop - start, want_eps)
        if torch.is_tensor(idx):
            idx = idx.item() if idx.dim() == 0 else idx.tolist()
        if isinstance(idx, (list, tuple)):
            return torch.stack([self._get(int(i), want_eps) for i in idx])
        i = int(idx)
        if i < 0:
            i += self.population_size
        if not 0 <= i < self.population_size:
            raise IndexError(f"member {idx} out of range for population of {self.population_size}")
        return self._build(i, 1, want_eps)[0]

    def rows(self, begin, count, want_eps=False):
        return self._build(begin, count, want_eps)

    def __iter__(self):
        for i in range(self.population_size):
            yield self[i]


class LazyPopulation(_NoiseRows):
    """``population_parameters``: row i = theta +- sigma*T[off] (estorch.py:192)."""

    def __getitem__(self, idx):
        return self._get(idx, False)

    def materialize(self):
        """The dense ``[P, n]`` tensor of estorch.py:192 (can be tens of GB)."""
        return self._build(0, self.population_size, False)


class NoiseHandle(_NoiseRows):
    """``epsilon``: row i = +-sigma*T[off] (estorch.py:193)."""

    def __getitem__(self, idx):
        return self._get(idx, True)

    def materialize(self):
        return self._build(0, self.population_size, True)
