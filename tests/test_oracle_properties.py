"""Size-independent properties of the restated algorithm (hypothesis, CPU only): the same
identities the full-size GPU tests check on the kernels (tests/test_kernels_gpu.py::
test_full_size_north_star_gradient_property), here on the oracle itself, plus the oracle's
Adam against torch.optim.Adam on random inputs (the reference's optimizer, estorch.py:245)."""
import numpy as np
import torch
from hypothesis import given, settings, strategies as st

from oracle import es_oracle as orc

SET = settings(max_examples=40, deadline=None)


def _distinct_returns(rng, P):
    r = rng.standard_normal(P).astype(np.float32)
    while len(np.unique(r)) != P:                       # ties are unspecified in the reference (numpy argsort)
        r = rng.standard_normal(P).astype(np.float32)
    return r


@SET
@given(st.integers(2, 200), st.integers(0, 2 ** 31 - 1))
def test_rank_transform_properties(P, seed):
    rng = np.random.RandomState(seed)
    r = _distinct_returns(rng, P)
    c = orc.rank_transformation(r)
    assert c.dtype == np.float64 and c.shape == (P,)
    np.testing.assert_allclose(np.sort(c), np.arange(P) / (P - 1) - 0.5, rtol=0, atol=1e-15)   # estorch.py:17-19
    assert c[np.argmin(r)] == -0.5 and c[np.argmax(r)] == 0.5
    perm = rng.permutation(P)
    np.testing.assert_array_equal(orc.rank_transformation(r[perm]), c[perm])                   # equivariance
    np.testing.assert_array_equal(orc.rank_transformation(3.0 * r + 7.0), c)                   # monotone invariance


@SET
@given(st.integers(1, 24), st.integers(1, 96), st.integers(0, 2 ** 31 - 1))
def test_gradient_estimate_identities(pairs, n, seed):
    rng = np.random.RandomState(seed)
    P, sigma = 2 * pairs, 0.05
    table = rng.standard_normal(4096 + n).astype(np.float32)
    offs = (rng.randint(0, 4096 // 32, size=pairs) * 32).astype(np.int64)
    r = _distinct_returns(rng, P)
    g = orc.calculate_grad_pairs(r, table, offs, n)
    # the reference's P x n matmul form on the sigma-carrying epsilon is the same sum (estorch.py:177-178)
    eps = np.stack([np.float32(sigma) * table[o:o + n] for o in offs])
    g_ref = orc.calculate_grad(r, np.concatenate([eps, -eps]), sigma)
    assert np.max(np.abs(g - g_ref)) <= 2e-5 * max(np.max(np.abs(g)), 1e-6) + 1e-7
    # mirrored sampling: swapping the + and - halves of the returns negates the estimate
    swapped = np.concatenate([r[pairs:], r[:pairs]])
    np.testing.assert_allclose(orc.calculate_grad_pairs(swapped, table, offs, n), -g, rtol=0, atol=1e-12)
    # it only depends on the order of the returns
    np.testing.assert_array_equal(orc.calculate_grad_pairs(np.exp(r), table, offs, n), g)
    # and it is linear in the noise rows
    g2 = orc.calculate_grad_pairs(r, (2.0 * table).astype(np.float32), offs, n)
    np.testing.assert_allclose(g2, 2.0 * g, rtol=1e-12, atol=0)


@SET
@given(st.integers(1, 300), st.integers(0, 2 ** 31 - 1), st.integers(1, 6))
def test_negate_clamp_adam_matches_torch(n, seed, steps):
    rng = np.random.RandomState(seed)
    theta0 = rng.standard_normal(n).astype(np.float32)
    p = torch.nn.Parameter(torch.from_numpy(theta0.copy()))
    opt = torch.optim.Adam([p], lr=0.01)
    theta, m, v = theta0.copy(), np.zeros(n, np.float32), np.zeros(n, np.float32)
    for step in range(1, steps + 1):
        g = (rng.standard_normal(n) * rng.choice([1e-3, 0.5, 3.0])).astype(np.float32)
        p.grad = torch.from_numpy(-g).clamp_(-1, 1)                      # estorch.py:239-244
        opt.step()
        theta, m, v = orc.adam_step(theta, m, v, orc.negate_clamp(g), step)
    assert np.max(np.abs(theta - p.detach().numpy())) <= 1e-6 * max(1.0, np.max(np.abs(theta)))
