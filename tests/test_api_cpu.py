"""Host logic of the estorch-compatible classes, on CPU, with the test-only
oracle stand-in for the kernels (tests/_oracle_backend.py).  Results are checked
against the goldens produced by the unmodified reference."""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err
from _oracle_backend import OracleBackend
import estorch_b200 as E


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


def _load_theta(module, flat):
    torch.nn.utils.vector_to_parameters(torch.from_numpy(flat.copy()), module.parameters())


class Rec:
    """mixin: record per-generation values in log()."""
    def log(self):
        self.rec.append(dict(returns=self.population_returns.copy(), episode=self.episode_reward,
                             best=self.best_reward))


def _make(cls, g, P, sigma, **kw):
    dims = [int(d) for d in g["dims"]]
    obs, tgt = torch.from_numpy(g["obs"]), torch.from_numpy(g["target"])

    class R(Rec, cls):
        pass
    es = R(MLP, E.DeviceAgent, torch.optim.Adam, population_size=P, sigma=sigma,
           policy_kwargs={"dims": dims},
           agent_kwargs=dict(obs=obs, target=tgt, bc_obs=int(g["bc_obs"]), bc_dim=int(g["bc_dim"])),
           optimizer_kwargs={"lr": 0.01}, noise_table_size=len(g["table"]),
           noise_seed=int(g["noise_seed"]), _backend=OracleBackend(), **kw)
    es.rec = []
    es._table.copy_(torch.from_numpy(g["table"]))       # the goldens' table (torch.randn, not Philox)
    return es


def test_public_names():
    for name in ("ES", "NS_ES", "NSR_ES", "NSRA_ES", "rank_transformation", "VirtualBatchNorm"):
        assert hasattr(E, name)                       # docs/index.rst:4-15 of the reference
    got = E.rank_transformation([-123, -50, 3, -5, 20, 10, 100])          # estorch.py:31-35
    np.testing.assert_allclose(got, [-0.5, -1 / 3, 0., -1 / 6, 1 / 3, 1 / 6, 0.5], atol=1e-12)


def test_es_fused_matches_reference_golden():
    g = load_golden("es_cartpole_p64.npz")
    es = _make(E.ES, g, 64, 0.1)
    assert es._fused and es.n_parameters == 4610 and es.n_workers == 1 and es.rank == 0
    _load_theta(es.policy, g["theta0"])
    es._slots[0].ensure_flat()        # vector_to_parameters re-pointed .data; train() re-aliases too
    assert rel_err(es._slots[0].theta.numpy(), g["theta0"]) == 0      # parameters are views of flat theta
    es.train(n_steps=3)
    assert es.step == 3 and len(es.rec) == 3
    for gen in range(3):
        assert es.rec[gen]["returns"].shape == (64, 1) and es.rec[gen]["returns"].dtype == np.float32
        assert rel_err(es.rec[gen]["returns"][:, 0], g["returns"][gen][:, 0]) < 1e-5
        assert abs(es.rec[gen]["episode"] - float(g["episode_reward"][gen])) < 2e-5
        assert abs(es.rec[gen]["best"] - float(g["best_reward"][gen])) < 2e-5
    theta = torch.nn.utils.parameters_to_vector(es.policy.parameters()).detach().numpy()
    assert rel_err(theta, g["theta_after"][2]) < 2e-4      # 3 chained generations on oracle returns
    # Adam moments are exposed through the torch optimizer object
    st = es.optimizer.state[next(es.policy.parameters())]
    assert float(st["step"]) == 3.0 and st["exp_avg"].shape == (64, 4)
    bp = es.best_policy_dict
    assert list(bp.keys()) == list(es.policy.state_dict().keys())
    assert rel_err(np.concatenate([v.reshape(-1).numpy() for v in bp.values()]), g["best_theta"]) < 2e-4
    # lazy population rows of the LAST generation are built around the pre-update theta
    pop = es.population_parameters
    assert pop.shape == (64, 4610)
    row = pop[5].numpy()
    t = g["table"][g["offsets"][2][5]: g["offsets"][2][5] + 4610]
    assert rel_err(row, g["theta_before"][2] + np.float32(0.1) * t) < 1e-4


def test_terminate_and_population_indexing_like_early_stopping_example():
    g = load_golden("es_cartpole_p64.npz")

    class Stopper(E.ES):                      # examples/early_stopping.py:7-23
        def log(self):
            idx = np.argmax(self.population_returns)
            self.best = self.population_parameters[idx]
            if self.step == 1:
                self.terminate()
    dims = [4, 64, 64, 2]
    es = Stopper(MLP, E.DeviceAgent, torch.optim.Adam, population_size=64, sigma=0.1,
                 policy_kwargs={"dims": dims},
                 agent_kwargs=dict(obs=torch.from_numpy(g["obs"]), target=torch.from_numpy(g["target"])),
                 optimizer_kwargs={"lr": 0.01}, noise_table_size=1 << 15, _backend=OracleBackend())
    es.train(n_steps=10)
    assert es.step == 2 and es.best.shape == (4610,)


class HostAgent:
    """A plain reference-protocol agent (not a DeviceAgent): forces the hooks path."""
    def __init__(self, obs, target):
        self.obs, self.target = obs, target

    def rollout(self, policy):
        with torch.no_grad():
            return float(-((policy(self.obs) - self.target) ** 2).mean())


def test_host_agent_hooks_path_matches_reference_golden():
    g = load_golden("es_cartpole_p64.npz")
    obs, tgt = torch.from_numpy(g["obs"]), torch.from_numpy(g["target"])

    class R(Rec, E.ES):
        pass
    es = R(MLP, HostAgent, torch.optim.Adam, population_size=64, sigma=0.1,
           policy_kwargs={"dims": [4, 64, 64, 2]}, agent_kwargs=dict(obs=obs, target=tgt),
           optimizer_kwargs={"lr": 0.01}, noise_table_size=len(g["table"]), noise_seed=int(g["noise_seed"]),
           _backend=OracleBackend())
    es.rec = []
    assert not es._fused
    es._table.copy_(torch.from_numpy(g["table"]))
    _load_theta(es.policy, g["theta0"])
    es.train(n_steps=2)
    for gen in range(2):
        assert rel_err(es.rec[gen]["returns"][:, 0], g["returns"][gen][:, 0]) < 1e-5
    theta = torch.nn.utils.parameters_to_vector(es.policy.parameters()).detach().numpy()
    assert rel_err(theta, g["theta_after"][1]) < 1e-4
    assert abs(es.best_reward - float(g["best_reward"][1])) < 2e-5
    assert rel_err(np.concatenate([v.reshape(-1).numpy() for v in es.best_policy_dict.values()]),
                   g["theta_after"][int(np.argmax(g["episode_reward"][:2]))]) < 1e-4


def test_custom_subclass_hooks_are_honoured():
    """examples/custom_es.py:8-27: overriding _sample_policy/_calculate_grad with
    dense tensors must bypass the fused kernels and still train."""
    g = load_golden("es_cartpole_p64.npz")
    calls = {"sample": 0, "grad": 0}

    class SymmetricES(E.ES):
        def _sample_policy(self, policy):
            calls["sample"] += 1
            parameters = torch.nn.utils.parameters_to_vector(policy.parameters())
            gen = torch.Generator().manual_seed(self.step)
            epsilon = torch.randn(self.population_size // 2, parameters.shape[0], generator=gen) * self.sigma
            parameters = parameters.detach().cpu()
            return torch.cat((parameters + epsilon, parameters - epsilon)), epsilon

        def _calculate_grad(self, epsilon):
            calls["grad"] += 1
            ranked = torch.from_numpy(E.rank_transformation(self.population_returns.squeeze())).unsqueeze(0).float()
            batch = self.population_size // 2
            return (torch.mm((ranked[0, :batch] - ranked[0, batch:]).unsqueeze(0), epsilon) /
                    (batch * self.sigma)).squeeze()

        def log(self):
            pass
    es = SymmetricES(MLP, E.DeviceAgent, torch.optim.Adam, population_size=16, sigma=0.05,
                     policy_kwargs={"dims": [4, 64, 64, 2]},
                     agent_kwargs=dict(obs=torch.from_numpy(g["obs"]), target=torch.from_numpy(g["target"])),
                     optimizer_kwargs={"lr": 0.01}, noise_table_size=1 << 15, _backend=OracleBackend())
    assert not es._fused
    before = torch.nn.utils.parameters_to_vector(es.policy.parameters()).detach().clone()
    es.train(n_steps=2)
    assert calls == {"sample": 2, "grad": 2}
    after = torch.nn.utils.parameters_to_vector(es.policy.parameters()).detach()
    assert float((after - before).abs().max()) > 1e-3 and es.population_returns.shape == (16, 1)


def test_other_optimizer_uses_hooks_path():
    g = load_golden("es_tiny_p8.npz")
    es = E.ES(MLP, E.DeviceAgent, torch.optim.SGD, population_size=8, sigma=0.05,
              policy_kwargs={"dims": [3, 2]},
              agent_kwargs=dict(obs=torch.from_numpy(g["obs"]), target=torch.from_numpy(g["target"])),
              optimizer_kwargs={"lr": 0.1}, noise_table_size=1 << 10, _backend=OracleBackend())
    es.log = lambda: None
    assert not es._fused
    es.train(n_steps=2)
    assert np.isfinite(es.episode_reward)


def test_argument_errors():
    g = load_golden("es_tiny_p8.npz")
    kw = dict(policy_kwargs={"dims": [3, 2]},
              agent_kwargs=dict(obs=torch.from_numpy(g["obs"]), target=torch.from_numpy(g["target"])),
              _backend=OracleBackend(), noise_table_size=1 << 10)
    with pytest.raises(ValueError):
        E.ES(MLP, E.DeviceAgent, torch.optim.Adam, population_size=7, **kw)
    es = E.ES(MLP, E.DeviceAgent, torch.optim.Adam, population_size=8, **kw)
    with pytest.raises(NotImplementedError):
        es.train(1, hostfile="hosts")


@pytest.mark.parametrize("algo,cls", [("ns", "NS_ES"), ("nsr", "NSR_ES"), ("nsra", "NSRA_ES")])
@pytest.mark.parametrize("mode", ["fused", "hooks"])
def test_ns_family_matches_reference_golden(algo, cls, mode):
    g = load_golden(f"{algo}_bipedal_p32.npz")
    kw = {"weight_t": 2} if algo == "nsra" else {}
    dims = [int(d) for d in g["dims"]]
    obs, tgt = torch.from_numpy(g["obs"]), torch.from_numpy(g["target"])
    base = getattr(E, cls)

    class R(base):
        def log(self):
            self.rec.append(dict(returns=self.population_returns.copy(), episode=self.episode_reward,
                                 idx=self.idx, weight=getattr(self, "weight", None), t=getattr(self, "t", None),
                                 archive=len(self._archive)))
    agent_cls = E.DeviceAgent
    if mode == "hooks":
        class agent_cls:                          # plain reference-protocol agent: same maths on the host
            def __init__(self, obs, target, bc_obs, bc_dim):
                self.inner = E.DeviceAgent(obs, target, bc_obs, bc_dim)

            def rollout(self, policy):
                return self.inner.rollout(policy)
    np.random.seed(123)                                # the golden run's np.random.choice stream
    es = R(MLP, agent_cls, torch.optim.Adam, population_size=32, sigma=0.02,
           policy_kwargs={"dims": dims},
           agent_kwargs=dict(obs=obs, target=tgt, bc_obs=64, bc_dim=256), optimizer_kwargs={"lr": 0.01},
           noise_table_size=len(g["table"]), noise_seed=int(g["noise_seed"]), _backend=OracleBackend(), **kw)
    es.rec = []
    assert not hasattr(es, "policy")                   # NS objects have no .policy (estorch.py:135-137)
    assert len(es.meta_population) == 3 and len(es._archive) == 3
    es._table.copy_(torch.from_numpy(g["table"]))
    for i, (p, _) in enumerate(es.meta_population):
        _load_theta(p, g["meta_theta0"][i])
        es._slots[i].push_theta()
    es._archive = [a.copy() for a in g["archive0"]]
    np.random.seed(123)
    # the golden run drew 0 numbers before training, so the choice stream lines up
    assert es._fused == (mode == "fused")
    es.train(n_steps=len(g["grad"]))
    for gen in range(len(g["grad"])):
        r = es.rec[gen]
        assert r["idx"] == int(g["idx"][gen])
        assert r["returns"].shape == (32, 2)
        assert rel_err(r["returns"][:, 0], g["returns"][gen][:, 0]) < 1e-4
        assert rel_err(r["returns"][:, 1], g["returns"][gen][:, 1]) < 1e-4
        assert abs(r["episode"] - float(g["episode_reward"][gen])) < 1e-4
        assert r["archive"] == int(g["archive_len"][gen])
        if algo == "nsra":
            assert r["weight"] == pytest.approx(float(g["weight"][gen])) and r["t"] == int(g["t"][gen])
    final = np.stack([torch.nn.utils.parameters_to_vector(p.parameters()).detach().numpy()
                      for p, _ in es.meta_population])
    assert rel_err(final, g["meta_theta_final"]) < 5e-3   # chained Adam steps, sign flips at g~0 allowed
    assert abs(es.best_reward - float(max(g["episode_reward"]))) < 1e-4


def test_policy_spec_detection():
    from estorch_b200.policy_spec import mlp_spec_from_module

    class Example(torch.nn.Module):            # examples/cartpole_es.py:5-20
        def __init__(self):
            super().__init__()
            self.linear_1 = torch.nn.Linear(4, 64)
            self.activation_1 = torch.nn.ReLU()
            self.linear_2 = torch.nn.Linear(64, 64)
            self.activation_2 = torch.nn.ReLU()
            self.linear_3 = torch.nn.Linear(64, 2)

        def forward(self, x):
            return self.linear_3(self.activation_2(self.linear_2(self.activation_1(self.linear_1(x)))))
    assert mlp_spec_from_module(Example()).dims == (4, 64, 64, 2)
    assert mlp_spec_from_module(MLP([24, 64, 64, 4])).n_parameters == 6020

    class Tanh(Example):
        def forward(self, x):
            return self.linear_3(torch.tanh(self.linear_2(torch.tanh(self.linear_1(x)))))
    assert mlp_spec_from_module(Tanh()) is None
    assert mlp_spec_from_module(torch.nn.Sequential(torch.nn.Conv2d(1, 1, 1))) is None


def test_virtual_batch_norm_matches_reference_golden():
    g = load_golden("vbn.npz")
    vbn = E.VirtualBatchNorm(3)
    with torch.no_grad():
        vbn.weight.copy_(torch.from_numpy(g["gamma"]))
        vbn.bias.copy_(torch.from_numpy(g["beta"]))
        y_ref = vbn(torch.from_numpy(g["xref"]))
        assert vbn.mean is not None and vbn.mean.shape == (1, 3, 5, 4)
        y = vbn(torch.from_numpy(g["x"]))
        assert vbn.mean is None and vbn.var is None
    np.testing.assert_allclose(y_ref.numpy(), g["y_ref"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(y.numpy(), g["y"], rtol=1e-6, atol=1e-6)
    assert sorted(vbn.state_dict().keys()) == ["bias", "weight"]   # stats are not in state_dict


@pytest.mark.parametrize("algo", ["es", "nsra"])
def test_checkpoint_resume_is_bit_identical(algo, tmp_path):
    """train(2) + save + train(2)  ==  fresh instance + load + train(2)."""
    g = load_golden("es_cartpole_p64.npz" if algo == "es" else "nsra_bipedal_p32.npz")

    def make():
        np.random.seed(7)
        torch.manual_seed(3)
        if algo == "es":
            es = _make(E.ES, g, 64, 0.1)
        else:
            es = _make(E.NSRA_ES, g, 32, 0.02, weight_t=2)
        es.log = lambda: None
        return es
    a = make()
    a.train(n_steps=2)
    path = str(tmp_path / "ck.pt")
    a.save_checkpoint(path)
    a.train(n_steps=2)
    b = make()
    b.load_checkpoint(path)
    assert b._generation == 2
    b.train(n_steps=2)
    for sa, sb in zip(a._slots, b._slots):
        np.testing.assert_array_equal(sa.theta.numpy(), sb.theta.numpy())
        np.testing.assert_array_equal(sa.m.numpy(), sb.m.numpy())
        np.testing.assert_array_equal(sa.v.numpy(), sb.v.numpy())
    assert a.best_reward == b.best_reward and a._generation == b._generation == 4
    if algo == "nsra":
        assert a.weight == b.weight and a.t == b.t and len(a._archive) == len(b._archive)
        np.testing.assert_array_equal(np.stack(a._archive), np.stack(b._archive))
    # a second train() call continues with NEW noise (generation counter is not reset)
    assert a.step == 2


def test_lazy_population_indexing_and_noise_handle():
    g = load_golden("es_cartpole_p64.npz")
    es = _make(E.ES, g, 64, 0.1)
    es.log = lambda: None
    _load_theta(es.policy, g["theta0"])
    es.train(n_steps=1)
    pop = es.population_parameters
    assert len(pop) == 64 and pop.shape == (64, 4610) and pop.size(1) == 4610
    full = pop.materialize()
    assert full.shape == (64, 4610)
    np.testing.assert_array_equal(pop[3].numpy(), full[3].numpy())
    np.testing.assert_array_equal(pop[-1].numpy(), full[63].numpy())
    np.testing.assert_array_equal(pop[10:13].numpy(), full[10:13].numpy())
    np.testing.assert_array_equal(pop[[1, 40]].numpy(), full[[1, 40]].numpy())
    np.testing.assert_array_equal(pop[torch.tensor(7)].numpy(), full[7].numpy())
    with pytest.raises(IndexError):
        pop[64]
    # mirrored layout of estorch.py:192: row j+P/2 = 2*theta - row j
    theta = g["theta_before"][0]
    assert rel_err((full[5] + full[5 + 32]).numpy() / 2, theta) < 1e-6


def test_checkpoint_resume_hooks_mode(tmp_path):
    g = load_golden("es_cartpole_p64.npz")
    obs, tgt = torch.from_numpy(g["obs"]), torch.from_numpy(g["target"])

    def make():
        torch.manual_seed(3)
        es = E.ES(MLP, HostAgent, torch.optim.SGD, population_size=16, sigma=0.1,
                  policy_kwargs={"dims": [4, 64, 64, 2]}, agent_kwargs=dict(obs=obs, target=tgt),
                  optimizer_kwargs={"lr": 0.05, "momentum": 0.9}, noise_table_size=1 << 15, _backend=OracleBackend())
        es.log = lambda: None
        return es
    a = make()
    a.train(n_steps=2)
    a.save_checkpoint(str(tmp_path / "h.pt"))
    a.train(n_steps=2)
    b = make()
    b.load_checkpoint(str(tmp_path / "h.pt"))
    b.train(n_steps=2)
    ta = torch.nn.utils.parameters_to_vector(a.policy.parameters()).detach().numpy()
    tb = torch.nn.utils.parameters_to_vector(b.policy.parameters()).detach().numpy()
    np.testing.assert_array_equal(ta, tb)             # incl. the SGD momentum buffers restored
    assert a.best_reward == b.best_reward


# ------------------------------------------------------------------ folded post-update rollout (host logic)
def _tc_es(log_interval, seen=None, n=16, precision="auto"):
    """Fused ES on the emulated tensor-core backend (the modes that fold the post-update rollout)."""
    dims = [64, 64, 32]
    g = torch.Generator().manual_seed(2)
    obs, tgt = torch.randn(256, 64, generator=g), torch.randn(256, 32, generator=g)

    class Q(E.ES):
        def log(self):
            if seen is not None:
                seen.append((self.step, self.episode_reward, self.best_reward))
    torch.manual_seed(4)
    be = OracleBackend(tensor_core=True)
    es = Q(MLP, E.DeviceAgent, torch.optim.Adam, population_size=n, sigma=0.02, policy_kwargs={"dims": dims},
           agent_kwargs=dict(obs=obs, target=tgt), optimizer_kwargs={"lr": 0.01}, noise_table_size=1 << 14,
           log_interval=log_interval, eval_precision=precision, _backend=be)
    assert es._fused and es._precision == ("f16" if precision == "auto" else precision)
    return es, be


@pytest.mark.parametrize("precision", ["auto", "bf16s"])
def test_deferred_post_update_rollout_is_equivalent_cpu(precision):
    """estorch.py:181-185 runs one rollout of the updated policy per generation.  With
    log_interval > 1 that rollout is folded into the next generation's evaluate launch;
    everything observable must equal the log_interval = 1 run."""
    out = {}
    for li in (1, 3):
        seen = []
        es, be = _tc_es(li, seen, precision=precision)
        es.train(n_steps=7)
        out[li] = dict(theta=es._slots[0].theta.clone(), best=es._slots[0].best_theta.clone(), seen=seen,
                       ep=es.episode_reward, br=es.best_reward, ret=es.population_returns.copy(), folds=be.centre_folds)
    a, b = out[1], out[3]
    assert a["folds"] == 0 and b["folds"] == 4          # generations 0, 1, 3, 4 defer; 2, 5 log; 6 is the last
    assert torch.equal(a["theta"], b["theta"]) and torch.equal(a["best"], b["best"])
    assert a["ep"] == b["ep"] and a["br"] == b["br"]
    np.testing.assert_array_equal(a["ret"], b["ret"])
    assert [s for s, _, _ in b["seen"]] == [2, 5] and len(a["seen"]) == 7
    for step, ep, br in b["seen"]:
        assert (step, ep, br) == a["seen"][step]


def test_observing_mid_training_flushes_the_deferred_rollout():
    """Reading episode_reward / best_reward / state_dict() while a rollout is still deferred
    runs it first (nobody may see the previous generation's value)."""
    ref, _ = _tc_es(1)
    ref.train(n_steps=1)
    es, be = _tc_es(100)
    es.n_steps, es.step = 10, 0
    es._fused_generation(es._slots[0])            # what _master does for one generation
    assert es._pending_centre
    assert es.episode_reward == ref.episode_reward and not es._pending_centre
    assert es.best_reward == ref.best_reward
    es.step += 1; es._generation += 1             # ... and what it does after it
    es._fused_generation(es._slots[0])
    assert es._pending_centre and be.centre_folds == 0
    sd = es.state_dict()
    assert not es._pending_centre
    ref.train(n_steps=1)                           # a second generation (the generation counter keeps running)
    assert sd["slots"][0]["state"]["episode_reward"] == ref.episode_reward
    assert torch.equal(sd["slots"][0]["theta"], ref._slots[0].theta)
