"""Worker for tests/test_dist_cpu.py: one rank of a world_size-N gloo job running
the fused ES host logic with the oracle stand-in backend.  Writes its final theta."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from _oracle_backend import OracleBackend  # noqa: E402
import estorch_b200 as E  # noqa: E402
from test_api_cpu import MLP, _load_theta  # noqa: E402


def main():
    out_dir, algo = sys.argv[1], sys.argv[2]
    if algo == "es_fold":
        # bf16s on the emulated tensor-core backend, log_interval 3: the post-update rollout of
        # most generations rides in the next generation's evaluate launch, on every rank
        from test_api_cpu import _tc_es
        seen = []
        es, be = _tc_es(3, seen)
        es.train(n_steps=5)
        np.savez(os.path.join(out_dir, f"rank{es.rank}.npz"), theta=es._slots[0].theta.numpy(),
                 best=es._slots[0].best_theta.numpy(), step=es.step, n_logs=len(seen), folds=be.centre_folds,
                 returns=es.population_returns, world=es.n_workers, pairs_local=es._pairs_local,
                 pair_begin=es._pair_begin, episode=es.episode_reward, best_reward=es.best_reward)
        return
    if algo in ("es_two_calls", "es_one_call"):
        # two consecutive train() calls of a multi-rank job (the second one does not broadcast the replicas again)
        # must leave what one call of the same length leaves
        dims = [4, 16, 2]
        gg = torch.Generator().manual_seed(5)
        obs, tgt = torch.randn(32, 4, generator=gg), torch.randn(32, 2, generator=gg)
        torch.manual_seed(100 + int(os.environ["RANK"]))      # different construction-time policies per rank
        es = E.ES(MLP, E.DeviceAgent, torch.optim.Adam, population_size=16, sigma=0.05, policy_kwargs={"dims": dims},
                  agent_kwargs=dict(obs=obs, target=tgt), optimizer_kwargs={"lr": 0.01}, noise_table_size=1 << 12,
                  noise_seed=3, _backend=OracleBackend())
        es.log = lambda: None
        if algo == "es_two_calls":
            es.train(n_steps=3)
            es.train(n_steps=2)
        else:
            es.train(n_steps=5)
        np.savez(os.path.join(out_dir, f"rank{es.rank}.npz"), theta=es._slots[0].theta.numpy(),
                 m=es._slots[0].m.numpy(), step=es.step, returns=es.population_returns,
                 episode=float(es.episode_reward), best=float(es.best_reward), synced=bool(es._replicas_synced))
        return
    if algo == "es_p8192":
        # BASELINE config 3's population (8192 members = 4096 antithetic pairs) sharded over the ranks, small policy
        dims = [4, 16, 2]
        gg = torch.Generator().manual_seed(5)
        obs, tgt = torch.randn(32, 4, generator=gg), torch.randn(32, 2, generator=gg)
        torch.manual_seed(21)
        es = E.ES(MLP, E.DeviceAgent, torch.optim.Adam, population_size=8192, sigma=0.02, policy_kwargs={"dims": dims},
                  agent_kwargs=dict(obs=obs, target=tgt), optimizer_kwargs={"lr": 0.01}, noise_table_size=1 << 14,
                  noise_seed=3, _backend=OracleBackend())
        es.log = lambda: None
        es.train(n_steps=2)
        np.savez(os.path.join(out_dir, f"rank{es.rank}.npz"), theta=es._slots[0].theta.numpy(), step=es.step,
                 returns=es.population_returns, ranks=es._ranks.numpy(), pairs_local=es._pairs_local,
                 pair_begin=es._pair_begin, episode=float(es.episode_reward))
        return
    if algo in ("es_unsynced", "nsra_unsynced", "ns_hooks_unsynced"):
        # Every rank seeds torch / numpy DIFFERENTLY and nothing preloads theta: what a user's script
        # does under torchrun.  The reference keeps one master copy (estorch.py:136, :401-408, :444-456);
        # here rank 0's construction-time state must win on every rank.
        rank = int(os.environ["RANK"])
        torch.manual_seed(1000 + rank)
        np.random.seed(77 + rank)
        dims = [4, 16, 2]
        gg = torch.Generator().manual_seed(5)
        obs, tgt = torch.randn(32, 4, generator=gg), torch.randn(32, 2, generator=gg)
        common = dict(population_size=16, sigma=0.05, policy_kwargs={"dims": dims}, optimizer_kwargs={"lr": 0.01},
                      noise_table_size=1 << 12, noise_seed=3, _backend=OracleBackend())
        if algo == "es_unsynced":
            es = E.ES(MLP, E.DeviceAgent, torch.optim.Adam, agent_kwargs=dict(obs=obs, target=tgt), **common)
        elif algo == "nsra_unsynced":
            es = E.NSRA_ES(MLP, E.DeviceAgent, torch.optim.Adam, weight_t=1,
                           agent_kwargs=dict(obs=obs, target=tgt, bc_obs=8, bc_dim=16), **common)
        else:
            class Noisy:                      # a host agent whose rollouts differ from rank to rank
                def __init__(self):
                    self.rng = np.random.RandomState(500 + rank)

                def rollout(self, policy):
                    with torch.no_grad():
                        out = policy(obs)
                    r = float(-((out - tgt) ** 2).mean()) + 1e-3 * float(self.rng.randn())
                    return r, out[:8].flatten()[:16].numpy().copy() + 1e-3 * self.rng.randn(16).astype(np.float32)
            es = E.NS_ES(MLP, Noisy, torch.optim.Adam, **common)
            assert not es._fused
        es.log = lambda: None
        es.train(n_steps=3)
        mods = [es.policy] if algo == "es_unsynced" else [p for p, _ in es.meta_population]
        theta = np.stack([torch.nn.utils.parameters_to_vector(m.parameters()).detach().numpy() for m in mods])
        extra = {}
        if algo != "es_unsynced":
            extra = dict(archive=np.stack(es._archive), idx=int(es.idx), best=float(es.best_reward))
        np.savez(os.path.join(out_dir, f"rank{es.rank}.npz"), theta=theta, step=es.step,
                 returns=es.population_returns, episode=float(es.episode_reward), **extra)
        return
    g = np.load(os.path.join(ROOT, "tests", "golden",
                             "es_cartpole_p64.npz" if algo == "es" else "nsra_bipedal_p32.npz"))
    dims = [int(d) for d in g["dims"]]
    obs, tgt = torch.from_numpy(g["obs"]), torch.from_numpy(g["target"])
    seen = []

    if algo == "es":
        class R(E.ES):
            def log(self):
                seen.append(self.population_returns.copy())
                if self.step == 1:
                    self.terminate()       # must stop EVERY rank after generation 1
        es = R(MLP, E.DeviceAgent, torch.optim.Adam, population_size=64, sigma=0.1,
               policy_kwargs={"dims": dims}, agent_kwargs=dict(obs=obs, target=tgt),
               optimizer_kwargs={"lr": 0.01}, noise_table_size=len(g["table"]),
               noise_seed=int(g["noise_seed"]), _backend=OracleBackend())
        es._table.copy_(torch.from_numpy(g["table"]))
        _load_theta(es.policy, g["theta0"])
        es.train(n_steps=5)
        theta = torch.nn.utils.parameters_to_vector(es.policy.parameters()).detach().numpy()
    else:
        class R(E.NSRA_ES):
            def log(self):
                seen.append(self.population_returns.copy())
        np.random.seed(123)
        es = R(MLP, E.DeviceAgent, torch.optim.Adam, population_size=32, sigma=0.02, weight_t=2,
               policy_kwargs={"dims": dims}, agent_kwargs=dict(obs=obs, target=tgt, bc_obs=64, bc_dim=256),
               optimizer_kwargs={"lr": 0.01}, noise_table_size=len(g["table"]),
               noise_seed=int(g["noise_seed"]), _backend=OracleBackend())
        es._table.copy_(torch.from_numpy(g["table"]))
        for i, (p, _) in enumerate(es.meta_population):
            _load_theta(p, g["meta_theta0"][i])
        es._archive = [a.copy() for a in g["archive0"]]
        np.random.seed(123)
        es.train(n_steps=2)
        theta = np.stack([torch.nn.utils.parameters_to_vector(p.parameters()).detach().numpy()
                          for p, _ in es.meta_population])
    np.savez(os.path.join(out_dir, f"rank{es.rank}.npz"), theta=theta, step=es.step,
             n_logs=len(seen), returns=es.population_returns, world=es.n_workers,
             pairs_local=es._pairs_local, pair_begin=es._pair_begin)


if __name__ == "__main__":
    main()
