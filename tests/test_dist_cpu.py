"""Multi-process path (one process per GPU in production) exercised with
world_size=2 over gloo on CPU: pair sharding, all-gather of returns, all-reduce
of the partial gradient, rank-0-only logging and terminate() propagation."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from conftest import load_golden, rel_err

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run(world, algo, tmp_path):
    port = _free_port()
    procs = []
    for rank in range(world):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                   MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "_dist_worker.py"),
                                       str(tmp_path), algo], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT))
    outs = [p.communicate(timeout=600)[0].decode() for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o[-3000:]
    return [np.load(os.path.join(tmp_path, f"rank{r}.npz")) for r in range(world)]


def test_world2_es_sharded_generation(tmp_path):
    g = load_golden("es_cartpole_p64.npz")
    r0, r1 = _run(2, "es", tmp_path)
    assert int(r0["world"]) == 2 and int(r0["pairs_local"]) == 16
    assert int(r0["pair_begin"]) == 0 and int(r1["pair_begin"]) == 16
    # terminate() on rank 0 at step 1 stopped both ranks after 2 generations
    assert int(r0["step"]) == 2 and int(r1["step"]) == 2
    assert int(r0["n_logs"]) == 2 and int(r1["n_logs"]) == 0         # only rank 0 logs
    np.testing.assert_array_equal(r0["theta"], r1["theta"])            # replicas stay bit-identical
    np.testing.assert_array_equal(r0["returns"], r1["returns"])
    assert rel_err(r0["returns"][:, 0], g["returns"][1][:, 0]) < 1e-5
    assert rel_err(r0["theta"], g["theta_after"][1]) < 1e-4


def test_world2_nsra_sharded_generation(tmp_path):
    g = load_golden("nsra_bipedal_p32.npz")
    r0, r1 = _run(2, "nsra", tmp_path)
    np.testing.assert_array_equal(r0["theta"], r1["theta"])
    assert rel_err(r0["returns"][:, 1], g["returns"][1][:, 1]) < 1e-4


def test_world2_folded_post_update_rollout(tmp_path):
    """The deferred post-update rollout (folded into the next generation's evaluate launch) on a
    sharded population: ranks stay bit-identical and agree with the single-process run up to the
    fp32 summation order of the all-reduced gradient."""
    from test_api_cpu import _tc_es
    r0, r1 = _run(2, "es_fold", tmp_path)
    assert int(r0["world"]) == 2 and int(r0["pairs_local"]) == 4 and int(r1["pair_begin"]) == 4
    assert int(r0["folds"]) == 3 and int(r1["folds"]) == 3       # generations 0, 1, 3 defer; 2 logs; 4 is the last
    assert int(r0["n_logs"]) == 1 and int(r1["n_logs"]) == 0
    for k in ("theta", "best", "returns", "episode", "best_reward"):
        np.testing.assert_array_equal(r0[k], r1[k])
    es, _ = _tc_es(3)
    es.train(n_steps=5)
    assert rel_err(r0["theta"], es._slots[0].theta.numpy()) < 1e-5
    assert abs(float(r0["episode"]) - es.episode_reward) < 1e-5 * abs(es.episode_reward)


@pytest.mark.parametrize("algo", ["es_unsynced", "nsra_unsynced", "ns_hooks_unsynced"])
def test_world2_replicas_agree_without_preloaded_parameters(tmp_path, algo):
    """Each rank builds its own policy / meta-population from a different torch seed (and, for the
    hooks-mode NS run, has a host agent with rank-dependent noise and its own numpy RNG): rank 0's
    state is broadcast before the loop and only rank 0 selects the meta-policy / feeds the archive,
    as the reference's single master does (estorch.py:136, :401-408, :444-456, :458-471)."""
    r0, r1 = _run(2, algo, tmp_path)
    assert int(r0["step"]) == 3 and int(r1["step"]) == 3
    np.testing.assert_array_equal(r0["theta"], r1["theta"])
    if algo != "es_unsynced":
        np.testing.assert_array_equal(r0["archive"], r1["archive"])
        assert int(r0["idx"]) == int(r1["idx"]) and float(r0["best"]) == float(r1["best"])
    if algo != "ns_hooks_unsynced":          # device agents are deterministic: identical returns everywhere
        np.testing.assert_array_equal(r0["returns"], r1["returns"])
        assert float(r0["episode"]) == float(r1["episode"])


def test_world2_two_train_calls_equal_one(tmp_path):
    """A second train() call of a multi-rank fused job skips the replica broadcast (the replicas are bit-identical
    by construction): 3 + 2 generations leave exactly what 5 leave, on both ranks."""
    (tmp_path / "a").mkdir(); (tmp_path / "b").mkdir()
    a0, a1 = _run(2, "es_two_calls", tmp_path / "a")
    b0, b1 = _run(2, "es_one_call", tmp_path / "b")
    assert bool(a0["synced"]) and bool(a1["synced"])
    for k in ("theta", "m", "returns"):
        np.testing.assert_array_equal(a0[k], a1[k])
        np.testing.assert_array_equal(a0[k], b0[k])
    assert float(a0["episode"]) == float(b0["episode"]) and float(a0["best"]) == float(b0["best"])


def test_world2_population_8192_sharded(tmp_path):
    """BASELINE config 3's population size (8192 members, sigma 0.02) sharded over two ranks: 2048 pairs per
    rank, rank-major all-gather of the returns, every rank ranks all 8192, bit-identical replicas, and the
    same update as the single-process run up to the fp32 summation order of the all-reduced gradient."""
    import torch
    import estorch_b200 as E
    from _oracle_backend import OracleBackend
    from test_api_cpu import MLP
    r0, r1 = _run(2, "es_p8192", tmp_path)
    assert int(r0["pairs_local"]) == 2048 and int(r1["pair_begin"]) == 2048
    for k in ("theta", "returns", "ranks", "episode"):
        np.testing.assert_array_equal(r0[k], r1[k])
    assert sorted(r0["ranks"].tolist()) == list(range(8192))
    gg = torch.Generator().manual_seed(5)
    obs, tgt = torch.randn(32, 4, generator=gg), torch.randn(32, 2, generator=gg)
    torch.manual_seed(21)
    es = E.ES(MLP, E.DeviceAgent, torch.optim.Adam, population_size=8192, sigma=0.02, policy_kwargs={"dims": [4, 16, 2]},
              agent_kwargs=dict(obs=obs, target=tgt), optimizer_kwargs={"lr": 0.01}, noise_table_size=1 << 14,
              noise_seed=3, _backend=OracleBackend())
    es.log = lambda: None
    es.train(n_steps=2)
    assert rel_err(r0["returns"], es.population_returns) < 1e-6     # generation 2 starts from a theta that differs in
    assert rel_err(r0["theta"], es._slots[0].theta.numpy()) < 1e-5  # the last bits (order of the all-reduce sum)


def test_train_n_proc_2_reexecs_under_torchrun_cpu(tmp_path):
    """``train(n_proc=2)`` from a plain ``python script.py`` (the reference's usage: ``_fork`` re-executes the
    calling script under mpirun, estorch.py:41-56, :305): the script is re-executed under
    ``torch.distributed.run`` with two ranks (gloo here, the oracle stand-in as backend), both ranks finish
    with the same parameters, rank 0 alone logs, the parent exits 0 and nobody hangs at exit."""
    import textwrap
    script = tmp_path / "user_script.py"
    script.write_text(textwrap.dedent(f"""
        import os, sys, numpy as np, torch
        sys.path.insert(0, {ROOT!r}); sys.path.insert(0, os.path.join({ROOT!r}, "tests"))
        from _oracle_backend import OracleBackend
        import estorch_b200 as E
        from test_api_cpu import MLP
        g = torch.Generator().manual_seed(1)
        obs, tgt = torch.randn(32, 4, generator=g), torch.randn(32, 2, generator=g)
        class Q(E.ES):
            def log(self):
                open(os.path.join({str(tmp_path)!r}, f"log_rank{{self.rank}}.txt"), "a").write(f"{{self.step}}\\n")
        es = Q(MLP, E.DeviceAgent, torch.optim.Adam, population_size=16, sigma=0.05, policy_kwargs={{"dims": [4, 16, 2]}},
               agent_kwargs=dict(obs=obs, target=tgt), optimizer_kwargs={{"lr": 0.01}}, noise_table_size=1 << 12,
               _backend=OracleBackend())
        es.train(n_steps=3, n_proc=2)
        np.save(os.path.join({str(tmp_path)!r}, f"theta_rank{{es.rank}}.npy"), es._slots[0].theta.numpy())
    """))
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "ESTORCH_B200_PARENT")}
    r = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stderr[-3000:]
    t0, t1 = np.load(tmp_path / "theta_rank0.npy"), np.load(tmp_path / "theta_rank1.npy")
    np.testing.assert_array_equal(t0, t1)
    assert (tmp_path / "log_rank0.txt").read_text().split() == ["0", "1", "2"]
    assert not (tmp_path / "log_rank1.txt").exists()
    # the parent (no WORLD_SIZE) never trains: it exits right after the children did
    assert not (tmp_path / "theta_rankNone.npy").exists()
