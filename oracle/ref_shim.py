"""Import shims that let the UNMODIFIED reference package (goktug97/estorch, installed by
``__graft_entry__.build()`` into ``baseline/_ref`` -- git-ignored, never copied into this repo's
sources) run in a single process: a one-rank ``mpi4py`` stand-in (the reference imports it at
estorch.py:10; with ``n_proc=1`` no MPI call is made, :207-209/:228-233 loop over zero workers) and
``np.int = int`` (estorch.py:453 uses the alias numpy removed).  TEST / BENCH INFRASTRUCTURE ONLY:
used by tests/golden/make_golden.py-style scripts and by ``bench.py --impl reference``."""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "baseline", "_ref")


def install_shims():
    if "mpi4py" not in sys.modules:
        m = types.ModuleType("mpi4py")
        MPI = types.ModuleType("mpi4py.MPI")

        class _Comm:
            def Get_rank(self):
                return 0

            def Get_size(self):
                return 1

            def send(self, *a, **k):
                pass

            def bcast(self, x, root=0):
                return x

        class Status:
            def Get_tag(self):
                return 0

        MPI.COMM_WORLD, MPI.Status, m.MPI = _Comm(), Status, MPI
        sys.modules["mpi4py"], sys.modules["mpi4py.MPI"] = m, MPI
    if not hasattr(np, "int"):
        np.int = int


def import_reference():
    """The reference's ``estorch`` module from baseline/_ref, or None when it is not installed."""
    if not os.path.isdir(os.path.join(REF_DIR, "estorch")):
        return None
    install_shims()
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    try:
        import estorch  # noqa: F401  (the reference)
    except Exception:
        return None
    return sys.modules["estorch"]
