"""Process launcher: the torchrun analogue of the reference's ``_fork``
(estorch.py:41-56), which re-executes the *calling script* under
``mpirun -np N`` guarded by the ``MPI_PARENT`` environment variable.  Here the
script is re-executed under ``python -m torch.distributed.run`` with one
process per GPU, guarded by ``ESTORCH_B200_PARENT``."""
import inspect
import os
import socket
import subprocess
import sys


def _free_port():
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def fork_under_torchrun(n_proc: int) -> bool:
    """Return True in the parent after the children finished (the caller then
    exits, like estorch.py:305); False inside a child."""
    if os.getenv("ESTORCH_B200_PARENT") is not None or os.getenv("WORLD_SIZE") is not None:
        return False
    frame = inspect.stack()[2]                 # the user's script calling train()
    module = inspect.getmodule(frame[0])
    script = os.path.abspath(module.__file__)
    env = os.environ.copy()
    env["ESTORCH_B200_PARENT"] = "1"
    command = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
               f"--nproc-per-node={n_proc}", "--master-addr", "127.0.0.1",
               "--master-port", str(_free_port()), script] + sys.argv[1:]
    subprocess.call(command, env=env)
    return True
