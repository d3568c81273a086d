"""CPU oracle for the ES hot path — TEST INFRASTRUCTURE ONLY.

Nothing under ``estorch_b200/`` may import this package.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs use it, and only as the checker / CPU baseline.
See ``oracle/es_oracle.py`` for the restatement and its pinning status.
"""
from .es_oracle import *  # noqa: F401,F403
