"""CPU oracle for the IC3Net rollout hot path.  TEST INFRASTRUCTURE ONLY.

Everything under ``oracle/`` is a checker: a CPU restatement (numpy, float64)
of the reference algorithm for the rollout hot path
(``ic3net_envs/predator_prey_env.py``, ``ic3net_envs/traffic_junction_env.py``,
``ic3net_envs/traffic_helper.py``, ``env_wrappers.py``, ``comm.py``,
``action_utils.py``, ``trainer.py:26-126``).  Each function cites the reference
file:line it follows.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline
legs (``cpu_baseline`` and ``--impl reference``) may import this package.  The
product (``ic3net_b200/``) never imports it and has no CPU fallback: it raises
if the CUDA library is missing.

Parity status: PINNED.  The restatement is checked against the *unmodified*
reference (imported from /root/reference through ``oracle/ref_shims.py``) by
``oracle/gen_golden.py``, which also writes the committed fixtures under
``tests/golden/``; ``tests/test_oracle_golden.py`` re-checks the oracle against
those fixtures on every run (no /root/reference needed at test time).
"""
