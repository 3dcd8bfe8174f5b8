"""CPU oracle for the CURIOUS training hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under `curious_b200/` may import this package.  It is imported by
`tests/`, by `__graft_entry__.smoke()` and by `bench.py`'s CPU-baseline legs, and
only as the checker / the timed CPU baseline - never as the product path.

What is pinned and what is not
------------------------------
* `her_oracle.py`, `replay_oracle.py` restate baselines/her/her.py and
  baselines/her/replay_buffer.py.  They are PINNED: `oracle/gen_golden.py` imports
  the unmodified reference modules from /root/reference (the reference is pure Python),
  runs them on seeded inputs and commits inputs, RNG stream and outputs under
  tests/golden/; tests/test_oracle_golden.py checks the restatement bit-for-bit.
* `reward_oracle.py` restates gym_flowers' `compute_reward`, a third-party package
  that is not vendored under /root/reference and is not version-pinned anywhere
  (reference baselines/her/experiment/config.py:3,158-159, readme.md:7).
  REWARD PARITY IS UNPINNED: only the call contract (argument names, shapes,
  (B,1) return, one-hot module selection, truncated ag slice) is pinned by the
  reference's own call sites.
* `ddpg_oracle.py` restates Normalizer / ActorCritic / MultiTaskActorCritic / DDPG graph /
  MpiAdam (TF1 + mpi4py, neither installable here).  PARITY UNPINNED by the reference's
  own tests (it has none for baselines/her/); the restatement is cross-checked against
  torch-CPU autograd and, for Adam, against the closed-form of the reference's
  `test_MpiAdam` problem (baselines/common/mpi_adam.py:52-78).
"""
