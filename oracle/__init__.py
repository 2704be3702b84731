"""CPU oracle for the MargiPose hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain-PyTorch (CPU, fp32/fp64) restatement of the reference algorithm for the
path named in BASELINE.json.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import this package,
and only as the checker (or the timed CPU baseline) -- never as part of what
`margipose_b200` ships or measures as its own.

Parity status: PINNED.  Every function here is checked against the unmodified
reference (imported through `oracle/ref_shim.py` where /root/reference exists)
by `tests/test_oracle_pin.py`, and against committed golden vectors generated
from the reference by `tests/golden/make_golden.py` (`tests/test_oracle_golden.py`),
including the reference's own known-answer test
(`/root/reference/tests/test_models.py:39-46`).
"""
