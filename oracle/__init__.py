"""CPU oracle for the ALD reconstruction hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain torch-CPU / numpy restatement of the reference algorithm
(10258392511/InverseProblemWithDiffusionModel), each function citing the reference file:line
it follows.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline /
`--impl reference` legs may import this package, and only as the checker (or as the thing timed
as "the reference on the host cores").  The product package
`inverseproblemwithdiffusionmodel_b200` never imports it and has no CPU fallback.

Parity status: the reference ships no tests or golden vectors, so the oracle
is pinned against outputs of the reference itself, generated in the authoring container by
`oracle/make_golden.py` and committed under `tests/golden/` (checked by
`tests/test_oracle_golden.py`).
"""
