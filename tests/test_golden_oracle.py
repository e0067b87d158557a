"""The oracle reproduces the committed fixtures (tests/golden/make_golden.py made them).
These are regression pins of the restatement, not reference-issued vectors: the reference has none
for this path.  Tolerance 1e-6 relative: the only arithmetic that may differ
between hosts is libm's sin/cos/atan2/acos (ifunc-selected variants can differ in the last ulp)."""
import numpy as np
import pytest

from oracle import Oracle
from tests.golden import make_golden as G

RTOL, ATOL = 1e-6, 1e-7


@pytest.mark.parametrize("name", sorted(G.cases().keys()))
def test_oracle_matches_golden(name):
    sc, gen, params, steps = G.cases()[name]
    gold = G.load(name)
    res = G.run_case(Oracle(), None, sc, gen, params, steps)
    assert len(gold["states"]) == steps
    for k in range(steps):
        assert np.allclose(res["states"][k]["position"], gold["states"][k]["position"], rtol=RTOL, atol=ATOL), (name, k)
        assert np.allclose(res["states"][k]["velocity"], gold["states"][k]["velocity"], rtol=RTOL, atol=ATOL), (name, k)
        assert res["impulses"][k].shape == gold["impulses"][k].shape
        assert np.allclose(res["impulses"][k], gold["impulses"][k], rtol=RTOL, atol=ATOL), (name, k)
    if len(sc.joints):
        assert np.allclose(res["joints"]["impulses"], gold["joint_impulses"], rtol=RTOL, atol=ATOL)
        assert np.array_equal(res["joints"]["broken"], gold["joint_broken"])
