"""CPU: the oracles against the committed golden vectors (tests/golden/*.npz, generated from the REFERENCE's own code
by tests/golden/make_golden.py). For oracle/port this is the pin to reference outputs; for oracle/_ref it checks that
the build in this container still reproduces them. Integers and all fp64 arrays must match bit for bit (both oracles
are built with -ffp-contract=off; sin() comes from the same libm in GeometryGrid's deformation)."""
import glob
import os

import numpy as np
import pytest

from tests.common import oracle_kinds
from tests.golden_cases import CASES, snapshot
from oracle import oracle as orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_fixtures_present():
    names = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz")))
    assert names == sorted(CASES), "regenerate with tests/golden/make_golden.py"


@pytest.mark.parametrize("kind", ["ref", "port"])
@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_reproduces_golden(kind, name):
    if kind not in oracle_kinds():
        pytest.skip(f"oracle '{kind}' is not built here")
    gold = np.load(os.path.join(GOLDEN, name + ".npz"))
    snap = snapshot(orc.load(kind), name)
    assert sorted(snap) == sorted(gold.files)
    for k in gold.files:
        a, b = snap[k], gold[k]
        assert a.shape == b.shape and a.dtype == b.dtype, k
        if a.dtype.kind == "f":
            # libm's sin may differ between the box that generated the fixtures and this one: GeometryGrid arrays
            # are compared to 1e-14 relative, everything else bit for bit
            if "/geogrid/" in k:
                scale = max(1.0, np.abs(b).max(initial=0.0))
                assert np.abs(a - b).max(initial=0.0) <= 1e-14 * scale, k
            else:
                assert np.array_equal(a, b, equal_nan=True), k
        else:
            assert np.array_equal(a, b), k
