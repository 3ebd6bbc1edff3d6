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


def test_port_equals_reference_on_random_configurations():
    """beyond the committed fixtures: where both oracles are built, the restatement must equal the reference's own code on
    seeded random grids (descriptors, sweeps, communicate, FV) bit for bit"""
    if sorted(oracle_kinds()) != ["port", "ref"]:
        pytest.skip("needs both oracles")
    from tests.common import make_cfg, oracle_world
    from tests import golden_cases as gc
    rng = np.random.default_rng(20261019)
    done = 0
    for trial in range(40):
        dim = int(rng.integers(2, 4))
        N = tuple(int(x) for x in rng.integers(4, 9, size=dim))
        cfg = make_cfg(dim=dim, N=N, mode=int(rng.integers(0, 3)), periodic=int(rng.integers(0, 1 << dim)), overlap=int(rng.integers(0, 3)),
                       refine=int(rng.integers(0, 2)), keep=int(rng.integers(0, 2)), lower=tuple(rng.uniform(-1, 0, dim)),
                       upper=tuple(rng.uniform(0.5, 2, dim)))
        P = int(rng.choice([1, 2, 3, 4, 6, 8]))
        if P > 1 and cfg["overlap"] == 0:
            # the reference's own iterators return garbage entity counts for some empty boxes of overlap-0 partitions
            # (negative box sizes are not "empty", ygrid.hh:141-147): undefined upstream behaviour, not a parity target
            cfg["overlap"] = 1
        layout = [int(x) for x in rng.integers(0, 3, size=dim + 1)]
        layout[0] = max(layout[0], 1)
        name = f"random{trial}"
        gc.CASES[name] = (cfg, P, layout, (int(rng.integers(0, 2)), [1.0, -0.5, 0.25, 0.1, 0.5, 0.5, 0.5, 0.1, 0.45]))
        try:
            try:
                a = snapshot(orc.load("ref"), name)
            except Exception:
                with pytest.raises(Exception):          # the reference rejects this grid: so must the restatement
                    snapshot(orc.load("port"), name)
                continue
            b = snapshot(orc.load("port"), name)
        finally:
            del gc.CASES[name]
        assert sorted(a) == sorted(b)
        for k in a:
            assert a[k].shape == b[k].shape and np.array_equal(a[k], b[k], equal_nan=True), (trial, cfg, P, k)
        done += 1
    assert done >= 15
