"""Host-side descriptor arithmetic of libdgb (no GPU): processor grid, nested partition boxes, index offsets,
send/recv lists, coordinates and refinement must equal the CPU oracle bit for bit.
With oracle/_ref built, the oracle IS the reference's own yaspgrid.hh running with threads as ranks."""
import itertools

import numpy as np
import pytest

from tests.common import best_oracle, make_cfg, oracle_world, product_grid

from tests.common import compare_world  # noqa: E402


def compare(cfg, P):
    o = best_oracle()
    w = oracle_world(o, cfg, P)
    compare_world(w, lambda r: product_grid(cfg, r, P), cfg["dim"], P, cfg["refine"] + 1)
    w.close()


CASES = [
    # the reference's own fixtures: 8x4(x4), overlap 1 (test/yasp/test-yaspgrid.hh:48-54), 1 and 2 ranks
    (make_cfg(dim=2, N=(8, 4)), 1), (make_cfg(dim=2, N=(8, 4)), 2),
    (make_cfg(dim=3, N=(8, 4, 4)), 1), (make_cfg(dim=3, N=(8, 4, 4)), 2),
    (make_cfg(dim=2, N=(8, 4), mode=1, lower=(-1.0, -0.5), upper=(1.0, 0.5)), 2),
    (make_cfg(dim=3, N=(8, 4, 4), mode=2), 2),
    # refinement with both overlap options (test-yaspgrid-yaspfactory-2d.cc:33-42)
    (make_cfg(dim=2, N=(8, 4), refine=2, keep=1), 2), (make_cfg(dim=2, N=(8, 4), refine=2, keep=0), 2),
    (make_cfg(dim=2, N=(8, 4), mode=2, refine=2, keep=0), 2), (make_cfg(dim=2, N=(8, 4), mode=1, refine=1, keep=0, overlap=2), 2),
    # periodic, uneven splits, wider overlap, several ranks
    (make_cfg(dim=3, N=(12, 10, 9), periodic=5, overlap=2), 4),
    (make_cfg(dim=3, N=(12, 10, 9), upper=(1.0, 2.0, 3.0)), 8),
    (make_cfg(dim=2, N=(13, 7), periodic=3), 1), (make_cfg(dim=2, N=(13, 7), periodic=3, refine=2), 6),
    (make_cfg(dim=3, N=(10, 6, 7), periodic=2, upper=(1.0, 2.0, 3.0), refine=1, keep=0), 4),
    (make_cfg(dim=3, N=(6, 6, 6), periodic=7), 8), (make_cfg(dim=3, N=(6, 6, 6), periodic=7, mode=2), 1),
    (make_cfg(dim=3, N=(9, 6, 5), periodic=7, mode=2, overlap=2, refine=1), 2),
    (make_cfg(dim=3, N=(9, 8, 8), overlap=0), 4), (make_cfg(dim=2, N=(9, 8), overlap=0, periodic=1), 3),
    (make_cfg(dim=3, N=(8, 8, 8), partitioner=2, fixed_dims=(1, 2, 2)), 4),
    (make_cfg(dim=2, N=(12, 12), partitioner=1), 9),
    (make_cfg(dim=3, N=(5, 4, 3), mode=1, lower=(0.1, -2.0, 3.0), upper=(0.7, 1.0, 4.5), periodic=4, refine=1), 1),
]


@pytest.mark.parametrize("cfg,P", CASES, ids=[f"{i}" for i in range(len(CASES))])
def test_descriptors_match_oracle(cfg, P):
    compare(cfg, P)


def test_descriptors_random_sweep():
    rng = np.random.default_rng(20261019)
    done = 0
    for _ in range(60):
        dim = int(rng.integers(2, 4))
        N = tuple(int(x) for x in rng.integers(4, 12, size=dim))
        cfg = make_cfg(dim=dim, N=N, mode=int(rng.integers(0, 3)), periodic=int(rng.integers(0, 1 << dim)),
                       overlap=int(rng.integers(0, 3)), refine=int(rng.integers(0, 2)), keep=int(rng.integers(0, 2)),
                       lower=tuple(rng.uniform(-1, 0, dim)), upper=tuple(rng.uniform(0.5, 2, dim)))
        P = int(rng.choice([1, 2, 3, 4, 6, 8]))
        o = best_oracle()
        try:
            w = oracle_world(o, cfg, P)
            w.close()
        except Exception:
            # the reference rejects this combination (GridError): the product must reject it too
            with pytest.raises(Exception):
                product_grid(cfg, 0, P)
            continue
        compare(cfg, P)
        done += 1
    assert done >= 20


def test_baseline_partitions():
    """the BASELINE configs"""
    import dune_grid_b200 as dg
    for P, dims in ((2, (2, 1, 1)), (4, (4, 1, 1)), (8, (2, 2, 2))):
        g = dg.YaspGrid(dim=3, N=(1024, 1024, 1024), upper=(1, 1, 1), rank=P - 1, nranks=P)
        assert g.torusDims() == dims
    g = dg.YaspGrid(dim=3, N=(1024, 1024, 1024), upper=(1, 1, 1), rank=0, nranks=8)
    gv = g.leafGridView()
    assert gv.size(0) == 513 ** 3 and gv.size(0, dg.Interior_Partition) == 512 ** 3
    l = g.commList(0, 0, 6)
    assert len(l) == 7 and int(np.prod(l[:, 7:10], axis=1).sum()) == 787969
