"""Dune::GlobalIndexSet (dune/grid/utility/globalindexset.hh). CPU part: the self-contained
restatement (oracle/port) against the reference's own class running in oracle/_ref. GPU part: dgb_global_index on one
rank (plain and periodic grids, every codim) against the oracle; the multi-rank case runs in tests/multigpu_check.py."""
import numpy as np
import pytest

from oracle import oracle as orc
from tests.common import best_oracle, make_cfg, oracle_world, product_grid

CASES = [
    (make_cfg(dim=2, N=(8, 4)), 1), (make_cfg(dim=2, N=(8, 4)), 2), (make_cfg(dim=3, N=(8, 4, 4)), 2),
    (make_cfg(dim=3, N=(12, 10, 9), overlap=2), 4), (make_cfg(dim=3, N=(8, 8, 8)), 8),
    (make_cfg(dim=2, N=(13, 7), periodic=3, refine=1), 6), (make_cfg(dim=3, N=(6, 6, 6), periodic=7), 8),
    (make_cfg(dim=3, N=(6, 5, 4), periodic=5), 1), (make_cfg(dim=3, N=(9, 8, 8), overlap=0), 4),
]


@pytest.mark.skipif(not (orc.available("ref") and orc.available("port")), reason="needs both oracles")
@pytest.mark.parametrize("cfg,P", CASES, ids=[str(i) for i in range(len(CASES))])
def test_port_restatement_equals_the_reference_class(cfg, P):
    wr, wp = oracle_world(orc.load("ref"), cfg, P), oracle_world(orc.load("port"), cfg, P)
    for lvl in range(cfg["refine"] + 1):
        for codim in range(cfg["dim"] + 1):
            a, na = wr.global_index(lvl, codim)
            b, nb = wp.global_index(lvl, codim)
            assert na == nb
            for r in range(P):
                assert np.array_equal(a[r], b[r]), (lvl, codim, r)
            # without periodic wrap every global index 0..size-1 occurs (with it the reference's class lets the copies of a
            # wrapped entity compete: mirrored here, not "fixed")
            if cfg["periodic"] == 0:
                allv = np.concatenate(a)
                assert set(allv[allv >= 0].tolist()) == set(range(na))
    wr.close(); wp.close()


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", [c for c, P in CASES if P == 1] + [make_cfg(dim=3, N=(17, 9, 12), periodic=2, refine=1),
                                                                  make_cfg(dim=2, N=(33, 20), periodic=1, overlap=2)])
def test_global_index_single_rank(cfg):
    o = best_oracle()
    w = oracle_world(o, cfg, 1)
    g = product_grid(cfg, 0, 1)
    for lvl in range(cfg["refine"] + 1):
        gv = g.levelGridView(lvl)
        for codim in range(cfg["dim"] + 1):
            ref, n = w.global_index(lvl, codim)
            got, m = gv.globalIndexSet(codim)
            assert m == n
            assert np.array_equal(got.cpu().numpy(), ref[0]), (lvl, codim)
    w.close()


@pytest.mark.skipif(not (orc.available("ref") and orc.available("port")), reason="needs both oracles")
def test_port_restatement_random_sweep():
    """seeded random non-periodic configurations: the restatement equals the reference's class on every rank, level and codim"""
    rng = np.random.default_rng(20261021)
    ref, port = orc.load("ref"), orc.load("port")
    done = 0
    for _ in range(40):
        dim = int(rng.integers(2, 4))
        N = tuple(int(x) for x in rng.integers(4, 11, size=dim))
        cfg = make_cfg(dim=dim, N=N, mode=int(rng.integers(0, 3)), overlap=int(rng.integers(0, 3)), refine=int(rng.integers(0, 2)),
                       keep=int(rng.integers(0, 2)))
        P = int(rng.choice([1, 2, 3, 4, 6, 8]))
        try:
            wr = oracle_world(ref, cfg, P)
        except Exception:
            continue
        if any(n < t for n, t in zip(N, wr.torus_dims())):
            wr.close()                                    # a rank without cells (overlap 0 lets the reference build such grids): its
            continue                                      # entities are never visited and the reference's result is whatever std::vector<int>(n) holds
        wp = oracle_world(port, cfg, P)
        for lvl in range(cfg["refine"] + 1):
            for codim in range(dim + 1):
                a, na = wr.global_index(lvl, codim)
                b, nb = wp.global_index(lvl, codim)
                assert na == nb and all(np.array_equal(x, y) for x, y in zip(a, b)), (cfg, P, lvl, codim)
        wr.close(); wp.close()
        done += 1
    assert done >= 15
