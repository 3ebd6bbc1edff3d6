"""BackupRestoreFacility<YaspGrid> (dune/grid/yaspgrid/backuprestore.hh): the backup text of the
product must equal, byte for byte, what the reference's facility writes (oracle/_ref runs the reference's own
backuprestore.hh; golden texts generated from it are committed under tests/golden/backup/), and a grid restored from such
a text must equal the grid the reference restores, descriptor for descriptor. Host arithmetic only: no GPU needed."""
import glob
import json
import os

import numpy as np
import pytest

import dune_grid_b200 as dg
from tests.common import best_oracle, compare_world, make_cfg, oracle_world, product_grid

MODES = ("equidistant", "equidistantoffset", "tensorproduct")
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "backup")

CASES = [
    (make_cfg(dim=2, N=(8, 4)), 1), (make_cfg(dim=3, N=(8, 4, 4)), 2),
    (make_cfg(dim=2, N=(13, 7), periodic=3, refine=2), 6),
    (make_cfg(dim=3, N=(12, 10, 9), periodic=5, overlap=2, upper=(1.0, 2.0 / 3.0, 0.3)), 4),
    (make_cfg(dim=2, N=(8, 4), mode=1, lower=(-1.0, -0.5), upper=(1.0, 1.0 / 3.0), refine=1, keep=0), 2),
    (make_cfg(dim=3, N=(5, 4, 3), mode=1, lower=(0.1, -2.0, 3.0), upper=(0.7, 1.0, 4.5), periodic=4, refine=1), 1),
    (make_cfg(dim=3, N=(8, 4, 4), mode=2), 2),
    (make_cfg(dim=3, N=(9, 6, 5), periodic=7, mode=2, overlap=2, refine=1), 2),
    (make_cfg(dim=2, N=(8, 4), mode=2, refine=2, keep=0), 2),
    (make_cfg(dim=3, N=(8, 8, 8), partitioner=2, fixed_dims=(1, 2, 2), refine=1, keep=0), 4),
]
IDS = [str(i) for i in range(len(CASES))]


@pytest.mark.parametrize("cfg,P", CASES, ids=IDS)
def test_backup_text_equals_the_reference(cfg, P):
    o = best_oracle()
    w = oracle_world(o, cfg, P)
    for r in range(P):
        assert product_grid(cfg, r, P).backup() == w.backup(r), (r,)
    w.close()


def test_backup_text_equals_golden_files():
    """texts written by the reference's own BackupRestoreFacility (tests/golden/make_golden.py), committed"""
    files = sorted(glob.glob(os.path.join(GOLDEN, "*.json")))
    assert files, "golden backup texts missing"
    for f in files:
        rec = json.load(open(f))
        cfg = rec["cfg"]
        if cfg["tensor"] is not None:
            cfg["tensor"] = [np.array(t) for t in cfg["tensor"]]
        for r, text in enumerate(rec["texts"]):
            assert product_grid(cfg, r, rec["P"]).backup() == text, (f, r)


@pytest.mark.parametrize("cfg,P", CASES, ids=IDS)
def test_restored_grid_equals_the_reference(cfg, P):
    o = best_oracle()
    w0 = oracle_world(o, cfg, P)
    texts = [w0.backup(r) for r in range(P)]
    w0.close()
    w = o.restore(cfg["dim"], cfg["mode"], texts)
    compare_world(w, lambda r: dg.YaspGrid.restore(texts[r], MODES[cfg["mode"]], r, P), cfg["dim"], P, cfg["refine"] + 1)
    w.close()


def test_restore_rejects_bad_input():
    with pytest.raises(Exception):
        dg.YaspGrid.restore("YaspGrid BackupRestore Format Version: 1\n", "equidistant")
    with pytest.raises(Exception):
        dg.YaspGrid.restore("YaspGrid BackupRestore Format Version: 2\nTorus structure: 1 1 \nRefinement", "equidistant")
    text = product_grid(make_cfg(dim=3, N=(8, 4, 4)), 0, 1).backup()
    with pytest.raises(Exception):                       # stored torus 1x1x1 cannot serve 2 ranks (torus.hh:87)
        dg.YaspGrid.restore(text, "equidistant", 0, 2)


def test_backup_files_follow_the_reference_naming(tmp_path):
    g = product_grid(make_cfg(dim=2, N=(8, 4)), 0, 1)
    g.backup(str(tmp_path / "eq"))
    r = dg.YaspGrid.restore(str(tmp_path / "eq"), "equidistant", is_file=True)
    assert r.backup() == g.backup()
    cfg = make_cfg(dim=2, N=(8, 4), mode=2)
    for rank in range(2):
        product_grid(cfg, rank, 2).backup(str(tmp_path / "tp"))
    assert os.path.exists(tmp_path / "tp0") and os.path.exists(tmp_path / "tp1")
    r1 = dg.YaspGrid.restore(str(tmp_path / "tp"), "tensorproduct", 1, 2, is_file=True)
    assert r1.backup() == product_grid(cfg, 1, 2).backup()


def test_backup_restore_random_sweep():
    """seeded random configurations (all coordinate containers, periodicity, overlaps, refinement options, rank counts):
    backup text byte-equal to the reference's on every rank, restored grid equal to the reference's restored grid"""
    rng = np.random.default_rng(20261020)
    o = best_oracle()
    done = 0
    for _ in range(50):
        dim = int(rng.integers(2, 4))
        N = tuple(int(x) for x in rng.integers(4, 11, size=dim))
        cfg = make_cfg(dim=dim, N=N, mode=int(rng.integers(0, 3)), periodic=int(rng.integers(0, 1 << dim)),
                       overlap=int(rng.integers(0, 3)), refine=int(rng.integers(0, 3)), keep=int(rng.integers(0, 2)),
                       lower=tuple(rng.uniform(-1, 0, dim)), upper=tuple(rng.uniform(0.5, 2, dim)))
        P = int(rng.choice([1, 2, 3, 4, 6]))
        try:
            w0 = oracle_world(o, cfg, P)
        except Exception:
            continue                                      # the reference rejects the combination (covered in test_host_descriptors)
        texts = [w0.backup(r) for r in range(P)]
        w0.close()
        for r in range(P):
            assert product_grid(cfg, r, P).backup() == texts[r], (cfg, P, r)
        try:
            w = o.restore(cfg["dim"], cfg["mode"], texts)
        except Exception:
            # six significant digits can make the restored tensor coordinates non-monotone or the torus too fine for
            # the coarse size: the product must reject the text too
            with pytest.raises(Exception):
                dg.YaspGrid.restore(texts[0], MODES[cfg["mode"]], 0, P)
            continue
        compare_world(w, lambda r: dg.YaspGrid.restore(texts[r], MODES[cfg["mode"]], r, P), cfg["dim"], P, cfg["refine"] + 1)
        w.close()
        done += 1
    assert done >= 20
