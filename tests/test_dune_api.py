"""The dune.grid-style façade (dune_grid_b200/dune_api.py): names and argument meaning of
python/dune/grid/_grids.py / core.py / mapper.hh. Host-side behaviour on CPU, mapper.communicate on the GPU against the
oracle (including the reference's add-falls-through-to-set behaviour, dune/python/grid/mapper.hh:124-137)."""
import numpy as np
import pytest

from dune_grid_b200.dune_api import (CommOp, cartesianDomain, equidistantCoordinates, equidistantOffsetCoordinates,
                                     structuredGrid, tensorProductCoordinates, yaspGrid)
from tests.common import best_oracle, make_cfg, oracle_world


def test_constructors_and_sizes_match_the_oracle():
    o = best_oracle()
    cases = [
        (yaspGrid(cartesianDomain([0, 0], [1, 2], [8, 4])), make_cfg(dim=2, N=(8, 4), mode=1, upper=(1.0, 2.0))),
        (yaspGrid(equidistantCoordinates([1.0, 2.0, 3.0], [6, 5, 4]), periodic=[True, False, True], overlap=2),
         make_cfg(dim=3, N=(6, 5, 4), upper=(1.0, 2.0, 3.0), periodic=5, overlap=2)),
        (yaspGrid(equidistantOffsetCoordinates([-1, 0], [1, 1], [8, 4])), make_cfg(dim=2, N=(8, 4), mode=1, lower=(-1.0, 0.0), upper=(1.0, 1.0))),
        (yaspGrid(tensorProductCoordinates([[0, .25, .5, 1], [0, .5, 1]])), make_cfg(dim=2, N=(3, 2), mode=2, tensor=[np.array([0, .25, .5, 1.]), np.array([0, .5, 1.])])),
        (structuredGrid([0, 0, 0], [1, 1, 1], [4, 4, 4]), make_cfg(dim=3, N=(4, 4, 4), mode=1)),
        (yaspGrid(cartesianDomain([0, 0], [1, 1], [9, 8], periodic=[True, True], overlap=1)), make_cfg(dim=2, N=(9, 8), mode=1, periodic=3)),
    ]
    for view, cfg in cases:
        w = oracle_world(o, cfg, 1)
        d = cfg["dim"]
        assert view.dimension == d and view.comm.rank == 0 and view.comm.size == 1
        for c in range(d + 1):
            assert view.size(c) == w.size(0, 0, c)
            assert view.interiorPartition.size(c) <= view.allPartition.size(c)
        for layout, block in ((lambda gt: gt.dim == d, [1] + [0] * d), (lambda gt: gt.dim == 0, [0] * d + [1]),
                              (lambda gt: 2 if gt.dim == d - 1 else 1, [1, 2] + [1] * (d - 1))):
            m = view.mapper(layout)
            assert len(m) == m.size == w.mapper_size(0, 0, block)
        view.hierarchicalGrid.globalRefine(1)
        assert view.hierarchicalGrid.maxLevel == 1
        assert view.size(0) == view.hierarchicalGrid.leafView.size(0) > view.hierarchicalGrid.levelView(0).size(0)
        w.close()


def test_argument_errors_like_the_reference():
    with pytest.raises(ValueError):
        yaspGrid("not a constructor")                                        # _grids.py:283
    with pytest.raises(ValueError):
        yaspGrid(equidistantCoordinates([1, 1], [4, 4]), dimgrid=3)          # _grids.py:252
    with pytest.raises(ValueError):
        tensorProductCoordinates([[0, 1], [0, 1]], offset=[0])               # _grids.py:191
    with pytest.raises(Exception):
        tensorProductCoordinates([[0, 1, 0.5], [0, 1]]) and yaspGrid(tensorProductCoordinates([[0, 1, 0.5], [0, 1]]))   # not monotone


@pytest.mark.gpu
def test_mapper_communicate_through_the_facade():
    import torch
    o = best_oracle()
    cfg = make_cfg(dim=3, N=(7, 6, 5), mode=1, periodic=5, overlap=1)
    view = yaspGrid(cartesianDomain([0, 0, 0], [1, 1, 1], [7, 6, 5]), periodic=[True, False, True])
    w = oracle_world(o, cfg, 1)
    layout = [1, 0, 0, 1]
    m = view.mapper(lambda gt: gt.dim in (0, 3))
    assert len(m) == w.mapper_size(0, 0, layout)
    rng = np.random.default_rng(5)
    pairs = [("interiorBorderPartition", "allPartition", 1, 0), ("allPartition", "interiorBorderPartition", 1, 1),
             ("interiorBorderPartition", "interiorBorderPartition", 0, 0), ("overlapPartition", "allPartition", 3, 0),
             ("allPartition", "allPartition", 4, 0), ("overlapFrontPartition", "overlapPartition", 2, 1)]
    for frm, to, iftype, direction in pairs:
        for op in (CommOp.set, CommOp.add):
            a, b = rng.standard_normal(len(m)), rng.standard_normal((len(m), 2))
            ref = [[a.copy(), b.copy().reshape(-1)]]
            if op == CommOp.add:                                  # mapper.hh:124-137: add, then (no break) set
                w.communicate(0, layout, iftype, direction, 1, ref, [1, 2])
            w.communicate(0, layout, iftype, direction, 0, ref, [1, 2])
            tb = torch.from_numpy(b.copy()).cuda()
            m.communicate(getattr(view, frm), getattr(view, to), op, a, tb)        # NumPy array in place + CUDA tensor
            assert np.array_equal(a, ref[0][0]), (frm, to, op)
            assert np.array_equal(tb.cpu().numpy().reshape(-1), ref[0][1]), (frm, to, op)
    with pytest.raises(TypeError):
        m.communicate(view.interiorPartition, view.ghostPartition, CommOp.set, np.zeros(len(m)))
    assert view.coordinates().shape == (view.size(3), 3)
    w.close()
