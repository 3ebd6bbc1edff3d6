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


@pytest.mark.gpu
def test_bulk_geometry_names_of_the_python_bindings():
    """view.geometries(codim): toGlobal / jacobianTransposed / jacobianInverseTransposed / integrationElement / corners / center /
    volume / toLocal with the (vectorised) names of dune/python/grid/geometry.hh:168-349, against the oracle's per-entity geometry;
    coordinates() and polygons() of dune/python/grid/numpy.hh:46-70,184-212"""
    import torch
    o = best_oracle()
    for view, cfg in [(yaspGrid(tensorProductCoordinates([[0, .25, .5, 1], [0, .5, 1, 1.5, 3]]), periodic=[True, False]),
                       make_cfg(dim=2, N=(3, 4), mode=2, tensor=[np.array([0, .25, .5, 1.]), np.array([0, .5, 1., 1.5, 3.])], periodic=1)),
                      (yaspGrid(equidistantCoordinates([1.0, 2.0, 3.0], [6, 5, 4])), make_cfg(dim=3, N=(6, 5, 4), upper=(1.0, 2.0, 3.0)))]:
        w = oracle_world(o, cfg, 1)
        d = cfg["dim"]
        for codim in range(d + 1):
            m = d - codim
            geo = view.geometries(codim)
            rng = np.random.default_rng(codim)
            x = rng.random((3, max(1, m)))
            ref = w.geometry_eval_codim(0, 0, codim, x)
            assert len(geo) == ref["global"].shape[0] and geo.affine
            assert np.array_equal(geo.toGlobal(x).cpu().numpy().transpose(2, 0, 1), ref["global"])
            assert np.array_equal(geo.integrationElement(x).cpu().numpy().T, ref["detJ"])
            if m > 0:
                assert np.array_equal(geo.jacobianTransposed(x).cpu().numpy().transpose(3, 0, 1, 2), ref["jt"])
                assert np.array_equal(geo.jacobianInverseTransposed(x).cpu().numpy().transpose(3, 0, 1, 2), ref["jit"])
                assert np.array_equal(geo.jacobian(x).cpu().numpy().transpose(3, 0, 2, 1), ref["jt"])
                assert np.array_equal(geo.jacobianInverse(x).cpu().numpy().transpose(3, 0, 2, 1), ref["jit"])
                # the single-point form drops the point axis (geometry.hh: scalar overloads)
                assert np.array_equal(geo.toGlobal(x[0]).cpu().numpy().T, ref["global"][:, 0])
            e = w.entities(0, 0, codim, 4)
            assert np.array_equal(geo.center.cpu().numpy().T, e["center"]) and np.array_equal(geo.volume.cpu().numpy(), e["volume"])
            c = geo.corners.cpu().numpy()                                   # (2^m, d, n): corner 0 = lower, last = upper
            assert c.shape == (1 << m, d, len(geo))
            assert np.array_equal(c[0].T, e["lower"]) and np.array_equal(c[-1].T, e["upper"])
        cells = view.geometries(0)
        y = np.full(d, 0.3)
        loc = cells.toLocal(y).cpu().numpy()
        back = cells.toGlobal(np.zeros(d)).cpu().numpy() + loc * (cells.corners[-1] - cells.corners[0]).cpu().numpy()
        assert np.abs(back - y[:, None]).max() <= 1e-13
        assert np.array_equal(view.coordinates().cpu().numpy(), w.entities(0, 0, d, 4)["center"])
        if d == 2:
            poly = view.polygons().cpu().numpy()
            cc = cells.corners.cpu().numpy()
            assert poly.shape == (len(cells), 4, 2) and np.array_equal(poly[:, 0], cc[1].T) and np.array_equal(poly[:, 1], cc[0].T)
        else:
            with pytest.raises(ValueError):
                view.polygons()
        assert isinstance(cells.toGlobal(np.zeros(d)), torch.Tensor)
        w.close()
