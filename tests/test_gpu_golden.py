"""GPU: the CUDA path (through the C ABI) against the committed golden vectors generated from the reference's own code
(tests/golden/make_golden.py) - no oracle library is needed at run time. Bars as in DESIGN.md §2: integers and
axis-aligned geometry bit-exact, FV exact mode bit-exact, GeometryGrid and FV separable mode <= 1e-13 relative."""
import os

import numpy as np
import pytest

from tests.common import product_grid
from tests.golden_cases import CASES, comm_data, quad_points
from tests.test_gpu_comm_fv import emulated_exchange, run_fv

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def np_(t):
    return t.cpu().numpy()


@pytest.mark.parametrize("name", sorted(CASES))
def test_sweeps_against_golden(name):
    import dune_grid_b200 as dg
    cfg, P, layout, _ = CASES[name]
    gold = np.load(os.path.join(GOLDEN, name + ".npz"))
    d = cfg["dim"]
    qp = quad_points(d)
    for r in range(P):
        g = product_grid(cfg, r, P)
        assert tuple(gold["torus_dims"]) == g.torusDims() and int(gold["max_level"][0]) == g.maxLevel
        assert int(gold[f"r{r}/nbs"][0]) == g.numBoundarySegments()
        for lvl in range(cfg["refine"] + 1):
            gv = g.levelGridView(lvl)
            k = f"r{r}/l{lvl}"
            for codim in range(d + 1):
                for which in range(8):
                    assert np.array_equal(g.commList(lvl, codim, which), gold[f"{k}/c{codim}/list{which}"])
                for pit in range(5):
                    e = gv.elements(codim, pit, coord=True)
                    kk = f"{k}/c{codim}/p{pit}"
                    assert np.array_equal(np_(e["index"]).view(np.uint32), gold[kk + "/index"])
                    assert np.array_equal(np_(e["partitionType"]), gold[kk + "/ptype"])
                    assert np.array_equal(np_(e["coord"]).T, gold[kk + "/coord"])
                    for q in ("lower", "upper", "center"):
                        assert np.array_equal(np_(e[q]).T, gold[kk + "/" + q]), (kk, q)
                    assert np.array_equal(np_(e["volume"]), gold[kk + "/volume"])
            m = dg.MultipleCodimMultipleGeomTypeMapper(gv, layout)
            assert m.size == int(gold[f"{k}/mapper/size"][0])
            for codim in range(d + 1):
                if layout[codim]:
                    assert np.array_equal(np_(m.index(codim, 4)).view(np.uint32), gold[f"{k}/mapper/index{codim}"])
                    assert np.array_equal(np_(m.subIndex(codim, 4)).view(np.uint32).T, gold[f"{k}/mapper/subindex{codim}"])
            for cc in range(d + 1):
                lay = [0] * (d + 1)
                lay[cc] = 1
                mm = dg.MultipleCodimMultipleGeomTypeMapper(gv, lay)
                assert np.array_equal(np_(mm.subIndex(cc, 4)).view(np.uint32).T, gold[f"{k}/subindex{cc}"])
            for pit in (0, 4):
                it = gv.intersections(pit)
                kk = f"{k}/is/p{pit}"
                assert np.array_equal(np_(it["neighbor"]).T, gold[kk + "/neighbor"])
                assert np.array_equal(np_(it["boundary"]).T, gold[kk + "/boundary"])
                assert np.array_equal(np_(it["outside"]).T, gold[kk + "/outside"])
                assert np.array_equal(np_(it["boundarySegmentIndex"]).T, gold[kk + "/bsi"])
                assert np.array_equal(np.transpose(np_(it["center"]), (2, 0, 1)), gold[kk + "/center"])
                assert np.array_equal(np_(it["volume"]).T, gold[kk + "/volume"])
            ge = gv.geometry(qp, 4)
            assert np.array_equal(np.transpose(np_(ge["global"]), (2, 0, 1)), gold[f"{k}/geo/global"])
            assert np.array_equal(np.transpose(np_(ge["jacobianInverseTransposed"]), (3, 0, 1, 2)), gold[f"{k}/geo/jit"])
            assert np.array_equal(np_(ge["integrationElement"]).T, gold[f"{k}/geo/detJ"])
            gg = gv.geometryGrid(1, [0.05], qp, 4)
            for a, b in ((np.transpose(np_(gg["corners"]), (2, 0, 1)), gold[f"{k}/geogrid/corners"]),
                         (np.transpose(np_(gg["global"]), (2, 0, 1)), gold[f"{k}/geogrid/global"]),
                         (np.transpose(np_(gg["jacobianTransposed"]), (3, 0, 1, 2)), gold[f"{k}/geogrid/jt"]),
                         (np.transpose(np_(gg["jacobianInverseTransposed"]), (3, 0, 1, 2)), gold[f"{k}/geogrid/jit"]),
                         (np_(gg["integrationElement"]).T, gold[f"{k}/geogrid/detJ"])):
                assert np.abs(a - b).max(initial=0.0) <= 1e-13 * max(1.0, np.abs(b).max(initial=0.0))     # tolerance: 1e-13 relative


@pytest.mark.parametrize("name", sorted(CASES))
def test_communicate_against_golden(name):
    import torch
    import dune_grid_b200 as dg
    cfg, P, layout, _ = CASES[name]
    gold = np.load(os.path.join(GOLDEN, name + ".npz"))
    lvl = cfg["refine"]
    grids = [product_grid(cfg, r, P) for r in range(P)]
    views = [g.levelGridView(lvl) for g in grids]
    mappers = [dg.MultipleCodimMultipleGeomTypeMapper(v, layout) for v in views]
    for iftype in range(5):
        for direction in (0, 1):
            for op in (0, 1):
                dev = [[torch.from_numpy(a).cuda() for a in comm_data(mappers[r].size, 1000 + 10 * r)] for r in range(P)]
                sent = emulated_exchange(grids, views, mappers, dev, [1, 3], iftype, direction, "add" if op else "set")
                key = f"comm/i{iftype}d{direction}o{op}"
                assert sent == int(gold[key + "/sent"][0])
                for r in range(P):
                    for a in range(2):
                        assert np.array_equal(np_(dev[r][a]), gold[f"{key}/r{r}a{a}"]), (key, r, a)


@pytest.mark.parametrize("name", sorted(CASES))
def test_communicate_variable_against_golden(name):
    """the variable-size data handle (fixedSize() == false) against the scatter-call sequences the reference itself produced"""
    import torch
    import dune_grid_b200 as dg
    from tests.golden_cases import var_data, var_layout
    from tests.test_gpu_comm_fv import emulated_variable_exchange
    cfg, P, layout, _ = CASES[name]
    gold = np.load(os.path.join(GOLDEN, name + ".npz"))
    lvl = cfg["refine"]
    grids = [product_grid(cfg, r, P) for r in range(P)]
    views = [g.levelGridView(lvl) for g in grids]
    mappers = [dg.MultipleCodimMultipleGeomTypeMapper(v, var_layout(layout)) for v in views]
    for iftype in range(5):
        for direction in (0, 1):
            vd = [var_data(mappers[r].size, 2000 + 10 * r) for r in range(P)]
            got = emulated_variable_exchange(views, mappers, [torch.from_numpy(v[0]).cuda() for v in vd], [torch.from_numpy(v[1]).cuda() for v in vd],
                                             iftype, direction)
            for r in range(P):
                for q, a in zip(("index", "n", "items"), got[r]):
                    assert np.array_equal(a, gold[f"commvar/i{iftype}d{direction}/r{r}/{q}"]), (iftype, direction, r, q)


@pytest.mark.parametrize("name", sorted(n for n in CASES if CASES[n][3] is not None))
def test_fv_against_golden(name):
    cfg, P, _, (problem, params) = CASES[name]
    gold = np.load(os.path.join(GOLDEN, name + ".npz"))
    got, dts, _ = run_fv(cfg, P, problem, params, "exact", 3)
    assert np.array_equal(np.array(dts), gold["fv/dt"])
    for r in range(P):
        assert np.array_equal(got[r], gold[f"fv/step3/r{r}"])
    got, dts, _ = run_fv(cfg, P, problem, params, "separable", 3)
    scale = max(np.abs(gold[f"fv/step3/r{r}"]).max() for r in range(P))
    assert np.abs(np.array(dts) - gold["fv/dt"]).max() <= 1e-13 * np.abs(gold["fv/dt"]).max()      # tolerance: 1e-13 relative
    for r in range(P):
        assert np.abs(got[r] - gold[f"fv/step3/r{r}"]).max() <= 1e-13 * scale
