"""GPU parity tests proper: every bulk sweep of libdgb (called through the C ABI) against the CPU oracle on the
same grids. Integers (indices, mapper numbering, partition types, neighbour tables, boundary-segment ids) must be
bit-exact; axis-aligned geometry is bit-exact too (same expressions, FMA contraction off on both sides);
GeometryGrid's J^-T and the separable FV mode are checked to 1e-13 relative as north_star allows.
The reference's own invariants (test/checkindexset.hh, checkintersectionit.hh, gridcheck.hh:844-925,
checkpartition.hh) are re-expressed as bulk checks in test_gpu_invariants.py."""
import math

import numpy as np
import pytest

from tests.common import best_oracle, make_cfg, oracle_world, product_grid

pytestmark = pytest.mark.gpu

ALL_PARTITIONS = [0, 1, 2, 3, 4, 5]

CASES = [
    (make_cfg(dim=2, N=(8, 4)), 1), (make_cfg(dim=2, N=(8, 4)), 2),
    (make_cfg(dim=3, N=(8, 4, 4)), 2),
    (make_cfg(dim=2, N=(8, 4), mode=1, lower=(-1.0, -0.5), upper=(1.0, 0.5), refine=1, keep=0), 2),
    (make_cfg(dim=3, N=(8, 4, 4), mode=2, tensor=[np.array([-1, -.5, -.25, -.125, 0, .125, .25, .5, 1.0]),
                                                  np.array([-1, -.5, 0, .5, 1.0]), np.array([-1, -.5, 0, .5, 1.0])]), 2),
    (make_cfg(dim=3, N=(12, 10, 9), periodic=5, overlap=2, upper=(1.0, 2.0, 3.0)), 4),
    (make_cfg(dim=2, N=(13, 7), periodic=3, refine=1), 1),
    (make_cfg(dim=3, N=(6, 5, 4), periodic=7, mode=2), 1),
    (make_cfg(dim=3, N=(10, 6, 7), periodic=2, mode=2, refine=1, keep=0), 4),
    (make_cfg(dim=3, N=(9, 8, 8), overlap=0), 4),
    (make_cfg(dim=3, N=(10, 10, 10), upper=(0.3, 0.7, 1.9)), 8),
]
IDS = [f"case{i}" for i in range(len(CASES))]


def np_(t):
    return t.cpu().numpy()


@pytest.fixture(scope="module")
def oracle():
    return best_oracle()


@pytest.mark.parametrize("cfg,P", CASES, ids=IDS)
def test_entities(oracle, cfg, P):
    w = oracle_world(oracle, cfg, P)
    d = cfg["dim"]
    for r in range(P):
        g = product_grid(cfg, r, P)
        for lvl in range(cfg["refine"] + 1):
            gv = g.levelGridView(lvl)
            for codim in range(d + 1):
                for pit in ALL_PARTITIONS:
                    ref = w.entities(r, lvl, codim, pit)
                    got = gv.elements(codim, pit, coord=True)
                    n = len(ref["index"])
                    assert gv.size(codim, pit) == n
                    assert np.array_equal(np_(got["index"]).view(np.uint32), ref["index"]), (r, lvl, codim, pit)
                    assert np.array_equal(np_(got["partitionType"]), ref["ptype"])
                    assert np.array_equal(np_(got["coord"]).T, ref["coord"])
                    for k in ("lower", "upper", "center"):
                        assert np.array_equal(np_(got[k]).T, ref[k]), (r, lvl, codim, pit, k)
                    assert np.array_equal(np_(got["volume"]), ref["volume"])
    w.close()


@pytest.mark.parametrize("cfg,P", CASES, ids=IDS)
def test_mapper_and_subindices(oracle, cfg, P):
    import dune_grid_b200 as dg
    w = oracle_world(oracle, cfg, P)
    d = cfg["dim"]
    layouts = [dg.mcmgElementLayout(d), dg.mcmgVertexLayout(d), [1] * (d + 1), [2, 0, 3, 1][:d + 1]]
    for r in range(P):
        g = product_grid(cfg, r, P)
        lvl = cfg["refine"]
        gv = g.levelGridView(lvl)
        for layout in layouts:
            m = dg.MultipleCodimMultipleGeomTypeMapper(gv, layout)
            assert m.size == w.mapper_size(r, lvl, layout)
            for codim in range(d + 1):
                if layout[codim] == 0:
                    with pytest.raises(dg.capi.DgbError):
                        m.index(codim)
                    continue
                for pit in (0, 1, 4):
                    assert np.array_equal(np_(m.index(codim, pit)).view(np.uint32), w.mapper_index(r, lvl, layout, codim, pit)), (layout, codim, pit)
                for pit in (0, 4):
                    assert np.array_equal(np_(m.subIndex(codim, pit)).view(np.uint32).T, w.mapper_subindex(r, lvl, layout, codim, pit))
        # plain index-set subIndex (layout all ones has offsets; use per-codim layouts with block 1)
        for cc in range(d + 1):
            lay = [0] * (d + 1)
            lay[cc] = 1
            m = dg.MultipleCodimMultipleGeomTypeMapper(gv, lay)
            assert np.array_equal(np_(m.subIndex(cc, 4)).view(np.uint32).T, w.subindices(r, lvl, cc, 4))
    w.close()


VEC_CASES = [
    (make_cfg(dim=3, N=(131, 5, 3)), 1), (make_cfg(dim=3, N=(5, 70, 9), periodic=7, overlap=2), 1), (make_cfg(dim=3, N=(3, 70, 9), periodic=7), 1), (make_cfg(dim=3, N=(600, 3, 2)), 2),
    (make_cfg(dim=2, N=(1030, 3)), 1), (make_cfg(dim=3, N=(33, 17, 6), periodic=2), 4), (make_cfg(dim=3, N=(64, 8, 8)), 1),
]


@pytest.mark.parametrize("cfg,P", VEC_CASES, ids=[f"vec{i}" for i in range(len(VEC_CASES))])
def test_mapper_index_vector_kernel(oracle, cfg, P):
    """k_index_vec (16-byte stores, one launch for all streams; taken for 16-byte aligned outputs) against the oracle AND against
    the column / linear kernels (taken for an output that is only 4-byte aligned), over every codim, partition range and stream
    alignment: rows longer and shorter than a CTA's 512 entities, rows that are no multiple of 4, one-wide boxes, streams that
    start at any offset modulo 4, guard words either side of the output."""
    import torch
    import dune_grid_b200 as dg
    w = oracle_world(oracle, cfg, P)
    d = cfg["dim"]
    for r in range(P):
        gv = product_grid(cfg, r, P).leafGridView()
        m = dg.MultipleCodimMultipleGeomTypeMapper(gv, [2, 1, 3, 1][:d + 1])
        for codim in range(d + 1):
            for pit in range(5):
                for sub in (False, True):
                    if sub and pit not in (0, 4):
                        continue
                    n = gv.size(0 if sub else codim, pit)
                    ns = (math.comb(d, codim) << codim) if sub else 1
                    ref = (w.mapper_subindex(r, 0, [2, 1, 3, 1][:d + 1], codim, pit).T if sub else
                           w.mapper_index(r, 0, [2, 1, 3, 1][:d + 1], codim, pit)).reshape(-1)
                    for misalign in (0, 1):
                        buf = torch.full((ns * n + 16,), -7, dtype=torch.int32, device="cuda")
                        out = buf[4 + misalign:4 + misalign + ns * n]
                        if sub:
                            m.subIndex(codim, pit, out=out)
                        else:
                            m.index(codim, pit, out=out)
                        got = np_(buf)
                        assert np.array_equal(got[4 + misalign:4 + misalign + ns * n].view(np.uint32), ref), (r, codim, pit, sub, misalign)
                        assert (got[:4 + misalign] == -7).all() and (got[4 + misalign + ns * n:] == -7).all(), "wrote outside the output"
    w.close()


@pytest.mark.parametrize("cfg,P", CASES, ids=IDS)
def test_intersections(oracle, cfg, P):
    w = oracle_world(oracle, cfg, P)
    d = cfg["dim"]
    for r in range(P):
        g = product_grid(cfg, r, P)
        for lvl in range(cfg["refine"] + 1):
            gv = g.levelGridView(lvl)
            for pit in (0, 2, 4):
                ref = w.intersections(r, lvl, pit)
                got = gv.intersections(pit)
                assert np.array_equal(np_(got["neighbor"]).T, ref["neighbor"]), (r, lvl, pit)
                assert np.array_equal(np_(got["boundary"]).T, ref["boundary"])
                assert np.array_equal(np_(got["outside"]).T, ref["outside"])
                assert np.array_equal(np_(got["boundarySegmentIndex"]).T, ref["bsi"])
                assert np.array_equal(np.transpose(np_(got["center"]), (2, 0, 1)), ref["center"])
                assert np.array_equal(np_(got["volume"]).T, ref["volume"])
                for k in range(2 * d):
                    assert np.array_equal(ref["normal"][:, k, :], np.broadcast_to(gv.centerUnitOuterNormal(k), ref["normal"][:, k, :].shape))
    w.close()


# more box shapes for the element / intersection sweeps: rows of 131 / 132 / 38 / 1030 / 33 / 3 entities (column and linear walks)
QUAD_CASES = [
    (make_cfg(dim=3, N=(131, 4, 4)), 1), (make_cfg(dim=3, N=(130, 6, 4), periodic=7), 1), (make_cfg(dim=3, N=(64, 8, 8)), 2),
    (make_cfg(dim=2, N=(1030, 4)), 1), (make_cfg(dim=3, N=(36, 8, 4), periodic=1, mode=2), 1), (make_cfg(dim=3, N=(33, 16, 8), periodic=2, refine=1), 4),
    (make_cfg(dim=3, N=(3, 4, 8), periodic=7), 1),
]
QUAD_IDS = [f"quad{i}" for i in range(len(QUAD_CASES))]


@pytest.mark.parametrize("cfg,P", QUAD_CASES, ids=QUAD_IDS)
def test_entities_row_shapes(oracle, cfg, P):
    test_entities(oracle, cfg, P)


@pytest.mark.parametrize("cfg,P", QUAD_CASES, ids=QUAD_IDS)
def test_intersections_row_shapes(oracle, cfg, P):
    test_intersections(oracle, cfg, P)


def gauss_points(dim, n1d):
    x, wts = np.polynomial.legendre.leggauss(n1d)
    x = 0.5 * (x + 1.0)
    wts = 0.5 * wts
    grids = np.meshgrid(*([x] * dim), indexing="ij")
    pts = np.stack([gq.T.reshape(-1) for gq in grids], axis=1)       # axis 0 fastest
    wg = np.meshgrid(*([wts] * dim), indexing="ij")
    ww = np.prod(np.stack([gq.T.reshape(-1) for gq in wg], axis=1), axis=1)
    return pts, ww


@pytest.mark.parametrize("cfg,P", CASES[:8], ids=IDS[:8])
def test_geometry_at_quadrature_points(oracle, cfg, P):
    w = oracle_world(oracle, cfg, P)
    d = cfg["dim"]
    qp, _ = gauss_points(d, 2)
    for r in range(P):
        g = product_grid(cfg, r, P)
        lvl = cfg["refine"]
        gv = g.levelGridView(lvl)
        ref = w.geometry_eval(r, lvl, qp, 4)
        got = gv.geometry(qp, 4)
        assert np.array_equal(np.transpose(np_(got["global"]), (2, 0, 1)), ref["global"])
        assert np.array_equal(np.transpose(np_(got["jacobianInverseTransposed"]), (3, 0, 1, 2)), ref["jit"])
        assert np.array_equal(np_(got["integrationElement"]).T, ref["detJ"])
    w.close()


@pytest.mark.parametrize("cfg,P", [CASES[0], CASES[2], CASES[4], CASES[6], CASES[8]], ids=["c0", "c2", "c4", "c6", "c8"])
def test_geometrygrid(oracle, cfg, P):
    """GeometryGrid over the Yasp host view (config 5 of BASELINE.json at small size). global / jacobianTransposed /
    integrationElement follow the same operation order as the oracle's MultiLinearGeometry restatement; the analytic
    deformation uses sin(), whose device implementation differs from glibc's by <= 2 ulp, so everything is compared
    to 1e-13 relative (of the largest entry of the quantity), as north_star states."""
    w = oracle_world(oracle, cfg, P)
    d = cfg["dim"]
    qp, wts = gauss_points(d, 2)
    for deformation, params in ((0, [0.0]), (1, [0.05])):
        for r in range(P):
            g = product_grid(cfg, r, P)
            for lvl in range(cfg["refine"] + 1):
                gv = g.levelGridView(lvl)
                ref = w.geogrid_eval(r, lvl, deformation, params, qp, 0)
                got = gv.geometryGrid(deformation, params, qp, 0)
                pairs = [(np.transpose(np_(got["corners"]), (2, 0, 1)), ref["corners"]),
                         (np.transpose(np_(got["global"]), (2, 0, 1)), ref["global"]),
                         (np.transpose(np_(got["jacobianTransposed"]), (3, 0, 1, 2)), ref["jt"]),
                         (np.transpose(np_(got["jacobianInverseTransposed"]), (3, 0, 1, 2)), ref["jit"]),
                         (np_(got["integrationElement"]).T, ref["detJ"])]
                for a, b in pairs:
                    scale = np.abs(b).max() if b.size else 1.0
                    assert np.abs(a - b).max(initial=0.0) <= 1e-13 * scale
                if deformation == 0:
                    assert np.array_equal(pairs[0][0], pairs[0][1]) and np.array_equal(pairs[1][0], pairs[1][1])
                    assert np.array_equal(pairs[2][0], pairs[2][1]) and np.array_equal(pairs[4][0], pairs[4][1])
                vol = float(np_(gv.geometryGridVolume(deformation, params, qp, wts, 0))[0])
                refvol = float((ref["detJ"] * wts[None, :]).sum())
                assert abs(vol - refvol) <= 1e-12 * max(1.0, abs(refvol))
    w.close()


GEOFACE_CASES = [
    make_cfg(dim=3, N=(7, 6, 5)),
    make_cfg(dim=3, N=(6, 5, 4), mode=2, periodic=1),
    make_cfg(dim=2, N=(9, 8), periodic=2),
    make_cfg(dim=3, N=(4, 4, 4), refine=1),
]


@pytest.mark.parametrize("cfg", GEOFACE_CASES, ids=[f"geoface{i}" for i in range(len(GEOFACE_CASES))])
def test_geometrygrid_intersections(cfg):
    """GeoGrid::Intersection (geometrygrid/intersection.hh:103-172, cornerstorage.hh:122-165) as a bulk sweep against the oracle
    - the reference's own intersection / corner-vector headers compiled over the restated multilinear map: face corners,
    face centre and volume bit for bit for the identity deformation (same expressions), within 1e-13 relative with the sine
    deformation (device sin() differs from glibc's by <= 2 ulp), normals within 1e-13 relative (the product inverts J by the
    adjugate, dune-geometry by a Cholesky right inverse) - plus the reference's own invariants: unit normals have norm 1 and
    are parallel to J^-T n_ref (checkintersectionit.hh:329-400), centerUnitOuterNormal == unitOuterNormal(face centre), and the
    integration outer normals of a cell, integrated with the face's Gauss rule, sum to zero (checkintersectionit.hh:643-654)."""
    import torch
    import dune_grid_b200 as dg
    d, lvl = cfg["dim"], cfg["refine"]
    o = best_oracle()
    w = oracle_world(o, cfg, 1)
    g = product_grid(cfg, 0, 1)
    gv = g.levelGridView(lvl)
    gauss = 0.5 + np.array([-0.5, 0.5]) / np.sqrt(3.0)
    qp = np.array([[a] for a in gauss]) if d == 2 else np.array([[a, b] for b in gauss for a in gauss])
    wq = np.full(len(qp), 1.0 / len(qp))
    for deformation, params in ((1, [0.05]), (0, [0.0])):
        got = gv.geometryGridIntersections(deformation, params, qp, dg.All_Partition)
        ref = w.geogrid_intersections(0, lvl, deformation, params, qp)
        t = lambda a, axes: np.transpose(a.cpu().numpy(), axes)                                  # noqa: E731
        for name, axes in (("corners", (3, 0, 1, 2)), ("center", (2, 0, 1)), ("volume", (1, 0))):
            a, b = t(got[name], axes), ref[name]
            if deformation == 0:
                assert np.array_equal(a, b), name
            else:
                assert np.abs(a - b).max() <= 1e-13 * np.abs(b).max(), name                      # tolerance: 1e-13 relative
        for name in ("integrationOuterNormal", "outerNormal", "unitOuterNormal"):
            a, b = t(got[name], (3, 0, 1, 2)), ref[name]
            assert np.abs(a - b).max() <= 1e-13 * max(1.0, np.abs(b).max()), name                # tolerance: 1e-13 relative
        a, b = t(got["centerUnitOuterNormal"], (2, 0, 1)), ref["centerUnitOuterNormal"]
        assert np.abs(a - b).max() <= 1e-13
        # invariants of the reference's intersection check
        uon = got["unitOuterNormal"]                                                             # (2d, nq, d, n)
        assert float((torch.linalg.vector_norm(uon, dim=2) - 1.0).abs().max()) <= 1e-14
        on = got["outerNormal"]
        cross = (on / torch.linalg.vector_norm(on, dim=2, keepdim=True) - uon).abs().max()
        assert float(cross) <= 1e-14
        # sum_faces int_F n dS = 0 for every cell: the Gauss rule integrates the (bi)quadratic cofactor field exactly in 2-D and
        # 3-D for multilinear cells; tolerance 1e-13 x the face areas
        ion = got["integrationOuterNormal"]
        wt = torch.from_numpy(wq).cuda()
        total = (ion * wt[None, :, None, None]).sum(dim=(0, 1))                                 # (d, n)
        area = got["volume"].sum(0)
        assert float((total.abs().max(dim=0).values / area).max()) <= 1e-13
        # the face centre among the points: centerUnitOuterNormal is taken from that point's evaluation - identical values
        qc = np.vstack([qp[:1], np.full((1, d - 1), 0.5), qp[1:]])
        got_c = gv.geometryGridIntersections(deformation, params, qc, dg.All_Partition)
        assert torch.equal(got_c["centerUnitOuterNormal"], got["centerUnitOuterNormal"])
        assert torch.equal(got_c["unitOuterNormal"][:, 1], got["centerUnitOuterNormal"])
        assert torch.equal(got_c["unitOuterNormal"][:, 2:], got["unitOuterNormal"][:, 1:]) and torch.equal(got_c["corners"], got["corners"])
        if deformation == 0:                                                                     # identity: the host's own normals
            for k in range(2 * d):
                e = np.zeros(d); e[k // 2] = 1.0 if k % 2 else -1.0
                assert np.abs(got["centerUnitOuterNormal"][k].cpu().numpy() - e[:, None]).max() <= 1e-15
    w.close()


@pytest.mark.parametrize("cfg", [make_cfg(dim=3, N=(6, 5, 4), mode=2, periodic=1), make_cfg(dim=2, N=(9, 8), periodic=2, upper=(2.0, 0.5))],
                         ids=["tensor3d", "periodic2d"])
def test_yasp_intersection_normals_and_local_geometries(cfg):
    """YaspIntersection::unitOuterNormal / integrationOuterNormal (yaspgridintersection.hh:163-190) for all faces, bit for bit
    against the oracle's centerUnitOuterNormal() and geometry().volume(); geometryInInside / geometryInOutside (:195-228) are the
    unit faces of the reference cube"""
    import dune_grid_b200 as dg
    d = cfg["dim"]
    o = best_oracle()
    w = oracle_world(o, cfg, 1)
    g = product_grid(cfg, 0, 1)
    gv = g.leafGridView()
    u, wn = gv.intersectionNormals(dg.All_Partition)
    ri = w.intersections(0, 0, dg.All_Partition)
    assert np.array_equal(np.transpose(np_(u), (2, 0, 1)), ri["normal"])
    assert np.array_equal(np.transpose(np_(wn), (2, 0, 1)), ri["volume"][:, :, None] * ri["normal"])
    for k in range(2 * d):
        lo, hi = gv.geometryInInside(k)
        e_lo, e_hi = np.zeros(d), np.ones(d)
        e_lo[k // 2] = e_hi[k // 2] = float(k % 2)
        assert np.array_equal(lo, e_lo) and np.array_equal(hi, e_hi)
        lo, hi = gv.geometryInOutside(k)
        e_lo[k // 2] = e_hi[k // 2] = float(1 - k % 2)
        assert np.array_equal(lo, e_lo) and np.array_equal(hi, e_hi)
    w.close()


@pytest.mark.parametrize("cfg", [make_cfg(dim=3, N=(6, 5, 4), mode=2, periodic=1), make_cfg(dim=2, N=(9, 8), periodic=2, upper=(2.0, 0.5)),
                                 make_cfg(dim=3, N=(4, 4, 4), refine=1, mode=1, lower=(-1.0, 0.0, 0.5), upper=(1.0, 1.0, 2.0))],
                         ids=["tensor3d", "periodic2d", "offset3d_refined"])
def test_geometry_of_all_codimensions_at_local_points(cfg):
    """Geometry::global / jacobianTransposed / jacobianInverseTransposed / integrationElement of the entities of EVERY codimension
    (faces and edges at quadrature points, vertices) bit for bit against the oracle's e.geometry()"""
    d, lvl = cfg["dim"], cfg["refine"]
    o = best_oracle()
    w = oracle_world(o, cfg, 1)
    g = product_grid(cfg, 0, 1)
    gv = g.levelGridView(lvl)
    for codim in range(d + 1):
        m = d - codim
        qp = np.array([[0.3, 0.7, 0.1][:max(1, m)], [0.5] * max(1, m), [1.0, 0.0, 0.25][:max(1, m)]])
        for pit in (4, 0):
            ref = w.geometry_eval_codim(0, lvl, codim, qp, pit)
            got = gv.geometryCodim(codim, qp, pit)
            assert np.array_equal(np.transpose(np_(got["global"]), (2, 0, 1)), ref["global"]), (codim, pit)
            assert np.array_equal(np.transpose(np_(got["jacobianTransposed"]), (3, 0, 1, 2)), ref["jt"]), (codim, pit)
            assert np.array_equal(np.transpose(np_(got["jacobianInverseTransposed"]), (3, 0, 1, 2)), ref["jit"]), (codim, pit)
            assert np.array_equal(np_(got["integrationElement"]).T, ref["detJ"]), (codim, pit)
    w.close()
