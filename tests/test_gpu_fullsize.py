"""BASELINE.json configurations at FULL size, where no CPU oracle finishes in seconds: size-independent properties of the
domain (the reference's own invariants, test/checkindexset.hh, test/gridcheck.hh:844-925, test/checkintersectionit.hh,
checkcommunicate.hh) evaluated in bulk on the device. Small-size bit parity with the oracle is in test_gpu_parity.py /
test_gpu_golden.py; the arithmetic is size independent."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def graded(n, r=1.005):
    w = r ** np.arange(n)
    x = np.concatenate([[0.0], np.cumsum(w)])
    return x / x[-1]


def need(gib):
    import torch
    free, _ = torch.cuda.mem_get_info()
    if free < gib * 2 ** 30:
        pytest.skip(f"needs {gib} GiB of device memory")


def test_config4_tensor_384_periodic_boundary_segments_and_wrap():
    """config 4: TensorProductCoordinates graded 384^3, x periodic, boundary-segment indexing"""
    import torch
    import dune_grid_b200 as dg
    need(40)
    n = 384
    g = dg.YaspGrid(coords=[graded(n), graded(n), graded(n)], periodic=1, coordinates="tensorproduct")
    gv = g.leafGridView()
    assert gv.size(0) == (n + 2) * n * n and gv.size(0, dg.Interior_Partition) == n ** 3
    assert g.numBoundarySegments() == 4 * (n + 2) * n                       # y and z faces of the stored (n+2) x n x n box
    e = gv.elements(0, dg.All_Partition)
    assert torch.equal(e["index"].long(), torch.arange(gv.size(0), device="cuda"))                       # consecutive, iteration order
    interior = e["partitionType"] == dg.InteriorEntity
    assert int(interior.sum()) == n ** 3 and int((e["partitionType"] == dg.OverlapEntity).sum()) == 2 * n * n
    vol = e["volume"][interior].sum().item()
    assert abs(vol - 1.0) <= 1e-12                                           # cells tile the unit cube
    c = e["center"]
    assert float(c[0].min()) > 0.0 and float(c[0].max()) < 1.0              # periodic overlap cells are wrapped into the domain
    del e, c
    it = gv.intersections(dg.All_Partition, geometry=False)
    bsi = it["boundarySegmentIndex"][it["boundary"].bool()]
    assert bsi.numel() == g.numBoundarySegments()
    assert torch.equal(torch.sort(bsi.long()).values, torch.arange(g.numBoundarySegments(), device="cuda"))  # bijection (gridcheck.hh:844-925)
    assert int((it["boundarySegmentIndex"][~it["boundary"].bool()] != -1).sum()) == 0
    out, nb = it["outside"].long(), it["neighbor"].bool()
    N = gv.size(0)
    for k in range(6):                                                       # outside(outside(e)) == e through the opposite face
        src = torch.nonzero(nb[k]).squeeze(1)
        o = out[k][src]
        assert bool(nb[k ^ 1][o].all()) and torch.equal(out[k ^ 1][o], src)
    assert N == out.shape[1]
    del it, out, nb
    # periodic self-exchange: idempotent, and overlap cells equal their interior images
    m = dg.MultipleCodimMultipleGeomTypeMapper(gv, dg.mcmgElementLayout(3))
    f = torch.rand(N, dtype=torch.float64, device="cuda")
    gv.communicate(m, [f], dg.InteriorBorder_All_Interface, dg.ForwardCommunication, "set")
    f1 = f.clone()
    gv.communicate(m, [f], dg.InteriorBorder_All_Interface, dg.ForwardCommunication, "set")
    assert torch.equal(f, f1)
    f3 = f.view(n, n, n + 2)
    assert torch.equal(f3[:, :, 0], f3[:, :, n]) and torch.equal(f3[:, :, n + 1], f3[:, :, 1])


def test_config5_geometrygrid_256_volume_and_levels():
    """config 5: GeometryGrid over YaspGrid<3> 128^3 refined once (256^3 leaf), non-affine deformation that fixes the
    boundary: the deformed domain is the unit cube, so sum_q w_q |det J| over all cells is 1 on every level"""
    import dune_grid_b200 as dg
    need(30)
    g = dg.YaspGrid(dim=3, N=(128, 128, 128), upper=(1.0, 1.0, 1.0))
    g.globalRefine(1)
    gauss = 0.5 + np.array([-0.5, 0.5]) / np.sqrt(3.0)
    grids = np.meshgrid(gauss, gauss, gauss, indexing="ij")
    qp = np.stack([x.T.reshape(-1) for x in grids], axis=1)
    wts = np.full(8, 0.125)
    vols = []
    for lvl in (0, 1):
        gv = g.levelGridView(lvl)
        assert gv.size(0) == (128 << lvl) ** 3
        vols.append(gv.geometryGridVolume(1, [0.05], qp, wts, dg.All_Partition).item())
    # 2x2x2 Gauss integrates the trilinear-interpolated map's det J (degree 2 per variable) exactly: both levels give the
    # volume of the interpolated domain, whose boundary is the cube's -> 1 up to summation rounding (tolerance 1e-12)
    assert abs(vols[0] - 1.0) <= 1e-12 and abs(vols[1] - 1.0) <= 1e-12
    gv = g.levelGridView(1)
    geo = gv.geometryGrid(1, [0.05], qp, dg.All_Partition, corners=False)
    detJ = geo["integrationElement"]
    assert float(detJ.min()) > 0.0                                            # orientation preserving
    assert abs((detJ * 0.125).sum().item() - 1.0) <= 1e-12                    # materialised path == reduction path
    jt, jit = geo["jacobianTransposed"], geo["jacobianInverseTransposed"]
    # J^-T J^T = I at every point (checkjacobians.hh), sampled on the first Gauss point; tolerance 1e-12
    prod = (jit[0].permute(2, 0, 1) @ jt[0].permute(2, 1, 0).transpose(1, 2))
    import torch
    eye = torch.eye(3, dtype=torch.float64, device="cuda")
    assert float((prod - eye).abs().max()) <= 1e-12


def test_config3_rank_of_1024_cube_descriptor_and_sweep():
    """config 3: one rank (corner rank 0 and interior-most rank 7) of 1024^3 on 2x2x2: 513^3 stored cells, index set and
    partition types of the stored box, and an FV step on the shell/bulk split equals the unsplit launch bit for bit"""
    import os
    import torch
    import dune_grid_b200 as dg
    need(12)
    for rank in (0, 7):
        g = dg.YaspGrid(dim=3, N=(1024, 1024, 1024), upper=(1.0, 1.0, 1.0), rank=rank, nranks=8)
        gv = g.leafGridView()
        assert g.torusDims() == (2, 2, 2) and gv.size(0) == 513 ** 3 and gv.size(0, dg.Interior_Partition) == 512 ** 3
        e = gv.elements(0, dg.All_Partition, geometry=False)
        assert int((e["partitionType"] == dg.OverlapEntity).sum()) == 513 ** 3 - 512 ** 3
        del e

        class _Fake:
            import ctypes
            h = ctypes.c_void_p(1)
        g.comm = _Fake()
        res = []
        for split in ("", "1"):
            if split:
                os.environ["DGB_FV_SPLIT_LOCAL"] = "1"
            else:
                os.environ.pop("DGB_FV_SPLIT_LOCAL", None)
            fv = dg.FiniteVolume(g, 0, 0, [1.0, 1.0, 1.0, 0.0, 0.25, 0.25, 0.25, 0.125, 0.2], "separable")
            c = torch.rand(gv.size(0), dtype=torch.float64, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
            cn = torch.zeros_like(c)
            fv.bootstrap_local()
            fv.step_local(c, cn)
            res.append(cn.clone())
            del fv, c, cn
        os.environ.pop("DGB_FV_SPLIT_LOCAL", None)
        assert torch.equal(res[0], res[1])
        del res


def test_config1_recipes_256x256_against_the_oracle():
    """config 1: YaspGrid<2> 256x256, the loops of doc/recipes/recipe-integration.cc:90-126 (midpoint rule, 3x3 Gauss rule
    of order 5, boundary flux of f(x) = x) and the MCMG mapper sweeps of recipe-iterate-over-grid, as bulk sweeps. The
    per-entity factors are bit-equal to the oracle's; the sums are compared with tolerance 1e-12 relative (the device sum
    and the CPU loop add in different orders)."""
    import torch
    import dune_grid_b200 as dg
    from tests.common import best_oracle, make_cfg, oracle_world, product_grid
    cfg = make_cfg(dim=2, N=(256, 256))
    o = best_oracle()
    w = oracle_world(o, cfg, 1)
    g = product_grid(cfg, 0, 1)
    gv = g.leafGridView()
    assert gv.size(0) == 65536 and gv.size(1) == 131584 and gv.size(2) == 66049            # sizes of config 1
    # midpoint rule: sum_e u(center) * volume, u(x) = exp(|x|)
    e, r = gv.elements(0, dg.All_Partition), w.entities(0, 0, 0)
    center, vol = e["center"].cpu().numpy().T, e["volume"].cpu().numpy()
    assert np.array_equal(center, r["center"]) and np.array_equal(vol, r["volume"])
    ref1 = 0.0
    for c, v in zip(r["center"], r["volume"]):                                             # the recipe's loop, in its order
        ref1 += np.exp(np.sqrt(c @ c)) * v
    got1 = (torch.exp(torch.linalg.vector_norm(e["center"], dim=0)) * e["volume"]).sum().item()
    assert abs(got1 - ref1) <= 1e-12 * abs(ref1)
    # Gauss rule of order 5 (3 x 3 points): sum_e sum_q u(global(x_q)) * integrationElement * w_q
    x1, w1 = np.polynomial.legendre.leggauss(3)
    x1, w1 = 0.5 * (x1 + 1.0), 0.5 * w1
    qp = np.array([[a, b] for b in x1 for a in x1])
    wq = np.array([wa * wb for wb in w1 for wa in w1])
    geo, rg = gv.geometry(qp, dg.All_Partition), w.geometry_eval(0, 0, qp)
    glob = geo["global"].cpu().numpy()                                                     # (nq, dim, n)
    assert np.array_equal(np.transpose(glob, (2, 0, 1)), rg["global"])
    assert np.array_equal(geo["integrationElement"].cpu().numpy().T, rg["detJ"])
    ref2 = float(np.sum(np.exp(np.linalg.norm(rg["global"], axis=2)) * rg["detJ"] * wq[None, :]))
    wt = torch.from_numpy(wq).cuda()
    got2 = (torch.exp(torch.linalg.vector_norm(geo["global"], dim=1)) * geo["integrationElement"] * wt[:, None]).sum().item()
    assert abs(got2 - ref2) <= 1e-12 * abs(ref2)
    assert abs(got1 - got2) <= 1e-4 * abs(got2)                                            # the two rules agree to discretisation error
    # boundary flux: sum over faces without neighbour of f(center) . n * |F|, f(x) = x  (= dim * |domain| = 2)
    it, ri = gv.intersections(dg.All_Partition), w.intersections(0, 0)
    nb = it["neighbor"].cpu().numpy().T
    assert np.array_equal(nb, ri["neighbor"])
    fc = it["center"].cpu().numpy()                                                        # (faces, dim, n)
    assert np.array_equal(np.transpose(fc, (2, 0, 1)), ri["center"]) and np.array_equal(it["volume"].cpu().numpy().T, ri["volume"])
    ref3 = float(np.sum(np.where(ri["neighbor"] == 0, np.einsum("efd,efd->ef", ri["center"], ri["normal"]) * ri["volume"], 0.0)))
    got3 = 0.0
    for k in range(4):
        nrm = torch.tensor(gv.centerUnitOuterNormal(k), dtype=torch.float64, device="cuda")
        flux = (it["center"][k] * nrm[:, None]).sum(0) * it["volume"][k]
        got3 += flux[~it["neighbor"][k].bool()].sum().item()
    assert abs(got3 - ref3) <= 1e-12 * abs(ref3) and abs(got3 - 2.0) <= 1e-12
    # mapper sweeps: element / vertex / all-codim layouts, index and subIndex (bit-exact)
    for layout in ([1, 0, 0], [0, 0, 1], [1, 1, 1]):
        m = dg.MultipleCodimMultipleGeomTypeMapper(gv, layout)
        assert m.size == w.mapper_size(0, 0, layout)
        for codim in range(3):
            if layout[codim]:
                assert np.array_equal(m.index(codim).cpu().numpy().view(np.uint32), w.mapper_index(0, 0, layout, codim))
                assert np.array_equal(m.subIndex(codim).cpu().numpy().view(np.uint32).T, w.mapper_subindex(0, 0, layout, codim))
    w.close()
