"""VTK output of the bulk view (dune/grid/io/file/vtk/vtkwriter.hh restated in csrc/dgb_vtk.cu). The
upstream writer cannot be built here, so these tests check the format rules cited in dgb_vtk.cu and that the files describe
the grid and the fields: XML structure, counts, first-visit vertex numbering, VTK corner order, Float32 formatting, vector
padding, parallel piece / header names. Grid-only output runs on CPU; fields live on the device (GPU tests)."""
import os
import xml.etree.ElementTree as ET

import numpy as np
import pytest

import dune_grid_b200 as dg
from tests.common import make_cfg, product_grid


def arrays(piece):
    out = {}
    for da in piece.iter("DataArray"):
        txt = da.text.split()
        kind = da.get("type")
        a = np.array(txt, dtype=np.float64 if kind.startswith("Float") else np.int64)
        nc = int(da.get("NumberOfComponents"))
        out[da.get("Name")] = (a.reshape(-1, nc) if nc > 1 else a, kind, da.text)
    return out


def check_grid(path, gv, dim):
    root = ET.parse(path).getroot()
    assert root.tag == "VTKFile" and root.get("type") == "UnstructuredGrid" and root.get("byte_order") == "LittleEndian"
    piece = root.find("UnstructuredGrid/Piece")
    ncells, npoints = int(piece.get("NumberOfCells")), int(piece.get("NumberOfPoints"))
    assert ncells == gv.size(0, dg.Interior_Partition)
    a = arrays(piece)
    pts, conn, off, types = a["Coordinates"][0], a["connectivity"][0], a["offsets"][0], a["types"][0]
    nc = 1 << dim
    assert pts.shape == (npoints, 3) and conn.size == ncells * nc and off.size == ncells and types.size == ncells
    assert np.array_equal(off, nc * np.arange(1, ncells + 1)) and np.all(types == (9 if dim == 2 else 12))
    # vertices are numbered in the order the walk over cells and corners meets them: the running maximum grows by <= 1
    ren = [0, 1, 3, 2] if dim == 2 else [0, 1, 3, 2, 4, 5, 7, 6]
    dune_order = conn.reshape(ncells, nc)[:, np.argsort(ren)]        # connectivity[k] = corner ren[k]
    flat = dune_order.reshape(-1)
    runmax = np.maximum.accumulate(flat)
    assert flat[0] == 0 and np.all(np.diff(runmax) <= 1) and runmax[-1] == npoints - 1
    # every cell is an axis-aligned box whose Dune corner k sits at bit pattern k; coordinates are the view's (as Float32)
    e = gv.elements(0, dg.Interior_Partition) if False else None
    cell_pts = pts[dune_order]                                        # (ncells, nc, 3)
    for k in range(nc):
        for i in range(dim):
            lo, hi = cell_pts[:, 0, i], cell_pts[:, nc - 1, i]
            assert np.all(hi > lo) and np.array_equal(cell_pts[:, k, i], hi if (k >> i) & 1 else lo)
    if dim == 2:
        assert np.all(pts[:, 2] == 0)
    # 12 values per line, two-space indentation (dataarraywriter.hh:124,162-168)
    lines = [l for l in a["connectivity"][2].split("\n") if l.strip()]
    assert all(len(l.split()) == 12 for l in lines[:-1]) and lines[0].startswith("          ") and not lines[0].startswith("           ")
    return piece, a, dune_order, pts


@pytest.mark.parametrize("cfg", [make_cfg(dim=2, N=(5, 3), upper=(1.0, 0.3)), make_cfg(dim=3, N=(4, 3, 2), mode=2),
                                 make_cfg(dim=3, N=(4, 4, 3), periodic=5, mode=1, lower=(-1.0, 0.0, 2.0), upper=(1.0, 1.0, 2.5))])
def test_grid_only_file(cfg, tmp_path):
    g = product_grid(cfg, 0, 1)
    gv = g.leafGridView()
    f = gv.writeVTK("mesh", path=str(tmp_path))
    assert f == str(tmp_path / "mesh.vtu")
    piece, a, order, pts = check_grid(f, gv, cfg["dim"])
    assert piece.find("CellData") is None and piece.find("PointData") is None      # vtkwriter.hh:1342,1356
    # the points are the vertex coordinates of the view (Float32, 6 significant digits)
    for i in range(cfg["dim"]):
        first, coords = gv.coordinates(i)
        lo = np.float32(coords[-first] if cfg["periodic"] >> i & 1 else coords[0])
        assert np.isclose(pts[:, i].min(), lo, rtol=1e-5, atol=1e-7)


def test_parallel_names(tmp_path):
    cfg = make_cfg(dim=2, N=(8, 4))
    files = [product_grid(cfg, r, 2).leafGridView().writeVTK("run", path=str(tmp_path)) for r in range(2)]
    assert files[0] == files[1] == str(tmp_path / "s0002-run.pvtu")
    assert sorted(os.listdir(tmp_path)) == ["s0002-p0000-run.vtu", "s0002-p0001-run.vtu", "s0002-run.pvtu"]
    root = ET.parse(files[0]).getroot()
    assert root.get("type") == "PUnstructuredGrid"
    assert [p.get("Source") for p in root.iter("Piece")] == ["s0002-p0000-run.vtu", "s0002-p0001-run.vtu"]
    total = 0
    for r in range(2):
        gv = product_grid(cfg, r, 2).leafGridView()
        piece, *_ = check_grid(str(tmp_path / f"s0002-p{r:04d}-run.vtu"), gv, 2)
        total += int(piece.get("NumberOfCells"))
    assert total == 32                                               # interior cells only: no cell is written twice


@pytest.mark.gpu
def test_fields_from_the_device(tmp_path):
    import torch
    cfg = make_cfg(dim=3, N=(6, 5, 4), periodic=1, upper=(1.0, 2.0, 3.0))
    g = product_grid(cfg, 0, 1)
    gv = g.leafGridView()
    rng = np.random.default_rng(11)
    c = torch.from_numpy(rng.standard_normal(gv.size(0))).cuda()
    vel = torch.from_numpy(rng.standard_normal((gv.size(0), 2))).cuda()
    p = torch.from_numpy(rng.standard_normal(gv.size(3))).cuda()
    f = gv.writeVTK("fields", celldata={"c": c, "velocity": vel}, pointdata={"p": p}, path=str(tmp_path))
    piece, a, order, pts = check_grid(f, gv, 3)
    assert [s.tag for s in piece] == ["PointData", "CellData", "Points", "Cells"]              # writeAllData, :1205-1217
    assert piece.find("CellData").get("Scalars") == "c" and piece.find("CellData").get("Vectors") == "velocity"
    assert piece.find("PointData").get("Scalars") == "p"
    idx = gv.elements(0, dg.Interior_Partition, geometry=False)["index"].cpu().numpy().astype(np.int64)
    got_c, kind, _ = a["c"]
    assert kind == "Float32" and np.allclose(got_c, c.cpu().numpy()[idx].astype(np.float32), rtol=1e-5, atol=0)  # 6 significant digits in the text
    got_v = a["velocity"][0]
    assert got_v.shape == (idx.size, 3) and np.all(got_v[:, 2] == 0)                            # padded to 3 components
    assert np.allclose(got_v[:, :2], vel.cpu().numpy()[idx].astype(np.float32), rtol=1e-5)
    # point data follow the first-visit vertex numbering: value of the vertex each point is
    sub = dg.MultipleCodimMultipleGeomTypeMapper(gv, dg.mcmgVertexLayout(3)).subIndex(3, dg.Interior_Partition).cpu().numpy().astype(np.int64)
    vert_of_point = np.zeros(pts.shape[0], dtype=np.int64)
    vert_of_point[order.reshape(-1)] = sub.T.reshape(-1)              # sub: (8, cells) -> per cell Dune corner order
    assert np.allclose(a["p"][0], p.cpu().numpy()[vert_of_point].astype(np.float32), rtol=1e-5)
    # Float64 output keeps 15 digits
    f64 = gv.writeVTK("fields64", celldata={"c": c}, path=str(tmp_path), float64=True)
    a64 = arrays(ET.parse(f64).getroot().find("UnstructuredGrid/Piece"))
    assert a64["c"][1] == "Float64" and np.allclose(a64["c"][0], c.cpu().numpy()[idx], rtol=1e-14)


# ---- the four output types against the reference's own writer classes ----------------------------------------------------------------
# oracle/_ref compiles dune/grid/io/file/vtk/{vtuwriter,dataarraywriter,streams,b64enc}.hh unmodified (orc_vtu_piece): the product's
# file must be byte-identical to what those classes write for the same arrays. The arrays are read back from the product's ASCII
# file; all test values are short dyadic numbers, so the text holds them exactly.
OUTPUT_TYPES = ["ascii", "base64", "appendedraw", "appendedbase64"]
PREC = {"Float32": 0, "Float64": 1, "Int32": 2, "UInt8": 3}
SECTION = {"PointData": 0, "CellData": 1, "Points": 2, "Cells": 3}


def piece_arrays(path):
    piece = ET.parse(path).getroot().find("UnstructuredGrid/Piece")
    out, attrs = [], {"PointData": ("", ""), "CellData": ("", "")}
    for sec in piece:
        if sec.tag in attrs:
            attrs[sec.tag] = (sec.get("Scalars") or "", sec.get("Vectors") or "")
        for da in sec.iter("DataArray"):
            out.append((SECTION[sec.tag], da.get("Name"), int(da.get("NumberOfComponents")), PREC[da.get("type")],
                        np.array(da.text.split(), dtype=np.float64)))
    return int(piece.get("NumberOfCells")), int(piece.get("NumberOfPoints")), out, attrs


def ref_lib():
    import oracle.oracle as orc
    if not orc.available("ref"):
        pytest.skip("oracle/_ref not built (the reference's VTK writer classes)")
    return orc.load("ref")


def check_against_reference_writer(gv, tmp_path, **kw):
    o = ref_lib()
    f = gv.writeVTK("t_ascii", path=str(tmp_path), **kw)
    ncells, npoints, arrs, attrs = piece_arrays(f)
    for k, ot in enumerate(OUTPUT_TYPES):
        got = open(gv.writeVTK("t_" + ot, path=str(tmp_path), outputType=ot, **kw), "rb").read()
        want = o.vtu_piece(k, ncells, npoints, arrs, attrs["PointData"], attrs["CellData"])
        assert got == want, f"{ot}: file differs from the reference VTUWriter's ({len(got)} vs {len(want)} bytes)"
    return arrs


@pytest.mark.parametrize("cfg", [make_cfg(dim=2, N=(4, 2), upper=(1.0, 0.5)), make_cfg(dim=3, N=(4, 2, 3), upper=(2.0, 0.5, 0.75)),
                                 make_cfg(dim=3, N=(5, 3, 2), upper=(5.0, 3.0, 2.0))])
def test_grid_only_files_match_the_reference_writer(cfg, tmp_path):
    arrs = check_against_reference_writer(product_grid(cfg, 0, 1).leafGridView(), tmp_path)
    assert [a[1] for a in arrs] == ["Coordinates", "connectivity", "offsets", "types"]


def test_unknown_output_type(tmp_path):
    gv = product_grid(make_cfg(dim=2, N=(2, 2)), 0, 1).leafGridView()
    with pytest.raises(KeyError):
        gv.writeVTK("x", path=str(tmp_path), outputType="binary")


@pytest.mark.gpu
@pytest.mark.parametrize("float64", [False, True])
def test_field_files_match_the_reference_writer(tmp_path, float64):
    import torch
    cfg = make_cfg(dim=3, N=(5, 4, 3), upper=(5.0, 2.0, 0.75))
    gv = product_grid(cfg, 0, 1).leafGridView()
    rng = np.random.default_rng(5)
    c = torch.from_numpy(rng.integers(-64, 64, gv.size(0)) / 8.0).cuda()
    vel = torch.from_numpy(rng.integers(-64, 64, (gv.size(0), 2)) / 4.0).cuda()
    p = torch.from_numpy(rng.integers(-64, 64, gv.size(3)) / 2.0).cuda()
    arrs = check_against_reference_writer(gv, tmp_path, celldata={"c": c, "velocity": vel}, pointdata={"p": p}, float64=float64)
    assert [a[1] for a in arrs] == ["p", "c", "velocity", "Coordinates", "connectivity", "offsets", "types"]


def test_sequence_writer_files_and_collection(tmp_path, monkeypatch):
    """VTKSequenceWriter (io/file/vtk/vtksequencewriterbase.hh:105-151): files `name-%05d.vtu` (serial) / `s####-name-%05d.pvtu` +
    pieces (parallel) and the .pvd collection rank 0 rewrites at every step, character for character; setTimeSteps continues a
    sequence after a restart"""
    import dune_grid_b200 as dg
    monkeypatch.chdir(tmp_path)
    cfg = make_cfg(dim=2, N=(8, 4))
    gv = product_grid(cfg, 0, 1).leafGridView()
    w = dg.VTKSequenceWriter(gv, "run", path="out")
    os.makedirs("out")
    for t in (0.0, 0.125, 1e-7):
        w.write(t)
    assert sorted(os.listdir("out")) == ["run-00000.vtu", "run-00001.vtu", "run-00002.vtu"]
    expect = ('<?xml version="1.0"?> \n<VTKFile type="Collection" version="0.1" byte_order="LittleEndian"> \n<Collection> \n'
              '<DataSet timestep="0" group="" part="0" name="" file="out/run-00000.vtu"/> \n'
              '<DataSet timestep="0.125" group="" part="0" name="" file="out/run-00001.vtu"/> \n'
              '<DataSet timestep="1e-07" group="" part="0" name="" file="out/run-00002.vtu"/> \n'
              '</Collection> \n</VTKFile> \n')
    assert open("run.pvd").read() == expect
    check_grid("out/run-00001.vtu", gv, 2)
    # restart: a new writer continues the numbering
    w2 = dg.VTKSequenceWriter(gv, "run", path="out")
    w2.setTimeSteps(w.getTimeSteps())
    assert w2.write(2.5) == "out/run-00003.vtu" and open("run.pvd").read().count("<DataSet") == 4
    # two ranks: pieces + header per step under path/extendpath, the collection names the headers
    ws = [dg.VTKSequenceWriter(product_grid(cfg, r, 2).leafGridView(), "par", path="out", extendpath="pieces") for r in range(2)]
    os.makedirs("out/pieces")
    for t in (0.0, 1.0):
        for x in ws:
            x.write(t)
    assert sorted(os.listdir("out/pieces")) == ["s0002-p0000-par-00000.vtu", "s0002-p0000-par-00001.vtu", "s0002-p0001-par-00000.vtu",
                                                "s0002-p0001-par-00001.vtu", "s0002-par-00000.pvtu", "s0002-par-00001.pvtu"]
    assert 'file="out/pieces/s0002-par-00001.pvtu"' in open("par.pvd").read()
