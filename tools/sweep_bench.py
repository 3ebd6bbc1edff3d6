#!/usr/bin/env python
"""Throughput of the materialising sweeps at the BASELINE.json sizes (developer tool -> profiles/rNN_sweeps.json):
bytes written per launch group / CUDA-event time, against the measured HBM peak. One JSON line per kernel group."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import dune_grid_b200 as dg  # noqa: E402


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def timeit(f, reps=10):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def nbytes(d):
    return sum(t.numel() * t.element_size() for t in d.values())


def report(name, config, ms, out_bytes, extra=None):
    gbs = out_bytes / ms / 1e6
    line = {"kernel": name, "config": config, "ms": ms, "bytes_written": out_bytes, "GBps": gbs, "frac_of_measured_hbm": gbs / peak()}
    if extra:
        line.update(extra)
    print(json.dumps(line), flush=True)


def graded(n, r=1.005):
    w = r ** np.arange(n)
    x = np.concatenate([[0.0], np.cumsum(w)])
    return x / x[-1]


def main():
    gauss = 0.5 + np.array([-0.5, 0.5]) / np.sqrt(3.0)
    # config 2: YaspGrid<3> 512^3 equidistant - element, mapper and intersection sweeps
    g = dg.YaspGrid(dim=3, N=(512, 512, 512), upper=(1.0, 1.0, 1.0))
    gv = g.leafGridView()
    out = gv.elements(0, dg.All_Partition)
    report("k_entities (codim 0: index, ptype, lower, upper, center, volume)", "512^3 equidistant",
           timeit(lambda: gv.elements(0, dg.All_Partition)), nbytes(out))
    del out
    m = dg.MultipleCodimMultipleGeomTypeMapper(gv, [1, 1, 1, 1])
    for cc in (1, 3):
        out = m.subIndex(cc)
        report(f"k_subindex (codim {cc})", "512^3 equidistant", timeit(lambda: m.subIndex(cc), 5), out.numel() * 4)
        del out
    out = m.index(1)
    report("k_entities (mapper.index, codim 1, 403M faces)", "512^3 equidistant", timeit(lambda: m.index(1), 5), out.numel() * 4)
    del out
    torch.cuda.empty_cache()
    gq = dg.YaspGrid(dim=3, N=(256, 256, 256), upper=(1.0, 1.0, 1.0))
    gvq = gq.leafGridView()
    out = gvq.intersections(dg.All_Partition)
    report("k_intersections (neighbor, boundary, outside, bsi, center, volume x 6 faces)", "256^3 equidistant",
           timeit(lambda: gvq.intersections(dg.All_Partition), 5), nbytes(out))
    del out
    torch.cuda.empty_cache()
    # config 4: TensorProductCoordinates graded 384^3, x periodic
    c = [graded(384), graded(384), graded(384)]
    g4 = dg.YaspGrid(coords=c, periodic=1, coordinates="tensorproduct")
    gv4 = g4.leafGridView()
    out = gv4.elements(0, dg.All_Partition)
    report("k_entities (codim 0)", "384^3 tensor-product graded, x periodic", timeit(lambda: gv4.elements(0, dg.All_Partition)), nbytes(out))
    del out
    out = gv4.intersections(dg.All_Partition)
    report("k_intersections", "384^3 tensor-product graded, x periodic", timeit(lambda: gv4.intersections(dg.All_Partition), 3), nbytes(out),
           {"numBoundarySegments": g4.numBoundarySegments()})
    del out
    torch.cuda.empty_cache()
    # config 5: GeometryGrid over YaspGrid<3> 128^3 refined once -> 256^3, 2x2x2 Gauss points
    g5 = dg.YaspGrid(dim=3, N=(128, 128, 128), upper=(1.0, 1.0, 1.0))
    g5.globalRefine(1)
    gv5 = g5.levelGridView(1)
    grids = np.meshgrid(gauss, gauss, gauss, indexing="ij")
    qp = np.stack([x.T.reshape(-1) for x in grids], axis=1)
    wts = np.full(8, 0.125)
    out = gv5.geometryGrid(1, [0.05], qp, dg.All_Partition, corners=False)
    n = gv5.size(0)
    report("k_geogrid (global + J^T + J^-T + detJ at 8 Gauss points)", "GeometryGrid over 256^3 (128^3 refined once)",
           timeit(lambda: gv5.geometryGrid(1, [0.05], qp, dg.All_Partition, corners=False), 3), nbytes(out), {"bytes_per_cell": nbytes(out) / n})
    del out
    torch.cuda.empty_cache()
    # GeometryGrid intersections (a9): face corners, centre, volume, three normals at the face centre + centerUnitOuterNormal
    for label, fq in (("face centre", np.array([[0.5, 0.5]])), ("2x2 Gauss points", np.stack([x.T.reshape(-1) for x in np.meshgrid(gauss, gauss, indexing="ij")], axis=1))):
        out = gv5.geometryGridIntersections(1, [0.05], fq, dg.All_Partition)
        report(f"k_geogrid_faces (corners, centre, volume, normals at {label})", "GeometryGrid over 256^3",
               timeit(lambda: gv5.geometryGridIntersections(1, [0.05], fq, dg.All_Partition), 3), nbytes(out), {"bytes_per_cell": nbytes(out) / n})
        del out
        torch.cuda.empty_cache()
    ms = timeit(lambda: gv5.geometryGridVolume(1, [0.05], qp, wts, dg.All_Partition), 5)
    vol = gv5.geometryGridVolume(1, [0.05], qp, wts, dg.All_Partition).item()
    print(json.dumps({"kernel": "k_geogrid_volume (reduction)", "config": "GeometryGrid over 256^3", "ms": ms, "cells_per_s": n / ms * 1e3,
                      "volume": vol}), flush=True)
    out = gv5.geometry(qp, dg.All_Partition)
    report("k_geometry (global + J^-T + detJ at 8 points, axis-aligned)", "256^3", timeit(lambda: gv5.geometry(qp, dg.All_Partition), 3), nbytes(out))


if __name__ == "__main__":
    main()
