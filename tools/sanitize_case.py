#!/usr/bin/env python
"""Small invocations of every FV exchange form on ONE GPU for `compute-sanitizer --tool memcheck|racecheck` (developer tool;
logs -> profiles/rNN_sanitizer_*.log): single-rank periodic grids, so that the fused / direct / hybrid peer-memory exchange
(flags, fences, in-kernel pack, side cells of the ring) runs with the rank as its own receiver; odd row and plane strides;
the three staging forms of the ring kernel; the general kernel; the pipelined host entry point; communicate and the sweeps."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import dune_grid_b200 as dg  # noqa: E402

PARAMS = [0.8, -0.5, 0.25, 0.1, 0.5, 0.3, 0.45, 0.1, 0.45]


def run(N, periodic, mode, staging, register, steps=3, upper=(1.0, 0.6, 0.9)):
    os.environ["DGB_FV_STAGING"] = staging
    g = dg.YaspGrid(dim=3, N=N, upper=upper, periodic=periodic, overlap=1)
    fv = dg.FiniteVolume(g, 0, 0, PARAMS, mode)
    c = fv.initial()
    cn = torch.empty_like(c)
    if register:
        fv.register_fields(c, cn)
    for _ in range(steps):
        fv.step(c, cn)
        c, cn = cn, c
    fv.fence()
    dt = fv.last_dt()
    print(f"N={N} periodic={periodic} mode={mode} staging={staging} registered={register}: exchange={fv.exchange_mode()} "
          f"kernel={fv.kernel_info()['kind']} dt={dt:.6g} sum={c.sum().item():.12g}", flush=True)
    fv.close()


def main():
    small = len(sys.argv) > 1 and sys.argv[1] == "small"
    for staging in ("pair", "8", "a16"):
        run((131, 21, 19), 7, "separable", staging, False)
        run((131, 21, 19), 7, "separable", staging, True)
    run((260, 10, 18), 1, "separable", "auto", True, upper=(1.0, 0.05, 0.1))       # wide tile, x faces only: hybrid side cells
    run((40, 23, 37), 7, "exact", "auto", True)
    run((40, 23, 37), 5, "exact", "auto", False)
    # pipelined host entry point with the fused exchange (unpack into the mapped host result)
    g = dg.YaspGrid(dim=3, N=(20, 12, 126), upper=(1.0, 0.6, 4.0), periodic=7)
    fv = dg.FiniteVolume(g, 0, 0, [0.8, -0.5, 1.25, 0.1, 0.5, 0.3, 2.0, 0.3, 1.45], "separable")
    h_in = fv.initial().cpu().pin_memory()
    h_out = torch.empty_like(h_in).pin_memory()
    for _ in range(2):
        fv.step_host(h_in, h_out)
        h_in, h_out = h_out, h_in
    print("step_host:", fv.host_path(), fv.exchange_mode(), flush=True)
    fv.close()
    if small:
        return
    # communicate of every interface on a periodic grid, the materialising sweeps, the functor sweeps
    g = dg.YaspGrid(dim=3, N=(9, 8, 7), upper=(1.0, 1.0, 1.0), periodic=5, overlap=2)
    gv = g.leafGridView()
    m = dg.MultipleCodimMultipleGeomTypeMapper(gv, [1, 1, 0, 2])
    a = torch.rand(m.size, dtype=torch.float64, device="cuda")
    for iftype in range(5):
        for direction in (0, 1):
            gv.communicate(m, [a], iftype, direction, "add")
    gv.elements(0); gv.elements(2); gv.intersections(); m.index(1); m.subIndex(3)
    gv.forEachElement("gauss5_exp_norm"); gv.forEachIntersection("boundary_flux_x")
    qp = np.array([[0.5, 0.5], [0.2, 0.7]])
    gv.geometryGridIntersections(1, [0.05], qp)
    torch.cuda.synchronize()
    print("communicate / sweeps done", flush=True)


if __name__ == "__main__":
    main()
