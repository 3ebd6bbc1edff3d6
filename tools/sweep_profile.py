#!/usr/bin/env python
"""One launch of each materialising sweep kernel on config 4 (384^3 tensor-product, x periodic) and on 512^3, for
`ncu --set full -k regex:k_intersections|k_index_rows|k_entities` (developer tool -> profiles/rNN_sweeps_ncu.json)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import dune_grid_b200 as dg  # noqa: E402


def graded(n, r=1.005):
    w = r ** np.arange(n)
    x = np.concatenate([[0.0], np.cumsum(w)])
    return x / x[-1]


def main():
    c = [graded(384), graded(384), graded(384)]
    g4 = dg.YaspGrid(coords=c, periodic=1, coordinates="tensorproduct")
    gv4 = g4.leafGridView()
    for _ in range(2):
        out = gv4.intersections(dg.All_Partition)
        torch.cuda.synchronize()
        del out
    for _ in range(2):
        out = gv4.elements(0, dg.All_Partition)
        torch.cuda.synchronize()
        del out
    m4 = dg.MultipleCodimMultipleGeomTypeMapper(gv4, [1, 1, 1, 1])
    for _ in range(2):
        out = m4.index(1)
        torch.cuda.synchronize()
        del out
    torch.cuda.empty_cache()
    g = dg.YaspGrid(dim=3, N=(512, 512, 512), upper=(1.0, 1.0, 1.0))
    gv = g.leafGridView()
    m = dg.MultipleCodimMultipleGeomTypeMapper(gv, [1, 1, 1, 1])
    for _ in range(2):
        out = m.index(1)
        torch.cuda.synchronize()
        del out
    for _ in range(2):
        out = gv.elements(0, dg.All_Partition)
        torch.cuda.synchronize()
        del out
    print("sweep_profile done")


if __name__ == "__main__":
    main()
