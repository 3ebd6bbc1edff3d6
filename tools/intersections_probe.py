#!/usr/bin/env python
"""k_intersections throughput against grid size, coordinate mode and periodicity (developer tool): which property of BASELINE
config 4 (384^3 tensor-product, x periodic) costs the bandwidth?"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import dune_grid_b200 as dg  # noqa: E402
from tools.sweep_bench import graded, nbytes, peak, timeit  # noqa: E402


def main():
    cases = []
    for n in (256, 384, 512):
        cases.append((f"{n}^3 equidistant", dict(dim=3, N=(n, n, n), upper=(1.0, 1.0, 1.0))))
    cases.append(("384^3 equidistant x periodic", dict(dim=3, N=(384, 384, 384), upper=(1.0, 1.0, 1.0), periodic=1)))
    cases.append(("384^3 equidistant y periodic", dict(dim=3, N=(384, 384, 384), upper=(1.0, 1.0, 1.0), periodic=2)))
    cases.append(("384^3 tensor", dict(coords=[graded(384)] * 3, coordinates="tensorproduct")))
    cases.append(("384^3 tensor x periodic (config 4)", dict(coords=[graded(384)] * 3, coordinates="tensorproduct", periodic=1)))
    cases.append(("382x384x384 equidistant x periodic (384 stored)", dict(dim=3, N=(382, 384, 384), upper=(1.0, 1.0, 1.0), periodic=1)))
    for name, kw in cases:
        g = dg.YaspGrid(**kw)
        gv = g.leafGridView()
        out = gv.intersections(dg.All_Partition)
        b = nbytes(out)
        ms = timeit(lambda: gv.intersections(dg.All_Partition), 3)
        geo = {k: out[k] for k in ("neighbor", "boundary", "outside", "boundarySegmentIndex")}
        del out
        ms_topo = timeit(lambda: gv.intersections(dg.All_Partition, geometry=False), 3)
        print(json.dumps({"case": name, "cells": gv.size(0), "ms": ms, "GBps": b / ms / 1e6, "frac": b / ms / 1e6 / peak(),
                          "topology_only_ms": ms_topo, "topology_only_frac": nbytes(geo) / ms_topo / 1e6 / peak()}), flush=True)
        del geo, g, gv
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
