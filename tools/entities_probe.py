#!/usr/bin/env python
"""Throughput of the entity sweep (index, partition type, lower, upper, centre, volume: 85 B per entity in 3-D) per codimension;
developer tool. Run once per setting of DGB_SWEEP_LINEAR (read once per process)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import dune_grid_b200 as dg  # noqa: E402
from tools.sweep_bench import graded, nbytes, peak, timeit  # noqa: E402


def main():
    tag = os.environ.get("DGB_SWEEP_LINEAR", "auto")
    cases = [("384^3 equidistant", dg.YaspGrid(dim=3, N=(384, 384, 384), upper=(1.0, 1.0, 1.0))),
             ("384^3 tensor, x periodic", dg.YaspGrid(coords=[graded(384)] * 3, periodic=1, coordinates="tensorproduct"))]
    for name, g in cases:
        gv = g.leafGridView()
        for codim in range(4):
            out = gv.elements(codim, dg.All_Partition)
            nb = nbytes(out)
            del out
            torch.cuda.empty_cache()
            ms = timeit(lambda: gv.elements(codim, dg.All_Partition), 5)
            print(json.dumps({"linear": tag, "grid": name, "codim": codim, "ms": round(ms, 4), "GBps": round(nb / ms / 1e6, 1),
                              "frac_of_measured_hbm": round(nb / ms / 1e6 / peak(), 3)}), flush=True)
        del gv
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
