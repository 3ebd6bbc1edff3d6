#!/usr/bin/env python
"""Throughput of mapper.index / mapper.subIndex (4 B per entity written) at the BASELINE sizes; developer tool.
Run once per setting of DGB_IDX_VEC (the switch is read once per process)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import dune_grid_b200 as dg  # noqa: E402
from tools.sweep_bench import graded, peak, timeit  # noqa: E402


def main():
    tag = os.environ.get("DGB_IDX_VEC", "1")
    cases = [("512^3 equidistant", dg.YaspGrid(dim=3, N=(512, 512, 512), upper=(1.0, 1.0, 1.0))),
             ("384^3 tensor, x periodic", dg.YaspGrid(coords=[graded(384)] * 3, periodic=1, coordinates="tensorproduct"))]
    for name, g in cases:
        gv = g.leafGridView()
        m = dg.MultipleCodimMultipleGeomTypeMapper(gv, [1, 1, 1, 1])
        for what, f in [("index codim 0", lambda: m.index(0)), ("index codim 1", lambda: m.index(1)), ("index codim 2", lambda: m.index(2)),
                        ("index codim 3", lambda: m.index(3)), ("subIndex codim 1", lambda: m.subIndex(1)),
                        ("subIndex codim 2", lambda: m.subIndex(2)), ("subIndex codim 3", lambda: m.subIndex(3))]:
            out = f()
            nb = out.numel() * 4
            del out
            ms = timeit(f, 10)
            print(json.dumps({"vec": tag, "grid": name, "what": what, "ms": round(ms, 4), "GBps": round(nb / ms / 1e6, 1),
                              "frac_of_measured_hbm": round(nb / ms / 1e6 / peak(), 3)}), flush=True)
        del m, gv
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
