#!/usr/bin/env python
"""One-GPU proxy of dgb_communicate (developer tool): a single-rank grid that is periodic in one direction exchanges two faces
with itself through the same pack / unpack kernels a partitioned grid uses. usage: python tools/comm_proxy.py [reps]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import dune_grid_b200 as dg  # noqa: E402
from tools.fv_staging_bench import timeit  # noqa: E402


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    for name, N, per in [("x faces 2 x 1M cells", (510, 1024, 1024), 1), ("y faces 2 x 1M cells", (1024, 510, 1024), 2), ("z faces 2 x 1M cells", (1024, 1024, 510), 4),
                         ("six faces of 510^3", (510, 510, 510), 7)]:
        g = dg.YaspGrid(dim=3, N=N, upper=(1.0, 1.0, 1.0), overlap=1, periodic=per)
        gv = g.leafGridView()
        m = dg.MultipleCodimMultipleGeomTypeMapper(gv, dg.mcmgElementLayout(3))
        x = torch.zeros(m.size, dtype=torch.float64, device="cuda")
        ms = min(timeit(lambda: gv.communicate(m, [x], dg.InteriorBorder_All_Interface, dg.ForwardCommunication, "set"), reps) for _ in range(3))
        sends, _ = gv.commPlan(m, dg.InteriorBorder_All_Interface, dg.ForwardCommunication, [1])
        cells = sum(s[2] for s in sends)
        print(json.dumps({"case": name, "peer_form": os.environ.get("DGB_COMM_PEER", "1") != "0", "ms": round(ms, 5), "cells": cells,
                          "GBps_payload": round(16 * cells / ms / 1e6, 1)}), flush=True)
        del x, m, gv, g
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
