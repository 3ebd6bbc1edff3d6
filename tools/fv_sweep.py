#!/usr/bin/env python
"""Tuning sweep of the 3-D FV kernel (developer tool, not part of the product): times dgb_fv_step on an n^3 grid for
every (DGB_FV_VARIANT, DGB_FV_ZCHUNK) pair given on the command line and prints ms/step and GB/s (16 B/cell)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import dune_grid_b200 as dg  # noqa: E402

PARAMS = [1.0, 1.0, 1.0, 0.0, 0.25, 0.25, 0.25, 0.125, 0.2]


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    variants = [int(x) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["0"])]
    chunks = [int(x) for x in (sys.argv[3].split(",") if len(sys.argv) > 3 else ["32"])]
    nranks = int(sys.argv[4]) if len(sys.argv) > 4 else 1          # rank 0 of an nranks job (kernel only, no exchange)
    g = dg.YaspGrid(dim=3, N=(n, n, n), upper=(1.0, 1.0, 1.0), overlap=1, rank=0, nranks=nranks)
    gv = g.leafGridView()
    cells = gv.size(0, dg.Interior_Partition)
    for var in variants:
        for zc in chunks:
            os.environ["DGB_FV_VARIANT"] = str(var)
            os.environ["DGB_FV_ZCHUNK"] = str(zc)
            if nranks > 1:
                class _C:
                    h = None
                g.comm = None
            fv = dg.FiniteVolume(g, 0, 0, PARAMS, "separable") if nranks == 1 else None
            if fv is None:
                break
            c = fv.initial()
            cn = torch.empty_like(c)
            for _ in range(5):
                fv.step(c, cn)
                c, cn = cn, c
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            steps = 30
            e0.record()
            for _ in range(steps):
                fv.step(c, cn)
                c, cn = cn, c
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            print(f"n={n} variant={var} zchunk={zc} ms/step={ms:.4f} GB/s={16 * cells / ms / 1e6:.1f} "
                  f"frac_of_6549.8={16 * cells / ms / 1e6 / 6549.8:.3f} sum={c.sum().item():.6f}", flush=True)
            del fv, c, cn


if __name__ == "__main__":
    main()
