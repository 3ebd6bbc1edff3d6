#!/usr/bin/env python
"""Step time of the production FV kernel at 512^3 (single rank) against the planes marched per CTA (DGB_FV_ZCHUNK is read when the
FiniteVolume object is created); developer tool.  usage: python tools/zchunk_sweep.py [n] [reps] [zc ...]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import dune_grid_b200 as dg  # noqa: E402
from tools.fv_staging_bench import PARAMS, PEAK, timeit  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    zcs = [int(x) for x in sys.argv[3:]] or [0, 512, 256, 192, 171, 128, 103, 86, 64, 32]
    g = dg.YaspGrid(dim=3, N=(n, n, n), upper=(1.0, 1.0, 1.0))
    for zc in zcs:
        if zc:
            os.environ["DGB_FV_ZCHUNK"] = str(zc)
        else:
            os.environ.pop("DGB_FV_ZCHUNK", None)
        fv = dg.FiniteVolume(g, 0, 0, PARAMS, "separable")
        state = [fv.initial(), None]
        state[1] = torch.empty_like(state[0])

        def step():
            fv.step(state[0], state[1])
            state.reverse()
        best = min(timeit(step, reps) for _ in range(3))
        print(json.dumps({"n": n, "zchunk": zc, "info": fv.kernel_info(), "ms": round(best, 5), "frac": round(16 * n ** 3 / best / 1e6 / PEAK, 4)}), flush=True)
        del fv, state
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
