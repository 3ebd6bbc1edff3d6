#!/usr/bin/env python
"""Step time at n^3 of the FV paths next to the production kernel (developer tool): the general z-march kernel (problems whose
velocity depends on z, here problem 2 = shear), the bit-exact mode, and the production ring kernel for reference."""
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
    g = dg.YaspGrid(dim=3, N=(n, n, n), upper=(1.0, 1.0, 1.0))
    for label, problem, params, mode in (("ring (problem 0, separable)", 0, PARAMS, "separable"),
                                         ("march (problem 2 shear, separable)", 2, [1.0, 0.5, 0.7, 0.0, 0.25, 0.25, 0.25, 0.125, 0.2], "separable"),
                                         ("march (problem 0, exact)", 0, PARAMS, "exact")):
        fv = dg.FiniteVolume(g, 0, problem, params, mode)
        state = [fv.initial(), None]
        state[1] = torch.empty_like(state[0])

        def step():
            fv.step(state[0], state[1])
            state.reverse()
        ms = min(timeit(step, 50) for _ in range(3))
        print(json.dumps({"n": n, "path": label, "info": fv.kernel_info(), "ms": round(ms, 4), "frac_of_hbm_roofline": round(16 * n ** 3 / ms / 1e6 / PEAK, 3)}), flush=True)
        del fv, state
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
