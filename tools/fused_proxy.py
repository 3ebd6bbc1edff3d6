#!/usr/bin/env python
"""One-GPU proxy of the partitioned FV step (developer tool): a fully periodic single-rank n^3 grid has an overlap layer on
all six sides that is refreshed from the rank itself, so dgb_fv_step runs the same kernels as a rank of a multi-GPU job
(fused: step kernel with in-kernel pack + flag, wait, unpack; DGB_FV_FUSED=0: shell + bulk + copies + unpack).
usage: python tools/fused_proxy.py [n] [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import dune_grid_b200 as dg  # noqa: E402

PARAMS = [1.0, 1.0, 1.0, 0.0, 0.25, 0.25, 0.25, 0.125, 0.2]


def timeit(f, reps):
    for _ in range(5):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    if len(sys.argv) > 4:                    # nx ny nz periodic-mask [reps [labels]]
        N = tuple(int(x) for x in sys.argv[1:4]); per = int(sys.argv[4])
        reps = int(sys.argv[5]) if len(sys.argv) > 5 else 200
        only = sys.argv[6].split(",") if len(sys.argv) > 6 else None
    else:
        n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
        reps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
        N, per, only = (n, n, n), 7, None
    n = N
    g = dg.YaspGrid(dim=3, N=N, upper=(1.0, 1.0, 1.0), overlap=1, periodic=per)
    gv = g.leafGridView()
    interior = gv.size(0, dg.Interior_Partition)
    out = {}
    for label, env in (("direct", {"DGB_FV_FUSED": "1"}), ("fused", {"DGB_FV_FUSED": "1"}), ("overlapped", {"DGB_FV_FUSED": "0"}),
                       ("serial", {"DGB_FV_FUSED": "0", "DGB_FV_NO_SPLIT": "1"})):
        if only and label not in only:
            continue
        for k in ("DGB_FV_FUSED", "DGB_FV_NO_SPLIT"):
            os.environ.pop(k, None)
        os.environ.update(env)
        fv = dg.FiniteVolume(g, 0, 0, PARAMS, "separable")
        c = fv.initial()
        cn = torch.empty_like(c)
        state = [c, cn]
        if label == "direct":
            assert fv.register_fields(c, cn) == 2

        def step():
            fv.step(state[0], state[1])
            state.reverse()
        ms = timeit(step, reps)
        out[label] = (ms, fv.exchange_mode())
        if label in ("fused", "direct") and "kernel only" not in out:
            ms_local = timeit(lambda: fv.step_local(c, cn), reps)
            out["kernel only"] = (ms_local, "-")
        del fv
    print(f"n={n} stored={gv.size(0)} interior={interior}")
    for k, (ms, mode) in out.items():
        print(f"{k:12s} {ms:.4f} ms  {16 * interior / ms / 1e6:7.0f} GB/s  mode={mode}")


if __name__ == "__main__":
    main()
