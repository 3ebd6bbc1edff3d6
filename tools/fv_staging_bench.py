#!/usr/bin/env python
"""Step-kernel timings on ONE GPU for the box shapes of the BASELINE configs (developer tool -> profiles/rNN_fv_staging.jsonl):
the aligned 512^3 single-rank box, rank boxes of the partitioned 1024^3 grid (kernel only: dgb_fv_step_local of rank r of an
N-rank grid, nothing exchanged), and the fully periodic single-rank proxy whose stored box is 513^3 (fused self-exchange),
each with the staging forms of k_fv_ring (DGB_FV_STAGING = pair | a16 | 8).
usage: python tools/fv_staging_bench.py [reps]"""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import dune_grid_b200 as dg  # noqa: E402

PARAMS = [1.0, 1.0, 1.0, 0.0, 0.25, 0.25, 0.25, 0.125, 0.2]
PEAK = 6549.8
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


class _FakeComm:
    h = ctypes.c_void_p(1)       # never dereferenced: only the *_local entry points are called


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


def run(label, grid, staging, reps, local):
    os.environ["DGB_FV_STAGING"] = staging
    fv = dg.FiniteVolume(grid, 0, 0, PARAMS, "separable")
    gv = grid.leafGridView()
    interior = gv.size(0, dg.Interior_Partition)
    state = [fv.initial(), None]
    state[1] = torch.empty_like(state[0])
    if local:
        fv.bootstrap_local()

    def step():
        (fv.step_local if local else fv.step)(state[0], state[1])
        state.reverse()
    ms = timeit(step, reps)
    info = fv.kernel_info()
    b = gv.boxes(0)[0]
    print(json.dumps({"case": label, "staging": staging, "stored_box": b["overlapfront"][1], "interior_origin_in_box":
                      [b["interior"][0][i] - b["overlapfront"][0][i] for i in range(3)], "interior_cells": interior, "ms": ms,
                      "GBps": 16 * interior / ms / 1e6, "frac_of_measured_hbm": 16 * interior / ms / 1e6 / PEAK,
                      "kernel": info, "exchange": fv.exchange_mode() if not local else "none (kernel only)"}), flush=True)
    del fv, state


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    quick = len(sys.argv) > 2 and sys.argv[2] == "quick"
    g1 = dg.YaspGrid(dim=3, N=(512, 512, 512), upper=(1.0, 1.0, 1.0), overlap=1)
    for st in ("pair", "a16", "8", "tma", "a16", "tma"):
        run("512^3 single rank (config 2)", g1, st, reps, False)
    del g1
    if len(sys.argv) > 2 and sys.argv[2] == "tma":              # the boxes tensor-map staging applies to: 512^3 and the even rank boxes of P = 4
        for r in (0, 1):
            g = dg.YaspGrid(dim=3, N=(1024, 1024, 1024), upper=(1.0, 1.0, 1.0), overlap=1, rank=r, nranks=4)
            g.comm = _FakeComm()
            for st in ("pair", "tma"):
                run(f"1024^3 rank {r} of 4 {g.torusDims()} (config 3), kernel only", g, st, max(10, reps // 2), True)
            del g
            torch.cuda.empty_cache()
        return
    for P, ranks in (((8, (7,)), (2, (1,)), (4, (1,))) if quick else ((8, (0, 7)), (2, (0, 1)), (4, (0, 1)))):
        for r in ranks:
            g = dg.YaspGrid(dim=3, N=(1024, 1024, 1024), upper=(1.0, 1.0, 1.0), overlap=1, rank=r, nranks=P)
            g.comm = _FakeComm()
            for st in (("pair", "8", "tma") if P == 4 else ("pair", "8")):
                run(f"1024^3 rank {r} of {P} {g.torusDims()} (config 3), kernel only", g, st, max(10, reps // (8 // P)), True)
            del g
            torch.cuda.empty_cache()
    gp = dg.YaspGrid(dim=3, N=(511, 511, 511), upper=(1.0, 1.0, 1.0), overlap=1, periodic=7)
    for st in ("pair", "8"):
        run("511^3 fully periodic single rank: 513^3 stored, fused self-exchange", gp, st, reps, False)


if __name__ == "__main__":
    main()
