#!/usr/bin/env python
"""Times the per-rank pieces of the partitioned FV step on ONE GPU (developer tool): the local kernel of rank `r` of a
P-rank n^3 job, and the pack / unpack kernels of communicate(InteriorBorder_All, Forward) for one double per cell.
usage: python tools/halo_timing.py n P [rank]"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import dune_grid_b200 as dg  # noqa: E402

PARAMS = [1.0, 1.0, 1.0, 0.0, 0.25, 0.25, 0.25, 0.125, 0.2]


class _FakeComm:
    h = ctypes.c_void_p(1)


def timeit(f, reps=20):
    for _ in range(3):
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
    n, P = int(sys.argv[1]), int(sys.argv[2])
    r = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    g = dg.YaspGrid(dim=3, N=(n, n, n), upper=(1.0, 1.0, 1.0), overlap=1, rank=r, nranks=P)
    if P > 1:
        g.comm = _FakeComm()
    gv = g.leafGridView()
    fv = dg.FiniteVolume(g, 0, 0, PARAMS, "separable")
    c = fv.initial()
    cn = torch.empty_like(c)
    fv.bootstrap_local()
    interior = gv.size(0, dg.Interior_Partition)
    ms = timeit(lambda: fv.step_local(c, cn))
    print(f"n={n} P={P} rank={r} dims={g.torusDims()} stored={gv.size(0)} interior={interior}")
    print(f"kernel   {ms:.4f} ms  {16 * interior / ms / 1e6:.0f} GB/s")
    m = dg.MultipleCodimMultipleGeomTypeMapper(gv, dg.mcmgElementLayout(3))
    send, recv = gv.commPlan(m, dg.InteriorBorder_All_Interface, dg.ForwardCommunication, [1])
    ns, nr = sum(s[2] for s in send), sum(s[2] for s in recv)
    ms = timeit(lambda: gv.commPack(m, [cn], dg.InteriorBorder_All_Interface, dg.ForwardCommunication, [1]))
    print(f"pack     {ms:.4f} ms  {ns} doubles in {len(send)} segments  {16 * ns / ms / 1e6:.0f} GB/s")
    ms = timeit(lambda: gv.commUnpack(m, [cn], dg.InteriorBorder_All_Interface, dg.ForwardCommunication, "set", [1]))
    print(f"unpack   {ms:.4f} ms  {nr} doubles in {len(recv)} segments  {16 * nr / ms / 1e6:.0f} GB/s")


if __name__ == "__main__":
    main()
