#!/usr/bin/env python
"""A few step-kernel launches of one box shape with one staging form, for `ncu --set full -k regex:k_fv_ring -s 4 -c 1`
(developer tool). usage: python tools/fv_profile_case.py aligned512|rank7of8|rank1of4|rank0of2 pair|a16|8 [steps]"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
case, staging = sys.argv[1], sys.argv[2]
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 8
os.environ["DGB_FV_STAGING"] = staging
import torch  # noqa: E402
import dune_grid_b200 as dg  # noqa: E402

PARAMS = [1.0, 1.0, 1.0, 0.0, 0.25, 0.25, 0.25, 0.125, 0.2]


class _FakeComm:
    h = ctypes.c_void_p(1)


if case == "aligned512":
    g = dg.YaspGrid(dim=3, N=(512, 512, 512), upper=(1.0, 1.0, 1.0), overlap=1)
else:
    r, P = {"rank7of8": (7, 8), "rank1of4": (1, 4), "rank0of2": (0, 2), "rank0of8": (0, 8)}[case]
    g = dg.YaspGrid(dim=3, N=(1024, 1024, 1024), upper=(1.0, 1.0, 1.0), overlap=1, rank=r, nranks=P)
    g.comm = _FakeComm()
fv = dg.FiniteVolume(g, 0, 0, PARAMS, "separable")
c = fv.initial()
cn = torch.empty_like(c)
fv.bootstrap_local()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(steps):
    if i == steps // 2:
        e0.record()
    fv.step_local(c, cn)
    c, cn = cn, c
e1.record()
torch.cuda.synchronize()
print(case, staging, fv.kernel_info(), "ms/step", e0.elapsed_time(e1) / (steps - steps // 2))
