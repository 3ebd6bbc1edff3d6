#!/usr/bin/env python
"""Phase timing of the partitioned FV step on N GPUs (developer tool; run under torchrun, one rank per GPU):
local kernel, scalar all-reduce, communicate(InteriorBorder_All), full dgb_fv_step. Max over ranks, CUDA events."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import dune_grid_b200 as dg  # noqa: E402
from dune_grid_b200.capi import lib, check  # noqa: E402

PARAMS = [1.0, 1.0, 1.0, 0.0, 0.25, 0.25, 0.25, 0.125, 0.2]


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    comm = dg.Communicator.from_torch_distributed()
    g = dg.YaspGrid(dim=3, N=(n, n, n), upper=(1.0, 1.0, 1.0), overlap=1, rank=rank, nranks=world, comm=comm)
    gv = g.leafGridView()
    fv = dg.FiniteVolume(g, 0, 0, PARAMS, "separable")
    m = dg.MultipleCodimMultipleGeomTypeMapper(gv, dg.mcmgElementLayout(3))
    if os.environ.get("DGB_SIDE_STREAM"):
        torch.cuda.set_stream(torch.cuda.Stream())
    c = fv.initial()
    cn = torch.empty_like(c)
    fv.step(c, cn)
    one = torch.ones(1, dtype=torch.float64, device="cuda")
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def timeit(f, reps=20):
        for _ in range(3):
            f()
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            f()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    res = {}
    res["kernel"] = timeit(lambda: fv.step_local(c, cn))
    res["allreduce"] = timeit(lambda: check(lib.dgb_allreduce_max(C.c_void_p(one.data_ptr()), comm.h, stream)))
    res["communicate"] = timeit(lambda: gv.communicate(m, [cn], dg.InteriorBorder_All_Interface, dg.ForwardCommunication, "set"))
    res["step"] = timeit(lambda: fv.step(c, cn))
    if rank == 0:
        print(f"n={n} world={world} dims={g.torusDims()} " + " ".join(f"{k}={v:.4f}ms" for k, v in res.items()), flush=True)
    dist.barrier()
    comm.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
