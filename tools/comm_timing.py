#!/usr/bin/env python
"""Stand-alone dgb_communicate on N GPUs (developer tool; run under torchrun, one rank per GPU): the cell field of the partitioned
n^3 grid over InteriorBorder_All (set) and a vertex field over InteriorBorder_InteriorBorder (add). Max over ranks, CUDA events.
DGB_COMM_PEER=0 selects the NCCL form (pack kernel, grouped ncclSend/ncclRecv, unpack kernel)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import dune_grid_b200 as dg  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    comm = dg.Communicator.from_torch_distributed()
    g = dg.YaspGrid(dim=3, N=(n, n, n), upper=(1.0, 1.0, 1.0), overlap=1, rank=rank, nranks=world, comm=comm)
    gv = g.leafGridView()
    cases = [("cells, InteriorBorder_All, set", dg.mcmgElementLayout(3), dg.InteriorBorder_All_Interface, "set"),
             ("vertices, InteriorBorder_InteriorBorder, add", dg.mcmgVertexLayout(3), dg.InteriorBorder_InteriorBorder_Interface, "add")]
    for name, layout, iftype, op in cases:
        m = dg.MultipleCodimMultipleGeomTypeMapper(gv, layout)
        x = torch.zeros(m.size, dtype=torch.float64, device="cuda")

        def f():
            gv.communicate(m, [x], iftype, dg.ForwardCommunication, op)
        for _ in range(10):
            f()
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            f()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps, float(g.lastBytesSent())], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            ms, nbytes = t[0].item(), t[1].item()
            print(json.dumps({"what": name, "grid": n, "ranks": world, "peer_form": os.environ.get("DGB_COMM_PEER", "1") != "0", "ms": round(ms, 5),
                              "bytes_per_rank": int(nbytes), "nvlink_gbs": round(nbytes / ms / 1e6, 1), "frac_of_900": round(nbytes / ms / 1e6 / 900, 4)}), flush=True)
    dist.barrier()
    comm.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
