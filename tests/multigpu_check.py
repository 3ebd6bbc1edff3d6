"""Multi-GPU parity of the public FV step (dgb_fv_step: shell/bulk split, NCCL exchange overlapped with the bulk kernel,
all-reduce of the CFL value) and of dgb_communicate against the CPU oracle. Run under torchrun, one rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tests/multigpu_check.py

Each rank builds its own grid, rank 0's oracle world simulates all N ranks; every rank compares its own arrays."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import dune_grid_b200 as dg  # noqa: E402
from tests.common import best_oracle, make_cfg, oracle_world, product_grid  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    comm = dg.Communicator.from_torch_distributed()
    o = best_oracle()
    cases = [
        (make_cfg(dim=3, N=(24, 20, 16)), 0, [1.0, 1.0, 1.0, 0.0, 0.25, 0.25, 0.25, 0.125, 0.2]),
        (make_cfg(dim=3, N=(20, 24, 18), periodic=5, upper=(1.0, 2.0, 1.5)), 0, [-1.0, 0.7, 0.3, 0.0, 0.5, 1.0, 0.7, 0.1, 0.45]),
        (make_cfg(dim=3, N=(16, 16, 24), mode=2), 1, [1.5, 0, 0, 0.25, 0.5, 0.5, 0.5, 0.1, 0.6]),
        (make_cfg(dim=2, N=(32, 24), periodic=1), 1, [1.0, 0, 0, 0.0, 0.5, 0.5, 0.0, 0.1, 0.3]),
        # overlap 2, fully periodic, uneven split, partial tiles: send boxes two cells thick, wrap-around partners
        (make_cfg(dim=3, N=(36, 22, 26), periodic=7, overlap=2), 0, [0.6, -1.0, 0.8, 0.0, 0.5, 0.5, 0.5, 0.1, 0.45]),
        # wide and tall rank boxes: the 128x4 tile with odd row strides, and z extents that take the pipelined host path
        (make_cfg(dim=3, N=(520, 12, 300), upper=(1.0, 0.05, 0.6)), 0, [1.0, -0.05, 0.4, 0.2, 0.5, 0.02, 0.3, 0.01, 0.2]),
        # split along z only: rank boxes with EVEN rows (16-byte strides) -> the tensor-map (TMA) staging form together with the
        # in-kernel pack, peers storing straight into the registered arrays (no x faces: no side cells)
        (make_cfg(dim=3, N=(256, 8, 130 * max(2, world)), upper=(1.0, 0.03, 0.5 * max(2, world))), 0, [0.3, -0.05, 1.0, 0.2, 0.5, 0.02, 0.3, 0.01, 0.2]),
    ]
    failures = 0
    modes = set()
    stagings = set()
    host_paths = set()
    for cfg, problem, params in cases:
        w = oracle_world(o, cfg, world)
        g = product_grid(cfg, rank, world, comm=comm)
        gv = g.leafGridView()
        for mode, registered in (("exact", False), ("separable", False), ("exact", True), ("separable", True)):
            if registered and cfg["dim"] != 3:
                continue
            ref = [w.fv_init(r, 0, problem, params) for r in range(world)]
            fv = dg.FiniteVolume(g, 0, problem, params, mode)
            c = fv.initial()
            cn = torch.empty_like(c)
            if registered:                    # peers store the halo cells straight into these two arrays (CUDA IPC)
                nreg = fv.register_fields(c, cn)
                if nreg != 2:
                    print(f"rank {rank}: register_fields declined ({nreg}) cfg={cfg['N']}", flush=True)
            ok = np.array_equal(c.cpu().numpy(), ref[rank])
            for _ in range(4):
                dt_ref = w.fv_step(0, problem, params, ref)
                fv.step(c, cn)
                c, cn = cn, c
                if not registered:            # registered runs are completed once, at the end (dgb_fv_fence): the hybrid form
                    dt = fv.last_dt()         # reads its x halo from the exchange buffer for 3 of the 4 steps
                    ok = ok and (dt == dt_ref if mode == "exact" else abs(dt - dt_ref) <= 1e-13 * dt_ref)
            fv.fence()
            got = c.cpu().numpy()
            scale = max(np.abs(x).max() for x in ref)
            if mode == "exact":
                ok = ok and np.array_equal(got, ref[rank])
            else:
                ok = ok and np.abs(got - ref[rank]).max() <= 1e-13 * scale           # tolerance: 1e-13 relative
            modes.add(fv.exchange_mode())
            stagings.add(fv.kernel_info()["staging"])
            if not ok:
                failures += 1
                print(f"rank {rank}: FV mismatch cfg={cfg['N']} mode={mode} registered={registered} exchange={fv.exchange_mode()} "
                      f"maxdiff={np.abs(got - ref[rank]).max()}", flush=True)
            if cfg["dim"] == 3 and not registered:
                # the host-buffer entry point (pipelined over z-slabs where the rank box is tall enough; unpack into the host result)
                h_in = torch.from_numpy(ref[rank].copy()).pin_memory()
                h_out = torch.full_like(h_in, float("nan")).pin_memory()
                for _ in range(2):
                    dt_ref = w.fv_step(0, problem, params, ref)
                    dt = fv.step_host(h_in, h_out)
                    h_in, h_out = h_out, h_in
                hgot = h_in.numpy()
                hok = (np.array_equal(hgot, ref[rank]) and dt == dt_ref) if mode == "exact" else \
                      (np.abs(hgot - ref[rank]).max() <= 1e-13 * scale and abs(dt - dt_ref) <= 1e-13 * dt_ref)
                host_paths.add(fv.host_path())
                if not hok:
                    failures += 1
                    print(f"rank {rank}: step_host mismatch cfg={cfg['N']} mode={mode} path={fv.host_path()} maxdiff={np.abs(hgot - ref[rank]).max()}", flush=True)
            fv.close()
        # communicate through the public call (NCCL transport), all interfaces, set and add
        layout = [1] + [0] * (cfg["dim"] - 1) + [1]
        m = dg.MultipleCodimMultipleGeomTypeMapper(gv, layout)
        rng = np.random.default_rng(99)
        for iftype in range(5):
            for direction in (0, 1):
                for op in ("set", "add", "min_nonneg", "set_nonneg"):
                    sizes = [w.mapper_size(r, 0, layout) for r in range(world)]
                    host = [[rng.standard_normal(sizes[r]), rng.standard_normal(sizes[r] * 2)] for r in range(world)]
                    dev = [torch.from_numpy(a.copy()).cuda() for a in host[rank]]
                    w.communicate(0, layout, iftype, direction, ("set", "add", "min_nonneg", "set_nonneg").index(op), host, [1, 2])
                    gv.communicate(m, dev, iftype, direction, op)
                    torch.cuda.synchronize()
                    for a in range(2):
                        if not np.array_equal(dev[a].cpu().numpy(), host[rank][a]):
                            failures += 1
                            print(f"rank {rank}: communicate mismatch cfg={cfg['N']} if={iftype} dir={direction} op={op}", flush=True)
        # data handle of variable size (fixedSize() == false) through the public call: both message rounds over NCCL
        from tests.golden_cases import var_data
        for iftype in range(5):
            for direction in (0, 1):
                vd = [var_data(w.mapper_size(r, 0, layout), 500 + 10 * r + iftype) for r in range(world)]
                ref = w.communicate_variable(0, layout, iftype, direction, [v[0] for v in vd], [v[1] for v in vd])[rank]
                index, roff, items = gv.communicateVariable(m, torch.from_numpy(vd[rank][0]).cuda(), torch.from_numpy(vd[rank][1]).cuda(), iftype, direction)
                torch.cuda.synchronize()
                if not (np.array_equal(index.cpu().numpy().view(np.uint32), ref[0]) and np.array_equal(np.diff(roff.cpu().numpy()), ref[1])
                        and np.array_equal(items.cpu().numpy(), ref[2])):
                    failures += 1
                    print(f"rank {rank}: variable-size communicate mismatch cfg={cfg['N']} if={iftype} dir={direction}", flush=True)
        # Dune::GlobalIndexSet of every codim (collective: all-gather of the counts, two exchanges)
        for codim in range(cfg["dim"] + 1):
            ref, n = w.global_index(0, codim)
            got, m = gv.globalIndexSet(codim)
            if m != n or not np.array_equal(got.cpu().numpy(), ref[rank]):
                failures += 1
                print(f"rank {rank}: global index mismatch cfg={cfg['N']} codim={codim} size {m} vs {n}", flush=True)
        w.close()
    t = torch.tensor([failures], dtype=torch.int64, device="cuda")
    dist.all_reduce(t)
    if rank == 0:
        print(f"multigpu_check world={world}: {'OK' if t.item() == 0 else 'FAILED'} ({int(t.item())} mismatches), "
              f"FV exchange modes on rank 0: {sorted(modes)}, staging forms: {sorted(stagings)}, host paths: {sorted(host_paths)}", flush=True)
    dist.barrier()
    comm.close()
    dist.destroy_process_group()
    sys.exit(0 if t.item() == 0 else 1)


if __name__ == "__main__":
    main()
