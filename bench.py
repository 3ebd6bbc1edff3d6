#!/usr/bin/env python
"""bench.py — leaf cells/s of the element+intersection finite-volume sweep (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--mode separable|exact] [--size n]

N=1: BASELINE configs[1] — YaspGrid<3> 512^3 equidistant, explicit upwind FV transport step (elements +
intersections + mapper + geometry + CFL min-reduction) on 1 B200. N>1 (launched by torchrun, one rank per GPU):
configs[2] — 1024^3, overlap 1, DefaultPartitioning over N ranks, communicate(InteriorBorder_All) halo exchange
over NVLink each step. A "step" is one full FV time step over the whole grid.

--impl reference times the reference's own CPU implementation (oracle/_ref: the unmodified dune-grid YaspGrid
headers, one thread per rank like its MPI model) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PARAMS = [1.0, 1.0, 1.0, 0.0, 0.25, 0.25, 0.25, 0.125, 0.2]     # u=(1,1,1), b=0, shell centre/radii
METRIC = "leaf cells/sec (element+intersection FV sweep)"
BYTES_PER_CELL = 16                                               # read c (8) + write c_new (8), DESIGN.md


def ncu_traffic(kernel_prefix):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` summary (profiles/), or None"""
    try:
        import glob
        best = None
        for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_fv_ring_ncu_full.json"))):
            for rec in json.load(open(path)):
                if kernel_prefix in rec.get("kernel", "") and "traffic_bytes" in rec:
                    best = (rec["traffic_bytes"], os.path.basename(path))
        return best
    except Exception:
        return None


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)"""

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        """start sampling and wait until the first sample has been written, so that the samples counted later
        (those after `self.skip`) lie inside the timed region"""
        self._start()
        self.skip = 0
        if not self.proc:
            return
        t0 = time.time()
        while time.time() - t0 < 2.0:
            try:
                self.skip = sum(1 for _ in open(self.path))
            except Exception:
                self.skip = 0
            if self.skip > 0:
                break
            time.sleep(0.02)

    def _start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
                 "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.index)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for ln, line in enumerate(open(self.path)):
                if ln < getattr(self, "skip", 0):
                    continue
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            sm.sort()
            out = {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}
        return out


def cpu_reference_run(n, P, steps, warmup):
    """the reference's CPU path (oracle/_ref, else the port) on an n^3 grid with P thread-ranks; returns cells/s"""
    from oracle import oracle as orc
    kind = "ref" if orc.available("ref") else "port"
    o = orc.load(kind)
    w = o.world(orc.Spec(dim=3, N=(n, n, n)), P)
    c = [w.fv_init(r, 0, 0, PARAMS) for r in range(P)]
    if warmup:
        w.fv_run(0, 0, PARAMS, c, warmup)
    secs, _ = w.fv_run(0, 0, PARAMS, c, steps)
    w.close()
    return n ** 3 * steps / secs, ("reference" if kind == "ref" else "port"), secs


def feasible_ranks(n, maxp):
    """largest rank count <= maxp that DefaultPartitioning accepts for an n^3 grid and that is a power of two"""
    p = 1
    while p * 2 <= maxp:
        p *= 2
    return p


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    P = feasible_ranks(args.ref_size, min(cores, 64))
    n = args.ref_size
    t0 = time.time()
    val, kind, secs = cpu_reference_run(n, P, max(1, args.steps), min(args.warmup, 1))
    workload = ("YaspGrid<3> 512^3 equidistant FV sweep" if args.gpus == 1 else "YaspGrid<3> 1024^3 overlap=1 FV sweep + halo exchange")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "cells/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * secs / max(1, args.steps), "higher_is_better": True,
            "scaling": "weak" if args.gpus == 1 else "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload, "sample": f"{n}^3 cells per step on {P} thread-ranks"},
            "cpu_baseline": {"value": val, "unit": "cells/s", "cores": P, "kind": kind,
                             "sample": f"{max(1, args.steps)} FV steps on a {n}^3 YaspGrid, {P} ranks (threads) of {cores} host cores"},
            "e2e": {"value": val, "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": time.time() - t0}
    print(json.dumps(line), flush=True)


def run_gpu(args):
    import torch
    import dune_grid_b200 as dg
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    comm = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        comm = dg.Communicator.from_torch_distributed()
    n = args.size or (512 if world == 1 else 1024)
    g = dg.YaspGrid(dim=3, N=(n, n, n), upper=(1.0, 1.0, 1.0), overlap=1, rank=rank, nranks=world, comm=comm)
    fv = dg.FiniteVolume(g, 0, 0, PARAMS, args.mode)
    gv = g.leafGridView()
    interior = gv.size(0, dg.Interior_Partition)
    c = fv.initial()
    cn = torch.empty_like(c)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        fv.step(c, cn)
        c, cn = cn, c
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()                                   # the other ranks must not enter the timed region while rank 0 starts nvidia-smi
    l0 = fv.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        fv.step(c, cn)
        c, cn = cn, c
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = fv.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
        tot = torch.tensor([float(interior)], dtype=torch.float64, device="cuda")
        dist.all_reduce(tot)
        total_cells = tot.item()
    else:
        total_cells = float(interior)
    assert abs(total_cells - n ** 3) < 0.5
    value = total_cells * args.steps / (ms * 1e-3)

    # end to end through the reference-facing C-ABI call with HOST buffers (copies inside the timed region)
    e2e = None
    if not args.no_e2e:
        ncell = gv.size(0)
        h_in = torch.empty(ncell, dtype=torch.float64).pin_memory()
        h_out = torch.empty(ncell, dtype=torch.float64).pin_memory()
        h_in.copy_(c)
        ksteps = max(1, min(args.steps, args.e2e_steps))
        fv.step_host(h_in, h_out)                 # warm-up (allocates the staging buffers)
        barrier()
        t0 = time.perf_counter()
        for _ in range(ksteps):
            fv.step_host(h_in, h_out)
            h_in, h_out = h_out, h_in
        barrier()
        secs = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([secs], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            secs = t.item()
        e2e = {"value": total_cells * ksteps / secs, "unit": "cells/s", "h2d_bytes_per_step": ncell * 8 * world,
               "d2h_bytes_per_step": ncell * 8 * world + 8 * world, "steps": ksteps}

    if rank == 0:
        peak, peak_src = peaks()
        kernel_ms = ms / args.steps
        traffic = ncu_traffic("k_fv_ring") if (world == 1 and n == 512 and args.mode == "separable") else None
        achieved = BYTES_PER_CELL * (total_cells / world) / (kernel_ms * 1e-3) / 1e9
        line = {"metric": METRIC, "value": value, "unit": "cells/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak" if world == 1 else "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": (f"YaspGrid<3> {n}^3 equidistant: explicit first-order FV transport sweep "
                                        "(elements+intersections, CFL min-reduction)" + ("" if world == 1 else
                                        f", overlap=1, DefaultPartitioning {g.torusDims()}, communicate(InteriorBorder_All) each step")),
                           "fv_mode": args.mode, "l2": "inputs (2 x %.2f GB fields per GPU) larger than L2; no flush needed" % (gv.size(0) * 8 / 1e9),
                           "bytes_per_cell": BYTES_PER_CELL},
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": traffic[0] if traffic else None,
                             "traffic_source": (traffic[1] + " (ncu --set full, 512^3 launch of k_fv_ring)") if traffic else None,
                             "algorithmic_bytes_per_launch": BYTES_PER_CELL * (total_cells / world), "peak_source": peak_src,
                             "note": "per-GPU algorithmic bytes (16 B x interior cells) / mean step time (one fused kernel per step"
                                     + ("" if world == 1 else " + pack/unpack + NCCL") + ")"},
                "e2e": e2e, "gpu_launches": launches, "clocks": clocks}
        if world == 1 and not args.no_cpu:
            cores = os.cpu_count() or 1
            P = feasible_ranks(args.ref_size, min(cores, 64))
            val, kind, secs = cpu_reference_run(args.ref_size, P, 24, 1)
            line["cpu_baseline"] = {"value": val, "unit": "cells/s", "cores": P, "kind": kind,
                                    "sample": f"24 FV steps on a {args.ref_size}^3 YaspGrid, {P} ranks (threads) of {cores} host cores, {secs:.1f} s"}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        comm.close()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="separable", choices=["separable", "exact"])
    ap.add_argument("--size", type=int, default=0, help="override the cube edge (default 512 at N=1, 1024 at N>1)")
    ap.add_argument("--ref-size", type=int, default=256, help="cube edge of the bounded CPU sample")
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
