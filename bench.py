#!/usr/bin/env python
"""bench.py — leaf cells/s of the element+intersection finite-volume sweep (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--mode separable|exact] [--size n]

N=1: BASELINE configs[1] — YaspGrid<3> 512^3 equidistant, explicit upwind FV transport step (elements +
intersections + mapper + geometry + CFL min-reduction) on 1 B200. N>1 (launched by torchrun, one rank per GPU):
configs[2] — 1024^3, overlap 1, DefaultPartitioning over N ranks, communicate(InteriorBorder_All) halo exchange
over NVLink each step. A "step" is one full FV time step over the whole grid.

--impl reference times the reference's own CPU implementation (oracle/_ref: the unmodified dune-grid YaspGrid
headers, one thread per rank like its MPI model, all host cores) on THE SAME grid and step (`config` is identical in
both arms); --size overrides the cube edge of both arms for local experiments only.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PARAMS = [1.0, 1.0, 1.0, 0.0, 0.25, 0.25, 0.25, 0.125, 0.2]     # u=(1,1,1), b=0, shell centre/radii
METRIC = "leaf cells/sec (element+intersection FV sweep)"
BYTES_PER_CELL = 16                                               # read c (8) + write c_new (8), DESIGN.md
NVLINK_GBS = 900.0                                                # nominal per direction per GPU (B200_PROFILING.md)


def workload_config(n, n_gpus):
    """the `config` object of the JSON line: a function of the workload only, so that both arms print the same one"""
    per_gpu = n ** 3 * 8 / max(1, n_gpus) / 1e9
    if n_gpus == 1:
        w = (f"YaspGrid<3> {n}^3 equidistant, overlap 1: explicit first-order finite-volume transport sweep "
             "(elements+intersections, MCMG mapper, geometry, CFL min-reduction)")
    else:
        w = (f"YaspGrid<3> {n}^3 equidistant, overlap 1, DefaultPartitioning over the ranks of the arm: explicit first-order "
             "finite-volume transport sweep (elements+intersections, MCMG mapper, geometry, CFL min-reduction) + "
             "communicate(InteriorBorder_All_Interface, ForwardCommunication) halo exchange each step")
    return {"workload": w, "grid": [n, n, n], "cells": n ** 3, "problem": "u=(1,1,1), b=0, c0 = spherical shell (SURVEY §8d)",
            "bytes_per_cell": BYTES_PER_CELL,
            "l2": "inputs larger than L2 (two %.2f GB fields per GPU, ping-pong): no flush needed" % per_gpu}


def ncu_traffic(pattern, kernel_prefix):
    """DRAM bytes per launch of the dominant kernel from the newest committed `ncu --set full` summary (profiles/), or None"""
    try:
        import glob
        best = None
        for path in sorted(glob.glob(os.path.join(ROOT, "profiles", pattern))):
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
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "5",
                                          "-i", str(self.index)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if not self.proc:
            return out
        # a timed region shorter than the sampling period (20 steps of 0.35 ms) may have seen no sample: wait for the next one - the
        # clocks right after the region - and say so
        late = False
        t0 = time.time()
        while time.time() - t0 < 0.2:
            try:
                if sum(1 for _ in open(self.path)) > getattr(self, "skip", 0):
                    break
            except Exception:
                break
            late = True
            time.sleep(0.005)
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
            if late:
                out["note"] = "the timed region was shorter than the sampling period: first sample after it"
        return out


def bind_to_gpu_numa(local):
    """pin this process to the CPUs next to its GPU BEFORE any pinned host buffer is allocated, so that the buffers of the
    end-to-end path are NUMA-local to the GPU (first touch); best effort, reports what it did"""
    try:
        import pynvml
        pynvml.nvmlInit()
        index = local
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [x for x in vis.split(",") if x.strip() != ""]
            if local < len(ids) and ids[local].strip().isdigit():
                index = int(ids[local])
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        near = {i * 64 + b for i, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cpus = sorted(near & allowed)
        if not cpus:
            return {"bound": False, "why": "no allowed CPU next to the GPU", "allowed": len(allowed)}
        os.sched_setaffinity(0, cpus)
        return {"bound": True, "cpus": f"{cpus[0]}-{cpus[-1]}", "count": len(cpus)}
    except Exception as e:                                      # noqa: BLE001
        return {"bound": False, "why": str(e)[:80]}


def cpu_reference_run(n, P, steps, warmup):
    """the reference's CPU path (oracle/_ref, else the port) on an n^3 grid with P thread-ranks; returns cells/s.
    This function and run_reference are the ONLY places of bench.py that execute anything under oracle/."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle as orc
    kind = "ref" if orc.available("ref") else "port"
    o = orc.load(kind)
    w = o.world(orc.Spec(dim=3, N=(n, n, n)), P)
    with ThreadPoolExecutor(max_workers=P) as ex:                # the per-rank initialisation runs concurrently (world is only read)
        c = list(ex.map(lambda r: w.fv_init(r, 0, 0, PARAMS), range(P)))
    if warmup:
        w.fv_run(0, 0, PARAMS, c, warmup)
    secs, _ = w.fv_run(0, 0, PARAMS, c, steps)
    w.close()
    return n ** 3 * steps / secs, ("reference" if kind == "ref" else "port"), secs


def feasible_ranks(n, maxp):
    """largest power of two <= maxp (DefaultPartitioning accepts every power of two for a power-of-two cube)"""
    p = 1
    while p * 2 <= maxp:
        p *= 2
    return p


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_cores()
    n = args.size or (512 if args.gpus == 1 else 1024)
    P = feasible_ranks(n, min(cores, 64))
    t0 = time.time()
    steps = max(1, args.steps)
    val, kind, secs = cpu_reference_run(n, P, steps, min(args.warmup, 1))
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "cells/s", "n_gpus": args.gpus, "steps": steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * secs / steps, "higher_is_better": True,
            "scaling": "weak" if args.gpus == 1 else "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(n, args.gpus),
            "cpu_baseline": {"value": val, "unit": "cells/s", "cores": P, "kind": kind,
                             "sample": f"{steps} FV steps (after {min(args.warmup, 1)} warm-up) on the full {n}^3 YaspGrid, "
                                       f"{P} ranks (threads) of {cores} host cores, {secs:.1f} s"},
            "e2e": {"value": val, "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "ranks": P, "wall_s": time.time() - t0}
    print(json.dumps(line), flush=True)


def parity_check(dg, torch, dist, rank, world, comm, dims, mode):
    """multi-GPU parity inside the scaling run: 3 steps of the SAME public step (fused peer-memory exchange) on a smaller grid
    with the partition of the timed grid (FixedSizePartitioning with its torus; 256 x 64 x 64 cells per rank, so the wide
    128x4 tile and the odd row strides of the timed run are what is exercised), every rank compares ALL its stored cells
    (interior + refreshed overlap) with the CPU oracle's rank of the same number. Oracle = checker only."""
    import numpy as np
    from oracle import oracle as orc
    kind = "ref" if orc.available("ref") else "port"
    o = orc.load(kind)
    N = (256 * dims[0], 64 * dims[1], 64 * dims[2])
    upper = (1.0, 0.25 * dims[1] / dims[0], 0.25 * dims[2] / dims[0])        # cubic cells
    params = [1.0, 0.6, -0.8, 0.0, 0.3, upper[1] / 2, upper[2] / 2, 0.05, 0.2]
    w = o.world(orc.Spec(dim=3, N=N, upper=upper, partitioner=2, fixed_dims=dims), world)
    ref = [w.fv_init(r, 0, 0, params) for r in range(world)]
    g = dg.YaspGrid(dim=3, N=N, upper=upper, overlap=1, rank=rank, nranks=world, comm=comm, partitioner="fixed", fixed_dims=dims)
    fv = dg.FiniteVolume(g, 0, 0, params, mode)
    c = fv.initial()
    ok = bool(np.array_equal(c.cpu().numpy(), ref[rank]))
    cn = torch.empty_like(c)
    steps, dt_rel = 3, 0.0
    for _ in range(steps):
        dt_ref = w.fv_step(0, 0, params, ref)
        fv.step(c, cn)
        c, cn = cn, c
        dt_rel = max(dt_rel, abs(fv.last_dt() - dt_ref) / dt_ref)
    scale = max(float(np.abs(x).max()) for x in ref)
    err = float(np.abs(c.cpu().numpy() - ref[rank]).max()) / scale
    moved = bool(max(float(np.abs(x).max()) for x in ref) > 0)        # somewhere in the domain (a rank far from the shell holds zeros)
    tol = 0.0 if mode == "exact" else 1e-13
    ok = ok and err <= tol and dt_rel <= tol and moved
    info, xmode = fv.kernel_info(), fv.exchange_mode()
    fv.close()
    w.close()
    t = torch.tensor([err, dt_rel, 0.0 if ok else 1.0], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return {"grid": list(N), "partition": list(dims), "steps": steps, "max_rel": t[0].item(), "dt_rel": t[1].item(), "tol": tol,
            "ok": t[2].item() == 0.0, "oracle": "reference" if kind == "ref" else "port", "exchange": xmode,
            "kernel": f"{info['kind']} {info['tile'][0]}x{info['tile'][1]} staging {info['staging']} ({info['staging_bytes']} B)"}


def run_gpu(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    numa = bind_to_gpu_numa(local) if world > 1 else {"bound": False, "why": "single GPU: all cores stay available to the CPU baseline"}
    import torch
    import dune_grid_b200 as dg
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
    if world > 1 and not args.no_register:
        fv.register_fields(c, cn)               # collective: peers may then store halo cells straight into these two arrays

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
    fv.fence()                                  # registered arrays: completes the halo of the last step (inside the timed region)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = fv.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    kinfo, xmode = fv.kernel_info(), fv.exchange_mode()
    fv.last_dt()                                # raises if the fused exchange timed out anywhere in the run
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

    # the halo exchange on its own (N>1): (i) exposed cost inside the step = step time - time of the same kernel without
    # exchange; (ii) the same halo through the stand-alone dgb_communicate (pack, NCCL send/recv over NVLink, unpack)
    halo = None
    if dist is not None and not args.no_halo:
        hs = max(5, min(args.steps, 50))
        sent = g.lastBytesSent()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        barrier()
        ev[0].record()
        for _ in range(hs):
            fv.step_local(c, cn)
            c, cn = cn, c
        ev[1].record()
        barrier()
        m = dg.MultipleCodimMultipleGeomTypeMapper(gv, dg.mcmgElementLayout(3))
        for _ in range(3):
            gv.communicate(m, [c], dg.InteriorBorder_All_Interface, dg.ForwardCommunication, "set")
        barrier()
        ev[2].record()
        for _ in range(hs):
            gv.communicate(m, [c], dg.InteriorBorder_All_Interface, dg.ForwardCommunication, "set")
        ev[3].record()
        barrier()
        t = torch.tensor([ev[0].elapsed_time(ev[1]) / hs, ev[2].elapsed_time(ev[3]) / hs, float(sent)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        local_ms, comm_ms, max_sent = t[0].item(), t[1].item(), t[2].item()
        exposed = max(ms / args.steps - local_ms, 1e-6)
        halo = {"bytes_per_rank": int(max_sent), "note": "max over ranks of the bytes one rank sends per step (= receives)",
                "step_ms": ms / args.steps, "kernel_only_ms": local_ms, "exchange_ms": exposed,
                "nvlink_gbs": max_sent / (exposed * 1e-3) / 1e9, "frac_of_900": max_sent / (exposed * 1e-3) / 1e9 / NVLINK_GBS,
                "how": "exchange_ms = step time minus the time of the same step kernel without exchange (the fused kernel stores into "
                       "the peers' memory while it computes, so only this exposed part costs time); nvlink_gbs = bytes_per_rank / exchange_ms",
                "standalone": {"ms": comm_ms, "nvlink_gbs": max_sent / (comm_ms * 1e-3) / 1e9,
                               "frac_of_900": max_sent / (comm_ms * 1e-3) / 1e9 / NVLINK_GBS,
                               "how": "dgb_communicate(InteriorBorder_All, Forward) of the same field alone, nothing overlapped: pack kernel "
                                      "that stores into the peers' receive buffers over NVLink + flag, unpack kernel that waits for the "
                                      "flags (DGB_COMM_PEER=0: pack, grouped ncclSend/ncclRecv, unpack)"}}
        # the local steps left the overlap cells stale: refresh them before anything else uses the field
        gv.communicate(m, [c], dg.InteriorBorder_All_Interface, dg.ForwardCommunication, "set")
        barrier()

    # end to end through the reference-facing C-ABI call with HOST buffers (copies inside the timed region)
    e2e = None
    if not args.no_e2e:
        ncell = gv.size(0)
        h_in = torch.empty(ncell, dtype=torch.float64).pin_memory()
        h_out = torch.empty(ncell, dtype=torch.float64).pin_memory()
        h_in.copy_(c)
        ksteps = max(1, min(args.steps, args.e2e_steps))
        fv.step_host(h_in, h_out)                 # warm-up (allocates the staging buffers)
        h_in, h_out = h_out, h_in
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
               "d2h_bytes_per_step": ncell * 8 * world + 8 * world, "steps": ksteps,
               "pcie_gbs_per_gpu_per_direction": ncell * 8 * ksteps / secs / 1e9, "host_path": fv.host_path()}
        del h_in, h_out

    parity = None
    if dist is not None and not args.no_parity:
        parity = parity_check(dg, torch, dist, rank, world, comm, g.torusDims(), args.mode)

    if rank == 0:
        peak, peak_src = peaks()
        kernel_ms = ms / args.steps
        if world == 1 and n == 512 and args.mode == "separable":
            traffic = ncu_traffic("r*_fv_ring_ncu_full.json", "k_fv_ring")
        elif world > 1 and n == 1024 and args.mode == "separable":
            traffic = ncu_traffic("r*_fv_ring_stride513_ncu_full.json", "k_fv_ring")
        else:
            traffic = None
        achieved = BYTES_PER_CELL * (total_cells / world) / (kernel_ms * 1e-3) / 1e9
        cfg = workload_config(n, world)
        line = {"metric": METRIC, "value": value, "unit": "cells/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak" if world == 1 else "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
                "fv_mode": args.mode, "partition": list(g.torusDims()),
                "kernel": {"kind": kinfo["kind"], "tile": list(kinfo["tile"]), "staging": kinfo["staging"], "staging_bytes": kinfo["staging_bytes"],
                           "zchunk": kinfo["zchunk"], "exchange": xmode, "registered_fields": bool(world > 1 and not args.no_register)},
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": traffic[0] if traffic else None,
                             "traffic_source": (traffic[1] + " (ncu --set full capture of one k_fv_ring launch on a box of this shape, "
                                                "committed under profiles/; not re-measured in this run)") if traffic else None,
                             "algorithmic_bytes_per_launch": BYTES_PER_CELL * (total_cells / world), "peak_source": peak_src,
                             "note": "per-GPU algorithmic bytes (16 B x interior cells) / mean step time; one fused kernel per step"
                                     + ("" if world == 1 else " that also stores the halo into the peers' memory over NVLink (no NCCL call in the "
                                        "steady state) + a one-CTA flag wait (+ an unpack kernel when the fields are not registered)")},
                "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "numa": numa}
        if halo:
            line["halo"] = halo
        if parity:
            line["parity_check"] = parity
        if not args.no_cpu:
            cores = host_cores()
            P = feasible_ranks(n, min(cores, 64))
            csteps = args.cpu_steps or (8 if n <= 512 else 2)
            val, kind, secs = cpu_reference_run(n, P, csteps, 1 if n <= 512 else 0)
            line["cpu_baseline"] = {"value": val, "unit": "cells/s", "cores": P, "kind": kind,
                                    "sample": f"{csteps} FV steps on the full {n}^3 YaspGrid, {P} ranks (threads) of {cores} host cores, {secs:.1f} s"}
        print(json.dumps(line), flush=True)
    if dist is not None:
        # the other ranks wait for rank 0's CPU baseline in a socket wait (no spinning next to the CPU threads being timed)
        try:
            store = dist.distributed_c10d._get_default_store()
            if rank == 0:
                store.set("dgb_bench_done", "1")
            else:
                store.wait(["dgb_bench_done"])
        except Exception:                        # noqa: BLE001
            pass
        dist.barrier()
        fv.close()                               # collective teardown of the fused exchange: before the communicator goes away
        comm.close()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="separable", choices=["separable", "exact"])
    ap.add_argument("--size", type=int, default=0, help="override the cube edge (default 512 at N=1, 1024 at N>1) - local experiments only")
    ap.add_argument("--cpu-steps", type=int, default=0, help="steps of the cpu_baseline sample (default 8 at 512^3, 2 at 1024^3)")
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-halo", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-register", action="store_true", help="N>1: keep the receive-buffer + unpack form of the fused exchange")
    args = ap.parse_args()
    if args.impl == "reference":
        if args.steps == 1000 and args.warmup == 20:      # no flags: a CPU-sized default
            args.steps, args.warmup = 10, 1
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
