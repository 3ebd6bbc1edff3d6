"""GPU parity of communicate (row a10) and of the finite-volume step (row a11) against the CPU oracle.

Several ranks are emulated on ONE device: every rank has its own grid object / fields, the pack and unpack
kernels run per rank through dgb_comm_pack / dgb_comm_unpack, and the test moves the packed segments between the
ranks' exchange buffers (what NCCL send/recv does on a multi-GPU box; see tests/test_multigpu.py for the real
transport). This follows B200_PROFILING.md: never run ranks that wait on each other as separate processes on one GPU."""
import ctypes

import numpy as np
import pytest

from tests.common import best_oracle, make_cfg, oracle_world, product_grid

pytestmark = pytest.mark.gpu


def np_(t):
    return t.cpu().numpy()


def dev_view(ptr, count, typestr="<f8"):
    """torch view of `count` doubles (or int64: typestr "<i8") of library-owned device memory"""
    import torch

    class _Holder:
        pass

    h = _Holder()
    h.__cuda_array_interface__ = {"shape": (count,), "typestr": typestr, "data": (ptr, False), "version": 2}
    return torch.as_tensor(h, device="cuda")


def emulated_exchange(grids, views, mappers, arrays_per_rank, item_sizes, iftype, direction, op):
    """pack on every rank, route segments, unpack on every rank"""
    import torch
    P = len(grids)
    plans, bufs = [], []
    for r in range(P):
        plans.append(views[r].commPlan(mappers[r], iftype, direction, item_sizes))
        bufs.append(views[r].commPack(mappers[r], arrays_per_rank[r], iftype, direction, item_sizes))
    torch.cuda.synchronize()
    sent = 0
    for r in range(P):
        send_segs, _ = plans[r]
        for (peer, off, cnt) in send_segs:
            # the k-th message r -> peer pairs with the k-th message peer <- r; with one segment per peer, k = 0
            rsegs = [s for s in plans[peer][1] if s[0] == r]
            assert len(rsegs) == 1 and rsegs[0][2] == cnt, (r, peer, cnt, rsegs)
            src = dev_view(bufs[r][0] + 8 * off, cnt)
            dst = dev_view(bufs[peer][1] + 8 * rsegs[0][1], cnt)
            dst.copy_(src)
            sent += cnt
    torch.cuda.synchronize()
    for r in range(P):
        views[r].commUnpack(mappers[r], arrays_per_rank[r], iftype, direction, op, item_sizes)
    torch.cuda.synchronize()
    return sent


def emulated_variable_exchange(views, mappers, offsets, data, iftype, direction):
    """variable-size data handle on ranks emulated on one device: dgb_commvar_create per rank, the test moves the size
    segments, dgb_commvar_sizes_received, the test moves the item segments, finish. Returns per rank (index, n, items)."""
    import torch
    import dune_grid_b200 as dg
    P = len(views)
    ex = [dg.VariableExchange(views[r], mappers[r], offsets[r], data[r], iftype, direction) for r in range(P)]
    torch.cuda.synchronize()

    def route(col_first, col_count, send_buf, recv_buf, typestr):
        ssegs = [e.segments(False) for e in ex]
        rsegs = [e.segments(True) for e in ex]
        bufs = [e.buffers() for e in ex]
        for r in range(P):
            for sg in ssegs[r]:
                peer, cnt = sg[0], sg[col_count]
                back = [q for q in rsegs[peer] if q[0] == r]
                assert len(back) == 1 and back[0][col_count] == cnt, (r, peer, sg, back)
                if cnt:
                    dev_view(bufs[peer][recv_buf] + 8 * back[0][col_first], cnt, typestr).copy_(dev_view(bufs[r][send_buf] + 8 * sg[col_first], cnt, typestr))
        torch.cuda.synchronize()

    route(1, 2, 0, 2, "<i8")                   # round 1: per-entity sizes (yaspgrid.hh:1566-1612)
    for e in ex:
        e.sizesReceived()
    route(3, 4, 1, 3, "<f8")                   # round 2: the items (:1637-1681)
    out = []
    for e in ex:
        index, off, items = e.result()
        torch.cuda.synchronize()
        off = np_(off)
        out.append((np_(index).view(np.uint32), np.diff(off), np_(items)))
        e.close()
    return out


VAR_CASES = [
    (make_cfg(dim=2, N=(8, 4)), 2, [1, 0, 0]), (make_cfg(dim=3, N=(8, 6, 4)), 4, [1, 1, 1, 1]), (make_cfg(dim=2, N=(9, 8), periodic=3), 1, [1, 0, 1]),
    (make_cfg(dim=3, N=(8, 8, 8), periodic=5, overlap=2), 8, [1, 0, 0, 1]), (make_cfg(dim=2, N=(12, 10), periodic=1, refine=1), 6, [0, 1, 1]),
    (make_cfg(dim=3, N=(6, 5, 4), periodic=7, mode=2), 1, [0, 1, 0, 0]),
]


@pytest.mark.parametrize("cfg,P,layout", VAR_CASES, ids=[f"var{i}" for i in range(len(VAR_CASES))])
def test_communicate_variable_size_handle(cfg, P, layout):
    """CommDataHandleIF with fixedSize() == false (yaspgrid.hh:1562-1634; what the reference's own checkCommunication uses,
    test/checkcommunicate.hh:96-102): every interface, both directions; the sequence of scatter(buff, e, n) calls - entity,
    n, items - must equal the reference's bit for bit, in its order."""
    import torch
    import dune_grid_b200 as dg
    from tests.golden_cases import var_data
    o = best_oracle()
    w = oracle_world(o, cfg, P)
    lvl = cfg["refine"]
    grids = [product_grid(cfg, r, P) for r in range(P)]
    views = [g.levelGridView(lvl) for g in grids]
    mappers = [dg.MultipleCodimMultipleGeomTypeMapper(v, layout) for v in views]
    for iftype in range(5):
        for direction in (0, 1):
            vd = [var_data(mappers[r].size, 31 * iftype + 7 * direction + r) for r in range(P)]
            ref = w.communicate_variable(lvl, layout, iftype, direction, [v[0] for v in vd], [v[1] for v in vd])
            got = emulated_variable_exchange(views, mappers, [torch.from_numpy(v[0]).cuda() for v in vd], [torch.from_numpy(v[1]).cuda() for v in vd],
                                             iftype, direction)
            for r in range(P):
                for q, a, b in zip(("index", "n", "items"), got[r], ref[r]):
                    assert np.array_equal(a, b), (iftype, direction, r, q)
    w.close()
    # a mapper layout with blocks > 1 is refused
    m2 = dg.MultipleCodimMultipleGeomTypeMapper(views[0], [2] + [0] * cfg["dim"])
    with pytest.raises(dg.capi.DgbError):
        dg.VariableExchange(views[0], m2, torch.zeros(m2.size + 1, dtype=torch.int64, device="cuda"), torch.zeros(1, dtype=torch.float64, device="cuda"), 4, 0)


def test_communicate_variable_single_rank_periodic_through_public_call():
    """one rank, fully periodic: the self messages travel as device copies inside dgb_commvar_exchange (no communicator)"""
    import torch
    import dune_grid_b200 as dg
    from tests.golden_cases import var_data
    cfg = make_cfg(dim=3, N=(7, 6, 5), periodic=7)
    o = best_oracle()
    w = oracle_world(o, cfg, 1)
    gv = product_grid(cfg, 0, 1).leafGridView()
    layout = [1, 0, 1, 1]
    m = dg.MultipleCodimMultipleGeomTypeMapper(gv, layout)
    off, data = var_data(m.size, 99)
    for iftype in (0, 1, 4):
        ref = w.communicate_variable(0, layout, iftype, 0, [off], [data])[0]
        index, roff, items = gv.communicateVariable(m, torch.from_numpy(off).cuda(), torch.from_numpy(data).cuda(), iftype, dg.ForwardCommunication)
        torch.cuda.synchronize()
        assert np.array_equal(np_(index).view(np.uint32), ref[0]) and np.array_equal(np.diff(np_(roff)), ref[1]) and np.array_equal(np_(items), ref[2])
    w.close()


COMM_CASES = [
    (make_cfg(dim=2, N=(8, 4)), 2, [1, 0, 0]),
    (make_cfg(dim=3, N=(8, 4, 4)), 2, [1, 0, 0, 0]),
    (make_cfg(dim=3, N=(8, 6, 4)), 4, [1, 1, 1, 1]),
    (make_cfg(dim=2, N=(9, 8), periodic=3), 1, [1, 0, 1]),
    (make_cfg(dim=3, N=(6, 5, 4), periodic=7, mode=2), 1, [1, 0, 0, 0]),
    (make_cfg(dim=3, N=(8, 8, 8), periodic=5, overlap=2), 8, [2, 0, 0, 1]),
    (make_cfg(dim=2, N=(12, 10), periodic=1, refine=1), 6, [0, 2, 1]),
    (make_cfg(dim=3, N=(8, 8, 6), periodic=7), 4, [1, 0, 0, 1]),
]


@pytest.mark.parametrize("cfg,P,layout", COMM_CASES, ids=[f"comm{i}" for i in range(len(COMM_CASES))])
def test_communicate_all_interfaces(cfg, P, layout):
    """checkcomcorrectness.hh-style sweep: every interface, both directions, set and add, two arrays with different
    item sizes, seeded random data; result must equal the oracle's communicate bit for bit."""
    import torch
    import dune_grid_b200 as dg
    o = best_oracle()
    w = oracle_world(o, cfg, P)
    lvl = cfg["refine"]
    grids = [product_grid(cfg, r, P) for r in range(P)]
    views = [g.levelGridView(lvl) for g in grids]
    mappers = [dg.MultipleCodimMultipleGeomTypeMapper(v, layout) for v in views]
    items = [1, 3]
    rng = np.random.default_rng(20261019)
    for iftype in range(5):
        for direction in (0, 1):
            # set / add (dune/python/grid/mapper.hh commOp) and the two combine rules of utility/globalindexset.hh
            for op in ("set", "add", "min_nonneg", "set_nonneg"):
                host = [[rng.standard_normal(mappers[r].size * it) for it in items] for r in range(P)]
                dev = [[torch.from_numpy(a.copy()).cuda() for a in host[r]] for r in range(P)]
                n_ref = w.communicate(lvl, layout, iftype, direction, ("set", "add", "min_nonneg", "set_nonneg").index(op), host, items)
                n_got = emulated_exchange(grids, views, mappers, dev, items, iftype, direction, op)
                assert n_ref == n_got, (iftype, direction, op, n_ref, n_got)
                for r in range(P):
                    for a in range(len(items)):
                        assert np.array_equal(np_(dev[r][a]), host[r][a]), (iftype, direction, op, r, a)
    w.close()


def test_communicate_single_rank_periodic_through_public_call():
    """one rank, periodic: the 3^d-1 self messages of torus.hh:395-403 become one device-local exchange inside
    dgb_communicate (no communicator needed)"""
    import torch
    import dune_grid_b200 as dg
    cfg = make_cfg(dim=3, N=(7, 6, 5), periodic=7)
    o = best_oracle()
    w = oracle_world(o, cfg, 1)
    g = product_grid(cfg, 0, 1)
    gv = g.leafGridView()
    layout = [1, 0, 0, 1]
    m = dg.MultipleCodimMultipleGeomTypeMapper(gv, layout)
    rng = np.random.default_rng(7)
    host = [[rng.standard_normal(m.size)]]
    dev = torch.from_numpy(host[0][0].copy()).cuda()
    w.communicate(0, layout, dg.InteriorBorder_All_Interface, dg.ForwardCommunication, 0, host, [1])
    gv.communicate(m, [dev], dg.InteriorBorder_All_Interface, dg.ForwardCommunication, "set")
    torch.cuda.synchronize()
    assert np.array_equal(np_(dev), host[0][0])
    w.close()


PARAMS0 = [1.0, 1.0, 1.0, 0.0, 0.25, 0.25, 0.25, 0.125, 0.2]

FV_CASES = [
    (make_cfg(dim=3, N=(16, 16, 16)), 1, 0, PARAMS0),
    (make_cfg(dim=2, N=(32, 24)), 1, 0, [1.0, 0.5, 0.0, 0.3, 0.25, 0.25, 0.0, 0.125, 0.2]),
    (make_cfg(dim=3, N=(12, 10, 14), upper=(0.3, 0.7, 1.9)), 1, 1, [2.0, 0, 0, 0.5, 0.15, 0.35, 0.9, 0.05, 0.4]),
    (make_cfg(dim=3, N=(10, 12, 9), mode=2), 1, 0, [1.0, -0.5, 0.25, 0.1, 0.5, 0.5, 0.5, 0.1, 0.45]),
    (make_cfg(dim=3, N=(10, 9, 8), periodic=7), 1, 0, [-1.0, 0.7, 0.3, 0.0, 0.5, 0.5, 0.5, 0.1, 0.45]),
    (make_cfg(dim=3, N=(16, 12, 8)), 2, 0, PARAMS0),
    (make_cfg(dim=3, N=(12, 12, 12), periodic=3, mode=1, lower=(-1.0, 0.0, 0.5), upper=(1.0, 1.0, 2.0)), 8, 1,
     [1.5, 0, 0, 0.25, 0.0, 0.5, 1.2, 0.1, 0.6]),
    (make_cfg(dim=2, N=(16, 16), refine=1), 4, 1, [1.0, 0, 0, 0.0, 0.5, 0.5, 0.0, 0.1, 0.3]),
    # problem 2 of the FV registry: shear flow whose velocity depends on z - not column invariant, runs on the general kernel
    (make_cfg(dim=3, N=(14, 12, 16), periodic=2), 1, 2, [0.9, -0.4, 0.7, 0.3, 0.5, 0.5, 0.5, 0.1, 0.45]),
    (make_cfg(dim=3, N=(16, 8, 12)), 2, 2, [-1.1, 0.5, 0.8, 0.2, 0.4, 0.5, 0.6, 0.1, 0.4]),
    (make_cfg(dim=2, N=(24, 18)), 1, 2, [0.9, -0.4, 0.7, 0.3, 0.5, 0.5, 0.0, 0.1, 0.45]),
]


def run_fv(cfg, P, problem, params, mode, steps, want_info=False):
    """product FV over `steps` steps with emulated ranks; returns per-rank fields (numpy), dts and first update
    (+ the kernel_info of every rank's last step with want_info)"""
    import torch
    import dune_grid_b200 as dg
    lvl = cfg["refine"]
    grids = [product_grid(cfg, r, P) for r in range(P)]
    # a communicator object is required by dgb_fv_create on multi-rank grids; the emulation never uses it
    fvs = []
    for g in grids:
        if P > 1:
            g.comm = _FakeComm()
        fvs.append(dg.FiniteVolume(g, lvl, problem, params, mode))
    views = [g.levelGridView(lvl) for g in grids]
    mappers = [dg.MultipleCodimMultipleGeomTypeMapper(v, dg.mcmgElementLayout(cfg["dim"])) for v in views]
    c = [fv.initial() for fv in fvs]
    cn = [torch.empty_like(x) for x in c]
    upd0 = [np_(fv_update_local(fvs, r, c)) for r in range(P)] if P == 1 else None

    def reduce_pending(ptrs):
        vals = [dev_view(p, 1) for p in ptrs]
        m = torch.stack(vals).max()
        for v in vals:
            v.fill_(m.item())

    reduce_pending([fv.bootstrap_local() for fv in fvs])
    dts = []
    for _ in range(steps):
        ptrs = [fvs[r].step_local(c[r], cn[r]) for r in range(P)]
        torch.cuda.synchronize()
        reduce_pending(ptrs)
        emulated_exchange(grids, views, mappers, [[x] for x in cn], [1], dg.InteriorBorder_All_Interface,
                          dg.ForwardCommunication, "set")
        dts.append(fvs[0].last_dt())
        c, cn = cn, c
    if want_info:
        return [np_(x) for x in c], dts, upd0, [fv.kernel_info() for fv in fvs]
    return [np_(x) for x in c], dts, upd0


class _FakeComm:
    h = ctypes.c_void_p(1)       # never dereferenced: the emulation only calls the *_local entry points


def fv_update_local(fvs, r, c):
    return fvs[r].update(c[r])


@pytest.mark.parametrize("cfg,P,problem,params", FV_CASES, ids=[f"fv{i}" for i in range(len(FV_CASES))])
def test_fv_exact_mode_is_bit_identical(cfg, P, problem, params):
    o = best_oracle()
    w = oracle_world(o, cfg, P)
    lvl = cfg["refine"]
    steps = 4
    ref = [w.fv_init(r, lvl, problem, params) for r in range(P)]
    init = [x.copy() for x in ref]
    ref_dts, ref_upd0 = [], None
    for s in range(steps):
        if s == 0:
            dt, upd = w.fv_step(lvl, problem, params, ref, want_update=True)
            ref_upd0 = upd
        else:
            dt = w.fv_step(lvl, problem, params, ref)
        ref_dts.append(dt)
    got, dts, upd0 = run_fv(cfg, P, problem, params, "exact", steps)
    assert dts == ref_dts
    if upd0 is not None:
        assert np.array_equal(upd0[0], ref_upd0[0])
    for r in range(P):
        assert np.array_equal(got[r], ref[r]), (r, np.abs(got[r] - ref[r]).max())
    assert any(not np.array_equal(init[r], ref[r]) for r in range(P))          # the field actually moved
    w.close()


@pytest.mark.parametrize("cfg,P,problem,params", FV_CASES, ids=[f"fv{i}" for i in range(len(FV_CASES))])
def test_fv_separable_mode_within_1e13(cfg, P, problem, params):
    """tolerance: |c_gpu - c_oracle| <= 1e-13 * max|c_oracle| after 4 steps, dt to 1e-13 relative"""
    o = best_oracle()
    w = oracle_world(o, cfg, P)
    lvl = cfg["refine"]
    steps = 4
    ref = [w.fv_init(r, lvl, problem, params) for r in range(P)]
    ref_dts = [w.fv_step(lvl, problem, params, ref) for _ in range(steps)]
    got, dts, _, infos = run_fv(cfg, P, problem, params, "separable", steps, want_info=True)
    if cfg["dim"] == 3:         # every registered problem runs on the ring kernel: column-invariant ones with hoisted x/y face factors,
        assert all(i["kind"] == "ring" for i in infos), infos      # the others (problem 2) in its general mode (velocity functor per plane)
    for a, b in zip(dts, ref_dts):
        assert abs(a - b) <= 1e-13 * abs(b)
    scale = max(np.abs(x).max() for x in ref)
    for r in range(P):
        assert np.abs(got[r] - ref[r]).max() <= 1e-13 * scale
    w.close()


def test_fv_public_step_and_host_entry_single_rank():
    """dgb_fv_step (device buffers) and dgb_fv_step_host (host buffers, copies inside) on one rank"""
    import torch
    import dune_grid_b200 as dg
    cfg = make_cfg(dim=3, N=(20, 16, 12))
    o = best_oracle()
    w = oracle_world(o, cfg, 1)
    ref = [w.fv_init(0, 0, 0, PARAMS0)]
    g = product_grid(cfg, 0, 1)
    fv = dg.FiniteVolume(g, 0, 0, PARAMS0, "exact")
    c = fv.initial()
    assert np.array_equal(np_(c), ref[0])
    cn = torch.empty_like(c)
    for _ in range(3):
        dt_ref = w.fv_step(0, 0, PARAMS0, ref)
        fv.step(c, cn)
        assert fv.last_dt() == dt_ref
        c, cn = cn, c
    assert np.array_equal(np_(c), ref[0])
    h_in = torch.from_numpy(ref[0].copy()).pin_memory()
    h_out = torch.empty_like(h_in).pin_memory()
    dt = fv.step_host(h_in, h_out)
    dt_ref = w.fv_step(0, 0, PARAMS0, ref)
    assert dt == dt_ref and np.array_equal(h_out.numpy(), ref[0])
    assert fv.launch_count() > 0
    w.close()


def test_fv_production_kernel_against_oracle_at_baseline_size():
    """BASELINE config 2 at full size (512^3 on one GPU; 256^3 if the device lacks memory): the kernels `bench.py` times
    - k_fv_ring, 128x4 tile, 16-byte staging, selected because the region is >= 256 cells wide - compared DIRECTLY with the
    reference oracle on the whole field after 3 steps: exact mode bit for bit, separable mode (production) within
    1e-13 relative (tolerance of north_star), dt likewise. The oracle runs the same grid on 8 thread-ranks (the FV result
    does not depend on the partition); plus the size-independent properties: dt = 0.99*h/3 exactly for u=(1,1,1) on the
    2^-k grid, maximum principle, discrete mass balance."""
    import torch
    import dune_grid_b200 as dg
    from tests.common import assemble_global, fv_init_all
    free, _ = torch.cuda.mem_get_info()
    n = 512 if free > 8 * 512 ** 3 * 8 else 256
    cfg = make_cfg(dim=3, N=(n, n, n))
    P, steps = 8, 3
    o = best_oracle()
    w = oracle_world(o, cfg, P)
    ref = fv_init_all(w, P, 0, 0, PARAMS0)
    c0_ref = assemble_global(w, cfg, P, 0, ref)
    _, dt_ref = w.fv_run(0, 0, PARAMS0, ref, steps)
    c_ref = assemble_global(w, cfg, P, 0, ref)
    w.close()
    del ref
    g = product_grid(cfg, 0, 1)
    for mode in ("separable", "exact"):
        fv = dg.FiniteVolume(g, 0, 0, PARAMS0, mode)
        c = fv.initial()
        assert np.array_equal(np_(c), c0_ref)
        cn = torch.empty_like(c)
        m0 = c.sum().item()
        for _ in range(steps):
            fv.step(c, cn)
            c, cn = cn, c
        info = fv.kernel_info()
        if mode == "separable":       # the production selection, not a small-grid variant
            assert info["kind"] == "ring" and info["tile"] == (128, 4) and info["staging"] == "tma", info
        else:
            assert info["kind"] == "march", info
        dt = fv.last_dt()
        assert dt == 0.99 * (1.0 / (3.0 * n))
        assert dt == dt_ref if mode == "exact" else abs(dt - dt_ref) <= 1e-13 * dt_ref
        assert c.min().item() >= 0.0 and c.max().item() <= 1.0
        assert abs(c.sum().item() - m0) <= 1e-9 * m0
        got = np_(c)
        if mode == "exact":
            assert np.array_equal(got, c_ref)
        else:
            assert np.abs(got - c_ref).max() <= 1e-13 * np.abs(c_ref).max()          # tolerance: 1e-13 relative
        del fv, c, cn, got


ODD_STRIDE_PARAMS = [1.0, -0.5, 0.25, 0.1, 0.5, 0.5, 0.5, 0.1, 0.45]


@pytest.mark.parametrize("mode", ["exact", "separable"])
def test_fv_wide_rank_boxes_with_odd_row_stride_against_oracle(mode):
    """the rank-local boxes of a partitioned grid (BASELINE config 3: 513 cells per row) have odd row strides and interior
    regions that start at an odd offset: rows are only 8-byte aligned. Here 512 x 48 x 40 on 2 emulated ranks (257-wide
    stored boxes, 256-wide regions -> the wide 128x4 tile is selected) against the oracle, mixed flow directions, b != 0"""
    cfg = make_cfg(dim=3, N=(512, 48, 40))
    o = best_oracle()
    w = oracle_world(o, cfg, 2)
    steps = 4
    ref = [w.fv_init(r, 0, 0, ODD_STRIDE_PARAMS) for r in range(2)]
    ref_dts = [w.fv_step(0, 0, ODD_STRIDE_PARAMS, ref) for _ in range(steps)]
    got, dts, _, infos = run_fv(cfg, 2, 0, ODD_STRIDE_PARAMS, mode, steps, want_info=True)
    if mode == "separable":
        for info in infos:
            assert info["kind"] == "ring" and info["tile"] == (128, 4), info
    scale = max(np.abs(x).max() for x in ref)
    for a, b in zip(dts, ref_dts):
        assert a == b if mode == "exact" else abs(a - b) <= 1e-13 * abs(b)
    for r in range(2):
        if mode == "exact":
            assert np.array_equal(got[r], ref[r])
        else:
            assert np.abs(got[r] - ref[r]).max() <= 1e-13 * scale                       # tolerance: 1e-13 relative
    w.close()


@pytest.mark.parametrize("N,periodic", [((511, 24, 40), 7), ((510, 21, 37), 5), ((255, 40, 33), 3), ((510, 10, 70), 5), ((256, 9, 66), 7)])
def test_fv_fused_exchange_on_wide_odd_boxes_against_oracle(N, periodic):
    """the fused step (in-kernel pack + peer stores + flags) on single-rank periodic grids whose stored boxes are 513 / 512
    / 257 cells wide with the interior starting at offset 1 - the shapes of the 2x2x2, 4x1x1 and 2x1x1 rank boxes of
    config 3 - and on 512- / 256-wide boxes of at least 64 planes (tensor-map staging together with the pack) in production
    (separable) mode against the oracle within 1e-13, and in exact mode bit for bit"""
    import torch
    import dune_grid_b200 as dg
    cfg = make_cfg(dim=3, N=N, periodic=periodic, upper=(1.0, 0.1, 0.1))
    o = best_oracle()
    w = oracle_world(o, cfg, 1)
    params = [1.0, -0.07, 0.03, 0.2, 0.5, 0.05, 0.05, 0.02, 0.3]
    for mode in ("separable", "exact"):
        ref = [w.fv_init(0, 0, 0, params)]
        g = product_grid(cfg, 0, 1)
        fv = dg.FiniteVolume(g, 0, 0, params, mode)
        c = fv.initial()
        cn = torch.empty_like(c)
        for _ in range(4):
            dt_ref = w.fv_step(0, 0, params, ref)
            fv.step(c, cn)
            c, cn = cn, c
            dt = fv.last_dt()
            assert dt == dt_ref if mode == "exact" else abs(dt - dt_ref) <= 1e-13 * dt_ref
        assert fv.exchange_mode() == "fused"
        info = fv.kernel_info()
        assert info["packs"]
        if mode == "separable" and N[0] >= 256:
            assert info["kind"] == "ring" and info["tile"] == (128, 4), info
        if mode == "separable" and N[0] >= 256 and N[0] % 2 == 0 and N[2] >= 64:        # even rows, tall enough: tensor-map staging + pack
            assert info["staging"] == "tma", info
        if mode == "exact":
            assert np.array_equal(np_(c), ref[0])
        else:
            assert np.abs(np_(c) - ref[0]).max() <= 1e-13 * np.abs(ref[0]).max()         # tolerance: 1e-13 relative
        fv.close()
    w.close()


@pytest.mark.parametrize("problem,params", [(0, [1.0, -0.5, 0.25, 0.1, 0.5, 0.5, 0.5, 0.1, 0.45]), (0, [-0.3, 0.0, 1.0, 0.7, 0.5, 0.5, 0.5, 0.1, 0.45]),
                                            (1, [2.0, 0, 0, 0.5, 0.15, 0.35, 0.9, 0.05, 0.4])])
def test_fv_ring_kernels_equal_general_kernel_bitwise(problem, params, monkeypatch):
    """the production ring kernel - every tile shape, and all three staging forms: phase-aware 16-byte pairs (default),
    aligned pairs, 8 bytes per cell - with partial tiles, mixed flow directions, boundary inflow b != 0, graded and periodic
    coordinates, odd and even row / plane strides, fields that are only 8-byte aligned, chunks that end at the array's last
    element, against the general z-march kernel on a random field: same arithmetic, same bits"""
    import torch
    import dune_grid_b200 as dg
    cfgs = (make_cfg(dim=3, N=(150, 37, 70), mode=2),                                        # even rows, odd offsets nowhere
            make_cfg(dim=3, N=(70, 45, 33), periodic=3, upper=(1.0, 0.7, 1.3)),              # 72 x 47: even row, odd plane stride... (72*47 even)
            make_cfg(dim=3, N=(131, 21, 19), periodic=7, upper=(1.0, 0.2, 0.2)),             # 133 x 23 x 21: odd rows, odd planes, interior at offset 1
            make_cfg(dim=3, N=(259, 10, 18), periodic=1, upper=(1.0, 0.05, 0.1)),            # 261 wide: the 128x4 tile with a partial last tile
            make_cfg(dim=3, N=(258, 11, 21), periodic=5, upper=(1.0, 0.05, 0.1)),            # 260 x 11 x 23: even rows, three tiles, interior at offset 1 (tensor-map staging)
            make_cfg(dim=3, N=(384, 9, 150), upper=(1.0, 0.03, 0.4)))                        # long columns: the ring wraps many times
    for ci, cfg in enumerate(cfgs):
        g = product_grid(cfg, 0, 1)
        n = g.leafGridView().size(0)
        for misalign in (0, 1):                                  # field (and result) start at an odd multiple of 8 bytes
            buf0 = torch.zeros(n + 2, dtype=torch.float64, device="cuda")
            buf1 = torch.zeros(n + 2, dtype=torch.float64, device="cuda")
            c0 = torch.from_numpy(np.random.default_rng(3).uniform(0.0, 1.0, n)).cuda()
            res = {}
            for variant, staging in (("99", "pair"), ("0", "pair"), ("3", "pair"), ("0", "8"), ("3", "8"), ("3", "a16"), ("7", "pair"),
                                     ("2", "pair"), ("9", "pair"), ("3", "tma"), ("3", "tma:64")):
                if misalign and (variant, staging) not in (("99", "pair"), ("0", "pair"), ("3", "pair"), ("3", "a16"), ("3", "tma")):
                    continue
                monkeypatch.setenv("DGB_FV_VARIANT", variant)
                monkeypatch.setenv("DGB_FV_STAGING", staging.split(":")[0])
                monkeypatch.setenv("DGB_FV_ZCHUNK", staging.split(":")[1] if ":" in staging else "16")
                monkeypatch.setenv("DGB_FV_FUSED", "0")         # the serial exchange: the step kernel alone is compared here
                monkeypatch.setenv("DGB_FV_NO_SPLIT", "1")
                fv = dg.FiniteVolume(g, 0, problem, params, "separable")
                c, cn = buf0[misalign:misalign + n], buf1[misalign:misalign + n]
                c.copy_(c0)
                for _ in range(2):
                    fv.step(c, cn)
                    c, cn = cn, c
                info = fv.kernel_info()
                if variant != "99":
                    assert info["kind"] == "ring" and info["variant"] == int(variant), info
                    if staging == "pair":
                        assert info["staging_bytes"] == 16
                    if staging == "8":
                        assert info["staging_bytes"] == 8
                    if staging.startswith("tma"):      # tensor-map staging needs 16-byte strides and an aligned field; else the automatic choice
                        box = g.leafGridView().boxes(0)[0]["overlapfront"][1]
                        assert info["staging"] == ("tma" if box[0] % 2 == 0 and not misalign else "pair"), (info, box)
                res[(variant, staging)] = (c.clone(), fv.last_dt())
            for key, val in res.items():
                assert val[1] == res[("99", "pair")][1]
                assert torch.equal(val[0], res[("99", "pair")][0]), (cfg["N"], misalign, key)


def test_fv_ring_general_mode_equals_march_kernel_bitwise(monkeypatch):
    """a problem whose velocity depends on z (problem 2, column_invariant = false) on the ring kernel's general mode - the
    velocity functor and the upwind decisions evaluated per plane, both tile shapes, aligned and misaligned fields, partial
    tiles, periodic wraps, odd strides, every flow direction inside one grid - against the z-march kernel (DGB_FV_VARIANT=99): same bits"""
    import torch
    import dune_grid_b200 as dg
    params = [1.3, -0.4, 0.9, 0.35, 0.5, 0.5, 0.5, 0.1, 0.45]          # u = (1.3 (z-0.5), -0.4, 0.9 (x-0.5)): all sign patterns occur
    cfgs = (make_cfg(dim=3, N=(150, 37, 70), mode=2), make_cfg(dim=3, N=(70, 45, 33), periodic=3, upper=(1.0, 0.7, 1.3)),
            make_cfg(dim=3, N=(131, 21, 19), periodic=7, upper=(1.0, 0.2, 0.2)), make_cfg(dim=3, N=(259, 10, 18), periodic=1, upper=(1.0, 0.05, 0.1)),
            make_cfg(dim=3, N=(260, 14, 40), upper=(1.0, 0.05, 0.15)))
    for cfg in cfgs:
        g = product_grid(cfg, 0, 1)
        n = g.leafGridView().size(0)
        for misalign in (0, 1):
            buf0 = torch.zeros(n + 2, dtype=torch.float64, device="cuda")
            buf1 = torch.zeros(n + 2, dtype=torch.float64, device="cuda")
            c0 = torch.from_numpy(np.random.default_rng(5).uniform(0.0, 1.0, n)).cuda()
            res = {}
            for variant, zchunk in (("99", "16"), ("0", "16"), ("3", "16"), ("3", "7"), ("0", "256")):
                monkeypatch.setenv("DGB_FV_VARIANT", variant)
                monkeypatch.setenv("DGB_FV_ZCHUNK", zchunk)
                monkeypatch.setenv("DGB_FV_FUSED", "0")
                monkeypatch.setenv("DGB_FV_NO_SPLIT", "1")
                fv = dg.FiniteVolume(g, 0, 2, params, "separable")
                c, cn = buf0[misalign:misalign + n], buf1[misalign:misalign + n]
                c.copy_(c0)
                for _ in range(3):
                    fv.step(c, cn)
                    c, cn = cn, c
                info = fv.kernel_info()
                assert info["kind"] == ("march" if variant == "99" else "ring") and (variant == "99" or info["staging_bytes"] == 16), info
                res[(variant, zchunk)] = (c.clone(), fv.last_dt())
            for key, val in res.items():
                assert val[1] == res[("99", "16")][1], (cfg["N"], misalign, key)
                assert torch.equal(val[0], res[("99", "16")][0]), (cfg["N"], misalign, key)


@pytest.mark.parametrize("mode", ["exact", "separable"])
@pytest.mark.parametrize("cfg,problem,params", [
    (make_cfg(dim=3, N=(40, 23, 37), periodic=7, upper=(1.0, 0.7, 1.3)), 0, [1.0, -0.5, 0.25, 0.1, 0.5, 0.35, 0.6, 0.1, 0.45]),
    (make_cfg(dim=3, N=(150, 9, 40), periodic=5, mode=2), 1, [2.0, 0, 0, 0.5, 0.15, 0.35, 0.9, 0.05, 0.4]),
    (make_cfg(dim=3, N=(12, 300, 5), periodic=2, overlap=2), 0, [-0.3, 0.6, 1.0, 0.7, 0.5, 0.5, 0.5, 0.1, 0.45]),
])
def test_fv_fused_exchange_single_rank(cfg, problem, params, mode, monkeypatch):
    """the fused step (pack + stores into the receiver's buffer + flag inside the step kernel; here the receiver is the
    rank itself through the periodic wrap) against the CPU oracle and, bit for bit, against the serial form
    (step kernel, then communicate)"""
    import torch
    import dune_grid_b200 as dg
    o = best_oracle()
    w = oracle_world(o, cfg, 1)
    ref = [w.fv_init(0, 0, problem, params)]
    res = {}
    for fused in ("1", "0"):
        monkeypatch.setenv("DGB_FV_FUSED", fused)
        monkeypatch.setenv("DGB_FV_NO_SPLIT", "1")
        g = product_grid(cfg, 0, 1)
        fv = dg.FiniteVolume(g, 0, problem, params, mode)
        c = fv.initial()
        cn = torch.empty_like(c)
        dts = []
        for _ in range(5):
            fv.step(c, cn)
            c, cn = cn, c
            dts.append(fv.last_dt())
        assert fv.exchange_mode() == ("fused" if fused == "1" else "serial")
        res[fused] = (np_(c), dts)
    assert np.array_equal(res["1"][0], res["0"][0]) and res["1"][1] == res["0"][1]
    for k in range(5):
        dt_ref = w.fv_step(0, problem, params, ref)
        assert res["1"][1][k] == dt_ref if mode == "exact" else abs(res["1"][1][k] - dt_ref) <= 1e-13 * dt_ref
    if mode == "exact":
        assert np.array_equal(res["1"][0], ref[0])
    else:
        assert np.abs(res["1"][0] - ref[0]).max() <= 1e-13 * np.abs(ref[0]).max()       # tolerance: 1e-13 relative
    w.close()


@pytest.mark.parametrize("mode", ["exact", "separable"])
@pytest.mark.parametrize("N,periodic,overlap", [((40, 23, 37), 7, 1), ((131, 9, 20), 5, 1), ((12, 30, 16), 2, 2), ((257, 12, 18), 1, 1)])
def test_fv_registered_fields_peers_store_into_the_arrays(N, periodic, overlap, mode):
    """dgb_fv_register_fields: the step kernel stores the halo cells straight into the receiver's registered ARRAY (here the rank
    itself through the periodic wrap) - no receive buffer, no unpack launch - and must give the same bits as the buffered
    fused form, and the oracle's result"""
    import torch
    import dune_grid_b200 as dg
    cfg = make_cfg(dim=3, N=N, periodic=periodic, overlap=overlap, upper=(1.0, 0.6, 0.9))
    params = [0.8, -0.5, 0.25, 0.1, 0.5, 0.3, 0.45, 0.1, 0.45]
    o = best_oracle()
    w = oracle_world(o, cfg, 1)
    ref = [w.fv_init(0, 0, 0, params)]
    res = {}
    for registered in (True, False):
        g = product_grid(cfg, 0, 1)
        fv = dg.FiniteVolume(g, 0, 0, params, mode)
        c = fv.initial()
        cn = torch.empty_like(c)
        third = torch.empty_like(c)
        if registered:
            assert fv.register_fields(c, cn) == 2
        l0 = fv.launch_count()
        for _ in range(4):
            fv.step(c, cn)
            c, cn = cn, c
        # x faces one cell thick + production kernel: they stay in the exchange buffer (hybrid); otherwise straight into the array
        hybrid = mode == "separable" and overlap == 1 and (periodic & 1)
        assert fv.exchange_mode() == (("hybrid" if hybrid else "direct") if registered else "fused")
        launches = fv.launch_count() - l0
        if registered:
            assert launches <= 4 * 2 + 1                         # step kernel + one-CTA wait (+ bootstrap), no unpack launches
            if hybrid:
                stale = np_(c)
                fv.fence()                                       # on-demand unpack of the x-overlap column
                assert not np.array_equal(stale, np_(c))
            fv.step(c, third)                                    # an unregistered output array: the buffered form, same object
            assert fv.exchange_mode() == "fused"
            res["mixed"] = np_(third)
        else:
            fv.step(c, cn)
            res["buffered5"] = np_(cn)
        res[registered] = np_(c)
        if registered and hybrid:
            # a longer hybrid run that is only completed at the end: 5 more steps reading the x halo from the exchange buffer
            a, b = c.clone(), torch.empty_like(c)
            fv2 = dg.FiniteVolume(g, 0, 0, params, mode)
            assert fv2.register_fields(a, b) == 2
            for _ in range(5):
                fv2.step(a, b)
                a, b = b, a
            assert fv2.exchange_mode() == "hybrid"
            fv2.fence()
            res["hybrid9"] = np_(a)
            fv2.close()
        elif not registered:
            a, b = c.clone(), torch.empty_like(c)
            for _ in range(5):
                fv.step(a, b)
                a, b = b, a
            res["buffered9"] = np_(a)
        fv.close()
    assert np.array_equal(res[True], res[False])
    assert np.array_equal(res["mixed"], res["buffered5"])
    if "hybrid9" in res:
        assert np.array_equal(res["hybrid9"], res["buffered9"])
    for _ in range(4):
        w.fv_step(0, 0, params, ref)
    if mode == "exact":
        assert np.array_equal(res[True], ref[0])
    else:
        assert np.abs(res[True] - ref[0]).max() <= 1e-13 * np.abs(ref[0]).max()            # tolerance: 1e-13 relative
    w.close()


@pytest.mark.parametrize("N,periodic", [((20, 12, 126), 7), ((36, 10, 128), 4), ((24, 16, 140), 0)])
def test_fv_step_host_pipelined(N, periodic):
    """dgb_fv_step_host on pinned buffers: z-slab pipeline (upload / kernel / download on three streams); with periodic
    directions the fused exchange rides on the slab launches and the received overlap cells are unpacked straight into the
    host result. Bit-equal to the oracle (exact mode), all stored cells including the refreshed overlap; pageable buffers
    take the serial path and give the same bits."""
    import torch
    import dune_grid_b200 as dg
    cfg = make_cfg(dim=3, N=N, periodic=periodic, upper=(1.0, 0.6, 4.0))
    params = [0.8, -0.5, 1.25, 0.1, 0.5, 0.3, 2.0, 0.3, 1.45]
    o = best_oracle()
    w = oracle_world(o, cfg, 1)
    ref = [w.fv_init(0, 0, 0, params)]
    g = product_grid(cfg, 0, 1)
    fv = dg.FiniteVolume(g, 0, 0, params, "exact")
    h_in = torch.from_numpy(ref[0].copy()).pin_memory()
    h_out = torch.full_like(h_in, float("nan")).pin_memory()
    for _ in range(3):
        dt = fv.step_host(h_in, h_out)
        dt_ref = w.fv_step(0, 0, params, ref)
        assert fv.host_path() == "pipelined"
        assert dt == dt_ref
        assert np.array_equal(h_out.numpy(), ref[0])
        h_in, h_out = h_out, h_in
    if periodic:
        assert fv.exchange_mode() == "fused"
    p_in = torch.from_numpy(ref[0].copy())                       # pageable
    p_out = torch.empty_like(p_in)
    fv.step_host(p_in, p_out)
    w.fv_step(0, 0, params, ref)
    assert fv.host_path() == ("serial" if periodic else "pipelined")
    assert np.array_equal(p_out.numpy(), ref[0])
    fv.close()
    w.close()


@pytest.mark.parametrize("registered", [False, True])
@pytest.mark.parametrize("N,periodic", [((131, 21, 19), 7), ((260, 10, 18), 1), ((70, 45, 33), 6)])
def test_fv_guard_bands_stay_untouched(N, periodic, registered, monkeypatch):
    """compute-sanitizer is closed on the GPU pool (profiles/r02_sanitizer_note.md), so the bounds of the step kernels are checked
    here: both fields live between NaN guard bands inside ONE allocation (a store outside a field changes a guard; a load
    outside a field that reaches the arithmetic poisons the result with NaN). Every staging form, buffered / direct / hybrid
    exchange, odd row and plane strides; results must equal the oracle."""
    import torch
    import dune_grid_b200 as dg
    cfg = make_cfg(dim=3, N=N, periodic=periodic, upper=(1.0, 0.3, 0.4))
    params = [0.8, -0.5, 0.25, 0.1, 0.5, 0.15, 0.2, 0.05, 0.3]
    o = best_oracle()
    w = oracle_world(o, cfg, 1)
    ref = [w.fv_init(0, 0, 0, params)]
    for _ in range(4):
        w.fv_step(0, 0, params, ref)
    g = product_grid(cfg, 0, 1)
    n = g.leafGridView().size(0)
    G = 1031                                                     # guard band length (odd: the fields start at odd multiples of 8 bytes)
    for staging in ("pair", "8", "a16"):
        monkeypatch.setenv("DGB_FV_STAGING", staging)
        big = torch.full((3 * G + 2 * n,), float("nan"), dtype=torch.float64, device="cuda")
        c, cn = big[G:G + n], big[2 * G + n:2 * G + 2 * n]
        fv = dg.FiniteVolume(g, 0, 0, params, "separable")
        c.copy_(fv.initial())
        if registered:
            assert fv.register_fields(c, cn) == 2
        for _ in range(4):
            fv.step(c, cn)
            c, cn = cn, c
        fv.fence()
        assert fv.last_dt() > 0
        guards = torch.cat([big[:G], big[G + n:2 * G + n], big[2 * G + 2 * n:]])
        assert bool(torch.isnan(guards).all()), staging                       # no store outside the fields
        got = np_(c)
        assert not np.isnan(got).any(), staging                              # no guard value reached the arithmetic
        assert np.abs(got - ref[0]).max() <= 1e-13 * np.abs(ref[0]).max(), staging
        fv.close()
    w.close()


def test_fv_restart_from_backup(tmp_path):
    """restart of an FV run: grid through BackupRestoreFacility text, field through the checkpoint file;
    the restarted run must continue bit-identically, and the checkpoint must refuse a different grid"""
    import torch
    import dune_grid_b200 as dg
    # mesh sizes that survive the reference's 6-significant-digit text format (0.05, 0.125, 0.0625): the restored grid is then
    # the same grid bit for bit; with e.g. h = 1/12 the reference itself restores a slightly different geometry
    cfg = make_cfg(dim=3, N=(20, 8, 16), periodic=1, mode=1, lower=(0.0, -1.0, 0.5), upper=(1.0, 0.0, 1.5), refine=1)
    params = [1.0, -0.5, 0.25, 0.1, 0.5, -0.5, 1.0, 0.1, 0.45]
    g = product_grid(cfg, 0, 1)
    fv = dg.FiniteVolume(g, None, 0, params, "exact")
    c = fv.initial()
    cn = torch.empty_like(c)
    for _ in range(3):
        fv.step(c, cn)
        c, cn = cn, c
    m = dg.MultipleCodimMultipleGeomTypeMapper(g.leafGridView(), dg.mcmgElementLayout(3))
    m.backupField(c, tmp_path / "c")
    text = g.backup()
    for _ in range(2):
        fv.step(c, cn)
        c, cn = cn, c
    # restart
    g2 = dg.YaspGrid.restore(text, "equidistantoffset")
    m2 = dg.MultipleCodimMultipleGeomTypeMapper(g2.leafGridView(), dg.mcmgElementLayout(3))
    c2 = torch.zeros(m2.size, dtype=torch.float64, device="cuda")
    m2.restoreField(c2, tmp_path / "c")
    fv2 = dg.FiniteVolume(g2, None, 0, params, "exact")
    c2n = torch.empty_like(c2)
    for _ in range(2):
        fv2.step(c2, c2n)
        c2, c2n = c2n, c2
    assert torch.equal(c, c2)
    other = product_grid(make_cfg(dim=3, N=(20, 8, 16), refine=1), 0, 1)
    mo = dg.MultipleCodimMultipleGeomTypeMapper(other.leafGridView(), dg.mcmgElementLayout(3))
    with pytest.raises(Exception):
        mo.restoreField(torch.zeros(mo.size, dtype=torch.float64, device="cuda"), tmp_path / "c")
