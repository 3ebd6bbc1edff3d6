"""N>1 host-side logic on CPU: two processes (gloo, 127.0.0.1), one rank each. Every rank builds ITS OWN grid and exchange
plan through the C ABI (descriptor arithmetic only - no GPU), the test packs / unpacks with numpy from the rank's send /
recv box lists, ships the plan's per-peer segments with torch.distributed send/recv in plan order, and compares the result
with the CPU oracle's communicate for all ranks. This checks what the NCCL path relies on: that the plans of different ranks
agree (segment sizes, peer order, message pairing) without any set-up communication."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WHICH = {0: (4, 5), 1: (6, 7), 2: (2, 3), 3: (2, 3), 4: (0, 1)}      # interface -> (send list, recv list), yaspgrid.hh:1506-1525


def box_indices(entry, of_box, dim):
    """consecutive indices of the entities of one list entry (x fastest), ygrid.hh:472-479"""
    off, origin, size = int(entry[3]), entry[4:4 + dim], entry[7:7 + dim]
    idx = np.zeros([int(s) for s in size[::-1]], dtype=np.int64)
    stride = 1
    for i in range(dim):
        shape = [1] * dim
        shape[dim - 1 - i] = int(size[i])
        idx = idx + ((np.arange(size[i]) + origin[i] - of_box[0][i]) * stride).reshape(shape)
        stride *= of_box[1][i]
    return off + idx.reshape(-1)


def worker(rank, world, port, cfgs, out):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import dune_grid_b200 as dg
    from tests.common import best_oracle, oracle_world, product_grid
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ok = True
    try:
        o = best_oracle()
        for cfg, layout in cfgs:
            d = cfg["dim"]
            w = oracle_world(o, cfg, world)
            g = product_grid(cfg, rank, world)
            gv = g.leafGridView()
            m = dg.MultipleCodimMultipleGeomTypeMapper(gv, layout)
            items = [1, 2]
            for iftype in range(5):
                for direction in (0, 1):
                    rng = np.random.default_rng(5)
                    host = [[rng.standard_normal(w.mapper_size(r, 0, layout) * it) for it in items] for r in range(world)]
                    mine = [a.copy() for a in host[rank]]
                    w.communicate(0, layout, iftype, direction, 1, host, items)             # op = add
                    sw, rw = WHICH[iftype]
                    if direction == 1:
                        sw, rw = rw, sw
                    send_segs, recv_segs = gv.commPlan(m, iftype, direction, items)
                    # pack per peer: codim d..0, list order inside a codim, array-major / block-major / item-minor per entity
                    def walk(which):
                        per_peer = {}
                        for codim in range(d, -1, -1):
                            if not layout[codim]:
                                continue
                            boxes = {b["shift"]: b for b in gv.boxes(codim)}
                            for e in g.commList(0, codim, which):
                                idx = box_indices(e, boxes[int(e[2])]["overlapfront"], d)
                                base = idx * layout[codim] + int(m.m.offset[codim])
                                per_peer.setdefault(int(e[0]), []).append((base, layout[codim]))
                        return per_peer
                    sends, recvs = walk(sw), walk(rw)
                    def flat_positions(chunks):
                        """(array, position) of every double of a segment in message order"""
                        pos = []
                        for base, block in chunks:
                            for ent in base:
                                for a, it in enumerate(items):
                                    for b in range(block):
                                        for i in range(it):
                                            pos.append((a, (ent + b) * it + i))
                        return pos
                    reqs, inbox = [], {}
                    for peer, off, cnt in send_segs:
                        pos = flat_positions(sends.get(peer, []))
                        ok = ok and len(pos) == cnt
                        buf = torch.tensor([mine[a][p] for a, p in pos], dtype=torch.float64)
                        if peer == rank:
                            inbox[peer] = buf.clone()
                        else:
                            reqs.append(dist.isend(buf, peer))
                    for peer, off, cnt in recv_segs:
                        if peer != rank:
                            inbox[peer] = torch.empty(cnt, dtype=torch.float64)
                            reqs.append(dist.irecv(inbox[peer], peer))
                    for r in reqs:
                        r.wait()
                    for peer, off, cnt in recv_segs:
                        ok = ok and len(flat_positions(recvs.get(peer, []))) == cnt
                    # scatter in the reference's order: recv-list order across ALL peers (yaspgrid.hh:1694-1729), each entry taking
                    # the next chunk of its peer's segment - entities fed by several neighbours accumulate in that order
                    cursor = {peer: 0 for peer in inbox}
                    for codim in range(d, -1, -1):
                        if not layout[codim]:
                            continue
                        boxes = {b["shift"]: b for b in gv.boxes(codim)}
                        for e in g.commList(0, codim, rw):
                            idx = box_indices(e, boxes[int(e[2])]["overlapfront"], d)
                            pos = flat_positions([(idx * layout[codim] + int(m.m.offset[codim]), layout[codim])])
                            peer = int(e[0])
                            chunk = inbox[peer].numpy()[cursor[peer]:cursor[peer] + len(pos)]
                            cursor[peer] += len(pos)
                            for (a, p), v in zip(pos, chunk):
                                mine[a][p] = v + mine[a][p]
                    for a in range(len(items)):
                        ok = ok and np.array_equal(mine[a], host[rank][a])
                    dist.barrier()
            w.close()
    finally:
        t = torch.tensor([1 if ok else 0])
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        out[rank] = int(t.item())
        dist.destroy_process_group()


def test_two_rank_exchange_over_gloo():
    import torch.multiprocessing as mp
    from tests.common import make_cfg
    cfgs = [(make_cfg(dim=2, N=(8, 4)), [1, 0, 1]),
            (make_cfg(dim=3, N=(8, 4, 4)), [1, 0, 0, 1]),
            (make_cfg(dim=3, N=(8, 5, 4), periodic=5, overlap=2, upper=(1.0, 2.0, 3.0)), [2, 1, 0, 1])]
    ctx = mp.get_context("spawn")
    out = ctx.Manager().dict()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=worker, args=(r, 2, port, cfgs, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert out.get(0) == 1 and out.get(1) == 1
