"""Shared helpers of the test-suite: build the same grid in the product (dune_grid_b200, through the C ABI)
and in a CPU oracle (oracle/oracle.py) from one configuration dict."""
import numpy as np

from oracle import oracle as orc


def oracle_kinds():
    """oracles that are built here: 'ref' = the reference's own headers (oracle/_ref), 'port' = the restatement"""
    return [k for k in ("ref", "port") if orc.available(k)]


def best_oracle():
    """the reference's own headers when built (oracle/_ref), else the self-contained restatement; ORACLE_KIND=ref|port forces one"""
    import os
    kinds = oracle_kinds()
    assert kinds, "no CPU oracle built: run `make -C oracle` (python -c 'import __graft_entry__ as g; g.build()')"
    want = os.environ.get("ORACLE_KIND")
    if want:
        assert want in kinds, f"oracle '{want}' is not built"
        return orc.load(want)
    return orc.load(kinds[0])


def graded(n, r=1.005, lo=0.0, hi=1.0):
    """geometric grading x_{i+1} = x_i + h0 r^i scaled to [lo, hi] (cf. utility/tensorgridfactory.hh:197)"""
    w = r ** np.arange(n)
    x = np.concatenate([[0.0], np.cumsum(w)])
    return lo + (hi - lo) * x / x[-1]


def make_cfg(dim=3, N=(4, 4, 4), mode=0, lower=None, upper=None, tensor=None, periodic=0, overlap=1,
             partitioner=0, fixed_dims=None, refine=0, keep=1):
    lower = tuple(lower) if lower is not None else (0.0,) * dim
    upper = tuple(upper) if upper is not None else (1.0,) * dim
    if mode == 2 and tensor is None:
        tensor = [graded(N[i], 1.0 + 0.01 * (i + 1), lower[i], upper[i]) for i in range(dim)]
    return dict(dim=dim, N=tuple(N[:dim]), mode=mode, lower=lower, upper=upper, tensor=tensor, periodic=periodic,
                overlap=overlap, partitioner=partitioner, fixed_dims=fixed_dims, refine=refine, keep=keep)


def oracle_world(o, cfg, P):
    s = orc.Spec(dim=cfg["dim"], N=cfg["N"], coord_mode=cfg["mode"], lower=cfg["lower"], upper=cfg["upper"],
                 tensor=cfg["tensor"], periodic=cfg["periodic"], overlap=cfg["overlap"], partitioner=cfg["partitioner"],
                 fixed_dims=cfg["fixed_dims"] or (1, 1, 1), refine=cfg["refine"], keep_overlap=cfg["keep"])
    return o.world(s, P)


def product_grid(cfg, rank, P, comm=None):
    import dune_grid_b200 as dg
    g = dg.YaspGrid(dim=cfg["dim"], N=cfg["N"], lower=cfg["lower"], upper=cfg["upper"], coords=cfg["tensor"] if cfg["mode"] == 2 else None,
                    periodic=cfg["periodic"], overlap=cfg["overlap"],
                    coordinates=["equidistant", "equidistantoffset", "tensorproduct"][cfg["mode"]],
                    rank=rank, nranks=P, partitioner=["default", "powerd", "fixed"][cfg["partitioner"]],
                    fixed_dims=cfg["fixed_dims"], comm=comm)
    g.refineOptions(bool(cfg["keep"]))
    if cfg["refine"]:
        g.globalRefine(cfg["refine"])
    return g


BOXES = ("overlapfront", "overlap", "interiorborder", "interior")


def compare_world(w, grid_of_rank, d, P, nlevels):
    """host descriptors of the product grids (grid_of_rank(r)) against an oracle world: torus, boundary segments, levels,
    nested boxes, index offsets, sizes, all eight send/recv lists and the coordinates, bit for bit"""
    for r in range(P):
        g = grid_of_rank(r)
        assert g.torusDims() == w.torus_dims()
        assert g.numBoundarySegments() == w.num_boundary_segments(r)
        assert g.maxLevel == w.max_level
        assert np.array_equal(g.domainSize(), w.domain_size()) or r > 0      # the oracle reports rank 0's domain size
        for lvl in range(nlevels):
            gv = g.levelGridView(lvl)
            assert gv.overlapSize(0) == w.overlap_size(lvl)
            for c in range(d + 1):
                a, b = gv.boxes(c), w.boxes(r, lvl, c)
                assert len(a) == len(b)
                for x, y in zip(a, b):
                    assert x["shift"] == y["shift"] and x["offset"] == y["offset"]
                    for k in BOXES:
                        assert x[k][0][:d] == y[k][0][:d] and x[k][1][:d] == y[k][1][:d], (r, lvl, c, k, x, y)
                assert gv.size(c) == w.size(r, lvl, c)
                for which in range(8):
                    la, lb = g.commList(lvl, c, which), w.lists(r, lvl, c, which)
                    assert la.shape == lb.shape and np.array_equal(la, lb), (r, lvl, c, which, la, lb)
            for ax in range(d):
                fa, ca = gv.coordinates(ax)
                fb, cb = w.coordinates(r, lvl, ax)
                assert fa == fb and np.array_equal(ca, cb, equal_nan=True), (r, lvl, ax)


def fv_init_all(w, P, level, problem, params):
    """initial fields of all ranks; the per-rank oracle calls run concurrently (ctypes releases the GIL, the world is only read)"""
    from concurrent.futures import ThreadPoolExecutor
    if P == 1:
        return [w.fv_init(0, level, problem, params)]
    with ThreadPoolExecutor(max_workers=min(P, 32)) as ex:
        return list(ex.map(lambda r: w.fv_init(r, level, problem, params), range(P)))


def assemble_global(w, cfg, P, level, per_rank):
    """one array over all cells of the level (x fastest) from the per-rank arrays of an oracle world: every rank contributes
    its INTERIOR cells (the FV result does not depend on the partition: same cells, same arithmetic, global dt)"""
    d = cfg["dim"]
    N = [n << level for n in cfg["N"]]
    shape = tuple(reversed(N))
    out = np.full(shape, np.nan)
    for r in range(P):
        b = w.boxes(r, level, 0)[0]
        (oo, os_), (io, is_) = b["overlapfront"], b["interior"]
        a = per_rank[r].reshape(tuple(reversed(os_[:d])))
        src = tuple(slice(io[i] - oo[i], io[i] - oo[i] + is_[i]) for i in reversed(range(d)))
        dst = tuple(slice(io[i], io[i] + is_[i]) for i in reversed(range(d)))
        out[dst] = a[src]
    assert not np.isnan(out).any()
    return out.reshape(-1)
