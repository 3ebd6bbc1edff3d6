"""Golden cases shared by tests/golden/make_golden.py (generator, needs oracle/_ref = the reference's own headers)
and the tests that consume the committed fixtures. A snapshot is a flat dict name -> numpy array holding everything the
oracle API returns for one configuration: descriptors, sweeps, mapper numbering, communicate results and FV steps."""
import numpy as np

from tests.common import make_cfg, oracle_world

TENSOR_X = np.array([-1, -.5, -.25, -.125, 0, .125, .25, .5, 1.0])     # test/yasp/test-yaspgrid.hh:107-136
TENSOR_Y = np.array([-1, -.5, 0, .5, 1.0])

# name -> (cfg, ranks, mapper layout used for communicate, FV (problem, params) or None)
CASES = {
    "yasp2d_8x4_p2": (make_cfg(dim=2, N=(8, 4)), 2, [1, 0, 1], (0, [1.0, 0.5, 0.0, 0.3, 0.25, 0.25, 0.0, 0.125, 0.2])),
    "yasp3d_8x4x4_p2": (make_cfg(dim=3, N=(8, 4, 4)), 2, [1, 0, 0, 1], (0, [1.0, 1.0, 1.0, 0.0, 0.25, 0.25, 0.25, 0.125, 0.2])),
    "yasp3d_tensor_8x4x4_p2": (make_cfg(dim=3, N=(8, 4, 4), mode=2, tensor=[TENSOR_X, TENSOR_Y, TENSOR_Y]), 2, [1, 1, 1, 1],
                               (1, [2.0, 0, 0, 0.5, 0.1, 0.1, 0.1, 0.3, 1.2])),
    "yasp3d_periodic_6x5x4_p1": (make_cfg(dim=3, N=(6, 5, 4), periodic=7, upper=(1.0, 2.0, 0.5)), 1, [1, 0, 0, 0],
                                 (0, [-1.0, 0.7, 0.3, 0.0, 0.5, 1.0, 0.25, 0.1, 0.45])),
    "yasp2d_offset_refine_p2": (make_cfg(dim=2, N=(8, 4), mode=1, lower=(-1.0, -0.5), upper=(1.0, 0.5), refine=1, keep=0), 2,
                                [2, 0, 1], None),
    "yasp3d_6x4x5_per_y_refine_p2": (make_cfg(dim=3, N=(6, 4, 5), periodic=2, upper=(1.0, 2.0, 3.0), refine=1, keep=0), 2,
                                     [1, 0, 0, 0], (1, [1.5, 0, 0, 0.25, 0.5, 1.0, 1.5, 0.2, 0.9])),
}
GAUSS2 = 0.5 + np.array([-0.5, 0.5]) / np.sqrt(3.0)


def quad_points(dim):
    g = np.meshgrid(*([GAUSS2] * dim), indexing="ij")
    return np.stack([x.T.reshape(-1) for x in g], axis=1)


def comm_data(size, seed):
    rng = np.random.default_rng(seed)
    return [rng.standard_normal(size), rng.standard_normal(size * 3)]


def var_data(size, seed):
    """a variable-size data handle as CSR arrays: entity i carries 0..3 items"""
    rng = np.random.default_rng(seed)
    n = rng.integers(0, 4, size=size)
    off = np.zeros(size + 1, dtype=np.int64)
    np.cumsum(n, out=off[1:])
    return off, rng.standard_normal(int(off[-1]))


def var_layout(layout):
    return [min(1, b) for b in layout]


def snapshot(o, name):
    """everything `o` (an oracle.OracleLib) computes for case `name`"""
    cfg, P, layout, fv = CASES[name]
    w = oracle_world(o, cfg, P)
    d = cfg["dim"]
    s = {}
    s["torus_dims"] = np.array(w.torus_dims())
    s["max_level"] = np.array([w.max_level])
    s["domain_size"] = w.domain_size()
    qp = quad_points(d)
    for r in range(P):
        s[f"r{r}/nbs"] = np.array([w.num_boundary_segments(r)])
        for lvl in range(cfg["refine"] + 1):
            k = f"r{r}/l{lvl}"
            for ax in range(d):
                first, c = w.coordinates(r, lvl, ax)
                s[f"{k}/coord{ax}/first"] = np.array([first])
                s[f"{k}/coord{ax}/x"] = c
            for codim in range(d + 1):
                bx = w.boxes(r, lvl, codim)
                s[f"{k}/c{codim}/boxes"] = np.array([[b["shift"], b["offset"]] + [x for q in ("overlapfront", "overlap", "interiorborder", "interior")
                                                                                   for x in (b[q][0] + b[q][1])] for b in bx], dtype=np.int32)
                for which in range(8):
                    s[f"{k}/c{codim}/list{which}"] = w.lists(r, lvl, codim, which)
                for pit in range(5):
                    e = w.entities(r, lvl, codim, pit)
                    for q, a in e.items():
                        s[f"{k}/c{codim}/p{pit}/{q}"] = a
            for cc in range(d + 1):
                s[f"{k}/subindex{cc}"] = w.subindices(r, lvl, cc, 4)
            s[f"{k}/mapper/size"] = np.array([w.mapper_size(r, lvl, layout)])
            for codim in range(d + 1):
                if layout[codim]:
                    s[f"{k}/mapper/index{codim}"] = w.mapper_index(r, lvl, layout, codim, 4)
                    s[f"{k}/mapper/subindex{codim}"] = w.mapper_subindex(r, lvl, layout, codim, 4)
            for pit in (0, 4):
                it = w.intersections(r, lvl, pit)
                for q, a in it.items():
                    s[f"{k}/is/p{pit}/{q}"] = a
            g = w.geometry_eval(r, lvl, qp, 4)
            for q, a in g.items():
                s[f"{k}/geo/{q}"] = a
            gg = w.geogrid_eval(r, lvl, 1, [0.05], qp, 4)
            for q, a in gg.items():
                s[f"{k}/geogrid/{q}"] = a
    lvl = cfg["refine"]
    for iftype in range(5):
        for direction in (0, 1):
            for op in (0, 1):
                data = [comm_data(w.mapper_size(r, lvl, layout), 1000 + 10 * r) for r in range(P)]
                n = w.communicate(lvl, layout, iftype, direction, op, data, [1, 3])
                s[f"comm/i{iftype}d{direction}o{op}/sent"] = np.array([n])
                for r in range(P):
                    for a in range(2):
                        s[f"comm/i{iftype}d{direction}o{op}/r{r}a{a}"] = data[r][a]
    # data handle of variable size (fixedSize() == false): the scatter calls each rank receives, in the reference's order
    vl = var_layout(layout)
    for iftype in range(5):
        for direction in (0, 1):
            vd = [var_data(w.mapper_size(r, lvl, vl), 2000 + 10 * r) for r in range(P)]
            got = w.communicate_variable(lvl, vl, iftype, direction, [v[0] for v in vd], [v[1] for v in vd])
            for r in range(P):
                for q, a in zip(("index", "n", "items"), got[r]):
                    s[f"commvar/i{iftype}d{direction}/r{r}/{q}"] = a
    if fv is not None:
        problem, params = fv
        c = [w.fv_init(r, lvl, problem, params) for r in range(P)]
        for r in range(P):
            s[f"fv/init/r{r}"] = c[r].copy()
        dts = []
        for step in range(3):
            dts.append(w.fv_step(lvl, problem, params, c))
        s["fv/dt"] = np.array(dts)
        for r in range(P):
            s[f"fv/step3/r{r}"] = c[r].copy()
    w.close()
    return s
