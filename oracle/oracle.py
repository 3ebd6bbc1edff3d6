"""ctypes front end of the CPU oracles (oracle/oracle_api.h).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs. The product package (dune_grid_b200) never imports this module.

    from oracle.oracle import load, Spec
    orc = load("ref")            # oracle/_ref/liboracle_ref.so  (the reference's own headers)
    orc = load("port")           # oracle/liboracle_port.so      (self-contained restatement)
    w = orc.world(Spec(dim=3, N=(8,4,4)), nranks=2)
"""
import ctypes as C
import os
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIBS = {"ref": os.path.join(HERE, "_ref", "liboracle_ref.so"),
        "port": os.path.join(HERE, "liboracle_port.so")}

# dune/grid/common/gridenums.hh
Interior_Partition, InteriorBorder_Partition, Overlap_Partition, OverlapFront_Partition, All_Partition, Ghost_Partition = range(6)
InteriorEntity, BorderEntity, OverlapEntity, FrontEntity, GhostEntity = range(5)
(InteriorBorder_InteriorBorder_Interface, InteriorBorder_All_Interface, Overlap_OverlapFront_Interface,
 Overlap_All_Interface, All_All_Interface) = range(5)
ForwardCommunication, BackwardCommunication = 0, 1
LIST_NAMES = ["send_of_of", "recv_of_of", "send_o_of", "recv_of_o", "send_ib_ib", "recv_ib_ib", "send_ib_of", "recv_of_ib"]


class _CSpec(C.Structure):
    _fields_ = [("dim", C.c_int), ("coord_mode", C.c_int), ("lower", C.c_double * 3), ("upper", C.c_double * 3),
                ("N", C.c_int * 3), ("tensor", C.POINTER(C.c_double) * 3), ("periodic", C.c_uint), ("overlap", C.c_int),
                ("partitioner", C.c_int), ("fixed_dims", C.c_int * 3), ("refine", C.c_int), ("keep_overlap", C.c_int)]


@dataclass
class Spec:
    dim: int = 3
    N: Sequence[int] = (4, 4, 4)
    coord_mode: int = 0                       # 0 equidistant, 1 equidistant+offset, 2 tensor product
    lower: Sequence[float] = (0.0, 0.0, 0.0)
    upper: Sequence[float] = (1.0, 1.0, 1.0)
    tensor: Optional[Sequence[np.ndarray]] = None
    periodic: int = 0
    overlap: int = 1
    partitioner: int = 0
    fixed_dims: Sequence[int] = (1, 1, 1)
    refine: int = 0
    keep_overlap: int = 1
    _keep: list = field(default_factory=list, repr=False)

    def c(self):
        s = _CSpec()
        s.dim, s.coord_mode = self.dim, self.coord_mode
        for i in range(3):
            s.lower[i] = self.lower[i] if i < len(self.lower) else 0.0
            s.upper[i] = self.upper[i] if i < len(self.upper) else 1.0
            s.N[i] = self.N[i] if i < self.dim else 1
            s.fixed_dims[i] = self.fixed_dims[i] if i < len(self.fixed_dims) else 1
        if self.coord_mode == 2:
            for i in range(self.dim):
                a = np.ascontiguousarray(self.tensor[i], dtype=np.float64)
                assert a.size == self.N[i] + 1
                self._keep.append(a)
                s.tensor[i] = a.ctypes.data_as(C.POINTER(C.c_double))
        s.periodic, s.overlap, s.partitioner = self.periodic, self.overlap, self.partitioner
        s.refine, s.keep_overlap = self.refine, self.keep_overlap
        return s


def _p(a, t):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


class OracleError(RuntimeError):
    pass


class OracleLib:
    def __init__(self, path):
        self.path = path
        L = self.lib = C.CDLL(path)
        vp, i32, i64, dbl = C.c_void_p, C.c_int, C.c_int64, C.c_double
        pi, pd = C.POINTER(C.c_int), C.POINTER(C.c_double)
        pu32, pu8, pi32 = C.POINTER(C.c_uint32), C.POINTER(C.c_uint8), C.POINTER(C.c_int32)
        ppd = C.POINTER(pd)
        sig = {
            "orc_kind": (C.c_char_p, []), "orc_last_error": (C.c_char_p, []),
            "orc_partition": (i32, [i32, pi, i32, i32, i32, pi, pi]),
            "orc_entity_shift": (i32, [i32, i32, i32]), "orc_entity_move": (i32, [i32, i32, i32]),
            "orc_sub_entities": (i32, [i32, i32, i32]),
            "orc_create": (vp, [C.POINTER(_CSpec), i32]), "orc_destroy": (None, [vp]),
            "orc_max_level": (i32, [vp]), "orc_torus_dims": (None, [vp, pi]), "orc_domain_size": (None, [vp, pd]),
            "orc_size": (i64, [vp, i32, i32, i32]), "orc_num_boundary_segments": (i32, [vp, i32]),
            "orc_overlap_size": (i32, [vp, i32]),
            "orc_boxes": (i32, [vp, i32, i32, i32, pi]), "orc_list": (i32, [vp, i32, i32, i32, i32, pi, i32]),
            "orc_coordinates": (i32, [vp, i32, i32, i32, pd, pi]),
            "orc_entities": (i64, [vp, i32, i32, i32, i32, pu32, pu8, pi32, pd, pd, pd, pd]),
            "orc_subindices": (i64, [vp, i32, i32, i32, i32, pu32]),
            "orc_mapper_size": (i64, [vp, i32, i32, pu32]),
            "orc_mapper_index": (i64, [vp, i32, i32, pu32, i32, i32, pu32]),
            "orc_mapper_subindex": (i64, [vp, i32, i32, pu32, i32, i32, pu32]),
            "orc_intersections": (i64, [vp, i32, i32, i32, pu8, pu8, pi32, pi32, pd, pd, pd]),
            "orc_geometry_eval": (i64, [vp, i32, i32, i32, i32, pd, pd, pd, pd]),
            "orc_geometry_eval_codim": (i64, [vp, i32, i32, i32, i32, i32, pd, pd, pd, pd, pd]),
            "orc_geogrid_eval": (i64, [vp, i32, i32, i32, i32, pd, i32, pd, pd, pd, pd, pd, pd]),
            "orc_geogrid_intersections": (i64, [vp, i32, i32, i32, i32, pd, i32, pd, pd, pd, pd, pd, pd, pd, pd]),
            "orc_communicate": (i64, [vp, i32, pu32, i32, i32, i32, i32, ppd, pi]),
            "orc_communicate_variable": (i64, [vp, i32, pu32, i32, i32, C.POINTER(C.POINTER(C.c_int64)), ppd, C.POINTER(C.c_int64),
                                               C.POINTER(C.c_int64), C.POINTER(pu32), C.POINTER(C.POINTER(C.c_int64)), ppd]),
            "orc_global_index": (i64, [vp, i32, i32, C.POINTER(pi32)]),
            "orc_backup": (i32, [vp, i32, C.c_char_p, i32]),
            "orc_restore": (vp, [i32, i32, i32, C.POINTER(C.c_char_p)]),
            "orc_vtu_piece": (i64, [i32, C.c_uint, C.c_uint, i32, pi, C.POINTER(C.c_char_p), pi, pi, pi, ppd, C.c_char_p, C.c_char_p,
                                    C.c_char_p, C.c_char_p, C.c_char_p, i64]),
            "orc_fv_init": (i32, [vp, i32, i32, i32, pd, pd]),
            "orc_fv_step": (dbl, [vp, i32, i32, pd, ppd, ppd]),
            "orc_fv_run": (dbl, [vp, i32, i32, pd, ppd, i32, pd]),
        }
        for name, (res, args) in sig.items():
            f = getattr(L, name)
            f.restype, f.argtypes = res, args
        self.kind = L.orc_kind().decode()

    def err(self):
        return self.lib.orc_last_error().decode()

    def vtu_piece(self, outputtype, ncells, npoints, arrays, point_attrs=("", ""), cell_attrs=("", "")):
        """One .vtu file from the reference's VTUWriter classes (REFERENCE library only). arrays: list of
        (section, name, ncomps, precision, values) in file order; section 0 PointData, 1 CellData, 2 Points, 3 Cells;
        precision 0 Float32, 1 Float64, 2 Int32, 3 UInt8. Returns bytes."""
        n = len(arrays)
        vals = [np.ascontiguousarray(a[4], dtype=np.float64).reshape(-1) for a in arrays]
        sec = (C.c_int * n)(*[a[0] for a in arrays])
        names = (C.c_char_p * n)(*[a[1].encode() for a in arrays])
        ncomps = (C.c_int * n)(*[a[2] for a in arrays])
        nitems = (C.c_int * n)(*[v.size // a[2] for a, v in zip(arrays, vals)])
        prec = (C.c_int * n)(*[a[3] for a in arrays])
        ptrs = (C.POINTER(C.c_double) * n)(*[v.ctypes.data_as(C.POINTER(C.c_double)) for v in vals])
        attrs = [x.encode() for x in (*point_attrs, *cell_attrs)]
        size = self.lib.orc_vtu_piece(outputtype, ncells, npoints, n, sec, names, ncomps, nitems, prec, ptrs, *attrs, None, 0)
        if size < 0:
            raise OracleError(self.err())
        buf = C.create_string_buffer(size)
        self.lib.orc_vtu_piece(outputtype, ncells, npoints, n, sec, names, ncomps, nitems, prec, ptrs, *attrs, buf, size)
        return buf.raw[:size]

    def partition(self, N, P, overlap=1, partitioner=0, fixed=None):
        d = len(N)
        Na = (C.c_int * d)(*N)
        fa = (C.c_int * d)(*(fixed or [1] * d))
        out = (C.c_int * d)()
        rc = self.lib.orc_partition(d, Na, P, overlap, partitioner, fa, out)
        if rc != 0:
            raise OracleError(self.err())
        return tuple(out)

    def entity_shift(self, dim, i, cc):
        return self.lib.orc_entity_shift(dim, i, cc)

    def entity_move(self, dim, i, cc):
        return self.lib.orc_entity_move(dim, i, cc)

    def sub_entities(self, dim, d, c):
        return self.lib.orc_sub_entities(dim, d, c)

    def world(self, spec: Spec, nranks: int = 1):
        return World(self, spec, nranks)

    def restore(self, dim, coord_mode, texts):
        """a world whose rank r is restored by BackupRestoreFacility::restore from texts[r]"""
        arr = (C.c_char_p * len(texts))(*[t.encode() for t in texts])
        h = self.lib.orc_restore(dim, coord_mode, len(texts), arr)
        if not h:
            raise OracleError(self.err())
        w = World.__new__(World)
        w.o, w.spec, w.P, w.dim, w.h, w.L = self, None, len(texts), dim, h, self.lib
        return w


def _block(block):
    b = list(block) + [0] * (4 - len(block))
    return (C.c_uint32 * 4)(*b)


class World:
    def __init__(self, lib: OracleLib, spec: Spec, nranks: int):
        self.o, self.spec, self.P, self.dim = lib, spec, nranks, spec.dim
        cs = spec.c()
        self.h = lib.lib.orc_create(C.byref(cs), nranks)
        if not self.h:
            raise OracleError(lib.err())
        self.L = lib.lib

    def close(self):
        if self.h:
            self.L.orc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def global_index(self, level, codim):
        """Dune::GlobalIndexSet(levelGridView(level), codim) on all ranks: ([per-rank int32 array in local index order], size)"""
        outs = [np.full(self.size(r, level, codim), -7, dtype=np.int32) for r in range(self.P)]
        ptrs = (C.POINTER(C.c_int32) * self.P)(*[_p(a, C.c_int32) for a in outs])
        n = self._chk(self.L.orc_global_index(self.h, level, codim, ptrs))
        return outs, n

    def backup(self, rank):
        """the text BackupRestoreFacility::backup(grid, stream) writes on `rank`"""
        n = self._chk(self.L.orc_backup(self.h, rank, None, 0))
        buf = C.create_string_buffer(n + 1)
        self._chk(self.L.orc_backup(self.h, rank, buf, n + 1))
        return buf.value.decode()

    def _chk(self, rc):
        if rc < 0:
            raise OracleError(self.o.err())
        return rc

    @property
    def max_level(self):
        return self.L.orc_max_level(self.h)

    def torus_dims(self):
        d = (C.c_int * 3)(1, 1, 1)
        self.L.orc_torus_dims(self.h, d)
        return tuple(d)[:self.dim]

    def domain_size(self):
        d = (C.c_double * 3)()
        self.L.orc_domain_size(self.h, d)
        return np.array(d[:self.dim])

    def size(self, rank, level, codim):
        return self._chk(self.L.orc_size(self.h, rank, level, codim))

    def num_boundary_segments(self, rank):
        return self.L.orc_num_boundary_segments(self.h, rank)

    def overlap_size(self, level):
        return self._chk(self.L.orc_overlap_size(self.h, level))

    def boxes(self, rank, level, codim):
        out = np.zeros(8 * 26, dtype=np.int32)
        n = self._chk(self.L.orc_boxes(self.h, rank, level, codim, _p(out, C.c_int)))
        res = []
        for k in range(n):
            o = out[26 * k:26 * (k + 1)]
            b = {"shift": int(o[0]), "offset": int(o[1])}
            for q, name in enumerate(["overlapfront", "overlap", "interiorborder", "interior"]):
                b[name] = (tuple(int(x) for x in o[2 + 6 * q:5 + 6 * q]), tuple(int(x) for x in o[5 + 6 * q:8 + 6 * q]))
            res.append(b)
        return res

    def lists(self, rank, level, codim, which):
        cap = 4096
        out = np.zeros(cap * 10, dtype=np.int32)
        n = self._chk(self.L.orc_list(self.h, rank, level, codim, which, _p(out, C.c_int), cap))
        return out[:n * 10].reshape(n, 10).copy()

    def coordinates(self, rank, level, axis):
        out = np.zeros(1 << 22, dtype=np.float64)
        first = C.c_int(0)
        n = self._chk(self.L.orc_coordinates(self.h, rank, level, axis, _p(out, C.c_double), C.byref(first)))
        return first.value, out[:n].copy()

    def entities(self, rank, level, codim, pitype=All_Partition, geometry=True):
        n = self._chk(self.L.orc_entities(self.h, rank, level, codim, pitype, None, None, None, None, None, None, None))
        d = self.dim
        r = {"index": np.zeros(n, np.uint32), "ptype": np.zeros(n, np.uint8), "coord": np.zeros((n, d), np.int32)}
        if geometry:
            r.update(lower=np.zeros((n, d)), upper=np.zeros((n, d)), center=np.zeros((n, d)), volume=np.zeros(n))
        self._chk(self.L.orc_entities(self.h, rank, level, codim, pitype, _p(r["index"], C.c_uint32), _p(r["ptype"], C.c_uint8),
                                      _p(r["coord"], C.c_int32), _p(r.get("lower"), C.c_double), _p(r.get("upper"), C.c_double),
                                      _p(r.get("center"), C.c_double), _p(r.get("volume"), C.c_double)))
        return r

    def ncells(self, rank, level, pitype):
        return self._chk(self.L.orc_entities(self.h, rank, level, 0, pitype, None, None, None, None, None, None, None))

    def subindices(self, rank, level, cc, pitype=All_Partition):
        n = self.ncells(rank, level, pitype)
        ns = self.o.sub_entities(self.dim, self.dim, cc)
        out = np.zeros((n, ns), np.uint32)
        self._chk(self.L.orc_subindices(self.h, rank, level, pitype, cc, _p(out, C.c_uint32)))
        return out

    def mapper_size(self, rank, level, block):
        return self._chk(self.L.orc_mapper_size(self.h, rank, level, _block(block)))

    def mapper_index(self, rank, level, block, codim, pitype=All_Partition):
        n = self._chk(self.L.orc_entities(self.h, rank, level, codim, pitype, None, None, None, None, None, None, None))
        out = np.zeros(n, np.uint32)
        self._chk(self.L.orc_mapper_index(self.h, rank, level, _block(block), codim, pitype, _p(out, C.c_uint32)))
        return out

    def mapper_subindex(self, rank, level, block, cc, pitype=All_Partition):
        n = self.ncells(rank, level, pitype)
        ns = self.o.sub_entities(self.dim, self.dim, cc)
        out = np.zeros((n, ns), np.uint32)
        self._chk(self.L.orc_mapper_subindex(self.h, rank, level, _block(block), pitype, cc, _p(out, C.c_uint32)))
        return out

    def intersections(self, rank, level, pitype=All_Partition):
        n = self.ncells(rank, level, pitype)
        d = self.dim
        r = {"neighbor": np.zeros((n, 2 * d), np.uint8), "boundary": np.zeros((n, 2 * d), np.uint8),
             "outside": np.zeros((n, 2 * d), np.int32), "bsi": np.zeros((n, 2 * d), np.int32),
             "normal": np.zeros((n, 2 * d, d)), "center": np.zeros((n, 2 * d, d)), "volume": np.zeros((n, 2 * d))}
        self._chk(self.L.orc_intersections(self.h, rank, level, pitype, _p(r["neighbor"], C.c_uint8), _p(r["boundary"], C.c_uint8),
                                           _p(r["outside"], C.c_int32), _p(r["bsi"], C.c_int32), _p(r["normal"], C.c_double),
                                           _p(r["center"], C.c_double), _p(r["volume"], C.c_double)))
        return r

    def geometry_eval(self, rank, level, qp, pitype=All_Partition):
        qp = np.ascontiguousarray(qp, dtype=np.float64).reshape(-1, self.dim)
        n, nq, d = self.ncells(rank, level, pitype), qp.shape[0], self.dim
        r = {"global": np.zeros((n, nq, d)), "jit": np.zeros((n, nq, d, d)), "detJ": np.zeros((n, nq))}
        self._chk(self.L.orc_geometry_eval(self.h, rank, level, pitype, nq, _p(qp, C.c_double), _p(r["global"], C.c_double),
                                           _p(r["jit"], C.c_double), _p(r["detJ"], C.c_double)))
        return r

    def geometry_eval_codim(self, rank, level, codim, qp, pitype=All_Partition):
        """e.geometry() of the entities of `codim` at local points qp (nqp x (dim-codim)): global (n, nq, d), jt (n, nq, m, d),
        jit (n, nq, d, m), detJ (n, nq)"""
        d, m = self.dim, self.dim - codim
        qp = np.ascontiguousarray(qp, dtype=np.float64).reshape(-1, max(1, m))
        nq = qp.shape[0]
        n = self._chk(self.L.orc_entities(self.h, rank, level, codim, pitype, None, None, None, None, None, None, None))
        r = {"global": np.zeros((n, nq, d)), "jt": np.zeros((n, nq, m, d)), "jit": np.zeros((n, nq, d, m)), "detJ": np.zeros((n, nq))}
        self._chk(self.L.orc_geometry_eval_codim(self.h, rank, level, codim, pitype, nq, _p(qp, C.c_double), _p(r["global"], C.c_double),
                                                 _p(r["jt"], C.c_double), _p(r["jit"], C.c_double), _p(r["detJ"], C.c_double)))
        return r

    def geogrid_eval(self, rank, level, deformation, params, qp, pitype=All_Partition):
        qp = np.ascontiguousarray(qp, dtype=np.float64).reshape(-1, self.dim)
        params = np.ascontiguousarray(params, dtype=np.float64)
        n, nq, d = self.ncells(rank, level, pitype), qp.shape[0], self.dim
        r = {"corners": np.zeros((n, 1 << d, d)), "global": np.zeros((n, nq, d)), "jt": np.zeros((n, nq, d, d)),
             "jit": np.zeros((n, nq, d, d)), "detJ": np.zeros((n, nq))}
        self._chk(self.L.orc_geogrid_eval(self.h, rank, level, pitype, deformation, _p(params, C.c_double), nq, _p(qp, C.c_double),
                                          _p(r["corners"], C.c_double), _p(r["global"], C.c_double), _p(r["jt"], C.c_double),
                                          _p(r["jit"], C.c_double), _p(r["detJ"], C.c_double)))
        return r

    def geogrid_intersections(self, rank, level, deformation, params, qp_face, pitype=All_Partition):
        """GeoGrid::Intersection of every face of every cell: corners (n, 2d, 2^(d-1), d), center (n, 2d, d), volume (n, 2d),
        integrationOuterNormal / outerNormal / unitOuterNormal (n, 2d, nqp, d) at the face-local points, centerUnitOuterNormal (n, 2d, d)"""
        d = self.dim
        qp = np.ascontiguousarray(qp_face, dtype=np.float64).reshape(-1, max(1, d - 1))
        params = np.ascontiguousarray(params, dtype=np.float64)
        n, nq = self.ncells(rank, level, pitype), qp.shape[0]
        r = {"corners": np.zeros((n, 2 * d, 1 << (d - 1), d)), "center": np.zeros((n, 2 * d, d)), "volume": np.zeros((n, 2 * d)),
             "integrationOuterNormal": np.zeros((n, 2 * d, nq, d)), "outerNormal": np.zeros((n, 2 * d, nq, d)),
             "unitOuterNormal": np.zeros((n, 2 * d, nq, d)), "centerUnitOuterNormal": np.zeros((n, 2 * d, d))}
        self._chk(self.L.orc_geogrid_intersections(self.h, rank, level, pitype, deformation, _p(params, C.c_double), nq, _p(qp, C.c_double),
                                                   _p(r["corners"], C.c_double), _p(r["center"], C.c_double), _p(r["volume"], C.c_double),
                                                   _p(r["integrationOuterNormal"], C.c_double), _p(r["outerNormal"], C.c_double),
                                                   _p(r["unitOuterNormal"], C.c_double), _p(r["centerUnitOuterNormal"], C.c_double)))
        return r

    def _pp(self, per_rank_arrays):
        """per_rank_arrays[rank][a] -> double** laid out [rank*narrays + a]"""
        flat = [a for r in per_rank_arrays for a in r]
        for a in flat:
            assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
        arr = (C.POINTER(C.c_double) * len(flat))(*[a.ctypes.data_as(C.POINTER(C.c_double)) for a in flat])
        return arr

    def communicate(self, level, block, iftype, direction, op, per_rank_arrays, item_sizes):
        na = len(item_sizes)
        arr = self._pp(per_rank_arrays)
        items = (C.c_int * na)(*item_sizes)
        return self._chk(self.L.orc_communicate(self.h, level, _block(block), iftype, direction, op, na, arr, items))

    def communicate_variable(self, level, block, iftype, direction, per_rank_offsets, per_rank_data):
        """variable-size data handle on all ranks: returns per rank (index uint32[k], n int64[k], items float64[...]) = the
        scatter(buff, e, n) calls in the reference's order"""
        P = len(per_rank_offsets)
        offs = [np.ascontiguousarray(o, dtype=np.int64) for o in per_rank_offsets]
        dats = [np.ascontiguousarray(d, dtype=np.float64) for d in per_rank_data]
        po = (C.POINTER(C.c_int64) * P)(*[o.ctypes.data_as(C.POINTER(C.c_int64)) for o in offs])
        pdat = (C.POINTER(C.c_double) * P)(*[d.ctypes.data_as(C.POINTER(C.c_double)) for d in dats])
        nc, ni = (C.c_int64 * P)(), (C.c_int64 * P)()
        self._chk(self.L.orc_communicate_variable(self.h, level, _block(block), iftype, direction, po, pdat, nc, ni, None, None, None))
        idx = [np.zeros(nc[p], dtype=np.uint32) for p in range(P)]
        n = [np.zeros(nc[p], dtype=np.int64) for p in range(P)]
        items = [np.zeros(ni[p], dtype=np.float64) for p in range(P)]
        pi_ = (C.POINTER(C.c_uint32) * P)(*[a.ctypes.data_as(C.POINTER(C.c_uint32)) for a in idx])
        pn = (C.POINTER(C.c_int64) * P)(*[a.ctypes.data_as(C.POINTER(C.c_int64)) for a in n])
        pit = (C.POINTER(C.c_double) * P)(*[a.ctypes.data_as(C.POINTER(C.c_double)) for a in items])
        self._chk(self.L.orc_communicate_variable(self.h, level, _block(block), iftype, direction, po, pdat, nc, ni, pi_, pn, pit))
        return [(idx[p], n[p], items[p]) for p in range(P)]

    def fv_init(self, rank, level, problem, params):
        params = np.ascontiguousarray(params, dtype=np.float64)
        c = np.zeros(self.size(rank, level, 0))
        self._chk(self.L.orc_fv_init(self.h, rank, level, problem, _p(params, C.c_double), _p(c, C.c_double)))
        return c

    def fv_step(self, level, problem, params, c_per_rank, want_update=False):
        params = np.ascontiguousarray(params, dtype=np.float64)
        cs = self._pp([[c] for c in c_per_rank])
        upd = [np.zeros_like(c) for c in c_per_rank] if want_update else None
        us = self._pp([[u] for u in upd]) if want_update else None
        dt = self.L.orc_fv_step(self.h, level, problem, _p(params, C.c_double), cs, us)
        if dt < 0:
            raise OracleError(self.o.err())
        return (dt, upd) if want_update else dt

    def fv_run(self, level, problem, params, c_per_rank, steps):
        params = np.ascontiguousarray(params, dtype=np.float64)
        cs = self._pp([[c] for c in c_per_rank])
        dt = C.c_double(0)
        secs = self.L.orc_fv_run(self.h, level, problem, _p(params, C.c_double), cs, steps, C.byref(dt))
        if secs < 0:
            raise OracleError(self.o.err())
        return secs, dt.value


_cache = {}


def load(kind="ref"):
    if kind not in _cache:
        if not os.path.exists(LIBS[kind]):
            raise FileNotFoundError(f"{LIBS[kind]} not built: run `make -C oracle` (or __graft_entry__.build())")
        _cache[kind] = OracleLib(LIBS[kind])
    return _cache[kind]


def available(kind):
    return os.path.exists(LIBS[kind])
