"""Host-side mirror of the reference's operator surface for the YaspGrid leaf-sweep path, on top of the
C ABI (include/dgb.h). Names follow dune-grid: YaspGrid(...).leafGridView()/levelGridView(l), GridView.size,
.elements(), .intersections(), .communicate(), MultipleCodimMultipleGeomTypeMapper(gv, layout).index()/size(),
geometry().global_()/jacobianInverseTransposed()/integrationElement(). Per-entity calls of the reference
(dune/grid/common/gridview.hh, entity.hh, intersection.hh, geometry.hh, mcmgmapper.hh) become per-view bulk
calls that return CUDA tensors in the reference's iteration order. torch is only the allocator / stream
provider here; all arithmetic happens in libdgb.so's CUDA kernels."""
import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import capi
from .capi import check, lib

# dune/grid/common/gridenums.hh
(Interior_Partition, InteriorBorder_Partition, Overlap_Partition, OverlapFront_Partition, All_Partition,
 Ghost_Partition) = range(6)
InteriorEntity, BorderEntity, OverlapEntity, FrontEntity, GhostEntity = range(5)
(InteriorBorder_InteriorBorder_Interface, InteriorBorder_All_Interface, Overlap_OverlapFront_Interface,
 Overlap_All_Interface, All_All_Interface) = range(5)
ForwardCommunication, BackwardCommunication = 0, 1


# compiled-in sweep functors (include/dgb.h DGB_FUNCTOR_*)
FUNCTORS = {"midpoint_exp_norm": 0, "gauss5_exp_norm": 1, "boundary_flux_x": 2, "measure": 3}


def functors():
    """the registry as the library reports it: {id: (name, has_element_body, has_intersection_body)}"""
    n = C.c_int(0)
    check(lib.dgb_functor_count(C.byref(n)))
    out = {}
    for i in range(64):
        name, e, f = C.c_char_p(), C.c_int(0), C.c_int(0)
        if lib.dgb_functor_info(i, C.byref(name), C.byref(e), C.byref(f)) == 0:
            out[i] = (name.value.decode(), bool(e.value), bool(f.value))
        if len(out) == n.value:
            break
    return out


def _torch():
    import torch
    return torch


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    torch = _torch()
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _require_cuda():
    torch = _torch()
    if not torch.cuda.is_available():
        raise capi.DgbError(capi.ECUDA, "no CUDA device: dune_grid_b200 has no CPU fallback")


class YaspGrid:
    """Dune::YaspGrid<dim, Coordinates> (dune/grid/yaspgrid.hh:164). coordinates: 'equidistant'
    (L=upper), 'equidistantoffset' (lower, upper) or 'tensorproduct' (coords = list of arrays)."""

    def __init__(self, dim=None, N=None, upper=None, lower=None, coords=None, periodic=0, overlap=1,
                 coordinates="equidistant", rank=0, nranks=1, partitioner="default", fixed_dims=None, comm=None):
        mode = {"equidistant": 0, "equidistantoffset": 1, "tensorproduct": 2}[coordinates]
        s = capi.GridSpec()
        self._keep = []
        if mode == 2:
            dim = len(coords)
            N = [len(c) - 1 for c in coords]
            for i, c in enumerate(coords):
                a = np.ascontiguousarray(c, dtype=np.float64)
                self._keep.append(a)
                s.tensor[i] = a.ctypes.data_as(C.POINTER(C.c_double))
        s.dim, s.coord_mode = dim, mode
        if isinstance(periodic, (list, tuple)):
            periodic = sum(1 << i for i, p in enumerate(periodic) if p)
        for i in range(3):
            s.N[i] = N[i] if i < dim else 1
            s.lower[i] = (lower[i] if lower is not None and i < dim else 0.0)
            s.upper[i] = (upper[i] if upper is not None and i < dim else 1.0)
            s.fixed_dims[i] = fixed_dims[i] if fixed_dims is not None and i < dim else 1
        s.periodic, s.overlap = periodic, overlap
        s.partitioner = {"default": 0, "powerd": 1, "fixed": 2}[partitioner]
        self.dim, self.rank, self.nranks, self.comm = dim, rank, nranks, comm
        self.spec = s
        self.h = C.c_void_p()
        check(lib.dgb_grid_create(C.byref(s), rank, nranks, C.byref(self.h)))

    def __del__(self):
        try:
            if getattr(self, "h", None):
                lib.dgb_grid_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # ---- Dune::BackupRestoreFacility<YaspGrid> (dune/grid/yaspgrid/backuprestore.hh) ----
    def backup(self, filename=None):
        """the text BackupRestoreFacility::backup writes; with a filename also written like the reference does: one file
        from rank 0 (equidistant containers, :91-104) or `filename + rank` from every rank (tensor product, :208-220)"""
        need = C.c_int64(0)
        check(lib.dgb_grid_backup(self.h, None, 0, C.byref(need)))
        buf = C.create_string_buffer(need.value)
        check(lib.dgb_grid_backup(self.h, buf, need.value, C.byref(need)))
        text = buf.value.decode()
        if filename is not None:
            if self.spec.coord_mode == 2:
                with open(f"{filename}{self.rank}", "w") as f:
                    f.write(text)
            elif self.rank == 0:
                with open(filename, "w") as f:
                    f.write(text)
        return text

    @classmethod
    def restore(cls, text_or_filename, coordinates="equidistant", rank=0, nranks=1, comm=None, is_file=False):
        """BackupRestoreFacility::restore (:131-218, :249-346); tensor-product files are per rank (`filename + rank`)"""
        mode = {"equidistant": 0, "equidistantoffset": 1, "tensorproduct": 2}[coordinates]
        text = text_or_filename
        if is_file:
            with open(f"{text_or_filename}{rank}" if mode == 2 else text_or_filename) as f:
                text = f.read()
        self = cls.__new__(cls)
        self._keep = []
        self.h = C.c_void_p()
        check(lib.dgb_grid_restore(text.encode(), mode, rank, nranks, C.byref(self.h)))
        self.rank, self.nranks, self.comm = rank, nranks, comm
        s = capi.GridSpec()
        s.coord_mode = mode
        self.spec = s
        v = capi.View()
        check(lib.dgb_view_level(self.h, 0, C.byref(v)))
        self.dim = s.dim = v.dim
        return self

    def refineOptions(self, keepPhysicalOverlap: bool):
        check(lib.dgb_grid_refine_options(self.h, int(keepPhysicalOverlap)))

    def globalRefine(self, refCount: int):
        check(lib.dgb_grid_global_refine(self.h, refCount))

    @property
    def maxLevel(self):
        m = C.c_int(0)
        check(lib.dgb_grid_max_level(self.h, C.byref(m)))
        return m.value

    def torusDims(self):
        d = (C.c_int32 * 3)(1, 1, 1)
        check(lib.dgb_grid_torus_dims(self.h, d))
        return tuple(d)[:self.dim]

    def numBoundarySegments(self):
        n = C.c_int64(0)
        check(lib.dgb_grid_num_boundary_segments(self.h, C.byref(n)))
        return n.value

    def domainSize(self):
        d = (C.c_double * 3)()
        check(lib.dgb_grid_domain_size(self.h, d))
        return np.array(d[:self.dim])

    def leafGridView(self):
        return GridView(self, None)

    def levelGridView(self, level):
        return GridView(self, level)

    def lastBytesSent(self):
        """bytes this rank sent to OTHER ranks in its most recent exchange (dgb_communicate or the FV step)"""
        n = C.c_int64(0)
        check(lib.dgb_comm_last_bytes(self.h, C.byref(n)))
        return n.value

    def commList(self, level, codim, which):
        cap = 4096
        out = (C.c_int32 * (cap * 10))()
        n = C.c_int(0)
        check(lib.dgb_view_comm_list(self.h, level, codim, which, out, cap, C.byref(n)))
        return np.ctypeslib.as_array(out)[:n.value * 10].reshape(n.value, 10).copy()


class GridView:
    """Dune::GridView of a YaspGrid level (leaf == finest level; dune/grid/common/gridview.hh:65-398)."""

    def __init__(self, grid: YaspGrid, level: Optional[int]):
        self.grid = grid
        self.v = capi.View()
        if level is None:
            check(lib.dgb_view_leaf(grid.h, C.byref(self.v)))
        else:
            check(lib.dgb_view_level(grid.h, level, C.byref(self.v)))
        self.dim, self.level = self.v.dim, self.v.level

    def size(self, codim, partition=All_Partition):
        n = C.c_int64(0)
        check(lib.dgb_view_partition_size(C.byref(self.v), codim, partition, C.byref(n)))
        return n.value

    def overlapSize(self, codim=0):
        n = C.c_int(0)
        check(lib.dgb_view_overlap_size(C.byref(self.v), codim, C.byref(n)))
        return n.value

    def ghostSize(self, codim=0):
        return 0

    def boxes(self, codim):
        """host view of the nested partition boxes (for tests): list of dicts per shift component"""
        v, res = self.v, []
        for j in range(v.comp_begin[codim], v.comp_begin[codim + 1]):
            c = v.comp[j]
            d = {"shift": c.shift, "offset": c.index_offset[0]}
            for q, name in enumerate(["overlapfront", "overlap", "interiorborder", "interior"]):
                d[name] = (tuple(c.box[q].origin), tuple(c.box[q].size))
            res.append(d)
        return res

    def coordinates(self, axis):
        """vertex coordinates along `axis` over the stored vertex range (host arithmetic of the descriptor)"""
        v = self.v
        b = v.comp[v.comp_begin[self.dim]].box[0]
        first, n = b.origin[axis], b.size[axis]
        idx = np.arange(first, first + n)
        if v.coord_mode == 0:
            return first, idx * v.h[axis]
        if v.coord_mode == 1:
            return first, v.origin[axis] + idx * v.h[axis]
        arr = np.ctypeslib.as_array(v.tensor_host[axis], shape=(v.tensor_size[axis],))
        return first, arr[idx - v.tensor_offset[axis]].copy()

    # ---- bulk sweeps -----------------------------------------------------------------------------------------
    def elements(self, codim=0, partition=All_Partition, geometry=True, coord=False):
        """for (e : elements(gv, partition)) -> dict of CUDA tensors: index, partitionType, [coord], lower, upper,
        center (dim x n), volume"""
        _require_cuda()
        torch = _torch()
        n, d = self.size(codim, partition), self.dim
        dev = "cuda"
        r = {"index": torch.empty(n, dtype=torch.int32, device=dev), "partitionType": torch.empty(n, dtype=torch.uint8, device=dev)}
        if coord:
            r["coord"] = torch.empty((d, n), dtype=torch.int32, device=dev)
        if geometry:
            for k in ("lower", "upper", "center"):
                r[k] = torch.empty((d, n), dtype=torch.float64, device=dev)
            r["volume"] = torch.empty(n, dtype=torch.float64, device=dev)
        check(lib.dgb_elements_materialize(C.byref(self.v), codim, partition, _ptr(r["index"]), _ptr(r["partitionType"]),
                                           _ptr(r.get("coord")), _ptr(r.get("lower")), _ptr(r.get("upper")),
                                           _ptr(r.get("center")), _ptr(r.get("volume")), _stream()))
        return r

    def intersections(self, partition=All_Partition, geometry=True):
        """for (is : intersections(gv, e)) for all cells: neighbor, boundary, outside (index or -1),
        boundarySegmentIndex (or -1): (2dim x n); center (2dim x dim x n), volume (2dim x n)"""
        _require_cuda()
        torch = _torch()
        n, d = self.size(0, partition), self.dim
        dev = "cuda"
        r = {"neighbor": torch.empty((2 * d, n), dtype=torch.uint8, device=dev),
             "boundary": torch.empty((2 * d, n), dtype=torch.uint8, device=dev),
             "outside": torch.empty((2 * d, n), dtype=torch.int32, device=dev),
             "boundarySegmentIndex": torch.empty((2 * d, n), dtype=torch.int32, device=dev)}
        if geometry:
            r["center"] = torch.empty((2 * d, d, n), dtype=torch.float64, device=dev)
            r["volume"] = torch.empty((2 * d, n), dtype=torch.float64, device=dev)
        check(lib.dgb_intersections_materialize(C.byref(self.v), partition, _ptr(r["neighbor"]), _ptr(r["boundary"]),
                                                _ptr(r["outside"]), _ptr(r["boundarySegmentIndex"]),
                                                _ptr(r.get("center")), _ptr(r.get("volume")), _stream()))
        return r

    # ---- functor sweeps (compiled-in device functors, csrc/dgb_functors.cuh) --------------------------------------
    def _for_each(self, faces, functor, args, partition, values):
        _require_cuda()
        torch = _torch()
        fid = FUNCTORS[functor] if isinstance(functor, str) else int(functor)
        a = np.ascontiguousarray(args if args is not None else [], dtype=np.float64)
        n = self.size(0, partition)
        per = torch.empty((2 * self.dim, n) if faces else (n,), dtype=torch.float64, device="cuda") if values else None
        total = torch.zeros(1, dtype=torch.float64, device="cuda")
        f = lib.dgb_for_each_intersection if faces else lib.dgb_for_each_element
        check(f(C.byref(self.v), partition, fid, a.ctypes.data_as(C.POINTER(C.c_double)), a.size, _ptr(per), _ptr(total), _stream()))
        return (total, per) if values else total

    def forEachElement(self, functor, args=None, partition=All_Partition, values=False):
        """sum over `for e in elements(gv, partition)` of the compiled-in functor's element body, as ONE fused kernel; returns a
        1-element CUDA tensor (and the per-element values with values=True). functor: id or a name of FUNCTORS"""
        return self._for_each(False, functor, args, partition, values)

    def forEachIntersection(self, functor, args=None, partition=All_Partition, values=False):
        """sum over `for e in elements(gv): for is in intersections(gv, e)` of the functor's intersection body (values: (2dim, n))"""
        return self._for_each(True, functor, args, partition, values)

    def intersectionNormals(self, partition=All_Partition):
        """unitOuterNormal (== outerNormal == centerUnitOuterNormal) and integrationOuterNormal of every face of every cell:
        two CUDA tensors of shape (2dim, dim, n)"""
        _require_cuda()
        torch = _torch()
        n, d = self.size(0, partition), self.dim
        u = torch.empty((2 * d, d, n), dtype=torch.float64, device="cuda")
        w = torch.empty((2 * d, d, n), dtype=torch.float64, device="cuda")
        check(lib.dgb_intersection_normals(C.byref(self.v), partition, _ptr(u), _ptr(w), _stream()))
        return u, w

    def geometryInInside(self, k):
        """(lower, upper) of YaspIntersection::geometryInInside() of face k (the same for every cell)"""
        lo, hi = (C.c_double * 3)(), (C.c_double * 3)()
        check(lib.dgb_intersection_local_geometry(self.dim, k, 0, lo, hi))
        return np.array(lo[:self.dim]), np.array(hi[:self.dim])

    def geometryInOutside(self, k):
        lo, hi = (C.c_double * 3)(), (C.c_double * 3)()
        check(lib.dgb_intersection_local_geometry(self.dim, k, 1, lo, hi))
        return np.array(lo[:self.dim]), np.array(hi[:self.dim])

    def centerUnitOuterNormal(self, k):
        """YaspIntersection::centerUnitOuterNormal (yaspgridintersection.hh:177-182): +-e_dir, independent of the cell"""
        n = np.zeros(self.dim)
        n[k // 2] = 1.0 if k % 2 else -1.0
        return n

    def geometry(self, qp, partition=All_Partition):
        """Geometry::global / jacobianInverseTransposed / integrationElement at local points qp (nqp x dim)"""
        _require_cuda()
        torch = _torch()
        qp = np.ascontiguousarray(qp, dtype=np.float64).reshape(-1, self.dim)
        n, d, nq = self.size(0, partition), self.dim, qp.shape[0]
        r = {"global": torch.empty((nq, d, n), dtype=torch.float64, device="cuda"),
             "jacobianInverseTransposed": torch.empty((nq, d, d, n), dtype=torch.float64, device="cuda"),
             "integrationElement": torch.empty((nq, n), dtype=torch.float64, device="cuda")}
        check(lib.dgb_geometry_eval(C.byref(self.v), partition, nq, qp.ctypes.data_as(C.POINTER(C.c_double)), _ptr(r["global"]),
                                    _ptr(r["jacobianInverseTransposed"]), _ptr(r["integrationElement"]), _stream()))
        return r

    def geometryCodim(self, codim, qp, partition=All_Partition):
        """e.geometry() of the entities of `codim` at local points qp (nqp x (dim-codim)): global (nq, d, n), jacobianTransposed
        (nq, m, d, n), jacobianInverseTransposed (nq, d, m, n), integrationElement (nq, n)"""
        _require_cuda()
        torch = _torch()
        d, m = self.dim, self.dim - codim
        qp = np.ascontiguousarray(qp, dtype=np.float64).reshape(-1, max(1, m))
        n, nq = self.size(codim, partition), qp.shape[0]
        z = lambda *shape: torch.empty(shape, dtype=torch.float64, device="cuda")       # noqa: E731
        r = {"global": z(nq, d, n), "jacobianTransposed": z(nq, m, d, n), "jacobianInverseTransposed": z(nq, d, m, n), "integrationElement": z(nq, n)}
        check(lib.dgb_geometry_eval_codim(C.byref(self.v), codim, partition, nq, qp.ctypes.data_as(C.POINTER(C.c_double)), _ptr(r["global"]),
                                          _ptr(r["jacobianTransposed"]), _ptr(r["jacobianInverseTransposed"]), _ptr(r["integrationElement"]), _stream()))
        return r

    def geometryGrid(self, deformation, params, qp, partition=All_Partition, corners=True):
        """GeometryGrid<YaspGrid, F> geometries: corners (2^dim x dim x n), global, jacobianTransposed,
        jacobianInverseTransposed, integrationElement at local points qp"""
        _require_cuda()
        torch = _torch()
        qp = np.ascontiguousarray(qp, dtype=np.float64).reshape(-1, self.dim)
        params = np.ascontiguousarray(params, dtype=np.float64)
        n, d, nq = self.size(0, partition), self.dim, qp.shape[0]
        r = {"global": torch.empty((nq, d, n), dtype=torch.float64, device="cuda"),
             "jacobianTransposed": torch.empty((nq, d, d, n), dtype=torch.float64, device="cuda"),
             "jacobianInverseTransposed": torch.empty((nq, d, d, n), dtype=torch.float64, device="cuda"),
             "integrationElement": torch.empty((nq, n), dtype=torch.float64, device="cuda")}
        if corners:
            r["corners"] = torch.empty((1 << d, d, n), dtype=torch.float64, device="cuda")
        check(lib.dgb_geogrid_eval(C.byref(self.v), partition, deformation, params.ctypes.data_as(C.POINTER(C.c_double)), params.size,
                                   nq, qp.ctypes.data_as(C.POINTER(C.c_double)), _ptr(r.get("corners")), _ptr(r["global"]),
                                   _ptr(r["jacobianTransposed"]), _ptr(r["jacobianInverseTransposed"]),
                                   _ptr(r["integrationElement"]), _stream()))
        return r

    def geometryGridIntersections(self, deformation, params, qp_face, partition=All_Partition, corners=True):
        """GeometryGrid<YaspGrid, F> intersections of every cell (GeoGrid::Intersection): corners (2dim, 2^(dim-1), dim, n), center
        (2dim, dim, n), volume (2dim, n); at the face-local points qp_face (nqp x (dim-1)): integrationOuterNormal, outerNormal,
        unitOuterNormal (2dim, nqp, dim, n); centerUnitOuterNormal (2dim, dim, n)"""
        _require_cuda()
        torch = _torch()
        d = self.dim
        qp = np.ascontiguousarray(qp_face, dtype=np.float64).reshape(-1, max(1, d - 1))
        params = np.ascontiguousarray(params, dtype=np.float64)
        n, nq = self.size(0, partition), qp.shape[0]
        z = lambda *shape: torch.empty(shape, dtype=torch.float64, device="cuda")       # noqa: E731
        r = {"center": z(2 * d, d, n), "volume": z(2 * d, n), "integrationOuterNormal": z(2 * d, nq, d, n), "outerNormal": z(2 * d, nq, d, n),
             "unitOuterNormal": z(2 * d, nq, d, n), "centerUnitOuterNormal": z(2 * d, d, n)}
        if corners:
            r["corners"] = z(2 * d, 1 << (d - 1), d, n)
        check(lib.dgb_geogrid_intersections(C.byref(self.v), partition, deformation, params.ctypes.data_as(C.POINTER(C.c_double)), params.size,
                                            nq, qp.ctypes.data_as(C.POINTER(C.c_double)), _ptr(r.get("corners")), _ptr(r["center"]), _ptr(r["volume"]),
                                            _ptr(r["integrationOuterNormal"]), _ptr(r["outerNormal"]), _ptr(r["unitOuterNormal"]),
                                            _ptr(r["centerUnitOuterNormal"]), _stream()))
        return r

    def geometryGridVolume(self, deformation, params, qp, weights, partition=All_Partition):
        _require_cuda()
        torch = _torch()
        qp = np.ascontiguousarray(qp, dtype=np.float64).reshape(-1, self.dim)
        w = np.ascontiguousarray(weights, dtype=np.float64)
        params = np.ascontiguousarray(params, dtype=np.float64)
        out = torch.zeros(1, dtype=torch.float64, device="cuda")
        check(lib.dgb_geogrid_volume(C.byref(self.v), partition, deformation, params.ctypes.data_as(C.POINTER(C.c_double)), params.size,
                                     qp.shape[0], qp.ctypes.data_as(C.POINTER(C.c_double)), w.ctypes.data_as(C.POINTER(C.c_double)),
                                     _ptr(out), _stream()))
        return out

    def writeVTK(self, name, celldata=None, pointdata=None, path="", float64=False, outputType="ascii"):
        """VTKWriter(gridView).addCellData/addVertexData + write(name) (dune/grid/io/file/vtk/vtkwriter.hh; dune-python:
        gridView.writeVTK(name, celldata={...}, pointdata={...})). Values: CUDA tensors indexed by the element / vertex index,
        shape (n,) scalar, (n, k) with k <= 3 vector. Returns the file to open (.vtu, or the .pvtu header on several ranks)."""
        def pack(d, nent, kind):
            arr = (capi.VtkField * max(1, len(d or {})))()
            keep = []
            for k, (nm, t) in enumerate((d or {}).items()):
                # dgb_vtk_write reads ncomp*entities DOUBLES from DEVICE memory: convert / validate instead of passing garbage
                torch = _torch()
                if not isinstance(t, torch.Tensor):
                    t = torch.as_tensor(np.asarray(t))
                t = t.to(device="cuda", dtype=torch.float64).contiguous()
                if t.dim() not in (1, 2) or (t.dim() == 2 and not 1 <= t.shape[1] <= 3):
                    raise ValueError(f"writeVTK: field '{nm}' must have shape (n,) or (n, k) with k <= 3, got {tuple(t.shape)}")
                if t.shape[0] != nent:
                    raise ValueError(f"writeVTK: field '{nm}' has {t.shape[0]} entries, the view has {nent} {kind}")
                keep.append(t)
                arr[k].name = nm.encode()
                arr[k].d_data = t.data_ptr()
                arr[k].ncomp = 1 if t.dim() == 1 else int(t.shape[1])
                arr[k].is_vector = 0 if t.dim() == 1 else 1
                arr[k].precision = 1 if float64 else 0
            return arr, len(d or {}), keep
        if celldata or pointdata:
            _require_cuda()
        ca, nc, k1 = pack(celldata, self.size(0), "cells")
        va, nv, k2 = pack(pointdata, self.size(self.dim), "vertices")
        out = C.create_string_buffer(4096)
        ot = {"ascii": 0, "base64": 1, "appendedraw": 2, "appendedbase64": 3}[outputType]        # VTK::OutputType (io/file/vtk/common.hh:43-51)
        check(lib.dgb_vtk_write_ex(self.grid.h, self.level, str(name).encode(), str(path).encode(), ca, nc, va, nv, 1 if float64 else 0,
                                   ot, out, 4096, _stream() if (nc or nv) else None))
        return out.value.decode()

    def globalIndexSet(self, codim):
        """Dune::GlobalIndexSet(gridView, codim) (dune/grid/utility/globalindexset.hh): (int32 tensor with index(e) for every
        entity of `codim` in local index order, size(codim) over all ranks). Collective."""
        _require_cuda()
        torch = _torch()
        out = torch.empty(self.size(codim), dtype=torch.int32, device="cuda")
        n = C.c_int64(0)
        comm = self.grid.comm.h if self.grid.comm is not None else None
        check(lib.dgb_global_index(self.grid.h, self.level, codim, _ptr(out), C.byref(n), comm, _stream()))
        return out, n.value

    def communicate(self, mapper, arrays, iftype, direction, op="set", item_sizes=None):
        """gv.communicate(dataHandle, iftype, dir) with the data handle 'arrays indexed by mapper' and the
        combine operation op in {'set','add'} (dune/python/grid/mapper.hh:117-137)"""
        _require_cuda()
        arrays = list(arrays)
        if item_sizes is None:
            item_sizes = [int(a.numel() // max(1, mapper.size)) for a in arrays]
        ptrs = (C.c_void_p * len(arrays))(*[a.data_ptr() for a in arrays])
        items = (C.c_int32 * len(arrays))(*item_sizes)
        comm = self.grid.comm.h if self.grid.comm is not None else None
        check(lib.dgb_communicate(self.grid.h, self.level, C.byref(mapper.m), iftype, direction, {"set": 0, "add": 1, "min_nonneg": 2, "set_nonneg": 3}[op],
                                  ptrs, items, len(arrays), comm, _stream()))


    def communicateVariable(self, mapper, offsets, data, iftype, direction):
        """gv.communicate(dataHandle, iftype, dir) for a data handle with fixedSize() == false (yaspgrid.hh:1562-1634):
        size(e) = offsets[i+1]-offsets[i] and gather(e) = data[offsets[i]:offsets[i+1]] with i = mapper.index(e) (int64 / float64
        CUDA tensors). Returns the scatter(buff, e, n) calls of the reference in its order as (index, recv_offsets, recv_data)."""
        ex = VariableExchange(self, mapper, offsets, data, iftype, direction)
        try:
            ex.exchange()
            return ex.result()
        finally:
            ex.close()

    # ---- the phases of communicate, for other transports and single-device rank emulation -----------------
    def _items(self, mapper, arrays, item_sizes):
        arrays = list(arrays)
        if item_sizes is None:
            item_sizes = [int(a.numel() // max(1, mapper.size)) for a in arrays]
        ptrs = (C.c_void_p * len(arrays))(*[a.data_ptr() for a in arrays])
        return arrays, ptrs, (C.c_int32 * len(arrays))(*item_sizes)

    def commPlan(self, mapper, iftype, direction, item_sizes):
        """(send segments, recv segments), each a list of (peer, offset, count) in doubles"""
        items = (C.c_int32 * len(item_sizes))(*item_sizes)
        out = (C.c_int64 * 1024)()
        check(lib.dgb_comm_plan_info(self.grid.h, self.level, C.byref(mapper.m), iftype, direction, items, len(item_sizes), out, 1024))
        ns, nr = out[0], out[1]
        seg = [(int(out[2 + 3 * k]), int(out[3 + 3 * k]), int(out[4 + 3 * k])) for k in range(ns + nr)]
        return seg[:ns], seg[ns:]

    def commPack(self, mapper, arrays, iftype, direction, item_sizes=None):
        """returns (sendbuf_ptr, recvbuf_ptr) device addresses of the library-owned exchange buffers"""
        _require_cuda()
        arrays, ptrs, items = self._items(mapper, arrays, item_sizes)
        sb, rb = C.c_void_p(), C.c_void_p()
        check(lib.dgb_comm_pack(self.grid.h, self.level, C.byref(mapper.m), iftype, direction, ptrs, items, len(arrays),
                                C.byref(sb), C.byref(rb), _stream()))
        return sb.value or 0, rb.value or 0

    def commUnpack(self, mapper, arrays, iftype, direction, op="set", item_sizes=None):
        _require_cuda()
        arrays, ptrs, items = self._items(mapper, arrays, item_sizes)
        check(lib.dgb_comm_unpack(self.grid.h, self.level, C.byref(mapper.m), iftype, direction, {"set": 0, "add": 1, "min_nonneg": 2, "set_nonneg": 3}[op],
                                  ptrs, items, len(arrays), _stream()))


class MultipleCodimMultipleGeomTypeMapper:
    """Dune::MultipleCodimMultipleGeomTypeMapper<GV> (dune/grid/common/mcmgmapper.hh). layout = block size per
    codimension (Yasp has one geometry type per codim)."""

    def __init__(self, gv: GridView, layout: Sequence[int]):
        self.gv = gv
        b = list(layout) + [0] * (4 - len(layout))
        self.m = capi.Mapper()
        check(lib.dgb_mapper_init(C.byref(gv.v), (C.c_uint32 * 4)(*b), C.byref(self.m)))

    @property
    def size(self):
        return int(self.m.size)

    def index(self, codim=0, partition=All_Partition, out=None):
        """mapper.index(e) of every entity of the range, in iteration order (mcmgmapper.hh:212-226). `out`: caller-owned int32
        CUDA tensor (any 4-byte alignment; 16-byte aligned buffers take the vectorised kernel)."""
        _require_cuda()
        torch = _torch()
        n = self.gv.size(codim, partition)
        if out is None:
            out = torch.empty(n, dtype=torch.int32, device="cuda")
        elif out.dtype != torch.int32 or not out.is_cuda or not out.is_contiguous() or out.numel() != n:
            raise ValueError(f"index: out must be a contiguous int32 CUDA tensor with {n} entries")
        check(lib.dgb_mapper_index(C.byref(self.gv.v), C.byref(self.m), codim, partition, _ptr(out), _stream()))
        return out

    def subIndex(self, cc, partition=All_Partition, out=None):
        _require_cuda()
        torch = _torch()
        from math import comb
        ns = comb(self.gv.dim, cc) << cc
        n = self.gv.size(0, partition)
        if out is None:
            out = torch.empty((ns, n), dtype=torch.int32, device="cuda")
        elif out.dtype != torch.int32 or not out.is_cuda or not out.is_contiguous() or out.numel() != ns*n:
            raise ValueError(f"subIndex: out must be a contiguous int32 CUDA tensor with {ns}x{n} entries")
        check(lib.dgb_mapper_subindex(C.byref(self.gv.v), C.byref(self.m), cc, partition, _ptr(out), _stream()))
        return out

    def communicate(self, iftype, direction, op, *arrays):
        self.gv.communicate(self, arrays, iftype, direction, op)

    def backupField(self, array, filename, item_size=1):
        """restart checkpoint of a device array indexed by this mapper: file `filename.rank` (header, grid backup text, data)"""
        _require_cuda()
        check(lib.dgb_field_backup(self.gv.grid.h, self.gv.v.level, C.byref(self.m), _ptr(array), item_size,
                                   str(filename).encode(), _stream()))

    def restoreField(self, array, filename, item_size=1):
        """inverse of backupField; refuses files written for another grid, rank, level, mapper or item size"""
        _require_cuda()
        check(lib.dgb_field_restore(self.gv.grid.h, self.gv.v.level, C.byref(self.m), _ptr(array), item_size,
                                    str(filename).encode(), _stream()))


class VariableExchange:
    """dgb_commvar: one exchange of a variable-size data handle (dgb.h). exchange() runs both message rounds over NCCL; the
    phase methods let another transport (or a test that emulates several ranks on one device) move the segments itself."""

    def __init__(self, gv, mapper, offsets, data, iftype, direction):
        _require_cuda()
        torch = _torch()
        if offsets.dtype != torch.int64 or not offsets.is_cuda or offsets.numel() != mapper.size + 1:
            raise ValueError(f"offsets must be an int64 CUDA tensor with mapper.size+1 = {mapper.size + 1} entries")
        if data.dtype != torch.float64 or not data.is_cuda:
            raise ValueError("data must be a float64 CUDA tensor")
        self.gv, self.keep = gv, (offsets.contiguous(), data.contiguous())
        self.h = C.c_void_p()
        check(lib.dgb_commvar_create(gv.grid.h, gv.level, C.byref(mapper.m), iftype, direction, _ptr(self.keep[0]), _ptr(self.keep[1]),
                                     _stream(), C.byref(self.h)))

    def exchange(self):
        comm = self.gv.grid.comm.h if self.gv.grid.comm is not None else None
        check(lib.dgb_commvar_exchange(self.h, comm, _stream()))

    def segments(self, recv_side):
        """list of (peer, first entity slot, entity slots, first item, items)"""
        out = (C.c_int64 * 512)()
        check(lib.dgb_commvar_segments(self.h, 1 if recv_side else 0, out, 512))
        return [tuple(int(out[1 + 5 * k + j]) for j in range(5)) for k in range(out[0])]

    def buffers(self):
        """device addresses (send sizes, send items, recv sizes, recv items) of the library-owned buffers"""
        p = [C.c_void_p() for _ in range(4)]
        check(lib.dgb_commvar_buffers(self.h, *[C.byref(x) for x in p]))
        return tuple(x.value or 0 for x in p)

    def sizesReceived(self):
        check(lib.dgb_commvar_sizes_received(self.h, _stream()))

    def result(self):
        torch = _torch()
        ne, ni = C.c_int64(0), C.c_int64(0)
        check(lib.dgb_commvar_finish(self.h, _stream(), C.byref(ne), C.byref(ni)))
        index = torch.empty(ne.value, dtype=torch.int32, device="cuda")
        offsets = torch.empty(ne.value + 1, dtype=torch.int64, device="cuda")
        data = torch.empty(ni.value, dtype=torch.float64, device="cuda")
        check(lib.dgb_commvar_result(self.h, _ptr(index), _ptr(offsets), _ptr(data), _stream()))
        return index, offsets, data

    def close(self):
        if self.h:
            _torch().cuda.synchronize()
            lib.dgb_commvar_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class VTKSequenceWriter:
    """Dune::VTKSequenceWriter(gridView, name, path, extendpath) (io/file/vtk/vtksequencewriter.hh, vtksequencewriterbase.hh:105-151):
    one VTK file per write(time, ...) named `name-%05d` through the view's writer, and - on rank 0 - the ParaView collection
    `name.pvd` that lists every time step written so far (same text as the reference's). With a non-empty `extendpath` on several
    ranks the pieces AND the .pvtu header go to path/extendpath, which is where the collection points; the reference's pwrite
    puts the header into `path` (vtkwriter.hh:1079) while its collection names path/extendpath (vtksequencewriterbase.hh:140-141)."""

    def __init__(self, gv, name, path="", extendpath=""):
        self.gv, self.name, self.path, self.extendpath = gv, str(name), str(path), str(extendpath)
        self.timesteps = []

    def _seq(self, count):                                   # seqName, :190-196
        return f"{self.name}-{count:05d}"

    @staticmethod
    def _concat(base, p):                                    # the cases of concatPaths the writer meets
        if not base:
            return p
        if not p:
            return base
        return base + p if base.endswith("/") else base + "/" + p

    def getTimeSteps(self):
        return list(self.timesteps)

    def setTimeSteps(self, timesteps):
        """continue a sequence after a restart (:176-185)"""
        self.timesteps = [float(t) for t in timesteps]

    def write(self, time, celldata=None, pointdata=None, outputType="ascii", float64=False):
        g = self.gv.grid
        count = len(self.timesteps)
        self.timesteps.append(float(time))
        piecepath = self.path if g.nranks == 1 else self._concat(self.path, self.extendpath)
        written = self.gv.writeVTK(self._seq(count), celldata=celldata, pointdata=pointdata, path=piecepath, float64=float64, outputType=outputType)
        if g.rank == 0:
            lines = ['<?xml version="1.0"?> \n', '<VTKFile type="Collection" version="0.1" byte_order="LittleEndian"> \n', "<Collection> \n"]
            for i, t in enumerate(self.timesteps):
                if g.nranks == 1:
                    full = self._concat(piecepath, self._seq(i) + ".vtu")
                else:
                    full = self._concat(piecepath, f"s{g.nranks:04d}-{self._seq(i)}.pvtu")
                lines.append(f'<DataSet timestep="{_cxx_double(t)}" group="" part="0" name="" file="{full}"/> \n')
            lines += ["</Collection> \n", "</VTKFile> \n"]
            with open(self.name + ".pvd", "w") as f:
                f.write("".join(lines))
        return written


def _cxx_double(x):
    """operator<<(std::ostream&, double) with the default precision 6 (%g)"""
    return "%g" % x


def mcmgElementLayout(dim):
    return [1] + [0] * dim


def mcmgVertexLayout(dim):
    return [0] * dim + [1]


class Communicator:
    """NCCL communicator for the torus exchange (replaces the MPI communicator handed to YaspGrid)."""

    def __init__(self, rank, nranks, unique_id: bytes):
        self.h = C.c_void_p()
        buf = C.create_string_buffer(unique_id, 128)
        check(lib.dgb_comm_init(buf, rank, nranks, C.byref(self.h)))
        self.rank, self.nranks = rank, nranks

    @staticmethod
    def unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        check(lib.dgb_comm_unique_id(buf))
        return buf.raw

    @classmethod
    def from_torch_distributed(cls):
        """bootstrap over an initialised torch.distributed process group (plumbing only)"""
        torch = _torch()
        import torch.distributed as dist
        rank, n = dist.get_rank(), dist.get_world_size()
        ids = [cls.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        return cls(rank, n, ids[0])

    def close(self):
        if self.h:
            lib.dgb_comm_destroy(self.h)
            self.h = None


class FiniteVolume:
    """The explicit upwind finite-volume transport functor of SURVEY.md §8d on a level view."""

    def __init__(self, grid: YaspGrid, level=None, problem=0, params=(1, 1, 1, 0, .25, .25, .25, .125, .2), mode="separable"):
        _require_cuda()
        self.grid = grid
        level = grid.maxLevel if level is None else level
        p = np.ascontiguousarray(params, dtype=np.float64)
        self.h = C.c_void_p()
        comm = grid.comm.h if grid.comm is not None else None
        check(lib.dgb_fv_create(grid.h, level, problem, p.ctypes.data_as(C.POINTER(C.c_double)), p.size,
                                {"exact": 0, "separable": 1}[mode], comm, C.byref(self.h)))
        self.gv = grid.levelGridView(level)
        self.ncells = self.gv.size(0)

    def close(self):
        """dgb_fv_destroy: COLLECTIVE on a multi-rank grid (all ranks close their mappings of the peers' exchange blocks, meet in
        a barrier on the communicator, then free) - call it on every rank before Communicator.close()"""
        if getattr(self, "h", None):
            h, self.h = self.h, None
            check(lib.dgb_fv_destroy(h))

    def __del__(self):
        # garbage collection is not collective: free what is local, leave the exported exchange block to process exit
        try:
            if getattr(self, "h", None):
                (lib.dgb_fv_abandon if self.grid.nranks > 1 else lib.dgb_fv_destroy)(self.h)
                self.h = None
        except Exception:
            pass

    def kernel_info(self):
        """which kernel the last step ran: dict(kind='ring'|'ring_bulk'|'march'|'step2d', variant, tile, staging_bytes, zchunk, packs,
        staging='8'|'a16'|'pair'|'tma')"""
        a = (C.c_int32 * 8)()
        check(lib.dgb_fv_kernel_info(self.h, a))
        return {"kind": ("none", "step2d", "march", "ring", "ring_bulk")[a[0]], "variant": a[1], "tile": (a[2], a[3]),
                "staging_bytes": a[4], "zchunk": a[5], "packs": bool(a[6]),
                "staging": ("8", "a16", "pair", "tma")[a[7]] if a[0] >= 3 else "none"}

    def initial(self):
        torch = _torch()
        c = torch.empty(self.ncells, dtype=torch.float64, device="cuda")
        check(lib.dgb_fv_init_field(self.h, _ptr(c), _stream()))
        return c

    def step(self, c_in, c_out):
        check(lib.dgb_fv_step(self.h, _ptr(c_in), _ptr(c_out), _stream()))

    def fence(self):
        """dgb_fv_fence: completes the halo of the most recent REGISTERED output array (its x-overlap column is refreshed on
        demand only); call before reading overlap cells of that array with anything but step()"""
        check(lib.dgb_fv_fence(self.h, _stream()))

    def register_fields(self, *fields):
        """dgb_fv_register_fields (collective): the arrays the time loop alternates between; steps that write one of them let
        the peers store halo cells straight into it (no receive buffer, no unpack). Returns the number of registered arrays
        (0: declined on some rank - the steps then keep the buffered form)."""
        self._registered = list(fields)                 # keep them alive
        ptrs = (C.c_void_p * len(fields))(*[f.data_ptr() for f in fields])
        check(lib.dgb_fv_register_fields(self.h, ptrs, len(fields), _stream()))
        n = C.c_int(0)
        check(lib.dgb_fv_registered_fields(self.h, C.byref(n)))
        return n.value

    def bootstrap_local(self):
        p = C.c_void_p()
        check(lib.dgb_fv_bootstrap_local(self.h, C.byref(p), _stream()))
        return p.value

    def step_local(self, c_in, c_out):
        """kernel only (no exchange, no reduction over ranks); returns the device address of the pending max"""
        p = C.c_void_p()
        check(lib.dgb_fv_step_local(self.h, _ptr(c_in), _ptr(c_out), C.byref(p), _stream()))
        return p.value

    def update(self, c_in):
        torch = _torch()
        u = torch.zeros(self.ncells, dtype=torch.float64, device="cuda")
        check(lib.dgb_fv_update(self.h, _ptr(c_in), _ptr(u), _stream()))
        return u

    def last_dt(self):
        dt = C.c_double(0)
        check(lib.dgb_fv_last_dt(self.h, C.byref(dt), _stream()))
        return dt.value

    def step_host(self, h_in, h_out):
        """same step on host tensors (pinned or pageable): H2D + step + D2H inside the call"""
        dt = C.c_double(0)
        check(lib.dgb_fv_step_host(self.h, C.c_void_p(h_in.data_ptr()), C.c_void_p(h_out.data_ptr()), C.byref(dt), _stream()))
        return dt.value

    def host_path(self):
        """how the last step_host moved the field: 'none', 'serial' or 'pipelined' (z-slabs on three streams)"""
        m = C.c_int(0)
        check(lib.dgb_fv_host_path(self.h, C.byref(m)))
        return ("none", "serial", "pipelined")[m.value]

    def launch_count(self):
        n = C.c_int64(0)
        check(lib.dgb_fv_launch_count(self.h, C.byref(n)))
        return n.value

    def exchange_mode(self):
        """how the last step refreshed the overlap cells: 'none', 'serial', 'overlapped' (NCCL), 'fused' (peer stores into the
        receivers' buffers) or 'direct' (peer stores into the receivers' registered arrays)"""
        m = C.c_int(0)
        check(lib.dgb_fv_exchange_mode(self.h, C.byref(m)))
        return ("none", "serial", "overlapped", "fused", "direct", "hybrid")[m.value]
