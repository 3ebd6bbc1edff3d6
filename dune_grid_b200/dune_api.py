"""`dune.grid`-style façade over the bulk GPU path: the names a dune-python script uses for YaspGrid —
`cartesianDomain`, `equidistantCoordinates`, `equidistantOffsetCoordinates`, `tensorProductCoordinates`, `yaspGrid`,
`structuredGrid`, `view.hierarchicalGrid.globalRefine`, `view.mapper(layout)`, `mapper.communicate(from, to, commOp,
*arrays)` — with the reference's argument meaning (python/dune/grid/_grids.py:172-321, python/dune/grid/core.py:7-13,
dune/python/grid/mapper.hh:117-249, dune/python/grid/gridview.hh:276-310). Entities are not Python objects here: whatever
the reference returns per entity comes back as whole arrays (torch CUDA tensors) in the view's index order.

    from dune_grid_b200.dune_api import cartesianDomain, yaspGrid, CommOp
    view = yaspGrid(cartesianDomain([0, 0, 0], [1, 1, 1], [64, 64, 64]), periodic=[True, False, False], overlap=1)
    mapper = view.mapper(lambda gt: gt.dim == view.dimension)          # one dof per element
    mapper.communicate(view.interiorBorderPartition, view.allPartition, CommOp.set, data)
"""
import enum
from typing import Sequence

import numpy as np

from . import grid as _g


class CommOp(enum.IntEnum):                     # dune/python/grid/mapper.hh:117 (enum CommOp { set, add })
    set = 0
    add = 1


# The reference's mapperCommunicate has no `break` after `case add` (dune/python/grid/mapper.hh:124-137): an `add` exchange
# is followed by a `set` exchange of the summed values. Scripts written against dune-python see that behaviour, so the
# façade reproduces it; set this to False for a plain add.
ADD_FALLS_THROUGH_TO_SET = True


class GeometryType:
    """the one geometry type per dimension a YaspGrid has (cubes; yaspgridindexsets.hh:90-93)"""

    def __init__(self, dim):
        self.dim = dim
        self.isCube = True
        self.isSimplex = dim <= 1
        self.isVertex = dim == 0
        self.isLine = dim == 1

    def __repr__(self):
        return f"GeometryType(cube, {self.dim})"


class _Coordinates:
    ctype = "double"

    def __init__(self, kind, dim, **kw):
        self.kind, self.dimgrid = kind, dim
        self.typeName = f"Dune::{kind}< double, {dim} >"
        self.__dict__.update(kw)

    @classmethod
    def name(cls):
        return cls.__name__


def equidistantCoordinates(upperright, elements):
    """EquidistantCoordinates(upperRight, s) (_grids.py:172-182)"""
    upperright = np.array(upperright, dtype=np.float64)
    assert len(upperright) == len(elements)
    return _Coordinates("EquidistantCoordinates", len(elements), upper=upperright, elements=list(elements))


def equidistantOffsetCoordinates(lowerleft, upperright, elements, ctype="double"):
    """EquidistantOffsetCoordinates(lowerLeft, upperRight, s) (_grids.py:159-170)"""
    if ctype != "double":
        raise ValueError("only ctype double is built")
    lowerleft, upperright = np.array(lowerleft, dtype=np.float64), np.array(upperright, dtype=np.float64)
    assert len(lowerleft) == len(elements) and len(upperright) == len(elements)
    return _Coordinates("EquidistantOffsetCoordinates", len(elements), lower=lowerleft, upper=upperright, elements=list(elements))


def tensorProductCoordinates(coords, offset=None, ctype="double"):
    """TensorProductCoordinates(c, offset) (_grids.py:184-196)"""
    if ctype != "double":
        raise ValueError("only ctype double is built")
    dim = len(coords)
    if offset is None:
        offset = [0] * dim
    if len(offset) != dim:
        raise ValueError("tensorProductCoordinates: offset parameter has wrong size")
    return _Coordinates("TensorProductCoordinates", dim, coords=[np.array(c, dtype=np.float64) for c in coords], offset=list(offset))


class CartesianDomain:
    """CartesianDomain(lower, upper, division, **parameters) (_grids.py:6-80); `periodic` and `overlap` parameters are honoured
    by yaspGrid like in the reference (:228-245)"""

    def __init__(self, lower, upper, division, boundary=True, **parameters):
        self.dimgrid = len(lower)
        self.lower, self.upper, self.division, self.param = list(lower), list(upper), list(division), dict(parameters)


def cartesianDomain(lower, upper, division, **parameters):          # python/dune/grid/core.py:7-8
    return CartesianDomain(lower, upper, division, **parameters)


class Partition:
    """GridViewPartition (dune/python/grid/gridview.hh:276-310): the entities of one PartitionIteratorType"""

    def __init__(self, view, pitype):
        self.gridView, self.pitype = view, pitype

    def size(self, codim):
        return self.gridView._gv.size(codim, self.pitype)


class HierarchicalGrid:
    def __init__(self, yasp, comm):
        self._yasp, self._comm = yasp, comm

    def globalRefine(self, steps=1):
        self._yasp.globalRefine(steps)

    def refineOptions(self, keepPhysicalOverlap):
        self._yasp.refineOptions(keepPhysicalOverlap)

    @property
    def maxLevel(self):
        return self._yasp.maxLevel

    @property
    def leafView(self):
        return GridView(self, None)

    def levelView(self, level):
        return GridView(self, level)

    @property
    def dimension(self):
        return self._yasp.dim

    def backup(self, filename=None):
        return self._yasp.backup(filename)


class _Comm:
    def __init__(self, rank, size):
        self.rank, self.size = rank, size


class GridView:
    """leaf or level view; bulk counterparts of the per-entity calls are methods returning arrays"""

    def __init__(self, hgrid, level):
        self.hierarchicalGrid = hgrid
        self._level = level

    @property
    def _gv(self):                      # re-created on demand: globalRefine changes what "leaf" means
        y = self.hierarchicalGrid._yasp
        return y.leafGridView() if self._level is None else y.levelGridView(self._level)

    @property
    def dimension(self):
        return self.hierarchicalGrid._yasp.dim

    dimGrid = dimWorld = dimension

    @property
    def comm(self):
        y = self.hierarchicalGrid._yasp
        return _Comm(y.rank, y.nranks)

    def size(self, codim):
        return self._gv.size(codim)

    def mapper(self, layout):
        return Mapper(self, layout)

    def coordinates(self):
        """vertex positions, shape (size(dim), dim) (dune/python/grid/numpy.hh:46-70)"""
        e = self._gv.elements(self.dimension, _g.All_Partition)
        return e["center"].T.contiguous()

    def globalIndexSet(self, codim):
        return self._gv.globalIndexSet(codim)

    def geometries(self, codim=0, partition=None):
        """the geometries of ALL entities of a codimension at once: the bulk form of `for e in view.entities(codim): e.geometry`,
        with the vectorised method names of dune/python/grid/geometry.hh:214-349"""
        return BulkGeometry(self, codim, _g.All_Partition if partition is None else partition.pitype if isinstance(partition, Partition) else partition)

    def polygons(self):
        """corner loops of the cells of a 2-D view, shape (n, 4, 2): numpy.hh:184-212 (corners 0 and 1 swapped, as there)"""
        if self.dimension != 2:
            raise ValueError("polygons: two-dimensional views only (numpy.hh:201)")
        c = self.geometries(0).corners                                     # (4, 2, n)
        return c[[1, 0, 2, 3]].permute(2, 0, 1).contiguous()

    def writeVTK(self, name, celldata=None, pointdata=None, **kw):
        """gridView.writeVTK(name, celldata={...}, pointdata={...}) of dune-python (python/dune/grid/grid_generator.py): arrays
        indexed by the element / vertex index"""
        return self._gv.writeVTK(name, celldata=celldata, pointdata=pointdata, **kw)

    interiorPartition = property(lambda self: Partition(self, _g.Interior_Partition))
    interiorBorderPartition = property(lambda self: Partition(self, _g.InteriorBorder_Partition))
    overlapPartition = property(lambda self: Partition(self, _g.Overlap_Partition))
    overlapFrontPartition = property(lambda self: Partition(self, _g.OverlapFront_Partition))
    allPartition = property(lambda self: Partition(self, _g.All_Partition))
    ghostPartition = property(lambda self: Partition(self, _g.Ghost_Partition))


class BulkGeometry:
    """Geometry of every entity of one codimension of a view (dune/python/grid/geometry.hh:168-349, vectorised forms): local
    points x are (mydim,) or (npoints, mydim) arrays; results are CUDA tensors whose LAST axis runs over the entities in
    iteration order (the order of mapper.index). Axis-aligned cubes (yaspgrid/yaspgridgeometry.hh:30-77), so affine is True."""

    affine = True

    def __init__(self, view, codim, pitype):
        self._view, self.codim, self._pit = view, codim, pitype
        self.dimension = self.mydimension = view.dimension - codim
        self.coorddimension = view.dimension

    def _x(self, x):
        x = np.asarray(x, dtype=np.float64)
        single = x.ndim <= 1
        return np.ascontiguousarray(x.reshape(-1, max(1, self.mydimension))), single

    def _eval(self, x, key):
        qp, single = self._x(x)
        r = self._view._gv.geometryCodim(self.codim, qp, self._pit)[key]
        return r[0] if single else r

    def toGlobal(self, x):                                       # geometry.hh:200-214
        return self._eval(x, "global")

    def integrationElement(self, x):                             # :232-250
        return self._eval(x, "integrationElement")

    def jacobianTransposed(self, x):                             # :268-279
        return self._eval(x, "jacobianTransposed")

    def jacobianInverseTransposed(self, x):                      # :289-300
        return self._eval(x, "jacobianInverseTransposed")

    def jacobian(self, x):                                       # :252-266: transpose of jacobianTransposed
        jt = self.jacobianTransposed(x)
        return jt.transpose(-3, -2)

    def jacobianInverse(self, x):                                # :280-288
        jit = self.jacobianInverseTransposed(x)
        return jit.transpose(-3, -2)

    def toLocal(self, y):
        """local coordinates of ONE global point y in every cell: shape (d, n); AxisAlignedCubeGeometry::local = (y - lower) / (upper - lower)"""
        if self.codim != 0:
            raise NotImplementedError("toLocal: cells only")
        torch = _g._torch()
        e = self._view._gv.elements(0, self._pit)
        y = torch.as_tensor(np.asarray(y, dtype=np.float64), device="cuda").reshape(-1, 1)
        return (y - e["lower"]) / (e["upper"] - e["lower"])

    @property
    def center(self):                                            # (d, n)
        return self._view._gv.elements(self.codim, self._pit)["center"]

    @property
    def volume(self):                                            # (n,)
        return self._view._gv.elements(self.codim, self._pit)["volume"]

    @property
    def corners(self):
        """corner(i) of every entity, reference-cube numbering: shape (2^mydim, d, n)"""
        m = self.mydimension
        ref = np.array([[(i >> k) & 1 for k in range(max(1, m))] for i in range(1 << m)], dtype=np.float64).reshape(1 << m, max(1, m))
        return self._view._gv.geometryCodim(self.codim, ref, self._pit)["global"]

    def __len__(self):
        return self._view._gv.size(self.codim, self._pit)


# registerMapperCommunicate instantiations (dune/python/grid/mapper.hh:196-249 and its callers): (from, to) -> interface;
# the swapped pair selects BackwardCommunication
_INTERFACES = {
    (_g.InteriorBorder_Partition, _g.InteriorBorder_Partition): _g.InteriorBorder_InteriorBorder_Interface,
    (_g.InteriorBorder_Partition, _g.All_Partition): _g.InteriorBorder_All_Interface,
    (_g.Overlap_Partition, _g.OverlapFront_Partition): _g.Overlap_OverlapFront_Interface,
    (_g.Overlap_Partition, _g.All_Partition): _g.Overlap_All_Interface,
    (_g.All_Partition, _g.All_Partition): _g.All_All_Interface,
}


class Mapper:
    """MultipleCodimMultipleGeomTypeMapper(gridView, layout) (python/dune/grid/map.py, common/mcmgmapper.hh:212-249): layout is
    a callable GeometryType -> number of dofs (bool counts as 0/1), or a sequence / dict indexed by the dimension of the type"""

    def __init__(self, view, layout):
        self.gridView = view
        dim = view.dimension
        block = []
        for codim in range(dim + 1):
            gt = GeometryType(dim - codim)
            if callable(layout):
                n = layout(gt)
            elif isinstance(layout, dict):
                n = layout.get(gt.dim, 0)
            else:
                n = layout[gt.dim] if gt.dim < len(layout) else 0
            block.append(int(n))
        self.layout = block
        self._m = _g.MultipleCodimMultipleGeomTypeMapper(view._gv, block)

    def __len__(self):
        return self._m.size

    @property
    def size(self):
        return self._m.size

    def types(self, codim):
        return [GeometryType(self.gridView.dimension - codim)] if self.layout[codim] else []

    def index(self, codim=0, partition=None):
        """mapper.index(e) for all entities of the codim (first dof of each entity), index order"""
        return self._m.index(codim, _g.All_Partition if partition is None else partition.pitype)

    def subIndex(self, codim, partition=None):
        """mapper.subIndex(e, i, codim) for all cells: shape (subEntities, cells)"""
        return self._m.subIndex(codim, _g.All_Partition if partition is None else partition.pitype)

    def update(self, view=None):
        """mapper.update(gridView) after a grid modification (mapper.hh:318-325)"""
        view = view or self.gridView
        block, dim = list(self.layout), view.dimension
        self.__init__(view, lambda gt: block[dim - gt.dim])

    def communicate(self, from_, to, commOp, *arrays):
        """mapper.communicate(from, to, commOp, *arrays): arrays are indexed by the mapper (first extent == len(mapper));
        torch CUDA tensors are exchanged in place, NumPy arrays are copied to the device and back"""
        key = (from_.pitype, to.pitype)
        if key in _INTERFACES:
            iftype, direction = _INTERFACES[key], _g.ForwardCommunication
        elif key[::-1] in _INTERFACES:
            iftype, direction = _INTERFACES[key[::-1]], _g.BackwardCommunication
        else:
            raise TypeError("communicate: no interface for this pair of partitions")
        if callable(commOp) and not isinstance(commOp, CommOp):
            raise NotImplementedError("communicate with a Python combine function: host callbacks are not supported on the bulk path")
        import torch
        dev, back = [], []
        for a in arrays:
            if isinstance(a, np.ndarray):
                if a.dtype != np.float64 or not a.flags.c_contiguous:
                    raise TypeError("communicate: NumPy arrays must be C-contiguous float64 (they are updated in place)")
                t = torch.from_numpy(a).cuda()
                back.append((a, t))
            else:
                t = a
            if t.shape[0] != len(self):
                raise ValueError("communicate: first extent of every array must equal the mapper size")
            dev.append(t)
        items = [int(np.prod(t.shape[1:])) if t.dim() > 1 else 1 for t in dev]
        gv = self.gridView._gv
        ops = ["add", "set"] if (CommOp(commOp) == CommOp.add and ADD_FALLS_THROUGH_TO_SET) else [CommOp(commOp).name]
        for op in ops:
            gv.communicate(self._m, dev, iftype, direction, op, items)
        torch.cuda.synchronize()
        for a, t in back:
            a[...] = t.cpu().numpy()


def yaspGrid(constructor, dimgrid=None, coordinates="equidistant", ctype=None, periodic=None, overlap=None,
             rank=0, nranks=1, comm=None, **param):
    """create a YaspGrid and return its leaf view (python/dune/grid/_grids.py:198-321). `constructor` is a coordinates object or
    a CartesianDomain. One process per GPU: `rank` / `nranks` / `comm` replace the MPI communicator of the reference."""
    if isinstance(constructor, CartesianDomain):
        if periodic is None:
            periodic = constructor.param.get("periodic", None)
        if overlap is None:
            overlap = constructor.param.get("overlap", None)
        constructor = equidistantOffsetCoordinates(constructor.lower, constructor.upper, constructor.division)
    if not isinstance(constructor, _Coordinates):
        raise ValueError("yaspGrid: unsupported constructor parameter " + str(constructor))
    if ctype not in (None, "double"):
        raise ValueError("yaspGrid: ctype is specified via coordinate object and can not be overwritten")
    dim = constructor.dimgrid
    if dimgrid is not None and dimgrid != dim:
        raise ValueError("yaspGrid: parameter dimgrid can only be specified, when using a reader")
    if periodic is None:
        periodic = [False] * dim
    if overlap is None:
        overlap = 1
    if constructor.kind == "EquidistantCoordinates":
        y = _g.YaspGrid(dim=dim, N=constructor.elements, upper=list(constructor.upper), periodic=list(periodic), overlap=overlap,
                        coordinates="equidistant", rank=rank, nranks=nranks, comm=comm)
    elif constructor.kind == "EquidistantOffsetCoordinates":
        y = _g.YaspGrid(dim=dim, N=constructor.elements, lower=list(constructor.lower), upper=list(constructor.upper),
                        periodic=list(periodic), overlap=overlap, coordinates="equidistantoffset", rank=rank, nranks=nranks, comm=comm)
    else:
        y = _g.YaspGrid(coords=constructor.coords, periodic=list(periodic), overlap=overlap, coordinates="tensorproduct",
                        rank=rank, nranks=nranks, comm=comm)
    return HierarchicalGrid(y, comm).leafView


def structuredGrid(lower, upper, division, **parameters):           # python/dune/grid/core.py:10-13
    return yaspGrid(cartesianDomain(lower, upper, division, **parameters), dimgrid=len(lower), coordinates="equidistantoffset")


__all__: Sequence[str] = ["CommOp", "GeometryType", "CartesianDomain", "cartesianDomain", "equidistantCoordinates",
                          "equidistantOffsetCoordinates", "tensorProductCoordinates", "yaspGrid", "structuredGrid", "Mapper",
                          "GridView", "HierarchicalGrid", "Partition"]
