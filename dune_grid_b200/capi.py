"""ctypes binding of libdgb.so (include/dgb.h). There is no Python or CPU fallback: if the shared
library is missing this module raises at import time, and device entry points raise DgbError(DGB_ECUDA)
on a box without a usable GPU."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DGB_LIBRARY") or os.path.join(HERE, "libdgb.so")      # DGB_LIBRARY: an alternative build (developer A/B runs)

DGB_MAX_DIM, DGB_MAX_COMP = 3, 8
OK, EGRID, ERANGE, EARG, ECUDA, ENCCL, ENOTIMPL = 0, -1, -2, -3, -4, -5, -6
COORD_EQUIDISTANT, COORD_EQUIDISTANT_OFFSET, COORD_TENSORPRODUCT = 0, 1, 2
PART_DEFAULT, PART_POWER_D, PART_FIXED = 0, 1, 2
BOX_OVERLAPFRONT, BOX_OVERLAP, BOX_INTERIORBORDER, BOX_INTERIOR = 0, 1, 2, 3
FV_EXACT, FV_SEPARABLE = 0, 1
OP_SET, OP_ADD = 0, 1


class GridSpec(C.Structure):
    _fields_ = [("dim", C.c_int32), ("coord_mode", C.c_int32), ("lower", C.c_double * 3), ("upper", C.c_double * 3),
                ("N", C.c_int32 * 3), ("tensor", C.POINTER(C.c_double) * 3), ("periodic", C.c_uint32),
                ("overlap", C.c_int32), ("partitioner", C.c_int32), ("fixed_dims", C.c_int32 * 3)]


class Box(C.Structure):
    _fields_ = [("origin", C.c_int32 * 3), ("size", C.c_int32 * 3)]


class Component(C.Structure):
    _fields_ = [("shift", C.c_uint32), ("codim", C.c_int32), ("index_offset", C.c_int32 * 4), ("box", Box * 4)]


class View(C.Structure):
    _fields_ = [("dim", C.c_int32), ("level", C.c_int32), ("rank", C.c_int32), ("nranks", C.c_int32),
                ("N", C.c_int32 * 3), ("periodic", C.c_uint32), ("overlap", C.c_int32), ("ncomp", C.c_int32),
                ("comp_begin", C.c_int32 * 5), ("shiftmap", C.c_int32 * 8), ("comp", Component * 8),
                ("coord_mode", C.c_int32), ("h", C.c_double * 3), ("origin", C.c_double * 3),
                ("tensor", C.c_void_p * 3), ("tensor_host", C.POINTER(C.c_double) * 3),
                ("tensor_offset", C.c_int32 * 3), ("tensor_size", C.c_int32 * 3), ("L", C.c_double * 3),
                ("bs_origin", C.c_int32 * 3), ("bs_size", C.c_int32 * 3), ("bs_sides", C.c_int32 * 3), ("N0", C.c_int32 * 3)]


class VtkField(C.Structure):
    _fields_ = [("name", C.c_char_p), ("d_data", C.c_void_p), ("ncomp", C.c_int32), ("is_vector", C.c_int32), ("precision", C.c_int32)]


class Mapper(C.Structure):
    _fields_ = [("dim", C.c_int32), ("block", C.c_uint32 * 4), ("offset", C.c_uint32 * 4), ("size", C.c_uint32)]


class DgbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"dgb error {code}: {msg}")
        self.code = code


if not os.path.exists(LIB_PATH):
    raise ImportError(f"{LIB_PATH} is missing: build it with ./build.sh (or __graft_entry__.build()). "
                      "dune_grid_b200 has no CPU or pure-Python fallback.")

lib = C.CDLL(LIB_PATH)

_vp, _i, _i64p = C.c_void_p, C.c_int, C.POINTER(C.c_int64)
_pv, _pm = C.POINTER(View), C.POINTER(Mapper)
_SIG = {
    "dgb_last_error": (C.c_char_p, []), "dgb_version": (C.c_char_p, []),
    "dgb_device_count": (_i, [C.POINTER(_i)]),
    "dgb_partition": (_i, [_i, C.POINTER(C.c_int32), _i, _i, _i, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "dgb_grid_create": (_i, [C.POINTER(GridSpec), _i, _i, C.POINTER(_vp)]),
    "dgb_grid_backup": (_i, [_vp, C.c_char_p, C.c_int64, C.POINTER(C.c_int64)]),
    "dgb_grid_restore": (_i, [C.c_char_p, _i, _i, _i, C.POINTER(_vp)]),
    "dgb_field_backup": (_i, [_vp, _i, C.POINTER(Mapper), _vp, _i, C.c_char_p, _vp]),
    "dgb_field_restore": (_i, [_vp, _i, C.POINTER(Mapper), _vp, _i, C.c_char_p, _vp]),
    "dgb_global_index": (_i, [_vp, _i, _i, _vp, _i64p, _vp, _vp]),
    "dgb_vtk_write": (_i, [_vp, _i, C.c_char_p, C.c_char_p, C.POINTER(VtkField), _i, C.POINTER(VtkField), _i, _i, C.c_char_p, _i, _vp]),
    "dgb_vtk_write_ex": (_i, [_vp, _i, C.c_char_p, C.c_char_p, C.POINTER(VtkField), _i, C.POINTER(VtkField), _i, _i, _i, C.c_char_p, _i, _vp]),
    "dgb_grid_destroy": (_i, [_vp]), "dgb_grid_refine_options": (_i, [_vp, _i]), "dgb_grid_global_refine": (_i, [_vp, _i]),
    "dgb_grid_max_level": (_i, [_vp, C.POINTER(_i)]), "dgb_grid_torus_dims": (_i, [_vp, C.POINTER(C.c_int32)]),
    "dgb_grid_num_boundary_segments": (_i, [_vp, _i64p]), "dgb_grid_domain_size": (_i, [_vp, C.POINTER(C.c_double)]),
    "dgb_view_leaf": (_i, [_vp, _pv]), "dgb_view_level": (_i, [_vp, _i, _pv]),
    "dgb_view_size": (_i, [_pv, _i, _i64p]), "dgb_view_partition_size": (_i, [_pv, _i, _i, _i64p]),
    "dgb_view_overlap_size": (_i, [_pv, _i, C.POINTER(_i)]), "dgb_view_ghost_size": (_i, [_pv, _i, C.POINTER(_i)]),
    "dgb_view_comm_list": (_i, [_vp, _i, _i, _i, C.POINTER(C.c_int32), _i, C.POINTER(_i)]),
    "dgb_mapper_init": (_i, [_pv, C.POINTER(C.c_uint32), _pm]), "dgb_mapper_size": (_i, [_pm, _i64p]),
    "dgb_mapper_index": (_i, [_pv, _pm, _i, _i, _vp, _vp]), "dgb_mapper_subindex": (_i, [_pv, _pm, _i, _i, _vp, _vp]),
    "dgb_elements_materialize": (_i, [_pv, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "dgb_intersections_materialize": (_i, [_pv, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "dgb_intersection_normals": (_i, [_pv, _i, _vp, _vp, _vp]),
    "dgb_intersection_local_geometry": (_i, [_i, _i, _i, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "dgb_geometry_eval": (_i, [_pv, _i, _i, C.POINTER(C.c_double), _vp, _vp, _vp, _vp]),
    "dgb_geometry_eval_codim": (_i, [_pv, _i, _i, _i, C.POINTER(C.c_double), _vp, _vp, _vp, _vp, _vp]),
    "dgb_geogrid_eval": (_i, [_pv, _i, _i, C.POINTER(C.c_double), _i, _i, C.POINTER(C.c_double), _vp, _vp, _vp, _vp, _vp, _vp]),
    "dgb_geogrid_intersections": (_i, [_pv, _i, _i, C.POINTER(C.c_double), _i, _i, C.POINTER(C.c_double), _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "dgb_geogrid_volume": (_i, [_pv, _i, _i, C.POINTER(C.c_double), _i, _i, C.POINTER(C.c_double), C.POINTER(C.c_double), _vp, _vp]),
    "dgb_functor_count": (_i, [C.POINTER(_i)]),
    "dgb_functor_info": (_i, [_i, C.POINTER(C.c_char_p), C.POINTER(_i), C.POINTER(_i)]),
    "dgb_for_each_element": (_i, [_pv, _i, _i, C.POINTER(C.c_double), _i, _vp, _vp, _vp]),
    "dgb_for_each_intersection": (_i, [_pv, _i, _i, C.POINTER(C.c_double), _i, _vp, _vp, _vp]),
    "dgb_comm_unique_id": (_i, [_vp]), "dgb_comm_init": (_i, [_vp, _i, _i, C.POINTER(_vp)]), "dgb_comm_destroy": (_i, [_vp]),
    "dgb_communicate": (_i, [_vp, _i, _pm, _i, _i, _i, C.POINTER(_vp), C.POINTER(C.c_int32), _i, _vp, _vp]),
    "dgb_comm_plan_info": (_i, [_vp, _i, _pm, _i, _i, C.POINTER(C.c_int32), _i, _i64p, _i]),
    "dgb_comm_pack": (_i, [_vp, _i, _pm, _i, _i, C.POINTER(_vp), C.POINTER(C.c_int32), _i, C.POINTER(_vp), C.POINTER(_vp), _vp]),
    "dgb_comm_unpack": (_i, [_vp, _i, _pm, _i, _i, _i, C.POINTER(_vp), C.POINTER(C.c_int32), _i, _vp]),
    "dgb_commvar_create": (_i, [_vp, _i, _pm, _i, _i, _vp, _vp, _vp, C.POINTER(_vp)]),
    "dgb_commvar_exchange": (_i, [_vp, _vp, _vp]),
    "dgb_commvar_segments": (_i, [_vp, _i, _i64p, _i]),
    "dgb_commvar_buffers": (_i, [_vp, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp)]),
    "dgb_commvar_sizes_received": (_i, [_vp, _vp]),
    "dgb_commvar_finish": (_i, [_vp, _vp, _i64p, _i64p]),
    "dgb_commvar_result": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "dgb_commvar_destroy": (_i, [_vp]),
    "dgb_allreduce_min": (_i, [_vp, _vp, _vp]), "dgb_allreduce_max": (_i, [_vp, _vp, _vp]),
    "dgb_comm_last_bytes": (_i, [_vp, _i64p]),
    "dgb_fv_create": (_i, [_vp, _i, _i, C.POINTER(C.c_double), _i, _i, _vp, C.POINTER(_vp)]),
    "dgb_fv_destroy": (_i, [_vp]), "dgb_fv_abandon": (_i, [_vp]), "dgb_fv_init_field": (_i, [_vp, _vp, _vp]),
    "dgb_fv_register_fields": (_i, [_vp, C.POINTER(_vp), _i, _vp]), "dgb_fv_registered_fields": (_i, [_vp, C.POINTER(_i)]),
    "dgb_fv_fence": (_i, [_vp, _vp]),
    "dgb_fv_kernel_info": (_i, [_vp, C.POINTER(C.c_int32)]), "dgb_fv_host_path": (_i, [_vp, C.POINTER(_i)]),
    "dgb_fv_chunk_rule": (_i, [_i, _i, _i, _i, _i, C.POINTER(_i)]),
    "dgb_fv_step": (_i, [_vp, _vp, _vp, _vp]), "dgb_fv_last_dt": (_i, [_vp, C.POINTER(C.c_double), _vp]),
    "dgb_fv_update": (_i, [_vp, _vp, _vp, _vp]),
    "dgb_fv_step_local": (_i, [_vp, _vp, _vp, C.POINTER(_vp), _vp]),
    "dgb_fv_bootstrap_local": (_i, [_vp, C.POINTER(_vp), _vp]),
    "dgb_fv_step_host": (_i, [_vp, _vp, _vp, C.POINTER(C.c_double), _vp]),
    "dgb_fv_launch_count": (_i, [_vp, _i64p]),
    "dgb_fv_exchange_mode": (_i, [_vp, C.POINTER(_i)]),
}
EXPORTS = sorted(_SIG)
for _name, (_res, _args) in _SIG.items():
    _f = getattr(lib, _name)
    _f.restype, _f.argtypes = _res, _args


def check(rc):
    if rc != OK:
        raise DgbError(rc, lib.dgb_last_error().decode())
    return rc


def device_count():
    n = C.c_int(0)
    rc = lib.dgb_device_count(C.byref(n))
    return n.value if rc == OK else 0
