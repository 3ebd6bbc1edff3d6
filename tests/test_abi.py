"""The drop-in boundary: libdgb.so must load without a GPU and export every function include/dgb.h declares, and the
ctypes stub (dune_grid_b200/capi.py) must bind every one of them. No compute call is made here."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "dgb.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)                      # comments mention function names too
    names = re.findall(r"^\s*(?:int|const\s+char\s*\*)\s+(dgb_\w+)\s*\(", text, flags=re.M)
    assert len(names) > 40, names
    return sorted(set(names))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(os.path.join(ROOT, "dune_grid_b200", "libdgb.so"))
    missing = [n for n in declared_functions() if not hasattr(lib, n)]
    assert not missing, missing


def test_ctypes_stub_binds_every_declared_symbol():
    from dune_grid_b200 import capi
    unbound = [n for n in declared_functions() if n not in capi.EXPORTS]
    assert not unbound, unbound
    undeclared = [n for n in capi.EXPORTS if n not in declared_functions()]
    assert not undeclared, undeclared


def test_no_cpu_fallback_without_a_device():
    """device entry points fail with DGB_ECUDA (and a message) where no GPU is usable; host queries keep working"""
    import torch
    from dune_grid_b200 import capi
    n = ctypes.c_int(0)
    rc = capi.lib.dgb_device_count(ctypes.byref(n))
    if torch.cuda.is_available():
        assert rc == 0 and n.value >= 1
        return
    import dune_grid_b200 as dg
    g = dg.YaspGrid(dim=2, N=(8, 4), upper=(1.0, 1.0))
    assert g.leafGridView().size(0) == 32                                   # host arithmetic works
    m = dg.MultipleCodimMultipleGeomTypeMapper(g.leafGridView(), [1, 0, 0])
    out = (ctypes.c_uint32 * 32)()
    rc = capi.lib.dgb_mapper_index(ctypes.byref(g.leafGridView().v), ctypes.byref(m.m), 0, 4, out, None)
    assert rc == -4 and b"no CPU fallback" in capi.lib.dgb_last_error()     # DGB_ECUDA
