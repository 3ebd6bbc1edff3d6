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


def test_fv_chunk_rule_host_arithmetic():
    """the z-chunk rules of the 3-D step kernel (host arithmetic, no device): a short LAST chunk fills the ragged last wave -
    512^3 single GPU (tensor-map form): 160 + 160 + 160 + 32; 512^3 interior of a rank of 1024^3 / 8 with the in-kernel pack:
    224 + 224 + 64; boxes with >= 2048 CTAs at 256 planes keep 256; small grids are cut until two waves of CTAs exist"""
    import ctypes as C
    from dune_grid_b200 import capi

    def rule(nx, ny, nz, packs, tma):
        z = C.c_int(0)
        assert capi.lib.dgb_fv_chunk_rule(nx, ny, nz, packs, tma, C.byref(z)) == 0
        return z.value
    assert rule(512, 512, 512, 0, 1) == 160
    assert rule(512, 512, 512, 0, 0) == 256
    assert rule(512, 512, 512, 1, 0) == 224
    assert rule(511, 511, 511, 1, 0) == 224
    assert rule(256, 1024, 1024, 1, 0) == 256          # rank box of 1024^3 / 4: 2048 CTAs at 256 planes
    assert rule(512, 1024, 1024, 1, 0) == 256          # rank box of 1024^3 / 2
    for nx, ny, nz, packs, tma in ((256, 8, 64, 0, 1), (256, 8, 32, 1, 0), (300, 40, 100, 0, 0), (1024, 1024, 1024, 0, 1)):
        z = rule(nx, ny, nz, packs, tma)
        assert 1 <= z <= max(nz, 16), (nx, ny, nz, z)
        chunks = -(-nz // z)
        assert chunks * z >= nz
        if tma and nz >= 512:                           # the last chunk is the short one
            assert nz - (chunks - 1) * z <= z // 2, (nz, z)
    assert capi.lib.dgb_fv_chunk_rule(0, 1, 1, 0, 0, None) != 0
