"""The header-only C++ bulk GridView (include/dune_grid_b200/bulkgrid.hh) compiled with plain g++ against libdgb.so:
host-side descriptor queries on CPU, the reference-style invariants (index set, sub-indices, intersections, boundary
segments, geometry, communicate, FV functor) on the GPU."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")


def build():
    out = os.path.join(ROOT, "build", "bulk_gridview_test")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    cxx = shutil.which("/usr/bin/g++") or shutil.which("g++")
    lib = os.path.join(ROOT, "dune_grid_b200")
    cmd = [cxx, "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(CUDA, "include"),
           os.path.join(ROOT, "tests", "cpp", "bulk_gridview_test.cc"), "-o", out, "-L", lib, "-ldgb",
           "-L", os.path.join(CUDA, "lib64"), "-lcudart", f"-Wl,-rpath,{lib}", f"-Wl,-rpath,{os.path.join(CUDA, 'lib64')}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return out


def test_bulk_gridview_host_queries():
    r = subprocess.run([build()], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_bulk_gridview_device_invariants():
    r = subprocess.run([build(), "--device"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout + r.stderr
