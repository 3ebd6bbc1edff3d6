#!/usr/bin/env python
"""Generates tests/golden/*.npz from the REFERENCE's own code (oracle/_ref = dune-grid's yaspgrid headers compiled in
place, see oracle/Makefile). Run in the build container (needs /root/reference to build oracle/_ref):

    make -C oracle ref && python tests/golden/make_golden.py

The fixtures pin the CPU restatement (oracle/port) and the CUDA path to outputs of the reference itself; they are
consumed by tests/test_golden.py (CPU) and tests/test_gpu_golden.py (GPU). Geometry values come from the stand-in
AxisAlignedCubeGeometry / MultiLinearGeometry of oracle/ref/shim (dune-geometry is not available): see DESIGN.md §2."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import oracle as orc  # noqa: E402
from tests.golden_cases import CASES, snapshot  # noqa: E402


def main():
    o = orc.load("ref")
    assert o.kind == "reference"
    for name in CASES:
        s = snapshot(o, name)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **s)
        print(f"{name}: {len(s)} arrays, {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
