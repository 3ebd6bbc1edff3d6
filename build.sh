#!/bin/bash
# Builds dune_grid_b200/libdgb.so (CUDA kernels + C ABI) for sm_100a, in tree.
set -e
cd "$(dirname "$0")"
SRC=dune_grid_b200/csrc
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC,-Wall,-Wno-unused-function ${DGB_NVCC_EXTRA}"
mkdir -p build
for f in dgb_capi dgb_sweeps dgb_comm dgb_fv dgb_gis dgb_vtk dgb_foreach; do
  if [ ! -f build/$f.o ] || [ -n "$(find $SRC include -newer build/$f.o -type f | head -1)" ]; then
    echo "nvcc $f.cu"
    $NVCC $FLAGS -c $SRC/$f.cu -o build/$f.o &
  fi
done
wait
$NVCC -shared -o dune_grid_b200/libdgb.so build/dgb_capi.o build/dgb_sweeps.o build/dgb_comm.o build/dgb_fv.o build/dgb_gis.o build/dgb_vtk.o build/dgb_foreach.o -ldl
echo "built dune_grid_b200/libdgb.so"
