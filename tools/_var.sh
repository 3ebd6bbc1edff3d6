timeout 300 python -m pytest tests/test_gpu_comm_fv.py -x -q -m gpu -k "ring_kernels or fused or public_step or full_size" 2>&1 | tail -3
for a in 1 0 1 0; do echo "A16=$a"; DGB_FV_A16=$a timeout 100 python tools/halo_timing.py 512 1 0 2>&1 | grep kernel; done
