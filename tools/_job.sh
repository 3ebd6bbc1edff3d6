python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
python tools/intersections_probe.py > gpurun_out/r2h_intersections_probe.jsonl 2>&1; cat gpurun_out/r2h_intersections_probe.jsonl
