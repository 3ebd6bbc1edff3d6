python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python tools/fused_proxy.py 511 511 511 7 60 direct,fused 2>&1 | tail -4
DGB_FV_MERGED_WAIT=0 python tools/fused_proxy.py 511 511 511 7 60 direct 2>&1 | tail -3
python tools/sweep_bench.py > gpurun_out/r2j_sweeps.jsonl 2>&1
python tools/intersections_probe.py > gpurun_out/r2j_intersections_probe.jsonl 2>&1; cat gpurun_out/r2j_intersections_probe.jsonl | cut -c1-200
