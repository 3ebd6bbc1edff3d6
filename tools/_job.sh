python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/r2g_pytest.log; tail -5 gpurun_out/r2g_pytest.log
for cfg in "2 64" "4 32" "8 128" "1 64"; do set -- $cfg; DGB_FACE_ROWS=$1 DGB_IDX_ROWS=$2 python tools/sweep_bench.py > gpurun_out/r2g_sweeps_f$1_i$2.jsonl 2>&1; done
