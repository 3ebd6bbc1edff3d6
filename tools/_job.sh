python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2g_pytest.log; tail -6 gpurun_out/r2g_pytest.log
for cfg in "2 64" "4 32" "8 128" "1 64"; do set -- $cfg; DGB_FACE_ROWS=$1 DGB_IDX_ROWS=$2 python tools/sweep_bench.py > gpurun_out/r2g_sweeps_f$1_i$2.jsonl 2>&1; done
python tools/fused_proxy.py 510 1024 1024 1 30 direct,fused > gpurun_out/r2g_px_hybrid.log 2>&1
DGB_FV_HYBRID=0 python tools/fused_proxy.py 510 1024 1024 1 30 direct > gpurun_out/r2g_px_direct.log 2>&1
python tools/fused_proxy.py 511 511 511 7 60 direct,fused > gpurun_out/r2g_p7_hybrid.log 2>&1
cat gpurun_out/r2g_px_hybrid.log gpurun_out/r2g_px_direct.log gpurun_out/r2g_p7_hybrid.log
