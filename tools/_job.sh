python -m pytest tests/test_gpu_comm_fv.py tests/test_multigpu.py -m gpu -x -q 2>&1 | tail -3
for m in 0 63; do echo "skip mask $m"; DGB_FV_SKIP_FACES=$m python tools/fused_proxy.py 511 511 511 7 60 direct,fused 2>&1 | grep -E "^direct|^fused|kernel"; done
python tools/fused_proxy.py 510 1024 1024 1 30 direct 2>&1 | grep -E "^direct|kernel"
