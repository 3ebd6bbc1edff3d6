python -m pytest tests/test_gpu_comm_fv.py -m gpu -x -q 2>&1 | tail -3
for m in 0 63; do echo "skip mask $m"; DGB_FV_SKIP_FACES=$m python tools/fused_proxy.py 511 511 511 7 60 direct 2>&1 | grep -E "^direct|kernel"; done
echo "merged wait off"; DGB_FV_MERGED_WAIT=0 DGB_FV_SKIP_FACES=63 python tools/fused_proxy.py 511 511 511 7 60 direct 2>&1 | grep -E "^direct"
echo "hybrid off"; DGB_FV_HYBRID=0 DGB_FV_SKIP_FACES=63 python tools/fused_proxy.py 511 511 511 7 60 direct 2>&1 | grep -E "^direct"
echo "zchunk 256"; DGB_FV_ZCHUNK=256 DGB_FV_SKIP_FACES=63 python tools/fused_proxy.py 511 511 511 7 60 direct 2>&1 | grep -E "^direct"
