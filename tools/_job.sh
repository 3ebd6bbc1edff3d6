python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
echo "x-proxy default"; python tools/fused_proxy.py 510 1024 1024 1 30 direct 2>&1 | grep -E "^direct"
echo "x-proxy zc256"; DGB_FV_ZCHUNK=256 python tools/fused_proxy.py 510 1024 1024 1 30 direct 2>&1 | grep -E "^direct"
echo "x-proxy zc256 merged off"; DGB_FV_MERGED_WAIT=0 DGB_FV_ZCHUNK=256 python tools/fused_proxy.py 510 1024 1024 1 30 direct 2>&1 | grep -E "^direct"
echo "x-proxy zc128 merged off"; DGB_FV_MERGED_WAIT=0 python tools/fused_proxy.py 510 1024 1024 1 30 direct 2>&1 | grep -E "^direct"
echo "p7 zc256"; DGB_FV_ZCHUNK=256 python tools/fused_proxy.py 511 511 511 7 60 direct 2>&1 | grep -E "^direct"
echo "N4-like x-proxy 254x1024x1024"; python tools/fused_proxy.py 254 1024 1024 1 30 direct 2>&1 | grep -E "^direct|kernel"
echo "N4-like zc256"; DGB_FV_ZCHUNK=256 python tools/fused_proxy.py 254 1024 1024 1 30 direct 2>&1 | grep -E "^direct"
