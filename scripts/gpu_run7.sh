cd /root/repo
GCRL_EDGE_IMPL=tc timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x --timeout 300 -k "backward or learn or forward" 2>&1 | tail -1
GCRL_NODE_IMPL=ffma timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x --timeout 300 -k "forward" 2>&1 | tail -1
GCRL_FWD_IMPL=ws timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_api_gpu.py -q -x --timeout 300 2>&1 | tail -1
