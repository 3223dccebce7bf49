"""The alternative edge-kernel implementations are selected by environment variables that the library reads once
per process, so they are exercised in a child pytest run: the default dispatch is the multi-group pair
(k_edge_fwd_wg / k_edge_bwd_wg); the single-group and warp-specialised forward (GCRL_FWD_IMPL=tc / ws), the
warp-specialised and single-group backward (GCRL_EDGE_IMPL=ws / tc) and the fp32 vertex side under the tensor-core
edge kernels (GCRL_NODE_IMPL=ffma) must pass the same tensor-core parity tests."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize('env', [{'GCRL_FWD_IMPL': 'ws'}, {'GCRL_FWD_IMPL': 'tc'}, {'GCRL_EDGE_IMPL': 'ws'},
                                 {'GCRL_EDGE_IMPL': 'tc'}, {'GCRL_NODE_IMPL': 'ffma'}])
def test_alternative_implementation_passes_tensor_core_parity(env):
    if os.environ.get('GCRL_ALT_CHILD'):
        pytest.skip('child run')
    e = dict(os.environ, GCRL_ALT_CHILD='1', **env)
    r = subprocess.run([sys.executable, '-m', 'pytest', os.path.join(ROOT, 'tests', 'test_kernels_gpu.py'), '-q', '-x',
                        '-k', 'tensor_core or learn_step', '--timeout', '300', '-p', 'no:cacheprovider'],
                       cwd=ROOT, env=e, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
