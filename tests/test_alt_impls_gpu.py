"""The fp32 vertex side under the tensor-core edge kernels (GCRL_NODE_IMPL=ffma) is selected by an environment variable
the library reads once per process, so it is exercised in a child pytest run and must pass the same tensor-core parity
tests.  (Round 1's superseded edge kernels -- single-group and warp-specialised -- were removed in round 2.)"""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize('env', [{'GCRL_NODE_IMPL': 'ffma'}])
def test_alternative_implementation_passes_tensor_core_parity(env):
    if os.environ.get('GCRL_ALT_CHILD'):
        pytest.skip('child run')
    e = dict(os.environ, GCRL_ALT_CHILD='1', **env)
    r = subprocess.run([sys.executable, '-m', 'pytest', os.path.join(ROOT, 'tests', 'test_kernels_gpu.py'), '-q', '-x',
                        '-k', 'tensor_core or learn_step', '--timeout', '300', '-p', 'no:cacheprovider'],
                       cwd=ROOT, env=e, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
