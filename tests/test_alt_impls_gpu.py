"""The A/B switches of the library are environment variables read once per process, so each is exercised in a child pytest
run that must pass the same tensor-core parity tests as the default path: the fp32 vertex side under the tensor-core edge
kernels (GCRL_NODE_IMPL=ffma), the unfused vertex-side launch chains of the forward (GCRL_VERTEX_IMPL=split) and of the
backward (GCRL_VBWD_IMPL=split), plain launches without programmatic dependent launch (GCRL_PDL=0), and block 1's edge kernels
through the projection + gather path instead of straight from the raw inputs (GCRL_FIRST_DIRECT=0), and the edge forward with its
MMA A operands as shared-memory tiles instead of tensor-memory columns (GCRL_FWD_ATM=0).
(Round 1's superseded edge kernels -- single-group and warp-specialised -- were removed in round 2.)"""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize('env', [{'GCRL_NODE_IMPL': 'ffma'}, {'GCRL_VERTEX_IMPL': 'split'}, {'GCRL_VBWD_IMPL': 'split'},
                                 {'GCRL_PDL': '0'}, {'GCRL_FIRST_DIRECT': '0'}, {'GCRL_FWD_ATM': '0'}])
def test_alternative_implementation_passes_tensor_core_parity(env):
    if os.environ.get('GCRL_ALT_CHILD'):
        pytest.skip('child run')
    e = dict(os.environ, GCRL_ALT_CHILD='1', **env)
    r = subprocess.run([sys.executable, '-m', 'pytest', os.path.join(ROOT, 'tests', 'test_kernels_gpu.py'),
                        os.path.join(ROOT, 'tests', 'test_headline_gpu.py'), '-q', '-x',
                        '-k', 'tensor_core or learn_step or bf16x3', '--timeout', '300', '-p', 'no:cacheprovider'],
                       cwd=ROOT, env=e, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
