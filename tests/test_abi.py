"""CPU tests of the C-ABI boundary: the library loads, exports exactly what include/gcrl_b200.h
declares, and its argument checks / host-side helpers work without a GPU (no compute calls)."""
import ctypes
import os
import subprocess

import pytest

from graph_colouring_with_rl_b200 import _lib, build


def test_library_builds_and_exports_every_declared_symbol():
    path = build.build()
    assert os.path.isfile(path)
    lib = _lib.load(build_if_missing=False)
    declared = _lib.header_functions()
    assert len(declared) >= 28
    assert sorted(declared) == sorted(_lib.SIGNATURES)          # binding table == header
    for name in declared:
        assert hasattr(lib, name), name
    out = subprocess.run(['nm', '-D', '--defined-only', path], capture_output=True, text=True).stdout
    exported = sorted(l.split()[-1] for l in out.splitlines() if ' T gcrl_' in l)
    assert exported == sorted(declared)                          # nothing else leaks out (-fvisibility=hidden)


def test_sass_targets_sm_100a():
    out = subprocess.run(['cuobjdump', '-lelf', build.LIB_PATH], capture_output=True, text=True).stdout
    assert 'sm_100a' in out


def test_param_layout_matches_state_dict():
    from oracle import gn_oracle
    lib = _lib.load()
    off = (ctypes.c_int64 * 47)()
    assert lib.gcrl_param_layout(2, 1, 0, off) == 0
    shapes = gn_oracle.param_shapes(2, 1, 0)
    sizes = [int(__import__('numpy').prod(s)) for s in shapes.values()]
    assert [off[i + 1] - off[i] for i in range(46)] == sizes
    assert off[46] == 198529
    assert lib.gcrl_param_layout(2, 1, 3, off) == 0 and off[46] == 198529 + 64 * 3


def test_argument_errors_do_not_throw_across_the_abi():
    lib = _lib.load()
    assert lib.gcrl_version() >= 100
    assert lib.gcrl_param_layout(2, 1, 0, None) == -1           # GCRL_ERR_INVALID
    assert b'invalid argument' in lib.gcrl_last_error()
    assert lib.gcrl_pack(None, 5, 3, None, None, None, None, None, None, 0, None) == -1
    assert lib.gcrl_pna_forward(None, None, -1, 0, 64, None, None, None, None, None) == -1
    assert lib.gcrl_gn_forward(None, 0, 1, 0, 4, 0, None, None, None, None, None, None, None, None, None, None,
                               None, 0, 0, None) == -1
    assert lib.gcrl_gn_forward_workspace_bytes(100, 1000, 0) > lib.gcrl_gn_forward_workspace_bytes(100, 1000, 1)
    assert lib.gcrl_gn_stash_bytes(100, 1000) >= 4 * 1000 * 256
    assert lib.gcrl_launch_count() == 0


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip('CUDA present')
    from graph_colouring_with_rl_b200.architecture import GNNetwork
    from graph_colouring_with_rl_b200 import ops
    with pytest.raises(_lib.GcrlError):
        GNNetwork(2, 1, 0, 'GN', 1e-3)
    with pytest.raises(_lib.GcrlError):
        ops.pack(torch.zeros(2, 3, dtype=torch.int64), 3)
    with pytest.raises(_lib.GcrlError):
        ops.pna_forward(torch.zeros(3, 64), torch.zeros(2, dtype=torch.int32), 1)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.dirname(build.__file__)
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith('.py'):
                src = open(os.path.join(root, f)).read()
                assert 'import oracle' not in src and 'from oracle' not in src, f
