"""ctypes binding of libgcrl_b200.so (include/gcrl_b200.h).

There is no CPU fallback: if the library is missing it is built with nvcc; if that fails, or
a kernel entry point is called without a CUDA device, an exception is raised.
"""
import ctypes
import os
import re

import torch

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
# GCRL_LIB: load a diagnostic build (e.g. a trace build made with GCRL_LIB_OUT + GCRL_NVCC_FLAGS) instead
LIB_PATH = os.environ.get('GCRL_LIB') or os.path.join(_PKG_DIR, 'libgcrl_b200.so')
HEADER_PATH = os.path.join(os.path.dirname(_PKG_DIR), 'include', 'gcrl_b200.h')

GCRL_MATH_FP32 = 0
GCRL_MATH_BF16 = 1
GCRL_MATH_BF16X3 = 2

_lib = None

c_void_p, c_int, c_int32, c_int64, c_size_t, c_float = (
    ctypes.c_void_p, ctypes.c_int, ctypes.c_int32, ctypes.c_int64, ctypes.c_size_t, ctypes.c_float)
P = c_void_p

# name -> (restype, argtypes); must list every function include/gcrl_b200.h declares
# (tests/test_abi.py parses the header and checks both directions).
SIGNATURES = {
    'gcrl_last_error': (ctypes.c_char_p, []),
    'gcrl_version': (c_int, []),
    'gcrl_sm_count': (c_int, []),
    'gcrl_launch_count': (c_int64, []),
    'gcrl_prof_enable': (c_int, [c_int]),
    'gcrl_prof_read': (c_int, [P, P, c_int]),
    'gcrl_selftest_tc': (c_int, [P, P, P, c_int64, c_int, P, P, P, c_int, P]),
    'gcrl_param_layout': (c_int, [c_int, c_int, c_int, P]),
    'gcrl_pack_workspace_bytes': (c_size_t, [c_int64, c_int64]),
    'gcrl_pack': (c_int, [P, c_int64, c_int64, P, P, P, P, P, P, c_size_t, P]),
    'gcrl_pack_dense_workspace_bytes': (c_size_t, [c_int32]),
    'gcrl_pack_dense': (c_int, [P, P, P, c_int32, c_int64, c_int64, P, P, P, P, P, P, c_size_t, P]),
    'gcrl_permute_rows': (c_int, [P, P, P, c_int64, c_int32, c_int, P]),
    'gcrl_pna_forward': (c_int, [P, P, c_int64, c_int64, c_int32, P, P, P, P, P]),
    'gcrl_pna_backward': (c_int, [P, P, P, P, P, P, P, P, c_int64, c_int64, c_int32, P, P]),
    'gcrl_segment_reduce': (c_int, [P, P, c_int64, c_int64, c_int32, c_int, P, P]),
    'gcrl_gn_stash_bytes': (c_size_t, [c_int64, c_int64]),
    'gcrl_gn_forward_workspace_bytes': (c_size_t, [c_int64, c_int64, c_int]),
    'gcrl_gn_forward': (c_int, [P, c_int, c_int, c_int, c_int64, c_int64, P, P, P, P, P, P, P, P, P, P,
                                P, c_size_t, c_int, P]),
    'gcrl_gn_block_forward': (c_int, [P, c_int, c_int, c_int, c_int, c_int64, c_int64, P, P, P, P, P,
                                      P, P, P, c_size_t, c_int, P]),
    'gcrl_gn_backward_workspace_bytes': (c_size_t, [c_int64, c_int64]),
    'gcrl_gn_backward': (c_int, [P, c_int, c_int, c_int, c_int64, c_int64, P, P, P, P, P, P, P, P, P, P,
                                 P, c_int, P, c_size_t, c_int, P]),
    'gcrl_score_argmax': (c_int, [P, P, P, c_int32, P, P, P]),
    'gcrl_td_target_loss': (c_int, [P, P, P, P, c_int32, P, P, P, c_float, c_int32, P, P, P, P, P, P]),
    'gcrl_adam_soft_update': (c_int, [P, P, P, P, P, c_int64, c_float, c_float, c_float, c_float,
                                      c_int64, c_float, P]),
    'gcrl_soft_update': (c_int, [P, P, c_int64, c_float, P]),
    'gcrl_env_reset': (c_int, [P, P, P, P, c_int32, c_int64, P]),
    'gcrl_env_step': (c_int, [P, P, P, c_int32, P, P, P, P, P, P, P]),
    'gcrl_env_step_greedy': (c_int, [P, P, P, c_int32, P, P, P, P, P, P, P, P]),
    'gcrl_env_first_vertex': (c_int, [P, P, P, c_int32, P, P, P]),
    'gcrl_replay_store': (c_int, [c_int32, P, P, P, P, P, P, P, P, c_int32, P, P, P, P, P, P, P]),
    'gcrl_replay_gather': (c_int, [c_int32, P, P, P, P, P, P, P, c_int32, P, P, P, c_int32, P, P, P, P, P, P,
                                   P, P, P, P]),
}


def header_functions():
    """Names of the functions declared in include/gcrl_b200.h."""
    with open(HEADER_PATH) as f:
        text = f.read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(gcrl_[a-z0-9_]+)\s*\(', text)))


class GcrlError(RuntimeError):
    pass


def load(build_if_missing=True):
    """dlopen the library (building it first if absent) and attach the signatures."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        if not build_if_missing:
            raise GcrlError('libgcrl_b200.so not built (%s); run python -m graph_colouring_with_rl_b200.build' % LIB_PATH)
        from . import build as _build
        _build.build()
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error():
    return load().gcrl_last_error().decode(errors='replace')


def check(rc, what):
    if rc != 0:
        raise GcrlError('%s failed (%d): %s' % (what, rc, last_error()))


def require_cuda(t, name='tensor'):
    if not t.is_cuda:
        raise GcrlError('%s must be a CUDA tensor: graph_colouring_with_rl_b200 has no CPU path' % name)


def ptr(t):
    """Raw device pointer of a contiguous tensor (None -> NULL); a plain int, which ctypes takes for a void*."""
    if t is None:
        return None
    assert t.is_contiguous(), 'non-contiguous tensor passed to the C ABI'
    return t.data_ptr()


_raw_stream = getattr(torch._C, '_cuda_getCurrentRawStream', None)


def stream_ptr(device=None):
    """cudaStream_t of torch's current stream on `device` (the raw-handle fast path when this torch has it: a learn step
    at the reference's minibatch size makes ~15 library calls and is bound by host issue time)."""
    if _raw_stream is not None and device is not None and getattr(device, 'index', None) is not None:
        return _raw_stream(device.index)
    return torch.cuda.current_stream(device).cuda_stream


class _NoGuard(object):
    def __enter__(self):
        return None

    def __exit__(self, *a):
        return False


_NO_GUARD = _NoGuard()


def device_guard(dev):
    """`torch.cuda.device(dev)` only when dev is not already the current device: the context manager costs ~10 us per
    call, and a small learn step makes fifteen library calls (it is bound by host issue time)."""
    idx = dev.index
    if idx is None or idx == torch.cuda.current_device():
        return _NO_GUARD
    return torch.cuda.device(dev)


_ws_cache = {}


def workspace(nbytes, device, tag='ws'):
    """Grow-only byte workspace per (device, tag); stream-ordered reuse on the current stream."""
    key = (str(device), tag)
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf
