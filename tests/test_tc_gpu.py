"""GPU self-test of the tcgen05 / TMA / TMEM plumbing (csrc/tc.cuh) through the C ABI."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def run(rows_total, row0, use_tma, store):
    from graph_colouring_with_rl_b200 import _lib
    lib = _lib.load()
    dev = 'cuda:0'
    torch.manual_seed(rows_total + row0)
    A = torch.randn(rows_total, 64, device=dev)
    A2 = torch.randn(rows_total, 64, device=dev) * 3
    W = torch.randn(64, 64, device=dev) * 0.2
    D = torch.full((128, 64), float('nan'), device=dev)
    DW = torch.full((64, 64), float('nan'), device=dev)
    out = torch.zeros(rows_total, 64, device=dev) if store else None
    P = ctypes.c_void_p
    rc = lib.gcrl_selftest_tc(P(A.data_ptr()), P(A2.data_ptr()), P(W.data_ptr()), rows_total, row0, P(D.data_ptr()),
                              P(DW.data_ptr()), P(out.data_ptr()) if store else None, use_tma,
                              P(torch.cuda.current_stream().cuda_stream))
    assert rc == 0, _lib.last_error()
    torch.cuda.synchronize()
    n = min(128, rows_total - row0)
    At = torch.zeros(128, 64, dtype=torch.float64, device=dev)
    A2t = torch.zeros(128, 64, dtype=torch.float64, device=dev)
    At[:n] = A[row0:row0 + n].double()
    A2t[:n] = A2[row0:row0 + n].double()
    D_ref = At @ W.double().T
    DW_ref = At.T @ A2t
    # bf16x3: ~2^-16 relative per operand
    assert float((D.double() - D_ref).abs().max()) < 2e-4 * float(D_ref.abs().max())
    assert float((DW.double() - DW_ref).abs().max()) < 2e-4 * float(DW_ref.abs().max())
    if store:
        assert torch.equal(out[row0:row0 + n], A[row0:row0 + n])
        assert float(out[:row0].abs().max() if row0 else 0) == 0


@pytest.mark.parametrize('use_tma', [0, 1, 2, 3])
def test_selftest_full_tile(use_tma):
    """use_tma bit 0: operand tiles brought in by TMA; bit 1: D through the A-in-TMEM product (tcgen05.mma with the A operand
    read from tensor memory, written there by tcgen05.st in the hi | lo half-tile layout of issue_gemm_feat_ts)."""
    run(300, 37, use_tma, bool(use_tma & 1))


def test_selftest_out_of_bounds_rows_are_zero_filled():
    run(200, 100, 1, True)
    run(200, 100, 0, False)
