"""The C-ABI library loads and exports every symbol include/cd360.h declares (no compute calls:
this runs on the CPU-only box), and argument validation fails loudly without a GPU."""
import ctypes

import pytest

from custom_diffusion360_b200 import _lib


def test_header_and_binding_agree():
    assert sorted(_lib.SIGNATURES) == _lib.header_symbols()


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    for name in _lib.header_symbols():
        assert hasattr(lib, name), name
    assert lib.cd360_abi_version() == 7
    assert lib.cd360_strerror(0) == b"ok"
    assert b"aligned" in lib.cd360_strerror(-2)


def test_argument_validation_needs_no_gpu():
    lib = _lib.load()
    assert lib.cd360_gemm_bf16(None, None) == -5
    args = _lib.GemmArgs()
    assert lib.cd360_gemm_bf16(ctypes.byref(args), None) == -5
    assert lib.cd360_attention_bf16(None, 0, None, 0, None, 0, None, 0, 1, 1, 1, 1, None) == -5
    assert lib.cd360_layernorm_bf16(16, 16, 16, 16, 4, 63, 1e-5, None) == -1  # c not multiple of 8
    assert lib.cd360_geglu_pack_block(5120) == 128 and lib.cd360_geglu_pack_block(100) == -1
    # TN mode (weight gradients): plain epilogues only, 16-byte aligned rows — checked before any CUDA call
    tn = dict(a0=256, w=256, out=256, M=640, N=208, k0=1024, lda0=640, ldw=208, ldo=208, out_fp32=1, tn=1)
    assert lib.cd360_gemm_bf16(ctypes.byref(_lib.GemmArgs(**{**tn, "geglu": 1, "N": 256, "ldw": 256, "ldo": 128})), None) == -3
    assert lib.cd360_gemm_bf16(ctypes.byref(_lib.GemmArgs(**{**tn, "conv": 1})), None) == -3
    assert lib.cd360_gemm_bf16(ctypes.byref(_lib.GemmArgs(**{**tn, "ldw": 204})), None) == -2   # ldw < N / not 8-aligned
    assert lib.cd360_gemm_bf16(ctypes.byref(_lib.GemmArgs(**{**tn, "M": 636})), None) == -2     # M not a multiple of 8
    assert lib.cd360_gemm_bf16(ctypes.byref(_lib.GemmArgs(**{**tn, "lda0": 632})), None) == -2  # lda0 < M
    assert lib.cd360_groupnorm_workspace_floats(3, 4096) >= 3 * 64


def test_ops_reject_cpu_tensors():
    import torch

    from custom_diffusion360_b200 import ops
    with pytest.raises(_lib.Cd360Error):
        ops.layernorm(torch.zeros(4, 64, dtype=torch.bfloat16), torch.ones(64), torch.zeros(64))
