"""The C-ABI library without a GPU: it loads, exports every symbol include/stabletriton_b200.h declares,
and validates arguments before touching CUDA (no compute calls here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    with open(os.path.join(ROOT, "include", "stabletriton_b200.h")) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(st_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound(built_lib):
    from stabletriton_b200 import _cabi
    lib = ctypes.CDLL(built_lib)
    declared = _declared()
    assert len(declared) >= 20
    for sym in declared:
        assert hasattr(lib, sym), f"{sym} declared in the header but not exported"
        assert sym in _cabi.PROTOTYPES, f"{sym} has no ctypes prototype"
    assert sorted(_cabi.PROTOTYPES) == declared, "ctypes prototypes and header disagree"
    assert _cabi.lib().st_version() == 100


def test_argument_validation_returns_error_codes(built_lib):
    from stabletriton_b200 import _cabi
    L = _cabi.lib()
    # null pointers
    assert L.st_gemm_bf16(0, 64, 0, 64, 0, 64, 128, 128, 64, 0, 0, 0, 0, 0, 0, 0) == -1
    assert b"null" in L.st_last_error_string()
    # K not a multiple of 64
    assert L.st_gemm_bf16(16, 72, 16, 72, 16, 128, 128, 128, 72, 0, 0, 0, 0, 0, 0, 0) == -1
    assert b"multiple of 64" in L.st_last_error_string()
    # misaligned pointer
    assert L.st_gemm_bf16(8, 64, 16, 64, 16, 128, 128, 128, 64, 0, 0, 0, 0, 0, 0, 0) == -1
    assert b"aligned" in L.st_last_error_string()
    # GEGLU and SiLU are exclusive
    assert L.st_gemm_bf16(16, 64, 16, 64, 16, 128, 128, 128, 64, 0, 0, 0, 3, 0, 0, 0) == -1
    # GroupNorm: channels not divisible by groups; too few channels per group
    assert L.st_groupnorm_nhwc_bf16(16, 16, 16, 16, 16, 1, 64, 320, 33, 1e-5, 1, 0) == -1
    assert L.st_groupnorm_nhwc_bf16(16, 16, 16, 16, 16, 1, 64, 64, 32, 1e-5, 1, 0) == -1
    assert L.st_groupnorm_workspace_bytes(2, 16384, 320, 32) > 2 * 320 * 2 * 4
    assert L.st_groupnorm_workspace_bytes(2, 16384, 321, 32) == 0
    # LayerNorm width, conv channel / spatial constraints, attention strides
    assert L.st_layernorm_bf16(16, 644, 16, 644, 16, 16, 8, 644, 1e-5, 0) == -1
    assert L.st_conv3x3_nhwc_bf16(16, 16, 16, 16, 1, 16, 16, 60, 64, 0, 0, 0, 0, 0, 0, 0) == -1
    assert L.st_conv3x3_nhwc_bf16(16, 16, 16, 16, 1, 12, 12, 64, 64, 0, 0, 0, 0, 0, 0, 0) == -1
    assert L.st_attention_bf16(16, 64, 64, 60, 16, 64, 64, 64, 16, 64, 64, 64, 16, 64, 64, 64, 1, 1, 8, 8, 0.125, 0) == -1
    assert L.st_linear_small_m_bf16(16, 64, 16, 64, 0, 16, 64, 33, 64, 64, 0, 0, 0, 0) == -1


def test_python_wrappers_reject_cpu_tensors():
    import torch
    from stabletriton_b200 import kernels as K
    x = torch.zeros(2, 64, 8, 8, dtype=torch.bfloat16)
    with pytest.raises(ValueError, match="CUDA"):
        K.groupnorm_wrapper(x, 8, None, None, 1e-5)
    with pytest.raises(ValueError, match="CUDA"):
        K.linear(torch.zeros(4, 64, dtype=torch.bfloat16), torch.zeros(64, 64, dtype=torch.bfloat16))
    with pytest.raises(ValueError, match="CUDA"):
        K.attention(*(torch.zeros(1, 1, 8, 64, dtype=torch.bfloat16),) * 3, 0.125)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from stabletriton_b200 import _cabi
    monkeypatch.setattr(_cabi, "_lib", None)
    monkeypatch.setattr(_cabi, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_cabi.StableTritonError, match="no CPU or PyTorch fallback"):
        _cabi.lib()


def test_attention_kernel_choice_model():
    """The launcher sends a self-attention sweep to the resident kernel (three small CTAs per SM, every tile of the launch on
    the machine at once) or to the pipelined one (owns an SM) from a cycle model calibrated on B200 -- pure host arithmetic,
    checked here against the measured table in profiles/r02_attention_experiments.txt, section 9 (148 SMs)."""
    import ctypes
    from stabletriton_b200 import _cabi
    f = _cabi._load().st_debug_attention_prefers_resident
    f.restype, f.argtypes = ctypes.c_int, [ctypes.c_longlong, ctypes.c_int, ctypes.c_int]
    tiles = lambda b, h, tq: b * h * ((tq + 127) // 128)
    resident = [(2, 20, 1024, 1024), (3, 20, 1024, 1024), (4, 20, 1024, 1024), (16, 20, 1024, 1024), (2, 20, 1024, 256),
                (2, 20, 1024, 512)]
    pipelined = [(1, 20, 1024, 1024), (1, 37, 1024, 1024), (1, 5, 4096, 4096), (1, 10, 4096, 4096), (2, 10, 4096, 4096),
                 (1, 10, 16384, 16384), (2, 10, 2048, 2048)]
    for b, h, tq, tk in resident:
        assert f(tiles(b, h, tq), tk, 148) == (4 if tiles(b, h, tq) >= 8 * 148 else 1), (b, h, tq, tk)
    for b, h, tq, tk in [(8, 20, 1024, 1024), (4, 10, 4096, 4096), (16, 10, 4096, 4096)]:  # saturated: the four-CTA form
        assert f(tiles(b, h, tq), tk, 148) == 4, (b, h, tq, tk)
    for b, h, tq, tk in pipelined:
        assert f(tiles(b, h, tq), tk, 148) == 0, (b, h, tq, tk)
    assert f(320, 77, 148) == 0 and f(320, 128, 148) == 0  # one-block launches never reach the model
