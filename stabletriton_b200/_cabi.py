"""ctypes binding of libstabletriton_b200.so -- the only path from Python to the sm_100a kernels.

There is deliberately no fallback: if the library is missing or fails to load, every op raises.
Prototypes mirror include/stabletriton_b200.h one to one.
"""
from __future__ import annotations

import ctypes
import os
import threading
from ctypes import c_float, c_int, c_longlong, c_size_t, c_uint, c_ulonglong, c_void_p, c_char_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libstabletriton_b200.so")

_lib = None
_lock = threading.Lock()

P, I, LL, F, U = c_void_p, c_int, c_longlong, c_float, c_uint

# name -> (restype, argtypes); every symbol the header declares
PROTOTYPES = {
    "st_version": (I, []),
    "st_last_error_string": (c_char_p, []),
    "st_launch_count": (c_ulonglong, []),
    "st_reset_launch_count": (None, []),
    "st_workspace_bytes": (c_size_t, []),
    "st_set_workspace": (I, [P, c_size_t]),
    "st_groupnorm_workspace_bytes": (c_size_t, [I, I, I, I]),
    "st_groupnorm_nhwc_bf16": (I, [P, P, P, P, P, I, I, I, I, F, I, P]),
    "st_layernorm_bf16": (I, [P, I, P, I, P, P, I, I, F, P]),
    "st_geglu_bf16": (I, [P, I, P, I, P, I, I, I, P]),
    "st_groupnorm_from_partials_nhwc_bf16": (I, [P, P, P, P, P, I, I, I, I, F, I, P, I, P, I, P]),
    "st_gemm_bf16": (I, [P, I, P, I, P, I, I, I, I, P, P, I, U, I, P, P]),
    "st_linear_small_m_bf16": (I, [P, I, P, I, P, P, I, I, I, I, I, I, U, P]),
    "st_conv3x3_nhwc_bf16": (I, [P, P, P, P, I, I, I, I, I, P, I, P, U, I, P, P]),
    "st_conv3x3_direct_bf16": (I, [P, LL, LL, LL, LL, P, P, P, LL, LL, LL, LL, I, I, I, I, I, P]),
    "st_im2col3x3_nhwc_bf16": (I, [P, P, I, I, I, I, I, P]),
    "st_upsample_nearest2x_nhwc_bf16": (I, [P, P, I, I, I, I, P]),
    "st_im2col3x3_smallc_bf16": (I, [P, LL, LL, LL, LL, P, I, I, I, I, P]),
    "st_nhwc_to_nchw_bf16": (I, [P, I, P, I, I, I, P]),
    "st_attention_bf16": (I, [P, LL, LL, LL, P, LL, LL, LL, P, LL, LL, LL, P, LL, LL, LL, I, I, I, I, F, P]),
    "st_timestep_embedding_bf16": (I, [P, P, I, I, I, P]),
    "st_concat_channels_bf16": (I, [P, I, P, I, P, LL, P]),
    "st_scale_model_input": (I, [P, P, LL, I, P, P, P]),
    "st_euler_cfg_update": (I, [P, P, P, LL, F, P, P, P]),
    "st_advance_step": (I, [P, P, P, P]),
    "st_softmax_rows_f32_bf16": (I, [P, LL, P, LL, I, I, F, P]),
    "st_pointwise_conv_small_bf16": (I, [P, P, P, P, I, LL, I, I, F, P]),
    "st_peer_slab_bytes": (c_size_t, [LL]),
    "st_peer_alloc": (I, [c_size_t, ctypes.POINTER(c_void_p)]),
    "st_peer_free": (I, [P]),
    "st_peer_export": (I, [P, c_char_p]),
    "st_peer_import": (I, [c_char_p, ctypes.POINTER(c_void_p)]),
    "st_peer_close": (I, [P]),
    "st_cfg_exchange_euler_update": (I, [P, I, P, P, P, LL, F, P, P, P]),
    "st_peer_error": (I, [P, ctypes.POINTER(c_uint)]),
}

ST_EPI_SILU = 1
ST_EPI_GEGLU = 2
ST_W_STATIC = 4
ST_EPI_F32OUT = 8


class StableTritonError(RuntimeError):
    pass


_NOT_KERNELS = {"st_version", "st_last_error_string", "st_launch_count", "st_reset_launch_count",
                "st_groupnorm_workspace_bytes", "st_workspace_bytes", "st_set_workspace", "st_peer_slab_bytes",
                "st_peer_alloc", "st_peer_free", "st_peer_export", "st_peer_import", "st_peer_close", "st_peer_error"}
_recording = None  # list of (symbol, args) while a recording is active


class _Recorder:
    """Proxy over the CDLL that logs every kernel-launching call (used by bench.py to re-issue, and
    time in isolation, exactly the launches of one UNet forward)."""

    def __init__(self, handle):
        self._handle = handle

    def __getattr__(self, name):
        fn = getattr(self._handle, name)
        if name in _NOT_KERNELS:
            return fn

        def call(*args):
            _recording.append((name, args))
            return fn(*args)

        return call


def start_recording() -> None:
    global _recording
    _recording = []


def stop_recording() -> list:
    global _recording
    calls, _recording = _recording or [], None
    return calls


def replay(calls, stream: int) -> None:
    """Re-issue recorded launches on `stream` (every entry point takes the stream as its last argument)."""
    handle = _load()
    for name, args in calls:
        check(getattr(handle, name)(*args[:-1], stream), name)


def lib():
    """The shared library (a recording proxy while `start_recording()` is active)."""
    handle = _load()
    return _Recorder(handle) if _recording is not None else handle


def _load() -> ctypes.CDLL:
    """Load (once) and return the shared library; raises if it has not been built."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise StableTritonError(
                        f"{LIB_PATH} is missing: build it with `python -m stabletriton_b200.build` "
                        "(nvcc, sm_100a). There is no CPU or PyTorch fallback."
                    )
                handle = ctypes.CDLL(LIB_PATH)
                for name, (res, args) in PROTOTYPES.items():
                    fn = getattr(handle, name)  # AttributeError if the .so is stale
                    fn.restype = res
                    fn.argtypes = args
                _lib = handle
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = _load().st_last_error_string().decode("utf-8", "replace")
        kind = ValueError if rc == -1 else StableTritonError
        raise kind(f"{what or 'stabletriton_b200'} failed (code {rc}): {msg}")


def launch_count() -> int:
    return int(_load().st_launch_count())


def reset_launch_count() -> None:
    _load().st_reset_launch_count()
