"""ctypes binding of the C ABI declared in include/afa_b200.h."""
from __future__ import annotations

import ctypes
import os
import threading

from .build import library_path

AFA_DTYPE_F32 = 0
AFA_DTYPE_BF16 = 1
AFA_FLAG_LOGSCALE = 1
AFA_FLAG_SNAKE = 2
AFA_ERR_ALIGNMENT = -5

EXPORTED_SYMBOLS = (
    "afa_version",
    "afa_last_error",
    "afa_activation1d_fwd",
    "afa_activation1d_fwd_pitched",
    "afa_bwd_workspace_bytes",
    "afa_activation1d_bwd",
    "afa_amp_activation1d_fwd_cl",
    "afa_resblock_mean",
    "afa_tail_fwd_cl",
    "afa_amp_act_conv_supported",
    "afa_amp_act_conv_fwd_cl",
    "afa_logmel_num_frames",
    "afa_logmel_fwd",
    "afa_logmel_bwd_workspace_bytes",
    "afa_logmel_bwd",
    "afa_l1_partial_sums",
    "afa_compact_zero_frames",
    "afa_set_tuning",
    "afa_kernel_info",
    "afa_kernel_info_shape",
    "afa_launch_count",
)


class AfaError(RuntimeError):
    """A call into libafa_sm100.so failed (argument check or CUDA error)."""


_lock = threading.Lock()
_lib = None


def load_library(path: str | None = None) -> ctypes.CDLL:
    """Load libafa_sm100.so.  Raises if it has not been built -- there is no fallback path."""
    global _lib
    with _lock:
        if _lib is not None and path is None:
            return _lib
        p = path or os.environ.get("AFA_LIBRARY") or library_path()
        if not os.path.exists(p):
            raise AfaError(
                f"{p} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). The fused Activation1d has no CPU or eager fallback."
            )
        lib = ctypes.CDLL(p)
        vp, i64, i32, fp = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.POINTER(ctypes.c_float)
        lib.afa_version.restype = i32
        lib.afa_last_error.restype = ctypes.c_char_p
        lib.afa_activation1d_fwd.restype = i32
        lib.afa_activation1d_fwd.argtypes = [vp, vp, vp, vp, fp, fp, i64, i64, i64, i32, i32, vp]
        lib.afa_activation1d_fwd_pitched.restype = i32
        lib.afa_activation1d_fwd_pitched.argtypes = [vp, i64, vp, i64, vp, vp, fp, fp, i64, i64, i64, i32, i32, vp]
        lib.afa_bwd_workspace_bytes.restype = ctypes.c_size_t
        lib.afa_bwd_workspace_bytes.argtypes = [i64, i64, i64, i32]
        lib.afa_activation1d_bwd.restype = i32
        lib.afa_activation1d_bwd.argtypes = [vp, vp, vp, vp, vp, vp, vp, fp, fp, i64, i64, i64, i32, i32, vp,
                                             ctypes.c_size_t, vp]
        f32 = ctypes.c_float
        lib.afa_amp_activation1d_fwd_cl.restype = i32
        lib.afa_amp_activation1d_fwd_cl.argtypes = [vp, i64, vp, i64, vp, vp, i64, vp, i64, i64, vp, vp, fp, fp,
                                                    i64, i64, i64, i32, i32, vp]
        lib.afa_resblock_mean.restype = i32
        lib.afa_resblock_mean.argtypes = [ctypes.POINTER(vp), ctypes.POINTER(vp), i32, vp, f32, vp, i64, i64, i32, vp]
        lib.afa_tail_fwd_cl.restype = i32
        lib.afa_tail_fwd_cl.argtypes = [vp, i64, vp, vp, fp, fp, vp, vp, i32, vp, vp, i32, f32, vp, i32, i64,
                                        i64, i64, i64, i32, i32, vp]
        lib.afa_amp_act_conv_supported.restype = i32
        lib.afa_amp_act_conv_supported.argtypes = [i64, i32, i32, i32]
        lib.afa_amp_act_conv_fwd_cl.restype = i32
        lib.afa_amp_act_conv_fwd_cl.argtypes = [vp, i64, vp, i64, vp, vp, i64, vp, i64, vp, i64, vp, vp, fp, fp, vp, i32, i32,
                                                i64, i64, i64, i32, i32, vp]
        ip = ctypes.POINTER(ctypes.c_int32)
        lib.afa_logmel_num_frames.restype = i64
        lib.afa_logmel_num_frames.argtypes = [i64, i32, i32, i32]
        lib.afa_logmel_fwd.restype = i32
        lib.afa_logmel_fwd.argtypes = [fp, fp, i64, i64, i64, i32, i32, i32, i32, fp, fp, i32, ip, ip, ip, fp,
                                       f32, f32, f32, i32, vp]
        lib.afa_logmel_bwd_workspace_bytes.restype = ctypes.c_size_t
        lib.afa_logmel_bwd_workspace_bytes.argtypes = [i64, i64, i32, i32, i32]
        lib.afa_logmel_bwd.restype = i32
        lib.afa_logmel_bwd.argtypes = [fp, fp, fp, i64, i64, i64, i64, i32, i32, i32, i32, fp, fp, i32, ip, ip, ip, fp, ip, ip,
                                       f32, f32, f32, i32, fp, f32, fp, vp, ctypes.c_size_t, vp]
        lib.afa_l1_partial_sums.restype = i32
        lib.afa_l1_partial_sums.argtypes = [fp, fp, i64, fp, i32, vp]
        lib.afa_compact_zero_frames.restype = i32
        lib.afa_compact_zero_frames.argtypes = [vp, vp, vp, vp, i64, i32, i64, f32, vp]
        lib.afa_set_tuning.restype = i32
        lib.afa_set_tuning.argtypes = [i32, i32, i32]
        lib.afa_kernel_info.restype = i32
        lib.afa_kernel_info.argtypes = [i32, i32, i64, ctypes.POINTER(ctypes.c_int32)]
        lib.afa_kernel_info_shape.restype = i32
        lib.afa_kernel_info_shape.argtypes = [i32, i32, i64, i64, i64, ctypes.POINTER(ctypes.c_int32)]
        lib.afa_launch_count.restype = i64
        if path is None:
            _lib = lib
        return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load_library().afa_last_error().decode("utf-8", "replace")
        raise AfaError(f"{what} failed with code {rc}: {msg}")


def launch_count() -> int:
    return int(load_library().afa_launch_count())


def kernel_info(which: int, dtype_code: int, T: int, batch: int = 1, channels: int = 1) -> dict:
    """Resource usage of the kernel variant a [batch, channels, T] launch selects."""
    out = (ctypes.c_int32 * 6)()
    check(load_library().afa_kernel_info_shape(which, dtype_code, batch, channels, T, out), "afa_kernel_info_shape")
    keys = ("registers", "smem_bytes", "threads", "segment_elems", "ctas_per_sm", "launches")
    return dict(zip(keys, [int(v) for v in out]))


def set_tuning(which: int, chunks: int = 0, threads: int = 0) -> None:
    check(load_library().afa_set_tuning(which, chunks, threads), "afa_set_tuning")
