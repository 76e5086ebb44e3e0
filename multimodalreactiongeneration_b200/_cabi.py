"""ctypes binding of include/mrg_lstm.h (the C-ABI shared library built from csrc/)."""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_size_t, c_uint8, c_uint64, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "csrc", "libmrg_b200.so")

F_TRAIN, F_GENERIC_REC, F_SIMT_GEMM, F_ACCUMULATE, F_ZERO_STATE, F_TF32 = 1, 2, 4, 8, 16, 32
F_ACC_WEIGHTS = 256
F_BF16 = 64
F_GRU = 128
F_BWD_NO_WGRAD, F_BWD_WGRAD_ONLY = 2048, 4096
F_PACK_VALID = 512

EXPORTS = (
    "mrg_version", "mrg_last_error_string", "mrg_device_info", "mrg_lstm_workspace_bytes",
    "mrg_lstm_layer_forward", "mrg_lstm_layer_backward", "mrg_gemm_nt", "mrg_philox_mask",
    "mrg_launch_count", "mrg_profile_enable", "mrg_profile_read", "mrg_gemm_strided",
    "mrg_gemm_workspace_bytes", "mrg_layernorm_workspace_bytes", "mrg_residual_layernorm_forward",
    "mrg_residual_layernorm_backward", "mrg_adamw_flat", "mrg_debug_set_trace", "mrg_colsum", "mrg_colsum_workspace_bytes",
    "mrg_attention_forward", "mrg_attention_backward", "mrg_gru_forward", "mrg_gru_backward",
    "mrg_rollout_supported", "mrg_rollout_forward", "mrg_rollout_backward", "mrg_profile_kernel_name",
    "mrg_lstm_pack_floats", "mrg_split_tf32", "mrg_gemm_split_supported", "mrg_gemm_strided_split",
    "mrg_audio_features", "mrg_attention_set_mode", "mrg_copy_rows",
)


class DirWeights(ctypes.Structure):
    _fields_ = [("w_ih", c_void_p), ("w_hh", c_void_p), ("b_ih", c_void_p), ("b_hh", c_void_p),
                ("h0", c_void_p), ("c0", c_void_p)]


class DirGrads(ctypes.Structure):
    _fields_ = [("dw_ih", c_void_p), ("dw_hh", c_void_p), ("db", c_void_p), ("dh0", c_void_p),
                ("dc0", c_void_p)]


class RolloutWeights(ctypes.Structure):
    _fields_ = [("H", c_int), ("L", c_int), ("P", c_int), ("FB", c_int), ("relu", c_int), ("ln_eps", c_float),
                ("w_prev", c_void_p), ("w_prev_ld", ctypes.c_longlong),
                ("w_ih", c_void_p * 2), ("b_ih", c_void_p * 2), ("b_hh", c_void_p * 2),
                ("ln_g", c_void_p * 2), ("ln_b", c_void_p * 2),
                ("w1", c_void_p), ("b1", c_void_p), ("w2", c_void_p), ("b2", c_void_p)]


class RolloutReserve(ctypes.Structure):
    _fields_ = [("xs", c_void_p), ("gates", c_void_p), ("xhat", c_void_p), ("rstd", c_void_p), ("fact", c_void_p),
                ("prev", c_void_p)]


class RolloutGrads(ctypes.Structure):
    _fields_ = [("dy", c_void_p), ("df", c_void_p), ("dpre", c_void_p), ("dbase", c_void_p), ("dprev", c_void_p),
                ("dln_g", c_void_p), ("dln_b", c_void_p)]


class MrgError(RuntimeError):
    pass


_LIB = None
PROF_KINDS = ("rec_fwd", "rec_bwd", "gemm", "rollout_fwd", "rollout_bwd")


def lib() -> ctypes.CDLL:
    """Load the extension; fail loudly when it has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise MrgError(
            f"{LIB_PATH} is missing: build it with `python -m multimodalreactiongeneration_b200._build` "
            "(there is no CPU / PyTorch fallback for the LSTM path)")
    L = ctypes.CDLL(LIB_PATH)
    L.mrg_version.restype = c_int
    L.mrg_last_error_string.restype = c_char_p
    L.mrg_device_info.argtypes = [POINTER(c_int)] * 5
    L.mrg_device_info.restype = c_int
    L.mrg_lstm_workspace_bytes.argtypes = [c_int] * 5
    L.mrg_lstm_workspace_bytes.restype = c_size_t
    L.mrg_lstm_layer_forward.argtypes = [
        c_void_p, POINTER(DirWeights), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
        c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]
    L.mrg_lstm_layer_forward.restype = c_int
    L.mrg_lstm_layer_backward.argtypes = [
        c_void_p, POINTER(DirWeights), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
        c_void_p, c_void_p, POINTER(DirGrads), c_void_p, c_size_t,
        c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]
    L.mrg_lstm_layer_backward.restype = c_int
    L.mrg_gemm_nt.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p,
                              c_size_t, c_int, c_void_p]
    L.mrg_gemm_nt.restype = c_int
    L.mrg_philox_mask.argtypes = [c_uint64, c_uint64, c_float, c_int, c_int, c_int, c_void_p, c_void_p]
    L.mrg_philox_mask.restype = c_int
    LL = ctypes.c_longlong
    L.mrg_gemm_strided.argtypes = [c_void_p, LL, LL, c_void_p, LL, LL, c_void_p, c_void_p, LL, c_int, c_int,
                                   c_int, c_int, c_int, c_void_p, c_size_t, c_int, c_void_p]
    L.mrg_gemm_strided.restype = c_int
    L.mrg_lstm_pack_floats.argtypes = [c_int, c_int, c_int]
    L.mrg_lstm_pack_floats.restype = c_size_t
    L.mrg_split_tf32.argtypes = [c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]
    L.mrg_split_tf32.restype = c_int
    L.mrg_gemm_split_supported.argtypes = [c_int, c_int, c_int, LL, LL, LL, LL, LL]
    L.mrg_gemm_split_supported.restype = c_int
    L.mrg_gemm_strided_split.argtypes = [c_void_p, LL, LL, c_void_p, c_void_p, LL, LL, c_void_p, c_void_p, LL, c_int,
                                         c_int, c_int, c_int, c_void_p, c_size_t, c_int, c_void_p]
    L.mrg_gemm_strided_split.restype = c_int
    L.mrg_audio_features.argtypes = [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, LL, LL, c_int,
                                     c_int, c_int, c_int, c_void_p]
    L.mrg_audio_features.restype = c_int
    L.mrg_copy_rows.argtypes = [c_void_p, LL, LL, c_void_p, c_int, c_int, c_int, c_void_p]
    L.mrg_copy_rows.restype = c_int
    L.mrg_gemm_workspace_bytes.argtypes = [c_int, c_int, c_int]
    L.mrg_gemm_workspace_bytes.restype = c_size_t
    L.mrg_layernorm_workspace_bytes.argtypes = [c_int]
    L.mrg_layernorm_workspace_bytes.restype = c_size_t
    L.mrg_residual_layernorm_forward.argtypes = [c_void_p, LL, LL, c_void_p, LL, LL, c_void_p, c_void_p, c_void_p,
                                                 LL, LL, c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_void_p]
    L.mrg_residual_layernorm_forward.restype = c_int
    L.mrg_residual_layernorm_backward.argtypes = [c_void_p, LL, LL, c_void_p, LL, LL, c_void_p, LL, LL, c_void_p,
                                                  c_void_p, c_void_p, c_void_p, LL, LL, c_void_p, c_void_p, c_void_p,
                                                  c_size_t, c_int, c_int, c_int, c_void_p]
    L.mrg_residual_layernorm_backward.restype = c_int
    L.mrg_adamw_flat.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p, c_float, c_float, c_float,
                                 c_float, c_float, c_void_p, c_float, c_int, c_void_p]
    L.mrg_adamw_flat.restype = c_int
    L.mrg_debug_set_trace.argtypes = [c_void_p]
    L.mrg_colsum_workspace_bytes.argtypes = [c_int, c_int]
    L.mrg_colsum_workspace_bytes.restype = c_size_t
    L.mrg_colsum.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]
    L.mrg_colsum.restype = c_int
    L.mrg_attention_forward.argtypes = [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p,
                                        c_int, c_int, c_int, c_int, c_int, c_float, c_int, c_int, c_void_p, c_void_p,
                                        c_void_p]
    L.mrg_attention_forward.restype = c_int
    L.mrg_attention_set_mode.argtypes = [c_int]
    L.mrg_attention_set_mode.restype = c_int
    L.mrg_attention_backward.argtypes = [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p,
                                         c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p,
                                         c_int, c_int, c_int, c_int, c_int, c_float, c_int, c_int, c_void_p, c_void_p,
                                         c_void_p]
    L.mrg_attention_backward.restype = c_int
    L.mrg_gru_forward.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                  c_void_p]
    L.mrg_gru_forward.restype = c_int
    L.mrg_gru_backward.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                   c_int, c_int, c_void_p]
    L.mrg_gru_backward.restype = c_int
    L.mrg_rollout_supported.argtypes = [c_int] * 4
    L.mrg_rollout_supported.restype = c_int
    L.mrg_rollout_forward.argtypes = [c_void_p, c_void_p, c_void_p, POINTER(RolloutWeights), c_void_p,
                                      POINTER(RolloutReserve), c_int, c_int, c_void_p]
    L.mrg_rollout_forward.restype = c_int
    L.mrg_rollout_backward.argtypes = [c_void_p, c_void_p, POINTER(RolloutWeights), POINTER(RolloutReserve),
                                       POINTER(RolloutGrads), c_int, c_int, c_void_p]
    L.mrg_rollout_backward.restype = c_int
    L.mrg_launch_count.restype = ctypes.c_ulonglong
    L.mrg_profile_enable.argtypes = [c_int]
    L.mrg_profile_read.argtypes = [POINTER(c_float), POINTER(c_int)]
    L.mrg_profile_kernel_name.argtypes = [c_int]
    L.mrg_profile_kernel_name.restype = c_char_p
    _LIB = L
    if os.environ.get("MRG_PRECISION", "fp32") in ("tf32", "bf16"):   # reduced-precision mode chosen by the environment:
        L.mrg_attention_set_mode(1)                                      # the attention kernels run one tf32 pass too
    return L


def launch_count() -> int:
    return int(lib().mrg_launch_count())


def profile_enable(on: bool) -> None:
    lib().mrg_profile_enable(1 if on else 0)


def profile_read():
    """-> {"rec_fwd": (ms, n), "rec_bwd": .., "gemm": .., "rollout_fwd": .., "rollout_bwd": ..} since
    profile_enable(True)."""
    ms = (c_float * len(PROF_KINDS))()
    n = (c_int * len(PROF_KINDS))()
    check(lib().mrg_profile_read(ms, n), "mrg_profile_read")
    return {k: (float(ms[i]), int(n[i])) for i, k in enumerate(PROF_KINDS)}


def profile_kernel_name(kind: str) -> str:
    """Demangled name of the last timed kernel of ``kind`` (one of PROF_KINDS), as the library launched it."""
    return lib().mrg_profile_kernel_name(PROF_KINDS.index(kind)).decode()


def check(status: int, what: str) -> None:
    if status != 0:
        msg = lib().mrg_last_error_string().decode("utf-8", "replace")
        raise MrgError(f"{what} failed with status {status}: {msg}")


def contiguous3(t):
    """``t.contiguous()`` for a 3-D fp32 CUDA tensor whose rows (last dimension) are contiguous — the transposed views at
    the LSTM seam — through ``mrg_copy_rows`` (a row copy at HBM speed; torch's generic strided copy is ~6x slower on
    these shapes).  Anything else goes to torch."""
    if t.is_contiguous():
        return t
    if (t.dim() == 3 and t.is_cuda and t.dtype.is_floating_point and t.element_size() == 4 and t.stride(2) == 1
            and t.shape[2] % 4 == 0 and t.stride(0) % 4 == 0 and t.stride(1) % 4 == 0 and t.data_ptr() % 16 == 0
            and t.numel() > 0):
        import torch
        out = torch.empty(t.shape, dtype=t.dtype, device=t.device)
        with torch.cuda.device(t.device):
            st = lib().mrg_copy_rows(t.data_ptr(), t.stride(0), t.stride(1), out.data_ptr(), t.shape[0], t.shape[1],
                                     t.shape[2], torch.cuda.current_stream(t.device).cuda_stream)
        check(st, "mrg_copy_rows")
        return out
    return t.contiguous()


def ptr(t) -> int | None:
    return None if t is None else t.data_ptr()
