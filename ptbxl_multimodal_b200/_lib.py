"""ctypes binding of libecgb200.so (the C ABI declared in include/ecgb200.h).

There is NO fallback: if the shared library cannot be loaded (or built in-tree
with nvcc) importing this module raises, and every op raises on non-CUDA
tensors.  PyTorch is used only for device memory and streams."""
from __future__ import annotations

import ctypes as C
import os

import torch  # noqa: F401  (loads libcudart.so.12 that the library links against)

from . import _build

_P = C.c_void_p
_I = C.c_int
_F = C.c_float
_Z = C.c_size_t

_SIGS = {
    "ecgb200_version": (_I, []),
    "ecgb200_arch": (_I, []),
    "ecgb200_conv1d_prep_weights_f32": (_I, [_P, _P, _P, _I, _I, _P]),
    "ecgb200_conv1d_fwd_f32": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "ecgb200_conv1d_stat_tiles": (_I, [_I, _I]),
    "ecgb200_conv1d_wgrad_f32": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "ecgb200_conv1d_wgrad_ws_bytes": (_Z, [_I, _I, _I, _I]),
    "ecgb200_bn_train_stats_f32": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _F, _F, _P]),
    "ecgb200_bn_stats_ws_bytes": (_Z, [_I, _I, _I]),
    "ecgb200_bn_eval_state_f32": (_I, [_P, _P, _P, _P, _P, _I, _F, _P]),
    "ecgb200_bn_relu_pool_fwd_f32": (_I, [_P, _P, _P, _P, _I, _I, _I, _P]),
    "ecgb200_bn_relu_pool_bwd_f32": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "ecgb200_bn_bwd_ws_bytes": (_Z, [_I, _I]),
    "ecgb200_linear_fwd_f32": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "ecgb200_linear_bwd_f32": (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P]),
    "ecgb200_film_fwd_f32": (_I, [_P, _P, _P, _I, _I, _P]),
    "ecgb200_film_bwd_f32": (_I, [_P, _P, _P, _P, _P, _I, _I, _P]),
    "ecgb200_bce_logits_f32": (_I, [_P, _P, _P, _P, _P, _I, _F, _P]),
    "ecgb200_eval_counts_f32": (_I, [_P, _P, _P, _P, _P, _I, _I, _F, _P]),
    "ecgb200_adamw_f32": (_I, [_I, _P, _P, _P, _P, _P, _P, _P, _P]),
    "ecgb200_gradcam_f32": (_I, [_P, _P, _P, _I, _P, _P, _P, _I, _I, _I, _I, _I, _I, _F, _P]),
    "ecgb200_zscore_f32": (_I, [_P, _P, _I, _I, _P]),
    "ecgb200_wfdb16_zscore_f32": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "ecgb200_wfdb16_zscore_pack_bf16": (_I, [_P, _P, _P, _P, _I, _I, _I, _P]),
    "ecgb200_row_mean_f32": (_I, [_P, _P, _I, _I, _P]),
    "ecgb200_pack_input_bf16": (_I, [_P, _P, _I, _I, _I, _P]),
    "ecgb200_unpack_act_bf16": (_I, [_P, _P, _I, _I, _I, _P]),
    "ecgb200_conv1d_prep_weights_bf16": (_I, [_P, _P, _P, _I, _I, _P]),
    "ecgb200_conv1d_fwd_bf16": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "ecgb200_conv1d_wgrad_bf16": (_I, [_P, _P, _P, _P, _P, _I, _P, _I, _I, _I, _I, _P]),
    "ecgb200_conv1d_wgrad_bf16_ws_bytes": (_Z, [_I, _I, _I, _I]),
    "ecgb200_bn_train_stats_bf16": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _F, _F, _P]),
    "ecgb200_bn_relu_pool_fwd_bf16": (_I, [_P, _P, _P, _P, _I, _I, _I, _P]),
    "ecgb200_bn_relu_pool_bwd_bf16": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "ecgb200_bn_nsplit": (_I, [_I, _I]),
    "ecgb200_bn_relu_pool_bwd_reduce_bf16": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _P]),
    "ecgb200_conv1d_fwd_stats_bf16": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "ecgb200_conv1d_stat_parts_bf16": (_I, [_I, _I, _I, _I]),
    "ecgb200_bn_relu_pool_fwd_train_bf16": (_I, [_P, _P, _I, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _F, _F, _I, _P]),
    "ecgb200_bn_relu_pool_fwd_train_route_bf16": (_I, [_P, _P, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _F, _F, _I, _P]),
    "ecgb200_bn_relu_pool_bwd_route_bf16": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "ecgb200_step_prep_bf16": (_I, [_P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _I, _I, _P, _P]),
    "ecgb200_head_fwd_bwd_f32": (_I, [_P] * 13 + [_I, _I, _I, _I, _F, _P]),
    "ecgb200_head_loss_parts": (_I, [_I]),
    "ecgb200_mm_head_fwd_bwd_f32": (_I, [_P] * 27 + [_I, _I, _I, _I, _I, _I, _F, _P]),
    "ecgb200_head_wgrad_multi_f32": (_I, [_I, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _P]),
    "ecgb200_head_wgrad_f32": (_I, [_P] * 10 + [_I, _I, _I, _I, _P]),
    "ecgb200_dp_adamw_fused_f32": (_I, [_P, _P, _P, _P, _P, C.c_int64, _I, _I, _P, _P, _P]),
    "ecgb200_dp_flag_words": (_I, [_I]),
    "ecgb200_dp_adamw_fused_range_f32": (_I, [_P, _P, _P, _P, _P, C.c_int64, C.c_int64, _I, _I, _P, _P, _P]),
    "ecgb200_dp_bn_sync_f32": (_I, [_P, _I, _I, _P, _P, _P, _I, _I, _P]),
    "ecgb200_dp_adamw_ll_f32": (_I, [_P, _P, _P, _P, _P, _P, C.c_int64, C.c_int64, _I, _I, _P, _P, _P]),
    "ecgb200_dp_ll_inbox_words": (_Z, [C.c_int64]),
    "ecgb200_dp_bn_sync_ll_f32": (_I, [_P, _I, _I, _P, _P, _P, _I, _I, _P]),
    "ecgb200_set_pdl": (_I, [_I]),
    "ecgb200_debug_set_trace": (_I, [_P]),
    "ecgb200_debug_set_diag": (_I, [_P]),
    "ecgb200_debug_set_cta_span": (_I, [_P]),
    "ecgb200_debug_set_conv_pair": (_I, [_I]),
    "ecgb200_set_spin_timeout_ms": (_I, [C.c_uint, C.c_uint]),
    "ecgb200_debug_stamp": (_I, [_P, _I, _P]),
    "ecgb200_adamw_flat_f32": (_I, [_P, _P, _P, _P, C.c_int64, _P, _P, _P]),
    "ecgb200_bn_relu_pool_bwd_apply_bf16": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "ecgb200_bn_fold_f32": (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _F, _P]),
    "ecgb200_conv1d_bn_relu_pool_infer_bf16": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "ecgb200_split_channels": (_I, [_I]),
    "ecgb200_pack_input_split_bf16": (_I, [_P, _P, _I, _I, _I, _P]),
    "ecgb200_conv1d_prep_weights_split_bf16": (_I, [_P, _P, _I, _I, _P]),
    "ecgb200_conv1d_bn_relu_pool_infer_split_bf16": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "ecgb200_conv1d_fwd_split_f32": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "ecgb200_infer_head_f32": (_I, [_P, _I, _F] + [_P] * 14 + [_I] * 6 + [_P]),
    "ecgb200_transpose_f32": (_I, [_P, _P, _I, _I, _P]),
}

EXPORTED = tuple(_SIGS)


def _load() -> C.CDLL:
    path = _build.LIB
    if _build.needs_build():
        try:
            _build.build()
        except Exception as e:  # no nvcc / compile error: only fatal if no prebuilt library
            if not os.path.exists(path):
                raise ImportError(
                    f"libecgb200.so is missing and could not be built ({e}); "
                    "run `python -c 'import __graft_entry__ as g; g.build()'`") from e
    lib = C.CDLL(path)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)           # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


class EcgB200Error(RuntimeError):
    pass


def check(rc: int, what: str) -> None:
    if rc != 0:
        if rc > 0:
            raise EcgB200Error(f"{what}: CUDA error {rc} ({torch.cuda.get_device_name() if torch.cuda.is_available() else 'no device'})")
        raise EcgB200Error(f"{what}: invalid/unsupported arguments (code {rc})")


_timeouts_done = False


def configure_timeouts(force: bool = False) -> None:
    """Apply ECGB200_SPIN_TIMEOUT_MS="<mbarrier ms>[,<peer ms>]" (0 = never trap) to the bounded spin waits of the
    kernels; needs a CUDA context, so the engines call it when they are built.  Unset: the library defaults
    (30 s for in-kernel barriers, 10 min for cross-rank waits)."""
    global _timeouts_done
    if _timeouts_done and not force:
        return
    _timeouts_done = True
    spec = os.environ.get("ECGB200_SPIN_TIMEOUT_MS")
    if not spec:
        return
    parts = [int(v) for v in spec.split(",")]
    mbar, peer = parts[0], parts[1] if len(parts) > 1 else parts[0]
    check(lib.ecgb200_set_spin_timeout_ms(mbar, peer), "set_spin_timeout")


def ptr(t):
    """Device pointer of a tensor (None -> NULL).  Refuses anything not on CUDA."""
    if t is None:
        return None
    if not t.is_cuda:
        raise EcgB200Error("ecgb200 ops run on CUDA tensors only (no CPU fallback)")
    if not t.is_contiguous():
        raise EcgB200Error("ecgb200 ops need contiguous tensors")
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream
