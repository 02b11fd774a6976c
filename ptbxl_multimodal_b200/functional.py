"""Autograd wrappers over the libecgb200 C ABI (fp32 exact path).

Each ``torch.autograd.Function`` below replaces one ATen op of the reference's
hot path with hand-written sm_100a kernels; tensors are only used as device
buffers (``data_ptr``) on torch's current CUDA stream.  There is no CPU path."""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch.autograd import Function

from ._lib import lib, check, ptr, stream, EcgB200Error

KSIZE = 15

_scratch = {}


def scratch(nbytes: int, device, tag: str = "ws") -> torch.Tensor:
    """Grow-only per-(device, tag) scratch buffer (caller-provided `ws` of the C ABI)."""
    key = (device.index if device.index is not None else torch.cuda.current_device(), tag)
    buf = _scratch.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=device)
        _scratch[key] = buf
    return buf


def _need_f32_cuda(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise EcgB200Error(f"{name}: expected a CUDA tensor, got {t.device} (ecgb200 has no CPU fallback)")
    if t.dtype != torch.float32:
        raise EcgB200Error(f"{name}: expected float32, got {t.dtype}")
    return t.contiguous()


# ------------------------------------------------------------------ Conv1d k=15 pad=7
def prep_conv_weights(w: torch.Tensor, need_dgrad: bool):
    co, ci, k = w.shape
    if k != KSIZE:
        raise EcgB200Error("ecgb200 Conv1d supports kernel_size=15, padding=7 only")
    wt = torch.empty((ci, KSIZE, co), dtype=torch.float32, device=w.device)
    wd = torch.empty((co, KSIZE, ci), dtype=torch.float32, device=w.device) if need_dgrad else None
    check(lib.ecgb200_conv1d_prep_weights_f32(ptr(w), ptr(wt), ptr(wd), co, ci, stream()), "conv1d_prep_weights")
    return wt, wd


def conv1d_raw(x: torch.Tensor, wt: torch.Tensor, bias: Optional[torch.Tensor], co: int,
               want_stats: bool) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    b, ci, l = x.shape
    y = torch.empty((b, co, l), dtype=torch.float32, device=x.device)
    stat = None
    if want_stats:
        ntiles = lib.ecgb200_conv1d_stat_tiles(b, l)
        stat = torch.empty((2, co, ntiles), dtype=torch.float32, device=x.device)
    check(lib.ecgb200_conv1d_fwd_f32(ptr(x), ptr(wt), ptr(bias), ptr(y), ptr(stat), b, ci, co, l, stream()),
          "conv1d_fwd")
    return y, stat


class Conv1dK15Fn(Function):
    """y = conv1d(x, w, b, padding=7); second output = BatchNorm partial statistics
    (non-differentiable side product of the conv epilogue, or None)."""

    @staticmethod
    def forward(ctx, x, w, b, want_stats: bool):
        x = _need_f32_cuda(x, "conv1d input")
        w = _need_f32_cuda(w, "conv1d weight")
        if x.dim() != 3 or x.shape[1] != w.shape[1]:
            raise EcgB200Error(f"conv1d: input {tuple(x.shape)} does not match weight {tuple(w.shape)}")
        need_dx = ctx.needs_input_grad[0]
        wt, wd = prep_conv_weights(w, need_dx)
        y, stat = conv1d_raw(x, wt, b, w.shape[0], want_stats)
        ctx.save_for_backward(x, wd if wd is not None else x.new_empty(0))
        ctx.shape_w = tuple(w.shape)
        ctx.has_bias = b is not None
        if stat is None:
            stat = x.new_empty(0)
        ctx.mark_non_differentiable(stat)
        return y, stat

    @staticmethod
    def backward(ctx, dy, _dstat):
        x, wd = ctx.saved_tensors
        co, ci, _ = ctx.shape_w
        b, _, l = x.shape
        dy = _need_f32_cuda(dy, "conv1d grad_output")
        dw = db = dx = None
        if ctx.needs_input_grad[1] or ctx.needs_input_grad[2]:
            dw = torch.empty(ctx.shape_w, dtype=torch.float32, device=x.device)
            db = torch.empty((co,), dtype=torch.float32, device=x.device)
            ws = scratch(lib.ecgb200_conv1d_wgrad_ws_bytes(b, ci, co, l), x.device)
            check(lib.ecgb200_conv1d_wgrad_f32(ptr(dy), ptr(x), ptr(dw), ptr(db), ptr(ws), b, ci, co, l, stream()),
                  "conv1d_wgrad")
            if not ctx.has_bias:
                db = None
        if ctx.needs_input_grad[0]:
            dx, _ = conv1d_raw(dy, wd, None, ci, False)
        return dx, dw, db, None


# ------------------------------------------------------------------ BN + ReLU + MaxPool (+GAP)
class BnReluPoolFn(Function):
    """(p, gap) = maxpool2(relu(batchnorm(y))), mean_t p.  Train mode uses batch statistics
    (taken from the conv epilogue partials `stat` when given) and updates the running
    statistics in place, exactly once per forward."""

    @staticmethod
    def forward(ctx, y, gamma, beta, running_mean, running_var, nbt, stat, training: bool,
                momentum: float, eps: float, want_gap: bool):
        y = _need_f32_cuda(y, "batchnorm input")
        b, c, l = y.shape
        if l < 2:
            raise EcgB200Error("MaxPool1d(2) needs at least 2 time steps")
        dev = y.device
        bn_state = torch.empty((4, c), dtype=torch.float32, device=dev)
        if training:
            if b * l <= 1:
                raise ValueError("Expected more than 1 value per channel when training")
            use_stat = stat is not None and stat.numel() > 0
            ws = None if use_stat else scratch(lib.ecgb200_bn_stats_ws_bytes(b, c, l), dev)
            check(lib.ecgb200_bn_train_stats_f32(ptr(y), ptr(stat) if use_stat else None, ptr(gamma), ptr(beta),
                                                 ptr(running_mean), ptr(running_var), ptr(nbt), ptr(bn_state),
                                                 ptr(ws), b, c, l, momentum, eps, stream()), "bn_train_stats")
        else:
            check(lib.ecgb200_bn_eval_state_f32(ptr(gamma), ptr(beta), ptr(running_mean), ptr(running_var),
                                                ptr(bn_state), c, eps, stream()), "bn_eval_state")
        p = torch.empty((b, c, l // 2), dtype=torch.float32, device=dev)
        gap = torch.empty((b, c), dtype=torch.float32, device=dev) if want_gap else None
        check(lib.ecgb200_bn_relu_pool_fwd_f32(ptr(y), ptr(bn_state), ptr(p), ptr(gap), b, c, l, stream()),
              "bn_relu_pool_fwd")
        ctx.save_for_backward(y, bn_state, gamma)
        ctx.training = training
        ctx.set_materialize_grads(False)
        if gap is None:
            gap = y.new_empty(0)
            ctx.mark_non_differentiable(gap)
        return p, gap

    @staticmethod
    def backward(ctx, dp, dgap):
        y, bn_state, gamma = ctx.saved_tensors
        b, c, l = y.shape
        lp = l // 2
        if dp is not None and dgap is not None:
            dp = dp + (dgap / lp).unsqueeze(-1)        # both outputs consumed: rare, combine
            dgap = None
        if dp is None and dgap is None:
            return (None,) * 11
        if dp is not None:
            dp = _need_f32_cuda(dp, "pool grad")
        else:
            dgap = _need_f32_cuda(dgap, "gap grad")
        dy = torch.empty_like(y)
        dgamma = torch.empty((c,), dtype=torch.float32, device=y.device)
        dbeta = torch.empty((c,), dtype=torch.float32, device=y.device)
        ws = scratch(lib.ecgb200_bn_bwd_ws_bytes(b, c), y.device)
        check(lib.ecgb200_bn_relu_pool_bwd_f32(ptr(y), ptr(bn_state), ptr(gamma), ptr(dp), ptr(dgap), ptr(dy),
                                               ptr(dgamma), ptr(dbeta), ptr(ws), b, c, l,
                                               1 if ctx.training else 0, stream()), "bn_relu_pool_bwd")
        return dy, dgamma, dbeta, None, None, None, None, None, None, None, None


# ------------------------------------------------------------------ Linear / FiLM / BCE
class LinearFn(Function):
    @staticmethod
    def forward(ctx, x, w, b, act: int):
        x = _need_f32_cuda(x, "linear input")
        w = _need_f32_cuda(w, "linear weight")
        if x.dim() != 2 or x.shape[1] != w.shape[1]:
            raise EcgB200Error(f"linear: input {tuple(x.shape)} does not match weight {tuple(w.shape)}")
        m, k = x.shape
        n = w.shape[0]
        y = torch.empty((m, n), dtype=torch.float32, device=x.device)
        check(lib.ecgb200_linear_fwd_f32(ptr(x), ptr(w), ptr(b), ptr(y), m, k, n, act, stream()), "linear_fwd")
        ctx.save_for_backward(x, w, y if act == 1 else x.new_empty(0))
        ctx.act = act
        ctx.has_bias = b is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, yact = ctx.saved_tensors
        m, k = x.shape
        n = w.shape[0]
        dy = _need_f32_cuda(dy, "linear grad_output").clone()       # masked in place by the kernel
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        dw = torch.empty_like(w)
        db = torch.empty((n,), dtype=torch.float32, device=x.device) if ctx.has_bias else None
        check(lib.ecgb200_linear_bwd_f32(ptr(x), ptr(w), ptr(dy), ptr(yact) if ctx.act == 1 else None, ptr(dx),
                                         ptr(dw), ptr(db), m, k, n, stream()), "linear_bwd")
        return dx, dw, db, None


class FilmFn(Function):
    """zc = (1 + tanh(film[:, :F])) * z + film[:, F:]"""

    @staticmethod
    def forward(ctx, z, film):
        z = _need_f32_cuda(z, "film z")
        film = _need_f32_cuda(film, "film params")
        b, f = z.shape
        if film.shape != (b, 2 * f):
            raise EcgB200Error("film params must be (B, 2*feat_dim)")
        zc = torch.empty_like(z)
        check(lib.ecgb200_film_fwd_f32(ptr(z), ptr(film), ptr(zc), b, f, stream()), "film_fwd")
        ctx.save_for_backward(z, film)
        return zc

    @staticmethod
    def backward(ctx, dzc):
        z, film = ctx.saved_tensors
        b, f = z.shape
        dzc = _need_f32_cuda(dzc, "film grad")
        dz = torch.empty_like(z)
        dfilm = torch.empty_like(film)
        check(lib.ecgb200_film_bwd_f32(ptr(z), ptr(film), ptr(dzc), ptr(dz), ptr(dfilm), b, f, stream()), "film_bwd")
        return dz, dfilm


class BceWithLogitsFn(Function):
    @staticmethod
    def forward(ctx, logits, target):
        logits = _need_f32_cuda(logits, "logits")
        target = _need_f32_cuda(target, "target")
        if logits.shape != target.shape:
            raise ValueError(f"Target size ({tuple(target.shape)}) must be the same as input size ({tuple(logits.shape)})")
        loss = torch.empty((), dtype=torch.float32, device=logits.device)
        dlogits = torch.empty_like(logits)
        check(lib.ecgb200_bce_logits_f32(ptr(logits), ptr(target), ptr(loss), ptr(dlogits), None,
                                         logits.numel(), 1.0, stream()), "bce_logits")
        ctx.save_for_backward(dlogits)
        return loss

    @staticmethod
    def backward(ctx, dloss):
        (dlogits,) = ctx.saved_tensors
        return dlogits * dloss, None


def binary_cross_entropy_with_logits(logits: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """Drop-in for F.binary_cross_entropy_with_logits(logits, y) (mean reduction)."""
    return BceWithLogitsFn.apply(logits, target)


def sigmoid(logits: torch.Tensor) -> torch.Tensor:
    logits = _need_f32_cuda(logits, "logits")
    prob = torch.empty_like(logits)
    check(lib.ecgb200_bce_logits_f32(ptr(logits), None, None, None, ptr(prob), logits.numel(), 1.0, stream()),
          "sigmoid")
    return prob


def eval_counts(logits: torch.Tensor, target: Optional[torch.Tensor] = None, counts: Optional[torch.Tensor] = None,
                threshold: float = 0.5):
    """Device-side evaluation epilogue: returns (prob fp32, pred uint8) and, when `target` and `counts`
    (int32 [C,4] = tp, fp, fn, tn) are given, accumulates the confusion counts with no host synchronisation."""
    lg = _need_f32_cuda(logits, "logits")
    rows, c = lg.shape
    prob = torch.empty_like(lg)
    pred = torch.empty(rows, c, dtype=torch.uint8, device=lg.device)
    tg = _need_f32_cuda(target, "target") if target is not None else None
    if counts is not None and (counts.dtype != torch.int32 or tuple(counts.shape) != (c, 4) or not counts.is_cuda):
        raise EcgB200Error("counts must be a CUDA int32 tensor of shape (C, 4)")
    check(lib.ecgb200_eval_counts_f32(ptr(lg), ptr(tg), ptr(prob), ptr(pred), ptr(counts), rows, c, float(threshold),
                                      stream()), "eval_counts")
    return prob, pred


def linear(x, w, b=None, act: int = 0):
    return LinearFn.apply(x, w, b, act)


def zscore(x: torch.Tensor) -> torch.Tensor:
    """Per-lead (x-mean)/(std+1e-6) over the last dim (src/datasets/ptbxl.py:122-127)."""
    x = _need_f32_cuda(x, "zscore input")
    out = torch.empty_like(x)
    t = x.shape[-1]
    check(lib.ecgb200_zscore_f32(ptr(x), ptr(out), x.numel() // t, t, stream()), "zscore")
    return out
