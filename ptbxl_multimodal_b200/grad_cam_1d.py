"""Grad-CAM for the 1D ECG CNNs.

``GradCAM1D`` keeps the reference's API (src/interpretability/grad_cam_1d.py:7-103):
hooks on a Conv1d, ``generate_cam(input_tensor, class_idx, signal_length)``,
``.activations`` / ``.gradients`` populated after a call.  The forward / backward it
triggers run on the ecgb200 kernels, and the CAM itself (channel weights, weighted sum,
ReLU, min-max normalisation, linear upsample) is one more kernel instead of six ATen ops.

``gradcam_batch`` is the throughput path for "10k ECGs x 5 classes": one eval forward to
the raw 4th-conv output and ONE closed-form kernel for all classes -- no backward pass
(dScore/dA = v_k * s * mask / Lp in eval mode)."""
from __future__ import annotations

from typing import Optional

import torch

from . import functional as Fn
from ._lib import lib, check, ptr, stream, EcgB200Error
from .ecg_cnn import ECGCNN, ConvBlock
from .ecg_multimodal import ECGMultimodal


def _cam_kernel(A, bn_state, v, v_per_sample, T, variant, eps, want_lo=True):
    b, c, lq = A.shape
    k = v.shape[-2]
    dev = A.device
    up = T is not None and T != lq
    cam_lo = torch.empty((b, k, lq), dtype=torch.float32, device=dev) if (want_lo or not up) else None
    cam_hi = torch.empty((b, k, T), dtype=torch.float32, device=dev) if up else None
    arg = torch.empty((b, k), dtype=torch.int32, device=dev)
    check(lib.ecgb200_gradcam_f32(ptr(A), ptr(bn_state), ptr(v), 1 if v_per_sample else 0, ptr(cam_lo),
                                  ptr(cam_hi), ptr(arg), b, c, lq, k, T if up else 0, variant, eps, stream()),
          "gradcam")
    return cam_lo, cam_hi, arg


class GradCAM1D:
    def __init__(self, model, target_layer):
        self.model = model
        self.model.eval()
        self.target_layer = target_layer
        self.activations = None     # A: (N, C, L')
        self.gradients = None       # dY/dA: (N, C, L')
        self._register_hooks()

    def _forward_hook(self, module, input, output):
        self.activations = output.detach()

    def _backward_hook(self, module, grad_input, grad_output):
        self.gradients = grad_output[0].detach()

    def _register_hooks(self):
        import warnings
        self.target_layer.register_forward_hook(self._forward_hook)
        with warnings.catch_warnings():           # same (legacy) hook kind as the reference, :36
            warnings.simplefilter("ignore")
            self.target_layer.register_backward_hook(self._backward_hook)

    def generate_cam(self, input_tensor, class_idx, signal_length=None):
        """input_tensor (1, C, L) -> CAM (signal_length,) or (L',): normalise at L', then
        linear upsample (the reference's order, grad_cam_1d.py:92-101)."""
        self.model.zero_grad()
        output = self.model(input_tensor)
        logits = output[0] if isinstance(output, tuple) else output
        score = logits[0, class_idx]
        score.backward(retain_graph=True)
        A = self.activations[:1].contiguous()
        G = self.gradients[:1].contiguous()
        n, c, lq = A.shape
        w = torch.empty((1, 1, c), dtype=torch.float32, device=A.device)
        check(lib.ecgb200_row_mean_f32(ptr(G), ptr(w), c, lq, stream()), "row_mean")
        cam_lo, cam_hi, _ = _cam_kernel(A, None, w, True, signal_length, 1, 0.0)
        return (cam_hi if cam_hi is not None else cam_lo)[0, 0]


@torch.no_grad()
def conv4_activations(model, x: torch.Tensor):
    """Eval-mode forward up to the raw output of the 4th Conv1d: returns (A, bn_state4)."""
    backbone = model.backbone if isinstance(model, ECGCNN) else model.ecg_backbone.backbone
    h = x
    for blk in list(backbone)[:-1]:
        was = blk.net[1].training
        blk.net[1].training = False
        try:
            h = blk(h)
        finally:
            blk.net[1].training = was
    last: ConvBlock = backbone[-1]
    conv, bn = last.net[0], last.net[1]
    conv._want_stats = False
    A = conv(h)
    bn_state = torch.empty((4, bn.num_features), dtype=torch.float32, device=x.device)
    check(lib.ecgb200_bn_eval_state_f32(ptr(bn.weight), ptr(bn.bias), ptr(bn.running_mean), ptr(bn.running_var),
                                        ptr(bn_state), bn.num_features, float(bn.eps), stream()), "bn_eval_state")
    return A, bn_state


@torch.no_grad()
def gradcam_batch(model, x: torch.Tensor, x_demo: Optional[torch.Tensor] = None,
                  signal_length: Optional[int] = None, variant: str = "v1", eps: float = 1e-9,
                  return_lowres: bool = False, engine=None):
    """All-class Grad-CAM for a batch: returns (cam (N, K, T or L'), argmax (N, K) int32
    [, cam_lowres]).  variant 'v1' = GradCAM1D order; 'v2' = script order (upsample, then
    (cam-min)/(max+eps); eps 1e-9 in scripts 00/13, 1e-8 in script 12).  Per-sample
    normalisation, i.e. each row equals the reference's single-sample call.
    `engine`: optional InferStep built for this model -- the forward up to the raw conv-4 output then runs on the
    bf16 tensor-core path (throughput mode; bf16 tolerance instead of the fp32 path's exact peak indices)."""
    if isinstance(model, ECGCNN):
        wh, wp = model.head.weight, model.proj.weight
        if x_demo is not None:
            raise EcgB200Error("ECGCNN takes no demographic input")
    elif isinstance(model, ECGMultimodal):
        wh, wp = model.head.weight, model.ecg_backbone.proj.weight
        if x_demo is None:
            raise EcgB200Error("ECGMultimodal Grad-CAM needs x_demo")
    else:
        raise EcgB200Error("gradcam_batch supports ecgb200 ECGCNN / ECGMultimodal models")
    if engine is not None:
        if engine.model is not model:
            raise EcgB200Error("the InferStep engine was built for another model")
        A, bn_state = engine.conv4(x)
    else:
        A, bn_state = conv4_activations(model, x)
    n = x.shape[0]
    k, f = wh.shape
    wpt = wp.detach().t().contiguous()                      # (256, F): v = Wh @ Wp as linear(Wh, Wp^T)
    if x_demo is None:
        v = Fn.linear(wh.detach(), wpt)                     # (K, 256)
        per_sample = False
    else:
        h = model.demo_encoder(x_demo)
        film = Fn.linear(h, model.film_gen.weight, model.film_gen.bias).clone()
        film[:, f:] = 0                                     # only gamma = 1 + tanh(.) enters dz_cond/dz
        zc = Fn.FilmFn.apply(wh.detach().repeat(n, 1), film.repeat_interleave(k, dim=0))   # (N*K, F)
        v = Fn.linear(zc, wpt).reshape(n, k, -1)
        per_sample = True
    cam_lo, cam_hi, arg = _cam_kernel(A, bn_state, v.contiguous(), per_sample, signal_length,
                                      1 if variant == "v1" else 2, eps, want_lo=return_lowres)
    cam = cam_hi if cam_hi is not None else cam_lo
    return (cam, arg, cam_lo) if return_lowres else (cam, arg)


def compute_demo_importance(model, x_ecg, x_demo, class_idx):
    """abs(grad * input) on the demographic vector, max-normalised
    (scripts/12_grad_cam_ecg_demo.py:78-97)."""
    model.zero_grad()
    x_demo = x_demo.clone().detach().requires_grad_(True)
    logits = model(x_ecg, x_demo)
    score = logits[:, class_idx].sum()
    score.backward()
    grad = x_demo.grad[0].detach().cpu().numpy()
    val = x_demo.detach()[0].cpu().numpy()
    import numpy as np
    importance = np.abs(grad * val)
    if importance.max() > 0:
        importance = importance / importance.max()
    return importance
