"""FusedAdamW: torch.optim.AdamW semantics (the optimizer the reference builds at
scripts/03_train_ecg_baseline.py:133, 04:158-162, 05:130 and steps at
src/training/loop.py:34) as ONE multi-tensor sm_100a kernel launch.  Hyper-parameters and
the step counter live on the device, so the step is CUDA-graph capturable."""
from __future__ import annotations

import ctypes as C

import torch

from ._lib import lib, check, stream, EcgB200Error


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 1e-2, grad_scale: float = 1.0):
        if lr < 0 or eps < 0 or weight_decay < 0 or not (0 <= betas[0] < 1) or not (0 <= betas[1] < 1):
            raise ValueError("invalid AdamW hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.grad_scale = grad_scale

    # device-side hyper-parameter block / step counter of a param group -------------------
    def _hyper_key(self, group):
        return (group["lr"], group["betas"][0], group["betas"][1], group["eps"], group["weight_decay"],
                self.grad_scale)

    def device_state(self, group, device):
        """(hyper float[6], step_ctr int32[1]) device tensors of a param group; the hyper block
        is refreshed whenever a host-side value (e.g. an lr schedule) changed."""
        key = self._hyper_key(group)
        if group.get("_hyper_key") != key or group.get("_hyper") is None or group["_hyper"].device != device:
            group["_hyper"] = torch.tensor(key, dtype=torch.float32, device=device)
            group["_hyper_key"] = key
        if group.get("_step_dev") is None or group["_step_dev"].device != device:
            group["_step_dev"] = torch.tensor([group.get("step", 0)], dtype=torch.int32, device=device)
        return group["_hyper"], group["_step_dev"]

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            gs, ms, vs = [], [], []
            for p in ps:
                if not p.is_cuda or p.dtype != torch.float32:
                    raise EcgB200Error("FusedAdamW updates CUDA float32 parameters only (no CPU fallback)")
                st = self.state[p]
                if not st:
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                gs.append(g); ms.append(st["exp_avg"]); vs.append(st["exp_avg_sq"])
            hyper, step_dev = self.device_state(group, ps[0].device)
            n = len(ps)
            arr = C.c_void_p * n
            num = (C.c_int64 * n)(*[p.numel() for p in ps])
            check(lib.ecgb200_adamw_f32(n, arr(*[p.data_ptr() for p in ps]), arr(*[g.data_ptr() for g in gs]),
                                        arr(*[m.data_ptr() for m in ms]), arr(*[v.data_ptr() for v in vs]),
                                        num, hyper.data_ptr(), step_dev.data_ptr(), stream()), "adamw")
            group["step"] = group.get("step", 0) + 1
            for p in ps:
                torch.autograd.graph.increment_version(p)
        return loss
