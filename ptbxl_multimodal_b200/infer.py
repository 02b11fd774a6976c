"""InferStep: the model.eval() forward of the reference's evaluation / test loops
(src/training/loop.py:52-65, src/training/loop_demo.py:59-75, scripts/06_ecg_baseline_test.py:94-106,
07:92-107, 08:96-110) as ONE CUDA graph of six ecgb200 kernels on the tcgen05 tensor cores.

B200-first structure: the weights are static in eval mode, so everything that depends only on them is
done once per ``refresh()`` -- conv weights re-laid to the bf16 tcgen05 operand layout, eval-mode
BatchNorm1d and the conv bias folded into per-channel fp32 {scale, shift}, proj.weight transposed.
A batch then costs: input pack (fp32 NCL -> blocked channels-last bf16), four implicit-GEMM convs whose
epilogue applies scale/shift + ReLU + MaxPool1d(2) straight out of TMEM (only the pooled activations are
written: 108*T elements per window instead of 364*T, SURVEY 8d "fused inference"), the last of which
reduces over time for AdaptiveAvgPool1d instead of storing anything, and one fused head kernel
(proj -> [demo encoder -> FiLM] -> head -> sigmoid).  No host synchronisation; logits / probabilities
stay on the device until the caller reads them.

Numerics: bf16 operands, fp32 accumulation / BN / head (stated tolerance in tests/test_gpu_infer.py).
``precision="fp32x3"`` keeps the same kernels and graph but carries every activation and weight as two bf16 planes
(hi = bf16(x), lo = bf16(x - hi)) and computes hi*hi + lo*hi + hi*lo as three times the input channels of each
implicit GEMM (fp32 accumulation in TMEM): logits within ~1e-5 of the fp32 reference path -- the north_star's
1e-4 bar -- at about a third of the bf16 engine's rate and several times the CUDA-core fp32 module forward."""
from __future__ import annotations

from typing import Optional

import torch

from ._lib import lib, check, EcgB200Error, configure_timeouts
from .ecg_cnn import ECGCNN
from .ecg_multimodal import ECGMultimodal

F32 = torch.float32
BF16 = torch.bfloat16


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


class InferStep:
    def __init__(self, model, batch_size: int, seq_len: int, use_graph: bool = True, precision: str = "bf16"):
        if not isinstance(model, (ECGCNN, ECGMultimodal)):
            raise EcgB200Error("InferStep drives ecgb200 ECGCNN / ECGMultimodal models")
        self.model = model
        self.mm = isinstance(model, ECGMultimodal)
        self.bb = model.ecg_backbone if self.mm else model
        self.B, self.T = int(batch_size), int(seq_len)
        if self.B <= 0 or self.T < 16:
            raise EcgB200Error("InferStep needs batch_size >= 1 and seq_len >= 16 (four MaxPool1d(2) stages)")
        self.dev = next(model.parameters()).device
        if self.dev.type != "cuda":
            raise EcgB200Error("InferStep needs the model on a CUDA device (no CPU fallback)")
        if precision not in ("bf16", "fp32x3"):
            raise EcgB200Error("precision must be 'bf16' or 'fp32x3' (split-precision: fp32-accurate on the tensor cores)")
        self.x3 = precision == "fp32x3"
        self.precision = precision
        self.use_graph = use_graph
        self.graphs = None
        self.launches_per_batch = 6
        configure_timeouts()
        self._alloc()
        self.refresh()

    # ------------------------------------------------------------------ static buffers
    def _alloc(self):
        B, T, dev = self.B, self.T, self.dev
        e = lambda *s: torch.empty(*s, dtype=F32, device=dev)          # noqa: E731
        eb = lambda *s: torch.empty(*s, dtype=BF16, device=dev)        # noqa: E731
        blocks = list(self.bb.backbone)
        self.chan = [blocks[0].net[0].in_channels] + [b.net[0].out_channels for b in blocks]
        if any(c % 32 for c in self.chan[1:]) or max(self.chan[1:]) > 256 or self.chan[0] > 256:
            raise EcgB200Error("bf16 engine needs conv widths that are multiples of 32 and <= 256")
        self.cip = [(self.chan[0] + 15) // 16 * 16] + self.chan[1:4]
        self.L = [T, T // 2, T // 4, T // 8]
        self.nl = self.model.head.out_features
        self.feat = self.bb.proj.out_features
        if self.feat > 256 or self.chan[4] > 256:
            raise EcgB200Error("fused inference head covers feat_dim <= 256")
        # two input slots: the next batch can be copied in (H2D) while the graph of the other slot runs
        self.xs = [torch.zeros(B, self.chan[0], T, dtype=F32, device=dev) for _ in range(2)]
        self.demos = [None, None]
        if self.mm:
            dm = self.model.demo_encoder.mlp
            self.d0, self.hid = dm[0].in_features, dm[0].out_features
            if not (dm[2].in_features == dm[2].out_features == self.hid <= 64
                    and self.model.film_gen.in_features == self.hid
                    and self.model.film_gen.out_features == 2 * self.feat):
                raise EcgB200Error("fused inference head covers the reference's demo encoder (D0 -> H -> H, H <= 64)")
            self.demos = [torch.zeros(B, self.d0, dtype=F32, device=dev) for _ in range(2)]
            self.w2T = e(self.hid, self.hid)
            self.wfT = e(self.hid, 2 * self.feat)
        self.cur = 0
        self.rows = [B, B]                                         # live windows per slot (ragged last batch)
        if self.x3:
            # split tensors: [hi | lo | hi] planes of every activation, [w_hi | w_hi | w_lo] of every weight
            self.ct = [lib.ecgb200_split_channels(c) for c in self.chan[:4]]
            zb = lambda *s: torch.zeros(*s, dtype=BF16, device=dev)    # noqa: E731  (the padding plane is never written)
            self.acts = [zb(B, self.ct[0] // 8, T, 8)] + [zb(B, self.ct[l + 1] // 8, self.L[l] // 2, 8) for l in range(3)]
            self.wt = [eb(15, self.ct[l] // 8, self.chan[l + 1], 8) for l in range(4)]
        else:
            self.acts = [eb(B, self.cip[0] // 8, T, 8)] + \
                        [eb(B, self.chan[l + 1] // 8, self.L[l] // 2, 8) for l in range(3)]
            self.wt = [eb(15, self.cip[l] // 8, self.chan[l + 1], 8) for l in range(4)]
        self.scale = [e(self.chan[l + 1]) for l in range(4)]
        self.shift = [e(self.chan[l + 1]) for l in range(4)]
        self.nparts = 4 * ((self.L[3] + 127) // 128)
        self.gap_part = e(B, self.nparts, self.chan[4])
        self.wpT = e(self.chan[4], self.feat)
        self.z = e(B, self.feat)
        self.logits = e(B, self.nl)
        self.prob = e(B, self.nl)

    # ------------------------------------------------------------------ weight-dependent state
    def refresh(self):
        """Re-derive the folded / re-laid weights from the module's current parameters and BatchNorm running
        statistics.  Call after load_state_dict() or a training epoch; captured graphs stay valid (same buffers)."""
        st = torch.cuda.current_stream(self.dev).cuda_stream
        for l, blk in enumerate(self.bb.backbone):
            conv, bn = blk.net[0], blk.net[1]
            for t in (conv.weight, bn.weight, bn.bias, bn.running_mean, bn.running_var):
                if t.dtype != F32 or not t.is_cuda or not t.is_contiguous():
                    raise EcgB200Error("InferStep needs contiguous float32 CUDA parameters")
            if self.x3:
                check(lib.ecgb200_conv1d_prep_weights_split_bf16(conv.weight.data_ptr(), _p(self.wt[l]), self.chan[l + 1],
                                                                 self.chan[l], st), "prep_weights_split")
            else:
                check(lib.ecgb200_conv1d_prep_weights_bf16(conv.weight.data_ptr(), _p(self.wt[l]), None, self.chan[l + 1],
                                                           self.chan[l], st), "prep_weights")
            check(lib.ecgb200_bn_fold_f32(bn.weight.data_ptr(), bn.bias.data_ptr(), bn.running_mean.data_ptr(),
                                          bn.running_var.data_ptr(), _p(conv.bias), _p(self.scale[l]),
                                          _p(self.shift[l]), self.chan[l + 1], float(bn.eps), st), "bn_fold")
        check(lib.ecgb200_transpose_f32(self.bb.proj.weight.data_ptr(), _p(self.wpT), self.feat, self.chan[4], st),
              "transpose")
        if self.mm:
            dm = self.model.demo_encoder.mlp
            check(lib.ecgb200_transpose_f32(dm[2].weight.data_ptr(), _p(self.w2T), self.hid, self.hid, st), "transpose")
            check(lib.ecgb200_transpose_f32(self.model.film_gen.weight.data_ptr(), _p(self.wfT), 2 * self.feat,
                                            self.hid, st), "transpose")
        if self.graphs is not None and self._head_ptrs() != self._captured_ptrs:
            self.graphs = None                # parameters were re-allocated (e.g. adopted by a TrainStep): re-capture

    def _head_ptrs(self):
        """Device pointers of the parameters the head kernel reads directly (baked into captured graphs)."""
        m = self.model
        ts = [self.bb.proj.bias, m.head.weight, m.head.bias]
        if self.mm:
            dm = m.demo_encoder.mlp
            ts += [dm[0].weight, dm[0].bias, dm[2].bias, m.film_gen.bias]
        for t in ts:
            if t.dtype != F32 or not t.is_cuda or not t.is_contiguous():
                raise EcgB200Error("InferStep needs contiguous float32 CUDA parameters")
        return tuple(t.data_ptr() for t in ts)

    # ------------------------------------------------------------------ the kernel sequence
    def _enqueue(self, slot: int):
        st = torch.cuda.current_stream(self.dev).cuda_stream
        B = self.B
        if self.x3:
            check(lib.ecgb200_pack_input_split_bf16(_p(self.xs[slot]), _p(self.acts[0]), B, self.chan[0], self.T, st), "pack_split")
            for l in range(4):
                last = l == 3
                check(lib.ecgb200_conv1d_bn_relu_pool_infer_split_bf16(
                    _p(self.acts[l]), _p(self.wt[l]), _p(self.scale[l]), _p(self.shift[l]),
                    None if last else _p(self.acts[l + 1]), _p(self.gap_part) if last else None,
                    B, self.chan[l], self.chan[l + 1], self.L[l], st), f"conv_infer_split_L{l + 1}")
        else:
            check(lib.ecgb200_pack_input_bf16(_p(self.xs[slot]), _p(self.acts[0]), B, self.chan[0], self.T, st), "pack")
            for l in range(4):
                last = l == 3
                check(lib.ecgb200_conv1d_bn_relu_pool_infer_bf16(
                    _p(self.acts[l]), _p(self.wt[l]), _p(self.scale[l]), _p(self.shift[l]),
                    None if last else _p(self.acts[l + 1]), _p(self.gap_part) if last else None,
                    B, self.cip[l], self.chan[l + 1], self.L[l], st), f"conv_infer_L{l + 1}")
        m, bb = self.model, self.bb
        if self.mm:
            dm = m.demo_encoder.mlp
            args = (_p(self.demos[slot]), _p(dm[0].weight), _p(dm[0].bias), _p(self.w2T), _p(dm[2].bias),
                    _p(self.wfT), _p(m.film_gen.bias))
            d0, hid = self.d0, self.hid
        else:
            args = (None,) * 7
            d0 = hid = 0
        check(lib.ecgb200_infer_head_f32(_p(self.gap_part), self.nparts, 1.0 / (self.L[3] // 2), _p(self.wpT),
                                         _p(bb.proj.bias), *args, _p(m.head.weight), _p(m.head.bias), _p(self.z),
                                         _p(self.logits), _p(self.prob), B, self.chan[4], self.feat, d0, hid,
                                         self.nl, st), "infer_head")

    def conv4(self, x):
        """Grad-CAM front end on the tensor cores: eval forward of blocks 1-3 (fused epilogue), then the RAW output
        of the 4th Conv1d (what the reference's forward hook captures, src/interpretability/grad_cam_1d.py:36-43),
        returned as fp32 (rows, C4, T/8) together with block 4's eval bn_state {mean, rstd, scale, shift} for
        ecgb200_gradcam_f32.  Un-captured (seven launches); the buffers are overwritten by the next call."""
        n = int(x.shape[0]) if x.dim() == 3 else -1
        if n < 1 or n > self.B or tuple(x.shape[1:]) != tuple(self.xs[0].shape[1:]):
            raise EcgB200Error(f"InferStep was built for x{tuple(self.xs[0].shape)} (or fewer windows), "
                               f"got x{tuple(x.shape)}")
        st = torch.cuda.current_stream(self.dev).cuda_stream
        B, c4, L4 = self.B, self.chan[4], self.L[3]
        if not hasattr(self, "y4"):
            self.y4 = torch.empty(B, c4 // 8, L4, 8, dtype=BF16, device=self.dev)
            self.A = torch.empty(B, c4, L4, dtype=F32, device=self.dev)
            self.bnst4 = torch.empty(4, c4, dtype=F32, device=self.dev)
        self.xs[0][:n].copy_(x, non_blocking=True)
        self.rows[0] = n
        blk = self.bb.backbone[3]
        conv, bn = blk.net[0], blk.net[1]
        if self.x3:
            # split precision: blocks 1-3 fused, then the raw 4th conv straight to fp32 (B, C4, T/8)
            check(lib.ecgb200_pack_input_split_bf16(_p(self.xs[0]), _p(self.acts[0]), B, self.chan[0], self.T, st), "pack_split")
            for l in range(3):
                check(lib.ecgb200_conv1d_bn_relu_pool_infer_split_bf16(
                    _p(self.acts[l]), _p(self.wt[l]), _p(self.scale[l]), _p(self.shift[l]), _p(self.acts[l + 1]), None,
                    B, self.chan[l], self.chan[l + 1], self.L[l], st), f"conv_infer_split_L{l + 1}")
            check(lib.ecgb200_conv1d_fwd_split_f32(_p(self.acts[3]), _p(self.wt[3]), _p(conv.bias), _p(self.A), B, self.chan[3],
                                                   c4, L4, st), "conv4_split")
            check(lib.ecgb200_bn_eval_state_f32(bn.weight.data_ptr(), bn.bias.data_ptr(), bn.running_mean.data_ptr(),
                                                bn.running_var.data_ptr(), _p(self.bnst4), c4, float(bn.eps), st),
                  "bn_eval_state")
            return self.A[:n], self.bnst4
        check(lib.ecgb200_pack_input_bf16(_p(self.xs[0]), _p(self.acts[0]), B, self.chan[0], self.T, st), "pack")
        for l in range(3):
            check(lib.ecgb200_conv1d_bn_relu_pool_infer_bf16(
                _p(self.acts[l]), _p(self.wt[l]), _p(self.scale[l]), _p(self.shift[l]), _p(self.acts[l + 1]), None,
                B, self.cip[l], self.chan[l + 1], self.L[l], st), f"conv_infer_L{l + 1}")
        check(lib.ecgb200_conv1d_fwd_bf16(_p(self.acts[3]), _p(self.wt[3]), _p(conv.bias), _p(self.y4), B, self.cip[3],
                                          c4, L4, st), "conv4")
        check(lib.ecgb200_unpack_act_bf16(_p(self.y4), _p(self.A), B, c4, L4, st), "unpack")
        check(lib.ecgb200_bn_eval_state_f32(bn.weight.data_ptr(), bn.bias.data_ptr(), bn.running_mean.data_ptr(),
                                            bn.running_var.data_ptr(), _p(self.bnst4), c4, float(bn.eps), st),
              "bn_eval_state")
        return self.A[:n], self.bnst4

    def capture(self):
        if not self.use_graph:
            return
        torch.cuda.synchronize(self.dev)
        s = torch.cuda.Stream(device=self.dev)
        graphs = []
        for slot in (0, 1):
            with torch.cuda.stream(s):
                self._enqueue(slot)                      # warm-up (sets the kernels' shared-memory attributes)
            torch.cuda.synchronize(self.dev)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s):
                self._enqueue(slot)
            graphs.append(g)
        self.graphs = graphs
        self._captured_ptrs = self._head_ptrs()

    # ------------------------------------------------------------------ public API
    def load_batch(self, x, demo=None, slot=None):
        """Copy a batch (pinned host or device tensors) into an input slot, async on the current stream.
        slot=None: the idle slot, which then becomes the one run() uses."""
        n = int(x.shape[0]) if x.dim() == 3 else -1
        if n < 1 or n > self.B or tuple(x.shape[1:]) != tuple(self.xs[0].shape[1:]):
            raise EcgB200Error(f"InferStep was built for x{tuple(self.xs[0].shape)} (or fewer windows), "
                               f"got x{tuple(x.shape)}")
        s = (self.cur ^ 1) if slot is None else int(slot)
        # a ragged last batch fills the first n rows; the stale rows are computed and ignored (eval mode:
        # windows are independent)
        self.xs[s][:n].copy_(x, non_blocking=True)
        if self.mm:
            if demo is None or int(demo.shape[0]) != n:
                raise EcgB200Error("ECGMultimodal inference needs x_demo with one row per window")
            self.demos[s][:n].copy_(demo, non_blocking=True)
        self.rows[s] = n
        if slot is None:
            self.cur = s

    def run(self, slot=None):
        """Forward of whatever input slot `slot` (default: the current one) holds.  Returns the logits buffer
        (rows, num_labels) fp32 on the device, overwritten by the next run; ``self.prob`` = sigmoid(logits),
        ``self.z`` = the backbone features (B, feat_dim)."""
        if slot is not None:
            self.cur = int(slot)
        if self.use_graph:
            if self.graphs is None:
                self.capture()
            self.graphs[self.cur].replay()
        else:
            self._enqueue(self.cur)
        n = self.rows[self.cur]
        return self.logits if n == self.B else self.logits[:n]

    def __call__(self, x, demo=None):
        self.load_batch(x, demo)
        return self.run()
