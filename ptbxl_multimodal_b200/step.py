"""TrainStep: the whole training step of the reference loop body
(zero_grad -> model(x) -> BCE -> backward -> AdamW.step, src/training/loop.py:26-36 and
src/training/loop_demo.py:30-41) as ONE CUDA graph of ecgb200 kernels over static buffers.

B200-first structure: parameters, gradients and both Adam moments live in four flat fp32
buffers (the nn.Module's parameters are re-pointed to views of the flat parameter buffer, so
``state_dict`` / checkpoints are untouched); activations live in preallocated buffers sized for
(batch, seq_len); the kernel sequence is enqueued once through the C ABI, captured with
``torch.cuda.CUDAGraph`` and replayed per step, so a step costs one launch from the host and
no host<->device synchronisation (the loss stays on the device until the caller reads it).

Data parallel (one process per GPU): gradients are summed with NCCL all-reduce in two buckets
-- the 4th conv block (65 % of the bytes, ready first) on a side stream while blocks 3..1
still run backward, the rest at the end -- and the 1/world_size average is folded into the
AdamW kernel (``gscale``).  BatchNorm statistics are per rank (torch DDP semantics)."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import torch

from ._lib import lib, check, EcgB200Error
from .ecg_cnn import ECGCNN
from .ecg_multimodal import ECGMultimodal
from .optim import FusedAdamW
from .parallel import padded_size, shard_bounds

F32 = torch.float32


def _p(t: Optional[torch.Tensor], off: int = 0):
    return None if t is None else t.data_ptr() + t.element_size() * off


class _Seg:
    """A parameter's slot in the flat buffers."""
    __slots__ = ("name", "param", "off", "n")

    def __init__(self, name, param, off):
        self.name, self.param, self.off, self.n = name, param, off, param.numel()


class TrainStep:
    def __init__(self, model, optimizer: FusedAdamW, batch_size: int, seq_len: int,
                 process_group=None, use_graph: bool = True, precision: str = "fp32", dp_mode: str = "auto"):
        if not isinstance(model, (ECGCNN, ECGMultimodal)):
            raise EcgB200Error("TrainStep drives ecgb200 ECGCNN / ECGMultimodal models")
        if not isinstance(optimizer, FusedAdamW) or len(optimizer.param_groups) != 1:
            raise EcgB200Error("TrainStep needs a single-group FusedAdamW")
        self.model, self.opt = model, optimizer
        self.mm = isinstance(model, ECGMultimodal)
        self.bb = model.ecg_backbone if self.mm else model
        self.B, self.T = int(batch_size), int(seq_len)
        if precision not in ("fp32", "bf16"):
            raise EcgB200Error("precision must be 'fp32' (CUDA-core exact path) or 'bf16' (tcgen05 path)")
        self.bf16 = precision == "bf16"
        self.precision = precision
        if self.T < 16:
            raise EcgB200Error("seq_len must be >= 16 (four MaxPool1d(2) stages)")
        self.pg = process_group
        self.world = 1
        if process_group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            self.world = torch.distributed.get_world_size(process_group)
        self.dev = next(model.parameters()).device
        if self.dev.type != "cuda":
            raise EcgB200Error("TrainStep needs the model on a CUDA device (no CPU fallback)")
        if dp_mode not in ("auto", "fused", "nccl"):
            raise EcgB200Error("dp_mode must be 'auto', 'fused' (peer-memory reduce-scatter + AdamW + all-gather "
                               "kernel) or 'nccl' (all-reduce, then a replicated AdamW)")
        # gradient exchange: the fused NVLink kernel needs the flat-buffer optimizer of the bf16 engine
        self.dp_fused = self.world > 1 and (dp_mode == "fused" or (dp_mode == "auto" and self.bf16))
        if self.dp_fused and not self.bf16:
            raise EcgB200Error("dp_mode='fused' is implemented for precision='bf16'")
        self.use_graph = use_graph
        self.graph = None
        self.launches_per_step = 0
        self._prof = None
        self._prof_tag = ""
        self._flatten()
        self._alloc()

    # ------------------------------------------------------------------ flat parameter space
    def _flatten(self):
        named = list(self.model.named_parameters())
        group = self.opt.param_groups[0]
        if {id(p) for _, p in named} != {id(p) for p in group["params"]}:
            raise EcgB200Error("the optimizer must hold exactly the model's parameters")
        last = ("ecg_backbone." if self.mm else "") + "backbone.3."
        order = [(n, p) for n, p in named if not n.startswith(last)] + \
                [(n, p) for n, p in named if n.startswith(last)]
        total = sum(p.numel() for _, p in order)
        dev = self.dev
        # padded so that the space splits into 16-byte aligned shards for any world size <= 8
        self.total_pad = padded_size(total)
        if self.dp_fused:
            self._alloc_symmetric(self.total_pad)
        else:
            self.P = torch.zeros(self.total_pad, dtype=F32, device=dev)
            self.G = torch.zeros(self.total_pad, dtype=F32, device=dev)
        self.M = torch.zeros(self.total_pad, dtype=F32, device=dev)
        self.V = torch.zeros(self.total_pad, dtype=F32, device=dev)
        self.seg = {}
        off = 0
        with torch.no_grad():
            for n, p in order:
                if p.dtype != F32:
                    raise EcgB200Error("parameters must be float32")
                s = _Seg(n, p, off)
                self.P[off:off + s.n].copy_(p.detach().reshape(-1))
                st = self.opt.state[p]
                if st:                                      # adopt existing optimizer moments
                    self.M[off:off + s.n].copy_(st["exp_avg"].reshape(-1))
                    self.V[off:off + s.n].copy_(st["exp_avg_sq"].reshape(-1))
                p.data = self.P[off:off + s.n].view(p.shape)
                p.grad = self.G[off:off + s.n].view(p.shape)
                st["exp_avg"] = self.M[off:off + s.n].view(p.shape)
                st["exp_avg_sq"] = self.V[off:off + s.n].view(p.shape)
                self.seg[n] = s
                off += s.n
        self.total = total
        if self.world > 1:
            # replicas must start identical (DDP broadcasts too); cheap, once
            torch.distributed.broadcast(self.P, src=torch.distributed.get_global_rank(self.pg, 0) if self.pg is not None else 0,
                                        group=self.pg)
        self.bucket_a_off = self.seg[last + "net.0.weight"].off     # block 4 = tail of the buffers
        self.hyper, self.step_dev = self.opt.device_state(group, dev)
        if self.world > 1:
            self.opt.grad_scale = 1.0 / self.world
            self.hyper, self.step_dev = self.opt.device_state(group, dev)

    def _alloc_symmetric(self, n):
        """Parameter / gradient / flag buffers in NVLink peer-mapped (symmetric) memory: every rank gets the
        device pointers of every other rank's buffers for the fused exchange kernel."""
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        group = self.pg if self.pg is not None else dist.group.WORLD
        self.rank = dist.get_rank(group)
        nflag = lib.ecgb200_dp_flag_words(self.world)
        self.P = symm.empty(n, dtype=F32, device=self.dev)
        self.G = symm.empty(n, dtype=F32, device=self.dev)
        self.flags = symm.empty(max(nflag, 64), dtype=torch.int32, device=self.dev)
        self.P.zero_(); self.G.zero_(); self.flags.zero_()
        torch.cuda.synchronize(self.dev)
        hp, hg, hf = (symm.rendezvous(t, group) for t in (self.P, self.G, self.flags))
        self._symm_handles = (hp, hg, hf)                    # keep the mappings alive
        ptrs = lambda h: [int(h.buffer_ptrs[r]) for r in range(self.world)]      # noqa: E731
        self.peer_p, self.peer_g, self.peer_f = ptrs(hp), ptrs(hg), ptrs(hf)
        if self.peer_p[self.rank] != self.P.data_ptr() or self.peer_g[self.rank] != self.G.data_ptr():
            raise EcgB200Error("symmetric-memory rendezvous returned unexpected local pointers")
        dist.barrier(group)

    def gather_optimizer_state(self):
        """dp_mode='fused' shards the Adam moments (each rank updates 1/world of them).  Before saving an
        optimizer checkpoint, call this on every rank: all-gathers the shards so that opt.state is complete."""
        if not self.dp_fused:
            return
        lo, hi = shard_bounds(self.total_pad, self.world)[self.rank]
        for buf in (self.M, self.V):
            shard = buf[lo:hi].clone()
            torch.distributed.all_gather_into_tensor(buf, shard, group=self.pg)

    def _refresh_views(self):
        """Re-point parameters at the flat buffers if something (e.g. load_state_dict keeps
        them, .to() does not) replaced their storage."""
        with torch.no_grad():
            for s in self.seg.values():
                if s.param.data_ptr() != self.P.data_ptr() + 4 * s.off:
                    self.P[s.off:s.off + s.n].copy_(s.param.detach().reshape(-1))
                    s.param.data = self.P[s.off:s.off + s.n].view(s.param.shape)
                s.param.grad = self.G[s.off:s.off + s.n].view(s.param.shape)

    # ------------------------------------------------------------------ static buffers
    def _alloc(self):
        B, T, dev = self.B, self.T, self.dev
        e = lambda *s: torch.empty(*s, dtype=F32, device=dev)   # noqa: E731
        blocks = list(self.bb.backbone)
        self.chan = [blocks[0].net[0].in_channels] + [b.net[0].out_channels for b in blocks]
        self.L = [T, T // 2, T // 4, T // 8]                      # conv lengths; pooled = L // 2
        self.nl = self.model.head.out_features
        self.feat = self.bb.proj.out_features
        # two input slots: the next batch can be copied in (H2D) while the graph of the other slot runs
        self.xs = [e(B, self.chan[0], T), e(B, self.chan[0], T)]
        self.ys = [e(B, self.nl), e(B, self.nl)]
        self.demos = [None, None]
        self.cur = 0
        self.x, self.y, self.demo = self.xs[0], self.ys[0], None
        self.acts = [self.x]                                       # input of conv l
        self.ybuf, self.stat, self.bnst, self.wt, self.wd = [], [], [], [], []
        for l in range(4):
            ci, co, L = self.chan[l], self.chan[l + 1], self.L[l]
            self.bnst.append(e(4, co))
            if self.bf16:
                continue
            self.ybuf.append(e(B, co, L))
            self.stat.append(e(2, co, lib.ecgb200_conv1d_stat_tiles(B, L)))
            self.wt.append(e(ci, 15, co))
            self.wd.append(e(co, 15, ci) if l > 0 else None)
            if l < 3:
                self.acts.append(e(B, co, L // 2))
        if self.bf16:
            eb = lambda *s: torch.empty(*s, dtype=torch.bfloat16, device=dev)   # noqa: E731
            if any(c % 32 for c in self.chan[1:]) or max(self.chan[1:]) > 256:
                raise EcgB200Error("bf16 path needs conv widths that are multiples of 32 and <= 256")
            self.cip = [(self.chan[0] + 15) // 16 * 16] + self.chan[1:4]          # padded input widths
            # blocked channels-last bf16 activations  [B][C/8][L][8]
            self.acts = [eb(B, self.cip[0] // 8, T, 8)] + [eb(B, self.chan[l + 1] // 8, self.L[l] // 2, 8) for l in range(3)]
            self.ybuf = [eb(B, self.chan[l + 1] // 8, self.L[l], 8) for l in range(4)]
            self.wt = [eb(15, self.cip[l] // 8, self.chan[l + 1], 8) for l in range(4)]
            self.wd = [None] + [eb(15, self.chan[l + 1] // 8, self.cip[l], 8) for l in range(1, 4)]
            # BN backward: reduce + apply as two launches.  The single cooperative launch
            # (ecgb200_bn_relu_pool_bwd_fused_bf16) measured SLOWER inside the step (466 vs 447 us at B=256): its
            # 2-blocks-per-SM shared-memory slices lower occupancy and, needing the whole GPU at once, it stops
            # overlapping with the weight-gradient branch.  Kept behind ECGB200_BN_FUSED=1 for experiments.
            import os
            use = os.environ.get("ECGB200_BN_FUSED") == "1"
            self.bn_fused = [lib.ecgb200_bn_bwd_fused_nsplit(B, self.chan[l + 1], self.L[l], 1 if l < 3 else 0) if use else 0
                             for l in range(4)]
            self.ndb = [self.bn_fused[l] or lib.ecgb200_bn_nsplit(B, self.chan[l + 1]) for l in range(4)]
            self.dbpart = [e(self.chan[l + 1], self.ndb[l]) for l in range(4)]
            # Option (off): dgrad of block l+1 also produces block l's BatchNorm-backward sums in its epilogue
            # (conv_tc_kernel<4>), so that the separate reduce launch disappears.  Measured SLOWER at batch 256 (0.462 vs
            # 0.431 ms/step): each dgrad grows by 10.5 us -- as much as the reduce kernel it replaces -- and that time is
            # spent in the exposed, SM-exclusive epilogue of a tensor kernel, while the reduce kernel could share the SMs
            # with the weight-gradient branch.  ECGB200_BN_DGRAD_FUSE=1 enables it.
            self.bn_dgrad_fuse = os.environ.get("ECGB200_BN_DGRAD_FUSE", "0") == "1" and not use
            self.nbw = [lib.ecgb200_conv1d_stat_parts_bf16(B, self.chan[l + 2], self.chan[l + 1], self.L[l + 1])
                        if self.bn_dgrad_fuse else 0 for l in range(3)]
            if self.bn_dgrad_fuse and min(self.nbw) <= 0:
                self.bn_dgrad_fuse = False
            self.bwpart = [e(self.nbw[l], 2, self.chan[l + 1]) if self.bn_dgrad_fuse else None for l in range(3)]
            self.stat = [None] * 4
            # per-CTA {sum, sumsq} partials written by the conv epilogue
            self.nstat = [lib.ecgb200_conv1d_stat_parts_bf16(B, self.cip[l], self.chan[l + 1], self.L[l]) for l in range(4)]
            if min(self.nstat) <= 0:
                raise EcgB200Error("bf16 conv kernel does not support this shape")
            self.statp = [e(self.nstat[l], 2, self.chan[l + 1]) for l in range(4)]
            self.wpT = e(self.chan[4], self.bb.proj.out_features)
            self.loss_part = e(lib.ecgb200_head_loss_parts(B))
        c4 = self.chan[4]
        self.gap = e(B, c4)
        self.dgap = e(B, c4)
        self.z = e(B, self.feat)
        self.dz = e(B, self.feat)
        self.logits = e(B, self.nl)
        self.dlogits = e(B, self.nl)
        self.loss = torch.zeros((), dtype=F32, device=dev)
        if self.mm:
            dm = self.model.demo_encoder.mlp
            self.demos = [e(B, dm[0].in_features), e(B, dm[0].in_features)]
            self.demo = self.demos[0]
            self.h1, self.dh1 = e(B, dm[0].out_features), e(B, dm[0].out_features)
            self.h2, self.dh2 = e(B, dm[2].out_features), e(B, dm[2].out_features)
            self.film, self.dfilm = e(B, 2 * self.feat), e(B, 2 * self.feat)
            self.zc, self.dzc = e(B, self.feat), e(B, self.feat)
        big = B * 32 * T                                          # every conv output has 32*T elems/sample
        act_dt = torch.bfloat16 if self.bf16 else F32
        self.dy = torch.empty(max(B * co * L for co, L in zip(self.chan[1:], self.L)), dtype=act_dt, device=dev)
        # bf16 engine: wgrad of block l runs on the side stream while block l-1's BN backward writes its
        # own dy, so dy ping-pongs between two buffers
        self.dy2 = torch.empty_like(self.dy) if self.bf16 else None
        self.dp = torch.empty(max(B * self.chan[l] * self.L[l] for l in range(1, 4)), dtype=act_dt, device=dev)
        wsfn = lib.ecgb200_conv1d_wgrad_bf16_ws_bytes if self.bf16 else lib.ecgb200_conv1d_wgrad_ws_bytes
        ws = max(wsfn(B, self.chan[l], self.chan[l + 1], self.L[l]) for l in range(4))
        if self.bf16:
            ws = max(ws, max(lib.ecgb200_bn_bwd_ws_bytes(B, c) for c in self.chan[1:]))
            self.ws2 = torch.empty(max(lib.ecgb200_bn_bwd_ws_bytes(B, c) for c in self.chan[1:]), dtype=torch.uint8, device=dev)
        ws = max(ws, max(lib.ecgb200_bn_bwd_ws_bytes(B, c) for c in self.chan[1:]))
        self.ws = torch.empty(ws, dtype=torch.uint8, device=dev)
        del big
        self.side = torch.cuda.Stream(device=dev) if (self.world > 1 or self.bf16) else None
        self.linear = False
        self.comm = torch.cuda.Stream(device=dev) if (self.world > 1 and self.bf16 and not self.dp_fused) else None

    # ------------------------------------------------------------------ the kernel sequence
    def _k(self, name, fn, *args):
        """One C-ABI call; with self._prof set, bracket it with CUDA events on the launch stream."""
        prof = self._prof
        if prof is None:
            check(fn(*args), name)
            return
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(fn(*args), name)
        e1.record()
        prof.append((name + self._prof_tag, e0, e1))

    def _seg_ptr(self, name, buf):
        return buf.data_ptr() + 4 * self.seg[name].off

    def _fwd_blocks_fp32(self, st, pre, Pp, blocks):
        B = self.B
        n = 0
        for l in range(4):
            ci, co, L = self.chan[l], self.chan[l + 1], self.L[l]
            k = f"{pre}backbone.{l}.net."
            bn = blocks[l].net[1]
            self._prof_tag = f"_L{l + 1}"
            self._k("prep", lib.ecgb200_conv1d_prep_weights_f32, Pp(k + "0.weight"), _p(self.wt[l]), _p(self.wd[l]), co, ci, st)
            self._k("conv_fwd", lib.ecgb200_conv1d_fwd_f32, _p(self.acts[l]), _p(self.wt[l]), Pp(k + "0.bias"), _p(self.ybuf[l]),
                                             _p(self.stat[l]), B, ci, co, L, st)
            self._k("bn_stats", lib.ecgb200_bn_train_stats_f32, _p(self.ybuf[l]), _p(self.stat[l]), Pp(k + "1.weight"), Pp(k + "1.bias"),
                                                 bn.running_mean.data_ptr(), bn.running_var.data_ptr(),
                                                 bn.num_batches_tracked.data_ptr(), _p(self.bnst[l]), None,
                                                 B, co, L, float(bn.momentum), float(bn.eps), st)
            self._k("bn_relu_pool", lib.ecgb200_bn_relu_pool_fwd_f32, _p(self.ybuf[l]), _p(self.bnst[l]),
                                                   _p(self.acts[l + 1]) if l < 3 else None,
                                                   _p(self.gap) if l == 3 else None, B, co, L, st)
            n += 4
        return n

    def _fwd_blocks_bf16(self, st, pre, Pp, blocks):
        """tcgen05 path: ONE prologue launch (input pack + the four weight re-layouts + proj transpose +
        step counter), then per block the persistent implicit-GEMM conv with BatchNorm statistics in
        its epilogue and one BN-finalise + ReLU + pool pass, on blocked channels-last bf16."""
        B = self.B
        n = 0
        self._prof_tag = ""
        PV, I4 = C.c_void_p * 4, C.c_int * 4
        wkeys = [f"{pre}backbone.{l}.net.0.weight" for l in range(4)]
        main = torch.cuda.current_stream(self.dev)
        # critical path: pack the input + block-1 weights (+ step counter); blocks 2-4 and the proj transpose
        # are re-laid beside the first conv on the side stream
        self._k("prep", lib.ecgb200_step_prep_bf16, _p(self.x), _p(self.acts[0]), B, self.chan[0], self.T, 1,
                PV(Pp(wkeys[0]), None, None, None), PV(_p(self.wt[0]), None, None, None), PV(None, None, None, None),
                I4(self.chan[1], 0, 0, 0), I4(self.chan[0], 0, 0, 0), None, None, 0, 0, self.step_dev.data_ptr(), st)
        ev_prep = torch.cuda.Event()
        ev_prep.record(main)

        def prep_rest():
            # released by `prep`, but enqueued AFTER conv 1 so that the critical node is created (and launched) first
            if self.linear:
                self.side = main
            else:
                self.side.wait_event(ev_prep)
            with torch.cuda.stream(self.side):
                self._k("prep_w", lib.ecgb200_step_prep_bf16, None, None, 0, 0, 0, 3,
                        PV(*[Pp(k) for k in wkeys[1:]], None), PV(*[_p(w) for w in self.wt[1:]], None),
                        PV(*[_p(w) for w in self.wd[1:]], None), I4(*self.chan[2:5], 0), I4(*self.chan[1:4], 0),
                        Pp(pre + "proj.weight"), _p(self.wpT),
                        self.feat, self.chan[4], None, self.side.cuda_stream)
                done = torch.cuda.Event()
                done.record(self.side)
            return done
        import os
        # creating prep_w's node after conv 1's measured slower (462 vs 448 us): conv 2 then waits for it
        prep_late = os.environ.get("ECGB200_PREPW_LATE", "0") == "1"
        prep_done = None if prep_late else prep_rest()
        n += 2
        for l in range(4):
            cip, co, L = self.cip[l], self.chan[l + 1], self.L[l]
            k = f"{pre}backbone.{l}.net."
            bn = blocks[l].net[1]
            self._prof_tag = f"_L{l + 1}"
            if l == 1:
                main.wait_event(prep_done)
            self._k("conv_fwd", lib.ecgb200_conv1d_fwd_stats_bf16, _p(self.acts[l]), _p(self.wt[l]), Pp(k + "0.bias"),
                    _p(self.ybuf[l]), _p(self.statp[l]), B, cip, co, L, st)
            if l == 0 and prep_late:
                self._prof_tag = ""
                prep_done = prep_rest()
                self._prof_tag = "_L1"
            self._k("bn_relu_pool", lib.ecgb200_bn_relu_pool_fwd_train_bf16, _p(self.ybuf[l]), _p(self.statp[l]),
                    self.nstat[l], Pp(k + "1.weight"), Pp(k + "1.bias"), bn.running_mean.data_ptr(),
                    bn.running_var.data_ptr(), bn.num_batches_tracked.data_ptr(), _p(self.bnst[l]),
                    _p(self.acts[l + 1]) if l < 3 else None, _p(self.gap) if l == 3 else None, B, co, L,
                    float(bn.momentum), float(bn.eps), st)
            n += 2
        return n

    def _fork_side(self, main):
        if self.linear:                      # single-stream schedule: "side" work is enqueued in line
            self.side = main
            return
        ev = torch.cuda.Event()
        ev.record(main)
        self.side.wait_event(ev)

    def _bwd_blocks_bf16(self, pre, Gp):
        """Critical path on the main stream: BN/ReLU/pool backward -> dgrad, block 4 down to 1.  The weight
        gradients (only AdamW needs them) run on the side stream beside it; dy ping-pongs so block l-1's
        BN backward never overwrites what block l's wgrad still reads."""
        B = self.B
        n = 0
        main = torch.cuda.current_stream(self.dev)
        st = main.cuda_stream
        dys = [self.dy, self.dy2]
        wg_done = [None, None]
        for l in (3, 2, 1, 0):
            ci, co, L = self.chan[l], self.chan[l + 1], self.L[l]
            k = f"{pre}backbone.{l}.net."
            self._prof_tag = f"_L{l + 1}"
            dy = dys[l & 1]
            if wg_done[l & 1] is not None:
                main.wait_event(wg_done[l & 1])                # wgrad of block l+2 has finished reading this dy
            if self.bn_dgrad_fuse and l < 3:
                # the sums came out of dgrad_{l+1}'s epilogue: second pass only
                self._k("bn_bwd", lib.ecgb200_bn_relu_pool_bwd_apply_bf16, _p(self.ybuf[l]), _p(self.bnst[l]), _p(self.dp),
                        _p(self.bwpart[l]), self.nbw[l], _p(dy), Gp(k + "1.weight"), Gp(k + "1.bias"), _p(self.dbpart[l]),
                        B, co, L, 1, st)
                n += 1
            else:
                self._k("bn_bwd", lib.ecgb200_bn_relu_pool_bwd_fused_bf16 if self.bn_fused[l] else lib.ecgb200_bn_relu_pool_bwd_bf16,
                        _p(self.ybuf[l]), _p(self.bnst[l]),
                        _p(self.dp) if l < 3 else None, _p(self.dgap) if l == 3 else None, _p(dy),
                        Gp(k + "1.weight"), Gp(k + "1.bias"), _p(self.dbpart[l]), _p(self.ws2), B, co, L, 1, st)
                n += 1 if self.bn_fused[l] else 2
            # Order matters: the tensor kernels cannot share an SM (TMEM + shared memory), so whichever of dgrad_l
            # (critical path) and wgrad_l (side) the graph launches first takes the GPU, and ready nodes are launched in
            # creation order.  Measured (same box, us/step): wgrad node created first ("early") 476 -- or 456 when a
            # slower head_wgrad happened to delay it; wgrad released only after dgrad completes ("after") 466;
            # both released by bn_bwd_l with dgrad's node created first ("both", the default) 453.
            import os
            mode = os.environ.get("ECGB200_WGRAD_ORDER", "both")
            early = mode == "early"
            ev_bn = None
            if mode == "both":                       # both released by bn_bwd_l, but dgrad's node is created first
                ev_bn = torch.cuda.Event()
                ev_bn.record(main)
            if l > 0 and not early:
                self._dgrad(l, dy, st)
                n += 1
            if ev_bn is not None and not self.linear:
                self.side.wait_event(ev_bn)
            else:
                self._fork_side(main)
            with torch.cuda.stream(self.side):
                self._k("wgrad", lib.ecgb200_conv1d_wgrad_bf16, _p(dy), _p(self.acts[l]), Gp(k + "0.weight"),
                        Gp(k + "0.bias"), _p(self.dbpart[l]), self.ndb[l], _p(self.ws), B, ci, co, L,
                        self.side.cuda_stream)
                wg_done[l & 1] = torch.cuda.Event()
                wg_done[l & 1].record(self.side)
                if l == 3 and self.world > 1 and not self.dp_fused:
                    # bucket A (block 4, the tail of G) is final: all-reduce it while blocks 3..1 run
                    self.comm.wait_event(wg_done[l & 1])
                    with torch.cuda.stream(self.comm):
                        torch.distributed.all_reduce(self.G[self.bucket_a_off:], group=self.pg)
            n += 2
            if l > 0 and early:
                self._dgrad(l, dy, st)
                n += 1
        return n

    def _dgrad(self, l, dy, st):
        """Input gradient of block l (= dp of block l-1); with bn_dgrad_fuse its epilogue also leaves block l-1's
        BatchNorm-backward sums in bwpart[l-1]."""
        B, ci, co, L = self.B, self.chan[l], self.chan[l + 1], self.L[l]
        if self.bn_dgrad_fuse:
            self._k("dgrad", lib.ecgb200_conv1d_dgrad_bnstats_bf16, _p(dy), _p(self.wd[l]), _p(self.dp), _p(self.ybuf[l - 1]),
                    _p(self.bnst[l - 1]), _p(self.bwpart[l - 1]), B, co, ci, L, self.L[l - 1], st)
        else:
            self._k("dgrad", lib.ecgb200_conv1d_fwd_bf16, _p(dy), _p(self.wd[l]), None, _p(self.dp), B, co, ci, L, st)

    def _bwd_blocks(self, st, pre, Pp, Gp):
        B = self.B
        n = 0
        main = torch.cuda.current_stream(self.dev)
        for l in (3, 2, 1, 0):
            ci, co, L = self.chan[l], self.chan[l + 1], self.L[l]
            k = f"{pre}backbone.{l}.net."
            self._prof_tag = f"_L{l + 1}"
            if self.bf16:
                self._k("bn_bwd", lib.ecgb200_bn_relu_pool_bwd_bf16, _p(self.ybuf[l]), _p(self.bnst[l]),
                        _p(self.dp) if l < 3 else None, _p(self.dgap) if l == 3 else None, _p(self.dy),
                        Gp(k + "1.weight"), Gp(k + "1.bias"), _p(self.dbpart[l]), _p(self.ws2), B, co, L, 1, st)
                self._k("wgrad", lib.ecgb200_conv1d_wgrad_bf16, _p(self.dy), _p(self.acts[l]), Gp(k + "0.weight"),
                        Gp(k + "0.bias"), _p(self.dbpart[l]), self.ndb[l], _p(self.ws), B, ci, co, L, st)
            else:
                self._k("bn_bwd", lib.ecgb200_bn_relu_pool_bwd_f32, _p(self.ybuf[l]), _p(self.bnst[l]), Pp(k + "1.weight"),
                        _p(self.dp) if l < 3 else None, _p(self.dgap) if l == 3 else None,
                        _p(self.dy), Gp(k + "1.weight"), Gp(k + "1.bias"), _p(self.ws), B, co, L, 1, st)
                self._k("wgrad", lib.ecgb200_conv1d_wgrad_f32, _p(self.dy), _p(self.acts[l]), Gp(k + "0.weight"),
                        Gp(k + "0.bias"), _p(self.ws), B, ci, co, L, st)
            n += 5
            if l == 3 and self.world > 1:
                # bucket A (block 4, the tail of G) is final: all-reduce it while blocks 3..1 run
                ev = torch.cuda.Event()
                ev.record(main)
                self.side.wait_event(ev)
                with torch.cuda.stream(self.side):
                    torch.distributed.all_reduce(self.G[self.bucket_a_off:], group=self.pg)
            if l > 0:
                if self.bf16:
                    self._k("dgrad", lib.ecgb200_conv1d_fwd_bf16, _p(self.dy), _p(self.wd[l]), None, _p(self.dp),
                            B, co, ci, L, st)
                else:
                    self._k("dgrad", lib.ecgb200_conv1d_fwd_f32, _p(self.dy), _p(self.wd[l]), None, _p(self.dp), None,
                            B, co, ci, L, st)
                n += 1
        return n

    def _select(self, slot: int):
        """Make input slot `slot` the one the next _enqueue() / run() reads."""
        self.cur = slot
        self.x, self.y, self.demo = self.xs[slot], self.ys[slot], self.demos[slot]
        if not self.bf16:
            self.acts[0] = self.x                      # fp32 mode convolves the input buffer directly

    def _enqueue(self):
        main = torch.cuda.current_stream(self.dev)
        st = main.cuda_stream
        B = self.B
        n = 0
        pre = "ecg_backbone." if self.mm else ""
        Pp = lambda k: self._seg_ptr(k, self.P)       # noqa: E731
        Gp = lambda k: self._seg_ptr(k, self.G)       # noqa: E731
        blocks = list(self.bb.backbone)
        # ---- forward
        if self.bf16:
            # ECGB200_PDL=fwd: programmatic dependent launch for the forward chain only (no weight-gradient branch
            # competes for SM slots there)
            import os
            fwd_pdl = self.use_graph and os.environ.get("ECGB200_PDL", "0") == "fwd"
            if fwd_pdl:
                old_pdl = lib.ecgb200_set_pdl(1)
            try:
                n += self._fwd_blocks_bf16(st, pre, Pp, blocks)
            finally:
                if fwd_pdl:
                    lib.ecgb200_set_pdl(old_pdl)
        else:
            n += self._fwd_blocks_fp32(st, pre, Pp, blocks)
        self._prof_tag = ""
        c4, F_, NL = self.chan[4], self.feat, self.nl
        if self.bf16 and not self.mm:
            # fused head: forward + BCE + input gradients in one launch; weight gradients + loss on the side
            self._k("head_fwd_bwd", lib.ecgb200_head_fwd_bwd_f32, _p(self.gap), _p(self.wpT), Pp("proj.weight"),
                    Pp("proj.bias"), Pp("head.weight"), Pp("head.bias"), _p(self.y), _p(self.z), _p(self.logits),
                    _p(self.dlogits), _p(self.dz), _p(self.dgap), _p(self.loss_part), B, c4, F_, NL, 1.0, st)
            self._fork_side(main)
            with torch.cuda.stream(self.side):
                V2, I2 = C.c_void_p * 2, C.c_int * 2
                self._k("head_wgrad", lib.ecgb200_head_wgrad_multi_f32, 2, V2(_p(self.dz), _p(self.dlogits)),
                        V2(_p(self.gap), _p(self.z)), V2(Gp("proj.weight"), Gp("head.weight")),
                        V2(Gp("proj.bias"), Gp("head.bias")), I2(F_, NL), I2(c4, F_), _p(self.loss_part), _p(self.loss),
                        B, NL, self.side.cuda_stream)
            n += 2
        elif self.bf16 and self.mm and self._mm_head_ok():
            # fused FiLM head (demo encoder, film, head, BCE and the whole chain rule) + one launch for the five
            # Linear layers' weight gradients on the side branch
            d0, hn = self.demo.shape[1], self.h1.shape[1]
            self._k("mm_head_fwd_bwd", lib.ecgb200_mm_head_fwd_bwd_f32, _p(self.gap), _p(self.demo), _p(self.wpT),
                    Pp(pre + "proj.weight"), Pp(pre + "proj.bias"), Pp("demo_encoder.mlp.0.weight"),
                    Pp("demo_encoder.mlp.0.bias"), Pp("demo_encoder.mlp.2.weight"), Pp("demo_encoder.mlp.2.bias"),
                    Pp("film_gen.weight"), Pp("film_gen.bias"), Pp("head.weight"), Pp("head.bias"), _p(self.y),
                    _p(self.z), _p(self.h1), _p(self.h2), _p(self.film), _p(self.zc), _p(self.logits), _p(self.dlogits),
                    _p(self.dz), _p(self.dfilm), _p(self.dh2), _p(self.dh1), _p(self.dgap), _p(self.loss_part),
                    B, c4, F_, d0, hn, NL, 1.0, st)
            self._fork_side(main)
            with torch.cuda.stream(self.side):
                V5, I5 = C.c_void_p * 5, C.c_int * 5
                self._k("head_wgrad", lib.ecgb200_head_wgrad_multi_f32, 5,
                        V5(_p(self.dz), _p(self.dlogits), _p(self.dfilm), _p(self.dh2), _p(self.dh1)),
                        V5(_p(self.gap), _p(self.zc), _p(self.h2), _p(self.h1), _p(self.demo)),
                        V5(Gp(pre + "proj.weight"), Gp("head.weight"), Gp("film_gen.weight"),
                           Gp("demo_encoder.mlp.2.weight"), Gp("demo_encoder.mlp.0.weight")),
                        V5(Gp(pre + "proj.bias"), Gp("head.bias"), Gp("film_gen.bias"), Gp("demo_encoder.mlp.2.bias"),
                           Gp("demo_encoder.mlp.0.bias")),
                        I5(F_, NL, 2 * F_, hn, hn), I5(c4, F_, hn, hn, d0), _p(self.loss_part), _p(self.loss), B, NL,
                        self.side.cuda_stream)
            n += 2
        else:
            n += self._head_unfused(st, pre, Pp, Gp)
        # ---- backward: conv blocks 4..1
        if self.bf16:
            n += self._bwd_blocks_bf16(pre, Gp)
        else:
            n += self._bwd_blocks(st, pre, Pp, Gp)
        # ---- gradient exchange + optimizer
        self._prof_tag = ""
        if self.world > 1 and not self.dp_fused:
            if self.bf16:
                ev = torch.cuda.Event()
                ev.record(self.side)
                self.comm.wait_event(ev)                        # all weight gradients are final
                with torch.cuda.stream(self.comm):
                    torch.distributed.all_reduce(self.G[:self.bucket_a_off], group=self.pg)
                ev2 = torch.cuda.Event()
                ev2.record(self.comm)
                main.wait_event(ev2)
            else:
                torch.distributed.all_reduce(self.G[:self.bucket_a_off], group=self.pg)
                ev2 = torch.cuda.Event()
                ev2.record(self.side)
                main.wait_event(ev2)
        if self.bf16:
            ev3 = torch.cuda.Event()
            ev3.record(self.side)
            main.wait_event(ev3)                                # join the weight-gradient branch
            if self.dp_fused:
                # reduce-scatter + AdamW on the owned shard + all-gather, one kernel over NVLink peer memory
                W = C.c_void_p * self.world
                self._k("dp_adamw_fused", lib.ecgb200_dp_adamw_fused_f32, W(*self.peer_p), W(*self.peer_g),
                        W(*self.peer_f), self.M.data_ptr(), self.V.data_ptr(), self.total_pad, self.rank, self.world,
                        self.hyper.data_ptr(), self.step_dev.data_ptr(), st)
            else:
                self._k("adamw", lib.ecgb200_adamw_flat_f32, self.P.data_ptr(), self.G.data_ptr(), self.M.data_ptr(),
                        self.V.data_ptr(), self.total_pad, self.hyper.data_ptr(), self.step_dev.data_ptr(), st)
            n += 1
        else:
            one = C.c_void_p * 1
            num = (C.c_int64 * 1)(self.total)
            self._k("adamw", lib.ecgb200_adamw_f32, 1, one(self.P.data_ptr()), one(self.G.data_ptr()), one(self.M.data_ptr()),
                    one(self.V.data_ptr()), num, self.hyper.data_ptr(), self.step_dev.data_ptr(), st)
            n += 2
        self.launches_per_step = n

    def _mm_head_ok(self):
        """Shapes the fused multimodal head kernel covers (the reference's defaults: demo 5 -> 64 -> 64, feat 256)."""
        dm = self.model.demo_encoder.mlp
        return (dm[0].out_features == dm[2].out_features == dm[2].in_features and dm[0].out_features <= 64
                and dm[0].in_features <= 8 and self.feat <= 256 and self.chan[4] <= 256 and self.nl <= 8
                and self.model.film_gen.in_features == dm[2].out_features)

    def _head_unfused(self, st, pre, Pp, Gp):
        """proj / (demo encoder, FiLM) / head / BCE and their backward as separate launches (fp32 engine and
        the multimodal model)."""
        B = self.B
        n = 0
        c4, F_, NL = self.chan[4], self.feat, self.nl
        self._k("proj", lib.ecgb200_linear_fwd_f32, _p(self.gap), Pp(pre + "proj.weight"), Pp(pre + "proj.bias"), _p(self.z),
                                         B, c4, F_, 0, st)
        n += 1
        zin = self.z
        if self.mm:
            d0, h1n, h2n = self.demo.shape[1], self.h1.shape[1], self.h2.shape[1]
            self._k("demo0", lib.ecgb200_linear_fwd_f32, _p(self.demo), Pp("demo_encoder.mlp.0.weight"), Pp("demo_encoder.mlp.0.bias"),
                                             _p(self.h1), B, d0, h1n, 1, st)
            self._k("demo2", lib.ecgb200_linear_fwd_f32, _p(self.h1), Pp("demo_encoder.mlp.2.weight"), Pp("demo_encoder.mlp.2.bias"),
                                             _p(self.h2), B, h1n, h2n, 1, st)
            self._k("film_gen", lib.ecgb200_linear_fwd_f32, _p(self.h2), Pp("film_gen.weight"), Pp("film_gen.bias"), _p(self.film),
                                             B, h2n, 2 * F_, 0, st)
            self._k("film", lib.ecgb200_film_fwd_f32, _p(self.z), _p(self.film), _p(self.zc), B, F_, st)
            zin = self.zc
            n += 4
        self._k("head", lib.ecgb200_linear_fwd_f32, _p(zin), Pp("head.weight"), Pp("head.bias"), _p(self.logits), B, F_, NL, 0, st)
        self._k("bce", lib.ecgb200_bce_logits_f32, _p(self.logits), _p(self.y), _p(self.loss), _p(self.dlogits), None,
                                         B * NL, 1.0, st)
        n += 2
        # ---- backward: head
        if self.mm:
            self._k("head_bwd", lib.ecgb200_linear_bwd_f32, _p(self.zc), Pp("head.weight"), _p(self.dlogits), None, _p(self.dzc),
                                             Gp("head.weight"), Gp("head.bias"), B, F_, NL, st)
            self._k("film_bwd", lib.ecgb200_film_bwd_f32, _p(self.z), _p(self.film), _p(self.dzc), _p(self.dz), _p(self.dfilm), B, F_, st)
            self._k("film_gen_bwd", lib.ecgb200_linear_bwd_f32, _p(self.h2), Pp("film_gen.weight"), _p(self.dfilm), None, _p(self.dh2),
                                             Gp("film_gen.weight"), Gp("film_gen.bias"), B, h2n, 2 * F_, st)
            self._k("demo2_bwd", lib.ecgb200_linear_bwd_f32, _p(self.h1), Pp("demo_encoder.mlp.2.weight"), _p(self.dh2), _p(self.h2), _p(self.dh1),
                                             Gp("demo_encoder.mlp.2.weight"), Gp("demo_encoder.mlp.2.bias"), B, h1n, h2n, st)
            self._k("demo0_bwd", lib.ecgb200_linear_bwd_f32, _p(self.demo), Pp("demo_encoder.mlp.0.weight"), _p(self.dh1), _p(self.h1), None,
                                             Gp("demo_encoder.mlp.0.weight"), Gp("demo_encoder.mlp.0.bias"), B, d0, h1n, st)
            n += 3 + 1 + 3 + 3 + 2
        else:
            self._k("head_bwd", lib.ecgb200_linear_bwd_f32, _p(self.z), Pp("head.weight"), _p(self.dlogits), None, _p(self.dz),
                                             Gp("head.weight"), Gp("head.bias"), B, F_, NL, st)
            n += 3
        self._k("proj_bwd", lib.ecgb200_linear_bwd_f32, _p(self.gap), Pp(pre + "proj.weight"), _p(self.dz), None, _p(self.dgap),
                                         Gp(pre + "proj.weight"), Gp(pre + "proj.bias"), B, c4, F_, st)
        n += 3
        return n

    # ------------------------------------------------------------------ public API
    def capture(self):
        self._refresh_views()
        group = self.opt.param_groups[0]
        self.hyper, self.step_dev = self.opt.device_state(group, self.dev)
        if not self.use_graph:
            return
        # warm-up outside capture would advance the optimizer; capture directly instead
        torch.cuda.synchronize(self.dev)
        import os
        # programmatic dependent launch of the conv / BN-forward chain measured slower in the step (460 vs 449 us:
        # early-scheduled dependents take SM slots from the weight-gradient branch), so it is opt-in
        old = lib.ecgb200_set_pdl(1 if (self.bf16 and os.environ.get("ECGB200_PDL", "0") == "1") else 0)
        # Capturing the critical path on a higher-priority stream than the weight-gradient branch (so that dgrad
        # wins the SMs over wgrad when both become ready) measured slightly SLOWER (460 vs 451 us): the step is
        # bound by the sum of the tensor kernels, which cannot share an SM (TMEM / shared memory), not by their
        # order.  Equal priorities by default.
        prio = os.environ.get("ECGB200_MAIN_PRIORITY", "0")
        self.capture_stream = torch.cuda.Stream(device=self.dev, priority=int(prio)) if self.bf16 else None
        keep = self.cur
        graphs = []
        try:
            for slot in (0, 1):                      # one graph per input slot (same kernels, other input pointers)
                self._select(slot)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=self.capture_stream):
                    self._enqueue()
                graphs.append(g)
        finally:
            lib.ecgb200_set_pdl(old)
            self._select(keep)
        self.graphs = graphs
        self.graph = graphs[0]

    def profile_kernels(self, iters: int = 5):
        """Per-C-ABI-call device time (ms, mean over `iters` un-graphed passes, CUDA events on the
        launch stream).  Advances training like `iters` ordinary steps."""
        self._refresh_views()
        acc = {}
        order = []
        for _ in range(iters):
            self._prof = []
            self._enqueue()
            torch.cuda.synchronize(self.dev)
            for name, e0, e1 in self._prof:
                if name not in acc:
                    acc[name] = 0.0
                    order.append(name)
                acc[name] += e0.elapsed_time(e1)
            self._prof = None
            self.opt.param_groups[0]["step"] = self.opt.param_groups[0].get("step", 0) + 1
        return [(n, acc[n] / iters) for n in order]

    def time_kernels(self, iters: int = 10):
        """Device time of every C-ABI call of the step, free of host launch overhead: each call is captured
        `iters` times back to back in its own CUDA graph, replayed, and timed with CUDA events on the replay
        stream (warm caches: in the real step a kernel's inputs were just produced by its predecessor).
        Returns [(name, ms per launch)].  Advances nothing that matters: buffers are reused, the optimizer
        kernels run on the live state (call it after the measurement you care about)."""
        if self.world > 1:
            raise EcgB200Error("time_kernels() is a single-GPU diagnostic")
        self._refresh_views()
        calls = []
        saved = self._k
        self._k = lambda name, fn, *args: calls.append((name + self._prof_tag, fn, args))
        try:
            self._enqueue()
        finally:
            self._k = saved
        torch.cuda.synchronize(self.dev)
        out = []
        s = torch.cuda.Stream(device=self.dev)
        for name, fn, args in calls:
            args = list(args)
            args[-1] = s.cuda_stream
            with torch.cuda.stream(s):
                check(fn(*args), name)
                torch.cuda.synchronize(self.dev)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=s):
                    for _ in range(iters):
                        check(fn(*args), name)
                g.replay()
                torch.cuda.synchronize(self.dev)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(s)
                g.replay()
                g.replay()
                e1.record(s)
                torch.cuda.synchronize(self.dev)
            out.append((name, e0.elapsed_time(e1) / (2 * iters)))
        return out

    def load_batch(self, x, y, demo=None, slot=None):
        """Copy a batch (pinned host or device tensors) into an input slot (async on the current stream).
        slot=None: the idle slot, which then becomes the one run() uses; an explicit slot (0/1) is only filled
        (pipelined use: fill slot s on a copy stream while the graph of slot 1-s runs, then run(slot=s))."""
        if tuple(x.shape) != tuple(self.x.shape) or tuple(y.shape) != tuple(self.y.shape):
            raise EcgB200Error(f"TrainStep was built for x{tuple(self.x.shape)} y{tuple(self.y.shape)}, "
                               f"got x{tuple(x.shape)} y{tuple(y.shape)}")
        s = (self.cur ^ 1) if slot is None else int(slot)
        self.xs[s].copy_(x, non_blocking=True)
        self.ys[s].copy_(y, non_blocking=True)
        if self.mm:
            if demo is None:
                raise EcgB200Error("ECGMultimodal step needs x_demo")
            self.demos[s].copy_(demo, non_blocking=True)
        if slot is None:
            self._select(s)

    def run(self, slot=None):
        """One optimizer step on whatever input slot `slot` (default: the current one) holds.  Returns the loss
        buffer (device scalar, overwritten by the next step)."""
        if slot is not None and int(slot) != self.cur:
            self._select(int(slot))
        group = self.opt.param_groups[0]
        hyper, _ = self.opt.device_state(group, self.dev)
        if hyper.data_ptr() != self.hyper.data_ptr():          # lr / betas changed on the host
            self.hyper.copy_(hyper)
            group["_hyper"] = self.hyper
        if self.use_graph:
            if self.graph is None:
                self.capture()
            self.graphs[self.cur].replay()
        else:
            self._refresh_views()
            self._enqueue()
        group["step"] = group.get("step", 0) + 1
        return self.loss

    def __call__(self, x, y, demo=None):
        self.load_batch(x, y, demo)
        return self.run()
