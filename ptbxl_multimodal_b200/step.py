"""TrainStep: the whole training step of the reference loop body
(zero_grad -> model(x) -> BCE -> backward -> AdamW.step, src/training/loop.py:26-36 and
src/training/loop_demo.py:30-41) as ONE CUDA graph of ecgb200 kernels over static buffers.

B200-first structure: parameters, gradients and both Adam moments live in four flat fp32
buffers (the nn.Module's parameters are re-pointed to views of the flat parameter buffer, so
``state_dict`` / checkpoints are untouched); activations live in preallocated buffers sized for
(batch, seq_len); the kernel sequence is enqueued once through the C ABI, captured with
``torch.cuda.CUDAGraph`` and replayed per step, so a step costs one launch from the host and
no host<->device synchronisation (the loss stays on the device until the caller reads it).

Data parallel (one process per GPU): the gradient mean and the optimizer are ONE exchange, in two buckets --
the 4th conv block (65 % of the bytes, final as soon as its weight gradient is) while blocks 3..1 still run
backward, the rest at the end.  bf16 engine: reduce-scatter + AdamW on the owned shard + all-gather as one
kernel per bucket over NVLink peer memory (csrc/dp_fused.cu, ``dp_mode="fused"``); ``dp_mode="nccl"`` keeps NCCL
all-reduce + a replicated AdamW (the 1/world_size average folded into the AdamW kernel) for A/B runs and for the
fp32 engine.  BatchNorm statistics are per rank (torch DDP semantics) unless ``sync_bn=True``: then every
BatchNorm pass exchanges its partial sums over peer memory and all ranks normalise with the statistics of the
GLOBAL batch -- the single-device reference's result on the concatenated batch."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import torch

from ._lib import lib, check, EcgB200Error, configure_timeouts
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
                 process_group=None, use_graph: bool = True, precision: str = "fp32", dp_mode: str = "auto",
                 raw_input: bool = False, sync_bn: bool = False, pdl: bool = False, input_slots: int = 2):
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
        if dp_mode not in ("auto", "fused", "barrier", "nccl"):
            raise EcgB200Error("dp_mode must be 'auto', 'fused' (peer-memory reduce-scatter + AdamW + all-gather kernel, "
                               "one-hop {value, epoch} words), 'barrier' (the same with flag barriers) or 'nccl' (all-reduce, "
                               "then a replicated AdamW)")
        # gradient exchange: the fused NVLink kernels need the flat-buffer optimizer of the bf16 engine
        self.dp_fused = self.world > 1 and (dp_mode in ("fused", "barrier") or (dp_mode == "auto" and self.bf16))
        self.dp_ll = self.dp_fused and dp_mode != "barrier"        # one-hop words instead of barriers
        if self.dp_fused and not self.bf16:
            raise EcgB200Error("dp_mode='fused' is implemented for precision='bf16'")
        # raw_input: the step starts from raw WFDB format-16 frames (B, T, leads) int16 (load_frames) -- decode, per-lead
        # z-score and the bf16 pack are one kernel at the head of the graph (SURVEY 8f N1 + N2)
        self.raw_input = bool(raw_input)
        if self.raw_input and not self.bf16:
            raise EcgB200Error("raw_input=True is implemented for precision='bf16'")
        # pdl: programmatic dependent launch along the main-stream chain (conv -> BatchNorm -> conv ..., head, BatchNorm
        # backward -> dgrad): a kernel's launch latency and prologue overlap the tail of the one before
        self.pdl = bool(pdl) and self.bf16
        self.sync_bn = bool(sync_bn) and self.world > 1
        if self.sync_bn and not self.dp_fused:
            raise EcgB200Error("sync_bn=True needs the peer-memory exchange (precision='bf16', dp_mode 'auto' or 'fused')")
        # input slots: batch i+1 .. i+slots-1 can be copied in (H2D) while the graph of slot i runs; one captured graph per
        # slot.  Two are enough on one GPU; under data parallel every step ends in a cross-rank exchange, so one late
        # host copy on ANY rank stalls all of them -- a deeper prefetch queue absorbs that jitter.
        self.nslots = int(input_slots)
        if self.nslots < 2 or self.nslots > 8:
            raise EcgB200Error("input_slots must be in 2..8")
        self.use_graph = use_graph
        self.graph = None
        self.launches_per_step = 0
        self._prof = None
        self._prof_tag = ""
        self._stamps = None
        configure_timeouts()
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
        # Two buckets, each padded so that it splits into 16-byte aligned shards for any world size <= 8:
        # B = [0, bucket_a_off) everything but conv block 4, A = [bucket_a_off, total_pad) block 4 (the tail)
        n_rest = sum(p.numel() for n, p in order if not n.startswith(last))
        self.bucket_a_off = padded_size(n_rest)
        self.total_pad = self.bucket_a_off + padded_size(total - n_rest)
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
                if n.startswith(last) and off < self.bucket_a_off:
                    off = self.bucket_a_off                  # block 4 starts on its own bucket
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
        self._probe = next(iter(self.seg.values()))
        if self.world > 1:
            # replicas must start identical (DDP broadcasts too); cheap, once
            torch.distributed.broadcast(self.P, src=torch.distributed.get_global_rank(self.pg, 0) if self.pg is not None else 0,
                                        group=self.pg)
        assert self.seg[last + "net.0.weight"].off == self.bucket_a_off
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
        self.flag_stride = max(nflag, 64)                    # words between flag pads (256-byte aligned)
        self.P = symm.empty(n, dtype=F32, device=self.dev)
        self.G = symm.empty(n, dtype=F32, device=self.dev)
        # three flag pads (bucket A, bucket B, BatchNorm exchanges) + 8 BatchNorm exchange slots of 2 * 256 floats
        self.flags = symm.empty(3 * self.flag_stride, dtype=torch.int32, device=self.dev)
        self.bnx = symm.empty(8 * 512, dtype=F32, device=self.dev)
        # one-hop exchange: per bucket an inbox of {value, epoch} words (gradients for the owned shard from every rank +
        # the new parameters of the whole bucket, both double-buffered by epoch parity), per BatchNorm exchange slot
        # [2][world][512] words; epoch counters are plain local memory
        self.ll_words = [int(lib.ecgb200_dp_ll_inbox_words(cnt)) for _, cnt, _ in self._buckets()]
        self.bn_ll_words = 2 * self.world * 512
        self.inbox = symm.empty(sum(self.ll_words) + 8 * self.bn_ll_words, dtype=torch.int64, device=self.dev)
        self.ll_ctr = torch.zeros(2 * 2 + 8, dtype=torch.int32, device=self.dev)
        self.P.zero_(); self.G.zero_(); self.flags.zero_(); self.bnx.zero_(); self.inbox.zero_()
        torch.cuda.synchronize(self.dev)
        hp, hg, hf, hb, hi = (symm.rendezvous(t, group) for t in (self.P, self.G, self.flags, self.bnx, self.inbox))
        self._symm_handles = (hp, hg, hf, hb, hi)            # keep the mappings alive
        ptrs = lambda h: [int(h.buffer_ptrs[r]) for r in range(self.world)]      # noqa: E731
        self.peer_p, self.peer_g, self.peer_f, self.peer_bnx, self.peer_inbox = ptrs(hp), ptrs(hg), ptrs(hf), ptrs(hb), ptrs(hi)
        if self.peer_p[self.rank] != self.P.data_ptr() or self.peer_g[self.rank] != self.G.data_ptr():
            raise EcgB200Error("symmetric-memory rendezvous returned unexpected local pointers")
        dist.barrier(group)

    def _flag_pad(self, which: int):
        """Peer pointers of flag pad `which` (0: bucket A, 1: bucket B, 2: BatchNorm exchanges)."""
        return [p + 4 * which * self.flag_stride for p in self.peer_f]

    def _buckets(self):
        """(offset, length, flag pad) of the exchange buckets, in launch order."""
        return [(self.bucket_a_off, self.total_pad - self.bucket_a_off, 0), (0, self.bucket_a_off, 1)]

    def gather_optimizer_state(self):
        """dp_mode='fused' shards the Adam moments (each rank updates 1/world of them).  Before saving an
        optimizer checkpoint, call this on every rank: all-gathers the shards so that opt.state is complete."""
        if not self.dp_fused:
            return
        for off, n, _ in self._buckets():
            lo, hi = shard_bounds(n, self.world)[self.rank]
            for buf in (self.M, self.V):
                shard = buf[off + lo:off + hi].clone()
                torch.distributed.all_gather_into_tensor(buf[off:off + n], shard, group=self.pg)

    def _refresh_views(self):
        """Re-point parameters at the flat buffers if something (e.g. load_state_dict keeps
        them, .to() does not) replaced their storage."""
        with torch.no_grad():
            for s in self.seg.values():
                if s.param.data_ptr() != self.P.data_ptr() + 4 * s.off:
                    self.P[s.off:s.off + s.n].copy_(s.param.detach().reshape(-1))
                    s.param.data = self.P[s.off:s.off + s.n].view(s.param.shape)
                s.param.grad = self.G[s.off:s.off + s.n].view(s.param.shape)
                # optimizer.load_state_dict() replaces the moment tensors: adopt the new values and re-alias, so that
                # the captured graphs (which update M / V) and opt.state_dict() stay one and the same
                st = self.opt.state[s.param]
                for key, flat in (("exp_avg", self.M), ("exp_avg_sq", self.V)):
                    t = st.get(key)
                    if t is None or t.data_ptr() != flat.data_ptr() + 4 * s.off:
                        if t is not None:
                            flat[s.off:s.off + s.n].copy_(t.detach().reshape(-1).to(flat.device, F32))
                        st[key] = flat[s.off:s.off + s.n].view(s.param.shape)
        group = self.opt.param_groups[0]
        sd = group.get("_step_dev")
        if sd is None or sd.data_ptr() != self.step_dev.data_ptr():
            # the step counter the graphs increment must be the optimizer's: re-adopt the host-side step count
            self.step_dev.fill_(int(group.get("step", 0)))
            group["_step_dev"] = self.step_dev

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
        self.xs = [e(B, self.chan[0], T) for _ in range(self.nslots)]
        self.ys = [e(B, self.nl) for _ in range(self.nslots)]
        self.demos = [None] * self.nslots
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
            # BN backward = reduce + apply launches (the apply pass is launched programmatically dependent)
            self.ndb = [lib.ecgb200_bn_nsplit(B, self.chan[l + 1]) for l in range(4)]
            self.dbpart = [e(self.chan[l + 1], self.ndb[l]) for l in range(4)]
            self.stat = [None] * 4
            # per-CTA {sum, sumsq} partials written by the conv epilogue
            self.nstat = [lib.ecgb200_conv1d_stat_parts_bf16(B, self.cip[l], self.chan[l + 1], self.L[l]) for l in range(4)]
            if min(self.nstat) <= 0:
                raise EcgB200Error("bf16 conv kernel does not support this shape")
            self.statp = [e(self.nstat[l], 2, self.chan[l + 1]) for l in range(4)]
            self.wpT = e(self.chan[4], self.bb.proj.out_features)
            self.loss_part = e(lib.ecgb200_head_loss_parts(B))
            if self.raw_input:
                # raw WFDB format-16 frames per input slot + per-lead calibration (.hea gain / baseline)
                nlead = self.chan[0]
                self.frames = [torch.zeros(B, T, nlead, dtype=torch.int16, device=dev) for _ in range(self.nslots)]
                self.gain = torch.full((nlead,), 200.0, dtype=F32, device=dev)
                self.baseline = torch.zeros(nlead, dtype=torch.int32, device=dev)
            if self.sync_bn:
                # exchanged statistics: one {sum, sum of squares} / {sum g, sum g*a} pair per replica
                self.bnsync = [e(self.world, 2, self.chan[l + 1]) for l in range(4)]
        c4 = self.chan[4]
        self.gap = e(B, c4)
        self.dgap = e(B, c4)
        self.route = e(2, B, c4)                # block 4's pool routing summary (BatchNorm backward in one pass)
        self.z = e(B, self.feat)
        self.dz = e(B, self.feat)
        self.logits = e(B, self.nl)
        self.dlogits = e(B, self.nl)
        # one loss scalar per input slot: the step of slot s leaves its loss in losses[s], where it stays valid until slot s
        # is run again -- a caller can read it back on ANOTHER stream while the next steps run (bench.py's e2e leg does)
        self.losses = [torch.zeros((), dtype=F32, device=dev) for _ in range(self.nslots)]
        self.loss = self.losses[0]
        if self.mm:
            dm = self.model.demo_encoder.mlp
            self.demos = [e(B, dm[0].in_features) for _ in range(self.nslots)]
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
        self.split_adamw = True                 # one GPU: optimizer for everything but block 1 beside wgrad_1 (A/B switch)
        self.comm = torch.cuda.Stream(device=dev) if (self.world > 1 and self.bf16) else None

    # ------------------------------------------------------------------ the kernel sequence
    def _k(self, name, fn, *args):
        """One C-ABI call; with self._prof set, bracket it with CUDA events on the launch stream."""
        prof = self._prof
        if self._stamps is not None:
            # schedule trace: %globaltimer stamps in stream order before / after the call (args[-1] is its stream)
            buf, names = self._stamps
            i = len(names)
            names.append((name + self._prof_tag, args[-1]))
            check(lib.ecgb200_debug_stamp(buf.data_ptr(), 2 * i, args[-1]), "stamp")
            check(fn(*args), name)
            check(lib.ecgb200_debug_stamp(buf.data_ptr(), 2 * i + 1, args[-1]), "stamp")
            return
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
        if self.raw_input:
            # raw frames: the decode + z-score + pack kernel is the critical path; ALL weight re-layouts (+ proj transpose +
            # step counter) are one launch beside it on the side stream, joined before conv 1
            ev0 = torch.cuda.Event()
            ev0.record(main)
            # the critical path's node is created FIRST: ready graph nodes are launched in creation order, and with the side
            # branch's root created first the whole weight-gradient branch won every later race for the SMs (multimodal
            # B=1024: head_wgrad starved behind dgrad_4, step 1.20 -> 1.27 ms)
            self._k("decode", lib.ecgb200_wfdb16_zscore_pack_bf16, _p(self.frames[self.cur]), _p(self.gain),
                    _p(self.baseline), _p(self.acts[0]), B, self.chan[0], self.T, st)
            if self.linear:
                self.side = main
            else:
                self.side.wait_event(ev0)
            with torch.cuda.stream(self.side):
                self._k("prep_w", lib.ecgb200_step_prep_bf16, None, None, 0, 0, 0, 4,
                        PV(*[Pp(k) for k in wkeys]), PV(*[_p(w) for w in self.wt]),
                        PV(*[_p(w) for w in self.wd]), I4(*self.chan[1:5]), I4(*self.chan[0:4]),
                        Pp(pre + "proj.weight"), _p(self.wpT),
                        self.feat, self.chan[4], self.step_dev.data_ptr(), self.side.cuda_stream)
                prep_done = torch.cuda.Event()
                prep_done.record(self.side)
            main.wait_event(prep_done)
        else:
            # critical path: pack the input + block-1 weights (+ step counter); blocks 2-4 and the proj transpose
            # are re-laid beside the first conv on the side stream
            self._k("prep", lib.ecgb200_step_prep_bf16, _p(self.x), _p(self.acts[0]), B, self.chan[0],
                    self.T, 1, PV(Pp(wkeys[0]), None, None, None), PV(_p(self.wt[0]), None, None, None), PV(None, None, None, None),
                    I4(self.chan[1], 0, 0, 0), I4(self.chan[0], 0, 0, 0), None, None, 0, 0, self.step_dev.data_ptr(), st)
            ev_prep = torch.cuda.Event()
            ev_prep.record(main)
            # released by `prep`; its node is created before conv 1's (created after it measured slower: conv 2 then waits)
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
                prep_done = torch.cuda.Event()
                prep_done.record(self.side)
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
            stat, nparts, nrep = self.statp[l], self.nstat[l], 1
            if self.sync_bn:
                stat, nparts, nrep = self._bn_exchange(l, 0, self.statp[l], self.nstat[l], st), self.world, self.world
                n += 1
            if l == 3 and not self.sync_bn:
                # last block: also the routing summary that lets its BatchNorm backward skip the reduce pass over y
                self._k("bn_relu_pool", lib.ecgb200_bn_relu_pool_fwd_train_route_bf16, _p(self.ybuf[l]), _p(stat),
                        nparts, Pp(k + "1.weight"), Pp(k + "1.bias"), bn.running_mean.data_ptr(),
                        bn.running_var.data_ptr(), bn.num_batches_tracked.data_ptr(), _p(self.bnst[l]), None,
                        _p(self.gap), _p(self.route), B, co, L, float(bn.momentum), float(bn.eps), nrep, st)
            else:
                self._k("bn_relu_pool", lib.ecgb200_bn_relu_pool_fwd_train_bf16, _p(self.ybuf[l]), _p(stat),
                        nparts, Pp(k + "1.weight"), Pp(k + "1.bias"), bn.running_mean.data_ptr(),
                        bn.running_var.data_ptr(), bn.num_batches_tracked.data_ptr(), _p(self.bnst[l]),
                        _p(self.acts[l + 1]) if l < 3 else None, _p(self.gap) if l == 3 else None, B, co, L,
                        float(bn.momentum), float(bn.eps), nrep, st)
            n += 2
        return n

    def _bn_exchange(self, l, direction, part, nparts, st):
        """SyncBN: all replicas' partial pairs of block l (direction 0 forward statistics, 1 backward sums) gathered
        into bnsync[l] (world, 2, C) over peer memory."""
        W = C.c_void_p * self.world
        k = 2 * l + direction
        if self.dp_ll:
            off = 8 * (sum(self.ll_words) + k * self.bn_ll_words)        # byte offset of this exchange's inbox
            self._k("bn_sync", lib.ecgb200_dp_bn_sync_ll_f32, _p(part), nparts, self.chan[l + 1],
                    W(*[p + off for p in self.peer_inbox]), self.ll_ctr.data_ptr() + 4 * (4 + k), _p(self.bnsync[l]),
                    self.rank, self.world, st)
            return self.bnsync[l]
        slot = 512 * 4 * k                                      # byte offset of this exchange's slot
        self._k("bn_sync", lib.ecgb200_dp_bn_sync_f32, _p(part), nparts, self.chan[l + 1],
                W(*[p + slot for p in self.peer_bnx]), W(*self._flag_pad(2)), _p(self.bnsync[l]), self.rank, self.world, st)
        return self.bnsync[l]

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
            dpb, dgap = (_p(self.dp), None) if l < 3 else (None, _p(self.dgap))
            if self.sync_bn:
                # reduce -> exchange over peer memory -> apply with the sums of the GLOBAL batch
                self._k("bn_bwd", lib.ecgb200_bn_relu_pool_bwd_reduce_bf16, _p(self.ybuf[l]), _p(self.bnst[l]), dpb, dgap,
                        _p(self.ws2), B, co, L, st)
                merged = self._bn_exchange(l, 1, self.ws2, self.ndb[l], st)
                self._k("bn_bwd_apply", lib.ecgb200_bn_relu_pool_bwd_apply_bf16, _p(self.ybuf[l]), _p(self.bnst[l]), dpb, dgap,
                        _p(merged), self.world, self.rank, self.world, _p(dy), Gp(k + "1.weight"), Gp(k + "1.bias"),
                        _p(self.dbpart[l]), B, co, L, 1, st)
                n += 3
            elif l == 3:
                # one pass: the batch reductions come from the forward pass's routing summary (GAP gradient)
                self._k("bn_bwd", lib.ecgb200_bn_relu_pool_bwd_route_bf16, _p(self.ybuf[l]), _p(self.bnst[l]), dgap,
                        _p(self.route), _p(dy), Gp(k + "1.weight"), Gp(k + "1.bias"), _p(self.dbpart[l]), B, co, L, 1, st)
                n += 1
            else:
                self._k("bn_bwd", lib.ecgb200_bn_relu_pool_bwd_bf16, _p(self.ybuf[l]), _p(self.bnst[l]), dpb, dgap, _p(dy),
                        Gp(k + "1.weight"), Gp(k + "1.bias"), _p(self.dbpart[l]), _p(self.ws2), B, co, L, 1, st)
                n += 2
            # Order matters: the tensor kernels of the two branches share the SMs one CTA at a time (TMEM + shared
            # memory), and ready graph nodes are launched in creation order.  Both dgrad_l (critical path) and wgrad_l
            # (side) are released by bn_bwd_l, with dgrad's node created first (measured, us/step: wgrad first 476,
            # wgrad released only after dgrad completes 466, this order 453).
            ev_bn = torch.cuda.Event()
            ev_bn.record(main)
            if l > 0:
                self._k("dgrad", lib.ecgb200_conv1d_fwd_bf16, _p(dy), _p(self.wd[l]), None, _p(self.dp), B, co, ci, L, st)
                n += 1
            if l == 2 and self.world > 1:
                # Bucket A (block 4 = the tail of the flat buffers; bn_bwd_4 left its BatchNorm gradients there, wgrad_4 the
                # rest) goes out on the communication stream under the backward of blocks 2..1.  It is released only once
                # dgrad_3 has COMPLETED: released together with dgrad_3 / wgrad_3 (right after wgrad_4) the extra ready node
                # flipped the launch order of those two, wgrad_3's CTAs took the SMs first and dgrad_3 -- the critical path --
                # ran 24 us late (traced schedule, 2 GPUs).
                ev_d3 = torch.cuda.Event()
                ev_d3.record(main)
                self.comm.wait_event(ev_wg4)
                self.comm.wait_event(ev_d3)
                with torch.cuda.stream(self.comm):
                    if self.dp_fused:
                        self._dp_exchange(0, self.comm.cuda_stream)
                        self._prof_tag = f"_L{l + 1}"
                        n += 1
                    else:
                        torch.distributed.all_reduce(self.G[self.bucket_a_off:], group=self.pg)
            if self.linear:
                self.side = main
            else:
                self.side.wait_event(ev_bn)
            with torch.cuda.stream(self.side):
                self._k("wgrad", lib.ecgb200_conv1d_wgrad_bf16, _p(dy), _p(self.acts[l]), Gp(k + "0.weight"),
                        Gp(k + "0.bias"), _p(self.dbpart[l]), self.ndb[l], _p(self.ws), B, ci, co, L,
                        self.side.cuda_stream)
                wg_done[l & 1] = torch.cuda.Event()
                wg_done[l & 1].record(self.side)
            n += 2
            if l == 3:
                ev_wg4 = wg_done[l & 1]
            if l == 0 and self._adamw_split():
                # One GPU: every gradient but conv 1's is final once wgrad_2 and bn_bwd_1 are.  The optimizer takes
                # [n1, total) on the main stream now, beside wgrad_1 -- a thin tensor kernel that leaves the memory system
                # idle -- and only block 1's 5.9 k parameters after it (beside bn_bwd_1 instead, the 20 MB AdamW pass slowed
                # that kernel on the critical path: 0.403 -> 0.412 ms).
                n1 = self._block1_span()
                main.wait_event(wg_done[1])                    # wgrad_2 (and everything before it on that stream)
                self._prof_tag = ""
                self._k("adamw_rest", lib.ecgb200_adamw_flat_f32, self.P.data_ptr() + 4 * n1, self.G.data_ptr() + 4 * n1,
                        self.M.data_ptr() + 4 * n1, self.V.data_ptr() + 4 * n1, self.total_pad - n1, self.hyper.data_ptr(),
                        self.step_dev.data_ptr(), st)
                n += 1
        return n

    def _block1_span(self):
        """Length of the flat-buffer prefix that holds exactly block 1's parameters (0 if they are not a 16-byte aligned prefix)."""
        pre = ("ecg_backbone." if self.mm else "") + "backbone.0."
        segs = sorted(self.seg.values(), key=lambda s: s.off)
        n1 = 0
        for s in segs:
            if s.name.startswith(pre):
                if s.off != n1:
                    return 0
                n1 = s.off + s.n
        return n1 if n1 % 4 == 0 and all(s.off >= n1 or s.name.startswith(pre) for s in segs) else 0

    def _adamw_split(self):
        return (self.split_adamw and self.bf16 and self.world == 1 and not self.linear and not self.sync_bn
                and self._block1_span() > 0)

    def _dp_exchange(self, which, st):
        """Bucket `which` (0 = block 4, 1 = the rest): reduce-scatter + AdamW on the owned shard + all-gather, one
        kernel over NVLink peer memory."""
        off, cnt, pad = self._buckets()[which]
        W = C.c_void_p * self.world
        self._prof_tag = "_A" if which == 0 else "_B"
        if self.dp_ll:
            boff = 8 * sum(self.ll_words[:which])
            self._k("dp_adamw_fused", lib.ecgb200_dp_adamw_ll_f32, self.P.data_ptr(), self.G.data_ptr(), self.M.data_ptr(),
                    self.V.data_ptr(), W(*[p + boff for p in self.peer_inbox]), self.ll_ctr.data_ptr() + 8 * which, off, cnt,
                    self.rank, self.world, self.hyper.data_ptr(), self.step_dev.data_ptr(), st)
        else:
            self._k("dp_adamw_fused", lib.ecgb200_dp_adamw_fused_range_f32, W(*self.peer_p), W(*self.peer_g),
                    W(*self._flag_pad(pad)), self.M.data_ptr(), self.V.data_ptr(), off, cnt, self.rank, self.world,
                    self.hyper.data_ptr(), self.step_dev.data_ptr(), st)
        self._prof_tag = ""

    def _bwd_blocks(self, st, pre, Pp, Gp):
        """fp32 engine: BN/ReLU/pool backward -> wgrad -> dgrad, block 4 down to 1, on one stream."""
        B = self.B
        n = 0
        main = torch.cuda.current_stream(self.dev)
        for l in (3, 2, 1, 0):
            ci, co, L = self.chan[l], self.chan[l + 1], self.L[l]
            k = f"{pre}backbone.{l}.net."
            self._prof_tag = f"_L{l + 1}"
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
                self._k("dgrad", lib.ecgb200_conv1d_fwd_f32, _p(self.dy), _p(self.wd[l]), None, _p(self.dp), None,
                        B, co, ci, L, st)
                n += 1
        return n

    def _select(self, slot: int):
        """Make input slot `slot` the one the next _enqueue() / run() reads."""
        self.cur = slot
        self.x, self.y, self.demo = self.xs[slot], self.ys[slot], self.demos[slot]
        self.loss = self.losses[slot]
        if not self.bf16:
            self.acts[0] = self.x                      # fp32 mode convolves the input buffer directly

    def _enqueue(self):
        old = lib.ecgb200_set_pdl(1 if self.pdl else 0)
        try:
            self._enqueue_step()
        finally:
            lib.ecgb200_set_pdl(old)

    def _enqueue_step(self):
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
            n += self._fwd_blocks_bf16(st, pre, Pp, blocks)
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
                # bucket B (everything but block 4); bucket A went out under the backward of blocks 3..1
                self._dp_exchange(1, st)
                ev4 = torch.cuda.Event()
                ev4.record(self.comm)
                main.wait_event(ev4)                            # join the bucket-A exchange
            elif self._adamw_split():
                self._k("adamw", lib.ecgb200_adamw_flat_f32, self.P.data_ptr(), self.G.data_ptr(), self.M.data_ptr(),
                        self.V.data_ptr(), self._block1_span(), self.hyper.data_ptr(), self.step_dev.data_ptr(), st)
            else:
                self._k("adamw", lib.ecgb200_adamw_flat_f32, self.P.data_ptr(), self.G.data_ptr(), self.M.data_ptr(),
                        self.V.data_ptr(), self.total_pad, self.hyper.data_ptr(), self.step_dev.data_ptr(), st)
            n += 1
        else:
            one = C.c_void_p * 1
            num = (C.c_int64 * 1)(self.total_pad)        # pads hold p = g = 0 and stay 0
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
        self.capture_stream = torch.cuda.Stream(device=self.dev) if self.bf16 else None
        keep = self.cur
        graphs = []
        try:
            for slot in range(self.nslots):          # one graph per input slot (same kernels, other input pointers)
                self._select(slot)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=self.capture_stream):
                    self._enqueue()
                graphs.append(g)
        finally:
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

    def trace_schedule(self, replays: int = 5):
        """The schedule the captured two-stream step REALLY runs (no nsys on the box): the step is captured once more with
        a one-thread %globaltimer stamp kernel before and after every C-ABI call on that call's stream, replayed, and the
        last replay's stamps are returned as [(name, stream id, start us, end us)] relative to the first stamp.  Every stamp
        costs ~1.5 us of its stream, so the traced step is slower than the real one: read overlaps and gaps, not totals.
        Advances training by `replays` steps."""
        self._refresh_views()
        buf = torch.zeros(512, dtype=torch.int64, device=self.dev)
        names = []
        self._stamps = (buf, names)
        try:
            torch.cuda.synchronize(self.dev)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=torch.cuda.Stream(device=self.dev)):
                self._enqueue()
        finally:
            self._stamps = None
        for _ in range(replays):
            g.replay()
        torch.cuda.synchronize(self.dev)
        self.opt.param_groups[0]["step"] = self.opt.param_groups[0].get("step", 0) + replays
        t = buf.cpu().tolist()
        t0 = min(v for v in t[:2 * len(names)] if v)
        streams = {}
        out = []
        for i, (n, st) in enumerate(names):
            sid = streams.setdefault(st, len(streams))
            out.append((n, sid, (t[2 * i] - t0) / 1000.0, (t[2 * i + 1] - t0) / 1000.0))
        return out

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
        slot=None: the next slot, which then becomes the one run() uses; an explicit slot (0 .. input_slots-1) is only
        filled (pipelined use: fill slot s on a copy stream while the graphs of the other slots run, then run(slot=s))."""
        if tuple(x.shape) != tuple(self.x.shape) or tuple(y.shape) != tuple(self.y.shape):
            raise EcgB200Error(f"TrainStep was built for x{tuple(self.x.shape)} y{tuple(self.y.shape)}, "
                               f"got x{tuple(x.shape)} y{tuple(y.shape)}")
        s = ((self.cur + 1) % self.nslots) if slot is None else int(slot)
        self.xs[s].copy_(x, non_blocking=True)
        self.ys[s].copy_(y, non_blocking=True)
        if self.mm:
            if demo is None:
                raise EcgB200Error("ECGMultimodal step needs x_demo")
            self.demos[s].copy_(demo, non_blocking=True)
        if slot is None:
            self._select(s)

    def load_frames(self, frames, y, demo=None, slot=None):
        """raw_input engines: copy a batch of raw WFDB format-16 frames (B, T, leads) int16 (pinned host or device) into an
        input slot; the graph decodes, z-scores and packs them on the device (set_calibration for gain / baseline)."""
        if not self.raw_input:
            raise EcgB200Error("load_frames() needs TrainStep(..., raw_input=True)")
        s = ((self.cur + 1) % self.nslots) if slot is None else int(slot)
        if tuple(frames.shape) != tuple(self.frames[s].shape) or frames.dtype != torch.int16:
            raise EcgB200Error(f"expected int16 frames {tuple(self.frames[s].shape)}, got {frames.dtype} {tuple(frames.shape)}")
        self.frames[s].copy_(frames, non_blocking=True)
        self.ys[s].copy_(y, non_blocking=True)
        if self.mm:
            if demo is None:
                raise EcgB200Error("ECGMultimodal step needs x_demo")
            self.demos[s].copy_(demo, non_blocking=True)
        if slot is None:
            self._select(s)

    def set_calibration(self, gain, baseline):
        """Per-lead ADC gain (units per mV) and baseline of the .hea header for raw_input engines."""
        self.gain.copy_(torch.as_tensor(gain, dtype=F32).reshape(-1))
        self.baseline.copy_(torch.as_tensor(baseline, dtype=torch.int32).reshape(-1))

    def close(self):
        """Drop the captured graphs and the NVLink peer mappings (call before destroy_process_group())."""
        self.graph = None
        self.graphs = None
        self._closed = True
        torch.cuda.synchronize(self.dev)
        for name in ("_symm_handles", "peer_p", "peer_g", "peer_f", "peer_bnx", "peer_inbox"):
            if hasattr(self, name):
                delattr(self, name)

    def run(self, slot=None):
        """One optimizer step on whatever input slot `slot` (default: the current one) holds.  Returns that slot's loss
        buffer (device scalar, overwritten the next time the same slot is run)."""
        if getattr(self, "_closed", False):
            raise EcgB200Error("TrainStep.close() was called: the graphs and peer mappings are gone, build a new engine")
        if slot is not None and int(slot) != self.cur:
            self._select(int(slot))
        group = self.opt.param_groups[0]
        if group.get("_step_dev") is not self.step_dev or \
                self.opt.state[self._probe.param].get("exp_avg", self.M).data_ptr() != self.M.data_ptr() + 4 * self._probe.off:
            self._refresh_views()                               # optimizer.load_state_dict() after the engine was built
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
