#!/usr/bin/env python
"""bench.py -- ECG samples/s of the training step (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step = one pass of the hot path over one synthetic batch: forward (train-mode BN), BCE,
backward, AdamW (+ gradient all-reduce when N > 1).  Workload = BASELINE.json configs[1]:
ECGCNN(12, 256, 5), random init seed 42, batch 256 x 12 x 1000 per GPU, AdamW(1.5e-3, 1e-4).
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ECG samples/sec train step (12x1000)"
UNIT = "samples/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="per-GPU batch (weak scaling)")
    ap.add_argument("--seq-len", type=int, default=1000)
    ap.add_argument("--precision", default="auto", choices=["auto", "fp32", "bf16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "tf_burst": p["bf16_tflops"], "tf_sustained": p["bf16_tflops_sustained"],
                "src": "measured"}
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "src": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [c.strip() for c in r.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------- CPU reference arm
def cpu_train_steps(batch: int, seq_len: int, steps: int, warmup: int, budget_s: float):
    """Reference CPU implementation of the path (oracle port of the reference modules: same
    ATen ops, fp32) on all host cores.  Returns (samples/s, per-step batch used, steps timed)."""
    import torch
    from oracle import ecg_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    sd = O.init_state_dict("cnn", 5, seed=42)
    st = O.AdamWState(sd, 1.5e-3, 1e-4)
    b = batch
    x, y = O.synth_batch(b, seq_len, 5, seed=0)
    O.train_step(sd, x, y, st)                     # cold (thread pool, allocator): not representative
    t0 = time.perf_counter()
    O.train_step(sd, x, y, st)
    one = time.perf_counter() - t0
    # bound the per-step sample so that warmup+steps fits the budget
    while b > 8 and one * (b / batch) * (steps + warmup) > budget_s:
        b //= 2
    if b != batch:
        x, y = x[:b].contiguous(), y[:b].contiguous()
    for _ in range(max(0, warmup - 2)):
        O.train_step(sd, x, y, st)
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        O.train_step(sd, x, y, st)
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return done * b / dt, b, done, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    val, b, done, dt = cpu_train_steps(args.batch, args.seq_len, args.steps, args.warmup, budget_s=150.0)
    cores = torch.get_num_threads()
    sample = f"{done} steps of batch {b} x 12 x {args.seq_len} (oracle port of the reference modules, fp32, {cores} threads)"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": done,
        "warmup": args.warmup, "ms_per_step": 1000.0 * dt / max(done, 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[1]: ECGCNN(12,256,5) train step, 12x1000, AdamW(1.5e-3,1e-4)",
                   "batch_per_step": b, "seq_len": args.seq_len},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- roofline helpers
def kernel_model(name: str, B: int, chan, Ls, dtype_bytes: int = 4, cin_pad: int = 0):
    """Algorithmic FLOPs and HBM bytes of one C-ABI call of the step (SURVEY 8d model)."""
    if "_L" not in name:
        return None
    base, l = name.rsplit("_L", 1)
    l = int(l) - 1
    ci, co, L = chan[l], chan[l + 1], Ls[l]
    flops = 2.0 * B * L * co * ci * 15                    # algorithmic: real input channels only
    xin, yout = B * ci * L * dtype_bytes, B * co * L * dtype_bytes
    if base in ("conv_fwd", "wgrad", "dgrad"):
        return {"flops": flops, "bytes": xin + yout}
    if base == "bn_relu_pool":
        return {"flops": 0.0, "bytes": yout + yout / 2}
    if base == "bn_bwd":
        return {"flops": 0.0, "bytes": 2 * (yout + yout / 2) + yout}
    if base == "bn_stats":
        return {"flops": 0.0, "bytes": 0.0}
    return None


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import ptbxl_multimodal_b200 as P
    from ptbxl_multimodal_b200.step import TrainStep

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the ecgb200 path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, T, K, W = args.batch, args.seq_len, args.steps, max(args.warmup, 3)
    precision = "bf16" if args.precision == "auto" else args.precision      # BASELINE metric is quoted in bf16

    torch.manual_seed(42)
    model = P.ECGCNN(12, 256, 5).to(dev).train()
    opt = P.FusedAdamW(model.parameters(), lr=1.5e-3, weight_decay=1e-4)
    eng = TrainStep(model, opt, B, T, precision=precision)

    # synthetic data: NB distinct batches, resident on device and in pinned host memory
    NB = 8
    g = torch.Generator().manual_seed(1000 + rank)
    prev = torch.tensor([0.25, 0.24, 0.12, 0.23, 0.44])
    hx = [torch.randn(B, 12, T, generator=g).pin_memory() for _ in range(NB)]
    hy = [(torch.rand(B, 5, generator=g) < prev).float().pin_memory() for _ in range(NB)]
    dx = [t.to(dev) for t in hx]
    dy = [t.to(dev) for t in hy]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    # ---- device-resident arm (value): the batches live in the engine's two input slots (HBM); a step = one graph
    # replay on the slot's graph.  The 12 MB input is evicted from L2 between steps by the step's own ~0.27 GB of
    # activation traffic.
    eng.load_batch(dx[0], dy[0], slot=0)
    eng.load_batch(dx[1], dy[1], slot=1)

    def step_dev(i):
        eng.run(slot=i & 1)

    for i in range(W):
        step_dev(i)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms = timed(step_dev, K)
    clocks = sampler.stop() if rank == 0 else None
    value = world * B * K / (ms / 1000.0)

    # ---- end-to-end arm: every step its own pinned-host batch in (H2D on a copy stream, straight into the idle
    # input slot while the other slot's graph runs) and the loss out (D2H)
    copy_stream = torch.cuda.Stream(device=dev)
    staged = [torch.cuda.Event() for _ in range(2)]
    freed = [torch.cuda.Event() for _ in range(2)]
    host_loss = torch.zeros(64, dtype=torch.float32).pin_memory()
    main_stream = torch.cuda.current_stream(dev)

    def prefetch(i):
        s = i & 1
        copy_stream.wait_event(freed[s])                  # the graph that read slot s has finished
        with torch.cuda.stream(copy_stream):
            eng.load_batch(hx[i % NB], hy[i % NB], slot=s)
            staged[s].record(copy_stream)

    def step_e2e(i):
        s = i & 1
        if i == 0:
            prefetch(0)
        prefetch(i + 1)                                   # next batch streams in under this step's compute
        main_stream.wait_event(staged[s])
        loss = eng.run(slot=s)
        freed[s].record(main_stream)
        host_loss[i % 64].copy_(loss, non_blocking=True)  # D2H read of the step's result

    for s in range(2):
        freed[s].record(main_stream)
    for i in range(3):
        step_e2e(i)
    torch.cuda.synchronize(dev)
    for s in range(2):
        freed[s].record(main_stream)
    ms_e2e = timed(step_e2e, K)
    e2e_value = world * B * K / (ms_e2e / 1000.0)
    h2d = hx[0].numel() * 4 + hy[0].numel() * 4
    last_loss = float(host_loss[(K - 1) % 64])

    # ---- roofline.  N == 1: the dominant kernel, timed live (CUDA events around a graph of 10 back-to-back
    # launches on the replay stream, after the timed region); N > 1: no rank-local kernel replay is possible
    # (the optimizer kernel is a cross-rank barrier), so the whole step is held against the per-layer model.
    line_extra = {}
    ns_model = (929.0 if precision == "fp32" else 693.0) * T / 1000.0       # ns per window, T = 1000
    measured_ns = 1e6 * ms / (K * B)                                        # per window per GPU
    if rank == 0:
        pk = peaks()
        line_extra["step_roofline"] = {"model_ns_per_sample": ns_model, "measured_ns_per_sample": measured_ns,
                                       "frac": ns_model / measured_ns}
        if world == 1:
            prof = eng.time_kernels(iters=10)
            tot = sum(t for _, t in prof)
            name, t_ms = max(prof, key=lambda kv: kv[1])
            km = kernel_model(name, B, eng.chan, eng.L, dtype_bytes=2 if precision == "bf16" else 4)
            ridge = pk["tf_burst"] * 1e12 / (pk["hbm_gbs"] * 1e9)
            if km and km["flops"] > 0 and km["flops"] / max(km["bytes"], 1.0) > ridge:
                ach = km["flops"] / (t_ms * 1e-3) / 1e12
                roof = {"kernel": name, "bound": "tensor", "achieved": ach, "peak": pk["tf_burst"], "unit": "TFLOP/s",
                        "frac": ach / pk["tf_burst"], "traffic": None}
            else:
                by = km["bytes"] if km else 0.0
                ach = by / (t_ms * 1e-3) / 1e9
                roof = {"kernel": name, "bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s",
                        "frac": ach / pk["hbm_gbs"], "traffic": None}
            # dram__bytes_read.sum + dram__bytes_write.sum per launch from the ncu --set full capture of the same kernels
            # at this workload (profiles/r01_ncu_full_v3.md); cold-cache, i.e. compulsory traffic
            ncu_traffic = {"wgrad_L4": 24.9e6, "wgrad_L3": 25.0e6, "wgrad_L2": 24.6e6, "wgrad_L1": 24.6e6, "conv_fwd_L4": 9.3e6,
                           "conv_fwd_L3": 8.5e6, "conv_fwd_L2": 8.4e6, "conv_fwd_L1": 8.3e6, "dgrad_L4": 17.5e6,
                           "dgrad_L3": 16.7e6, "dgrad_L2": 16.5e6}
            if B == 256 and T == 1000 and precision == "bf16":
                roof["traffic"] = ncu_traffic.get(name)
            roof.update({"peak_source": pk["src"] + " (burst: kernel timed alone, CUDA events around a graph of 10 launches)",
                         "kernel_ms": t_ms, "share_of_step": t_ms / tot,
                         "algorithmic_flops": km["flops"] if km else None,
                         "algorithmic_bytes": km["bytes"] if km else None})
            line_extra["kernels_ms"] = {n: round(t, 4) for n, t in sorted(prof, key=lambda kv: -kv[1])[:12]}
            line_extra["kernels_total_ms"] = tot
        else:
            flops = 2.0 * 334080.0 * T                                      # conv FLOPs of one train step per window
            ach = flops * value / 1e12
            peak = pk["tf_sustained"] * world
            roof = {"kernel": "whole train step (all ranks)", "bound": "tensor", "achieved": ach, "peak": peak,
                    "unit": "TFLOP/s", "frac": ach / peak, "traffic": None,
                    "peak_source": pk["src"] + " (sustained, x n_gpus)", "algorithmic_flops": flops * B * world}
        line_extra["roofline"] = roof

    # ---- side figure (N == 1): the eval forward of the same model through the bf16 inference engine (InferStep:
    # BN/ReLU/pool/GAP fused into the conv epilogue), device-timed like `value`; not part of the train-step metric
    if rank == 0 and world == 1 and precision == "bf16":
        model.eval()
        inf = P.InferStep(model, B, T)
        inf.load_batch(dx[0], slot=0)
        inf.load_batch(dx[1], slot=1)
        inf.capture()
        for i in range(5):
            inf.run(slot=i & 1)
        ms_inf = timed(lambda i: inf.run(slot=i & 1), 100)
        line_extra["infer"] = {"value": B * 100 / (ms_inf / 1000.0), "unit": UNIT, "ms_per_batch": ms_inf / 100,
                               "launches_per_batch": inf.launches_per_batch,
                               "roofline_frac": 176.7 * T / 1000.0 * 1e-9 * B / (ms_inf / 100 * 1e-3),
                               "note": "eval forward, bf16 tcgen05 engine, inputs resident; 176.7 ns/window fused-inference model"}
        model.train()

    # ---- CPU baseline beside the GPU number (rank 0, N == 1 only; bounded sample)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, b, done, dt = cpu_train_steps(B, T, steps=80, warmup=2, budget_s=15.0)
        line_extra["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                                      "sample": f"{done} steps of batch {b} x 12 x {T}, oracle port of the reference "
                                                f"modules on the host CPU, fp32, {dt:.1f} s"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if precision == "fp32" else "bf16", "data": "synthetic",
            "config": {"workload": "configs[1]: ECGCNN(12,256,5) train step, synthetic 12x1000, 5-label BCE, "
                                   "AdamW(1.5e-3,1e-4)", "batch_per_gpu": B, "global_batch": B * world,
                       "seq_len": T, "parallelism": f"dp{world}", "precision": precision,
                       "l2": "step working set ~0.27 GB > 126 MB L2; inputs alternate between the engine's two resident input slots"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / K, "last_loss": last_loss},
            "gpu_launches": eng.launches_per_step * K,
        }
        line.update(line_extra)
        line["config"]["grad_exchange"] = ("fused peer-memory reduce-scatter + AdamW + all-gather kernel" if eng.dp_fused
                                           else ("nccl all-reduce" if world > 1 else "none"))
        print(json.dumps(line), flush=True)
    if world > 1:
        # NVLink peer mappings + NCCL teardown can block at interpreter exit: synchronise, then leave directly
        dist.barrier()
        torch.cuda.synchronize(dev)
        sys.stdout.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
