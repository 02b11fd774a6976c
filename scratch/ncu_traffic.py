"""From an `ncu --set full` capture of ONE step of scratch/ncu_step.py (`-k regex:<pattern> -s 58 -c 29`: the third step),
exported with `ncu -i X.ncu-rep --page raw --csv > X.csv`, write
  profiles/ncu_traffic.json   {"source": ..., "calls": {call name: dram__bytes_read.sum + dram__bytes_write.sum per call}}
  (bench.py reads it for roofline.traffic) and print the per-launch table (markdown) with the call each launch belongs to.

    python scratch/ncu_traffic.py /tmp/r02_full.csv profiles/r02_ncu_full.ncu-rep-name > profiles/r02_ncu_full.md
"""
import csv, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = list(csv.reader(open(sys.argv[1])))
src = sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3,
         "msecond": 1e3}


def g(r, k, scaled=False):
    try:
        v = float(r[idx[k]].replace(",", ""))
    except (ValueError, KeyError):
        return float("nan")
    return v * SCALE.get(units[idx[k]], 1.0) if scaled else v


# launch order of one bf16 step (host enqueue order = ncu's serialised order)
SEQ = {"wfdb16_zscore_pack_kernel": ["decode"], "conv_tc_kernel<0>": ["conv_fwd_L1", "conv_fwd_L2"],
       "conv_tc_pair_kernel<0>": ["conv_fwd_L3", "conv_fwd_L4"],
       "bn_fwd_train_bf16_kernel": [f"bn_relu_pool_L{l}" for l in (1, 2, 3, 4)], "head_fwd_bwd_kernel": ["head_fwd_bwd"],
       "bn_bwd_reduce_bf16_kernel": [f"bn_bwd_L{l}" for l in (3, 2, 1)], "bn_bwd_apply_bf16_kernel": [f"bn_bwd_L{l}" for l in (4, 3, 2, 1)],
       "conv_tc_pair_kernel<3>": ["dgrad_L4", "dgrad_L3"], "conv_tc_kernel<3>": ["dgrad_L2"],
       "wgrad_pair_kernel": ["wgrad_L4", "wgrad_L3"],
       "wgrad_thin_kernel": ["wgrad_L2", "wgrad_L1"], "wgrad_tc_reduce_kernel": [f"wgrad_L{l}" for l in (4, 3, 2, 1)]}
seen = {}
calls = {}
aux = {}
lines = []
for r in rows[2:]:
    full = r[idx["Kernel Name"]].replace("void ", "")
    base = full.split("(")[0]
    key = next((k for k in SEQ if base.startswith(k)), None)
    call = "?"
    if key is not None:
        n = seen.get(key, 0)
        seen[key] = n + 1
        call = SEQ[key][n % len(SEQ[key])]
    rd, wr = g(r, "dram__bytes_read.sum", True), g(r, "dram__bytes_write.sum", True)
    if key == "wgrad_tc_reduce_kernel":      # the split-K partials are L2-resident in the real step; ncu flushes the caches
        aux[call + " (split-K reduce launch, cold caches)"] = rd + wr
    else:
        calls[call] = calls.get(call, 0.0) + rd + wr
    lines.append(f"| {call} | {base} | {r[idx['Grid Size']]} | {g(r, 'gpu__time_duration.sum', True):.1f} | {rd / 1e6:.1f} | {wr / 1e6:.2f} | "
                 f"{g(r, 'sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active'):.1f} | {g(r, 'lts__t_sector_hit_rate.pct'):.1f} | "
                 f"{g(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | {g(r, 'lts__throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | "
                 f"{g(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):.1f} | {r[idx['launch__registers_per_thread']]} | "
                 f"{g(r, 'smsp__inst_executed.sum') / 1e6:.2f} |")
print("| call | kernel | grid | us | DRAM read MB | DRAM write MB | tensor pipe % | L2 hit % | DRAM % | L2 % | warps active % | regs | warp-instr M |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|---|")
print("\n".join(lines))
calls.pop("?", None)
with open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w") as f:
    json.dump({"source": src, "workload": "configs[1]: 256 x 12 x 1000, bf16 engine, raw int16 input",
               "metric": "dram__bytes_read.sum + dram__bytes_write.sum per C-ABI call (all launches of the call), cold caches",
               "calls": {k: round(v) for k, v in calls.items()},
               "not_counted": {k: round(v) for k, v in aux.items()}}, f, indent=1)
