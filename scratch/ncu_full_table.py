"""Summarise an `ncu --set full` capture exported with `ncu -i X.ncu-rep --page raw --csv`: one row per launch."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
SCALE = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}


def g(r, k, scaled=False):
    try:
        v = float(r[idx[k]].replace(",", ""))
    except (ValueError, KeyError):
        return float("nan")
    return v * SCALE.get(units[idx[k]], 1.0) if scaled else v


print("| kernel | grid | us | DRAM read MB | DRAM write MB | tensor pipe % | L2 hit % | DRAM % | L2 % | warps active % | regs |")
print("|---|---|---|---|---|---|---|---|---|---|---|")
for r in rows[2:]:
    print(f"| {r[idx['Kernel Name']].split('(')[0].replace('void ', '')} | {r[idx['Grid Size']]} | "
          f"{g(r, 'gpu__time_duration.sum', True):.1f} | {g(r, 'dram__bytes_read.sum', True):.1f} | "
          f"{g(r, 'dram__bytes_write.sum', True):.2f} | {g(r, 'sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active'):.1f} | "
          f"{g(r, 'lts__t_sector_hit_rate.pct'):.1f} | {g(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | "
          f"{g(r, 'lts__throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | "
          f"{g(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):.1f} | {r[idx['launch__registers_per_thread']]} |")
