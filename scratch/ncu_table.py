"""Summarise an ncu gpu__time_duration launch list: per-kernel times of the LAST complete step."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = None
launches = []
for r in rows:
    if 'Kernel Name' in r:
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    d = dict(zip(hdr, r))
    try:
        v = float(d['Metric Value'].replace(',', ''))
    except ValueError:
        continue
    u = d['Metric Unit']
    v = v / 1000 if u == 'ns' else (v * 1000 if u == 'ms' else v)
    launches.append((d['Kernel Name'].split('(')[0].replace('void ', ''), d['Grid Size'], d['Block Size'], v))
# last step = from the last step_prep_kernel (or pack_input) launch on
start = max(i for i, l in enumerate(launches) if 'step_prep' in l[0] or 'pack_input' in l[0])
step = launches[start:]
tot = sum(l[3] for l in step)
print(f"| # | kernel | grid | block | us | share |\n|---|---|---|---|---|---|")
for i, (k, g, b, v) in enumerate(step):
    print(f"| {i} | {k} | {g} | {b} | {v:.1f} | {100 * v / tot:.1f}% |")
print(f"\nsum {tot:.1f} us over {len(step)} launches")
