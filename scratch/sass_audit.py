"""Static audit of the built library (no GPU needed): which Blackwell instructions the kernels really contain
(cuobjdump -sass) and registers / stack (= spill space) / static shared memory per kernel (cuobjdump -res-usage).
    python scratch/sass_audit.py > profiles/r02_sass_audit.md"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "ptbxl_multimodal_b200", "libecgb200.so")
MNEMONICS = [("UTCHMMA", "tcgen05.mma kind::f16 (bf16 operands, fp32 accumulate in TMEM)"),
             ("UTCHMMA.2CTA", "of which cta_group::2 (one instruction spans the CTA pair, M = 256)"),
             ("UTMALDG", "cp.async.bulk.tensor (TMA tensor-map loads)"),
             ("UBLKCP", "cp.async.bulk (1-D bulk copies)"),
             ("LDTM", "tcgen05.ld (TMEM -> registers, epilogues)"),
             ("UTCBAR", "tcgen05.commit (-> mbarrier, incl. multicast::cluster)"),
             ("SYNCS", "mbarrier arrive / try_wait / expect_tx"),
             ("UCGABAR", "barrier.cluster (CTA-pair set-up / tear-down)")]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout.splitlines()
    print("# Static audit of `ptbxl_multimodal_b200/libecgb200.so` (sm_100a; `python scratch/sass_audit.py`)\n")
    print("| SASS mnemonic | count | PTX it comes from |\n|---|---|---|")
    for m, what in MNEMONICS:
        n = len(re.findall(r"\s" + re.escape(m), sass))          # mnemonic as a prefix: suffixes (.4D, .x16, .2CTA ...) included
        print(f"| `{m}` | {n} | {what} |")
    print(f"| `HMMA` / `WGMMA` (legacy mma.sync / wgmma) | {len(re.findall(r' HMMA', sass))} / {len(re.findall(r'WGMMA', sass))} | none: every MMA is tcgen05 |")
    rows = []
    for i, line in enumerate(res):
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            u = res[i + 1]
            g = lambda k: int(re.search(k + r":(\d+)", u).group(1)) if re.search(k + r":(\d+)", u) else 0   # noqa: E731
            rows.append((m.group(1), g("REG"), g("STACK"), g("SHARED")))
    names = subprocess.run(["c++filt"] + [r[0] for r in rows], capture_output=True, text=True).stdout.splitlines()
    rows = sorted((re.sub(r"\(.*", "", n).replace("void ", ""), r, s, sh) for (_, r, s, sh), n in zip(rows, names))
    print(f"\n{len(rows)} kernels.  Registers / stack bytes per thread (stack > 0 = spill space or local arrays) / static shared "
          "memory (the tensor kernels take their 200+ KB dynamically):\n")
    print("| kernel | regs | stack B | static smem B |\n|---|---|---|---|")
    for n, r, s, sh in rows:
        print(f"| `{n}` | {r} | {s} | {sh} |")


if __name__ == "__main__":
    sys.exit(main())
