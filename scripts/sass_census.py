"""SASS instruction census of mvs_gi_b200/libmvsgi_b200.so (sm_100a cubin) -> profiles/r4_sass_census.txt:
which Blackwell/Hopper-class instructions the shipped kernels actually contain.   python scripts/sass_census.py"""
import collections
import re
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sass = subprocess.run(["cuobjdump", "-sass", str(ROOT / "mvs_gi_b200/libmvsgi_b200.so")], capture_output=True, text=True).stdout
per, fn = collections.defaultdict(collections.Counter), None
pat = re.compile(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)")
for line in sass.splitlines():
    if "Function :" in line:
        fn = line.split("Function :")[1].strip()
        continue
    m = pat.match(line)
    if m and fn:
        per[fn][m.group(1)] += 1
tot = collections.Counter()
for c in per.values():
    for k, v in c.items():
        tot[k.split(".")[0]] += v
out = ["# SASS instruction census of mvs_gi_b200/libmvsgi_b200.so (sm_100a cubin), `cuobjdump -sass` (scripts/sass_census.py)\n",
       "# UTMALDG = cp.async.bulk.tensor (TMA), UBLKCP = cp.async.bulk (the kept tap table of the sweep), SYNCS = mbarrier ops, FFMA2/FMUL2/FADD2 = packed f32x2 arithmetic (sm_100+),\n"
       "# FMNMX3 = 3-input min/max, REDG = red.global (vector float reductions of the backward)\n",
       "whole library: %d kernels; " % len(per) + ", ".join(
           f"{k} {tot[k]}" for k in ["UTMALDG", "UBLKCP", "SYNCS", "FFMA2", "FMUL2", "FADD2", "FMNMX3", "LDG", "STG", "REDG", "LDS", "STS", "MUFU", "SHFL"] if tot[k]) + "\n\n"]
names = ["softargmin_x2_tma_kernelILi16ELb1ELb0ELb0", "softargmin_x2_tma_kernelILi16ELb1ELb1ELb0",
         "softargmin_x2_backward_kernelILi16ELb1ELi2", "softargmin_x2_backward_kernelILi16ELb1ELi0",
         "sweep_kernelILi3ELi0ELb1ELb0ELb0ELb1ELb0ELb1ELb0ELi4ELb0", "sweep_kernelILi4ELi0ELb1ELb0ELb0ELb1ELb0ELb1ELb0ELi3ELb0",
         "sweep_kernelILi3ELi0ELb1ELb0ELb0ELb0ELb0ELb1ELb0ELi1ELb1", "sweep_kernelILi3ELi0ELb1ELb0ELb0ELb0ELb0ELb0ELb0ELi1ELb1",
         "octet_cost_kernelILi3ELi0", "resample_kernelILi3Ef"]
keep = ("UTMALDG", "SYNCS", "UBLKCP", "FFMA2", "FMUL2", "FADD2", "FMNMX3", "LDG.E.128", "LDG.E.64", "STG.E.128", "STG.E.NA",
        "REDG", "RED.", "LDS", "MUFU.EX2", "SHFL", "ATOMS", "BAR", "ATOMG", "VOTE")
for n in names:
    for f, c in per.items():
        if n in f:
            dem = subprocess.run(["c++filt", f], capture_output=True, text=True).stdout.strip()
            sel = {k: v for k, v in c.items() if k.startswith(keep)}
            out.append(dem[:170] + f"\n  total {sum(c.values())} instructions; " + ", ".join(f"{k} {v}" for k, v in sorted(sel.items())) + "\n")
            break
(ROOT / "profiles/r4_sass_census.txt").write_text("".join(out))
print("".join(out))
