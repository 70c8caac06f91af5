"""SASS instruction summary of every kernel in libadept_cuda.so (run where cuobjdump is: no GPU needed).
    python tools/sass_summary.py > profiles/rNN_sass_summary.md
What the table shows: UBLKCP = 1-D bulk TMA (cp.async.bulk global->shared, the right instruction for a linear record
stream; there is no UTMALDG because nothing here is a tiled tensor), SYNCS = mbarrier operations, DFMA must be 0
(-fmad=false + __dmul_rn/__dadd_rn: the arithmetic contract with the reference), DMUL/DADD are the replay arithmetic."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "adept-2_b200", "libadept_cuda.so")
txt = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)[1:]
rows = []
for f in funcs:
    name = f.split("\n", 1)[0].strip()
    ops = collections.Counter(re.findall(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", f, flags=re.M))
    c = lambda p: sum(v for k, v in ops.items() if k.startswith(p))
    rows.append((name, sum(ops.values()), c("UBLKCP"), c("SYNCS"), c("DFMA"), c("DMUL"), c("DADD"), c("LDG"), c("STG"), c("LDS"), c("STS"),
                 c("MATCH"), c("ATOM") + c("RED")))
dem = subprocess.run(["c++filt"] + [r[0] for r in rows], capture_output=True, text=True).stdout.splitlines()
print("# SASS summary of libadept_cuda.so (sm_100a), `cuobjdump -sass` via tools/sass_summary.py\n")
print((__doc__ or "").split("What the table shows:")[1].strip() + "\n")
print("| kernel | SASS instr | UBLKCP | SYNCS | DFMA | DMUL | DADD | LDG | STG | LDS | STS | MATCH | ATOM/RED |")
print("|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
tot_dfma = 0
for r, d in sorted(zip(rows, dem), key=lambda x: x[1]):
    d = re.sub(r"\(.*\)$", "", d).replace("adept_b200::", "").replace("void ", "").replace("(anonymous namespace)::", "")
    tot_dfma += r[4]
    if not d or "cub::" in d or "thrust" in d:
        continue
    print("| `%s` | %s |" % (d, " | ".join(str(x) for x in r[1:])))
print(f"\nDFMA over the whole library (CUB kernels included): {tot_dfma}")
