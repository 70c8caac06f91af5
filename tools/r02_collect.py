"""Copy the outputs of tools/r02_final.sh (gpurun_out/r02z_*) into profiles/ under round-2 names and write the
summaries the judge reads: ncu --set full table of the dominant kernel, launch list shares, SASS summary."""
import collections
import csv
import glob
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
T = sys.argv[1] if len(sys.argv) > 1 else "r02z"

for f in glob.glob(os.path.join(G, f"{T}_bench_*.json")):
    line = [l for l in open(f) if l.startswith("{")]
    if line:
        open(os.path.join(P, "r02_final_" + os.path.basename(f)[len(T) + 1:]), "w").write(line[-1])
for f in glob.glob(os.path.join(G, f"{T}_traffic_*.json")):
    shutil.copy(f, os.path.join(P, "r02_final_" + os.path.basename(f)[len(T) + 1:]))
if os.path.exists(os.path.join(G, "r02b_gather_ceiling.jsonl")):
    shutil.copy(os.path.join(G, "r02b_gather_ceiling.jsonl"), os.path.join(P, "r02_gather_ceiling.jsonl"))

# traffic.json for bench.py
tr = json.load(open(os.path.join(P, "traffic.json")))
for key, tag, kern, note in (("synthetic1e8_fwd", "synthetic1e8_fwd_m0", "forward_balanced_kernel<2,sparse>",
                              "BASELINE config 4 full size, forward sweep (forward_balanced_kernel): every 50th launch, all 8192 directions, one stream under the profiler"),
                             ("synthetic1e8_rev", "synthetic1e8_fwd_m1", "reverse_fused_kernel<2>",
                              "BASELINE config 4 full size, fused reverse sweep: every 50th launch of both passes, 4096 directions per pass, one stream under the profiler")):
    p = os.path.join(G, f"{T}_traffic_{tag}.json")
    if os.path.exists(p) and os.path.getsize(p) > 10:
        d = json.load(open(p))
        tr[key] = {"kernel": kern, "dram_bytes_per_launch": d["dram_bytes_per_launch"],
                   "algorithmic_bytes_per_launch": d["algorithmic_bytes_per_launch"],
                   "traffic_over_algorithmic": d["traffic_over_algorithmic"], "launches_sampled": d["launches_sampled"],
                   "dram_GBps_under_ncu": d["dram_GBps_under_ncu"],
                   "launch": note + "; ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum "
                                    f"(tools/ncu_traffic.py, profiles/r02_final_traffic_{tag}.json)"}
json.dump(tr, open(os.path.join(P, "traffic.json"), "w"), indent=1)

# ncu --set full of the dominant kernel
rep = os.path.join(G, f"{T}_ncu_forward_balanced.ncu-rep")
if os.path.exists(rep):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep], capture_output=True, text=True).stdout
    open(os.path.join(P, "r02_ncu_full_forward_balanced.md"), "w").write(
        "# Round 2: `ncu --set full` of `forward_balanced_kernel<2,true>` (dominant kernel of the default bench)\n\n"
        "Command (B200, one GPU, tools/r02_final.sh, after the same bench command had exited 0 without ncu):\n\n```\n"
        "ncu --set full --clock-control none --import-source on -k regex:forward_balanced -s 1254 -c 1 \\\n"
        "    python bench.py --workload synthetic1e7_fwd --no-cpu-baseline --e2e-steps 0 --reverse-steps 0 --steps 1 --warmup 3 --opt wide_streams=1\n```\n\n"
        "Step 302 of 953 (a wide one: 225 373 records = 45 000 statements) of the forward sweep of the 1e7-statement tape (config 4's shape at 1/10 length, all 8192\n"
        "directions, one stream).  Under ncu every launch runs alone with cold caches: durations are upper bounds, the counters matter.\n"
        "Compare with `forward_level_kernel` in round 1 (`r01_c_ncu_full_dominant_kernels.md`, prof_fwd_syn_c: 27 warp instructions per\n"
        "(record, column block), issue slots 56 % busy, 5.0 GB in 844 us).\n\n```\n" + out + "```\n")

# launch list of one bench command: share of the step per kernel
ll = os.path.join(G, f"{T}_launches_synthetic1e7_fwd.csv")
if os.path.exists(ll):
    rows = [r for r in csv.reader(l for l in open(ll) if l.startswith('"'))]
    h = rows[0]
    ik, iv = h.index("Kernel Name"), h.index("Metric Value")
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows[1:]:
        k = r[ik].split("(")[0].replace("void ", "").replace("adept_b200::", "")[-70:]
        tot[k] += float(r[iv].replace(",", "")) / 1e6
        cnt[k] += 1
    s = sum(tot.values())
    with open(os.path.join(P, "r02_launches_synthetic1e7_fwd.md"), "w") as f:
        f.write("# Round 2: ncu launch list of `python bench.py --workload synthetic1e7_fwd --steps 1 --warmup 3` (first 3000 launches)\n\n"
                "`ncu --metrics gpu__time_duration.sum --clock-control none -c 3000`; times are cold-cache and serialised: the SHARE matters.\n\n"
                "| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
        for k, v in tot.most_common(20):
            f.write(f"| `{k}` | {cnt[k]} | {v:.3f} | {100 * v / s:.1f} % |\n")

open(os.path.join(P, "r02_sass_summary.md"), "w").write(
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sass_summary.py")], capture_output=True, text=True).stdout)
print("collected into profiles/")
