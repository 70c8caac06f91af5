"""DRAM traffic of a whole sweep (or a stratified sample of its launches) next to the ALGORITHMIC bytes of the very
same launches ("back the config-4 fractions with counters").

Step 1 (on the GPU box, under ncu):
    ADEPT_CUDA_PROFILE_LOG=gpurun_out/prof_<tag>.log ncu --profile-from-start off --clock-control none \
        --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv \
        --log-file gpurun_out/traffic_<tag>.csv python tools/ncu_traffic.py run <workload> <mode> <every>
  `run` uploads the workload's tape, sweeps one warm-up Jacobian, then one Jacobian with the engine option
  "profile_every" = <every>: the engine brackets every <every>-th launch of the level sweep with
  cudaProfilerStart/Stop and logs, per bracketed launch, the step's statement and record counts.
Step 2 (anywhere):
    python tools/ncu_traffic.py summarise gpurun_out/traffic_<tag>.csv gpurun_out/prof_<tag>.log <mode>
  prints a JSON summary: launches sampled, sum of DRAM bytes, sum of algorithmic bytes of those launches
  (forward 8 B, reverse 16 B per record x direction), their ratio, and the same per launch.
"""
import csv
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(workload, mode, every, dirs=0):
    import numpy as np
    import torch
    sys.argv = [sys.argv[0]]
    import bench
    pkg = importlib.import_module("adept-2_b200")
    wl = bench.WORKLOADS[workload]
    t = bench.build_tape(wl)
    eng = pkg.Engine([0])
    eng.upload_tape(t["stmt"], t["mult"], t["idx"], t["max_gradient"])
    indep, dep = t["indep"], t["dep"]
    if dirs:
        if mode == 0:
            indep = np.ascontiguousarray(indep[:dirs])
        else:
            dep = np.ascontiguousarray(dep[:dirs])
    out = torch.empty(indep.size * dep.size, dtype=torch.float64, device="cuda")
    eng.jacobian_device(mode, indep, dep, out.data_ptr())          # warm-up (builds the reverse stream)
    eng.set_option("profile_every", every)
    eng.jacobian_device(mode, indep, dep, out.data_ptr())
    eng.set_option("profile_every", 0)
    print("profiled sweep: replay_ms", eng.stat("replay_ms"), "sweep_ms", eng.stat("sweep_ms"), "launches", eng.stat("sweep_launches"),
          "passes", eng.stat("n_passes"))
    eng.close()


def summarise(csv_path, log_path, mode, label=""):
    rows = [r for r in csv.reader(l for l in open(csv_path) if l.startswith('"'))]
    hdr = rows[0]
    iname, imetric, ival, ikern = hdr.index("ID"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Kernel Name")
    per = {}
    for r in rows[1:]:
        d = per.setdefault(int(r[iname]), {"kernel": r[ikern]})
        d[r[imetric]] = float(r[ival].replace(",", ""))
    launches = [per[k] for k in sorted(per)]
    log = [l.split() for l in open(log_path) if l.strip()]
    if len(log) == 2 * len(launches) and log[:len(launches)] == log[len(launches):]:
        log = log[:len(launches)]      # the log was flushed twice under ncu (same lines, same order)
    assert len(log) == len(launches), (len(log), len(launches))
    per_rec = 8.0 if mode == 0 else 16.0
    tot_dram = tot_alg = tot_ns = 0.0
    worst = None
    for l, p in zip(log, launches):
        recs, dirs = float(l[l.index("recs") + 1]), float(l[l.index("dirs") + 1])
        alg = per_rec * recs * dirs
        dram = p["dram__bytes_read.sum"] + p["dram__bytes_write.sum"]
        tot_dram += dram
        tot_alg += alg
        tot_ns += p["gpu__time_duration.sum"]
        if alg > 0 and (worst is None or dram / alg > worst[0]):
            worst = (dram / alg, l)
    n = len(launches)
    out = {"label": label, "kernel": launches[0]["kernel"].split("(")[0] if n else None, "mode": "reverse" if mode else "forward",
           "launches_sampled": n, "dram_bytes_sum": tot_dram, "algorithmic_bytes_sum": tot_alg,
           "traffic_over_algorithmic": tot_dram / tot_alg if tot_alg else None,
           "dram_bytes_per_launch": tot_dram / n if n else None, "algorithmic_bytes_per_launch": tot_alg / n if n else None,
           "gpu_time_sum_ms_under_ncu": tot_ns / 1e6, "dram_GBps_under_ncu": tot_dram / tot_ns if tot_ns else None,
           "worst_launch_ratio": worst[0] if worst else None}
    print(json.dumps(out))
    return out


if __name__ == "__main__":
    if sys.argv[1] == "run":
        run(sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]) if len(sys.argv) > 5 else 0)
    else:
        summarise(sys.argv[2], sys.argv[3], int(sys.argv[4]), sys.argv[5] if len(sys.argv) > 5 else "")
