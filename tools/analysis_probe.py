"""Upload one workload's tape (schedule construction included) N times and print the phase timings; run under
`ncu --metrics gpu__time_duration.sum` to get the per-kernel times of the device-side dependency analysis."""
import importlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.argv, args = [sys.argv[0]], sys.argv[1:]
import bench
pkg = importlib.import_module("adept-2_b200")
name, reps = args[0], int(args[1]) if len(args) > 1 else 2
if name == "rosenbrock1e7":
    t = pkg.tapegen.rosenbrock_tape(10_000_000)
else:
    t = bench.build_tape(bench.WORKLOADS[name])
eng = pkg.Engine([0])
for kv in args[2:]:
    k, v = kv.split("=")
    eng.set_option(k, int(v))
for _ in range(reps):
    eng.upload_tape(t["stmt"], t["mult"], t["idx"], t["max_gradient"])
    print(json.dumps({k: eng.stat(k) for k in ("upload_ms", "analysis_ms", "an_validate_ms", "an_window_ms", "an_pack_ms", "an_forward_ms", "n_levels")}))
eng.close()
