"""Times adept_cuda_upload_tape from PAGEABLE host arrays (what the C++ drop-in hands over) vs pinned ones."""
import importlib, sys, time
import numpy as np, torch
sys.path.insert(0, '.')
pkg = importlib.import_module("adept-2_b200")
t = pkg.tapegen.advection_tape("toon", 4096, 1000)
eng = pkg.Engine([0])
pin = {k: torch.from_numpy(np.ascontiguousarray(t[k])).pin_memory().numpy() for k in ("stmt", "mult", "idx")}
for name, arrs in (("pinned", pin), ("pageable", t)):
    for i in range(3):
        t0 = time.perf_counter()
        eng.upload_tape(arrs["stmt"], arrs["mult"], arrs["idx"], t["max_gradient"])
        print(name, i, "call ms %.1f" % (1e3 * (time.perf_counter() - t0)), "h2d ms %.1f" % eng.stat("upload_ms"),
              "schedule ms %.1f" % eng.stat("analysis_ms"), flush=True)
