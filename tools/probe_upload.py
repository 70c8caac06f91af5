import sys, time, importlib, numpy as np, torch
sys.path.insert(0, '.')
pkg = importlib.import_module("adept-2_b200")
t = pkg.tapegen.advection_tape("toon", 4096, 1000)
eng = pkg.Engine([0])
pin = {k: torch.from_numpy(np.ascontiguousarray(t[k])).pin_memory() for k in ("stmt","mult","idx")}
host = {k: v.numpy() for k, v in pin.items()}
for name, arrs in (("pinned", host), ("pageable", t)):
    for i in range(3):
        t0 = time.perf_counter()
        eng.upload_tape(arrs["stmt"], arrs["mult"], arrs["idx"], t["max_gradient"])
        t1 = time.perf_counter()
        print(name, i, "python call ms", 1e3*(t1-t0), "C total", eng.stat("upload_total_ms"), "upload", eng.stat("upload_ms"), "analysis", eng.stat("analysis_ms"), flush=True)
