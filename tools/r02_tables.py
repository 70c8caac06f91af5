"""Fill the three result tables of DESIGN.md section 6 from the bench lines archived under profiles/ (r02_final_*,
r02_scale_*): the numbers in the document are the numbers in the files."""
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, "profiles")


def L(name):
    p = os.path.join(P, name)
    return json.load(open(p)) if os.path.exists(p) and os.path.getsize(p) > 2 else None


def fmt(x, nd=2):
    return "—" if x is None else (f"{x:.{nd}e}" if abs(x) >= 1e5 else f"{x:.{nd}f}")


d = L("r02_final_bench_synthetic1e8_fwd.json")
ref = L("r02_final_bench_synthetic1e8_fwd_ref.json")
tf, tr = L("r02_final_traffic_synthetic1e8_fwd_m0.json"), L("r02_final_traffic_synthetic1e8_fwd_m1.json")
r, o, e = d["roofline"], d["reverse"], d["e2e"]
ph = e["phases_ms"]
cfg4 = f"""| quantity | forward (`value`, the default line) | reverse (same tape, `"reverse"` key) |
|---|---|---|
| replay, tape + schedule resident (zero-fill + seed + sweep + extraction) | **{fmt(d['value'])} op·dir/s**, {d['ms_per_step']:.0f} ms per Jacobian (round 1: 5.3e11, 6210 ms) | **{fmt(o['value'])} op·dir/s**, {o['ms_per_step']:.0f} ms (round 1: 3.6e11, 9080 ms), two passes of 4096 directions |
| `roofline.frac` = algorithmic bytes ÷ time of the replay kernels in the timed region ÷ 6549.4 GB/s | **{r['frac']:.3f}** ({r['achieved']:.0f} GB/s; {r['launches_per_step']:.0f} launches on {r['concurrent_streams']:.0f} streams; round 1: 0.805) | {o['roofline']['frac']:.3f} nominal |
| measured DRAM traffic ÷ algorithmic bytes of the same launches (ncu, every 50th launch) | **{tf['traffic_over_algorithmic']:.3f}** ({tf['launches_sampled']} launches, {tf['dram_GBps_under_ncu']:.0f} GB/s under ncu) | **{tr['traffic_over_algorithmic']:.3f}** ({tr['launches_sampled']} launches, {tr['dram_GBps_under_ncu']:.0f} GB/s under ncu) |
| ⇒ real DRAM bytes ÷ time ÷ 6549.4 GB/s | **{r['frac'] * tf['traffic_over_algorithmic']:.2f}** | **{o['roofline']['frac'] * tr['traffic_over_algorithmic']:.2f}** |
| `e2e`: pinned H2D of the 5.6 GB tape + schedule build + replay + D2H of the 537 MB Jacobian | **{fmt(e['value'])} op·dir/s**, {e['ms_per_step']:.0f} ms: statements+indices H2D {ph['upload_ms']:.0f} (multipliers overlapped with the analysis), analysis {ph['analysis_ms']:.0f} (walk {ph['an_window_ms']:.0f}, packing {ph['an_pack_ms']:.0f}, forward stream {ph['an_forward_ms']:.0f}), replay {ph['replay_ms']:.0f}, D2H {ph['d2h_ms']:.1f} (round 1: 6896 ms) | — |
| reference arm (`--impl reference`): OpenMP `Stack::jacobian_forward`, {ref['cpu_baseline']['cores']} host cores, AVX-512 build (P=8) | {fmt(ref['value'])} op·dir/s — `value` is **{d['value'] / ref['value']:.0f}×**, `e2e` **{e['value'] / ref['value']:.0f}×** | — |
| clocks during the timed region | SM {d['clocks']['sm_mhz']:.0f} of {d['clocks']['sm_max_mhz']:.0f} MHz, reasons {d['clocks']['reasons']} (the sweep runs into the power cap) | |
"""

rows = []
t = L("r02_final_bench_toon4096.json")
cpu2 = t['cpu_baseline'] or L('r02a_bench_toon4096.json')['cpu_baseline']
rows.append(f"| config 2, Toon nx=4096 nt=1000, 4096² **reverse** (`toon4096`), fused sweep, 4 streams, 3126 steps | {fmt(t['value'])} op·dir/s, {t['ms_per_step']:.0f} ms per Jacobian ({t['roofline']['sweep_ms_per_step']:.0f} ms of replay kernels; round 1: 81–89 ms) | effective {t['roofline']['frac']:.1f} (working set in L2: not a DRAM figure) | {fmt(cpu2['value'])} ({cpu2['cores']} cores) | {fmt(t['e2e']['value'])}, {t['e2e']['ms_per_step']:.0f} ms |")
s7 = L("r02_final_bench_synthetic1e7_fwd.json")
rows.append(f"| config 4 at 1/10 length, forward (`synthetic1e7_fwd`; reverse {s7['reverse']['ms_per_step']:.0f} ms) | {fmt(s7['value'])}, {s7['ms_per_step']:.0f} ms (round 1: 462–521) | {s7['roofline']['frac']:.2f} nominal; traffic ÷ algorithmic 0.63 | {fmt(s7['cpu_baseline']['value'])} | {fmt(s7['e2e']['value'])}, {s7['e2e']['ms_per_step']:.0f} ms |")
ro = L("r02_final_bench_rosenbrock1e7.json")
rp = ro["e2e"]["phases_ms"]
rows.append(f"| config 5, Rosenbrock N=1e7, one level-scheduled `compute_adjoint` (`rosenbrock1e7`) | {fmt(ro['value'])} entry/s, {ro['ms_per_step']:.2f} ms per sweep (gradient on the device) | {ro['roofline']['frac']:.2f} (37.6 B per entry) | {fmt(ro['cpu_baseline']['value'])} entry/s = {ro['e2e']['reference_sweep_ms']:.0f} ms per sweep, single-threaded by design | **a NEW tape per sweep: {ro['e2e']['ms_per_step']:.0f} ms** = H2D {rp['upload_ms']:.0f} + analysis {rp['analysis_ms']:.0f} + adjoint call {rp['adjoint_call_ms']:.0f} (reverse stream build, 480 MB of gradient over PCIe, sweep) — {'beats' if ro['e2e']['beats_the_reference_sweep'] else 'LOSES to'} the reference's {ro['e2e']['reference_sweep_ms']:.0f} ms (round 1: ≈ 280 ms) |")
en = L("r02_final_bench_ensemble1e5.json")
rows.append(f"| config 3, 1e5-member radiance ensemble, 2×26 reverse (`ensemble1e5`), real Adept recording of `simulate_radiances` on {en['config']['host_threads']} host threads | batch call {en['ms_per_step']:.2f} ms ({fmt(en['members_per_s'])} members/s; round 1: 17.3 ms) | — (PCIe/launch-bound) | reference's own `simulate_radiances()` under OpenMP: {fmt(en['cpu_baseline']['members_per_s'])} members/s (20.0 ms) | recording {en['e2e']['phases_ms']['host_recording_ms']:.1f} + batch {en['e2e']['phases_ms']['jacobian_batch_ms']:.1f} = **{en['e2e']['ms_per_step']:.1f} ms**, every member bit-identical |")
lw = L("r02_final_bench_lw100.json")
rows.append(f"| config 1, Lax-Wendroff nx=100 nt=2000, 100² forward (`lw100`), shared-memory kernel | {fmt(lw['value'])}, {lw['ms_per_step']:.1f} ms | {lw['roofline']['frac']:.2f} | **{fmt(lw['cpu_baseline']['value'])}** (the CPU wins: 100 directions, 6001 dependent steps) | {fmt(lw['e2e']['value'])} |")
ck = L("r02_final_bench_checkpoint.json")
cp = ck["e2e"]["phases_ms"]
rows.append(f"| checkpoint chain (`checkpoint`): `test/test_checkpoint.cpp` at its own size, NX=100, 100 blocks × 100 steps, recording inside the timed region | chain {cp['chain_total_ms']:.0f} ms, drop-in loop {cp['dropin_loop_total_ms']:.0f} ms | — | CPU sweeps of the same tapes: {cp['cpu_sweeps_total_ms']:.0f} ms in all | **the GPU loses at this size** ({ck['ms_per_step']:.1f} ms per 2e4-statement block: schedule build + ~600 dependent launches vs a 0.15 ms CPU sweep) |")
other = "| workload (`--workload`) | `value` | dominant kernel `frac` | CPU reference | `e2e` |\n|---|---|---|---|---|\n" + "\n".join(rows) + "\n"

one = d["ms_per_step"]
one_e = e["ms_per_step"]
one_r = o["ms_per_step"]
lines = [f"| 1 | **{one:.0f} ms** ({fmt(d['value'])}) | 1.00 | {r['frac']:.2f} | {one_e:.0f} ms | 1.00 | {one_r:.0f} ms | {t['ms_per_step']:.0f} ms | {en['ms_per_step']:.1f} / {en['e2e']['ms_per_step']:.1f} ms |"]
for n in (2, 4, 8):
    c = L(f"r02_scale_cfg4_n{n}.json")
    tt = L(f"r02_scale_toon4096_n{n}.json")
    ee = L(f"r02_scale_ensemble1e5_n{n}.json")
    if not c:
        continue
    lines.append(f"| {n} | **{c['ms_per_step']:.0f} ms** ({fmt(c['value'])}) | {one / c['ms_per_step'] / n:.2f} | {c['roofline']['frac']:.2f} | {c['e2e']['ms_per_step']:.0f} ms | {one_e / c['e2e']['ms_per_step']:.2f}× | "
                 f"{c['reverse']['ms_per_step']:.0f} ms ({one_r / c['reverse']['ms_per_step'] / n:.2f}) | {tt['ms_per_step']:.0f} ms | {ee['ms_per_step']:.1f} / {ee['e2e']['ms_per_step']:.1f} ms |" if tt and ee else "")
scale = ("| GPUs | config 4 forward, replay | efficiency | per-rank `frac` | config 4 `e2e` | e2e speed-up | config 4 reverse (efficiency) | config 2 (Toon 4096² reverse) | config 3 batch / e2e |\n"
         "|---|---|---|---|---|---|---|---|---|\n" + "\n".join(x for x in lines if x) + "\n")

p = os.path.join(ROOT, "DESIGN.md")
s = open(p).read()
for key, val in (("CFG4", cfg4), ("OTHER", other), ("SCALE", scale)):
    begin, end = f"<!-- TABLE_{key} -->", f"<!-- /TABLE_{key} -->"
    if f"RESULTS_TABLE_{key}" in s:
        s = s.replace(f"RESULTS_TABLE_{key}", begin + "\n" + val + end)
    else:
        s = re.sub(re.escape(begin) + ".*?" + re.escape(end), lambda m: begin + "\n" + val + end, s, flags=re.S)
open(p, "w").write(s)
print("DESIGN.md tables updated")
