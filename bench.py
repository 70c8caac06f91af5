#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric: Jacobian tape-entry x direction per second (+ % of HBM roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...        (N > 1, one rank per GPU)

A "step" is one pass of the hot path over one batch of synthetic input: one whole Jacobian of the
workload's tape.  Default workload: BASELINE config 4 at FULL size -- the configuration the metric is
quoted on ("at 1/2/4/8 B200", the HBM-bound one): seeded random tape of 1e8
statements / 4.0e8 operations / 2^20 slots, 8192x8192 Jacobian, by FORWARD sweeps (what Stack::jacobian
dispatches to for a square Jacobian, include/adept/Stack.h:374-385); the REVERSE sweep of the same tape
is measured in the same run and printed under "reverse".  With N > 1 the directions of the same Jacobian
are partitioned over the ranks: tape ncclBroadcast from rank 0, every rank analyses its copy at the same
time, result blocks gathered on rank 0 (total work fixed => "scaling": "strong").

One JSON line on rank 0:
  value        whole-job op*dir/s with the tape and its schedule already resident in HBM when the timed
               region starts (what the reference times around stack_.jacobian(), differentiator.h:338-348);
               the timed region holds zero-fill + seeding + sweep + extraction (+ NCCL gather).
  e2e          the same metric through the C-ABI call a user makes, with HOST buffers: pinned H2D of the
               tape (+ broadcast + device-side schedule construction) and D2H of the Jacobian inside the
               timed region.
  roofline     dominant replay kernel: ALGORITHMIC bytes per launch / its average launch duration, both from
               the TIMED REGION ITSELF (CUDA events around the replay kernels of every pass on the engine's
               own stream, production launch mode: PDL on, same streams), against the measured HBM copy
               bandwidth (MEASURED_PEAKS.json).  "traffic" = DRAM bytes per launch from the committed ncu
               capture of the same workload (profiles/traffic.json), when there is one.
  cpu_baseline the reference's own OpenMP Stack::jacobian_* (oracle/_ref, built from the unmodified
               reference) on this box's host cores, on a bounded sample of the same Jacobian's directions.
--impl reference prints the reference's CPU arm as the main line instead (same config keys).
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Jacobian tape-entry x direction per second (batched replay of the recorded derivative tape)"
UNIT = "op*dir/s"

WORKLOADS = {
    # name: (description, tape builder, mode)
    "toon4096": dict(desc="BASELINE config 2: Toon advection nx=4096 nt=1000 (8195096 stmts / 36858096 ops / 12287 slots), "
                          "4096x4096 Jacobian, reverse sweeps", scheme="toon", nx=4096, nt=1000, mode=1),
    "synthetic1e6_fwd": dict(desc="BASELINE config 4 at 1/100 length (smoke runs): seeded random tape, 1e6 statements, 2^16 slots, "
                                  "1024x1024 Jacobian, forward sweeps", scheme="synthetic", n=1_000_000, log2_slots=16, n_io=1024,
                             mode=0, also_reverse=True),
    "toon1024": dict(desc="reduced config 2: Toon advection nx=1024 nt=250, 1024x1024 Jacobian, reverse sweeps",
                     scheme="toon", nx=1024, nt=250, mode=1),
    "lw100": dict(desc="BASELINE config 1: Lax-Wendroff nx=100 nt=2000, 100x100 Jacobian, forward sweeps",
                  scheme="lw", nx=100, nt=2000, mode=0),
    "synthetic1e7_fwd": dict(desc="BASELINE config 4 at 1/10 length: seeded random tape, 1e7 statements (mean 4 operands), "
                                  "2^20 slots, 8192x8192 Jacobian, forward sweeps", scheme="synthetic", n=10_000_000,
                             log2_slots=20, n_io=8192, mode=0, also_reverse=True),
    "synthetic1e8_fwd": dict(desc="BASELINE config 4, full size: seeded random tape, 1e8 statements (mean 4 operands), "
                                  "2^20 slots, 8192x8192 Jacobian, forward sweeps", scheme="synthetic", n=100_000_000,
                             log2_slots=20, n_io=8192, mode=0, also_reverse=True),
    "synthetic1e8_rev": dict(desc="BASELINE config 4, full size: seeded random tape, 1e8 statements (mean 4 operands), "
                                  "2^20 slots, 8192x8192 Jacobian, reverse sweeps", scheme="synthetic", n=100_000_000,
                             log2_slots=20, n_io=8192, mode=1),
    "synthetic1e7_rev": dict(desc="BASELINE config 4 at 1/10 length: seeded random tape, 1e7 statements (mean 4 operands), "
                                  "2^20 slots, 8192x8192 Jacobian, reverse sweeps", scheme="synthetic", n=10_000_000,
                             log2_slots=20, n_io=8192, mode=1),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks and throttle reasons DURING the timed region.  The nvidia-smi process is started BEFORE the
    warm-up steps (its start-up takes driver locks for a few hundred ms, which would otherwise land inside a short timed
    region) and only the samples that arrive between mark_begin() and stop() are used."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index
        self.t_begin = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def mark_begin(self):
        self.t_begin = time.perf_counter()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def stop(self):
        t_end = time.perf_counter()
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)          # one more sample for very short regions
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        t0 = self.t_begin if self.t_begin is not None else 0.0
        inside = [r for (t, r) in self.rows if t0 <= t <= t_end + 0.15] or [r for (_, r) in self.rows[-3:]]
        sm, mx, reasons = [], [], set()
        for r in inside:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_tape(w):
    pkg = importlib.import_module("adept-2_b200")
    if w["scheme"] == "synthetic":
        return pkg.tapegen.synthetic_tape(w["n"], w["log2_slots"], w["n_io"], seed=42)
    return pkg.tapegen.advection_tape(w["scheme"], w["nx"], w["nt"])


def alg_bytes_per_opdir(S, O, mode):
    """forward 8(1+S/O), reverse 16(1+S/O) bytes of gradient traffic per op*dir."""
    return (8.0 if mode == 0 else 16.0) * (1.0 + S / O)


# ---------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation (oracle/_ref) on the host cores
# ---------------------------------------------------------------------------------------------
def reference_rate(tape, mode, target_seconds=12.0, variant=None):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle as po   # checker / CPU baseline only

    variant = po.best_ref_variant() if variant is None else variant
    kind = "reference"
    O = tape["idx"].size
    if po.ref_available(variant):
        rt = po.RefTape(variant)
        P = rt.packet_size
        threads = host_cores()      # set_max_jacobian_threads(0) = omp_get_num_procs() threads (adept/Stack.cpp:101-104)
        rt.inject(tape["stmt"], tape["mult"], tape["idx"], tape["max_gradient"] - 1, tape["indep"], tape["dep"])

        def run(dirs):
            if mode == 0:
                rt.set_io(tape["indep"][:dirs], tape["dep"])
            else:
                rt.set_io(tape["indep"], tape["dep"][:dirs])
            rt.jacobian(mode, n_threads=0)
            return rt.last_seconds
    else:   # the restatement, if the compiled reference did not travel
        kind = "port"
        P, threads = 8, host_cores()

        def run(dirs):
            ind = tape["indep"][:dirs] if mode == 0 else tape["indep"]
            dep = tape["dep"] if mode == 0 else tape["dep"][:dirs]
            t0 = time.perf_counter()
            po.jacobian(tape["stmt"], tape["mult"], tape["idx"], tape["max_gradient"], ind, dep, mode, P=P, n_threads=0)
            return time.perf_counter() - t0
    D_all = tape["indep"].size if mode == 0 else tape["dep"].size
    d = min(D_all, threads * P)              # calibration: one block per thread
    t = run(d)
    dirs = int(min(D_all, max(threads * P, (target_seconds / max(t, 1e-6)) * d // (threads * P) * threads * P)))
    return dict(run=run, dirs=dirs, P=P, threads=threads, kind=kind, variant=variant or "sse2", O=O)


RADIANCE_PROFILE = [294.0, 290.0, 285.0, 279.0, 273.0, 267.0, 261.0, 255.0, 248.0, 242.0, 235.0, 229.0, 222.0, 216.0, 216.0, 216.0,
                    216.0, 216.0, 216.0, 217.0, 218.0, 219.0, 220.0, 222.0, 223.0, 224.0]   # test/test_radiances.cpp:76-80


def ensemble_state(n_members, seed=0):
    """BASELINE config 3: the mid-latitude summer profile plus seeded N(0, 2 K) perturbations."""
    rng = np.random.default_rng(seed)
    return np.ascontiguousarray(np.array(RADIANCE_PROFILE)[None, :] + 2.0 * rng.standard_normal((n_members, len(RADIANCE_PROFILE))))


def run_ensemble_bench(state, n_devices, threads=0, reps=3, chunks=4, workdir=None):
    """adept-2_b200/dropin/_build/ensemble_bench: real Adept recording of simulate_radiances for every member on all host
    cores (adept_cuda_ensemble.h) + adept_cuda_jacobian_batch over n_devices GPUs.  Returns (jacobians, info)."""
    import tempfile
    exe = os.path.join(ROOT, "adept-2_b200", "dropin", "_build", "ensemble_bench")
    if not os.path.exists(exe):
        return None, {"error": "ensemble_bench not built"}
    with tempfile.TemporaryDirectory(dir=workdir) as d:
        sp, op = os.path.join(d, "state.bin"), os.path.join(d, "out.bin")
        state.tofile(sp)
        r = subprocess.run([exe, sp, str(state.shape[0]), str(n_devices), op, str(threads), str(reps), str(chunks)],
                           capture_output=True, text=True, timeout=900)
        if r.returncode != 0 or not os.path.exists(op):
            return None, {"error": (r.stdout + r.stderr)[-500:]}
        raw = open(op, "rb").read()
    hdr = np.frombuffer(raw, np.int64, 8)
    tm = np.frombuffer(raw, np.float64, 4, 64)
    B, ni, nd = int(hdr[0]), int(hdr[1]), int(hdr[2])
    jac = np.frombuffer(raw, np.float64, B * ni * nd, 96).reshape(B, ni, nd)
    info = {"members": B, "n_indep": ni, "n_dep": nd, "devices": int(hdr[3]), "host_threads": int(hdr[4]), "reps": int(hdr[5]),
            "pinned": bool(hdr[7]), "record_ms": float(tm[0]), "jacobian_ms": float(tm[1]), "total_ms": float(tm[2]),
            "mean_total_ms": float(tm[3]), "stdout": r.stdout.strip()}
    return jac, info


def ensemble_workload(args, eng, pkg, po, W, K):
    """BASELINE config 3: Jacobians of the radiance model for an ensemble of 1e5 state vectors.  End to end on both
    sides: our arm records every member with real Adept expression templates on all host cores (one Stack per thread)
    and replays the packed ensemble on the GPU(s), members sharded over the devices; the CPU arm is the reference's
    own simulate_radiances() (record + Stack::jacobian per member) under OpenMP on the same cores."""
    B = 100_000
    n_dev = max(1, args.gpus)
    state = ensemble_state(B)
    jac, info = run_ensemble_bench(state, n_dev, reps=max(K, 1))
    O, D = 106, 2.0
    cpu = None
    ok = None
    if po.ref_available(po.best_ref_variant()) and hasattr(po._ref(po.best_ref_variant()), "ref_ensemble_radiances"):
        po.ref_ensemble_radiances(state[:2000], 0, po.best_ref_variant())          # warm-up
        best, want = None, None
        for _ in range(max(1, min(K, 3))):
            want, sec = po.ref_ensemble_radiances(state, 0, po.best_ref_variant())
            best = sec if best is None else min(best, sec)
        cpu = {"value": O * D * B / best, "unit": "op*dir/s", "members_per_s": B / best, "cores": host_cores(), "kind": "reference",
               "sample": f"all {B} members: the reference's own simulate_radiances() (record + Stack::jacobian) under "
                         f"#pragma omp parallel for, one Stack per call; {1e3 * best:.1f} ms"}
        if jac is not None:
            ok = bool(np.array_equal(jac.view(np.uint64), want.view(np.uint64)))   # EVERY member, bit for bit
    if jac is None:
        return {"metric": "ensemble Jacobians: tape-entry x direction per second (BASELINE config 3)", "value": None, "unit": "op*dir/s",
                "error": info.get("error"), "cpu_baseline": cpu}
    wall = info["total_ms"] / 1e3
    work = O * D * B
    return {"metric": "ensemble Jacobians: tape-entry x direction per second (BASELINE config 3)", "value": work / (info["jacobian_ms"] / 1e3),
            "unit": "op*dir/s", "members_per_s": B / (info["jacobian_ms"] / 1e3), "n_gpus": n_dev, "steps": info["reps"], "warmup": 1,
            "ms_per_step": info["jacobian_ms"], "higher_is_better": True, "scaling": "strong (members sharded over devices, no collective)",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "BASELINE config 3: test/simulate_radiances.cpp (56 statements / 106 operations / 30 slots per member), "
                                   "2x26 reverse Jacobian for an ensemble of 1e5 state vectors (standard profile + N(0, 2 K))",
                       "parallelism": f"members/{n_dev}", "host_threads": info["host_threads"], "pinned_buffers": info["pinned"]},
            "value_is": "adept_cuda_jacobian_batch with host (page-locked) buffers: H2D of the multipliers, sweeps, D2H of the Jacobians, "
                        "pipelined in chunks per device",
            "e2e": {"value": work / wall, "unit": "op*dir/s", "members_per_s": B / wall, "ms_per_step": info["total_ms"],
                    "h2d_bytes_per_step": O * B * 8, "d2h_bytes_per_step": 52 * B * 8,
                    "phases_ms": {"host_recording_ms": info["record_ms"], "jacobian_batch_ms": info["jacobian_ms"]},
                    "what": "recording of all members on the host cores (plain Adept API, one Stack per thread) + packing + the batch call"},
            "gpu_launches": None,
            "cpu_baseline": cpu,
            "parity_bit_exact_vs_reference_every_member": ok,
            "note": info["stdout"]}


def checkpoint_workload(args, po, K):
    """the reference's checkpointing example at its own size (test/test_checkpoint.cpp:52-56: NX=100,
    100 blocks of 100 Toon steps), recorded with real Adept templates by adept-2_b200/dropin/_build/chain_check:
    (a) the reference's loop through the drop-in Stack (every reverse() a GPU sweep with host gradient vectors),
    (b) the device-resident chain (adept_cuda_chain_*), against the CPU sweeps of the same block tapes."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import tempfile
    import pathlib
    from test_chain_helper import EXE, run_check, oracle_chain
    if not os.path.exists(EXE):
        return {"metric": "checkpoint chain", "value": None, "error": "chain_check not built"}
    nblocks, nt = 100, 100
    best = None
    with tempfile.TemporaryDirectory() as d:
        for _ in range(max(1, min(K, 3)) + 1):
            r = run_check(pathlib.Path(d), nblocks, nt, "gpu")
            if best is None or r["ms_chain"] < best["ms_chain"]:
                best = r
    t0 = time.perf_counter()
    want = oracle_chain(po, best["blocks"])
    cpu_ms = 1e3 * (time.perf_counter() - t0)
    ops = sum(b["mult"].size for b in best["blocks"])
    ok = bool(np.array_equal(best["dj_chain"].view(np.uint64), want.view(np.uint64)) and
              np.array_equal(best["dj_loop"].view(np.uint64), want.view(np.uint64)))
    return {"metric": "tape entries per second through a checkpoint chain (reverse sweeps, one per freshly recorded block)",
            "value": ops / (best["ms_chain"] / 1e3), "unit": "entry/s", "n_gpus": 1, "steps": nblocks, "warmup": 1,
            "ms_per_step": best["ms_chain"] / nblocks, "higher_is_better": True, "scaling": "replicas only", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"test/test_checkpoint.cpp at its own size: NX=100, {nblocks} blocks x {nt} Toon steps, "
                                   f"{ops // nblocks} operations per block; recording on the host (real Adept templates) inside the timed region"},
            "e2e": {"value": ops / (best["ms_chain"] / 1e3), "unit": "entry/s", "ms_per_step": best["ms_chain"] / nblocks,
                    "h2d_bytes_per_step": int(sum(b["stmt"].nbytes + b["mult"].nbytes + b["idx"].nbytes for b in best["blocks"]) / nblocks),
                    "d2h_bytes_per_step": 0,
                    "phases_ms": {"chain_total_ms": best["ms_chain"], "dropin_loop_total_ms": best["ms_loop"],
                                  "host_recording_ms_in_the_loop_variant": best["ms_record"], "cpu_sweeps_total_ms": cpu_ms},
                    "verdict": ("At the reference's own size (20 000-statement blocks) the CPU sweeps take %.1f ms in all; the GPU chain takes "
                                "%.0f ms including %.0f ms of host recording: per block the device-side schedule build and ~600 dependent launches "
                                "of a 300-step level sweep cost more than the 0.2 ms CPU sweep.  The chain is %.1fx faster than calling the drop-in "
                                "block by block, and it only pays off against the CPU for blocks of millions of statements."
                                % (cpu_ms, best["ms_chain"], best["ms_record"], best["ms_loop"] / max(best["ms_chain"], 1e-9)))},
            "cpu_baseline": {"value": ops / (cpu_ms / 1e3), "unit": "entry/s", "cores": 1, "kind": "port",
                             "sample": "the same block tapes swept by the oracle's compute_adjoint, hand-off in host memory (no recording time)"},
            "parity_bit_exact_vs_oracle": ok}


def side_workload(args, name):
    """BASELINE configs 5 and 3: not Jacobians of one tape, so they have their own small drivers.
    Printed in the same JSON shape; `value` is from CUDA events around the device work, `e2e` the whole
    C-ABI call with host buffers."""
    import torch
    pkg = importlib.import_module("adept-2_b200")
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle as po
    W, K = max(args.warmup, 3), max(args.steps, 1)
    eng = pkg.Engine([0])
    peak, peak_src = peaks()
    if name == "rosenbrock1e7":
        n = 10_000_000
        t = pkg.tapegen.rosenbrock_tape(n)
        S, O, G = t["stmt"].shape[0] - 1, t["idx"].size, t["max_gradient"]
        pin = {k: torch.from_numpy(np.ascontiguousarray(t[k])).pin_memory() for k in ("stmt", "mult", "idx")}
        host = {k: v.numpy() for k, v in pin.items()}
        eng.upload_tape(host["stmt"], host["mult"], host["idx"], G)
        g0 = np.zeros(G); g0[t["dep"][0]] = 1.0
        g_pin = torch.from_numpy(g0.copy()).pin_memory()
        for _ in range(W):
            out = eng.adjoint(g0)
        # value: tape + schedule resident, the gradient vector stays on the device (adept_cuda_adjoint_device)
        gd0 = torch.from_numpy(g0).cuda()
        gd = gd0.clone()
        eng.adjoint_device(gd.data_ptr())
        dev_wall = 0.0
        ev = 0.0
        for _ in range(K):
            gd.copy_(gd0)
            torch.cuda.synchronize()
            t0 = time.perf_counter(); eng.adjoint_device(gd.data_ptr()); dev_wall += time.perf_counter() - t0
            ev += eng.stat("replay_ms") / 1e3
        # e2e: what ONE L-BFGS sample costs: a NEW tape every time -- pinned H2D of the 1.5 GB tape,
        # device-side schedule construction (level analysis + forward stream + reverse target stream), the gradient
        # vector in and out through host memory, and the sweep
        wall = 0.0
        ph = {"upload_ms": 0.0, "analysis_ms": 0.0, "an_window_ms": 0.0, "an_pack_ms": 0.0, "an_forward_ms": 0.0,
              "an_target_ms": 0.0, "adjoint_call_ms": 0.0, "sweep_ms": 0.0}
        n_e2e = max(1, min(K, 3))
        for _ in range(n_e2e):
            gh = g_pin.numpy()
            gh[:] = g0
            t0 = time.perf_counter()
            eng.upload_tape(host["stmt"], host["mult"], host["idx"], G)
            t1 = time.perf_counter()
            rc = eng.L.adept_cuda_adjoint(eng.h, gh.ctypes.data, 1, 0)     # in place on the pinned vector
            t2 = time.perf_counter()
            assert rc == 0
            wall += t2 - t0
            for k in ("upload_ms", "analysis_ms", "an_window_ms", "an_pack_ms", "an_forward_ms", "an_target_ms"):
                ph[k] += (eng.stat(k, 0.0) or 0.0) / n_e2e
            ph["adjoint_call_ms"] += 1e3 * (t2 - t1) / n_e2e
            ph["sweep_ms"] += (eng.stat("replay_ms", 0.0) or 0.0) / n_e2e
        out = g_pin.numpy().copy()
        wall /= n_e2e
        assert np.array_equal(gd.cpu().numpy().view(np.uint64), out.view(np.uint64))
        want = po.adjoint(t["stmt"], t["mult"], t["idx"], g0)
        t0 = time.perf_counter(); po.adjoint(t["stmt"], t["mult"], t["idx"], g0); cpu_s = time.perf_counter() - t0
        rt = None
        if po.ref_available(po.best_ref_variant()):
            rt = po.RefTape(po.best_ref_variant()).inject(t["stmt"], t["mult"], t["idx"], G - 1, t["indep"], t["dep"])
            rt.sweep(True, g0); cpu_s = rt.last_seconds
        bytes_per_entry = 12 + 8 * S / O + 16 * (1 + S / O)         # 37.6 B for this tape
        ach = O * bytes_per_entry / (ev / K) / 1e9
        h2d = (S + 1) * 8 + O * 12 + G * 8
        line = {"metric": "tape entries per second, one compute_adjoint sweep (BASELINE config 5)", "value": O * K / dev_wall,
                "unit": "entry/s", "n_gpus": 1, "steps": K, "warmup": W, "ms_per_step": 1e3 * dev_wall / K,
                "ms_per_step_cuda_events": 1e3 * ev / K,
                "higher_is_better": True, "scaling": "replicas only", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "BASELINE config 5: RosenbrockN N=1e7, one level-scheduled compute_adjoint "
                                       "(4e7 statements / 1e8 operations / 3e7 slots; one statement with 4e7 operands)",
                           "steps_in_schedule": eng.stat("n_levels")},
                "e2e": {"value": O / wall, "unit": "entry/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": G * 8,
                        "ms_per_step": 1e3 * wall, "phases_ms": ph,
                        "what": "a NEW tape per sweep, as an L-BFGS sample has it: pinned H2D of the tape + schedule build + "
                                "gradient vector in/out + sweep",
                        "reference_sweep_ms": 1e3 * cpu_s, "beats_the_reference_sweep": bool(wall < cpu_s),
                        "verdict": ("the GPU path wins end to end" if wall < cpu_s else
                                    "the GPU path LOSES end to end on this config: one sweep per freshly recorded tape cannot "
                                    "pay for the upload and the dependency analysis; it wins only when a tape is swept more "
                                    "than once (value) or when the gradient stays on the device")},
                "gpu_launches": int(eng.stat("kernel_launches", 0)),
                "roofline": {"bound": "hbm", "kernel": "sweep1_snapshot_kernel + sweep1_target_kernel", "achieved": ach,
                             "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None, "peak_source": peak_src},
                "cpu_baseline": {"value": O / cpu_s, "unit": "entry/s", "cores": 1, "kind": "reference" if rt else "port",
                                 "sample": "the whole sweep, single-threaded by design (adept/Stack.cpp:139-164)"},
                "parity_bit_exact_vs_oracle": bool(np.array_equal(out.view(np.uint64), want.view(np.uint64)))}
    elif name == "checkpoint":
        line = checkpoint_workload(args, po, K)
    else:  # ensemble1e5
        line = ensemble_workload(args, eng, pkg, po, W, K)
    print(json.dumps(line))
    return 0


def shared_config(wl, S, O, G, D, mode, world, dirs_note=""):
    """The workload description, identical (keys and values) for our arm and for --impl reference."""
    return {"workload": wl["desc"] + dirs_note, "statements": int(S), "operations": int(O), "slots": int(G),
            "directions": int(D), "mode": "reverse" if mode else "forward", "parallelism": f"directions/{world}"}


def _host_cores():
    try:
        return len(os.sched_getaffinity(0))      # what omp_get_num_procs() reports
    except Exception:
        return os.cpu_count() or 1


# taken at import, before any OpenMP runtime is loaded: libgomp may bind the initial thread (OMP_PROC_BIND), after
# which the affinity mask of this process no longer shows the cores the reference's OpenMP loop will use
_HOST_CORES = _host_cores()


def host_cores():
    return _HOST_CORES


def reference_main(args, wl, mode, W, K):
    """--impl reference: the reference's own CPU implementation of the path (oracle/_ref: the unmodified
    Stack::jacobian_forward/_reverse under OpenMP) on all host cores, each step a bounded sample of the workload."""
    os.environ["OMP_NUM_THREADS"] = str(host_cores())      # torchrun exports OMP_NUM_THREADS=1
    tape = build_tape(wl)
    S, O, G = tape["stmt"].shape[0] - 1, tape["idx"].size, tape["max_gradient"]
    D_all = tape["dep"].size if mode else tape["indep"].size
    # the whole --steps K --warmup W run should end within a few minutes: ~150 s of sweeps in all
    per_step = min(8.0, max(0.5, 150.0 / (K + min(W, 3))))
    r = reference_rate(tape, mode, target_seconds=per_step)
    for _ in range(min(W, 3)):
        r["run"](min(r["dirs"], r["threads"] * r["P"]))
    ts = [r["run"](r["dirs"]) for _ in range(K)]
    t = sum(ts) / K
    value = O * r["dirs"] / t
    sample = (f"{r['dirs']} of {D_all} directions ({'rows' if mode else 'columns'}) of the same Jacobian per step; "
              f"rate is linear in blocks (adept/jacobian.cpp:146-147); build {r['variant']} P={r['P']} -O3 -ffp-contract=off; "
              f"OpenMP threads = {r['threads']} (all host cores)")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": K,
        "warmup": W, "ms_per_step": 1e3 * t, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": shared_config(wl, S, O, G, D_all, mode, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": r["threads"], "kind": r["kind"], "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="synthetic1e8_fwd", choices=sorted(WORKLOADS) + ["rosenbrock1e7", "ensemble1e5", "checkpoint"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--reverse-steps", type=int, default=2, help="timed steps of the other sweep direction (workloads that name it)")
    ap.add_argument("--time-kernels", type=int, default=0, help="cross-check: one extra step with every launch bracketed by events "
                                                                 "(single stream, PDL off: a different, slower execution mode)")
    ap.add_argument("--schedule", type=int, default=0)
    ap.add_argument("--pass-dirs", type=int, default=0)
    ap.add_argument("--window", type=int, default=0)
    ap.add_argument("--warps", type=int, default=8)
    ap.add_argument("--persistent", type=int, default=0)
    ap.add_argument("--chunk", type=int, default=0)
    ap.add_argument("--streams", type=int, default=0)
    ap.add_argument("--fused", type=int, default=1, help="reverse level sweep as one kernel per step (0: snapshot + gather)")
    ap.add_argument("--cbs", type=int, default=0, help="fused kernel: column blocks per CTA (0 = automatic)")
    ap.add_argument("--pdl", type=int, default=1, help="programmatic dependent launch between the per-step kernels")
    ap.add_argument("--tail", type=int, default=1, help="fused kernel: tiny steps ride along with the previous launch")
    ap.add_argument("--opt", action="append", default=[], help="extra engine option key=value (repeatable)")
    ap.add_argument("--dirs", type=int, default=0,
                    help="EXPERIMENT ONLY: sweep just the first N directions (what one rank of a multi-GPU run holds); "
                         "the line is then not the named workload and says so in config")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.workload in ("rosenbrock1e7", "ensemble1e5", "checkpoint"):
        return side_workload(args, args.workload) if rank == 0 else 0
    W = max(args.warmup, 3)
    K = max(args.steps, 1)
    wl = WORKLOADS[args.workload]
    mode = wl["mode"]

    # ------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        return reference_main(args, wl, mode, W, K) if rank == 0 else 0

    # ------------------------------------------------------------------ our arm
    import torch
    import torch.distributed as dist

    pkg = importlib.import_module("adept-2_b200")
    torch.cuda.set_device(local_rank)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), timeout=datetime.timedelta(minutes=30))
    eng = pkg.Engine([local_rank])
    if world > 1:
        ids = [eng.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        eng.join(ids[0], rank, world)
    eng.set_option("schedule", args.schedule)
    eng.set_option("pass_dirs", args.pass_dirs)
    eng.set_option("window", args.window)
    eng.set_option("warps_per_cta", args.warps)
    eng.set_option("persistent", args.persistent)
    eng.set_option("chunk_recs", args.chunk)
    eng.set_option("fused", args.fused)
    eng.set_option("cbs", args.cbs)
    eng.set_option("tail", args.tail)
    eng.set_option("pdl", args.pdl)
    if args.streams:
        eng.set_option("streams", args.streams)
    for kv in args.opt:
        k, v = kv.split("=")
        eng.set_option(k, int(v))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(vals):
        t = torch.tensor(vals, dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t]

    t_gen = time.perf_counter()
    tape = build_tape(wl) if rank == 0 else None
    t_gen = time.perf_counter() - t_gen
    meta = [None]
    if rank == 0:
        meta = [dict(S=int(tape["stmt"].shape[0] - 1), O=int(tape["idx"].size), G=int(tape["max_gradient"]),
                     indep=tape["indep"], dep=tape["dep"])]
    if world > 1:
        dist.broadcast_object_list(meta, src=0)
    m = meta[0]
    S, O, G = m["S"], m["O"], m["G"]
    indep_all, dep_all = m["indep"], m["dep"]

    # pinned host copies of the tape (what a Stack shim would hand over) and of the result
    if rank == 0:
        pin = {k: torch.from_numpy(np.ascontiguousarray(tape[k])).pin_memory() for k in ("stmt", "mult", "idx")}
        host = {k: v.numpy() for k, v in pin.items()}
        out_pin = torch.empty(indep_all.size * dep_all.size, dtype=torch.float64).pin_memory()

    def upload():
        if rank == 0:
            eng.upload_tape(host["stmt"], host["mult"], host["idx"], G)
        if world > 1:
            eng.share_tape(0)

    upload()
    out_dev = torch.empty(indep_all.size * dep_all.size, dtype=torch.float64, device="cuda") if rank == 0 else None
    out_ptr = out_dev.data_ptr() if rank == 0 else 0
    peak, peak_src = peaks()

    def measure(mode, n_warm, n_steps, with_clocks):
        """W untimed + K timed Jacobians of one sweep direction, tape and schedule resident.  Everything reported
        comes from the timed region itself: wall clock between barriers, and the engine's CUDA events around each
        step (replay_ms) and around the replay kernels of every pass (sweep_ms) on its own stream."""
        indep, dep = indep_all, dep_all
        if args.dirs > 0:   # experiment: a rank's share of the directions
            if mode == 0:
                indep = np.ascontiguousarray(indep[:args.dirs])
            else:
                dep = np.ascontiguousarray(dep[:args.dirs])
        N, M = indep.size, dep.size
        D = N if mode == 0 else M
        work = float(O) * D

        def step():
            eng.jacobian_device(mode, indep, dep, out_ptr)   # collective when world > 1
            return eng.stat("replay_ms", 0.0), eng.stat("sweep_ms", 0.0), eng.stat("sweep_launches", 0.0)

        sampler = ClockSampler(local_rank)
        if rank == 0 and with_clocks:
            sampler.start()
        for _ in range(n_warm):
            step()
        launches0 = eng.stat("kernel_launches", 0.0)
        barrier()
        sampler.mark_begin()
        t0 = time.perf_counter()
        ev_ms = sw_ms = sw_launches = 0.0
        for _ in range(n_steps):
            a, b, c = step()
            ev_ms += a
            sw_ms += b
            sw_launches += c
        barrier()
        t1 = time.perf_counter()
        clocks = sampler.stop() if (rank == 0 and with_clocks) else None
        launches = eng.stat("kernel_launches", 0.0) - launches0
        wall_s, ev_s, sw_s = allmax([t1 - t0, ev_ms / 1e3, sw_ms / 1e3])
        used = eng.stat("used_schedule", 0.0)
        used_fused = bool(eng.stat("used_fused", 0.0)) and used == 2 and mode == 1
        n_passes = max(1.0, eng.stat("n_passes", 1.0) or 1.0)
        kname = {(1, 2.0): "reverse_target_kernel<2>", (0, 2.0): "forward_level_kernel<2,sparse>", (1, 1.0): "replay_reverse_kernel",
                 (0, 1.0): "replay_forward_kernel", (1, 3.0): "smallg_kernel<reverse>", (0, 3.0): "smallg_kernel<forward>"}
        if used_fused:
            kname[(1, 2.0)] = "reverse_fused_kernel<2>"
        used_resident = used_fused and bool(eng.stat("used_fused_persistent", 0.0))
        if used_resident:
            kname[(1, 2.0)] = "reverse_fused_persistent_kernel<2>"
        if eng.stat("used_fwd_balanced", 0.0):
            kname[(0, 2.0)] = "forward_balanced_kernel<2,sparse>"
        bpo = alg_bytes_per_opdir(S, O, mode)
        # algorithmic bytes of this rank's share of the step; the two-kernel reverse form splits them
        # between the snapshot kernel (per-statement part) and the gather (per-operation part): the events bracket both
        step_bytes = bpo * work / world
        n_l = sw_launches / n_steps if n_steps else 0.0
        n_streams = max(1.0, eng.stat("sweep_streams", 1.0) or 1.0)
        roof = None
        if sw_s > 0 and n_l > 0:
            sweep_s = sw_s / n_steps                  # replay kernels of one Jacobian, max over ranks
            ach = step_bytes / sweep_s / 1e9
            roof = {"bound": "hbm", "kernel": kname.get((mode, used), "?"), "achieved": ach, "peak": peak, "unit": "GB/s",
                    "frac": ach / peak, "traffic": None, "peak_source": peak_src,
                    "launches_per_step": n_l, "algorithmic_bytes_per_launch": step_bytes / n_l,
                    "avg_launch_us": 1e6 * sweep_s * n_streams / n_l, "concurrent_streams": n_streams,
                    "sweep_ms_per_step": 1e3 * sweep_s,
                    "source": "CUDA events of the timed region itself (engine stream, PDL and streams as in production), "
                              "around the replay kernels of every pass; zero-fill, seeding and extraction are outside them "
                              "but inside ms_per_step",
                    "frac_of_whole_step": step_bytes / (wall_s / n_steps) / 1e9 / peak,
                    "note": "gradient working set per pass is G*K*8 B; when it fits the 126 MB L2 the fraction is an "
                            "effective bandwidth and may exceed DRAM traffic (SURVEY.md 8d caveat)"}
            try:   # DRAM traffic per launch from the committed ncu capture of this workload, when there is one
                key = args.workload if mode == wl["mode"] else args.workload.replace("_fwd", "_rev")
                tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(key)
                if tr and tr.get("kernel") == roof["kernel"]:
                    roof["traffic"] = tr["dram_bytes_per_launch"]
                    roof["traffic_sample"] = tr["launch"]
                    if tr.get("algorithmic_bytes_per_launch"):
                        roof["traffic_over_algorithmic"] = tr["dram_bytes_per_launch"] / tr["algorithmic_bytes_per_launch"]
            except Exception:
                pass
        engine_info = {"schedule": {1.0: "tape-ordered", 2.0: "level", 3.0: "level, shared-memory gradients"}.get(used, "?"),
                       "fused_reverse": used_fused, "resident_kernel": used_resident,
                       "resident_grid": eng.stat("persistent_grid", None) if used_resident else None,
                       "fused_split": eng.stat("fused_split", None) if used_fused else None,
                       "fused_sub": eng.stat("fused_sub", None) if used_fused else None,
                       "steps_in_schedule": eng.stat("n_levels", 0),
                       "pass_dirs": eng.stat("pass_dirs", 0), "n_passes": n_passes,
                       "host_issue_ms": eng.stat("host_issue_ms", None),
                       "l2": "inputs larger than L2: record stream %.0f MB + workspace %.0f MB per pass" %
                             (16e-6 * (O + S), 8e-6 * G * (2 if used_fused else 1) * min(D, eng.stat("pass_dirs", D) or D))}
        return dict(N=N, M=M, D=D, work=work, wall_s=wall_s, ev_s=ev_s, value=work * n_steps / wall_s, launches=launches,
                    clocks=clocks, roof=roof, engine=engine_info, indep=indep, dep=dep, used=used, used_fused=used_fused)

    main_m = measure(mode, W, K, True)
    N, M, D, work = main_m["N"], main_m["M"], main_m["D"], main_m["work"]
    indep, dep = main_m["indep"], main_m["dep"]
    h2d = (S + 1) * 8 + O * 12 + (N + M) * 4
    d2h = N * M * 8

    # cross-check only (NOT the source of roofline): every launch bracketed by events, one stream, PDL off
    xcheck = None
    if args.time_kernels:
        eng.set_option("time_kernels", 1)
        eng.jacobian_device(mode, indep, dep, out_ptr)
        xcheck = {"dominant_ms": eng.stat("dominant_ms", 0.0), "dominant_launches": eng.stat("dominant_launches", 0.0),
                  "ms_by_quarter_of_sweep": [eng.stat(f"dominant_q{q}_ms", None) for q in (1, 2, 3, 4)],
                  "note": "single stream, PDL off: slower than the timed region by construction"}
        eng.set_option("time_kernels", 0)

    # end to end through the C-ABI with host buffers: H2D of the tape (+ broadcast) + schedule build + replay + D2H
    barrier()
    e2e_s = None
    phases = {}
    if args.e2e_steps > 0:
        t0 = time.perf_counter()
        t_up = t_jac = 0.0
        each = []
        for _ in range(args.e2e_steps):
            ta = time.perf_counter()
            upload()
            tb = time.perf_counter()
            if rank == 0:
                eng.jacobian(mode, indep, dep, out=out_pin.numpy())
            else:
                eng.jacobian_device(mode, indep, dep, 0)
            tc = time.perf_counter()
            t_up += tb - ta
            t_jac += tc - tb
            each.append([round(1e3 * (tb - ta), 2), round(1e3 * (tc - tb), 2), round(eng.stat("replay_ms", 0.0), 2)])
        phases["each_step_upload_jacobian_replay_ms"] = each
        barrier()
        e2e_s = (time.perf_counter() - t0) / args.e2e_steps
        phases["upload_call_ms"] = 1e3 * t_up / args.e2e_steps
        phases["jacobian_call_ms"] = 1e3 * t_jac / args.e2e_steps
        for k in ("upload_ms", "bcast_ms", "analysis_ms", "an_validate_ms", "an_window_ms", "an_pack_ms", "an_forward_ms", "an_target_ms",
                  "an_fused_ms", "replay_ms", "d2h_ms", "jacobian_wall_ms", "upload_total_ms"):
            phases[k] = eng.stat(k, None)
        e2e_s = allmax([e2e_s])[0]

    # rows/columns for the parity spot check below, taken now (the next measurement reuses the output buffers)
    sub = np.linspace(0, D - 1, 8).astype(np.int32)
    got = None
    if rank == 0:
        if args.e2e_steps > 0:
            full = out_pin.numpy()[:N * M].reshape(N, M)
        else:
            full = out_dev[:N * M].reshape(N, M)
        got = full[sub, :] if mode == 0 else full[:, sub]
        got = np.ascontiguousarray(got.cpu().numpy() if hasattr(got, "cpu") else got)

    # the other sweep direction of the same tape (config 4 names both), a short timed region of its own
    other = None
    if wl.get("also_reverse") and args.reverse_steps > 0 and args.dirs == 0:
        om = measure(1 - mode, 1, args.reverse_steps, False)
        other = {"value": om["value"], "unit": UNIT, "steps": args.reverse_steps, "warmup": 1,
                 "ms_per_step": 1e3 * om["wall_s"] / args.reverse_steps, "roofline": om["roof"], "engine": om["engine"],
                 "algorithmic_bytes_per_opdir": alg_bytes_per_opdir(S, O, 1 - mode), "gpu_launches": int(om["launches"]),
                 "an_fused_ms": eng.stat("an_fused_ms", None)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # parity spot check inside the bench (checker only): a few directions against the oracle, bit for bit
    # (full-matrix and full-size parity live in tests/: test_gpu_parity.py, test_gpu_fullsize.py)
    parity = None
    try:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import pyoracle as po
        ind = indep[sub] if mode == 0 else indep
        dp = dep if mode == 0 else dep[sub]
        rc, want = po.jacobian(tape["stmt"], tape["mult"], tape["idx"], G, ind, dp, mode, n_threads=0)
        parity = bool(rc == 0 and np.array_equal(got.ravel().view(np.uint64), want.view(np.uint64)))
    except Exception as e:  # the bench number does not depend on the checker being present
        parity = f"not checked: {e}"

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        try:
            r = reference_rate(tape, mode, target_seconds=12.0)
            t = r["run"](r["dirs"])
            cpu = {"value": O * r["dirs"] / t, "unit": UNIT, "cores": r["threads"], "kind": r["kind"],
                   "sample": f"{r['dirs']} of {D} directions of the same Jacobian, {t:.1f} s, build {r['variant']} "
                             f"P={r['P']}, OpenMP Stack::jacobian_{'reverse' if mode else 'forward'}"}
        except Exception as e:
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": str(e)}

    line = {
        "metric": METRIC, "value": main_m["value"], "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": 1e3 * main_m["wall_s"] / K, "ms_per_step_cuda_events": 1e3 * main_m["ev_s"] / K,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": shared_config(wl, S, O, G, D, mode, world,
                                f" -- EXPERIMENT: first {args.dirs} directions only" if args.dirs > 0 else ""),
        "engine": main_m["engine"],
        "clocks": main_m["clocks"],
        "e2e": {"value": (work / e2e_s) if e2e_s else None, "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * e2e_s if e2e_s else None, "phases_ms": phases},
        "gpu_launches": int(main_m["launches"]),
        "roofline": main_m["roof"],
        "algorithmic_bytes_per_opdir": alg_bytes_per_opdir(S, O, mode),
        "reverse" if mode == 0 else "forward": other,
        "time_kernels_cross_check": xcheck,
        "cpu_baseline": cpu,
        "parity_bit_exact_vs_oracle": parity,
        "tape_generation_s": round(t_gen, 1),
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
