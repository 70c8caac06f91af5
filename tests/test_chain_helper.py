"""The checkpoint-chain helper (adept-2_b200/dropin/adept_cuda_chain.h) on the reference's own
checkpointing example (test/test_checkpoint.cpp:136-225), recorded with real Adept expression templates and the
reference's Toon scheme by adept-2_b200/dropin/_build/chain_check.

CPU part: the dumped block tapes are well formed (sentinel, sizes, seed/output slot lists) and chain through the oracle.
GPU part (-m gpu): both the reference's own loop run through the drop-in Stack (every reverse() a GPU sweep) and
the device-resident chain equal the oracle's chain bit for bit.
The binary is built where /root/reference exists (dropin/Makefile) and travels; nothing here reads the reference tree."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, same_bits

EXE = os.path.join(ROOT, "adept-2_b200", "dropin", "_build", "chain_check")
needs_build = pytest.mark.skipif(not os.path.exists(EXE), reason="chain_check not built (needs /root/reference at build time)")


def run_check(tmp_path, nblocks, nt, mode):
    out = tmp_path / f"chain_{nblocks}_{nt}_{mode}.bin"
    r = subprocess.run([EXE, str(nblocks), str(nt), str(out), mode], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    raw = open(out, "rb").read()
    off = 0

    def take(dt, n):
        nonlocal off
        a = np.frombuffer(raw, dt, n, off)
        off += a.nbytes
        return a
    nb, nt_, nx, gpu = (int(x) for x in take(np.int64, 4))
    blocks = []
    for _ in range(nb):
        n_stmt, n_ops, G, n_seed, n_vals, n_out = (int(x) for x in take(np.int64, 6))
        b = dict(stmt=take(np.int32, 2 * n_stmt).reshape(-1, 2), mult=take(np.float64, n_ops), idx=take(np.int32, n_ops),
                 max_gradient=G, seed=take(np.int32, n_seed), seed_vals=take(np.float64, n_vals), out=take(np.int32, n_out))
        blocks.append(b)
    rec_us = int(take(np.int64, 1)[0])
    rc, us_loop, us_chain = (int(x) for x in take(np.int64, 3))
    dj_loop, dj_chain = take(np.float64, nx), take(np.float64, nx)
    assert off == len(raw)
    return dict(nx=nx, blocks=blocks, rc=rc, ms_loop=us_loop / 1e3, ms_chain=us_chain / 1e3, ms_record=rec_us / 1e3,
                dj_loop=dj_loop, dj_chain=dj_chain, stdout=r.stdout)


def oracle_chain(po, blocks):
    h = None
    for b in blocks:                                  # dumped last block first, as the sweep visits them
        g = np.zeros(b["max_gradient"])
        g[b["seed"]] = b["seed_vals"] if b["seed_vals"].size else h
        g = po.adjoint(b["stmt"], b["mult"], b["idx"], g)
        h = g[b["out"]].copy()
    return h


@needs_build
def test_recorded_blocks_are_well_formed_and_chain(po, tmp_path):
    d = run_check(tmp_path, 5, 8, "record")
    nx = d["nx"]
    assert nx == 100 and len(d["blocks"]) == 5
    for k, b in enumerate(d["blocks"]):
        assert b["stmt"][0].tolist() == [-1, 0] and b["stmt"][-1][1] == b["mult"].size
        assert b["out"].size == nx and b["seed"].size == (1 if k == 0 else nx)
    h = oracle_chain(po, d["blocks"])
    assert np.isfinite(h).all() and np.abs(h).max() > 0
    # the hand-off really chains: dropping the middle block changes the answer, and block k's seeds are block k-1's
    # outputs one-to-one (both are the NX field values)
    h2 = oracle_chain(po, d["blocks"][:2] + d["blocks"][3:])
    assert not np.array_equal(h, h2)


@needs_build
@pytest.mark.gpu
@pytest.mark.parametrize("nblocks,nt", [(1, 10), (7, 10), (20, 50)])
def test_zz_chain_and_dropin_loop_equal_the_oracle_chain(po, tmp_path, nblocks, nt):
    d = run_check(tmp_path, nblocks, nt, "gpu")
    assert d["rc"] == 0
    want = oracle_chain(po, d["blocks"])
    assert same_bits(d["dj_loop"], want), "the reference's loop through the drop-in Stack"
    assert same_bits(d["dj_chain"], want), "the device-resident chain"
    print(d["stdout"].strip(), f"(host recording {d['ms_record']:.2f} ms)")
