"""Synthetic derivative tapes of the shapes BASELINE.json names (bench.py and tests feed these
to the engine).  Every generator returns a dict

    stmt [n_stmt,2] int32 (sentinel {-1,0} first), mult [n_ops] f64, idx [n_ops] i32,
    max_gradient, indep [N] i32, dep [M] i32

i.e. exactly the arrays adept::Stack holds after a recording
(reference include/adept/StackStorageOrig.h:155-163, include/adept/Stack.h:782-800).

These are tape WRITERS, not replay code: they stand in for the reference's expression-template
recording (which stays on the host, in C++) so that the GPU box -- where /root/reference does
not exist -- can build the benchmark workloads.  tests/test_tapegen.py checks them against
tapes recorded by the real reference: identical structure, multipliers to rounding.
"""
from __future__ import annotations

import numpy as np

_GOLD = np.uint64(0x9E3779B97F4A7C15)


def _mix(z):
    """splitmix64 output function on a uint64 array (counter-based: state_i = seed + (i+1)*GOLD)."""
    z = z.copy()
    z ^= z >> np.uint64(30)
    z *= np.uint64(0xBF58476D1CE4E5B9)
    z ^= z >> np.uint64(27)
    z *= np.uint64(0x94D049BB133111EB)
    z ^= z >> np.uint64(31)
    return z


def _stream(seed, start, n):
    with np.errstate(over="ignore"):
        i = np.arange(start + 1, start + n + 1, dtype=np.uint64)
        return _mix(np.uint64(seed) + i * _GOLD)


def synthetic_tape(n_statements, log2_slots=20, n_io=8192, seed=42, chunk=1 << 22):
    """BASELINE config 4: seeded random tape.  S statements with 1 + (r mod 7)
    operands (mean 4); G = 2**log2_slots slots; slots [0,n_io) are the independents and are never
    written; the last n_io statements write the dependents [G-n_io, G); every other LHS is uniform
    in [n_io, G-n_io) (heavy slot reuse); operand slots are uniform in [0, G-n_io); multipliers are
    uniform in +-0.866 (unit gain).  PRNG: counter-based splitmix64."""
    S = int(n_statements)
    G = 1 << int(log2_slots)
    n_io = int(n_io)
    if S < n_io or G <= 2 * n_io:
        raise ValueError("need n_statements >= n_io and 2**log2_slots > 2*n_io")
    stmt = np.empty((S + 1, 2), np.int32)
    stmt[0] = (-1, 0)
    nops_all = np.empty(S, np.int64)
    for s0 in range(0, S, chunk):
        n = min(chunk, S - s0)
        r = _stream(seed, s0, n)
        nops_all[s0:s0 + n] = 1 + (r % np.uint64(7)).astype(np.int64)
        lhs = n_io + ((r >> np.uint64(8)) % np.uint64(G - 2 * n_io)).astype(np.int64)
        stmt[1 + s0:1 + s0 + n, 0] = lhs
    stmt[1 + S - n_io:, 0] = G - n_io + np.arange(n_io)
    ends = np.cumsum(nops_all)
    O = int(ends[-1])
    if O >= 2**31:
        raise ValueError("tape exceeds the reference's 32-bit operation index")
    stmt[1:, 1] = ends
    mult = np.empty(O, np.float64)
    idx = np.empty(O, np.int32)
    for o0 in range(0, O, chunk):
        n = min(chunk, O - o0)
        r = _stream(seed ^ 0x5DEECE66D, o0, n)
        idx[o0:o0 + n] = ((r >> np.uint64(20)) % np.uint64(G - n_io)).astype(np.int32)
        u = (_mix(r) >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53))
        mult[o0:o0 + n] = (2.0 * u - 1.0) * 0.866
    return dict(stmt=stmt, mult=mult, idx=idx, max_gradient=G + 1,
                indep=np.arange(n_io, dtype=np.int32), dep=(G - n_io + np.arange(n_io)).astype(np.int32))


def advection_q0(nx):
    """benchmark/autodiff_benchmark.cpp:294."""
    i = np.arange(nx, dtype=np.float64)
    return (0.5 + 0.5 * np.sin((i * 2.0 * np.pi) / (nx - 1.5))) + 1


def advection_tape(scheme, nx, nt, c=0.125, q0=None):
    """The tape Adept records for benchmark/advection_schemes.h:24-58 ('lw' = lax_wendroff,
    'toon'), set up as benchmark/differentiator.h:324-334: slots q_init = [0,nx), q = [nx,2nx),
    flux = [2nx, 3nx-1); max_gradient = 3nx-1.  S = nt(2nx-1)+nx statements; operations
    nt(7nx-8)+nx (lw) or nt(9nx-10)+nx (toon)  (SURVEY.md section 8)."""
    nx, nt = int(nx), int(nt)
    q = np.array(advection_q0(nx) if q0 is None else q0, dtype=np.float64)
    per_flux = 4 if scheme == "lw" else 6
    ops_step = per_flux * (nx - 1) + 3 * (nx - 2) + 2
    S = nt * (2 * nx - 1) + nx
    O = nt * ops_step + nx
    stmt = np.empty((S + 1, 2), np.int32)
    mult = np.empty(O, np.float64)
    idx = np.empty(O, np.int32)
    stmt[0] = (-1, 0)
    # q[i] = q_init[i]
    stmt[1:nx + 1, 0] = nx + np.arange(nx)
    stmt[1:nx + 1, 1] = 1 + np.arange(nx)
    mult[:nx] = 1.0
    idx[:nx] = np.arange(nx)
    # structure of one time step (identical every step)
    ia = nx + np.arange(nx - 1)          # slot of q[i]
    ib = ia + 1                          # slot of q[i+1]
    flux_idx = np.stack([ia, ib] * (per_flux // 2), axis=1).reshape(-1)
    iu = np.arange(1, nx - 1)
    upd_idx = np.stack([nx + iu, 2 * nx + iu - 1, 2 * nx + iu], axis=1).reshape(-1)
    step_idx = np.concatenate([flux_idx, upd_idx, [nx + nx - 2, nx + 1]]).astype(np.int32)
    step_lhs = np.concatenate([2 * nx + np.arange(nx - 1), nx + iu, [nx, 2 * nx - 1]]).astype(np.int32)
    step_nops = np.concatenate([np.full(nx - 1, per_flux), np.full(nx - 2, 3), [1, 1]])
    step_ends = np.cumsum(step_nops)
    upd_mult = np.tile(np.array([1.0, 1.0, -1.0]), nx - 2)
    n_s = 2 * nx - 1
    for j in range(nt):
        s0 = 1 + nx + j * n_s
        o0 = nx + j * ops_step
        stmt[s0:s0 + n_s, 0] = step_lhs
        stmt[s0:s0 + n_s, 1] = o0 + step_ends
        idx[o0:o0 + ops_step] = step_idx
        a, b = q[:-1], q[1:]
        if scheme == "lw":
            h = 0.5 * c
            fm = np.stack([np.full(nx - 1, h), np.full(nx - 1, h), np.full(nx - 1, h * c),
                           np.full(nx - 1, -(h * c))], axis=1)
            flux = h * (a + b + c * (a - b))
        else:
            r = a / b
            E = np.exp(c * np.log(r))
            A = E - 1.0
            Dn = a - b
            N = A * a * b
            f = N / Dn
            w = 1.0 / Dn
            t = w * b * a * E * c / r          # d f / d r  (through the exp/log chain)
            m1 = t / b
            m2 = -(t * r) / b
            m3 = w * b * A
            m4 = w * (A * a)
            m5 = -(f / Dn)
            fm = np.stack([m1, m2, m3, m4, m5, -m5], axis=1)
            flux = f
        nf = per_flux * (nx - 1)
        mult[o0:o0 + nf] = fm.reshape(-1)
        mult[o0 + nf:o0 + nf + 3 * (nx - 2)] = upd_mult
        mult[o0 + ops_step - 2:o0 + ops_step] = 1.0
        q[1:-1] += flux[:-1] - flux[1:]
        q[0] = q[nx - 2]
        q[nx - 1] = q[1]
    return dict(stmt=stmt, mult=mult, idx=idx, max_gradient=3 * nx - 1,
                indep=np.arange(nx, dtype=np.int32), dep=(nx + np.arange(nx)).astype(np.int32), q_final=q)


def rosenbrock_tape(n, x0=-3.0):
    """The tape of RosenbrockN (test/test_minimizer.cpp:39-49, :103-108): y = calc_y(x);
    cost = 0.5*sum(y*y).  Slots x = [0,n), y = [n, 3n-2), sum temporary 3n-2, cost 3n-1;
    S = 4(n-1)+2, O = 10(n-1)+1, max_gradient = 3n.  One statement has 4(n-1) operands."""
    n = int(n)
    x = np.full(n, float(x0))
    m = n - 1
    S = 4 * m + 2
    O = 10 * m + 1
    stmt = np.empty((S + 1, 2), np.int32)
    mult = np.empty(O, np.float64)
    idx = np.empty(O, np.int32)
    stmt[0] = (-1, 0)
    ix = np.arange(m)
    # y(2ix) = 10*(x(ix+1) - x(ix)*x(ix)) ; y(2ix+1) = 1 - x(ix)
    lhs1 = np.stack([n + 2 * ix, n + 2 * ix + 1], axis=1).reshape(-1)
    ends1 = (np.stack([4 * ix + 3, 4 * ix + 4], axis=1)).reshape(-1)
    stmt[1:2 * m + 1, 0] = lhs1
    stmt[1:2 * m + 1, 1] = ends1
    idx[:4 * m] = np.stack([ix + 1, ix, ix, ix], axis=1).reshape(-1)
    dx = -(10.0 * x[:-1])
    mult[:4 * m] = np.stack([np.full(m, 10.0), dx, dx, np.full(m, -1.0)], axis=1).reshape(-1)
    y = np.empty(2 * m)
    y[0::2] = 10.0 * (x[1:] - x[:-1] * x[:-1])
    y[1::2] = 1.0 - x[:-1]
    # y *= sqrt(2): one self-referencing statement per element
    k = np.arange(2 * m)
    stmt[2 * m + 1:4 * m + 1, 0] = n + k
    stmt[2 * m + 1:4 * m + 1, 1] = 4 * m + 1 + k
    idx[4 * m:6 * m] = n + k
    mult[4 * m:6 * m] = np.sqrt(2.0)
    y = y * np.sqrt(2.0)
    # sum(y*y): ONE statement, every element pushed twice with multiplier y
    stmt[4 * m + 1] = (3 * n - 2, 10 * m)
    idx[6 * m:10 * m] = np.repeat(n + k, 2)
    mult[6 * m:10 * m] = np.repeat(y, 2)
    stmt[4 * m + 2] = (3 * n - 1, 10 * m + 1)
    idx[10 * m] = 3 * n - 2
    mult[10 * m] = 0.5
    return dict(stmt=stmt, mult=mult, idx=idx, max_gradient=3 * n,
                indep=np.arange(n, dtype=np.int32), dep=np.array([3 * n - 1], np.int32))


def radiance_tape(surface_temperature=294.0, temperature=None):
    """The radiance model of test/simulate_radiances.cpp:22-49 as simulate_radiances records it
    (:84-128): slots st = 0, t = [1,n+1), r = [n+1,n+3), bt = n+3.  The model is linear, so the
    multipliers do not depend on the temperatures (they are kept as arguments for symmetry)."""
    if temperature is None:
        temperature = [290.0, 285.0, 279.0, 273.0, 267.0, 261.0, 255.0, 248.0, 242.0, 235.0, 229.0, 222.0,
                       216.0, 216.0, 216.0, 216.0, 216.0, 216.0, 217.0, 218.0, 219.0, 220.0, 222.0, 223.0, 224.0]
    n = len(temperature)
    wavelength = [0.006, 0.0061]
    mass_abs = [3.0e-5, 3.0e-3]
    dz = 1000.0
    kB, cl = 1.380648813e-23, 299792458.0
    dens = 1.2 * 0.21 * (16.0 / 29.0) * np.exp(-(np.arange(n) * dz) / 8000.0)
    st_rows, mult, idx = [(-1, 0)], [], []
    bt = n + 3
    for ch in range(2):
        em = 1.0 - np.exp(-dens * mass_abs[ch] * dz)
        idx.append(0); mult.append(1.0); st_rows.append((bt, len(idx)))          # bt = st
        for i in range(n):                                                         # bt = bt*(1-e) + e*t[i]
            idx += [bt, 1 + i]
            mult += [1.0 - em[i], em[i]]
            st_rows.append((bt, len(idx)))
        wl = wavelength[ch]
        idx.append(bt); mult.append(2.0 * cl * kB / (wl * wl * wl * wl)); st_rows.append((n + 1 + ch, len(idx)))
    return dict(stmt=np.array(st_rows, np.int32), mult=np.array(mult), idx=np.array(idx, np.int32),
                max_gradient=n + 4, indep=np.arange(n + 1, dtype=np.int32),
                dep=np.array([n + 1, n + 2], np.int32))


def random_small_tape(rng, n_statements, n_slots, max_ops=5, n_indep=3, n_dep=3, zero_op_prob=0.1,
                      self_ref_prob=0.2, dup_prob=0.2, lhs_any_prob=0.0):
    """Adversarial little tapes for parity tests: heavy slot reuse, statements that read their own
    LHS, repeated operands, zero-operand statements (``x = constant``).  With lhs_any_prob > 0 that share
    of the statements may write ANY slot, the independents included (an independent that is overwritten
    during the recording: its Jacobian column sees only what happened before the first overwrite)."""
    stmt = [(-1, 0)]
    mult, idx = [], []
    for _ in range(n_statements):
        lhs = int(rng.integers(n_indep, n_slots))
        if lhs_any_prob > 0.0 and rng.random() < lhs_any_prob:
            lhs = int(rng.integers(0, n_slots))
        if rng.random() >= zero_op_prob:
            k = int(rng.integers(1, max_ops + 1))
            for _ in range(k):
                if idx and rng.random() < dup_prob:
                    j = idx[-1]
                elif rng.random() < self_ref_prob:
                    j = lhs
                else:
                    j = int(rng.integers(0, n_slots))
                idx.append(j)
                mult.append(float(rng.uniform(-1.2, 1.2)))
        stmt.append((lhs, len(idx)))
    indep = np.arange(n_indep, dtype=np.int32)
    dep = rng.choice(np.arange(n_indep, n_slots), size=n_dep, replace=False).astype(np.int32)
    return dict(stmt=np.array(stmt, np.int32), mult=np.array(mult, np.float64), idx=np.array(idx, np.int32),
                max_gradient=n_slots + 1, indep=indep, dep=dep)
