"""GPU (-m gpu): parity of the CUDA path, through the C-ABI, against the oracle.

Bar (BASELINE.json north_star): Jacobians and sweeps are BIT-IDENTICAL in double to the reference
built without FMA contraction -- for the tape-ordered kernels AND for the level-scheduled ones
(this engine's level schedule keeps every accumulation order, so no tolerance is needed; the
1e-10 the north star would allow for a reordered adjoint is not used).  NaNs compare as a class.
"""
import numpy as np
import pytest

from conftest import golden_names, load_golden, same_bits

pytestmark = pytest.mark.gpu

ORDERED, LEVELS, SMALLG = 1, 2, 3   # SMALLG: level schedule with the shared-memory-resident kernel allowed
UNFUSED = 4                         # level schedule, reverse sweep as snapshot + gather (two kernels per step)
                                    # instead of the fused one-kernel-per-step sweep over versioned rows (LEVELS)


def upload(engine, t, schedule, window=0):
    engine.set_option("smallg", 1 if schedule in (0, SMALLG) else 0)
    engine.set_option("fused", 0 if schedule == UNFUSED else 1)
    schedule = LEVELS if schedule in (SMALLG, UNFUSED) else schedule
    engine.set_option("schedule", schedule)
    engine.set_option("window", window)
    engine.set_option("pass_dirs", 0)
    engine.set_option("ref_block", 2)
    engine.upload_tape(t["stmt"], t["mult"], t["idx"], t["max_gradient"])


def check_jacobians(engine, po, t, what=""):
    for mode in (0, 1):
        got = engine.jacobian(mode, t["indep"], t["dep"])
        rc, want = po.jacobian(t["stmt"], t["mult"], t["idx"], t["max_gradient"], t["indep"], t["dep"], mode,
                               n_threads=0)
        assert rc == 0
        assert same_bits(got, want), f"{what} mode {mode}: {int(np.sum(got != want))} of {got.size} elements differ"


# ---- golden vectors (made by the real reference) --------------------------------------------------
@pytest.mark.parametrize("name", golden_names())
@pytest.mark.parametrize("schedule,window", [(ORDERED, 0), (LEVELS, 0), (LEVELS, 64), (LEVELS, 1000), (SMALLG, 0), (SMALLG, 64),
                                             (UNFUSED, 0), (UNFUSED, 64)])
def test_golden_jacobians(engine, name, schedule, window):
    g = load_golden(name)
    upload(engine, g, schedule, window)
    assert same_bits(engine.jacobian(0, g["indep"], g["dep"]), g["jac_fwd"]), "forward"
    assert same_bits(engine.jacobian(1, g["indep"], g["dep"]), g["jac_rev"]), "reverse"
    used = engine.stat("used_schedule")
    if schedule in (LEVELS, UNFUSED):
        assert used == LEVELS
        # tapes without operations have nothing to fuse; everything else must take the path the test names
        assert engine.stat("used_fused") == (1 if schedule == LEVELS and g["idx"].size else 0)
    else:
        assert used == schedule or (schedule == SMALLG and used in (LEVELS, SMALLG))   # rosenbrock_n400 is too wide for smem


@pytest.mark.parametrize("name", golden_names())
@pytest.mark.parametrize("schedule,window", [(ORDERED, 0), (LEVELS, 0), (LEVELS, 100)])
def test_golden_sweeps(engine, pkg, name, schedule, window):
    g = load_golden(name)
    upload(engine, g, schedule, window)
    flags = pkg.SWEEP_ORDERED if schedule == ORDERED else 0
    assert same_bits(engine.tangent_linear(g["g_tl_in"], True, flags), g["g_tl_out"]), "tangent linear"
    assert same_bits(engine.adjoint(g["g_ad_in"], True, flags), g["g_ad_out"]), "adjoint"


@pytest.mark.parametrize("name", ["toon_nx128_nt20", "synthetic_s3000", "adversarial_2", "rosenbrock_n400"])
@pytest.mark.parametrize("persistent", [0, 1])
def test_reverse_level_sweep_per_step_and_persistent(engine, name, persistent):
    """The level-scheduled reverse sweep exists as two launches per step and as one cooperative kernel
    with grid-wide barriers; both must reproduce the reference's reverse Jacobian bit for bit."""
    g = load_golden(name)
    engine.set_option("persistent", persistent)
    engine.set_option("smallg", 0)
    try:
        for window in (0, 200):
            upload(engine, g, LEVELS, window)
            engine.set_option("smallg", 0)
            assert same_bits(engine.jacobian(1, g["indep"], g["dep"]), g["jac_rev"])
            engine.set_option("pass_dirs", 64)
            assert same_bits(engine.jacobian(1, g["indep"], g["dep"]), g["jac_rev"])
    finally:
        engine.set_option("persistent", 0)
        engine.set_option("pass_dirs", 0)


@pytest.mark.parametrize("name", golden_names())
@pytest.mark.parametrize("balanced,cbs,streams", [(0, 0, 4), (1, 0, 4), (1, 3, 1), (1, 32, 2)])
def test_forward_level_sweep_both_cta_shapes(engine, name, balanced, cbs, streams):
    """The level-scheduled forward sweep exists as forward_level_kernel (warp bound to a column block, chunks of any
    length) and forward_balanced_kernel (active column blocks handed to the warps, compacted walk, one staged piece);
    the host picks per step.  Either way, any column-block grouping, one pass or several: the reference's bits."""
    g = load_golden(name)
    try:
        engine.set_option("fwd_balanced", balanced)
        engine.set_option("streams", streams)
        for window in (0, 80):
            upload(engine, g, LEVELS, window)
            engine.set_option("cbs", cbs)
            for pass_dirs in (0, 64):
                engine.set_option("pass_dirs", pass_dirs)
                assert same_bits(engine.jacobian(0, g["indep"], g["dep"]), g["jac_fwd"])
    finally:
        engine.set_option("fwd_balanced", 1)
        engine.set_option("streams", 4)
        engine.set_option("cbs", 0)
        engine.set_option("pass_dirs", 0)


@pytest.mark.parametrize("name", golden_names())
def test_forward_balanced_kernel_with_four_doubles_per_lane(engine, po, pkg, name):
    """Option fwd_v4: the balanced forward kernel with 128-direction column blocks (32-byte lane vectors).  Tapes with an
    over-long step keep two doubles per lane (both kernels of a sweep must agree on the block width)."""
    g = load_golden(name)
    try:
        engine.set_option("fwd_v4", 1)
        for window, cbs, pass_dirs in ((0, 0, 0), (80, 3, 0), (0, 0, 128), (80, 1, 64)):
            upload(engine, g, LEVELS, window)
            engine.set_option("cbs", cbs)
            engine.set_option("pass_dirs", pass_dirs)
            assert same_bits(engine.jacobian(0, g["indep"], g["dep"]), g["jac_fwd"]), (window, cbs, pass_dirs)
            assert same_bits(engine.jacobian(1, g["indep"], g["dep"]), g["jac_rev"])     # reverse is unaffected
        if name.startswith("toon") or name.startswith("synthetic"):
            engine.set_option("cbs", 0)
            engine.set_option("pass_dirs", 0)
            engine.jacobian(0, g["indep"], g["dep"])
            assert engine.stat("fwd_lane_doubles") == 4
    finally:
        engine.set_option("fwd_v4", 0)
        engine.set_option("cbs", 0)
        engine.set_option("pass_dirs", 0)


def test_fwd_v4_on_config4_shape_and_non_finite(engine, po, pkg):
    t = pkg.tapegen.synthetic_tape(120_000, 13, 200, seed=21)     # 200 directions: a ragged last 128-direction block
    try:
        engine.set_option("fwd_v4", 1)
        upload(engine, t, LEVELS)
        check_jacobians(engine, po, t, "fwd_v4 synthetic")
        assert engine.stat("fwd_lane_doubles") == 4
        t["mult"][::997] = np.inf                                   # non-sparse variant of the kernel
        upload(engine, t, LEVELS)
        check_jacobians(engine, po, t, "fwd_v4 synthetic, non-finite multipliers")
    finally:
        engine.set_option("fwd_v4", 0)


def broadcast_tape(n=700, fan=64, seed=1):
    """x0, x1 feed EVERY statement of a layer (y_i = a_i x0 + b_i x1), and a second layer mixes the y's:
    in the reverse sweep x0 and x1 each receive n contributions within one step -- one target group of more
    than 256 records, i.e. a chunk the kernels must stream in several pieces."""
    rng = np.random.default_rng(seed)
    stmt, mult, idx = [[-1, 0]], [], []
    for i in range(n):                                   # slots 2 .. n+1
        idx += [0, 1]; mult += list(rng.uniform(-1, 1, 2))
        stmt.append([2 + i, len(idx)])
    for k in range(n):                                   # slots n+2 .. 2n+1
        src = rng.choice(n, fan, replace=False)
        idx += list(2 + src); mult += list(rng.uniform(-1, 1, fan))
        stmt.append([2 + n + k, len(idx)])
    return dict(stmt=np.array(stmt, np.int32), mult=np.array(mult), idx=np.array(idx, np.int32), max_gradient=2 * n + 2,
                indep=np.array([0, 1], np.int32), dep=np.arange(2 + n, 2 + 2 * n, dtype=np.int32))


@pytest.mark.parametrize("cbs", [0, 32, 3])
def test_target_groups_longer_than_a_staged_piece(engine, po, cbs):
    t = broadcast_tape()
    try:
        for schedule in (LEVELS, UNFUSED, ORDERED):
            upload(engine, t, schedule)
            engine.set_option("cbs", cbs)                # fused kernel: column blocks per CTA (32 > 8 warps: two sets)
            check_jacobians(engine, po, t, f"broadcast schedule {schedule} cbs {cbs}")
            g0 = np.zeros(t["max_gradient"]); g0[t["dep"]] = np.linspace(-1, 1, t["dep"].size)
            flags = 0 if schedule != ORDERED else 1
            assert same_bits(engine.adjoint(g0, True, flags), po.adjoint(t["stmt"], t["mult"], t["idx"], g0))
    finally:
        engine.set_option("cbs", 0)


@pytest.mark.parametrize("name", ["toon_nx128_nt20", "adversarial_1", "rosenbrock_n400", "lw_nx100_nt30"])
@pytest.mark.parametrize("tail,streams", [(0, 1), (1, 1), (1, 4)])
def test_tiny_steps_ride_along_with_the_previous_launch(engine, name, tail, streams):
    """Fused reverse sweep: steps of a chunk or two are run by the last CTA of the previous step's launch
    (ticket counter + fence) instead of a launch of their own; with and without, on one and on four streams,
    the Jacobian is the reference's bit for bit."""
    g = load_golden(name)
    try:
        engine.set_option("tail", tail)
        engine.set_option("streams", streams)
        for window in (0, 50):
            upload(engine, g, LEVELS, window)
            for pass_dirs in (0, 64):
                engine.set_option("pass_dirs", pass_dirs)
                assert same_bits(engine.jacobian(1, g["indep"], g["dep"]), g["jac_rev"])
            assert engine.stat("used_fused") == 1
    finally:
        engine.set_option("tail", 1)
        engine.set_option("streams", 4)
        engine.set_option("pass_dirs", 0)


# ---- the device-side dependency analysis ------------------------------------------------------------
@pytest.mark.parametrize("window", [0, 64, 777, 4096])
def test_levels_match_the_sequential_rule(engine, po, pkg, window):
    # the last four have more slots than the shared-memory table holds: hash-table walk, four statements per round --
    # random slots (hardly any conflict), a stencil (every neighbour conflicts), and a 71996-operand statement
    for t in (pkg.tapegen.advection_tape("toon", 96, 40), pkg.tapegen.synthetic_tape(30000, 11, 64, seed=5),
              pkg.tapegen.rosenbrock_tape(3000), pkg.tapegen.rosenbrock_tape(18000),
              pkg.tapegen.synthetic_tape(150000, 16, 64, seed=9), pkg.tapegen.advection_tape("toon", 8000, 4),
              pkg.tapegen.advection_tape("lw", 7000, 3)):
        W = window or 65536
        steps, want = po.levels(t["stmt"], t["idx"], t["max_gradient"], 2, window=W, huge=65536, pack=1)
        for coop in (1, 0):                  # packing as one cooperative kernel, and as three launches per window
            engine.set_option("coop_pack", coop)
            upload(engine, t, LEVELS, window)
            got = engine.levels()
            assert engine.stat("n_levels") == steps
            assert np.array_equal(got, want)     # global step of every statement, windows packed
        engine.set_option("coop_pack", 1)


def chain_tape(n=20000, n_slots=40000):
    """A serial chain y_k = a_k*y_{k-1} + b_k*x (every statement depends on the one before it) through fresh slots:
    ONE window holds more local steps than the packing pass keeps in shared memory (4096), and with 40 000 slots the
    walk uses the hash-table variant."""
    rng = np.random.default_rng(11)
    stmt = np.empty((n + 1, 2), np.int32)
    stmt[0] = (-1, 0)
    idx = np.empty(2 * n, np.int32)
    idx[0::2] = np.r_[0, 1 + np.arange(n - 1)]       # y_{k-1} (x for the first)
    idx[1::2] = 0                                     # x
    stmt[1:, 0] = 1 + np.arange(n)
    stmt[1:, 1] = 2 * (1 + np.arange(n))
    return dict(stmt=stmt, mult=rng.uniform(0.9, 1.1, 2 * n), idx=idx, max_gradient=n_slots,
                indep=np.array([0], np.int32), dep=np.array([n, n // 2, 7], np.int32))


@pytest.mark.parametrize("coop", [1, 0])
def test_a_window_deeper_than_the_shared_step_table(engine, po, coop):
    t = chain_tape()
    steps, want = po.levels(t["stmt"], t["idx"], t["max_gradient"], 2, window=65536, huge=65536, pack=1)
    assert steps == 20000
    try:
        engine.set_option("coop_pack", coop)
        upload(engine, t, LEVELS)
        assert engine.stat("n_levels") == steps and np.array_equal(engine.levels(), want)
        check_jacobians(engine, po, t, f"serial chain, coop_pack {coop}")
        g0 = np.zeros(t["max_gradient"]); g0[t["dep"]] = [1.0, -2.0, 0.5]
        assert same_bits(engine.adjoint(g0), po.adjoint(t["stmt"], t["mult"], t["idx"], g0))
    finally:
        engine.set_option("coop_pack", 1)


# ---- adversarial little tapes: slot reuse, self reference, repeated operands, empty statements -----
@pytest.mark.parametrize("seed", range(6))
def test_random_tapes(engine, po, pkg, seed):
    rng = np.random.default_rng(100 + seed)
    t = pkg.tapegen.random_small_tape(rng, int(rng.integers(50, 6000)), int(rng.integers(8, 300)),
                                      n_indep=int(rng.integers(1, 70)) if seed else 1,
                                      n_dep=int(rng.integers(1, 7)))
    for schedule, window in ((ORDERED, 0), (LEVELS, 0), (LEVELS, 97), (SMALLG, 97), (UNFUSED, 97)):
        upload(engine, t, schedule, window)
        check_jacobians(engine, po, t, f"seed {seed} schedule {schedule} window {window}")
        engine.set_option("pass_dirs", 32)                       # several passes through the same stream
        check_jacobians(engine, po, t, f"seed {seed} schedule {schedule} window {window} passes")
        engine.set_option("pass_dirs", 0)
        g0 = rng.uniform(-1, 1, t["max_gradient"])
        flags = pkg.SWEEP_ORDERED if schedule == ORDERED else 0
        assert same_bits(engine.adjoint(g0, True, flags), po.adjoint(t["stmt"], t["mult"], t["idx"], g0))
        assert same_bits(engine.tangent_linear(g0, True, flags), po.tangent_linear(t["stmt"], t["mult"], t["idx"], g0))


def test_overwritten_independents_leave_no_stale_rows(engine, po, pkg):
    """in the fused reverse sweep a ZMARK only clears the activity byte of the row it opens;
    when no active contribution lands on that row nothing is stored and the memory still shows the version
    before last.  X is an independent AND an LHS (s1 X=2X; s2 X=const; s3 X=3X; s4 Y=X): the reference returns
    dY/dX = 0, an extraction that ignores the byte returns the stale 3."""
    from test_fused_stream_model import overwritten_independent_tape
    t = overwritten_independent_tape()
    for schedule in (LEVELS, UNFUSED, ORDERED, SMALLG):
        upload(engine, t, schedule)
        check_jacobians(engine, po, t, f"overwritten independent, schedule {schedule}")
    upload(engine, t, LEVELS)
    assert engine.jacobian(1, t["indep"], t["dep"])[0] == 0.0 and engine.stat("used_fused") == 1


@pytest.mark.parametrize("seed", range(8))
def test_random_tapes_whose_independents_are_overwritten(engine, po, pkg, seed):
    """LHS drawn from ALL slots (independents included), many x = constant statements: chains of versions
    with breaks, in every schedule, single and multiple passes."""
    rng = np.random.default_rng(700 + seed)
    t = pkg.tapegen.random_small_tape(rng, int(rng.integers(30, 3000)), int(rng.integers(12, 200)),
                                      n_indep=int(rng.integers(1, 12)), n_dep=int(rng.integers(1, 7)),
                                      zero_op_prob=0.3, lhs_any_prob=0.5)
    for schedule, window in ((LEVELS, 0), (LEVELS, 53), (UNFUSED, 53), (SMALLG, 53), (ORDERED, 0)):
        upload(engine, t, schedule, window)
        check_jacobians(engine, po, t, f"seed {seed} schedule {schedule} window {window}")
        engine.set_option("pass_dirs", 32)
        check_jacobians(engine, po, t, f"seed {seed} schedule {schedule} window {window} passes")
        engine.set_option("pass_dirs", 0)


def test_empty_and_tiny_tapes(engine, po):
    # only the sentinel: every Jacobian entry is d(slot)/d(slot)
    t = dict(stmt=np.array([[-1, 0]], np.int32), mult=np.zeros(0), idx=np.zeros(0, np.int32), max_gradient=4,
             indep=np.array([0, 1], np.int32), dep=np.array([1, 2], np.int32))
    upload(engine, t, 0)
    check_jacobians(engine, po, t, "sentinel only")
    # one statement without operands (x = constant) and one with
    t = dict(stmt=np.array([[-1, 0], [2, 0], [1, 1]], np.int32), mult=np.array([3.0]), idx=np.array([0], np.int32),
             max_gradient=3, indep=np.array([0], np.int32), dep=np.array([1, 2], np.int32))
    for schedule in (ORDERED, LEVELS):
        upload(engine, t, schedule)
        check_jacobians(engine, po, t, "tiny")


# ---- output layout: dep_offset / indep_offset exactly as adept/jacobian.cpp:215-220 ---------------
def test_offsets(engine, po, pkg):
    rng = np.random.default_rng(7)
    t = pkg.tapegen.random_small_tape(rng, 800, 60, n_indep=37, n_dep=5)
    n, m = 37, 5
    for schedule in (LEVELS, SMALLG):
        upload(engine, t, schedule, 50)
        for mode in (0, 1):
            for do, io in ((1, 0), (0, 1), (n, 1), (1, m + 2), (n + 3, 1), (2, 2 * m + 1), (0, 0)):
                size = 1 + (m - 1) * (do if do > 0 else n) + (n - 1) * (io if io > 0 else m)
                got = engine.jacobian(mode, t["indep"], t["dep"], out=np.full(size, np.nan), dep_offset=do, indep_offset=io)
                rc, want = po.jacobian(t["stmt"], t["mult"], t["idx"], t["max_gradient"], t["indep"], t["dep"], mode,
                                       dep_offset=do, indep_offset=io, out=np.full(size, np.nan))
                assert rc == 0 and same_bits(got, want), (schedule, mode, do, io)


def test_multiple_passes(engine, po, pkg):
    t = pkg.tapegen.advection_tape("lw", 100, 30)
    for schedule in (ORDERED, LEVELS, UNFUSED):
        upload(engine, t, schedule)
        engine.set_option("pass_dirs", 32)        # 100 directions -> 4 passes, the last one partial
        check_jacobians(engine, po, t, f"passes schedule {schedule}")
        assert engine.stat("n_passes") == 4
    engine.set_option("pass_dirs", 0)


# ---- error behaviour (reference: exceptions thrown before the sweep; here: status codes) -----------
def test_errors(pkg):
    eng = pkg.Engine([0])
    with pytest.raises(pkg.cuda_replay_error):                                   # no tape yet
        eng.jacobian(0, [0], [1])
    t = pkg.tapegen.radiance_tape()
    eng.upload_tape(t["stmt"], t["mult"], t["idx"], t["max_gradient"])
    with pytest.raises(pkg.dependents_or_independents_not_identified):          # adept/jacobian.cpp:208-210
        eng.jacobian(0, [], t["dep"])
    with pytest.raises(pkg.dependents_or_independents_not_identified):
        eng.jacobian(1, t["indep"], [])
    with pytest.raises(pkg.gradients_not_initialized):                          # adept/Stack.cpp:161-163
        eng.adjoint(np.zeros(t["max_gradient"]), initialized=False)
    with pytest.raises(pkg.gradients_not_initialized):
        eng.tangent_linear(np.zeros(t["max_gradient"]), initialized=False)
    with pytest.raises(pkg.gradient_out_of_range):                              # slot outside the gradient list
        eng.jacobian(0, [t["max_gradient"] + 5], t["dep"])
    bad = t["idx"].copy(); bad[3] = t["max_gradient"] + 100
    with pytest.raises(pkg.gradient_out_of_range):                              # malformed tape is refused, not replayed
        eng.upload_tape(t["stmt"], t["mult"], bad, t["max_gradient"])
    eng.close()


# ---- non-finite multipliers: the only place where the reference's block width P is visible --------
@pytest.mark.parametrize("P", [1, 2, 4, 8])
def test_non_finite_multipliers_follow_the_reference_block_rule(engine, po, pkg, P):
    rng = np.random.default_rng(3)
    t = pkg.tapegen.random_small_tape(rng, 500, 40, n_indep=6, n_dep=9)
    t["mult"][rng.integers(0, t["mult"].size, 12)] = np.inf
    t["mult"][rng.integers(0, t["mult"].size, 6)] = np.nan
    for schedule in (ORDERED, LEVELS, SMALLG, UNFUSED):
        upload(engine, t, schedule, 64)
        engine.set_option("ref_block", P)
        for mode in (0, 1):
            got = engine.jacobian(mode, t["indep"], t["dep"])
            rc, want = po.jacobian(t["stmt"], t["mult"], t["idx"], t["max_gradient"], t["indep"], t["dep"], mode, P=P)
            assert rc == 0 and same_bits(got, want), (schedule, mode, P)
    engine.set_option("ref_block", 2)


# ---- the reference's own test workloads at full size ----------------------------------------------------
def test_thread_safe_workload(engine, po, pkg):
    """test/test_thread_safe.cpp:36-113: Toon NX=128, 200 steps, 128x128 Jacobian by forward AND reverse
    sweeps; the reference's criterion is RMSD(fwd, rev) <= 1e-5; ours is bit equality with the
    reference itself (oracle/_ref when it travelled to this box, else the pinned restatement)."""
    if po.ref_available(""):
        rt = po.RefTape().record_advection("toon", 128, 200)
        t = rt.arrays()
        want = {0: rt.jacobian(0, n_threads=0), 1: rt.jacobian(1, n_threads=0)}
    else:
        t = pkg.tapegen.advection_tape("toon", 128, 200)
        want = {m: po.jacobian(t["stmt"], t["mult"], t["idx"], t["max_gradient"], t["indep"], t["dep"], m)[1]
                for m in (0, 1)}
    for schedule in (ORDERED, LEVELS, SMALLG, UNFUSED):
        upload(engine, t, schedule)
        jf = engine.jacobian(0, t["indep"], t["dep"])
        jr = engine.jacobian(1, t["indep"], t["dep"])
        assert same_bits(jf, want[0]) and same_bits(jr, want[1])
        assert engine.stat("used_schedule") == (LEVELS if schedule == UNFUSED else schedule)
        assert np.sqrt(np.mean((jf - jr) ** 2)) <= 1e-5


def test_autodiff_benchmark_workload(engine, po, pkg):
    """BASELINE config 1: Lax-Wendroff NX=100, nt=2000, full 100x100 Jacobian (398100 / 1384100 / 299)."""
    t = pkg.tapegen.advection_tape("lw", 100, 2000)
    assert (t["stmt"].shape[0] - 1, t["idx"].size, t["max_gradient"]) == (398100, 1384100, 299)
    upload(engine, t, 0)
    check_jacobians(engine, po, t, "lw 100x2000")


def test_rosenbrock_adjoint_with_a_long_statement(engine, po, pkg):
    """BASELINE config 5 shape (test/test_minimizer.cpp:39-49,103-115): one statement holds 4(N-1) operands."""
    t = pkg.tapegen.rosenbrock_tape(20000)
    for schedule in (ORDERED, LEVELS):
        upload(engine, t, schedule)
        g0 = np.zeros(t["max_gradient"]); g0[t["dep"][0]] = 1.0      # cost.set_gradient(1.0)
        flags = pkg.SWEEP_ORDERED if schedule == ORDERED else 0
        got = engine.adjoint(g0, True, flags)
        assert same_bits(got, po.adjoint(t["stmt"], t["mult"], t["idx"], g0))
        x = np.zeros(t["max_gradient"]); x[t["indep"]] = np.linspace(-1, 1, t["indep"].size)
        assert same_bits(engine.tangent_linear(x, True, flags), po.tangent_linear(t["stmt"], t["mult"], t["idx"], x))
    check_jacobians(engine, po, t, "rosenbrock jacobian (1 x N: reverse; forward has N directions)")


def test_synthetic_config4_shape(engine, po, pkg):
    """BASELINE config 4 at reduced length: 2e5 statements, 2^14 slots, 256x256; parity on the full matrix."""
    t = pkg.tapegen.synthetic_tape(200_000, 14, 256, seed=42)
    for schedule in (LEVELS, UNFUSED):
        upload(engine, t, schedule)
        check_jacobians(engine, po, t, "synthetic")
        assert engine.stat("used_schedule") == LEVELS
        assert engine.stat("used_fused") == (1 if schedule == LEVELS else 0)


# ---- size-independent properties at a size the oracle would not finish quickly ------------------------
def test_large_advection_properties(engine, po, pkg):
    """NX=1024, nt=300 (1.06e6 statements), too big for a quick oracle run: size-independent properties.
    Toon: forward == reverse to rounding (the reference's own criterion, test/test_thread_safe.cpp:86-113)
    and level schedule == tape order, bit for bit.  Lax-Wendroff (a stable scheme: the Toon recurrence
    amplifies rounding differences between two summation paths): linearity, J v from one tangent-linear
    sweep equals the forward Jacobian applied to v, and v^T J from one adjoint sweep likewise."""
    t = pkg.tapegen.advection_tape("toon", 1024, 300)
    upload(engine, t, LEVELS)
    jf = engine.jacobian(0, t["indep"], t["dep"]).reshape(1024, 1024)      # [indep, dep]
    jr = engine.jacobian(1, t["indep"], t["dep"]).reshape(1024, 1024)
    assert np.sqrt(np.mean((jf - jr) ** 2)) <= 1e-5
    upload(engine, t, ORDERED)
    sub = np.arange(0, 1024, 16).astype(np.int32)                           # 64 columns through the ordered kernel
    jo = engine.jacobian(0, t["indep"][sub], t["dep"]).reshape(64, 1024)
    assert same_bits(jo, jf[sub])
    jo = engine.jacobian(1, t["indep"], t["dep"][sub]).reshape(1024, 64)
    assert same_bits(jo, jr[:, sub])
    # ... and both against the ORACLE on a subset of directions (not only against this engine's other kernel)
    sub = np.arange(3, 1024, 64).astype(np.int32)
    rc, want = po.jacobian(t["stmt"], t["mult"], t["idx"], t["max_gradient"], t["indep"][sub], t["dep"], 0, n_threads=0)
    assert rc == 0 and same_bits(np.ascontiguousarray(jf[sub]), want)
    rc, want = po.jacobian(t["stmt"], t["mult"], t["idx"], t["max_gradient"], t["indep"], t["dep"][sub], 1, n_threads=0)
    assert rc == 0 and same_bits(np.ascontiguousarray(jr[:, sub]), want)

    t = pkg.tapegen.advection_tape("lw", 1024, 300)
    upload(engine, t, LEVELS)
    jf = engine.jacobian(0, t["indep"], t["dep"]).reshape(1024, 1024)
    jr = engine.jacobian(1, t["indep"], t["dep"]).reshape(1024, 1024)
    np.testing.assert_allclose(jf, jr, rtol=0, atol=1e-13)
    rng = np.random.default_rng(0)
    v = np.zeros(t["max_gradient"]); v[t["indep"]] = rng.uniform(-1, 1, 1024)
    np.testing.assert_allclose(engine.tangent_linear(v)[t["dep"]], v[t["indep"]] @ jf, rtol=0, atol=1e-12)
    u = np.zeros(t["max_gradient"]); u[t["dep"]] = rng.uniform(-1, 1, 1024)
    np.testing.assert_allclose(engine.adjoint(u)[t["indep"]], jr @ u[t["dep"]], rtol=0, atol=1e-12)


# ---- the Stack mirror: reads like the reference's test_adept / test_thread_safe -----------------------
def test_stack_api(pkg, po):
    g = load_golden("algorithm")
    s = pkg.Stack()
    s.load_tape(g["stmt"], g["mult"], g["idx"], g["max_gradient"])
    s.independent(g["indep"]); s.dependent(g["dep"])
    assert same_bits(s.jacobian(), g["jac_rev"])                  # 2 independents > 1 dependent: reverse
    assert same_bits(s.jacobian_forward(), g["jac_fwd"])
    J = s.jacobian_matrix()
    assert J.shape == (1, 2)
    s.set_gradients(0, g["max_gradient"], g["g_ad_in"])
    s.reverse()
    assert same_bits(s.get_gradients(0, g["max_gradient"]), g["g_ad_out"])
    s.clear_gradients()
    s.set_gradients(0, g["max_gradient"], g["g_tl_in"])
    s.forward()
    assert same_bits(s.get_gradients(0, g["max_gradient"]), g["g_tl_out"])
    # recording continues after a Jacobian call: the device copy must follow (ownership)
    s2 = pkg.Stack()
    x = s2.register_gradients(2); y = s2.register_gradient()
    s2.new_recording()
    s2.push_rhs(2.0, x); s2.push_rhs(3.0, x + 1); s2.push_lhs(y)
    s2.independent([x, x + 1]); s2.dependent(y)
    assert s2.jacobian().tolist() == [2.0, 3.0]
    s2.push_rhs(10.0, y); s2.push_lhs(y)                          # y *= 10
    assert s2.jacobian().tolist() == [20.0, 30.0]


def test_two_stacks_sharing_one_engine(pkg):
    """an Engine holds ONE tape; a Stack must notice that another Stack (or an ensemble call) on the
    same engine replaced it, and upload again instead of replaying the stranger's tape."""
    eng = pkg.Engine([0])
    a, b = pkg.Stack(engine=eng), pkg.Stack(engine=eng)
    for s, m in ((a, 2.0), (b, 5.0)):
        x = s.register_gradient(); y = s.register_gradient()
        s.new_recording()
        s.push_rhs(m, x); s.push_lhs(y)
        s.independent(x); s.dependent(y)
    assert a.jacobian().tolist() == [2.0]
    assert b.jacobian().tolist() == [5.0]
    assert a.jacobian().tolist() == [2.0]          # a's tape was replaced by b's in between
    t = pkg.tapegen.radiance_tape()
    eng.jacobian_batch(1, t["stmt"], t["idx"], t["max_gradient"], t["mult"][:, None] * np.ones((1, 3)), t["indep"], t["dep"])
    assert b.jacobian().tolist() == [5.0]          # ... and by the ensemble's structure
    eng.close()


# ---- ensembles: B tapes sharing one structure (BASELINE config 3: simulate_radiances x 1e5 members) ----
@pytest.mark.parametrize("B", [1, 37, 1000])
def test_ensemble_jacobians(engine, po, pkg, B):
    """test/simulate_radiances.cpp:52-146 records the same 54-statement structure for every state vector;
    members differ only in their multipliers ([op][member]).  Each member's 2x26 Jacobian must equal the
    oracle's for that member's tape, bit for bit, in both sweep directions."""
    t = pkg.tapegen.radiance_tape()
    rng = np.random.default_rng(B)
    mm = t["mult"][:, None] * (1.0 + 0.05 * rng.standard_normal((t["mult"].size, B)))
    for mode in (1, 0):
        got = engine.jacobian_batch(mode, t["stmt"], t["idx"], t["max_gradient"], mm, t["indep"], t["dep"])
        assert got.shape == (B, t["indep"].size, t["dep"].size)
        for b in ([0] if B == 1 else list(rng.integers(0, B, 12)) + [0, B - 1]):
            rc, want = po.jacobian(t["stmt"], np.ascontiguousarray(mm[:, b]), t["idx"], t["max_gradient"], t["indep"],
                                   t["dep"], mode)
            assert rc == 0 and same_bits(got[b], want), (mode, b)
    # an adversarial structure too (self references, repeated operands, empty statements), odd direction counts
    r = pkg.tapegen.random_small_tape(np.random.default_rng(9), 200, 30, n_indep=5, n_dep=3)
    mm = r["mult"][:, None] * rng.uniform(0.5, 1.5, (r["mult"].size, 9))
    for mode in (0, 1):
        got = engine.jacobian_batch(mode, r["stmt"], r["idx"], r["max_gradient"], mm, r["indep"], r["dep"])
        for b in range(9):
            rc, want = po.jacobian(r["stmt"], np.ascontiguousarray(mm[:, b]), r["idx"], r["max_gradient"], r["indep"],
                                   r["dep"], mode)
            assert rc == 0 and same_bits(got[b], want), (mode, b)


def test_device_resident_sweeps(engine, po, pkg):
    """adept_cuda_adjoint_device / _tangent_linear_device: in-place sweeps of a gradient vector that stays on the
    GPU (optimiser loops, checkpoint chains): same bits as the host-vector entry points and the oracle."""
    import torch
    t = pkg.tapegen.rosenbrock_tape(5000)
    for schedule in (ORDERED, LEVELS):
        upload(engine, t, schedule)
        flags = pkg.SWEEP_ORDERED if schedule == ORDERED else 0
        g0 = np.zeros(t["max_gradient"]); g0[t["dep"][0]] = 1.0
        d = torch.from_numpy(g0).cuda()
        engine.adjoint_device(d.data_ptr(), flags)
        assert same_bits(d.cpu().numpy(), po.adjoint(t["stmt"], t["mult"], t["idx"], g0))
        x = np.zeros(t["max_gradient"]); x[t["indep"]] = np.linspace(-1, 1, t["indep"].size)
        d = torch.from_numpy(x).cuda()
        engine.tangent_linear_device(d.data_ptr(), flags)
        engine.tangent_linear_device(d.data_ptr(), flags)      # twice on the same vector, as test_derivatives.cpp:25-29 does
        want = po.tangent_linear(t["stmt"], t["mult"], t["idx"], po.tangent_linear(t["stmt"], t["mult"], t["idx"], x))
        assert same_bits(d.cpu().numpy(), want)


# ---- tape compaction on upload -----------------------------------------------------------------
def prune_null(t):
    """What a reference built with -DADEPT_REMOVE_NULL_STATEMENTS records (include/adept/StackStorageOrig.h:46-65):
    push_rhs drops a pair whose multiplier is zero; statements stay."""
    keep = t["mult"] != 0.0
    pos = np.concatenate([[0], np.cumsum(keep)])
    stmt = t["stmt"].copy()
    stmt[:, 1] = pos[stmt[:, 1]]
    return dict(t, stmt=stmt.astype(np.int32), mult=np.ascontiguousarray(t["mult"][keep]), idx=np.ascontiguousarray(t["idx"][keep]))


@pytest.mark.parametrize("seed", range(3))
def test_remove_null_matches_the_reference_built_with_remove_null_statements(engine, po, pkg, seed):
    rng = np.random.default_rng(40 + seed)
    t = pkg.tapegen.random_small_tape(rng, 4000 + 900 * seed, 150, n_indep=9, n_dep=7, lhs_any_prob=0.1)
    m = t["mult"]
    m[rng.random(m.size) < 0.25] = 0.0
    m[rng.random(m.size) < 0.05] = -0.0
    m[rng.random(m.size) < 0.1] = 1.0
    m[rng.random(m.size) < 0.1] = -1.0
    if seed == 2:                         # 0 * inf: the one thing a dropped term could have contributed
        m[rng.integers(0, m.size, 8)] = np.inf
    p = prune_null(t)
    assert p["idx"].size < t["idx"].size
    try:
        for schedule, window in ((LEVELS, 0), (LEVELS, 97), (UNFUSED, 97), (SMALLG, 0), (ORDERED, 0)):
            engine.set_option("remove_null", 1)
            upload(engine, t, schedule, window)
            assert engine.stat("ops_removed") == t["idx"].size - p["idx"].size
            assert engine.stat("frac_mult_zero") == np.mean(m == 0.0)                 # Stack::fraction_multipliers_equal_to
            assert engine.stat("frac_mult_one") == np.mean(m == 1.0) and engine.stat("frac_mult_minus_one") == np.mean(m == -1.0)
            for mode in (0, 1):
                got = engine.jacobian(mode, t["indep"], t["dep"])
                rc, want = po.jacobian(p["stmt"], p["mult"], p["idx"], p["max_gradient"], p["indep"], p["dep"], mode)
                assert rc == 0 and same_bits(got, want), (schedule, window, mode)
            g0 = rng.uniform(-1, 1, t["max_gradient"])
            flags = pkg.SWEEP_ORDERED if schedule == ORDERED else 0
            assert same_bits(engine.adjoint(g0, True, flags), po.adjoint(p["stmt"], p["mult"], p["idx"], g0))
            assert same_bits(engine.tangent_linear(g0, True, flags), po.tangent_linear(p["stmt"], p["mult"], p["idx"], g0))
            # the default (no compaction) keeps every recorded term
            engine.set_option("remove_null", 0)
            upload(engine, t, schedule, window)
            assert engine.stat("ops_removed") == 0
            check_jacobians(engine, po, t, "recorded tape")
    finally:
        engine.set_option("remove_null", 0)


def test_overlapped_multiplier_upload_from_pinned_memory(engine, po, pkg):
    """adept_cuda_upload_tape sends page-locked multipliers on a stream of their own while the level analysis (which
    reads only statements and indices) runs; the first kernel that reads them waits.  Same bits, same statistics,
    with the overlap on and off, level-scheduled and tape-ordered, with and without compaction."""
    import torch
    t = pkg.tapegen.synthetic_tape(300_000, 15, 128, seed=3)
    t["mult"][::7] = 1.0
    pin = {k: torch.from_numpy(np.ascontiguousarray(t[k])).pin_memory() for k in ("stmt", "mult", "idx")}
    rc, want = po.jacobian(t["stmt"], t["mult"], t["idx"], t["max_gradient"], t["indep"], t["dep"], 0, n_threads=0)
    try:
        for overlap in (1, 0):
            for schedule in (LEVELS, ORDERED):
                engine.set_option("overlap_upload", overlap)
                engine.set_option("schedule", schedule)
                engine.set_option("smallg", 0)
                engine.upload_tape(pin["stmt"].numpy(), pin["mult"].numpy(), pin["idx"].numpy(), t["max_gradient"])
                assert engine.stat("upload_overlapped") == overlap
                assert engine.stat("frac_mult_one") == np.mean(t["mult"] == 1.0)
                assert same_bits(engine.jacobian(0, t["indep"], t["dep"]), want), (overlap, schedule)
        # pageable memory and compaction take the sequential path
        engine.set_option("overlap_upload", 1)
        engine.upload_tape(t["stmt"], t["mult"], t["idx"], t["max_gradient"])
        assert engine.stat("upload_overlapped") == 0
        engine.set_option("remove_null", 1)
        engine.upload_tape(pin["stmt"].numpy(), pin["mult"].numpy(), pin["idx"].numpy(), t["max_gradient"])
        assert engine.stat("upload_overlapped") == 0
        assert same_bits(engine.jacobian(0, t["indep"], t["dep"]), want)
    finally:
        engine.set_option("remove_null", 0)
        engine.set_option("overlap_upload", 1)
        engine.set_option("schedule", 0)
        engine.set_option("smallg", 1)


@pytest.mark.parametrize("chunk", [0, 64, 128, 200, 256])
def test_statements_around_the_staging_limit(engine, po, chunk):
    """forward_balanced_kernel stages a chunk in ONE 256-record piece; steps that hold a statement long enough for a
    chunk to exceed that go to forward_level_kernel (the host decides per step, and refuses the schedule if a chunk of
    a step it did not flag is longer: stat max_short_chunk_recs).  Statements of 90..190 operands sit right on that
    line for every chunk length."""
    rng = np.random.default_rng(77 + chunk)
    n_in, n_mid = 260, 420
    stmt, mult, idx = [[-1, 0]], [], []
    for k in range(n_mid):                       # layer 1: long statements over the inputs
        ops = rng.choice(n_in, int(rng.integers(90, 191)), replace=False)
        idx += list(ops); mult += list(rng.uniform(-0.2, 0.2, ops.size))
        stmt.append([n_in + k, len(idx)])
    for k in range(300):                         # layer 2: short and long statements mixed, over layer 1
        ops = rng.choice(n_mid, int(rng.choice([3, 5, 120, 127, 128, 129, 160])), replace=False)
        idx += list(n_in + ops); mult += list(rng.uniform(-0.2, 0.2, ops.size))
        stmt.append([n_in + n_mid + k, len(idx)])
    t = dict(stmt=np.array(stmt, np.int32), mult=np.array(mult), idx=np.array(idx, np.int32), max_gradient=n_in + n_mid + 300,
             indep=np.arange(0, n_in, 3).astype(np.int32), dep=(n_in + n_mid + np.arange(0, 300, 7)).astype(np.int32))
    try:
        engine.set_option("chunk_recs", chunk)
        for balanced in (1, 0):
            engine.set_option("fwd_balanced", balanced)
            upload(engine, t, LEVELS)
            assert engine.stat("max_short_chunk_recs") <= 256
            check_jacobians(engine, po, t, f"chunk {chunk} balanced {balanced}")
    finally:
        engine.set_option("chunk_recs", 0)
        engine.set_option("fwd_balanced", 1)
