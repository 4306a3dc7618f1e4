"""GPU parity tests: the CUDA path, called through the C ABI
(include/edmd_cuda.h), against the oracle on the same seeded inputs and against
the golden fixtures generated from the unmodified reference.

Gates (BASELINE.json north_star): partner / dir / ctype bit-exact; event times
bit-exact (bar: 1e-12 relative); g(r) integer pair counts exact (=> g within
1e-10); psi6 within 1e-10, neighbour counts exact."""
import numpy as np
import pytest

from helpers import (ANALYSIS_ATOL, GOLDEN_CASES, assert_boop_close, assert_events_equal,
                     cfg_of, load_golden, pcf_counts_from_g)

pytestmark = pytest.mark.gpu


def gpu_sweep(pkg, c, *, t=None, cells=None, mode=0, allow_overlap=False):
    t = c.get("t", 0.0) if t is None else t
    with pkg.EdmdCuda(c["n"], c["lx"], c["ly"]) as ctx:
        ctx.upload(c["x"], c["y"], c["vx"], c["vy"], c["rad"], cell_xy=cells, t=t)
        return ctx.predict_all(mode=mode, vr=c.get("vr"), allow_overlap=allow_overlap)


def oracle_sweep(oracle, c, *, t=None, cells=None, mode=0):
    t = c.get("t", 0.0) if t is None else t
    return oracle.predict_all(c["n"], c["lx"], c["ly"], t, c["x"], c["y"], c["vx"], c["vy"],
                              c["rad"], vr=c.get("vr"), cell_xy=cells, mode=mode)


# ---------------------------------------------------------------- golden ----
@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_sweep_matches_reference_golden(pkg, name):
    g = load_golden(name)
    c = cfg_of(g)
    grow = bool(g["grow"])
    got = gpu_sweep(pkg, c, mode=int(grow))
    assert tuple(got["overlap"]) == (-1, -1)
    assert_events_equal(got, g, exact_times=not grow, prefix="first_")
    if not grow:
        assert_events_equal(got, g, exact_times=True, prefix="re_")


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_free_fly_matches_reference_golden(pkg, name):
    g = load_golden(name)
    c = cfg_of(g)
    grow = bool(g["grow"])
    with pkg.EdmdCuda(c["n"], c["lx"], c["ly"]) as ctx:
        ctx.upload(c["x"], c["y"], c["vx"], c["vy"], c["rad"], t=c["t"])
        if grow:
            ctx.set_growth(c["vr"])
        ctx.free_fly(float(g["ff_t"]), mode=int(grow))
        s = ctx.download_state()
    assert np.array_equal(s["x"], g["ff_x"])
    assert np.array_equal(s["y"], g["ff_y"])
    assert np.array_equal(s["rad"], g["ff_rad"])
    assert np.array_equal(s["vx"], c["vx"]) and np.array_equal(s["vy"], c["vy"])


@pytest.mark.parametrize("name", [n for n in GOLDEN_CASES if "grow" not in n])
def test_analysis_matches_reference_golden(pkg, name):
    g = load_golden(name)
    c = cfg_of(g)
    with pkg.EdmdCuda(c["n"], c["lx"], c["ly"]) as ctx:
        ctx.upload(c["x"], c["y"], c["vx"], c["vy"], c["rad"], t=c["t"])
        b = ctx.boop_cutoff(2.5)
        p = ctx.pcf(float(g["pcf_dr"]), float(g["pcf_max_r"]))
    assert_boop_close(b, g, prefix="boop_")
    assert abs(b["mean_q6"] - g["boop_q6"].mean()) <= ANALYSIS_ATOL
    assert p["num_bins"] == len(g["pcf_g"])
    want = pcf_counts_from_g(g["pcf_g"], c["n"], c["lx"], c["ly"], float(g["pcf_dr"]))
    assert np.array_equal(p["counts"], want)
    assert np.abs(p["g_r"] - g["pcf_g"]).max() <= ANALYSIS_ATOL


# ------------------------------------------------- oracle, BASELINE sizes ----
SWEEP_CASES = [
    # (N, phi, seed, small_fraction, shuffle)   BASELINE.json configs
    (2000, 0.70, 1, 0.3, True),      # configs[0]: reference CLI defaults, bidisperse
    (100000, 0.72, 2, 0.0, True),    # configs[1]
    (1000000, 0.70, 3, 0.0, True),   # configs[2]
    (1000000, 0.85, 1, 0.0, True),   # configs[4]
    (1000000, 0.70, 2, 0.3, False),  # bidisperse, lattice order
    (4000000, 0.70, 3, 0.0, True),   # configs[3] (whole system on one GPU)
]


@pytest.mark.parametrize("n,phi,seed,sf,shuffle", SWEEP_CASES)
def test_sweep_matches_oracle(pkg, oracle, n, phi, seed, sf, shuffle):
    c = pkg.synth.lattice_config(n, phi, seed, small_fraction=sf, shuffle=shuffle)
    t = 12.5 * seed
    got = gpu_sweep(pkg, c, t=t)
    want = oracle_sweep(oracle, c, t=t)
    assert want["rc"] == 0
    assert_events_equal(got, want)
    # size-independent properties
    assert got["dir"].min() >= 1 and got["dir"].max() <= 4
    assert (got["t_cross"] >= t).all() and (got["t_coll"] >= t).all()
    has = got["t_coll"] < 1e25
    i = np.nonzero(has)[0]
    j = got["partner"][i]
    mutual = got["t_coll"][j] == got["t_coll"][i]
    assert np.array_equal(got["partner"][j[mutual]], i[mutual])
    assert (got["partner"][~has] == 0).all()
    assert (got["t_coll"][~has] == t + 1e26).all()


def test_sweep_is_deterministic_and_reentrant(pkg):
    c = pkg.synth.lattice_config(200000, 0.72, seed=5)
    with pkg.EdmdCuda(c["n"], c["lx"], c["ly"]) as ctx:
        ctx.upload(c["x"], c["y"], c["vx"], c["vy"], c["rad"], t=1.0)
        a = ctx.predict_all()
        b = ctx.predict_all()      # idempotent on resident state
        ctx.upload(c["x"], c["y"], c["vx"], c["vy"], c["rad"], t=1.0)
        d = ctx.predict_all()
    for k in ("t_cross", "dir", "t_coll", "partner", "ctype"):
        assert np.array_equal(a[k], b[k]) and np.array_equal(a[k], d[k])


def test_upload_without_radii_keeps_them(pkg, oracle):
    """A thermostat tick re-uploads positions and velocities only (rad = NULL)."""
    c = pkg.synth.lattice_config(30000, 0.70, seed=6, small_fraction=0.3)
    rng = np.random.default_rng(6)
    with pkg.EdmdCuda(c["n"], c["lx"], c["ly"]) as ctx:
        with pytest.raises(pkg.EdmdError):
            ctx.upload(c["x"], c["y"], c["vx"], c["vy"], None, t=0.0)   # nothing resident yet
        ctx.upload(c["x"], c["y"], c["vx"], c["vy"], c["rad"], t=0.0)
        ctx.predict_all()
        c2 = dict(c, vx=c["vx"] * 0.5 + 0.1 * rng.standard_normal(c["n"]), vy=c["vy"] * 0.9)
        ctx.upload(c2["x"], c2["y"], c2["vx"], c2["vy"], None, t=1.0)
        got = ctx.predict_all()
        s = ctx.download_state()
    assert np.array_equal(s["rad"], c["rad"])
    assert_events_equal(got, oracle_sweep(oracle, c2, t=1.0))


def test_growth_sweep_matches_oracle(pkg, oracle):
    c = pkg.synth.growth_config(100000, 0.70, seed=7)
    rng = np.random.default_rng(7)
    c["vr"] = c["vr"] * (0.5 + rng.random(c["n"]))
    c["rad"] = c["rad"] * (0.7 + 0.3 * rng.random(c["n"]))
    got = gpu_sweep(pkg, c, mode=1)
    want = oracle_sweep(oracle, c, mode=1)
    assert_events_equal(got, want)   # same source order => bit-exact vs the restatement


def test_host_cells_and_aos_upload(pkg, oracle):
    """Mid-run form: the host supplies cell[2]; a few particles sit on the
    'wrong' side of their boundary (as after a crossing event at the boundary)."""
    c = pkg.synth.lattice_config(50000, 0.70, seed=8)
    n = c["n"]
    cells = oracle.cells(n, c["lx"], c["ly"], c["x"], c["y"]).reshape(n, 2).copy()
    box = oracle.box(n, c["lx"], c["ly"])
    # particles within 0.03 of their right-hand boundary are moved exactly onto it but
    # keep the old cell (0.03 cannot create an overlap at this density)
    edge = (cells[:, 0] + 1) * box.csx
    idx = np.nonzero((edge - c["x"] < 0.03) & (cells[:, 0] + 1 < box.nx))[0]
    assert idx.size > 100
    c["x"] = c["x"].copy()
    c["x"][idx] = edge[idx]
    assert not np.array_equal(oracle.cells(n, c["lx"], c["ly"], c["x"], c["y"]).reshape(n, 2), cells)
    want = oracle_sweep(oracle, c, t=2.0, cells=cells)
    got = gpu_sweep(pkg, c, t=2.0, cells=cells)
    assert_events_equal(got, want)
    # the same through the array-of-records entry point (reference struct stride)
    rec = np.zeros(n, dtype=np.dtype({
        "names": ["rad", "x", "y", "vx", "vy", "cell"],
        "formats": ["f8", "f8", "f8", "f8", "f8", ("i4", 2)],
        "offsets": [0, 8, 16, 24, 32, 136], "itemsize": 144}))
    for k in ("rad", "x", "y", "vx", "vy"):
        rec[k] = c[k]
    rec["cell"] = cells
    with pkg.EdmdCuda(n, c["lx"], c["ly"]) as ctx:
        ctx.upload_aos(rec, t=2.0, with_cells=True)
        got2 = ctx.predict_all()
    assert_events_equal(got2, want)


# ------------------------------------------------- calendar ingest plan ----
@pytest.mark.parametrize("n,phi,seed,sf,t", [(20000, 0.70, 51, 0.0, 0.0), (20000, 0.72, 52, 0.3, 37.25),
                                             (3000, 0.30, 53, 0.0, 2.0)])
def test_calendar_plan_matches_sequential_inserts(pkg, n, phi, seed, sf, t):
    """The device's bucket / next / prev / head equal what 2N sequential
    addEventToQueue calls leave in an empty calendar (the oracle walks them)."""
    from oracle.oracle_py import calendar_plan_oracle
    c = pkg.synth.lattice_config(n, phi, seed, small_fraction=sf)
    n = c["n"]
    dt_paul, paul_n = 5.0 / n, n                       # boxConstantHelper, src/EDMD.c:707-708
    actual = (7 * n) // 11
    paul_time = t - 0.4 * dt_paul                      # mid-bucket, like a running calendar
    with pkg.EdmdCuda(n, c["lx"], c["ly"]) as ctx:
        ctx.upload(c["x"], c["y"], c["vx"], c["vy"], c["rad"], t=t)
        ev = ctx.predict_all()
        got = ctx.calendar_plan(paul_time, dt_paul, paul_n, actual)
    want = calendar_plan_oracle(ev["t_cross"], ev["t_coll"], paul_time, dt_paul, paul_n, actual)
    assert got["n_tree"] == want["n_tree"]
    for k in ("bucket", "head"):
        assert np.array_equal(got[k], want[k]), k
    listed = want["bucket"] >= 0                        # next / prev are undefined for BST events
    assert np.array_equal(got["next"][listed], want["next"][listed])
    assert np.array_equal(got["prev"][listed], want["prev"][listed])
    assert (want["bucket"] == paul_n).sum() > 0 or phi > 0.5   # dilute: "never" events fill the overflow list


def test_calendar_plan_declines_crowded_buckets(pkg):
    """Few, wide buckets: more than 128 events per list => EDMD_EPLAN, the host
    then inserts one by one."""
    c = pkg.synth.lattice_config(5000, 0.70, seed=54)
    with pkg.EdmdCuda(c["n"], c["lx"], c["ly"]) as ctx:
        ctx.upload(c["x"], c["y"], c["vx"], c["vy"], c["rad"], t=0.0)
        ctx.predict_all()
        got = ctx.calendar_plan(0.0, 0.25, 4, 1, allow_declined=True)
    assert got["rc"] == pkg.binding.EPLAN


# ------------------------------------------------------------ edge cases ----
def test_empty_and_single_particle(pkg):
    e = np.zeros(0)
    with pkg.EdmdCuda(0, 10.0, 10.0) as ctx:
        ctx.upload(e, e, e, e, e, t=0.0)
        o = ctx.predict_all()
        assert o["t_coll"].size == 0
        assert ctx.pcf(0.1, 5.0)["counts"].sum() == 0
    with pkg.EdmdCuda(1, 10.0, 10.0) as ctx:
        ctx.upload(np.array([1.0]), np.array([1.0]), np.array([0.5]), np.array([-0.25]),
                   np.array([1.0]), t=2.0)
        o = ctx.predict_all()
        assert o["partner"][0] == 0 and o["t_coll"][0] == 2.0 + 1e26 and o["ctype"][0] == 1
        assert o["dir"][0] == 2 and o["t_cross"][0] == 2.0 + (2.0 - 1.0) / 0.5
        b = ctx.boop_cutoff(2.5)
        assert b["neighbors"][0] == 0 and b["q6"][0] == 0.0 and b["q6_arg"][0] == 0.0


@pytest.mark.parametrize("lx,ly", [(2.5, 2.5), (4.5, 9.0), (5.0, 2.2), (6.1, 6.1), (7.9, 40.0)])
def test_tiny_boxes_with_repeated_cells(pkg, oracle, lx, ly):
    """nx or ny < 3: the reference's 3x3 scan visits the same cell more than
    once (and psi6 double counts); the device must do the same."""
    rng = np.random.default_rng(int(lx * 10 + ly))
    n = max(2, int(lx * ly / 12))
    x = rng.random(n) * lx
    y = rng.random(n) * ly
    vx = rng.standard_normal(n)
    vy = rng.standard_normal(n)
    rad = np.full(n, 0.05)   # tiny disks: no overlap worries
    c = dict(n=n, lx=lx, ly=ly, x=x, y=y, vx=vx, vy=vy, rad=rad)
    got = gpu_sweep(pkg, c, t=0.5)
    want = oracle_sweep(oracle, c, t=0.5)
    assert_events_equal(got, want)
    with pkg.EdmdCuda(n, lx, ly) as ctx:
        ctx.upload(x, y, vx, vy, rad, t=0.5)
        b = ctx.boop_cutoff(2.5)
    assert_boop_close(b, oracle.boop_cutoff(n, lx, ly, x, y, 2.5))


def test_exact_ties_follow_reference_scan_order(pkg, oracle):
    """Two candidates with identical pair times: the first in the reference's
    scan order (rows, columns, then descending id inside a cell) wins."""
    lx = ly = 20.0
    # particle 0 at rest in the middle; mirrored partners approach with equal speed
    x = np.array([11.0, 8.5, 13.5, 11.0, 11.0, 11.5])
    y = np.array([11.0, 11.0, 11.0, 8.5, 13.5, 11.5])
    vx = np.array([0.0, 1.0, -1.0, 0.0, 0.0, 0.0])
    vy = np.array([0.0, 0.0, 0.0, 1.0, -1.0, 0.0])
    rad = np.array([1.0, 1.0, 1.0, 1.0, 1.0, 0.01])
    for perm in (np.arange(6), np.array([5, 4, 3, 2, 1, 0]), np.array([2, 0, 4, 1, 5, 3])):
        c = dict(n=6, lx=lx, ly=ly, x=x[perm], y=y[perm], vx=vx[perm], vy=vy[perm], rad=rad[perm])
        got = gpu_sweep(pkg, c)
        want = oracle_sweep(oracle, c)
        assert_events_equal(got, want)


def test_nan_candidates_lose(pkg, oracle):
    """Identical velocities: b = 0, v2 = 0 => 0/0 = NaN, `dt > NaN` is false."""
    lx = ly = 16.0
    x = np.array([5.0, 7.5, 11.0]); y = np.array([5.0, 5.0, 5.5])
    vx = np.array([0.25, 0.25, -1.0]); vy = np.array([0.5, 0.5, 0.0])
    rad = np.ones(3)
    c = dict(n=3, lx=lx, ly=ly, x=x, y=y, vx=vx, vy=vy, rad=rad)
    got = gpu_sweep(pkg, c)
    want = oracle_sweep(oracle, c)
    assert_events_equal(got, want)
    assert not np.isnan(got["t_coll"]).any()


def test_overlap_is_reported_not_fatal(pkg, oracle):
    c = pkg.synth.lattice_config(5000, 0.70, seed=9)
    c["x"] = c["x"].copy(); c["y"] = c["y"].copy()
    c["vx"] = c["vx"].copy(); c["vy"] = c["vy"].copy()
    # drop particle 4000 onto particle 17 (approaching) and 4500 onto 99
    for a, b in ((4000, 17), (4500, 99)):
        c["x"][a] = c["x"][b] + 1.2 if c["x"][b] + 1.2 < c["lx"] else c["x"][b] - 1.2
        c["y"][a] = c["y"][b]
        s = np.sign(c["x"][a] - c["x"][b])
        c["vx"][a], c["vx"][b] = -s, s
        c["vy"][a] = c["vy"][b] = 0.0
    want = oracle_sweep(oracle, c)
    assert want["rc"] == 1
    got = gpu_sweep(pkg, c, allow_overlap=True)
    assert got["rc"] == pkg.binding.EOVERLAP
    assert tuple(got["overlap"]) == tuple(want["overlap"])
    assert_events_equal(got, want)   # outputs are still filled, like the oracle's


def test_bad_cell_id_is_an_error(pkg):
    c = pkg.synth.lattice_config(1000, 0.5, seed=10)
    with pkg.EdmdCuda(c["n"], c["lx"], c["ly"]) as ctx:
        x = c["x"].copy()
        x[3] = c["lx"] * 1.5
        with pytest.raises(pkg.EdmdError) as ei:
            ctx.upload(x, c["y"], c["vx"], c["vy"], c["rad"], t=0.0)
        assert ei.value.code == pkg.binding.ECELL
        with pytest.raises(pkg.EdmdError) as ei:
            ctx.predict_all()
        assert ei.value.code == pkg.binding.ESTATE


# ------------------------------------------------------------- analysis ----
def test_boop_ignores_nan_coordinates_and_counts_coincident_disks(pkg, oracle):
    """The psi6 kernel on the cell slots runs every candidate through the whole chain (no branch): a NaN /
    infinite coordinate (the reference's `r2 < r_c*r_c` is false for it, src/boop.c:84) must not reach any sum,
    and two disks at the same point see each other at atan2(0, 0) = 0, i.e. z = 1 (src/boop.c:86-89)."""
    c = pkg.synth.lattice_config(40000, 0.70, seed=77)
    n = c["n"]
    cells = oracle.cells(n, c["lx"], c["ly"], c["x"], c["y"]).reshape(n, 2).copy()
    x, y = c["x"].copy(), c["y"].copy()
    x[17], y[4321] = np.nan, np.inf
    x[1000], y[1000] = x[2000], y[2000]          # coincident pair, filed under the cell of 2000
    cells[1000] = cells[2000]
    with pkg.EdmdCuda(n, c["lx"], c["ly"]) as ctx:
        ctx.upload(x, y, c["vx"], c["vy"], c["rad"], cell_xy=cells, t=0.0)
        b = ctx.boop_cutoff(2.5)
    want = oracle.boop_cutoff(n, c["lx"], c["ly"], x, y, 2.5, cell_xy=cells)
    assert want["neighbors"][17] == 0 and want["neighbors"][4321] == 0
    assert np.isfinite(b["q6"]).all() and np.isfinite(b["q6_arg"]).all()
    assert_boop_close(b, want)
    # ... and against the unmodified reference itself where oracle/_ref exists (the same state through its own
    # computeBOOPCutoff; tests/test_reference_edges.py pins the restatement on it in the CPU suite)
    from oracle.oracle_py import Reference
    if Reference.available():
        ref = Reference()
        ref.setup(n, c["lx"], c["ly"], 0.0, x, y, c["vx"], c["vy"], c["rad"], cell_xy=cells)
        try:
            assert_boop_close(b, ref.boop_cutoff(2.5))
        finally:
            ref.teardown()
    # the same without the strays (the plain variant of the tile kernel away from the box edges)
    x[17], y[4321] = c["x"][17], c["y"][4321]
    with pkg.EdmdCuda(n, c["lx"], c["ly"]) as ctx:
        ctx.upload(x, y, c["vx"], c["vy"], c["rad"], cell_xy=cells, t=0.0)
        b = ctx.boop_cutoff(2.5)
    assert_boop_close(b, oracle.boop_cutoff(n, c["lx"], c["ly"], x, y, 2.5, cell_xy=cells))


@pytest.mark.parametrize("n,phi,seed,sf", [(100000, 0.72, 2, 0.0), (1000000, 0.70, 3, 0.0),
                                           (1000000, 0.85, 1, 0.3)])
def test_boop_matches_oracle(pkg, oracle, n, phi, seed, sf):
    c = pkg.synth.lattice_config(n, phi, seed, small_fraction=sf)
    with pkg.EdmdCuda(c["n"], c["lx"], c["ly"]) as ctx:
        ctx.upload(c["x"], c["y"], c["vx"], c["vy"], c["rad"], t=0.0)
        b = ctx.boop_cutoff(2.5)
    want = oracle.boop_cutoff(c["n"], c["lx"], c["ly"], c["x"], c["y"], 2.5)
    assert_boop_close(b, want)
    assert abs(b["mean_q6"] - want["q6"].mean()) <= ANALYSIS_ATOL
    # the reference's truncation (3x3 cells of width ~2 with r_c = 2.5) is reproduced:
    # at phi = 0.72 on a near-perfect lattice not every particle sees 6 neighbours
    assert b["neighbors"].max() <= 8


@pytest.mark.parametrize("n,phi,seed,dr,frac", [(30000, 0.70, 4, 0.1, 0.5), (50000, 0.72, 5, 0.1, 0.5),
                                                (20000, 0.85, 6, 2.0, 0.5), (20000, 0.70, 7, 0.03, 0.1)])
def test_pcf_counts_match_oracle(pkg, oracle, n, phi, seed, dr, frac):
    c = pkg.synth.lattice_config(n, phi, seed)
    max_r = min(c["lx"], c["ly"]) * frac
    with pkg.EdmdCuda(c["n"], c["lx"], c["ly"]) as ctx:
        ctx.upload(c["x"], c["y"], c["vx"], c["vy"], c["rad"], t=0.0)
        p = ctx.pcf(dr, max_r)
    want = oracle.pcf(c["n"], c["lx"], c["ly"], c["x"], c["y"], dr, max_r)
    assert p["num_bins"] == want["num_bins"]
    assert np.array_equal(p["counts"], want["counts"])
    assert np.abs(p["g_r"] - want["g_r"]).max() <= ANALYSIS_ATOL


@pytest.mark.parametrize("n,phi,seed,dr,frac", [(120000, 0.70, 81, 0.1, 0.5), (60000, 0.85, 82, 0.013, 0.3),
                                                (100000, 0.30, 83, 2.0, 0.75), (20000, 0.70, 84, 0.0011, 0.2)])
def test_pcf_sorted_tiles_equal_plain_kernel(pkg, n, phi, seed, dr, frac):
    """Large systems take the sorted-tile kernels: by default the bin of a pair is
    decided in FP32 under a rigorous error bound and undecided pairs are redone with
    the reference's FP64 operations (k_pcf_f32); option 2 certifies every bin in FP64
    (k_pcf_sorted; also what very fine bins fall back to, last case).  The plain
    kernel (IEEE sqrt + division per pair) must count the same integers."""
    c = pkg.synth.lattice_config(n, phi, seed)
    max_r = min(c["lx"], c["ly"]) * frac
    with pkg.EdmdCuda(c["n"], c["lx"], c["ly"]) as ctx:
        ctx.upload(c["x"], c["y"], c["vx"], c["vy"], c["rad"], t=0.0)
        a = ctx.pcf(dr, max_r)
        exact = ctx.stat(pkg.binding.STAT_PCF_EXACT_PAIRS)
        ctx.set_option(pkg.binding.OPT_PCF_LEGACY, 2)
        a64 = ctx.pcf(dr, max_r)
        exact64 = ctx.stat(pkg.binding.STAT_PCF_EXACT_PAIRS) - exact
        ctx.set_option(pkg.binding.OPT_PCF_LEGACY, 1)
        b = ctx.pcf(dr, max_r)
    assert np.array_equal(a["counts"], b["counts"])
    assert np.array_equal(a64["counts"], b["counts"])
    npairs = c["n"] * (c["n"] - 1) // 2
    assert exact < 0.02 * npairs, (exact, npairs)      # FP32 settles almost every pair
    assert exact64 < 0.02 * npairs, (exact64, npairs)


def test_hardware_rsqrt_stays_inside_the_budget_of_the_pcf_kernel(pkg):
    """k_pcf_f32 budgets 1.28e-7 (PTX: 2^-22.9) for the relative error of rsqrt.approx.ftz.f32; measured
    here over every float in [2^-100, 2^64)."""
    with pkg.EdmdCuda(16, 30.0, 30.0) as ctx:
        worst = ctx.selftest_rsqrt()
    assert 0.0 < worst <= 1.28e-7, worst


@pytest.mark.parametrize("dr,spacing,frac", [(0.1, 0.5, 0.5), (0.25, 0.75, 0.75), (0.05, 1.0, 0.35)])
def test_pcf_points_on_bin_edges(pkg, oracle, dr, spacing, frac):
    """Adversarial for the FP32 decision: a square lattice whose spacing is a multiple
    of dr puts a large share of all distances exactly on bin edges, where only the
    reference's own FP64 rounding decides the bin (1.5/0.1 is not 15)."""
    m = 128
    n = m * m
    ix, iy = np.meshgrid(np.arange(m), np.arange(m), indexing="ij")
    x, y = (ix.ravel() * spacing).astype(np.float64), (iy.ravel() * spacing).astype(np.float64)
    lx = ly = m * spacing
    perm = np.random.default_rng(5).permutation(n)
    x, y = x[perm], y[perm]
    z = np.zeros(n)
    max_r = frac * lx
    with pkg.EdmdCuda(n, lx, ly) as ctx:
        ctx.upload(x, y, z + 1.0, z, z + 0.1, t=0.0)
        p = ctx.pcf(dr, max_r)
        exact = ctx.stat(pkg.binding.STAT_PCF_EXACT_PAIRS)
        ctx.set_option(pkg.binding.OPT_PCF_LEGACY, 2)
        p64 = ctx.pcf(dr, max_r)
    want = oracle.pcf(n, lx, ly, x, y, dr, max_r)
    assert np.array_equal(p["counts"], want["counts"])
    assert np.array_equal(p64["counts"], want["counts"])
    assert exact > 0          # the edge pairs did go through the FP64 path


def test_pcf_groups_sharing_one_histogram(pkg):
    """Large histograms (N >= 2*10^6) leave room for one or two CTAs per SM only, so a CTA
    then holds 2 or 4 groups of 256 threads that share the histogram, each with its own tile,
    queue, plans and named barrier.  Forced here at a size the test can afford: same counts."""
    c = pkg.synth.lattice_config(150000, 0.70, 171)
    max_r = min(c["lx"], c["ly"]) / 2
    got = {}
    with pkg.EdmdCuda(c["n"], c["lx"], c["ly"]) as ctx:
        ctx.upload(c["x"], c["y"], c["vx"], c["vy"], c["rad"], t=0.0)
        for g in (1, 2, 4):
            ctx.set_option(pkg.binding.OPT_PCF_GROUPS, g)
            got[g] = ctx.pcf(0.1, max_r)["counts"]
        ctx.set_option(pkg.binding.OPT_PCF_GROUPS, 0)
        ctx.set_option(pkg.binding.OPT_PCF_LEGACY, 2)
        ref = ctx.pcf(0.1, max_r)["counts"]
    for g in (1, 2, 4):
        assert np.array_equal(got[g], ref), g
    assert int(ref.sum()) > 0


def test_pcf_with_stray_and_non_finite_coordinates(pkg, oracle):
    """edmd_cuda_pcf_device takes any device array.  A NaN coordinate (the reference's
    `r < max_r` is false for it: the pairs are dropped) and coordinates outside the box
    (the reference wraps once, src/EDMD.c:5896-5913) must come out as the reference's
    arithmetic has them: tiles holding such a particle leave the FP32 path entirely."""
    import torch
    c = pkg.synth.lattice_config(20000, 0.70, 161)
    x, y = c["x"].copy(), c["y"].copy()
    x[17] = np.nan
    x[4711] = c["lx"] + 3.0
    y[9000] = -2.0 * c["ly"] - 1.0
    x[15000], y[15000] = 40.0 * c["lx"], 0.5
    dr, max_r = 0.1, 0.5 * min(c["lx"], c["ly"])
    xy = torch.from_numpy(np.stack([x, y], axis=1).copy()).cuda()
    nb = int(max_r / dr)
    counts = torch.zeros(nb, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    with pkg.EdmdCuda(c["n"], c["lx"], c["ly"]) as ctx:
        assert ctx.pcf_device(xy.data_ptr(), c["n"], dr, max_r, 0, 1, counts.data_ptr()) == nb
        exact = ctx.stat(pkg.binding.STAT_PCF_EXACT_PAIRS)
    want = oracle.pcf(c["n"], c["lx"], c["ly"], x, y, dr, max_r)
    assert np.array_equal(counts.cpu().numpy().astype(np.uint64), want["counts"])
    # the tiles of the NaN and of the far-out particle went through FP64 whole (the two strays within
    # 2 L of the box stay on the FP32 path: its error bound covers them)
    assert exact > 256 * (c["n"] - 512)


def test_pcf_fp32_decision_at_full_size(pkg):
    """N = 10^6 (BASELINE configs[2]) at a range the test can afford twice: the default
    kernel against the FP64-certified one, and the pair checksum."""
    c = pkg.synth.lattice_config(1000000, 0.70, seed=12345)
    n = c["n"]
    max_r = 0.12 * min(c["lx"], c["ly"])
    with pkg.EdmdCuda(n, c["lx"], c["ly"]) as ctx:
        ctx.upload(c["x"], c["y"], c["vx"], c["vy"], c["rad"], t=0.0)
        a = ctx.pcf(0.1, max_r)
        ctx.set_option(pkg.binding.OPT_PCF_LEGACY, 2)
        b = ctx.pcf(0.1, max_r)
    assert np.array_equal(a["counts"], b["counts"])
    assert int(a["counts"].sum()) > 0


def test_pcf_counts_every_pair_once(pkg):
    """Checksum at full size: with max_r beyond the half-diagonal every one of
    the N(N-1)/2 pairs lands in exactly one bin."""
    c = pkg.synth.lattice_config(300000, 0.70, seed=12)
    n = c["n"]
    with pkg.EdmdCuda(n, c["lx"], c["ly"]) as ctx:
        ctx.upload(c["x"], c["y"], c["vx"], c["vy"], c["rad"], t=0.0)
        p = ctx.pcf(0.1, 0.75 * max(c["lx"], c["ly"]))
    assert int(p["counts"].sum()) == n * (n - 1) // 2
    # hard disks: nothing below contact
    assert p["counts"][: int(1.99 / 0.1)].sum() == 0


@pytest.mark.parametrize("n,phi,seed,dr,frac,nparts", [(60000, 0.70, 91, 0.1, 0.5, 3), (40000, 0.85, 92, 0.05, 0.25, 2),
                                                       (5000, 0.70, 93, 0.1, 0.5, 2)])
def test_pcf_rank_shares_sum_to_the_whole(pkg, oracle, n, phi, seed, dr, frac, nparts):
    """Multi-GPU split of g(r) (edmd_cuda_pcf_device): every rank sorts the
    all-gathered positions itself and takes the tile pairs w = rank (mod nparts), so
    the sort must cut the same tiles on every rank.  Each share here comes from its
    own context (its own atomic arrival order); the shares must add up to the
    oracle's counts."""
    import torch
    c = pkg.synth.lattice_config(n, phi, seed)
    max_r = min(c["lx"], c["ly"]) * frac
    xy = torch.from_numpy(np.stack([c["x"], c["y"]], axis=1).copy()).cuda()
    nb = int(max_r / dr)
    total = torch.zeros(nb, dtype=torch.int64, device="cuda")
    for part in range(nparts):
        share = torch.zeros(nb, dtype=torch.int64, device="cuda")
        torch.cuda.synchronize()
        with pkg.EdmdCuda(c["n"], c["lx"], c["ly"]) as ctx:
            assert ctx.pcf_device(xy.data_ptr(), c["n"], dr, max_r, part, nparts, share.data_ptr()) == nb
        assert int(share.sum()) > 0
        total += share
    want = oracle.pcf(c["n"], c["lx"], c["ly"], c["x"], c["y"], dr, max_r)
    assert np.array_equal(total.cpu().numpy().astype(np.uint64), want["counts"])


# -------------------------------------------------- weighted g(r) family ----
@pytest.mark.parametrize("name", ["weighted_n1500_phi072", "weighted_n1200_phi060_bidisperse"])
def test_weighted_pcf_family_matches_reference_golden(pkg, name):
    """Bragg-peak search + cos(k.r)-weighted g(r) against the reference's own
    outputs (find_max_structure_factor_bragg, calculate_bond_order_pcf)."""
    g = load_golden(name)
    n = int(g["n"])
    z = np.zeros(n)
    with pkg.EdmdCuda(n, float(g["lx"]), float(g["ly"])) as ctx:
        ctx.upload(g["x"], g["y"], z, z, g["rad"], t=0.0)
        peak = ctx.bragg_peak(float(g["expected_bragg"]))
        bo = ctx.pcf_bond_order(float(g["dr"]), float(g["max_r"]), g["k"])
        plain = ctx.pcf(float(g["dr"]), float(g["max_r"]))
    assert np.allclose(peak["k"], g["k"], rtol=0, atol=1e-12)
    assert bo["num_bins"] == len(g["g_r"])
    assert np.abs(bo["g_r"] - g["g_r"]).max() <= ANALYSIS_ATOL
    assert np.abs(bo["g6_r"] - g["g6_r"]).max() <= ANALYSIS_ATOL
    assert np.array_equal(bo["counts"], plain["counts"])


@pytest.mark.parametrize("n,phi,seed,dr", [(20000, 0.72, 71, 2.0), (12000, 0.85, 72, 0.25)])
def test_weighted_pcf_family_matches_oracle(pkg, oracle, n, phi, seed, dr):
    c = pkg.synth.lattice_config(n, phi, seed)
    n = c["n"]
    max_r = min(c["lx"], c["ly"]) / 2
    expected = float(np.sqrt(8 * np.pi * phi / np.sqrt(3)))
    with pkg.EdmdCuda(n, c["lx"], c["ly"]) as ctx:
        ctx.upload(c["x"], c["y"], c["vx"], c["vy"], c["rad"], t=0.0)
        peak = ctx.bragg_peak(expected)
        bo = ctx.pcf_bond_order(dr, max_r, peak["k"])
    want_peak = oracle.bragg_peak(n, c["lx"], c["ly"], c["x"], c["y"], expected)
    assert np.allclose(peak["k"], want_peak["k"], rtol=0, atol=1e-12)
    assert abs(peak["s_max"] - want_peak["s_max"]) <= 1e-10 * want_peak["s_max"]
    assert peak["s_max"] > 0.1 * n          # a (jittered) crystal: S at the Bragg peak is O(N)
    want = oracle.bond_order_pcf(n, c["lx"], c["ly"], c["x"], c["y"], dr, max_r, peak["k"])
    assert np.array_equal(bo["counts"], want["counts"])
    assert np.abs(bo["g_r"] - want["g_r"]).max() <= ANALYSIS_ATOL
    assert np.abs(bo["g6_r"] - want["g6_r"]).max() <= ANALYSIS_ATOL
    # size-independent properties: k = 0 weights every pair by 1; the average is bounded
    with pkg.EdmdCuda(n, c["lx"], c["ly"]) as ctx:
        ctx.upload(c["x"], c["y"], c["vx"], c["vy"], c["rad"], t=0.0)
        one = ctx.pcf_bond_order(dr, max_r, np.zeros(2))
    assert np.array_equal(one["g6_r"][one["counts"] > 0], np.ones((one["counts"] > 0).sum()))
    assert np.abs(bo["g6_r"]).max() <= 1.0


# ------------------------------------------- tiled kernel vs generic kernel ----
@pytest.mark.parametrize("n,phi,seed,sf", [(300000, 0.70, 21, 0.3), (300000, 0.85, 22, 0.0),
                                           (50000, 0.30, 23, 0.0)])
def test_tiled_and_generic_kernels_agree(pkg, n, phi, seed, sf):
    """The two-phase tiled kernel and the plain exact kernel are two
    implementations of the same function; the exact re-scan must be rare."""
    c = pkg.synth.lattice_config(n, phi, seed, small_fraction=sf)
    with pkg.EdmdCuda(c["n"], c["lx"], c["ly"]) as ctx:
        ctx.upload(c["x"], c["y"], c["vx"], c["vy"], c["rad"], t=4.0)
        a = ctx.predict_all()
        rescans = ctx.stat(pkg.binding.STAT_EXACT_RESCANS)
        ctx.set_option(pkg.binding.OPT_FORCE_GENERIC, 1)
        b = ctx.predict_all()
    for k in ("t_cross", "dir", "t_coll", "partner", "ctype"):
        assert np.array_equal(a[k], b[k]), k
    assert rescans < 0.02 * c["n"], rescans


# ------------------------------- lean (certified FP32 screening) vs full FP64 ----
@pytest.mark.parametrize("n,phi,seed,vscale", [(300000, 0.70, 41, 1.0), (300000, 0.85, 42, 1.0),
                                               (200000, 0.72, 43, 1e-6), (200000, 0.55, 44, 3e4),
                                               (50000, 0.30, 45, 1.0)])
def test_lean_and_full_sweeps_agree(pkg, n, phi, seed, vscale):
    """Monodisperse NORMAL-mode sweeps take the lean path: FP32 lower bounds pick
    the one pair that gets the exact FP64 evaluation.  Same bits as the full FP64
    kernel at any velocity scale; the exact re-scan must stay rare."""
    c = pkg.synth.lattice_config(n, phi, seed)
    with pkg.EdmdCuda(c["n"], c["lx"], c["ly"]) as ctx:
        ctx.upload(c["x"], c["y"], c["vx"] * vscale, c["vy"] * vscale, c["rad"], t=4.0)
        r0 = ctx.stat(pkg.binding.STAT_EXACT_RESCANS)
        a = ctx.predict_all()
        rescans = ctx.stat(pkg.binding.STAT_EXACT_RESCANS) - r0
        assert ctx.stat(pkg.binding.STAT_LEAN_SWEEPS) == 1
        ctx.set_option(pkg.binding.OPT_NO_LEAN, 1)
        b = ctx.predict_all()
        assert ctx.stat(pkg.binding.STAT_LEAN_SWEEPS) == 1
    for k in ("t_cross", "dir", "t_coll", "partner", "ctype"):
        assert np.array_equal(a[k], b[k]), k
    assert rescans < 0.02 * c["n"], rescans


@pytest.mark.parametrize("n,phi,seed,ratio", [(200000, 0.70, 49, 0.4), (120000, 0.80, 50, 0.85)])
def test_lean_sweep_with_two_radius_classes(pkg, oracle, n, phi, seed, ratio):
    """Bidisperse systems (the reference's default) stay on the lean path: the radius
    class rides in the last mantissa bit of the FP32 vy; three radii do not."""
    c = pkg.synth.lattice_config(n, phi, seed, small_fraction=0.3)
    rad = np.where(c["rad"] < 1.0, ratio, 1.0)
    c = dict(c, rad=rad)
    with pkg.EdmdCuda(c["n"], c["lx"], c["ly"]) as ctx:
        ctx.upload(c["x"], c["y"], c["vx"], c["vy"], c["rad"], t=2.0)
        a = ctx.predict_all()
        assert ctx.stat(pkg.binding.STAT_LEAN_SWEEPS) == 1
        rescans = ctx.stat(pkg.binding.STAT_EXACT_RESCANS)
        ctx.set_option(pkg.binding.OPT_NO_LEAN, 1)
        b = ctx.predict_all()
        # a third radius: not eligible any more
        rad3 = rad.copy()
        rad3[5::11] = 0.5 * ratio
        ctx.set_option(pkg.binding.OPT_NO_LEAN, 0)
        ctx.upload(c["x"], c["y"], c["vx"], c["vy"], rad3, t=2.0)
        d = ctx.predict_all()
        assert ctx.stat(pkg.binding.STAT_LEAN_SWEEPS) == 1
    for k in ("t_cross", "dir", "t_coll", "partner", "ctype"):
        assert np.array_equal(a[k], b[k]), k
    assert rescans < 0.02 * c["n"], rescans
    assert_events_equal(a, oracle_sweep(oracle, c, t=2.0))
    assert_events_equal(d, oracle_sweep(oracle, dict(c, rad=rad3), t=2.0))


def _adversarial_clusters(seed, n_clusters, vscale):
    """Isolated 3-particle clusters on a coarse grid, built so that the FP32 screening
    of the lean sweep is at its limits: two partners whose collision times with the
    centre particle differ by a relative 10^-k (k = 1 .. 12, incl. exact ties), partners
    almost in contact (gap 10^-k), grazing trajectories (det ~ 0 within 10^-k), and
    nearly parallel motion (b ~ 0)."""
    rng = np.random.default_rng(seed)
    pitch = 16.0     # centres >= 10 apart, clusters <= 3.3 in radius: no overlaps between clusters
    m = int(np.ceil(np.sqrt(n_clusters)))
    lx = ly = m * pitch
    x, y, vx, vy = [], [], [], []
    for q in range(n_clusters):
        cx, cy = (q % m + 0.5) * pitch + rng.uniform(-3, 3), (q // m + 0.5) * pitch + rng.uniform(-3, 3)
        kind = q % 4
        eps = 10.0 ** -rng.integers(1, 13)
        v0 = rng.uniform(0.2, 3.0) * vscale
        ang = rng.uniform(0, 2 * np.pi)
        ca, sa = np.cos(ang), np.sin(ang)
        pts = [(0.0, 0.0, 0.0, 0.0)]
        if kind == 0:      # two head-on partners, times t and t (1 + eps) (eps = 1e-12 ~ a tie)
            d1 = rng.uniform(2.05, 3.2)
            t1 = (d1 - 2.0) / v0
            d2 = 2.0 + v0 * t1 * (1.0 + (eps if rng.random() < 0.8 else 0.0))
            pts += [(d1, 0.0, -v0, 0.0), (-d2, 0.0, v0, 0.0)]
        elif kind == 1:    # a partner almost in contact, approaching slowly or fast
            pts += [(2.0 + eps, 0.0, -v0 * rng.choice([1e-3, 1.0]), 0.0),
                    (0.0, rng.uniform(2.3, 3.0), 0.0, -v0)]
        elif kind == 2:    # grazing: impact parameter 2 (1 -+ eps)
            b_imp = 2.0 * (1.0 + eps * rng.choice([-1.0, 1.0]))
            pts += [(rng.uniform(2.5, 3.3), b_imp if b_imp < 3.2 else 2.0, -v0, 0.0),
                    (-rng.uniform(2.2, 3.0), 0.0, 0.3 * v0, 0.0)]
        else:              # nearly parallel motion (b ~ 0) next to a real partner
            pts += [(0.0, 2.0 + rng.uniform(0.01, 0.8), v0 * eps, v0 * eps * rng.choice([-1.0, 1.0])),
                    (rng.uniform(2.2, 3.2), 0.0, -v0, 0.0)]
        for (dx, dy, ux, uy) in pts:
            x.append(cx + ca * dx - sa * dy)
            y.append(cy + sa * dx + ca * dy)
            vx.append(ca * ux - sa * uy)
            vy.append(sa * ux + ca * uy)
    n = len(x)
    return dict(n=n, lx=lx, ly=ly, x=np.array(x), y=np.array(y), vx=np.array(vx), vy=np.array(vy),
                rad=np.ones(n))


@pytest.mark.parametrize("n,phi,seed,sf", [(120000, 0.70, 141, 0.0), (120000, 0.72, 142, 0.3)])
def test_lean_sweep_with_radii_spread_inside_their_classes(pkg, oracle, n, phi, seed, sf):
    """The reference's growth phase leaves every disk at vr * t with its own rounding
    (src/EDMD.c:4992-5007; stopGrow :4740 does not reset it): its "monodisperse" and
    "bidisperse" runs hold many distinct radii within ~1e-15 of one or two values.  The
    lean path groups radii into classes (relative tolerance 1e-9), screens with the
    inflated class radius and reads every disk's own FP64 radius in the exact stage:
    same bits as the oracle, and the lean path is the one that ran."""
    c = pkg.synth.lattice_config(n, phi, seed, small_fraction=sf)
    ulps = np.random.default_rng(seed).integers(-4, 5, c["n"])
    rad = c["rad"] * (1.0 + ulps * 2.0 ** -52)
    assert len(np.unique(rad)) > 5
    with pkg.EdmdCuda(c["n"], c["lx"], c["ly"]) as ctx:
        ctx.upload(c["x"], c["y"], c["vx"], c["vy"], rad, t=1.5)
        got = ctx.predict_all()
        assert ctx.stat(pkg.binding.STAT_LEAN_SWEEPS) == 1
        assert ctx.stat(pkg.binding.STAT_LEAN_DECLINES) == 0
        rescans = ctx.stat(pkg.binding.STAT_EXACT_RESCANS)
    want = oracle.predict_all(c["n"], c["lx"], c["ly"], 1.5, c["x"], c["y"], c["vx"], c["vy"], rad)
    assert_events_equal(got, want)
    assert rescans < 0.02 * c["n"], rescans


def test_three_radius_classes_take_the_full_path(pkg, oracle):
    c = pkg.synth.lattice_config(30000, 0.65, 143, small_fraction=0.3)
    rad = c["rad"].copy()
    rad[::7] *= 0.9          # a third species
    with pkg.EdmdCuda(c["n"], c["lx"], c["ly"]) as ctx:
        ctx.upload(c["x"], c["y"], c["vx"], c["vy"], rad, t=0.0)
        got = ctx.predict_all()
        assert ctx.stat(pkg.binding.STAT_LEAN_SWEEPS) == 0
    want = oracle.predict_all(c["n"], c["lx"], c["ly"], 0.0, c["x"], c["y"], c["vx"], c["vy"], rad)
    assert_events_equal(got, want)


@pytest.mark.parametrize("k,seed", [(3, 151), (2, 152)])
def test_reference_grown_liquid_tiled(pkg, oracle, k, seed):
    """SURVEY.md 8d, second input family: a liquid grown and equilibrated by the reference
    itself (tests/golden/liquid_n10000_phi070.npz, made by make_liquid.py from the
    reference's own main()), tiled k x k with fresh velocities.  Sweep bit-exact on the
    lean path (its radii are spread inside one class), psi6 and g(r) against the oracle."""
    base = load_golden("liquid_n10000_phi070")
    c = pkg.synth.tiled_config(base, k, seed)
    n = c["n"]
    assert len(np.unique(c["rad"])) > 1
    max_r = min(c["lx"], c["ly"]) / 2
    with pkg.EdmdCuda(n, c["lx"], c["ly"]) as ctx:
        ctx.upload(c["x"], c["y"], c["vx"], c["vy"], c["rad"], t=0.0)
        got = ctx.predict_all()
        assert ctx.stat(pkg.binding.STAT_LEAN_SWEEPS) == 1
        b = ctx.boop_cutoff(2.5)
        p = ctx.pcf(0.1, max_r)
    want = oracle.predict_all(n, c["lx"], c["ly"], 0.0, c["x"], c["y"], c["vx"], c["vy"], c["rad"])
    assert_events_equal(got, want)
    assert_boop_close(b, oracle.boop_cutoff(n, c["lx"], c["ly"], c["x"], c["y"], 2.5))
    wp = oracle.pcf(n, c["lx"], c["ly"], c["x"], c["y"], 0.1, max_r)
    assert np.array_equal(p["counts"], wp["counts"])
    # a liquid: the first peak of g(r) sits at contact and g -> 1 at long range
    assert p["g_r"][20:23].max() > 3.0 and abs(p["g_r"][-200:].mean() - 1.0) < 0.02


@pytest.mark.parametrize("seed,vscale", [(61, 1.0), (62, 1.0), (63, 1e-5), (64, 2e3), (65, 1.0)])
def test_lean_certificate_on_adversarial_pairs(pkg, oracle, seed, vscale):
    """Near-ties, near-contacts, grazing and nearly parallel pairs at relative
    margins from 10^-1 down to 10^-12: whatever the FP32 bounds cannot decide must
    reach the exact re-scan, so the lean sweep still equals the oracle bit for bit."""
    c = _adversarial_clusters(seed, 20000, vscale)
    if seed == 65:   # two radius classes: isolated small disks between the clusters
        rng = np.random.default_rng(seed)
        m = int(np.ceil(np.sqrt(20000)))
        gx, gy = np.meshgrid(np.arange(m), np.arange(m))
        ex, ey = (gx.ravel() * 16.0 + 0.3)[:8000], (gy.ravel() * 16.0 + 0.3)[:8000]   # cluster-free corners
        c = dict(n=c["n"] + 8000, lx=c["lx"], ly=c["ly"], x=np.concatenate([c["x"], ex]),
                 y=np.concatenate([c["y"], ey]), vx=np.concatenate([c["vx"], rng.standard_normal(8000)]),
                 vy=np.concatenate([c["vy"], rng.standard_normal(8000)]),
                 rad=np.concatenate([c["rad"], np.full(8000, 0.25)]))
    with pkg.EdmdCuda(c["n"], c["lx"], c["ly"]) as ctx:
        ctx.upload(c["x"], c["y"], c["vx"], c["vy"], c["rad"], t=0.5)
        r0 = ctx.stat(pkg.binding.STAT_EXACT_RESCANS)
        got = ctx.predict_all(allow_overlap=True)
        rescans = ctx.stat(pkg.binding.STAT_EXACT_RESCANS) - r0
        assert ctx.stat(pkg.binding.STAT_LEAN_SWEEPS) == 1       # it did run on the lean path
    want = oracle_sweep(oracle, c, t=0.5)
    assert tuple(got["overlap"]) == tuple(want["overlap"])
    assert_events_equal(got, want)
    assert 0 < rescans < 0.6 * c["n"], rescans     # the hard cases do take the exact path, the rest do not


def test_lean_handles_heavy_velocity_tails_and_rest(pkg, oracle):
    """A few very fast particles set the FP32 velocity scale (loose bounds => more
    re-scans, same answers); particles at rest never collide by themselves."""
    c = pkg.synth.lattice_config(60000, 0.70, seed=46)
    vx, vy = c["vx"].copy(), c["vy"].copy()
    vx[::1000] *= 500.0
    vy[::1000] *= -300.0
    vx[1::7] = 0.0
    vy[1::7] = 0.0
    c["vx"], c["vy"] = vx, vy
    got = gpu_sweep(pkg, c, t=3.0)
    want = oracle_sweep(oracle, c, t=3.0)
    assert_events_equal(got, want)


def test_normal_sweep_after_growth_free_flight(pkg, oracle):
    """Growth free flight changes the radii on the device (rad += dt vr); a NORMAL
    sweep afterwards must use them (and must not assume they are still equal)."""
    c = pkg.synth.lattice_config(30000, 0.55, seed=48)
    rng = np.random.default_rng(48)
    vr = 0.02 * rng.random(c["n"])
    with pkg.EdmdCuda(c["n"], c["lx"], c["ly"]) as ctx:
        ctx.upload(c["x"], c["y"], c["vx"], c["vy"], c["rad"], t=0.0)
        ctx.set_growth(vr)
        ctx.free_fly(0.125, mode=1)
        s = ctx.download_state()
        cells = oracle.cells(c["n"], c["lx"], c["ly"], c["x"], c["y"]).reshape(c["n"], 2)
        got = ctx.predict_all(allow_overlap=True)
        assert ctx.stat(pkg.binding.STAT_LEAN_SWEEPS) == 0
    assert not np.array_equal(s["rad"], c["rad"])
    c2 = dict(c, x=s["x"], y=s["y"], rad=s["rad"])
    want = oracle_sweep(oracle, c2, t=0.125, cells=cells)
    assert_events_equal(got, want)


@pytest.mark.parametrize("n,phi,seed,sf,vscale", [(300000, 0.70, 241, 0.0, 1.0), (300000, 0.85, 242, 0.0, 1e-4),
                                                  (200000, 0.70, 243, 0.3, 1.0), (50000, 0.40, 244, 0.0, 300.0),
                                                  (2700, 0.70, 245, 0.3, 1.0), (1000000, 0.70, 246, 0.0, 1.0)])
def test_tile_sweep_equals_lean_chain_and_full_path(pkg, n, phi, seed, sf, vscale):
    """The two-kernel cell-slot sweep (cell_sweep.cu), the five-kernel lean chain and the full
    FP64 path produce the same bits (one and two radius classes, tiles narrower than a
    full tile at the grid edge, grids of a single tile)."""
    c = pkg.synth.lattice_config(n, phi, seed, small_fraction=sf, shuffle=True)
    with pkg.EdmdCuda(c["n"], c["lx"], c["ly"]) as ctx:
        ctx.upload(c["x"], c["y"], c["vx"] * vscale, c["vy"] * vscale, c["rad"], t=1.25)
        a = ctx.predict_all()
        assert ctx.stat(pkg.binding.STAT_LEAN_SWEEPS) == 1
        a2 = ctx.predict_all()          # the per-tile cursors clean themselves
        ctx.set_option(pkg.binding.OPT_NO_TILE, 1)
        b = ctx.predict_all()
        assert ctx.stat(pkg.binding.STAT_LEAN_SWEEPS) == 3
        ctx.set_option(pkg.binding.OPT_NO_LEAN, 1)
        d = ctx.predict_all()
        assert ctx.stat(pkg.binding.STAT_LEAN_SWEEPS) == 3
    for k in ("t_cross", "dir", "t_coll", "partner", "ctype"):
        assert np.array_equal(a[k], d[k]), k
        assert np.array_equal(a2[k], d[k]), k
        assert np.array_equal(b[k], d[k]), k


@pytest.mark.parametrize("n,phi,seed,sf", [(300000, 0.70, 251, 0.0), (200000, 0.85, 252, 0.0), (100000, 0.72, 253, 0.3)])
def test_tile_psi6_equals_row_kernel(pkg, oracle, n, phi, seed, sf):
    """psi6 on the tile buckets (fresh partition, and re-using the buckets of a sweep) gives the
    same values as the row kernel over the full cell index (to 1e-13), and the oracle's."""
    c = pkg.synth.lattice_config(n, phi, seed, small_fraction=sf, shuffle=True)
    with pkg.EdmdCuda(c["n"], c["lx"], c["ly"]) as ctx:
        ctx.upload(c["x"], c["y"], c["vx"], c["vy"], c["rad"], t=0.0)
        a = ctx.boop_cutoff(2.5)              # partition + tile kernel
        ctx.predict_all()
        b = ctx.boop_cutoff(2.5)              # buckets of the sweep, run lengths from the kept copy
        b2 = ctx.boop_cutoff(2.5)
        ctx.set_option(pkg.binding.OPT_NO_TILE_BOOP, 1)
        d = ctx.boop_cutoff(2.5)              # row kernel
    # the tile kernel is reproducible bit for bit (cells sorted by id, fixed summation order) ...
    for k in ("q5", "q6", "q7", "q6_arg", "neighbors"):
        assert np.array_equal(a[k], b[k]), k
        assert np.array_equal(a[k], b2[k]), k
    # ... and agrees with the row kernel (2^-48 fixed-point sums) far inside the 1e-10 gate
    assert np.array_equal(a["neighbors"], d["neighbors"])
    for k in ("q5", "q6", "q7"):
        assert np.abs(a[k] - d[k]).max() < 1e-13, k
    assert_boop_close(a, d)
    assert_boop_close(a, oracle.boop_cutoff(c["n"], c["lx"], c["ly"], c["x"], c["y"], 2.5))


@pytest.mark.parametrize("rad,spacing", [(0.1, 0.3), (0.3, 0.95)])
def test_cell_slot_sweep_declines_when_cells_overflow(pkg, oracle, rad, spacing):
    """Tiny disks crowded into one corner of a dilute box (cells are ~2 wide whatever the radius):
    spacing 0.3 puts ~40 disks in a cell -- more than the kSlotK slots a cell has --, spacing 0.95 about
    four -- more extras than a frame's list holds.  The sweep declines on the device and the call
    re-runs on the full path; the results are the reference's either way."""
    rng = np.random.default_rng(77)
    lx, ly, n = 400.0, 300.0, 6000
    side = int(np.ceil(np.sqrt(n)))
    k = np.arange(n)
    x = 1.0 + spacing * (k % side) + 0.02 * rng.random(n)
    y = 1.0 + spacing * (k // side) + 0.02 * rng.random(n)
    vx, vy = rng.standard_normal(n), rng.standard_normal(n)
    c = dict(n=n, lx=lx, ly=ly, x=x, y=y, vx=vx, vy=vy, rad=np.full(n, rad))
    with pkg.EdmdCuda(n, lx, ly) as ctx:
        ctx.upload(x, y, vx, vy, c["rad"], t=0.5)
        got = ctx.predict_all()
        assert ctx.stat(pkg.binding.STAT_LEAN_DECLINES) >= 1
        got2 = ctx.predict_all()
    want = oracle_sweep(oracle, c, t=0.5)
    assert_events_equal(got, want)
    assert_events_equal(got2, want)


def test_cell_slot_sweep_takes_a_crowded_corner(pkg, oracle):
    """One disk per cell, all in one corner of the box (the tile buckets of round 2's first sweep
    overflowed on this): the cell slots take it."""
    rng = np.random.default_rng(78)
    lx, ly, n = 400.0, 300.0, 6000
    side = int(np.ceil(np.sqrt(n)))
    k = np.arange(n)
    x = 1.0 + 2.05 * (k % side) + 0.02 * rng.random(n)
    y = 1.0 + 2.05 * (k // side) + 0.02 * rng.random(n)
    vx, vy = rng.standard_normal(n), rng.standard_normal(n)
    c = dict(n=n, lx=lx, ly=ly, x=x, y=y, vx=vx, vy=vy, rad=np.ones(n))
    with pkg.EdmdCuda(n, lx, ly) as ctx:
        ctx.upload(x, y, vx, vy, c["rad"], t=0.5)
        got = ctx.predict_all()
        assert ctx.stat(pkg.binding.STAT_LEAN_DECLINES) == 0
    assert_events_equal(got, oracle_sweep(oracle, c, t=0.5))


def test_rescale_after_growth_free_flight_stays_off_the_lean_path(pkg, oracle):
    """free_fly(GROW) changes the radii on the device; a velocity rescale afterwards
    refreshes the lean facts from flags that predate the growth -- it must not put the
    sweep back on the equal-radii path (the device-resident stopGrow tick)."""
    c = pkg.synth.lattice_config(30000, 0.55, seed=148)
    rng = np.random.default_rng(148)
    vr = 0.02 * rng.random(c["n"])
    with pkg.EdmdCuda(c["n"], c["lx"], c["ly"]) as ctx:
        ctx.upload(c["x"], c["y"], c["vx"], c["vy"], c["rad"], t=0.0)
        ctx.set_growth(vr)
        ctx.free_fly(0.125, mode=1)
        ctx.rescale_velocities(1.0)
        s = ctx.download_state()
        cells = oracle.cells(c["n"], c["lx"], c["ly"], c["x"], c["y"]).reshape(c["n"], 2)
        got = ctx.predict_all(allow_overlap=True)
        assert ctx.stat(pkg.binding.STAT_LEAN_SWEEPS) == 0
        # an upload that carries radii re-derives the classes: eligible again
        ctx.upload(c["x"], c["y"], c["vx"], c["vy"], c["rad"], t=0.0)
        ctx.predict_all()
        assert ctx.stat(pkg.binding.STAT_LEAN_SWEEPS) == 1
    c2 = dict(c, x=s["x"], y=s["y"], vx=s["vx"], vy=s["vy"], rad=s["rad"])
    want = oracle_sweep(oracle, c2, t=0.125, cells=cells)
    assert_events_equal(got, want)


def test_analysis_between_a_declined_lean_sweep_and_its_fetch(pkg, oracle):
    """predict_device on the lean path that declines on the device, then an analysis call
    (which rebuilds the full index), then the fetch: the pending decline must still be
    resolved -- no stale predictions with return code 0."""
    c = pkg.synth.lattice_config(20000, 0.30, seed=147)
    with pkg.EdmdCuda(c["n"], c["lx"], c["ly"]) as ctx:
        ctx.upload(c["x"], c["y"], c["vx"], c["vy"], c["rad"], t=0.0)
        ctx.free_fly(1.5)
        s = ctx.download_state()
        cells = oracle.cells(c["n"], c["lx"], c["ly"], c["x"], c["y"]).reshape(c["n"], 2)
        ctx.predict_device()
        ctx.boop_cutoff(2.5)
        got = ctx.fetch_predictions(allow_overlap=True)
        assert ctx.stat(pkg.binding.STAT_LEAN_SWEEPS) == 0
    c2 = dict(c, x=s["x"], y=s["y"])
    want = oracle_sweep(oracle, c2, t=1.5, cells=cells)
    assert_events_equal(got, want)


def test_lean_declines_after_free_flight_out_of_the_cells(pkg, oracle):
    """Free flight moves particles but not their (host-owned) cells; once some
    particle is more than a cell away from where it is filed the lean sweep
    declines on the device and the call falls back to the full path."""
    c = pkg.synth.lattice_config(20000, 0.30, seed=47)
    with pkg.EdmdCuda(c["n"], c["lx"], c["ly"]) as ctx:
        ctx.upload(c["x"], c["y"], c["vx"], c["vy"], c["rad"], t=0.0)
        ctx.free_fly(1.5)
        s = ctx.download_state()
        cells = oracle.cells(c["n"], c["lx"], c["ly"], c["x"], c["y"]).reshape(c["n"], 2)
        got = ctx.predict_all(allow_overlap=True)
        assert ctx.stat(pkg.binding.STAT_LEAN_SWEEPS) == 0       # declined on the device
    c2 = dict(c, x=s["x"], y=s["y"])
    want = oracle_sweep(oracle, c2, t=1.5, cells=cells)
    assert_events_equal(got, want)


def test_dense_small_disks_overflow_the_tile_buffer(pkg, oracle):
    """Many tiny disks per cell: more particles than a tile's staging buffer
    holds, so CTAs take the global-memory path; results unchanged."""
    rng = np.random.default_rng(31)
    lx, ly = 64.0, 48.0
    n = 40000          # ~52 per cell
    x = rng.random(n) * lx
    y = rng.random(n) * ly
    vx = rng.standard_normal(n)
    vy = rng.standard_normal(n)
    rad = np.full(n, 1e-4)
    c = dict(n=n, lx=lx, ly=ly, x=x, y=y, vx=vx, vy=vy, rad=rad)
    got = gpu_sweep(pkg, c, t=0.25)
    want = oracle_sweep(oracle, c, t=0.25)
    assert_events_equal(got, want)
    with pkg.EdmdCuda(n, lx, ly) as ctx:
        ctx.upload(x, y, vx, vy, rad, t=0.25)
        b = ctx.boop_cutoff(0.5)
    assert_boop_close(b, oracle.boop_cutoff(n, lx, ly, x, y, 0.5))


# ------------------------------------ Voronoi family, g6 correlation, S(q) ----
from helpers import VORONOI_CASES, VORONOI_GEOM_ATOL, random_points  # noqa: E402


def _upload_points(pkg, c):
    ctx = pkg.EdmdCuda(c["n"], c["lx"], c["ly"])
    ctx.upload(c["x"], c["y"], c["vx"], c["vy"], c["rad"], t=0.0)
    return ctx


@pytest.mark.parametrize("name", VORONOI_CASES)
def test_voronoi_family_matches_reference_golden(pkg, name):
    """K5 (one locally clipped Voronoi cell per particle) against the reference's
    global jc_voronoi diagram: neighbour counts exact, psi within 1e-10; cell area
    and perimeter; g6 correlation with the device's own Voronoi psi6 and with the
    reference's psi6 handed in."""
    g = load_golden(name)
    c = dict(n=int(g["n"]), lx=float(g["lx"]), ly=float(g["ly"]), x=g["x"], y=g["y"], vx=g["vx"], vy=g["vy"],
             rad=g["rad"])
    with _upload_points(pkg, c) as ctx:
        b = ctx.boop_voronoi()
        cells = ctx.voronoi_cells()
        g6 = ctx.g6_correlation(float(g["g6_dr"]), float(g["g6_max_r"]))
        g6b = ctx.g6_correlation(float(g["g6_dr"]), float(g["g6_max_r"]),
                                 g["vor_q6"] * np.cos(g["vor_q6_arg"]), g["vor_q6"] * np.sin(g["vor_q6_arg"]))
    assert_boop_close(b, g, prefix="vor_")
    assert abs(b["mean_q6"] - g["vor_q6"].mean()) <= ANALYSIS_ATOL
    assert np.array_equal(cells["neighbors"], g["vor_neighbors"])
    assert np.abs(cells["area"] - g["vor_area"]).max() <= VORONOI_GEOM_ATOL
    assert np.abs(cells["perimeter"] - g["vor_perimeter"]).max() <= VORONOI_GEOM_ATOL
    for r in (g6, g6b):
        assert np.array_equal(r["counts"], g["g6_counts"])
        assert np.abs(r["g6_corr"] - g["g6_corr"]).max() <= ANALYSIS_ATOL


@pytest.mark.parametrize("name", VORONOI_CASES)
def test_structure_factor_matches_reference_golden(pkg, name):
    g = load_golden(name)
    c = dict(n=int(g["n"]), lx=float(g["lx"]), ly=float(g["ly"]), x=g["x"], y=g["y"], vx=g["vx"], vy=g["vy"],
             rad=g["rad"])
    with _upload_points(pkg, c) as ctx:
        s = ctx.structure_factor(float(g["sq_qmax"]))
        v = ctx.structure_factor(float(g["sq_qmax"]), velocity=True)
    assert np.array_equal(s["qx"], g["sq_qx"]) and np.array_equal(s["qy"], g["sq_qy"])
    assert (np.abs(s["s"] - g["sq_s"]) / np.maximum(1.0, g["sq_s"])).max() <= ANALYSIS_ATOL
    assert (np.abs(v["s"] - g["sq_s_velocity"]) / np.maximum(1.0, g["sq_s_velocity"])).max() <= ANALYSIS_ATOL
    assert np.abs(s["s"] - (s["re"] ** 2 + s["im"] ** 2) / c["n"]).max() <= 1e-9


@pytest.mark.parametrize("kind,n,seed", [("liquid", 40000, 21), ("poisson", 30000, 22), ("jitter", 40000, 23),
                                         ("dense", 30000, 24)])
def test_voronoi_family_matches_oracle(pkg, oracle, kind, n, seed):
    if kind == "liquid":
        c = pkg.synth.lattice_config(n, 0.70, seed=seed, small_fraction=0.3)
    elif kind == "dense":
        c = pkg.synth.lattice_config(n, 0.88, seed=seed)
    elif kind == "poisson":
        c = random_points(n, 310.0, 290.0, seed)
    else:
        c = random_points(n, 400.0, 380.0, seed, jitter=0.95)
    n, lx, ly = c["n"], c["lx"], c["ly"]
    with _upload_points(pkg, c) as ctx:
        b = ctx.boop_voronoi()
        b2 = ctx.boop_voronoi()
        cells = ctx.voronoi_cells()
    want = oracle.boop_voronoi(n, lx, ly, c["x"], c["y"])
    assert_boop_close(b, want)
    assert b["neighbors"].sum() == 6 * n
    for k in ("q5", "q6", "q7", "q6_arg", "neighbors"):      # reproducible bit for bit
        assert np.array_equal(b[k], b2[k]), k
    wa = oracle.voronoi_area(n, lx, ly, c["x"], c["y"])
    assert np.abs(cells["area"] - wa["area"]).max() <= VORONOI_GEOM_ATOL
    assert np.abs(cells["perimeter"] - wa["perimeter"]).max() <= VORONOI_GEOM_ATOL
    assert abs(cells["area"].sum() - lx * ly) <= 1e-10 * lx * ly   # the cells tile the periodic box


def test_voronoi_cells_tile_the_box_at_full_size(pkg):
    """Size-independent checks at N = 10^6: the cells tile the box, and a periodic
    triangulation has exactly 3N edges (mean coordination 6)."""
    c = pkg.synth.lattice_config(1000000, 0.70, seed=31)
    with _upload_points(pkg, c) as ctx:
        cells = ctx.voronoi_cells()
        b = ctx.boop_voronoi()
    assert int(cells["neighbors"].sum()) == 6 * c["n"]
    assert abs(cells["area"].sum() - c["lx"] * c["ly"]) <= 1e-10 * c["lx"] * c["ly"]
    assert np.array_equal(b["neighbors"], cells["neighbors"])
    assert 0.0 < b["mean_q6"] <= 1.0


def test_voronoi_declines_systems_too_small_to_close_locally(pkg):
    x = np.array([1.0, 4.0, 2.5, 7.0])
    y = np.array([1.0, 2.0, 6.0, 7.5])
    with pkg.EdmdCuda(4, 9.0, 9.0) as ctx:
        ctx.upload(x, y, np.zeros(4), np.zeros(4), np.full(4, 0.1), t=0.0)
        with pytest.raises(pkg.EdmdError) as e:
            ctx.boop_voronoi()
    assert e.value.code == pkg.binding.EVORONOI


@pytest.mark.parametrize("n,seed,qmax", [(20000, 41, 0.6), (3000, 42, 2.0)])
def test_structure_factor_and_g6_match_oracle(pkg, oracle, n, seed, qmax):
    c = pkg.synth.lattice_config(n, 0.72, seed=seed)
    n, lx, ly = c["n"], c["lx"], c["ly"]
    rng = np.random.default_rng(seed)
    psi = rng.standard_normal((2, n)) * 0.5
    with _upload_points(pkg, c) as ctx:
        s = ctx.structure_factor(qmax)
        v = ctx.structure_factor(qmax, velocity=True)
        g6 = ctx.g6_correlation(0.25, min(lx, ly) / 2, psi[0], psi[1])
    ws = oracle.structure_factor(n, lx, ly, c["x"], c["y"], qmax)
    wv = oracle.structure_factor(n, lx, ly, c["x"], c["y"], qmax, c["vx"], c["vy"])
    assert np.array_equal(s["qx"], ws["qx"]) and np.array_equal(s["qy"], ws["qy"])
    assert (np.abs(s["s"] - ws["s"]) / np.maximum(1.0, ws["s"])).max() <= ANALYSIS_ATOL
    assert (np.abs(v["s"] - wv["s"]) / np.maximum(1.0, wv["s"])).max() <= ANALYSIS_ATOL
    wg = oracle.g6_correlation(n, lx, ly, c["x"], c["y"], psi[0], psi[1], 0.25, min(lx, ly) / 2)
    assert np.array_equal(g6["counts"], wg["counts"])
    assert np.abs(g6["g6_corr"] - wg["g6_corr"]).max() <= ANALYSIS_ATOL


# ------------------------------------------ thermostat tick on resident state ----
from helpers import TICK_CASES, TICK_RTOL, assert_tick_close  # noqa: E402


def _device_tick(ctx, t_new, T):
    ctx.free_fly(t_new)
    r = ctx.rescale_velocities(T)
    out = ctx.predict_all()
    out.update(ctx.download_state())
    out.update(r)
    return out


@pytest.mark.parametrize("name", TICK_CASES)
def test_device_thermostat_tick_matches_reference_golden(pkg, name):
    """A whole tick without the state leaving the device -- free flight, kinetic
    energy, velocity rescale, re-predict -- against the reference's physicalQ +
    addNoise on the same snapshot."""
    g = load_golden(name)
    c = cfg_of(g)
    with pkg.EdmdCuda(c["n"], c["lx"], c["ly"]) as ctx:
        ctx.upload(c["x"], c["y"], c["vx"], c["vy"], c["rad"], t=c["t"])
        k = ctx.kinetic()
        got = _device_tick(ctx, float(g["t_new"]), float(g["T"]))
        after = ctx.kinetic()
    assert abs(k["E"] - float(g["E_before"])) <= TICK_RTOL * float(g["E_before"])
    assert abs(k["px"] - c["vx"].sum()) < 1e-9 and abs(k["py"] - c["vy"].sum()) < 1e-9
    assert abs(got["E_before"] - k["E"]) == 0.0          # reproducible sum
    assert_tick_close(got, g)
    assert abs(after["E"] / c["n"] - float(g["T"])) < 1e-12


@pytest.mark.parametrize("n,phi,seed,sf", [(200000, 0.60, 51, 0.0), (1000000, 0.70, 52, 0.0), (100000, 0.55, 53, 0.3)])
def test_device_thermostat_tick_matches_oracle(pkg, oracle, n, phi, seed, sf):
    c = pkg.synth.lattice_config(n, phi, seed, small_fraction=sf)
    n = c["n"]
    t0, t1, T = 2.0, 2.0 + 1.0 / 64, 1.3
    with pkg.EdmdCuda(n, c["lx"], c["ly"]) as ctx:
        ctx.upload(c["x"], c["y"], c["vx"], c["vy"], c["rad"], t=t0)
        ctx.predict_all()
        got = _device_tick(ctx, t1, T)
    want = oracle.tick_rescale(n, c["lx"], c["ly"], t0, t1, T, c["x"], c["y"], c["vx"], c["vy"], c["rad"])
    assert want["rc"] == 0
    assert abs(got["E_before"] - want["E_before"]) <= TICK_RTOL * want["E_before"]
    assert np.array_equal(got["x"], want["x"]) and np.array_equal(got["y"], want["y"])
    for k in ("vx", "vy"):
        assert (np.abs(got[k] - want[k]) <= TICK_RTOL * np.abs(want[k]).max()).all()
    # a velocity that differs in the last bits can flip an exact near-tie: allow a handful
    same = got["partner"] == want["partner"]
    assert (~same).sum() <= max(2, n // 200000)
    assert np.array_equal(got["dir"], want["dir"])
    for k in ("t_cross", "t_coll"):
        ok = np.abs(got[k] - want[k]) <= TICK_RTOL * np.abs(want[k])
        assert (ok | ~same).all(), k


def test_rescale_keeps_the_lean_path_and_the_sweep_exact(pkg, oracle):
    """The rescale refreshes the velocity bound of the lean sweep's error model on the
    device: the next sweep still runs (and certifies) on the lean path, bit-exact
    for the rescaled velocities.  (A batched free flight that carries particles
    across the periodic edge without their crossing events -- as the synthetic ticks
    above do -- makes the device decline the lean path: positions then sit a box
    length from their filed cells.)"""
    c = pkg.synth.lattice_config(300000, 0.70, seed=54)
    B = pkg.binding
    with pkg.EdmdCuda(c["n"], c["lx"], c["ly"]) as ctx:
        ctx.upload(c["x"], c["y"], c["vx"] * 7.0, c["vy"] * 7.0, c["rad"], t=1.0)
        ctx.predict_all()
        r = ctx.rescale_velocities(0.01)          # a 70-fold change of the velocity scale
        assert ctx.stat(B.STAT_LEAN_ELIGIBLE) == 1
        got = ctx.predict_all()
        s = ctx.download_state()
        assert ctx.stat(B.STAT_LEAN_SWEEPS) == 2 and ctx.stat(B.STAT_LEAN_DECLINES) == 0
    assert abs(r["divisor"] - np.sqrt(r["E_before"] / c["n"] / 0.01)) <= 1e-15 * r["divisor"]
    assert np.array_equal(s["vx"], c["vx"] * 7.0 / r["divisor"])
    want = oracle.predict_all(c["n"], c["lx"], c["ly"], 1.0, c["x"], c["y"], s["vx"], s["vy"], c["rad"])
    assert_events_equal(got, want)


@pytest.mark.parametrize("n,phi,seed,T,gdt", [(200000, 0.70, 181, 1.0, 0.5), (50000, 0.55, 182, 0.25, 3.0)])
def test_device_langevin_kick(pkg, oracle, n, phi, seed, T, gdt):
    """edmd_cuda_langevin_kick = the Langevin branch of addNoise (randomGaussian,
    src/EDMD.c:5802-5826) with a counter-based generator instead of the reference's
    sequential MT19937.  (1) exact to rounding against the numpy restatement of the same
    generator (synth.uniform01) and formula; (2) the reference's statistics: per kick
    <v^2> -> c^2 <v^2> + (1 - c^2) T and no drift; (3) the re-predict sweep on the kicked,
    resident state is the oracle's on the downloaded state, bit for bit."""
    c = pkg.synth.lattice_config(n, phi, seed)
    n = c["n"]
    vx0, vy0 = 1.7 * c["vx"], 1.7 * c["vy"]            # start hot: T0 = 2.89
    gamma, dt, tick = 2.0, gdt / 2.0, 7
    with pkg.EdmdCuda(n, c["lx"], c["ly"]) as ctx:
        ctx.upload(c["x"], c["y"], vx0, vy0, c["rad"], t=0.0)
        e0 = ctx.kinetic()["E"]
        ctx.langevin_kick(T, gamma, dt, seed, tick)
        e1 = ctx.kinetic()
        ctx.predict_device()
        got = ctx.fetch_predictions()
        st = ctx.download_state()
    ids = np.arange(n, dtype=np.int64)
    u1 = 1.0 - pkg.synth.uniform01(seed, ids, 2 * tick)
    u2 = pkg.synth.uniform01(seed, ids, 2 * tick + 1)
    a, b = np.sqrt(-2.0 * np.log(u1)), 2.0 * np.pi * u2
    cc = np.exp(-gamma * dt)
    std = np.sqrt(T * (1.0 - cc * cc))
    wx, wy = std * a * np.cos(b) + vx0 * cc, std * a * np.sin(b) + vy0 * cc
    assert np.abs(st["vx"] - wx).max() <= 1e-13 and np.abs(st["vy"] - wy).max() <= 1e-13
    assert np.array_equal(st["x"], c["x"]) and np.array_equal(st["y"], c["y"])
    want_e = cc * cc * e0 / n + (1.0 - cc * cc) * T          # energy per particle = <vx^2> = T in 2-D, m = 1
    assert abs(e1["E"] / n - want_e) <= 0.02 * want_e          # 1 / sqrt(n) = 0.2-0.5 %
    drift = 5.0 * np.sqrt(want_e / n)                          # five standard errors of a mean velocity
    assert abs(e1["px"]) / n < drift and abs(e1["py"]) / n < drift
    want = oracle.predict_all(n, c["lx"], c["ly"], 0.0, st["x"], st["y"], st["vx"], st["vy"], c["rad"])
    assert_events_equal(got, want)


# ------------------------------------------------- slabs on ONE GPU (no IPC, no NCCL) ----
@pytest.mark.parametrize("n,phi,seed,sf,nslabs", [(1000000, 0.70, 301, 0.0, 4), (1000000, 0.85, 302, 0.0, 2),
                                                  (300000, 0.70, 303, 0.3, 3), (60000, 0.72, 304, 0.0, 8)])
def test_slab_decomposition_on_one_gpu_matches_oracle(pkg, oracle, n, phi, seed, sf, nslabs):
    """The multi-GPU path with every slab context on the same device: ownership by cell row, the
    one-row halo through edmd_cuda_halo_pack / edmd_cuda_halo_append (device buffers, no IPC), the
    sweep (tile path) and psi6 of every slab against the whole-system oracle, bit for bit; the
    weighted mean of the per-slab mean q6 is the whole system's."""
    import torch
    slab = pkg.slab
    c = pkg.synth.lattice_config(n, phi, seed, small_fraction=sf, shuffle=True)
    N, lx, ly, t = c["n"], c["lx"], c["ly"], 0.75
    cells = oracle.cells(N, lx, ly, c["x"], c["y"]).reshape(N, 2)
    want = oracle_sweep(oracle, c, t=t)
    wb = oracle.boop_cutoff(N, lx, ly, c["x"], c["y"], 2.5)
    ranks = [slab.SlabRank(pkg, N, lx, ly, r, nslabs, 0) for r in range(nslabs)]
    try:
        gids = [sr.load_owned(c, cells, t) for sr in ranks]
        rec = slab.HALO_REC.itemsize
        packed = []
        for sr in ranks:
            n_lo = sr.ctx.halo_pack(0, sr.send[0].data_ptr(), sr.halo_capacity)
            n_hi = sr.ctx.halo_pack(1, sr.send[1].data_ptr(), sr.halo_capacity)
            packed.append((n_lo, n_hi))
        torch.cuda.synchronize()
        for r, sr in enumerate(ranks):
            lower, upper = slab.neighbours(r, nslabs)
            # my local row 0 is the lower neighbour's LAST row (its side 1); my last row its upper's FIRST
            sr.ctx.halo_append(0, ranks[lower].send[1].data_ptr(), packed[lower][1])
            sr.ctx.halo_append(1, ranks[upper].send[0].data_ptr(), packed[upper][0])
        seen = np.zeros(N, bool)
        q6sum = 0.0
        for sr, gid in zip(ranks, gids):
            got = sr.predict()
            assert sr.ctx.stat(pkg.binding.STAT_LEAN_SWEEPS) == 1
            for k in ("t_cross", "dir", "t_coll", "partner", "ctype"):
                assert np.array_equal(got[k], want[k][gid]), (k, sr.rank)
            b = sr.ctx.boop_cutoff(2.5)
            assert np.array_equal(b["neighbors"], wb["neighbors"][gid])
            for k in ("q5", "q6", "q7"):
                assert np.abs(b[k] - wb[k][gid]).max() <= 1e-10, k
            q6sum += b["mean_q6"] * len(gid)
            seen[gid] = True
        assert seen.all()
        assert abs(q6sum / N - wb["q6"].mean()) < 1e-12
    finally:
        for sr in ranks:
            sr.close()


from helpers import GOLDEN, NORMALIZE_CASES  # noqa: E402


@pytest.mark.gpu
@pytest.mark.parametrize("name", NORMALIZE_CASES)
def test_device_normalize_matches_reference_golden(pkg, name):
    """edmd_cuda_normalize_velocities against the reference's normalizePhysicalQ (src/EDMD.c:5723-5764):
    velocities within 1e-12 (parallel sums), the sweep on the normalised resident state follows."""
    g = np.load(GOLDEN / f"{name}.npz")
    n = int(g["n"])
    with pkg.EdmdCuda(n, float(g["lx"]), float(g["ly"])) as ctx:
        ctx.upload(g["x"], g["y"], g["vx"], g["vy"], g["rad"], t=0.0)
        info = ctx.normalize_velocities(float(g["e_init"]))
        st = ctx.download_state()
    for k in ("vx", "vy"):
        w = g["norm_" + k]
        assert (np.abs(st[k] - w) <= TICK_RTOL * np.abs(w).max()).all(), k
    assert abs(info["px_before"] - float(g["px_before"])) <= 1e-12 * abs(float(g["px_before"]))
    assert abs(info["E_shifted"] - float(g["E_shifted"])) <= 1e-12 * float(g["E_shifted"])


@pytest.mark.gpu
def test_device_normalize_then_sweep_matches_oracle(pkg, oracle):
    """stopGrow's order on the device: normalise the resident velocities, re-predict -- against the oracle's
    restatement of both at N = 10^5; shift_scale with caller-supplied sums (the slab form) gives the same state."""
    c = pkg.synth.lattice_config(100000, 0.70, seed=31, shuffle=True)
    vx, vy = c["vx"] + 0.25, c["vy"] - 0.4
    want = oracle.normalize(vx, vy, 1.3)
    with pkg.EdmdCuda(c["n"], c["lx"], c["ly"]) as ctx:
        ctx.upload(c["x"], c["y"], vx, vy, c["rad"], t=0.0)
        ctx.normalize_velocities(1.3)
        st = ctx.download_state()
        got = ctx.predict_all()
        # the slab form: the caller's sums
        ctx.upload(c["x"], c["y"], vx, vy, c["rad"], t=0.0)
        k1 = ctx.kinetic()
        ctx.shift_scale_velocities(k1["px"] / c["n"], k1["py"] / c["n"], 1.0)
        k2 = ctx.kinetic()
        ctx.shift_scale_velocities(0.0, 0.0, np.sqrt(k2["E"] / c["n"] / 1.3))
        st2 = ctx.download_state()
    for k in ("vx", "vy"):
        assert (np.abs(st[k] - want[k]) <= TICK_RTOL * np.abs(want[k]).max()).all(), k
        assert (np.abs(st2[k] - want[k]) <= TICK_RTOL * np.abs(want[k]).max()).all(), k
    ref = oracle.predict_all(c["n"], c["lx"], c["ly"], 0.0, c["x"], c["y"], st["vx"], st["vy"], c["rad"])
    assert_events_equal(got, ref)


@pytest.mark.gpu
@pytest.mark.parametrize("nslabs,n,phi,sf", [(1, 30000, 0.70, 0.0), (2, 60000, 0.70, 0.0), (3, 150000, 0.72, 0.0), (4, 1000000, 0.70, 0.0),
                                             (2, 40000, 0.55, 0.3)])
def test_multi_gpu_c_entry_matches_oracle(pkg, oracle, nslabs, n, phi, sf):
    """edmd_cuda_create_mg (csrc/multi_gpu.cu): the single-GPU interface over several slab contexts --
    here several slabs on ONE device, the halo by direct stores between them: sweep bit-exact against the
    whole-system oracle, psi6 and its mean within 1e-10, g(r) counts exact.  Two ticks: particles that
    changed rows are dealt to another slab by the second upload."""
    import torch
    ndev = torch.cuda.device_count()
    devices = [k % ndev for k in range(nslabs)]
    c = pkg.synth.lattice_config(n, phi, seed=41, small_fraction=sf, shuffle=True)
    with pkg.EdmdMg(c["n"], c["lx"], c["ly"], devices) as mg:
        mg.upload(c["x"], c["y"], c["vx"], c["vy"], c["rad"], t=0.0)
        assert sum(mg.slab_sizes) == c["n"] and min(mg.slab_sizes) > 0
        got = mg.predict_all()
        assert_events_equal(got, oracle_sweep(oracle, c, t=0.0))
        b = mg.boop_cutoff(2.5)
        want_b = oracle.boop_cutoff(c["n"], c["lx"], c["ly"], c["x"], c["y"], 2.5)
        assert_boop_close(b, want_b)
        assert abs(b["mean_q6"] - want_b["q6"].mean()) < 1e-10
        if c["n"] <= 200000:
            g = mg.pcf(c["x"], c["y"], 0.1, 12.0)
            want_g = oracle.pcf(c["n"], c["lx"], c["ly"], c["x"], c["y"], 0.1, 12.0)
            assert np.array_equal(g["counts"], want_g["counts"])
            assert np.abs(g["g_r"] - want_g["g_r"]).max() <= 1e-12 * want_g["g_r"].max()
        # second tick: free flight on the host, new cells
        dt = 0.05
        c2 = dict(c)
        c2["x"] = np.mod(c["x"] + dt * c["vx"], c["lx"])
        c2["y"] = np.mod(c["y"] + dt * c["vy"], c["ly"])
        mg.upload(c2["x"], c2["y"], c2["vx"], c2["vy"], c2["rad"], t=dt)
        got2 = mg.predict_all(allow_overlap=True)
        want2 = oracle_sweep(oracle, c2, t=dt)
        assert_events_equal(got2, want2)


@pytest.mark.gpu
def test_pcf_g_r_alone_and_slab_rejection(pkg, oracle):
    """edmd_cuda_pcf with counts == NULL but g_r wanted computes g(r) (it used to return early);
    a slab context -- which holds copies of its neighbours' rows -- refuses the whole-system call."""
    import ctypes as C
    c = pkg.synth.lattice_config(20000, 0.7, seed=5)
    want = oracle.pcf(c["n"], c["lx"], c["ly"], c["x"], c["y"], 0.1, 9.0)
    with pkg.EdmdCuda(c["n"], c["lx"], c["ly"]) as ctx:
        ctx.upload(c["x"], c["y"], c["vx"], c["vy"], c["rad"], t=0.0)
        nb = C.c_int(0)
        g = np.zeros(want["num_bins"], np.float64)
        rc = ctx.lib.edmd_cuda_pcf(ctx._h, 0.1, 9.0, None, g.ctypes.data_as(C.c_void_p), C.byref(nb))
        assert rc == 0 and nb.value == want["num_bins"]
        assert np.abs(g - want["g_r"]).max() <= 1e-12 * want["g_r"].max()
    ny = int(c["ly"] / 2)
    with pkg.EdmdCuda(c["n"] + 4096, c["lx"], c["ly"], slab_rows=(0, ny // 2)) as sl:
        cells = oracle.cells(c["n"], c["lx"], c["ly"], c["x"], c["y"]).reshape(-1, 2)
        own = np.nonzero(cells[:, 1] < ny // 2)[0].astype(np.int32)
        sl.upload_owned(c["x"][own], c["y"][own], c["vx"][own], c["vy"][own], c["rad"][own], cells[own], own, t=0.0)
        cnt = np.zeros(want["num_bins"], np.uint64)
        rc = sl.lib.edmd_cuda_pcf(sl._h, 0.1, 9.0, cnt.ctypes.data_as(C.c_void_p), None, C.byref(nb))
        assert rc == pkg.binding.ESTATE


@pytest.mark.gpu
def test_multi_gpu_entry_with_an_empty_slab(pkg, oracle):
    """All particles in one corner of the box: the third of three slabs owns nothing (its sweep still has to
    keep the cell counters of its rows clean for the tick in which particles arrive there)."""
    import torch
    ndev = torch.cuda.device_count()
    rng = np.random.default_rng(79)
    lx, ly, n = 400.0, 300.0, 6000
    side = int(np.ceil(np.sqrt(n)))
    k = np.arange(n)
    x = 1.0 + 2.05 * (k % side) + 0.02 * rng.random(n)
    y = 1.0 + 2.05 * (k // side) + 0.02 * rng.random(n)
    c = dict(n=n, lx=lx, ly=ly, x=x, y=y, vx=rng.standard_normal(n), vy=rng.standard_normal(n), rad=np.ones(n))
    with pkg.EdmdMg(n, lx, ly, [j % ndev for j in range(3)]) as mg:
        for tick in range(3):
            mg.upload(c["x"], c["y"], c["vx"], c["vy"], c["rad"], t=0.5 * tick)
            assert mg.slab_sizes[2] == 0 or tick == 2
            assert_events_equal(mg.predict_all(), oracle_sweep(oracle, c, t=0.5 * tick))
            if tick == 1:   # for the last tick the blob moves up: the empty slab gets particles
                c = dict(c, y=np.mod(c["y"] + 130.0, ly))
        assert mg.slab_sizes[2] > 0
