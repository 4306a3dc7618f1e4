"""CPU tests: the oracle restatement against the golden fixtures generated from
the unmodified reference, against the live reference where oracle/_ref exists,
and its own invariants."""
import numpy as np
import pytest

from helpers import (GOLDEN_CASES, ANALYSIS_ATOL, assert_boop_close, assert_events_equal,
                     cfg_of, load_golden, pcf_counts_from_g)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_oracle_sweep_matches_golden(oracle, name):
    g = load_golden(name)
    c = cfg_of(g)
    grow = bool(g["grow"])
    box = oracle.box(c["n"], c["lx"], c["ly"])
    assert (box.nx, box.ny) == (int(g["nx"]), int(g["ny"]))
    assert box.csx == float(g["csx"]) and box.csy == float(g["csy"])
    cells = oracle.cells(c["n"], c["lx"], c["ly"], c["x"], c["y"])
    assert np.array_equal(cells, g["cells"])
    got = oracle.predict_all(c["n"], c["lx"], c["ly"], c["t"], c["x"], c["y"], c["vx"], c["vy"],
                             c["rad"], vr=c.get("vr"), mode=int(grow))
    assert got["rc"] == 0 and tuple(got["overlap"]) == (-1, -1)
    # growth quadratic: the reference's -ffast-math build reassociates it (1e-15 rel)
    assert_events_equal(got, g, exact_times=not grow, prefix="first_")
    if not grow:
        # the reference's own thermostat tick (addNoise) re-predicts identically
        assert_events_equal(got, g, exact_times=True, prefix="re_")


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_oracle_free_fly_matches_golden(oracle, name):
    g = load_golden(name)
    c = cfg_of(g)
    grow = bool(g["grow"])
    got = oracle.free_fly(c["n"], c["lx"], c["ly"], c["t"], float(g["ff_t"]), c["x"], c["y"],
                          c["vx"], c["vy"], rad=c["rad"], vr=c.get("vr"), mode=int(grow))
    assert np.array_equal(got["x"], g["ff_x"])
    assert np.array_equal(got["y"], g["ff_y"])
    assert np.array_equal(got["rad"], g["ff_rad"])


@pytest.mark.parametrize("name", [n for n in GOLDEN_CASES if "grow" not in n])
def test_oracle_analysis_matches_golden(oracle, name):
    g = load_golden(name)
    c = cfg_of(g)
    b = oracle.boop_cutoff(c["n"], c["lx"], c["ly"], c["x"], c["y"], 2.5)
    assert_boop_close(b, g, prefix="boop_")
    p = oracle.pcf(c["n"], c["lx"], c["ly"], c["x"], c["y"], float(g["pcf_dr"]), float(g["pcf_max_r"]))
    assert p["num_bins"] == len(g["pcf_g"])
    assert np.abs(p["g_r"] - g["pcf_g"]).max() <= ANALYSIS_ATOL
    want_counts = pcf_counts_from_g(g["pcf_g"], c["n"], c["lx"], c["ly"], float(g["pcf_dr"]))
    assert np.array_equal(p["counts"], want_counts)


def test_oracle_vs_live_reference(oracle, reference, pkg):
    """Where the unmodified reference is built, re-pin on fresh seeds/sizes."""
    for n, phi, seed, sf in [(5000, 0.70, 11, 0.3), (20000, 0.72, 12, 0.0), (20000, 0.85, 13, 0.0)]:
        c = pkg.synth.lattice_config(n, phi, seed, small_fraction=sf)
        t = 0.75 * seed
        reference.setup(c["n"], c["lx"], c["ly"], t, c["x"], c["y"], c["vx"], c["vy"], c["rad"])
        want = reference.predict_first()
        got = oracle.predict_all(c["n"], c["lx"], c["ly"], t, c["x"], c["y"], c["vx"], c["vy"], c["rad"])
        assert_events_equal(got, want)
        assert_events_equal(got, reference.repredict())
        assert_boop_close(oracle.boop_cutoff(c["n"], c["lx"], c["ly"], c["x"], c["y"], 2.5),
                          reference.boop_cutoff(2.5))
    reference.teardown()


def test_oracle_host_cells_override(oracle, pkg):
    """Host-owned cell ids win over coordinates: a particle one
    ulp across its boundary keeps the host's cell."""
    c = pkg.synth.lattice_config(400, 0.6, seed=5)
    n = c["n"]
    cells = oracle.cells(n, c["lx"], c["ly"], c["x"], c["y"])
    a = oracle.predict_all(n, c["lx"], c["ly"], 0.0, c["x"], c["y"], c["vx"], c["vy"], c["rad"])
    b = oracle.predict_all(n, c["lx"], c["ly"], 0.0, c["x"], c["y"], c["vx"], c["vy"], c["rad"],
                           cell_xy=cells)
    assert_events_equal(a, b)


def test_oracle_symmetry_property(oracle, pkg):
    """If i's earliest partner is j at time tau and j's minimum is also tau,
    then j's partner is i (same pair time from both sides)."""
    c = pkg.synth.lattice_config(5000, 0.72, seed=6)
    o = oracle.predict_all(c["n"], c["lx"], c["ly"], 0.0, c["x"], c["y"], c["vx"], c["vy"], c["rad"])
    has = o["t_coll"] < 1e25
    i = np.nonzero(has)[0]
    j = o["partner"][i]
    mutual = o["t_coll"][j] == o["t_coll"][i]
    assert np.array_equal(o["partner"][j[mutual]], i[mutual])
    assert mutual.mean() > 0.3


def test_oracle_overlap_flag(oracle):
    lx = ly = 12.0
    x = np.array([3.0, 4.2, 9.0]); y = np.array([3.0, 3.0, 9.0])
    vx = np.array([1.0, -1.0, 0.3]); vy = np.array([0.0, 0.0, 0.1])
    rad = np.ones(3)
    o = oracle.predict_all(3, lx, ly, 0.0, x, y, vx, vy, rad)
    assert o["rc"] == 1 and tuple(o["overlap"]) == (0, 1)


def test_oracle_empty_and_single(oracle):
    e = np.zeros(0)
    o = oracle.predict_all(0, 10.0, 10.0, 0.0, e, e, e, e, e)
    assert o["rc"] == 0 and o["t_coll"].size == 0
    one = oracle.predict_all(1, 10.0, 10.0, 2.0, np.array([1.0]), np.array([1.0]),
                             np.array([0.5]), np.array([-0.25]), np.array([1.0]))
    assert one["partner"][0] == 0 and one["t_coll"][0] == 2.0 + 1e26
    assert one["dir"][0] == 2 and one["t_cross"][0] == 2.0 + (2.0 - 1.0) / 0.5


WEIGHTED_CASES = ["weighted_n1500_phi072", "weighted_n1200_phi060_bidisperse"]


@pytest.mark.parametrize("name", WEIGHTED_CASES)
def test_oracle_weighted_pcf_family_matches_golden(oracle, name):
    """Bragg-peak search and cos(k.r)-weighted g(r) (src/pcf.c:77-167, 405-467):
    the restatement against the reference's own outputs."""
    g = load_golden(name)
    n, lx, ly = int(g["n"]), float(g["lx"]), float(g["ly"])
    peak = oracle.bragg_peak(n, lx, ly, g["x"], g["y"], float(g["expected_bragg"]))
    assert np.allclose(peak["k"], g["k"], rtol=0, atol=1e-12)
    assert peak["s_max"] > 1.0
    bo = oracle.bond_order_pcf(n, lx, ly, g["x"], g["y"], float(g["dr"]), float(g["max_r"]), g["k"])
    assert bo["num_bins"] == len(g["g_r"])
    assert np.abs(bo["g_r"] - g["g_r"]).max() <= ANALYSIS_ATOL
    assert np.abs(bo["g6_r"] - g["g6_r"]).max() <= ANALYSIS_ATOL
    # plain g(r) of the same snapshot: same counts
    p = oracle.pcf(n, lx, ly, g["x"], g["y"], float(g["dr"]), float(g["max_r"]))
    assert np.array_equal(p["counts"], bo["counts"])


from helpers import VORONOI_CASES, VORONOI_GEOM_ATOL  # noqa: E402


@pytest.mark.parametrize("name", VORONOI_CASES)
def test_oracle_voronoi_family_matches_golden(oracle, name):
    """computeBOOPVoronoi, Voronoi cell area / perimeter, compute_g6_correlation
    (src/boop.c:15-59, src/voronoi_edmd.c:33-149, src/pcf.c:169-230): the restatement
    (the reference's image-augmented point set, diagram from Qhull) against the
    reference's own outputs (jc_voronoi)."""
    g = load_golden(name)
    n, lx, ly = int(g["n"]), float(g["lx"]), float(g["ly"])
    b = oracle.boop_voronoi(n, lx, ly, g["x"], g["y"])
    assert_boop_close(b, g, prefix="vor_")
    assert b["neighbors"].sum() == 6 * n          # Euler: a periodic triangulation has 3N edges
    a = oracle.voronoi_area(n, lx, ly, g["x"], g["y"])
    assert np.abs(a["area"] - g["vor_area"]).max() <= VORONOI_GEOM_ATOL
    assert np.abs(a["perimeter"] - g["vor_perimeter"]).max() <= VORONOI_GEOM_ATOL
    assert abs(a["area"].sum() - lx * ly) <= 1e-9 * lx * ly
    psi_re, psi_im = b["q6"] * np.cos(b["q6_arg"]), b["q6"] * np.sin(b["q6_arg"])
    c = oracle.g6_correlation(n, lx, ly, g["x"], g["y"], psi_re, psi_im, float(g["g6_dr"]), float(g["g6_max_r"]))
    assert np.array_equal(c["counts"], g["g6_counts"])
    assert np.abs(c["g6_corr"] - g["g6_corr"]).max() <= ANALYSIS_ATOL


@pytest.mark.parametrize("name", VORONOI_CASES)
def test_oracle_structure_factor_matches_golden(oracle, name):
    """initStructureFactor's grid + computeStructureFactor / computeVelocityStructureFactor
    (src/struc.c:328-345, 364-408)."""
    g = load_golden(name)
    n, lx, ly = int(g["n"]), float(g["lx"]), float(g["ly"])
    s = oracle.structure_factor(n, lx, ly, g["x"], g["y"], float(g["sq_qmax"]))
    assert np.array_equal(s["qx"], g["sq_qx"]) and np.array_equal(s["qy"], g["sq_qy"])
    assert s["s"].shape == g["sq_s"].shape
    assert (np.abs(s["s"] - g["sq_s"]) / np.maximum(1.0, g["sq_s"])).max() <= ANALYSIS_ATOL
    i0, j0 = (len(s["qx"]) - 1) // 2, (len(s["qy"]) - 1) // 2
    assert s["s"][i0, j0] == n                       # q = 0: S = N
    v = oracle.structure_factor(n, lx, ly, g["x"], g["y"], float(g["sq_qmax"]), g["vx"], g["vy"])
    assert (np.abs(v["s"] - g["sq_s_velocity"]) / np.maximum(1.0, g["sq_s_velocity"])).max() <= ANALYSIS_ATOL


from helpers import TICK_CASES, TICK_RTOL, assert_tick_close  # noqa: E402


@pytest.mark.parametrize("name", TICK_CASES)
def test_oracle_thermostat_tick_matches_golden(oracle, name):
    """physicalQ + addNoise's velocity-rescale branch (src/EDMD.c:5968-5997, 4828-4923)."""
    g = load_golden(name)
    c = cfg_of(g)
    got = oracle.tick_rescale(c["n"], c["lx"], c["ly"], c["t"], float(g["t_new"]), float(g["T"]),
                              c["x"], c["y"], c["vx"], c["vy"], c["rad"])
    assert abs(got["E_before"] - float(g["E_before"])) <= TICK_RTOL * float(g["E_before"])
    assert_tick_close(got, g)
    e_after = 0.5 * (got["vx"] ** 2 + got["vy"] ** 2).sum()
    assert abs(e_after / c["n"] - float(g["T"])) < 1e-12          # the tick sets E/N = T


def test_reference_grown_liquid_fixture_and_its_tiling(pkg, oracle):
    """tests/golden/liquid_n10000_phi070.npz (the reference's own main() grew and
    equilibrated it; make_liquid.py) is a periodic hard-disk liquid at phi = 0.70 whose
    radii carry the growth phase's rounding; the k x k tiling keeps it overlap-free,
    and the oracle's sweep on it finds a real collision for (nearly) every disk."""
    from helpers import load_golden
    base = load_golden("liquid_n10000_phi070")
    n0, lx, ly = len(base["x"]), float(base["lx"]), float(base["ly"])
    assert n0 == 10000 and abs(np.pi * np.sum(base["rad"] ** 2) / (lx * ly) - 0.70) < 1e-9
    assert 1 < len(np.unique(base["rad"])) < 100 and np.abs(base["rad"] - 1).max() < 1e-12
    c = pkg.synth.tiled_config(base, 2, seed=5)
    assert c["n"] == 4 * n0 and (c["x"] < c["lx"]).all() and (c["y"] < c["ly"]).all()
    assert abs(c["vx"].mean()) < 1e-12 and abs(c["vy"].mean()) < 1e-12
    out = oracle.predict_all(c["n"], c["lx"], c["ly"], 0.0, c["x"], c["y"], c["vx"], c["vy"], c["rad"])
    assert out["overlap"][0] < 0                      # no overlapping pair anywhere
    assert (out["t_coll"] < 1e20).mean() > 0.95       # a dense liquid: nearly everyone has a partner
    # the tiling is periodic with the base's period: pair counts of the tiling = 4 x the base's
    pb = oracle.pcf(n0, lx, ly, base["x"], base["y"], 0.1, 8.0)
    pt = oracle.pcf(c["n"], c["lx"], c["ly"], c["x"], c["y"], 0.1, 8.0)
    assert np.array_equal(pt["counts"], 4 * pb["counts"])


from helpers import GOLDEN, NORMALIZE_CASES  # noqa: E402


@pytest.mark.parametrize("name", NORMALIZE_CASES)
def test_oracle_normalize_matches_golden(oracle, name):
    """normalizePhysicalQ (src/EDMD.c:5723-5764): the numpy restatement against the reference's own run
    (sequential sums in both: bit-exact but for the -ffast-math build's reassociation, <= 1e-15)."""
    g = np.load(GOLDEN / f"{name}.npz")
    got = oracle.normalize(g["vx"], g["vy"], float(g["e_init"]))
    for k in ("vx", "vy"):
        w = g["norm_" + k]
        assert (np.abs(got[k] - w) <= 1e-14 * np.abs(w).max()).all(), k
    assert abs(got["px_before"] - float(g["px_before"])) <= 1e-12 * abs(float(g["px_before"]))
    assert abs(got["E_shifted"] - float(g["E_shifted"])) <= 1e-13 * float(g["E_shifted"])
    # what the routine is for: no centre-of-mass motion, E/N = Einit
    assert abs(got["vx"].sum()) < 1e-9 and abs(got["vy"].sum()) < 1e-9
    e = 0.5 * (got["vx"] ** 2 + got["vy"] ** 2).sum() / len(got["vx"])
    assert abs(e - float(g["e_init"])) < 1e-12 * float(g["e_init"])
