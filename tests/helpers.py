"""Shared comparison helpers for the parity tests."""
from pathlib import Path

import numpy as np

GOLDEN = Path(__file__).resolve().parent / "golden"
GOLDEN_CASES = ["sweep_n2000_phi070_bidisperse", "sweep_n3000_phi085_mono",
                "sweep_n2500_phi072_ordered", "sweep_n2000_grow"]

# tolerances of BASELINE.json's north_star
TIME_RTOL = 1e-12     # event times (goal: bit-exact, asserted separately)
ANALYSIS_ATOL = 1e-10  # g(r), psi6


def load_golden(name):
    z = np.load(GOLDEN / f"{name}.npz")
    return {k: z[k] for k in z.files}


def cfg_of(g):
    c = dict(n=int(g["n"]), lx=float(g["lx"]), ly=float(g["ly"]), t=float(g["t"]),
             x=g["x"], y=g["y"], vx=g["vx"], vy=g["vy"], rad=g["rad"])
    if "vr" in g:
        c["vr"] = g["vr"]
    return c


def assert_events_equal(got, want, *, exact_times=True, prefix=""):
    """partner / dir / ctype bit-exact; times bit-exact or within 1e-12 rel."""
    for k in ("partner", "dir", "ctype"):
        w = want[prefix + k]
        assert np.array_equal(got[k], w), f"{k}: {(got[k] != w).sum()} mismatches"
    for k in ("t_cross", "t_coll"):
        w = want[prefix + k]
        if exact_times:
            assert np.array_equal(got[k], w), \
                f"{k}: {(got[k] != w).sum()} not bit-exact, max rel " \
                f"{np.max(np.abs(got[k] - w) / np.abs(w))}"
        else:
            rel = np.abs(got[k] - w) / np.maximum(np.abs(w), 1e-300)
            assert rel.max() <= TIME_RTOL, f"{k}: max rel err {rel.max()}"


def angle_diff(a, b):
    return np.abs(np.angle(np.exp(1j * (a - b))))


def assert_boop_close(got, want, prefix=""):
    assert np.array_equal(got["neighbors"], want[prefix + "neighbors"])
    for k in ("q5", "q6", "q7"):
        assert np.abs(got[k] - want[prefix + k]).max() <= ANALYSIS_ATOL, k
    # the argument is ill-conditioned where |sum6| ~ 0: weight by the modulus
    w6 = want[prefix + "q6"] * np.maximum(want[prefix + "neighbors"], 1)
    d = angle_diff(got["q6_arg"], want[prefix + "q6_arg"]) * np.minimum(w6, 1.0)
    assert d.max() <= ANALYSIS_ATOL, f"q6_arg {d.max()}"


def pcf_counts_from_g(g_r, n, lx, ly, dr):
    """Invert the reference normalisation (src/pcf.c:56-72) to integer pair counts."""
    nb = len(g_r)
    r = (np.arange(nb) + 0.5) * dr
    norm = 2 * np.pi * r * dr * (n / (lx * ly)) * n
    c = g_r * norm / 2.0
    ci = np.rint(c)
    assert np.abs(c - ci).max() < 1e-6
    return ci.astype(np.uint64)


VORONOI_CASES = ["voronoi_n1500_phi085", "voronoi_n2025_jittered", "voronoi_n1500_poisson"]
VORONOI_GEOM_ATOL = 1e-9   # area / perimeter: the reference sums cross products of absolute coordinates


def random_points(n, lx, ly, seed, jitter=None):
    """Point configurations for the Voronoi / S(q) family (radii are irrelevant to
    it): Poisson points, or a square lattice with a large jitter (units of the spacing)."""
    rng = np.random.default_rng(seed)
    if jitter is None:
        x, y = rng.random(n) * lx, rng.random(n) * ly
    else:
        m = int(round(np.sqrt(n)))
        n = m * m
        gx, gy = np.meshgrid(np.arange(m), np.arange(m))
        x = ((gx.ravel() + 0.5 + jitter * (rng.random(n) - 0.5)) * lx / m) % lx
        y = ((gy.ravel() + 0.5 + jitter * (rng.random(n) - 0.5)) * ly / m) % ly
        p = rng.permutation(n)
        x, y = x[p], y[p]
    return dict(n=n, lx=lx, ly=ly, x=x, y=y, vx=rng.standard_normal(n), vy=rng.standard_normal(n),
                rad=np.full(n, 0.05))


TICK_CASES = ["tick_n2000_phi060", "tick_n1500_phi045_bidisperse"]
NORMALIZE_CASES = ["normalize_n2000_phi060", "normalize_n1500_einit"]
TICK_RTOL = 1e-12   # the kinetic-energy sum is order dependent: rescaled velocities and event times to 1e-12


def assert_tick_close(got, want, prefix="tick_"):
    """State and events after a thermostat tick: positions bit-exact (free flight),
    velocities and times within 1e-12 relative, partners / directions exact."""
    assert np.array_equal(got["x"], want[prefix + "x"]) and np.array_equal(got["y"], want[prefix + "y"])
    for k in ("vx", "vy"):
        w = want[prefix + k]
        assert (np.abs(got[k] - w) <= TICK_RTOL * np.abs(w).max()).all(), k
    for k in ("partner", "dir", "ctype"):
        assert np.array_equal(got[k], want[prefix + k]), k
    for k in ("t_cross", "t_coll"):
        w = want[prefix + k]
        assert (np.abs(got[k] - w) <= TICK_RTOL * np.abs(w)).all(), k
