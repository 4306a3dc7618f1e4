"""Generate the golden fixtures in this directory from the UNMODIFIED reference
(oracle/_ref/libedmd_ref.so, built from /root/reference by oracle/Makefile).

Run in the build container (where /root/reference is mounted):
    python tests/golden/make_golden.py
The reference itself ships no tests or golden vectors;
these files pin the oracle restatement -- and through it the CUDA path -- to
outputs of the reference's own functions on fixed inputs.
"""
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as entry  # noqa: E402
from oracle.oracle_py import Reference  # noqa: E402

pkg = entry.load_package()
ref = Reference()
ONLY = set(sys.argv[1:])   # e.g. `make_golden.py voronoi` regenerates only that family


def want(family):
    return not ONLY or family in ONLY



def sweep_case(name, cfg, t, grow=False, with_analysis=True, pcf_max_r=None):
    n = cfg["n"]
    ref.setup(n, cfg["lx"], cfg["ly"], t, cfg["x"], cfg["y"], cfg["vx"], cfg["vy"], cfg["rad"],
              vr=cfg.get("vr"))
    out = dict(n=n, lx=cfg["lx"], ly=cfg["ly"], t=t, grow=int(grow),
               x=cfg["x"], y=cfg["y"], vx=cfg["vx"], vy=cfg["vy"], rad=cfg["rad"])
    if grow:
        out["vr"] = cfg["vr"]
    box = ref.box()
    out.update(nx=box["nx"], ny=box["ny"], csx=box["csx"], csy=box["csy"])
    out["cells"] = ref.cells()
    first = ref.predict_first(grow=grow)
    for k in ("t_cross", "dir", "t_coll", "partner", "ctype"):
        out["first_" + k] = first[k]
    if not grow:
        again = ref.repredict()  # the reference's own addNoise() tick
        for k in ("t_cross", "dir", "t_coll", "partner", "ctype"):
            out["re_" + k] = again[k]
    if with_analysis and not grow:
        b = ref.boop_cutoff(2.5)
        for k in ("q5", "q6", "q7", "q6_arg", "neighbors"):
            out["boop_" + k] = b[k]
        max_r = pcf_max_r if pcf_max_r else min(cfg["lx"], cfg["ly"]) / 2
        p = ref.pcf(0.1, max_r)
        out["pcf_dr"] = 0.1
        out["pcf_max_r"] = max_r
        out["pcf_g"] = p["g_r"]
        out["pcf_r"] = p["r"]
    ff = ref.free_fly(t + 0.4375, grow=grow)
    out["ff_t"] = t + 0.4375
    out["ff_x"], out["ff_y"], out["ff_rad"] = ff["x"], ff["y"], ff["rad"]
    np.savez_compressed(HERE / f"{name}.npz", **out)
    print("wrote", name, "n =", n)


def sweep_cases():
    # BASELINE config 1 flavour: N=2000, phi=0.7, 30 % small disks (reference CLI defaults)
    sweep_case("sweep_n2000_phi070_bidisperse",
               pkg.synth.lattice_config(2000, 0.70, seed=1, small_fraction=0.3), t=0.0)
    # dense monodisperse, non-zero sweep time
    sweep_case("sweep_n3000_phi085_mono",
               pkg.synth.lattice_config(3000, 0.85, seed=2), t=17.25)
    # lattice order (ids correlated with position), near liquid-hexatic density
    sweep_case("sweep_n2500_phi072_ordered",
               pkg.synth.lattice_config(2500, 0.72, seed=3, shuffle=False), t=3.0)
    # growth mode (setup sweep)
    g = pkg.synth.growth_config(2000, 0.70, seed=4)
    rng = np.random.default_rng(4)
    g["vr"] = g["vr"] * (0.5 + rng.random(g["n"]))
    g["rad"] = g["rad"] * (0.7 + 0.3 * rng.random(g["n"]))
    sweep_case("sweep_n2000_grow", g, t=g["t"], grow=True)


if want("sweep"):
    sweep_cases()


def weighted_case(name, cfg, dr, max_r):
    """The weighted g(r) family: the Bragg-peak search and the
    cos(k.r)-weighted pair correlation at that wave vector, from the reference."""
    n = cfg["n"]
    ref2 = ref
    ref2.setup(n, cfg["lx"], cfg["ly"], 0.0, cfg["x"], cfg["y"], cfg["vx"], cfg["vy"], cfg["rad"])
    phi = float(np.pi * (cfg["rad"] ** 2).sum() / (cfg["lx"] * cfg["ly"]))
    expected = float(np.sqrt(8 * np.pi * phi / np.sqrt(3)))   # save_pcf_boop, src/pcf.c:341
    k = ref2.bragg_peak(expected)["k"]
    bo = ref2.bond_order_pcf(dr, max_r, k)
    np.savez_compressed(HERE / f"{name}.npz", n=n, lx=cfg["lx"], ly=cfg["ly"], x=cfg["x"], y=cfg["y"],
                        rad=cfg["rad"], expected_bragg=expected, k=k, dr=dr, max_r=max_r,
                        g_r=bo["g_r"], g6_r=bo["g6_r"])
    print("wrote", name, "n =", n, "k =", k)


if want("weighted"):
    weighted_case("weighted_n1500_phi072", pkg.synth.lattice_config(1500, 0.72, seed=5), 2.0, 30.0)
    weighted_case("weighted_n1200_phi060_bidisperse",
                  pkg.synth.lattice_config(1200, 0.60, seed=6, small_fraction=0.3), 0.5, 20.0)


def voronoi_case(name, cfg, dr, q_max):
    """The Voronoi family and the structure factor:
    computeBOOPVoronoi, get_particle_voronoi_area/_perimeter, compute_g6_correlation,
    computeStructureFactor / computeVelocityStructureFactor, from the reference."""
    n, lx, ly = cfg["n"], cfg["lx"], cfg["ly"]
    ref.setup(n, lx, ly, 0.0, cfg["x"], cfg["y"], cfg["vx"], cfg["vy"], cfg["rad"])
    b = ref.boop_voronoi()
    a = ref.voronoi_area()
    max_r = min(lx, ly) / 2
    g6 = ref.g6_correlation(dr, max_r)
    sp = ref.structure_factor(q_max, velocity=False)
    sv = ref.structure_factor(q_max, velocity=True)
    np.savez_compressed(HERE / f"{name}.npz", n=n, lx=lx, ly=ly, x=cfg["x"], y=cfg["y"], vx=cfg["vx"],
                        vy=cfg["vy"], rad=cfg["rad"],
                        vor_q5=b["q5"], vor_q6=b["q6"], vor_q7=b["q7"], vor_q6_arg=b["q6_arg"],
                        vor_neighbors=b["neighbors"], vor_area=a["area"], vor_perimeter=a["perimeter"],
                        g6_dr=dr, g6_max_r=max_r, g6_corr=g6["g6_corr"], g6_counts=g6["counts"],
                        sq_qmax=q_max, sq_qx=sp["qx"], sq_qy=sp["qy"], sq_s=sp["s"], sq_s_velocity=sv["s"])
    print("wrote", name, "n =", n, "neighbours", np.bincount(b["neighbors"]))


def point_config(n, lx, ly, seed, jitter=None):
    """Positions only matter for this family (no overlap checks in the analysis
    functions): Poisson points (jitter None) or a square lattice with a large
    uniform jitter in units of the spacing -- many 5/7-fold and rarer 4/8-fold cells."""
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


if want("voronoi"):
    # dense nearly ordered monodisperse disks (every cell six-sided), a strongly
    # disordered lattice, and Poisson points (cells with 3 to 12 sides, big voids)
    voronoi_case("voronoi_n1500_phi085", pkg.synth.lattice_config(1500, 0.85, seed=8), 1.0, 0.8)
    voronoi_case("voronoi_n2025_jittered", point_config(2025, 97.0, 88.0, seed=7, jitter=0.9), 0.5, 1.0)
    voronoi_case("voronoi_n1500_poisson", point_config(1500, 80.0, 70.0, seed=9), 1.0, 0.5)

def tick_case(name, cfg, t0, t_new, T):
    """A whole thermostat tick: physicalQ + addNoise with the
    velocity-rescale branch, from the reference."""
    n = cfg["n"]
    ref.setup(n, cfg["lx"], cfg["ly"], t0, cfg["x"], cfg["y"], cfg["vx"], cfg["vy"], cfg["rad"])
    ref.predict_first()
    k = ref.tick_rescale(t_new, T)
    np.savez_compressed(HERE / f"{name}.npz", n=n, lx=cfg["lx"], ly=cfg["ly"], t=t0, t_new=t_new, T=T,
                        x=cfg["x"], y=cfg["y"], vx=cfg["vx"], vy=cfg["vy"], rad=cfg["rad"],
                        E_before=k["E_before"], tick_x=k["x"], tick_y=k["y"], tick_vx=k["vx"], tick_vy=k["vy"],
                        **{"tick_" + q: k[q] for q in ("t_cross", "dir", "t_coll", "partner", "ctype")})
    print("wrote", name, "n =", n, "E/N before", k["E_before"] / n)


if want("tick"):
    # the free flight must not create overlaps: short flights of a dilute-ish liquid
    tick_case("tick_n2000_phi060", pkg.synth.lattice_config(2000, 0.60, seed=10), 1.25, 1.2625, 0.7)
    tick_case("tick_n1500_phi045_bidisperse",
              pkg.synth.lattice_config(1500, 0.45, seed=11, small_fraction=0.3), 0.0, 0.03125, 1.6)
def normalize_case(name, cfg, drift, e_init):
    """normalizePhysicalQ (stopGrow / the initial conditions): a system with a centre-of-mass drift."""
    n = cfg["n"]
    vx, vy = cfg["vx"] + drift[0], cfg["vy"] + drift[1]
    ref.setup(n, cfg["lx"], cfg["ly"], 0.0, cfg["x"], cfg["y"], vx, vy, cfg["rad"])
    k = ref.normalize(e_init)
    np.savez_compressed(HERE / f"{name}.npz", n=n, lx=cfg["lx"], ly=cfg["ly"], x=cfg["x"], y=cfg["y"], vx=vx, vy=vy,
                        rad=cfg["rad"], e_init=e_init, norm_vx=k["vx"], norm_vy=k["vy"], px_before=k["px_before"],
                        py_before=k["py_before"], E_shifted=k["E_shifted"])
    print("wrote", name, "n =", n, "p/N before", k["px_before"] / n, k["py_before"] / n)


if want("normalize"):
    normalize_case("normalize_n2000_phi060", pkg.synth.lattice_config(2000, 0.60, seed=12), (0.37, -0.21), 1.0)
    normalize_case("normalize_n1500_einit", pkg.synth.lattice_config(1500, 0.45, seed=13), (-1.5, 0.02), 2.5)
ref.teardown()
