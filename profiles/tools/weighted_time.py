"""Wall time of the weighted g(r) family through the ABI (cos(k.r)-weighted g(r), g6 correlation with given psi).
usage: python profiles/tools/weighted_time.py [N ...]"""
import sys
import time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from __graft_entry__ import load_package  # noqa: E402
pkg = load_package()
for n in [int(a) for a in sys.argv[1:]] or [100000, 1000000]:
    c = pkg.synth.lattice_config(n, 0.70, 12345, shuffle=True)
    max_r = min(c["lx"], c["ly"]) / 2
    k = np.array([3.4, 0.2])
    rng = np.random.default_rng(1)
    th = rng.uniform(0, 2 * np.pi, c["n"])
    with pkg.EdmdCuda(c["n"], c["lx"], c["ly"]) as ctx:
        ctx.upload(c["x"], c["y"], c["vx"], c["vy"], c["rad"], t=0.0)
        ctx.pcf_bond_order(0.1, 5.0, k)   # warm-up (module load, scratch)
        t0 = time.perf_counter()
        bo = ctx.pcf_bond_order(0.1, max_r, k)
        t1 = time.perf_counter()
        g6 = ctx.g6_correlation(0.1, max_r, np.cos(th), np.sin(th))
        t2 = time.perf_counter()
        plain = ctx.pcf(0.1, max_r)
        t3 = time.perf_counter()
        peak = ctx.bragg_peak(float(np.sqrt(8 * np.pi * 0.70 / np.sqrt(3))))
        t4 = time.perf_counter()
        sq = ctx.structure_factor(1.0)
        t5 = time.perf_counter()
    assert np.array_equal(bo["counts"], plain["counts"]) and np.array_equal(g6["counts"], plain["counts"])
    print(f"N={c['n']}: cos-weighted g(r) {t1 - t0:.3f} s, g6 correlation (given psi) {t2 - t1:.3f} s, "
          f"{int(plain['counts'].sum())} pairs in range, counts equal the plain g(r)'s; Bragg search {t4 - t3:.3f} s "
          f"(k = {peak['k']}, S = {peak['s_max']:.1f}); S(q), q_max = 1: {t5 - t4:.3f} s ({sq['s'].size} wave vectors)", flush=True)
