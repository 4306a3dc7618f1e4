import sys, time
sys.path.insert(0, str(__import__("pathlib").Path(__file__).resolve().parents[2]))
import numpy as np
import __graft_entry__ as entry
pkg = entry.load_package()
B = pkg.binding
for n, phi in ((1000000, 0.70), (1000000, 0.85), (100000, 0.72)):
    c = pkg.synth.lattice_config(n, phi, seed=12345)
    with pkg.EdmdCuda(c["n"], c["lx"], c["ly"]) as ctx:
        ctx.upload(c["x"], c["y"], c["vx"], c["vy"], c["rad"], t=0.0)
        tot, main = ctx.bench(B.BENCH_VORONOI, warmup=2, iters=5, flush_bytes=256 << 20)
        print(f"K5 N={c['n']} phi={phi}: total {np.mean(tot)*1e3:.1f} us, k_voronoi {np.mean(main)*1e3:.1f} us")
        t0 = time.perf_counter(); s = ctx.structure_factor(0.3); t1 = time.perf_counter()
        print(f"  S(q) q_max=0.3 grid {s['s'].shape}: {(t1-t0)*1e3:.1f} ms")
        if n <= 100000:
            t0 = time.perf_counter(); g = ctx.g6_correlation(0.1, min(c["lx"], c["ly"]) / 2); t1 = time.perf_counter()
            print(f"  g6 correlation (Voronoi psi6 on device, all pairs): {(t1-t0)*1e3:.1f} ms")
