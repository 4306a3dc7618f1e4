"""A/B timing of the sweep paths on resident state (CUDA events inside edmd_cuda_bench,
cold + clean L2 between steps): tile sweep vs five-kernel lean chain vs full FP64 path.
usage: python profiles/tools/sweep_ab.py [N] [phi] [small_fraction] [iters]"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from __graft_entry__ import load_package  # noqa: E402

pkg = load_package()
B = pkg.binding
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
phi = float(sys.argv[2]) if len(sys.argv) > 2 else 0.70
sf = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 20
c = pkg.synth.lattice_config(n, phi, 12345, small_fraction=sf, shuffle=True)
with pkg.EdmdCuda(c["n"], c["lx"], c["ly"]) as ctx:
    ctx.upload(c["x"], c["y"], c["vx"], c["vy"], c["rad"], t=0.0)
    for name, opts in (("tile", {}), ("lean5", {B.OPT_NO_TILE: 1}), ("full", {B.OPT_NO_LEAN: 1})):
        for o, v in opts.items():
            ctx.set_option(o, v)
        r0 = ctx.stat(B.STAT_EXACT_RESCANS)
        tot, main = ctx.bench(B.BENCH_SWEEP, warmup=3, iters=iters, flush_bytes=256 << 20)
        resc = (ctx.stat(B.STAT_EXACT_RESCANS) - r0) / (iters + 3)
        print(f"{name:6s} N={c['n']} phi={phi} sf={sf}: step {np.median(tot)*1e3:7.1f} us (min {np.min(tot)*1e3:.1f}) "
              f"K1 {np.median(main)*1e3:7.1f} us  K0 {np.median(tot-main)*1e3:6.1f} us  rescans/sweep {resc:.0f}",
              flush=True)
