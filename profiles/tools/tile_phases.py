"""Timing experiments on k_tile_sweep (internal option 100): which part of the kernel costs what."""
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
c = pkg.synth.lattice_config(n, phi, 12345, shuffle=True)
with pkg.EdmdCuda(c["n"], c["lx"], c["ly"]) as ctx:
    ctx.upload(c["x"], c["y"], c["vx"], c["vy"], c["rad"], t=0.0)
    for name, dbg in (("all", 0), ("no screening loops", 1), ("no exact stage", 2), ("neither", 3), ("frame load only", 4), ("P1 without stores (+P2 frame load only)", 12), ("P1 without atomics (+P2 frame load only)", 20), ("P1 loads only", 28)):
        ctx.set_option(100, dbg)
        tot, main = ctx.bench(B.BENCH_SWEEP, warmup=3, iters=20, flush_bytes=256 << 20)
        print(f"{name:40s}: step {np.median(tot)*1e3:7.1f} us  P2 {np.median(main)*1e3:7.1f} us  P1 {np.median(tot-main)*1e3:6.1f} us", flush=True)
    ctx.set_option(100, 0)
    tot, _ = ctx.bench(B.BENCH_SWEEP, warmup=3, iters=20, flush_bytes=256 << 20, split=False)
    print(f"{'the chain as shipped (no event between)':40s}: step {np.median(tot)*1e3:7.1f} us (min {np.min(tot)*1e3:.1f})", flush=True)
