"""Sweep timing for A/B builds of the partition kernel: P1 / P2 between events, and the chain as shipped.
usage: python profiles/tools/part_ab.py [N] [phi] [small_fraction]"""
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
c = pkg.synth.lattice_config(n, phi, 12345, small_fraction=sf, shuffle=True)
with pkg.EdmdCuda(c["n"], c["lx"], c["ly"]) as ctx:
    ctx.upload(c["x"], c["y"], c["vx"], c["vy"], c["rad"], t=0.0)
    tot, main = ctx.bench(B.BENCH_SWEEP, warmup=5, iters=30, flush_bytes=256 << 20)
    chain, _ = ctx.bench(B.BENCH_SWEEP, warmup=5, iters=30, flush_bytes=256 << 20, split=False)
    print(f"N={c['n']} phi={phi} sf={sf}: P1 {np.median(tot-main)*1e3:6.1f} us  P2 {np.median(main)*1e3:6.1f} us  "
          f"chain {np.median(chain)*1e3:6.1f} us (min {np.min(chain)*1e3:.1f})", flush=True)
