"""psi6 timing: tile kernel (with / without the partition) vs row kernel. python profiles/tools/boop_ab.py [N] [phi]"""
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
    for name, off in (("tile", 0), ("rows", 1)):
        ctx.set_option(B.OPT_NO_TILE_BOOP, off)
        tot, main = ctx.bench(B.BENCH_BOOP, dr=2.5, warmup=3, iters=20, flush_bytes=256 << 20)
        print(f"psi6 {name}: total {np.median(tot)*1e3:7.1f} us  kernel {np.median(main)*1e3:7.1f} us", flush=True)
