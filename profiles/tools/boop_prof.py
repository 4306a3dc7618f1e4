"""Three psi6 frames (partition + tile kernel + unpack) for ncu."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from __graft_entry__ import load_package  # noqa: E402
pkg = load_package()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
c = pkg.synth.lattice_config(n, 0.70, 12345, shuffle=True)
with pkg.EdmdCuda(c["n"], c["lx"], c["ly"]) as ctx:
    ctx.upload(c["x"], c["y"], c["vx"], c["vy"], c["rad"], t=0.0)
    tot, main = ctx.bench(pkg.binding.BENCH_BOOP, dr=2.5, warmup=1, iters=2, flush_bytes=0)
    print(tot, main)
