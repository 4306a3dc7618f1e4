"""One K5 pass (Voronoi psi6 + cells) at N = 10^6 for ncu:
ncu --set full -k regex:k_vor -c 12 ... python profiles/tools/k5_prof.py"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import __graft_entry__ as entry
pkg = entry.load_package()
c = pkg.synth.lattice_config(1000000, 0.70, seed=12345)
with pkg.EdmdCuda(c["n"], c["lx"], c["ly"]) as ctx:
    ctx.upload(c["x"], c["y"], c["vx"], c["vy"], c["rad"], t=0.0)
    ctx.bench(pkg.binding.BENCH_VORONOI, warmup=1, iters=1, flush_bytes=256 << 20)
