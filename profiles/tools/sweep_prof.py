"""Three sweeps on resident state (for ncu): python profiles/tools/sweep_prof.py [N] [phi] [sf] [opt=value ...]"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from __graft_entry__ import load_package  # noqa: E402

pkg = load_package()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
phi = float(sys.argv[2]) if len(sys.argv) > 2 else 0.70
sf = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
c = pkg.synth.lattice_config(n, phi, 12345, small_fraction=sf, shuffle=True)
with pkg.EdmdCuda(c["n"], c["lx"], c["ly"]) as ctx:
    for kv in sys.argv[4:]:
        k, v = kv.split("=")
        ctx.set_option(int(k), int(v))
    ctx.upload(c["x"], c["y"], c["vx"], c["vy"], c["rad"], t=0.0)
    for _ in range(3):
        ctx.predict_device()
    ctx.fetch_predictions()
    print("lean sweeps", ctx.stat(pkg.binding.STAT_LEAN_SWEEPS))
