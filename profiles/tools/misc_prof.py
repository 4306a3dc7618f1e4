"""One pass through the kernels beside the sweep (for ncu): upload pack, event unpack, calendar plan,
free flight, thermostat (kinetic sums, rescale, normalize), Voronoi cells."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from __graft_entry__ import load_package  # noqa: E402
pkg = load_package()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
c = pkg.synth.lattice_config(n, 0.70, 12345, shuffle=True)
with pkg.EdmdCuda(c["n"], c["lx"], c["ly"]) as ctx:
    for _ in range(2):
        ctx.upload(c["x"], c["y"], c["vx"], c["vy"], c["rad"], t=0.0)
        ctx.predict_all()
        ctx.calendar_plan(0.0, 5.0 / c["n"], c["n"], 0)
        ctx.free_fly(0.01)
        ctx.rescale_velocities(1.0)
        ctx.normalize_velocities(1.0)
        ctx.boop_voronoi()
print("done")
