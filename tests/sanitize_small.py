"""compute-sanitizer target: small sweeps (one class, two classes, slabs on one device), psi6, normalize."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from __graft_entry__ import load_package  # noqa: E402
pkg = load_package()
for sf in (0.0, 0.3):
    c = pkg.synth.lattice_config(30000, 0.72, seed=3, small_fraction=sf, shuffle=True)
    with pkg.EdmdCuda(c["n"], c["lx"], c["ly"]) as ctx:
        ctx.upload(c["x"], c["y"], c["vx"], c["vy"], c["rad"], t=0.0)
        a = ctx.predict_all()
        b = ctx.predict_all()
        assert np.array_equal(a["t_coll"], b["t_coll"])
        ctx.boop_cutoff(2.5)
        ctx.normalize_velocities(1.0)
        ctx.predict_all()
    with pkg.EdmdMg(c["n"], c["lx"], c["ly"], [0, 0, 0]) as mg:
        mg.upload(c["x"], c["y"], c["vx"], c["vy"], c["rad"], t=0.0)
        g = mg.predict_all()
        assert np.array_equal(g["t_coll"], a["t_coll"]) and np.array_equal(g["partner"], a["partner"])
        mg.boop_cutoff(2.5)

# ---- the other paths: growth sweep, free flight, g(r), tiny boxes / overflow -> full path
for (n,phi,sf) in [(3000,0.7,0.3),(800,0.3,0.0)]:
    c=pkg.synth.lattice_config(n,phi,3,small_fraction=sf)
    with pkg.EdmdCuda(c['n'],c['lx'],c['ly']) as ctx:
        ctx.upload(c['x'],c['y'],c['vx'],c['vy'],c['rad'],t=1.0)
        a=ctx.predict_all(); b=ctx.boop_cutoff(2.5); g=ctx.pcf(0.1,10.0)
        ctx.free_fly(1.5); s=ctx.download_state()
        ctx.set_growth(np.full(c['n'],0.01)); ctx.predict_all(mode=1, vr=np.full(c['n'],0.01))
    print('ok',c['n'])
# the weighted g(r) family and the sums on grids of wave vectors
c = pkg.synth.lattice_config(900, 0.70, 5)   # (small: the pair and grid kernels are O(N^2) under the sanitizer)
with pkg.EdmdCuda(c['n'], c['lx'], c['ly']) as ctx:
    ctx.upload(c['x'], c['y'], c['vx'], c['vy'], c['rad'], t=0.0)
    pk = ctx.bragg_peak(3.2)
    bo = ctx.pcf_bond_order(0.1, min(c['lx'], c['ly']) / 2, pk['k'])
    g6 = ctx.g6_correlation(0.25, 20.0, np.cos(c['x']), np.sin(c['x']))
    ctx.structure_factor(1.0)
    ctx.structure_factor(1.0, velocity=True)
    assert np.array_equal(bo['counts'], ctx.pcf(0.1, min(c['lx'], c['ly']) / 2)['counts'])
print('ok weighted')
# tiny boxes / overflow path
rng=np.random.default_rng(1)
n=4000; lx,ly=20.0,16.0
x=rng.random(n)*lx; y=rng.random(n)*ly
with pkg.EdmdCuda(n,lx,ly) as ctx:
    ctx.upload(x,y,rng.standard_normal(n),rng.standard_normal(n),np.full(n,1e-4),t=0.0)
    ctx.predict_all(); ctx.boop_cutoff(0.5)
print('ok overflow')
print("sanitize target done")
