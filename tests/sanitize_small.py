import sys; sys.path.insert(0,'/root/repo')
import numpy as np, __graft_entry__ as e
pkg=e.load_package()
for (n,phi,sf) in [(3000,0.7,0.3),(800,0.3,0.0)]:
    c=pkg.synth.lattice_config(n,phi,3,small_fraction=sf)
    with pkg.EdmdCuda(c['n'],c['lx'],c['ly']) as ctx:
        ctx.upload(c['x'],c['y'],c['vx'],c['vy'],c['rad'],t=1.0)
        a=ctx.predict_all(); b=ctx.boop_cutoff(2.5); g=ctx.pcf(0.1,10.0)
        ctx.free_fly(1.5); s=ctx.download_state()
        ctx.set_growth(np.full(c['n'],0.01)); ctx.predict_all(mode=1, vr=np.full(c['n'],0.01))
    print('ok',c['n'])
# tiny boxes / overflow path
rng=np.random.default_rng(1)
n=4000; lx,ly=20.0,16.0
x=rng.random(n)*lx; y=rng.random(n)*ly
with pkg.EdmdCuda(n,lx,ly) as ctx:
    ctx.upload(x,y,rng.standard_normal(n),rng.standard_normal(n),np.full(n,1e-4),t=0.0)
    ctx.predict_all(); ctx.boop_cutoff(0.5)
print('ok overflow')
