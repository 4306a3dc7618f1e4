"""g(r) at N = 4*10^6 (BASELINE configs[3], whole system on one GPU): one group per CTA (the histogram
leaves room for 2 CTAs = 16 warps per SM) against the automatic choice (4 groups sharing it: 32 warps)."""
import os, sys; sys.path.insert(0, ".")
import numpy as np, __graft_entry__ as e
pkg = e.load_package(); B = pkg.binding
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000000
c = pkg.synth.lattice_config(n, 0.7, 12345)
max_r = min(c["lx"], c["ly"]) / 2
with pkg.EdmdCuda(c["n"], c["lx"], c["ly"]) as ctx:
    ctx.upload(c["x"], c["y"], c["vx"], c["vy"], c["rad"], t=0.0)
    for g in ("1", "2", "4", ""):
        ctx.set_option(B.OPT_PCF_GROUPS, int(g or 0))
        t = ctx.bench(B.BENCH_PCF, dr=0.1, max_r=max_r, warmup=0, iters=1)[0][0]
        print("N", c["n"], "bins", int(max_r / 0.1), "groups", g or "auto", "ms", float(t), flush=True)
