import sys; sys.path.insert(0,".")
import numpy as np, __graft_entry__ as e
pkg=e.load_package(); B=pkg.binding
c=pkg.synth.lattice_config(400000,0.7,12345)
with pkg.EdmdCuda(c["n"],c["lx"],c["ly"]) as ctx:
    ctx.upload(c["x"],c["y"],c["vx"],c["vy"],c["rad"],t=0.0)
    tot,main=ctx.bench(B.BENCH_PCF, dr=0.1, max_r=min(c["lx"],c["ly"])/2, warmup=0, iters=1)
    print(tot)
    print("rsqrt worst rel err", ctx.selftest_rsqrt(), "exact-path pairs", ctx.stat(B.STAT_PCF_EXACT_PAIRS))
