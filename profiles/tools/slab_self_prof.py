"""One rank of the 2-GPU slab workload on ONE GPU, its two neighbours being itself
(halo_connect(None, None)): lets ncu list the kernels of a multi-GPU step
(ncu must never wrap a multi-rank command).  Usage:
    ncu --metrics gpu__time_duration.sum ... python profiles/tools/slab_self_prof.py"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import __graft_entry__ as entry  # noqa: E402

pkg = entry.load_package()
B, slab = pkg.binding, pkg.slab
world, n_per = 2, 1000000
cfg = pkg.synth.lattice_config(n_per * world, 0.70, 12345)
N, lx, ly = cfg["n"], cfg["lx"], cfg["ly"]
fx, fy = 1.0 / (lx / int(lx / 2)), 1.0 / (ly / int(ly / 2))
cells = np.stack([(cfg["x"] * fx).astype(np.int32), (cfg["y"] * fy).astype(np.int32)], 1)
sr = slab.SlabRank(pkg, N, lx, ly, 0, world, 0)
sr.ctx.halo_export(sr.halo_capacity)
sr.ctx.halo_connect(None, None)
sr.p2p = True
sr.load_owned(cfg, cells, 0.0)
sr.exchange(None)
tot, main = sr.ctx.bench(B.BENCH_SWEEP, warmup=3, iters=int(sys.argv[1]) if len(sys.argv) > 1 else 5,
                         flush_bytes=256 << 20)
print("slab self-exchange: step %.1f us, K1 %.1f us" % (1e3 * tot.mean(), 1e3 * main.mean()))
# timeline of the last step (%globaltimer, ns): send start/end, receive start/end, partition first/last block start, end, sweep start
sr.ctx.set_option(100, 32)
sr.ctx.bench(B.BENCH_SWEEP, warmup=1, iters=1, flush_bytes=256 << 20)
ts = [sr.ctx.stat(100 + k) for k in range(8)]
t0 = min(t for t in ts if t)
names = ["send start", "send end", "recv start", "recv end", "partition first block", "partition last block", "partition end", "sweep start"]
for nme, t in zip(names, ts):
    print("  %-24s %8.1f us" % (nme, (t - t0) / 1e3))
sr.close()
