"""Timeline of worker 0 of k_cell_sweep (%globaltimer stamps, option 100 bit 32)."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from __graft_entry__ import load_package  # noqa: E402
pkg = load_package()
B = pkg.binding
c = pkg.synth.lattice_config(1000000, 0.70, 12345, shuffle=True)
with pkg.EdmdCuda(c["n"], c["lx"], c["ly"]) as ctx:
    ctx.upload(c["x"], c["y"], c["vx"], c["vy"], c["rad"], t=0.0)
    for dbg in (32, 32 + 4):
        ctx.set_option(100, dbg)
        ctx.bench(B.BENCH_SWEEP, warmup=3, iters=3, flush_bytes=256 << 20)
        ts = [ctx.stat(100 + k) for k in range(64)]
        t0 = ts[4]
        rel = lambda k: (ts[k] - t0) / 1e3
        print(f"dbg={dbg}: P1 first thread 0.0, P1 last {rel(6):.1f}, P2 after pdl wait {rel(7):.1f}, flags read {rel(8):.1f}, worker end {rel(15):.1f}")
        print(f"   last worker ends {rel(14):.1f}; most tiles per worker {ts[13]}; tiles in all (3+3 sweeps) {ts[12]}")
        print("   workers ending in [32+2k, 34+2k) us (cumulative over the sweeps run so far):", [int(v) for v in ts[40:64]])
        for it in range(6):
            print(f"   tile {it}: start {rel(16+4*it):7.1f}  frame landed {rel(17+4*it):7.1f}  converted {rel(18+4*it):7.1f}  swept {rel(19+4*it):7.1f}")
