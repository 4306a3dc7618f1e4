"""End-to-end tick through the single-process multi-GPU entry (edmd_cuda_create_mg): upload + predict_all with
whole-system host arrays, on 1 .. n GPUs of this box.  usage: python profiles/tools/mg_e2e.py [N]"""
import sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import torch
from __graft_entry__ import load_package  # noqa: E402
pkg = load_package()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000000
c = pkg.synth.lattice_config(n, 0.70, 12345, shuffle=True)
fx, fy = 1.0 / (c["lx"] / int(c["lx"] / 2)), 1.0 / (c["ly"] / int(c["ly"] / 2))
cells = np.stack([(c["x"] * fx).astype(np.int32), (c["y"] * fy).astype(np.int32)], 1)
import ctypes as C
B = pkg.binding
P = B._ptr
outs = dict(t_cross=np.zeros(c["n"]), dir=np.zeros(c["n"], np.uint8), t_coll=np.zeros(c["n"]),
            partner=np.zeros(c["n"], np.int32), ctype=np.zeros(c["n"], np.uint8))   # allocated and touched ONCE
ov = np.zeros(2, np.int32)
ref = None
for ndev in range(1, torch.cuda.device_count() + 1):
    with pkg.EdmdMg(c["n"], c["lx"], c["ly"], list(range(ndev))) as mg:
        ts = []
        for it in range(8):
            t0 = time.perf_counter()
            rc = mg.lib.edmd_cuda_mg_upload(mg._h, P(c["x"]), P(c["y"]), P(c["vx"]), P(c["vy"]), P(c["rad"]), P(cells), 0.0)
            assert rc == 0
            t1 = time.perf_counter()
            rc = mg.lib.edmd_cuda_mg_predict_all(mg._h, B.MODE_NORMAL, P(outs["t_cross"]), P(outs["dir"]), P(outs["t_coll"]),
                                                 P(outs["partner"]), P(outs["ctype"]), P(ov))
            assert rc == 0
            t2 = time.perf_counter()
            if it >= 3:
                ts.append((t1 - t0, t2 - t1))
        up, pr = np.median(ts, 0) * 1e3
        if ref is None:
            ref = {k: v.copy() for k, v in outs.items()}
        same = all(np.array_equal(outs[k], ref[k]) for k in ("t_cross", "t_coll", "partner", "dir"))
        print(f"N = {c['n']}  {ndev} GPU(s): upload (deal + copies) {up:.2f} ms, predict_all (exchange + sweep + gather) {pr:.2f} ms, "
              f"tick {up + pr:.2f} ms; same events as 1 GPU: {same}", flush=True)
