import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from __graft_entry__ import load_package
pkg = load_package(); B = pkg.binding
c = pkg.synth.lattice_config(1000000, 0.70, 12345, shuffle=True)
n = c["n"]
def pin(a):
    t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory(); return t, t.numpy()
keep = []
host = {}
for k in ("x","y","vx","vy","rad"):
    t, host[k] = pin(c[k]); keep.append(t)
fx, fy = 1.0/(c["lx"]/int(c["lx"]/2)), 1.0/(c["ly"]/int(c["ly"]/2))
cells = np.stack([(c["x"]*fx).astype(np.int32), (c["y"]*fy).astype(np.int32)], 1)
t, host["cells"] = pin(cells); keep.append(t)
outs = {}
for k, dt in (("t_cross", np.float64), ("dir", np.uint8), ("t_coll", np.float64), ("partner", np.int32)):
    t, outs[k] = pin(np.empty(n, dt)); keep.append(t)
# raw PCIe rates
d = torch.empty(40_000_000, dtype=torch.uint8, device="cuda"); h = torch.empty(40_000_000, dtype=torch.uint8).pin_memory()
for nm, f in (("H2D 40 MB", lambda: d.copy_(h, non_blocking=True)), ("D2H 21 MB", lambda: h[:21_000_000].copy_(d[:21_000_000], non_blocking=True))):
    for _ in range(3): f()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10): f()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
    print(f"{nm}: {dt*1e3:.3f} ms  ({(40 if 'H2D' in nm else 21)/dt/1e3:.1f} GB/s)")
with pkg.EdmdCuda(n, c["lx"], c["ly"]) as ctx:
    lib, hnd, P = ctx.lib, ctx._h, B._ptr
    ov = np.zeros(2, np.int32)
    ctx.upload(host["x"], host["y"], host["vx"], host["vy"], host["rad"], cell_xy=host["cells"], t=0.0)
    def up(): assert lib.edmd_cuda_upload(hnd, P(host["x"]), P(host["y"]), P(host["vx"]), P(host["vy"]), None, P(host["cells"]), 0.0) == 0
    def pred(): assert lib.edmd_cuda_predict_device(hnd, B.MODE_NORMAL) == 0; torch.cuda.synchronize()
    def fetch(): assert lib.edmd_cuda_fetch_predictions(hnd, P(outs["t_cross"]), P(outs["dir"]), P(outs["t_coll"]), P(outs["partner"]), None, P(ov)) == 0
    def allin(): assert lib.edmd_cuda_predict_all(hnd, B.MODE_NORMAL, None, P(outs["t_cross"]), P(outs["dir"]), P(outs["t_coll"]), P(outs["partner"]), None, P(ov)) == 0
    for _ in range(3): up(); pred(); fetch()
    for nm, f in (("upload", up), ("predict_device+sync", pred), ("fetch", fetch)):
        ts = []
        for _ in range(10):
            up() if nm != "upload" else None
            if nm == "fetch": pred()
            torch.cuda.synchronize(); t0 = time.perf_counter(); f(); ts.append(time.perf_counter() - t0)
        print(f"{nm}: {np.median(ts)*1e3:.3f} ms")
    ts = []
    for _ in range(10):
        t0 = time.perf_counter(); up(); allin(); ts.append(time.perf_counter() - t0)
    print(f"tick (upload + predict_all): {np.median(ts)*1e3:.3f} ms")
