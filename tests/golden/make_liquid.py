"""Generates tests/golden/liquid_n10000_phi070.npz: a hard-disk LIQUID grown and
equilibrated by the UNMODIFIED reference itself (second input family).

    ./a.out -N 10000 --phi 0.70 -x 0 -v 1 -t <tmax>

run through oracle/_ref (its own main(): random points, growth phase, event loop),
the final state taken IN MEMORY (ref_run_main + ref_export_state in oracle/ref_shim.c;
the reference's dump file rounds the coordinates).  graphical-edmd_b200/synth.py tiles
it k x k (a periodic tiling of a periodic configuration keeps every distance >= contact)
into the N = 10^5 ... 4*10^6 liquids of the parity tests and of `bench.py --input liquid`.

Run in the build container only (needs /root/reference):  python tests/golden/make_liquid.py
"""
import ctypes as C
import os
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
N0, PHI, TMAX = 10000, 0.70, 400.0

CHILD = r"""
import ctypes as C, sys, numpy as np
lib = C.CDLL(sys.argv[1])
a = [b'a.out', b'-N', sys.argv[2].encode(), b'--phi', sys.argv[3].encode(), b'-x', b'0', b'-v', b'1',
     b'-t', sys.argv[4].encode()]
argv = (C.c_char_p * (len(a) + 1))(*a, None)
lib.ref_run_main(len(a), argv)
n = int(sys.argv[2])
arr = [np.zeros(n) for _ in range(5)]
box = np.zeros(3)
p = lambda v: v.ctypes.data_as(C.c_void_p)
got = lib.ref_export_state(n, *[p(v) for v in arr], p(box))
assert got == n, got
np.savez(sys.argv[5], x=arr[0], y=arr[1], vx=arr[2], vy=arr[3], rad=arr[4], lx=box[0], ly=box[1], t=box[2])
"""


def main():
    so = ROOT / "oracle" / "_ref" / "libedmd_ref.so"
    subprocess.run(["make", "-C", str(ROOT / "oracle"), "ref"], check=True, stdout=subprocess.DEVNULL)
    with tempfile.TemporaryDirectory() as d:
        os.makedirs(os.path.join(d, "dump"), exist_ok=True)
        raw = os.path.join(d, "state.npz")
        subprocess.run([sys.executable, "-c", CHILD, str(so), str(N0), str(PHI), str(TMAX), raw], cwd=d,
                       check=True, stdout=subprocess.DEVNULL)
        s = np.load(raw)
        x, y, rad = s["x"], s["y"], s["rad"]
        lx, ly = float(s["lx"]), float(s["ly"])
    n = len(x)
    # the growth phase leaves every radius at vr * t with its own rounding: ~N distinct values within 1e-15 of 1
    assert n == N0 and np.abs(rad - 1.0).max() < 1e-12, (n, rad.min(), rad.max())
    assert (x >= 0).all() and (x < lx).all() and (y >= 0).all() and (y < ly).all()
    # hard disks: no pair closer than contact (checked against all pairs, minimum image)
    worst = np.inf
    for i0 in range(0, n, 500):
        dx = x[i0:i0 + 500, None] - x[None, :]
        dy = y[i0:i0 + 500, None] - y[None, :]
        dx -= lx * np.rint(dx / lx)
        dy -= ly * np.rint(dy / ly)
        d2 = dx * dx + dy * dy
        d2[np.arange(len(d2)), i0 + np.arange(len(d2))] = np.inf
        worst = min(worst, float(d2.min()))
    assert worst >= 4.0 * (1 - 1e-12), worst
    out = Path(__file__).with_name("liquid_n10000_phi070.npz")
    np.savez_compressed(out, x=x, y=y, rad=rad, lx=lx, ly=ly, t_end=float(s["t"]),
                        made_by="reference main(): -N 10000 --phi 0.70 -x 0 -v 1 -t %g" % TMAX)
    print(out, n, lx, ly, "closest pair", worst ** 0.5, "phi", np.pi * n / (lx * ly), "distinct radii", len(np.unique(rad)))


if __name__ == "__main__":
    main()
