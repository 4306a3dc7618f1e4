"""Edge-case fixtures generated from the UNMODIFIED reference (oracle/_ref/libedmd_ref.so):
exact ties, NaN candidates, boxes with fewer than three cells per axis, host-supplied cells
that differ from coordToCell, the default bidisperse mixture.  One file, tests/golden/edges.npz,
keys `<case>/<array>`.

Run in the build container (where /root/reference is mounted):
    python tests/golden/make_edge_golden.py
The overlap case (c < -0.01) cannot be generated: the reference calls exit(3) there
(src/EDMD.c:2714-2716).
"""
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as entry  # noqa: E402
from oracle.oracle_py import Reference  # noqa: E402

pkg = entry.load_package()
ref = Reference()
out = {}
names = []


def case(name, c, t=0.0, cells=None, boop=True):
    n = c["n"]
    ref.setup(n, c["lx"], c["ly"], t, c["x"], c["y"], c["vx"], c["vy"], c["rad"], cell_xy=cells)
    first = ref.predict_first()
    again = ref.repredict()          # the reference's own addNoise() tick
    d = dict(n=n, lx=c["lx"], ly=c["ly"], t=t, x=c["x"], y=c["y"], vx=c["vx"], vy=c["vy"], rad=c["rad"],
             cells=ref.cells(), has_cells=int(cells is not None))
    for k in ("t_cross", "dir", "t_coll", "partner", "ctype"):
        d["first_" + k] = first[k]
        d["re_" + k] = again[k]
    if boop:
        b = ref.boop_cutoff(2.5)
        for k in ("q5", "q6", "q7", "q6_arg", "neighbors"):
            d["boop_" + k] = b[k]
    ref.teardown()
    for k, v in d.items():
        out[f"{name}/{k}"] = np.asarray(v)
    names.append(name)


# exact ties: mirrored partners with identical pair times, three id orders
x = np.array([11.0, 8.5, 13.5, 11.0, 11.0, 11.5]); y = np.array([11.0, 11.0, 11.0, 8.5, 13.5, 11.5])
vx = np.array([0.0, 1.0, -1.0, 0.0, 0.0, 0.0]); vy = np.array([0.0, 0.0, 0.0, 1.0, -1.0, 0.0])
rad = np.array([1.0, 1.0, 1.0, 1.0, 1.0, 0.01])
for k, perm in enumerate((np.arange(6), np.array([5, 4, 3, 2, 1, 0]), np.array([2, 0, 4, 1, 5, 3]))):
    case(f"ties{k}", dict(n=6, lx=20.0, ly=20.0, x=x[perm], y=y[perm], vx=vx[perm], vy=vy[perm], rad=rad[perm]))
# ties inside one cell: two small disks in the same cell, both hit at the same time
case("ties_same_cell", dict(n=3, lx=20.0, ly=20.0, x=np.array([9.0, 11.25, 11.25]), y=np.array([11.0, 10.5, 11.5]),
                            vx=np.array([1.0, 0.0, 0.0]), vy=np.zeros(3), rad=np.array([0.5, 0.5, 0.5])))
# identical velocities: b = 0, v2 = 0 -> 0/0 = NaN loses every comparison
case("nan", dict(n=3, lx=16.0, ly=16.0, x=np.array([5.0, 7.5, 11.0]), y=np.array([5.0, 5.0, 5.5]),
                 vx=np.array([0.25, 0.25, -1.0]), vy=np.array([0.5, 0.5, 0.0]), rad=np.ones(3)))
# a particle at rest on an axis: vx == 0 -> tX = +-inf or NaN
case("rest", dict(n=4, lx=24.0, ly=18.0, x=np.array([3.0, 9.0, 15.0, 21.0]), y=np.array([4.0, 9.0, 14.0, 9.0]),
                  vx=np.array([0.0, 1.0, 0.0, -0.5]), vy=np.array([0.0, 0.0, 2.0, 0.0]), rad=np.ones(4)))
# fewer than three cells per axis: the 3x3 scan revisits cells (and psi6 double counts)
for lx, ly in [(2.5, 2.5), (4.5, 9.0), (5.0, 2.2), (6.1, 6.1), (7.9, 40.0)]:
    rng = np.random.default_rng(int(lx * 10 + ly))
    n = max(2, int(lx * ly / 12))
    case(f"tiny_{lx}x{ly}", dict(n=n, lx=lx, ly=ly, x=rng.random(n) * lx, y=rng.random(n) * ly,
                                 vx=rng.standard_normal(n), vy=rng.standard_normal(n), rad=np.full(n, 0.05)), t=0.5)
# host-supplied cells: particles moved exactly onto their right-hand boundary keep the old cell
c = pkg.synth.lattice_config(3000, 0.70, seed=8)
n = c["n"]
ref.setup(n, c["lx"], c["ly"], 0.0, c["x"], c["y"], c["vx"], c["vy"], c["rad"])
cells = ref.cells().reshape(n, 2).copy()
box = ref.box()
ref.teardown()
edge = (cells[:, 0] + 1) * box["csx"]
idx = np.nonzero((edge - c["x"] < 0.03) & (cells[:, 0] + 1 < box["nx"]))[0]
assert idx.size > 5
c["x"] = c["x"].copy()
c["x"][idx] = edge[idx]
case("host_cells", c, t=2.0, cells=cells, boop=False)
# the reference's default mixture (30 % of radius 0.4), through the lean / tile sweep's two classes
case("bidisperse", pkg.synth.lattice_config(2500, 0.70, seed=9, small_fraction=0.3), t=1.0)

out["names"] = np.array(names)
np.savez_compressed(HERE / "edges.npz", **out)
print("wrote", HERE / "edges.npz", "cases:", names)
