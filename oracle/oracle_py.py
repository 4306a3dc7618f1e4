"""ctypes wrappers around the two CPU checkers -- TEST INFRASTRUCTURE ONLY.

* ``Oracle``    -> oracle/liboracle.so      (our C restatement, edmd_oracle.c)
* ``Reference`` -> oracle/_ref/libedmd_ref.so (the unmodified reference + ref_shim.c)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.  The product never does.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ORACLE_SO = HERE / "liboracle.so"
REF_SO = HERE / "_ref" / "libedmd_ref.so"

MODE_NORMAL, MODE_GROW = 0, 1


def build(ref: bool = True) -> None:
    """Compile liboracle.so (always) and _ref (only if /root/reference exists)."""
    subprocess.run(["make", "-C", str(HERE), "liboracle.so"] + (["ref"] if ref else []),
                   check=True, capture_output=True)


class OBox(C.Structure):
    _fields_ = [("n", C.c_int), ("nx", C.c_int), ("ny", C.c_int), ("lx", C.c_double),
                ("ly", C.c_double), ("half_lx", C.c_double), ("half_ly", C.c_double),
                ("csx", C.c_double), ("csy", C.c_double), ("fx", C.c_double), ("fy", C.c_double)]


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


class Oracle:
    def __init__(self):
        if not ORACLE_SO.exists():
            build(ref=False)
        self.lib = lib = C.CDLL(str(ORACLE_SO))
        vp = C.c_void_p
        lib.oracle_box_init.argtypes = [C.POINTER(OBox), C.c_int, C.c_double, C.c_double]
        lib.oracle_box_init.restype = None
        lib.oracle_cells_from_coords.argtypes = [C.POINTER(OBox), C.c_int, vp, vp, vp]
        lib.oracle_cells_from_coords.restype = None
        lib.oracle_predict_all.argtypes = [C.POINTER(OBox), C.c_int, C.c_double] + [vp] * 7 + \
            [C.c_int] + [vp] * 6
        lib.oracle_free_fly.argtypes = [C.POINTER(OBox), C.c_int, C.c_int, C.c_double, vp,
                                        C.c_double] + [vp] * 6
        lib.oracle_free_fly.restype = None
        lib.oracle_pcf.argtypes = [C.POINTER(OBox), C.c_int, vp, vp, C.c_double, C.c_double,
                                   vp, vp, vp]
        lib.oracle_pcf_num_bins.argtypes = [C.c_double, C.c_double]
        lib.oracle_boop_cutoff.argtypes = [C.POINTER(OBox), C.c_int, vp, vp, vp, C.c_double] + [vp] * 5
        lib.oracle_boop_cutoff.restype = None

    def box(self, n, lx, ly) -> OBox:
        b = OBox()
        self.lib.oracle_box_init(C.byref(b), n, lx, ly)
        return b

    def cells(self, n, lx, ly, x, y):
        b = self.box(n, lx, ly)
        x, y = _f64(x), _f64(y)
        out = np.empty(2 * n, np.int32)
        self.lib.oracle_cells_from_coords(C.byref(b), n, _p(x), _p(y), _p(out))
        return out

    def predict_all(self, n, lx, ly, t, x, y, vx, vy, rad, vr=None, cell_xy=None,
                    mode=MODE_NORMAL):
        b = self.box(n, lx, ly)
        x, y, vx, vy, rad, vr = map(_f64, (x, y, vx, vy, rad, vr))
        cells = None if cell_xy is None else np.ascontiguousarray(cell_xy, np.int32).reshape(-1)
        tc = np.empty(n, np.float64)
        d = np.empty(n, np.uint8)
        tl = np.empty(n, np.float64)
        p = np.empty(n, np.int32)
        ct = np.empty(n, np.uint8)
        ov = np.full(2, -7, np.int32)
        rc = self.lib.oracle_predict_all(C.byref(b), n, float(t), _p(x), _p(y), _p(vx), _p(vy),
                                         _p(rad), _p(vr), _p(cells), mode, _p(tc), _p(d),
                                         _p(tl), _p(p), _p(ct), _p(ov))
        return dict(t_cross=tc, dir=d, t_coll=tl, partner=p, ctype=ct, overlap=ov, rc=rc)

    def free_fly(self, n, lx, ly, t_old, t_new, x, y, vx, vy, rad=None, vr=None,
                 mode=MODE_NORMAL, tp=None):
        b = self.box(n, lx, ly)
        x = _f64(x).copy()
        y = _f64(y).copy()
        rad = None if rad is None else _f64(rad).copy()
        vx, vy, vr, tp = map(_f64, (vx, vy, vr, tp))
        self.lib.oracle_free_fly(C.byref(b), n, mode, float(t_old), _p(tp), float(t_new),
                                 _p(x), _p(y), _p(vx), _p(vy), _p(rad), _p(vr))
        return dict(x=x, y=y, rad=rad)

    def pcf(self, n, lx, ly, x, y, dr, max_r):
        b = self.box(n, lx, ly)
        x, y = _f64(x), _f64(y)
        nb = self.lib.oracle_pcf_num_bins(dr, max_r)
        counts = np.zeros(max(nb, 1), np.uint64)
        g = np.zeros(max(nb, 1), np.float64)
        r = np.zeros(max(nb, 1), np.float64)
        self.lib.oracle_pcf(C.byref(b), n, _p(x), _p(y), dr, max_r, _p(counts), _p(g), _p(r))
        return dict(num_bins=nb, counts=counts[:nb], g_r=g[:nb], r=r[:nb])

    def boop_cutoff(self, n, lx, ly, x, y, r_c=2.5, cell_xy=None):
        b = self.box(n, lx, ly)
        x, y = _f64(x), _f64(y)
        cells = None if cell_xy is None else np.ascontiguousarray(cell_xy, np.int32).reshape(-1)
        q5, q6, q7, arg = (np.empty(n, np.float64) for _ in range(4))
        nb = np.empty(n, np.int32)
        self.lib.oracle_boop_cutoff(C.byref(b), n, _p(x), _p(y), _p(cells), r_c, _p(q5), _p(q6),
                                    _p(q7), _p(arg), _p(nb))
        return dict(q5=q5, q6=q6, q7=q7, q6_arg=arg, neighbors=nb)


    def bond_order_pcf(self, n, lx, ly, x, y, dr, max_r, k):
        b = self.box(n, lx, ly)
        x, y = _f64(x), _f64(y)
        nb = self.lib.oracle_pcf_num_bins(C.c_double(dr), C.c_double(max_r))
        counts = np.zeros(max(nb, 1), np.uint64)
        g = np.zeros(max(nb, 1), np.float64)
        g6 = np.zeros(max(nb, 1), np.float64)
        self.lib.oracle_bond_order_pcf(C.byref(b), n, _p(x), _p(y), C.c_double(dr), C.c_double(max_r),
                                       C.c_double(k[0]), C.c_double(k[1]), _p(counts), _p(g), _p(g6))
        return dict(num_bins=nb, counts=counts[:nb], g_r=g[:nb], g6_r=g6[:nb])

    def bragg_peak(self, n, lx, ly, x, y, expected_bragg):
        x, y = _f64(x), _f64(y)
        k = np.zeros(2, np.float64)
        s = C.c_double(0.0)
        self.lib.oracle_bragg_peak(n, _p(x), _p(y), C.c_double(lx), C.c_double(ly),
                                   C.c_double(expected_bragg), _p(k), C.byref(s))
        return dict(k=k, s_max=s.value)

    def tick_rescale(self, n, lx, ly, t_old, t_new, T, x, y, vx, vy, rad, cell_xy=None):
        """Thermostat tick, velocity-rescale branch: physicalQ's E (src/EDMD.c:5968-5997,
        sequential sum, unit masses), then addNoise :4889-4915 -- freeFly to t_new,
        `v /= sqrt(E/N/T)`, re-predict everything.  Free flight does not re-file
        particles: the sweep runs with the cell ids the particles had BEFORE the flight
        (cell_xy None = coordToCell of the old positions)."""
        if cell_xy is None:
            cell_xy = self.cells(n, lx, ly, x, y)
        vx, vy = _f64(vx), _f64(vy)
        terms = 0.5 * 1.0 * (vx * vx + vy * vy)
        E = float(np.cumsum(terms)[-1]) if n else 0.0      # left-to-right, like the loop
        s = np.sqrt(E / n / T)
        ff = self.free_fly(n, lx, ly, t_old, t_new, x, y, vx, vy, rad=rad)
        vx2, vy2 = vx / s, vy / s
        out = self.predict_all(n, lx, ly, t_new, ff["x"], ff["y"], vx2, vy2, rad, cell_xy=cell_xy)
        out.update(x=ff["x"], y=ff["y"], vx=vx2, vy=vy2, E_before=E, divisor=float(s))
        return out

    @staticmethod
    def normalize(vx, vy, e_init=1.0):
        """normalizePhysicalQ (src/EDMD.c:5723-5764, no circular wall, unit masses): physicalQ's
        sequential sums, `v -= p/(N*m)`, physicalQ again, `v /= sqrt(E/N/Einit)`."""
        vx, vy = _f64(vx).copy(), _f64(vy).copy()
        n = len(vx)
        px = float(np.cumsum(vx)[-1])            # left-to-right, like the loop
        py = float(np.cumsum(vy)[-1])
        vx -= px / (n * 1.0)
        vy -= py / (n * 1.0)
        E = float(np.cumsum(0.5 * 1.0 * (vx * vx + vy * vy))[-1])
        s = np.sqrt(E / n / e_init)
        return dict(vx=vx / s, vy=vy / s, px_before=px, py_before=py, E_shifted=E, divisor=float(s))

    def g6_correlation(self, n, lx, ly, x, y, psi_re, psi_im, dr, max_r):
        """Pair loop of compute_g6_correlation (src/pcf.c:189-228) for a given psi6."""
        b = self.box(n, lx, ly)
        x, y, pr, pi = map(_f64, (x, y, psi_re, psi_im))
        nb = self.lib.oracle_pcf_num_bins(C.c_double(dr), C.c_double(max_r))
        counts = np.zeros(max(nb, 1), np.uint64)
        g6 = np.zeros(max(nb, 1), np.float64)
        self.lib.oracle_g6_correlation(C.byref(b), n, _p(x), _p(y), _p(pr), _p(pi), C.c_double(dr),
                                       C.c_double(max_r), _p(counts), _p(g6))
        return dict(num_bins=nb, counts=counts[:nb], g6_corr=g6[:nb])

    def sq_grid(self, q_max, lx, ly):
        nqx, nqy = C.c_int(0), C.c_int(0)
        f = self.lib.oracle_sq_grid
        f(C.c_double(q_max), C.c_double(lx), C.c_double(ly), C.byref(nqx), C.byref(nqy), None, None)
        qx, qy = np.zeros(max(nqx.value, 1)), np.zeros(max(nqy.value, 1))
        f(C.c_double(q_max), C.c_double(lx), C.c_double(ly), C.byref(nqx), C.byref(nqy), _p(qx), _p(qy))
        return qx[:nqx.value], qy[:nqy.value]

    def structure_factor(self, n, lx, ly, x, y, q_max, vx=None, vy=None):
        """computeStructureFactor (vx None) / computeVelocityStructureFactor, src/struc.c:364-408."""
        qx, qy = self.sq_grid(q_max, lx, ly)
        x, y, vx, vy = map(_f64, (x, y, vx, vy))
        s = np.zeros(max(len(qx) * len(qy), 1), np.float64)
        self.lib.oracle_structure_factor(n, _p(x), _p(y), _p(vx), _p(vy), len(qx), _p(qx), len(qy),
                                         _p(qy), int(vx is not None), _p(s))
        return dict(qx=qx, qy=qy, s=s[:len(qx) * len(qy)].reshape(len(qx), len(qy)))

    # ---- Voronoi family.  The reference's algorithm is Fortune's sweep in the vendored
    # jc_voronoi.h (src/jc_voronoi.h, patched to double :19-22); what it outputs is the
    # Voronoi diagram of the point set built by get_particle_voronoi
    # (src/voronoi_edmd.c:33-121).  The restatement builds THAT point set as written and
    # takes the diagram from Qhull (scipy.spatial): Voronoi neighbours = Delaunay edges.
    @staticmethod
    def voronoi_points(n, lx, ly, x, y):
        """Particles + periodic images within 6.0 of the box edges, in the
        reference's order (src/voronoi_edmd.c:36-52, 63-111)."""
        bd = 6.0
        x, y = _f64(x)[:n], _f64(y)[:n]
        px, py, src = [x], [y], [np.arange(n)]
        # the reference appends, per particle, up to 8 images; the order of the
        # appended points does not matter for the diagram
        for cond, sx, sy in ((x < bd, lx, 0.0), (x > lx - bd, -lx, 0.0),
                             (y < bd, 0.0, ly), (y > ly - bd, 0.0, -ly),
                             ((x < bd) & (y < bd), lx, ly), ((x > lx - bd) & (y < bd), -lx, ly),
                             ((x < bd) & (y > ly - bd), lx, -ly),
                             ((x > lx - bd) & (y > ly - bd), -lx, -ly)):
            idx = np.nonzero(cond)[0]
            px.append(x[idx] + sx)
            py.append(y[idx] + sy)
            src.append(idx)
        return np.concatenate(px), np.concatenate(py), np.concatenate(src)

    def boop_voronoi(self, n, lx, ly, x, y):
        """computeBOOPVoronoi, src/boop.c:15-59."""
        from scipy.spatial import Delaunay
        px, py, _ = self.voronoi_points(n, lx, ly, x, y)
        tri = Delaunay(np.stack([px, py], axis=1))
        indptr, indices = tri.vertex_neighbor_vertices
        q5, q6, q7, arg = (np.zeros(n) for _ in range(4))
        nbr = np.zeros(n, np.int32)
        for i in range(n):
            nb = indices[indptr[i]:indptr[i + 1]]
            # `dx = e->neighbor->p.x - particles[index].x` :29-32 (image positions as built)
            th = np.arctan2(py[nb] - py[i], px[nb] - px[i])
            nbr[i] = len(nb)
            if len(nb):
                s5, s6, s7 = (np.exp(1j * k * th).sum() for k in (5, 6, 7))
                q5[i], q6[i], q7[i] = abs(s5) / len(nb), abs(s6) / len(nb), abs(s7) / len(nb)
                arg[i] = np.angle(s6)
        return dict(q5=q5, q6=q6, q7=q7, q6_arg=arg, neighbors=nbr)

    def voronoi_area(self, n, lx, ly, x, y):
        """get_particle_voronoi_area / _perimeter, src/voronoi_edmd.c:123-149."""
        from scipy.spatial import Voronoi
        px, py, _ = self.voronoi_points(n, lx, ly, x, y)
        vor = Voronoi(np.stack([px, py], axis=1))
        area, per = np.zeros(n), np.zeros(n)
        for i in range(n):
            reg = vor.regions[vor.point_region[i]]
            assert -1 not in reg and len(reg) >= 3
            v = vor.vertices[reg] - np.array([px[i], py[i]])
            v = v[np.argsort(np.arctan2(v[:, 1], v[:, 0]))]
            w = np.roll(v, -1, axis=0)
            area[i] = 0.5 * abs(np.sum(v[:, 0] * w[:, 1] - v[:, 1] * w[:, 0]))
            per[i] = np.sum(np.hypot(w[:, 0] - v[:, 0], w[:, 1] - v[:, 1]))
        return dict(area=area, perimeter=per)


class Reference:
    """The unmodified reference behind ref_shim.c.  Holds global state: one
    system at a time per process."""

    @staticmethod
    def available() -> bool:
        return REF_SO.exists()

    def __init__(self):
        if not REF_SO.exists():
            raise FileNotFoundError(f"{REF_SO} missing (built only where /root/reference exists)")
        self.lib = lib = C.CDLL(str(REF_SO))
        vp = C.c_void_p
        lib.ref_setup.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double] + [vp] * 7
        lib.ref_teardown.restype = None
        lib.ref_get_box.argtypes = [C.POINTER(C.c_int), C.POINTER(C.c_int)] + [C.POINTER(C.c_double)] * 4
        lib.ref_get_box.restype = None
        lib.ref_get_cells.argtypes = [vp]
        lib.ref_get_cells.restype = None
        lib.ref_predict_first.argtypes = [C.c_int] + [vp] * 5
        lib.ref_predict_first.restype = C.c_double
        lib.ref_repredict.argtypes = [vp] * 5
        lib.ref_repredict.restype = C.c_double
        lib.ref_calendar_only.restype = C.c_double
        lib.ref_free_fly.argtypes = [C.c_int, C.c_double, vp, vp, vp]
        lib.ref_free_fly.restype = C.c_double
        lib.ref_pcf.argtypes = [C.c_double, C.c_double, C.c_int, vp, vp, C.POINTER(C.c_int)]
        lib.ref_pcf.restype = C.c_double
        lib.ref_boop_cutoff.argtypes = [C.c_double] + [vp] * 5
        lib.ref_boop_cutoff.restype = C.c_double
        self.n = 0

    def setup(self, n, lx, ly, t, x, y, vx, vy, rad, vr=None, cell_xy=None):
        x, y, vx, vy, rad, vr = map(_f64, (x, y, vx, vy, rad, vr))
        cells = None if cell_xy is None else np.ascontiguousarray(cell_xy, np.int32).reshape(-1)
        rc = self.lib.ref_setup(n, lx, ly, float(t), _p(x), _p(y), _p(vx), _p(vy), _p(rad),
                                _p(vr), _p(cells))
        assert rc == 0
        self.n = n

    def teardown(self):
        self.lib.ref_teardown()

    def box(self):
        nx, ny = C.c_int(), C.c_int()
        d = [C.c_double() for _ in range(4)]
        self.lib.ref_get_box(C.byref(nx), C.byref(ny), *[C.byref(v) for v in d])
        return dict(nx=nx.value, ny=ny.value, csx=d[0].value, csy=d[1].value,
                    fx=d[2].value, fy=d[3].value)

    def cells(self):
        out = np.empty(2 * self.n, np.int32)
        self.lib.ref_get_cells(_p(out))
        return out

    def _outs(self):
        n = self.n
        return (np.empty(n, np.float64), np.empty(n, np.uint8), np.empty(n, np.float64),
                np.empty(n, np.int32), np.empty(n, np.uint8))

    def predict_first(self, grow=False):
        tc, d, tl, p, ct = self._outs()
        sec = self.lib.ref_predict_first(int(grow), _p(tc), _p(d), _p(tl), _p(p), _p(ct))
        return dict(t_cross=tc, dir=d, t_coll=tl, partner=p, ctype=ct, seconds=sec)

    def repredict(self):
        tc, d, tl, p, ct = self._outs()
        sec = self.lib.ref_repredict(_p(tc), _p(d), _p(tl), _p(p), _p(ct))
        return dict(t_cross=tc, dir=d, t_coll=tl, partner=p, ctype=ct, seconds=sec)

    def calendar_only(self) -> float:
        return self.lib.ref_calendar_only()

    def free_fly(self, t_new, grow=False):
        x, y, rad = (np.empty(self.n, np.float64) for _ in range(3))
        sec = self.lib.ref_free_fly(int(grow), float(t_new), _p(x), _p(y), _p(rad))
        return dict(x=x, y=y, rad=rad, seconds=sec)

    def pcf(self, dr, max_r, n_sub=0):
        nb_guess = int(max_r / dr) + 2
        g = np.zeros(nb_guess, np.float64)
        r = np.zeros(nb_guess, np.float64)
        nb = C.c_int(0)
        sec = self.lib.ref_pcf(dr, max_r, n_sub, _p(g), _p(r), C.byref(nb))
        return dict(num_bins=nb.value, g_r=g[:nb.value], r=r[:nb.value], seconds=sec)

    def boop_cutoff(self, r_c=2.5):
        n = self.n
        q5, q6, q7, arg = (np.empty(n, np.float64) for _ in range(4))
        nb = np.empty(n, np.int32)
        sec = self.lib.ref_boop_cutoff(r_c, _p(q5), _p(q6), _p(q7), _p(arg), _p(nb))
        return dict(q5=q5, q6=q6, q7=q7, q6_arg=arg, neighbors=nb, seconds=sec)

    def bond_order_pcf(self, dr, max_r, k):
        nb_guess = int(max_r / dr) + 2
        g = np.zeros(nb_guess, np.float64)
        g6 = np.zeros(nb_guess, np.float64)
        nb = C.c_int(0)
        self.lib.ref_bond_order_pcf.restype = C.c_double
        sec = self.lib.ref_bond_order_pcf(C.c_double(dr), C.c_double(max_r), C.c_double(k[0]),
                                          C.c_double(k[1]), _p(g), _p(g6), C.byref(nb))
        return dict(num_bins=nb.value, g_r=g[:nb.value], g6_r=g6[:nb.value], seconds=sec)

    def bragg_peak(self, expected_bragg):
        k = np.zeros(2, np.float64)
        self.lib.ref_bragg_peak.restype = C.c_double
        sec = self.lib.ref_bragg_peak(C.c_double(expected_bragg), _p(k))
        return dict(k=k, seconds=sec)

    def normalize(self, e_init):
        """normalizePhysicalQ() of the loaded system."""
        n = self.n
        vx, vy = np.empty(n, np.float64), np.empty(n, np.float64)
        px, py, e = C.c_double(0.0), C.c_double(0.0), C.c_double(0.0)
        self.lib.ref_normalize.restype = None
        self.lib.ref_normalize(C.c_double(e_init), _p(vx), _p(vy), C.byref(px), C.byref(py), C.byref(e))
        return dict(vx=vx, vy=vy, px_before=px.value, py_before=py.value, E_shifted=e.value)

    def tick_rescale(self, t_new, T):
        """physicalQ + addNoise (velocity-rescale branch) at time t_new."""
        n = self.n
        x, y, vx, vy = (np.empty(n, np.float64) for _ in range(4))
        tc, d, tl, p, ct = self._outs()
        e = C.c_double(0.0)
        f = self.lib.ref_tick_rescale
        f.restype = C.c_double
        sec = f(C.c_double(t_new), C.c_double(T), C.byref(e), _p(x), _p(y), _p(vx), _p(vy),
                _p(tc), _p(d), _p(tl), _p(p), _p(ct))
        return dict(t_cross=tc, dir=d, t_coll=tl, partner=p, ctype=ct, x=x, y=y, vx=vx, vy=vy,
                    E_before=e.value, seconds=sec)

    def boop_voronoi(self):
        n = self.n
        q5, q6, q7, arg = (np.empty(n, np.float64) for _ in range(4))
        nb = np.empty(n, np.int32)
        self.lib.ref_boop_voronoi.restype = C.c_double
        sec = self.lib.ref_boop_voronoi(_p(q5), _p(q6), _p(q7), _p(arg), _p(nb))
        return dict(q5=q5, q6=q6, q7=q7, q6_arg=arg, neighbors=nb, seconds=sec)

    def voronoi_area(self):
        area, per = np.empty(self.n), np.empty(self.n)
        self.lib.ref_voronoi_area.restype = C.c_double
        sec = self.lib.ref_voronoi_area(_p(area), _p(per))
        return dict(area=area, perimeter=per, seconds=sec)

    def g6_correlation(self, dr, max_r):
        nb_guess = int(max_r / dr) + 2
        g6 = np.zeros(nb_guess, np.float64)
        cnt = np.zeros(nb_guess, np.int32)
        nb = C.c_int(0)
        self.lib.ref_g6_correlation.restype = C.c_double
        sec = self.lib.ref_g6_correlation(C.c_double(dr), C.c_double(max_r), _p(g6), _p(cnt), C.byref(nb))
        return dict(num_bins=nb.value, g6_corr=g6[:nb.value], counts=cnt[:nb.value].astype(np.uint64),
                    seconds=sec)

    def structure_factor(self, q_max, velocity=False, lx=None, ly=None):
        cap = 4096
        qx, qy = np.zeros(cap), np.zeros(cap)
        nqx, nqy = C.c_int(0), C.c_int(0)
        f = self.lib.ref_structure_factor
        f.restype = C.c_double
        f(C.c_double(q_max), int(velocity), C.byref(nqx), C.byref(nqy), _p(qx), _p(qy), None)
        s = np.zeros(max(nqx.value * nqy.value, 1))
        sec = f(C.c_double(q_max), int(velocity), C.byref(nqx), C.byref(nqy), _p(qx), _p(qy), _p(s))
        return dict(qx=qx[:nqx.value].copy(), qy=qy[:nqy.value].copy(),
                    s=s[:nqx.value * nqy.value].reshape(nqx.value, nqy.value), seconds=sec)


def calendar_plan_oracle(t_cross, t_coll, paul_time, dt_paul, paul_n, actual_paul):
    """What 2N calls of the reference's addEventToQueue (src/EDMD.c:2144-2170), in the
    order of its batch loops (crossing 0, collision 0, crossing 1, ...;
    :2007-2012, :4909-4915), do to an empty calendar.  Test infrastructure: a plain
    Python/numpy restatement, event e = i (crossing) / N + i (collision)."""
    n = len(t_cross)
    bucket = np.empty(2 * n, np.int32)
    nxt = np.full(2 * n, -1, np.int32)
    prv = np.full(2 * n, -1, np.int32)
    head = np.full(paul_n + 1, -1, np.int32)
    n_tree = 0
    for i in range(n):
        for e, te in ((i, t_cross[i]), (n + i, t_coll[i])):
            dt = np.float64(te) - np.float64(paul_time)
            if dt < dt_paul:
                bucket[e] = -1
                n_tree += 1
                continue
            if dt >= np.float64(dt_paul) * paul_n:
                k = paul_n
            else:
                k = actual_paul + int(dt / np.float64(dt_paul))
                if k >= paul_n:
                    k -= paul_n
            bucket[e] = k
            nxt[e] = head[k]          # toAdd->rgt = eventPaul[k]
            if head[k] >= 0:
                prv[head[k]] = e      # toAdd->rgt->lft = toAdd
            head[k] = e               # eventPaul[k] = toAdd   (toAdd->lft = NULL)
    return dict(bucket=bucket, next=nxt, prev=prv, head=head, n_tree=n_tree)
