"""ctypes mirror of include/edmd_cuda.h (the drop-in C ABI).

Nothing here computes anything: every method forwards to libedmd_cuda.so and
raises EdmdError when the library reports a failure.  If the library is missing
the import of this module still works but ``load_library`` raises -- there is
no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "libedmd_cuda.so"

MODE_NORMAL, MODE_GROW = 0, 1
EV_CELLCROSS, EV_COLLISION = 0, 1
EINVAL, ESTATE, EOVERLAP, ECELL, ENOMEM = 1, 2, 3, 4, 5
BENCH_SWEEP, BENCH_FREEFLY, BENCH_BOOP, BENCH_PCF, BENCH_VORONOI = 0, 1, 2, 3, 4
OPT_FORCE_GENERIC = 1
OPT_NO_LEAN = 2
OPT_NO_PDL = 3
OPT_PCF_LEGACY = 4
OPT_PCF_GROUPS = 5
OPT_NO_TILE = 6
OPT_NO_TILE_BOOP = 7
EPLAN = 6
STAT_EXACT_RESCANS = 1
STAT_LEAN_SWEEPS = 2
STAT_PCF_EXACT_PAIRS = 3
STAT_PCF_SKIPPED_TILE_PAIRS = 4
STAT_LEAN_DECLINES = 5
STAT_LEAN_ELIGIBLE = 6

# every symbol include/edmd_cuda.h declares (+ edmd_cuda_bench from the bench-only header)
SYMBOLS = [
    "edmd_cuda_create", "edmd_cuda_destroy", "edmd_cuda_last_error",
    "edmd_cuda_get_box", "edmd_cuda_launch_count", "edmd_cuda_upload",
    "edmd_cuda_upload_aos", "edmd_cuda_predict_all", "edmd_cuda_predict_device",
    "edmd_cuda_fetch_predictions", "edmd_cuda_set_growth", "edmd_cuda_free_fly",
    "edmd_cuda_download_state", "edmd_cuda_pcf", "edmd_cuda_boop_cutoff",
    "edmd_cuda_bench", "edmd_cuda_set_option", "edmd_cuda_get_stat",
    "edmd_cuda_host_alloc", "edmd_cuda_host_free",
    "edmd_cuda_create_slab", "edmd_cuda_upload_owned", "edmd_cuda_halo_pack",
    "edmd_cuda_halo_append", "edmd_cuda_get_counts", "edmd_cuda_pcf_device",
    "edmd_cuda_halo_export", "edmd_cuda_halo_connect", "edmd_cuda_halo_exchange",
    "edmd_cuda_exchange_predict_device",
    "edmd_cuda_calendar_plan", "edmd_cuda_pcf_bond_order", "edmd_cuda_bragg_peak",
    "edmd_cuda_boop_voronoi", "edmd_cuda_voronoi_cells", "edmd_cuda_g6_correlation",
    "edmd_cuda_structure_factor", "edmd_cuda_kinetic", "edmd_cuda_rescale_velocities",
    "edmd_cuda_selftest_rsqrt", "edmd_cuda_langevin_kick",
    "edmd_cuda_normalize_velocities", "edmd_cuda_shift_scale_velocities",
    "edmd_cuda_create_mg", "edmd_cuda_destroy_mg", "edmd_cuda_mg_last_error", "edmd_cuda_mg_get_box",
    "edmd_cuda_mg_slab_sizes", "edmd_cuda_mg_upload", "edmd_cuda_mg_predict_all", "edmd_cuda_mg_boop_cutoff",
    "edmd_cuda_mg_pcf",
]
EVORONOI = 7
HALO_RECORD_BYTES = 48


class Box(C.Structure):
    _fields_ = [("n", C.c_int32), ("nxcells", C.c_int32), ("nycells", C.c_int32),
                ("lx", C.c_double), ("ly", C.c_double), ("half_lx", C.c_double),
                ("half_ly", C.c_double), ("cellx_size", C.c_double),
                ("celly_size", C.c_double), ("cellx_fac", C.c_double),
                ("celly_fac", C.c_double), ("dt_paul", C.c_double)]


class EdmdError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"edmd_cuda error {code}: {msg}")
        self.code = code


_lib = None


def load_library(path: os.PathLike | None = None) -> C.CDLL:
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = Path(path) if path else LIB_PATH
    if not p.exists():
        raise FileNotFoundError(
            f"{p} not built -- run `make cuda` (or __graft_entry__.build()); "
            "there is no CPU fallback for the hot path")
    lib = C.CDLL(str(p))
    dp, ip, u8p = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_uint8)
    vp = C.c_void_p
    lib.edmd_cuda_create.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.POINTER(vp)]
    lib.edmd_cuda_destroy.argtypes = [vp]
    lib.edmd_cuda_destroy.restype = None
    lib.edmd_cuda_last_error.argtypes = [vp]
    lib.edmd_cuda_last_error.restype = C.c_char_p
    lib.edmd_cuda_get_box.argtypes = [vp, C.POINTER(Box)]
    lib.edmd_cuda_launch_count.argtypes = [vp]
    lib.edmd_cuda_launch_count.restype = C.c_uint64
    lib.edmd_cuda_upload.argtypes = [vp, vp, vp, vp, vp, vp, vp, C.c_double]
    lib.edmd_cuda_upload_aos.argtypes = [vp, vp] + [C.c_size_t] * 7 + [C.c_double]
    lib.edmd_cuda_predict_all.argtypes = [vp, C.c_int, vp, vp, vp, vp, vp, vp, vp]
    lib.edmd_cuda_predict_device.argtypes = [vp, C.c_int]
    lib.edmd_cuda_fetch_predictions.argtypes = [vp, vp, vp, vp, vp, vp, vp]
    lib.edmd_cuda_set_growth.argtypes = [vp, vp]
    lib.edmd_cuda_free_fly.argtypes = [vp, C.c_int, C.c_double]
    lib.edmd_cuda_download_state.argtypes = [vp, vp, vp, vp, vp, vp]
    lib.edmd_cuda_pcf.argtypes = [vp, C.c_double, C.c_double, vp, vp, C.POINTER(C.c_int)]
    lib.edmd_cuda_boop_cutoff.argtypes = [vp, C.c_double, vp, vp, vp, vp, vp, vp]
    lib.edmd_cuda_bench.argtypes = [vp, C.c_int, C.c_int, C.c_double, C.c_double,
                                    C.c_int, C.c_int, C.c_size_t, vp, vp]
    lib.edmd_cuda_create_slab.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int,
                                          C.POINTER(vp)]
    lib.edmd_cuda_upload_owned.argtypes = [vp, C.c_int, vp, vp, vp, vp, vp, vp, vp, C.c_double]
    lib.edmd_cuda_halo_pack.argtypes = [vp, C.c_int, vp, C.c_int, C.POINTER(C.c_int)]
    lib.edmd_cuda_halo_append.argtypes = [vp, C.c_int, vp, C.c_int]
    lib.edmd_cuda_get_counts.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.edmd_cuda_halo_export.argtypes = [vp, C.c_int, C.c_char_p]
    lib.edmd_cuda_halo_connect.argtypes = [vp, C.c_char_p, C.c_char_p]
    lib.edmd_cuda_halo_exchange.argtypes = [vp]
    lib.edmd_cuda_exchange_predict_device.argtypes = [vp, C.c_int]
    lib.edmd_cuda_pcf_device.argtypes = [vp, vp, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, vp,
                                         C.POINTER(C.c_int)]
    lib.edmd_cuda_host_alloc.argtypes = [C.POINTER(vp), C.c_size_t]
    lib.edmd_cuda_host_free.argtypes = [vp]
    lib.edmd_cuda_host_free.restype = None
    lib.edmd_cuda_set_option.argtypes = [vp, C.c_int, C.c_int]
    lib.edmd_cuda_pcf_bond_order.argtypes = [vp, C.c_double, C.c_double, vp, vp, vp, vp, C.POINTER(C.c_int)]
    lib.edmd_cuda_bragg_peak.argtypes = [vp, C.c_double, vp, C.POINTER(C.c_double)]
    lib.edmd_cuda_calendar_plan.argtypes = [vp, C.c_double, C.c_double, C.c_int, C.c_int, vp, vp, vp, vp,
                                            C.POINTER(C.c_int32)]
    lib.edmd_cuda_get_stat.argtypes = [vp, C.c_int, C.POINTER(C.c_uint64)]
    lib.edmd_cuda_selftest_rsqrt.argtypes = [vp, C.POINTER(C.c_double)]
    lib.edmd_cuda_langevin_kick.argtypes = [vp, C.c_double, C.c_double, C.c_double, C.c_uint32, C.c_uint32]
    lib.edmd_cuda_kinetic.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.edmd_cuda_create_mg.argtypes = [C.c_int, C.POINTER(C.c_int), C.c_int, C.c_double, C.c_double, C.POINTER(vp)]
    lib.edmd_cuda_destroy_mg.argtypes = [vp]
    lib.edmd_cuda_destroy_mg.restype = None
    lib.edmd_cuda_mg_last_error.argtypes = [vp]
    lib.edmd_cuda_mg_last_error.restype = C.c_char_p
    lib.edmd_cuda_mg_get_box.argtypes = [vp, C.POINTER(Box)]
    lib.edmd_cuda_mg_slab_sizes.argtypes = [vp, C.POINTER(C.c_int)]
    lib.edmd_cuda_mg_upload.argtypes = [vp, vp, vp, vp, vp, vp, vp, C.c_double]
    lib.edmd_cuda_mg_predict_all.argtypes = [vp, C.c_int, vp, vp, vp, vp, vp, vp]
    lib.edmd_cuda_mg_boop_cutoff.argtypes = [vp, C.c_double, vp, vp, vp, vp, vp, C.POINTER(C.c_double)]
    lib.edmd_cuda_mg_pcf.argtypes = [vp, vp, vp, C.c_double, C.c_double, vp, vp, C.POINTER(C.c_int)]
    lib.edmd_cuda_normalize_velocities.argtypes = [vp, C.c_double] + [C.POINTER(C.c_double)] * 4
    lib.edmd_cuda_shift_scale_velocities.argtypes = [vp, C.c_double, C.c_double, C.c_double]
    lib.edmd_cuda_rescale_velocities.argtypes = [vp, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.edmd_cuda_boop_voronoi.argtypes = [vp, vp, vp, vp, vp, vp, vp]
    lib.edmd_cuda_voronoi_cells.argtypes = [vp, vp, vp, vp]
    lib.edmd_cuda_g6_correlation.argtypes = [vp, C.c_double, C.c_double, vp, vp, vp, vp, C.POINTER(C.c_int)]
    lib.edmd_cuda_structure_factor.argtypes = [vp, C.c_double, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                               vp, vp, vp, vp, vp]
    for name in SYMBOLS:
        fn = getattr(lib, name)
        if fn.restype is C.c_int:
            pass
    if path is None:
        _lib = lib
    return lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f64(a, n):
    a = np.ascontiguousarray(a, dtype=np.float64)
    assert a.shape == (n,), (a.shape, n)
    return a


class EdmdCuda:
    """One context = one GPU + one particle system (edmd_cuda_create)."""

    def __init__(self, n: int, lx: float, ly: float, device: int = 0, slab_rows=None):
        """slab_rows=(row_lo, row_hi): a slab context of capacity n (edmd_cuda_create_slab)."""
        self.lib = load_library()
        self.n = int(n)
        self.slab_rows = slab_rows
        h = C.c_void_p()
        if slab_rows is None:
            rc = self.lib.edmd_cuda_create(device, self.n, lx, ly, C.byref(h))
        else:
            rc = self.lib.edmd_cuda_create_slab(device, self.n, lx, ly, int(slab_rows[0]),
                                                int(slab_rows[1]), C.byref(h))
        self._h = h
        if rc != 0:
            msg = self.lib.edmd_cuda_last_error(h).decode() if h else "create failed"
            if h:
                self.lib.edmd_cuda_destroy(h)
            self._h = None
            raise EdmdError(rc, msg)

    def close(self):
        if getattr(self, "_h", None):
            self.lib.edmd_cuda_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc, allow=()):
        if rc != 0 and rc not in allow:
            raise EdmdError(rc, self.lib.edmd_cuda_last_error(self._h).decode())
        return rc

    @property
    def box(self) -> Box:
        b = Box()
        self._check(self.lib.edmd_cuda_get_box(self._h, C.byref(b)))
        return b

    @property
    def launches(self) -> int:
        return int(self.lib.edmd_cuda_launch_count(self._h))

    def set_option(self, option, value):
        self._check(self.lib.edmd_cuda_set_option(self._h, option, int(value)))

    def stat(self, which=STAT_EXACT_RESCANS) -> int:
        v = C.c_uint64(0)
        self._check(self.lib.edmd_cuda_get_stat(self._h, which, C.byref(v)))
        return int(v.value)

    def upload(self, x, y, vx, vy, rad, cell_xy=None, t=0.0):
        n = self.n
        # rad=None: radii unchanged since the last upload (kept resident)
        arrs = [None if a is None else _f64(a, n) for a in (x, y, vx, vy, rad)]
        cells = None
        if cell_xy is not None:
            cells = np.ascontiguousarray(cell_xy, dtype=np.int32).reshape(-1)
            assert cells.size == 2 * n
        self._check(self.lib.edmd_cuda_upload(self._h, *[_ptr(a) for a in arrs], _ptr(cells), float(t)))

    # ---- slab (multi-GPU) ----------------------------------------------------
    def upload_owned(self, x, y, vx, vy, rad, cell_xy, global_id, t=0.0):
        n = len(x)
        arrs = [_f64(a, n) for a in (x, y, vx, vy, rad)]
        cells = np.ascontiguousarray(cell_xy, dtype=np.int32).reshape(-1)
        gid = np.ascontiguousarray(global_id, dtype=np.int32)
        assert cells.size == 2 * n and gid.size == n
        self._check(self.lib.edmd_cuda_upload_owned(self._h, n, *[_ptr(a) for a in arrs], _ptr(cells),
                                                    _ptr(gid), float(t)))
        self.n = n   # outputs cover the owned particles

    def halo_pack(self, side: int, dev_ptr: int, capacity: int) -> int:
        cnt = C.c_int(0)
        self._check(self.lib.edmd_cuda_halo_pack(self._h, side, C.c_void_p(dev_ptr), capacity, C.byref(cnt)))
        return cnt.value

    def halo_append(self, side: int, dev_ptr: int, count: int):
        self._check(self.lib.edmd_cuda_halo_append(self._h, side, C.c_void_p(dev_ptr), count))

    def halo_export(self, halo_capacity: int) -> bytes:
        buf = C.create_string_buffer(64)
        self._check(self.lib.edmd_cuda_halo_export(self._h, halo_capacity, buf))
        return buf.raw

    def halo_connect(self, lower: bytes | None, upper: bytes | None):
        self._check(self.lib.edmd_cuda_halo_connect(self._h, lower, upper))

    def halo_exchange(self):
        self._check(self.lib.edmd_cuda_halo_exchange(self._h))

    def exchange_predict_device(self, mode=MODE_NORMAL):
        """Halo exchange + sweep as one stream-ordered sequence (transfer hidden behind the partition)."""
        self._check(self.lib.edmd_cuda_exchange_predict_device(self._h, mode))

    def counts(self):
        a, b = C.c_int(0), C.c_int(0)
        self._check(self.lib.edmd_cuda_get_counts(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def pcf_device(self, xy_dev_ptr: int, n_total: int, dr, max_r, part, nparts, counts_dev_ptr: int) -> int:
        nb = C.c_int(0)
        self._check(self.lib.edmd_cuda_pcf_device(self._h, C.c_void_p(xy_dev_ptr), n_total, dr, max_r,
                                                  part, nparts, C.c_void_p(counts_dev_ptr), C.byref(nb)))
        return nb.value

    def upload_aos(self, records: np.ndarray, t=0.0, with_cells=True):
        """records: structured array with fields x,y,vx,vy,rad[,cell (2 x i4)]."""
        rec = np.ascontiguousarray(records)
        off = {k: rec.dtype.fields[k][1] for k in ("x", "y", "vx", "vy", "rad")}
        off_cell = rec.dtype.fields["cell"][1] if with_cells else C.c_size_t(-1).value
        self._check(self.lib.edmd_cuda_upload_aos(
            self._h, _ptr(rec), rec.dtype.itemsize, off["x"], off["y"], off["vx"],
            off["vy"], off["rad"], off_cell, float(t)))

    def _out(self):
        n = self.n
        return (np.empty(n, np.float64), np.empty(n, np.uint8), np.empty(n, np.float64),
                np.empty(n, np.int32), np.empty(n, np.uint8), np.full(2, -7, np.int32))

    def predict_all(self, mode=MODE_NORMAL, vr=None, allow_overlap=False):
        tc, d, tl, p, ct, ov = self._out()
        vr_a = None if vr is None else _f64(vr, self.n)
        rc = self._check(self.lib.edmd_cuda_predict_all(
            self._h, mode, _ptr(vr_a), _ptr(tc), _ptr(d), _ptr(tl), _ptr(p), _ptr(ct), _ptr(ov)),
            allow=(EOVERLAP,) if allow_overlap else ())
        return dict(t_cross=tc, dir=d, t_coll=tl, partner=p, ctype=ct, overlap=ov, rc=rc)

    def set_growth(self, vr):
        vr_a = _f64(vr, self.n)
        self._check(self.lib.edmd_cuda_set_growth(self._h, _ptr(vr_a)))

    def predict_device(self, mode=MODE_NORMAL):
        self._check(self.lib.edmd_cuda_predict_device(self._h, mode))

    def fetch_predictions(self, allow_overlap=False):
        tc, d, tl, p, ct, ov = self._out()
        rc = self._check(self.lib.edmd_cuda_fetch_predictions(
            self._h, _ptr(tc), _ptr(d), _ptr(tl), _ptr(p), _ptr(ct), _ptr(ov)),
            allow=(EOVERLAP,) if allow_overlap else ())
        return dict(t_cross=tc, dir=d, t_coll=tl, partner=p, ctype=ct, overlap=ov, rc=rc)

    def free_fly(self, t_new, mode=MODE_NORMAL):
        self._check(self.lib.edmd_cuda_free_fly(self._h, mode, float(t_new)))

    def download_state(self):
        n = self.n
        out = [np.empty(n, np.float64) for _ in range(5)]
        self._check(self.lib.edmd_cuda_download_state(self._h, *[_ptr(a) for a in out]))
        return dict(zip(("x", "y", "vx", "vy", "rad"), out))

    def langevin_kick(self, T, gamma, dtnoise, seed, tick):
        """Langevin kick on the resident velocities (edmd_cuda_langevin_kick)."""
        self._check(self.lib.edmd_cuda_langevin_kick(self._h, float(T), float(gamma), float(dtnoise),
                                                     int(seed) & 0xFFFFFFFF, int(tick) & 0xFFFFFFFF))

    def normalize_velocities(self, e_init=1.0):
        """normalizePhysicalQ on the resident velocities (edmd_cuda_normalize_velocities)."""
        px, py, e, d = (C.c_double(0.0) for _ in range(4))
        self._check(self.lib.edmd_cuda_normalize_velocities(self._h, float(e_init), C.byref(px), C.byref(py),
                                                            C.byref(e), C.byref(d)))
        return dict(px_before=px.value, py_before=py.value, E_shifted=e.value, divisor=d.value)

    def shift_scale_velocities(self, dvx, dvy, divisor):
        """v <- (v - (dvx, dvy)) / divisor (edmd_cuda_shift_scale_velocities)."""
        self._check(self.lib.edmd_cuda_shift_scale_velocities(self._h, float(dvx), float(dvy), float(divisor)))

    def selftest_rsqrt(self) -> float:
        """Largest relative error of the hardware rsqrt over [2^-100, 2^64) (edmd_cuda_selftest_rsqrt)."""
        v = C.c_double(0.0)
        self._check(self.lib.edmd_cuda_selftest_rsqrt(self._h, C.byref(v)))
        return float(v.value)

    def pcf_num_bins(self, dr, max_r) -> int:
        nb = C.c_int(0)
        self._check(self.lib.edmd_cuda_pcf(self._h, dr, max_r, None, None, C.byref(nb)))
        return nb.value

    def pcf(self, dr, max_r):
        nb = self.pcf_num_bins(dr, max_r)
        counts = np.zeros(max(nb, 1), np.uint64)
        g = np.zeros(max(nb, 1), np.float64)
        nbc = C.c_int(0)
        self._check(self.lib.edmd_cuda_pcf(self._h, dr, max_r, _ptr(counts), _ptr(g), C.byref(nbc)))
        return dict(num_bins=nb, counts=counts[:nb], g_r=g[:nb],
                    r=(np.arange(nb) + 0.5) * dr)

    def boop_cutoff(self, r_c=2.5):
        n = self.n
        q5, q6, q7, arg = (np.empty(n, np.float64) for _ in range(4))
        nb = np.empty(n, np.int32)
        mean = C.c_double(0.0)
        self._check(self.lib.edmd_cuda_boop_cutoff(
            self._h, r_c, _ptr(q5), _ptr(q6), _ptr(q7), _ptr(arg), _ptr(nb), C.addressof(mean)))
        return dict(q5=q5, q6=q6, q7=q7, q6_arg=arg, neighbors=nb, mean_q6=mean.value)

    def kinetic(self):
        """physicalQ's sums (src/EDMD.c:5968-5997) over the resident velocities."""
        e, px, py = C.c_double(0), C.c_double(0), C.c_double(0)
        self._check(self.lib.edmd_cuda_kinetic(self._h, C.byref(e), C.byref(px), C.byref(py)))
        return dict(E=e.value, px=px.value, py=py.value)

    def rescale_velocities(self, T):
        """addNoise's velocity rescale (src/EDMD.c:4899-4902) on the resident velocities."""
        e, s = C.c_double(0), C.c_double(0)
        self._check(self.lib.edmd_cuda_rescale_velocities(self._h, float(T), C.byref(e), C.byref(s)))
        return dict(E_before=e.value, divisor=s.value)

    def boop_voronoi(self):
        """computeBOOPVoronoi (src/boop.c:15-59) on the resident positions."""
        n = self.n
        q5, q6, q7, arg = (np.empty(n, np.float64) for _ in range(4))
        nb = np.empty(n, np.int32)
        mean = C.c_double(0.0)
        self._check(self.lib.edmd_cuda_boop_voronoi(
            self._h, _ptr(q5), _ptr(q6), _ptr(q7), _ptr(arg), _ptr(nb), C.addressof(mean)))
        return dict(q5=q5, q6=q6, q7=q7, q6_arg=arg, neighbors=nb, mean_q6=mean.value)

    def voronoi_cells(self):
        """get_particle_voronoi_area / _perimeter (src/voronoi_edmd.c:123-149)."""
        n = self.n
        area, per = np.empty(n, np.float64), np.empty(n, np.float64)
        nb = np.empty(n, np.int32)
        self._check(self.lib.edmd_cuda_voronoi_cells(self._h, _ptr(area), _ptr(per), _ptr(nb)))
        return dict(area=area, perimeter=per, neighbors=nb)

    def g6_correlation(self, dr, max_r, psi_re=None, psi_im=None):
        """compute_g6_correlation (src/pcf.c:169-230); psi None = Voronoi psi6 from the device."""
        nb = C.c_int(0)
        self._check(self.lib.edmd_cuda_g6_correlation(self._h, dr, max_r, None, None, None, None, C.byref(nb)))
        counts = np.zeros(nb.value, np.uint64)
        g6 = np.zeros(nb.value, np.float64)
        pr = None if psi_re is None else _f64(psi_re, self.n)
        pi = None if psi_im is None else _f64(psi_im, self.n)
        self._check(self.lib.edmd_cuda_g6_correlation(self._h, dr, max_r, _ptr(pr), _ptr(pi), _ptr(counts),
                                                      _ptr(g6), C.byref(nb)))
        return dict(num_bins=nb.value, counts=counts, g6_corr=g6)

    def structure_factor(self, q_max, velocity=False):
        """initStructureFactor's grid + computeStructureFactor / computeVelocityStructureFactor
        (src/struc.c:328-345, 364-408); s[i, j] for (qx[i], qy[j])."""
        nqx, nqy = C.c_int(0), C.c_int(0)
        self._check(self.lib.edmd_cuda_structure_factor(self._h, q_max, int(velocity), C.byref(nqx), C.byref(nqy),
                                                        None, None, None, None, None))
        qx, qy = np.zeros(nqx.value), np.zeros(nqy.value)
        nk = nqx.value * nqy.value
        s, re, im = np.zeros(nk), np.zeros(nk), np.zeros(nk)
        self._check(self.lib.edmd_cuda_structure_factor(self._h, q_max, int(velocity), C.byref(nqx), C.byref(nqy),
                                                        _ptr(qx), _ptr(qy), _ptr(s), _ptr(re), _ptr(im)))
        shape = (nqx.value, nqy.value)
        return dict(qx=qx, qy=qy, s=s.reshape(shape), re=re.reshape(shape), im=im.reshape(shape))

    def pcf_bond_order(self, dr, max_r, k_vector):
        """calculate_bond_order_pcf (src/pcf.c:77-167) on the resident positions."""
        k = np.ascontiguousarray(k_vector, dtype=np.float64)
        nb = C.c_int(0)
        self._check(self.lib.edmd_cuda_pcf_bond_order(self._h, dr, max_r, _ptr(k), None, None, None,
                                                      C.byref(nb)))
        counts = np.zeros(nb.value, np.uint64)
        g = np.zeros(nb.value, np.float64)
        g6 = np.zeros(nb.value, np.float64)
        self._check(self.lib.edmd_cuda_pcf_bond_order(self._h, dr, max_r, _ptr(k), _ptr(counts), _ptr(g),
                                                      _ptr(g6), C.byref(nb)))
        return dict(num_bins=nb.value, counts=counts, g_r=g, g6_r=g6)

    def bragg_peak(self, expected_bragg):
        """find_max_structure_factor_bragg (src/pcf.c:405-467) on the resident positions."""
        k = np.zeros(2, np.float64)
        s = C.c_double(0.0)
        self._check(self.lib.edmd_cuda_bragg_peak(self._h, float(expected_bragg), _ptr(k), C.byref(s)))
        return dict(k=k, s_max=s.value)

    def calendar_plan(self, paul_time, dt_paul, paul_n, actual_paul, allow_declined=False):
        """Ingest plan of the last sweep's 2N events for an empty calendar
        (addEventToQueue, src/EDMD.c:2144-2170, applied in the batch loops' order)."""
        e2 = 2 * self.n
        bucket = np.empty(e2, np.int32)
        nxt = np.empty(e2, np.int32)
        prv = np.empty(e2, np.int32)
        head = np.empty(paul_n + 1, np.int32)
        n_tree = C.c_int32(0)
        rc = self.lib.edmd_cuda_calendar_plan(self._h, float(paul_time), float(dt_paul), int(paul_n),
                                              int(actual_paul), _ptr(bucket), _ptr(nxt), _ptr(prv),
                                              _ptr(head), C.byref(n_tree))
        self._check(rc, allow=(EPLAN,) if allow_declined else ())
        return dict(rc=rc, bucket=bucket, next=nxt, prev=prv, head=head, n_tree=n_tree.value)

    def bench(self, what, mode=MODE_NORMAL, dr=0.0, max_r=0.0, warmup=3, iters=10,
              flush_bytes=0, split=True):
        """split=False (sweep only): no event between K0 and K1 -- the chain as the product launches it;
        returns (ms_total, None)."""
        tot = np.zeros(iters, np.float32)
        main = np.zeros(iters, np.float32) if split else None
        self._check(self.lib.edmd_cuda_bench(self._h, what, mode, dr, max_r, warmup, iters,
                                             flush_bytes, _ptr(tot), _ptr(main) if split else None))
        return tot, main


class EdmdMg:
    """Several GPUs in one process behind the single-GPU interface (edmd_cuda_create_mg): whole-system
    host arrays in and out; `devices` may list one device several times (several slabs on one GPU)."""

    def __init__(self, n: int, lx: float, ly: float, devices=(0,)):
        self.lib = load_library()
        self.n = int(n)
        devs = (C.c_int * len(devices))(*[int(d) for d in devices])
        h = C.c_void_p()
        rc = self.lib.edmd_cuda_create_mg(len(devices), devs, self.n, lx, ly, C.byref(h))
        self._h = h if rc == 0 else None
        self.ndev = len(devices)
        if rc != 0:
            raise EdmdError(rc, "edmd_cuda_create_mg failed")

    def close(self):
        if getattr(self, "_h", None):
            self.lib.edmd_cuda_destroy_mg(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, allow=()):
        if rc != 0 and rc not in allow:
            raise EdmdError(rc, self.lib.edmd_cuda_mg_last_error(self._h).decode())
        return rc

    @property
    def slab_sizes(self):
        out = (C.c_int * self.ndev)()
        self._check(self.lib.edmd_cuda_mg_slab_sizes(self._h, out))
        return list(out)

    def upload(self, x, y, vx, vy, rad, cell_xy=None, t=0.0):
        n = self.n
        a = [_f64(v, n) for v in (x, y, vx, vy, rad)]
        cells = None if cell_xy is None else np.ascontiguousarray(cell_xy, dtype=np.int32)
        self._check(self.lib.edmd_cuda_mg_upload(self._h, *[_ptr(v) for v in a], _ptr(cells), float(t)))

    def predict_all(self, mode=MODE_NORMAL, allow_overlap=False):
        n = self.n
        tc, tl = np.empty(n, np.float64), np.empty(n, np.float64)
        d, ct = np.empty(n, np.uint8), np.empty(n, np.uint8)
        p, ov = np.empty(n, np.int32), np.full(2, -7, np.int32)
        rc = self._check(self.lib.edmd_cuda_mg_predict_all(self._h, mode, _ptr(tc), _ptr(d), _ptr(tl), _ptr(p),
                                                           _ptr(ct), _ptr(ov)),
                         allow=(EOVERLAP,) if allow_overlap else ())
        return dict(t_cross=tc, dir=d, t_coll=tl, partner=p, ctype=ct, overlap=ov, rc=rc)

    def boop_cutoff(self, r_c=2.5):
        n = self.n
        q5, q6, q7, arg = (np.empty(n, np.float64) for _ in range(4))
        nb = np.empty(n, np.int32)
        mean = C.c_double(0.0)
        self._check(self.lib.edmd_cuda_mg_boop_cutoff(self._h, r_c, _ptr(q5), _ptr(q6), _ptr(q7), _ptr(arg), _ptr(nb),
                                                      C.byref(mean)))
        return dict(q5=q5, q6=q6, q7=q7, q6_arg=arg, neighbors=nb, mean_q6=mean.value)

    def pcf(self, x, y, dr, max_r):
        nb = C.c_int(0)
        self._check(self.lib.edmd_cuda_mg_pcf(self._h, None, None, dr, max_r, None, None, C.byref(nb)))
        counts = np.zeros(max(nb.value, 1), np.uint64)
        g = np.zeros(max(nb.value, 1), np.float64)
        xa, ya = _f64(x, self.n), _f64(y, self.n)
        self._check(self.lib.edmd_cuda_mg_pcf(self._h, _ptr(xa), _ptr(ya), dr, max_r, _ptr(counts), _ptr(g),
                                              C.byref(nb)))
        return dict(counts=counts[:nb.value], g_r=g[:nb.value], num_bins=nb.value)
