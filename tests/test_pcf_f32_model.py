"""CPU model of the FP32 bin decision of the default g(r) kernel
(graphical-edmd_b200/csrc/analysis_pcf_sorted.cu, k_pcf_f32): the same plan per
tile pair (image shift or per-pair wrap, error bound eps, range limits) and the
same FP32 operations per pair in numpy float32 (q_lo = fma(s, rsqrt(s), -eps), certain
when frac(q_lo) < 1 - 2 eps), with the hardware's approximate
rsqrt replaced by an ADVERSARIAL one (correct value pushed to either end of the
1.28e-7 relative budget).  Every pair the model calls certain must land in the
bin / range decision of the reference's FP64 arithmetic (src/pcf.c:34-47 with
PBC, src/EDMD.c:5896-5913); the share of undecided pairs must stay small.

This checks the error analysis written at the top of k_pcf_f32, not the CUDA
code itself -- that is what the -m gpu parity tests do."""
import numpy as np
import pytest

U = 2.0 ** -24 * 1.001
MAGIC = np.float32(12582912.0)
f32 = np.float32


def min_image(d, half, length):
    return np.where(d >= half, d - length, np.where(d < -half, d + length, d))


def gap(lo, hi):
    return max(0.0, max(lo, -hi))


def plan_axis(d0, d1, half, length):
    shift, wrap = 0.0, False
    if d1 < half and d0 >= -half:
        g = gap(d0, d1)
    elif d0 >= half:
        shift, g = -length, gap(d0 - length, d1 - length)
    elif d1 < -half:
        shift, g = length, gap(d0 + length, d1 + length)
    else:
        wrap = True
        g = min(gap(d0, d1), gap(d0 - length, d1 - length), gap(d0 + length, d1 + length))
    m_un = max(abs(d0 + shift), abs(d1 + shift)) * (1.0 + 1e-12)
    m_w = max(half, m_un - length) if wrap else m_un
    return shift, wrap, m_un, m_w, g


def down32(x):
    y = f32(x)
    return y if float(y) <= x else np.nextafter(y, f32(-np.inf))


def up32(x):
    y = f32(x)
    return y if float(y) >= x else np.nextafter(y, f32(np.inf))


def model_tile_pair(xa, ya, xb, yb, lx, ly, dr, max_r, num_bins, rsqrt_bias):
    """Returns (take, drop, bin32, eps) per pair (i in a, j in b), or None when the tile pair is skipped."""
    inv_dr = 1.0 / dr
    lmax = max(lx, ly)
    slack = 2.0 ** -44 * (lmax * inv_dr + 1.0)
    A = (xa.min(), xa.max(), ya.min(), ya.max())
    B = (xb.min(), xb.max(), yb.min(), yb.max())
    sx, wx, mxu, mxw, gx = plan_axis(B[0] - A[1], B[1] - A[0], lx / 2, lx)
    sy, wy, myu, myw, gy = plan_axis(B[2] - A[3], B[3] - A[2], ly / 2, ly)
    rcut = max_r * (1.0 + 1e-12)
    if gx * gx + gy * gy >= rcut * rcut:
        return None
    hbx, hby = 0.5 * (B[1] - B[0]) * inv_dr, 0.5 * (B[3] - B[2]) * inv_dr
    ex = U * (2 * hbx + 2 * mxu * inv_dr) + (U * lx * inv_dr if wx else 0.0) + slack
    ey = U * (2 * hby + 2 * myu * inv_dr) + (U * ly * inv_dr if wy else 0.0) + slack
    e = np.hypot(ex, ey) * 1.001
    rmax = np.hypot(mxw, myw) * inv_dr * (1 + 1e-9)
    rho = 1.28e-7 + 2.1 * U
    eps = ((rmax + e) * rho + e + 1e-12 + slack) * 1.001
    eps32 = up32(eps)
    assert float(eps32) < 0.45           # else the kernel sends the whole tile pair to FP64
    cth = down32(1.0 - 2.0 * float(eps32) - 2.0 ** -22)
    inr = (rmax + e) * (1 + 1e-6) < float(num_bins) and not wx and not wy
    cb = (0.5 * (B[0] + B[1]), 0.5 * (B[2] + B[3]))
    bx = ((xb - cb[0]) * inv_dr).astype(f32)
    by = ((yb - cb[1]) * inv_dr).astype(f32)
    px = ((xa - (cb[0] + sx)) * inv_dr).astype(f32)
    py = ((ya - (cb[1] + sy)) * inv_dr).astype(f32)
    dx = bx[None, :] - px[:, None]
    dy = by[None, :] - py[:, None]
    lx32, ly32 = f32(lx * inv_dr), f32(ly * inv_dr)
    if wx:
        m = np.abs(dx)
        dx = np.where(m >= f32(0.5) * lx32, m - lx32, m)
    if wy:
        m = np.abs(dy)
        dy = np.where(m >= f32(0.5) * ly32, m - ly32, m)
    # fma in float64 then one rounding = the float32 fma (products of floats are exact in double)
    s1 = (dx.astype(np.float64) * dx + 1e-30).astype(f32)
    s = (dy.astype(np.float64) * dy + s1).astype(f32)
    y = ((1.0 / np.sqrt(s.astype(np.float64))) * (1.0 + rsqrt_bias)).astype(f32)
    # q_lo = fma(s, y, -eps): exact product in double, one rounding
    sy = s.astype(np.float64) * y.astype(np.float64)
    qlo = (sy - float(eps32)).astype(f32)
    assert float(qlo.max()) < 2 ** 22
    nf = np.floor(qlo).astype(f32)       # t = RD(q_lo + 1.5 * 2^23) holds floor(q_lo) exactly
    frac = qlo - nf
    cert = frac < cth
    # every pair is counted in floor(q_lo) clamped to num_bins (the dummy word); `bin < num_bins`
    # is the whole range test
    if inr:
        assert float(nf.max()) < num_bins
    take = cert & (nf < num_bins)
    drop = cert & (nf >= num_bins)
    return take, drop, nf.astype(np.int64), float(eps32)


def reference_bins(xa, ya, xb, yb, lx, ly, dr, max_r, num_bins):
    dx = min_image(xb[None, :] - xa[:, None], lx / 2, lx)
    dy = min_image(yb[None, :] - ya[:, None], ly / 2, ly)
    r = np.sqrt(dx * dx + dy * dy)
    b = (r / dr).astype(np.int64)
    ok = (r < max_r) & (b < num_bins)
    return np.where(ok, b, -1)


@pytest.mark.parametrize("lx,ly,dr,frac,tile", [(2276.46, 1971.47, 0.1, 0.5, 40.0), (730.0, 700.0, 0.1, 0.5, 35.0),
                                                (471.0, 471.0, 0.013, 0.3, 30.0), (300.0, 420.0, 2.0, 0.75, 60.0),
                                                (64.0, 64.0, 0.05, 0.5, 55.0)])
def test_certain_pairs_match_the_reference_bins(lx, ly, dr, frac, tile):
    rng = np.random.default_rng(int(lx * 7 + dr * 1000))
    max_r = min(lx, ly) * frac if frac <= 0.5 else max(lx, ly) * frac
    num_bins = int(max_r / dr)
    checked = undecided = pairs = 0
    for trial in range(120):
        # two tiles anywhere in the box (tile b sometimes across the periodic seam from a,
        # sometimes the same region); 48 particles each
        wa, wb = rng.uniform(0.2, 1.0, 2) * tile
        ax, ay = rng.uniform(0, lx - wa), rng.uniform(0, ly - wa)
        if trial % 6 == 0:
            bx_, by_ = ax, ay
            wb = wa
        else:
            bx_, by_ = rng.uniform(0, lx - wb), rng.uniform(0, ly - wb)
        xa, ya = ax + rng.uniform(0, wa, 48), ay + rng.uniform(0, wa, 48)
        xb, yb = bx_ + rng.uniform(0, wb, 48), by_ + rng.uniform(0, wb, 48)
        if trial % 5 == 0:      # distances sitting on bin edges: lattice offsets in multiples of dr
            xb, yb = xa[0] + dr * rng.integers(-200, 200, 48), ya[0] + dr * rng.integers(-200, 200, 48)
            xb, yb = np.mod(xb, lx), np.mod(yb, ly)
        want = reference_bins(xa, ya, xb, yb, lx, ly, dr, max_r, num_bins)
        for bias in (-1.28e-7, 0.0, 1.28e-7):
            got = model_tile_pair(xa, ya, xb, yb, lx, ly, dr, max_r, num_bins, bias)
            if got is None:
                assert (want < 0).all()      # a skipped tile pair holds no pair in range
                continue
            take, drop, b32, eps = got
            assert np.array_equal(b32[take], want[take]), (trial, bias, eps)
            assert (want[drop] < 0).all(), (trial, bias)
            checked += int(take.sum() + drop.sum())
            undecided += int((~take & ~drop).sum())
            pairs += take.size
    assert checked > 0.5 * pairs
    # the adversarial lattice tiles sit on bin edges on purpose; the rest must be decided in FP32
    assert undecided < 0.25 * pairs
