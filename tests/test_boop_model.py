"""CPU model of the arithmetic of the psi6 kernel on the cell slots
(graphical-edmd_b200/csrc/cell_sweep.cu, boop_tile / atan2_gate / rsqrt_gate) in numpy:

  * every candidate runs the whole chain; one that is not a neighbour at a distance gets 1/r = 0,
    its z = 0 then gives z^6 = -1 through the unit-modulus squarings (z^2 = (2 zr^2 - 1, 2 zr zi)),
    and the count of such visits is put back into sum6 at the end;
  * sum5 and sum7 from A = sum Re(z) z^6 and B = sum Im(z) z^6;
  * 1/r and the three moduli from a reciprocal-square-root SEED of ~2^-19 (the MUFU seed is better)
    plus one third-order step;
  * atan2 by octant folding, one rotation by pi/4 and the degree-8 polynomial of the kernel
    (the same hexadecimal coefficients), the quotient from a 2^-19 reciprocal seed + two Newton steps;
  * coincident disks count as z = 1 (atan2(0, 0) = 0 in the reference, src/boop.c:86-89).

Checked against the oracle's computeBOOPCutoff (src/boop.c:61-107) at the 1e-10 gate of the
north star.  This pins the formulas and their error budget, not the CUDA code itself -- that is
what the -m gpu parity tests do."""
import math

import numpy as np
import pytest
from scipy.spatial import cKDTree

from helpers import assert_boop_close

ATAN_C = [float.fromhex(h) for h in (
    "0x1.fffffffffff0fp-1", "-0x1.55555554e5fc5p-2", "0x1.99999911c1573p-3", "-0x1.2492291945813p-3",
    "0x1.c714d3e720df5p-4", "-0x1.73d9cba10d56bp-4", "0x1.35ced8de982a2p-4", "-0x1.e13ac280a9accp-5",
    "0x1.f657ae7e09908p-6")]
SEED_ERR = 2.0 ** -19


def noisy(v, rng):
    return v * (1.0 + rng.uniform(-1, 1, size=np.shape(v)) * SEED_ERR)


def rsqrt_gate(x, rng):
    with np.errstate(all="ignore"):
        y = noisy(1.0 / np.sqrt(x), rng)
        e = 1.0 - (x * y) * y
        return y * ((0.375 * e + 0.5) * e) + y


def atan2_gate(y, x, rng):
    ax, ay = np.abs(x), np.abs(y)
    hi, lo = np.maximum(ax, ay), np.minimum(ax, ay)
    rot = lo > 0.41421356237309503 * hi
    h2 = np.where(rot, hi + lo, hi)
    l2 = np.where(rot, lo - hi, lo)
    with np.errstate(all="ignore"):
        r = noisy(1.0 / h2, rng)
        r = r * (1.0 - h2 * r) + r
        r = r * (1.0 - h2 * r) + r
        t = l2 * r
        u = t * t
        p = np.full_like(u, ATAN_C[8])
        for k in range(7, -1, -1):
            p = p * u + ATAN_C[k]
        a = t * p
    a = np.where(rot, a + 0.78539816339744831, a)
    a = np.where(ay > ax, 1.5707963267948966 - a, a)
    a = np.where(x < 0.0, 3.1415926535897931 - a, a)
    a = np.where(h2 > 0.0, a, 0.0)
    return np.copysign(a, y)


def model(n, lx, ly, x, y, nxc, nyc, rc, rng, empty_visits=2):
    """All candidates of the reference's 3 x 3 scan (+ `empty_visits` visits of an empty cell per
    particle: a disk 1e300 away), every one through the branch-free chain."""
    cx = np.floor(x / (lx / nxc)).astype(int)
    cy = np.floor(y / (ly / nyc)).astype(int)
    tree = cKDTree(np.c_[x, y], boxsize=[lx, ly])
    pairs = tree.query_pairs(2.0 * rc, output_type="ndarray")   # candidates in and out of range
    i = np.r_[pairs[:, 0], pairs[:, 1], np.repeat(np.arange(n), empty_visits)]
    j = np.r_[pairs[:, 1], pairs[:, 0], np.repeat(np.arange(n), empty_visits)]
    far = np.r_[np.zeros(2 * len(pairs), bool), np.ones(n * empty_visits, bool)]
    with np.errstate(all="ignore"):
        dx = np.where(far, 1e300, x[j]) - x[i]
        dy = np.where(far, 1e300, y[j]) - y[i]
        dx = np.where(dx >= lx / 2, dx - lx, np.where(dx < -lx / 2, dx + lx, dx))
        dy = np.where(dy >= ly / 2, dy - ly, np.where(dy < -ly / 2, dy + ly, dy))
        dcx = (cx[j] - cx[i] + nxc // 2) % nxc - nxc // 2
        dcy = (cy[j] - cy[i] + nyc // 2) % nyc - nyc // 2
        scanned = far | ((np.abs(dcx) <= 1) & (np.abs(dcy) <= 1))
        i, dx, dy = i[scanned], dx[scanned], dy[scanned]
        r2 = dx * dx + dy * dy
        inr = r2 < rc * rc
        pos = inr & (r2 > 0.0)
        yv = np.where(pos, rsqrt_gate(r2, rng), 0.0)     # the bit mask
        zr, zi = dx * yv, dy * yv
        tr = zr + zr
        z2r, z2i = tr * zr - 1.0, tr * zi
        t2 = z2r + z2r
        z4r, z4i = t2 * z2r - 1.0, t2 * z2i
        z6r = z4r * z2r - z4i * z2i
        z6i = z4r * z2i + z4i * z2r
    assert np.all(z6r[~pos] == -1.0) and np.all(z6i[~pos] == 0.0)
    s6r, s6i = np.bincount(i, z6r, n), np.bincount(i, z6i, n)
    Ar, Ai = np.bincount(i, zr * z6r, n), np.bincount(i, zr * z6i, n)
    Br, Bi = np.bincount(i, zi * z6r, n), np.bincount(i, zi * z6i, n)
    nvis = np.bincount(i, minlength=n)
    nb = np.bincount(i, inr.astype(float), n).astype(np.int64)
    npos = np.bincount(i, pos.astype(float), n).astype(np.int64)
    nzero = nb - npos
    s6r = s6r + (nvis - npos + nzero)
    s5r, s5i = Ar + Bi + nzero, Ai - Br
    s7r, s7i = Ar - Bi + nzero, Ai + Br
    with np.errstate(all="ignore"):
        nd = np.maximum(nb, 1).astype(float)
        inv = noisy(1.0 / nd, rng)
        inv = inv * (1.0 - nd * inv) + inv
        inv = inv * (1.0 - nd * inv) + inv

        def modulus(re, im):
            m2 = re * re + im * im
            return np.where(m2 > 0.0, m2 * rsqrt_gate(m2, rng) * inv, 0.0)

        q5, q6, q7 = modulus(s5r, s5i), modulus(s6r, s6i), modulus(s7r, s7i)
        arg = atan2_gate(s6i, s6r, rng)
    # the record: the neighbour count travels in the eight low mantissa bits of the argument
    bits = arg.view(np.int64).copy()
    bits = (bits & ~np.int64(0xff)) | nb
    assert np.array_equal(bits & 0xff, nb)
    arg = (bits & ~np.int64(0xff)).view(np.float64)
    z = nb == 0
    return dict(q5=np.where(z, 0.0, q5), q6=np.where(z, 0.0, q6), q7=np.where(z, 0.0, q7),
                q6_arg=np.where(z, 0.0, arg), neighbors=nb.astype(np.int32))


@pytest.mark.parametrize("n,phi,seed", [(12000, 0.70, 3), (12000, 0.85, 4), (6000, 0.72, 5)])
def test_branch_free_psi6_arithmetic_meets_the_gate(pkg, oracle, n, phi, seed):
    c = pkg.synth.lattice_config(n, phi, seed)
    b = oracle.box(c["n"], c["lx"], c["ly"])
    want = oracle.boop_cutoff(c["n"], c["lx"], c["ly"], c["x"], c["y"], 2.5)
    got = model(c["n"], c["lx"], c["ly"], c["x"], c["y"], b.nx, b.ny, 2.5, np.random.default_rng(seed))
    assert_boop_close(got, want)
    # far inside the gate: the budget is rounding, not the seeds
    for k in ("q5", "q6", "q7"):
        assert np.abs(got[k] - want[k]).max() < 1e-13


def test_coincident_disks_count_as_unit_vectors(pkg, oracle):
    c = pkg.synth.lattice_config(3000, 0.70, 9)
    x, y = c["x"].copy(), c["y"].copy()
    x[1::50], y[1::50] = x[0::50][: len(x[1::50])], y[0::50][: len(y[1::50])]   # exact duplicates
    b = oracle.box(c["n"], c["lx"], c["ly"])
    want = oracle.boop_cutoff(c["n"], c["lx"], c["ly"], x, y, 2.5)
    got = model(c["n"], c["lx"], c["ly"], x, y, b.nx, b.ny, 2.5, np.random.default_rng(1))
    assert_boop_close(got, want)


def test_atan2_polynomial_is_far_inside_the_gate():
    rng = np.random.default_rng(7)
    th = np.r_[rng.uniform(-math.pi, math.pi, 200000), np.linspace(-math.pi, math.pi, 4001),
               np.arange(-8, 9) * math.pi / 8]
    r = 10.0 ** rng.uniform(-12, 1, size=th.shape)
    y, x = r * np.sin(th), r * np.cos(th)
    d = np.abs(np.angle(np.exp(1j * (atan2_gate(y, x, rng) - np.arctan2(y, x)))))
    assert d.max() < 1e-13
    assert atan2_gate(np.array([0.0]), np.array([0.0]), rng)[0] == 0.0
