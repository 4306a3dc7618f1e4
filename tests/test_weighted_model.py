"""CPU models of the arithmetic of graphical-edmd_b200/csrc/analysis_weighted.cu (numpy, float64):

  * k_pcf_bond_order: cos(k.d) of a pair from per-particle phases psi = e^{i k.r} and the phase factor of the
    periodic image shift PBC() applied, cos(k.d) = Re(conj(psi_i) psi_j e^{i k.S}), d = r_j - r_i + S
    (the reference takes `cos(k_vector[0]*dx + k_vector[1]*dy)` per pair, src/pcf.c:123);
  * its bin certificate (k dr)^2 (1 + 2^-48) <= s < ((k+1) dr)^2 (1 - 2^-48)  =>  (int)(sqrt(s) / dr) == k,
    tried on values of s within a few ulps of the bin edges;
  * the 64-bit fixed-point sum kept as two 32-bit words with the carry taken from the returned low word;
  * k_grid_sums: e^{i k.r} = e^{i kx x} e^{i ky y} on a product grid of wave vectors, the velocity weight
    (vx - i vy) folded into the y table (src/pcf.c:442-446, src/struc.c:364-408).

The gate of the north star is 1e-10; these pin the error budget on the CPU.  The CUDA code itself is
compared with the reference's fixtures and the oracle by the -m gpu tests."""
import numpy as np


def min_image_shift(d, half, length):
    """(wrapped d, shift S applied) of PBC(), src/EDMD.c:5896-5913."""
    s = np.where(d >= half, -length, np.where(d < -half, length, 0.0))
    return d + s, s


def test_pair_cosine_from_particle_phases(pkg):
    c = pkg.synth.lattice_config(1500, 0.70, seed=3)
    x, y, lx, ly = c["x"], c["y"], c["lx"], c["ly"]
    k = np.array([3.1870482, 0.0157])                    # not a reciprocal-lattice vector of the box
    i, j = np.triu_indices(c["n"], 1)
    dx, sx = min_image_shift(x[j] - x[i], lx / 2, lx)
    dy, sy = min_image_shift(y[j] - y[i], ly / 2, ly)
    want = np.cos(k[0] * dx + k[1] * dy)
    psi = np.exp(1j * (k[0] * x + k[1] * y))
    got = (np.conj(psi[i]) * psi[j] * np.exp(1j * (k[0] * sx + k[1] * sy))).real
    assert (sx != 0).any() and (sy != 0).any()
    assert np.abs(got - want).max() < 1e-12
    # a box a thousand disks wide: phases ~ 10^4
    big = 1000.0
    got = (np.conj(np.exp(1j * k[0] * (x + big))) * np.exp(1j * k[0] * (x[::-1] + big))).real
    assert np.abs(got - np.cos(k[0] * (x[::-1] - x))).max() < 1e-11


def test_bin_certificate_implies_the_reference_bin():
    rng = np.random.default_rng(11)
    for dr in (0.1, 0.25, 2.0, 0.037):
        dr_lo, dr_hi = dr * (1.0 + 2.0 ** -49), dr * (1.0 - 2.0 ** -49)
        kk = rng.integers(0, 20000, 200000).astype(np.float64)
        # s within a few ulps of either edge of bin k, and in its middle
        edge = np.where(rng.random(kk.size) < 0.5, kk, kk + 1.0) * dr
        s = edge * edge
        for _ in range(3):
            s = np.where(rng.random(kk.size) < 0.5, np.nextafter(s, np.inf), np.nextafter(s, 0.0))
        s = np.r_[s, ((kk + rng.random(kk.size)) * dr) ** 2]
        for k in (np.floor(np.sqrt(s) / dr), np.floor(np.sqrt(s) / dr) - 1.0, np.floor(np.sqrt(s) / dr) + 1.0):
            e0, e1 = k * dr_lo, (k + 1.0) * dr_hi
            ok = (s >= e0 * e0) & (s < e1 * e1) & (k >= 0)
            ref = np.sqrt(s) / dr                         # the reference: sqrt, then the division, then (int)
            assert np.array_equal(ref[ok].astype(np.int64), k[ok].astype(np.int64))
        # the certificate decides nearly everything away from the edges
        mid = ((kk + 0.5) * dr) ** 2
        k = np.floor(np.sqrt(mid) / dr)
        assert (((mid >= (k * dr_lo) ** 2) & (mid < ((k + 1) * dr_hi) ** 2)).mean()) > 0.999


def test_split_fixed_point_sum_equals_the_64_bit_sum():
    rng = np.random.default_rng(5)
    w = rng.uniform(-1, 1, 100000)
    q = np.rint(w * 2.0 ** 32).astype(np.int64).view(np.uint64)
    lo_sum, hi_sum = np.uint32(0), np.uint32(0)
    with np.errstate(over="ignore"):
        for v in q[:5000]:
            lo = np.uint32(v & np.uint64(0xffffffff))
            old = lo_sum
            lo_sum = np.uint32(lo_sum + lo)
            carry = np.uint32(1) if np.uint32(old + lo) < old else np.uint32(0)
            hi_sum = np.uint32(hi_sum + np.uint32(v >> np.uint64(32)) + carry)
        total = np.uint64(0)
        for v in q[:5000]:
            total = np.uint64(total + v)
    assert (np.uint64(hi_sum) << np.uint64(32)) + np.uint64(lo_sum) == total
    assert abs(float(total.view(np.int64)) / 2.0 ** 32 - w[:5000].sum()) < 5000 * 2.0 ** -33


def test_factorised_grid_sums(pkg):
    c = pkg.synth.lattice_config(4000, 0.72, seed=8)
    x, y, vx, vy = c["x"], c["y"], c["vx"], c["vy"]
    qx = 2 * np.pi / c["lx"] * np.arange(-20, 21)
    qy = 2 * np.pi / c["ly"] * np.arange(-15, 16) * 7.0
    ph = qx[:, None, None] * x[None, None, :] + qy[None, :, None] * y[None, None, :]
    A = np.exp(1j * qx[:, None] * x[None, :])
    B = np.exp(1j * qy[:, None] * y[None, :])
    # positions: re + i im = sum_p e^{i q.r}
    want = np.exp(1j * ph).sum(axis=2)
    got = A @ B.T
    assert np.abs(got - want).max() < 1e-10 * max(1.0, np.abs(want).max())
    s_want, s_got = np.abs(want) ** 2 / c["n"], np.abs(got) ** 2 / c["n"]
    assert np.abs(s_got - s_want).max() <= 1e-10 * s_want.max()
    # velocities: `re += vx cos + vy sin; im += vx sin - vy cos`  =  sum_p (vx - i vy) e^{i q.r}
    want_v = (vx * np.cos(ph) + vy * np.sin(ph)).sum(axis=2) + 1j * (vx * np.sin(ph) - vy * np.cos(ph)).sum(axis=2)
    got_v = A @ ((vx - 1j * vy)[None, :] * B).T
    assert np.abs(got_v - want_v).max() < 1e-10 * max(1.0, np.abs(want_v).max())
