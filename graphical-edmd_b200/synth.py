"""Synthetic non-overlapping hard-disk configurations.

Jittered triangular lattice in a near-square periodic box built with the
reference's ``Hex`` box formulas (src/EDMD.c:1026-1033), Maxwell velocities at
unit temperature with the centre-of-mass momentum removed.  All randomness is a
counter-based hash of (seed, particle id, stream) so any generator (numpy here,
C in the host) reproduces the same numbers.
"""
from __future__ import annotations

import math

import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _mix(z: np.ndarray) -> np.ndarray:
    """splitmix64 finaliser on uint64 arrays (wrap-around arithmetic)."""
    with np.errstate(over="ignore"):
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def uniform01(seed: int, ids: np.ndarray, stream: int) -> np.ndarray:
    """Uniform doubles in [0,1) keyed on (seed, id, stream)."""
    with np.errstate(over="ignore"):
        z = (ids.astype(np.uint64) + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
        z = z + np.uint64(seed & 0xFFFFFFFF) * np.uint64(0xD1B54A32D192ED03)
        z = z + np.uint64(stream + 1) * np.uint64(0x8CB92BA72F3D8DD7)
        z = _mix(_mix(z))
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def lattice_dims(n_target: int) -> tuple[int, int]:
    ny = int(math.floor(math.sqrt(n_target)))
    ny -= ny % 2
    ny = max(ny, 2)
    nx = max(n_target // ny, 1)
    return nx, ny


def lattice_config(n_target: int, phi: float, seed: int, *, shuffle: bool = True,
                   small_fraction: float = 0.0, size_ratio: float = 0.4,
                   jitter: float = 0.3) -> dict:
    """Jittered triangular lattice of (mostly) unit disks at packing fraction phi.

    shuffle=True permutes particle ids so memory order is uncorrelated with
    position, like a configuration produced by the reference itself (random
    insertion order, src/EDMD.c:1690-1829).
    """
    nx, ny = lattice_dims(n_target)
    n = nx * ny
    ly = math.sqrt(math.sqrt(3.0) / 2.0 * ny / nx * math.pi * n / phi)
    lx = 2.0 / math.sqrt(3.0) * nx / ny * ly
    dx = lx / nx
    dy = ly / ny
    if dx <= 2.0:
        raise ValueError("phi too large for a triangular lattice of unit disks")
    ids = np.arange(n, dtype=np.int64)
    site = ids
    if shuffle:
        # Fisher-Yates-free permutation: argsort of hashed keys (deterministic)
        site = np.argsort(uniform01(seed, ids, 7), kind="stable")
    i = (site % nx).astype(np.float64)
    j = (site // nx)
    amp = jitter * (dx - 2.0)
    x = i * dx + (j % 2) * (dx / 2.0) + amp * (2.0 * uniform01(seed, ids, 0) - 1.0)
    y = j.astype(np.float64) * dy + amp * (2.0 * uniform01(seed, ids, 1) - 1.0)
    x = np.where(x < 0, x + lx, x)
    x = np.where(x >= lx, x - lx, x)
    y = np.where(y < 0, y + ly, y)
    y = np.where(y >= ly, y - ly, y)
    # Box-Muller
    u1 = 1.0 - uniform01(seed, ids, 2)
    u2 = uniform01(seed, ids, 3)
    rr = np.sqrt(-2.0 * np.log(u1))
    vx = rr * np.cos(2.0 * math.pi * u2)
    vy = rr * np.sin(2.0 * math.pi * u2)
    vx -= vx.mean()
    vy -= vy.mean()
    rad = np.ones(n, dtype=np.float64)
    if small_fraction > 0:
        rad = np.where(uniform01(seed, ids, 4) < small_fraction, size_ratio, 1.0)
    return dict(n=n, lx=lx, ly=ly, phi=phi, seed=seed,
                x=np.ascontiguousarray(x), y=np.ascontiguousarray(y),
                vx=np.ascontiguousarray(vx), vy=np.ascontiguousarray(vy),
                rad=np.ascontiguousarray(rad.astype(np.float64)))


def growth_config(n: int, phi: float, seed: int) -> dict:
    """State like the reference's default start (random points, rad = 0,
    growing at vr; src/EDMD.c:1690-1829) but on a jittered lattice so a t>0
    snapshot is overlap-free: radii already grown to 60 % with growth rate vr."""
    cfg = lattice_config(n, phi, seed, shuffle=True)
    vr = 0.1
    cfg["rad"] = np.full(cfg["n"], 0.6)
    cfg["vr"] = np.full(cfg["n"], vr)
    cfg["t"] = 0.6 / vr
    return cfg


def tiled_config(base: dict, k: int, seed: int, *, shuffle: bool = True) -> dict:
    """k x k periodic tiling of a periodic hard-disk configuration `base` (keys x, y,
    rad, lx, ly: e.g. the reference-grown liquid of tests/golden/liquid_*.npz, SURVEY.md
    8d second input family).  Copies are shifted by whole box lengths, so every
    distance of the tiling is a distance (or a periodic image distance) of the base:
    no overlaps are created.  Fresh Maxwell velocities (unit temperature, zero total
    momentum) from the counter-based generator; shuffle=True permutes the particle ids.
    The radii are taken as they are -- a reference-grown system holds ~10 distinct radii
    within 1e-15 of each other (its growth phase rounds every disk on its own)."""
    bx, by, brad = (np.asarray(base[key], dtype=np.float64) for key in ("x", "y", "rad"))
    l0x, l0y = float(base["lx"]), float(base["ly"])
    n0 = len(bx)
    n = n0 * k * k
    ids = np.arange(n, dtype=np.int64)
    src = ids
    if shuffle:
        src = np.argsort(uniform01(seed, ids, 7), kind="stable")
    p, tile = src % n0, src // n0
    x = bx[p] + (tile % k).astype(np.float64) * l0x
    y = by[p] + (tile // k).astype(np.float64) * l0y
    lx, ly = k * l0x, k * l0y
    # a sum like 3 * l0x + x can round up to the box length itself
    x = np.where(x >= lx, np.nextafter(lx, 0.0), x)
    y = np.where(y >= ly, np.nextafter(ly, 0.0), y)
    u1 = 1.0 - uniform01(seed, ids, 2)
    u2 = uniform01(seed, ids, 3)
    rr = np.sqrt(-2.0 * np.log(u1))
    vx = rr * np.cos(2.0 * math.pi * u2)
    vy = rr * np.sin(2.0 * math.pi * u2)
    vx -= vx.mean()
    vy -= vy.mean()
    rad = brad[p]
    phi = float(math.pi * np.sum(rad * rad) / (lx * ly))
    return dict(n=n, lx=lx, ly=ly, phi=phi, seed=seed,
                x=np.ascontiguousarray(x), y=np.ascontiguousarray(y),
                vx=np.ascontiguousarray(vx), vy=np.ascontiguousarray(vy),
                rad=np.ascontiguousarray(rad))
