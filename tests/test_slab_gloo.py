"""CPU tests of the multi-GPU (N > 1) host logic with torch.distributed/gloo:
slab ownership, the one-cell-row halo exchange and result reassembly.  The
per-rank compute is done by the oracle here (no GPU in this suite); the check
is that "owned + halo" is enough to reproduce the whole-system sweep bit for
bit on every owned particle, for 2 and 3 ranks, including the periodic wrap; and that psi6 of the owned
particles plus the all-reduce of the slabs' q6 sums give the whole system's values and mean."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, phi, seed, out_dir):
    sys.path.insert(0, str(ROOT))
    import torch
    import torch.distributed as dist
    import __graft_entry__ as entry
    from oracle.oracle_py import Oracle

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg = entry.load_package()
    slab = pkg.slab
    orc = Oracle()
    cfg = pkg.synth.lattice_config(n, phi, seed, small_fraction=0.3)
    N, lx, ly, t = cfg["n"], cfg["lx"], cfg["ly"], 2.5
    cells = orc.cells(N, lx, ly, cfg["x"], cfg["y"]).reshape(N, 2)
    box = orc.box(N, lx, ly)
    lo, hi = slab.slab_rows(box.ny, world)[rank]
    gid = slab.owned_indices(cells[:, 1], box.ny, world, rank)
    # pack my first / last owned rows and exchange them
    send_lo = slab.boundary_records(cfg, cells, gid, lo)
    send_hi = slab.boundary_records(cfg, cells, gid, hi - 1)
    as_t = lambda rec: torch.from_numpy(rec.view(np.uint8).copy())
    from_lo, from_hi = slab.exchange_halo(as_t(send_lo), as_t(send_hi), rank, world, dist)
    halo_lo = from_lo.numpy().view(slab.HALO_REC)
    halo_hi = from_hi.numpy().view(slab.HALO_REC)
    # the halo must be exactly the neighbours' boundary rows
    row_below, row_above = (lo - 1) % box.ny, hi % box.ny
    assert set(halo_lo["gid"]) == set(np.nonzero(cells[:, 1] == row_below)[0])
    assert set(halo_hi["gid"]) == set(np.nonzero(cells[:, 1] == row_above)[0])
    # local system = owned + halo, cells stay global
    loc = {k: np.concatenate([cfg[k][gid], halo_lo[k], halo_hi[k]]) for k in ("x", "y", "vx", "vy", "rad")}
    lgid = np.concatenate([gid, halo_lo["gid"], halo_hi["gid"]])
    lcells = np.concatenate([cells[gid],
                             np.stack([halo_lo["cell"] - 1, np.full(len(halo_lo), row_below)], 1),
                             np.stack([halo_hi["cell"] - 1, np.full(len(halo_hi), row_above)], 1)]).astype(np.int32)
    res = orc.predict_all(len(lgid), lx, ly, t, loc["x"], loc["y"], loc["vx"], loc["vy"], loc["rad"],
                          cell_xy=lcells)
    no = len(gid)
    has = res["t_coll"][:no] < 1e25
    partner = np.where(has, lgid[res["partner"][:no]], 0)
    np.savez(Path(out_dir) / f"rank{rank}.npz", gid=gid, t_cross=res["t_cross"][:no], dir=res["dir"][:no],
             t_coll=res["t_coll"][:no], partner=partner)
    # psi6 (computeBOOPCutoff, src/boop.c:61-107): owned + halo rows hold every neighbour of an owned particle;
    # the mean q6 of the thermo column (src/EDMD.c:5521-5536) is the all-reduced sum of the slabs' sums over N
    bo = orc.boop_cutoff(len(lgid), lx, ly, loc["x"], loc["y"], 2.5, cell_xy=lcells)
    sums = torch.tensor([float(bo["q6"][:no].sum()), float(no)], dtype=torch.float64)
    dist.all_reduce(sums)
    np.savez(Path(out_dir) / f"boop{rank}.npz", gid=gid, q6=bo["q6"][:no], q6_arg=bo["q6_arg"][:no],
             neighbors=bo["neighbors"][:no], mean_q6=float(sums[0] / sums[1]), n_total=float(sums[1]))
    # g(r) (calculate_pcf, src/pcf.c:34-54): the positions go to every rank, rank k bins the pairs of the tile
    # pairs w = k (mod world) -- tiles of 256 particles, the upper triangle walked row by row, as the library
    # deals them (edmd_cuda_pcf_device) --, the integer histograms are all-reduced
    dr, max_r = 0.1, min(lx, ly) / 2
    nb = int(max_r / dr)
    tile = 256
    nt = (N + tile - 1) // tile
    counts = torch.zeros(nb, dtype=torch.int64)
    w = 0
    for ta in range(nt):
        for tb in range(ta, nt):
            if w % world == rank:
                ia = np.arange(ta * tile, min(N, (ta + 1) * tile))
                ib = np.arange(tb * tile, min(N, (tb + 1) * tile))
                ii, jj = np.meshgrid(ia, ib, indexing="ij")
                keep = ii < jj
                ii, jj = ii[keep], jj[keep]
                dx, dy = cfg["x"][jj] - cfg["x"][ii], cfg["y"][jj] - cfg["y"][ii]
                dx = np.where(dx >= lx / 2, dx - lx, np.where(dx < -lx / 2, dx + lx, dx))
                dy = np.where(dy >= ly / 2, dy - ly, np.where(dy < -ly / 2, dy + ly, dy))
                r = np.sqrt(dx * dx + dy * dy)
                b = (r / dr).astype(np.int64)
                ok = (r < max_r) & (b < nb)
                counts += torch.from_numpy(np.bincount(b[ok], minlength=nb))
            w += 1
    mine = int(counts.sum())
    dist.all_reduce(counts)
    if rank == 0:
        np.savez(Path(out_dir) / "pcf.npz", counts=counts.numpy(), share0=mine)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_slab_partition_and_halo_reproduce_the_global_sweep(tmp_path, world):
    import torch.multiprocessing as mp
    sys.path.insert(0, str(ROOT))
    import __graft_entry__ as entry
    from oracle.oracle_py import Oracle
    n, phi, seed = 6000, 0.70, 17
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n, phi, seed, str(tmp_path)), nprocs=world, join=True)
    pkg = entry.load_package()
    orc = Oracle()
    cfg = pkg.synth.lattice_config(n, phi, seed, small_fraction=0.3)
    want = orc.predict_all(cfg["n"], cfg["lx"], cfg["ly"], 2.5, cfg["x"], cfg["y"], cfg["vx"], cfg["vy"],
                           cfg["rad"])
    seen = np.zeros(cfg["n"], bool)
    for r in range(world):
        z = np.load(tmp_path / f"rank{r}.npz")
        g = z["gid"]
        assert not seen[g].any()
        seen[g] = True
        for k in ("t_cross", "dir", "t_coll", "partner"):
            assert np.array_equal(z[k], want[k][g]), (r, k)
    assert seen.all()
    # g(r): the all-reduced histogram of the ranks' shares is calculate_pcf's
    z = np.load(tmp_path / "pcf.npz")
    wg = orc.pcf(cfg["n"], cfg["lx"], cfg["ly"], cfg["x"], cfg["y"], 0.1, min(cfg["lx"], cfg["ly"]) / 2)
    assert np.array_equal(z["counts"].astype(np.uint64), wg["counts"])
    assert 0 < int(z["share0"]) < int(wg["counts"].sum())
    # psi6 per slab + the all-reduced mean
    wb = orc.boop_cutoff(cfg["n"], cfg["lx"], cfg["ly"], cfg["x"], cfg["y"], 2.5)
    for r in range(world):
        z = np.load(tmp_path / f"boop{r}.npz")
        g = z["gid"]
        assert np.array_equal(z["neighbors"], wb["neighbors"][g])
        assert np.abs(z["q6"] - wb["q6"][g]).max() <= 1e-10
        assert z["n_total"] == cfg["n"]
        assert abs(float(z["mean_q6"]) - wb["q6"].mean()) <= 1e-10


def test_slab_rows_are_balanced_and_cover_the_grid():
    sys.path.insert(0, str(ROOT))
    import __graft_entry__ as entry
    slab = entry.load_package().slab
    for ny, world in [(985, 1), (985, 2), (985, 8), (17, 5), (8, 8)]:
        rows = slab.slab_rows(ny, world)
        assert rows[0][0] == 0 and rows[-1][1] == ny
        assert all(a[1] == b[0] for a, b in zip(rows, rows[1:]))
        sizes = [b - a for a, b in rows]
        assert max(sizes) - min(sizes) <= 1 and min(sizes) >= 1
    with pytest.raises(ValueError):
        slab.slab_rows(2, 2)
