"""Row-slab decomposition of the cell grid for N GPUs (one process per GPU).

The prediction sweep and psi6 of a particle depend only on the 3x3 cell block
around it, so the cell grid is cut into contiguous row slabs (row-major cell
index Y*Nxcells + X, src/EDMD.c:2071, makes a slab one contiguous cell range;
periodic in y).  Each rank holds its owned particles plus a ONE-CELL-ROW halo
copied from its two neighbours; sweep outputs are disjoint per rank (no
collective), g(r) counts and psi6 sums are all-reduced.

This module is the host-side logic: which rows / particles belong to which
rank, the halo exchange over ``torch.distributed`` (NCCL on device buffers in
production, gloo on CPU tensors in the tests) and the reassembly of results.
"""
from __future__ import annotations

import numpy as np

HALO_REC = np.dtype([("x", "f8"), ("y", "f8"), ("vx", "f8"), ("vy", "f8"), ("rad", "f8"),
                     ("gid", "i4"), ("cell", "i4")])
assert HALO_REC.itemsize == 48


def slab_rows(ny: int, world: int) -> list[tuple[int, int]]:
    """Balanced contiguous row ranges [lo, hi); needs >= 1 row per rank and ny >= 3."""
    if world < 1 or ny < max(3, world):
        raise ValueError("grid too small for this many slabs")
    edges = [(ny * r) // world for r in range(world + 1)]
    return [(edges[r], edges[r + 1]) for r in range(world)]


def owner_of_rows(ny: int, world: int) -> np.ndarray:
    own = np.empty(ny, np.int32)
    for r, (lo, hi) in enumerate(slab_rows(ny, world)):
        own[lo:hi] = r
    return own


def owned_indices(cell_y: np.ndarray, ny: int, world: int, rank: int) -> np.ndarray:
    lo, hi = slab_rows(ny, world)[rank]
    return np.nonzero((cell_y >= lo) & (cell_y < hi))[0].astype(np.int32)


def neighbours(rank: int, world: int) -> tuple[int, int]:
    """(lower, upper) neighbour ranks, periodic."""
    return (rank - 1) % world, (rank + 1) % world


def boundary_records(cfg: dict, cells: np.ndarray, gid: np.ndarray, row: int) -> np.ndarray:
    """Host-side packing of one boundary row (what edmd_cuda_halo_pack does on the device)."""
    sel = np.nonzero(cells[gid, 1] == row)[0]
    g = gid[sel]
    rec = np.zeros(len(g), HALO_REC)
    for k in ("x", "y", "vx", "vy", "rad"):
        rec[k] = cfg[k][g]
    rec["gid"] = g
    rec["cell"] = cells[g, 0] + 1   # padded column
    return rec


def exchange_halo(send_lo, send_hi, rank: int, world: int, dist, device=None):
    """Send `send_lo` (first owned row) to the lower neighbour and `send_hi`
    (last owned row) to the upper one; return (recv_from_lower, recv_from_upper).
    Arguments are torch uint8 tensors (48-byte records), on the CPU for gloo or
    on the GPU for NCCL.  Counts travel first, then the payloads, as point-to-
    point messages (two small messages per neighbour; latency-bound)."""
    import torch
    lower, upper = neighbours(rank, world)
    dev = send_lo.device if device is None else device
    n_lo = torch.tensor([send_lo.numel()], dtype=torch.int64, device=dev)
    n_hi = torch.tensor([send_hi.numel()], dtype=torch.int64, device=dev)
    r_lo = torch.zeros(1, dtype=torch.int64, device=dev)
    r_hi = torch.zeros(1, dtype=torch.int64, device=dev)
    # tags are implied by order: (to lower, to upper) / (from upper, from lower)
    ops = [dist.P2POp(dist.isend, n_lo, lower), dist.P2POp(dist.isend, n_hi, upper),
           dist.P2POp(dist.irecv, r_hi, upper), dist.P2POp(dist.irecv, r_lo, lower)]
    if world == 2:
        # both neighbours are the same rank: keep (send first-row, recv its first-row) paired
        ops = [dist.P2POp(dist.isend, n_lo, lower), dist.P2POp(dist.irecv, r_hi, upper),
               dist.P2POp(dist.isend, n_hi, upper), dist.P2POp(dist.irecv, r_lo, lower)]
    for w in dist.batch_isend_irecv(ops):
        w.wait()
    from_lo = torch.empty(int(r_lo.item()), dtype=torch.uint8, device=dev)
    from_hi = torch.empty(int(r_hi.item()), dtype=torch.uint8, device=dev)
    ops = []
    pairs = [(dist.isend, send_lo, lower), (dist.irecv, from_hi, upper),
             (dist.isend, send_hi, upper), (dist.irecv, from_lo, lower)]
    for fn, t, peer in pairs:
        if t.numel() > 0:
            ops.append(dist.P2POp(fn, t, peer))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return from_lo, from_hi


class SlabRank:
    """One rank of the slab decomposition on a GPU: owns a slab context of the
    C ABI, torch only carries the halo buffers through NCCL."""

    def __init__(self, pkg, n_total: int, lx: float, ly: float, rank: int, world: int, device: int,
                 capacity: int | None = None, halo_capacity: int | None = None):
        import torch
        self.torch = torch
        self.rank, self.world, self.device = rank, world, device
        probe = pkg.binding.EdmdCuda.__new__(pkg.binding.EdmdCuda)  # box constants without a context
        ny = int(ly / 2)
        self.ny = ny
        self.rows = slab_rows(ny, world)[rank]
        nrows = self.rows[1] - self.rows[0]
        per_row = n_total / ny
        if halo_capacity is None:
            halo_capacity = int(per_row * 1.5) + 1024
        if capacity is None:
            capacity = int(per_row * nrows * 1.1) + 2 * halo_capacity + 4096
        self.halo_capacity = halo_capacity
        self.ctx = pkg.binding.EdmdCuda(capacity, lx, ly, device=device, slab_rows=self.rows)
        dev = torch.device("cuda", device)
        self.send = [torch.empty(halo_capacity * HALO_REC.itemsize, dtype=torch.uint8, device=dev)
                     for _ in range(2)]
        self.p2p = False
        del probe

    def connect_p2p(self, dist):
        """Map the neighbours' halo inboxes through CUDA IPC so that the exchange
        becomes peer stores over NVLink issued by the pack kernel itself."""
        mine = self.ctx.halo_export(self.halo_capacity)
        if self.world == 1:
            self.ctx.halo_connect(None, None)
        else:
            handles = [None] * self.world
            dist.all_gather_object(handles, mine)
            lower, upper = neighbours(self.rank, self.world)
            self.ctx.halo_connect(handles[lower], handles[upper])
            dist.barrier()
        self.p2p = True

    def close(self):
        self.ctx.close()

    def load_owned(self, cfg: dict, cells: np.ndarray, t: float):
        """cfg holds the WHOLE system on the host; this rank uploads only what it owns."""
        gid = owned_indices(cells[:, 1], self.ny, self.world, self.rank)
        self.gid = gid
        self.ctx.upload_owned(cfg["x"][gid], cfg["y"][gid], cfg["vx"][gid], cfg["vy"][gid],
                              cfg["rad"][gid], cells[gid], gid, t=t)
        return gid

    def exchange(self, dist):
        """One-cell-row halo exchange with the two neighbouring slabs over NCCL."""
        torch = self.torch
        if self.p2p:
            self.ctx.halo_exchange()        # asynchronous, on the context's stream
            return 2 * self.halo_capacity * HALO_REC.itemsize
        n_lo = self.ctx.halo_pack(0, self.send[0].data_ptr(), self.halo_capacity)
        n_hi = self.ctx.halo_pack(1, self.send[1].data_ptr(), self.halo_capacity)
        s_lo = self.send[0][: n_lo * HALO_REC.itemsize]
        s_hi = self.send[1][: n_hi * HALO_REC.itemsize]
        if self.world == 1:
            from_lo, from_hi = s_hi, s_lo      # periodic: my own last row is below my first
        else:
            from_lo, from_hi = exchange_halo(s_lo, s_hi, self.rank, self.world, dist)
            torch.cuda.synchronize()
        self.ctx.halo_append(0, from_lo.data_ptr(), from_lo.numel() // HALO_REC.itemsize)
        self.ctx.halo_append(1, from_hi.data_ptr(), from_hi.numel() // HALO_REC.itemsize)
        return (n_lo + n_hi) * HALO_REC.itemsize

    def predict(self):
        return self.ctx.predict_all()

    def exchange_predict(self, dist=None):
        """Halo exchange + sweep.  With the peer-to-peer halo this is ONE stream-ordered sequence on
        the device (edmd_cuda_exchange_predict_device): send, partition of the owned particles,
        receive + partition of the neighbours' rows, sweep kernel."""
        if self.p2p:
            self.ctx.exchange_predict_device()
            return self.ctx.fetch_predictions()
        self.exchange(dist)
        return self.predict()

    def boop(self, dist, n_total: int, r_c: float = 2.5):
        """psi6 of this rank's particles (computeBOOPCutoff, src/boop.c:61-107; the halo rows of the
        last exchange supply the neighbours across the slab boundary) and the mean q6 of the WHOLE
        system (the thermo column, src/EDMD.c:5521-5536): the per-rank sums are all-reduced."""
        torch = self.torch
        b = self.ctx.boop_cutoff(r_c)
        n_own = len(b["q6"])
        part = torch.tensor([b["mean_q6"] * n_own, float(n_own)], dtype=torch.float64,
                            device=torch.device("cuda", self.device))
        if self.world > 1:
            dist.all_reduce(part)
        tot = part.cpu().numpy()
        assert int(round(tot[1])) == n_total
        b["mean_q6_global"] = float(tot[0] / tot[1])
        return b

    def pcf(self, dist, x_owned, y_owned, n_total: int, dr: float, max_r: float):
        """g(r) of the whole system across the ranks (calculate_pcf, src/pcf.c:16-75):
        the owned positions are all-gathered (NCCL), every rank sorts them into the
        same tiles and bins the tile pairs w = rank (mod world)
        (edmd_cuda_pcf_device), the integer counts are all-reduced.  Returns the
        counts (uint64, every rank) and the device time of the three stages in ms."""
        torch = self.torch
        dev = torch.device("cuda", self.device)
        own_xy = torch.from_numpy(np.stack([x_owned, y_owned], 1)).to(dev)
        sizes = [None] * self.world
        if self.world > 1:
            dist.all_gather_object(sizes, int(own_xy.shape[0]))
        else:
            sizes = [int(own_xy.shape[0])]
        assert sum(sizes) == n_total
        cap = max(sizes)
        pad = torch.zeros(cap, 2, dtype=torch.float64, device=dev)
        pad[: own_xy.shape[0]] = own_xy
        nb = int(max_r / dr)
        counts = torch.zeros(max(nb, 1), dtype=torch.int64, device=dev)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier()
        ev[0].record()
        if self.world > 1:
            allpad = torch.empty(self.world * cap, 2, dtype=torch.float64, device=dev)
            dist.all_gather_into_tensor(allpad, pad)
            xy = torch.cat([allpad[r * cap: r * cap + sizes[r]] for r in range(self.world)]).contiguous()
        else:
            xy = own_xy.contiguous()
        torch.cuda.synchronize()      # the library bins on its own stream
        self.ctx.pcf_device(xy.data_ptr(), n_total, dr, max_r, self.rank, self.world, counts.data_ptr())
        if self.world > 1:
            dist.all_reduce(counts)
        ev[1].record()
        torch.cuda.synchronize()
        return counts[:nb].cpu().numpy().astype(np.uint64), ev[0].elapsed_time(ev[1])
