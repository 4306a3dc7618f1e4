// cell_index.cu -- K0: the reference's intrusive doubly linked cell list
// (cellListInit / addToCell, src/EDMD.c:1906-1920, 2053-2078) rebuilt on the
// device as a counting-sort cell index.
//
//   pack     : SoA upload staging -> resident records, cell id per particle
//              (coordToCell, src/EDMD.c:2098-2107, or the host's cell[2])
//   count    : histogram of cell ids
//   scan     : single-pass decoupled look-back exclusive scan over the cells
//   bucket   : particle ids into their cell's slot range (atomic cursor; the
//              histogram is counted back down to zero = self-cleaning)
//   gather   : rank-sort each cell's ids DESCENDING (the reference's list
//              order) and gather the state into cell order
//
// Everything here is integer / data movement: HBM- and L2-bound.
#include "edmd_internal.cuh"

namespace {

constexpr int kThreads = 256;

// ------------------------------------------------------------------ pack --
__global__ void __launch_bounds__(kThreads)
k_pack(int n, edmd_dev_box b, const double *__restrict__ soa,
       const int32_t *__restrict__ cell_xy, double4 *__restrict__ xv,
       double *__restrict__ rad, int32_t *__restrict__ cid,
       int32_t *__restrict__ flags)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    size_t N = (size_t)n;
    double x = soa[i], y = soa[N + i];
    xv[i] = make_double4(x, y, soa[2 * N + i], soa[3 * N + i]);
    rad[i] = soa[4 * N + i];
    int X, Y;
    if (cell_xy) {
        int2 c = reinterpret_cast<const int2 *>(cell_xy)[i];
        X = c.x;
        Y = c.y;
    } else {
        // coordToCell: multiply by the reciprocal, truncate toward zero
        X = (int)__dmul_rn(x, b.fx);
        Y = (int)__dmul_rn(y, b.fy);
    }
    if (X < 0 || X >= b.nx || Y < 0 || Y >= b.ny) {
        atomicOr(flags, 1);
        X = min(max(X, 0), b.nx - 1);
        Y = min(max(Y, 0), b.ny - 1);
    }
    cid[i] = Y * b.nx + X;
}

// ----------------------------------------------------------------- count --
__global__ void __launch_bounds__(kThreads)
k_count(int n, const int32_t *__restrict__ cid, int32_t *__restrict__ cnt)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) atomicAdd(&cnt[cid[i]], 1);
}

// ------------------------------------------------------------------ scan --
// Exclusive scan of cnt[0..nc) into start[0..nc]; start[nc] = total.
// Single pass, decoupled look-back.  Tile status word: 2 flag bits | 30 value
// bits (N < 2^30).  Tiles take tickets from an atomic counter so a tile never
// waits on one that has not started.  `state_next` / `ticket_next` belong to
// the NEXT sweep and are cleared here (ping-pong) so no memset is needed.
constexpr int kScanThreads = 256;
constexpr int kScanItems = 16;
constexpr int kScanTile = kScanThreads * kScanItems;
constexpr uint32_t kFlagAgg = 1u << 30;
constexpr uint32_t kFlagInc = 2u << 30;
constexpr uint32_t kValMask = (1u << 30) - 1;

__global__ void __launch_bounds__(kScanThreads)
k_scan(int nc, const int32_t *__restrict__ cnt, int32_t *__restrict__ start,
       uint32_t *state, int32_t *ticket, uint32_t *state_next,
       int32_t *ticket_next, int tiles)
{
    __shared__ int s_tile;
    __shared__ int s_warp[kScanThreads / 32];
    __shared__ int s_prefix;
    const int tid = threadIdx.x;
    if (tid == 0) s_tile = atomicAdd(ticket, 1);
    __syncthreads();
    const int tile = s_tile;

    // clear next sweep's bookkeeping (one tile does it)
    if (tile == 0) {
        for (int k = tid; k < tiles; k += kScanThreads) state_next[k] = 0;
        if (tid == 0) *ticket_next = 0;
    }

    const int base = tile * kScanTile + tid * kScanItems;
    int v[kScanItems];
    int sum = 0;
#pragma unroll
    for (int q = 0; q < kScanItems / 4; q++) {
        int idx = base + 4 * q;
        int4 w;
        if (idx + 3 < nc) {
            w = *reinterpret_cast<const int4 *>(cnt + idx);
        } else {
            w.x = idx < nc ? cnt[idx] : 0;
            w.y = idx + 1 < nc ? cnt[idx + 1] : 0;
            w.z = idx + 2 < nc ? cnt[idx + 2] : 0;
            w.w = 0;
        }
        v[4 * q] = w.x; v[4 * q + 1] = w.y; v[4 * q + 2] = w.z; v[4 * q + 3] = w.w;
        sum += w.x + w.y + w.z + w.w;
    }
    // block exclusive scan of the per-thread sums
    int incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int o = __shfl_up_sync(0xffffffffu, incl, d);
        if ((tid & 31) >= d) incl += o;
    }
    if ((tid & 31) == 31) s_warp[tid >> 5] = incl;
    __syncthreads();
    if (tid < 32) {
        int w = tid < kScanThreads / 32 ? s_warp[tid] : 0;
        int wi = w;
#pragma unroll
        for (int d = 1; d < kScanThreads / 32; d <<= 1) {
            int o = __shfl_up_sync(0xffffffffu, wi, d);
            if (tid >= d) wi += o;
        }
        if (tid < kScanThreads / 32) s_warp[tid] = wi - w;  // exclusive
        int total = __shfl_sync(0xffffffffu, wi, kScanThreads / 32 - 1);
        if (tid == 0) {
            volatile uint32_t *vs = state;
            int prefix = 0;
            if (tile == 0) {
                vs[0] = kFlagInc | (uint32_t)total;
            } else {
                vs[tile] = kFlagAgg | (uint32_t)total;
                __threadfence();
                int p = tile - 1;
                while (true) {
                    uint32_t s = vs[p];
                    if ((s >> 30) == 0) continue;  // predecessor not published yet
                    prefix += (int)(s & kValMask);
                    if (s & kFlagInc) break;
                    p--;
                }
                vs[tile] = kFlagInc | (uint32_t)(prefix + total);
            }
            s_prefix = prefix;
            if (tile == tiles - 1) start[nc] = prefix + total;
        }
    }
    __syncthreads();
    int run = s_prefix + s_warp[tid >> 5] + (incl - sum);
#pragma unroll
    for (int q = 0; q < kScanItems; q++) {
        int idx = base + q;
        if (idx < nc) start[idx] = run;
        run += v[q];
    }
}

// ---------------------------------------------------------------- bucket --
__global__ void __launch_bounds__(kThreads)
k_bucket(int n, const int32_t *__restrict__ cid, int32_t *__restrict__ cnt,
         const int32_t *__restrict__ start, int32_t *__restrict__ slot_id)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int c = cid[i];
    int r = atomicSub(&cnt[c], 1) - 1;  // counts back down to zero
    slot_id[start[c] + r] = i;
}

// ---------------------------------------------------------------- gather --
template <bool GROW>
__global__ void __launch_bounds__(kThreads)
k_gather(int n, const int32_t *__restrict__ slot_id,
         const int32_t *__restrict__ cid, const int32_t *__restrict__ start,
         const double4 *__restrict__ xv, const double *__restrict__ rad,
         const double *__restrict__ vr, double4 *__restrict__ sxv,
         double *__restrict__ srad, double *__restrict__ svr,
         int32_t *__restrict__ sid, int32_t *__restrict__ scid)
{
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    int id = slot_id[s];
    int c = cid[id];
    int lo = start[c], hi = start[c + 1];
    int rank = 0;  // ids in this cell larger than mine come first
    for (int p = lo; p < hi; p++) rank += (slot_id[p] > id);
    int d = lo + rank;
    sxv[d] = xv[id];
    srad[d] = rad[id];
    if (GROW) svr[d] = vr[id];
    sid[d] = id;
    scid[d] = c;
}

}  // namespace

int edmd_launch_pack(edmd_ctx *c, bool have_cells)
{
    int n = c->n;
    if (n == 0) return 0;
    k_pack<<<(n + kThreads - 1) / kThreads, kThreads, 0, c->stream>>>(
        n, c->dbox, c->in_soa, have_cells ? c->in_cell : nullptr, c->xv, c->rad,
        c->cid, c->flags);
    return 1;
}

int edmd_launch_cell_index(edmd_ctx *c, int mode)
{
    int n = c->n;
    int nc = c->dbox.nc;
    int blocks = (n + kThreads - 1) / kThreads;
    int par = c->scan_parity;
    int launched = 0;
    if (n > 0) {
        k_count<<<blocks, kThreads, 0, c->stream>>>(n, c->cid, c->cell_cnt);
        launched++;
    }
    k_scan<<<c->scan_tiles, kScanThreads, 0, c->stream>>>(
        nc, c->cell_cnt, c->cell_start, c->scan_state[par],
        c->scan_ticket + par, c->scan_state[par ^ 1], c->scan_ticket + (par ^ 1),
        c->scan_tiles);
    launched++;
    c->scan_parity = par ^ 1;
    if (n > 0) {
        k_bucket<<<blocks, kThreads, 0, c->stream>>>(n, c->cid, c->cell_cnt,
                                                     c->cell_start, c->slot_id);
        if (mode == EDMD_MODE_GROW)
            k_gather<true><<<blocks, kThreads, 0, c->stream>>>(
                n, c->slot_id, c->cid, c->cell_start, c->xv, c->rad, c->vr,
                c->sxv, c->srad, c->svr, c->sid, c->scid);
        else
            k_gather<false><<<blocks, kThreads, 0, c->stream>>>(
                n, c->slot_id, c->cid, c->cell_start, c->xv, c->rad, c->vr,
                c->sxv, c->srad, c->svr, c->sid, c->scid);
        launched += 2;
    }
    return launched;
}

int edmd_scan_tiles_for(int nc) { return (nc + kScanTile - 1) / kScanTile; }
