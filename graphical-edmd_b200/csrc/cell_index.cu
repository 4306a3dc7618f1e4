// cell_index.cu -- K0: the reference's intrusive doubly linked cell list
// (cellListInit / addToCell, src/EDMD.c:1906-1920, 2053-2078) rebuilt on the
// device as a counting sort over the padded cell grid described in
// edmd_internal.cuh.
//
//   pack     (upload)  SoA staging -> resident records; padded cell id per
//                      particle (coordToCell, src/EDMD.c:2098-2107, or the
//                      host's cell[2]); counts ghost entries; flags particles
//                      that are not near the cell they are filed under
//   count    histogram of cell ids; the atomic's return value is the
//            particle's arrival rank inside its cell
//   rowscan  one CTA per cell row: exclusive scan of the padded row (ghost
//            cells take the count of the cell they mirror), row total, and the
//            histogram is zeroed again (no memset between sweeps)
//   rowbase  one CTA: scan of the row totals rounded up to 32
//   chunkmeta one CTA per row: the staging plan (ChunkMeta) of every 32-slot chunk
//   scatter  every particle writes its 32+16-byte record (and its ghost copy when
//            it sits in an edge cell) to row_base + off + rank
//
// Integer / data movement only: HBM- and L2-bound.
#include "edmd_internal.cuh"
#include "rowstage.cuh"

namespace {

constexpr int kThreads = 256;

// ------------------------------------------------------------------ pack --
__global__ void __launch_bounds__(kThreads)
k_pack(int n, size_t N, edmd_dev_box b, int ps, const double *__restrict__ soa,
       const int32_t *__restrict__ cell_xy, double4 *__restrict__ xv,
       double *__restrict__ rad, int32_t *__restrict__ cid,
       int32_t *__restrict__ flags, double rad0, int keep_rad,
       int32_t *__restrict__ halo_list, int32_t *__restrict__ halo_cnt, int halo_cap, int first)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int ghosts = 0, insane = 0;
    float vm = 0.0f;
    if (i < n) {
        double x = soa[i], y = soa[N + i];
        const double vx = soa[2 * N + i], vy = soa[3 * N + i];
        const double r = keep_rad ? rad[i] : soa[4 * N + i];   // keep_rad: radii unchanged since the last upload
        xv[i] = make_double4(x, y, vx, vy);
        if (!keep_rad) rad[i] = r;
        // what the lean sweep needs to know (lean.cuh): one common radius, the speed scale
        edmd_note_radius(flags, r, rad0);
        vm = __double2float_ru(fmax(fabs(vx), fabs(vy)));
        if (!(vm == vm)) vm = __int_as_float(0x7f800000);
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
            atomicOr(&flags[kFlagBadCell], 1);
            X = min(max(X, 0), b.nx - 1);
            Y = min(max(Y, 0), b.ny - 1);
        }
        // within one cell width of the filed cell?  (lets the sweep skip the
        // minimum-image test away from the periodic edges)
        insane = !(fabs(x - ((double)X + 0.5) * b.csx) <= 1.5 * b.csx) ||
                 !(fabs(y - ((double)Y + 0.5) * b.csy) <= 1.5 * b.csy);
        int l = Y - b.yoff;   // row inside this context (slab: owned rows + halo)
        if (l < 0) l += b.ny;
        if (l >= b.nl) {
            atomicOr(&flags[kFlagBadCell], 1);
            l = b.nl - 1;
        }
        cid[i] = l * ps + X + 1;
        ghosts = (X == 0) + (X == b.nx - 1);
        // slab contexts with peer-to-peer halo: list the particles of the two boundary
        // rows now, so that every exchange until the next upload can skip that pass
        if (halo_list && (l == 1 || l == b.nl - 2)) {
#pragma unroll
            for (int side = 0; side < 2; side++) {
                if (l != (side == 0 ? 1 : b.nl - 2)) continue;
                const int k = atomicAdd(halo_cnt + side, 1);
                if (k < halo_cap) halo_list[side * halo_cap + k] = first + i;
                else atomicOr(&flags[kFlagBadCell], 2);
            }
        }
    }
    ghosts = __reduce_add_sync(0xffffffffu, ghosts);
    insane = __reduce_add_sync(0xffffffffu, insane);
    const unsigned vmb = __reduce_max_sync(0xffffffffu, (unsigned)__float_as_int(vm));   // vm >= 0: bits order like values
    if ((threadIdx.x & 31) == 0) {
        if (ghosts) atomicAdd(&flags[kFlagGhosts], ghosts);
        if (insane) atomicAdd(&flags[kFlagInsane], insane);
        atomicMax(reinterpret_cast<unsigned int *>(&flags[kFlagVmax]), vmb);
    }
}

// ----------------------------------------------------------------- count --
// four particles per thread: four independent L2 atomics in flight
__global__ void __launch_bounds__(kThreads)
k_count(int n, const int32_t *__restrict__ cid, int32_t *__restrict__ cnt,
        int32_t *__restrict__ rank)
{
    edmd_pdl_wait();
    const int i = 4 * (blockIdx.x * blockDim.x + threadIdx.x);
    if (i + 3 < n) {
        const int4 c = *reinterpret_cast<const int4 *>(cid + i);
        int4 r;
        // cid < 0: unused halo slot of a slab context
        r.x = c.x >= 0 ? atomicAdd(&cnt[c.x], 1) : 0;
        r.y = c.y >= 0 ? atomicAdd(&cnt[c.y], 1) : 0;
        r.z = c.z >= 0 ? atomicAdd(&cnt[c.z], 1) : 0;
        r.w = c.w >= 0 ? atomicAdd(&cnt[c.w], 1) : 0;
        *reinterpret_cast<int4 *>(rank + i) = r;
    } else {
        for (int k = i; k < n; k++)
            if (cid[k] >= 0) rank[k] = atomicAdd(&cnt[cid[k]], 1);
    }
}

// --------------------------------------------------------------- rowscan --
__global__ void __launch_bounds__(kThreads)
k_rowscan(int nx, int ps, int32_t *__restrict__ cnt, int32_t *__restrict__ off,
          int32_t *__restrict__ row_total)
{
    __shared__ int s_warp[kThreads / 32];
    __shared__ int s_carry;
    const int Y = blockIdx.x;
    const int tid = threadIdx.x;
    edmd_pdl_wait();
    int32_t *row = cnt + (size_t)Y * ps;
    int32_t *orow = off + (size_t)Y * ps;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < ps; base += kThreads) {
        const int pcx = base + tid;
        int v = 0;
        if (pcx < ps) {
            if (pcx == 0) v = row[nx];            // left ghost mirrors cell nx-1
            else if (pcx <= nx) v = row[pcx];
            else if (pcx == nx + 1) v = row[1];   // right ghost mirrors cell 0
        }
        int incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int o = __shfl_up_sync(0xffffffffu, incl, d);
            if ((tid & 31) >= d) incl += o;
        }
        if ((tid & 31) == 31) s_warp[tid >> 5] = incl;
        __syncthreads();
        int wbase = 0;
#pragma unroll
        for (int w = 0; w < kThreads / 32; w++)
            if (w < (tid >> 5)) wbase += s_warp[w];
        const int carry = s_carry;
        if (pcx < ps) orow[pcx] = carry + wbase + incl - v;
        __syncthreads();
        if (tid == kThreads - 1) s_carry = carry + wbase + incl;
        __syncthreads();
    }
    if (tid == 0) row_total[Y] = s_carry;
    // histogram back to zero for the next sweep
    for (int pcx = 1 + tid; pcx <= nx; pcx += kThreads) row[pcx] = 0;
}

// --------------------------------------------------------------- rowbase --
constexpr int kBaseThreads = 1024;

__global__ void __launch_bounds__(kBaseThreads)
k_rowbase(int ny, const int32_t *__restrict__ row_total, int32_t *__restrict__ row_base)
{
    __shared__ int s_warp[kBaseThreads / 32];
    __shared__ int s_carry;
    const int tid = threadIdx.x;
    edmd_pdl_wait();
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < ny; base += kBaseThreads) {
        const int Y = base + tid;
        const int v = Y < ny ? ((row_total[Y] + 31) & ~31) : 0;
        int incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int o = __shfl_up_sync(0xffffffffu, incl, d);
            if ((tid & 31) >= d) incl += o;
        }
        if ((tid & 31) == 31) s_warp[tid >> 5] = incl;
        __syncthreads();
        int wbase = 0;
        for (int w = 0; w < (tid >> 5); w++) wbase += s_warp[w];
        const int carry = s_carry;
        if (Y < ny) row_base[Y] = carry + wbase + incl - v;
        __syncthreads();
        if (tid == kBaseThreads - 1) s_carry = carry + wbase + incl;
        __syncthreads();
    }
    if (tid == 0) row_base[ny] = s_carry;
}

// ------------------------------------------------------------- chunkmeta --
// one CTA per cell row: the staging plan of each of the row's 32-slot chunks
constexpr int kMetaThreads = 128;

// smallest column pcx with off[pcx + 1] > slot  (the cell that holds `slot`)
__device__ __forceinline__ int cell_of_slot(const int32_t *__restrict__ o, int ps, int slot)
{
    int lo = 0, hi = ps - 2;  // off[ps-1] = row total > slot (columns past nx+1 are empty)
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (o[mid + 1] > slot) hi = mid;
        else lo = mid + 1;
    }
    return lo;
}

__global__ void __launch_bounds__(kMetaThreads)
k_chunkmeta(int nx, int ny, int ps, int gny, int yoff, int slab, const int32_t *__restrict__ off,
            const int32_t *__restrict__ row_total, const int32_t *__restrict__ row_base,
            ChunkMeta *__restrict__ meta, int32_t *__restrict__ cstart, int max_chunks, int cap_rec,
            int cap_off)
{
    const int Y = blockIdx.x;
    edmd_pdl_wait();
    const int rb = row_base[Y];
    const int tot = row_total[Y];
    const int nch = (tot + 31) >> 5;
    const int32_t *o = off + (size_t)Y * ps;
    // absolute first slot of every padded cell of this row (one gather for the scatter)
    for (int k = threadIdx.x; k < ps; k += kMetaThreads) cstart[(size_t)Y * ps + k] = rb + o[k];
    for (int q = threadIdx.x; q < nch; q += kMetaThreads) {
        ChunkMeta m;
        const int first = 32 * q, last = min(first + 31, tot - 1);
        const int ca = max(cell_of_slot(o, ps, first), 1);
        const int cb = min(cell_of_slot(o, ps, last), nx);
        m.Y = ca <= cb ? Y : -1;   // only ghost entries: nothing to do
        if (slab && (Y == 0 || Y == ny - 1)) m.Y = -1;   // halo rows are never predicted
        m.cfirst = ca;
        m.ncells = cb - ca + 1;
        m.row_end = rb + tot;
        m.flags = 0;
        m.wstart = (ca - 1) & ~3;
        m.wlen = ((cb + 2 - m.wstart + 1) + 3) & ~3;   // columns wstart .. cb+2, rounded up to 4
        if (m.wstart + m.wlen > ps) m.wlen = ps - m.wstart;
        if (m.Y >= 0) {
#pragma unroll
            for (int j = 0; j < 3; j++) {
                int Yr = Y - 1 + j;
                if (Yr < 0) Yr += ny;
                else if (Yr >= ny) Yr -= ny;
                const int rbr = row_base[Yr];
                const int32_t *orow = off + (size_t)Yr * ps;
                const int lo = rbr + orow[ca - 1];
                const int hi = rbr + orow[cb + 2];
                m.seg_lo[j] = lo;
                m.seg_len[j] = hi - lo;
                m.delta[j] = rbr - lo;
                if (hi - lo > cap_rec) m.flags |= kMetaOverflow;
            }
            if (m.wlen > cap_off) m.flags |= kMetaOverflow;
            int Yg = Y + yoff;   // global row: no periodic image in y away from rows 0 and gny-1
            if (Yg >= gny) Yg -= gny;
            if (nx >= 12 && gny >= 12 && Yg >= 1 && Yg <= gny - 2 && ca >= 2 && cb <= nx - 1)
                m.flags |= kMetaInterior;
        } else {
#pragma unroll
            for (int j = 0; j < 3; j++) m.seg_lo[j] = m.seg_len[j] = m.delta[j] = 0;
        }
        const int ch = (rb >> 5) + q;
        if (ch < max_chunks) meta[ch] = m;
    }
    if (Y == ny - 1) {  // chunks past the last row are empty
        ChunkMeta e;
        e.Y = -1; e.cfirst = 0; e.ncells = 0; e.row_end = 0; e.flags = 0; e.wstart = e.wlen = 0;
        for (int j = 0; j < 3; j++) e.seg_lo[j] = e.seg_len[j] = e.delta[j] = 0;
        for (int ch = ((rb + ((tot + 31) & ~31)) >> 5) + threadIdx.x; ch < max_chunks; ch += kMetaThreads)
            meta[ch] = e;
    }
}

// --------------------------------------------------------------- scatter --
__device__ __forceinline__ void put_rec(SPos *__restrict__ spos, SAux *__restrict__ saux,
                                        double *__restrict__ svr, int d, const SPos &p, const SAux &a,
                                        double g, bool grow)
{
    // one full 32-byte sector with a single 256-bit store + one 16-byte tag
    const uint32_t *w = reinterpret_cast<const uint32_t *>(&p);
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(spos + d), "r"(w[0]), "r"(w[1]),
                 "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
                 : "memory");
    *reinterpret_cast<uint4 *>(saux + d) = *reinterpret_cast<const uint4 *>(&a);
    if (grow) svr[d] = g;
}

// two particles per thread (more loads in flight); destination = one gather
// of the cell's absolute start + the particle's arrival rank
template <bool GROW>
__global__ void __launch_bounds__(kThreads)
k_scatter(int n, int nx, int ps, const int32_t *__restrict__ cid,
          const int32_t *__restrict__ rank, const int32_t *__restrict__ cstart,
          const double4 *__restrict__ xv, const double *__restrict__ rad,
          const double *__restrict__ vr, SPos *__restrict__ spos, SAux *__restrict__ saux,
          double *__restrict__ svr)
{
    const int i0 = 2 * (blockIdx.x * blockDim.x + threadIdx.x);
    edmd_pdl_wait();
    if (i0 >= n) return;
    const bool two = i0 + 1 < n;
    int pc[2], rk[2];
    if (two) {
        const int2 c2 = *reinterpret_cast<const int2 *>(cid + i0);
        const int2 r2 = *reinterpret_cast<const int2 *>(rank + i0);
        pc[0] = c2.x; pc[1] = c2.y; rk[0] = r2.x; rk[1] = r2.y;
    } else {
        pc[0] = pc[1] = cid[i0];
        rk[0] = rk[1] = rank[i0];
    }
    int base[2];
    base[0] = pc[0] >= 0 ? cstart[pc[0]] : 0;
    base[1] = pc[1] >= 0 ? cstart[pc[1]] : 0;
    SPos r[2];
    SAux t[2];
    double g[2] = {0.0, 0.0};
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const int i = i0 + (two ? k : 0);
        const double4 p = xv[i];
        r[k].x = p.x; r[k].y = p.y; r[k].vx = p.z; r[k].vy = p.w;
        t[k].rad = rad[i];
        t[k].id = i;
        t[k].pc = pc[k];
        if (GROW) g[k] = vr[i];
    }
#pragma unroll
    for (int k = 0; k < 2; k++) {
        if (k == 1 && !two) break;
        if (pc[k] < 0) continue;   // unused halo slot
        put_rec(spos, saux, svr, base[k] + rk[k], r[k], t[k], g[k], GROW);
        const int Y = pc[k] / ps;
        const int pcx = pc[k] - Y * ps;
        if (pcx == 1) {  // cell 0 is mirrored by the right ghost
            SAux q = t[k];
            q.pc = Y * ps + nx + 1;
            put_rec(spos, saux, svr, cstart[q.pc] + rk[k], r[k], q, g[k], GROW);
        }
        if (pcx == nx) {  // cell nx-1 is mirrored by the left ghost
            SAux q = t[k];
            q.pc = Y * ps;
            put_rec(spos, saux, svr, cstart[q.pc] + rk[k], r[k], q, g[k], GROW);
        }
    }
}

}  // namespace

// pack `count` staged particles into the resident arrays starting at `first`
int edmd_launch_pack(edmd_ctx *c, bool have_cells, int first, int count, bool keep_rad)
{
    if (count == 0) return 0;
    k_pack<<<(count + kThreads - 1) / kThreads, kThreads, 0, c->stream>>>(
        count, (size_t)c->n_cap, c->dbox, c->ps, c->in_soa, have_cells ? c->in_cell : nullptr,
        c->xv + first, c->rad + first, c->cid + first, c->flags, c->rad0, keep_rad ? 1 : 0,
        c->halo_list_at_pack ? c->halo_list : nullptr, c->halo_cnt, c->halo_cap, first);
    return 1;
}

// ---- slab halo: one 48-byte record per particle of a boundary row -------------
struct __align__(16) HaloRec {
    double x, y, vx, vy, rad;
    int gid;
    int cell;   // padded column 1..nx (the row is implied by which neighbour sent it)
};

namespace {
__global__ void __launch_bounds__(kThreads)
k_halo_pack(int n_owned, int ps, int row, const int32_t *__restrict__ cid,
            const double4 *__restrict__ xv, const double *__restrict__ rad,
            const int32_t *__restrict__ gid, HaloRec *__restrict__ out, int cap,
            int32_t *__restrict__ count)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_owned) return;
    const int pc = cid[i];
    if (pc / ps != row) return;
    const int k = atomicAdd(count, 1);
    if (k >= cap) return;
    const double4 p = xv[i];
    HaloRec r;
    r.x = p.x; r.y = p.y; r.vx = p.z; r.vy = p.w;
    r.rad = rad[i];
    r.gid = gid[i];
    r.cell = pc - row * ps;   // padded column 1..nx
    out[k] = r;
}

__global__ void __launch_bounds__(kThreads)
k_halo_append(int count, int first, int ps, int row, const HaloRec *__restrict__ in,
              double4 *__restrict__ xv, double *__restrict__ rad, int32_t *__restrict__ cid,
              int32_t *__restrict__ gid, int nx, int32_t *__restrict__ flags, double rad0)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    const HaloRec r = in[k];
    edmd_note_radius(flags, r.rad, rad0);
    {
        float vm = __double2float_ru(fmax(fabs(r.vx), fabs(r.vy)));
        if (!(vm == vm)) vm = __int_as_float(0x7f800000);
        atomicMax(reinterpret_cast<unsigned int *>(&flags[kFlagVmax]), (unsigned)__float_as_int(vm));
    }
    const int i = first + k;
    xv[i] = make_double4(r.x, r.y, r.vx, r.vy);
    rad[i] = r.rad;
    gid[i] = r.gid;
    cid[i] = row * ps + r.cell;
    const int g = (r.cell == 1) + (r.cell == nx);
    if (g) atomicAdd(&flags[kFlagGhosts], g);
}
}  // namespace

// side 0: first owned row (local row 1) -> lower neighbour; side 1: last owned row -> upper neighbour
int edmd_launch_halo_pack(edmd_ctx *c, int side, void *out, int cap, int32_t *count_dev)
{
    const int n = c->n_owned;
    if (n == 0) return 0;
    const int row = side == 0 ? 1 : c->dbox.nl - 2;
    k_halo_pack<<<(n + kThreads - 1) / kThreads, kThreads, 0, c->stream>>>(
        n, c->ps, row, c->cid, c->xv, c->rad, c->gid, (HaloRec *)out, cap, count_dev);
    return 1;
}

int edmd_launch_halo_append_row(edmd_ctx *c, const void *in, int count, int row)
{
    if (count == 0) return 0;
    k_halo_append<<<(count + kThreads - 1) / kThreads, kThreads, 0, c->stream>>>(
        count, c->n, c->ps, row, (const HaloRec *)in, c->xv, c->rad, c->cid, c->gid, c->dbox.nx, c->flags,
        c->rad0);
    return 1;
}

int edmd_launch_cell_index(edmd_ctx *c, int mode)
{
    const int n = c->n;
    const int blocks4 = ((n + 3) / 4 + kThreads - 1) / kThreads;
    const int blocks = ((n + 1) / 2 + kThreads - 1) / kThreads;
    int launched = 0;
    if (n > 0) {
        // first kernel of the chain: plain launch (everything before it on the stream completes first)
        edmd_launch(k_count, dim3(blocks4), dim3(kThreads), 0, c->stream, false, n, c->cid, c->cell_cnt, c->rank);
        launched++;
    }
    const bool pdl = c->lean_pdl;
    edmd_launch(k_rowscan, dim3(c->dbox.nl), dim3(kThreads), 0, c->stream, pdl && n > 0, c->dbox.nx, c->ps,
                c->cell_cnt, c->off, c->row_total);
    edmd_launch(k_rowbase, dim3(1), dim3(kBaseThreads), 0, c->stream, pdl, c->dbox.nl, c->row_total, c->row_base);
    edmd_launch(k_chunkmeta, dim3(c->dbox.nl), dim3(kMetaThreads), 0, c->stream, pdl, c->dbox.nx, c->dbox.nl,
                c->ps, c->dbox.ny, c->dbox.yoff, c->slab ? 1 : 0, c->off, c->row_total, c->row_base, c->meta,
                c->cstart, edmd_chunks_bound(c), kCapW, kOffW);
    launched += 3;
    if (n > 0) {
        if (mode == EDMD_MODE_GROW)
            edmd_launch(k_scatter<true>, dim3(blocks), dim3(kThreads), 0, c->stream, pdl, n, c->dbox.nx, c->ps,
                        c->cid, c->rank, c->cstart, c->xv, c->rad, c->vr, c->spos, c->saux, c->svr);
        else
            edmd_launch(k_scatter<false>, dim3(blocks), dim3(kThreads), 0, c->stream, pdl, n, c->dbox.nx, c->ps,
                        c->cid, c->rank, c->cstart, c->xv, c->rad, c->vr, c->spos, c->saux, c->svr);
        launched++;
    }
    c->index_has_vr = (mode == EDMD_MODE_GROW);
    return launched;
}
