// lean_index.cu -- K0-lean: the cell index of the lean sweep (lean.cuh).
//
// Same padded-grid counting sort as cell_index.cu (the reference's linked cell
// list, cellListInit / addToCell, src/EDMD.c:1906-1920, 2053-2078), in three
// launches instead of five and with 32 instead of 48 bytes scattered per
// particle (one full sector, one 256-bit store):
//   count    one RED atomic per particle (no return value, no rank array)
//   index    one CTA per cell row, rows independent (row Y owns slots
//            [Y * rowcap, (Y+1) * rowcap)): exclusive scan of the padded row,
//            absolute cell cursors, the plan of the row's 32-slot chunks, and the
//            row's chunks appended to the sweep's work list (one atomic per row)
//   scatter  slot = atomicAdd(cell cursor); writes the 32-byte record (FP32 screening
//            half + (id, cell) tag); edge cells also feed their ghost column
// Integer / data movement only.
#include "lean.cuh"

namespace {

constexpr int kThreads = 256;

struct CountArgs {
    int n;
    const int32_t *cid;
    int32_t *cnt, *flags;
};

__global__ void __launch_bounds__(kThreads)
k_lean_count(const __grid_constant__ CountArgs a)
{
    const int n = a.n;
    const int32_t *__restrict__ cid = a.cid;
    int32_t *__restrict__ cnt = a.cnt;
    int32_t *__restrict__ flags = a.flags;
    edmd_pdl_wait();   // the previous kernel on the stream may still be reading the histogram's inputs
    const int i = 4 * (blockIdx.x * blockDim.x + threadIdx.x);
    if (i == 0) flags[kFlagWork] = 0;   // the index kernel (next launch) appends to the work list
    if (i + 3 < n) {
        const int4 c = *reinterpret_cast<const int4 *>(cid + i);
        // cid < 0: unused halo slot of a slab context
        if (c.x >= 0) atomicAdd(&cnt[c.x], 1);
        if (c.y >= 0) atomicAdd(&cnt[c.y], 1);
        if (c.z >= 0) atomicAdd(&cnt[c.z], 1);
        if (c.w >= 0) atomicAdd(&cnt[c.w], 1);
    } else {
        for (int k = i; k < n; k++)
            if (cid[k] >= 0) atomicAdd(&cnt[cid[k]], 1);
    }
}

struct IndexArgs {
    int nx, nl, ps, slab, rowcap;
    int row_in_smem;   // the row's offsets fit the dynamic shared memory
    int32_t *cnt, *off, *cstart, *flags, *work;
    LeanChunk *chunks;
};

// smallest column pcx with off[pcx + 1] > slot  (the cell that holds `slot`)
__device__ __forceinline__ int cell_of_slot(const int32_t *o, int ps, int slot)
{
    int lo = 0, hi = ps - 2;  // off[ps-1] = row total > slot (columns past nx+1 are empty)
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (o[mid + 1] > slot) hi = mid;
        else lo = mid + 1;
    }
    return lo;
}

// One CTA per cell row, rows independent: exclusive scan of the padded row
// (ghost cells take the count of the cell they mirror), absolute cell cursors
// (row Y starts at slot Y * rowcap), the plan of the row's chunks; the
// histogram is zeroed again for the next sweep.
__global__ void __launch_bounds__(kThreads)
k_lean_index(const __grid_constant__ IndexArgs a)
{
    extern __shared__ int s_off[];   // the row's offsets (chunk planning searches them)
    __shared__ int s_warp[kThreads / 32];
    __shared__ int s_carry, s_wbase;
    const int tid = threadIdx.x;
    const int Y = blockIdx.x;
    const int nx = a.nx, ps = a.ps;
    edmd_pdl_wait();
    int32_t *row = a.cnt + (size_t)Y * ps;
    int32_t *orow = a.off + (size_t)Y * ps;
    const int rb = Y * a.rowcap;
    if (tid == 0) s_carry = 0;

    // all loads first (one memory round trip for rows up to kPre * 256 columns)
    constexpr int kPre = 8;
    int pre[kPre];
#pragma unroll
    for (int k = 0; k < kPre; k++) {
        const int pcx = k * kThreads + tid;
        int v = 0;
        if (pcx < ps) {
            if (pcx == 0) v = row[nx];            // left ghost mirrors cell nx-1
            else if (pcx <= nx) v = row[pcx];
            else if (pcx == nx + 1) v = row[1];   // right ghost mirrors cell 0
        }
        pre[k] = v;
    }
    __syncthreads();
    auto scan_block = [&](int pcx, int v) {
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
        if (pcx < ps) {
            const int o = carry + wbase + incl - v;
            orow[pcx] = o;
            a.cstart[(size_t)Y * ps + pcx] = rb + o;
            if (a.row_in_smem) s_off[pcx] = o;
        }
        __syncthreads();
        if (tid == kThreads - 1) s_carry = carry + wbase + incl;
        __syncthreads();
    };
#pragma unroll
    for (int k = 0; k < kPre; k++)
        if (k * kThreads < ps) scan_block(k * kThreads + tid, pre[k]);
    for (int base = kPre * kThreads; base < ps; base += kThreads) {   // very wide rows
        const int pcx = base + tid;
        int v = 0;
        if (pcx < ps) {
            if (pcx <= nx) v = row[pcx];
            else if (pcx == nx + 1) v = row[1];
        }
        scan_block(pcx, v);
    }
    const int tot = s_carry;
    // histogram back to zero for the next sweep
    for (int pcx = 1 + tid; pcx <= nx; pcx += kThreads) row[pcx] = 0;
    if (tot > a.rowcap) {   // denser than the lean layout provides for: decline
        if (tid == 0) atomicOr(&a.flags[kFlagLeanFail], 1);
        return;
    }

    // ---- plan of the row's chunks ---------------------------------------------
    const int32_t *srow = a.row_in_smem ? s_off : orow;
    const int nch = (tot + 31) >> 5;
    const bool halo_row = a.slab && (Y == 0 || Y == a.nl - 1);   // never predicted
    // the row's chunks join the work list of the sweep (one atomic per row, any order)
    if (tid == 0) s_wbase = (nch > 0 && !halo_row) ? atomicAdd(&a.flags[kFlagWork], nch) : 0;
    __syncthreads();
    if (!halo_row)
        for (int q = tid; q < nch; q += kThreads) a.work[s_wbase + q] = (rb >> 5) + q;
    for (int q = tid; q < (a.rowcap >> 5); q += kThreads) {
        LeanChunk m = make_int4(-1, 0, 0, 0);
        if (q < nch) {
            const int first = 32 * q, last = min(first + 31, tot - 1);
            const int ca = max(cell_of_slot(srow, ps, first), 1);
            const int cb = min(cell_of_slot(srow, ps, last), nx);
            m = make_int4((ca <= cb && !halo_row) ? Y : -1, ca, cb, rb + tot);
        }
        a.chunks[(rb >> 5) + q] = m;
    }
}

struct ScatterArgs {
    int n, nx, ps;
    edmd_dev_box b;
    const int32_t *cid;
    const double4 *xv;
    int32_t *cursor;
    const int32_t *flags;
    LeanRec *rec;
    const double *rad;   // two radius classes: class bit = (rad[i] != rad0), carried in vy's last mantissa bit
    double rad0;
};

// one full 32-byte sector with a single 256-bit store
__device__ __forceinline__ void put_lean(const ScatterArgs &a, int slot, int pc, int id, const float4 &r)
{
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(a.rec + slot),
                 "r"(__float_as_int(r.x)), "r"(__float_as_int(r.y)), "r"(__float_as_int(r.z)),
                 "r"(__float_as_int(r.w)), "r"(id), "r"(pc), "r"(0), "r"(0)
                 : "memory");
}

// four particles per thread; all loads, then all cursor atomics, then the stores
constexpr int kPer = 1;

__global__ void __launch_bounds__(kThreads)
k_lean_scatter(const __grid_constant__ ScatterArgs a)
{
    const int i0 = kPer * (blockIdx.x * blockDim.x + threadIdx.x);
    edmd_pdl_wait();
    if (i0 >= a.n || a.flags[kFlagLeanFail] != 0) return;   // a row overflowed its slot range: declined
    const bool two = a.flags[kFlagNotMono] == 1;
    int pc[kPer], slot[kPer];
    double4 p[kPer];
    if (kPer == 4 && i0 + kPer <= a.n) {
        const int4 c4 = *reinterpret_cast<const int4 *>(a.cid + i0);
        pc[0] = c4.x; pc[1] = c4.y; pc[kPer - 2] = c4.z; pc[kPer - 1] = c4.w;
    } else if (kPer == 2 && i0 + kPer <= a.n) {
        const int2 c2 = *reinterpret_cast<const int2 *>(a.cid + i0);
        pc[0] = c2.x; pc[1] = c2.y;
    } else {
#pragma unroll
        for (int k = 0; k < kPer; k++) pc[k] = i0 + k < a.n ? a.cid[i0 + k] : -1;
    }
#pragma unroll
    for (int k = 0; k < kPer; k++) p[k] = ld_sector(a.xv + min(i0 + k, a.n - 1));
#pragma unroll
    for (int k = 0; k < kPer; k++) slot[k] = pc[k] >= 0 ? atomicAdd(&a.cursor[pc[k]], 1) : -1;   // < 0: unused halo slot
#pragma unroll
    for (int k = 0; k < kPer; k++) {
        if (pc[k] < 0) continue;
        const int Yl = pc[k] / a.ps;
        const int pcx = pc[k] - Yl * a.ps;
        const int Yg = edmd_global_row(a.b, Yl);
        float4 r;
        r.x = __double2float_rn(__dsub_rn(p[k].x, __dmul_rn((double)(pcx - 1) + 0.5, a.b.csx)));
        r.y = __double2float_rn(__dsub_rn(p[k].y, __dmul_rn((double)Yg + 0.5, a.b.csy)));
        r.z = __double2float_rn(p[k].z);
        r.w = __double2float_rn(p[k].w);
        if (two) r.w = __int_as_float((__float_as_int(r.w) & ~1) | (edmd_same_class(a.rad[i0 + k], a.rad0) ? 0 : 1));
        put_lean(a, slot[k], pc[k], i0 + k, r);
        if (pcx == 1) {   // cell 0 -> right ghost
            const int g = Yl * a.ps + a.nx + 1;
            put_lean(a, atomicAdd(&a.cursor[g], 1), g, i0 + k, r);
        }
        if (pcx == a.nx) {   // cell nx-1 -> left ghost
            const int g = Yl * a.ps;
            put_lean(a, atomicAdd(&a.cursor[g], 1), g, i0 + k, r);
        }
    }
}

}  // namespace

int edmd_launch_lean_index(edmd_ctx *c)
{
    const int n = c->n;
    if (n == 0) return 0;
    CountArgs ca;
    ca.n = n; ca.cid = c->cid; ca.cnt = c->cell_cnt; ca.flags = c->flags;
    // the first kernel of the chain is launched plainly: whatever precedes it on the stream completes first
    edmd_launch(k_lean_count, dim3(((n + 3) / 4 + kThreads - 1) / kThreads), dim3(kThreads), 0, c->stream, false, ca);
    IndexArgs ia;
    ia.nx = c->dbox.nx; ia.nl = c->dbox.nl; ia.ps = c->ps; ia.slab = c->slab ? 1 : 0;
    ia.rowcap = c->rowcap;
    ia.cnt = c->cell_cnt; ia.off = c->off; ia.cstart = c->cstart; ia.flags = c->flags;
    ia.chunks = c->lchunks;
    ia.work = c->lwork;
    ia.row_in_smem = (size_t)c->ps * sizeof(int) <= 40960 ? 1 : 0;
    edmd_launch(k_lean_index, dim3(c->dbox.nl), dim3(kThreads), ia.row_in_smem ? (size_t)c->ps * sizeof(int) : 0,
                c->stream, c->lean_pdl, ia);
    ScatterArgs sa;
    sa.n = n; sa.nx = c->dbox.nx; sa.ps = c->ps; sa.b = c->dbox;
    sa.cid = c->cid; sa.xv = c->xv; sa.cursor = c->cstart; sa.flags = c->flags; sa.rec = c->lrec;
    sa.rad = c->rad; sa.rad0 = c->rad0;
    edmd_launch(k_lean_scatter, dim3(((n + kPer - 1) / kPer + kThreads - 1) / kThreads), dim3(kThreads), 0, c->stream,
                c->lean_pdl, sa);
    c->index_has_vr = false;
    return 3;
}
