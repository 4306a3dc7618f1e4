// lean_index.cu -- K0-lean: the cell index of the lean sweep (lean.cuh).
//
// Same padded-grid counting sort as cell_index.cu (the reference's linked cell
// list, cellListInit / addToCell, src/EDMD.c:1906-1920, 2053-2078), in three
// launches instead of five and with 32 instead of 48 bytes scattered per
// particle (one full sector, one 256-bit store):
//   count    one RED atomic per particle (no return value, no rank array)
//   index    one CTA per cell row, rows handed out by a ticket: exclusive scan
//            of the padded row, row total published in row_state[] (tagged
//            with the sweep's epoch), row base = sum of the totals of all
//            earlier rows (each waited for by its tag: earlier tickets are
//            running or done, so this cannot deadlock), absolute cell cursors,
//            and the plan of the row's 32-slot chunks
//   scatter  slot = atomicAdd(cell cursor); writes the 32-byte record (FP32 screening
//            half + (id, cell) tag) and slot_of[id]; edge cells also feed their ghost column
// Integer / data movement only.
#include "lean.cuh"

namespace {

constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads)
k_lean_count(int n, const int32_t *__restrict__ cid, int32_t *__restrict__ cnt)
{
    const int i = 4 * (blockIdx.x * blockDim.x + threadIdx.x);
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
    int nx, nl, ps, slab, max_chunks;
    int row_in_smem;   // the row's offsets fit the dynamic shared memory
    uint32_t epoch;
    int32_t *cnt, *off, *cstart, *row_base, *ticket;
    LeanChunk *chunks;
    unsigned long long *row_state;
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

__device__ __forceinline__ int block_sum(int v, int *s_warp)
{
    v = __reduce_add_sync(0xffffffffu, v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = v;
    __syncthreads();
    int t = 0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; w++) t += s_warp[w];
    return t;
}

__global__ void __launch_bounds__(kThreads)
k_lean_index(const __grid_constant__ IndexArgs a)
{
    extern __shared__ int s_off[];   // the row's offsets (chunk planning searches them)
    __shared__ int s_warp[kThreads / 32];
    __shared__ int s_carry, s_row;
    const int tid = threadIdx.x;
    if (tid == 0) {
        s_row = atomicAdd(a.ticket, 1);
        s_carry = 0;
    }
    __syncthreads();
    const int Y = s_row;
    const int nx = a.nx, ps = a.ps;
    int32_t *row = a.cnt + (size_t)Y * ps;
    int32_t *orow = a.off + (size_t)Y * ps;

    // ---- exclusive scan of the padded row ------------------------------------
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
        if (pcx < ps) {
            orow[pcx] = carry + wbase + incl - v;
            if (a.row_in_smem) s_off[pcx] = carry + wbase + incl - v;
        }
        __syncthreads();
        if (tid == kThreads - 1) s_carry = carry + wbase + incl;
        __syncthreads();
    }
    const int tot = s_carry;
    const int tot32 = (tot + 31) & ~31;
    if (tid == 0)
        *reinterpret_cast<volatile unsigned long long *>(a.row_state + Y) =
            ((unsigned long long)a.epoch << 32) | (unsigned int)tot32;
    // histogram back to zero for the next sweep
    for (int pcx = 1 + tid; pcx <= nx; pcx += kThreads) row[pcx] = 0;

    // ---- row base: totals of all earlier rows (earlier tickets: running or done) --
    int part = 0;
    for (int r = tid; r < Y; r += kThreads) {
        const volatile unsigned long long *p = a.row_state + r;
        unsigned long long st = *p;
        while ((uint32_t)(st >> 32) != a.epoch) {
            __nanosleep(20);
            st = *p;
        }
        part += (int)(uint32_t)st;
    }
    const int rb = block_sum(part, s_warp);   // also orders the off[] writes before the reads below
    if (tid == 0) {
        a.row_base[Y] = rb;
        if (Y == a.nl - 1) {
            a.row_base[a.nl] = rb + tot32;
            *a.ticket = 0;   // every row has taken its ticket by now
        }
    }
    const int32_t *srow = a.row_in_smem ? s_off : orow;
    for (int k = tid; k < ps; k += kThreads) a.cstart[(size_t)Y * ps + k] = rb + srow[k];

    // ---- plan of the row's chunks ---------------------------------------------
    const int nch = tot32 >> 5;
    const bool halo_row = a.slab && (Y == 0 || Y == a.nl - 1);   // never predicted
    for (int q = tid; q < nch; q += kThreads) {
        const int first = 32 * q, last = min(first + 31, tot - 1);
        const int ca = max(cell_of_slot(srow, ps, first), 1);
        const int cb = min(cell_of_slot(srow, ps, last), nx);
        const int ch = (rb >> 5) + q;
        if (ch < a.max_chunks)
            a.chunks[ch] = make_int4((ca <= cb && !halo_row) ? Y : -1, ca, cb, rb + tot);
    }
    if (Y == a.nl - 1)   // chunks past the last row are empty
        for (int ch = ((rb + tot32) >> 5) + tid; ch < a.max_chunks; ch += kThreads)
            a.chunks[ch] = make_int4(-1, 0, 0, 0);
}

struct ScatterArgs {
    int n, nx, ps;
    edmd_dev_box b;
    const int32_t *cid;
    const double4 *xv;
    int32_t *cursor;
    LeanRec *rec;
    int32_t *slot_of;   // slot of particle i's own (non-ghost) entry
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
constexpr int kPer = 4;

__global__ void __launch_bounds__(kThreads)
k_lean_scatter(const __grid_constant__ ScatterArgs a)
{
    const int i0 = kPer * (blockIdx.x * blockDim.x + threadIdx.x);
    if (i0 >= a.n) return;
    int pc[kPer], slot[kPer];
    double4 p[kPer];
    if (i0 + kPer <= a.n) {
        const int4 c4 = *reinterpret_cast<const int4 *>(a.cid + i0);
        pc[0] = c4.x; pc[1] = c4.y; pc[2] = c4.z; pc[3] = c4.w;
    } else {
#pragma unroll
        for (int k = 0; k < kPer; k++) pc[k] = i0 + k < a.n ? a.cid[i0 + k] : -1;
    }
#pragma unroll
    for (int k = 0; k < kPer; k++) p[k] = a.xv[min(i0 + k, a.n - 1)];
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
        put_lean(a, slot[k], pc[k], i0 + k, r);
        a.slot_of[i0 + k] = slot[k];
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
    k_lean_count<<<((n + 3) / 4 + kThreads - 1) / kThreads, kThreads, 0, c->stream>>>(n, c->cid, c->cell_cnt);
    IndexArgs ia;
    ia.nx = c->dbox.nx; ia.nl = c->dbox.nl; ia.ps = c->ps; ia.slab = c->slab ? 1 : 0;
    ia.max_chunks = edmd_chunks_bound(c);
    ia.epoch = ++c->index_epoch;
    ia.cnt = c->cell_cnt; ia.off = c->off; ia.cstart = c->cstart; ia.row_base = c->row_base;
    ia.ticket = c->flags + kFlagTicket;
    ia.chunks = c->lchunks;
    ia.row_state = c->row_state;
    ia.row_in_smem = (size_t)c->ps * sizeof(int) <= 40960 ? 1 : 0;
    k_lean_index<<<c->dbox.nl, kThreads, ia.row_in_smem ? (size_t)c->ps * sizeof(int) : 0, c->stream>>>(ia);
    ScatterArgs sa;
    sa.n = n; sa.nx = c->dbox.nx; sa.ps = c->ps; sa.b = c->dbox;
    sa.cid = c->cid; sa.xv = c->xv; sa.cursor = c->cstart; sa.rec = c->lrec; sa.slot_of = c->rank;
    k_lean_scatter<<<((n + kPer - 1) / kPer + kThreads - 1) / kThreads, kThreads, 0, c->stream>>>(sa);
    c->index_has_vr = false;
    return 3;
}
