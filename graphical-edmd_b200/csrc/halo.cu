// halo.cu -- one-cell-row halo exchange between neighbouring slabs, done by
// the GPUs themselves over NVLink peer memory: no NCCL call, no host
// synchronisation, everything ordered on the context's stream.
//
// Every rank owns one allocation `halo_mem` (exported to its two neighbours
// through CUDA IPC):
//     inbox[from][parity]   from = 0: records sent by the LOWER neighbour (its last
//                           owned row), 1: by the UPPER neighbour (its first row);
//                           two parities so consecutive exchanges do not collide
//         { int count; int epoch; int pad[2]; HaloRec rec[H]; }
//     ack[2]                ack[0] written by my lower neighbour, ack[1] by my upper
//                           one: "I have consumed your exchange number e"
// k_halo_collect lists the owned particles of the two boundary rows; k_halo_send
// writes their 48-byte records STRAIGHT INTO THE NEIGHBOURS' inboxes (peer
// stores over NVLink); its last block publishes count, then epoch, behind
// system fences.
// k_halo_recv waits for the epoch, unpacks the records into the fixed halo
// region [n_owned + from*H, n_owned + (from+1)*H) of the resident arrays (unused
// slots get cell id -1 and are skipped by the cell index), and acks.  A sender
// may run at most two exchanges ahead of its receiver (ack check).
#include "edmd_internal.cuh"
#include "halo.cuh"

namespace {

constexpr int kThreads = 256;

// Sending is two steps.  The indices of the owned particles in the two boundary
// rows are listed once per upload (by the pack kernel, or by k_halo_collect if the
// inboxes were exported after the upload; device-scope atomics only).  k_halo_send, a few blocks, builds their records and writes them
// straight into the neighbours' inboxes; only these few threads pay for
// system-scope fences.  (One kernel doing both cost 55 us at 10^6 owned
// particles -- a system fence in each of its 3900 blocks -- against 9 us.)
//   side 0: first owned row -> lower neighbour's inbox[from=1];
//   side 1: last owned row  -> upper neighbour's inbox[from=0].
struct CollectArgs {
    int n_owned, ps, row[2], H;
    const int32_t *cid;
    int32_t *list;      // [2][H] particle indices per side
    int32_t *cnt;       // [2]
    int32_t *flags;
};

__global__ void __launch_bounds__(kThreads)
k_halo_collect(const __grid_constant__ CollectArgs a)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n_owned) return;
    const int l = a.cid[i] / a.ps;
#pragma unroll
    for (int side = 0; side < 2; side++) {
        if (l != a.row[side]) continue;
        const int k = atomicAdd(a.cnt + side, 1);
        if (k < a.H) a.list[side * a.H + k] = i;
        else atomicOr(&a.flags[kFlagBadCell], 2);   // halo buffer too small
    }
}

struct SendArgs {
    int ps, H, epoch;
    const int32_t *cid;
    const double4 *xv;
    const double *rad;
    const int32_t *gid;
    const int32_t *list;
    char *peer_inbox[2];
    const int *ack;     // [2]
    int32_t *cnt;       // [2] records listed per side
    int32_t *done;      // finished blocks
    unsigned long long *ts;
};

// blockIdx.y = side
__global__ void __launch_bounds__(kThreads)
k_halo_send(const __grid_constant__ SendArgs a)
{
    __shared__ bool last;
    const int side = blockIdx.y;
    // In the fused exchange + sweep chain (edmd_cuda_exchange_predict_device) this kernel comes FIRST and
    // lets the rest of the chain (receive, partition: launched with the programmatic attribute) start at
    // once: a handful of small blocks here, mostly waiting on NVLink, beside the partition's thousands.
    edmd_pdl_trigger();
    if (blockIdx.x == 0 && side == 0 && threadIdx.x == 0) edmd_stamp(a.ts, 0);
    // do not overwrite a buffer the neighbour may still be reading
    if (threadIdx.x == 0)
        while (ld_volatile(a.ack + side) < a.epoch - 2) __nanosleep(50);
    __syncthreads();
    const int total = min(a.cnt[side], a.H);
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < total) {
        const int i = a.list[side * a.H + k];
        const int pc = a.cid[i];
        const double4 p = a.xv[i];
        HaloRec r;
        r.x = p.x; r.y = p.y; r.vx = p.z; r.vy = p.w;
        r.rad = a.rad[i];
        r.gid = a.gid[i];
        r.cell = pc - (pc / a.ps) * a.ps;
        HaloRec *rec = reinterpret_cast<HaloRec *>(a.peer_inbox[side] + sizeof(InboxHeader));
        const uint4 *src = reinterpret_cast<const uint4 *>(&r);
        uint4 *dst = reinterpret_cast<uint4 *>(rec + k);   // peer store over NVLink
        dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        // one system-scope fence per block: it is cumulative over the stores the barrier
        // ordered before it, so they are visible to the neighbour before `done` counts
        // this block (and hence before count / epoch are published)
        __threadfence_system();
        last = atomicAdd(a.done, 1) == (int)(gridDim.x * gridDim.y) - 1;
    }
    __syncthreads();
    if (last && threadIdx.x < 2) {
        const int sd = threadIdx.x;
        InboxHeader *hdr = reinterpret_cast<InboxHeader *>(a.peer_inbox[sd]);
        const int tot = min(ld_volatile(a.cnt + sd), a.H);
        *reinterpret_cast<volatile int *>(&hdr->count) = tot;
        __threadfence_system();
        *reinterpret_cast<volatile int *>(&hdr->epoch) = a.epoch;
        __threadfence_system();
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        *a.done = 0;   // cnt / list stay valid until the next upload
        edmd_stamp(a.ts, 1);
    }
}

struct RecvArgs {
    int H, first, ps, row[2], epoch;
    const char *inbox[2];
    double4 *xv;
    double *rad;
    int32_t *cid, *gid;
    int *peer_ack[2];
    int32_t *done;
    int32_t *flags;
    double rad0;
};

// blockIdx.y = from (0: lower neighbour's records -> local row 0, 1: upper -> row nl-1)
__global__ void __launch_bounds__(kThreads)
k_halo_recv(const __grid_constant__ RecvArgs a)
{
    __shared__ bool last;
    const int from = blockIdx.y;
    const InboxHeader *hdr = reinterpret_cast<const InboxHeader *>(a.inbox[from]);
    const HaloRec *rec = reinterpret_cast<const HaloRec *>(a.inbox[from] + sizeof(InboxHeader));
    if (threadIdx.x == 0)
        while (ld_volatile(&hdr->epoch) != a.epoch) __nanosleep(50);
    __syncthreads();
    const int count = ld_volatile(&hdr->count);
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < a.H) {
        const int i = a.first + from * a.H + k;
        if (k < count) {
            const uint4 *src = reinterpret_cast<const uint4 *>(rec + k);
            HaloRec r;
            uint4 *dst = reinterpret_cast<uint4 *>(&r);
            dst[0] = __ldcv(src); dst[1] = __ldcv(src + 1); dst[2] = __ldcv(src + 2);
            a.xv[i] = make_double4(r.x, r.y, r.vx, r.vy);
            a.rad[i] = r.rad;
            a.gid[i] = r.gid;
            a.cid[i] = a.row[from] * a.ps + r.cell;
            // keep the lean sweep's eligibility facts current (lean.cuh)
            edmd_note_radius(a.flags, r.rad, a.rad0);
            float vm = __double2float_ru(fmax(fabs(r.vx), fabs(r.vy)));
            if (!(vm == vm)) vm = __int_as_float(0x7f800000);
            if (__float_as_int(vm) > a.flags[kFlagVmax])
                atomicMax(reinterpret_cast<unsigned int *>(&a.flags[kFlagVmax]), (unsigned)__float_as_int(vm));
        } else {
            a.cid[i] = -1;   // unused slot: skipped by the cell index
            a.gid[i] = -1;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(a.done + from, 1) == (int)gridDim.x - 1;
    __syncthreads();
    if (last && threadIdx.x == 0) {
        __threadfence_system();
        *reinterpret_cast<volatile int *>(a.peer_ack[from]) = a.epoch;   // "consumed", peer store
        __threadfence_system();
        a.done[from] = 0;
    }
}

}  // namespace

size_t edmd_halo_mem_bytes(int halo_cap) { return ack_offset(halo_cap) + 256; }

// The exchange is two launches on the context's stream: send (peer stores into the neighbours'
// inboxes; starts a new epoch) and receive (waits for the neighbours' epoch, unpacks, acks).  Work that
// does not need the halo may be launched between the two (edmd_cuda_exchange_predict_device).
// The kernels of an exchange WAIT FOR EACH OTHER across streams / devices (receive spins on the neighbour's
// send).  With CUDA's lazy module loading the first launch of a kernel loads it, and that load may have to
// wait for the device to go idle -- which a spinning receive kernel of the same process never lets happen
// (a host thread that drives several slabs hung in its first exchange, measured).  So every kernel of the
// exchange chain is loaded BEFORE the first exchange: cudaFuncGetAttributes forces the load.
void edmd_preload_sweep_kernels();   // cell_sweep.cu
void edmd_preload_exchange_kernels()
{
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, k_halo_collect);
    cudaFuncGetAttributes(&fa, k_halo_send);
    cudaFuncGetAttributes(&fa, k_halo_recv);
    edmd_preload_sweep_kernels();
    cudaGetLastError();
}

int edmd_launch_halo_send(edmd_ctx *c, bool chained)
{
    cudaStream_t st = c->stream;
    const int H = c->halo_cap;
    const int e = ++c->halo_epoch;
    const int par = e & 1;
    const int n = c->n_owned;
    const int nl = c->dbox.nl;
    int32_t *cnt = c->halo_cnt;
    CollectArgs ca;
    ca.n_owned = n; ca.ps = c->ps; ca.H = H;
    ca.row[0] = 1;          // my first owned row -> LOWER neighbour, where I am its upper one (from = 1)
    ca.row[1] = nl - 2;     // my last owned row  -> UPPER neighbour, where I am its lower one (from = 0)
    ca.cid = c->cid; ca.list = c->halo_list; ca.cnt = cnt; ca.flags = c->flags;
    const bool collect = n > 0 && !c->halo_list_valid;   // else the upload's pack kernel listed them
    if (collect) {
        cudaMemsetAsync(cnt, 0, 2 * sizeof(int32_t), st);
        k_halo_collect<<<(n + kThreads - 1) / kThreads, kThreads, 0, st>>>(ca);
        c->halo_list_valid = true;
    }
    SendArgs sa;
    sa.ps = c->ps; sa.H = H; sa.epoch = e;
    sa.cid = c->cid; sa.xv = c->xv; sa.rad = c->rad; sa.gid = c->gid; sa.list = c->halo_list;
    sa.peer_inbox[0] = c->peer_mem[0] + inbox_offset(H, 1, par);
    sa.peer_inbox[1] = c->peer_mem[1] + inbox_offset(H, 0, par);
    sa.ack = reinterpret_cast<const int *>(c->halo_mem + ack_offset(H));
    sa.cnt = cnt; sa.done = cnt + 2;
    sa.ts = c->tile_dbg & 32 ? c->dbg_ts : nullptr;
    (void)chained;   // always the first kernel of its chain: plain launch
    edmd_launch(k_halo_send, dim3((H + kThreads - 1) / kThreads, 2), dim3(kThreads), 0, st, false, sa);
    return collect ? 2 : 1;
}

int edmd_launch_halo_recv(edmd_ctx *c)
{
    const int H = c->halo_cap;
    const int e = c->halo_epoch, par = e & 1;
    const int nl = c->dbox.nl;
    int32_t *cnt = c->halo_cnt;
    RecvArgs ra;
    ra.H = H; ra.first = c->n_owned; ra.ps = c->ps; ra.epoch = e;
    ra.row[0] = 0; ra.row[1] = nl - 1;
    ra.inbox[0] = c->halo_mem + inbox_offset(H, 0, par);
    ra.inbox[1] = c->halo_mem + inbox_offset(H, 1, par);
    ra.xv = c->xv; ra.rad = c->rad; ra.cid = c->cid; ra.gid = c->gid;
    // consuming the lower neighbour's records: I am its UPPER neighbour -> its ack[1]; and vice versa
    ra.peer_ack[0] = reinterpret_cast<int *>(c->peer_mem[0] + ack_offset(H)) + 1;
    ra.peer_ack[1] = reinterpret_cast<int *>(c->peer_mem[1] + ack_offset(H)) + 0;
    ra.done = cnt + 4;
    ra.flags = c->flags;
    ra.rad0 = c->rad0;
    dim3 rgrid((H + kThreads - 1) / kThreads, 2);
    k_halo_recv<<<rgrid, kThreads, 0, c->stream>>>(ra);
    return 1;
}

int edmd_launch_halo_p2p(edmd_ctx *c)
{
    int launched = edmd_launch_halo_send(c, false);
    return launched + edmd_launch_halo_recv(c);
}
