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
// k_halo_send packs a boundary row of the owned particles and writes the
// 48-byte records STRAIGHT INTO THE NEIGHBOUR'S inbox (peer stores over
// NVLink); the last block publishes count, then epoch, behind system fences.
// k_halo_recv waits for the epoch, unpacks the records into the fixed halo
// region [n_owned + from*H, n_owned + (from+1)*H) of the resident arrays (unused
// slots get cell id -1 and are skipped by the cell index), and acks.  A sender
// may run at most two exchanges ahead of its receiver (ack check).
#include "edmd_internal.cuh"

namespace {

constexpr int kThreads = 256;

struct __align__(16) HaloRec {
    double x, y, vx, vy, rad;
    int gid;
    int cell;   // padded column 1..nx (the row is implied by which neighbour sent it)
};
static_assert(sizeof(HaloRec) == 48, "halo record");

struct __align__(16) InboxHeader {
    int count, epoch, pad[2];
};

__host__ __device__ inline size_t inbox_bytes(int H)
{
    return (sizeof(InboxHeader) + sizeof(HaloRec) * (size_t)H + 255) & ~(size_t)255;
}
__host__ __device__ inline size_t inbox_offset(int H, int from, int parity)
{
    return inbox_bytes(H) * (size_t)(2 * from + parity);
}
__host__ __device__ inline size_t ack_offset(int H) { return inbox_bytes(H) * 4; }

__device__ __forceinline__ int ld_volatile(const int *p)
{
    return *reinterpret_cast<const volatile int *>(p);
}

// side 0: first owned row -> lower neighbour's inbox[from=1]; side 1: last owned row -> upper's inbox[from=0]
__global__ void __launch_bounds__(kThreads)
k_halo_send(int n_owned, int ps, int row, int H, int epoch, const int32_t *__restrict__ cid,
            const double4 *__restrict__ xv, const double *__restrict__ rad,
            const int32_t *__restrict__ gid, char *peer_inbox, const int *ack, int32_t *cnt,
            int32_t *done, int32_t *flags)
{
    __shared__ bool last;
    // do not overwrite a buffer the neighbour may still be reading
    if (threadIdx.x == 0)
        while (ld_volatile(ack) < epoch - 2) __nanosleep(50);
    __syncthreads();
    InboxHeader *hdr = reinterpret_cast<InboxHeader *>(peer_inbox);
    HaloRec *rec = reinterpret_cast<HaloRec *>(peer_inbox + sizeof(InboxHeader));
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_owned) {
        const int pc = cid[i];
        if (pc / ps == row) {
            const int k = atomicAdd(cnt, 1);
            if (k < H) {
                const double4 p = xv[i];
                HaloRec r;
                r.x = p.x; r.y = p.y; r.vx = p.z; r.vy = p.w;
                r.rad = rad[i];
                r.gid = gid[i];
                r.cell = pc - row * ps;
                const uint4 *src = reinterpret_cast<const uint4 *>(&r);
                uint4 *dst = reinterpret_cast<uint4 *>(rec + k);
                dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2];
            } else
                atomicOr(&flags[kFlagBadCell], 2);   // halo buffer too small
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(done, 1) == (int)gridDim.x - 1;
    __syncthreads();
    if (last && threadIdx.x == 0) {
        const int total = min(ld_volatile(cnt), H);
        *reinterpret_cast<volatile int *>(&hdr->count) = total;
        __threadfence_system();
        *reinterpret_cast<volatile int *>(&hdr->epoch) = epoch;
        __threadfence_system();
        *cnt = 0;
        *done = 0;
    }
}

__global__ void __launch_bounds__(kThreads)
k_halo_recv(int H, int first, int ps, int row, int nx, int epoch, const char *inbox,
            double4 *__restrict__ xv, double *__restrict__ rad, int32_t *__restrict__ cid,
            int32_t *__restrict__ gid, int *peer_ack, int32_t *done)
{
    __shared__ bool last;
    const InboxHeader *hdr = reinterpret_cast<const InboxHeader *>(inbox);
    const HaloRec *rec = reinterpret_cast<const HaloRec *>(inbox + sizeof(InboxHeader));
    if (threadIdx.x == 0)
        while (ld_volatile(&hdr->epoch) != epoch) __nanosleep(50);
    __syncthreads();
    const int count = ld_volatile(&hdr->count);
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < H) {
        const int i = first + k;
        if (k < count) {
            const uint4 *src = reinterpret_cast<const uint4 *>(rec + k);
            HaloRec r;
            uint4 *dst = reinterpret_cast<uint4 *>(&r);
            dst[0] = __ldcv(src); dst[1] = __ldcv(src + 1); dst[2] = __ldcv(src + 2);
            xv[i] = make_double4(r.x, r.y, r.vx, r.vy);
            rad[i] = r.rad;
            gid[i] = r.gid;
            cid[i] = row * ps + r.cell;
        } else {
            cid[i] = -1;
            gid[i] = -1;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(done, 1) == (int)gridDim.x - 1;
    __syncthreads();
    if (last && threadIdx.x == 0) {
        __threadfence_system();
        *reinterpret_cast<volatile int *>(peer_ack) = epoch;
        __threadfence_system();
        *done = 0;
    }
    (void)nx;
}

}  // namespace

size_t edmd_halo_mem_bytes(int halo_cap) { return ack_offset(halo_cap) + 256; }

// Launches send(both sides) then recv(both sides) on the context's stream.
int edmd_launch_halo_p2p(edmd_ctx *c)
{
    const int H = c->halo_cap;
    const int e = ++c->halo_epoch;
    const int par = e & 1;
    const int n = c->n_owned;
    const int nl = c->dbox.nl;
    int32_t *cnt = c->halo_cnt;
    int *my_ack = reinterpret_cast<int *>(c->halo_mem + ack_offset(H));
    const int sblocks = n > 0 ? (n + kThreads - 1) / kThreads : 1;
    // my first owned row goes to the LOWER neighbour, where I am its upper one (from = 1)
    k_halo_send<<<sblocks, kThreads, 0, c->stream>>>(n, c->ps, 1, H, e, c->cid, c->xv, c->rad, c->gid,
                                                     c->peer_mem[0] + inbox_offset(H, 1, par), my_ack + 0,
                                                     cnt + 0, cnt + 2, c->flags);
    // my last owned row goes to the UPPER neighbour, where I am its lower one (from = 0)
    k_halo_send<<<sblocks, kThreads, 0, c->stream>>>(n, c->ps, nl - 2, H, e, c->cid, c->xv, c->rad, c->gid,
                                                     c->peer_mem[1] + inbox_offset(H, 0, par), my_ack + 1,
                                                     cnt + 1, cnt + 3, c->flags);
    const int rblocks = (H + kThreads - 1) / kThreads;
    // from the lower neighbour: its last row = my local row 0; I am its UPPER neighbour -> its ack[1]
    int *ack_lower = reinterpret_cast<int *>(c->peer_mem[0] + ack_offset(H)) + 1;
    int *ack_upper = reinterpret_cast<int *>(c->peer_mem[1] + ack_offset(H)) + 0;
    k_halo_recv<<<rblocks, kThreads, 0, c->stream>>>(H, n, c->ps, 0, c->dbox.nx, e,
                                                     c->halo_mem + inbox_offset(H, 0, par), c->xv, c->rad,
                                                     c->cid, c->gid, ack_lower, cnt + 4);
    k_halo_recv<<<rblocks, kThreads, 0, c->stream>>>(H, n + H, c->ps, nl - 1, c->dbox.nx, e,
                                                     c->halo_mem + inbox_offset(H, 1, par), c->xv, c->rad,
                                                     c->cid, c->gid, ack_upper, cnt + 5);
    return 4;
}
