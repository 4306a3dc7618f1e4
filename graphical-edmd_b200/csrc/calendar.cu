// calendar.cu -- the ingest plan of a whole sweep for the host's event calendar.
//
// After a full re-predict the reference re-inserts all 2N events one by one
// (addEventToQueue, src/EDMD.c:2144-2170, called from the batch loops at
// :2007-2012, :4772-4779, :4909-4915): for each event a division to find its
// bucket ("Paul list"), then a head insertion into that bucket's intrusive
// list -- two cache misses per event on the host, ~50 ms per tick at N = 10^6,
// forty times the GPU sweep.  The calendar itself stays a sequential host
// structure (BASELINE.json north_star); what moves to the device is the part
// that is a data-parallel function of the 2N event times:
//     bucket[e]   the list index addEventToQueue computes (same FP64 operations:
//                 dt = t_e - paulTime; dt < dtPaul -> -1 = goes to the BST;
//                 dt >= dtPaul*paulListN -> the overflow list paulListN; else
//                 actualPaulList + (int)(dt/dtPaul), wrapped)
//     next/prev   each event's neighbours in its bucket list exactly as 2N head
//                 insertions in the reference's order (crossing 0, collision 0,
//                 crossing 1, ...) would leave them: from the head, descending
//                 insertion sequence
//     head[k]     first event of bucket k
// The host then fills its nodes in ONE streaming pass (no pointer chasing, no
// division) and inserts only the few BST events itself.
//
// Device algorithm: counting sort of the events by bucket (atomics), a scan of
// the bucket counts, and -- because atomics give an arbitrary order inside a
// bucket -- a per-bucket insertion sort by sequence number (buckets hold ~2
// events); the overflow list, which holds every "never" event, gets its order
// from a scan instead.  Integer / data movement only.
#include "edmd_internal.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kScanItems = 4;                         // per thread
constexpr int kScanBlock = kThreads * kScanItems;     // per CTA
constexpr int kMaxBucket = 128;                       // events one thread sorts; more => plan declined

// sequence number s = 2 i + (0 crossing | 1 collision)  <->  event index e = i | N + i
__device__ __forceinline__ int event_of(int s, int n) { return (s & 1) ? n + (s >> 1) : (s >> 1); }

struct PlanArgs {
    int n, paul_n, actual;
    double paul_time, dt_paul;
    const double *t_cross, *t_coll;
    int32_t *bs;        // [2n] bucket by sequence number
    int32_t *ovf;       // [2n] 1 when the event goes to the overflow list
    int32_t *cnt;       // [paul_n + 1]
    int32_t *counters;  // [0] events for the BST, [1] oversized bucket seen
};

__global__ void __launch_bounds__(kThreads)
k_cal_bucket(const __grid_constant__ PlanArgs a)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= 2 * a.n) return;
    const int i = s >> 1;
    const double te = (s & 1) ? a.t_coll[i] : a.t_cross[i];
    // addEventToQueue, src/EDMD.c:2146-2160, operation for operation
    const double dt = __dsub_rn(te, a.paul_time);
    int b;
    if (dt < a.dt_paul) {
        b = -1;
        atomicAdd(&a.counters[0], 1);
    } else if (dt >= __dmul_rn(a.dt_paul, (double)a.paul_n)) {
        b = a.paul_n;
    } else {
        b = a.actual + (int)__ddiv_rn(dt, a.dt_paul);
        if (b >= a.paul_n) b -= a.paul_n;
    }
    a.bs[s] = b;
    a.ovf[s] = b == a.paul_n;
    if (b >= 0 && b < a.paul_n) atomicAdd(&a.cnt[b], 1);
}

// ---- exclusive scan of an int array (three launches, any length) -----------------
__global__ void __launch_bounds__(kThreads)
k_scan_reduce(int n, const int32_t *__restrict__ in, int32_t *__restrict__ sums)
{
    __shared__ int s_w[kThreads / 32];
    const int base = blockIdx.x * kScanBlock;
    int v = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
        const int i = base + k * kThreads + threadIdx.x;
        if (i < n) v += in[i];
    }
    v = __reduce_add_sync(0xffffffffu, v);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < kThreads / 32; w++) t += s_w[w];
        sums[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(1024)
k_scan_sums(int nb, int32_t *__restrict__ sums, int32_t *__restrict__ total)
{
    __shared__ int s_w[32];
    __shared__ int s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < nb; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < nb ? sums[i] : 0;
        int incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_up_sync(0xffffffffu, incl, d);
            if ((threadIdx.x & 31) >= d) incl += o;
        }
        if ((threadIdx.x & 31) == 31) s_w[threadIdx.x >> 5] = incl;
        __syncthreads();
        int wb = 0;
        for (int w = 0; w < (int)(threadIdx.x >> 5); w++) wb += s_w[w];
        const int carry = s_carry;
        if (i < nb) sums[i] = carry + wb + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = carry + wb + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0 && total) *total = s_carry;
}

// each thread scans kScanItems CONSECUTIVE items
__global__ void __launch_bounds__(kThreads)
k_scan_apply(int n, const int32_t *__restrict__ in, const int32_t *__restrict__ sums,
             int32_t *__restrict__ out)
{
    __shared__ int s_w[kThreads / 32];
    const int first = blockIdx.x * kScanBlock + threadIdx.x * kScanItems;
    int v[kScanItems], tsum = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
        v[k] = first + k < n ? in[first + k] : 0;
        tsum += v[k];
    }
    int incl = tsum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, incl, d);
        if ((threadIdx.x & 31) >= d) incl += o;
    }
    if ((threadIdx.x & 31) == 31) s_w[threadIdx.x >> 5] = incl;
    __syncthreads();
    int wb = 0;
    for (int w = 0; w < (int)(threadIdx.x >> 5); w++) wb += s_w[w];
    int run = sums[blockIdx.x] + wb + incl - tsum;
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
        if (first + k < n) out[first + k] = run;
        run += v[k];
    }
}

struct SortArgs {
    int n, paul_n;
    const int32_t *bs, *ovrank;
    const int32_t *start;   // [paul_n + 1] exclusive scan of cnt (entry paul_n = events in regular buckets)
    int32_t *fill;          // [paul_n]
    int32_t *sorted;        // [2n] sequence numbers grouped by bucket; the overflow list follows the buckets
    int32_t *cnt, *counters;
    int32_t *bucket, *next, *prev, *head;   // outputs, by event index / bucket
};

__global__ void __launch_bounds__(kThreads)
k_cal_scatter(const __grid_constant__ SortArgs a)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= 2 * a.n) return;
    const int b = a.bs[s];
    a.bucket[event_of(s, a.n)] = b;
    if (b < 0) return;
    if (b == a.paul_n)
        a.sorted[a.start[a.paul_n] + a.ovrank[s]] = s;   // already in sequence order
    else
        a.sorted[a.start[b] + atomicAdd(&a.fill[b], 1)] = s;
}

// one thread per regular bucket: order its few events by sequence number, link them
__global__ void __launch_bounds__(kThreads)
k_cal_link(const __grid_constant__ SortArgs a)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= a.paul_n) return;
    const int c = a.cnt[k];
    a.cnt[k] = 0;    // ready for the next plan
    a.fill[k] = 0;
    if (c == 0) {
        a.head[k] = -1;
        return;
    }
    if (c > kMaxBucket) {
        a.counters[1] = 1;
        a.head[k] = -1;
        return;
    }
    int v[kMaxBucket];
    const int32_t *src = a.sorted + a.start[k];
    for (int m = 0; m < c; m++) {   // insertion sort, ascending
        const int x = src[m];
        int p = m;
        while (p > 0 && v[p - 1] > x) {
            v[p] = v[p - 1];
            p--;
        }
        v[p] = x;
    }
    // 2N head insertions in sequence order leave the LAST inserted event first
    for (int m = 0; m < c; m++) {
        const int e = event_of(v[m], a.n);
        a.next[e] = m > 0 ? event_of(v[m - 1], a.n) : -1;
        a.prev[e] = m + 1 < c ? event_of(v[m + 1], a.n) : -1;
    }
    a.head[k] = event_of(v[c - 1], a.n);
}

// the overflow list: positions are already in sequence order
__global__ void __launch_bounds__(kThreads)
k_cal_link_overflow(const __grid_constant__ SortArgs a, int first, int count)
{
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m == 0) a.head[a.paul_n] = count > 0 ? event_of(a.sorted[first + count - 1], a.n) : -1;
    if (m >= count) return;
    const int e = event_of(a.sorted[first + m], a.n);
    a.next[e] = m > 0 ? event_of(a.sorted[first + m - 1], a.n) : -1;
    a.prev[e] = m + 1 < count ? event_of(a.sorted[first + m + 1], a.n) : -1;
}

int scan_exclusive(edmd_ctx *c, const int32_t *in, int n, int32_t *out, int32_t *sums, int32_t *total)
{
    const int nb = (n + kScanBlock - 1) / kScanBlock;
    k_scan_reduce<<<nb, kThreads, 0, c->stream>>>(n, in, sums);
    k_scan_sums<<<1, 1024, 0, c->stream>>>(nb, sums, total);
    k_scan_apply<<<nb, kThreads, 0, c->stream>>>(n, in, sums, out);
    return 3;
}

}  // namespace

// Device part of edmd_cuda_calendar_plan.  Scratch layout (int32, allocated by the caller):
//   bs[2n] ovf[2n] ovrank[2n] sorted[2n] | cnt[pn+1] start[pn+1] fill[pn+1] | sums[..] | counters[4]
int edmd_launch_calendar_plan(edmd_ctx *c, double paul_time, double dt_paul, int paul_n, int actual,
                              int32_t *scratch, int32_t *bucket, int32_t *next, int32_t *prev,
                              int32_t *head, int *n_overflow_host_sync)
{
    const int n = c->n_owned;
    const size_t e2 = 2 * (size_t)n, pn1 = (size_t)paul_n + 1;
    int32_t *bs = scratch, *ovf = bs + e2, *ovrank = ovf + e2, *sorted = ovrank + e2;
    int32_t *cnt = sorted + e2, *start = cnt + pn1, *fill = start + pn1;
    const size_t nsum = (e2 > pn1 ? e2 : pn1) / kScanBlock + 2;
    int32_t *sums = fill + pn1, *counters = sums + nsum;
    int launched = 0;
    cudaMemsetAsync(counters, 0, 4 * sizeof(int32_t), c->stream);
    PlanArgs pa;
    pa.n = n; pa.paul_n = paul_n; pa.actual = actual; pa.paul_time = paul_time; pa.dt_paul = dt_paul;
    pa.t_cross = c->t_cross; pa.t_coll = c->t_coll;
    pa.bs = bs; pa.ovf = ovf; pa.cnt = cnt; pa.counters = counters;
    const int eb = (int)((e2 + kThreads - 1) / kThreads);
    k_cal_bucket<<<eb, kThreads, 0, c->stream>>>(pa);
    launched++;
    launched += scan_exclusive(c, cnt, paul_n + 1, start, sums, nullptr);       // cnt[paul_n] == 0 here
    launched += scan_exclusive(c, ovf, (int)e2, ovrank, sums, counters + 2);    // counters[2] = overflow events
    SortArgs sa;
    sa.n = n; sa.paul_n = paul_n; sa.bs = bs; sa.ovrank = ovrank; sa.start = start; sa.fill = fill;
    sa.sorted = sorted; sa.cnt = cnt; sa.counters = counters;
    sa.bucket = bucket; sa.next = next; sa.prev = prev; sa.head = head;
    k_cal_scatter<<<eb, kThreads, 0, c->stream>>>(sa);
    k_cal_link<<<(paul_n + kThreads - 1) / kThreads, kThreads, 0, c->stream>>>(sa);
    launched += 2;
    // the overflow list's extent is needed for the last launch: one small synchronous read
    int32_t h[4] = {0, 0, 0, 0};
    int32_t first = 0;
    cudaMemcpyAsync(h, counters, sizeof(h), cudaMemcpyDeviceToHost, c->stream);
    cudaMemcpyAsync(&first, start + paul_n, sizeof(first), cudaMemcpyDeviceToHost, c->stream);
    cudaStreamSynchronize(c->stream);
    const int nov = h[2];
    k_cal_link_overflow<<<(nov > 0 ? nov + kThreads - 1 : kThreads) / kThreads, kThreads, 0, c->stream>>>(sa, first, nov);
    launched++;
    c->cal_tree = h[0];
    c->cal_declined = h[1] != 0;
    if (n_overflow_host_sync) *n_overflow_host_sync = nov;
    return launched;
}
