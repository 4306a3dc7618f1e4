// analysis_pcf_sorted.cu -- K3s: full-range g(r) (calculate_pcf, src/pcf.c:16-75)
// over spatially sorted tiles, with the bin of a pair CERTIFIED instead of
// computed by an IEEE square root and division.
//
// The reference bins  bin = (int)(sqrt(dx^2 + dy^2) / dr)  for all N(N-1)/2 pairs
// after the minimum-image adjustment (PBC, src/EDMD.c:5896-5913).  The integer
// counts must be the reference's.  Per pair this kernel computes only
//     s = fl(fl(dx*dx) + fl(dy*dy))          exactly as the reference (unfused FP64)
// and then
//   * range test in s-space: r < max_r  <=>  s < S_max, with S_max the smallest
//     double whose correctly rounded square root reaches max_r (found on the host);
//   * an FP32 estimate k of the bin (float taken from the bits of s, MUFU rsqrt);
//   * the certificate  (k dr)^2 (1 + 2^-48) <= s < ((k+1) dr)^2 (1 - 2^-48)  in FP64
//     (5 multiplications, 2 additions, 2 compares): the two roundings of
//     sqrt-then-divide move r/dr by at most 2^-52 relative, so inside that
//     interval the reference's truncation gives k.  A pair that fails (the FP32
//     estimate was off by one -- ~0.1-0.5 % of the pairs -- or s sits within
//     2^-48 of an edge) takes the reference's sqrt and division.
// 14 FP64 operations per pair instead of ~35.
//
// Tiles are 256 consecutive particles of an array sorted by coarse cell
// (row-major), with exact bounding boxes.  For a tile pair the boxes tell whether
// any pair can need the periodic image in x / in y (else the adjustment is
// skipped: identical result) and whether every pair is farther than max_r (the
// tile pair is skipped: for max_r = L/2 that is ~20 % of them).
#include "edmd_internal.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kTile = 256;

__device__ __forceinline__ double min_image(double d, double half, double len)
{
    if (d >= half) return __dsub_rn(d, len);
    if (d < -half) return __dadd_rn(d, len);
    return d;
}

struct SortArgs {
    int n, stride, gx, gy;
    double fx, fy;       // coarse cells per unit length
    const double *xy;
    int32_t *cnt;        // [gx * gy + 1]
    int32_t *cell;       // [n]
    double2 *sorted;
    int32_t *sidx;       // [n] original index of sorted[k] (nullptr: not needed)
};

__device__ __forceinline__ int coarse_cell(const SortArgs &a, double x, double y)
{
    int cx = (int)(x * a.fx), cy = (int)(y * a.fy);
    cx = min(max(cx, 0), a.gx - 1);
    cy = min(max(cy, 0), a.gy - 1);
    return cy * a.gx + cx;
}

__global__ void __launch_bounds__(kThreads)
k_coarse_count(const __grid_constant__ SortArgs a)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const double2 p = *reinterpret_cast<const double2 *>(a.xy + (size_t)i * a.stride);
    const int c = coarse_cell(a, p.x, p.y);
    a.cell[i] = c;
    atomicAdd(&a.cnt[c], 1);
}

// exclusive scan of a short array in place, one block
__global__ void __launch_bounds__(1024)
k_small_scan(int m, int32_t *__restrict__ v)
{
    __shared__ int s_w[32];
    __shared__ int s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < m; base += 1024) {
        const int i = base + threadIdx.x;
        const int x = i < m ? v[i] : 0;
        int incl = x;
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
        if (i < m) v[i] = carry + wb + incl - x;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = carry + wb + incl;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kThreads)
k_coarse_scatter(const __grid_constant__ SortArgs a)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const double2 p = *reinterpret_cast<const double2 *>(a.xy + (size_t)i * a.stride);
    const int k = atomicAdd(&a.cnt[a.cell[i]], 1);   // cnt holds the running cursors now
    a.sorted[k] = p;
    if (a.sidx) a.sidx[k] = i;
}

// Multi-GPU split: every rank must cut the SAME tiles, so the arrival order inside a
// coarse cell (atomics) is replaced by the order of the original indices.  One CTA
// per coarse cell; rank of an entry = number of entries of the cell with a smaller index.
__global__ void __launch_bounds__(kThreads)
k_coarse_order(int ncoarse, const int32_t *__restrict__ cend, const double2 *__restrict__ in,
               const int32_t *__restrict__ idx, double2 *__restrict__ out)
{
    __shared__ int s_idx[1024];
    const int c = blockIdx.x;
    const int lo = c == 0 ? 0 : cend[c - 1], hi = cend[c];   // cursors ended at the cell ends
    const int m = hi - lo;
    const bool cached = m <= 1024;
    if (cached)
        for (int k = threadIdx.x; k < m; k += kThreads) s_idx[k] = idx[lo + k];
    __syncthreads();
    for (int k = threadIdx.x; k < m; k += kThreads) {
        const int mine = idx[lo + k];
        int rank = 0;
        for (int q = 0; q < m; q++) rank += (cached ? s_idx[q] : idx[lo + q]) < mine;
        out[lo + rank] = in[lo + k];
    }
}

// exact bounding box of each tile
__global__ void __launch_bounds__(kTile)
k_tile_bbox(int n, const double2 *__restrict__ sorted, double4 *__restrict__ bbox)
{
    __shared__ double s[4][kTile / 32];
    const int i = blockIdx.x * kTile + threadIdx.x;
    const double big = 1e300;
    double x0 = big, x1 = -big, y0 = big, y1 = -big;
    if (i < n) {
        const double2 p = sorted[i];
        x0 = x1 = p.x;
        y0 = y1 = p.y;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        x0 = fmin(x0, __shfl_xor_sync(0xffffffffu, x0, d));
        x1 = fmax(x1, __shfl_xor_sync(0xffffffffu, x1, d));
        y0 = fmin(y0, __shfl_xor_sync(0xffffffffu, y0, d));
        y1 = fmax(y1, __shfl_xor_sync(0xffffffffu, y1, d));
    }
    if ((threadIdx.x & 31) == 0) {
        s[0][threadIdx.x >> 5] = x0; s[1][threadIdx.x >> 5] = x1;
        s[2][threadIdx.x >> 5] = y0; s[3][threadIdx.x >> 5] = y1;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < kTile / 32; w++) {
            x0 = fmin(x0, s[0][w]); x1 = fmax(x1, s[1][w]);
            y0 = fmin(y0, s[2][w]); y1 = fmax(y1, s[3][w]);
        }
        bbox[blockIdx.x] = make_double4(x0, x1, y0, y1);
    }
}

struct PcfArgs {
    int n, num_bins, use_smem;
    edmd_dev_box b;
    double dr, max_r;
    double s_max;        // r < max_r  <=>  s < s_max
    double dr_lo, dr_hi; // dr (1 + 2^-49) and dr (1 - 2^-49)
    float inv_dr;
    const double2 *sorted;
    const double4 *bbox;
    unsigned long long *counts;
    unsigned long long *stats;   // [0] pairs that took the exact path, [1] tile pairs skipped
    int part, nparts;            // this launch takes the tile pairs w = part (mod nparts)
};

// distance of the interval [lo, hi] from 0
__device__ __forceinline__ double gap(double lo, double hi) { return fmax(0.0, fmax(lo, -hi)); }

// 32-bit shared-window addresses and explicit ld.shared / red.shared keep the
// address arithmetic of the inner loop to one add per pair.
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ double2 lds_double2(uint32_t addr)
{
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}

__device__ __forceinline__ void bin_add(const PcfArgs &a, uint32_t hist_addr, int bin)
{
    if (hist_addr) asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(hist_addr + 4u * (uint32_t)bin) : "memory");
    else atomicAdd(&a.counts[bin], 1ull);
}

// predicated shared-memory increment: no branch in the instruction stream
__device__ __forceinline__ void bin_add_if(uint32_t hist_addr, int bin, bool p)
{
    asm volatile("{\n.reg .pred q;\nsetp.ne.s32 q, %2, 0;\n@q red.shared.add.u32 [%0], %1;\n}" ::"r"(
                     hist_addr + 4u * (uint32_t)bin),
                 "r"(1u), "r"((int)p)
                 : "memory");
}

// One pair: s exactly as the reference, the range test, the FP32 bin estimate and its
// FP64 certificate.  Returns the estimate k; `take` = in range, certified and a real
// bin; `redo` = in range but not certified.
template <bool WX, bool WY>
__device__ __forceinline__ int pair_bin(const PcfArgs &a, const double2 pi, const double2 pj, double s_max,
                                        double dr_lo, double dr_hi, float inv_dr, int num_bins, double &s,
                                        bool &take, bool &redo)
{
    double dx = __dsub_rn(pj.x, pi.x);
    double dy = __dsub_rn(pj.y, pi.y);
    if (WX) dx = min_image(dx, a.b.half_lx, a.b.lx);
    if (WY) dy = min_image(dy, a.b.half_ly, a.b.ly);
    s = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
    // FP32 estimate of r / dr from the bits of s (truncated to float).  For s outside
    // the float range the estimate is garbage and the certificate rejects it (or k = 0
    // is right anyway).
    const float sf = __int_as_float(((__double2hiint(s) - 0x38000000) << 3) |
                                    (int)((unsigned)__double2loint(s) >> 29));
    float rs;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(rs) : "f"(sf));
    const int k = __float2int_rz(sf * rs * inv_dr);
    // certificate in FP64: (k dr)^2 (1 + 2^-48) <= s < ((k+1) dr)^2 (1 - 2^-48)
    const double kd = __dsub_rn(__hiloint2double(0x43300000, k), 4503599627370496.0);
    const double e0 = __dmul_rn(kd, dr_lo);
    const double e1 = __dmul_rn(__dadd_rn(kd, 1.0), dr_hi);
    const bool in_range = s < s_max;   // r < max_r
    const bool ok = (s >= __dmul_rn(e0, e0)) && (s < __dmul_rn(e1, e1));
    take = in_range && ok && k < num_bins;
    redo = in_range && !ok;
    return k;
}

// Four pairs per trip, branch-free (their dependency chains interleave); the rare
// uncertified pair is redone with the reference's sqrt and division after the batch.
template <bool WX, bool WY>
__device__ __forceinline__ void pair_loop(const PcfArgs &a, const double2 pi, uint32_t tile_addr, int jstart,
                                          int jcount, uint32_t hist_addr, unsigned int &slow)
{
    const double s_max = a.s_max, dr_lo = a.dr_lo, dr_hi = a.dr_hi;
    const float inv_dr = a.inv_dr;
    const int num_bins = a.num_bins;
    uint32_t addr = tile_addr + 16u * (uint32_t)jstart;
    auto exact = [&](double s) {   // the reference's own arithmetic
        const int bin = (int)__ddiv_rn(__dsqrt_rn(s), a.dr);
        slow++;
        if (bin < num_bins) bin_add(a, hist_addr, bin);
    };
    int jj = jstart;
    if (hist_addr) {
        for (; jj + 4 <= jcount; jj += 4, addr += 64u) {
            double s[4];
            int k[4];
            bool take[4], redo[4];
#pragma unroll
            for (int u = 0; u < 4; u++)
                k[u] = pair_bin<WX, WY>(a, pi, lds_double2(addr + 16u * u), s_max, dr_lo, dr_hi, inv_dr, num_bins,
                                        s[u], take[u], redo[u]);
#pragma unroll
            for (int u = 0; u < 4; u++) bin_add_if(hist_addr, take[u] ? k[u] : 0, take[u]);
            if (redo[0] | redo[1] | redo[2] | redo[3]) {
#pragma unroll
                for (int u = 0; u < 4; u++)
                    if (redo[u]) exact(s[u]);
            }
        }
    }
    for (; jj < jcount; jj++, addr += 16u) {
        double s;
        bool take, redo;
        const int k = pair_bin<WX, WY>(a, pi, lds_double2(addr), s_max, dr_lo, dr_hi, inv_dr, num_bins, s, take, redo);
        if (take) bin_add(a, hist_addr, k);
        else if (redo) exact(s);
    }
}

__global__ void __launch_bounds__(kThreads)
k_pcf_sorted(const __grid_constant__ PcfArgs a)
{
    extern __shared__ unsigned char smem_raw[];
    double2 *tile = reinterpret_cast<double2 *>(smem_raw);
    unsigned int *hist = a.use_smem ? reinterpret_cast<unsigned int *>(smem_raw + kTile * sizeof(double2)) : nullptr;
    __shared__ int s_flags;
    if (hist)
        for (int k = threadIdx.x; k < a.num_bins; k += kThreads) hist[k] = 0;
    const int nt = (a.n + kTile - 1) / kTile;
    const long long npairs = (long long)nt * (nt + 1) / 2;
    unsigned int slow = 0, skipped = 0;
    const double rcut = a.max_r * (1.0 + 1e-12);
    for (long long w = (long long)blockIdx.x * a.nparts + a.part; w < npairs; w += (long long)gridDim.x * a.nparts) {
        // unrank w -> (ta, tb), ta <= tb, row-major over the upper triangle
        const double fn = (double)nt + 0.5;
        long long ta = (long long)(fn - sqrt(fn * fn - 2.0 * (double)w));
        while (ta * nt - ta * (ta - 1) / 2 > w) ta--;
        while ((ta + 1) * nt - (ta + 1) * ta / 2 <= w) ta++;
        const long long tb = ta + (w - (ta * nt - ta * (ta - 1) / 2));
        __syncthreads();   // previous tile / flags fully consumed
        if (threadIdx.x == 0) {
            // what the two boxes say about all pairs of this tile pair
            const double4 A = a.bbox[ta], B = a.bbox[tb];
            const double dx0 = B.x - A.y, dx1 = B.y - A.x;   // range of x_b - x_a
            const double dy0 = B.z - A.w, dy1 = B.w - A.z;
            const bool wx = !(dx1 < a.b.half_lx && dx0 >= -a.b.half_lx);
            const bool wy = !(dy1 < a.b.half_ly && dy0 >= -a.b.half_ly);
            double gx = gap(dx0, dx1), gy = gap(dy0, dy1);
            if (wx) gx = fmin(gx, fmin(gap(dx0 - a.b.lx, dx1 - a.b.lx), gap(dx0 + a.b.lx, dx1 + a.b.lx)));
            if (wy) gy = fmin(gy, fmin(gap(dy0 - a.b.ly, dy1 - a.b.ly), gap(dy0 + a.b.ly, dy1 + a.b.ly)));
            const bool skip = gx * gx + gy * gy >= rcut * rcut;
            s_flags = skip ? 4 : ((wx ? 1 : 0) | (wy ? 2 : 0));
        }
        __syncthreads();
        const int flags = s_flags;
        if (flags == 4) {
            skipped++;
            continue;
        }
        {
            const int jj = (int)tb * kTile + threadIdx.x;
            if (jj < a.n) tile[threadIdx.x] = a.sorted[jj];
        }
        __syncthreads();
        const int i = (int)ta * kTile + threadIdx.x;
        if (i < a.n) {
            const double2 pi = a.sorted[i];
            const int jcount = min(kTile, a.n - (int)tb * kTile);
            const int jstart = (ta == tb) ? threadIdx.x + 1 : 0;
            const uint32_t ta_addr = smem_addr(tile), h_addr = hist ? smem_addr(hist) : 0u;
            switch (flags) {
            case 0: pair_loop<false, false>(a, pi, ta_addr, jstart, jcount, h_addr, slow); break;
            case 1: pair_loop<true, false>(a, pi, ta_addr, jstart, jcount, h_addr, slow); break;
            case 2: pair_loop<false, true>(a, pi, ta_addr, jstart, jcount, h_addr, slow); break;
            default: pair_loop<true, true>(a, pi, ta_addr, jstart, jcount, h_addr, slow); break;
            }
        }
    }
    if (hist) {
        __syncthreads();
        for (int k = threadIdx.x; k < a.num_bins; k += kThreads) {
            const unsigned int v = hist[k];
            if (v) atomicAdd(&a.counts[k], (unsigned long long)v);
        }
    }
    if (a.stats) {
        if (slow) atomicAdd(&a.stats[0], (unsigned long long)slow);
        if (threadIdx.x == 0 && skipped) atomicAdd(&a.stats[1], (unsigned long long)skipped);
    }
}

}  // namespace

// smallest double s with correctly rounded sqrt(s) >= max_r
static double s_threshold(double max_r)
{
    double s = max_r * max_r;
    while (sqrt(s) >= max_r && s > 0) s = nextafter(s, 0.0);
    while (sqrt(s) < max_r) s = nextafter(s, INFINITY);
    return s;
}

// Sorted-tile g(r); ADDS into counts the tile pairs w = part (mod nparts).  Returns launches.
int edmd_launch_pcf_sorted(edmd_ctx *c, double dr, double max_r, int num_bins, const double *xy, int stride,
                           int n, int part, int nparts, unsigned long long *counts)
{
    if (n < 2 || num_bins <= 0) return 0;
    // coarse cells of ~192 particles, row-major
    const double area = c->box.lx * c->box.ly;
    const double side = sqrt(192.0 * area / (double)n);
    int gx = (int)(c->box.lx / side), gy = (int)(c->box.ly / side);
    if (gx < 1) gx = 1;
    if (gy < 1) gy = 1;
    const int ncoarse = gx * gy;
    const int nt = (n + kTile - 1) / kTile;
    const bool ordered = nparts > 1;   // every rank must cut identical tiles
    const size_t need = (size_t)n * sizeof(double2) * (ordered ? 2 : 1) + (size_t)nt * sizeof(double4) +
                        ((size_t)n * (ordered ? 2 : 1) + ncoarse + 8) * sizeof(int32_t) + 64;
    if (need > c->pcfs_bytes) {
        if (c->pcfs_mem) cudaFree(c->pcfs_mem);
        c->pcfs_mem = nullptr;
        c->pcfs_bytes = 0;
        if (cudaMalloc((void **)&c->pcfs_mem, need) != cudaSuccess) return -1;
        c->pcfs_bytes = need;
    }
    if (!c->pcfs_stats) {
        if (cudaMalloc((void **)&c->pcfs_stats, 2 * sizeof(unsigned long long)) != cudaSuccess) return -1;
        cudaMemsetAsync(c->pcfs_stats, 0, 2 * sizeof(unsigned long long), c->stream);
    }
    char *m = c->pcfs_mem;
    double2 *sorted = reinterpret_cast<double2 *>(m);
    m += (size_t)n * sizeof(double2);
    double2 *sorted2 = reinterpret_cast<double2 *>(m);
    if (ordered) m += (size_t)n * sizeof(double2);
    double4 *bbox = reinterpret_cast<double4 *>(m);
    m += (size_t)nt * sizeof(double4);
    int32_t *cnt = reinterpret_cast<int32_t *>(m);
    int32_t *cell = cnt + ncoarse + 8;
    int32_t *sidx = ordered ? cell + n : nullptr;
    unsigned long long *stats = reinterpret_cast<unsigned long long *>(c->pcfs_stats);
    SortArgs sa;
    sa.n = n; sa.stride = stride; sa.gx = gx; sa.gy = gy;
    sa.fx = gx / c->box.lx; sa.fy = gy / c->box.ly;
    sa.xy = xy; sa.cnt = cnt; sa.cell = cell; sa.sorted = sorted; sa.sidx = sidx;
    cudaMemsetAsync(cnt, 0, ((size_t)ncoarse + 8) * sizeof(int32_t), c->stream);
    const int pb = (n + kThreads - 1) / kThreads;
    k_coarse_count<<<pb, kThreads, 0, c->stream>>>(sa);
    k_small_scan<<<1, 1024, 0, c->stream>>>(ncoarse + 1, cnt);
    k_coarse_scatter<<<pb, kThreads, 0, c->stream>>>(sa);
    int launched = 5;
    if (ordered) {
        k_coarse_order<<<ncoarse, kThreads, 0, c->stream>>>(ncoarse, cnt, sorted, sidx, sorted2);
        sorted = sorted2;
        launched++;
    }
    k_tile_bbox<<<nt, kTile, 0, c->stream>>>(n, sorted, bbox);
    PcfArgs a;
    a.n = n; a.num_bins = num_bins;
    a.b = c->dbox;
    a.dr = dr; a.max_r = max_r;
    a.s_max = s_threshold(max_r);
    a.dr_lo = dr * (1.0 + 1.7763568394002505e-15);   // 2^-49
    a.dr_hi = dr * (1.0 - 1.7763568394002505e-15);
    a.inv_dr = (float)(1.0 / dr);
    a.sorted = sorted; a.bbox = bbox; a.counts = counts; a.stats = stats;
    a.part = part; a.nparts = nparts;
    const size_t tile_bytes = kTile * sizeof(double2);
    const size_t hist_bytes = (size_t)num_bins * sizeof(unsigned int);
    a.use_smem = (tile_bytes + hist_bytes) <= 200 * 1024;
    const size_t smem = tile_bytes + (a.use_smem ? hist_bytes : 0);
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(k_pcf_sorted, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024);
        attr = true;
    }
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pcf_sorted, kThreads, smem);
    if (per_sm < 1) per_sm = 1;
    long long grid = (long long)(c->sm_count > 0 ? c->sm_count : 148) * per_sm;
    const long long npairs = ((long long)nt * (nt + 1) / 2 + nparts - 1) / nparts;
    if (grid > npairs) grid = npairs;
    if (grid < 1) grid = 1;
    k_pcf_sorted<<<(int)grid, kThreads, smem, c->stream>>>(a);
    return launched;
}
