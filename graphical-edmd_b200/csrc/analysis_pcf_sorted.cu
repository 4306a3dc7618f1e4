// analysis_pcf_sorted.cu -- K3: full-range g(r) (calculate_pcf, src/pcf.c:16-75) over
// spatially sorted tiles.  The reference bins  bin = (int)(sqrt(dx^2 + dy^2) / dr)  for all
// N(N-1)/2 pairs after the minimum-image adjustment (PBC, src/EDMD.c:5896-5913); the integer
// counts must be the reference's.  Two kernels share the tiles:
//
//   k_pcf_f32     (default, second half of this file) decides the bin of a pair in packed FP32
//                 under a rigorous error bound and redoes the ~0.4 % of pairs FP32 cannot settle
//                 with the reference's own FP64 operations; ~10 instructions per pair.
//   k_pcf_sorted  (EDMD_OPT_PCF_LEGACY = 2; very fine bins) computes per pair only
//                     s = fl(fl(dx*dx) + fl(dy*dy))      exactly as the reference (unfused FP64)
//                 and then
//                   * range test in s-space: r < max_r  <=>  s < S_max, with S_max the smallest
//                     double whose correctly rounded square root reaches max_r (found on the host);
//                   * an FP32 estimate k of the bin (float taken from the bits of s, MUFU rsqrt);
//                   * the certificate  (k dr)^2 (1 + 2^-48) <= s < ((k+1) dr)^2 (1 - 2^-48)  in FP64
//                     (5 multiplications, 2 additions, 2 compares): the two roundings of
//                     sqrt-then-divide move r/dr by at most 2^-52 relative, so inside that
//                     interval the reference's truncation gives k.  A pair that fails takes the
//                     reference's sqrt and division.  14 FP64 operations per pair instead of ~35;
//                     ~42 instructions per pair.
//
// Tiles are 256 consecutive particles of an array sorted by coarse cell (row-major, alternate
// rows reversed), with exact bounding boxes.  For a tile pair the boxes tell whether any pair can
// need the periodic image in x / in y and whether every pair is farther than max_r (the tile pair
// is skipped: for max_r = L/2 that is ~23 % of them).
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
    if (cy & 1) cx = a.gx - 1 - cx;   // boustrophedon: consecutive tiles stay compact at row ends
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

// exact bounding box of each tile; optionally also the tile centre and the FP32
// coordinates of its particles relative to that centre in units of dr (k_pcf_f32)
__global__ void __launch_bounds__(kTile)
k_tile_bbox(int n, const double2 *__restrict__ sorted, double4 *__restrict__ bbox, double inv_dr,
            double2 *__restrict__ ctr, float2 *__restrict__ rel, double coord_max)
{
    __shared__ double s[4][kTile / 32];
    __shared__ double2 s_ctr;
    const int i = blockIdx.x * kTile + threadIdx.x;
    const double big = 1e300;
    double x0 = big, x1 = -big, y0 = big, y1 = -big;
    double2 p = make_double2(0.0, 0.0);
    int bad = 0;
    if (i < n) {
        p = sorted[i];
        x0 = x1 = p.x;
        y0 = y1 = p.y;
        // non-finite or far outside the box: the FP32 offsets of this tile mean nothing
        bad = !(fabs(p.x) <= coord_max && fabs(p.y) <= coord_max);
    }
    bad = __syncthreads_or(bad);
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
        const double qnan = __longlong_as_double(0x7ff8000000000000ll);
        bbox[blockIdx.x] = (bad && rel) ? make_double4(qnan, qnan, qnan, qnan) : make_double4(x0, x1, y0, y1);
        if (rel) {
            s_ctr = make_double2(0.5 * (x0 + x1), 0.5 * (y0 + y1));
            ctr[blockIdx.x] = s_ctr;
        }
    }
    if (rel) {
        __syncthreads();
        if (i < n) {
            const double2 c = s_ctr;
            rel[i] = make_float2(__double2float_rn((p.x - c.x) * inv_dr), __double2float_rn((p.y - c.y) * inv_dr));
        }
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

// ---------------------------------------------------------------------------
// k_pcf_f32 -- the bin of a pair decided in FP32 with a RIGOROUS error bound,
// the FP64 arithmetic of the reference only for the pairs FP32 cannot settle.
//
// Per tile the particles are stored as FP32 offsets from the tile centre in units
// of dr (b_j, by k_tile_bbox).  Per tile pair (a, b) one thread derives from the two
// exact bounding boxes
//   * the periodic image every pair of the tile pair takes (S = 0, -L or +L per
//     axis), or that the pairs are mixed (then |d| >= L/2 -> |d| - L per pair, in
//     FP32: the magnitude after the reference's single wrap is a 1-Lipschitz
//     function of d, so a different decision next to the threshold costs no more
//     than the rounding error that caused it);
//   * a bound eps on |q32 - q*|, q* = the reference's FP64 sqrt(dx^2+dy^2)/dr and
//     q32 its FP32 estimate (derivation below);
//   * whether every pair is in range (then the range test is dropped).
// Thread i holds p_i = fl32(((x_i - S) - C_b) / dr); per pair
//     d = b_j - p_i,  s = fma(dy, dy, fma(dx, dx, 1e-30)),  y = rsqrt(s),
//     q_lo = fma(s, y, -eps)                                  (q* lies in [q_lo, q_lo + 2 eps])
//     t = RD(q_lo + 1.5 * 2^23)     (bits of t = 0x4B400000 + floor(q_lo), exact for |q| < 2^22)
//     frac = q_lo - (t - 1.5 * 2^23)                          (exact)
// and the pair is CERTAIN when frac < 1 - 2 eps -- no integer in (q_lo, q_lo + 2 eps], so
// floor(q*) = floor(q_lo).  Every pair is counted in floor(q_lo) at once (clamped to one dummy
// word at num_bins: `bin < num_bins` is the reference's whole range test, see pair32);
// everything else goes to a shared-memory queue that the CTA drains with the
// reference's own FP64 operations (~0.6 % of the pairs at N = 10^6, dr = 0.1) and moves the
// count to the exact bin when it differs.
//
// Error bound (u = 2^-24, lengths in units of dr):
//   b_j, p_i are roundings of FP64 values: |err| <= u |b_j| <= u h_b and
//   u |p_i| <= u (h_b + M) with h_b the half extent of tile b and M the largest
//   |d| of the tile pair; the subtraction adds u |d| <= u M:
//       E_axis = u (2 h_b + 2 M) (+ u L when the FP32 wrap subtracts fl32(L)),
//   E = |(E_x, E_y)|; by the triangle inequality the exact norm of the FP32 vector is
//   within E of the true distance.  s carries two roundings (<= 2u), the MUFU
//   rsqrt <= 1.28e-7 relative (PTX documents 2^-22.9 = 1.2775e-7; edmd_cuda_selftest_rsqrt measures it
//   exhaustively -- 1.2467e-7 on B200 -- and the GPU test asserts <= 1.28e-7), the product one more:
//       |s y - q_true| <= (R + E) (1.28e-7 + 1.01u) + E + 1e-15 (the 1e-30 floor)
//   with R the largest distance of the tile pair; the rounding of the fma adds
//   u (R + E) more: eps = (R + E) (1.28e-7 + 2.1u) + E + slack.  FP64 roundings on either side
//   (ours and the reference's) stay below 2^-48 L/dr; 2^-44 (L/dr + 1) is budgeted.
// ---------------------------------------------------------------------------
constexpr int kQueue = 2048;                 // undecided pairs parked per tile pair (u32 each: counted bin | i | j)
constexpr float kMagic = 12582912.0f;        // 1.5 * 2^23
constexpr uint32_t kMagicBits = 0x4B400000u;

struct TilePairPlan {
    double2 cbs;         // C_b + S: p_i = ((x_i - cbs.x) / dr, ...)
    float eps;           // q* in [q_lo, q_lo + 2 eps], q_lo = fma(s, y, -eps)
    float cth;           // certain  <=>  frac(q_lo) < cth   (cth = 1 - 2 eps, rounded down)
    int ta, tb;          // the tile pair
    int flags;           // 1 wrap x per pair, 2 wrap y per pair, 4 every pair certainly below num_bins, 8 skip,
                         // 16 every pair by the exact path (bad coordinates, or eps too large to decide anything)
};

struct F32Args {
    PcfArgs p;
    const float2 *rel;
    const double2 *ctr;
    double inv_dr, slack;
    float tmax;                     // 1.5 * 2^23 + num_bins: t of the dummy word behind the histogram
    float lx32, ly32, hx32, hy32;   // fl32(L / dr) and exactly half of it
};

__device__ __forceinline__ float lds_float(uint32_t addr)
{
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}

// ---- packed FP32 (sm_100: FADD2 / FFMA2, two lanes per instruction and per issue slot) ----
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi)
{
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float &lo, float &hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b)
{
    f32x2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b)
{
    f32x2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 add2_rd(f32x2 a, f32x2 b)
{
    f32x2 r;
    asm("add.rm.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c)
{
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ void lds_2x64(uint32_t addr, f32x2 &a, f32x2 &b)
{
    asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "r"(addr));
}

__device__ __forceinline__ float4 lds_float4(uint32_t addr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

// the reference's own arithmetic for one pair (pcf.c:34-47 + PBC), into the shared histogram;
// `counted` = the FP32 bin this pair was already counted in (kNotCounted: none)
constexpr unsigned int kNotCounted = 0xFFFFu;

__device__ __forceinline__ void exact_pair(const PcfArgs &a, int gi, int gj, uint32_t hist_addr, unsigned int counted)
{
    const double2 pi = a.sorted[gi], pj = a.sorted[gj];
    const double dx = min_image(__dsub_rn(pj.x, pi.x), a.b.half_lx, a.b.lx);
    const double dy = min_image(__dsub_rn(pj.y, pi.y), a.b.half_ly, a.b.ly);
    const double s = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
    unsigned int bin = kNotCounted;
    if (s < a.s_max) {
        const int k = (int)__ddiv_rn(__dsqrt_rn(s), a.dr);
        if (k < a.num_bins) bin = (unsigned int)k;
    }
    if (bin == counted) return;
    if (counted != kNotCounted) asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(hist_addr + 4u * counted), "r"(0xFFFFFFFFu) : "memory");
    if (bin != kNotCounted) asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(hist_addr + 4u * bin), "r"(1u) : "memory");
}

// One pair in FP32.  The pair is counted in its FP32 bin unconditionally -- no predicate, no
// branch -- and an undecided pair (frac >= cth) is corrected by the exact path afterwards
// (minus one here, plus one there).  `bin < num_bins` is the whole range test: num_bins =
// (int)(max_r / dr) and rounding is monotone, so fl(r / dr) < num_bins implies r < max_r
// (src/pcf.c:44-47).  Unless every pair of the tile pair is certainly below num_bins (INR), t is
// clamped to num_bins: one dummy word behind the histogram takes every pair beyond the range.
// Returns frac(q_lo); tbits = bits of the (clamped) t, bin = tbits - kMagicBits.
template <bool WX, bool WY, bool INR>
__device__ __forceinline__ float pair32(float bx, float by, float px, float py, float eps, float tmax, float lx32,
                                        float ly32, float hx32, float hy32, uint32_t hbase, uint32_t &tbits)
{
    float dx = __fsub_rn(bx, px), dy = __fsub_rn(by, py);
    if (WX) {
        const float m = fabsf(dx);
        dx = m >= hx32 ? __fsub_rn(m, lx32) : m;
    }
    if (WY) {
        const float m = fabsf(dy);
        dy = m >= hy32 ? __fsub_rn(m, ly32) : m;
    }
    const float s = __fmaf_rn(dy, dy, __fmaf_rn(dx, dx, 1e-30f));
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(s));
    const float qlo = __fmaf_rn(s, y, -eps);                 // q_lo <= q* <= q_lo + 2 eps
    const float t = __fadd_rd(qlo, kMagic);
    const float frac = __fsub_rn(qlo, __fsub_rn(t, kMagic));   // exact, in [0, 1)
    tbits = __float_as_uint(INR ? t : fminf(t, tmax));
    // q_lo < 0 (a pair closer than eps dr) gives bin -1: the word in front of the histogram is a pad
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(hbase + (tbits << 2)), "r"(1u) : "memory");
    return frac;
}

// Two pairs (j, j+1) of one i at once with the packed instructions: lane for lane the same
// round-to-nearest / round-down operations as pair32, so the same bits.  No wrap variant.
struct Packed32 {
    f32x2 px, py, neps, tiny, magic, nmagic;
};
// (Measured: doing the two subtractions as scalar FADDs, which either FMA pipe takes while the packed
// forms only run on the heavy one, is slower -- 283 against 278 ms at N = 10^6.)
template <bool INR>
__device__ __forceinline__ void pair32x2(f32x2 bx, f32x2 by, const Packed32 &k, float tmax, uint32_t hbase,
                                         float &f0, float &f1, uint32_t &t0, uint32_t &t1)
{
    const f32x2 dx = sub2(bx, k.px), dy = sub2(by, k.py);
    const f32x2 s = fma2(dy, dy, fma2(dx, dx, k.tiny));
    float s0, s1, y0, y1;
    unpack2(s, s0, s1);
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y0) : "f"(s0));
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y1) : "f"(s1));
    const f32x2 qlo = fma2(s, pack2(y0, y1), k.neps);
    const f32x2 t = add2_rd(qlo, k.magic);
    const f32x2 frac = sub2(qlo, add2(t, k.nmagic));
    float ta, tb;
    unpack2(t, ta, tb);
    unpack2(frac, f0, f1);
    t0 = __float_as_uint(INR ? ta : fminf(ta, tmax));
    t1 = __float_as_uint(INR ? tb : fminf(tb, tmax));
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(hbase + (t0 << 2)), "r"(1u) : "memory");
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(hbase + (t1 << 2)), "r"(1u) : "memory");
}

template <bool WX, bool WY, bool INR>
__device__ __forceinline__ void pair_loop32(const F32Args &a, const TilePairPlan &tp, float px, float py, int gi0,
                                            int gj0, uint32_t tile_addr, int jstart, int jcount, uint32_t hist_addr,
                                            uint32_t queue_addr, uint32_t qn, unsigned int &slow, unsigned int tid)
{
    const float c = tp.eps, cth = tp.cth, tmax = a.tmax;
    const float lx32 = a.lx32, ly32 = a.ly32, hx32 = a.hx32, hy32 = a.hy32;
    const uint32_t hbase = hist_addr - (kMagicBits << 2);
    const uint32_t tmaxbits = __float_as_uint(tmax);
    auto park = [&](int j, uint32_t tbits) {
        const unsigned int counted = (tbits - kMagicBits) & 0xFFFFu;   // -1 (the pad) reads as kNotCounted
        uint32_t slot;
        // (tbits >> 31) is always 0 (t lies in [2^23, 2^24)), but an address ptxas can prove uniform makes it
        // emit the warp-aggregated form of the atomic: ~25 instructions for one or two lanes at a time
        asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(slot) : "r"(qn + ((tbits >> 31) << 2)) : "memory");
        if (slot < (uint32_t)kQueue) {
            asm volatile("st.shared.u32 [%0], %1;" ::"r"(queue_addr + 4u * (uint32_t)slot),
                         "r"((counted << 16) | (tid << 8) | (unsigned int)j)
                         : "memory");
        } else {
            exact_pair(a.p, gi0 + (int)tid, gj0 + j, hist_addr, counted);
            slow++;
        }
    };
    int jj = jstart;
    uint32_t tb;
    constexpr uint32_t kYOff = 4u * kTile;   // the tile is stored as x[kTile], y[kTile]
    for (; jj < jcount && (jj & 3); jj++) {
        const uint32_t ad = tile_addr + 4u * (uint32_t)jj;
        if (pair32<WX, WY, INR>(lds_float(ad), lds_float(ad + kYOff), px, py, c, tmax, lx32, ly32, hx32, hy32, hbase, tb) >= cth && (INR || tb != tmaxbits)) park(jj, tb);
    }
    uint32_t addr = tile_addr + 4u * (uint32_t)jj;
    Packed32 pk;
    if (!WX && !WY) {
        pk.px = pack2(px, px); pk.py = pack2(py, py);
        pk.neps = pack2(-c, -c); pk.tiny = pack2(1e-30f, 1e-30f);
        pk.magic = pack2(kMagic, kMagic); pk.nmagic = pack2(-kMagic, -kMagic);
    }
    // measured at N = 10^6: scalar 293 ms, packed four pairs per trip 278 ms, packed eight per trip 271 ms
    if (!WX && !WY) {
        for (; jj + 8 <= jcount; jj += 8, addr += 32u) {
            uint32_t t[8];
            float u[8];
            f32x2 x01, x23, x45, x67, y01, y23, y45, y67;
            lds_2x64(addr, x01, x23);
            lds_2x64(addr + 16u, x45, x67);
            lds_2x64(addr + kYOff, y01, y23);
            lds_2x64(addr + kYOff + 16u, y45, y67);
            pair32x2<INR>(x01, y01, pk, tmax, hbase, u[0], u[1], t[0], t[1]);
            pair32x2<INR>(x23, y23, pk, tmax, hbase, u[2], u[3], t[2], t[3]);
            pair32x2<INR>(x45, y45, pk, tmax, hbase, u[4], u[5], t[4], t[5]);
            pair32x2<INR>(x67, y67, pk, tmax, hbase, u[6], u[7], t[6], t[7]);
            if (fmaxf(fmaxf(fmaxf(u[0], u[1]), fmaxf(u[2], u[3])), fmaxf(fmaxf(u[4], u[5]), fmaxf(u[6], u[7]))) >= cth) {
#pragma unroll
                for (int q = 0; q < 8; q++)
                    if (u[q] >= cth && (INR || t[q] != tmaxbits)) park(jj + q, t[q]);
            }
        }
    }
    for (; jj + 4 <= jcount; jj += 4, addr += 16u) {
        uint32_t t0, t1, t2, t3;
        float u0, u1, u2, u3;
        if (!WX && !WY) {
            f32x2 x01, x23, y01, y23;
            lds_2x64(addr, x01, x23);
            lds_2x64(addr + kYOff, y01, y23);
            pair32x2<INR>(x01, y01, pk, tmax, hbase, u0, u1, t0, t1);
            pair32x2<INR>(x23, y23, pk, tmax, hbase, u2, u3, t2, t3);
        } else {
            const float4 bx = lds_float4(addr), by = lds_float4(addr + kYOff);
            u0 = pair32<WX, WY, INR>(bx.x, by.x, px, py, c, tmax, lx32, ly32, hx32, hy32, hbase, t0);
            u1 = pair32<WX, WY, INR>(bx.y, by.y, px, py, c, tmax, lx32, ly32, hx32, hy32, hbase, t1);
            u2 = pair32<WX, WY, INR>(bx.z, by.z, px, py, c, tmax, lx32, ly32, hx32, hy32, hbase, t2);
            u3 = pair32<WX, WY, INR>(bx.w, by.w, px, py, c, tmax, lx32, ly32, hx32, hy32, hbase, t3);
        }
        if (fmaxf(fmaxf(u0, u1), fmaxf(u2, u3)) >= cth) {
            // (a pair clamped to the dummy word is certainly beyond the range: q* >= q_lo >= num_bins)
            if (u0 >= cth && (INR || t0 != tmaxbits)) park(jj, t0);
            if (u1 >= cth && (INR || t1 != tmaxbits)) park(jj + 1, t1);
            if (u2 >= cth && (INR || t2 != tmaxbits)) park(jj + 2, t2);
            if (u3 >= cth && (INR || t3 != tmaxbits)) park(jj + 3, t3);
        }
    }
    for (; jj < jcount; jj++) {
        const uint32_t ad = tile_addr + 4u * (uint32_t)jj;
        if (pair32<WX, WY, INR>(lds_float(ad), lds_float(ad + kYOff), px, py, c, tmax, lx32, ly32, hx32, hy32, hbase, tb) >= cth && (INR || tb != tmaxbits)) park(jj, tb);
    }
}

// One axis of the tile-pair plan: the image shift S every pair takes (or per-pair
// wrapping), the largest |d| before (m_un) and after (m_w) the wrap, the gap for
// the skip test.  d0, d1 = range of fl(x_b - x_a) over the tile pair.
__device__ __forceinline__ void plan_axis(double d0, double d1, double half, double len, double &shift, bool &wrap,
                                          double &m_un, double &m_w, double &g)
{
    shift = 0.0;
    wrap = false;
    if (d1 < half && d0 >= -half) {
        g = gap(d0, d1);
    } else if (d0 >= half) {
        shift = -len;
        g = gap(d0 - len, d1 - len);
    } else if (d1 < -half) {
        shift = len;
        g = gap(d0 + len, d1 + len);
    } else {
        wrap = true;
        g = fmin(gap(d0, d1), fmin(gap(d0 - len, d1 - len), gap(d0 + len, d1 + len)));
    }
    m_un = fmax(fabs(d0 + shift), fabs(d1 + shift)) * (1.0 + 1e-12);
    m_w = wrap ? fmax(half, m_un - len) : m_un;
}

// The plan of tile pair w (one thread): skip / image shifts / eps / flags, see the header comment.
__device__ __forceinline__ TilePairPlan make_plan(const F32Args &a, long long w, int nt, double rcut)
{
    const PcfArgs &p = a.p;
    TilePairPlan tp;
    // unrank w -> (ta, tb), ta <= tb, row-major over the upper triangle
    const double fn = (double)nt + 0.5;
    long long ta = (long long)(fn - sqrt(fn * fn - 2.0 * (double)w));
    while (ta * nt - ta * (ta - 1) / 2 > w) ta--;
    while ((ta + 1) * nt - (ta + 1) * ta / 2 <= w) ta++;
    const long long tb = ta + (w - (ta * nt - ta * (ta - 1) / 2));
    const double4 A = p.bbox[ta], B = p.bbox[tb];
    double sx, sy, mxu, mxw, myu, myw, gx, gy;
    bool wx, wy;
    plan_axis(B.x - A.y, B.y - A.x, p.b.half_lx, p.b.lx, sx, wx, mxu, mxw, gx);
    plan_axis(B.z - A.w, B.w - A.z, p.b.half_ly, p.b.ly, sy, wy, myu, myw, gy);
    tp.ta = (int)ta;
    tp.tb = (int)tb;
    tp.eps = 0.0f;
    tp.cth = 0.0f;
    tp.cbs = make_double2(0.0, 0.0);
    if (A.x != A.x || B.x != B.x) {
        tp.flags = 16;   // a tile with non-finite or far-out coordinates (k_tile_bbox): FP64 for every pair
        return tp;
    }
    if (gx * gx + gy * gy >= rcut * rcut) {
        tp.flags = 8;
        return tp;
    }
    const double u = 5.9604644775390625e-08 * 1.001;   // 2^-24, padded
    const double id = a.inv_dr;
    const double hbx = 0.5 * (B.y - B.x) * id, hby = 0.5 * (B.w - B.z) * id;
    const double ex = u * (2.0 * hbx + 2.0 * mxu * id) + (wx ? u * p.b.lx * id : 0.0) + a.slack;
    const double ey = u * (2.0 * hby + 2.0 * myu * id) + (wy ? u * p.b.ly * id : 0.0) + a.slack;
    const double e = sqrt(ex * ex + ey * ey) * 1.001;
    const double rmax = sqrt(mxw * mxw + myw * myw) * id * (1.0 + 1e-9);
    // rsqrt.approx <= 2^-22.9 relative (PTX; measured 1.2467e-7 on B200, asserted by the GPU test);
    // 1.01u from the two roundings of s under the square root, u from the fma of q_lo
    const double rho = 1.28e-07 + 2.1 * u;
    const double eps = ((rmax + e) * rho + e + 1e-12 + a.slack) * 1.001;
    const float eps32 = __double2float_ru(eps);
    tp.eps = eps32;
    tp.cth = __double2float_rd(1.0 - 2.0 * (double)eps32 - 2.384185791015625e-07);
    const bool inr = (rmax + e) * (1.0 + 1e-6) < (double)p.num_bins;   // every q_lo below num_bins
    tp.flags = !(eps < 0.45) ? 16 : ((wx ? 1 : 0) | (wy ? 2 : 0) | ((inr && !wx && !wy) ? 4 : 0));
    const double2 cb = a.ctr[tb];
    tp.cbs = make_double2(cb.x + sx, cb.y + sy);
    return tp;
}

constexpr int kPlanBatch = 64;   // tile pairs planned at once (one thread each)
constexpr size_t kGroupSmem = kTile * sizeof(float2) + kQueue * sizeof(unsigned int);   // tile + queue of one group

__device__ __forceinline__ void group_barrier(int group)
{
    asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "r"(kThreads) : "memory");
}

// A CTA is G groups of 256 threads that share ONE histogram (the histogram is what limits the CTAs
// per SM: 39 KB at N = 10^6, 79 KB at 4*10^6, 112 KB at 8*10^6 -- with G = 4 one resident CTA still
// gives 32 warps).  Each group is what a CTA used to be: its own tile, queue, plans and named barrier,
// and it takes the tile pairs w = first + k * stride of its virtual CTA id.  Plans are made kPlanBatch
// at a time, one thread each (the plan is a chain of ~60 dependent FP64 operations: made by one thread
// per tile pair it was 4 % of the kernel and held eight warps at a barrier).  Per tile pair two
// barriers: the drain of the previous pair's queue shares a phase with loading this pair's tile.
template <int G>
__global__ void __launch_bounds__(kThreads * G, G == 1 ? 4 : (G == 2 ? 2 : 1))   // 64 registers
k_pcf_f32(const __grid_constant__ F32Args a)
{
    extern __shared__ unsigned char smem_raw[];
    const int group = threadIdx.x / kThreads;
    const unsigned int tid = threadIdx.x % kThreads;
    float *tile = reinterpret_cast<float *>(smem_raw + group * kGroupSmem);   // x[kTile], y[kTile]
    unsigned int *queue = reinterpret_cast<unsigned int *>(smem_raw + group * kGroupSmem + kTile * sizeof(float2));
    // [pad: 4 words][histogram][dummy words]: bin -1 (see pair32) lands in the pad, bins >= num_bins in hist[num_bins]
    unsigned int *hist = reinterpret_cast<unsigned int *>(smem_raw + G * kGroupSmem) + 4;
    __shared__ TilePairPlan s_plans_all[G][kPlanBatch];
    __shared__ int s_qn_all[G][2];
    TilePairPlan *s_plans = s_plans_all[group];
    int *s_qn = s_qn_all[group];
    const PcfArgs &p = a.p;
    for (int k = threadIdx.x; k < p.num_bins; k += kThreads * G) hist[k] = 0;
    if (tid < 2) s_qn[tid] = 0;
    __syncthreads();
    const int nt = (p.n + kTile - 1) / kTile;
    const long long npairs = (long long)nt * (nt + 1) / 2;
    unsigned int slow = 0;
    unsigned long long skipped = 0;
    const double rcut = p.max_r * (1.0 + 1e-12);
    const uint32_t tile_addr = smem_addr(tile), hist_addr = smem_addr(hist), queue_addr = smem_addr(queue);
    const long long wstride = (long long)gridDim.x * G * p.nparts;
    long long wbase = ((long long)blockIdx.x * G + group) * p.nparts + p.part;
    int par = 0;                 // which queue counter the next tile pair parks with
    bool have_prev = false;      // a tile pair whose queue is not drained yet
    int pgi0 = 0, pgj0 = 0;
    // drain: the parked pairs of the previous tile pair with the reference's FP64 operations
    auto drain = [&](int counter) {
        const int qn = min(s_qn[counter], kQueue);
        for (int k = tid; k < qn; k += kThreads) {
            const unsigned int e = queue[k];
            exact_pair(p, pgi0 + (int)((e >> 8) & 255u), pgj0 + (int)(e & 255u), hist_addr, e >> 16);
            slow++;
        }
    };
    for (bool done = false; !done; wbase += (long long)kPlanBatch * wstride) {
        if (have_prev) {
            drain(par ^ 1);
            have_prev = false;
        }
        group_barrier(group);   // everybody is through with the previous batch of plans (and the drain)
        if (tid < kPlanBatch) {
            const long long w = wbase + (long long)tid * wstride;
            TilePairPlan tp;
            tp.flags = 32;   // no tile pair left
            if (w < npairs) {
                tp = make_plan(a, w, nt, rcut);
                if (tp.flags == 8) skipped++;
            }
            s_plans[tid] = tp;
        }
        group_barrier(group);
        for (int k = 0; k < kPlanBatch; k++) {
            const TilePairPlan &tp = s_plans[k];
            const int flags = tp.flags;
            if (flags == 32) {
                done = true;
                break;
            }
            if (flags == 8) continue;
            const int ta = tp.ta, tb = tp.tb;
            const int gi0 = ta * kTile, gj0 = tb * kTile;
            // phase 1: the previous pair's queue is drained while this pair's tile comes in
            if (have_prev) drain(par ^ 1);
            if (gj0 + (int)tid < p.n) {
                const float2 b = a.rel[gj0 + tid];
                tile[tid] = b.x;
                tile[kTile + tid] = b.y;
            }
            const int i = gi0 + tid;
            float px = 0.0f, py = 0.0f;
            if (i < p.n && flags != 16) {
                const double2 pi = p.sorted[i];
                px = __double2float_rn((pi.x - tp.cbs.x) * a.inv_dr);
                py = __double2float_rn((pi.y - tp.cbs.y) * a.inv_dr);
            }
            group_barrier(group);
            // phase 2: the pairs of this tile pair
            if (tid == 0) s_qn[par ^ 1] = 0;   // drained in phase 1; next used by the next tile pair
            if (i < p.n) {
                const int jcount = min(kTile, p.n - gj0);
                const int jstart = (ta == tb) ? (int)tid + 1 : 0;
                const uint32_t qn_addr = smem_addr(&s_qn[par]);
                switch (flags) {
                case 16:
                    for (int jj = jstart; jj < jcount; jj++) exact_pair(p, i, gj0 + jj, hist_addr, kNotCounted);
                    slow += (unsigned int)max(jcount - jstart, 0);
                    break;
                case 4: pair_loop32<false, false, true>(a, tp, px, py, gi0, gj0, tile_addr, jstart, jcount, hist_addr, queue_addr, qn_addr, slow, tid); break;
                case 0: pair_loop32<false, false, false>(a, tp, px, py, gi0, gj0, tile_addr, jstart, jcount, hist_addr, queue_addr, qn_addr, slow, tid); break;
                case 1: pair_loop32<true, false, false>(a, tp, px, py, gi0, gj0, tile_addr, jstart, jcount, hist_addr, queue_addr, qn_addr, slow, tid); break;
                case 2: pair_loop32<false, true, false>(a, tp, px, py, gi0, gj0, tile_addr, jstart, jcount, hist_addr, queue_addr, qn_addr, slow, tid); break;
                default: pair_loop32<true, true, false>(a, tp, px, py, gi0, gj0, tile_addr, jstart, jcount, hist_addr, queue_addr, qn_addr, slow, tid); break;
                }
            }
            group_barrier(group);
            pgi0 = gi0;
            pgj0 = gj0;
            have_prev = true;
            par ^= 1;
        }
    }
    if (have_prev) drain(par ^ 1);
    __syncthreads();
    for (int k = threadIdx.x; k < p.num_bins; k += kThreads * G) {
        const unsigned int v = hist[k];
        if (v) atomicAdd(&p.counts[k], (unsigned long long)v);
    }
    if (p.stats) {
        if (slow) atomicAdd(&p.stats[0], (unsigned long long)slow);
        if (skipped) atomicAdd(&p.stats[1], skipped);
    }
}

// largest relative error of the MUFU reciprocal square root over every float in
// [2^-100, 2^64) (the kernel's s lies in [1e-30, 2^46)): the measured side of the
// 1.28e-7 budgeted in k_pcf_f32
__global__ void __launch_bounds__(256)
k_rsqrt_selftest(unsigned long long *worst_bits)
{
    double worst = 0.0;
    const uint32_t lo = (127u - 100u) << 23, hi = (127u + 64u) << 23;
    for (uint64_t b = lo + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; b < hi; b += (uint64_t)gridDim.x * blockDim.x) {
        const float x = __uint_as_float((uint32_t)b);
        float y;
        asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
        const double rel = fabs((double)y * sqrt((double)x) - 1.0);
        worst = fmax(worst, rel);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) worst = fmax(worst, __shfl_xor_sync(0xffffffffu, worst, d));
    if ((threadIdx.x & 31) == 0) atomicMax(worst_bits, (unsigned long long)__double_as_longlong(worst));
}

}  // namespace

int edmd_launch_rsqrt_selftest(edmd_ctx *c, unsigned long long *worst_bits_dev)
{
    k_rsqrt_selftest<<<(c->sm_count > 0 ? c->sm_count : 148) * 8, 256, 0, c->stream>>>(worst_bits_dev);
    return 1;
}

// smallest double s with correctly rounded sqrt(s) >= max_r
double edmd_pcf_s_threshold(double max_r)
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
    // coarse cells of ~192 particles, row-major (alternate rows reversed)
    const double area = c->box.lx * c->box.ly;
    const double side = sqrt(192.0 * area / (double)n);
    int gx = (int)(c->box.lx / side), gy = (int)(c->box.ly / side);
    if (gx < 1) gx = 1;
    if (gy < 1) gy = 1;
    const int ncoarse = gx * gy;
    const int nt = (n + kTile - 1) / kTile;
    const bool ordered = nparts > 1;   // every rank must cut identical tiles
    // FP32-decided bins (k_pcf_f32) when the histogram fits in shared memory next to the
    // queue, q = r/dr stays far below 2^22 and the expected share of undecided pairs
    // (~2 q_max * 5.4e-7) is small; else bins certified in FP64 (k_pcf_sorted)
    const double lmax = c->box.lx > c->box.ly ? c->box.lx : c->box.ly;
    const size_t f32_hist = ((size_t)num_bins + 4 + 32) * sizeof(unsigned int);
    const size_t f32_smem = kGroupSmem + f32_hist;   // with one group per CTA
    const bool f32 = c->pcf_mode == 0 && f32_smem <= 200 * 1024 && num_bins < 65535 && dr > 0.0 && max_r > 0.0 &&
                     num_bins <= (int)(max_r / dr) &&   // then `bin < num_bins` implies r < max_r (see pair32)
                     1.5 * lmax / dr < 2097152.0 && max_r / dr <= 60000.0 && lmax / dr <= 120000.0;
    const size_t need = (size_t)n * sizeof(double2) * (ordered ? 2 : 1) + (size_t)nt * (sizeof(double4) + sizeof(double2)) +
                        (size_t)n * sizeof(float2) + ((size_t)n * (ordered ? 2 : 1) + ncoarse + 8) * sizeof(int32_t) + 64;
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
    double2 *ctr = reinterpret_cast<double2 *>(m);
    m += (size_t)nt * sizeof(double2);
    float2 *rel = reinterpret_cast<float2 *>(m);
    m += (size_t)n * sizeof(float2);
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
    const double inv_dr = 1.0 / dr;
    k_tile_bbox<<<nt, kTile, 0, c->stream>>>(n, sorted, bbox, inv_dr, f32 ? ctr : nullptr, f32 ? rel : nullptr, 2.0 * lmax);
    PcfArgs a;
    a.n = n; a.num_bins = num_bins;
    a.b = c->dbox;
    a.dr = dr; a.max_r = max_r;
    a.s_max = edmd_pcf_s_threshold(max_r);
    a.dr_lo = dr * (1.0 + 1.7763568394002505e-15);   // 2^-49
    a.dr_hi = dr * (1.0 - 1.7763568394002505e-15);
    a.inv_dr = (float)(1.0 / dr);
    a.sorted = sorted; a.bbox = bbox; a.counts = counts; a.stats = stats;
    a.part = part; a.nparts = nparts;
    const long long npairs = ((long long)nt * (nt + 1) / 2 + nparts - 1) / nparts;
    const long long sms = c->sm_count > 0 ? c->sm_count : 148;
    if (f32) {
        F32Args fa;
        a.use_smem = 1;
        fa.p = a;
        fa.rel = rel; fa.ctr = ctr;
        fa.inv_dr = inv_dr;
        fa.tmax = kMagic + (float)num_bins;
        fa.slack = 5.684341886080802e-14 * (lmax * inv_dr + 1.0);   // 2^-44 (L/dr + 1)
        fa.lx32 = (float)(c->box.lx * inv_dr);
        fa.ly32 = (float)(c->box.ly * inv_dr);
        fa.hx32 = 0.5f * fa.lx32;
        fa.hy32 = 0.5f * fa.ly32;
        // groups of 256 threads per CTA (they share the histogram): the G with the most resident warps
        static unsigned long long attr32 = 0;   // devices of this process the attributes are set on
        if (edmd_first_on_device(&attr32)) {
            cudaFuncSetAttribute(k_pcf_f32<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024);
            cudaFuncSetAttribute(k_pcf_f32<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024);
            cudaFuncSetAttribute(k_pcf_f32<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024);
        }
        int best_g = 1, best_warps = 0, best_per_sm = 1;
        for (int g = 1; g <= 4; g *= 2) {
            const size_t sm = g * kGroupSmem + f32_hist;
            if (sm > 208 * 1024) break;
            int per_sm = 0;
            if (g == 1) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pcf_f32<1>, kThreads, sm);
            else if (g == 2) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pcf_f32<2>, 2 * kThreads, sm);
            else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pcf_f32<4>, 4 * kThreads, sm);
            const int warps = per_sm * g * (kThreads / 32);
            const bool forced = c->pcf_groups == g && per_sm > 0;   // EDMD_OPT_PCF_GROUPS
            // ties go to the larger CTA (measured at N = 10^6, 32 warps each way: 274 / 267 / 256 ms for 1 / 2 / 4 groups;
            // N = 2*10^6: 1240 / 1117 / 1071 ms)
            if (warps >= best_warps || forced) {
                best_warps = forced ? 1 << 20 : warps;
                best_g = g;
                best_per_sm = per_sm;
            }
        }
        if (best_per_sm < 1) best_per_sm = 1;
        const size_t smem = best_g * kGroupSmem + f32_hist;
        long long grid = sms * best_per_sm;
        const long long vctas = (npairs + best_g - 1) / best_g;
        if (grid > vctas) grid = vctas;
        if (grid < 1) grid = 1;
        if (best_g == 1) k_pcf_f32<1><<<(int)grid, kThreads, smem, c->stream>>>(fa);
        else if (best_g == 2) k_pcf_f32<2><<<(int)grid, 2 * kThreads, smem, c->stream>>>(fa);
        else k_pcf_f32<4><<<(int)grid, 4 * kThreads, smem, c->stream>>>(fa);
        return launched;
    }
    const size_t tile_bytes = kTile * sizeof(double2);
    const size_t hist_bytes = (size_t)num_bins * sizeof(unsigned int);
    a.use_smem = (tile_bytes + hist_bytes) <= 200 * 1024;
    const size_t smem = tile_bytes + (a.use_smem ? hist_bytes : 0);
    static unsigned long long attr = 0;   // devices of this process the attributes are set on
    if (edmd_first_on_device(&attr)) {
        cudaFuncSetAttribute(k_pcf_sorted, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024);
    }
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pcf_sorted, kThreads, smem);
    if (per_sm < 1) per_sm = 1;
    long long grid = sms * per_sm;
    if (grid > npairs) grid = npairs;
    if (grid < 1) grid = 1;
    k_pcf_sorted<<<(int)grid, kThreads, smem, c->stream>>>(a);
    return launched;
}
