// analysis.cu -- per-frame structure analysis.
//
//   K4 boop_cutoff : computeBOOPCutoff, src/boop.c:61-107
//   K3 pcf_hist    : calculate_pcf,     src/pcf.c:16-75
//
// Parity-critical arithmetic (the r2 < r_c^2 neighbour test, the pair
// distance and its bin) uses explicit unfused __d*_rn operations in the
// reference's order so that neighbour counts and histogram counts are
// integers identical to the reference's.
#include "edmd_internal.cuh"
#include "rowstage.cuh"

namespace {

__device__ __forceinline__ double min_image(double d, double half, double len)
{
    if (d >= half) return __dsub_rn(d, len);
    if (d < -half) return __dadd_rn(d, len);
    return d;
}


// ------------------------------------------------------------------ K4 ----
// Same per-warp row staging and 3x3 traversal as the sweep (and the same
// truncation the reference has: cells are ~2.0 wide, r_c = 2.5, so neighbours
// two cells away are never seen -- reproduced on purpose).
// e^{ik theta} = ((dx + i dy)/r)^k by complex powers instead of atan2 + cexp;
// agrees with libm to a few ulp (gate: 1e-10).  The six sums are accumulated
// in 2^-48 fixed point: exact integer addition makes them independent of the
// order in which a cell's particles are stored (error <= 8 * 2^-49 per sum).
struct BoopArgs {
    edmd_dev_box b;
    CellIndex g;
    int max_chunks;
    double rc2;
    double *q5, *q6, *q7, *q6arg;
    int32_t *nbr;
};

struct BoopAcc {
    long long s5r = 0, s5i = 0, s6r = 0, s6i = 0, s7r = 0, s7i = 0;
    int nb = 0;
};

constexpr double kFix = 281474976710656.0;        // 2^48
constexpr double kUnfix = 1.0 / 281474976710656.0;

template <bool WRAP>
__device__ __forceinline__ void boop_pair(const edmd_dev_box &b, double rc2, const SRec &p1,
                                          const SRec &p2, BoopAcc &acc)
{
    double dx = __dsub_rn(p2.x, p1.x);
    double dy = __dsub_rn(p2.y, p1.y);
    if (WRAP) {
        dx = min_image(dx, b.half_lx, b.lx);
        dy = min_image(dy, b.half_ly, b.ly);
    }
    const double r2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
    if (r2 < rc2) {
        acc.nb++;
        double zr = 1.0, zi = 0.0;  // atan2(0,0) = 0 in the reference
        if (r2 > 0) {
            const double inv = rsqrt(r2);
            zr = dx * inv;
            zi = dy * inv;
        }
        const double z2r = zr * zr - zi * zi, z2i = 2.0 * zr * zi;
        const double z4r = z2r * z2r - z2i * z2i, z4i = 2.0 * z2r * z2i;
        const double z5r = z4r * zr - z4i * zi, z5i = z4r * zi + z4i * zr;
        const double z6r = z4r * z2r - z4i * z2i, z6i = z4r * z2i + z4i * z2r;
        const double z7r = z6r * zr - z6i * zi, z7i = z6r * zi + z6i * zr;
        acc.s5r += __double2ll_rn(z5r * kFix); acc.s5i += __double2ll_rn(z5i * kFix);
        acc.s6r += __double2ll_rn(z6r * kFix); acc.s6i += __double2ll_rn(z6i * kFix);
        acc.s7r += __double2ll_rn(z7r * kFix); acc.s7i += __double2ll_rn(z7i * kFix);
    }
}

__device__ __forceinline__ void boop_emit(const BoopArgs &a, int id, const BoopAcc &acc)
{
    a.nbr[id] = acc.nb;
    if (acc.nb > 0) {
        const double dn = (double)acc.nb;
        const double s6r = (double)acc.s6r * kUnfix, s6i = (double)acc.s6i * kUnfix;
        a.q5[id] = hypot((double)acc.s5r * kUnfix, (double)acc.s5i * kUnfix) / dn;
        a.q6[id] = hypot(s6r, s6i) / dn;
        a.q7[id] = hypot((double)acc.s7r * kUnfix, (double)acc.s7i * kUnfix) / dn;
        a.q6arg[id] = atan2(s6i, s6r);
    } else {
        a.q5[id] = 0.0;
        a.q6[id] = 0.0;
        a.q7[id] = 0.0;
        a.q6arg[id] = 0.0;
    }
}

template <bool WRAP>
__device__ __forceinline__ void boop_range(const edmd_dev_box &b, double rc2, const SRec &p1,
                                           const SPos *pos, const SAux *aux, int lo, int hi, BoopAcc &acc)
{
#pragma unroll 1
    for (int p = lo; p < hi; p++) {
        if (aux[p].id == p1.id) continue;  // `p2->num != p1->num`
        const SPos q = pos[p];
        SRec p2;
        p2.x = q.x; p2.y = q.y;
        boop_pair<WRAP>(b, rc2, p1, p2, acc);
    }
}

__global__ void __launch_bounds__(kStageThreads, kStageCtasPerSm)
k_boop_rows(const __grid_constant__ BoopArgs a)
{
    const bool sane = a.g.flags[kFlagInsane] == 0;
    row_pipeline(a.g, a.g.meta, a.max_chunks,
                 [&](const StageBuf &buf, const ChunkMeta &m, const RowLane &rl, int status) {
                     if (!rl.active) return;
                     BoopAcc acc;
                     if (status == 2) {  // segments do not fit: straight from global memory
                         const SRec p1 = make_rec(a.g.spos[rl.s], a.g.saux[rl.s]);
                         if (p1.id >= a.g.n_owned) return;
#pragma unroll
                         for (int j = 0; j < 3; j++)
                             boop_range<true>(a.b, a.rc2, p1, a.g.spos, a.g.saux, rl.lo[j], rl.hi[j], acc);
                         boop_emit(a, p1.id, acc);
                         return;
                     }
                     const bool fast = sane && (m.flags & kMetaInterior);
                     const SRec p1 = make_rec(buf.pos[1][rl.self], buf.aux[1][rl.self]);
                     if (p1.id >= a.g.n_owned) return;
#pragma unroll
                     for (int j = 0; j < 3; j++) {
                         if (fast) boop_range<false>(a.b, a.rc2, p1, buf.pos[j], buf.aux[j], rl.lo[j], rl.hi[j], acc);
                         else boop_range<true>(a.b, a.rc2, p1, buf.pos[j], buf.aux[j], rl.lo[j], rl.hi[j], acc);
                     }
                     boop_emit(a, p1.id, acc);
                 });
}

// deterministic two-stage sum: fixed block partials, then one block in order
constexpr int kRedThreads = 256;

__global__ void __launch_bounds__(kRedThreads)
k_sum_partial(int n, const double *__restrict__ v, double *__restrict__ partial)
{
    __shared__ double sh[kRedThreads];
    double acc = 0;
    for (int i = blockIdx.x * kRedThreads + threadIdx.x; i < n;
         i += gridDim.x * kRedThreads)
        acc += v[i];
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int d = kRedThreads / 2; d > 0; d >>= 1) {
        if (threadIdx.x < d) sh[threadIdx.x] += sh[threadIdx.x + d];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}

__global__ void __launch_bounds__(kRedThreads)
k_sum_final(int m, const double *__restrict__ partial, double scale,
            double *__restrict__ out)
{
    __shared__ double sh[kRedThreads];
    double acc = 0;
    for (int i = threadIdx.x; i < m; i += kRedThreads) acc += partial[i];
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int d = kRedThreads / 2; d > 0; d >>= 1) {
        if (threadIdx.x < d) sh[threadIdx.x] += sh[threadIdx.x + d];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = sh[0] * scale;
}

// ------------------------------------------------------------------ K3 ----
// All unordered pairs i < j (src/pcf.c:39-54).  Particles are cut into tiles
// of kTile; a CTA walks tile pairs (a <= b) from a persistent work list, keeps
// tile b in shared memory (broadcast reads) and one particle of tile a per
// thread in registers, and bins into a per-CTA shared u32 histogram that is
// merged into the global u64 histogram once at the end.  FP64-pipe bound
// (positions fit in L2; HBM traffic is negligible).
constexpr int kPcfThreads = 256;
constexpr int kTile = 256;

__global__ void __launch_bounds__(kPcfThreads)
k_pcf(int n, edmd_dev_box b, double bin_width, double max_r, int num_bins,
      const double *__restrict__ xy, int stride, int part, int nparts,
      unsigned long long *__restrict__ counts, int use_smem_hist)
{
    extern __shared__ unsigned char smem_raw[];
    double2 *tile = reinterpret_cast<double2 *>(smem_raw);
    unsigned int *hist = reinterpret_cast<unsigned int *>(smem_raw + kTile * sizeof(double2));
    if (use_smem_hist) {
        for (int k = threadIdx.x; k < num_bins; k += kPcfThreads) hist[k] = 0;
    }
    const int nt = (n + kTile - 1) / kTile;
    const long long npairs = (long long)nt * (nt + 1) / 2;
    for (long long w = (long long)blockIdx.x * nparts + part; w < npairs; w += (long long)gridDim.x * nparts) {
        // unrank w -> (ta, tb) with ta <= tb, row-major over the upper triangle
        // row ta starts at ta*nt - ta*(ta-1)/2
        double fn = (double)nt + 0.5;
        long long ta = (long long)(fn - sqrt(fn * fn - 2.0 * (double)w));
        while (ta * nt - ta * (ta - 1) / 2 > w) ta--;
        while ((ta + 1) * nt - (ta + 1) * ta / 2 <= w) ta++;
        long long tb = ta + (w - (ta * nt - ta * (ta - 1) / 2));
        __syncthreads();
        {
            int jj = (int)tb * kTile + threadIdx.x;
            if (jj < n) tile[threadIdx.x] = *reinterpret_cast<const double2 *>(xy + (size_t)jj * stride);
        }
        __syncthreads();
        const int i = (int)ta * kTile + threadIdx.x;
        if (i < n) {
            const double2 pi = *reinterpret_cast<const double2 *>(xy + (size_t)i * stride);
            const int jbase = (int)tb * kTile;
            const int jcount = min(kTile, n - jbase);
            const int jstart = (ta == tb) ? threadIdx.x + 1 : 0;
            for (int jj = jstart; jj < jcount; jj++) {
                double2 pj = tile[jj];
                double dx = min_image(__dsub_rn(pj.x, pi.x), b.half_lx, b.lx);
                double dy = min_image(__dsub_rn(pj.y, pi.y), b.half_ly, b.ly);
                double r = __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
                if (r < max_r) {
                    int bin = (int)__ddiv_rn(r, bin_width);
                    if (bin < num_bins) {
                        if (use_smem_hist) atomicAdd(&hist[bin], 1u);
                        else atomicAdd(&counts[bin], 1ull);
                    }
                }
            }
        }
    }
    if (use_smem_hist) {
        __syncthreads();
        for (int k = threadIdx.x; k < num_bins; k += kPcfThreads) {
            unsigned int v = hist[k];
            if (v) atomicAdd(&counts[k], (unsigned long long)v);
        }
    }
}

}  // namespace

int edmd_launch_boop(edmd_ctx *c, double r_c)
{
    int n = c->n;
    if (n == 0) return 0;
    size_t N = (size_t)c->n_cap;   // the four psi6 arrays lie n_cap apart
    BoopArgs a;
    a.b = c->dbox;
    a.g = edmd_cell_index(c);
    a.max_chunks = edmd_chunks_bound(c);
    a.rc2 = r_c * r_c;  // `r_c*r_c`, a single rounded product
    a.q5 = c->boop;
    a.q6 = c->boop + N;
    a.q7 = c->boop + 2 * N;
    a.q6arg = c->boop + 3 * N;
    a.nbr = c->boop_nb;
    const int blocks = (a.max_chunks + kStageWarps - 1) / kStageWarps;
    static unsigned long long attr = 0;   // devices of this process the attributes are set on
    if (edmd_first_on_device(&attr)) {
        cudaFuncSetAttribute(k_boop_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kStageSmem);
        cudaFuncSetAttribute(k_boop_rows, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    }
    k_boop_rows<<<min(blocks, edmd_persistent_blocks(c)), kStageThreads, kStageSmem, c->stream>>>(a);
    return 1;
}

int edmd_launch_mean(edmd_ctx *c, const double *v, int n, double *out_dev)
{
    int blocks = c->red_cap;
    k_sum_partial<<<blocks, kRedThreads, 0, c->stream>>>(n, v, c->red_partial);
    k_sum_final<<<1, kRedThreads, 0, c->stream>>>(blocks, c->red_partial,
                                                  n > 0 ? 1.0 / (double)n : 0.0,
                                                  out_dev);
    return 2;
}

// xy: positions as (x, y) pairs every `stride` doubles (the resident xv array: stride 4);
// this launch handles tile pairs w = part (mod nparts) and ADDS into counts.
int edmd_launch_pcf(edmd_ctx *c, double dr, double max_r, int num_bins, const double *xy, int stride,
                    int n, int part, int nparts, unsigned long long *counts)
{
    if (n < 2 || num_bins <= 0) return 0;
    // large systems: spatially sorted tiles + certified bins (analysis_pcf_sorted.cu); for
    // the multi-GPU split the sort is made deterministic so that every rank cuts the same tiles
    if (n >= 8192 && c->pcf_mode != 1) {
        const int r = edmd_launch_pcf_sorted(c, dr, max_r, num_bins, xy, stride, n, part, nparts, counts);
        if (r >= 0) return r;
    }
    size_t tile_bytes = kTile * sizeof(double2);
    size_t hist_bytes = (size_t)num_bins * sizeof(unsigned int);
    int use_smem = (tile_bytes + hist_bytes) <= 200 * 1024;
    size_t smem = tile_bytes + (use_smem ? hist_bytes : 0);
    static unsigned long long attr_set = 0;   // devices of this process the attributes are set on
    if (edmd_first_on_device(&attr_set)) {
        cudaFuncSetAttribute(k_pcf, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             227 * 1024);
    }
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pcf, kPcfThreads, smem);
    if (per_sm < 1) per_sm = 1;
    long long nt = (n + kTile - 1) / kTile;
    long long npairs = (nt * (nt + 1) / 2 + nparts - 1) / nparts;
    long long grid = (long long)(c->sm_count > 0 ? c->sm_count : 148) * per_sm;
    if (grid > npairs) grid = npairs;
    if (grid < 1) grid = 1;
    k_pcf<<<(int)grid, kPcfThreads, smem, c->stream>>>(n, c->dbox, dr, max_r, num_bins, xy, stride, part,
                                                       nparts, counts, use_smem);
    return 1;
}
