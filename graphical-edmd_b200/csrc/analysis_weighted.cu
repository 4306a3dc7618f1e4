// analysis_weighted.cu -- the weighted g(r) family of the reference's frame
// analysis:
//   calculate_bond_order_pcf          src/pcf.c:77-167    g(r) and <cos(k.r)>(r)
//   find_max_structure_factor_bragg   src/pcf.c:405-467   argmax_k S(k) in a wedge
//   compute_g6_correlation (pair loop) src/pcf.c:189-228   <Re psi6_i^* psi6_j>(r)
//   computeStructureFactor / computeVelocityStructureFactor  src/struc.c:364-408
// Same pair tiling as K3 (analysis.cu).  The pair distance and its bin use the
// reference's unfused FP64 operations (integer counts identical to
// calculate_pcf's); the weights are accumulated in 2^-32 fixed point (exact
// integer addition: the result does not depend on the order of the atomics;
// quantisation 1.2e-10 per pair, ~1e-14 on the per-bin average), so the sums
// are reproducible run to run -- unlike the reference's OpenMP merge order.
#include "edmd_internal.cuh"

namespace {

__device__ __forceinline__ double min_image(double d, double half, double len)
{
    if (d >= half) return __dsub_rn(d, len);
    if (d < -half) return __dadd_rn(d, len);
    return d;
}

constexpr int kThreads = 256;
constexpr int kTile = 256;
constexpr double kWFix = 4294967296.0;   // 2^32

struct WPcfArgs {
    int n, num_bins, stride, use_smem;
    edmd_dev_box b;
    double bin_width, max_r, kx, ky;
    const double *xy;
    const double2 *psi;           // PSI weights: psi6 of every particle (re, im)
    unsigned long long *counts;   // [num_bins] unordered pairs
    unsigned long long *wsum;     // [num_bins] sum of the weights * 2^32 (two's complement)
};

// weight of a pair: cos(k_vector . d) (calculate_bond_order_pcf) or Re(conj(psi_i) psi_j)
// (compute_g6_correlation)
enum { kWeightCos = 0, kWeightPsi = 1 };

// All unordered pairs i < j; the reference walks ordered pairs, which doubles
// both sums (cos is even, d_ji = -d_ij) and leaves their ratio unchanged.
template <int WEIGHT>
__global__ void __launch_bounds__(kThreads)
k_pcf_bond_order(const __grid_constant__ WPcfArgs a)
{
    extern __shared__ unsigned char smem_raw[];
    double2 *tile = reinterpret_cast<double2 *>(smem_raw);
    double2 *ptile = tile + kTile;
    unsigned long long *hw = reinterpret_cast<unsigned long long *>(smem_raw + 2 * kTile * sizeof(double2));
    unsigned int *hc = reinterpret_cast<unsigned int *>(hw + a.num_bins);
    if (a.use_smem)
        for (int k = threadIdx.x; k < a.num_bins; k += kThreads) {
            hw[k] = 0;
            hc[k] = 0;
        }
    const int nt = (a.n + kTile - 1) / kTile;
    const long long npairs = (long long)nt * (nt + 1) / 2;
    for (long long w = blockIdx.x; w < npairs; w += gridDim.x) {
        // unrank w -> (ta, tb), ta <= tb, row-major over the upper triangle
        const double fn = (double)nt + 0.5;
        long long ta = (long long)(fn - sqrt(fn * fn - 2.0 * (double)w));
        while (ta * nt - ta * (ta - 1) / 2 > w) ta--;
        while ((ta + 1) * nt - (ta + 1) * ta / 2 <= w) ta++;
        const long long tb = ta + (w - (ta * nt - ta * (ta - 1) / 2));
        __syncthreads();
        {
            const int jj = (int)tb * kTile + threadIdx.x;
            if (jj < a.n) {
                tile[threadIdx.x] = *reinterpret_cast<const double2 *>(a.xy + (size_t)jj * a.stride);
                if (WEIGHT == kWeightPsi) ptile[threadIdx.x] = a.psi[jj];
            }
        }
        __syncthreads();
        const int i = (int)ta * kTile + threadIdx.x;
        if (i < a.n) {
            const double2 pi = *reinterpret_cast<const double2 *>(a.xy + (size_t)i * a.stride);
            const double2 si = WEIGHT == kWeightPsi ? a.psi[i] : make_double2(0.0, 0.0);
            const int jbase = (int)tb * kTile;
            const int jcount = min(kTile, a.n - jbase);
            const int jstart = (ta == tb) ? threadIdx.x + 1 : 0;
            for (int jj = jstart; jj < jcount; jj++) {
                const double2 pj = tile[jj];
                const double dx = min_image(__dsub_rn(pj.x, pi.x), a.b.half_lx, a.b.lx);
                const double dy = min_image(__dsub_rn(pj.y, pi.y), a.b.half_ly, a.b.ly);
                const double r = __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
                if (r < a.max_r) {
                    const int bin = (int)__ddiv_rn(r, a.bin_width);
                    if (bin < a.num_bins) {
                        // `cos(k_vector[0]*dx + k_vector[1]*dy)` src/pcf.c:123
                        // `creal(conj(psi6[i]) * psi6[j])` src/pcf.c:204
                        const double cw = WEIGHT == kWeightPsi
                                              ? __dadd_rn(__dmul_rn(si.x, ptile[jj].x), __dmul_rn(si.y, ptile[jj].y))
                                              : cos(__dadd_rn(__dmul_rn(a.kx, dx), __dmul_rn(a.ky, dy)));
                        const unsigned long long q = (unsigned long long)__double2ll_rn(cw * kWFix);
                        if (a.use_smem) {
                            atomicAdd(&hc[bin], 1u);
                            atomicAdd(&hw[bin], q);
                        } else {
                            atomicAdd(&a.counts[bin], 1ull);
                            atomicAdd(&a.wsum[bin], q);
                        }
                    }
                }
            }
        }
    }
    if (a.use_smem) {
        __syncthreads();
        for (int k = threadIdx.x; k < a.num_bins; k += kThreads) {
            if (hc[k]) {
                atomicAdd(&a.counts[k], (unsigned long long)hc[k]);
                atomicAdd(&a.wsum[k], hw[k]);
            }
        }
    }
}

// ---- Bragg peak: S(k) = |sum_j exp(i k.r_j)|^2 / N over a list of wave vectors ------
struct BraggArgs {
    int n, nk, stride, chunk;
    const double *xy;
    const double2 *kvec;   // [nk] in the reference's loop order
    double *re, *im;       // [nk] accumulated with atomics over particle chunks
};

// blockIdx.x: tile of kThreads wave vectors (one per thread); blockIdx.y: chunk of
// particles, staged through shared memory and broadcast
__global__ void __launch_bounds__(kThreads)
k_bragg_sums(const __grid_constant__ BraggArgs a)
{
    __shared__ double2 tile[kTile];
    const int ik = blockIdx.x * kThreads + threadIdx.x;
    const double2 kv = ik < a.nk ? a.kvec[ik] : make_double2(0, 0);
    double re = 0.0, im = 0.0;
    const int j0 = blockIdx.y * a.chunk, j1 = min(a.n, j0 + a.chunk);
    for (int base = j0; base < j1; base += kTile) {
        __syncthreads();
        const int jj = base + threadIdx.x;
        if (jj < j1) tile[threadIdx.x] = *reinterpret_cast<const double2 *>(a.xy + (size_t)jj * a.stride);
        __syncthreads();
        const int cnt = min(kTile, j1 - base);
        for (int q = 0; q < cnt; q++) {
            // `phase = kx * x + ky * y; re += cos(phase); im += sin(phase);` src/pcf.c:442-446
            const double phase = __dadd_rn(__dmul_rn(kv.x, tile[q].x), __dmul_rn(kv.y, tile[q].y));
            double s, c;
            sincos(phase, &s, &c);
            re += c;
            im += s;
        }
    }
    if (ik < a.nk) {
        atomicAdd(&a.re[ik], re);
        atomicAdd(&a.im[ik], im);
    }
}

// Structure factor on the reference's wave-vector grid: the same sums with the
// weights of computeStructureFactor (1) or computeVelocityStructureFactor
// (`re += vx cos + vy sin; im += vx sin - vy cos`, src/struc.c:395-396).  Wave
// vector ik = i * nqy + j  ->  (qx[i], qy[j]).
struct SqArgs {
    int n, nqx, nqy, chunk;
    const double4 *xv;
    const double *qx, *qy;
    double *re, *im;   // [nqx * nqy]
};

template <bool VEL>
__global__ void __launch_bounds__(kThreads)
k_sq_sums(const __grid_constant__ SqArgs a)
{
    __shared__ double4 tile[kTile];
    const int nk = a.nqx * a.nqy;
    const int ik = blockIdx.x * kThreads + threadIdx.x;
    const double kx = ik < nk ? a.qx[ik / a.nqy] : 0.0, ky = ik < nk ? a.qy[ik % a.nqy] : 0.0;
    double re = 0.0, im = 0.0;
    const int j0 = blockIdx.y * a.chunk, j1 = min(a.n, j0 + a.chunk);
    for (int base = j0; base < j1; base += kTile) {
        __syncthreads();
        const int jj = base + threadIdx.x;
        if (jj < j1) tile[threadIdx.x] = a.xv[jj];
        __syncthreads();
        const int cnt = min(kTile, j1 - base);
        for (int q = 0; q < cnt; q++) {
            const double4 p = tile[q];
            // `qr = qx[i]*p->x + qy[j]*p->y` src/struc.c:373
            const double qr = __dadd_rn(__dmul_rn(kx, p.x), __dmul_rn(ky, p.y));
            double s, c;
            sincos(qr, &s, &c);
            if (VEL) {
                re += p.z * c + p.w * s;
                im += p.z * s - p.w * c;
            } else {
                re += c;
                im += s;
            }
        }
    }
    if (ik < nk) {
        atomicAdd(&a.re[ik], re);
        atomicAdd(&a.im[ik], im);
    }
}

// one block: S per wave vector, the first maximum in list order wins
__global__ void __launch_bounds__(1024)
k_bragg_argmax(int n, int nk, const double *__restrict__ re, const double *__restrict__ im,
               double *__restrict__ best_s, int *__restrict__ best_i)
{
    __shared__ double ss[1024];
    __shared__ int si[1024];
    double bs = -1.0;
    int bi = -1;
    for (int k = threadIdx.x; k < nk; k += 1024) {
        const double S = (re[k] * re[k] + im[k] * im[k]) / (double)n;
        if (S > bs) {   // ascending k within a thread: the earlier index stays on ties
            bs = S;
            bi = k;
        }
    }
    ss[threadIdx.x] = bs;
    si[threadIdx.x] = bi;
    __syncthreads();
    for (int d = 512; d > 0; d >>= 1) {
        if (threadIdx.x < d) {
            const double os = ss[threadIdx.x + d];
            const int oi = si[threadIdx.x + d];
            if (oi >= 0 && (os > ss[threadIdx.x] || (os == ss[threadIdx.x] && (si[threadIdx.x] < 0 || oi < si[threadIdx.x])))) {
                ss[threadIdx.x] = os;
                si[threadIdx.x] = oi;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        *best_s = ss[0];
        *best_i = si[0];
    }
}

}  // namespace

// psi == nullptr: weights cos(k.d) (calculate_bond_order_pcf); else Re(conj(psi_i) psi_j)
// (compute_g6_correlation)
int edmd_launch_pcf_bond_order(edmd_ctx *c, double dr, double max_r, int num_bins, double kx, double ky,
                               const double2 *psi, unsigned long long *counts, unsigned long long *wsum)
{
    const int n = c->n;
    if (n < 2 || num_bins <= 0) return 0;
    WPcfArgs a;
    a.n = n; a.num_bins = num_bins; a.stride = 4;
    a.b = c->dbox;
    a.bin_width = dr; a.max_r = max_r; a.kx = kx; a.ky = ky;
    a.xy = reinterpret_cast<const double *>(c->xv);
    a.counts = counts; a.wsum = wsum; a.psi = psi;
    auto kern = psi ? k_pcf_bond_order<kWeightPsi> : k_pcf_bond_order<kWeightCos>;
    const size_t tile_bytes = 2 * kTile * sizeof(double2);
    const size_t hist_bytes = (size_t)num_bins * (sizeof(unsigned long long) + sizeof(unsigned int));
    a.use_smem = (tile_bytes + hist_bytes) <= 200 * 1024;
    const size_t smem = tile_bytes + (a.use_smem ? hist_bytes : 0);
    static unsigned long long attr = 0;   // devices of this process the attributes are set on
    if (edmd_first_on_device(&attr)) {
        cudaFuncSetAttribute(k_pcf_bond_order<kWeightCos>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        cudaFuncSetAttribute(k_pcf_bond_order<kWeightPsi>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    }
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, smem);
    if (per_sm < 1) per_sm = 1;
    const long long nt = (n + kTile - 1) / kTile;
    long long grid = (long long)(c->sm_count > 0 ? c->sm_count : 148) * per_sm;
    if (grid > nt * (nt + 1) / 2) grid = nt * (nt + 1) / 2;
    kern<<<(int)grid, kThreads, smem, c->stream>>>(a);
    return 1;
}

// kvec / re / im / out live in caller-provided device scratch
int edmd_launch_bragg(edmd_ctx *c, int nk, const double2 *kvec, double *re, double *im, double *best_s,
                      int *best_i)
{
    const int n = c->n;
    if (n < 1 || nk < 1) return 0;
    BraggArgs a;
    a.n = n; a.nk = nk; a.stride = 4;
    a.xy = reinterpret_cast<const double *>(c->xv);
    a.kvec = kvec; a.re = re; a.im = im;
    const int kb = (nk + kThreads - 1) / kThreads;
    // enough CTAs to fill the GPU: split the particles when there are few wave vectors
    int parts = (2 * (c->sm_count > 0 ? c->sm_count : 148) + kb - 1) / kb;
    const int max_parts = (n + kTile - 1) / kTile;
    if (parts > max_parts) parts = max_parts;
    if (parts < 1) parts = 1;
    a.chunk = (((n + parts - 1) / parts) + kTile - 1) / kTile * kTile;
    parts = (n + a.chunk - 1) / a.chunk;
    k_bragg_sums<<<dim3(kb, parts), kThreads, 0, c->stream>>>(a);
    k_bragg_argmax<<<1, 1024, 0, c->stream>>>(n, nk, re, im, best_s, best_i);
    return 2;
}

// S(q) sums on the grid qx[nqx] x qy[nqy] (device arrays); re / im [nqx*nqy] zeroed by the caller
int edmd_launch_structure_factor(edmd_ctx *c, int velocity, int nqx, int nqy, const double *qx, const double *qy,
                                 double *re, double *im)
{
    const int n = c->n, nk = nqx * nqy;
    if (n < 1 || nk < 1) return 0;
    SqArgs a;
    a.n = n; a.nqx = nqx; a.nqy = nqy;
    a.xv = c->xv; a.qx = qx; a.qy = qy; a.re = re; a.im = im;
    const int kb = (nk + kThreads - 1) / kThreads;
    int parts = (2 * (c->sm_count > 0 ? c->sm_count : 148) + kb - 1) / kb;
    const int max_parts = (n + kTile - 1) / kTile;
    if (parts > max_parts) parts = max_parts;
    if (parts < 1) parts = 1;
    a.chunk = (((n + parts - 1) / parts) + kTile - 1) / kTile * kTile;
    parts = (n + a.chunk - 1) / a.chunk;
    if (velocity) k_sq_sums<true><<<dim3(kb, parts), kThreads, 0, c->stream>>>(a);
    else k_sq_sums<false><<<dim3(kb, parts), kThreads, 0, c->stream>>>(a);
    return 1;
}
