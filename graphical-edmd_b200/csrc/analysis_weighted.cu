// analysis_weighted.cu -- the weighted g(r) family of the reference's frame
// analysis:
//   calculate_bond_order_pcf          src/pcf.c:77-167    g(r) and <cos(k.r)>(r)
//   find_max_structure_factor_bragg   src/pcf.c:405-467   argmax_k S(k) in a wedge
//   compute_g6_correlation (pair loop) src/pcf.c:189-228   <Re psi6_i^* psi6_j>(r)
//   computeStructureFactor / computeVelocityStructureFactor  src/struc.c:364-408
// Same pair tiling as K3 (analysis.cu).  The squared pair distance uses the
// reference's unfused FP64 operations and its bin is certified in FP64 (the reference's
// sqrt and division only where the certificate fails: integer counts identical to
// calculate_pcf's); the cosine of a pair comes from per-particle phases e^{i k.r} (two
// multiply-adds per pair instead of a cos); the weights are accumulated in 2^-32 fixed point (exact
// integer addition: the result does not depend on the order of the atomics;
// quantisation 1.2e-10 per pair, ~1e-14 on the per-bin average), so the sums
// are reproducible run to run -- unlike the reference's OpenMP merge order.
#include "edmd_internal.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kTile = 256;
constexpr double kWFix = 4294967296.0;   // 2^32

struct WPcfArgs {
    int n, num_bins, stride, use_smem;
    edmd_dev_box b;
    double bin_width, max_r;
    double s_max;          // r < max_r  <=>  s < s_max  (s = fl(fl(dx*dx) + fl(dy*dy)), see s_threshold)
    double dr_lo, dr_hi;   // dr (1 + 2^-49) and dr (1 - 2^-49)
    float inv_dr;
    double2 ph[9];         // COS weights: e^{i k.S} of the nine image shifts S; index 3 ix + iy, 1 = no shift
    const double *xy;
    const double2 *psi;           // per-particle weights: psi6 (PSI) or e^{i k.r} (COS)
    unsigned long long *counts;   // [num_bins] unordered pairs
    unsigned long long *wsum;     // [num_bins] sum of the weights * 2^32 (two's complement)
};

// weight of a pair: cos(k_vector . d) (calculate_bond_order_pcf) or Re(conj(psi_i) psi_j)
// (compute_g6_correlation)
enum { kWeightCos = 0, kWeightPsi = 1 };

// psi_i = e^{i k.r_i}: the cosine of a pair is then two multiply-adds, cos(k.d) = Re(conj(psi_i) psi_j e^{i k.S})
// with S the periodic image shift the pair took (d = r_j - r_i + S) -- one FP64 sincos per PARTICLE instead of
// one cos per PAIR (the reference's `cos(k_vector[0]*dx + k_vector[1]*dy)`, src/pcf.c:123; the phases reach
// k L ~ 10^4, so the two forms agree to ~10^-12: gate 1e-10).
__global__ void __launch_bounds__(kThreads)
k_pair_phase(int n, const double *__restrict__ xy, int stride, double kx, double ky, double2 *__restrict__ psi)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double2 p = *reinterpret_cast<const double2 *>(xy + (size_t)i * stride);
    double sn, cs;
    sincos(__dadd_rn(__dmul_rn(kx, p.x), __dmul_rn(ky, p.y)), &sn, &cs);
    psi[i] = make_double2(cs, sn);
}

// All unordered pairs i < j; the reference walks ordered pairs, which doubles
// both sums (cos is even, d_ji = -d_ij) and leaves their ratio unchanged.
// Per pair: s = fl(fl(dx*dx) + fl(dy*dy)) with the reference's unfused operations, the range test in s-space,
// an FP32 estimate k of the bin and its FP64 certificate (k dr)^2 (1 + 2^-48) <= s < ((k+1) dr)^2 (1 - 2^-48)
// -- inside that interval the reference's sqrt-then-divide truncates to k (analysis_pcf_sorted.cu, pair_bin) --,
// the reference's own sqrt and division for the few pairs that fail it: integer counts identical to
// calculate_pcf's.  The weights are accumulated in 2^-32 fixed point, the 64-bit sums of a CTA as two 32-bit
// shared-memory words with the carry taken from the returned low word (a 64-bit shared-memory atomic is a
// compare-and-swap loop).
// A CTA is kGroups groups of kThreads threads that share ONE histogram (at N = 10^6, dr = 0.1 it takes 118 KB:
// one CTA per SM; with a single group that is 8 warps per SM and the kernel waits on its own FP64 chains and
// on the returned atomics -- measured 5.9 s for 5*10^11 pairs).  Every group walks its own tile pairs with its
// own tile and named barrier; four pairs per trip, branch-free up to the atomics, so that the chains interleave.
constexpr int kGroups = 4;

template <int WEIGHT>
__global__ void __launch_bounds__(kThreads * kGroups)
k_pcf_bond_order(const __grid_constant__ WPcfArgs a)
{
    extern __shared__ unsigned char smem_raw[];
    const int group = threadIdx.x / kThreads, tid = threadIdx.x % kThreads;
    double2 *tile = reinterpret_cast<double2 *>(smem_raw) + 2 * kTile * group;
    double2 *ptile = tile + kTile;
    unsigned int *hlo = reinterpret_cast<unsigned int *>(smem_raw + kGroups * 2 * kTile * sizeof(double2));
    unsigned int *hhi = hlo + a.num_bins;
    unsigned int *hc = hhi + a.num_bins;
    if (a.use_smem)
        for (int k = threadIdx.x; k < a.num_bins; k += kThreads * kGroups) {
            hlo[k] = 0;
            hhi[k] = 0;
            hc[k] = 0;
        }
    __syncthreads();
    auto group_barrier = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "r"(kThreads) : "memory"); };
    const double half_lx = a.b.half_lx, half_ly = a.b.half_ly, lx = a.b.lx, ly = a.b.ly;
    const double s_max = a.s_max, dr_lo = a.dr_lo, dr_hi = a.dr_hi;
    const float inv_dr = a.inv_dr;
    const int num_bins = a.num_bins;
    const int nt = (a.n + kTile - 1) / kTile;
    const long long npairs = (long long)nt * (nt + 1) / 2;
    // one pair's weight into bin `bin` (fixed point, see the header comment)
    auto add = [&](int bin, double cw) {
        const unsigned long long q = (unsigned long long)__double2ll_rn(cw);
        if (a.use_smem) {
            atomicAdd(&hc[bin], 1u);
            const unsigned int lo = (unsigned int)q;
            const unsigned int old = atomicAdd(&hlo[bin], lo);
            const unsigned int hi = (unsigned int)(q >> 32) + ((old + lo) < old ? 1u : 0u);
            if (hi) atomicAdd(&hhi[bin], hi);
        } else {
            atomicAdd(&a.counts[bin], 1ull);
            atomicAdd(&a.wsum[bin], q);
        }
    };
    for (long long w = (long long)blockIdx.x * kGroups + group; w < npairs; w += (long long)gridDim.x * kGroups) {
        // unrank w -> (ta, tb), ta <= tb, row-major over the upper triangle
        const double fn = (double)nt + 0.5;
        long long ta = (long long)(fn - sqrt(fn * fn - 2.0 * (double)w));
        while (ta * nt - ta * (ta - 1) / 2 > w) ta--;
        while ((ta + 1) * nt - (ta + 1) * ta / 2 <= w) ta++;
        const long long tb = ta + (w - (ta * nt - ta * (ta - 1) / 2));
        group_barrier();
        {
            const int jj = (int)tb * kTile + tid;
            if (jj < a.n) {
                tile[tid] = *reinterpret_cast<const double2 *>(a.xy + (size_t)jj * a.stride);
                ptile[tid] = a.psi[jj];
            }
        }
        group_barrier();
        const int i = (int)ta * kTile + tid;
        if (i >= a.n) continue;
        const double2 pi = *reinterpret_cast<const double2 *>(a.xy + (size_t)i * a.stride);
        // the own weight, times 2^32 (exact): the pair's weight comes out in fixed-point units
        const double2 s0 = a.psi[i];
        const double2 si = make_double2(__dmul_rn(s0.x, kWFix), __dmul_rn(s0.y, kWFix));
        const int jbase = (int)tb * kTile;
        const int jcount = min(kTile, a.n - jbase);
        const int jstart = (ta == tb) ? tid + 1 : 0;
        // s, the weight and the certified bin of pair (i, jj): take = in range, certified, a real bin;
        // redo = in range, not certified
        auto pair = [&](int jj, double &s, double &cw, int &bin, bool &take, bool &redo) {
            const double2 pj = tile[jj];
            // PBC(), src/EDMD.c:5896-5913: `if (d >= half) d -= L; else if (d < -half) d += L;`
            const double dxr = __dsub_rn(pj.x, pi.x), dyr = __dsub_rn(pj.y, pi.y);
            const int ix = dxr >= half_lx ? 0 : (dxr < -half_lx ? 2 : 1);
            const int iy = dyr >= half_ly ? 0 : (dyr < -half_ly ? 2 : 1);
            const double dx = __dadd_rn(dxr, ix == 1 ? 0.0 : (ix == 0 ? -lx : lx));   // (d + 0.0 = d: the same bits)
            const double dy = __dadd_rn(dyr, iy == 1 ? 0.0 : (iy == 0 ? -ly : ly));
            s = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
            // FP32 estimate of r / dr from the bits of s; garbage outside the float range -- the certificate
            // rejects it
            const float sf = __int_as_float(((__double2hiint(s) - 0x38000000) << 3) |
                                            (int)((unsigned)__double2loint(s) >> 29));
            float rs;
            asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(rs) : "f"(sf));
            bin = __float2int_rz(sf * rs * inv_dr);
            const double kd = __dsub_rn(__hiloint2double(0x43300000, bin), 4503599627370496.0);
            const double e0 = __dmul_rn(kd, dr_lo), e1 = __dmul_rn(__dadd_rn(kd, 1.0), dr_hi);
            const bool in_range = s < s_max;   // `if (r < max_r)`
            const bool ok = (s >= __dmul_rn(e0, e0)) && (s < __dmul_rn(e1, e1));
            take = in_range && ok && bin >= 0 && bin < num_bins;
            redo = in_range && !ok;
            const double2 sj = ptile[jj];
            // `creal(conj(psi6[i]) * psi6[j])` src/pcf.c:204 (the reference's unfused operations)
            cw = __dadd_rn(__dmul_rn(si.x, sj.x), __dmul_rn(si.y, sj.y));
            if (WEIGHT == kWeightCos) {   // times e^{i k.S} of the image shift (no shift: (1, 0), the same bits)
                const double2 P = a.ph[3 * ix + iy];
                const double im = __dsub_rn(__dmul_rn(si.x, sj.y), __dmul_rn(si.y, sj.x));
                cw = __dsub_rn(__dmul_rn(cw, P.x), __dmul_rn(im, P.y));
            }
        };
        auto exact = [&](double s, double cw) {   // the reference's own sqrt and division
            const int bin = (int)__ddiv_rn(__dsqrt_rn(s), a.bin_width);   // `(int)(r / data->bin_width)`
            if (bin >= 0 && bin < num_bins) add(bin, cw);
        };
        int jj = jstart;
        for (; jj + 4 <= jcount; jj += 4) {
            double s[4], cw[4];
            int bin[4];
            bool take[4], redo[4];
#pragma unroll
            for (int u = 0; u < 4; u++) pair(jj + u, s[u], cw[u], bin[u], take[u], redo[u]);
#pragma unroll
            for (int u = 0; u < 4; u++)
                if (take[u]) add(bin[u], cw[u]);
            if (redo[0] | redo[1] | redo[2] | redo[3]) {
#pragma unroll
                for (int u = 0; u < 4; u++)
                    if (redo[u]) exact(s[u], cw[u]);
            }
        }
        for (; jj < jcount; jj++) {
            double s, cw;
            int bin;
            bool take, redo;
            pair(jj, s, cw, bin, take, redo);
            if (take) add(bin, cw);
            else if (redo) exact(s, cw);
        }
    }
    if (a.use_smem) {
        __syncthreads();
        for (int k = threadIdx.x; k < a.num_bins; k += kThreads * kGroups) {
            if (hc[k]) {
                atomicAdd(&a.counts[k], (unsigned long long)hc[k]);
                atomicAdd(&a.wsum[k], ((unsigned long long)hhi[k] << 32) + hlo[k]);
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

// phase: weights cos(k.d) (calculate_bond_order_pcf) -- psi is scratch for e^{i k.r}, filled here --; else
// Re(conj(psi_i) psi_j) of the caller's psi (compute_g6_correlation)
int edmd_launch_pcf_bond_order(edmd_ctx *c, double dr, double max_r, int num_bins, double kx, double ky,
                               double2 *psi, bool phase, unsigned long long *counts, unsigned long long *wsum)
{
    const int n = c->n;
    if (n < 2 || num_bins <= 0) return 0;
    int launches = 0;
    WPcfArgs a;
    a.n = n; a.num_bins = num_bins; a.stride = 4;
    a.b = c->dbox;
    a.bin_width = dr; a.max_r = max_r;
    a.s_max = edmd_pcf_s_threshold(max_r);
    a.dr_lo = dr * (1.0 + 1.7763568394002505e-15);   // 2^-49
    a.dr_hi = dr * (1.0 - 1.7763568394002505e-15);
    a.inv_dr = (float)(1.0 / dr);
    a.xy = reinterpret_cast<const double *>(c->xv);
    a.counts = counts; a.wsum = wsum; a.psi = psi;
    for (int ix = 0; ix < 3; ix++)
        for (int iy = 0; iy < 3; iy++) {
            // the shift PBC() applied: ix = 0: d -= Lx, 2: d += Lx (likewise iy)
            const double sx = ix == 0 ? -c->box.lx : (ix == 2 ? c->box.lx : 0.0);
            const double sy = iy == 0 ? -c->box.ly : (iy == 2 ? c->box.ly : 0.0);
            const double th = kx * sx + ky * sy;
            a.ph[3 * ix + iy] = make_double2(cos(th), sin(th));
        }
    if (phase) {
        k_pair_phase<<<(n + kThreads - 1) / kThreads, kThreads, 0, c->stream>>>(n, a.xy, a.stride, kx, ky, psi);
        launches++;
    }
    auto kern = phase ? k_pcf_bond_order<kWeightCos> : k_pcf_bond_order<kWeightPsi>;
    const size_t tile_bytes = kGroups * 2 * kTile * sizeof(double2);
    const size_t hist_bytes = (size_t)num_bins * 3 * sizeof(unsigned int);
    a.use_smem = (tile_bytes + hist_bytes) <= 200 * 1024;
    const size_t smem = tile_bytes + (a.use_smem ? hist_bytes : 0);
    static unsigned long long attr = 0;   // devices of this process the attributes are set on
    if (edmd_first_on_device(&attr)) {
        cudaFuncSetAttribute(k_pcf_bond_order<kWeightCos>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        cudaFuncSetAttribute(k_pcf_bond_order<kWeightPsi>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    }
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads * kGroups, smem);
    if (per_sm < 1) per_sm = 1;
    const long long nt = (n + kTile - 1) / kTile;
    long long grid = (long long)(c->sm_count > 0 ? c->sm_count : 148) * per_sm;
    const long long need = (nt * (nt + 1) / 2 + kGroups - 1) / kGroups;
    if (grid > need) grid = need;
    kern<<<(int)grid, kThreads * kGroups, smem, c->stream>>>(a);
    return launches + 1;
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
