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

// ---- sums over the particles on a GRID of wave vectors ------------------------------------
//   find_max_structure_factor_bragg   src/pcf.c:405-467   S(k) = |sum_p e^{i k.r_p}|^2 / N, k = (kx_i, ky_j)
//   computeStructureFactor            src/struc.c:364-384  the same on (qx[i], qy[j])
//   computeVelocityStructureFactor    src/struc.c:386-408  `re += vx cos + vy sin; im += vx sin - vy cos`
// The wave vectors are a product grid, so e^{i k.r} = e^{i kx x} e^{i ky y}: per block of 32 x 32 wave vectors and
// 32 particles, 2 x 32 x 32 FP64 sincos build two tables (the weight (vx - i vy) of the velocity form folded into
// the y table) and every (wave vector, particle) term is then one complex multiply-add -- four fused
// multiply-adds -- instead of a sincos (the reference's `phase = kx*x + ky*y; cos(phase); sin(phase)`; the two
// forms agree to a few 1e-16 per term, gate 1e-10 on S).  A thread holds one column and four rows of the block:
// the x phasor is read once per particle (a 128-bit load per lane), the y phasors are warp-wide broadcasts.
constexpr int kGT = 32;    // wave vectors per block side
constexpr int kGP = 32;    // particles per table
constexpr int kGR = kGT * kGT / kThreads;   // rows per thread

struct GridArgs {
    int n, ni, nj, chunk;
    long long si, sj;            // output index of wave vector (i, j) = i * si + j * sj
    const double4 *xv;
    const double *qi, *qj;       // the two axes: wave vector (i, j) = (qi[i], qj[j])
    const unsigned char *mask;   // by output index, nullptr: every wave vector wanted
    double *re, *im;             // by output index, accumulated with atomics over particle chunks
};

// blockIdx.x, .y: block of wave vectors; blockIdx.z: chunk of particles
template <bool VEL>
__global__ void __launch_bounds__(kThreads)
k_grid_sums(const __grid_constant__ GridArgs a)
{
    __shared__ double2 A[kGP][kGT], B[kGP][kGT];
    __shared__ double4 pos[kGP];
    const int tid = threadIdx.x, ti = tid & 31, w = tid >> 5;
    const int i = blockIdx.x * kGT + ti, jb = blockIdx.y * kGT + kGR * w;
    bool want[kGR];
    bool any = false;
#pragma unroll
    for (int r = 0; r < kGR; r++) {
        const int j = jb + r;
        want[r] = i < a.ni && j < a.nj && (!a.mask || a.mask[i * a.si + j * a.sj] != 0);
        any = any || want[r];
    }
    if (!__syncthreads_or(any)) return;   // (a block outside the wedge of the Bragg search)
    // the table entries this thread builds: column `col` of A (tid < 128: wave vectors i) or of B (j), particles
    // prow, prow + 4, ...
    const int col = tid & 31, tab = (tid >> 5) & 1, prow = tid >> 6;
    const int qidx = tab == 0 ? blockIdx.x * kGT + col : blockIdx.y * kGT + col;
    const double q = tab == 0 ? (qidx < a.ni ? a.qi[qidx] : 0.0) : (qidx < a.nj ? a.qj[qidx] : 0.0);
    double re[kGR], im[kGR];
#pragma unroll
    for (int r = 0; r < kGR; r++) re[r] = im[r] = 0.0;
    const int p0 = blockIdx.z * a.chunk, p1 = min(a.n, p0 + a.chunk);
    for (int base = p0; base < p1; base += kGP) {
        __syncthreads();
        if (tid < kGP) pos[tid] = base + tid < p1 ? a.xv[base + tid] : make_double4(0, 0, 0, 0);
        __syncthreads();
        const int cnt = min(kGP, p1 - base);
#pragma unroll
        for (int m = 0; m < kGP / 4; m++) {
            const int p = prow + 4 * m;
            const double4 P = pos[p];
            double sn, cs;
            sincos(__dmul_rn(q, tab == 0 ? P.x : P.y), &sn, &cs);
            if (tab == 0) {
                A[p][col] = make_double2(cs, sn);
            } else if (VEL) {   // (vx - i vy) e^{i qy y}
                B[p][col] = make_double2(__fma_rn(P.z, cs, __dmul_rn(P.w, sn)), __fma_rn(P.z, sn, -__dmul_rn(P.w, cs)));
            } else {
                B[p][col] = make_double2(cs, sn);
            }
        }
        __syncthreads();
        for (int p = 0; p < cnt; p++) {
            const double2 x = A[p][ti];
#pragma unroll
            for (int r = 0; r < kGR; r++) {
                const double2 y = B[p][kGR * w + r];
                re[r] = __fma_rn(x.x, y.x, re[r]);
                re[r] = __fma_rn(-x.y, y.y, re[r]);
                im[r] = __fma_rn(x.x, y.y, im[r]);
                im[r] = __fma_rn(x.y, y.x, im[r]);
            }
        }
    }
#pragma unroll
    for (int r = 0; r < kGR; r++)
        if (want[r]) {
            const long long idx = i * a.si + (jb + r) * a.sj;
            atomicAdd(&a.re[idx], re[r]);
            atomicAdd(&a.im[idx], im[r]);
        }
}

// one block: S per wave vector (mask: only the marked ones), the first maximum in index order wins
__global__ void __launch_bounds__(1024)
k_bragg_argmax(int n, long long nk, const double *__restrict__ re, const double *__restrict__ im,
               const unsigned char *__restrict__ mask, double *__restrict__ best_s, long long *__restrict__ best_i)
{
    __shared__ double ss[1024];
    __shared__ long long si[1024];
    double bs = -1.0;
    long long bi = -1;
    for (long long k = threadIdx.x; k < nk; k += 1024) {
        if (mask && !mask[k]) continue;
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
            const long long oi = si[threadIdx.x + d];
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

// sums on the grid (qi[i], qj[j]) (device arrays); re / im by output index i * si + j * sj, zeroed by the caller
static int launch_grid_sums(edmd_ctx *c, bool velocity, int ni, int nj, long long si, long long sj, const double *qi,
                            const double *qj, const unsigned char *mask, double *re, double *im)
{
    const int n = c->n;
    GridArgs a;
    a.n = n; a.ni = ni; a.nj = nj; a.si = si; a.sj = sj;
    a.xv = c->xv; a.qi = qi; a.qj = qj; a.mask = mask; a.re = re; a.im = im;
    const int bi = (ni + kGT - 1) / kGT, bj = (nj + kGT - 1) / kGT;
    // enough CTAs to fill the GPU: split the particles when there are few wave vectors
    const long long blocks = (long long)bi * bj;
    long long parts = (8LL * (c->sm_count > 0 ? c->sm_count : 148) + blocks - 1) / blocks;
    const long long max_parts = (n + kGP - 1) / kGP;
    if (parts > max_parts) parts = max_parts;
    if (parts > 65535) parts = 65535;
    if (parts < 1) parts = 1;
    a.chunk = (int)((((n + parts - 1) / parts) + kGP - 1) / kGP * kGP);
    parts = (n + a.chunk - 1) / a.chunk;
    const dim3 grid(bi, bj, (unsigned)parts);
    if (velocity) k_grid_sums<true><<<grid, kThreads, 0, c->stream>>>(a);
    else k_grid_sums<false><<<grid, kThreads, 0, c->stream>>>(a);
    return 1;
}

// Bragg search on the grid (kx[ikx], ky[iky]); wave vector (ikx, iky) has index iky * nkx + ikx (the reference's
// loop order), mask marks the ones inside its wedge; re / im zeroed by the caller
int edmd_launch_bragg(edmd_ctx *c, int nkx, int nky, const double *kx, const double *ky, const unsigned char *mask,
                      double *re, double *im, double *best_s, long long *best_i)
{
    const int n = c->n;
    if (n < 1 || nkx < 1 || nky < 1) return 0;
    launch_grid_sums(c, false, nkx, nky, 1, nkx, kx, ky, mask, re, im);
    k_bragg_argmax<<<1, 1024, 0, c->stream>>>(n, (long long)nkx * nky, re, im, mask, best_s, best_i);
    return 2;
}

// S(q) sums on the grid qx[nqx] x qy[nqy] (device arrays); re / im [nqx*nqy] zeroed by the caller
int edmd_launch_structure_factor(edmd_ctx *c, int velocity, int nqx, int nqy, const double *qx, const double *qy,
                                 double *re, double *im)
{
    if (c->n < 1 || nqx * nqy < 1) return 0;
    return launch_grid_sums(c, velocity != 0, nqx, nqy, nqy, 1, qx, qy, nullptr, re, im);
}
