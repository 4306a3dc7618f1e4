// analysis_voronoi.cu -- K5: the Voronoi family of the reference's frame analysis
// (SURVEY.md 8 a12 / 8f rank 4):
//   computeBOOPVoronoi              src/boop.c:15-59          psi_5,6,7 over Voronoi neighbours
//   get_particle_voronoi_area       src/voronoi_edmd.c:123-135
//   get_particle_voronoi_perimeter  src/voronoi_edmd.c:137-149
//
// The reference builds ONE global diagram with Fortune's sweep (jc_voronoi.h) over
// the particles plus the periodic images within 6.0 of the box edges
// (get_particle_voronoi, src/voronoi_edmd.c:33-121) -- a sequential algorithm.  The
// Voronoi cell of one particle, however, is a local object: the intersection of the
// half-planes  { p : p.d_j <= |d_j|^2 / 2 }  over the offsets d_j to the other
// particles, and a particle farther than twice the cell's circumradius cannot cut
// it.  So here every particle builds its own cell (one thread per particle):
//   * positions are counting-sorted into a uniform grid of ~2 particles per cell
//     (cell width w = 1.5 mean spacings); the periodic images are the cells reached
//     by wrapping the grid index, with the shift +-L added to the candidate exactly
//     like the reference builds its image points (`x + Lx`, :68-110);
//   * the polygon (<= kMaxV vertices, relative to the particle, in shared memory)
//     starts as a large square and is clipped by the bisector of every candidate in
//     growing Chebyshev rings of grid cells; after ring k every unvisited particle is
//     at least `reach` >= k*w away, so the cell is final as soon as
//     2 R_max <= reach  (R_max = farthest vertex).  Candidates with |d| >= 2 R_max are
//     skipped without clipping.  The first-shell candidates are listed first and then
//     clipped round by round, so that the lanes of a warp clip at the same time;
//   * every edge remembers the candidate that made it: the edges of the final
//     polygon ARE the Voronoi neighbours (jcv_graphedge::neighbor).
// Cells of one grid cell are put in ascending particle id first, so the clipping
// order -- and with it every output bit -- is the same run after run.  The psi sums
// are accumulated like K4 (complex powers of the unit bond vector, 2^-48 fixed point).
// Degenerate inputs (four cocircular neighbours, e.g. an exact square lattice) have no
// unique diagram: the reference's answer there depends on its sweep's rounding, ours
// on the clipping's.
#include "edmd_internal.cuh"

namespace {

constexpr int kThreads = 128;
constexpr int kMaxV = 20;   // vertices of a cell under construction (Poisson points: cells of up to ~13 sides)
constexpr int kNear = 10;   // first-shell candidates clipped in lockstep

struct VorGrid {
    int n, gx, gy, kmax;
    double lx, ly;
    double fx, fy;        // grid cells per unit length
    double cmin;          // min(cell width, cell height)
    const double4 *xv;    // resident state
    int32_t *cnt;         // [gx*gy + 1] histogram -> cell starts -> (after scatter) cell ends
    int32_t *cell;        // [n]
    double2 *spos;        // [n] positions in cell order
    int32_t *sid;         // [n] particle id of spos[k]
};

__device__ __forceinline__ int vor_cell(const VorGrid &g, double x, double y)
{
    int cx = (int)(x * g.fx), cy = (int)(y * g.fy);
    cx = min(max(cx, 0), g.gx - 1);
    cy = min(max(cy, 0), g.gy - 1);
    return cy * g.gx + cx;
}

__global__ void __launch_bounds__(256)
k_vor_count(const __grid_constant__ VorGrid g)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.n) return;
    const double4 p = g.xv[i];
    const int c = vor_cell(g, p.x, p.y);
    g.cell[i] = c;
    atomicAdd(&g.cnt[c], 1);
}

// exclusive scan of cnt[0..m) in place (m ~ N/2 grid cells): block sums, scan of the
// sums by one block, block-local scan + offset.  (A single-block scan of the whole
// array took 462 us at N = 10^6.)
constexpr int kScanThreads = 256;
constexpr int kScanPer = 8;                               // consecutive items per thread
constexpr int kScanTile = kScanThreads * kScanPer;        // items per block

__device__ __forceinline__ int block_excl_scan(int x, int &total)
{
    __shared__ int s_w[kScanThreads / 32];
    int incl = x;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, incl, d);
        if ((threadIdx.x & 31) >= d) incl += o;
    }
    __syncthreads();   // s_w free
    if ((threadIdx.x & 31) == 31) s_w[threadIdx.x >> 5] = incl;
    __syncthreads();
    int wb = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < kScanThreads / 32; w++) {
        if (w < (int)(threadIdx.x >> 5)) wb += s_w[w];
        tot += s_w[w];
    }
    total = tot;
    return wb + incl - x;
}

__global__ void __launch_bounds__(kScanThreads)
k_scan_sums(int m, const int32_t *__restrict__ v, int32_t *__restrict__ sums)
{
    const int base = blockIdx.x * kScanTile + threadIdx.x * kScanPer;
    int acc = 0;
#pragma unroll
    for (int u = 0; u < kScanPer; u++) acc += base + u < m ? v[base + u] : 0;
    int tot;
    block_excl_scan(acc, tot);
    if (threadIdx.x == 0) sums[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(kScanThreads)
k_scan_top(int nb, int32_t *__restrict__ sums)   // one block: exclusive scan of the block sums
{
    __shared__ int s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < nb; base += kScanThreads) {
        const int i = base + threadIdx.x;
        const int x = i < nb ? sums[i] : 0;
        int tot;
        const int ex = block_excl_scan(x, tot);
        const int carry = s_carry;
        if (i < nb) sums[i] = carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) s_carry = carry + tot;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kScanThreads)
k_scan_apply(int m, int32_t *__restrict__ v, const int32_t *__restrict__ sums)
{
    const int base = blockIdx.x * kScanTile + threadIdx.x * kScanPer;
    int x[kScanPer], acc = 0;
#pragma unroll
    for (int u = 0; u < kScanPer; u++) {
        x[u] = base + u < m ? v[base + u] : 0;
        acc += x[u];
    }
    int tot;
    int run = block_excl_scan(acc, tot) + sums[blockIdx.x];
#pragma unroll
    for (int u = 0; u < kScanPer; u++) {
        if (base + u < m) v[base + u] = run;
        run += x[u];
    }
}

__global__ void __launch_bounds__(256)
k_vor_scatter(const __grid_constant__ VorGrid g)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.n) return;
    const int k = atomicAdd(&g.cnt[g.cell[i]], 1);   // cnt = running cursors; they end at the cell ends
    g.sid[k] = i;
}

// one thread per grid cell: ascending particle id inside the cell (insertion sort of a
// handful of entries), then the positions in that order
__global__ void __launch_bounds__(256)
k_vor_order(const __grid_constant__ VorGrid g)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= g.gx * g.gy) return;
    const int lo = c == 0 ? 0 : g.cnt[c - 1], hi = g.cnt[c];
    for (int k = lo + 1; k < hi; k++) {
        const int v = g.sid[k];
        int q = k - 1;
        while (q >= lo && g.sid[q] > v) {
            g.sid[q + 1] = g.sid[q];
            q--;
        }
        g.sid[q + 1] = v;
    }
    for (int k = lo; k < hi; k++) {
        const double4 p = g.xv[g.sid[k]];
        g.spos[k] = make_double2(p.x, p.y);
    }
}

struct VorArgs {
    VorGrid g;
    double *q5, *q6, *q7, *q6arg;   // by particle id (any may be null together: boop == 0)
    int32_t *nbr;
    double *area, *perim;           // nullable
    int boop;
    double near2;                   // candidates closer than this (squared) are clipped first, in lockstep
    int32_t *fail;                  // [0] cells not closed within kmax rings, [1] polygon overflow
};

constexpr double kFix = 281474976710656.0;        // 2^48
constexpr double kUnfix = 1.0 / 281474976710656.0;

// The polygon of one thread lives in shared memory, one column per thread
// ([vertex][thread]: conflict-free), and is clipped IN PLACE: a half-plane cuts a
// contiguous arc off a convex polygon, so the arc is replaced by the two new
// vertices and the tail moved.  (First version: two polygons per thread in local
// memory, copied by every clip -- 412 MB of DRAM write-back and 568 M
// warp-instructions at N = 10^6; see profiles/.)
struct Poly {
    double *x, *y;   // [kMaxV][kThreads] columns of this thread
    int *lab;
    int m;
    __device__ __forceinline__ double &X(int k) { return x[k * kThreads]; }
    __device__ __forceinline__ double &Y(int k) { return y[k * kThreads]; }
    __device__ __forceinline__ int &L(int k) { return lab[k * kThreads]; }
};

// Clip with the half-plane p.d <= hd (edge k = vertex k -> k+1 carries the label of the
// candidate that made it).  Returns false when the polygon would not fit.
__device__ __forceinline__ bool vor_clip(Poly &P, double dx, double dy, double hd, int label)
{
    const int m = P.m;
    int a = -1, b = -1, nout = 0;
    double sprev = P.X(m - 1) * dx + P.Y(m - 1) * dy - hd;
    for (int k = 0; k < m; k++) {
        const double sk = P.X(k) * dx + P.Y(k) * dy - hd;
        const bool ok = sk > 0.0, op = sprev > 0.0;
        if (!op && ok) a = k == 0 ? m - 1 : k - 1;   // vertex a inside, a+1 outside
        if (op && !ok) b = k == 0 ? m - 1 : k - 1;   // vertex b outside, b+1 inside
        nout += ok;
        sprev = sk;
    }
    if (nout == 0 || a < 0 || b < 0) return true;    // nothing to cut (the polygon always holds the origin)
    const int a1 = a + 1 == m ? 0 : a + 1, b1 = b + 1 == m ? 0 : b + 1;
    double i1x, i1y, i2x, i2y;
    {
        const double ax = P.X(a), ay = P.Y(a), cx = P.X(a1), cy = P.Y(a1);
        const double sa = ax * dx + ay * dy - hd, sc = cx * dx + cy * dy - hd;
        const double t = sa / (sa - sc);
        i1x = ax + t * (cx - ax); i1y = ay + t * (cy - ay);
    }
    {
        const double bx = P.X(b), by = P.Y(b), cx = P.X(b1), cy = P.Y(b1);
        const double sb = bx * dx + by * dy - hd, sc = cx * dx + cy * dy - hd;
        const double t = sb / (sb - sc);
        i2x = bx + t * (cx - bx); i2y = by + t * (cy - by);
    }
    const int labb = P.L(b);
    if (a < b) {   // the cut arc a+1 .. b does not wrap
        const int shift = 2 - nout;
        if (shift == 1) {
            if (m + 1 > kMaxV) return false;
            for (int k = m - 1; k > b; k--) { P.X(k + 1) = P.X(k); P.Y(k + 1) = P.Y(k); P.L(k + 1) = P.L(k); }
        } else if (shift < 0) {
            for (int k = b + 1; k < m; k++) { P.X(k + shift) = P.X(k); P.Y(k + shift) = P.Y(k); P.L(k + shift) = P.L(k); }
        }
        P.X(a + 1) = i1x; P.Y(a + 1) = i1y; P.L(a + 1) = label;      // the new edge runs along the bisector
        P.X(a + 2) = i2x; P.Y(a + 2) = i2y; P.L(a + 2) = labb;       // the rest of old edge b
        P.m = m + shift;
    } else {       // the cut arc wraps through vertex 0: keep b+1 .. a
        const int cnt = a - b;
        if (cnt + 2 > kMaxV) return false;
        if (b + 1 > 0)
            for (int k = 0; k < cnt; k++) { P.X(k) = P.X(k + b + 1); P.Y(k) = P.Y(k + b + 1); P.L(k) = P.L(k + b + 1); }
        P.X(cnt) = i1x; P.Y(cnt) = i1y; P.L(cnt) = label;
        P.X(cnt + 1) = i2x; P.Y(cnt + 1) = i2y; P.L(cnt + 1) = labb;
        P.m = cnt + 2;
    }
    return true;
}

__device__ __forceinline__ double poly_r2(Poly &P)
{
    double r2 = 0.0;
    for (int v = 0; v < P.m; v++) r2 = fmax(r2, P.X(v) * P.X(v) + P.Y(v) * P.Y(v));
    return r2;
}

constexpr size_t kVorSmem = (size_t)kThreads * (kMaxV * (2 * sizeof(double) + sizeof(int)) + kNear * sizeof(int));

__global__ void __launch_bounds__(kThreads)
k_voronoi(const __grid_constant__ VorArgs a)
{
    extern __shared__ __align__(16) unsigned char vor_smem[];
    const VorGrid &g = a.g;
    const int s = blockIdx.x * blockDim.x + threadIdx.x;   // slot in cell order: neighbours in memory are neighbours in space
    if (s >= g.n) return;
    Poly P;
    P.x = reinterpret_cast<double *>(vor_smem) + threadIdx.x;
    P.y = P.x + kMaxV * kThreads;
    P.lab = reinterpret_cast<int *>(vor_smem + 2 * sizeof(double) * kMaxV * kThreads) + threadIdx.x;
    int *near = P.lab + kMaxV * kThreads;   // [kNear][kThreads]
    const int id = g.sid[s];
    const double2 p = g.spos[s];
    const int c0 = vor_cell(g, p.x, p.y);
    const int cx0 = c0 % g.gx, cy0 = c0 / g.gx;
    {
        const double H = 0.5 * fmax(g.lx, g.ly);
        P.X(0) = -H; P.Y(0) = -H; P.X(1) = H; P.Y(1) = -H;
        P.X(2) = H; P.Y(2) = H; P.X(3) = -H; P.Y(3) = H;
        P.L(0) = P.L(1) = P.L(2) = P.L(3) = -1;
        P.m = 4;
    }
    double R2 = 0.5 * fmax(g.lx, g.ly) * fmax(g.lx, g.ly);
    bool overflow = false;

    // candidate of a packed label: image position first, then the difference -- `points[].x = x + Lx`
    // (src/voronoi_edmd.c:68-110), `dx = e->neighbor->p.x - particles[index].x` (src/boop.c:31)
    auto offset_of = [&](int label, double &dx, double &dy) {
        const int q = label / 9, sh = label - q * 9;
        const int sx = sh / 3, sy = sh - sx * 3;
        const double2 pj = g.spos[q];
        dx = __dsub_rn(__dadd_rn(pj.x, sx == 0 ? -g.lx : (sx == 2 ? g.lx : 0.0)), p.x);
        dy = __dsub_rn(__dadd_rn(pj.y, sy == 0 ? -g.ly : (sy == 2 ? g.ly : 0.0)), p.y);
    };
    // the shell of Chebyshev ring k; f(label, dx, dy, d2) per candidate
    auto ring = [&](int k, auto &&f) {
        for (int oy = -k; oy <= k && !overflow; oy++) {
            const bool edge_row = oy == -k || oy == k;
            int cy = cy0 + oy, sy = 1;
            if (cy < 0) { cy += g.gy; sy = 0; }
            else if (cy >= g.gy) { cy -= g.gy; sy = 2; }
            const double shy = sy == 0 ? -g.ly : (sy == 2 ? g.ly : 0.0);
            for (int ox = -k; ox <= k && !overflow; ox += (edge_row || k == 0) ? 1 : 2 * k) {
                int cx = cx0 + ox, sx = 1;
                if (cx < 0) { cx += g.gx; sx = 0; }
                else if (cx >= g.gx) { cx -= g.gx; sx = 2; }
                const double shx = sx == 0 ? -g.lx : (sx == 2 ? g.lx : 0.0);
                const int c = cy * g.gx + cx;
                const int lo = c == 0 ? 0 : g.cnt[c - 1], hi = g.cnt[c];
                for (int q = lo; q < hi && !overflow; q++) {
                    if (q == s && sx == 1 && sy == 1) continue;
                    const double2 pj = g.spos[q];
                    const double dx = __dsub_rn(__dadd_rn(pj.x, shx), p.x);
                    const double dy = __dsub_rn(__dadd_rn(pj.y, shy), p.y);
                    f(q * 9 + sx * 3 + sy, dx, dy, dx * dx + dy * dy);
                }
            }
        }
    };
    auto clip_if_near = [&](int label, double dx, double dy, double d2) {
        if (0.25 * d2 >= R2) return;   // |d| >= 2 R_max: cannot cut
        if (!vor_clip(P, dx, dy, 0.5 * d2, label)) {
            overflow = true;
            return;
        }
        R2 = poly_r2(P);
    };

    // 1. the nearest candidates of rings 0 and 1 (the first shell in a dense system) are only
    //    LISTED ...
    int nn = 0;
    for (int k = 0; k <= 1 && k <= g.kmax; k++)
        ring(k, [&](int label, double, double, double d2) {
            if (d2 < a.near2 && nn < kNear) near[(nn++) * kThreads] = label;
        });
    // 2. ... and clipped round by round, so that the lanes of a warp clip together
    for (int t = 0; t < nn && !overflow; t++) {
        double dx, dy;
        const int label = near[t * kThreads];
        offset_of(label, dx, dy);
        if (!vor_clip(P, dx, dy, 0.5 * (dx * dx + dy * dy), label)) overflow = true;
    }
    R2 = poly_r2(P);
    // 3. the rest of rings 0 and 1 (almost all of it too far to cut), then further rings until the
    //    cell is closed: after ring k every unvisited particle is at least `reach` away
    bool closed = false;
    const double oxc = p.x - cx0 * (g.lx / g.gx), oyc = p.y - cy0 * (g.ly / g.gy);
    const double cwx = g.lx / g.gx, cwy = g.ly / g.gy;
    const double inx = fmax(0.0, fmin(oxc, cwx - oxc)), iny = fmax(0.0, fmin(oyc, cwy - oyc));
    int seen = 0;
    for (int k = 0; k <= g.kmax && !closed && !overflow; k++) {
        if (k <= 1)
            ring(k, [&](int label, double dx, double dy, double d2) {
                if (d2 < a.near2 && seen < kNear) {   // listed in step 1 (same order, same counter)
                    seen++;
                    return;
                }
                clip_if_near(label, dx, dy, d2);
            });
        else
            ring(k, clip_if_near);
        const double reach = fmin((double)k * cwx + inx, (double)k * cwy + iny);
        closed = k >= 1 && 4.0 * R2 <= reach * reach;
    }
    if (overflow) atomicAdd(&a.fail[1], 1);
    else if (!closed) atomicAdd(&a.fail[0], 1);

    long long s5r = 0, s5i = 0, s6r = 0, s6i = 0, s7r = 0, s7i = 0;
    int nb = 0;
    double area2 = 0.0, per = 0.0;
    for (int k = 0; k < P.m; k++) {
        const int kn = k + 1 == P.m ? 0 : k + 1;
        const double ex = P.X(kn) - P.X(k), ey = P.Y(kn) - P.Y(k);
        area2 += P.X(k) * P.Y(kn) - P.Y(k) * P.X(kn);
        per += sqrt(ex * ex + ey * ey);
        const int label = P.L(k);
        if (label < 0 || (ex == 0.0 && ey == 0.0)) continue;
        nb++;
        if (!a.boop) continue;
        double dx, dy;
        offset_of(label, dx, dy);
        const double r2 = dx * dx + dy * dy;
        double zr = 1.0, zi = 0.0;   // atan2(0,0) = 0
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
        s5r += __double2ll_rn(z5r * kFix); s5i += __double2ll_rn(z5i * kFix);
        s6r += __double2ll_rn(z6r * kFix); s6i += __double2ll_rn(z6i * kFix);
        s7r += __double2ll_rn(z7r * kFix); s7i += __double2ll_rn(z7i * kFix);
    }
    if (a.nbr) a.nbr[id] = nb;
    if (a.area) a.area[id] = 0.5 * fabs(area2);
    if (a.perim) a.perim[id] = per;
    if (a.boop) {
        if (nb > 0) {
            const double dn = (double)nb;
            const double r6 = (double)s6r * kUnfix, i6 = (double)s6i * kUnfix;
            a.q5[id] = hypot((double)s5r * kUnfix, (double)s5i * kUnfix) / dn;
            a.q6[id] = hypot(r6, i6) / dn;
            a.q7[id] = hypot((double)s7r * kUnfix, (double)s7i * kUnfix) / dn;
            a.q6arg[id] = atan2(i6, r6);
        } else {
            a.q5[id] = a.q6[id] = a.q7[id] = a.q6arg[id] = 0.0;
        }
    }
}

// psi6_i = q6_i e^{i arg_i} as the reference forms it (src/pcf.c:183-186)
__global__ void __launch_bounds__(256)
k_psi6_from_boop(int n, const double *__restrict__ q6, const double *__restrict__ arg, double2 *__restrict__ psi)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s, c;
    sincos(arg[i], &s, &c);
    psi[i] = make_double2(q6[i] * c, q6[i] * s);
}

}  // namespace

size_t edmd_voronoi_scratch_bytes(const edmd_ctx *c, int *gx_out, int *gy_out)
{
    const int n = c->n;
    const double w = 1.5 * sqrt(c->box.lx * c->box.ly / (double)(n > 0 ? n : 1));
    int gx = (int)(c->box.lx / w), gy = (int)(c->box.ly / w);
    if (gx < 1) gx = 1;
    if (gy < 1) gy = 1;
    if (gx > 8192) gx = 8192;
    if (gy > 8192) gy = 8192;
    *gx_out = gx;
    *gy_out = gy;
    const size_t ncell = (size_t)gx * gy + 8;
    return (size_t)n * (sizeof(double2) + 2 * sizeof(int32_t)) + (ncell + ncell / kScanTile + 8) * sizeof(int32_t) + 64;
}

// Voronoi cell of every resident particle.  scratch: edmd_voronoi_scratch_bytes().  fail_dev[2] is
// zeroed here and holds the particles whose cell could not be closed / whose polygon overflowed.
int edmd_launch_voronoi(edmd_ctx *c, char *scratch, int boop, double *q5, double *q6, double *q7, double *q6arg,
                        int32_t *nbr, double *area, double *perim, int32_t *fail_dev, cudaEvent_t before_cells)
{
    const int n = c->n;
    int gx, gy;
    edmd_voronoi_scratch_bytes(c, &gx, &gy);
    VorGrid g;
    g.n = n; g.gx = gx; g.gy = gy;
    g.kmax = ((gx < gy ? gx : gy) - 1) / 2;
    g.lx = c->box.lx; g.ly = c->box.ly;
    g.fx = gx / c->box.lx; g.fy = gy / c->box.ly;
    const double cwx = c->box.lx / gx, cwy = c->box.ly / gy;
    g.cmin = cwx < cwy ? cwx : cwy;
    g.xv = c->xv;
    char *m = scratch;
    g.spos = reinterpret_cast<double2 *>(m);
    m += (size_t)n * sizeof(double2);
    g.sid = reinterpret_cast<int32_t *>(m);
    m += (size_t)n * sizeof(int32_t);
    g.cell = reinterpret_cast<int32_t *>(m);
    m += (size_t)n * sizeof(int32_t);
    g.cnt = reinterpret_cast<int32_t *>(m);
    const int ncell = gx * gy;
    cudaMemsetAsync(g.cnt, 0, ((size_t)ncell + 8) * sizeof(int32_t), c->stream);
    cudaMemsetAsync(fail_dev, 0, 2 * sizeof(int32_t), c->stream);
    const int pb = (n + 255) / 256;
    k_vor_count<<<pb, 256, 0, c->stream>>>(g);
    {
        int32_t *sums = g.cnt + ncell + 8;
        const int m = ncell + 1, nb = (m + kScanTile - 1) / kScanTile;
        k_scan_sums<<<nb, kScanThreads, 0, c->stream>>>(m, g.cnt, sums);
        k_scan_top<<<1, kScanThreads, 0, c->stream>>>(nb, sums);
        k_scan_apply<<<nb, kScanThreads, 0, c->stream>>>(m, g.cnt, sums);
    }
    k_vor_scatter<<<pb, 256, 0, c->stream>>>(g);
    k_vor_order<<<(ncell + 255) / 256, 256, 0, c->stream>>>(g);
    VorArgs a;
    a.g = g;
    a.q5 = q5; a.q6 = q6; a.q7 = q7; a.q6arg = q6arg; a.nbr = nbr;
    a.area = area; a.perim = perim; a.boop = boop; a.fail = fail_dev;
    {
        // first shell of a dense system: 1.45 mean spacings (triangular lattice: neighbours at 1.07, 1.86)
        const double sp = sqrt(c->box.lx * c->box.ly / (double)(n > 0 ? n : 1));
        a.near2 = (1.45 * sp) * (1.45 * sp);
    }
    static unsigned long long attr = 0;   // devices of this process the attributes are set on
    if (edmd_first_on_device(&attr)) {
        cudaFuncSetAttribute(k_voronoi, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kVorSmem);
    }
    if (before_cells) cudaEventRecord(before_cells, c->stream);
    k_voronoi<<<(n + kThreads - 1) / kThreads, kThreads, kVorSmem, c->stream>>>(a);
    return 7;
}

int edmd_launch_psi6(edmd_ctx *c, const double *q6, const double *arg, double2 *psi)
{
    k_psi6_from_boop<<<(c->n + 255) / 256, 256, 0, c->stream>>>(c->n, q6, arg, psi);
    return 1;
}
