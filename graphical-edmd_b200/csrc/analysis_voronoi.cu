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
//   * the polygon (<= kMaxV vertices, relative to the particle) starts as a large
//     square and is clipped by the bisector of every candidate in growing Chebyshev
//     rings of grid cells; after ring k every unvisited particle is at least k*w
//     away, so the cell is final as soon as  2 R_max <= k w  (R_max = farthest
//     vertex).  Candidates with |d| >= 2 R_max are skipped without clipping;
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
constexpr int kMaxV = 24;

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

// exclusive scan in place, one block (cells ~ N/2: a few hundred trips)
__global__ void __launch_bounds__(1024)
k_vor_scan(int m, int32_t *__restrict__ v)
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
    int32_t *fail;                  // [0] cells not closed within kmax rings, [1] polygon overflow
};

struct Poly {
    double x[kMaxV], y[kMaxV];
    int lab[kMaxV];
    int m;
};

constexpr double kFix = 281474976710656.0;        // 2^48
constexpr double kUnfix = 1.0 / 281474976710656.0;

// Clip the convex polygon `in` (counter-clockwise, edge k = vertex k -> k+1, made by
// candidate lab[k]) with the half-plane p.d <= hd; returns false when it does not fit.
__device__ __forceinline__ bool vor_clip(const Poly &in, Poly &out, double dx, double dy, double hd, int label)
{
    int m = 0;
    double ax = in.x[0], ay = in.y[0];
    double sa = ax * dx + ay * dy - hd;
    for (int k = 0; k < in.m; k++) {
        const int kn = k + 1 == in.m ? 0 : k + 1;
        const double bx = in.x[kn], by = in.y[kn];
        const double sb = bx * dx + by * dy - hd;
        if (sa <= 0.0) {
            if (m >= kMaxV) return false;
            out.x[m] = ax; out.y[m] = ay; out.lab[m] = in.lab[k];
            m++;
            if (sb > 0.0) {   // leaving: the new edge runs along the bisector
                if (m >= kMaxV) return false;
                const double t = sa / (sa - sb);
                out.x[m] = ax + t * (bx - ax); out.y[m] = ay + t * (by - ay); out.lab[m] = label;
                m++;
            }
        } else if (sb <= 0.0) {   // entering: the rest of old edge k
            if (m >= kMaxV) return false;
            const double t = sa / (sa - sb);
            out.x[m] = ax + t * (bx - ax); out.y[m] = ay + t * (by - ay); out.lab[m] = in.lab[k];
            m++;
        }
        ax = bx; ay = by; sa = sb;
    }
    out.m = m;
    return true;
}

__global__ void __launch_bounds__(kThreads)
k_voronoi(const __grid_constant__ VorArgs a)
{
    const VorGrid &g = a.g;
    const int s = blockIdx.x * blockDim.x + threadIdx.x;   // slot in cell order: neighbours in memory are neighbours in space
    if (s >= g.n) return;
    const int id = g.sid[s];
    const double2 p = g.spos[s];
    const int c0 = vor_cell(g, p.x, p.y);
    const int cx0 = c0 % g.gx, cy0 = c0 / g.gx;

    Poly P[2];
    int cur = 0;
    {
        const double H = 0.5 * fmax(g.lx, g.ly);
        P[0].x[0] = -H; P[0].y[0] = -H; P[0].x[1] = H; P[0].y[1] = -H;
        P[0].x[2] = H; P[0].y[2] = H; P[0].x[3] = -H; P[0].y[3] = H;
        P[0].lab[0] = P[0].lab[1] = P[0].lab[2] = P[0].lab[3] = -1;
        P[0].m = 4;
    }
    double R2 = 2.0 * (0.5 * fmax(g.lx, g.ly)) * (0.5 * fmax(g.lx, g.ly));
    bool closed = false, overflow = false;
    for (int k = 0; k <= g.kmax && !closed && !overflow; k++) {
        for (int oy = -k; oy <= k; oy++) {
            const bool edge_row = oy == -k || oy == k;
            int cy = cy0 + oy, sy = 1;
            if (cy < 0) { cy += g.gy; sy = 0; }
            else if (cy >= g.gy) { cy -= g.gy; sy = 2; }
            const double shy = sy == 0 ? -g.ly : (sy == 2 ? g.ly : 0.0);
            for (int ox = -k; ox <= k; ox += (edge_row || k == 0) ? 1 : 2 * k) {   // the ring's shell only
                int cx = cx0 + ox, sx = 1;
                if (cx < 0) { cx += g.gx; sx = 0; }
                else if (cx >= g.gx) { cx -= g.gx; sx = 2; }
                const double shx = sx == 0 ? -g.lx : (sx == 2 ? g.lx : 0.0);
                const int c = cy * g.gx + cx;
                const int lo = c == 0 ? 0 : g.cnt[c - 1], hi = g.cnt[c];
                for (int q = lo; q < hi; q++) {
                    if (q == s && sx == 1 && sy == 1) continue;
                    const double2 pj = g.spos[q];
                    // image position first, then the difference: `points[].x = x + Lx` (src/voronoi_edmd.c:68-110),
                    // `dx = e->neighbor->p.x - particles[index].x` (src/boop.c:31)
                    const double dx = __dsub_rn(__dadd_rn(pj.x, shx), p.x);
                    const double dy = __dsub_rn(__dadd_rn(pj.y, shy), p.y);
                    const double hd = 0.5 * (dx * dx + dy * dy);
                    if (0.5 * hd >= R2) continue;   // |d| >= 2 R_max: cannot cut
                    if (!vor_clip(P[cur], P[cur ^ 1], dx, dy, hd, q * 9 + sx * 3 + sy)) {
                        overflow = true;
                        break;
                    }
                    cur ^= 1;
                    double r2 = 0.0;
                    for (int v = 0; v < P[cur].m; v++) r2 = fmax(r2, P[cur].x[v] * P[cur].x[v] + P[cur].y[v] * P[cur].y[v]);
                    R2 = r2;
                }
                if (overflow) break;
            }
            if (overflow) break;
        }
        const double reach = (double)k * g.cmin;
        closed = 4.0 * R2 <= reach * reach;
    }
    if (overflow) atomicAdd(&a.fail[1], 1);
    else if (!closed) atomicAdd(&a.fail[0], 1);

    const Poly &F = P[cur];
    long long s5r = 0, s5i = 0, s6r = 0, s6i = 0, s7r = 0, s7i = 0;
    int nb = 0;
    double area2 = 0.0, per = 0.0;
    for (int k = 0; k < F.m; k++) {
        const int kn = k + 1 == F.m ? 0 : k + 1;
        const double ex = F.x[kn] - F.x[k], ey = F.y[kn] - F.y[k];
        area2 += F.x[k] * F.y[kn] - F.y[k] * F.x[kn];
        per += sqrt(ex * ex + ey * ey);
        if (F.lab[k] < 0 || (ex == 0.0 && ey == 0.0)) continue;
        nb++;
        if (!a.boop) continue;
        const int q = F.lab[k] / 9, sh = F.lab[k] - q * 9;
        const int sx = sh / 3, sy = sh - sx * 3;
        const double2 pj = g.spos[q];
        const double dx = __dsub_rn(__dadd_rn(pj.x, sx == 0 ? -g.lx : (sx == 2 ? g.lx : 0.0)), p.x);
        const double dy = __dsub_rn(__dadd_rn(pj.y, sy == 0 ? -g.ly : (sy == 2 ? g.ly : 0.0)), p.y);
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
    return (size_t)n * (sizeof(double2) + 2 * sizeof(int32_t)) + ((size_t)gx * gy + 8) * sizeof(int32_t) + 64;
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
    k_vor_scan<<<1, 1024, 0, c->stream>>>(ncell + 1, g.cnt);
    k_vor_scatter<<<pb, 256, 0, c->stream>>>(g);
    k_vor_order<<<(ncell + 255) / 256, 256, 0, c->stream>>>(g);
    VorArgs a;
    a.g = g;
    a.q5 = q5; a.q6 = q6; a.q7 = q7; a.q6arg = q6arg; a.nbr = nbr;
    a.area = area; a.perim = perim; a.boop = boop; a.fail = fail_dev;
    if (before_cells) cudaEventRecord(before_cells, c->stream);
    k_voronoi<<<(n + kThreads - 1) / kThreads, kThreads, 0, c->stream>>>(a);
    return 5;
}

int edmd_launch_psi6(edmd_ctx *c, const double *q6, const double *arg, double2 *psi)
{
    k_psi6_from_boop<<<(c->n + 255) / 256, 256, 0, c->stream>>>(c->n, q6, arg, psi);
    return 1;
}
