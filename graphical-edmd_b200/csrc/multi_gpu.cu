// multi_gpu.cu -- the C-ABI entry for SEVERAL GPUs in ONE process (last row; 8e).
//
// An edmd_mg owns one slab context per device: contiguous row slabs of the cell grid (row-major
// cell index Y*Nxcells + X, src/EDMD.c:2071; periodic in y like PBCcellY :2118-2124), each with a
// one-cell-row halo.  The caller sees the single-GPU interface -- whole-system host arrays in,
// whole-system host arrays out, indexed by particle id:
//
//   upload       the particles are dealt to the slabs by the cell row they are filed under (the
//                host's cell_xy, or coordToCell of the positions) and uploaded with their global ids;
//   predict_all  every slab runs the fused exchange + sweep (edmd_cuda_exchange_predict_device:
//                boundary rows stored straight into the neighbour's inbox -- peer access over NVLink
//                between devices, plain device memory between slabs of one device --, partition,
//                persistent sweep kernel); all devices are launched before any is waited for; the
//                outputs of the slabs are disjoint and are scattered to the caller's arrays by id.
//                No collective on the data path.
//   boop_cutoff  per slab (the halo rows supply the neighbours across the boundary); the mean q6 of
//                the thermo column (src/EDMD.c:5521-5536) is the sum of the slabs' sums over N.
//   pcf          the positions go to every device, device k bins the tile pairs w = k (mod ndev)
//                (edmd_cuda_pcf_device), the integer histograms are added.
//
// The two reductions (q6 sums, g(r) counts) end in HOST arrays, so they are done on the host as part of
// the gather: a device-side all-reduce would only add a step.  The multi-PROCESS form of the same
// decomposition (one rank per GPU, NCCL all-reduce / all-gather through torch.distributed) is
// graphical-edmd_b200/slab.py; both drive the same slab contexts.
//
// NORMAL mode only: the growth sweep is a once-per-run setup step (single-GPU context).
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <new>
#include <thread>
#include <vector>

#include "edmd_internal.cuh"

struct edmd_mg {
    int ndev, n;
    double lx, ly;
    edmd_box box;
    std::vector<int> devices;
    std::vector<edmd_ctx *> ctx;
    std::vector<int> row_lo, row_hi;     // owned rows of every slab
    std::vector<int> owner_of_row;       // [ny]
    int halo_cap;
    // host staging, one set per slab (pinned)
    struct Stage {
        double *x, *y, *vx, *vy, *rad, *t_cross, *t_coll, *q[4];
        int32_t *cells, *gid, *partner, *nb;
        uint8_t *dir;
        int n, cap;
    };
    std::vector<Stage> st;
    bool have_state;
    bool whole;   // ndev == 1: one ordinary whole-system context (a slab needs at least two rows of somebody else)
    double t;
    char err[512];
};

namespace {

int mg_fail(edmd_mg *m, int code, const char *what, const edmd_ctx *c = nullptr)
{
    if (c) snprintf(m->err, sizeof(m->err), "%s: %s", what, edmd_cuda_last_error(c));
    else snprintf(m->err, sizeof(m->err), "%s", what);
    return code;
}

template <typename T>
bool pin(T **p, size_t n)
{
    return cudaHostAlloc(reinterpret_cast<void **>(p), (n ? n : 1) * sizeof(T), cudaHostAllocPortable) == cudaSuccess;
}

// f(part, nparts) on `nparts` host threads (the caller's included).  The dealing of the particles to the slabs,
// the scatter of the slabs' outputs and the per-device uploads / fetches are memory- or PCIe-bound loops over the
// whole system: one thread would make the multi-GPU tick slower than the single-GPU one.
template <class F>
void parallel_parts(int nparts, F f)
{
    std::vector<std::thread> th;
    for (int t = 1; t < nparts; t++) th.emplace_back([&f, t, nparts] { f(t, nparts); });
    f(0, nparts);
    for (auto &x : th) x.join();
}

int host_threads()
{
    const unsigned hc = std::thread::hardware_concurrency();
    return hc < 2 ? 1 : (hc > 16 ? 16 : (int)hc);
}

// Scatter the slabs' per-particle outputs into whole-system arrays BY RANGES OF PARTICLE IDS: thread p handles
// the ids [lo, hi) of every slab whose `ok` is set (a slab lists its ids in ascending order: binary search for lo).
// write(stage, j, i): element j of the slab's staging is particle i.  Scattering slab by slab on one thread each
// had the threads write neighbouring elements of the same cache lines (ids are dealt to the slabs at random):
// 7-26 ms instead of ~1 ms at N = 2*10^6.
template <class W>
void scatter_by_id_ranges(const edmd_mg *m, const std::vector<char> &ok, W write)
{
    parallel_parts(host_threads(), [&](int p, int np) {
        const int lo = (int)((long long)m->n * p / np), hi = (int)((long long)m->n * (p + 1) / np);
        for (int k = 0; k < m->ndev; k++) {
            if (!ok[k]) continue;
            const edmd_mg::Stage &s = m->st[k];
            int a = 0, b = s.n;
            while (a < b) {   // first j with gid[j] >= lo
                const int mid = (a + b) / 2;
                if (s.gid[mid] < lo) a = mid + 1;
                else b = mid;
            }
            for (int j = a; j < s.n && s.gid[j] < hi; j++) write(s, j, s.gid[j]);
        }
    });
}

void free_stage(edmd_mg::Stage &s)
{
    void *ps[] = {s.x, s.y, s.vx, s.vy, s.rad, s.t_cross, s.t_coll, s.q[0], s.q[1], s.q[2], s.q[3],
                  s.cells, s.gid, s.partner, s.nb, s.dir};
    for (void *p : ps)
        if (p) cudaFreeHost(p);
    memset(&s, 0, sizeof(s));
}

}  // namespace

// halo.cu's connect for contexts of ONE process: the neighbours' inboxes are addressed directly
// (peer access between devices, the same memory between slabs of one device) -- no IPC handle
int edmd_halo_connect_direct(edmd_ctx *c, edmd_ctx *lower, edmd_ctx *upper)
{
    edmd_ctx *nb[2] = {lower, upper};
    for (int k = 0; k < 2; k++) {
        if (!nb[k]->halo_mem) return EDMD_ESTATE;
        if (nb[k]->device != c->device) {
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, c->device, nb[k]->device) != cudaSuccess || !can) return EDMD_ESTATE;
            if (cudaSetDevice(c->device) != cudaSuccess) return -1;
            const cudaError_t e = cudaDeviceEnablePeerAccess(nb[k]->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return -(int)e;
            cudaGetLastError();
        }
        c->peer_mem[k] = nb[k]->halo_mem;
        c->peer_opened[k] = false;
    }
    if (cudaSetDevice(c->device) != cudaSuccess) return -1;
    edmd_preload_exchange_kernels();   // the kernels of an exchange wait for each other: none may load lazily
    return 0;
}

extern "C" {

const char *edmd_cuda_mg_last_error(const edmd_mg *m) { return m ? m->err : "null handle"; }

void edmd_cuda_destroy_mg(edmd_mg *m)
{
    if (!m) return;
    for (auto &s : m->st) free_stage(s);
    for (edmd_ctx *c : m->ctx)
        if (c) edmd_cuda_destroy(c);
    delete m;
}

int edmd_cuda_create_mg(int ndev, const int *devices, int n, double lx, double ly, edmd_mg **out)
{
    if (!out || ndev < 1 || !devices || n < 0) return EDMD_EINVAL;
    *out = nullptr;
    edmd_mg *m = new (std::nothrow) edmd_mg();
    if (!m) return -2;
    m->ndev = ndev; m->n = n; m->lx = lx; m->ly = ly;
    m->have_state = false;
    m->err[0] = 0;
    m->devices.assign(devices, devices + ndev);
    m->ctx.assign(ndev, nullptr);
    m->st.resize(ndev);
    for (auto &s : m->st) memset(&s, 0, sizeof(s));
    // the global box, as boxConstantHelper derives it (src/EDMD.c:679-706)
    const int nx = (int)(lx / 2), ny = (int)(ly / 2);
    if (nx < 1 || ny < 3 * ndev || ny < 3) {
        delete m;
        return EDMD_EINVAL;   // the decomposition needs Nycells >= 3 per slab
    }
    m->whole = ndev == 1;
    if (m->whole) {
        const int rc1 = edmd_cuda_create(devices[0], n, lx, ly, &m->ctx[0]);
        if (rc1) {
            edmd_cuda_destroy_mg(m);
            return rc1;
        }
        edmd_cuda_get_box(m->ctx[0], &m->box);
        m->st[0].n = 0;
        *out = m;
        return 0;
    }
    m->owner_of_row.resize(ny);
    const double per_row = (double)n / ny;
    m->halo_cap = (int)(per_row * 1.5) + 1024;
    int rc = 0;
    for (int k = 0; k < ndev && !rc; k++) {
        const int lo = (int)(((long long)ny * k) / ndev), hi = (int)(((long long)ny * (k + 1)) / ndev);
        m->row_lo.push_back(lo);
        m->row_hi.push_back(hi);
        for (int y = lo; y < hi; y++) m->owner_of_row[y] = k;
        // capacity: the slab's share + 30 % (particles migrate between slabs during a run) + the halo
        const int cap = (int)(per_row * (hi - lo) * 1.3) + 2 * m->halo_cap + 4096;
        rc = edmd_cuda_create_slab(devices[k], cap, lx, ly, lo, hi, &m->ctx[k]);
        if (rc) break;
        unsigned char handle[64];
        rc = edmd_cuda_halo_export(m->ctx[k], m->halo_cap, handle);   // allocates the inboxes
        edmd_mg::Stage &s = m->st[k];
        s.cap = cap;
        bool ok = pin(&s.x, cap) && pin(&s.y, cap) && pin(&s.vx, cap) && pin(&s.vy, cap) && pin(&s.rad, cap) &&
                  pin(&s.t_cross, cap) && pin(&s.t_coll, cap) && pin(&s.cells, 2 * (size_t)cap) && pin(&s.gid, cap) &&
                  pin(&s.partner, cap) && pin(&s.nb, cap) && pin(&s.dir, cap);
        for (int q = 0; q < 4; q++) ok = ok && pin(&s.q[q], cap);
        if (!ok) rc = -2;
    }
    for (int k = 0; k < ndev && !rc; k++)
        rc = edmd_halo_connect_direct(m->ctx[k], m->ctx[(k + ndev - 1) % ndev], m->ctx[(k + 1) % ndev]);
    if (rc) {
        edmd_cuda_destroy_mg(m);
        return rc;
    }
    edmd_cuda_get_box(m->ctx[0], &m->box);
    *out = m;
    return 0;
}

int edmd_cuda_mg_get_box(const edmd_mg *m, edmd_box *box)
{
    if (!m || !box) return EDMD_EINVAL;
    *box = m->box;
    return 0;
}

int edmd_cuda_mg_slab_sizes(const edmd_mg *m, int *n_owned /* [ndev] */)
{
    if (!m || !n_owned) return EDMD_EINVAL;
    for (int k = 0; k < m->ndev; k++) n_owned[k] = m->st[k].n;
    return 0;
}

int edmd_cuda_mg_upload(edmd_mg *m, const double *x, const double *y, const double *vx, const double *vy,
                        const double *rad, const int32_t *cell_xy, double t)
{
    if (!m) return EDMD_EINVAL;
    if (m->n > 0 && (!x || !y || !vx || !vy || !rad)) return mg_fail(m, EDMD_EINVAL, "mg_upload: null array");
    if (m->whole) {
        const int rc = edmd_cuda_upload(m->ctx[0], x, y, vx, vy, rad, cell_xy, t);
        if (rc) return mg_fail(m, rc, "mg_upload", m->ctx[0]);
        m->st[0].n = m->n;
        m->have_state = true;
        m->t = t;
        return 0;
    }
    const int ny = m->box.nycells, nx = m->box.nxcells, nd = m->ndev, N = m->n;
    // Deal the particles to the slabs by the row they are filed under, in ascending particle id inside every
    // slab, on several host threads: thread p counts its range per slab, a prefix over (thread, slab) gives every
    // thread its first slot in every slab, then every thread files its range.
    const int P = host_threads();
    std::vector<int> cnt((size_t)P * nd, 0);
    std::atomic<int> bad(0);
    auto cell_of = [&](int i, int &cx, int &cy) {
        if (cell_xy) {
            cx = cell_xy[2 * i];
            cy = cell_xy[2 * i + 1];
        } else {   // coordToCell, src/EDMD.c:2098-2107
            cx = (int)(x[i] * m->box.cellx_fac);
            cy = (int)(y[i] * m->box.celly_fac);
        }
        return cx >= 0 && cx < nx && cy >= 0 && cy < ny;
    };
    parallel_parts(P, [&](int p, int np) {
        const int lo = (int)((long long)N * p / np), hi = (int)((long long)N * (p + 1) / np);
        std::vector<int> mine(nd, 0);   // (thread-local: neighbouring threads' counters would share cache lines)
        for (int i = lo; i < hi; i++) {
            int cx, cy;
            if (!cell_of(i, cx, cy)) { bad = 1; continue; }
            mine[m->owner_of_row[cy]]++;
        }
        for (int k = 0; k < nd; k++) cnt[(size_t)p * nd + k] = mine[k];
    });
    if (bad) return mg_fail(m, EDMD_ECELL, "mg_upload: a particle's cell lies outside the cell grid");
    std::vector<int> base((size_t)P * nd, 0);
    for (int k = 0; k < nd; k++) {
        int run = 0;
        for (int p = 0; p < P; p++) {
            base[(size_t)p * nd + k] = run;
            run += cnt[(size_t)p * nd + k];
        }
        if (run > m->st[k].cap - 2 * m->halo_cap) return mg_fail(m, EDMD_EINVAL, "mg_upload: a slab holds more particles than its capacity (very uneven density)");
        m->st[k].n = run;
    }
    parallel_parts(P, [&](int p, int np) {
        const int lo = (int)((long long)N * p / np), hi = (int)((long long)N * (p + 1) / np);
        std::vector<int> pos(base.begin() + (size_t)p * nd, base.begin() + (size_t)(p + 1) * nd);   // thread-local cursors
        for (int i = lo; i < hi; i++) {
            int cx, cy;
            cell_of(i, cx, cy);
            edmd_mg::Stage &s = m->st[m->owner_of_row[cy]];
            const int k = pos[m->owner_of_row[cy]]++;
            s.x[k] = x[i]; s.y[k] = y[i]; s.vx[k] = vx[i]; s.vy[k] = vy[i]; s.rad[k] = rad[i];
            s.cells[2 * k] = cx; s.cells[2 * k + 1] = cy;
            s.gid[k] = i;
        }
    });
    // one host thread per device: the copies of the slabs run side by side on the devices' own PCIe links
    std::vector<int> rcs(nd, 0);
    parallel_parts(nd, [&](int k, int) {
        const edmd_mg::Stage &s = m->st[k];
        rcs[k] = edmd_cuda_upload_owned(m->ctx[k], s.n, s.x, s.y, s.vx, s.vy, s.rad, s.cells, s.gid, t);
    });
    for (int k = 0; k < nd; k++)
        if (rcs[k]) return mg_fail(m, rcs[k], "mg_upload (slab upload)", m->ctx[k]);
    m->have_state = true;
    m->t = t;
    return 0;
}

int edmd_cuda_mg_predict_all(edmd_mg *m, int mode, double *t_cross, uint8_t *dir, double *t_coll, int32_t *partner,
                             uint8_t *ctype, int32_t *overlap_pair)
{
    if (!m) return EDMD_EINVAL;
    if (mode != EDMD_MODE_NORMAL) return mg_fail(m, EDMD_EINVAL, "mg_predict_all: NORMAL mode only (the growth sweep is a single-GPU setup step)");
    if (!m->have_state) return mg_fail(m, EDMD_ESTATE, "mg_predict_all before mg_upload");
    if (overlap_pair) overlap_pair[0] = overlap_pair[1] = -1;
    if (m->whole) {
        const int rc = edmd_cuda_predict_all(m->ctx[0], mode, nullptr, t_cross, dir, t_coll, partner, ctype, overlap_pair);
        if (rc) mg_fail(m, rc, "mg_predict_all", m->ctx[0]);
        return rc;
    }
    // every device is launched before any is waited for: the slabs run side by side, their halo
    // kernels meet over NVLink
    for (int k = 0; k < m->ndev; k++) {
        const int rc = edmd_cuda_exchange_predict_device(m->ctx[k], mode);
        if (rc) return mg_fail(m, rc, "mg_predict_all (exchange + sweep)", m->ctx[k]);
    }
    // one host thread per device fetches its slab's outputs (pinned staging) and scatters them by particle id
    std::vector<int> rcs(m->ndev, 0);
    std::vector<int32_t> ovs(2 * (size_t)m->ndev, -1);
    parallel_parts(m->ndev, [&](int k, int) {
        edmd_mg::Stage &s = m->st[k];
        rcs[k] = edmd_cuda_fetch_predictions(m->ctx[k], s.t_cross, s.dir, s.t_coll, s.partner, nullptr, &ovs[2 * k]);
    });
    std::vector<char> ok(m->ndev);
    for (int k = 0; k < m->ndev; k++) ok[k] = rcs[k] == 0 || rcs[k] == EDMD_EOVERLAP;
    scatter_by_id_ranges(m, ok, [&](const edmd_mg::Stage &s, int j, int i) {
        if (t_cross) t_cross[i] = s.t_cross[j];
        if (dir) dir[i] = s.dir[j];
        if (t_coll) t_coll[i] = s.t_coll[j];
        if (partner) partner[i] = s.partner[j];
        if (ctype) ctype[i] = EDMD_EV_COLLISION;
    });
    int result = 0;
    for (int k = 0; k < m->ndev; k++) {
        if (rcs[k] == EDMD_EOVERLAP) {
            // the FIRST overlapping pair in sweep order = the smallest (i, j) over the slabs
            const int32_t *ov = &ovs[2 * k];
            if (overlap_pair && (result != EDMD_EOVERLAP || ov[0] < overlap_pair[0] ||
                                 (ov[0] == overlap_pair[0] && ov[1] < overlap_pair[1]))) {
                overlap_pair[0] = ov[0];
                overlap_pair[1] = ov[1];
            }
            result = EDMD_EOVERLAP;
            snprintf(m->err, sizeof(m->err), "mg_predict_all: %s", edmd_cuda_last_error(m->ctx[k]));
        } else if (rcs[k]) {
            return mg_fail(m, rcs[k], "mg_predict_all (fetch)", m->ctx[k]);
        }
    }
    return result;
}

int edmd_cuda_mg_boop_cutoff(edmd_mg *m, double r_c, double *q5, double *q6, double *q7, double *q6_arg,
                             int32_t *neighbors, double *mean_q6)
{
    if (!m) return EDMD_EINVAL;
    if (!m->have_state) return mg_fail(m, EDMD_ESTATE, "mg_boop_cutoff before mg_upload");
    if (m->whole) {
        const int rc = edmd_cuda_boop_cutoff(m->ctx[0], r_c, q5, q6, q7, q6_arg, neighbors, mean_q6);
        if (rc) mg_fail(m, rc, "mg_boop_cutoff", m->ctx[0]);
        return rc;
    }
    // the halo rows must match the resident state: exchange them (peer stores; asynchronous)
    for (int k = 0; k < m->ndev; k++) {
        const int rc = edmd_cuda_halo_exchange(m->ctx[k]);
        if (rc) return mg_fail(m, rc, "mg_boop_cutoff (halo exchange)", m->ctx[k]);
    }
    // one host thread per device: the slabs' psi6 kernels run side by side
    std::vector<int> rcs(m->ndev, 0);
    std::vector<double> means(m->ndev, 0.0);
    parallel_parts(m->ndev, [&](int k, int) {
        edmd_mg::Stage &s = m->st[k];
        rcs[k] = edmd_cuda_boop_cutoff(m->ctx[k], r_c, s.q[0], s.q[1], s.q[2], s.q[3], s.nb, &means[k]);
    });
    double sum = 0.0;
    for (int k = 0; k < m->ndev; k++) {
        if (rcs[k]) return mg_fail(m, rcs[k], "mg_boop_cutoff", m->ctx[k]);
        sum += means[k] * m->st[k].n;   // the slab's sum of q6 (its mean is sum / n_owned); added in slab order
    }
    scatter_by_id_ranges(m, std::vector<char>(m->ndev, 1), [&](const edmd_mg::Stage &s, int j, int i) {
        if (q5) q5[i] = s.q[0][j];
        if (q6) q6[i] = s.q[1][j];
        if (q7) q7[i] = s.q[2][j];
        if (q6_arg) q6_arg[i] = s.q[3][j];
        if (neighbors) neighbors[i] = s.nb[j];
    });
    if (mean_q6) *mean_q6 = m->n > 0 ? sum / m->n : 0.0;
    return 0;
}

int edmd_cuda_mg_pcf(edmd_mg *m, const double *x, const double *y, double dr, double max_r, uint64_t *counts,
                     double *g_r, int *num_bins)
{
    if (!m || !num_bins) return EDMD_EINVAL;
    if (!(dr > 0) || !(max_r > 0)) return mg_fail(m, EDMD_EINVAL, "mg_pcf: dr and max_r must be positive");
    const int nb = (int)(max_r / dr);   // `int num_bins = (int)(max_r / dr)` pcf.c:20
    *num_bins = nb;
    if (!counts && !g_r) return 0;
    if (m->n > 0 && (!x || !y)) return mg_fail(m, EDMD_EINVAL, "mg_pcf: positions needed");
    const size_t N = (size_t)m->n;
    std::vector<double> xy(2 * N);
    for (size_t i = 0; i < N; i++) {
        xy[2 * i] = x[i];
        xy[2 * i + 1] = y[i];
    }
    std::vector<double *> dxy(m->ndev, nullptr);
    std::vector<unsigned long long *> dcnt(m->ndev, nullptr);
    std::vector<unsigned long long> total((size_t)(nb > 0 ? nb : 1), 0ull), part((size_t)(nb > 0 ? nb : 1));
    int rc = 0;
    for (int k = 0; k < m->ndev && !rc; k++) {   // positions to every device, its share of the tile pairs launched
        if (cudaSetDevice(m->devices[k]) != cudaSuccess || cudaMalloc((void **)&dxy[k], (2 * N + 2) * sizeof(double)) != cudaSuccess ||
            cudaMalloc((void **)&dcnt[k], total.size() * sizeof(unsigned long long)) != cudaSuccess) {
            rc = -2;
            break;
        }
        cudaMemcpy(dxy[k], xy.data(), 2 * N * sizeof(double), cudaMemcpyHostToDevice);
        cudaMemset(dcnt[k], 0, total.size() * sizeof(unsigned long long));
        int nb2 = 0;
        rc = edmd_cuda_pcf_device(m->ctx[k], dxy[k], m->n, dr, max_r, k, m->ndev,
                                  reinterpret_cast<uint64_t *>(dcnt[k]), &nb2);
        if (rc) mg_fail(m, rc, "mg_pcf", m->ctx[k]);
    }
    for (int k = 0; k < m->ndev; k++) {   // the integer histograms are added on the host
        if (dcnt[k] && !rc) {
            cudaSetDevice(m->devices[k]);
            cudaDeviceSynchronize();
            if (cudaMemcpy(part.data(), dcnt[k], total.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost) != cudaSuccess)
                rc = -1;
            for (size_t b = 0; b < total.size(); b++) total[b] += part[b];
        }
        if (dxy[k]) cudaFree(dxy[k]);
        if (dcnt[k]) cudaFree(dcnt[k]);
    }
    if (rc) return rc;
    for (int b = 0; b < nb; b++) {
        if (counts) counts[b] = total[b];
        if (g_r) {
            // normalisation, src/pcf.c:56-72 (host side: num_bins values)
            const double volume = m->lx * m->ly;
            const double density = m->n / volume;
            const double r = (b + 0.5) * dr;
            const double shell_volume = 2 * M_PI * r * dr;
            const double norm = shell_volume * density * m->n;
            const double g = 2.0 * (double)total[b];
            g_r[b] = norm > 0 ? g / norm : 0.0;
        }
    }
    return 0;
}

}  // extern "C"
