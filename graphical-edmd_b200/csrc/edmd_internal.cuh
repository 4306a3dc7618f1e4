// edmd_internal.cuh -- device-side data layout and kernel launchers shared by
// the translation units of libedmd_cuda.so (sm_100a only).
//
// HBM layout (N particles, NC = nx*ny cells), all allocated once per context:
//
//   resident state (original particle order, written by upload / free_fly)
//     xv    double4[N]   (x, y, vx, vy)   one 32-byte sector per particle
//     rad   double [N]
//     vr    double [N]   growth rates (GROW mode only)
//     cid   int32  [N]   row-major cell id Y*nx + X   (src/EDMD.c:2071)
//
//   cell index (rebuilt by every sweep, K0)
//     cell_cnt   int32[NC]     histogram, self-cleaning (returns to zero)
//     cell_start int32[NC+1]   exclusive scan; cell c owns [start[c], start[c+1])
//     slot_id    int32[N]      particle ids bucketed by cell (arbitrary in-cell order)
//     sxv/srad/svr/sid/scid    state gathered into cell order; inside a cell
//                              the order is DESCENDING particle id, i.e. the
//                              reference's linked-list order after
//                              cellListInit (head insertion, src/EDMD.c:1906-1920,
//                              2071-2072), so a plain strict-> running minimum
//                              reproduces the reference's tie-breaking.
//
//   sweep outputs (original particle order)
//     t_cross f64[N], t_coll f64[N], partner i32[N], dir u8[N], ctype u8[N]
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "edmd_cuda.h"

struct edmd_dev_box {
    int n, nx, ny, nc;
    double lx, ly, half_lx, half_ly;
    double csx, csy, fx, fy;
};

struct edmd_ctx {
    int device;
    int n;
    edmd_box box;
    edmd_dev_box dbox;
    cudaStream_t stream;
    cudaEvent_t ev[4];
    char err[512];
    uint64_t launches;
    bool force_generic;  // EDMD_OPT_FORCE_GENERIC: global-memory exact kernel only

    bool have_state;     // upload done
    bool have_pred;      // device predictions valid
    bool have_index;     // cell index matches resident state
    bool have_vr;
    double t;            // time of the resident snapshot

    // upload staging (device SoA mirror of the host arrays)
    double *in_soa;      // 5*N doubles: x | y | vx | vy | rad
    int32_t *in_cell;    // 2*N
    // pinned host staging
    void *h_pin;
    size_t h_pin_bytes;

    // resident state
    double4 *xv;
    double *rad, *vr;
    int32_t *cid;

    // cell index
    int32_t *cell_cnt, *cell_start, *slot_id;
    uint32_t *scan_state[2];   // decoupled look-back tile states, ping-pong
    int32_t *scan_ticket;      // [2] dynamic tile counters, ping-pong
    int scan_tiles;
    int scan_parity;
    double4 *sxv;
    double *srad, *svr;
    int32_t *sid, *scid;

    // outputs
    double *t_cross, *t_coll;
    int32_t *partner;
    uint8_t *dir, *ctype;
    unsigned long long *overlap_key;  // min over (i<<32 | j), ~0ull = none
    int32_t *flags;                   // [0] bad cell id seen

    // analysis scratch
    unsigned long long *pcf_counts;   // capacity pcf_cap bins
    int pcf_cap;
    double *boop;                     // 4*N doubles q5|q6|q7|arg
    int32_t *boop_nb;                 // N
    double *red_partial;              // block partials for deterministic sums
    int red_cap;
    char *flush_buf;                  // L2 flush scratch (bench only)
    size_t flush_cap;
};

// ---- launchers (each returns the number of kernels it launched) ----------
int edmd_launch_pack(edmd_ctx *c, bool have_cells);
int edmd_launch_cell_index(edmd_ctx *c, int mode);
int edmd_launch_predict(edmd_ctx *c, int mode);
int edmd_launch_free_fly(edmd_ctx *c, int mode, double dt);
int edmd_launch_boop(edmd_ctx *c, double r_c);
int edmd_launch_mean(edmd_ctx *c, const double *v, int n, double *out_dev);
int edmd_launch_pcf(edmd_ctx *c, double dr, double max_r, int num_bins);
