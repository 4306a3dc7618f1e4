// rowstage.cuh -- per-warp shared-memory staging of the three cell rows a
// 32-slot chunk of the cell-ordered array needs; shared by the prediction
// sweep (K1) and the psi6 kernel (K4).
//
// A warp owns the 32 consecutive slots of one chunk; all of them lie in one
// cell row Y (rows start on 32-slot boundaries).  For row Y+j (j = -1, 0, 1,
// periodic in y: PBCcellY, src/EDMD.c:2118-2124) the candidates of lane L are
// the cells pcx-1 .. pcx+1 of its padded cell index, i.e. ONE contiguous range
// [lo_j(L), hi_j(L)) of the record array (ghost cells make this true at the
// periodic x edge too), and because slots are in cell order the ranges of the
// 32 lanes overlap and are monotone: their union is the contiguous segment
// [lo_j(first active lane), hi_j(last active lane)).  The warp copies each of
// the three segments with fully coalesced 16-byte loads (a 48-byte record is
// three uint4) into its private shared-memory window; no block barrier is
// involved.  The 48-byte record stride is conflict-free for 128-bit shared
// loads of consecutive records.
//
// Row order j = -1, 0, 1 and ascending cell order inside a segment are exactly
// the reference's scan order (src/EDMD.c:2959-2965).
#pragma once

#include "edmd_internal.cuh"

constexpr int kStageThreads = 128;               // 4 warps per CTA
constexpr int kStageWarps = kStageThreads / 32;
constexpr int kCapW = 48;                         // records per staged row segment

struct WarpStage {
    SRec rec[3][kCapW];
};

struct RowLane {
    int Y;          // cell row of the chunk
    int s;          // own slot
    int pc;         // own padded cell id
    int pcx;        // ... and its column 0..nx+1
    bool active;    // valid slot and not a ghost entry
    int lo[3], hi[3];   // candidate ranges, relative to the staged segment
    int self;       // own index inside staged row 1
};

__device__ __forceinline__ int row_wrap(int a, int n)
{
    if (a < 0) return a + n;
    if (a >= n) return a - n;
    return a;
}

// Returns 0 when the warp has nothing to do, 1 when staged, 2 when a segment
// does not fit (caller uses the global-memory path; lo/hi are then ABSOLUTE
// slot ranges).  Warp-uniform result; all 32 lanes must call.
__device__ __forceinline__ int row_stage(WarpStage &w, const CellIndex &g, int chunk, RowLane &rl)
{
    const int lane = threadIdx.x & 31;
    rl.Y = g.chunk_row[chunk];
    if (rl.Y < 0) return 0;
    rl.s = chunk * 32 + lane;
    const int rbY = g.row_base[rl.Y];
    const bool valid = rl.s < rbY + g.row_total[rl.Y];
    rl.pc = valid ? g.srec[rl.s].pc : 0;
    rl.pcx = rl.pc - rl.Y * g.ps;
    rl.active = valid && rl.pcx >= 1 && rl.pcx <= g.nx;
    if (!__any_sync(0xffffffffu, rl.active)) return 0;

    int seg_lo[3], seg_len[3];
    bool fits = true;
#pragma unroll
    for (int j = 0; j < 3; j++) {
        const int Yr = row_wrap(rl.Y - 1 + j, g.ny);
        const int rb = g.row_base[Yr];
        const int32_t *o = g.off + (size_t)Yr * g.ps;
        int lo = 0x7fffffff, hi = -1;
        if (rl.active) {
            lo = rb + o[rl.pcx - 1];
            hi = rb + o[rl.pcx + 2];
        }
        seg_lo[j] = __reduce_min_sync(0xffffffffu, lo);
        const int seg_hi = __reduce_max_sync(0xffffffffu, hi);
        seg_len[j] = seg_hi - seg_lo[j];
        fits = fits && (seg_len[j] <= kCapW);
        rl.lo[j] = lo;
        rl.hi[j] = hi;
    }
    if (!fits) return 2;
#pragma unroll
    for (int j = 0; j < 3; j++) {
        const uint4 *src = reinterpret_cast<const uint4 *>(g.srec + seg_lo[j]);
        uint4 *dst = reinterpret_cast<uint4 *>(&w.rec[j][0]);
        const int words = 3 * seg_len[j];
        for (int k = lane; k < words; k += 32) dst[k] = src[k];
        rl.lo[j] -= seg_lo[j];
        rl.hi[j] -= seg_lo[j];
    }
    rl.self = rl.s - seg_lo[1];
    __syncwarp();
    return 1;
}
