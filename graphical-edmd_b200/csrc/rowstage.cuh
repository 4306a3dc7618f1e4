// rowstage.cuh -- the per-warp, double-buffered, asynchronously fed pipeline
// shared by the prediction sweep (K1) and the psi6 kernel (K4).
//
// Work unit = one 32-slot CHUNK of the cell-ordered record array (all 32 slots
// lie in one cell row Y; rows start on 32-slot boundaries).  For row Y+j
// (j = -1, 0, 1, periodic in y: PBCcellY, src/EDMD.c:2118-2124) the candidates
// of a lane are the cells pcx-1 .. pcx+1 of its padded cell column, ONE
// contiguous range of the record array (the ghost cells make that true at the
// periodic x edge too), and the union over the 32 lanes is one contiguous
// segment per row.  K0 precomputes those three segments per chunk (ChunkMeta).
//
// Each warp is persistent and runs its own two-stage pipeline, no block
// barriers anywhere:
//     issue(next chunk):  one lane arms an mbarrier with the byte count and
//                         fires three cp.async.bulk (TMA 1-D bulk copies,
//                         48-byte records -> 16-byte aligned, any length) from
//                         the record array into the stage's shared buffer;
//                         all lanes cp.async the window of per-cell offsets
//     wait(current chunk): mbarrier try_wait + cp.async.wait_group
//     compute(current chunk) out of shared memory
// so the HBM/L2 latency of chunk k+1 hides behind the FP64 work of chunk k.
// The 48-byte record stride is bank-conflict-free for 128-bit shared loads of
// consecutive records.  Row order j = -1, 0, 1 and ascending cell order inside
// a segment are the reference's scan order (src/EDMD.c:2959-2965).
#pragma once

#include "edmd_internal.cuh"

constexpr int kStageThreads = 128;               // 4 independent warps per CTA
constexpr int kStageWarps = kStageThreads / 32;
constexpr int kCapW = 40;                         // records per staged row segment
constexpr int kOffW = 96;                         // cell-offset window per row

struct __align__(16) StageBuf {
    SRec rec[3][kCapW];
    int offw[3][kOffW];
};

struct RowLane {
    int Y;          // cell row of the chunk
    int s;          // own slot
    int pcx;        // own padded cell column 0..nx+1
    bool active;    // valid slot and not a ghost entry
    int lo[3], hi[3];   // candidate ranges (staged: indices into rec[j]; else absolute slots)
    int self;       // own index inside staged row 1
};

__device__ __forceinline__ int row_wrap(int a, int n)
{
    if (a < 0) return a + n;
    if (a >= n) return a - n;
    return a;
}

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// TMA 1-D bulk copy global -> shared, completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

__device__ __forceinline__ void cp_async4(void *dst, const void *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// warp-uniform copy of a chunk's metadata
__device__ __forceinline__ ChunkMeta load_meta(const ChunkMeta *m, int chunk)
{
    ChunkMeta r;
    const int4 *p = reinterpret_cast<const int4 *>(m + chunk);
    int4 *q = reinterpret_cast<int4 *>(&r);
    q[0] = p[0]; q[1] = p[1]; q[2] = p[2]; q[3] = p[3];
    return r;
}

// Fire the asynchronous copies of a chunk into `buf`.  Warp-uniform.
__device__ __forceinline__ void stage_issue(StageBuf &buf, uint64_t *bar, const CellIndex &g,
                                            const ChunkMeta &m, int lane)
{
    if (m.Y < 0 || (m.flags & kMetaOverflow)) {
        cp_async_commit();  // keep the per-thread group count in step
        return;
    }
    if (lane == 0) {
        const uint32_t bytes = (uint32_t)sizeof(SRec) * (uint32_t)(m.seg_len[0] + m.seg_len[1] + m.seg_len[2]);
        mbar_expect_tx(bar, bytes);
#pragma unroll
        for (int j = 0; j < 3; j++)
            if (m.seg_len[j] > 0)
                bulk_g2s(&buf.rec[j][0], g.srec + m.seg_lo[j], (uint32_t)sizeof(SRec) * m.seg_len[j], bar);
    }
    // per-cell offsets of columns cfirst-1 .. cfirst+ncells+1 (ncells + 3 values) of the three rows
    const int nwin = m.ncells + 3;
#pragma unroll
    for (int j = 0; j < 3; j++) {
        const int32_t *o = g.off + (size_t)row_wrap(m.Y - 1 + j, g.ny) * g.ps + (m.cfirst - 1);
        for (int k = lane; k < nwin; k += 32) cp_async4(&buf.offw[j][k], o + k);
    }
    cp_async_commit();
}

// Wait for a chunk's copies and derive this lane's view of it.  Returns 0 when
// the chunk is empty, 1 when staged, 2 on overflow (ranges are then absolute
// slots read from global memory).  Warp-uniform result.
__device__ __forceinline__ int stage_wait(const StageBuf &buf, uint64_t *bar, uint32_t parity,
                                          const CellIndex &g, const ChunkMeta &m, int chunk, int lane,
                                          bool more_in_flight, RowLane &rl)
{
    if (more_in_flight) cp_async_wait<1>();
    else cp_async_wait<0>();
    if (m.Y < 0) return 0;
    rl.Y = m.Y;
    rl.s = chunk * 32 + lane;
    const bool valid = rl.s < m.row_end;
    if (m.flags & kMetaOverflow) {
        const int pc = valid ? g.srec[rl.s].pc : 0;
        rl.pcx = pc - m.Y * g.ps;
        rl.active = valid && rl.pcx >= 1 && rl.pcx <= g.nx;
#pragma unroll
        for (int j = 0; j < 3; j++) {
            const int Yr = row_wrap(m.Y - 1 + j, g.ny);
            const int rb = g.row_base[Yr];
            const int32_t *o = g.off + (size_t)Yr * g.ps;
            rl.lo[j] = rl.active ? rb + o[rl.pcx - 1] : 0;
            rl.hi[j] = rl.active ? rb + o[rl.pcx + 2] : 0;
        }
        return 2;
    }
    mbar_wait(bar, parity);
    __syncwarp();
    rl.self = rl.s - m.seg_lo[1];
    const int pc = valid ? buf.rec[1][rl.self].pc : 0;
    rl.pcx = pc - m.Y * g.ps;
    rl.active = valid && rl.pcx >= 1 && rl.pcx <= g.nx;
    const int w = rl.active ? rl.pcx - m.cfirst : 0;   // window index of column pcx-1
#pragma unroll
    for (int j = 0; j < 3; j++) {
        rl.lo[j] = buf.offw[j][w] + m.delta[j];
        rl.hi[j] = buf.offw[j][w + 3] + m.delta[j];
    }
    return 1;
}

// dynamic shared memory a row_pipeline kernel must be launched with
constexpr size_t kStageSmem = sizeof(StageBuf) * kStageWarps * 2 + sizeof(uint64_t) * kStageWarps * 2;

// Persistent per-warp pipeline.  `body(buf, meta, rl, status)` is called with
// all 32 lanes converged for every non-empty chunk.
template <class Body>
__device__ __forceinline__ void row_pipeline(const CellIndex &g, const ChunkMeta *meta, int nchunks,
                                             Body body)
{
    extern __shared__ __align__(128) unsigned char stage_smem[];
    StageBuf(*bufs)[2] = reinterpret_cast<StageBuf(*)[2]>(stage_smem);
    uint64_t(*bars)[2] = reinterpret_cast<uint64_t(*)[2]>(stage_smem + sizeof(StageBuf) * kStageWarps * 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int stride = gridDim.x * kStageWarps;
    int c = blockIdx.x * kStageWarps + warp;
    if (c >= nchunks) return;
    if (lane == 0) {
        mbar_init(&bars[warp][0], 1);
        mbar_init(&bars[warp][1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    uint32_t parity = 0;  // bit st = phase parity of stage st
    int st = 0;
    ChunkMeta m = load_meta(meta, c);
    stage_issue(bufs[warp][0], &bars[warp][0], g, m, lane);
    while (true) {
        const int cn = c + stride;
        const bool more = cn < nchunks;
        ChunkMeta mn;
        if (more) {
            mn = load_meta(meta, cn);
            stage_issue(bufs[warp][st ^ 1], &bars[warp][st ^ 1], g, mn, lane);
        }
        RowLane rl;
        const int status = stage_wait(bufs[warp][st], &bars[warp][st], (parity >> st) & 1u, g, m, c, lane,
                                      more, rl);
        if (status == 1) parity ^= 1u << st;
        if (status != 0) body(bufs[warp][st], m, rl, status);
        __syncwarp();  // every lane is done with this stage before it is refilled
        if (!more) break;
        m = mn;
        c = cn;
        st ^= 1;
    }
}
