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
// Each warp is persistent and independent (no block barriers anywhere):
//     issue:   one lane arms an mbarrier with the byte count and fires three
//              cp.async.bulk (TMA 1-D bulk copies) per record array (32-byte
//              kinematics, 16-byte tags) into the warp's shared buffer and
//              three more for the windows of per-cell offsets
//              (rows of off[] are 16-byte aligned); the next chunk's plan is
//              prefetched into L1
//     wait:    mbarrier try_wait
//     compute: out of shared memory
// A warp's instruction stream is a chain of dependent FP64 operations (one
// issue every ~7 cycles), so the SM is filled with 32 such warps (8 CTAs x 4):
// the copy latency of one warp hides behind the arithmetic of the other seven
// on its scheduler, which is worth more here than a second buffer per warp
// (measured: double buffering at 16 warps/SM was no faster).
// Row order j = -1, 0, 1 and ascending cell order inside
// a segment are the reference's scan order (src/EDMD.c:2959-2965).
#pragma once

#include "edmd_internal.cuh"

constexpr int kStageThreads = 128;               // 4 independent warps per CTA
constexpr int kStageWarps = kStageThreads / 32;
constexpr int kCapW = 40;                         // records per staged row segment (48 B each)
constexpr int kOffW = 64;                         // cell-offset window per row

struct __align__(32) StageBuf {
    SPos pos[3][kCapW];
    SAux aux[3][kCapW];
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

// shared-window addresses of one warp's buffer, computed once
struct StageAddr {
    uint32_t pos[3], aux[3], offw[3], bar;
};

// Fire the asynchronous copies of a chunk.  Warp-uniform; one lane issues nine
// TMA bulk copies (three segments of each record array, three off[] windows).
__device__ __forceinline__ void stage_issue(const StageAddr &sa, const CellIndex &g, const ChunkMeta &m,
                                            int lane)
{
    if (m.Y < 0 || (m.flags & kMetaOverflow)) return;
    if (lane == 0) {
        const uint32_t rbytes = (uint32_t)(sizeof(SPos) + sizeof(SAux)) *
                                (uint32_t)(m.seg_len[0] + m.seg_len[1] + m.seg_len[2]);
        const uint32_t wbytes = 4u * (uint32_t)m.wlen;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sa.bar),
                     "r"(rbytes + 3u * wbytes)
                     : "memory");
#pragma unroll
        for (int j = 0; j < 3; j++) {
            if (m.seg_len[j] > 0) {
                asm volatile(
                    "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                        sa.pos[j]),
                    "l"(g.spos + m.seg_lo[j]), "r"((uint32_t)sizeof(SPos) * (uint32_t)m.seg_len[j]), "r"(sa.bar)
                    : "memory");
                asm volatile(
                    "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                        sa.aux[j]),
                    "l"(g.saux + m.seg_lo[j]), "r"((uint32_t)sizeof(SAux) * (uint32_t)m.seg_len[j]), "r"(sa.bar)
                    : "memory");
            }
            const int32_t *o = g.off + (size_t)row_wrap(m.Y - 1 + j, g.ny) * g.ps + m.wstart;
            asm volatile(
                "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                    sa.offw[j]),
                "l"(o), "r"(wbytes), "r"(sa.bar)
                : "memory");
        }
    }
}

// Wait for a chunk's copies and derive this lane's view of it.  Returns 0 when
// the chunk is empty, 1 when staged, 2 on overflow (ranges are then absolute
// slots read from global memory).  Warp-uniform result.
__device__ __forceinline__ int stage_wait(const StageBuf &buf, uint64_t *bar, uint32_t parity,
                                          const CellIndex &g, const ChunkMeta &m, int chunk, int lane,
                                          RowLane &rl)
{
    if (m.Y < 0) return 0;
    rl.Y = m.Y;
    rl.s = chunk * 32 + lane;
    const bool valid = rl.s < m.row_end;
    if (m.flags & kMetaOverflow) {
        const int pc = valid ? g.saux[rl.s].pc : 0;
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
    rl.self = rl.s - m.seg_lo[1];
    const int pc = valid ? buf.aux[1][rl.self].pc : 0;
    rl.pcx = pc - m.Y * g.ps;
    rl.active = valid && rl.pcx >= 1 && rl.pcx <= g.nx;
    const int w = rl.active ? rl.pcx - 1 - m.wstart : 0;   // window index of column pcx-1
#pragma unroll
    for (int j = 0; j < 3; j++) {
        rl.lo[j] = buf.offw[j][w] + m.delta[j];
        rl.hi[j] = buf.offw[j][w + 3] + m.delta[j];
    }
    return 1;
}

// dynamic shared memory a row_pipeline kernel must be launched with
constexpr size_t kStageSmem = sizeof(StageBuf) * kStageWarps + sizeof(uint64_t) * kStageWarps;
constexpr int kStageCtasPerSm = 8;

// Persistent per-warp pipeline.  `body(buf, meta, rl, status)` is called with
// all 32 lanes converged for every non-empty chunk.
template <class Body>
__device__ __forceinline__ void row_pipeline(const CellIndex &g, const ChunkMeta *meta, int nchunks,
                                             Body body)
{
    extern __shared__ __align__(128) unsigned char stage_smem[];
    StageBuf *bufs = reinterpret_cast<StageBuf *>(stage_smem);
    uint64_t *bars = reinterpret_cast<uint64_t *>(stage_smem + sizeof(StageBuf) * kStageWarps);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int stride = gridDim.x * kStageWarps;
    int c = blockIdx.x * kStageWarps + warp;
    if (c >= nchunks) return;
    StageBuf &buf = bufs[warp];
    uint64_t *bar = &bars[warp];
    StageAddr sa;
#pragma unroll
    for (int j = 0; j < 3; j++) {
        sa.pos[j] = smem_u32(&buf.pos[j][0]);
        sa.aux[j] = smem_u32(&buf.aux[j][0]);
        sa.offw[j] = smem_u32(&buf.offw[j][0]);
    }
    sa.bar = smem_u32(bar);
    if (lane == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    uint32_t parity = 0;
    ChunkMeta m = load_meta(meta, c);
    stage_issue(sa, g, m, lane);
    while (true) {
        const int cn = c + stride;
        const bool more = cn < nchunks;
        if (more && lane == 0) asm volatile("prefetch.global.L1 [%0];" ::"l"(meta + cn));
        RowLane rl;
        const int status = stage_wait(buf, bar, parity, g, m, c, lane, rl);
        if (status == 1) parity ^= 1u;
        if (status != 0) body(buf, m, rl, status);
        __syncwarp();  // every lane is done with the buffer before it is refilled
        if (!more) break;
        c = cn;
        m = load_meta(meta, c);
        stage_issue(sa, g, m, lane);
    }
}
