/* edmd_cuda_bench.h -- measurement helper of the repository's own bench.py and profiling tools.
 * NOT part of the drop-in interface (include/edmd_cuda.h): a maintainer of the reference never
 * includes this file.  The symbol lives in the same shared library so that the timed kernels are
 * the shipped ones, launched on the context's own stream (torch's events would not see it). */
#ifndef EDMD_CUDA_BENCH_H
#define EDMD_CUDA_BENCH_H
#include "edmd_cuda.h"
#ifdef __cplusplus
extern "C" {
#endif

/* kernel ids for edmd_cuda_bench */
#define EDMD_BENCH_SWEEP 0  /* K0 (cell index) + K1 (predict) */
#define EDMD_BENCH_FREEFLY 1
#define EDMD_BENCH_BOOP 2
#define EDMD_BENCH_PCF 3    /* uses dr / max_r arguments */
#define EDMD_BENCH_VORONOI 4 /* K5: grid sort + one Voronoi cell per particle (psi, area, perimeter) */

/* Runs `warmup` untimed and `iters` timed passes of the chosen device path,
 * writing `flush_bytes` of scratch between passes (L2 flush, outside the timed
 * events) when flush_bytes > 0.  ms_total[iters] = whole pass,
 * ms_main[iters] = the dominant kernel alone (K1 for the sweep).  EDMD_BENCH_SWEEP with ms_main == NULL: no
 * event is recorded between K0 and K1, the two run as the product calls launch them (programmatic dependent
 * launch: K1's launch latency and prologue overlap K0's tail) -- ms_total is then the step as shipped. */
int edmd_cuda_bench(edmd_ctx *ctx, int what, int mode, double dr, double max_r,
                    int warmup, int iters, size_t flush_bytes, float *ms_total,
                    float *ms_main);

#ifdef __cplusplus
}
#endif
#endif /* EDMD_CUDA_BENCH_H */
