// halo.cuh -- record and inbox layout of the peer-to-peer halo exchange (halo.cu), shared with the
// tile sweep's fused receive + partition kernel (tile_sweep.cu).
#pragma once

#include "edmd_internal.cuh"

struct __align__(16) HaloRec {
    double x, y, vx, vy, rad;
    int gid;
    int cell;   // padded column 1..nx (the row is implied by which neighbour sent it)
};
static_assert(sizeof(HaloRec) == 48, "halo record");

struct __align__(16) InboxHeader {
    int count, epoch, pad[2];
};

__host__ __device__ inline size_t inbox_bytes(int H)
{
    return (sizeof(InboxHeader) + sizeof(HaloRec) * (size_t)H + 255) & ~(size_t)255;
}
__host__ __device__ inline size_t inbox_offset(int H, int from, int parity)
{
    return inbox_bytes(H) * (size_t)(2 * from + parity);
}
__host__ __device__ inline size_t ack_offset(int H) { return inbox_bytes(H) * 4; }

__device__ __forceinline__ int ld_volatile(const int *p)
{
    return *reinterpret_cast<const volatile int *>(p);
}
