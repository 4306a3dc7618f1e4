// Microbenchmark behind the record layout choice: cost of scattering
// 1M records to random slots with different store shapes.  Build & run:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/scatter_stores scatter_stores.cu && /tmp/scatter_stores
#include <cstdio>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <random>
#include <cuda_runtime.h>

struct __align__(16) R48 { double a[5]; int i, j; };
struct __align__(32) R64 { double a[6]; int i, j; double pad; };
struct __align__(32) R32 { double a[4]; };
struct __align__(16) R16 { double r; int i, j; };

__device__ __forceinline__ void st256(void* p, const void* s) {
    const uint32_t* r = (const uint32_t*)s;
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" :: "l"(p), "r"(r[0]),"r"(r[1]),"r"(r[2]),"r"(r[3]),"r"(r[4]),"r"(r[5]),"r"(r[6]),"r"(r[7]) : "memory");
}

__global__ void k48(int n, const int* perm, const double4* xv, const double* rad, R48* out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= n) return;
    double4 p = xv[i]; R48 r; r.a[0]=p.x; r.a[1]=p.y; r.a[2]=p.z; r.a[3]=p.w; r.a[4]=rad[i]; r.i=i; r.j=perm[i];
    uint4* d = (uint4*)(out + perm[i]); const uint4* s = (const uint4*)&r; d[0]=s[0]; d[1]=s[1]; d[2]=s[2];
}
__global__ void k64(int n, const int* perm, const double4* xv, const double* rad, R64* out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= n) return;
    double4 p = xv[i]; R64 r; r.a[0]=p.x; r.a[1]=p.y; r.a[2]=p.z; r.a[3]=p.w; r.a[4]=rad[i]; r.a[5]=0; r.i=i; r.j=perm[i]; r.pad=0;
    st256(out + perm[i], &r); st256((char*)(out + perm[i]) + 32, (char*)&r + 32);
}
__global__ void ksplit(int n, const int* perm, const double4* xv, const double* rad, R32* o32, R16* o16) {
    int i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= n) return;
    double4 p = xv[i]; R32 a; a.a[0]=p.x; a.a[1]=p.y; a.a[2]=p.z; a.a[3]=p.w; R16 b; b.r=rad[i]; b.i=i; b.j=perm[i];
    st256(o32 + perm[i], &a); *(uint4*)(o16 + perm[i]) = *(uint4*)&b;
}
__global__ void k32only(int n, const int* perm, const double4* xv, R32* o32) {
    int i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= n) return;
    double4 p = xv[i]; R32 a; a.a[0]=p.x; a.a[1]=p.y; a.a[2]=p.z; a.a[3]=p.w; st256(o32 + perm[i], &a);
}
__global__ void kgather48(int n, const int* perm, const R48* in, R48* out) {   // random READ, coalesced write
    int i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= n) return;
    const uint4* s = (const uint4*)(in + perm[i]); uint4* d = (uint4*)(out + i); d[0]=s[0]; d[1]=s[1]; d[2]=s[2];
}

int main() {
    const int n = 1000000;
    std::vector<int> perm(n); for (int i = 0; i < n; i++) perm[i] = i;
    std::mt19937 g(1); std::shuffle(perm.begin(), perm.end(), g);
    int* dperm; double4* xv; double* rad; void *o1, *o2, *flush;
    cudaMalloc(&dperm, n * 4); cudaMalloc(&xv, n * 32); cudaMalloc(&rad, n * 8);
    cudaMalloc(&o1, (size_t)n * 64); cudaMalloc(&o2, (size_t)n * 64); cudaMalloc(&flush, 256 << 20);
    cudaMemcpy(dperm, perm.data(), n * 4, cudaMemcpyHostToDevice);
    cudaMemset(xv, 0, n * 32); cudaMemset(rad, 0, n * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const char* names[] = {"3xSTG.128 (48 B record)", "2xSTG.256 (64 B record)", "STG.256 + STG.128 (32+16 split)", "STG.256 only (32 B)", "random 48 B gather, coalesced write"};
    for (int v = 0; v < 5; v++) {
        float best = 1e9;
        for (int it = 0; it < 6; it++) {
            cudaMemsetAsync(flush, it, 256 << 20);
            cudaEventRecord(e0);
            int b = (n + 255) / 256;
            if (v == 0) k48<<<b, 256>>>(n, dperm, xv, rad, (R48*)o1);
            if (v == 1) k64<<<b, 256>>>(n, dperm, xv, rad, (R64*)o1);
            if (v == 2) ksplit<<<b, 256>>>(n, dperm, xv, rad, (R32*)o1, (R16*)o2);
            if (v == 3) k32only<<<b, 256>>>(n, dperm, xv, (R32*)o1);
            if (v == 4) kgather48<<<b, 256>>>(n, dperm, (const R48*)o1, (R48*)o2);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (it > 1 && ms < best) best = ms;
        }
        printf("%-40s %7.1f us\n", names[v], best * 1e3);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
