// thermostat.cu -- the velocity part of a thermostat tick on the RESIDENT state
// (SURVEY.md 8f rank 2):
//   physicalQ        src/EDMD.c:5968-5997   E = sum 1/2 m v^2, px, py   (m = 1: the
//                                           reference's default initial conditions, :1467-1548)
//   addNoise, velocity-rescale branch  :4830-4832, :4899-4902   v /= sqrt(E/N/T)
// With edmd_cuda_free_fly (the `freeFly(p)` of the same loop, :4893) and
// edmd_cuda_predict_device (the re-predict loop :4909-4915) the whole tick runs on
// the device without the state crossing PCIe in between.
//
// The reference adds the N terms in particle order; a parallel sum cannot reproduce
// that rounding, so E (and with it every rescaled velocity) agrees to ~1e-15
// relative, not bit for bit -- inside the 1e-12 bar of the event times.  The sum IS
// reproducible run to run: fixed partition (kBlocks x kThreads strided), fixed tree.
#include "edmd_internal.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kBlocks = 592;   // 4 per SM

__device__ __forceinline__ void block_sum3(double &a, double &b, double &c)
{
    __shared__ double sh[3][kThreads];
    sh[0][threadIdx.x] = a; sh[1][threadIdx.x] = b; sh[2][threadIdx.x] = c;
    __syncthreads();
    for (int d = kThreads / 2; d > 0; d >>= 1) {
        if (threadIdx.x < d) {
            sh[0][threadIdx.x] += sh[0][threadIdx.x + d];
            sh[1][threadIdx.x] += sh[1][threadIdx.x + d];
            sh[2][threadIdx.x] += sh[2][threadIdx.x + d];
        }
        __syncthreads();
    }
    a = sh[0][0]; b = sh[1][0]; c = sh[2][0];
}

__global__ void __launch_bounds__(kThreads)
k_kinetic_partial(int n, const double4 *__restrict__ xv, double *__restrict__ partial)
{
    double e = 0, px = 0, py = 0;
    for (int i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads) {
        const double4 p = xv[i];
        // `E += 0.5 * particles[i].m * (vx * vx + particles[i].vy * particles[i].vy)` :5988
        e += 0.5 * (p.z * p.z + p.w * p.w);
        px += p.z;
        py += p.w;
    }
    block_sum3(e, px, py);
    if (threadIdx.x == 0) {
        partial[blockIdx.x] = e;
        partial[kBlocks + blockIdx.x] = px;
        partial[2 * kBlocks + blockIdx.x] = py;
    }
}

// out[0..2] = E, px, py;  out[3] = sqrt(E/N/T) (the divisor of the rescale), 0 when T <= 0
__global__ void __launch_bounds__(kThreads)
k_kinetic_final(int n, double T, const double *__restrict__ partial, double *__restrict__ out)
{
    double e = 0, px = 0, py = 0;
    for (int i = threadIdx.x; i < kBlocks; i += kThreads) {
        e += partial[i];
        px += partial[kBlocks + i];
        py += partial[2 * kBlocks + i];
    }
    block_sum3(e, px, py);
    if (threadIdx.x == 0) {
        out[0] = e; out[1] = px; out[2] = py;
        out[3] = T > 0 ? sqrt(e / n / T) : 0.0;   // `sqrt(E/N/T)` :4899
    }
}

// addNoise's rescale (`p->vx /= sqrt(E/N/T)`, src/EDMD.c:4899-4900) and
// normalizePhysicalQ (src/EDMD.c:5723-5764, the branch without a circular wall; unit masses):
//   `vx -= px/(N*m); vy -= py/(N*m);`  then, with the E of the shifted velocities,  `vx /= sqrt(E/N/Einit)`.
// shift = (dvx, dvy) subtracted first, then -- when divisor != 1 -- the division: the reference's two
// roundings.  `red` (nullable): take dvx = red[1]/n, dvy = red[2]/n from the device's sums instead.
// Optionally accumulates the kinetic sums of the NEW velocities (partial != nullptr) for the step that follows.
__global__ void __launch_bounds__(kThreads)
k_shift_scale(int n, int n_total, double dvx, double dvy, double divisor, const double *__restrict__ red, int div_from_red,
              double4 *__restrict__ xv, int32_t *__restrict__ flags, double *__restrict__ partial)
{
    if (red && !div_from_red) {
        // `px/(N*particles[i].m)` with m == 1
        dvx = __ddiv_rn(red[1], __dmul_rn((double)n_total, 1.0));
        dvy = __ddiv_rn(red[2], __dmul_rn((double)n_total, 1.0));
    }
    if (red && div_from_red) divisor = red[3];
    double e = 0, px = 0, py = 0;
    float vm = 0.0f;
    for (int i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads) {
        double4 p = xv[i];
        p.z = __dsub_rn(p.z, dvx);
        p.w = __dsub_rn(p.w, dvy);
        if (divisor != 1.0 && divisor > 0) {   // (0: a system without kinetic energy / T <= 0: left alone)
            p.z = __ddiv_rn(p.z, divisor);
            p.w = __ddiv_rn(p.w, divisor);
        }
        xv[i] = p;
        e += 0.5 * (p.z * p.z + p.w * p.w);
        px += p.z;
        py += p.w;
        float v = __double2float_ru(fmax(fabs(p.z), fabs(p.w)));
        if (!(v == v)) v = __int_as_float(0x7f800000);
        vm = fmaxf(vm, v);
    }
    const unsigned vmb = __reduce_max_sync(0xffffffffu, (unsigned)__float_as_int(vm));
    if ((threadIdx.x & 31) == 0 && vmb) atomicMax(reinterpret_cast<unsigned int *>(&flags[kFlagVmax]), vmb);
    if (partial) {
        block_sum3(e, px, py);
        if (threadIdx.x == 0) {
            partial[blockIdx.x] = e;
            partial[kBlocks + blockIdx.x] = px;
            partial[2 * kBlocks + blockIdx.x] = py;
        }
    }
}

// ---- Langevin kick (addNoise with noise == 1: randomGaussian, src/EDMD.c:5802-5826) ----
// Counter-based uniforms: the generator of graphical-edmd_b200/synth.py (splitmix64 finaliser over
// (seed, id, stream)), so that numpy reproduces every draw.
__device__ __forceinline__ unsigned long long mix64(unsigned long long z)
{
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__device__ __forceinline__ double uniform01(unsigned int seed, unsigned long long id, unsigned long long stream)
{
    unsigned long long z = (id + 1ull) * 0x9E3779B97F4A7C15ull;
    z += (unsigned long long)seed * 0xD1B54A32D192ED03ull;
    z += (stream + 1ull) * 0x8CB92BA72F3D8DD7ull;
    z = mix64(mix64(z));
    return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}

// non-Euler branch, no damping, unit masses:
//   c = exp(-gamm*dtnoise); std = sqrt(T*(1 - c*c)/m);
//   vx = std*a*cos(b) + vx*c;  vy = std*a*sin(b) + vy*c;     a = sqrt(-2 log u1), b = 2 pi u2
__global__ void __launch_bounds__(kThreads)
k_langevin(int n, double std, double c, unsigned int seed, unsigned int tick, const int32_t *__restrict__ gid,
           double4 *__restrict__ xv, int32_t *__restrict__ flags)
{
    const int i = blockIdx.x * kThreads + threadIdx.x;
    float vm = 0.0f;
    if (i < n) {
        const unsigned long long id = (unsigned long long)(gid ? gid[i] : i);
        const double u1 = 1.0 - uniform01(seed, id, 2ull * tick);        // (0, 1]: log is finite
        const double u2 = uniform01(seed, id, 2ull * tick + 1ull);
        const double a = sqrt(-2.0 * log(u1));
        const double b = 2.0 * 3.14159265358979323846 * u2;
        double sb, cb;
        sincos(b, &sb, &cb);
        double4 p = xv[i];
        p.z = std * a * cb + p.z * c;
        p.w = std * a * sb + p.w * c;
        xv[i] = p;
        vm = __double2float_ru(fmax(fabs(p.z), fabs(p.w)));
        if (!(vm == vm)) vm = __int_as_float(0x7f800000);
    }
    const unsigned vmb = __reduce_max_sync(0xffffffffu, (unsigned)__float_as_int(vm));
    if ((threadIdx.x & 31) == 0 && vmb) atomicMax(reinterpret_cast<unsigned int *>(&flags[kFlagVmax]), vmb);
}

}  // namespace

int edmd_launch_langevin(edmd_ctx *c, double T, double gamma, double dtnoise, unsigned int seed, unsigned int tick)
{
    const int n = c->n_owned;
    if (n == 0) return 0;
    const double cc = exp(-gamma * dtnoise);
    const double std = sqrt(T * (1.0 - cc * cc));
    cudaMemsetAsync(c->flags + kFlagVmax, 0, sizeof(int32_t), c->stream);
    k_langevin<<<(n + kThreads - 1) / kThreads, kThreads, 0, c->stream>>>(n, std, cc, seed, tick,
                                                                      c->slab ? c->gid : nullptr, c->xv, c->flags);
    return 1;
}

size_t edmd_thermostat_scratch_doubles() { return 3 * (size_t)kBlocks + 8; }

// scratch: edmd_thermostat_scratch_doubles(); results in scratch[3*kBlocks .. +3]
double *edmd_launch_kinetic(edmd_ctx *c, double T, double *scratch, int *launched)
{
    double *out = scratch + 3 * kBlocks;
    k_kinetic_partial<<<kBlocks, kThreads, 0, c->stream>>>(c->n_owned, c->xv, scratch);
    k_kinetic_final<<<1, kThreads, 0, c->stream>>>(c->n_owned, T, scratch, out);
    *launched += 2;
    return out;
}

// v <- (v - shift) / divisor on the owned particles.  red != nullptr: the shift is the centre-of-mass velocity
// red[1..2] / n_total (normalizePhysicalQ's first loop), or -- div_from_red -- the divisor is red[3].
// sums != nullptr: the kinetic sums of the new velocities go to sums (scratch of edmd_launch_kinetic's layout),
// finished by edmd_launch_kinetic_final.
int edmd_launch_shift_scale(edmd_ctx *c, double dvx, double dvy, double divisor, const double *red, bool div_from_red,
                            int n_total, double *sums)
{
    const int n = c->n_owned;
    if (n == 0) return 0;
    cudaMemsetAsync(c->flags + kFlagVmax, 0, sizeof(int32_t), c->stream);
    k_shift_scale<<<kBlocks, kThreads, 0, c->stream>>>(n, n_total, dvx, dvy, divisor, red, div_from_red ? 1 : 0, c->xv,
                                                       c->flags, sums);
    return 1;
}

// second half of edmd_launch_kinetic for sums accumulated elsewhere
double *edmd_launch_kinetic_final(edmd_ctx *c, double T, double *scratch, int n_total)
{
    double *out = scratch + 3 * kBlocks;
    k_kinetic_final<<<1, kThreads, 0, c->stream>>>(n_total, T, scratch, out);
    return out;
}

// `p->vx /= sqrt(E/N/T); p->vy /= sqrt(E/N/T);` (addNoise, src/EDMD.c:4899-4900) with the divisor red[3] of the
// kinetic sums: the shift-and-scale kernel with a zero shift (x - 0.0 is x, bit for bit).  (A thread-per-particle
// kernel for this took 29.8 us at N = 10^6 against 12.9 us for the grid-stride form: ncu, profiles/r2m_ncu_misc.json.)
int edmd_launch_rescale(edmd_ctx *c, const double *red)
{
    return edmd_launch_shift_scale(c, 0.0, 0.0, 1.0, red, true, c->n_owned, nullptr);
}
