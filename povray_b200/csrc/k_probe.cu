// Component-level harness kernels: the device Solve_Polynomial and Noise / DNoise / Turbulence on explicit inputs
// (pvgpu_solve_polynomial, pvgpu_noise in include/pvgpu.h), so that they can be checked against known-answer vectors
// taken from the reference's own functions (tests/golden/make_golden_probe.py).
#include "pv_solver.cuh"
#include "pv_noise.cuh"
#include "pv_kernels.hpp"

namespace pvgpu {

__global__ void k_probe_solver(uint32_t n, const int32_t* degree, const int32_t* sturm, const double* epsilon, const double* coeffs,
                               double* roots, int32_t* counts)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        double c[5], r[4] = { 0.0, 0.0, 0.0, 0.0 };
        const int deg = degree[i];
        for (int k = 0; k <= deg; k++) c[k] = coeffs[5 * (size_t)i + (4 - deg) + k];
        const int cnt = solve_polynomial(deg, c, r, sturm[i], epsilon[i]);
        counts[i] = cnt;
        for (int k = 0; k < 4; k++) roots[4 * (size_t)i + k] = (k < cnt) ? r[k] : 0.0;
    }
}

__global__ void k_probe_noise(NoiseTables nt, uint32_t n, const double* xyz, const int32_t* gen, const int32_t* octaves, double* out)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const V3 p = mk(xyz[3 * (size_t)i], xyz[3 * (size_t)i + 1], xyz[3 * (size_t)i + 2]);
        double* o = out + 5 * (size_t)i;
        o[0] = noise3(nt, p, gen[i]);
        const V3 dn = dnoise3(nt, p);
        o[1] = dn.x; o[2] = dn.y; o[3] = dn.z;
        o[4] = turbulence(nt, p, octaves[i], 2.0, 0.5, gen[i]);
    }
}

// Initialize_Waves (noise.cpp:189-205): wave sources = normalised DNoise(<i, 0, 0>), frequencies from the reference's LCG
__global__ void k_init_waves(NoiseTables nt, uint32_t n, double* sources, double* freqs)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const V3 s = normalized(dnoise3(nt, mk((double)i, 0.0, 0.0)));
        sources[3 * i] = s.x; sources[3 * i + 1] = s.y; sources[3 * i + 2] = s.z;
        int next_rand = -560851967;
        for (uint32_t k = 0; k <= i; k++) next_rand = (int)((long long)next_rand * 1812433253LL + 12345LL);
        freqs[i] = ((double)((int)(next_rand >> 16) & 0x7FFF) * 0.000030518509476) + 0.01;
    }
}

// FP64 vector peak: eight independent DFMA chains per thread (explicit __fma_rn: the library is compiled with -fmad=false).
__global__ void __launch_bounds__(256)
k_fp64_peak(double* out, int iters)
{
    double a0 = 1.0 + threadIdx.x * 1e-9, a1 = a0 + 1e-9, a2 = a0 + 2e-9, a3 = a0 + 3e-9, a4 = a0 + 4e-9, a5 = a0 + 5e-9, a6 = a0 + 6e-9, a7 = a0 + 7e-9;
    const double m = 1.0 - 1e-12, c = 1e-12;
    for (int i = 0; i < iters; i++) {
        a0 = __fma_rn(a0, m, c); a1 = __fma_rn(a1, m, c); a2 = __fma_rn(a2, m, c); a3 = __fma_rn(a3, m, c);
        a4 = __fma_rn(a4, m, c); a5 = __fma_rn(a5, m, c); a6 = __fma_rn(a6, m, c); a7 = __fma_rn(a7, m, c);
    }
    const double r = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (r == 12345.678) *out = r;       // never true: keeps the chains alive
}

unsigned long long launch_fp64_peak(double* out, int iters, cudaStream_t st)
{
    const int blocks = sm_count() * 8;
    k_fp64_peak<<<blocks, 256, 0, st>>>(out, iters);
    return 2ull * 8ull * (unsigned long long)iters * 256ull * (unsigned long long)blocks;
}

void launch_init_waves(const NoiseTables& nt, uint32_t n, double* sources, double* freqs, cudaStream_t st)
{ if (n) k_init_waves<<<grid_for(n, 128, 8), 128, 0, st>>>(nt, n, sources, freqs); }

void launch_probe_solver(uint32_t n, const int32_t* degree, const int32_t* sturm, const double* epsilon, const double* coeffs,
                         double* roots, int32_t* counts, cudaStream_t st)
{ k_probe_solver<<<grid_for(n, 128, 8), 128, 0, st>>>(n, degree, sturm, epsilon, coeffs, roots, counts); }
void launch_probe_noise(const NoiseTables& nt, uint32_t n, const double* xyz, const int32_t* gen, const int32_t* octaves, double* out, cudaStream_t st)
{ k_probe_noise<<<grid_for(n, 128, 8), 128, 0, st>>>(nt, n, xyz, gen, octaves, out); }

}  // namespace pvgpu
