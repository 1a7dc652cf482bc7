// Shadow-ray kernel for scenes with transparent shadow casters: closest-hit loop towards the light, the light
// colour is filtered through every transparent blocker (ComputeShadowColour, trace.cpp:2274-2439).
#include "pv_shadow.cuh"

namespace pvgpu {

__global__ void __launch_bounds__(128)
k_shadow_filter(DScene sc, const SRay* __restrict__ rays, const PRay* __restrict__ wave, float4* accum, Counters* cnt)
{
    uint2 stack[PV_STACK_SIZE];
    const uint32_t n = cnt->n_shadow;
    unsigned long long tests = 0;
    if (blockIdx.x == 0 && threadIdx.x == 0 && n) atomicAdd(&cnt->shadow_rays, (unsigned long long)n);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const SRay s = rays[i];
        float f[3];
        trace_shadow<false>(sc, ld3(s.o), ld3(s.d), s.depth, wave, s.parent, stack, cnt, f, tests);
        accum_add(accum, s.sample, s.a[0] * f[0], s.a[1] * f[1], s.a[2] * f[2], 0.0f);
    }
    for (int off = 16; off > 0; off >>= 1) tests += __shfl_down_sync(0xffffffffu, tests, off);
    if ((threadIdx.x & 31) == 0 && tests) atomicAdd(&cnt->shadow_tests, tests);
}

void launch_shadow_filter(const DScene& sc, const SRay* rays, const PRay* wave, uint32_t n_max, float4* accum, Counters* cnt, cudaStream_t st)
{
    k_shadow_filter<<<grid_for(n_max, 128, 8), 128, 0, st>>>(sc, rays, wave, accum, cnt);
}

}  // namespace pvgpu
