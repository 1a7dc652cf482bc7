// Shadow-ray kernel for scenes with transparent shadow casters: closest-hit loop towards the light, the light
// colour is filtered through every transparent blocker (ComputeShadowColour, trace.cpp:2274-2439).
#include "pv_shadow.cuh"

namespace pvgpu {

__global__ void __launch_bounds__(PV_TRAV_BLOCK, PV_TRAV_MIN_BLOCKS)
PV_VARIANT(k_shadow_filter)(DScene sc, const SRay* __restrict__ rays, const PRay* __restrict__ wave, float4* accum, Counters* cnt)
{
#if PV_SSTACK > 0
    __shared__ uint2 stack_sh[PV_SSTACK * PV_TRAV_BLOCK];
    uint2 stack_lo[PV_STACK_SIZE - PV_SSTACK];
    const TStack stack{ stack_sh + threadIdx.x, stack_lo, PV_SSTACK };
#else
    uint2 stack_lo[PV_STACK_SIZE];
    const TStack stack{ nullptr, stack_lo, 0 };
#endif
    const uint32_t n = cnt->n_shadow;
    unsigned long long tests = 0;
    if (blockIdx.x == 0 && threadIdx.x == 0 && n) atomicAdd(&cnt->shadow_rays, (unsigned long long)n);
    const uint32_t lane = threadIdx.x & 31u;
    for (uint32_t i0 = blockIdx.x * blockDim.x + (threadIdx.x - lane); i0 < n; i0 += gridDim.x * blockDim.x) {
        const uint32_t i = i0 + lane;
        bool alive = i < n;
        const SRay s = rays[alive ? i : 0u];
#if PV_HEAVY
        if (sc.has_area_lights && (sc.lights[s.light].flags & PVGPU_LIGHT_AREA)) alive = false;        // served by k_shadow_area
#endif
        float f[3];
        trace_shadow<false>(alive, sc, ld3(s.o), ld3(s.d), s.depth, wave, s.parent, stack, cnt, f, tests);
        if (alive) accum_add(accum, s.sample, s.a[0] * f[0], s.a[1] * f[1], s.a[2] * f[2], 0.0f);
    }
    for (int off = 16; off > 0; off >>= 1) tests += __shfl_down_sync(0xffffffffu, tests, off);
    if ((threadIdx.x & 31) == 0 && tests) atomicAdd(&cnt->shadow_tests, tests);
}

void PV_VARIANT(launch_shadow_filter)(const DScene& sc, const SRay* rays, const PRay* wave, uint32_t n_max, float4* accum, Counters* cnt, cudaStream_t st)
{
    PV_VARIANT(k_shadow_filter)<<<grid_for(n_max, PV_TRAV_BLOCK, PV_TRAV_MIN_BLOCKS), PV_TRAV_BLOCK, 0, st>>>(sc, rays, wave, accum, cnt);
}

}  // namespace pvgpu
