// Shadow-ray kernel for scenes in which every shadow caster is opaque: any-hit search, no filtering.
#include "pv_shadow.cuh"

namespace pvgpu {

__global__ void __launch_bounds__(PV_TRAV_BLOCK, PV_TRAV_MIN_BLOCKS)
PV_VARIANT(k_shadow_opaque)(DScene sc, const SRay* __restrict__ rays, float4* accum, Counters* cnt)
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
        const SRay* sp = rays + (alive ? i : 0u);
#if PV_HEAVY
        if (sc.has_area_lights && (sc.lights[sp->light].flags & PVGPU_LIGHT_AREA)) alive = false;      // served by k_shadow_area
#endif
        const V3 o = ld3(sp->o), d = ld3(sp->d);
        const double depth = sp->depth;
        float f[3];
        trace_shadow<true>(alive, sc, o, d, depth, nullptr, 0u, stack, cnt, f, tests);
        if (alive && f[0] != 0.0f) accum_add(accum, sp->sample, sp->a[0], sp->a[1], sp->a[2], 0.0f);
    }
    for (int off = 16; off > 0; off >>= 1) tests += __shfl_down_sync(0xffffffffu, tests, off);
    if ((threadIdx.x & 31) == 0 && tests) atomicAdd(&cnt->shadow_tests, tests);
}

void PV_VARIANT(launch_shadow_opaque)(const DScene& sc, const SRay* rays, uint32_t n_max, float4* accum, Counters* cnt, cudaStream_t st)
{
    PV_VARIANT(k_shadow_opaque)<<<grid_for(n_max, PV_TRAV_BLOCK, PV_TRAV_MIN_BLOCKS), PV_TRAV_BLOCK, 0, st>>>(sc, rays, accum, cnt);
}

}  // namespace pvgpu
