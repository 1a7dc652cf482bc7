// Shadow-ray kernel for scenes in which every shadow caster is opaque: any-hit search, no filtering.
#include "pv_shadow.cuh"

namespace pvgpu {

__global__ void __launch_bounds__(128)
k_shadow_opaque(DScene sc, const SRay* __restrict__ rays, float4* accum, Counters* cnt)
{
    uint2 stack[PV_STACK_SIZE];
    const uint32_t n = cnt->n_shadow;
    unsigned long long tests = 0;
    if (blockIdx.x == 0 && threadIdx.x == 0 && n) atomicAdd(&cnt->shadow_rays, (unsigned long long)n);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const SRay* sp = rays + i;
        const V3 o = ld3(sp->o), d = ld3(sp->d);
        const double depth = sp->depth;
        float f[3];
        trace_shadow<true>(sc, o, d, depth, nullptr, 0u, stack, cnt, f, tests);
        if (f[0] != 0.0f) accum_add(accum, sp->sample, sp->a[0], sp->a[1], sp->a[2], 0.0f);
    }
    for (int off = 16; off > 0; off >>= 1) tests += __shfl_down_sync(0xffffffffu, tests, off);
    if ((threadIdx.x & 31) == 0 && tests) atomicAdd(&cnt->shadow_tests, tests);
}

void launch_shadow_opaque(const DScene& sc, const SRay* rays, uint32_t n_max, float4* accum, Counters* cnt, cudaStream_t st)
{
    k_shadow_opaque<<<grid_for(n_max, 128, 8), 128, 0, st>>>(sc, rays, accum, cnt);
}

}  // namespace pvgpu
