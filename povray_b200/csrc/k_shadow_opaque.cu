// Shadow-ray kernel for scenes in which every shadow caster is opaque: any-hit search, no filtering.
#include "pv_shadow.cuh"
#include <algorithm>

namespace pvgpu {

__global__ void __launch_bounds__(PV_TRAV_BLOCK, PV_TRAV_MIN_BLOCKS)
PV_VARIANT(k_shadow_opaque)(DScene sc, const SRay* __restrict__ rays, WaveCounts* wc, uint32_t cap, uint32_t max_chunks, float4* accum, Counters* cnt)
{
#if PV_SSTACK > 0
    __shared__ uint2 stack_sh[PV_SSTACK * PV_TRAV_BLOCK];
    uint2 stack_lo[PV_STACK_SIZE - PV_SSTACK];
    const TStack stack{ stack_sh + threadIdx.x, stack_lo, PV_SSTACK };
#else
    uint2 stack_lo[PV_STACK_SIZE];
    const TStack stack{ nullptr, stack_lo, 0 };
#endif
    // the queue never holds more than `cap` records: push_shadow drops (and flags) what does not fit, the count keeps running
    PV_TREELET_STAGE(sc);
    const uint32_t n = min(wc->n_shadow, cap);
    unsigned long long tests = 0;
    TravCount tc{ 0u, 0u };
    // (max_chunks: a warp leaves after that many chunks - see the launcher; 0 = stay until the queue is exhausted)
    if (blockIdx.x == 0 && threadIdx.x == 0 && n && atomicExch(&wc->counted, 1u) == 0u) atomicAdd(&cnt->shadow_rays, (unsigned long long)n);
    const uint32_t cs = chunk_size(n);
    uint32_t i;
    uint32_t taken = 0;
    while ((max_chunks == 0u || taken++ < max_chunks) && next_chunk(&wc->cur_shadow, n, cs, i)) {
        bool alive = i < n;
        const SRay sr = load_cs(rays + (alive ? i : 0u));
        const SRay* sp = &sr;
#if PV_HEAVY
        if (sc.has_area_lights && (sc.lights[sp->light].flags & PVGPU_LIGHT_AREA)) alive = false;      // served by k_shadow_area
#endif
        const V3 o = ld3(sp->o), d = ld3(sp->d);
        const double depth = sp->depth;
        float f[3];
        trace_shadow<true>(alive, sc, o, d, depth, nullptr, 0u, stack, cnt, f, tests, tc);
        if (alive && f[0] != 0.0f) accum_add(accum, sp->sample, sp->a[0], sp->a[1], sp->a[2], 0.0f);
    }
    unsigned long long n_nodes = tc.nodes, n_prims = tc.prims;
    for (int off = 16; off > 0; off >>= 1) {
        tests += __shfl_down_sync(0xffffffffu, tests, off);
        n_nodes += __shfl_down_sync(0xffffffffu, n_nodes, off);
        n_prims += __shfl_down_sync(0xffffffffu, n_prims, off);
    }
    if ((threadIdx.x & 31) == 0) {
        if (tests) atomicAdd(&cnt->shadow_tests, tests);
        if (n_nodes) atomicAdd(&cnt->node_tests[1], n_nodes);
        if (n_prims) atomicAdd(&cnt->prim_tests[1], n_prims);
    }
}

#ifndef PV_SHADOW_CHUNKS_PER_WARP
#define PV_SHADOW_CHUNKS_PER_WARP 2
#endif
// grid sized for PV_SHADOW_CHUNKS_PER_WARP chunks per warp: see launch_shadow_filter (k_shadow_filter.cu)
void PV_VARIANT(launch_shadow_opaque)(const DScene& sc, const SRay* rays, WaveCounts* wc, uint32_t n_bound, uint32_t cap, float4* accum, Counters* cnt, cudaStream_t st)
{
    const int grid = grid_for(n_bound / PV_SHADOW_CHUNKS_PER_WARP + 1u, PV_TRAV_BLOCK, PV_TRAV_MIN_BLOCKS);
    PV_VARIANT(k_shadow_opaque)<<<grid, PV_TRAV_BLOCK, PV_TREELET_SMEM, st>>>(sc, rays, wc, cap, 0u, accum, cnt);
}

}  // namespace pvgpu
