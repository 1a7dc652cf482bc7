// Shadow-ray kernel for scenes with transparent shadow casters: closest-hit loop towards the light, the light
// colour is filtered through every transparent blocker (ComputeShadowColour, trace.cpp:2274-2439).
#include "pv_shadow.cuh"
#include <algorithm>

namespace pvgpu {

__global__ void __launch_bounds__(PV_TRAV_BLOCK, PV_TRAV_MIN_BLOCKS)
PV_VARIANT(k_shadow_filter)(DScene sc, const SRay* __restrict__ rays, WaveCounts* wc, uint32_t cap, uint32_t max_chunks, const PRay* __restrict__ wave, float4* accum, Counters* cnt)
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
        const SRay s = load_cs(rays + (alive ? i : 0u));
#if PV_HEAVY
        if (sc.has_area_lights && (sc.lights[s.light].flags & PVGPU_LIGHT_AREA)) alive = false;        // served by k_shadow_area
#endif
        float f[3];
        trace_shadow<false>(alive, sc, ld3(s.o), ld3(s.d), s.depth, wave, s.parent, stack, cnt, f, tests, tc);
        if (alive) accum_add(accum, s.sample, s.a[0] * f[0], s.a[1] * f[1], s.a[2] * f[2], 0.0f);
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

// The shadow kernels run on a low-priority stream next to k_closest / k_shade of the following wave, which are on the frame's
// critical path.  A small wave does not spread over the whole machine: the grid is sized so that every warp has
// PV_SHADOW_CHUNKS_PER_WARP chunks to work through, which leaves block slots for the critical kernels; large waves fill the machine
// as before.  (Tried and rejected: rounds of short-lived blocks, which would let the block scheduler's stream priorities arbitrate,
// lose more in the drain of every round - config 2: 10.0 -> 17.2 ms.)
#ifndef PV_SHADOW_CHUNKS_PER_WARP
#define PV_SHADOW_CHUNKS_PER_WARP 2
#endif
void PV_VARIANT(launch_shadow_filter)(const DScene& sc, const SRay* rays, WaveCounts* wc, uint32_t n_bound, uint32_t cap, const PRay* wave, float4* accum, Counters* cnt, cudaStream_t st)
{
    const int grid = grid_for(n_bound / PV_SHADOW_CHUNKS_PER_WARP + 1u, PV_TRAV_BLOCK, PV_TRAV_MIN_BLOCKS);
    PV_VARIANT(k_shadow_filter)<<<grid, PV_TRAV_BLOCK, PV_TREELET_SMEM, st>>>(sc, rays, wc, cap, 0u, wave, accum, cnt);
}

}  // namespace pvgpu
