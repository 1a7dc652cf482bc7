// Shadow rays towards area lights: Trace::TraceAreaLightShadowRay / TraceAreaLightSubsetShadowRay (trace.cpp:2078-2271).
// The reference recurses over the light's Area_Size1 x Area_Size2 grid: the four corners of a region are sampled with point-light
// shadow rays (cached in `lightGrid`), the region is split in four while it is coarser than Adaptive_Level or its corner colours
// differ by more than 0.1, and the four results are averaged.  Here every lane runs that recursion as an explicit state machine and
// the warp meets at each sample, because the traversal underneath (trace_shadow) is warp-synchronous.
#define PV_FULL 1          // area lights imply the full-material variant (see PV_FULL_MATERIALS in pv_common.cuh)
#include "pv_shadow.cuh"

namespace pvgpu {

#define PV_AREA_MAX_DEPTH 12          // regions halve per level: grids up to 2048 x 2048

struct AreaFrame {
    int16_t u1, v1, u2, v2;
    uint8_t i;                        // next corner (phase 0) or next child (phase 1)
    uint8_t phase;
    float   col[4][3];                // sample_Colour
};

template <bool ALL_OPAQUE>
__global__ void __launch_bounds__(PV_TRAV_BLOCK, PV_TRAV_MIN_BLOCKS)
k_shadow_area(DScene sc, const SRay* __restrict__ rays, WaveCounts* wc, uint32_t cap, const PRay* __restrict__ wave, float4* accum, Counters* cnt, float* __restrict__ grid_mem, uint32_t n_threads)
{
    uint2 stack_lo[PV_STACK_SIZE];
    const TStack stack{ nullptr, stack_lo, 0 };
    AreaFrame fr[PV_AREA_MAX_DEPTH];
    PV_TREELET_STAGE(sc);
    const uint32_t n = min(wc->n_shadow, cap);
    unsigned long long tests = 0;
    TravCount tc{ 0u, 0u };
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    float* const grid = grid_mem + tid;                 // lightGrid of this thread: cell c, channel k at grid[(3 * c + k) * n_threads]
    const uint32_t cs = chunk_size(n);
    uint32_t i;
    while (next_chunk(&wc->cur_area, n, cs, i)) {
        const SRay s = rays[(i < n) ? i : 0u];
        const pvgpu_light& Lt = sc.lights[s.light];
        bool active = (i < n) && (Lt.flags & PVGPU_LIGHT_AREA);
        const V3 ipoint = ld3(s.o);
        const int n1 = Lt.area_size1, n2 = Lt.area_size2;
        float lcol[3] = { 0.0f, 0.0f, 0.0f }, result[3] = { 0.0f, 0.0f, 0.0f };
        V3 axis1 = ld3(Lt.axis1), axis2 = ld3(Lt.axis2);
        int sp = 0;
        if (active) {
            // the light colour the samples start from: ComputeOneLightRay at the light's centre (s.d, s.depth)
            const double latt = attenuate_light(Lt, ipoint, ld3(s.d), s.depth);
            #pragma unroll
            for (int k = 0; k < 3; k++) lcol[k] = (float)(Lt.colour[k] * latt);
            if (Lt.flags & PVGPU_LIGHT_ORIENT) {            // trace.cpp:2103-2128
                const V3 ldir = ld3(s.d);
                const double axis1_length = length(axis1);
                const V3 temp = (fabs(fabs(ldir.z) - 1.0) < 0.01) ? mk(0.0, 1.0, 0.0) : mk(0.0, 0.0, 1.0);
                axis1 = normalized(cross(ldir, temp));
                axis2 = normalized(cross(ldir, axis1));
                axis1 = axis1 * axis1_length;
                axis2 = axis2 * axis1_length;
            }
            for (int c = 0; c < n1 * n2; c++) grid[(size_t)(3 * c) * n_threads] = __int_as_float(0x7fc00000);     // Invalidate()
            fr[0].u1 = 0; fr[0].v1 = 0; fr[0].u2 = (int16_t)(n1 - 1); fr[0].v2 = (int16_t)(n2 - 1); fr[0].i = 0; fr[0].phase = 0;
        }
        for (;;) {
            // advance this lane's recursion until it needs a shadow ray (need) or is finished (!active)
            bool need = false;
            int su = 0, sv = 0;
            while (active && !need) {
                AreaFrame& F = fr[sp];
                bool finish = false;
                if (F.phase == 0) {
                    if (F.i < 4) {
                        su = (F.i == 1 || F.i == 3) ? F.u2 : F.u1;
                        sv = (F.i >= 2) ? F.v2 : F.v1;
                        const size_t cell = (size_t)(3 * (su * n2 + sv)) * n_threads;
                        const float r = grid[cell];
                        if (r == r) { F.col[F.i][0] = r; F.col[F.i][1] = grid[cell + n_threads]; F.col[F.i][2] = grid[cell + 2 * (size_t)n_threads]; F.i++; }
                        else need = true;
                    } else {
                        auto dist = [&](int a, int b) { return fabsf(F.col[a][0] - F.col[b][0]) + fabsf(F.col[a][1] - F.col[b][1]) + fabsf(F.col[a][2] - F.col[b][2]); };
                        if (((F.u2 - F.u1 > 1) || (F.v2 - F.v1 > 1)) && (sp + 1 < PV_AREA_MAX_DEPTH) &&
                            ((sp < Lt.adaptive_level) || ((double)dist(0, 1) > 0.1) || ((double)dist(1, 3) > 0.1) || ((double)dist(3, 2) > 0.1) || ((double)dist(2, 0) > 0.1))) {
                            F.phase = 1; F.i = 0;
                        } else finish = true;
                    }
                } else {
                    if (F.i < 4) {
                        AreaFrame& C = fr[sp + 1];
                        const int lo_u = (int)floor((F.u1 + F.u2) / 2.0), hi_u = (int)ceil((F.u1 + F.u2) / 2.0);
                        const int lo_v = (int)floor((F.v1 + F.v2) / 2.0), hi_v = (int)ceil((F.v1 + F.v2) / 2.0);
                        C.u1 = (int16_t)((F.i & 1) ? hi_u : F.u1); C.u2 = (int16_t)((F.i & 1) ? F.u2 : lo_u);
                        C.v1 = (int16_t)((F.i & 2) ? hi_v : F.v1); C.v2 = (int16_t)((F.i & 2) ? F.v2 : lo_v);
                        C.i = 0; C.phase = 0;
                        sp++;
                    } else finish = true;
                }
                if (finish) {
                    float avg[3];
                    #pragma unroll
                    for (int k = 0; k < 3; k++) avg[k] = (((F.col[0][k] + F.col[1][k]) + F.col[2][k]) + F.col[3][k]) * 0.25f;
                    if (sp == 0) { result[0] = avg[0]; result[1] = avg[1]; result[2] = avg[2]; active = false; }
                    else {
                        sp--;
                        AreaFrame& P = fr[sp];
                        P.col[P.i][0] = avg[0]; P.col[P.i][1] = avg[1]; P.col[P.i][2] = avg[2];
                        P.i++;
                    }
                }
            }
            if (!vote_any(need)) break;
            // the sample's light ray (trace.cpp:2163-2213)
            V3 ldir = mk(0.0, 0.0, 1.0);
            double ldepth = 1.0;
            if (need) {
                double ju = (double)su, jv = (double)sv;
                V3 j1, j2;
                if (Lt.flags & PVGPU_LIGHT_CIRCULAR) {
                    ju = ju / (n1 - 1) - 0.5 + 0.001;
                    jv = jv / (n2 - 1) - 0.5 + 0.001;
                    double scale = (fabs(ju) > fabs(jv)) ? fabs(ju) : fabs(jv);
                    scale /= sqrt(ju * ju + jv * jv);
                    ju *= scale; jv *= scale;
                    j1 = axis1 * ju; j2 = axis2 * jv;
                } else {
                    j1 = (n1 > 1) ? axis1 * (ju / (double)(n1 - 1) - 0.5) : mk(0.0, 0.0, 0.0);
                    j2 = (n2 > 1) ? axis2 * (jv / (double)(n2 - 1) - 0.5) : mk(0.0, 0.0, 0.0);
                }
                light_ray(Lt, ipoint, ldir, ldepth, j1 + j2);
            }
            float f[3];
            trace_shadow<ALL_OPAQUE>(need, sc, ipoint, ldir, ldepth, wave, s.parent, stack, cnt, f, tests, tc);
            if (need) {
                AreaFrame& F = fr[sp];
                const size_t cell = (size_t)(3 * (su * n2 + sv)) * n_threads;
                #pragma unroll
                for (int k = 0; k < 3; k++) { const float c = lcol[k] * f[k]; F.col[F.i][k] = c; grid[cell + (size_t)k * n_threads] = c; }
                F.i++;
            }
        }
        if ((i < n) && (Lt.flags & PVGPU_LIGHT_AREA))
            accum_add(accum, s.sample, s.a[0] * result[0], s.a[1] * result[1], s.a[2] * result[2], 0.0f);
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

// grid_mem: 3 floats x area_grid_max cells for every thread of the launch (area_threads() threads)
uint32_t area_threads() { return (uint32_t)(sm_count() * PV_TRAV_MIN_BLOCKS * PV_TRAV_BLOCK); }

void launch_shadow_area(const DScene& sc, const SRay* rays, WaveCounts* wc, uint32_t n_bound, uint32_t cap, const PRay* wave, float4* accum, Counters* cnt, float* grid_mem, cudaStream_t st)
{
    const int blocks = grid_for(trav_grid_bound(n_bound), PV_TRAV_BLOCK, PV_TRAV_MIN_BLOCKS);
    if (sc.all_opaque) k_shadow_area<true><<<blocks, PV_TRAV_BLOCK, PV_TREELET_SMEM, st>>>(sc, rays, wc, cap, wave, accum, cnt, grid_mem, area_threads());
    else k_shadow_area<false><<<blocks, PV_TRAV_BLOCK, PV_TREELET_SMEM, st>>>(sc, rays, wc, cap, wave, accum, cnt, grid_mem, area_threads());
}

}  // namespace pvgpu
