// Host-callable launchers of the wavefront kernels.  Each heavy kernel lives in its own translation unit
// (k_closest.cu, k_shade.cu, k_shadow_opaque.cu, k_shadow_filter.cu) so that ptxas works on them in parallel.
#pragma once
#include "pv_common.cuh"

namespace pvgpu {

struct WaveCtx {
    float4*   accum;          // per-sample RGBT accumulators
    PRay*     next;           // next wave's queue
    SRay*     shadow;         // shadow-ray queue of the current chunk
    Counters* cnt;
    uint32_t  next_cap, shadow_cap;
};

// Where the samples of a batch come from.
struct SampleSource {
    const pvgpu_rect* rects;      // mode 0: pixel centres of rectangles, rect-major (SimpleSamplingM0, tracetask.cpp:438)
    const uint32_t*   rect_off;
    uint32_t          n_rects;
    const double2*    coords;     // mode 1: explicit image-plane coordinates (anti-aliasing passes), slot = slots[i] or first + i
    const uint32_t*   slots;
};

int  sm_count();
int  grid_for(uint32_t n, int block, int per_sm);

void launch_container_state(const DScene& sc, uint16_t* out, Counters* cnt, cudaStream_t st);
void launch_primary(const DScene& sc, const SampleSource& src, uint32_t first, uint32_t n, double width, double height,
                    PRay* out, Counters* cnt, cudaStream_t st);
void launch_closest(const DScene& sc, const PRay* cur, uint32_t n, HitRec* hits, Counters* cnt, cudaStream_t st);
void launch_shade(const DScene& sc, const PRay* cur, const HitRec* hits, uint32_t n, const WaveCtx& ctx, cudaStream_t st);
// n_max: upper bound of the shadow-ray count (the exact count is read from cnt->n_shadow on the device)
void launch_shadow_opaque(const DScene& sc, const SRay* rays, uint32_t n_max, float4* accum, Counters* cnt, cudaStream_t st);
void launch_shadow_filter(const DScene& sc, const SRay* rays, const PRay* wave, uint32_t n_max, float4* accum, Counters* cnt, cudaStream_t st);
void launch_probe_rays(const double* org_dir, uint32_t n, PRay* out, cudaStream_t st);
void launch_probe_results(const HitRec* hits, uint32_t n, uint32_t* obj, double* depth, uint32_t* aux, cudaStream_t st);
void launch_camera_rays(const DScene& sc, const double* xy, uint32_t n, double width, double height, double* org_dir, cudaStream_t st);

}  // namespace pvgpu
