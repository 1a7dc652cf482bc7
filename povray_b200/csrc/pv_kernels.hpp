// Host-callable launchers of the wavefront kernels.  Each heavy kernel lives in its own translation unit
// (k_closest.cu, k_shade.cu, k_shadow_opaque.cu, k_shadow_filter.cu) so that ptxas works on them in parallel.
#pragma once
#include "pv_common.cuh"

namespace pvgpu {

struct WaveCtx {
    float4*   accum;          // per-sample RGBT accumulators
    PRay*     next;           // next wave's queue
    SRay*     shadow;         // shadow-ray queue of this wave
    Counters* cnt;
    unsigned int* n_next;     // ring[wave + 1].n_rays
    unsigned int* n_shadow;   // ring[wave].n_shadow
    uint32_t  cur_cap, next_cap, shadow_cap;
    Cont*     conts;          // continuation records (scenes with a reflection exponent != 1), else nullptr
    uint32_t  cont_cap, cont_base, wave;     // capacity, accumulator slot of record 0, number of the current wave
};

// Where the samples of a batch come from.
struct SampleSource {
    const pvgpu_rect* rects;      // mode 0: pixel centres of rectangles, rect-major (SimpleSamplingM0, tracetask.cpp:438)
    const uint32_t*   rect_off;
    uint32_t          n_rects;
    const double2*    coords;     // mode 1: explicit image-plane coordinates (anti-aliasing passes)
    const uint32_t*   slots;      // accumulator slot per sample, or nullptr: slot = slot_base + sample index
    uint32_t          slot_base;
};

// Anti-aliasing parameters as the kernels use them (TraceTask members, tracetask.h:107-118)
struct AAParams {
    int      method;
    uint32_t depth;               // aaDepth
    double   threshold;           // aaThreshold
    double   jitter_scale;        // jitterScale after the per-method division (tracetask.cpp:526, 611); 0 = no jitter
    float    enc_gamma;           // encoding exponent of the power-law aaGamma curve
    int      neutral;             // aaGamma is neutral (no encoding)
};

// Slot layout of one anti-aliased render call
struct AALayout {
    const pvgpu_rect* rects;
    const uint32_t*   rect_off;   // pixels of the rectangles before rectangle r
    const uint32_t*   frame_off;  // method 1: frame samples (w + h per rectangle) before rectangle r
    const uint32_t*   corner_off; // method 2: corner samples ((w+1)(h+1) per rectangle) before rectangle r
    uint32_t n_rects, n_px, n_frame, n_corner;
    uint32_t s_base;              // first slot of the supersampling sums (method 1) / sample buffers (method 2)
};

int  sm_count();
int  grid_for(uint32_t n, int block, int per_sm);
// grid bound of a traversal kernel: small waves are spread over the warps in chunks of down to 8 rays (chunk_size, pv_traverse.cuh)
#ifndef PV_GRID_PER_RAY
#define PV_GRID_PER_RAY 4u
#endif
inline uint32_t trav_grid_bound(uint32_t n_bound) { return n_bound > 0xFFFFFFFFu / PV_GRID_PER_RAY ? 0xFFFFFFFFu : n_bound * PV_GRID_PER_RAY; }

void launch_container_state(const DScene& sc, uint16_t* out, Counters* cnt, cudaStream_t st);
void launch_primary(const DScene& sc, const SampleSource& src, uint32_t first, uint32_t n, double width, double height,
                    PRay* out, Counters* cnt, float4* accum, cudaStream_t st);
// The wave kernels read their ray counts on the device (wc = this wave's record in the ring, clamped to the queue capacity `cap`);
// n_bound is the host's upper bound of that count and only sizes the grid.
void launch_wave_init(WaveCounts* ring, uint32_t n_slots, uint32_t n0, cudaStream_t st);
void launch_closest(const DScene& sc, const PRay* cur, WaveCounts* wc, uint32_t n_bound, uint32_t cap, HitRec* hits, Counters* cnt, cudaStream_t st);
void launch_shade(const DScene& sc, const PRay* cur, const HitRec* hits, const WaveCounts* wc, uint32_t n_bound, const WaveCtx& ctx, cudaStream_t st);
void launch_shadow_opaque(const DScene& sc, const SRay* rays, WaveCounts* wc, uint32_t n_bound, uint32_t cap, float4* accum, Counters* cnt, cudaStream_t st);
void launch_shadow_filter(const DScene& sc, const SRay* rays, WaveCounts* wc, uint32_t n_bound, uint32_t cap, const PRay* wave, float4* accum, Counters* cnt, cudaStream_t st);
// lean variants (spheres, boxes, planes, meshes only; compiled from the same sources with -DPV_LEAN)
void launch_closest_lean(const DScene& sc, const PRay* cur, WaveCounts* wc, uint32_t n_bound, uint32_t cap, HitRec* hits, Counters* cnt, cudaStream_t st);
void launch_shade_lean(const DScene& sc, const PRay* cur, const HitRec* hits, const WaveCounts* wc, uint32_t n_bound, const WaveCtx& ctx, cudaStream_t st);
void launch_shadow_opaque_lean(const DScene& sc, const SRay* rays, WaveCounts* wc, uint32_t n_bound, uint32_t cap, float4* accum, Counters* cnt, cudaStream_t st);
void launch_shadow_filter_lean(const DScene& sc, const SRay* rays, WaveCounts* wc, uint32_t n_bound, uint32_t cap, const PRay* wave, float4* accum, Counters* cnt, cudaStream_t st);
// quadric-class + CSG variants of the traversal kernels (-DPV_CSG)
void launch_camera_normal_rays(const DScene& sc, const SampleSource& src, uint32_t first, uint32_t n, double width, double height, PRay* rays, cudaStream_t st);
void launch_camera_normal_probe(const DScene& sc, const double* xy, uint32_t n, double width, double height, double* org_dir, cudaStream_t st);
void launch_closest_quartic(const DScene& sc, const PRay* cur, WaveCounts* wc, uint32_t n_bound, uint32_t cap, HitRec* hits, Counters* cnt, cudaStream_t st);
void launch_shadow_opaque_quartic(const DScene& sc, const SRay* rays, WaveCounts* wc, uint32_t n_bound, uint32_t cap, float4* accum, Counters* cnt, cudaStream_t st);
void launch_shadow_filter_quartic(const DScene& sc, const SRay* rays, WaveCounts* wc, uint32_t n_bound, uint32_t cap, const PRay* wave, float4* accum, Counters* cnt, cudaStream_t st);
void launch_shade_csg(const DScene& sc, const PRay* cur, const HitRec* hits, const WaveCounts* wc, uint32_t n_bound, const WaveCtx& ctx, cudaStream_t st);
void launch_closest_csg(const DScene& sc, const PRay* cur, WaveCounts* wc, uint32_t n_bound, uint32_t cap, HitRec* hits, Counters* cnt, cudaStream_t st);
void launch_shadow_opaque_csg(const DScene& sc, const SRay* rays, WaveCounts* wc, uint32_t n_bound, uint32_t cap, float4* accum, Counters* cnt, cudaStream_t st);
void launch_shadow_filter_csg(const DScene& sc, const SRay* rays, WaveCounts* wc, uint32_t n_bound, uint32_t cap, const PRay* wave, float4* accum, Counters* cnt, cudaStream_t st);
uint32_t area_threads();
void launch_shadow_area(const DScene& sc, const SRay* rays, WaveCounts* wc, uint32_t n_bound, uint32_t cap, const PRay* wave, float4* accum, Counters* cnt, float* grid_mem, cudaStream_t st);
// full-material variants (normal perturbation, pigment maps, sky_sphere, fog, area lights; -DPV_FULL)
void launch_shade_full(const DScene& sc, const PRay* cur, const HitRec* hits, const WaveCounts* wc, uint32_t n_bound, const WaveCtx& ctx, cudaStream_t st);
void launch_shadow_filter_full(const DScene& sc, const SRay* rays, WaveCounts* wc, uint32_t n_bound, uint32_t cap, const PRay* wave, float4* accum, Counters* cnt, cudaStream_t st);
void launch_resolve_conts(float4* accum, const Cont* conts, const Counters* cnt, uint32_t cont_cap, uint32_t cont_base, uint32_t wave, cudaStream_t st);
void launch_clear_slots(const SampleSource& src, uint32_t first, uint32_t n, float4* accum, cudaStream_t st);
void launch_probe_rays(const double* org_dir, uint32_t n, PRay* out, cudaStream_t st);
void launch_probe_results(const HitRec* hits, uint32_t n, uint32_t* obj, double* depth, uint32_t* aux, cudaStream_t st);
void launch_aa1_frame_coords(const AALayout& L, double2* coords, cudaStream_t st);
void launch_aa1_candidates(const AALayout& L, const AAParams& aa, const float4* accum, int32_t* s_slot, uint32_t* cand_list, unsigned int* n_cand, cudaStream_t st);
void launch_aa1_sample_coords(const AALayout& L, const AAParams& aa, const uint16_t* hash, const uint32_t* cand_list, uint32_t first, uint32_t n,
                              const double2* offsets, uint32_t n_off, double2* coords, uint32_t* slots, cudaStream_t st);
void launch_aa1_decide(const AALayout& L, const AAParams& aa, const float4* accum, int32_t* s_slot, uint32_t* cand_list, unsigned int* n_cand,
                       float4* out, uint8_t* flag, unsigned int* n_supersampled, cudaStream_t st);
void launch_aa2_corner_coords(const AALayout& L, double2* coords, cudaStream_t st);
void launch_aa2_mark(const AALayout& L, const AAParams& aa, const float4* accum, int32_t* act_idx, uint32_t* act_list, unsigned int* n_active, cudaStream_t st);
void launch_aa2_expand(const AALayout& L, const AAParams& aa, const uint16_t* hash, const float4* accum, const uint32_t* act_list, uint32_t n_active,
                       int target, uint32_t* sampled, double2* coords, uint32_t* slots, unsigned int* n_samples, uint32_t cap, cudaStream_t st);
void launch_aa2_resolve(const AALayout& L, const AAParams& aa, const float4* accum, const int32_t* act_idx, float4* out, cudaStream_t st,
                        const uint32_t* group = nullptr, uint32_t n_group = 0, int skip_active = 0);
void launch_probe_solver(uint32_t n, const int32_t* degree, const int32_t* sturm, const double* epsilon, const double* coeffs,
                         double* roots, int32_t* counts, cudaStream_t st);
void launch_probe_noise(const NoiseTables& nt, uint32_t n, const double* xyz, const int32_t* gen, const int32_t* octaves, double* out, cudaStream_t st);
void launch_init_waves(const NoiseTables& nt, uint32_t n, double* sources, double* freqs, cudaStream_t st);
unsigned long long launch_fp64_peak(double* out, int iters, cudaStream_t st);      // returns the flop the launch performs
void launch_camera_rays(const DScene& sc, const double* xy, uint32_t n, double width, double height, double* org_dir, cudaStream_t st);

}  // namespace pvgpu
