// The shading kernel of the wavefront: for every ray of the wave and its closest hit, Trace::ComputeTextureColour /
// ComputeLightedTexture (trace.cpp:457-1179) or ComputeSky (trace.cpp:2769-2890); emits shadow rays and the
// reflection / refraction rays of the next wave.
#include "pv_shade.cuh"

#ifndef PV_SHADE_MIN_BLOCKS
#define PV_SHADE_MIN_BLOCKS 4       // 128 registers: measured best (2 / 3 / 4 / 6 CTAs per SM: 1.63 / 1.5 / 1.28 / 1.57 ms on config 2)
#endif

namespace pvgpu {

#if PV_FULL_MATERIALS
// A ray that travelled through fog (trace.cpp:207-216): colour' = sum_col + sum_att * colour, so the fog's own light is added
// here and everything the ray still collects is weighted by sum_att.  Out of line, so that scenes without fog keep the ray record
// a read-only copy the compiler may re-read instead of spilling.
static __device__ __noinline__ void shade_fogged(const DScene& sc, const PRay& ray0, uint32_t i, const HitRec& h, WaveCtx& ctx)
{
    PRay ray = ray0;
    float sum_att[3], sum_col[3];
    compute_fog(sc, ld3(ray.o), ld3(ray.d), (h.obj == PV_HIT_MISS) ? PV_BOUND_HUGE : h.depth, sum_att, sum_col);
    accum_add(ctx.accum, ray.sample, ray.w[0] * sum_col[0], ray.w[1] * sum_col[1], ray.w[2] * sum_col[2], 0.0f);
    ray.w[0] *= sum_att[0]; ray.w[1] *= sum_att[1]; ray.w[2] *= sum_att[2];
    ray.wt *= greyscale(sum_att);
    if (h.obj == PV_HIT_MISS) {
        float col[3], transm;
        compute_sky(sc, ray, col, transm);
        accum_add(ctx.accum, ray.sample, ray.w[0] * col[0], ray.w[1] * col[1], ray.w[2] * col[2], ray.wt * transm);
        return;
    }
    Hit hit;
    hit.depth = h.depth; hit.ip = mk(h.ip[0], h.ip[1], h.ip[2]); hit.obj = h.obj; hit.aux = h.aux; hit.csg = h.csg;
    shade_hit(sc, ray, i, hit, ctx);
}
#endif

__global__ void __launch_bounds__(128, PV_SHADE_MIN_BLOCKS)
PV_VARIANT(k_shade)(DScene sc, const PRay* __restrict__ cur, const HitRec* __restrict__ hits, const WaveCounts* wc, WaveCtx ctx)
{
    const uint32_t n = min(wc->n_rays, ctx.cur_cap);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const HitRec h = load_cs(hits + i);
        if (h.obj == PV_HIT_STOPPED) continue;
        const PRay ray = cur[i];
#if PV_FULL_MATERIALS
        if (sc.n_fogs && (sc.g.quality_flags & PVGPU_Q_MEDIA) && ray_is_hollow(sc, ray)) { shade_fogged(sc, ray, i, h, ctx); continue; }
#endif
        if (h.obj == PV_HIT_MISS) {
            float col[3], transm;
            compute_sky(sc, ray, col, transm);
            accum_add(ctx.accum, ray.sample, ray.w[0] * col[0], ray.w[1] * col[1], ray.w[2] * col[2], ray.wt * transm);
            continue;
        }
        Hit hit;
        hit.depth = h.depth; hit.ip = mk(h.ip[0], h.ip[1], h.ip[2]); hit.obj = h.obj; hit.aux = h.aux; hit.csg = h.csg;
        shade_hit(sc, ray, i, hit, ctx);
    }
}

#ifdef PV_FULL
// camera { normal { ... } }: the tail of TracePixel::CreateCameraRay (tracepixel.cpp:917-924) as a pass over the primary rays - the
// direction is perturbed like a surface normal at the point (x0, y0, 0) of the image plane and normalised again.  Perspective and
// orthographic cameras (x0 = x / width - 0.5, y0 = 0.5 - y / height); lives here because Perturb_Normal is shading code.
__device__ __forceinline__ V3 camera_normal_dir(const DScene& sc, const V3& d, double x, double y, double width, double height)
{
    const double x0 = x / width - 0.5, y0 = 0.5 - y / height;
    // (d has been normalised once by camera_ray, like the reference does before Perturb_Normal)
    return normalized(perturb_normal(sc, (int32_t)sc.cam.reserved - 1, d, mk(x0, y0, 0.0)));
}
__global__ void k_camera_normal_rays(DScene sc, SampleSource src, uint32_t first, uint32_t n, double width, double height, PRay* rays)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        double x, y;
        uint32_t slot;
        if (src.coords) { const double2 c = src.coords[first + i]; x = c.x; y = c.y; }
        else sample_xy(src.rects, src.rect_off, src.n_rects, first + i, x, y, slot);
        PRay& r = rays[i];
        if (r.flags & PV_RAY_DEAD) continue;
        const V3 d = camera_normal_dir(sc, ld3(r.d), x, y, width, height);
        r.d[0] = d.x; r.d[1] = d.y; r.d[2] = d.z;
    }
}
__global__ void k_camera_normal_probe(DScene sc, const double* xy, uint32_t n, double width, double height, double* org_dir)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        double* r = org_dir + 6 * (size_t)i;
        if (r[3] == 0.0 && r[4] == 0.0 && r[5] == 0.0) continue;       // no ray for this pixel
        const V3 d = camera_normal_dir(sc, mk(r[3], r[4], r[5]), xy[2 * (size_t)i], xy[2 * (size_t)i + 1], width, height);
        r[3] = d.x; r[4] = d.y; r[5] = d.z;
    }
}
void launch_camera_normal_rays(const DScene& sc, const SampleSource& src, uint32_t first, uint32_t n, double width, double height, PRay* rays, cudaStream_t st)
{
    k_camera_normal_rays<<<grid_for(n, 128, 8), 128, 0, st>>>(sc, src, first, n, width, height, rays);
}
void launch_camera_normal_probe(const DScene& sc, const double* xy, uint32_t n, double width, double height, double* org_dir, cudaStream_t st)
{
    k_camera_normal_probe<<<grid_for(n, 128, 8), 128, 0, st>>>(sc, xy, n, width, height, org_dir);
}
#endif

void PV_VARIANT(launch_shade)(const DScene& sc, const PRay* cur, const HitRec* hits, const WaveCounts* wc, uint32_t n_bound, const WaveCtx& ctx, cudaStream_t st)
{
    PV_VARIANT(k_shade)<<<grid_for(n_bound, 128, 8), 128, 0, st>>>(sc, cur, hits, wc, ctx);
}

}  // namespace pvgpu
