// Camera-ray generation and the closest-hit kernel (Trace::TraceRay's level / ADC test and
// Trace::FindIntersection, trace.cpp:135-228, 285-344) of the wavefront, plus the small kernels of the
// ray-level harness.
#include "pv_traverse.cuh"
#include "pv_kernels.hpp"

namespace pvgpu {

#if !PV_SECONDARY_TU

int sm_count()
{
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

int grid_for(uint32_t n, int block, int per_sm)
{
    long long blocks = ((long long)n + block - 1) / block;
    long long cap = (long long)sm_count() * per_sm;      // a multiple of the SM count; grid-stride loops cover the rest
    return (int)(blocks < 1 ? 1 : (blocks < cap ? blocks : cap));
}

// TracePixel::InitRayContainerState (tracepixel.cpp:929-1006): interiors of all objects containing `p`.
__device__ inline void container_state(const DScene& sc, const V3& p, uint16_t* out, uint32_t& n, TStack stack, unsigned int* overflow)
{
    n = 0;
    auto inside_bbox = [&](const float* lo, const float* size) {       // Inside_BBox (boundingbox.h:139-155)
        if (p.x < (double)lo[0] || p.y < (double)lo[1] || p.z < (double)lo[2]) return false;
        if (p.x > (double)lo[0] + (double)size[0] || p.y > (double)lo[1] + (double)size[1] || p.z > (double)lo[2] + (double)size[2]) return false;
        return true;
    };
    auto test_object = [&](uint32_t idx, int sp) {
        const pvgpu_object& o = sc.objs[idx];
        if (o.interior >= 0 && inside_object(sc, idx, p, stack, sp, false)) {
            if (n < PV_MAX_INTERIORS) out[n++] = (uint16_t)o.interior; else atomicOr(overflow, 4u);
        }
    };
    if (!sc.use_tree) {
        for (uint32_t i = 0; i < sc.n_frame; i++) {
            const pvgpu_object& o = sc.objs[sc.frame[i]];
            if (o.interior >= 0 && inside_bbox(o.bbox, o.bbox + 3)) test_object(sc.frame[i], 0);
        }
        return;
    }
    // InitRayContainerStateTree: children visited in order (pre-order), so push them reversed
    int sp = 0;
    stack.set(sp++, make_uint2(0u, 0u));
    while (sp > 0) {
        const uint32_t ni = stack.get(--sp).y;
        const pvgpu_node nd = sc.nodes[ni];
        if (!inside_bbox(nd.lo, nd.size)) continue;
        if (nd.count == 0) test_object(nd.first, sp);
        else for (uint32_t c = nd.count; c-- > 0 && sp < PV_STACK_SIZE;) stack.set(sp++, make_uint2(0u, nd.first + c));
    }
}

__global__ void k_container_state(DScene sc, uint16_t* out, Counters* cnt)
{
    PV_TREELET_STAGE(sc);
    if (threadIdx.x || blockIdx.x) return;
    uint2 stack_mem[PV_STACK_SIZE];
    const TStack stack{ nullptr, stack_mem, 0 };
    uint16_t ints[PV_MAX_INTERIORS];
    uint32_t n;
    container_state(sc, ld3(sc.cam.location), ints, n, stack, &cnt->overflow);
    for (uint32_t i = 0; i < n; i++) out[i] = ints[i];
    out[PV_MAX_INTERIORS] = (uint16_t)n;
}

// TracePixel::CreateCameraRay (tracepixel.cpp:341-674, 917-927): every camera type except mesh_camera / user_defined.
// Returns false where the reference traces no ray (fisheye / omnimax pixels outside the image circle).
__device__ inline bool camera_ray(const DScene& sc, double x, double y, double width, double height, V3& o, V3& d)
{
    const pvgpu_camera& cam = sc.cam;
    const V3 loc = ld3(cam.location);
    o = loc;
    if (cam.type <= PVGPU_CAMERA_ORTHOGRAPHIC) {
        const double x0 = x / width - 0.5;
        const double y0 = 0.5 - y / height;
        const V3 dirv = ld3(cam.direction), right = ld3(cam.right), up = ld3(cam.up);
        if (cam.type == PVGPU_CAMERA_ORTHOGRAPHIC) {
            d = dirv;
            o = (loc + x0 * right) + y0 * up;
        } else d = (dirv + x0 * right) + y0 * up;
        d = normalized(d);
        return true;
    }
#ifndef PV_LEAN
    const V3 right = ld3(sc.cam_right), up = ld3(sc.cam_up), dirv = ld3(sc.cam_dir);
    const double pi = 3.1415926535897932384626, pi_180 = 0.01745329251994329576, pi_360 = 0.00872664625997164788, pi_2 = 1.57079632679489661923;
    double x0, y0, cx, sx, cy, sy;
    switch (cam.type) {
        case PVGPU_CAMERA_FISHEYE:
        case PVGPU_CAMERA_OMNIMAX: {
            x0 = 2.0 * (x / width - 0.5);
            y0 = 2.0 * (0.5 - y / height);
            if (cam.type == PVGPU_CAMERA_FISHEYE) { x0 *= sc.cam_len_right; y0 *= sc.cam_len_up; }
            else if (sc.cam_aspect > 1.0) {
                if (sc.cam_aspect > 1.283458) { x0 *= sc.cam_aspect / 1.283458; y0 = (y0 - 1.0) / 1.283458 + 1.0; }
                else y0 = (y0 - 1.0) / sc.cam_aspect + 1.0;
            } else y0 /= sc.cam_aspect;
            const double rad = sqrt(x0 * x0 + y0 * y0);
            if (rad > 1.0) return false;
            double phi;
            if (rad == 0.0) phi = 0.0;
            else if (x0 < 0.0) phi = pi - asin(y0 / rad);
            else phi = asin(y0 / rad);
            x0 = phi;
            if (cam.type == PVGPU_CAMERA_FISHEYE) y0 = rad * sc.cam_angle * pi_360;
            else y0 = 1.411269 * rad - 0.09439 * rad * rad * rad + 0.25674 * rad * rad * rad * rad * rad;
            cx = cos(x0); sx = sin(x0); cy = cos(y0); sy = sin(y0);
            if (cam.type == PVGPU_CAMERA_OMNIMAX && (sx * sy < tan(135.0 * pi_180) * cy)) return false;
            d = ((cx * sy) * right + (sx * sy) * up) + cy * dirv;
            break;
        }
        case PVGPU_CAMERA_PANORAMIC: {
            x0 = x / width;
            y0 = 2.0 * (0.5 - y / height);
            x0 = (1.0 - x0) * pi;
            y0 = pi_2 * y0;
            cx = cos(x0); sx = sin(x0);
            double ty;
            if (fabs(pi_2 - fabs(y0)) < PV_EPSILON) ty = (y0 > 0.0) ? PV_BOUND_HUGE : -PV_BOUND_HUGE;
            else ty = tan(y0);
            d = (cx * right + ty * up) + sx * dirv;
            break;
        }
        case PVGPU_CAMERA_ULTRA_WIDE_ANGLE:
            x0 = x / width - 0.5; y0 = 0.5 - y / height;
            x0 *= sc.cam_angle * pi_180;
            y0 *= sc.cam_angle * sc.cam_aspect * pi_180;
            cx = cos(x0); sx = sin(x0); cy = cos(y0); sy = sin(y0);
            d = (sx * right + sy * up) + (cx * cy) * dirv;
            break;
        case PVGPU_CAMERA_CYL_1:
        case PVGPU_CAMERA_CYL_3:
            x0 = x / width - 0.5; y0 = 0.5 - y / height;
            x0 *= sc.cam_angle * pi_180;
            y0 *= sc.cam_aspect;
            cx = cos(x0); sx = sin(x0);
            if (cam.type == PVGPU_CAMERA_CYL_1) d = (sx * right + y0 * up) + cx * dirv;
            else { d = sx * right + cx * dirv; o = loc + y0 * up; }
            break;
        case PVGPU_CAMERA_CYL_2:
        case PVGPU_CAMERA_CYL_4:
            x0 = x / width - 0.5; y0 = 0.5 - y / height;
            y0 *= sc.cam_angle * pi_180;
            x0 *= sc.cam_aspect;
            cy = cos(y0); sy = sin(y0);
            if (cam.type == PVGPU_CAMERA_CYL_2) d = (x0 * right + sy * up) + cy * dirv;
            else { d = sy * up + cy * dirv; o = loc + x0 * right; }
            break;
        default: {      // PVGPU_CAMERA_SPHERICAL: two axis rotations (Compute_Axis_Rotation_Transform, matrix.cpp:825-850)
            x0 = x / width - 0.5; y0 = 0.5 - y / height;
            y0 *= (sc.cam_v_angle / 360) * 6.283185307179586476925286766560;
            x0 *= (sc.cam_h_angle / 360) * 6.283185307179586476925286766560;
            auto rotate = [](const V3& axis, double angle, const V3& p) {
                const V3 a = normalized(axis);
                const double cosx = cos(angle), sinx = sin(angle);
                const double m00 = a.x * a.x + cosx * (1.0 - a.x * a.x), m01 = a.x * a.y * (1.0 - cosx) + a.z * sinx, m02 = a.x * a.z * (1.0 - cosx) - a.y * sinx;
                const double m10 = a.x * a.y * (1.0 - cosx) - a.z * sinx, m11 = a.y * a.y + cosx * (1.0 - a.y * a.y), m12 = a.y * a.z * (1.0 - cosx) + a.x * sinx;
                const double m20 = a.x * a.z * (1.0 - cosx) + a.y * sinx, m21 = a.y * a.z * (1.0 - cosx) - a.x * sinx, m22 = a.z * a.z + cosx * (1.0 - a.z * a.z);
                // MTransPoint with a zero translation row (matrix.cpp:415-430)
                return mk(p.x * m00 + p.y * m10 + p.z * m20 + 0.0, p.x * m01 + p.y * m11 + p.z * m21 + 0.0, p.x * m02 + p.y * m12 + p.z * m22 + 0.0);
            };
            const V3 v1 = rotate(right, -y0, dirv);
            d = rotate(up, x0, v1);
            break;
        }
    }
    d = normalized(d);
#endif
    return true;
}

// TracePixel::operator() (tracepixel.cpp:311-339): one new ticket + camera ray per sample.
__global__ void __launch_bounds__(256)
k_primary(DScene sc, SampleSource src, uint32_t first, uint32_t n, double width, double height, PRay* out, Counters* cnt, float4* accum)
{
    uint2 stack_mem[PV_STACK_SIZE];
    const TStack stack{ nullptr, stack_mem, 0 };
    PV_TREELET_STAGE(sc);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        double x, y;
        uint32_t slot = src.slot_base + first + i;
        if (src.coords) {
            const double2 c = src.coords[first + i];
            x = c.x; y = c.y;
            if (src.slots) slot = src.slots[first + i];
        } else {
            sample_xy(src.rects, src.rect_off, src.n_rects, first + i, x, y, slot);
            slot += src.slot_base;
        }
        V3 o = mk(0.0, 0.0, 0.0), d = mk(0.0, 0.0, 1.0);
        const bool have_ray = camera_ray(sc, x, y, width, height, o, d);
        PRay r;
        r.o[0] = o.x; r.o[1] = o.y; r.o[2] = o.z;
        r.d[0] = d.x; r.d[1] = d.y; r.d[2] = d.z;
        r.w[0] = r.w[1] = r.w[2] = 1.0f;
        r.wt = 1.0f;
        r.adc = 1.0f;
        r.sample = slot;
        r.level = 0;
        r.flags = (uint8_t)(PV_RAY_PRIMARY | (sc.g.output_alpha ? PV_RAY_ALPHA_BG : 0));
        if (!have_ray) {        // TracePixel::operator(): numTraced == 0 -> colour stays black, transm = 1 (tracepixel.cpp:332-335)
            r.flags |= PV_RAY_DEAD;
            atomicAdd(reinterpret_cast<float*>(accum + slot) + 3, 1.0f);
        }
        r.n_int = (uint8_t)sc.n_cam_interiors;
        r.pad = 0;
        #pragma unroll
        for (int k = 0; k < PV_MAX_INTERIORS; k++) r.interiors[k] = sc.cam_interiors[k];
        if ((sc.cam.type == PVGPU_CAMERA_ORTHOGRAPHIC || sc.cam.type == PVGPU_CAMERA_CYL_3 || sc.cam.type == PVGPU_CAMERA_CYL_4) && sc.has_interiors) {
            // InitRayContainerState(ray, true): recomputed per ray when the origin moves with the pixel
            uint32_t nci;
            container_state(sc, o, r.interiors, nci, stack, &cnt->overflow);
            r.n_int = (uint8_t)nci;
        }
        store_cs(out + i, r);
    }
}

// Start of a batch: the ring of per-wave records is cleared and wave 0 gets its ray count.
__global__ void k_wave_init(WaveCounts* ring, uint32_t n_slots, uint32_t n0)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_slots * (uint32_t)(sizeof(WaveCounts) / 4); i += gridDim.x * blockDim.x)
        reinterpret_cast<unsigned int*>(ring)[i] = (i == 0u) ? n0 : 0u;
}

// Clears the accumulator slots of samples [first, first + n) of a sample source (retry of a batch after a queue overflow): the slots
// are recomputed like k_primary does, because a rectangle's samples are dealt to its pixels in 8 x 4 blocks (sample_xy).
__global__ void k_clear_slots(SampleSource src, uint32_t first, uint32_t n, float4* accum)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        uint32_t slot = src.slot_base + first + i;
        if (src.coords) { if (src.slots) slot = src.slots[first + i]; }
        else {
            double x, y;
            sample_xy(src.rects, src.rect_off, src.n_rects, first + i, x, y, slot);
            slot += src.slot_base;
        }
        accum[slot] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
}

// Continuation records of wave `wave` (see Cont in pv_common.cuh): parent += w * Pow(slot colour, exponent), colour.h Pow = powf per channel.
__global__ void k_resolve_conts(float4* accum, const Cont* __restrict__ conts, const Counters* cnt, uint32_t cont_cap, uint32_t cont_base, uint32_t wave)
{
    const uint32_t n = min(cnt->n_cont, cont_cap);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const Cont c = conts[i];
        if (c.wave != wave) continue;
        const float4 v = accum[cont_base + i];
        float* a = reinterpret_cast<float*>(accum + c.parent);
        const float r = c.w[0] * powf(v.x, c.exponent), g = c.w[1] * powf(v.y, c.exponent), b = c.w[2] * powf(v.z, c.exponent);
        if (r != 0.0f) atomicAdd(a + 0, r);
        if (g != 0.0f) atomicAdd(a + 1, g);
        if (b != 0.0f) atomicAdd(a + 2, b);
    }
}

#endif  // !PV_LEAN

// Trace::TraceRay's entry (trace.cpp:142-160) + FindIntersection for every ray of the wave.
__global__ void __launch_bounds__(PV_TRAV_BLOCK, PV_TRAV_MIN_BLOCKS)
PV_VARIANT(k_closest)(DScene sc, const PRay* __restrict__ cur, WaveCounts* wc, uint32_t cap, HitRec* __restrict__ hits, Counters* cnt)
{
#if PV_SSTACK > 0
    __shared__ uint2 stack_sh[PV_SSTACK * PV_TRAV_BLOCK];
    uint2 stack_lo[PV_STACK_SIZE - PV_SSTACK];
    const TStack stack{ stack_sh + threadIdx.x, stack_lo, PV_SSTACK };
#else
    uint2 stack_lo[PV_STACK_SIZE];
    const TStack stack{ nullptr, stack_lo, 0 };
#endif
    PV_TREELET_STAGE(sc);
    const uint32_t n = min(wc->n_rays, cap);
    unsigned long long n_rays = 0, n_adc = 0;
    unsigned int max_level = 0;
    TravCount tc{ 0u, 0u };
    // all lanes of a warp (all threads of the block with PV_CTA_SYNC) stay in the loop and in the phase-voting traversal together
    const uint32_t cs = chunk_size(n);
    uint32_t i;
    while (next_chunk(&wc->cur_closest, n, cs, i)) {
        bool alive = i < n;
        V3 o = mk(0.0, 0.0, 0.0), d = mk(0.0, 0.0, 1.0);
        uint32_t flags = 0;
        HitRec out;
        out.pad = 0; out.csg = -1; out.aux = 0; out.depth = 0.0; out.ip[0] = out.ip[1] = out.ip[2] = 0.0; out.obj = PV_HIT_MISS;
        if (alive) {
            const PRay r = load_cs(cur + i);
            o = ld3(r.o); d = ld3(r.d);
            const float adcw = r.adc;
            const uint32_t level = r.level;
            flags = r.flags;
            if (flags & PV_RAY_DEAD) { out.obj = PV_HIT_STOPPED; alive = false; }
            else if (!(flags & PV_RAY_PROBE)) {
                n_rays++;
                // max. trace level / ADC bailout (trace.cpp:147-155)
                if ((level >= sc.g.max_trace_level) || ((double)adcw < sc.g.adc_bailout)) {
                    if ((double)adcw < sc.g.adc_bailout) n_adc++;
                    out.obj = PV_HIT_STOPPED;
                    alive = false;
                } else {
                    const unsigned int lvl = (flags & PV_RAY_CONTINUED) ? level : level + 1u;
                    if (lvl > max_level) max_level = lvl;
                }
            }
        }
        Hit best;
        best.depth = ((flags & PV_RAY_PRIMARY) && !(flags & PV_RAY_PROBE) && sc.cam.max_ray_distance >= PV_EPSILON) ? sc.cam.max_ray_distance : PV_BOUND_HUGE;
        best.obj = PV_NO_OBJECT;
        best.aux = 0; best.csg = -1;
#ifdef PV_DIAG_MAX_VISITS
        const uint32_t diag_nodes0 = tc.nodes;
#endif
        const bool found = find_intersection_sync<false>(alive, sc, o, d, flags & ~PV_RAY_PROBE, false, -1.0, best, stack, &cnt->overflow, tc);
#ifdef PV_DIAG_MAX_VISITS
        // diagnostic build: the largest number of box tests any single ray of the frame needed (Counters::pad, printed by PVGPU_TRACE_WAVES)
        if (alive) atomicMax(&cnt->pad, tc.nodes - diag_nodes0);
#endif
        if (found) {
            out.depth = best.depth; out.ip[0] = best.ip.x; out.ip[1] = best.ip.y; out.ip[2] = best.ip.z;
            out.obj = best.obj; out.aux = best.aux; out.csg = best.csg;
        }
        if (i < n) store_cs(hits + i, out);
    }
    // one atomic per warp for the statistics
    unsigned long long n_nodes = tc.nodes, n_prims = tc.prims;
    for (int off = 16; off > 0; off >>= 1) {
        n_rays += __shfl_down_sync(0xffffffffu, n_rays, off);
        n_adc += __shfl_down_sync(0xffffffffu, n_adc, off);
        n_nodes += __shfl_down_sync(0xffffffffu, n_nodes, off);
        n_prims += __shfl_down_sync(0xffffffffu, n_prims, off);
        max_level = max(max_level, __shfl_down_sync(0xffffffffu, max_level, off));
    }
    if ((threadIdx.x & 31) == 0) {
        if (n_nodes) atomicAdd(&cnt->node_tests[0], n_nodes);
        if (n_prims) atomicAdd(&cnt->prim_tests[0], n_prims);
        if (n_rays) atomicAdd(&cnt->rays, n_rays);
        if (n_adc) atomicAdd(&cnt->adc_saves, n_adc);
        if (max_level) atomicMax(&cnt->max_level, max_level);
    }
}

#if !PV_SECONDARY_TU
// ray-level harness: explicit rays under primary-ray conditions (Trace::FindIntersection(Intersection&, const Ray&), trace.h:255)
__global__ void k_probe_rays(const double* org_dir, uint32_t n, PRay* out)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        PRay r;
        #pragma unroll
        for (int k = 0; k < 3; k++) { r.o[k] = org_dir[6 * (size_t)i + k]; r.d[k] = org_dir[6 * (size_t)i + 3 + k]; }
        r.w[0] = r.w[1] = r.w[2] = 0.0f; r.wt = 0.0f; r.adc = 1.0f;
        r.sample = i; r.level = 0; r.flags = (uint8_t)(PV_RAY_PRIMARY | PV_RAY_PROBE); r.n_int = 0; r.pad = 0;
        #pragma unroll
        for (int k = 0; k < PV_MAX_INTERIORS; k++) r.interiors[k] = 0;
        out[i] = r;
    }
}

__global__ void k_probe_results(const HitRec* hits, uint32_t n, uint32_t* obj, double* depth, uint32_t* aux)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const HitRec h = hits[i];
        const bool found = h.obj < PV_HIT_STOPPED;
        obj[i] = found ? h.obj : PV_NO_OBJECT;
        depth[i] = found ? h.depth : PV_BOUND_HUGE;
        if (aux) aux[i] = found ? h.aux : 0u;
    }
}

__global__ void k_camera_rays(DScene sc, const double* xy, uint32_t n, double width, double height, double* org_dir)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        V3 o, d;
        if (!camera_ray(sc, xy[2 * (size_t)i], xy[2 * (size_t)i + 1], width, height, o, d)) { o = mk(0.0, 0.0, 0.0); d = mk(0.0, 0.0, 0.0); }
        double* r = org_dir + 6 * (size_t)i;
        r[0] = o.x; r[1] = o.y; r[2] = o.z; r[3] = d.x; r[4] = d.y; r[5] = d.z;
    }
}

void launch_container_state(const DScene& sc, uint16_t* out, Counters* cnt, cudaStream_t st)
{
    k_container_state<<<1, 32, PV_TREELET_SMEM, st>>>(sc, out, cnt);
}
void launch_wave_init(WaveCounts* ring, uint32_t n_slots, uint32_t n0, cudaStream_t st)
{
    k_wave_init<<<1, 256, 0, st>>>(ring, n_slots, n0);
}
void launch_resolve_conts(float4* accum, const Cont* conts, const Counters* cnt, uint32_t cont_cap, uint32_t cont_base, uint32_t wave, cudaStream_t st)
{
    k_resolve_conts<<<grid_for(cont_cap, 256, 8), 256, 0, st>>>(accum, conts, cnt, cont_cap, cont_base, wave);
}
void launch_clear_slots(const SampleSource& src, uint32_t first, uint32_t n, float4* accum, cudaStream_t st)
{
    k_clear_slots<<<grid_for(n, 256, 8), 256, 0, st>>>(src, first, n, accum);
}
void launch_primary(const DScene& sc, const SampleSource& src, uint32_t first, uint32_t n, double width, double height,
                    PRay* out, Counters* cnt, float4* accum, cudaStream_t st)
{
    k_primary<<<grid_for(n, 256, 8), 256, PV_TREELET_SMEM, st>>>(sc, src, first, n, width, height, out, cnt, accum);
}
#endif  // !PV_LEAN
void PV_VARIANT(launch_closest)(const DScene& sc, const PRay* cur, WaveCounts* wc, uint32_t n_bound, uint32_t cap, HitRec* hits, Counters* cnt, cudaStream_t st)
{
    PV_VARIANT(k_closest)<<<grid_for(trav_grid_bound(n_bound), PV_TRAV_BLOCK, PV_TRAV_MIN_BLOCKS), PV_TRAV_BLOCK, PV_TREELET_SMEM, st>>>(sc, cur, wc, cap, hits, cnt);
}
#if !PV_SECONDARY_TU
void launch_probe_rays(const double* org_dir, uint32_t n, PRay* out, cudaStream_t st)
{
    k_probe_rays<<<grid_for(n, 256, 8), 256, 0, st>>>(org_dir, n, out);
}
void launch_probe_results(const HitRec* hits, uint32_t n, uint32_t* obj, double* depth, uint32_t* aux, cudaStream_t st)
{
    k_probe_results<<<grid_for(n, 256, 8), 256, 0, st>>>(hits, n, obj, depth, aux);
}
void launch_camera_rays(const DScene& sc, const double* xy, uint32_t n, double width, double height, double* org_dir, cudaStream_t st)
{
    k_camera_rays<<<grid_for(n, 256, 8), 256, 0, st>>>(sc, xy, n, width, height, org_dir);
}

#endif  // !PV_LEAN

}  // namespace pvgpu
