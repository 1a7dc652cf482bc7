// Device side of pvgpu: scene upload, the wavefront kernels and the render / ray-harness entry points.
//
// One frame = a loop of "waves".  Wave 0 holds one camera ray per sample; every wave runs
//   k_trace   : per ray  - level / ADC test, closest hit (tree walk + FP64 primitive tests), shading of the
//                          hit, emission of shadow rays (one per light) and of reflection / refraction rays
//   k_shadow  : per shadow ray - blocker search towards the light, filtered through transparent objects,
//                          then the un-shadowed contribution x filter is added to the sample
// and the reflection / refraction rays it produced form the next wave.  Queues live in HBM; slots are
// handed out with warp-aggregated atomics (the compiler turns the uniform-address atomicAdd into
// REDUX + one atomic per warp).  There is no CPU fallback anywhere in this file.
#include "pvgpu_scene.hpp"
#include "pv_shade.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

namespace pvgpu {

#define CUDA_TRY(expr) \
    do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) return fail(PVGPU_E_CUDA, "%s: %s", #expr, cudaGetErrorString(e__)); } while (0)

// ------------------------------------------------------------------------------------------------
// device buffers
// ------------------------------------------------------------------------------------------------
struct DeviceScene {
    DScene view{};
    std::vector<void*> allocs;
    // work buffers (grown on demand)
    PRay* q[2] = { nullptr, nullptr };
    SRay* sq = nullptr;
    Counters* cnt = nullptr;
    pvgpu_rect* rects = nullptr;
    uint32_t* rect_off = nullptr;
    size_t q_cap = 0, sq_cap = 0, rect_cap = 0;
    uint16_t* d_cam_int = nullptr;       // container-state result
    unsigned long long kernel_launches = 0;
    bool camera_dirty = true;
};

template <class T>
static int upload(DeviceScene& d, const std::vector<T>& v, const T*& out, size_t& total)
{
    out = nullptr;
    size_t bytes = std::max<size_t>(v.size(), 1) * sizeof(T);
    void* p = nullptr;
    CUDA_TRY(cudaMalloc(&p, bytes));
    d.allocs.push_back(p);
    if (!v.empty()) CUDA_TRY(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    total += v.size() * sizeof(T);
    out = reinterpret_cast<const T*>(p);
    return PVGPU_OK;
}

// Noise tables: InitTextureTable (noise.cpp:231-255), RTable fill (noise.cpp:181-182),
// InitSolidNoise (noise.cpp:306-348).
static const double kRTableEven[267] = {
#include "pv_rtable.inc"
};

static void build_noise_tables(std::vector<uint16_t>& hash, std::vector<double>& rtable,
                               std::vector<uint16_t>& perm, std::vector<double>& grad)
{
    hash.resize(8192);
    for (int i = 0; i < 4096; i++) hash[i] = (uint16_t)i;
    int next_rand = 0;
    for (int i = 4095; i >= 0; i--) {
        next_rand = (int)((long long)next_rand * 1812433253LL + 12345LL);
        uint16_t j = (uint16_t)(((int)(next_rand >> 16) & 0x7FFF) % 4096);
        std::swap(hash[i], hash[j]);
    }
    for (int i = 0; i < 4096; i++) hash[4096 + i] = hash[i];

    rtable.resize(534);
    for (int i = 0; i < 267; i++) { rtable[2 * i] = kRTableEven[i]; rtable[2 * i + 1] = kRTableEven[i] * 0.5; }

    const int NE = 2048;
    std::vector<int> p(2 * (NE + 1), 0);
    grad.assign(3 * 2 * (NE + 1), 0.0);
    next_rand = 1;
    for (int i = 0; i < NE; i++) {
        double v[3], s;
        do {
            for (int j = 0; j < 3; j++) {
                next_rand = (int)((long long)next_rand * 1812433253LL + 12345LL);
                v[j] = (double)((((int)(next_rand >> 16) & 0x7FFF) % (NE << 1)) - NE) / (double)NE;
            }
            s = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
        } while ((s > 1.0) || (s < 1.0e-5));
        double l = std::sqrt(s);
        for (int j = 0; j < 3; j++) grad[3 * i + j] = v[j] / l;
    }
    for (int i = 0; i < NE; i++) p[i] = i;
    for (int i = NE; i > 0; i -= 2) {
        int k = p[i];
        next_rand = (int)((long long)next_rand * 1812433253LL + 12345LL);
        int j = ((int)(next_rand >> 16) & 0x7FFF) % NE;
        p[i] = p[j];
        p[j] = k;
    }
    for (int i = 0; i < NE + 2; i++) {
        p[NE + i] = p[i];
        for (int j = 0; j < 3; j++) grad[3 * (NE + i) + j] = grad[3 * i + j];
    }
    perm.resize(p.size());
    for (size_t i = 0; i < p.size(); i++) perm[i] = (uint16_t)p[i];
}

void device_release(Scene& s)
{
    if (!s.dev) return;
    if (s.device >= 0) cudaSetDevice(s.device);
    for (void* p : s.dev->allocs) cudaFree(p);
    cudaFree(s.dev->q[0]); cudaFree(s.dev->q[1]); cudaFree(s.dev->sq); cudaFree(s.dev->cnt);
    cudaFree(s.dev->rects); cudaFree(s.dev->rect_off); cudaFree(s.dev->d_cam_int);
    delete s.dev;
    s.dev = nullptr;
}

int device_upload(Scene& s, int device)
{
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
        return fail(PVGPU_E_NO_DEVICE, "no CUDA device available (pvgpu has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(PVGPU_E_INVALID, "device %d out of range (have %d)", device, ndev);
    CUDA_TRY(cudaSetDevice(device));
    s.device = device;
    DeviceScene* d = new DeviceScene();
    s.dev = d;
    size_t total = 0;
    DScene& v = d->view;

    // CSG leaf lists: primitive descendants of every parentless CSG object, depth first (csg.cpp walks children in order)
    std::vector<uint32_t> leaves;
    std::vector<uint2> leaf_range(s.objects.size(), make_uint2(0, 0));
    for (size_t i = 0; i < s.objects.size(); i++) {
        const pvgpu_object& o = s.objects[i];
        if (o.type < PVGPU_OBJ_CSG_UNION || o.parent >= 0) continue;
        uint32_t first = (uint32_t)leaves.size();
        std::vector<uint32_t> st{ (uint32_t)i };
        while (!st.empty()) {
            uint32_t c = st.back(); st.pop_back();
            const pvgpu_object& co = s.objects[c];
            if (co.type >= PVGPU_OBJ_CSG_UNION) {
                for (uint32_t k = co.child_count; k-- > 0;) {
                    uint32_t ch = s.index_list[co.child_first + k];
                    if (ch >= s.objects.size() || s.objects[ch].parent != (int32_t)c) {
                        device_release(s);
                        return fail(PVGPU_E_INVALID, "CSG object %u: child %u has inconsistent parent", c, ch);
                    }
                    st.push_back(ch);
                }
            } else {
                if (co.type == PVGPU_OBJ_MESH) { device_release(s); return fail(PVGPU_E_UNSUPPORTED, "mesh inside CSG is outside the hot-path scope"); }
                if (co.bound_count) { device_release(s); return fail(PVGPU_E_UNSUPPORTED, "bounded_by on a CSG child is outside the hot-path scope"); }
                leaves.push_back(c);
            }
        }
        leaf_range[i] = make_uint2(first, (uint32_t)leaves.size() - first);
    }

    // packed triangles + resolved mesh descriptors
    std::vector<DTri> dtris(s.triangles.size());
    std::vector<DMesh> dmeshes(s.meshes.size());
    for (size_t m = 0; m < s.meshes.size(); m++) {
        const pvgpu_mesh& me = s.meshes[m];
        DMesh& dm = dmeshes[m];
        dm.tri_first = me.triangle_first; dm.tri_count = me.triangle_count;
        dm.node_first = me.node_first; dm.node_count = me.node_count;
        dm.vertex_first = me.vertex_first; dm.normal_first = me.normal_first;
        dm.texture_first = me.texture_first; dm.texture_count = me.texture_count;
        dm.has_inside_vector = me.has_inside_vector;
        std::memcpy(dm.inside_vector, me.inside_vector, sizeof dm.inside_vector);
        const float* V = s.vertices.data() + 3 * (size_t)me.vertex_first;
        const float* N = s.normals.data() + 3 * (size_t)me.normal_first;
        for (uint32_t t = 0; t < me.triangle_count; t++) {
            const pvgpu_triangle& tr = s.triangles[me.triangle_first + t];
            DTri& dt = dtris[me.triangle_first + t];
            std::memcpy(dt.p1, V + 3 * tr.p1, 12); std::memcpy(dt.p2, V + 3 * tr.p2, 12); std::memcpy(dt.p3, V + 3 * tr.p3, 12);
            std::memcpy(dt.n, N + 3 * tr.normal_ind, 12);
            dt.dist = tr.distance; dt.dom = tr.dominant_axis; dt.pad[0] = dt.pad[1] = 0;
        }
    }
    std::vector<uint16_t> hash, perm;
    std::vector<double> rtable, grad;
    build_noise_tables(hash, rtable, perm, grad);

    int rc = PVGPU_OK;
    #define UP(vec, field) if (rc == PVGPU_OK) rc = upload(*d, vec, field, total)
    UP(s.objects, v.objs); UP(s.transforms, v.xf); UP(s.index_list, v.index_list); UP(s.frame, v.frame);
    UP(s.nodes, v.nodes); UP(dmeshes, v.meshes); UP(dtris, v.dtris); UP(s.triangles, v.tris);
    UP(s.vertices, v.verts); UP(s.normals, v.norms); UP(s.mesh_nodes, v.mnodes); UP(s.lights, v.lights);
    UP(s.textures, v.textures); UP(s.pigments, v.pigments); UP(s.finishes, v.finishes); UP(s.blend_maps, v.maps);
    UP(s.blend_entries, v.entries); UP(s.warps, v.warps); UP(s.interiors, v.interiors);
    UP(leaves, v.csg_leaves); UP(leaf_range, v.csg_leaf_range);
    UP(hash, v.noise.hash); UP(rtable, v.noise.rtable); UP(perm, v.noise.perm); UP(grad, v.noise.grad);
    #undef UP
    if (rc != PVGPU_OK) { device_release(s); return rc; }

    v.n_objs = (uint32_t)s.objects.size();
    v.n_frame = (uint32_t)s.frame.size();
    v.n_nodes = (uint32_t)s.nodes.size();
    v.n_lights = (uint32_t)s.lights.size();
    v.use_tree = (s.globals.bounding_method == 1 && !s.nodes.empty()) ? 1u : 0u;
    v.all_opaque = s.all_shadow_casters_opaque ? 1u : 0u;
    v.g = s.globals;
    v.cam = s.camera;
    v.n_cam_interiors = 0;
    s.device_bytes = total;

    if (cudaMalloc(&d->cnt, sizeof(Counters)) != cudaSuccess || cudaMalloc(&d->d_cam_int, 64) != cudaSuccess) {
        device_release(s);
        return fail(PVGPU_E_CUDA, "cudaMalloc of counters failed");
    }
    // the deepest Inside()/sturm paths keep a few small arrays per thread
    cudaDeviceSetLimit(cudaLimitStackSize, 4096);
    d->camera_dirty = true;
    return PVGPU_OK;
}

// ------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------

// TracePixel::InitRayContainerState (tracepixel.cpp:929-1006): interiors of all objects containing `p`.
__device__ inline void container_state(const DScene& sc, const V3& p, uint16_t* out, uint32_t& n, uint2* stack, unsigned int* overflow)
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
    stack[sp++] = make_uint2(0u, 0u);
    while (sp > 0) {
        const uint32_t ni = stack[--sp].y;
        const pvgpu_node nd = sc.nodes[ni];
        if (!inside_bbox(nd.lo, nd.size)) continue;
        if (nd.count == 0) test_object(nd.first, sp);
        else for (uint32_t c = nd.count; c-- > 0 && sp < PV_STACK_SIZE;) stack[sp++] = make_uint2(0u, nd.first + c);
    }
}

__global__ void k_container_state(DScene sc, uint16_t* out, Counters* cnt)
{
    if (threadIdx.x || blockIdx.x) return;
    uint2 stack[PV_STACK_SIZE];
    uint16_t ints[PV_MAX_INTERIORS];
    uint32_t n;
    container_state(sc, ld3(sc.cam.location), ints, n, stack, &cnt->overflow);
    for (uint32_t i = 0; i < n; i++) out[i] = ints[i];
    out[PV_MAX_INTERIORS] = (uint16_t)n;
}

// TracePixel::CreateCameraRay (tracepixel.cpp:341-391, 917-927): perspective and orthographic cameras.
__device__ __forceinline__ void camera_ray(const pvgpu_camera& cam, double x, double y, double width, double height, V3& o, V3& d)
{
    const double x0 = x / width - 0.5;
    const double y0 = 0.5 - y / height;
    const V3 loc = ld3(cam.location), dirv = ld3(cam.direction), right = ld3(cam.right), up = ld3(cam.up);
    if (cam.type == PVGPU_CAMERA_ORTHOGRAPHIC) {
        d = dirv;
        o = (loc + x0 * right) + y0 * up;
    } else {
        o = loc;
        d = (dirv + x0 * right) + y0 * up;
    }
    d = normalized(d);
}

// sample i -> (rectangle, x, y): rectangles are row-major runs, rect_off holds their prefix sums
__device__ __forceinline__ void sample_xy(const pvgpu_rect* rects, const uint32_t* rect_off, uint32_t n_rects, uint32_t i, double& x, double& y)
{
    uint32_t lo = 0, hi = n_rects;
    while (hi - lo > 1) { uint32_t mid = (lo + hi) >> 1; if (rect_off[mid] <= i) lo = mid; else hi = mid; }
    const pvgpu_rect r = rects[lo];
    const uint32_t k = i - rect_off[lo], w = (uint32_t)(r.right - r.left + 1);
    x = (double)(r.left + (int)(k % w)) + 0.5;       // SimpleSamplingM0: pixel centres (tracetask.cpp:438)
    y = (double)(r.top + (int)(k / w)) + 0.5;
}

__global__ void __launch_bounds__(256)
k_primary(DScene sc, const pvgpu_rect* rects, const uint32_t* rect_off, uint32_t n_rects, uint32_t first, uint32_t n,
          double width, double height, PRay* out)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        double x, y;
        sample_xy(rects, rect_off, n_rects, first + i, x, y);
        V3 o, d;
        camera_ray(sc.cam, x, y, width, height, o, d);
        PRay r;
        r.o[0] = o.x; r.o[1] = o.y; r.o[2] = o.z;
        r.d[0] = d.x; r.d[1] = d.y; r.d[2] = d.z;
        r.w[0] = r.w[1] = r.w[2] = 1.0f;
        r.wt = 1.0f;
        r.adc = 1.0f;
        r.sample = first + i;
        r.level = 0;
        r.flags = (uint8_t)(PV_RAY_PRIMARY | (sc.g.output_alpha ? PV_RAY_ALPHA_BG : 0));
        r.n_int = (uint8_t)sc.n_cam_interiors;
        r.pad = 0;
        #pragma unroll
        for (int k = 0; k < PV_MAX_INTERIORS; k++) r.interiors[k] = sc.cam_interiors[k];
        out[i] = r;
    }
}

// Trace::TraceRay (trace.cpp:135-228) for every ray of the wave.
__global__ void __launch_bounds__(128)
k_trace(DScene sc, const PRay* __restrict__ cur, uint32_t n, WaveCtx ctx)
{
    uint2 stack[PV_STACK_SIZE];
    unsigned long long n_rays = 0, n_adc = 0;
    unsigned int max_level = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        PRay ray = cur[i];
        n_rays++;
        // max. trace level / ADC bailout (trace.cpp:147-155)
        if ((ray.level >= sc.g.max_trace_level) || ((double)ray.adc < sc.g.adc_bailout)) {
            if ((double)ray.adc < sc.g.adc_bailout) n_adc++;
            continue;
        }
        const unsigned int lvl = (ray.flags & PV_RAY_CONTINUED) ? ray.level : ray.level + 1u;
        if (lvl > max_level) max_level = lvl;
        if ((ray.flags & PV_RAY_PRIMARY) && sc.cam.type == PVGPU_CAMERA_ORTHOGRAPHIC) {
            // InitRayContainerState(ray, true): recomputed per ray when the origin moves with the pixel
            uint32_t nci;
            container_state(sc, ld3(ray.o), ray.interiors, nci, stack, &ctx.cnt->overflow);
            ray.n_int = (uint8_t)nci;
        }
        Hit best;
        best.depth = ((ray.flags & PV_RAY_PRIMARY) && sc.cam.max_ray_distance >= PV_EPSILON) ? sc.cam.max_ray_distance : PV_BOUND_HUGE;
        best.obj = PV_NO_OBJECT;
        const bool found = find_intersection<false>(sc, ld3(ray.o), ld3(ray.d), ray.flags, false, -1.0, best, stack, &ctx.cnt->overflow);
        if (found) shade_hit(sc, ray, i, best, ctx);
        else {
            float col[3], transm;
            compute_sky(sc, ray, col, transm);
            accum_add(ctx.accum, ray.sample, ray.w[0] * col[0], ray.w[1] * col[1], ray.w[2] * col[2], ray.wt * transm);
        }
    }
    // one atomic per warp for the statistics
    for (int off = 16; off > 0; off >>= 1) {
        n_rays += __shfl_down_sync(0xffffffffu, n_rays, off);
        n_adc += __shfl_down_sync(0xffffffffu, n_adc, off);
        max_level = max(max_level, __shfl_down_sync(0xffffffffu, max_level, off));
    }
    if ((threadIdx.x & 31) == 0) {
        if (n_rays) atomicAdd(&ctx.cnt->rays, n_rays);
        if (n_adc) atomicAdd(&ctx.cnt->adc_saves, n_adc);
        if (max_level) atomicMax(&ctx.cnt->max_level, max_level);
    }
}

// Trace::TraceShadowRay (trace.cpp:1892-1940) for every shadow ray the chunk produced.
__global__ void __launch_bounds__(128)
k_shadow(DScene sc, const SRay* __restrict__ rays, const PRay* __restrict__ wave, float4* accum, Counters* cnt)
{
    uint2 stack[PV_STACK_SIZE];
    const uint32_t n = cnt->n_shadow;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const SRay s = rays[i];
        float f[3];
        trace_shadow(sc, s, wave, stack, cnt, f);
        accum_add(accum, s.sample, s.a[0] * f[0], s.a[1] * f[1], s.a[2] * f[2], 0.0f);
    }
}

// ray-level harness: Trace::FindIntersection for explicit rays under primary-ray conditions
__global__ void __launch_bounds__(128)
k_probe(DScene sc, const double* org_dir, uint32_t n, uint32_t* obj, double* depth, uint32_t* aux, Counters* cnt)
{
    uint2 stack[PV_STACK_SIZE];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const V3 o = ld3(org_dir + 6 * (size_t)i), d = ld3(org_dir + 6 * (size_t)i + 3);
        Hit best;
        best.depth = PV_BOUND_HUGE;
        best.obj = PV_NO_OBJECT;
        best.aux = 0;
        const bool found = find_intersection<false>(sc, o, d, PV_RAY_PRIMARY, false, -1.0, best, stack, &cnt->overflow);
        obj[i] = found ? best.obj : PV_NO_OBJECT;
        depth[i] = found ? best.depth : PV_BOUND_HUGE;
        if (aux) aux[i] = found ? best.aux : 0u;
    }
}

__global__ void k_camera_rays(DScene sc, const double* xy, uint32_t n, double width, double height, double* org_dir)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        V3 o, d;
        camera_ray(sc.cam, xy[2 * (size_t)i], xy[2 * (size_t)i + 1], width, height, o, d);
        double* r = org_dir + 6 * (size_t)i;
        r[0] = o.x; r[1] = o.y; r[2] = o.z; r[3] = d.x; r[4] = d.y; r[5] = d.z;
    }
}

// ------------------------------------------------------------------------------------------------
// host orchestration
// ------------------------------------------------------------------------------------------------
static int g_sm_count = 0;
static int grid_for(uint32_t n, int block, int per_sm)
{
    if (g_sm_count == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
        if (g_sm_count <= 0) g_sm_count = 148;
    }
    long long blocks = ((long long)n + block - 1) / block;
    long long cap = (long long)g_sm_count * per_sm;      // a multiple of the SM count; grid-stride loops cover the rest
    return (int)std::max<long long>(1, std::min(blocks, cap));
}

static int ensure_work_buffers(DeviceScene& d, size_t q_cap, size_t sq_cap, size_t n_rects)
{
    if (q_cap > d.q_cap) {
        cudaFree(d.q[0]); cudaFree(d.q[1]); d.q[0] = d.q[1] = nullptr; d.q_cap = 0;
        CUDA_TRY(cudaMalloc(&d.q[0], q_cap * sizeof(PRay)));
        CUDA_TRY(cudaMalloc(&d.q[1], q_cap * sizeof(PRay)));
        d.q_cap = q_cap;
    }
    if (sq_cap > d.sq_cap) {
        cudaFree(d.sq); d.sq = nullptr; d.sq_cap = 0;
        CUDA_TRY(cudaMalloc(&d.sq, sq_cap * sizeof(SRay)));
        d.sq_cap = sq_cap;
    }
    if (n_rects > d.rect_cap) {
        cudaFree(d.rects); cudaFree(d.rect_off); d.rects = nullptr; d.rect_off = nullptr; d.rect_cap = 0;
        CUDA_TRY(cudaMalloc(&d.rects, n_rects * sizeof(pvgpu_rect)));
        CUDA_TRY(cudaMalloc(&d.rect_off, (n_rects + 1) * sizeof(uint32_t)));
        d.rect_cap = n_rects;
    }
    return PVGPU_OK;
}

// camera-dependent part of the device view (camera interiors for the pinhole case)
static int refresh_camera(Scene& s, cudaStream_t stream)
{
    DeviceScene& d = *s.dev;
    d.view.cam = s.camera;
    d.view.n_cam_interiors = 0;
    if (!s.interiors.empty() && s.camera.type == PVGPU_CAMERA_PERSPECTIVE) {
        k_container_state<<<1, 32, 0, stream>>>(d.view, d.d_cam_int, d.cnt);
        d.kernel_launches++;
        uint16_t h[PV_MAX_INTERIORS + 1];
        CUDA_TRY(cudaMemcpyAsync(h, d.d_cam_int, sizeof h, cudaMemcpyDeviceToHost, stream));
        CUDA_TRY(cudaStreamSynchronize(stream));
        d.view.n_cam_interiors = h[PV_MAX_INTERIORS];
        for (int i = 0; i < PV_MAX_INTERIORS; i++) d.view.cam_interiors[i] = h[i];
    }
    d.camera_dirty = false;
    return PVGPU_OK;
}

static const size_t kBatchSamples = 1u << 22;          // samples traced per batch of waves
static const size_t kShadowCap = 1u << 23;             // shadow-ray queue capacity (records)

// Runs all waves for samples [first, first+n).  Returns PVGPU_E_OVERFLOW if a queue was too small.
static int run_batch(Scene& s, uint32_t first, uint32_t n, uint32_t n_rects, int width, int height,
                     float4* d_accum, cudaStream_t stream, pvgpu_stats& st)
{
    DeviceScene& d = *s.dev;
    const uint32_t q_cap = (uint32_t)d.q_cap, sq_cap = (uint32_t)d.sq_cap;
    const uint32_t n_lights = std::max<uint32_t>(1, (uint32_t)s.lights.size());
    const uint32_t chunk_max = std::max<uint32_t>(1, sq_cap / n_lights);
    k_primary<<<grid_for(n, 256, 8), 256, 0, stream>>>(d.view, d.rects, d.rect_off, n_rects, first, n, (double)width, (double)height, d.q[0]);
    d.kernel_launches++;
    uint32_t n_cur = n;
    int cur = 0;
    const uint32_t max_waves = s.globals.max_trace_level + 64;     // continued rays do not consume a level
    for (uint32_t wave = 0; n_cur > 0 && wave < max_waves; wave++) {
        CUDA_TRY(cudaMemsetAsync(&d.cnt->n_next, 0, sizeof(unsigned int), stream));
        for (uint32_t c0 = 0; c0 < n_cur; c0 += chunk_max) {
            const uint32_t cn = std::min(chunk_max, n_cur - c0);
            CUDA_TRY(cudaMemsetAsync(&d.cnt->n_shadow, 0, sizeof(unsigned int), stream));
            WaveCtx ctx;
            ctx.accum = d_accum; ctx.next = d.q[cur ^ 1]; ctx.shadow = d.sq; ctx.cnt = d.cnt;
            ctx.next_cap = q_cap; ctx.shadow_cap = sq_cap;
            k_trace<<<grid_for(cn, 128, 8), 128, 0, stream>>>(d.view, d.q[cur] + c0, cn, ctx);
            // the shadow kernel reads its count on the device; its grid is sized for the worst case of this chunk
            const uint32_t worst = (uint32_t)std::min<unsigned long long>((unsigned long long)cn * n_lights, sq_cap);
            k_shadow<<<grid_for(worst, 128, 8), 128, 0, stream>>>(d.view, d.sq, d.q[cur] + c0, d_accum, d.cnt);
            d.kernel_launches += 2;
        }
        unsigned int h[4];     // n_next, n_shadow, max_level, overflow
        CUDA_TRY(cudaMemcpyAsync(h, &d.cnt->n_next, sizeof h, cudaMemcpyDeviceToHost, stream));
        CUDA_TRY(cudaStreamSynchronize(stream));
        st.waves++;
        if (h[3] & (8u | 16u)) return PVGPU_E_OVERFLOW;
        n_cur = h[0];
        cur ^= 1;
    }
    return PVGPU_OK;
}

static int render_impl(Scene& s, const pvgpu_aa* aa, int width, int height, const pvgpu_rect* rects, size_t n_rects,
                       float* d_out, pvgpu_stats* stats, cudaStream_t stream, int (*cooperate)(void*), void* user)
{
    if (!s.dev) return fail(PVGPU_E_INVALID, "scene not finalized");
    if (width <= 0 || height <= 0 || !rects || !n_rects || !d_out) return fail(PVGPU_E_INVALID, "pvgpu_render: bad arguments");
    if (aa && aa->method != 0) return fail(PVGPU_E_UNSUPPORTED, "anti-aliasing method %u not available yet", aa->method);
    CUDA_TRY(cudaSetDevice(s.device));
    DeviceScene& d = *s.dev;
    std::vector<uint32_t> off(n_rects + 1, 0);
    for (size_t i = 0; i < n_rects; i++) {
        const pvgpu_rect& r = rects[i];
        if (r.right < r.left || r.bottom < r.top) return fail(PVGPU_E_INVALID, "rectangle %zu is empty", i);
        unsigned long long area = (unsigned long long)(r.right - r.left + 1) * (unsigned long long)(r.bottom - r.top + 1);
        if (off[i] + area > 0xFFFFFFF0ull) return fail(PVGPU_E_INVALID, "too many pixels in one call");
        off[i + 1] = off[i] + (uint32_t)area;
    }
    const uint32_t n_samples = off[n_rects];
    pvgpu_stats st{};
    const unsigned long long launches0 = d.kernel_launches;
    cudaEvent_t ev0, ev1;
    CUDA_TRY(cudaEventCreate(&ev0));
    CUDA_TRY(cudaEventCreate(&ev1));
    CUDA_TRY(cudaEventRecord(ev0, stream));
    CUDA_TRY(cudaMemsetAsync(d.cnt, 0, sizeof(Counters), stream));
    if (d.camera_dirty || std::memcmp(&d.view.cam, &s.camera, sizeof s.camera) != 0) {
        int rc = refresh_camera(s, stream);
        if (rc != PVGPU_OK) return rc;
    }
    size_t batch = std::min<size_t>(kBatchSamples, n_samples);
    int rc = ensure_work_buffers(d, std::max<size_t>(3 * batch, 1024), kShadowCap, n_rects);
    if (rc != PVGPU_OK) return rc;
    CUDA_TRY(cudaMemcpyAsync(d.rects, rects, n_rects * sizeof(pvgpu_rect), cudaMemcpyHostToDevice, stream));
    CUDA_TRY(cudaMemcpyAsync(d.rect_off, off.data(), (n_rects + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, stream));
    CUDA_TRY(cudaMemsetAsync(d_out, 0, (size_t)n_samples * 4 * sizeof(float), stream));
    float4* accum = reinterpret_cast<float4*>(d_out);

    // batches; a batch whose ray queues overflow is retried as two halves (accumulators of the batch are cleared first)
    struct Span { uint32_t first, n; };
    std::vector<Span> todo;
    for (uint32_t f = 0; f < n_samples; f += (uint32_t)batch) todo.push_back({ f, (uint32_t)std::min<size_t>(batch, n_samples - f) });
    std::reverse(todo.begin(), todo.end());
    while (!todo.empty()) {
        Span sp = todo.back(); todo.pop_back();
        if (cooperate && cooperate(user)) { cudaEventDestroy(ev0); cudaEventDestroy(ev1); return fail(PVGPU_E_ABORTED, "render aborted by the cooperate callback"); }
        rc = run_batch(s, sp.first, sp.n, (uint32_t)n_rects, width, height, accum, stream, st);
        if (rc == PVGPU_E_OVERFLOW && sp.n > 1) {
            CUDA_TRY(cudaMemsetAsync(accum + sp.first, 0, (size_t)sp.n * sizeof(float4), stream));
            CUDA_TRY(cudaMemsetAsync(&d.cnt->overflow, 0, sizeof(unsigned int), stream));
            todo.push_back({ sp.first + sp.n / 2, sp.n - sp.n / 2 });
            todo.push_back({ sp.first, sp.n / 2 });
            continue;
        }
        if (rc != PVGPU_OK) { cudaEventDestroy(ev0); cudaEventDestroy(ev1); return rc == PVGPU_E_OVERFLOW ? fail(rc, "ray queue overflow") : rc; }
    }
    Counters hc;
    CUDA_TRY(cudaMemcpyAsync(&hc, d.cnt, sizeof hc, cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaEventRecord(ev1, stream));
    CUDA_TRY(cudaEventSynchronize(ev1));
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, ev0, ev1);
    cudaEventDestroy(ev0); cudaEventDestroy(ev1);
    st.rays = hc.rays; st.shadow_ray_tests = hc.shadow_tests; st.reflected_rays = hc.reflected;
    st.refracted_rays = hc.refracted; st.transmitted_rays = hc.transmitted; st.tir_rays = hc.tir;
    st.adc_saves = hc.adc_saves; st.samples = 0; st.max_trace_level = hc.max_level; st.overflow = hc.overflow;
    st.kernel_launches = d.kernel_launches - launches0;
    st.device_ms = ms;
    if (stats) *stats = st;
    if (hc.overflow & ~(8u | 16u))
        return fail(PVGPU_E_OVERFLOW, "device capacity exceeded (flags 0x%x: 1 traversal stack, 2 mesh in CSG, 4 interior list)", hc.overflow);
    return PVGPU_OK;
}

}  // namespace pvgpu

using namespace pvgpu;

extern "C" {

int pvgpu_render_device(pvgpu_scene* sc, const pvgpu_aa* aa, int width, int height,
                        const pvgpu_rect* rects, size_t n_rects, float* d_rgbt_out,
                        pvgpu_stats* stats, void* cuda_stream)
{
    clear_error();
    if (!sc) return fail(PVGPU_E_INVALID, "pvgpu_render_device: null scene");
    return render_impl(*reinterpret_cast<Scene*>(sc), aa, width, height, rects, n_rects, d_rgbt_out, stats,
                       reinterpret_cast<cudaStream_t>(cuda_stream), nullptr, nullptr);
}

int pvgpu_render(pvgpu_scene* sc, const pvgpu_aa* aa, int width, int height,
                 const pvgpu_rect* rects, size_t n_rects, float* rgbt_out,
                 pvgpu_stats* stats, int (*cooperate)(void*), void* user)
{
    clear_error();
    if (!sc || !rgbt_out || !rects) return fail(PVGPU_E_INVALID, "pvgpu_render: null argument");
    Scene& s = *reinterpret_cast<Scene*>(sc);
    if (!s.dev) return fail(PVGPU_E_INVALID, "scene not finalized");
    CUDA_TRY(cudaSetDevice(s.device));
    size_t n = 0;
    for (size_t i = 0; i < n_rects; i++)
        if (rects[i].right >= rects[i].left && rects[i].bottom >= rects[i].top)
            n += (size_t)(rects[i].right - rects[i].left + 1) * (size_t)(rects[i].bottom - rects[i].top + 1);
    float* d_out = nullptr;
    CUDA_TRY(cudaMalloc(&d_out, std::max<size_t>(n, 1) * 4 * sizeof(float)));
    int rc = render_impl(s, aa, width, height, rects, n_rects, d_out, stats, 0, cooperate, user);
    if (rc == PVGPU_OK) {
        cudaError_t e = cudaMemcpy(rgbt_out, d_out, n * 4 * sizeof(float), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = fail(PVGPU_E_CUDA, "copy of the frame to the host failed: %s", cudaGetErrorString(e));
    }
    cudaFree(d_out);
    return rc;
}

int pvgpu_trace_rays(pvgpu_scene* sc, const double* org_dir, size_t n, uint32_t* obj, double* depth, uint32_t* aux)
{
    clear_error();
    if (!sc || !org_dir || !obj || !depth) return fail(PVGPU_E_INVALID, "pvgpu_trace_rays: null argument");
    Scene& s = *reinterpret_cast<Scene*>(sc);
    if (!s.dev) return fail(PVGPU_E_INVALID, "scene not finalized");
    if (n == 0) return PVGPU_OK;
    if (n > 0xFFFFFFF0ull) return fail(PVGPU_E_INVALID, "too many rays");
    CUDA_TRY(cudaSetDevice(s.device));
    DeviceScene& d = *s.dev;
    double *d_rays = nullptr, *d_depth = nullptr;
    uint32_t *d_obj = nullptr, *d_aux = nullptr;
    int rc = PVGPU_OK;
    auto cleanup = [&]() { cudaFree(d_rays); cudaFree(d_depth); cudaFree(d_obj); cudaFree(d_aux); };
    if (cudaMalloc(&d_rays, n * 6 * sizeof(double)) != cudaSuccess || cudaMalloc(&d_depth, n * sizeof(double)) != cudaSuccess ||
        cudaMalloc(&d_obj, n * sizeof(uint32_t)) != cudaSuccess || cudaMalloc(&d_aux, n * sizeof(uint32_t)) != cudaSuccess) {
        cleanup();
        return fail(PVGPU_E_CUDA, "cudaMalloc failed in pvgpu_trace_rays");
    }
    cudaMemcpy(d_rays, org_dir, n * 6 * sizeof(double), cudaMemcpyHostToDevice);
    cudaMemset(d.cnt, 0, sizeof(Counters));
    k_probe<<<grid_for((uint32_t)n, 128, 8), 128>>>(d.view, d_rays, (uint32_t)n, d_obj, d_depth, d_aux, d.cnt);
    d.kernel_launches++;
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) rc = fail(PVGPU_E_CUDA, "k_probe failed: %s", cudaGetErrorString(e));
    if (rc == PVGPU_OK) {
        cudaMemcpy(obj, d_obj, n * sizeof(uint32_t), cudaMemcpyDeviceToHost);
        cudaMemcpy(depth, d_depth, n * sizeof(double), cudaMemcpyDeviceToHost);
        if (aux) cudaMemcpy(aux, d_aux, n * sizeof(uint32_t), cudaMemcpyDeviceToHost);
        Counters hc;
        cudaMemcpy(&hc, d.cnt, sizeof hc, cudaMemcpyDeviceToHost);
        if (hc.overflow) rc = fail(PVGPU_E_OVERFLOW, "device capacity exceeded (flags 0x%x)", hc.overflow);
    }
    cleanup();
    return rc;
}

int pvgpu_camera_rays(pvgpu_scene* sc, int width, int height, const double* xy, size_t n, double* org_dir)
{
    clear_error();
    if (!sc || !xy || !org_dir) return fail(PVGPU_E_INVALID, "pvgpu_camera_rays: null argument");
    Scene& s = *reinterpret_cast<Scene*>(sc);
    if (!s.dev) return fail(PVGPU_E_INVALID, "scene not finalized");
    if (n == 0) return PVGPU_OK;
    CUDA_TRY(cudaSetDevice(s.device));
    DeviceScene& d = *s.dev;
    d.view.cam = s.camera;
    double *d_xy = nullptr, *d_out = nullptr;
    if (cudaMalloc(&d_xy, n * 2 * sizeof(double)) != cudaSuccess || cudaMalloc(&d_out, n * 6 * sizeof(double)) != cudaSuccess) {
        cudaFree(d_xy); cudaFree(d_out);
        return fail(PVGPU_E_CUDA, "cudaMalloc failed in pvgpu_camera_rays");
    }
    cudaMemcpy(d_xy, xy, n * 2 * sizeof(double), cudaMemcpyHostToDevice);
    k_camera_rays<<<grid_for((uint32_t)n, 256, 8), 256>>>(d.view, d_xy, (uint32_t)n, (double)width, (double)height, d_out);
    d.kernel_launches++;
    cudaError_t e = cudaMemcpy(org_dir, d_out, n * 6 * sizeof(double), cudaMemcpyDeviceToHost);
    cudaFree(d_xy); cudaFree(d_out);
    if (e != cudaSuccess) return fail(PVGPU_E_CUDA, "k_camera_rays failed: %s", cudaGetErrorString(e));
    return PVGPU_OK;
}

}  // extern "C"
