// Device side of pvgpu: scene upload, the wavefront kernels and the render / ray-harness entry points.
//
// One frame = a loop of "waves".  Wave 0 holds one camera ray per sample (k_primary); every wave runs
//   k_closest : per ray  - level / ADC test, closest hit (tree walk + FP64 primitive tests) -> HitRec
//   k_shade   : per ray  - shading of the hit (or the sky), emission of shadow rays (one per light) and of
//                          the reflection / refraction rays that form the next wave
//   k_shadow_*: per shadow ray - blocker search towards the light (any-hit when every caster is opaque,
//                          else filtered through transparent objects), then the contribution is added
// Queues live in HBM; slots are handed out with warp-aggregated atomics.  The ray counts stay on the device (a ring of
// WaveCounts records): the host launches the kernels of wave k + 1 before it knows how many rays wave k produced and reads the
// counts one wave behind only to know when to stop, so the GPU never waits for the host between waves.  k_shadow_* of wave k runs
// on a second stream next to k_closest / k_shade of wave k + 1 (they are independent; shadow queues are double-, ray queues
// triple-buffered for that).  A scene may be replicated on several devices; then one host thread per device pulls chunks of
// rectangles from one atomic counter (the GetNextRectangle contract, view.cpp:236-271) and delivers its tiles into the caller's
// frame.  The kernels live in k_*.cu; this file is the host side.  There is no CPU fallback anywhere.
#include "pvgpu_scene.hpp"
#include "pv_kernels.hpp"

#include <algorithm>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <chrono>
#include <memory>
#include <string>
#include <thread>
#include <random>
#include <vector>

namespace pvgpu {

#define CUDA_TRY(expr) \
    do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) return fail(PVGPU_E_CUDA, "%s: %s", #expr, cudaGetErrorString(e__)); } while (0)

// ------------------------------------------------------------------------------------------------
// device buffers
// ------------------------------------------------------------------------------------------------
#define PV_RING_SLOTS 512          // WaveCounts records per batch: max_trace_level (<= 256) + continued rays + slack

// Everything one in-flight render call needs on one device: streams, ray queues, counters, staging.  A device has one or more
// of these (PVGPU_CTX_PER_DEVICE); each serves one call at a time.
struct WorkCtx {
    int device = -1;
    cudaStream_t s_main = nullptr, s_shadow = nullptr;     // own non-blocking streams; pvgpu_render_device uses the caller's as main
    PRay* q[4] = { nullptr, nullptr, nullptr, nullptr };   // wave k lives in q[k % 4]
    HitRec* hits = nullptr;
    SRay* sq[3] = { nullptr, nullptr, nullptr };           // shadow rays of wave k in sq[k % 3]: k_shadow_* may lag two waves behind
    Counters* cnt = nullptr;             // [0] live counters, [1] their state at the start of the current batch (restored when it is retried)
    WaveCounts* ring = nullptr;
    unsigned int* h_counts = nullptr;                      // pinned: n_rays of wave k + 1 as read back after k_shade of wave k
    pvgpu_rect* rects = nullptr;
    uint32_t* rect_off = nullptr;
    size_t q_cap = 0, sq_cap = 0, rect_cap = 0;
    float* area_grid = nullptr;          // lightGrid scratch of k_shadow_area (3 floats x area_grid_max per resident thread)
    Cont* conts = nullptr;               // continuation records (reflection exponent != 1), cont_cap of them
    size_t cont_cap = 0;
    float4* accum_ext = nullptr;         // frame accumulators + continuation slots of a non-anti-aliased call with such records
    size_t accum_ext_cap = 0;
    // host-side staging of pvgpu_render (pinned) and its device frame
    float* d_frame = nullptr;
    float* h_frame = nullptr;
    size_t frame_cap = 0, h_frame_cap = 0;
    unsigned long long kernel_launches = 0;
    // per-launch timing (CUDA events on the launching stream)
    std::vector<cudaEvent_t> ev_pool;
    struct Timed { int kind; size_t e0, e1; unsigned long long items; };
    std::vector<Timed> timed;
    size_t ev_used = 0;
    std::mutex in_use;
};

struct DeviceScene {
    int device = -1;
    DScene view{};
    std::vector<void*> allocs;
    std::vector<std::unique_ptr<WorkCtx>> ctx;
    uint16_t* d_cam_int = nullptr;       // container-state result
    bool lean = false;                   // only spheres, boxes, planes, meshes and no clipped_by / bounded_by: lean kernel variants
    bool full = false;                   // normal{}, pigment_map / average, sky_sphere, fog or area lights: full-material shading variants
    bool csg = false;                    // quadric-class primitives (+ CSG) only: the _csg traversal variants
    bool quartic = false;                // stand-alone spheres, boxes, planes, quadrics, tori, blobs: the _quartic traversal variants
    bool camera_dirty = true;
    uint32_t spawn_factor = 0;           // upper bound of the rays one shaded ray adds to the next wave (0: the frame is one wave)
    uint32_t shadow_factor = 1;          // upper bound of the shadow rays one shaded ray emits
    bool has_reflect_exp = false;        // some reflective finish has Reflect_Exp != 1: continuation records (Cont, pv_common.cuh)
    size_t bytes = 0;
    std::mutex camera_mutex;
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

// Traversal copy of a node array (see DNode in pv_common.cuh).  `base` is added to child indices (mesh trees are stored
// back to back in one array with indices relative to their own root).  Nodes with more than 14 children are split
// into groups of <= 14 appended at the end of this tree's range; the groups repeat the parent's box and flags.
static int make_dnodes(const pvgpu_node* nodes, size_t n, std::vector<DNode>& out)
{
    const size_t base = out.size();
    out.resize(base + n);
    // appended group nodes need contiguous children: children of `nodes` already are, so a group is a sub-range
    std::vector<DNode> extra;
    auto fill = [&](DNode& dn, const pvgpu_node& nd) {
        for (int k = 0; k < 3; k++) { dn.lo[k] = nd.lo[k]; dn.hi[k] = nd.lo[k] + nd.size[k]; }     // FP32 add, as the reference does per test
        dn.aux = 0;
    };
    for (size_t i = 0; i < n; i++) {
        const pvgpu_node& nd = nodes[i];
        DNode& dn = out[base + i];
        fill(dn, nd);
        const uint32_t inf = (nd.flags & PVGPU_NODE_INFINITE) ? PV_CODE_INFINITE : 0u;
        if (nd.count == 0) {
            if (nd.first > PV_CODE_INDEX) return fail(PVGPU_E_UNSUPPORTED, "more than 2^27 objects / triangles");
            dn.code = inf | nd.first;
        } else if (nd.count <= 14) {
            dn.code = ((uint32_t)nd.count << 28) | inf | (uint32_t)(nd.first);
        } else {
            // groups of <= 14 children; a group node has the parent's box.  Up to 14 groups (196 children) per level.
            std::vector<std::pair<uint32_t, uint32_t>> ranges;      // (first, count) at the current level, indices into this tree
            for (uint32_t c = 0; c < nd.count; c += 14) ranges.push_back({ nd.first + c, std::min<uint32_t>(14u, nd.count - c) });
            while (ranges.size() > 14) return fail(PVGPU_E_UNSUPPORTED, "a bounding node with more than 196 children");
            const uint32_t gfirst = (uint32_t)(n + extra.size());
            for (auto& r : ranges) {
                DNode g;
                fill(g, nd);
                g.code = (r.second << 28) | inf | r.first;
                extra.push_back(g);
            }
            dn.code = ((uint32_t)ranges.size() << 28) | inf | gfirst;
        }
    }
    out.insert(out.end(), extra.begin(), extra.end());
    if (out.size() - base > PV_CODE_INDEX) return fail(PVGPU_E_UNSUPPORTED, "more than 2^27 bounding nodes");
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

static void release_ctx(WorkCtx& c)
{
    cudaFree(c.q[0]); cudaFree(c.q[1]); cudaFree(c.q[2]); cudaFree(c.q[3]); cudaFree(c.sq[0]); cudaFree(c.sq[1]); cudaFree(c.sq[2]); cudaFree(c.cnt); cudaFree(c.hits); cudaFree(c.ring);
    cudaFree(c.rects); cudaFree(c.rect_off); cudaFree(c.area_grid); cudaFree(c.d_frame); cudaFree(c.conts); cudaFree(c.accum_ext);
    if (c.h_frame) cudaFreeHost(c.h_frame);
    if (c.h_counts) cudaFreeHost(c.h_counts);
    for (cudaEvent_t e : c.ev_pool) cudaEventDestroy(e);
    if (c.s_main) cudaStreamDestroy(c.s_main);
    if (c.s_shadow) cudaStreamDestroy(c.s_shadow);
}

static void release_one(DeviceScene* d)
{
    if (!d) return;
    if (d->device >= 0) cudaSetDevice(d->device);
    for (auto& c : d->ctx) release_ctx(*c);
    for (void* p : d->allocs) cudaFree(p);
    cudaFree(d->d_cam_int);
    if (d->device >= 0) {       // hand the pooled anti-aliasing scratch back with the scene
        cudaMemPool_t pool;
        if (cudaDeviceSynchronize() != cudaSuccess || cudaDeviceGetDefaultMemPool(&pool, d->device) != cudaSuccess || cudaMemPoolTrimTo(pool, 0) != cudaSuccess) cudaGetLastError();
    }
    delete d;
}

void device_release(Scene& s)
{
    for (DeviceScene* d : s.devs) release_one(d);
    s.devs.clear();
    s.dev = nullptr;
}

// Replicates the host tables of `s` on `device` (called once per device of the scene).
static double now_s()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

static int upload_one(Scene& s, int device, DeviceScene*& out)
{
    out = nullptr;
    const bool trace_t = getenv("PVGPU_TRACE_TIMING") != nullptr;
    const double t_begin = now_s();
    double t_mark = t_begin;
    auto mark = [&](const char* what) { if (trace_t) { const double t = now_s(); fprintf(stderr, "pvgpu upload: %-28s %.3f s\n", what, t - t_mark); t_mark = t; } };
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
        return fail(PVGPU_E_NO_DEVICE, "no CUDA device available (pvgpu has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(PVGPU_E_INVALID, "device %d out of range (have %d)", device, ndev);
    CUDA_TRY(cudaSetDevice(device));
    {   // The traversal / shading kernels have different (large) local-memory frames; ask the driver to keep the device's local-memory
        // pool at its high-water mark instead of re-sizing it between launches (cudaDeviceLmemResizeToMax).  Harmless where the
        // context already exists (torch created it first in bench.py: the flag is then ignored, and no difference was measured).
        unsigned int flags = 0;
        if (cudaGetDeviceFlags(&flags) == cudaSuccess && !(flags & cudaDeviceLmemResizeToMax)) {
            if (cudaSetDeviceFlags(flags | cudaDeviceLmemResizeToMax) != cudaSuccess) cudaGetLastError();
        }
    }
    // the FP32 stand-in for EPSILON in the slab test must be the smallest float >= 1e-10 (pv_traverse.cuh)
    if (!((double)1.0e-10f >= 1.0e-10 && (double)std::nextafterf(1.0e-10f, 0.0f) < 1.0e-10))
        return fail(PVGPU_E_INVALID, "internal: FP32 epsilon of the slab test is not the smallest float >= 1e-10");
    if (s.nodes.size() >= (1u << 27) || s.mesh_nodes.size() >= (1u << 27) || s.triangles.size() >= (1u << 27) || s.objects.size() >= (1u << 27))
        return fail(PVGPU_E_UNSUPPORTED, "more than 2^27 nodes / triangles / objects");
    DeviceScene* d = new DeviceScene();
    d->device = device;
    size_t total = 0;
    DScene& v = d->view;

    {   // the anti-aliasing scratch comes from the default memory pool (Scratch): keep what it frees
        cudaMemPool_t pool;
        unsigned long long keep = ~0ull;
        if (cudaDeviceGetDefaultMemPool(&pool, device) != cudaSuccess || cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep) != cudaSuccess) cudaGetLastError();
    }
    mark("context + device flags");
    // CSG leaf lists: primitive descendants of every parentless CSG object, depth first (csg.cpp walks children in order)
    std::vector<uint32_t> leaves;
    std::vector<uint2> leaf_range(s.objects.size(), make_uint2(0, 0));
    for (size_t i = 0; i < s.objects.size(); i++) {
        const pvgpu_object& o = s.objects[i];
        if (!PVGPU_IS_CSG(o.type) || o.parent >= 0) continue;
        uint32_t first = (uint32_t)leaves.size();
        std::vector<uint32_t> st{ (uint32_t)i };
        while (!st.empty()) {
            uint32_t c = st.back(); st.pop_back();
            const pvgpu_object& co = s.objects[c];
            if (PVGPU_IS_CSG(co.type)) {
                for (uint32_t k = co.child_count; k-- > 0;) {
                    uint32_t ch = s.index_list[co.child_first + k];
                    if (ch >= s.objects.size() || s.objects[ch].parent != (int32_t)c) {
                        release_one(d);
                        return fail(PVGPU_E_INVALID, "CSG object %u: child %u has inconsistent parent", c, ch);
                    }
                    st.push_back(ch);
                }
            } else {
                leaves.push_back(c);
            }
        }
        // "flat" CSG: every child is a simple primitive without its own clipped_by list -> csg_hits' fast path (bit 31)
        bool flat = true;
        for (uint32_t k = 0; k < o.child_count && flat; k++) {
            const pvgpu_object& co = s.objects[s.index_list[o.child_first + k]];
            flat = ((co.type >= PVGPU_OBJ_SPHERE && co.type <= PVGPU_OBJ_TORUS) || (co.type >= PVGPU_OBJ_CONE && co.type <= PVGPU_OBJ_POLY)) && co.clip_count == 0 && co.bound_count == 0;
        }
        leaf_range[i] = make_uint2(first, ((uint32_t)leaves.size() - first) | (flat ? 0x80000000u : 0u));
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
    // traversal copies of the trees; every mesh tree gets its own index space (children relative to its root)
    std::vector<DNode> dnodes, dmnodes;
    {
        int rc = make_dnodes(s.nodes.data(), s.nodes.size(), dnodes);
        for (size_t m = 0; m < s.meshes.size() && rc == PVGPU_OK; m++) {
            const pvgpu_mesh& me = s.meshes[m];
            dmeshes[m].node_first = (uint32_t)dmnodes.size();
            rc = make_dnodes(s.mesh_nodes.data() + me.node_first, me.node_count, dmnodes);
        }
        if (rc != PVGPU_OK) { release_one(d); return rc; }
    }

    mark("host-side derived tables");
    int rc = PVGPU_OK;
    #define UP(vec, field) if (rc == PVGPU_OK) rc = upload(*d, vec, field, total)
    UP(s.objects, v.objs); UP(s.transforms, v.xf); UP(s.index_list, v.index_list); UP(s.frame, v.frame);
    UP(s.nodes, v.nodes); UP(dnodes, v.dnodes); UP(dmnodes, v.dmnodes); UP(dmeshes, v.meshes); UP(dtris, v.dtris); UP(s.triangles, v.tris);
    UP(s.vertices, v.verts); UP(s.normals, v.norms); UP(s.lights, v.lights);
    UP(s.textures, v.textures); UP(s.pigments, v.pigments); UP(s.finishes, v.finishes); UP(s.blend_maps, v.maps);
    UP(s.blend_entries, v.entries); UP(s.warps, v.warps); UP(s.interiors, v.interiors);
    UP(s.blobs, v.blobs); UP(s.blob_elements, v.blob_elements); UP(s.blob_nodes, v.blob_nodes); UP(s.shape_data, v.shape_data);
    if (!s.blob_textures.empty()) { UP(s.blob_textures, v.blob_textures); } else v.blob_textures = nullptr;
    if (!s.tri_uv.empty()) { UP(s.mesh_uv, v.mesh_uv); UP(s.tri_uv, v.tri_uv); } else { v.mesh_uv = nullptr; v.tri_uv = nullptr; }
    UP(s.images, v.images); UP(s.texels, v.texels);
    UP(s.tnormals, v.tnormals); UP(s.slope_entries, v.slopes); UP(s.fogs, v.fogs);
    std::vector<double> pattern_rands;
    for (const pvgpu_pigment& pg : s.pigments)
        if ((pg.pattern == PVGPU_PAT_CRACKLE || pg.pattern == PVGPU_PAT_CELLS) && pattern_rands.empty()) {
            std::mt19937 gen;                                   // RandomDoubles (randomsequence.cpp:138-149): boost mt19937 + uniform_real<double>(0, 1)
            pattern_rands.resize(32768);
            for (double& r : pattern_rands) r = gen() / 4294967296.0;
        }
    UP(pattern_rands, v.pattern_rands);
    UP(leaves, v.csg_leaves); UP(leaf_range, v.csg_leaf_range);
    UP(hash, v.noise.hash); UP(rtable, v.noise.rtable); UP(perm, v.noise.perm); UP(grad, v.noise.grad);
    #undef UP
    if (rc != PVGPU_OK) { release_one(d); return rc; }
    mark("cudaMalloc + H2D of the tables");
    if (s.globals.number_of_waves) {       // TraceThreadData::waveSources / waveFrequencies, computed with the device's own DNoise
        const uint32_t nw = s.globals.number_of_waves;
        std::vector<double> zeros(4 * (size_t)nw, 0.0);
        const double* d_waves = nullptr;
        rc = upload(*d, zeros, d_waves, total);
        if (rc != PVGPU_OK) { release_one(d); return rc; }
        v.wave_sources = d_waves; v.wave_freqs = d_waves + 3 * (size_t)nw;
        launch_init_waves(v.noise, nw, const_cast<double*>(v.wave_sources), const_cast<double*>(v.wave_freqs), 0);
        if (cudaDeviceSynchronize() != cudaSuccess) { release_one(d); return fail(PVGPU_E_CUDA, "wave source initialisation failed: %s", cudaGetErrorString(cudaGetLastError())); }
    }

    d->lean = true;
    for (const pvgpu_object& o : s.objects)
        if (!(o.type == PVGPU_OBJ_SPHERE || o.type == PVGPU_OBJ_BOX || o.type == PVGPU_OBJ_PLANE || o.type == PVGPU_OBJ_MESH) ||
            o.clip_count || o.bound_count) d->lean = false;
    for (const pvgpu_texture& t : s.textures) if (t.tnormal >= 0 || t.type != PVGPU_PAT_PLAIN) d->lean = false;
    for (const pvgpu_pigment& pg : s.pigments) if (pg.pattern > PVGPU_PAT_AGATE || pg.pattern == PVGPU_PAT_BRICK || pg.pattern == PVGPU_PAT_HEXAGON) d->lean = false;
    if (!s.fogs.empty() || !s.sky_spheres.empty()) d->lean = false;
    for (const pvgpu_blend_map& m : s.blend_maps) if (m.blend_mode & PVGPU_BLEND_PIGMENT_MAP) d->lean = false;
    for (const pvgpu_pigment& pg : s.pigments) if (pg.pattern == PVGPU_PAT_AVERAGE || pg.pattern == PVGPU_PAT_IMAGE_MAP) d->lean = false;
    for (int k = 0; k < 3; k++) v.irid_wavelengths[k] = (s.irid_wavelengths.size() == 3) ? s.irid_wavelengths[k] : 1.0f;
    v.has_tnormals = 0;
    for (const pvgpu_texture& t : s.textures) if (t.tnormal >= 0) v.has_tnormals = 1;
    v.has_area_lights = 0; v.area_grid_max = 0;
    if (s.globals.quality_flags & PVGPU_Q_AREA_LIGHTS)
        for (const pvgpu_light& l : s.lights)
            if (l.flags & PVGPU_LIGHT_AREA) { v.has_area_lights = 1; v.area_grid_max = std::max<uint32_t>(v.area_grid_max, (uint32_t)(l.area_size1 * l.area_size2)); }
    if (v.has_area_lights) d->lean = false;
    v.n_fogs = (uint32_t)s.fogs.size();
    v.has_sky = s.sky_spheres.empty() ? 0u : 1u;
    if (v.has_sky) v.sky = s.sky_spheres[0];
    for (const pvgpu_mesh& me : s.meshes) if (me.node_count == 0) d->lean = false;       // `hierarchy off` meshes take the generic walk
    if (const char* e = getenv("PVGPU_LEAN")) if (e[0] == '0') d->lean = false;
    d->full = false;
    for (const pvgpu_texture& t : s.textures) if (t.tnormal >= 0 || t.type != PVGPU_PAT_PLAIN) d->full = true;
    for (const pvgpu_blend_map& m : s.blend_maps) if (m.blend_mode & PVGPU_BLEND_PIGMENT_MAP) d->full = true;
    for (const pvgpu_pigment& pg : s.pigments) if (pg.pattern >= PVGPU_PAT_AVERAGE) d->full = true;      // average, crackle, cells
    if (!s.fogs.empty() || !s.sky_spheres.empty() || v.has_area_lights || !s.blob_textures.empty() || !s.images.empty()) d->full = true;
    for (const pvgpu_object& o : s.objects) if (o.type == PVGPU_OBJ_GLYPH || o.type == PVGPU_OBJ_PRISM || o.type == PVGPU_OBJ_SUPERELLIPSOID) d->full = true;      // their normals live in the full shading kernels
    v.has_uv = 0; v.pad_uv = 0;
    for (const pvgpu_object& o : s.objects) if (o.flags & PVGPU_UV_FLAG) v.has_uv = 1;
    for (const pvgpu_pigment& pg : s.pigments) if (pg.pattern == PVGPU_PAT_UV_MAP) v.has_uv = 1;
    for (const pvgpu_warp& w : s.warps) if (w.type > PVGPU_WARP_CLASSIC_TURBULENCE) { d->full = true; d->lean = false; }      // point-mapping warps
    if (s.camera.reserved) { d->full = true; d->lean = false; }      // camera normal: Perturb_Normal is full-variant code
    if (v.has_uv) { d->full = true; d->lean = false; }      // hit_uv / uv_mapping pigments are full-variant code
    for (const pvgpu_finish& fi : s.finishes) if (fi.irid > 0.0f) { d->full = true; d->lean = false; }
    for (const pvgpu_finish& fi : s.finishes) {
        const bool reflective = fi.reflection_max[0] != 0 || fi.reflection_max[1] != 0 || fi.reflection_max[2] != 0 ||
                                fi.reflection_min[0] != 0 || fi.reflection_min[1] != 0 || fi.reflection_min[2] != 0;
        if (reflective && fi.reflect_exp != 1.0f) { d->has_reflect_exp = true; d->full = true; d->lean = false; }
    }
    if (const char* e = getenv("PVGPU_FULL")) if (e[0] == '1') d->full = true;
    d->csg = !d->lean && !d->full;
    for (const pvgpu_object& o : s.objects) {
        const bool in_class = o.type == PVGPU_OBJ_SPHERE || o.type == PVGPU_OBJ_BOX || o.type == PVGPU_OBJ_PLANE || o.type == PVGPU_OBJ_QUADRIC ||
                              o.type == PVGPU_OBJ_CONE || o.type == PVGPU_OBJ_DISC || PVGPU_IS_CSG(o.type);
        if (!in_class) d->csg = false;
    }
    if (const char* e = getenv("PVGPU_CSG")) if (e[0] == '0') d->csg = false;
    d->quartic = !d->lean && !d->full && !d->csg;
    for (const pvgpu_object& o : s.objects) {
        const bool in_class = o.type == PVGPU_OBJ_SPHERE || o.type == PVGPU_OBJ_BOX || o.type == PVGPU_OBJ_PLANE || o.type == PVGPU_OBJ_QUADRIC ||
                              o.type == PVGPU_OBJ_TORUS || o.type == PVGPU_OBJ_BLOB;
        if (!in_class) d->quartic = false;
    }
    if (const char* e = getenv("PVGPU_QUARTIC")) if (e[0] == '0') d->quartic = false;
    v.n_objs = (uint32_t)s.objects.size();
    v.n_frame = (uint32_t)s.frame.size();
    v.n_nodes = (uint32_t)s.nodes.size();
    v.n_mnodes = (uint32_t)dmnodes.size();
    v.n_lights = (uint32_t)s.lights.size();
    v.use_tree = (s.globals.bounding_method == 1 && !s.nodes.empty()) ? 1u : 0u;
    v.all_opaque = s.all_shadow_casters_opaque ? 1u : 0u;
    v.has_interiors = s.interiors.empty() ? 0u : 1u;
    v.g = s.globals;
    v.cam = s.camera;
    v.n_cam_interiors = 0;
    d->bytes = total;
    // how a wave can grow: a shaded ray adds at most one transmitted ray (or its total internal reflection) when the scene has
    // interiors, plus one reflected ray per reflective layer; and one shadow ray per light (per resolved texture_map leaf)
    {
        uint32_t refl_layers = 0;
        for (size_t t = 0; t < s.textures.size(); t++) {
            uint32_t nl = 0;
            for (int32_t li = (int32_t)t; li >= 0 && nl < PV_MAX_LAYERS; li = s.textures[li].next) {
                const pvgpu_texture& tx = s.textures[li];
                if (tx.type != PVGPU_PAT_PLAIN || tx.finish < 0 || (size_t)tx.finish >= s.finishes.size()) { nl = PV_MAX_LAYERS; break; }
                const pvgpu_finish& fi = s.finishes[tx.finish];
                if (fi.reflection_max[0] != 0.0f || fi.reflection_max[1] != 0.0f || fi.reflection_max[2] != 0.0f ||
                    fi.reflection_min[0] != 0.0f || fi.reflection_min[1] != 0.0f || fi.reflection_min[2] != 0.0f) nl++;
            }
            refl_layers = std::max(refl_layers, nl);
        }
        const bool leaves = d->full;     // texture_map: up to PV_MAX_TEX_LEAVES plain textures are shaded per hit
        d->spawn_factor = (refl_layers + (s.interiors.empty() ? 0u : 1u)) * (leaves ? 16u : 1u);
        d->shadow_factor = std::max<uint32_t>(1u, (uint32_t)s.lights.size()) * (leaves ? 16u : 1u);
    }

    if (cudaMalloc(&d->d_cam_int, 64) != cudaSuccess) {
        release_one(d);
        return fail(PVGPU_E_CUDA, "cudaMalloc of the camera state failed");
    }
    // the deepest Inside()/sturm paths keep a few small arrays per thread; blobs add their per-ray interval lists
    cudaDeviceSetLimit(cudaLimitStackSize, s.blobs.empty() ? 4096 : 12288);
    d->camera_dirty = true;
    // work contexts (streams, queues) of this device
    int n_ctx = 1;
    if (const char* e = getenv("PVGPU_CTX_PER_DEVICE")) n_ctx = std::max(1, std::min(4, atoi(e)));
    for (int i = 0; i < n_ctx; i++) {
        std::unique_ptr<WorkCtx> c(new WorkCtx());
        c->device = device;
        // the shadow stream is a background filler: its kernels give way to k_closest / k_shade of the next wave (critical path)
        int prio_low = 0, prio_high = 0;
        cudaDeviceGetStreamPriorityRange(&prio_low, &prio_high);
        bool ok = cudaStreamCreateWithPriority(&c->s_main, cudaStreamNonBlocking, prio_high) == cudaSuccess &&
                  cudaStreamCreateWithPriority(&c->s_shadow, cudaStreamNonBlocking, prio_low) == cudaSuccess &&
                  cudaMalloc(&c->cnt, 2 * sizeof(Counters)) == cudaSuccess &&
                  cudaMalloc(&c->ring, PV_RING_SLOTS * sizeof(WaveCounts)) == cudaSuccess &&
                  cudaMallocHost(&c->h_counts, PV_RING_SLOTS * sizeof(unsigned int)) == cudaSuccess;
        if (ok && v.has_area_lights)
            ok = cudaMalloc(&c->area_grid, (size_t)area_threads() * v.area_grid_max * 3 * sizeof(float)) == cudaSuccess;
        d->ctx.push_back(std::move(c));
        if (!ok) { release_one(d); return fail(PVGPU_E_CUDA, "allocation of the work context failed: %s", cudaGetErrorString(cudaGetLastError())); }
    }
    mark("work contexts");
    if (trace_t) fprintf(stderr, "pvgpu upload: total %.3f s\n", now_s() - t_begin);
    out = d;
    return PVGPU_OK;
}

static std::mutex g_prewarm_mutex;
static std::thread g_prewarm_thread;
static struct PrewarmGuard { ~PrewarmGuard() { if (g_prewarm_thread.joinable()) g_prewarm_thread.join(); } } g_prewarm_guard;

static void join_prewarm()
{
    std::lock_guard<std::mutex> lock(g_prewarm_mutex);
    if (g_prewarm_thread.joinable()) g_prewarm_thread.join();
}

int device_upload(Scene& s, const int* devices, int n_devices)
{
    join_prewarm();
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
        return fail(PVGPU_E_NO_DEVICE, "no CUDA device available (pvgpu has no CPU fallback)");
    std::vector<int> list;
    if (n_devices <= 0) for (int i = 0; i < ndev; i++) list.push_back(i);                 // all visible devices
    else for (int i = 0; i < n_devices; i++) list.push_back(devices ? devices[i] : i);
    for (size_t i = 0; i < list.size(); i++) {
        if (list[i] < 0 || list[i] >= ndev) return fail(PVGPU_E_INVALID, "device %d out of range (have %d)", list[i], ndev);
        for (size_t j = 0; j < i; j++) if (list[j] == list[i]) return fail(PVGPU_E_INVALID, "device %d listed twice", list[i]);
    }
    for (int dev : list) {
        DeviceScene* d = nullptr;
        int rc = upload_one(s, dev, d);
        if (rc != PVGPU_OK) { device_release(s); return rc; }
        s.devs.push_back(d);
    }
    s.dev = s.devs[0];
    s.device = list[0];
    s.device_bytes = s.devs[0]->bytes;
    // finished tiles travel to the first device of the list (or to the host) with peer copies
    for (size_t i = 1; i < list.size(); i++) {
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, list[i], list[0]) == cudaSuccess && can) {
            cudaSetDevice(list[i]);
            if (cudaDeviceEnablePeerAccess(list[0], 0) != cudaSuccess) cudaGetLastError();
            cudaSetDevice(list[0]);
            if (cudaDeviceEnablePeerAccess(list[i], 0) != cudaSuccess) cudaGetLastError();
        }
    }
    cudaSetDevice(list[0]);
    return PVGPU_OK;
}

// ------------------------------------------------------------------------------------------------
// host orchestration
// ------------------------------------------------------------------------------------------------
enum { KIND_PRIMARY = 0, KIND_CLOSEST = 1, KIND_SHADE = 2, KIND_SHADOW = 3, KIND_AA = 4, KIND_COUNT = 5 };

static size_t next_event(WorkCtx& c)
{
    if (c.ev_used == c.ev_pool.size()) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        c.ev_pool.push_back(e);
    }
    return c.ev_used++;
}

// Brackets one kernel launch with CUDA events on the launching stream; the elapsed times are summed per kernel
// kind when the frame is complete (pvgpu_stats::kernel_ms).
struct TimedLaunch {
    WorkCtx& c; cudaStream_t st; size_t e0;
    int kind; unsigned long long items;
    TimedLaunch(WorkCtx& c_, cudaStream_t st_, int kind_, unsigned long long items_) : c(c_), st(st_), kind(kind_), items(items_)
    {
        e0 = next_event(c);
        cudaEventRecord(c.ev_pool[e0], st);
    }
    ~TimedLaunch()
    {
        size_t e1 = next_event(c);
        cudaEventRecord(c.ev_pool[e1], st);
        c.timed.push_back({ kind, e0, e1, items });
        c.kernel_launches++;
    }
};

static int ensure_work_buffers(WorkCtx& c, size_t q_cap, size_t sq_cap, size_t n_rects)
{
    if (q_cap > c.q_cap) {
        for (int k = 0; k < 4; k++) { cudaFree(c.q[k]); c.q[k] = nullptr; }
        cudaFree(c.hits); c.hits = nullptr; c.q_cap = 0;
        for (int k = 0; k < 4; k++) CUDA_TRY(cudaMalloc(&c.q[k], q_cap * sizeof(PRay)));
        CUDA_TRY(cudaMalloc(&c.hits, q_cap * sizeof(HitRec)));
        c.q_cap = q_cap;
    }
    if (sq_cap > c.sq_cap) {
        for (int k = 0; k < 3; k++) { cudaFree(c.sq[k]); c.sq[k] = nullptr; }
        c.sq_cap = 0;
        for (int k = 0; k < 3; k++) CUDA_TRY(cudaMalloc(&c.sq[k], sq_cap * sizeof(SRay)));
        c.sq_cap = sq_cap;
    }
    if (n_rects > c.rect_cap) {
        cudaFree(c.rects); cudaFree(c.rect_off); c.rects = nullptr; c.rect_off = nullptr; c.rect_cap = 0;
        CUDA_TRY(cudaMalloc(&c.rects, n_rects * sizeof(pvgpu_rect)));
        CUDA_TRY(cudaMalloc(&c.rect_off, (n_rects + 1) * sizeof(uint32_t)));
        c.rect_cap = n_rects;
    }
    return PVGPU_OK;
}

// camera-dependent part of the device view (camera interiors for the pinhole case)
// TracePixel::SetupCamera (tracepixel.cpp:235-309): normalised axes, aspectRatio and axis lengths of the non-pinhole cameras
static void setup_camera(const Scene& s, DScene& v)
{
        auto len3 = [](const double* a) { return std::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); };
        auto norm3 = [&](const double* a, double* out) { double l = len3(a); for (int k = 0; k < 3; k++) out[k] = (l != 0.0) ? a[k] / l : a[k]; };
        v.cam_len_right = len3(s.camera.right);
        v.cam_len_up = len3(s.camera.up);
        bool normalise = true;
        switch (s.camera.type) {
            case PVGPU_CAMERA_CYL_1: case PVGPU_CAMERA_CYL_3: v.cam_aspect = v.cam_len_up; break;
            case PVGPU_CAMERA_CYL_2: case PVGPU_CAMERA_CYL_4: v.cam_aspect = v.cam_len_right; break;
            case PVGPU_CAMERA_ULTRA_WIDE_ANGLE: v.cam_aspect = v.cam_len_up / v.cam_len_right; break;
            case PVGPU_CAMERA_OMNIMAX: case PVGPU_CAMERA_FISHEYE: v.cam_aspect = v.cam_len_right / v.cam_len_up; break;
            default: v.cam_aspect = v.cam_len_right / v.cam_len_up; normalise = false; break;
        }
        for (int k = 0; k < 3; k++) { v.cam_right[k] = s.camera.right[k]; v.cam_up[k] = s.camera.up[k]; v.cam_dir[k] = s.camera.direction[k]; }
        if (normalise) { norm3(s.camera.right, v.cam_right); norm3(s.camera.up, v.cam_up); norm3(s.camera.direction, v.cam_dir); }
        v.cam_angle = s.camera_ext.size() == 3 ? s.camera_ext[0] : 0.0;
        v.cam_h_angle = s.camera_ext.size() == 3 ? s.camera_ext[1] : 0.0;
        v.cam_v_angle = s.camera_ext.size() == 3 ? s.camera_ext[2] : 0.0;
}

static int refresh_camera(Scene& s, DeviceScene& d, WorkCtx& c, cudaStream_t stream)
{
    d.view.cam = s.camera;
    d.view.n_cam_interiors = 0;
    setup_camera(s, d.view);
    // pinhole-style cameras: the containing interiors are found once (InitRayContainerState(ray, false)); orthographic and the
    // cylinder cameras 3 / 4 move the origin with the pixel and recompute per ray (k_primary)
    if (!s.interiors.empty() && s.camera.type != PVGPU_CAMERA_ORTHOGRAPHIC && s.camera.type != PVGPU_CAMERA_CYL_3 && s.camera.type != PVGPU_CAMERA_CYL_4) {
        launch_container_state(d.view, d.d_cam_int, c.cnt, stream);
        c.kernel_launches++;
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
static const size_t kShadowCapMax = 1u << 24;          // shadow-ray queue capacity limit (records per buffer)

struct FrameCtx {
    Scene& s;
    DeviceScene& d;
    WorkCtx& c;
    cudaStream_t stream;
    int width, height;
    float4* accum;
    pvgpu_stats st{};
    int (*cooperate)(void*);
    void* user;
    uint32_t cont_base = 0;          // first continuation slot in `accum` (scenes with a reflection exponent != 1): room for c.cont_cap slots
};

// Runs all waves for samples [first, first+n) of `src`.  Returns PVGPU_E_OVERFLOW if a queue was too small.
// PVGPU_SHADE_CSG=0: scenes of the `_csg` class shade with the plain heavy k_shade (A/B switch)
static bool shade_csg_enabled()
{
    static const bool on = [] { const char* e = getenv("PVGPU_SHADE_CSG"); return !(e && e[0] == '0'); }();
    return on;
}

static int run_batch(FrameCtx& f, const SampleSource& src, uint32_t first, uint32_t n)
{
    Scene& s = f.s;
    DeviceScene& d = f.d;
    WorkCtx& c = f.c;
    cudaStream_t S1 = f.stream, S2 = c.s_shadow;
    const uint32_t q_cap = (uint32_t)c.q_cap, sq_cap = (uint32_t)c.sq_cap;
    const uint32_t max_waves = std::min<uint32_t>(s.globals.max_trace_level + 64u, PV_RING_SLOTS - 2);     // continued rays do not consume a level
    const bool have_lights = !s.lights.empty();
    launch_wave_init(c.ring, PV_RING_SLOTS, n, S1);
    c.kernel_launches++;
    const bool conts = d.has_reflect_exp && c.conts != nullptr;
    if (conts) {
        CUDA_TRY(cudaMemsetAsync(&c.cnt->n_cont, 0, sizeof(unsigned int), S1));
        CUDA_TRY(cudaMemsetAsync(f.accum + f.cont_base, 0, c.cont_cap * sizeof(float4), S1));
    }
    {
        TimedLaunch t(c, S1, KIND_PRIMARY, n);
        launch_primary(d.view, src, first, n, (double)f.width, (double)f.height, c.q[0], c.cnt, f.accum, S1);
        if (d.view.cam.reserved) { launch_camera_normal_rays(d.view, src, first, n, (double)f.width, (double)f.height, c.q[0], S1); c.kernel_launches++; }
    }
    std::vector<size_t> ev_rb, ev_shadow;           // per wave: count read back / shadow kernels done
    unsigned long long bound = n;                   // upper bound of the wave's ray count (sizes the grids only)
    uint32_t known = n;                             // last ray count the host has seen
    uint32_t wave = 0;
    int rc = PVGPU_OK;
    for (;; wave++) {
        if (wave >= 2) {
            // n_rays of wave - 1 (read back after k_shade of wave - 2, which finished while wave - 1 was being traced)
            CUDA_TRY(cudaEventSynchronize(c.ev_pool[ev_rb[wave - 2]]));
            known = c.h_counts[wave - 2];
            if (known == 0) break;                  // wave - 1 was empty, so is this one
            bound = std::min<unsigned long long>(bound, (unsigned long long)std::min(known, q_cap) * std::max<uint32_t>(d.spawn_factor, 1u));
            if (f.cooperate && f.cooperate(f.user)) { rc = fail(PVGPU_E_ABORTED, "render aborted by the cooperate callback"); break; }
        }
        if (wave >= max_waves) { rc = fail(PVGPU_E_OVERFLOW, "rays still alive after %u waves (max_trace_level %u)", wave, s.globals.max_trace_level); break; }
        WaveCounts* wc = c.ring + wave;
        PRay* cur = c.q[wave % 4];
        const uint32_t nb = (uint32_t)std::min<unsigned long long>(std::max<unsigned long long>(bound, 1), q_cap);
        WaveCtx ctx;
        ctx.accum = f.accum; ctx.next = c.q[(wave + 1) % 4]; ctx.shadow = c.sq[wave % 3]; ctx.cnt = c.cnt;
        ctx.n_next = &wc[1].n_rays; ctx.n_shadow = &wc->n_shadow;
        ctx.cur_cap = q_cap; ctx.next_cap = q_cap; ctx.shadow_cap = sq_cap;
        ctx.conts = conts ? c.conts : nullptr; ctx.cont_cap = (uint32_t)c.cont_cap; ctx.cont_base = f.cont_base; ctx.wave = wave;
        {
            TimedLaunch t(c, S1, KIND_CLOSEST, 0);
            (d.lean ? launch_closest_lean : d.csg ? launch_closest_csg : d.quartic ? launch_closest_quartic : launch_closest)(d.view, cur, wc, nb, q_cap, c.hits, c.cnt, S1);
        }
        // k_shade of this wave writes sq[wave % 3] and q[(wave + 1) % 4]: k_shadow_* of wave - 3 read both (queue and parent rays)
        if (wave >= 3 && have_lights) CUDA_TRY(cudaStreamWaitEvent(S1, c.ev_pool[ev_shadow[wave - 3]], 0));
        {
            TimedLaunch t(c, S1, KIND_SHADE, 0);
            (d.lean ? launch_shade_lean : d.full ? launch_shade_full : (d.csg && shade_csg_enabled()) ? launch_shade_csg : launch_shade)(d.view, cur, c.hits, wc, nb, ctx, S1);
        }
        CUDA_TRY(cudaMemcpyAsync(&c.h_counts[wave], &wc[1].n_rays, sizeof(unsigned int), cudaMemcpyDeviceToHost, S1));
        ev_rb.push_back(next_event(c));
        CUDA_TRY(cudaEventRecord(c.ev_pool[ev_rb.back()], S1));
        if (have_lights) {
            CUDA_TRY(cudaStreamWaitEvent(S2, c.ev_pool[ev_rb.back()], 0));
            const uint32_t sb = (uint32_t)std::min<unsigned long long>((unsigned long long)nb * d.shadow_factor, sq_cap);
            {
                TimedLaunch t(c, S2, KIND_SHADOW, 0);
                if (d.view.all_opaque) (d.lean ? launch_shadow_opaque_lean : d.csg ? launch_shadow_opaque_csg : d.quartic ? launch_shadow_opaque_quartic : launch_shadow_opaque)(d.view, c.sq[wave % 3], wc, sb, sq_cap, f.accum, c.cnt, S2);
                else (d.lean ? launch_shadow_filter_lean : d.full ? launch_shadow_filter_full : d.csg ? launch_shadow_filter_csg : d.quartic ? launch_shadow_filter_quartic : launch_shadow_filter)(d.view, c.sq[wave % 3], wc, sb, sq_cap, cur, f.accum, c.cnt, S2);
                if (d.view.has_area_lights) { c.kernel_launches++; launch_shadow_area(d.view, c.sq[wave % 3], wc, sb, sq_cap, cur, f.accum, c.cnt, c.area_grid, S2); }
            }
            ev_shadow.push_back(next_event(c));
            CUDA_TRY(cudaEventRecord(c.ev_pool[ev_shadow.back()], S2));
        }
        f.st.waves++;
        if (d.spawn_factor == 0) { wave++; break; }         // no reflective or refractive material: the frame is this one wave
        bound = std::min<unsigned long long>(bound * d.spawn_factor, q_cap);
    }
    // the shadow stream joins the main stream; queue overflows are known once everything has run
    if (have_lights) for (size_t k = (ev_shadow.size() > 3 ? ev_shadow.size() - 3 : 0); k < ev_shadow.size(); k++) CUDA_TRY(cudaStreamWaitEvent(S1, c.ev_pool[ev_shadow[k]], 0));
    if (rc != PVGPU_OK) { cudaStreamSynchronize(S1); return rc; }
    if (conts)       // everything has been gathered: fold the continuation slots into their parents, last wave first
        for (uint32_t w = wave; w-- > 0;) { launch_resolve_conts(f.accum, c.conts, c.cnt, (uint32_t)c.cont_cap, f.cont_base, w, S1); c.kernel_launches++; }
    unsigned int ovf = 0;
    CUDA_TRY(cudaMemcpyAsync(&ovf, &c.cnt->overflow, sizeof ovf, cudaMemcpyDeviceToHost, S1));
    CUDA_TRY(cudaStreamSynchronize(S1));
    if (ovf & (8u | 16u)) return PVGPU_E_OVERFLOW;
    return PVGPU_OK;
}

// Traces samples [0, n) of `src` into the accumulators: batches of kBatchSamples; a batch whose ray queues overflow is
// retried as two halves (its accumulator slots are cleared by recomputing them, k_clear_slots), unless several samples share
// one slot (`clear_slots` = false, anti-aliasing sums) - then an overflow is fatal.
static int trace_samples(FrameCtx& f, const SampleSource& src, uint32_t n_samples, bool clear_slots)
{
    DeviceScene& d = f.d;
    WorkCtx& c = f.c;
    if (n_samples == 0) return PVGPU_OK;
    const size_t batch = std::min<size_t>(kBatchSamples, n_samples);
    {
        // a wave cannot outgrow the batch when a ray spawns at most one ray; otherwise leave room for growth
        size_t q_cap = std::max<size_t>((d.spawn_factor <= 1 ? 1 : 4) * batch, 1024);
        size_t sq_cap = std::min<size_t>(kShadowCapMax, std::max<size_t>(q_cap * d.shadow_factor, 1024));
        if (const char* e = getenv("PVGPU_TEST_QUEUE_CAP")) {       // tests: force small queues to exercise the clamp + overflow retry
            const size_t cap = (size_t)std::max(64, atoi(e));
            if (c.q_cap == 0) q_cap = std::min(q_cap, cap);
            if (c.sq_cap == 0) sq_cap = std::min(sq_cap, cap);
        }
        int rc = ensure_work_buffers(c, q_cap, f.s.lights.empty() ? 0 : sq_cap, 0);
        if (rc != PVGPU_OK) return rc;
        if (d.has_reflect_exp && c.conts == nullptr) return fail(PVGPU_E_INVALID, "internal: continuation records not allocated");
    }
    struct Span { uint32_t first, n; };
    std::vector<Span> todo;
    for (uint32_t b = 0; b < n_samples; b += (uint32_t)batch) todo.push_back({ b, (uint32_t)std::min<size_t>(batch, n_samples - b) });
    std::reverse(todo.begin(), todo.end());
    while (!todo.empty()) {
        Span sp = todo.back(); todo.pop_back();
        if (f.cooperate && f.cooperate(f.user)) return fail(PVGPU_E_ABORTED, "render aborted by the cooperate callback");
        CUDA_TRY(cudaMemcpyAsync(c.cnt + 1, c.cnt, sizeof(Counters), cudaMemcpyDeviceToDevice, f.stream));
        int rc = run_batch(f, src, sp.first, sp.n);
        if (rc == PVGPU_E_OVERFLOW && sp.n > 1 && clear_slots) {
            // the batch is traced again in two halves: its pixels and its share of the statistics are taken back
            launch_clear_slots(src, sp.first, sp.n, f.accum, f.stream);
            c.kernel_launches++;
            CUDA_TRY(cudaMemcpyAsync(c.cnt, c.cnt + 1, sizeof(Counters), cudaMemcpyDeviceToDevice, f.stream));
            todo.push_back({ sp.first + sp.n / 2, sp.n - sp.n / 2 });
            todo.push_back({ sp.first, sp.n / 2 });
            continue;
        }
        if (rc != PVGPU_OK) return rc == PVGPU_E_OVERFLOW ? fail(rc, "ray queue overflow") : rc;
    }
    return PVGPU_OK;
}

// Accumulator slots to add behind the slots of a call for the continuation records of scenes with a reflection exponent != 1
static size_t cont_extra(const DeviceScene& d, const WorkCtx& c) { return d.has_reflect_exp ? c.cont_cap : 0; }

// Device scratch of one anti-aliased call, freed on scope exit.  Stream-ordered allocations from the device's default pool, whose
// release threshold device_upload raises: after the first frame the buffers of a call (GBs at 1080p +R3) come out of the pool
// instead of cudaMalloc / cudaFree, which cost config 3 with AA half a second of idle GPU per frame.  The buffers are freed in the
// order of `stream` after it has been made to wait for the shadow stream (a wave loop that returns early with an error has not
// joined the two).
struct Scratch {
    cudaStream_t stream, other;
    std::vector<void*> ptrs;
    Scratch(cudaStream_t s, cudaStream_t o) : stream(s), other(o) {}
    ~Scratch()
    {
        if (ptrs.empty()) return;
        cudaEvent_t ev;
        if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) == cudaSuccess) {
            if (cudaEventRecord(ev, other) != cudaSuccess || cudaStreamWaitEvent(stream, ev, 0) != cudaSuccess) { cudaGetLastError(); cudaStreamSynchronize(other); }
            cudaEventDestroy(ev);
        } else { cudaGetLastError(); cudaStreamSynchronize(other); }
        for (void* p : ptrs) cudaFreeAsync(p, stream);
    }
    template <class T> int alloc(T*& out, size_t n)
    {
        void* p = nullptr;
        cudaError_t e = cudaMallocAsync(&p, std::max<size_t>(n, 1) * sizeof(T), stream);
        if (e != cudaSuccess) return fail(PVGPU_E_CUDA, "cudaMallocAsync of %zu bytes failed: %s", n * sizeof(T), cudaGetErrorString(e));
        ptrs.push_back(p);
        out = reinterpret_cast<T*>(p);
        return PVGPU_OK;
    }
};
#define AA_TRY(expr) do { int rc__ = (expr); if (rc__ != PVGPU_OK) return rc__; } while (0)

static int read_counter(const unsigned int* d_ptr, cudaStream_t stream, unsigned int& out)
{
    CUDA_TRY(cudaMemcpyAsync(&out, d_ptr, sizeof(unsigned int), cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    return PVGPU_OK;
}

static AAParams make_aa_params(const pvgpu_aa& aa)
{
    AAParams p{};
    p.method = (int)aa.method;
    p.depth = aa.depth;
    p.threshold = aa.threshold;
    // jitterScale = jitterScale / aaDepth (tracetask.cpp:526) resp. / ((1 << aaDepth) + 1) (tracetask.cpp:611)
    p.jitter_scale = (aa.method == 1) ? aa.jitter_scale / (double)aa.depth : aa.jitter_scale / (double)((1u << aa.depth) + 1u);
    p.neutral = !(aa.gamma > 0.0 && aa.gamma != 1.0);
    p.enc_gamma = p.neutral ? 1.0f : 1.0f / (float)aa.gamma;       // PowerLawGammaCurve::GetByDecodingGamma (colourspace.cpp:306)
    return p;
}

// NonAdaptiveSupersamplingM1 (tracetask.cpp:521-602) for all rectangles of the call; see k_aa.cu for the scheme.
static int render_aa1(FrameCtx& f, const pvgpu_aa& aa, const pvgpu_rect* rects, size_t n_rects, const std::vector<uint32_t>& off, float4* d_out,
                      unsigned long long& n_extra_samples)
{
    DeviceScene& d = f.d;
    WorkCtx& c = f.c;
    cudaStream_t stream = f.stream;
    const uint32_t n_px = off[n_rects];
    std::vector<uint32_t> foff(n_rects + 1, 0);
    for (size_t i = 0; i < n_rects; i++)
        foff[i + 1] = foff[i] + (uint32_t)(rects[i].right - rects[i].left + 1) + (uint32_t)(rects[i].bottom - rects[i].top + 1);
    const uint32_t n_frame = foff[n_rects];
    if ((unsigned long long)n_px * 2 + n_frame > 0xFFFFFFF0ull) return fail(PVGPU_E_INVALID, "too many pixels in one anti-aliased call");
    // SupersampleOnePixel's offsets: the reference's own floating-point loop (tracetask.cpp:862-869)
    std::vector<double2> offsets;
    {
        const double step = 1.0 / (double)aa.depth, range = 0.5 - (step * 0.5);
        for (double yy = -range; yy <= (range + PV_EPSILON); yy += step)
            for (double xx = -range; xx <= (range + PV_EPSILON); xx += step) offsets.push_back(make_double2(xx, yy));
    }
    const uint32_t n_off = (uint32_t)offsets.size();
    const AAParams ap = make_aa_params(aa);

    Scratch sc(stream, c.s_shadow);
    float4* accum = nullptr; uint32_t* d_foff = nullptr; double2* d_fcoords = nullptr; int32_t* s_slot = nullptr; uint32_t* cand = nullptr;
    unsigned int* counters = nullptr; uint8_t* flag = nullptr; double2* d_offsets = nullptr; double2* d_coords = nullptr; uint32_t* d_slots = nullptr;
    const size_t n_slots = (size_t)n_px * 2 + n_frame;
    const uint32_t cand_chunk = std::max<uint32_t>(1u, (uint32_t)(kBatchSamples / n_off));
    AA_TRY(sc.alloc(accum, n_slots + cont_extra(d, c))); AA_TRY(sc.alloc(d_foff, n_rects + 1)); AA_TRY(sc.alloc(d_fcoords, n_frame));
    AA_TRY(sc.alloc(s_slot, n_px)); AA_TRY(sc.alloc(cand, n_px)); AA_TRY(sc.alloc(counters, 4)); AA_TRY(sc.alloc(flag, n_px));
    AA_TRY(sc.alloc(d_offsets, n_off)); AA_TRY(sc.alloc(d_coords, (size_t)cand_chunk * n_off)); AA_TRY(sc.alloc(d_slots, (size_t)cand_chunk * n_off));
    CUDA_TRY(cudaMemsetAsync(accum, 0, n_slots * sizeof(float4), stream));
    CUDA_TRY(cudaMemsetAsync(counters, 0, 4 * sizeof(unsigned int), stream));
    CUDA_TRY(cudaMemcpyAsync(d_foff, foff.data(), foff.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, stream));
    CUDA_TRY(cudaMemcpyAsync(d_offsets, offsets.data(), n_off * sizeof(double2), cudaMemcpyHostToDevice, stream));
    f.accum = accum; f.cont_base = (uint32_t)n_slots;
    AALayout L{};
    L.rects = c.rects; L.rect_off = c.rect_off; L.frame_off = d_foff; L.corner_off = nullptr;
    L.n_rects = (uint32_t)n_rects; L.n_px = n_px; L.n_frame = n_frame; L.n_corner = 0; L.s_base = n_px + n_frame;

    // 1. pixel centres, 2. the frame above / left of every rectangle
    SampleSource src{};
    src.rects = c.rects; src.rect_off = c.rect_off; src.n_rects = (uint32_t)n_rects;
    AA_TRY(trace_samples(f, src, n_px, true));
    { TimedLaunch t(c, stream, KIND_AA, n_frame); launch_aa1_frame_coords(L, d_fcoords, stream); }
    SampleSource fsrc{};
    fsrc.coords = d_fcoords; fsrc.slot_base = n_px;
    AA_TRY(trace_samples(f, fsrc, n_frame, true));
    // 3. candidates from un-supersampled colours
    { TimedLaunch t(c, stream, KIND_AA, n_px); launch_aa1_candidates(L, ap, accum, s_slot, cand, counters, stream); }
    unsigned int n_cand = 0, n_done = 0;
    AA_TRY(read_counter(counters, stream, n_cand));
    // 4. trace what is queued, replay the sequential walk, repeat while the walk queues more
    for (int round = 0; round < 64; round++) {
        while (n_done < n_cand) {
            const uint32_t cn = std::min<uint32_t>(cand_chunk, n_cand - n_done);
            { TimedLaunch t(c, stream, KIND_AA, cn); launch_aa1_sample_coords(L, ap, d.view.noise.hash, cand, n_done, cn, d_offsets, n_off, d_coords, d_slots, stream); }
            SampleSource ssrc{};
            ssrc.coords = d_coords; ssrc.slots = d_slots;
            AA_TRY(trace_samples(f, ssrc, cn * n_off, false));
            n_done += cn;
            n_extra_samples += (unsigned long long)cn * n_off;
        }
        CUDA_TRY(cudaMemsetAsync(counters + 1, 0, sizeof(unsigned int), stream));
        { TimedLaunch t(c, stream, KIND_AA, n_rects); launch_aa1_decide(L, ap, accum, s_slot, cand, counters, d_out, flag, counters + 1, stream); }
        AA_TRY(read_counter(counters, stream, n_cand));
        if (n_cand == n_done) return PVGPU_OK;
    }
    return fail(PVGPU_E_OVERFLOW, "anti-aliasing method 1 did not settle");
}

// AdaptiveSupersamplingM2 (tracetask.cpp:604-657, 892-1074) for all rectangles of the call; see k_aa.cu.
static int render_aa2(FrameCtx& f, const pvgpu_aa& aa, const pvgpu_rect* rects, size_t n_rects, const std::vector<uint32_t>& off, float4* d_out,
                      unsigned long long& n_extra_samples)
{
    DeviceScene& d = f.d;
    WorkCtx& c = f.c;
    cudaStream_t stream = f.stream;
    const uint32_t n_px = off[n_rects];
    std::vector<uint32_t> coff(n_rects + 1, 0);
    for (size_t i = 0; i < n_rects; i++) {
        unsigned long long c = (unsigned long long)(rects[i].right - rects[i].left + 2) * (unsigned long long)(rects[i].bottom - rects[i].top + 2);
        if (coff[i] + c > 0xFFFFFFF0ull) return fail(PVGPU_E_INVALID, "too many pixels in one anti-aliased call");
        coff[i + 1] = coff[i] + (uint32_t)c;
    }
    const uint32_t n_corner = coff[n_rects];
    const AAParams ap = make_aa_params(aa);
    const uint32_t S1 = (1u << aa.depth) + 1u, per = S1 * S1, words = (per + 31) / 32;

    Scratch sc(stream, c.s_shadow);
    float4* corners = nullptr; uint32_t* d_coff = nullptr; double2* d_ccoords = nullptr; int32_t* act_idx = nullptr; uint32_t* act_list = nullptr;
    unsigned int* counters = nullptr;
    AA_TRY(sc.alloc(corners, n_corner + cont_extra(d, c))); AA_TRY(sc.alloc(d_coff, n_rects + 1)); AA_TRY(sc.alloc(d_ccoords, n_corner));
    AA_TRY(sc.alloc(act_idx, n_px)); AA_TRY(sc.alloc(act_list, n_px)); AA_TRY(sc.alloc(counters, 4));
    CUDA_TRY(cudaMemsetAsync(corners, 0, (size_t)n_corner * sizeof(float4), stream));
    CUDA_TRY(cudaMemsetAsync(counters, 0, 4 * sizeof(unsigned int), stream));
    CUDA_TRY(cudaMemcpyAsync(d_coff, coff.data(), coff.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, stream));
    AALayout L{};
    L.rects = c.rects; L.rect_off = c.rect_off; L.frame_off = nullptr; L.corner_off = d_coff;
    L.n_rects = (uint32_t)n_rects; L.n_px = n_px; L.n_frame = 0; L.n_corner = n_corner; L.s_base = n_corner;

    // 1. pixel corners
    { TimedLaunch t(c, stream, KIND_AA, n_corner); launch_aa2_corner_coords(L, d_ccoords, stream); }
    f.accum = corners; f.cont_base = n_corner;
    SampleSource csrc{};
    csrc.coords = d_ccoords;
    AA_TRY(trace_samples(f, csrc, n_corner, true));
    // 2. pixels that subdivide get a sample buffer
    { TimedLaunch t(c, stream, KIND_AA, n_px); launch_aa2_mark(L, ap, corners, act_idx, act_list, counters, stream); }
    unsigned int n_active = 0;
    AA_TRY(read_counter(counters, stream, n_active));
    // The sample buffers ((2^depth + 1)^2 slots per subdividing pixel, like the reference's SubdivisionBuffer) of all subdividing
    // pixels are held at once when they fit the slot budget; deep levels (+R6 .. +R9: 4 k .. 263 k slots per pixel) work the pixels
    // off in groups, each with its own buffers, rounds and resolve pass - the result does not depend on the grouping.
    unsigned long long kSlotBudget = 1ull << 26;                // 1 GB of accumulators
    if (const char* e = getenv("PVGPU_TEST_AA2_SLOTS")) kSlotBudget = (unsigned long long)std::max(1, atoi(e));      // tests: force the grouped path
    const uint32_t group_max = (uint32_t)std::max<unsigned long long>(1, std::min<unsigned long long>(n_active ? n_active : 1, (kSlotBudget > n_corner ? kSlotBudget - n_corner : 1) / per));
    const bool grouped = n_active > group_max;
    auto refine = [&](float4* accum, uint32_t* sampled, const uint32_t* list, uint32_t n_list) -> int {
        // one tracing round per subdivision level for the pixels of `list`
        unsigned long long per_pixel = 5;
        for (uint32_t round = 0; round + 1 < aa.depth; round++, per_pixel = std::min<unsigned long long>(per_pixel * 4, per)) {
            const unsigned long long cap64 = std::min<unsigned long long>((unsigned long long)n_list * per_pixel, 0xFFFFFFF0ull);
            const uint32_t cap = (uint32_t)cap64;
            double2* d_coords = nullptr; uint32_t* d_slots = nullptr;
            Scratch rs(stream, c.s_shadow);
            AA_TRY(rs.alloc(d_coords, cap)); AA_TRY(rs.alloc(d_slots, cap));
            CUDA_TRY(cudaMemsetAsync(counters + 1, 0, sizeof(unsigned int), stream));
            { TimedLaunch t(c, stream, KIND_AA, n_list); launch_aa2_expand(L, ap, d.view.noise.hash, accum, list, n_list, (int)round, sampled, d_coords, d_slots, counters + 1, cap, stream); }
            unsigned int n_new = 0;
            AA_TRY(read_counter(counters + 1, stream, n_new));
            if (n_new > cap) return fail(PVGPU_E_OVERFLOW, "anti-aliasing method 2: sample list overflow");
            if (n_new == 0) break;
            SampleSource ssrc{};
            ssrc.coords = d_coords; ssrc.slots = d_slots;
            AA_TRY(trace_samples(f, ssrc, n_new, false));
            n_extra_samples += n_new;
            CUDA_TRY(cudaStreamSynchronize(stream));
        }
        return PVGPU_OK;
    };
    if (!grouped) {
        float4* accum = corners;
        if (n_active) {
            const unsigned long long n_slots = (unsigned long long)n_corner + (unsigned long long)n_active * per;
            if (n_slots > 0xFFFFFFF0ull) return fail(PVGPU_E_OVERFLOW, "anti-aliasing method 2: %u subdividing pixels x %u samples exceed the slot range; render fewer rectangles per call", n_active, per);
            uint32_t* sampled = nullptr;
            AA_TRY(sc.alloc(accum, (size_t)n_slots + cont_extra(d, c))); AA_TRY(sc.alloc(sampled, (size_t)n_active * words));
            CUDA_TRY(cudaMemsetAsync(accum, 0, (size_t)n_slots * sizeof(float4), stream));
            CUDA_TRY(cudaMemcpyAsync(accum, corners, (size_t)n_corner * sizeof(float4), cudaMemcpyDeviceToDevice, stream));
            CUDA_TRY(cudaMemsetAsync(sampled, 0, (size_t)n_active * words * sizeof(uint32_t), stream));
            f.accum = accum; f.cont_base = (uint32_t)n_slots;
            AA_TRY(refine(accum, sampled, act_list, n_active));
        }
        // 4. combine
        { TimedLaunch t(c, stream, KIND_AA, n_px); launch_aa2_resolve(L, ap, accum, act_idx, d_out, stream); }
        CUDA_TRY(cudaStreamSynchronize(stream));
        return PVGPU_OK;
    }
    // pixels that do not subdivide: the mean of their corners
    { TimedLaunch t(c, stream, KIND_AA, n_px); launch_aa2_resolve(L, ap, corners, act_idx, d_out, stream, nullptr, 0, 1); }
    const unsigned long long g_slots = (unsigned long long)n_corner + (unsigned long long)group_max * per;
    if (g_slots > 0xFFFFFFF0ull) return fail(PVGPU_E_OVERFLOW, "anti-aliasing method 2: sample buffer of one pixel exceeds the slot range");
    float4* accum = nullptr; uint32_t* sampled = nullptr;
    AA_TRY(sc.alloc(accum, (size_t)g_slots + cont_extra(d, c))); AA_TRY(sc.alloc(sampled, (size_t)group_max * words));
    f.accum = accum; f.cont_base = (uint32_t)g_slots;
    for (uint32_t a0 = 0; a0 < n_active; a0 += group_max) {
        const uint32_t ng = std::min(group_max, n_active - a0);
        CUDA_TRY(cudaMemsetAsync(accum, 0, ((size_t)n_corner + (size_t)ng * per) * sizeof(float4), stream));
        CUDA_TRY(cudaMemcpyAsync(accum, corners, (size_t)n_corner * sizeof(float4), cudaMemcpyDeviceToDevice, stream));
        CUDA_TRY(cudaMemsetAsync(sampled, 0, (size_t)ng * words * sizeof(uint32_t), stream));
        AA_TRY(refine(accum, sampled, act_list + a0, ng));
        { TimedLaunch t(c, stream, KIND_AA, ng); launch_aa2_resolve(L, ap, accum, act_idx, d_out, stream, act_list + a0, ng, 0); }
        if (f.cooperate && f.cooperate(f.user)) return fail(PVGPU_E_ABORTED, "render aborted by the cooperate callback");
    }
    CUDA_TRY(cudaStreamSynchronize(stream));
    return PVGPU_OK;
}

static int render_impl(Scene& s, DeviceScene& d, WorkCtx& c, const pvgpu_aa* aa, int width, int height, const pvgpu_rect* rects, size_t n_rects,
                       float* d_out, pvgpu_stats* stats, cudaStream_t stream, int (*cooperate)(void*), void* user)
{
    if (width <= 0 || height <= 0 || !rects || !n_rects || !d_out) return fail(PVGPU_E_INVALID, "pvgpu_render: bad arguments");
    const unsigned int method = aa ? aa->method : 0u;
    if (method > 2) return fail(PVGPU_E_UNSUPPORTED, "anti-aliasing method %u (stochastic supersampling) is outside the GPU trace path", method);
    if (method && (aa->depth < 1 || aa->depth > 9)) return fail(PVGPU_E_INVALID, "anti-aliasing depth %u out of range 1..9", aa->depth);
    CUDA_TRY(cudaSetDevice(d.device));
    std::vector<uint32_t> off(n_rects + 1, 0);
    for (size_t i = 0; i < n_rects; i++) {
        const pvgpu_rect& r = rects[i];
        if (r.right < r.left || r.bottom < r.top) return fail(PVGPU_E_INVALID, "rectangle %zu is empty", i);
        unsigned long long area = (unsigned long long)(r.right - r.left + 1) * (unsigned long long)(r.bottom - r.top + 1);
        if (off[i] + area > 0xFFFFFFF0ull) return fail(PVGPU_E_INVALID, "too many pixels in one call");
        off[i + 1] = off[i] + (uint32_t)area;
    }
    const uint32_t n_samples = off[n_rects];
    const unsigned long long launches0 = c.kernel_launches;
    c.ev_used = 0;
    c.timed.clear();
    const size_t ev0 = next_event(c);
    CUDA_TRY(cudaEventRecord(c.ev_pool[ev0], stream));
    CUDA_TRY(cudaMemsetAsync(c.cnt, 0, sizeof(Counters), stream));
    {
        std::lock_guard<std::mutex> cam_lock(d.camera_mutex);
        if (d.camera_dirty || std::memcmp(&d.view.cam, &s.camera, sizeof s.camera) != 0) {
            int rc = refresh_camera(s, d, c, stream);
            if (rc != PVGPU_OK) return rc;
        }
    }
    int rc = ensure_work_buffers(c, 1024, 0, n_rects);
    if (rc != PVGPU_OK) return rc;
    CUDA_TRY(cudaMemcpyAsync(c.rects, rects, n_rects * sizeof(pvgpu_rect), cudaMemcpyHostToDevice, stream));
    CUDA_TRY(cudaMemcpyAsync(c.rect_off, off.data(), (n_rects + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, stream));
    CUDA_TRY(cudaMemsetAsync(d_out, 0, (size_t)n_samples * 4 * sizeof(float), stream));
    CUDA_TRY(cudaStreamSynchronize(stream));        // `off` and (possibly pageable) `rects` are consumed

    FrameCtx f{ s, d, c, stream, width, height, reinterpret_cast<float4*>(d_out), pvgpu_stats{}, cooperate, user };
    if (d.has_reflect_exp && c.conts == nullptr) {
        c.cont_cap = std::min<size_t>(1u << 23, std::max<size_t>(8 * (size_t)std::min<size_t>(kBatchSamples, n_samples), 65536));
        if (const char* e = getenv("PVGPU_TEST_QUEUE_CAP")) c.cont_cap = std::min<size_t>(c.cont_cap, (size_t)std::max(64, atoi(e)));
        CUDA_TRY(cudaMalloc(&c.conts, c.cont_cap * sizeof(Cont)));
    }
    unsigned long long n_extra_samples = 0;
    if (method == 0) {
        SampleSource src{};
        src.rects = c.rects; src.rect_off = c.rect_off; src.n_rects = (uint32_t)n_rects;
        if (d.has_reflect_exp) {
            // the frame is gathered in a buffer with room for the continuation slots behind the pixels, then copied out
            const size_t need = (size_t)n_samples + c.cont_cap;
            if (need > c.accum_ext_cap) {
                cudaFree(c.accum_ext); c.accum_ext = nullptr; c.accum_ext_cap = 0;
                CUDA_TRY(cudaMalloc(&c.accum_ext, need * sizeof(float4)));
                c.accum_ext_cap = need;
            }
            CUDA_TRY(cudaMemsetAsync(c.accum_ext, 0, (size_t)n_samples * sizeof(float4), stream));
            f.accum = c.accum_ext; f.cont_base = n_samples;
        }
        rc = trace_samples(f, src, n_samples, true);
        if (rc == PVGPU_OK && d.has_reflect_exp) CUDA_TRY(cudaMemcpyAsync(d_out, c.accum_ext, (size_t)n_samples * sizeof(float4), cudaMemcpyDeviceToDevice, stream));
    } else if (method == 1) rc = render_aa1(f, *aa, rects, n_rects, off, reinterpret_cast<float4*>(d_out), n_extra_samples);
    else rc = render_aa2(f, *aa, rects, n_rects, off, reinterpret_cast<float4*>(d_out), n_extra_samples);
    if (rc != PVGPU_OK) { cudaStreamSynchronize(stream); cudaStreamSynchronize(c.s_shadow); return rc; }

    Counters hc;
    CUDA_TRY(cudaMemcpyAsync(&hc, c.cnt, sizeof hc, cudaMemcpyDeviceToHost, stream));
    const size_t ev1 = next_event(c);
    CUDA_TRY(cudaEventRecord(c.ev_pool[ev1], stream));
    CUDA_TRY(cudaEventSynchronize(c.ev_pool[ev1]));
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, c.ev_pool[ev0], c.ev_pool[ev1]);
    pvgpu_stats& st = f.st;
    st.rays = hc.rays; st.shadow_ray_tests = hc.shadow_tests; st.reflected_rays = hc.reflected;
    st.refracted_rays = hc.refracted; st.transmitted_rays = hc.transmitted; st.tir_rays = hc.tir;
    st.adc_saves = hc.adc_saves; st.samples = n_extra_samples; st.max_trace_level = hc.max_level; st.overflow = hc.overflow;
    st.node_tests_closest = hc.node_tests[0]; st.prim_tests_closest = hc.prim_tests[0]; st.node_tests_shadow = hc.node_tests[1]; st.prim_tests_shadow = hc.prim_tests[1];
    st.kernel_launches = c.kernel_launches - launches0;
    st.device_ms = ms;
    for (const WorkCtx::Timed& t : c.timed) {
        float k = 0.0f;
        cudaEventElapsedTime(&k, c.ev_pool[t.e0], c.ev_pool[t.e1]);
        st.kernel_ms[t.kind] += k;
        st.kernel_count[t.kind] += 1;
        st.kernel_items[t.kind] += t.items;
    }
    if (getenv("PVGPU_TRACE_WAVES")) {       // development aid: every timed launch of the call, in launch order
        static const char* names[KIND_COUNT] = { "primary", "closest", "shade", "shadow", "aa" };
        for (const WorkCtx::Timed& t : c.timed) {
            float k = 0.0f, t0 = 0.0f;
            cudaEventElapsedTime(&k, c.ev_pool[t.e0], c.ev_pool[t.e1]);
            cudaEventElapsedTime(&t0, c.ev_pool[ev0], c.ev_pool[t.e0]);
            fprintf(stderr, "pvgpu wave trace: %-8s start %8.3f ms  dur %8.3f ms\n", names[t.kind], t0, k);
        }
        fprintf(stderr, "pvgpu wave trace: rays of waves 1..: ");
        for (uint32_t k = 0; k + 1 < (uint32_t)st.waves && k < 16; k++) fprintf(stderr, "%u ", c.h_counts[k]);
        fprintf(stderr, "| frame %.3f ms", ms);
        if (hc.pad) fprintf(stderr, " | most box tests of one ray (diagnostic build): %u", hc.pad);
        fprintf(stderr, "\n");
    }
    st.kernel_items[KIND_CLOSEST] = hc.rays;
    st.kernel_items[KIND_SHADE] = hc.rays;
    st.kernel_items[KIND_SHADOW] = hc.shadow_rays;
    if (stats) *stats = st;
    if (hc.overflow & ~(8u | 16u))
        return fail(PVGPU_E_OVERFLOW, "device capacity exceeded (flags 0x%x: 1 traversal stack, 2 mesh in CSG, 4 interior list, 32 blob components per ray, 128 glyph hit batches)", hc.overflow);
    return PVGPU_OK;
}

static void merge_stats(pvgpu_stats& a, const pvgpu_stats& b)
{
    a.rays += b.rays; a.shadow_ray_tests += b.shadow_ray_tests; a.reflected_rays += b.reflected_rays; a.refracted_rays += b.refracted_rays;
    a.transmitted_rays += b.transmitted_rays; a.tir_rays += b.tir_rays; a.adc_saves += b.adc_saves; a.samples += b.samples;
    a.waves += b.waves; a.kernel_launches += b.kernel_launches;
    a.node_tests_closest += b.node_tests_closest; a.prim_tests_closest += b.prim_tests_closest; a.node_tests_shadow += b.node_tests_shadow; a.prim_tests_shadow += b.prim_tests_shadow;
    a.max_trace_level = std::max(a.max_trace_level, b.max_trace_level); a.overflow |= b.overflow;
    for (int k = 0; k < 5; k++) { a.kernel_ms[k] += b.kernel_ms[k]; a.kernel_count[k] += b.kernel_count[k]; a.kernel_items[k] += b.kernel_items[k]; }
}

// D2H of `bytes` from a device buffer of context `c` into caller memory: one DMA when the destination is page-locked
// (pvgpu_host_alloc), else through the context's pinned staging buffer in slices that a few host threads copy out while the
// next ones are in flight.
static int deliver_to_host(WorkCtx& c, const float* d_src, float* dst, size_t bytes, cudaStream_t stream)
{
    if (bytes == 0) return PVGPU_OK;
    cudaPointerAttributes attr{};
    const bool pinned = cudaPointerGetAttributes(&attr, dst) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    cudaGetLastError();
    if (pinned) {
        cudaError_t e = cudaMemcpyAsync(dst, d_src, bytes, cudaMemcpyDeviceToHost, stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
        if (e != cudaSuccess) return fail(PVGPU_E_CUDA, "copy of the frame to the host failed: %s", cudaGetErrorString(e));
        return PVGPU_OK;
    }
    if (c.h_frame_cap < bytes) {
        if (c.h_frame) cudaFreeHost(c.h_frame);
        c.h_frame = nullptr; c.h_frame_cap = 0;
        CUDA_TRY(cudaMallocHost(&c.h_frame, bytes));
        c.h_frame_cap = bytes;
    }
    const size_t slice = 2u << 20;
    const size_t n_slices = (bytes + slice - 1) / slice;
    std::vector<cudaEvent_t> evs(n_slices, nullptr);
    cudaError_t e = cudaSuccess;
    for (size_t k = 0; k < n_slices && e == cudaSuccess; k++) {
        const size_t off = k * slice, len = std::min(slice, bytes - off);
        cudaEventCreateWithFlags(&evs[k], cudaEventDisableTiming);
        e = cudaMemcpyAsync(reinterpret_cast<char*>(c.h_frame) + off, reinterpret_cast<const char*>(d_src) + off, len, cudaMemcpyDeviceToHost, stream);
        cudaEventRecord(evs[k], stream);
    }
    if (e == cudaSuccess) {
        const int n_workers = (int)std::min<size_t>(4, n_slices);
        std::atomic<int> failed(0);
        const int device = c.device;
        auto worker = [&](int w) {
            cudaSetDevice(device);
            for (size_t k = (size_t)w; k < n_slices; k += (size_t)n_workers) {
                if (cudaEventSynchronize(evs[k]) != cudaSuccess) { failed = 1; return; }
                const size_t off = k * slice, len = std::min(slice, bytes - off);
                std::memcpy(reinterpret_cast<char*>(dst) + off, reinterpret_cast<const char*>(c.h_frame) + off, len);
            }
        };
        std::vector<std::thread> pool;
        for (int w = 1; w < n_workers; w++) pool.emplace_back(worker, w);
        worker(0);
        for (auto& t : pool) t.join();
        if (failed) e = cudaErrorUnknown;
    }
    for (size_t k = 0; k < n_slices; k++) if (evs[k]) cudaEventDestroy(evs[k]);
    if (e != cudaSuccess) return fail(PVGPU_E_CUDA, "copy of the frame to the host failed: %s", cudaGetErrorString(e));
    return PVGPU_OK;
}

static int ensure_frame(WorkCtx& c, size_t bytes)
{
    if (bytes > c.frame_cap) {
        cudaFree(c.d_frame); c.d_frame = nullptr; c.frame_cap = 0;
        CUDA_TRY(cudaMalloc(&c.d_frame, bytes));
        c.frame_cap = bytes;
    }
    return PVGPU_OK;
}

static int coop_flag(void* p) { return reinterpret_cast<std::atomic<int>*>(p)->load(std::memory_order_relaxed); }

// One frame on all devices / work contexts of the scene.  The rectangles are cut into units of consecutive rectangles (a unit is
// contiguous in the rect-major output); chunk j is made of the units j, j + C, j + 2C ... so that every chunk samples the whole
// frame (sky and geometry alike), and the workers - one host thread per work context - take chunk numbers from ONE atomic
// counter in page-locked host memory until none is left: whoever finishes early takes the next chunk, like the reference's
// render threads do with ViewData::GetNextRectangle (view.cpp:236-271).  Each worker renders its chunk into its own device frame
// and sends every unit to its place in the caller's frame: D2H for a host frame, cudaMemcpyPeerAsync into the first device's
// memory for a device frame.
static int render_multi(Scene& s, const pvgpu_aa* aa, int width, int height, const pvgpu_rect* rects, size_t n_rects,
                        float* host_out, float* dev_out, pvgpu_stats* stats, int (*cooperate)(void*), void* user)
{
    std::vector<std::pair<DeviceScene*, WorkCtx*>> workers;
    for (DeviceScene* d : s.devs) for (auto& c : d->ctx) workers.push_back({ d, c.get() });
    const size_t n_workers = workers.size();
    std::vector<size_t> off(n_rects + 1, 0);
    for (size_t i = 0; i < n_rects; i++) {
        const pvgpu_rect& r = rects[i];
        if (r.right < r.left || r.bottom < r.top) return fail(PVGPU_E_INVALID, "rectangle %zu is empty", i);
        off[i + 1] = off[i] + (size_t)(r.right - r.left + 1) * (size_t)(r.bottom - r.top + 1);
    }
    // chunk plan
    // chunks per worker: every chunk costs a full sequence of waves, each with a latency floor of a few hundred microseconds, so a
    // worker gets ONE chunk (an interleaved sample of the whole frame: statistically balanced) unless its share is large enough
    // for that floor not to matter (measured, 2 GPUs, config 2 at 1080p: 6.8 ms with one chunk per worker, 8.9 ms with two)
    size_t grabs = std::max<size_t>(1, std::min<size_t>(8, off[n_rects] / (n_workers * (size_t)(4u << 20))));
    if (const char* e = getenv("PVGPU_CHUNKS_PER_WORKER")) grabs = (size_t)std::max(1, std::min(64, atoi(e)));
    const size_t n_chunks = std::max<size_t>(1, std::min(n_rects, n_workers * grabs));
    const size_t unit = std::max<size_t>(1, std::min<size_t>(16, n_rects / (n_chunks * 8)));       // rectangles per unit
    const size_t n_units = (n_rects + unit - 1) / unit;
    unsigned int* counter = nullptr;                            // the work queue: next chunk to hand out
    CUDA_TRY(cudaHostAlloc(reinterpret_cast<void**>(&counter), sizeof(unsigned int), cudaHostAllocPortable));
    *counter = 0;
    std::atomic<int> abort_flag(0);
    std::mutex merge_mutex;
    pvgpu_stats total{};
    int first_rc = PVGPU_OK;
    std::string first_msg;
    std::vector<double> worker_ms(n_workers, 0.0);
    const int primary = s.devs[0]->device;

    auto work = [&](size_t w) {
        DeviceScene& d = *workers[w].first;
        WorkCtx& c = *workers[w].second;
        std::lock_guard<std::mutex> ctx_lock(c.in_use);
        cudaSetDevice(d.device);
        std::vector<pvgpu_rect> mine;
        std::vector<std::pair<size_t, size_t>> runs;          // (first rectangle, count) of every unit of the chunk
        for (;;) {
            if (abort_flag.load()) break;
            const unsigned int j = __atomic_fetch_add(counter, 1u, __ATOMIC_RELAXED);
            if (j >= n_chunks) break;
            mine.clear(); runs.clear();
            size_t px = 0;
            for (size_t u = j; u < n_units; u += n_chunks) {
                const size_t r0 = u * unit, r1 = std::min(n_rects, r0 + unit);
                runs.push_back({ r0, r1 - r0 });
                mine.insert(mine.end(), rects + r0, rects + r1);
                px += off[r1] - off[r0];
            }
            if (mine.empty()) continue;
            pvgpu_stats st{};
            int rc = ensure_frame(c, px * 4 * sizeof(float));
            if (rc == PVGPU_OK) rc = render_impl(s, d, c, aa, width, height, mine.data(), mine.size(), c.d_frame, &st, c.s_main, coop_flag, &abort_flag);
            // deliver the units
            size_t src_px = 0;
            for (size_t k = 0; k < runs.size() && rc == PVGPU_OK; k++) {
                const size_t r0 = runs[k].first, r1 = r0 + runs[k].second, n_px = off[r1] - off[r0];
                const float* src = c.d_frame + src_px * 4;
                if (dev_out) {
                    cudaError_t e = (d.device == primary) ? cudaMemcpyAsync(dev_out + off[r0] * 4, src, n_px * 16, cudaMemcpyDeviceToDevice, c.s_main)
                                                          : cudaMemcpyPeerAsync(dev_out + off[r0] * 4, primary, src, d.device, n_px * 16, c.s_main);
                    if (e != cudaSuccess) rc = fail(PVGPU_E_CUDA, "peer copy of finished tiles failed: %s", cudaGetErrorString(e));
                } else {
                    cudaError_t e = cudaMemcpyAsync(host_out + off[r0] * 4, src, n_px * 16, cudaMemcpyDeviceToHost, c.s_main);      // pinned or pageable (then staged by the driver)
                    if (e != cudaSuccess) rc = fail(PVGPU_E_CUDA, "copy of finished tiles to the host failed: %s", cudaGetErrorString(e));
                }
                src_px += n_px;
            }
            if (rc == PVGPU_OK && cudaStreamSynchronize(c.s_main) != cudaSuccess) rc = fail(PVGPU_E_CUDA, "delivery of finished tiles failed: %s", cudaGetErrorString(cudaGetLastError()));
            std::lock_guard<std::mutex> lock(merge_mutex);
            if (rc != PVGPU_OK) {
                if (first_rc == PVGPU_OK) { first_rc = rc; first_msg = pvgpu_last_error(); }
                abort_flag = 1;
                break;
            }
            merge_stats(total, st);
            worker_ms[w] += st.device_ms;
        }
    };
    std::vector<std::thread> pool;
    for (size_t w = 1; w < n_workers; w++) pool.emplace_back(work, w);
    if (cooperate) {
        // the caller's thread keeps polling its cooperate callback (Task::Cooperate is bound to that thread) and works as well
        std::thread w0(work, 0);
        std::atomic<int> done(0);
        std::thread waiter([&] { w0.join(); for (auto& t : pool) t.join(); done = 1; });
        while (!done.load()) {
            if (!abort_flag.load() && cooperate(user)) {
                std::lock_guard<std::mutex> lock(merge_mutex);
                if (first_rc == PVGPU_OK) { first_rc = PVGPU_E_ABORTED; first_msg = "render aborted by the cooperate callback"; }
                abort_flag = 1;
            }
            std::this_thread::sleep_for(std::chrono::milliseconds(2));
        }
        waiter.join();
    } else {
        work(0);
        for (auto& t : pool) t.join();
    }
    cudaFreeHost(counter);
    cudaSetDevice(primary);
    if (first_rc != PVGPU_OK) return fail(first_rc, "%s", first_msg.c_str());
    total.device_ms = *std::max_element(worker_ms.begin(), worker_ms.end());
    if (stats) *stats = total;
    return PVGPU_OK;
}

}  // namespace pvgpu

// device buffer holding a copy of a host array; freed on scope exit
template <class T> struct DevCopy {
    T* p = nullptr;
    cudaError_t e;
    DevCopy(const T* h, size_t n, bool upload) { e = cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)); if (e == cudaSuccess && upload && n) e = cudaMemcpy(p, h, n * sizeof(T), cudaMemcpyHostToDevice); }
    ~DevCopy() { cudaFree(p); }
};

using namespace pvgpu;

extern "C" {

static bool multi_worker(const Scene& s)
{
    return s.devs.size() > 1 || (s.devs.size() == 1 && s.devs[0]->ctx.size() > 1);
}

int pvgpu_render_device(pvgpu_scene* sc, const pvgpu_aa* aa, int width, int height,
                        const pvgpu_rect* rects, size_t n_rects, float* d_rgbt_out,
                        pvgpu_stats* stats, void* cuda_stream)
{
    clear_error();
    if (!sc) return fail(PVGPU_E_INVALID, "pvgpu_render_device: null scene");
    Scene& s = *reinterpret_cast<Scene*>(sc);
    if (!s.dev) return fail(PVGPU_E_INVALID, "scene not finalized");
    if (multi_worker(s)) {
        if (width <= 0 || height <= 0 || !rects || !n_rects || !d_rgbt_out) return fail(PVGPU_E_INVALID, "pvgpu_render_device: bad arguments");
        // the output belongs to the first device of the scene; work already queued on the caller's stream comes first
        CUDA_TRY(cudaSetDevice(s.devs[0]->device));
        CUDA_TRY(cudaStreamSynchronize(reinterpret_cast<cudaStream_t>(cuda_stream)));
        return render_multi(s, aa, width, height, rects, n_rects, nullptr, d_rgbt_out, stats, nullptr, nullptr);
    }
    WorkCtx& c = *s.dev->ctx[0];
    std::lock_guard<std::mutex> lock(c.in_use);
    return render_impl(s, *s.dev, c, aa, width, height, rects, n_rects, d_rgbt_out, stats,
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
    if (multi_worker(s)) {
        if (width <= 0 || height <= 0 || !n_rects) return fail(PVGPU_E_INVALID, "pvgpu_render: bad arguments");
        return render_multi(s, aa, width, height, rects, n_rects, rgbt_out, nullptr, stats, cooperate, user);
    }
    DeviceScene& d = *s.dev;
    WorkCtx& c = *d.ctx[0];
    std::lock_guard<std::mutex> frame_lock(c.in_use);      // the device frame + staging buffer serve one call at a time
    CUDA_TRY(cudaSetDevice(d.device));
    size_t n = 0;
    for (size_t i = 0; i < n_rects; i++)
        if (rects[i].right >= rects[i].left && rects[i].bottom >= rects[i].top)
            n += (size_t)(rects[i].right - rects[i].left + 1) * (size_t)(rects[i].bottom - rects[i].top + 1);
    // device frame + pinned staging buffer are kept between calls (a frame sequence reuses them)
    int rc = ensure_frame(c, std::max<size_t>(n, 1) * 4 * sizeof(float));
    if (rc != PVGPU_OK) return rc;
    rc = render_impl(s, d, c, aa, width, height, rects, n_rects, c.d_frame, stats, c.s_main, cooperate, user);
    if (rc != PVGPU_OK) return rc;
    return deliver_to_host(c, c.d_frame, rgbt_out, n * 4 * sizeof(float), c.s_main);
}

void* pvgpu_host_alloc(size_t bytes)
{
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}

void pvgpu_host_free(void* p)
{
    if (p) cudaFreeHost(p);
}

int pvgpu_trace_rays(pvgpu_scene* sc, const double* org_dir, size_t n, uint32_t* obj, double* depth, uint32_t* aux)
{
    clear_error();
    if (!sc || !org_dir || !obj || !depth) return fail(PVGPU_E_INVALID, "pvgpu_trace_rays: null argument");
    Scene& s = *reinterpret_cast<Scene*>(sc);
    if (!s.dev) return fail(PVGPU_E_INVALID, "scene not finalized");
    if (n == 0) return PVGPU_OK;
    if (n > 0xFFFFFFF0ull) return fail(PVGPU_E_INVALID, "too many rays");
    DeviceScene& d = *s.dev;
    WorkCtx& c = *d.ctx[0];
    std::lock_guard<std::mutex> device_lock(c.in_use);
    CUDA_TRY(cudaSetDevice(d.device));
    cudaStream_t st = c.s_main;
    const size_t chunk = 1u << 22;
    int rc = ensure_work_buffers(c, std::max<size_t>(std::min(n, chunk), 1024), 0, 0);
    if (rc != PVGPU_OK) return rc;
    double *d_rays = nullptr, *d_depth = nullptr;
    uint32_t *d_obj = nullptr, *d_aux = nullptr;
    auto cleanup = [&]() { cudaFree(d_rays); cudaFree(d_depth); cudaFree(d_obj); cudaFree(d_aux); };
    const size_t cn_max = std::min(n, chunk);
    if (cudaMalloc(&d_rays, cn_max * 6 * sizeof(double)) != cudaSuccess || cudaMalloc(&d_depth, cn_max * sizeof(double)) != cudaSuccess ||
        cudaMalloc(&d_obj, cn_max * sizeof(uint32_t)) != cudaSuccess || cudaMalloc(&d_aux, cn_max * sizeof(uint32_t)) != cudaSuccess) {
        cleanup();
        return fail(PVGPU_E_CUDA, "cudaMalloc failed in pvgpu_trace_rays");
    }
    cudaMemsetAsync(c.cnt, 0, sizeof(Counters), st);
    for (size_t c0 = 0; c0 < n && rc == PVGPU_OK; c0 += chunk) {
        const uint32_t cn = (uint32_t)std::min(chunk, n - c0);
        cudaMemcpyAsync(d_rays, org_dir + 6 * c0, (size_t)cn * 6 * sizeof(double), cudaMemcpyHostToDevice, st);
        launch_wave_init(c.ring, 2, cn, st);
        launch_probe_rays(d_rays, cn, c.q[0], st);
        launch_closest(d.view, c.q[0], c.ring, cn, (uint32_t)c.q_cap, c.hits, c.cnt, st);
        launch_probe_results(c.hits, cn, d_obj, d_depth, d_aux, st);
        c.kernel_launches += 4;
        cudaError_t e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { rc = fail(PVGPU_E_CUDA, "k_closest failed: %s", cudaGetErrorString(e)); break; }
        cudaMemcpy(obj + c0, d_obj, (size_t)cn * sizeof(uint32_t), cudaMemcpyDeviceToHost);
        cudaMemcpy(depth + c0, d_depth, (size_t)cn * sizeof(double), cudaMemcpyDeviceToHost);
        if (aux) cudaMemcpy(aux + c0, d_aux, (size_t)cn * sizeof(uint32_t), cudaMemcpyDeviceToHost);
    }
    if (rc == PVGPU_OK) {
        Counters hc;
        cudaMemcpy(&hc, c.cnt, sizeof hc, cudaMemcpyDeviceToHost);
        if (hc.overflow) rc = fail(PVGPU_E_OVERFLOW, "device capacity exceeded (flags 0x%x)", hc.overflow);
    }
    cleanup();
    return rc;
}

int pvgpu_solve_polynomial(pvgpu_scene* sc, size_t n, const int32_t* degree, const int32_t* sturm, const double* epsilon,
                           const double* coeffs, double* roots, int32_t* counts)
{
    clear_error();
    if (!sc || !degree || !sturm || !epsilon || !coeffs || !roots || !counts) return fail(PVGPU_E_INVALID, "pvgpu_solve_polynomial: null argument");
    Scene& s = *reinterpret_cast<Scene*>(sc);
    if (!s.dev) return fail(PVGPU_E_INVALID, "scene not finalized");
    if (n == 0) return PVGPU_OK;
    for (size_t i = 0; i < n; i++) if (degree[i] < 1 || degree[i] > 4) return fail(PVGPU_E_INVALID, "polynomial %zu: degree %d outside 1..4", i, degree[i]);
    std::lock_guard<std::mutex> lock(s.dev->ctx[0]->in_use);
    CUDA_TRY(cudaSetDevice(s.device));
    DevCopy<int32_t> d_deg(degree, n, true), d_st(sturm, n, true), d_cnt(nullptr, n, false);
    DevCopy<double> d_eps(epsilon, n, true), d_c(coeffs, 5 * n, true), d_r(nullptr, 4 * n, false);
    if (d_deg.e || d_st.e || d_cnt.e || d_eps.e || d_c.e || d_r.e) return fail(PVGPU_E_CUDA, "pvgpu_solve_polynomial: device allocation failed");
    launch_probe_solver((uint32_t)n, d_deg.p, d_st.p, d_eps.p, d_c.p, d_r.p, d_cnt.p, 0);
    s.dev->ctx[0]->kernel_launches++;
    CUDA_TRY(cudaMemcpy(roots, d_r.p, 4 * n * sizeof(double), cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(counts, d_cnt.p, n * sizeof(int32_t), cudaMemcpyDeviceToHost));
    return PVGPU_OK;
}

int pvgpu_noise(pvgpu_scene* sc, size_t n, const double* xyz, const int32_t* generator, const int32_t* octaves, double* out)
{
    clear_error();
    if (!sc || !xyz || !generator || !octaves || !out) return fail(PVGPU_E_INVALID, "pvgpu_noise: null argument");
    Scene& s = *reinterpret_cast<Scene*>(sc);
    if (!s.dev) return fail(PVGPU_E_INVALID, "scene not finalized");
    if (n == 0) return PVGPU_OK;
    std::lock_guard<std::mutex> lock(s.dev->ctx[0]->in_use);
    CUDA_TRY(cudaSetDevice(s.device));
    DevCopy<double> d_p(xyz, 3 * n, true), d_o(nullptr, 5 * n, false);
    DevCopy<int32_t> d_g(generator, n, true), d_oct(octaves, n, true);
    if (d_p.e || d_o.e || d_g.e || d_oct.e) return fail(PVGPU_E_CUDA, "pvgpu_noise: device allocation failed");
    launch_probe_noise(s.dev->view.noise, (uint32_t)n, d_p.p, d_g.p, d_oct.p, d_o.p, 0);
    s.dev->ctx[0]->kernel_launches++;
    CUDA_TRY(cudaMemcpy(out, d_o.p, 5 * n * sizeof(double), cudaMemcpyDeviceToHost));
    return PVGPU_OK;
}

void pvgpu_prewarm(int device)
{
    std::lock_guard<std::mutex> lock(g_prewarm_mutex);
    if (g_prewarm_thread.joinable()) return;
    g_prewarm_thread = std::thread([device] {
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) { cudaGetLastError(); return; }
        if (cudaSetDevice(device) == cudaSuccess) cudaFree(nullptr);      // creates the primary context
        cudaGetLastError();
    });
}

int pvgpu_fp64_peak(int device, double* tflops)
{
    clear_error();
    if (!tflops) return fail(PVGPU_E_INVALID, "pvgpu_fp64_peak: null argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return fail(PVGPU_E_NO_DEVICE, "no CUDA device available");
    if (device < 0 || device >= ndev) return fail(PVGPU_E_INVALID, "device %d out of range (have %d)", device, ndev);
    CUDA_TRY(cudaSetDevice(device));
    double* d_out = nullptr;
    CUDA_TRY(cudaMalloc(&d_out, sizeof(double)));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4096;
    double best = 0.0;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0, 0);
        const unsigned long long flop = launch_fp64_peak(d_out, iters, 0);
        cudaEventRecord(e1, 0);
        if (cudaEventSynchronize(e1) != cudaSuccess) break;
        float ms = 0.0f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms > 0.0f) best = std::max(best, (double)flop / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d_out);
    if (cudaGetLastError() != cudaSuccess || best == 0.0) return fail(PVGPU_E_CUDA, "pvgpu_fp64_peak: measurement kernel failed");
    *tflops = best;
    return PVGPU_OK;
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
    std::lock_guard<std::mutex> lock(d.ctx[0]->in_use);
    d.view.cam = s.camera;
    setup_camera(s, d.view);
    d.camera_dirty = true;
    double *d_xy = nullptr, *d_out = nullptr;
    if (cudaMalloc(&d_xy, n * 2 * sizeof(double)) != cudaSuccess || cudaMalloc(&d_out, n * 6 * sizeof(double)) != cudaSuccess) {
        cudaFree(d_xy); cudaFree(d_out);
        return fail(PVGPU_E_CUDA, "cudaMalloc failed in pvgpu_camera_rays");
    }
    cudaMemcpy(d_xy, xy, n * 2 * sizeof(double), cudaMemcpyHostToDevice);
    launch_camera_rays(d.view, d_xy, (uint32_t)n, (double)width, (double)height, d_out, 0);
    if (d.view.cam.reserved) launch_camera_normal_probe(d.view, d_xy, (uint32_t)n, (double)width, (double)height, d_out, 0);
    d.ctx[0]->kernel_launches++;
    cudaError_t e = cudaMemcpy(org_dir, d_out, n * 6 * sizeof(double), cudaMemcpyDeviceToHost);
    cudaFree(d_xy); cudaFree(d_out);
    if (e != cudaSuccess) return fail(PVGPU_E_CUDA, "k_camera_rays failed: %s", cudaGetErrorString(e));
    return PVGPU_OK;
}

}  // extern "C"
