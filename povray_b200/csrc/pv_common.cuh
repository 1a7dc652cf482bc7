// Device-side view of the flattened scene, the wavefront records that live in HBM, and the numeric
// constants of the reference's trace path.  sm_100a only; compiled with -fmad=false so that every FP64
// expression is evaluated with the same roundings as the -ffp-contract=off reference build.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "pvgpu.h"

namespace pvgpu {

// ---- constants (source/core/configcore.h) --------------------------------------------------------
#define PV_EPSILON          1.0e-10     // EPSILON          configcore.h:153
#define PV_HUGE_VAL         1.0e17      // HUGE_VAL         configcore.h:164
#define PV_BOUND_HUGE       2.0e10      // BOUND_HUGE       configcore.h:184
#define PV_SMALL_TOLERANCE  1.0e-3      // SMALL_TOLERANCE  configcore.h:193
#define PV_MAX_DISTANCE     1.0e7       // MAX_DISTANCE     configcore.h:198
#define PV_MIN_ISECT_DEPTH  1.0e-4      // MIN_ISECT_DEPTH  configcore.h:206
#define PV_SHADOW_TOLERANCE 1.0e-3      // SHADOW_TOLERANCE trace.cpp:81
#define PV_COORDINATE_LIMIT 1.0e17      // COORDINATE_LIMIT warp.h

// Resident CTAs per SM the traversal kernels are compiled for (register budget).  Measured per variant: the lean kernels are best
// at 8 (64 registers; 6 / 7 / 10: slower), the heavy ones at 16 (32 registers, full occupancy): their hot path is ~50 KB of
// instructions, two thirds of the warp samples wait on instruction fetch (profiles/r1_ncu_source_footprint_cfg3.txt), and more
// resident warps hide that better than registers help (config 3: 212 / 175 / 157 / 135 / 133 ms at 4 / 6 / 8 / 12 / 16).
#ifndef PV_TRAV_MIN_BLOCKS_LEAN
#define PV_TRAV_MIN_BLOCKS_LEAN 8
#endif
#ifndef PV_TRAV_MIN_BLOCKS_HEAVY
#define PV_TRAV_MIN_BLOCKS_HEAVY 12
#endif
#ifndef PV_TRAV_MIN_BLOCKS
#ifdef PV_LEAN
#define PV_TRAV_MIN_BLOCKS PV_TRAV_MIN_BLOCKS_LEAN
#else
#define PV_TRAV_MIN_BLOCKS PV_TRAV_MIN_BLOCKS_HEAVY
#endif
#endif
#define PV_STACK_SIZE     96            // traversal stack entries per ray (scene tree + nested mesh tree)
#define PV_MAX_LAYERS     8             // layers of a layered texture
#define PV_MAX_INTERIORS  10            // interiors a ray can be inside of at once
#define PV_CSG_STACK      16            // nesting depth of Inside() evaluation
#define PV_NO_OBJECT      0xFFFFFFFFu

// Every hot kernel is compiled twice: the full variant knows all primitives, the lean variant (-DPV_LEAN) only spheres,
// boxes, planes and meshes - scenes made of those (BASELINE configs 1 and 2) then run kernels that carry no quartic
// solver, CSG, blob or cone code (fewer registers spilled, smaller local frames).  The host picks per scene.
// The shading-side kernels (k_shade, k_shadow_filter) exist a third time (-DPV_FULL): the material features few scenes use -
// normal perturbation, pigment_map / average pigments, sky_sphere, fog, area lights - are compiled only there.  Their mere
// presence costs the other scenes dearly (config 3: 171 -> 305 ms per frame with them in the one heavy variant: larger local
// frames and code in the kernels that write the next wave's queues).
#ifdef PV_LEAN
#define PV_HEAVY 0
#define PV_FULL_MATERIALS 0
#define PV_VARIANT(name) name##_lean
#elif defined(PV_CSG)
// traversal kernels for scenes made of spheres, boxes, planes, quadrics, cones / cylinders and discs, alone or in CSG
// (BASELINE config 3): no polynomial solver, blob, mesh or polygon code
#define PV_HEAVY 1
#define PV_FULL_MATERIALS 0
#define PV_VARIANT(name) name##_csg
#define PV_TYPES ((1u << PVGPU_OBJ_SPHERE) | (1u << PVGPU_OBJ_BOX) | (1u << PVGPU_OBJ_PLANE) | (1u << PVGPU_OBJ_QUADRIC) | (1u << PVGPU_OBJ_CONE) | \
                  (1u << PVGPU_OBJ_DISC) | (1u << PVGPU_OBJ_CSG_UNION) | (1u << PVGPU_OBJ_CSG_INTERSECTION) | (1u << PVGPU_OBJ_CSG_MERGE))
#elif defined(PV_QUARTIC)
// traversal kernels for scenes made of stand-alone spheres, boxes, planes, quadrics, tori and blobs (BASELINE config 4): the
// polynomial solver stays, CSG, mesh, cone, disc, polygon, poly, glyph and prism code goes
#define PV_HEAVY 1
#define PV_FULL_MATERIALS 0
#define PV_VARIANT(name) name##_quartic
#define PV_TYPES ((1u << PVGPU_OBJ_SPHERE) | (1u << PVGPU_OBJ_BOX) | (1u << PVGPU_OBJ_PLANE) | (1u << PVGPU_OBJ_QUADRIC) | (1u << PVGPU_OBJ_TORUS) | (1u << PVGPU_OBJ_BLOB))
#elif defined(PV_FULL)
#define PV_HEAVY 1
#define PV_FULL_MATERIALS 1
#define PV_VARIANT(name) name##_full
#else
#define PV_HEAVY 1
#define PV_FULL_MATERIALS 0
#define PV_VARIANT(name) name
#endif

// Primitive kinds a kernel variant is compiled for (bit n = PVGPU_OBJ_* value n).  The heavy variants' hot path is bound by
// instruction fetch, so a variant that serves a class of scenes carries only that class's code: tests on the mask are constant
// folded and the code of the other primitives never reaches the kernel.
#ifndef PV_TYPES
#if PV_HEAVY
#define PV_TYPES 0x7FFFFu
#else
#define PV_TYPES ((1u << PVGPU_OBJ_SPHERE) | (1u << PVGPU_OBJ_BOX) | (1u << PVGPU_OBJ_PLANE) | (1u << PVGPU_OBJ_MESH))
#endif
#endif
#define PV_HAS(type) ((PV_TYPES & (1u << (type))) != 0u)
// Pattern set of a variant.  PV_BASIC_PATTERNS: plain and checker pigments with transform warps only - no noise, turbulence or
// waveform code reaches the shading-side kernels.
#ifndef PV_BASIC_PATTERNS
#define PV_BASIC_PATTERNS 0
#endif
#ifndef PV_CLIPBOUND
#define PV_CLIPBOUND PV_HEAVY          // clipped_by / bounded_by lists are served
#endif

// k_closest.cu also holds the kernels that exist once (camera rays, probes, queue bookkeeping): compiled in its default variant only
#if defined(PV_LEAN) || defined(PV_CSG) || defined(PV_QUARTIC)
#define PV_SECONDARY_TU 1
#else
#define PV_SECONDARY_TU 0
#endif

struct V3 { double x, y, z; };

// Per-ray traversal stack.  The first `nsh` entries of a thread live in shared memory (entry-major, one 8-byte
// column per thread: lanes never collide on a bank whatever their stack depths are), deeper entries in local memory.
#ifndef PV_TRAV_BLOCK
#define PV_TRAV_BLOCK   128            // threads per CTA of the traversal kernels
#endif
#ifndef PV_SSTACK
#define PV_SSTACK       0              // shared-memory entries per thread.  0: the whole stack stays in local memory - measured
                                       // faster on B200: 24 entries (24 KB per CTA, 8 CTAs per SM) shrink the L1 that serves the
                                       // node fetches and the register spills (cfg2 28.4 -> 30.0 ms)
#endif
struct TStack {
    uint2* sh;      // shared column of this thread (stride PV_TRAV_BLOCK), or nullptr
    uint2* lo;      // local-memory part
    int    nsh;     // entries held in shared memory (0 or PV_SSTACK)
    __device__ __forceinline__ uint2 get(int i) const { return (i < nsh) ? sh[i * PV_TRAV_BLOCK] : lo[i - nsh]; }
    __device__ __forceinline__ void set(int i, uint2 v) const { if (i < nsh) sh[i * PV_TRAV_BLOCK] = v; else lo[i - nsh] = v; }
};

// Packed mesh triangle for the intersection test: everything intersect_mesh_triangle
// (mesh.cpp:1040-1127) reads, in one 64-byte record (one L2 sector pair) instead of five gathers.
struct __align__(16) DTri {
    float p1[3], p2[3], p3[3];   // Vertices[P1..P3]
    float n[3];                  // Normals[Normal_Ind]
    float dist;                  // Distance
    uint32_t dom;                // Dominant_Axis
    uint32_t pad[2];
};
static_assert(sizeof(DTri) == 64, "DTri must be 64 bytes");

// Traversal copy of a BBOX_TREE node (derived at upload from pvgpu_node): 32 bytes = one sector.
//   hi   = lowerLeft + size rounded in FP32 - the very sum the reference forms at every slab test (boundingbox.cpp:576-591)
//   code = what a traversal stack entry needs: [31:28] number of children (0 = leaf), [27] BBOX_TREE::Infinite,
//          [26:0] first child (inner node) or object / triangle index (leaf).  Nodes with more than 14 children (only
//          the node collecting the infinite objects can have that many) are split at upload into groups that repeat
//          the parent's box, which changes no test result.
#define PV_CODE_INFINITE 0x08000000u
#define PV_CODE_INDEX    0x07FFFFFFu
struct __align__(16) DNode {
    float    lo[3], hi[3];
    uint32_t code;
    uint32_t aux;
};
static_assert(sizeof(DNode) == 32, "DNode must be 32 bytes");

struct DMesh {                   // pvgpu_mesh with absolute offsets resolved
    uint32_t tri_first, tri_count, node_first, node_count;    // node_first: root of the mesh tree in dmnodes
    uint32_t vertex_first, normal_first, texture_first, texture_count;
    uint32_t has_inside_vector;
    double inside_vector[3];
};

// Noise tables (built on the host with the reference's LCG, noise.cpp:231-255, 306-348).
struct NoiseTables {
    const uint16_t* hash;      // hashTable[8192]
    const double*   rtable;    // RTable[534]                            noise.cpp:98-145,181-182
    const uint16_t* perm;      // NoisePermutation[2*(NoiseEntries+1)]
    const double*   grad;      // NoiseGradients[2*(NoiseEntries+1)][3]
};

// Read-only scene tables in HBM (uploaded once by pvgpu_scene_finalize).
struct DScene {
    const pvgpu_object*      objs;
    const pvgpu_transform*   xf;
    const uint32_t*          index_list;
    const uint32_t*          frame;
    const pvgpu_node*        nodes;         // scene tree, root 0 (verbatim; container state only)
    const DNode*             dnodes;        // scene tree, traversal copy
    const DNode*             dmnodes;       // mesh trees, traversal copy (node_first of a mesh indexes this array)
    const DMesh*             meshes;
    const DTri*              dtris;
    const pvgpu_triangle*    tris;
    const float*             verts;
    const float*             norms;
    const pvgpu_light*       lights;
    const pvgpu_texture*     textures;
    const pvgpu_pigment*     pigments;
    const pvgpu_finish*      finishes;
    const pvgpu_blend_map*   maps;
    const pvgpu_blend_entry* entries;
    const pvgpu_warp*        warps;
    const pvgpu_interior*    interiors;
    const pvgpu_blob*        blobs;
    const pvgpu_blob_element* blob_elements;
    const pvgpu_blob_node*   blob_nodes;
    const pvgpu_image*       images;        // image_map pigments
    const float*             texels;        // r g b filter transmit per texel
    const int32_t*           blob_textures; // per blob element: texture or -1; nullptr: no blob has per-component textures
    const double*            mesh_uv;       // (u, v) pairs of the meshes' UV vectors
    const uint32_t*          tri_uv;        // per triangle: three indices into mesh_uv; nullptr: no mesh has UV vectors
    uint32_t                 has_uv;        // an object with PVGPU_UV_FLAG or a PVGPU_PAT_UV_MAP pigment exists: hits compute their (u, v)
    uint32_t                 pad_uv;
    const double*            shape_data;    // triangle / smooth_triangle / polygon parameters (pvgpu_object::mesh = offset)
    const pvgpu_tnormal*     tnormals;
    const pvgpu_slope_entry* slopes;
    const double*            wave_sources;  // TraceThreadData::waveSources (xyz per wave), Initialize_Waves (noise.cpp:189)
    const double*            wave_freqs;    // TraceThreadData::waveFrequencies
    const double*            pattern_rands; // gPatternRands (pattern.cpp:91): 32768 x mt19937 / 2^32; crackle and cells
    const pvgpu_fog*         fogs;          // SceneData::fog in list order
    uint32_t                 n_fogs, has_sky;
    float                    irid_wavelengths[3];   // SceneData::iridWavelengths
    uint32_t                 has_tnormals;      // some texture layer has a normal{} (per-layer normals are kept only then)
    uint32_t                 has_area_lights;   // some light is an area light and QualityFlags::areaLights is on
    uint32_t                 area_grid_max;     // largest Area_Size1 * Area_Size2
    pvgpu_sky_sphere         sky;           // SceneData::skysphere (has_sky)
    const uint32_t*          csg_leaves;    // per top-level CSG object: its primitive descendants (DFS order)
    const uint2*             csg_leaf_range;// per object: (first, count) into csg_leaves
    NoiseTables              noise;
    uint32_t n_objs, n_frame, n_nodes, n_lights;
    uint32_t n_mnodes;                      // nodes of all mesh trees (dmnodes)
    uint32_t use_tree;                      // boundingMethod == 1 && tree present
    uint32_t all_opaque;                    // every shadow caster has OPAQUE_FLAG
    uint32_t has_interiors;                 // the interior table is not empty
    pvgpu_globals g;
    pvgpu_camera  cam;
    // TracePixel::SetupCamera (tracepixel.cpp:235-309) for the non-pinhole cameras: normalised axes, aspectRatio, axis lengths, angles
    double cam_right[3], cam_up[3], cam_dir[3];
    double cam_aspect, cam_len_right, cam_len_up, cam_angle, cam_h_angle, cam_v_angle;
    uint16_t cam_interiors[PV_MAX_INTERIORS];   // TracePixel::InitRayContainerState result
    uint32_t n_cam_interiors;
};

// ---- wavefront records --------------------------------------------------------------------------
#define PV_RAY_PRIMARY     0x01u
#define PV_RAY_REFLECTION  0x02u
#define PV_RAY_REFRACTION  0x04u
#define PV_RAY_CONTINUED   0x08u    // TraceRay(..., continuedRay = true): trace level is not incremented
#define PV_RAY_ALPHA_BG    0x10u    // TraceTicket::alphaBackground
#define PV_RAY_PROBE       0x20u
#define PV_RAY_DEAD        0x40u    // CreateCameraRay returned false (fisheye / omnimax pixel outside the image circle): nothing is traced    // ray of the ray-level harness: Trace::FindIntersection without the camera's Max_Ray_Distance

// One pending TraceRay call (trace.cpp:135): 96 bytes.
struct __align__(16) PRay {
    double   o[3], d[3];
    float    w[3];               // linear RGB factor this ray's colour is multiplied with on its way to the pixel
    float    wt;                 // same for the transmittance (alpha) channel
    float    adc;                // TraceRay's `weight` argument (ADC bailout test)
    uint32_t sample;             // accumulation slot
    uint8_t  level;              // TraceTicket::traceLevel on entry
    uint8_t  flags;              // PV_RAY_*
    uint8_t  n_int;              // Ray::interiors.size()
    uint8_t  pad;
    uint16_t interiors[PV_MAX_INTERIORS];
};
static_assert(sizeof(PRay) == 96, "PRay must be 96 bytes");

// One pending TraceShadowRay call (trace.cpp:1892): 96 bytes.
struct __align__(16) SRay {
    double   o[3], d[3];
    double   depth;              // lightsourcedepth
    float    a[3];               // un-shadowed contribution (path weight x filter x light colour x BRDF terms)
    uint32_t sample;
    uint32_t parent;             // index of the PRay that spawned it (interior list for fade attenuation)
    uint32_t light;
    uint32_t pad[2];
};
static_assert(sizeof(SRay) == 96, "SRay must be 96 bytes");

struct Counters {
    unsigned long long rays, shadow_tests, reflected, refracted, transmitted, tir, adc_saves;
    unsigned long long shadow_rays;   // shadow rays traced (TraceShadowRay calls)
    unsigned long long node_tests[2]; // bounding-box slab tests (scene tree + mesh trees): [0] k_closest, [1] k_shadow_*
    unsigned long long prim_tests[2]; // top-level primitive tests + mesh triangle tests: [0] k_closest, [1] k_shadow_*
    unsigned int max_level;
    unsigned int overflow;
    unsigned int n_cont;              // continuation records taken in this batch (reflection exponent != 1)
    unsigned int pad;
};

// A reflected ray whose colour enters its parent NON-linearly: resultColour += reflec * Pow(rflCol, Reflect_Exp) (trace.cpp:1166-1168).
// Such a ray gets an accumulator slot of its own (behind the slots of the call), its whole subtree adds into that slot, and when
// the batch's waves are done the records are resolved from the last wave back to the first: the parent's slot receives
// w * pow(slot colour, exponent).  A continuation created in wave k has a parent from an earlier wave, so one pass per wave,
// latest first, sees every slot complete before it is read.
struct Cont {
    uint32_t parent;             // accumulator slot of the ray that was reflected
    uint32_t wave;               // wave whose k_shade created the record
    float    w[3];               // the parent's path weight x the layer's reflectivity
    float    exponent;           // FINISH::Reflect_Exp
};

// Per-wave hand-over between the kernels of one batch, in a device-resident ring indexed by the wave number.  The ray counts
// never travel through the host: k_shade of wave k appends into ring[k + 1].n_rays and ring[k].n_shadow, k_closest / k_shadow_*
// read their counts (clamped to the queue capacities) from the ring, and the warps of the traversal kernels take their 32-ray
// chunks from the cursors.  The host reads n_rays one wave behind the launches only to know when to stop (pvgpu_device.cu).
struct WaveCounts {
    unsigned int n_rays;          // rays of this wave
    unsigned int n_shadow;        // shadow rays this wave emitted
    unsigned int cur_closest;     // next unclaimed ray of k_closest
    unsigned int cur_shadow;      // next unclaimed shadow ray of k_shadow_*
    unsigned int cur_area;        // ... of k_shadow_area
    unsigned int counted;         // the wave's shadow rays have been added to Counters::shadow_rays (the queue is worked off by several launches)
    unsigned int pad[2];
};
static_assert(sizeof(WaveCounts) == 32, "WaveCounts must be 32 bytes");

// Queue records are written once and read once, megabytes apart in time: they go through the caches with the streaming hint
// (ld.global.cs / st.global.cs) so that they do not push the tree nodes and the per-ray traversal stacks out of L2.
template <class T> __device__ __forceinline__ T load_cs(const T* p)
{
    static_assert(sizeof(T) % 16 == 0, "16-byte records");
    T v;
    uint4* dst = reinterpret_cast<uint4*>(&v);
    const uint4* src = reinterpret_cast<const uint4*>(p);
    #pragma unroll
    for (int k = 0; k < (int)(sizeof(T) / 16); k++) dst[k] = __ldcs(src + k);
    return v;
}
template <class T> __device__ __forceinline__ void store_cs(T* p, const T& v)
{
    static_assert(sizeof(T) % 16 == 0, "16-byte records");
    const uint4* src = reinterpret_cast<const uint4*>(&v);
    uint4* dst = reinterpret_cast<uint4*>(p);
    #pragma unroll
    for (int k = 0; k < (int)(sizeof(T) / 16); k++) __stcs(dst + k, src[k]);
}

struct Hit {
    double   depth;
    V3       ip;                 // Intersection::IPoint
    uint32_t obj;                // Intersection::Object (primitive)
    uint32_t aux;                // Intersection::i1 (box side) or triangle index (Intersection::Pointer)
    int32_t  csg;                // Intersection::Csg or -1
};

// Closest-hit result of one PRay as it travels through HBM from k_closest to k_shade: 48 bytes.
#define PV_HIT_MISS     0xFFFFFFFFu      // no object hit: ComputeSky
#define PV_HIT_STOPPED  0xFFFFFFFEu      // max. trace level / ADC bailout stopped the ray (trace.cpp:147-155)
struct __align__(16) HitRec {
    double   depth;
    double   ip[3];
    uint32_t obj;
    uint32_t aux;
    int32_t  csg;
    uint32_t pad;
};
static_assert(sizeof(HitRec) == 48, "HitRec must be 48 bytes");

// sample i -> (rectangle, x, y) and its accumulator slot.  Slots are row-major inside a rectangle (the layout
// ViewData::CompletedRectangle takes), but the ORDER in which a rectangle's pixels become rays is by 8 x 4 pixel blocks
// when its size allows: the 32 rays of a warp then cover a compact patch of the image instead of a 32 x 1 strip and share
// more of their tree nodes (the shadow / reflection rays they spawn inherit the order).
__device__ __forceinline__ void sample_xy(const pvgpu_rect* rects, const uint32_t* rect_off, uint32_t n_rects, uint32_t i,
                                          double& x, double& y, uint32_t& slot)
{
    uint32_t lo = 0, hi = n_rects;
    while (hi - lo > 1) { uint32_t mid = (lo + hi) >> 1; if (rect_off[mid] <= i) lo = mid; else hi = mid; }
    const pvgpu_rect r = rects[lo];
    const uint32_t k = i - rect_off[lo], w = (uint32_t)(r.right - r.left + 1), h = (uint32_t)(r.bottom - r.top + 1);
    uint32_t px = k % w, py = k / w;
    if (((w & 15u) | (h & 7u)) == 0u) {
        // 8 x 4 blocks per warp, 2 x 2 such blocks (16 x 8 pixels) per CTA of four warps
        const uint32_t blk = k >> 5, in = k & 31u, grp = blk >> 2, gpr = w >> 4;
        px = (grp % gpr) * 16u + (blk & 1u) * 8u + (in & 7u);
        py = (grp / gpr) * 8u + ((blk >> 1) & 1u) * 4u + (in >> 3);
    } else if (((w & 7u) | (h & 3u)) == 0u) {
        const uint32_t blk = k >> 5, in = k & 31u, bpr = w >> 3;      // 8 x 4 blocks, row-major over the rectangle
        px = (blk % bpr) * 8u + (in & 7u);
        py = (blk / bpr) * 4u + (in >> 3);
    }
    slot = rect_off[lo] + py * w + px;
    x = (double)(r.left + (int)px) + 0.5;       // SimpleSamplingM0: pixel centres (tracetask.cpp:438)
    y = (double)(r.top + (int)py) + 0.5;
}

}  // namespace pvgpu
