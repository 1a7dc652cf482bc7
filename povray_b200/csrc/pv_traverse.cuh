// Closest-hit search: the flattened BBOX_TREE walk (Intersect_BBox_Tree + Check_And_Enqueue,
// boundingbox.cpp:485-648), the per-object Find_Intersection (object.cpp:172-224) with its FP32 box
// pre-test (object.cpp:917-941, 1074-1109), bounded_by / clipped_by (object.cpp:346-443), CSG
// candidate filtering (csg.cpp:128-375) and the nested mesh tree (mesh.cpp:1452-1528).
//
// The reference walks the tree best-first with a binary heap; here each ray keeps a small LIFO stack
// of (entry depth, node) pairs in local memory, pushes the children that pass the slab test nearest
// last, and prunes entries whose entry depth exceeds the best hit so far.  The set of leaves whose
// objects are tested against the final best depth is the same, so the closest hit is the same except
// for exact ties between different objects (first visited wins in both schemes, visit order differs).
#pragma once
#include "pv_shapes.cuh"
#if PV_HEAVY
#include "pv_blob.cuh"
#endif

namespace pvgpu {

// Rayinfo (boundingbox.h:182-216): FP32 origin and reciprocal direction, plus per-axis constants that let the
// slab test run without per-axis branches.
struct RayInfo {
    float org[3], inv[3];
    float cmin[3], cmax[3];       // 0 for axes the ray moves along; -/+BOUND_HUGE for axes with a zero direction component
    bool  nonzero[3], positive[3];
    bool  any_zero;               // some direction component is exactly zero (containment test needed)
    bool  special;                // any_zero, or a reciprocal overflowed to +-inf: such rays take the general slab test
};

__device__ __forceinline__ RayInfo make_rayinfo(const V3& o, const V3& d)
{
    RayInfo ri;
    const double dd[3] = { d.x, d.y, d.z };
    const double oo[3] = { o.x, o.y, o.z };
    ri.any_zero = false;
    #pragma unroll
    for (int k = 0; k < 3; k++) {
        ri.org[k] = (float)oo[k];
        ri.nonzero[k] = (dd[k] != 0.0);
        ri.inv[k] = ri.nonzero[k] ? (float)(1.0 / dd[k]) : 0.0f;
        ri.positive[k] = (dd[k] > 0.0);
        ri.cmin[k] = ri.nonzero[k] ? 0.0f : -2.0e10f;
        ri.cmax[k] = ri.nonzero[k] ? 0.0f : 2.0e10f;
        ri.any_zero = ri.any_zero || !ri.nonzero[k];
    }
    ri.special = ri.any_zero || isinf(ri.inv[0]) || isinf(ri.inv[1]) || isinf(ri.inv[2]);
    return ri;
}

// Check_And_Enqueue's slab test (boundingbox.cpp:566-636).  In the reference box data and ray info are FP32,
// the products (box - origin) * invdir are formed in FP32 and then compared as doubles against each other,
// against +-BOUND_HUGE (2e10, exactly representable in FP32) and against EPSILON.  Every operand of those
// comparisons is therefore an FP32 value, so the same decisions are taken here with FP32 compares; the one
// FP64 constant, EPSILON = 1e-10, is replaced by the smallest FP32 value that is >= 1e-10 (for FP32 t:
// (double)t < 1e-10  <=>  t < PV_EPSILON_F32).
#define PV_EPSILON_F32   1.0e-10f      // (double)1.0e-10f = 1.00000001335e-10 >= 1e-10; checked in device_upload()
#define PV_BOUND_HUGE_F  2.0e10f

// One traversal node (DNode, 32 bytes = one sector), fetched with two 128-bit loads.
struct NodeL {
    float lo[3], hi[3];
    uint32_t code;       // doubles as the stack entry: children count | infinite flag | first child or leaf payload
};
__device__ __forceinline__ NodeL unpack_node(const uint4 a, const uint4 b)
{
    NodeL n;
    n.lo[0] = __uint_as_float(a.x); n.lo[1] = __uint_as_float(a.y); n.lo[2] = __uint_as_float(a.z);
    n.hi[0] = __uint_as_float(a.w); n.hi[1] = __uint_as_float(b.x); n.hi[2] = __uint_as_float(b.y);
    n.code = b.z;
    return n;
}
__device__ __forceinline__ NodeL load_node(const DNode* p)
{
    const uint4* q = reinterpret_cast<const uint4*>(p);
    return unpack_node(__ldg(q), __ldg(q + 1));
}

// ---- upper tree levels in shared memory ------------------------------------------------------------
// The node arrays are laid out breadth first, so the first PV_TREELET_BYTES / 32 nodes of an array ARE the upper levels of the tree
// every ray walks through.  Each thread block of a traversal kernel copies that prefix of the scene's hottest tree (the mesh tree
// when the scene has meshes, else the scene tree) into shared memory with ONE bulk asynchronous copy of the TMA unit
// (cp.async.bulk.shared::cluster.global, completion on an mbarrier; UBLKCP in the SASS) before it takes its first ray; node
// fetches that fall into the prefix are then served from shared memory (LDS.128), the rest from global memory as before.
#ifndef PV_TREELET_BYTES
#define PV_TREELET_BYTES 0
#endif
#if PV_TREELET_BYTES > 0
extern __shared__ __align__(128) unsigned char pv_dyn_smem[];
struct TreeletHdr { const char* g_base; uint32_t bytes; uint32_t pad; unsigned long long mbar; };
#define PV_TREELET_HDR 128
__device__ __forceinline__ void treelet_stage(const DNode* g, uint32_t n_nodes)
{
    TreeletHdr* h = reinterpret_cast<TreeletHdr*>(pv_dyn_smem);
    const uint32_t bytes = min(n_nodes * (uint32_t)sizeof(DNode), (uint32_t)PV_TREELET_BYTES);
    const uint32_t mbar = (uint32_t)__cvta_generic_to_shared(&h->mbar);
    if (threadIdx.x == 0) {
        h->g_base = reinterpret_cast<const char*>(g);
        h->bytes = bytes;
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(mbar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (bytes) {
            const uint32_t dst = (uint32_t)__cvta_generic_to_shared(pv_dyn_smem + PV_TREELET_HDR);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(mbar), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         :: "r"(dst), "l"(g), "r"(bytes), "r"(mbar) : "memory");
        } else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(mbar) : "memory");
    }
    __syncthreads();                       // the barrier is initialised and armed
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(mbar) : "memory");
    } while (!done);
}
// the four children at p .. p + 3 (clamped duplicates included) lie inside the staged prefix
__device__ __forceinline__ bool treelet_holds(const DNode* p, uint32_t& off)
{
    const TreeletHdr* h = reinterpret_cast<const TreeletHdr*>(pv_dyn_smem);
    const size_t d = (size_t)(reinterpret_cast<const char*>(p) - h->g_base);
    off = (uint32_t)d;
    return d + 4u * sizeof(DNode) <= (size_t)h->bytes;
}
__device__ __forceinline__ NodeL load_node_sh(uint32_t off)
{
    const uint4* q = reinterpret_cast<const uint4*>(pv_dyn_smem + PV_TREELET_HDR + off);
    return unpack_node(q[0], q[1]);
}
#define PV_TREELET_SMEM (PV_TREELET_HDR + PV_TREELET_BYTES)
#define PV_TREELET_STAGE(sc) treelet_stage((sc).n_mnodes ? (sc).dmnodes : (sc).dnodes, (sc).n_mnodes ? (sc).n_mnodes : (sc).n_nodes)
#else
#define PV_TREELET_SMEM 0
#define PV_TREELET_STAGE(sc)
#endif

// Branch-free form: the reference's cascade of early exits is, as its own comment says (boundingbox.cpp:597-603),
// "if (tmax < dmax) dmax = tmax; if (tmin > dmin) dmin = tmin; if (dmin > dmax) return;" per axis.  dmin only grows and
// dmax only shrinks, so one final dmin > dmax test gives the same verdict; "tmax < EPSILON on some axis" is
// "dmax < EPSILON" at the end; fmaxf / fminf ignore a NaN product exactly like the reference's strict compares do.
// Axes with a zero direction component do not take part in the reference (containment test instead): their
// products are 0 * finite = 0 here and the per-ray constants cmin / cmax (-/+BOUND_HUGE) neutralise them, while for
// all other axes the constants are 0 and adding them changes nothing.
// Fast form for rays whose direction components are all non-zero with finite FP32 reciprocals (every ray in practice):
// no NaN can arise, (lo - org) * inv and (hi - org) * inv are ordered by the sign of inv (rounding is monotone), so
// tmin / tmax are the min / max of the pair - the same two products the reference assigns by the sign of the direction.
__device__ __forceinline__ bool slab_test_fast(const float* lo, const float* hi, const RayInfo& ri, float& dmin_out)
{
    float tn[3], tf[3];
    #pragma unroll
    for (int k = 0; k < 3; k++) {
        const float a = __fmul_rn(__fsub_rn(lo[k], ri.org[k]), ri.inv[k]);
        const float b = __fmul_rn(__fsub_rn(hi[k], ri.org[k]), ri.inv[k]);
        tn[k] = fminf(a, b);
        tf[k] = fmaxf(a, b);
    }
    const float dmin = fmaxf(fmaxf(fmaxf(tn[0], tn[1]), tn[2]), -PV_BOUND_HUGE_F);
    const float dmax = fminf(fminf(fminf(tf[0], tf[1]), tf[2]), PV_BOUND_HUGE_F);
    dmin_out = dmin;
    return !(dmax < PV_EPSILON_F32) && !(dmin > dmax);
}

__device__ __forceinline__ bool slab_test(const float* lo, const float* hi, const RayInfo& ri, float& dmin_out)
{
    float dmin = -PV_BOUND_HUGE_F, dmax = PV_BOUND_HUGE_F;
    #pragma unroll
    for (int k = 0; k < 3; k++) {
        const float near_p = ri.positive[k] ? lo[k] : hi[k];
        const float far_p = ri.positive[k] ? hi[k] : lo[k];
        const float tmin = __fadd_rn(__fmul_rn(__fsub_rn(near_p, ri.org[k]), ri.inv[k]), ri.cmin[k]);
        const float tmax = __fadd_rn(__fmul_rn(__fsub_rn(far_p, ri.org[k]), ri.inv[k]), ri.cmax[k]);
        dmin = fmaxf(dmin, tmin);
        dmax = fminf(dmax, tmax);
    }
    bool ok = !(dmax < PV_EPSILON_F32) && !(dmin > dmax);
    if (ri.any_zero) {
        #pragma unroll
        for (int k = 0; k < 3; k++)
            if (!ri.nonzero[k]) ok = ok && (lo[k] <= ri.org[k]) && (ri.org[k] <= hi[k]);
    }
    dmin_out = dmin;
    return ok;
}

// ObjectBase::Intersect_BBox -> Intersect_BBox_Dir (object.cpp:917-941, 1074-1109): all-FP32 test that
// gates All_Intersections for every primitive that does not override it (quadric, torus, mesh, CSG).
__device__ inline bool object_bbox_test(const float* bbox, const V3& o, const V3& d, float maxd)
{
    float org[3] = { (float)o.x, (float)o.y, (float)o.z };
    float inv[3] = { (float)(1.0 / d.x), (float)(1.0 / d.y), (float)(1.0 / d.z) };
    float b0[3] = { bbox[0], bbox[1], bbox[2] };
    float b1[3] = { __fadd_rn(bbox[0], bbox[3]), __fadd_rn(bbox[1], bbox[4]), __fadd_rn(bbox[2], bbox[5]) };
    const bool nx = inv[0] < 0.0f, ny = inv[1] < 0.0f, nz = inv[2] < 0.0f;
    float tmin  = __fmul_rn(__fsub_rn(nx ? b1[0] : b0[0], org[0]), inv[0]);
    float tmax  = __fmul_rn(__fsub_rn(nx ? b0[0] : b1[0], org[0]), inv[0]);
    float tymin = __fmul_rn(__fsub_rn(ny ? b1[1] : b0[1], org[1]), inv[1]);
    float tymax = __fmul_rn(__fsub_rn(ny ? b0[1] : b1[1], org[1]), inv[1]);
    if ((tmin > tymax) || (tymin > tmax)) return false;
    if (tymin > tmin) tmin = tymin;
    if (tymax < tmax) tmax = tymax;
    float tzmin = __fmul_rn(__fsub_rn(nz ? b1[2] : b0[2], org[2]), inv[2]);
    float tzmax = __fmul_rn(__fsub_rn(nz ? b0[2] : b1[2], org[2]), inv[2]);
    if ((tmin > tzmax) || (tzmin > tmax)) return false;
    if (tzmin > tmin) tmin = tzmin;
    if (tzmax < tmax) tmax = tzmax;
    return (tmin < maxd) && (tmax > (float)PV_MIN_ISECT_DEPTH);
}

__device__ __forceinline__ bool type_uses_bbox_test(uint32_t type)
{
    // Sphere, Box and Plane override Intersect_BBox to return true (sphere.cpp:753, box.cpp:1079, plane.cpp:629)
    // ... and so does Triangle (triangle.cpp:1419)
    // ... and Poly (polynomial.cpp:1510)
    return type >= PVGPU_OBJ_QUADRIC && type != PVGPU_OBJ_TRIANGLE && type != PVGPU_OBJ_POLY;
}

#define PV_MAX_DISTANCE_F 1.0e7f

// Tests the `count` children stored at nodes[first..] (Check_And_Enqueue per child, boundingbox.cpp:541-648) and
// pushes those the ray may hit so that the nearest ends up on top (equal depths: the later child on top, the order
// an insertion of one child after the other produces).  Children are fetched four at a time (the reference bunches
// <= 4 entries per node): 8 x LDG.128 in flight, four branch-free slab tests, ranks from the six pairwise comparisons, and
// predicated scattered pushes - no data-dependent loop, so the lanes of a warp stay converged through a node visit.
//   ORDERED = false (any-hit searches): no sorting, children are pushed in storage order.
//   limit: children entered beyond this depth are not pushed at all (a conservative FP32 upper bound of the best depth
//   so far; the exact test is repeated when an entry is popped, so this only saves stack traffic).
template <bool ALLOW_INFINITE, bool ORDERED>
__device__ __forceinline__ void push_children(const DNode* __restrict__ nodes, uint32_t first, uint32_t count, const RayInfo& ri,
                                              const TStack& stack, int& sp, unsigned int* overflow, float limit = 3.0e38f)
{
    const float kInvalid = __int_as_float(0x7fc00000);     // NaN: every comparison with it is false, never pushed
    for (uint32_t c0 = 0; c0 < count; c0 += 4) {
        NodeL ch[4];
#if PV_TREELET_BYTES > 0
        uint32_t sh_off;
        if (treelet_holds(nodes + first + c0, sh_off)) {
            #pragma unroll
            for (int k = 0; k < 4; k++) ch[k] = load_node_sh(sh_off + (uint32_t)sizeof(DNode) * min((uint32_t)k, count - 1u - c0));
        } else
#endif
        {
            #pragma unroll
            for (int k = 0; k < 4; k++) ch[k] = load_node(nodes + first + min(c0 + (uint32_t)k, count - 1u));
        }
        float key[4];
        if (!ri.special) {
            #pragma unroll
            for (int k = 0; k < 4; k++) {
                float dmin;
                bool ok = slab_test_fast(ch[k].lo, ch[k].hi, ri, dmin);
                if (ALLOW_INFINITE && (ch[k].code & PV_CODE_INFINITE)) { dmin = -PV_MAX_DISTANCE_F; ok = true; }   // boundingbox.cpp:643-647
                ok = ok && (c0 + k < count) && !(dmin > limit);
                key[k] = ok ? dmin : kInvalid;
            }
        } else {
            #pragma unroll
            for (int k = 0; k < 4; k++) {
                float dmin;
                bool ok = slab_test(ch[k].lo, ch[k].hi, ri, dmin);
                if (ALLOW_INFINITE && (ch[k].code & PV_CODE_INFINITE)) { dmin = -PV_MAX_DISTANCE_F; ok = true; }
                ok = ok && (c0 + k < count) && !(dmin > limit);
                key[k] = ok ? dmin : kInvalid;
            }
        }
        // Where each child lands above the current top: the farthest at the bottom, the nearest on top, equal depths in child
        // order (the later child on top) - the order a descending stable sort of the entry depths produces, computed as ranks from
        // the six pairwise comparisons instead of a sorting network.  ORDERED = false: storage order.
        int pos[4] = { 0, 0, 0, 0 };
        int n_valid = 0;
        if (ORDERED) {
            #pragma unroll
            for (int k = 1; k < 4; k++) {
                #pragma unroll
                for (int j = 0; j < k; j++) {
                    pos[k] += (key[j] >= key[k]) ? 1 : 0;      // j goes below k: farther, or as far and earlier
                    pos[j] += (key[j] < key[k]) ? 1 : 0;       // k goes below j
                }
            }
            #pragma unroll
            for (int k = 0; k < 4; k++) n_valid += (key[k] == key[k]) ? 1 : 0;
        } else {
            #pragma unroll
            for (int k = 0; k < 4; k++) { pos[k] = n_valid; n_valid += (key[k] == key[k]) ? 1 : 0; }
        }
#ifdef PV_PREFETCH_CHILDREN
        // the children of an inner node are one 128-byte line (4 x 32 B): ask L2 for the line of every pushed inner child now - the
        // nearest is visited next anyway, the others find theirs waiting when they are popped (a dependent DRAM fetch saved each)
        #pragma unroll
        for (int k = 0; k < 4; k++)
            if (key[k] == key[k] && (ch[k].code >> 28) != 0u) asm volatile("prefetch.global.L2 [%0];" :: "l"(nodes + (ch[k].code & PV_CODE_INDEX)));
#endif
        const int room = PV_STACK_SIZE - 1 - sp;               // the last slot is never used: an overflow is flagged instead
        #pragma unroll
        for (int k = 0; k < 4; k++)
            if (key[k] == key[k] && pos[k] < room) stack.set(sp + pos[k], make_uint2(__float_as_uint(key[k]), ch[k].code));
        if (n_valid > room) { atomicOr(overflow, 1u); n_valid = room > 0 ? room : 0; }
        sp += n_valid;
    }
}

// ---- Inside --------------------------------------------------------------------------------------
__device__ bool mesh_inside(const DScene& sc, const pvgpu_object& ob, const V3& p, TStack stack, int sp0);

#if PV_HEAVY
static __device__ __noinline__ bool prim_inside(const DScene& sc, const pvgpu_object& ob, const V3& p, TStack stack, int sp0)
#else
__device__ inline bool prim_inside(const DScene& sc, const pvgpu_object& ob, const V3& p, TStack stack, int sp0)
#endif
{
    switch (ob.type) {
        case PVGPU_OBJ_SPHERE: if (PV_HAS(PVGPU_OBJ_SPHERE)) return sphere_inside(sc, ob, p); break;
        case PVGPU_OBJ_BOX: if (PV_HAS(PVGPU_OBJ_BOX)) return box_inside(sc, ob, p); break;
        case PVGPU_OBJ_PLANE: if (PV_HAS(PVGPU_OBJ_PLANE)) return plane_inside(sc, ob, p); break;
        case PVGPU_OBJ_MESH: if (PV_HAS(PVGPU_OBJ_MESH)) return mesh_inside(sc, ob, p, stack, sp0); break;
#if PV_HEAVY
        case PVGPU_OBJ_QUADRIC: if (PV_HAS(PVGPU_OBJ_QUADRIC)) return quadric_inside(ob, p); break;
        case PVGPU_OBJ_TORUS: if (PV_HAS(PVGPU_OBJ_TORUS)) return torus_inside(sc, ob, p); break;
        case PVGPU_OBJ_BLOB: if (PV_HAS(PVGPU_OBJ_BLOB)) return blob_inside(sc, ob, p); break;
        case PVGPU_OBJ_CONE: if (PV_HAS(PVGPU_OBJ_CONE)) return cone_inside(sc, ob, p); break;
        case PVGPU_OBJ_DISC: if (PV_HAS(PVGPU_OBJ_DISC)) return disc_inside(sc, ob, p); break;
        case PVGPU_OBJ_POLY: if (PV_HAS(PVGPU_OBJ_POLY)) return poly_inside(sc, ob, p); break;
        case PVGPU_OBJ_GLYPH: if (PV_HAS(PVGPU_OBJ_GLYPH)) return glyph_inside(sc, ob, p); break;
        case PVGPU_OBJ_PRISM: if (PV_HAS(PVGPU_OBJ_PRISM)) return prism_inside(sc, ob, p); break;
        case PVGPU_OBJ_SUPERELLIPSOID: if (PV_HAS(PVGPU_OBJ_SUPERELLIPSOID)) return superellipsoid_inside(sc, ob, p); break;
#endif
    }
    return false;
}

// Inside() of the primitives that need no traversal stack (sphere.cpp:262, box.cpp:538, plane.cpp:208, quadric.cpp:248, torus.cpp:348)
__device__ __forceinline__ bool simple_inside(const DScene& sc, const pvgpu_object& ob, const V3& p)
{
    switch (ob.type) {
        case PVGPU_OBJ_SPHERE: if (PV_HAS(PVGPU_OBJ_SPHERE)) return sphere_inside(sc, ob, p); break;
        case PVGPU_OBJ_BOX: if (PV_HAS(PVGPU_OBJ_BOX)) return box_inside(sc, ob, p); break;
        case PVGPU_OBJ_PLANE: if (PV_HAS(PVGPU_OBJ_PLANE)) return plane_inside(sc, ob, p); break;
#if PV_HEAVY
        case PVGPU_OBJ_QUADRIC: if (PV_HAS(PVGPU_OBJ_QUADRIC)) return quadric_inside(ob, p); break;
        case PVGPU_OBJ_TORUS: if (PV_HAS(PVGPU_OBJ_TORUS)) return torus_inside(sc, ob, p); break;
        case PVGPU_OBJ_CONE: if (PV_HAS(PVGPU_OBJ_CONE)) return cone_inside(sc, ob, p); break;
        case PVGPU_OBJ_DISC: if (PV_HAS(PVGPU_OBJ_DISC)) return disc_inside(sc, ob, p); break;
        case PVGPU_OBJ_POLY: if (PV_HAS(PVGPU_OBJ_POLY)) return poly_inside(sc, ob, p); break;
#endif
    }
    return false;
}

// Inside_Object (object.cpp:346-355) = every clipped_by object contains the point AND Object->Inside();
// CSGUnion/CSGMerge::Inside = any child, CSGIntersection::Inside = all children (csg.cpp:393-454).
// Evaluated iteratively with short-circuit over the object graph (the reference recurses).
static __device__ __noinline__ bool inside_object(const DScene& sc, uint32_t root, const V3& p, TStack stack, int sp0, bool root_clip = true)
{
    struct Frame { uint32_t obj; uint32_t cur; };
    Frame st[PV_CSG_STACK];
    int sp = 0;
    st[0].obj = root; st[0].cur = 0;
    bool ret = false, have_ret = false;
    while (sp >= 0) {
        Frame& f = st[sp];
        const pvgpu_object& o = sc.objs[f.obj];
        const uint32_t nclip = (sp == 0 && !root_clip) ? 0u : o.clip_count;   // Object->Inside() alone for the root if asked
        const bool is_csg = PVGPU_IS_CSG(o.type);
        if (have_ret) {
            have_ret = false;
            const uint32_t e = f.cur - 1;
            if (e < nclip) { if (!ret) { have_ret = true; sp--; continue; } }           // a clip object excludes the point
            else if (o.type == PVGPU_OBJ_CSG_INTERSECTION) { if (!ret) { have_ret = true; sp--; continue; } }
            else if (ret) { have_ret = true; sp--; continue; }                            // union / merge: one child suffices
        }
        uint32_t next;
        if (f.cur < nclip) next = sc.index_list[o.clip_first + f.cur];
        else if (!is_csg) { ret = prim_inside(sc, o, p, stack, sp0); have_ret = true; sp--; continue; }
        else if (f.cur - nclip < o.child_count) next = sc.index_list[o.child_first + (f.cur - nclip)];
        else { ret = (o.type == PVGPU_OBJ_CSG_INTERSECTION); have_ret = true; sp--; continue; }
        f.cur++;
        if (sp + 1 >= PV_CSG_STACK) { ret = false; have_ret = true; sp--; continue; }     // deeper than supported (validated on the host)
        sp++;
        st[sp].obj = next; st[sp].cur = 0;
    }
    return ret;
}

// Point_In_Clip (object.cpp:430-443)
__device__ inline bool point_in_clip(const DScene& sc, const pvgpu_object& o, const V3& p, TStack stack, int sp0)
{
#if PV_CLIPBOUND
    for (uint32_t i = 0; i < o.clip_count; i++)
        if (!inside_object(sc, sc.index_list[o.clip_first + i], p, stack, sp0)) return false;
#endif
    return true;       // (the lean variant only serves scenes without clipped_by / bounded_by lists: device_upload)
}

// ---- per-object candidate collection ----------------------------------------------------------------
// Selection rule of Find_Intersection (object.cpp:203-215): the IStack is popped from the top with a
// strict `<`, so among equal depths of ONE object the hit pushed LAST wins -> `<=` in push order.
// Work counters of the hot traversal (per lane, summed into Counters::node_tests / prim_tests by the kernels): bounding-box slab tests
// (children of visited nodes, scene tree and mesh trees) and primitive tests (top-level All_Intersections calls and mesh triangles).
// bench.py derives the algorithmic bytes of the roofline from them.
struct TravCount { uint32_t nodes, prims; };

struct HitAcc {
    double closest;      // starts at HUGE_VAL per object
    double post_min;     // SmallToleranceRayObjectCondition: depth > post_min (shadow rays), else -1
    Hit    best;
    bool   found;
};

__device__ __forceinline__ void consider(HitAcc& acc, double depth, const V3& ip, uint32_t obj, uint32_t aux, int32_t csg)
{
    if (depth <= acc.closest && depth >= PV_MIN_ISECT_DEPTH && depth > acc.post_min) {
        acc.closest = depth;
        acc.best.depth = depth; acc.best.ip = ip; acc.best.obj = obj; acc.best.aux = aux; acc.best.csg = csg;
        acc.found = true;
    }
}

// (one out-of-line copy in the full variant: the switch carries every primitive incl. the quartic solver, and it is reached from
//  object_find and from both CSG paths - inlined copies made the heavy kernels instruction-fetch bound)
#if PV_HEAVY
static __device__ __noinline__ void prim_hits(const DScene& sc, const pvgpu_object& ob, const V3& o, const V3& d, PrimHits& h, unsigned int* overflow = nullptr, int* resume = nullptr)
#else
__device__ inline void prim_hits(const DScene& sc, const pvgpu_object& ob, const V3& o, const V3& d, PrimHits& h, unsigned int* overflow = nullptr, int* resume = nullptr)
#endif
{
    h.n = 0;
    switch (ob.type) {
#if PV_HEAVY
        case PVGPU_OBJ_BLOB: if (PV_HAS(PVGPU_OBJ_BLOB)) { if (!blob_hits(sc, ob, o, d, h, resume) && overflow) atomicOr(overflow, 32u); } break;
        case PVGPU_OBJ_QUADRIC: if (PV_HAS(PVGPU_OBJ_QUADRIC)) { quadric_hits(ob, o, d, h); } break;
        case PVGPU_OBJ_TORUS: if (PV_HAS(PVGPU_OBJ_TORUS)) { torus_hits(sc, ob, o, d, h); } break;
        case PVGPU_OBJ_CONE: if (PV_HAS(PVGPU_OBJ_CONE)) { cone_hits(sc, ob, o, d, h); } break;
        case PVGPU_OBJ_DISC: if (PV_HAS(PVGPU_OBJ_DISC)) { disc_hits(sc, ob, o, d, h); } break;
        case PVGPU_OBJ_TRIANGLE: if (PV_HAS(PVGPU_OBJ_TRIANGLE)) { triangle_hits(sc, ob, o, d, h); } break;
        case PVGPU_OBJ_POLYGON: if (PV_HAS(PVGPU_OBJ_POLYGON)) { polygon_hits(sc, ob, o, d, h); } break;
        case PVGPU_OBJ_POLY: if (PV_HAS(PVGPU_OBJ_POLY)) { poly_hits(sc, ob, o, d, h); } break;
        case PVGPU_OBJ_GLYPH: if (PV_HAS(PVGPU_OBJ_GLYPH)) { glyph_hits(sc, ob, o, d, h, resume); if (resume == nullptr && overflow) atomicOr(overflow, 128u); } break;
        case PVGPU_OBJ_PRISM: if (PV_HAS(PVGPU_OBJ_PRISM)) { prism_hits(sc, ob, o, d, h, resume); if (resume == nullptr && overflow) atomicOr(overflow, 128u); } break;
        case PVGPU_OBJ_SUPERELLIPSOID: if (PV_HAS(PVGPU_OBJ_SUPERELLIPSOID)) { superellipsoid_hits(sc, ob, o, d, h, resume); if (resume == nullptr && overflow) atomicOr(overflow, 128u); } break;
#endif
        case PVGPU_OBJ_SPHERE: if (PV_HAS(PVGPU_OBJ_SPHERE)) { sphere_hits(sc, ob, o, d, h); } break;
        case PVGPU_OBJ_BOX: if (PV_HAS(PVGPU_OBJ_BOX)) { box_hits(sc, ob, o, d, h); } break;
        case PVGPU_OBJ_PLANE: if (PV_HAS(PVGPU_OBJ_PLANE)) { plane_hits(sc, ob, o, d, h); } break;
        default: h.n = 0; break;
    }
}

// Test_Ray_Flags / Test_Ray_Flags_Shadow (csg.cpp:80-104) for the ray kinds this path traces.
__device__ __forceinline__ bool test_ray_flags(uint32_t oflags, uint32_t rflags, bool shadow_ray, bool shadow_variant)
{
    const bool primary = (rflags & PV_RAY_PRIMARY) != 0;
    const bool reflection = (rflags & PV_RAY_REFLECTION) != 0;
    const bool image = !shadow_ray && (primary || ((rflags & PV_RAY_REFRACTION) && !reflection));
    bool ok = (!(oflags & PVGPU_NO_IMAGE_FLAG) || !image || (!shadow_variant && primary)) &&
              (!(oflags & PVGPU_NO_REFLECTION_FLAG) || !reflection);
    if (shadow_variant && shadow_ray && !(oflags & PVGPU_NO_SHADOW_FLAG)) ok = true;
    return ok;
}

// Mesh::intersect_bbox_tree + test_hit (mesh.cpp:1452-1528, 1208-1243) with a LIFO stack.
//   ANY_HIT (shadow rays, every caster opaque): return at the first accepted hit inside (SHADOW_TOLERANCE, any_limit) -
//   the closest hit the reference would find is then inside the same window, i.e. the light is blocked either way.
//   accept(depth, point, triangle, stack top): takes a triangle hit - by default Point_In_Clip + the selection rule of consider(); a mesh
//   that is a CSG child passes the CSG filter instead (csg_hits).  Entries are pruned against acc.closest, which only accepted hits move.
template <bool ANY_HIT, class Accept>
__device__ inline void mesh_hits_f(const DScene& sc, uint32_t obj_index, const pvgpu_object& ob, const V3& o, const V3& d,
                                   HitAcc& acc, TStack stack, int sp0, unsigned int* overflow, double any_limit, Accept accept)
{
    const DMesh& me = sc.meshes[ob.mesh];
    V3 mo = o, md = d;
    double len = 1.0;
    if (ob.transform >= 0) {
        const pvgpu_transform& t = sc.xf[ob.transform];
        mo = inv_trans_point(t, o);
        md = inv_trans_direction(t, d);
        len = length(md);
        md = md / len;
    }
    if (me.node_count == 0) {
        for (uint32_t i = 0; i < me.tri_count; i++) {
            double t;
            if (tri_intersect(sc.dtris[me.tri_first + i], mo, md, t)) {
                double wd = t / len;
                V3 ip = evaluate(o, d, wd);
                accept(wd, ip, me.tri_first + i, sp0);
            }
        }
        return;
    }
    const RayInfo ri = make_rayinfo(mo, md);
    const DNode* __restrict__ nodes = sc.dmnodes + me.node_first;
    int sp = sp0;
    {
        const NodeL root = load_node(nodes);
        float dmin;
        if (!slab_test(root.lo, root.hi, ri, dmin)) return;
        stack.set(sp++, make_uint2(__float_as_uint(dmin), root.code));
    }
    // "while-while" traversal: every lane first walks inner nodes until it holds a leaf (or runs out of work); the
    // warp reconverges behind that loop and the lanes that found a triangle run the FP64 test together, instead of
    // paying for both code paths on every iteration.
    // Entries further than the closest accepted hit cannot improve it; compared in mesh space (t = world * len).
    const bool any_ok = ANY_HIT && (ob.flags & PVGPU_OPAQUE_FLAG) && ob.clip_count == 0;
    for (;;) {
        uint32_t leaf = 0xFFFFFFFFu;
        while (sp > sp0) {
            const uint2 e = stack.get(--sp);
            if ((double)__uint_as_float(e.x) > acc.closest * len) continue;
            const uint32_t code = e.y >> 28, idx = e.y & PV_CODE_INDEX;
            if (code == 0u) { leaf = idx; break; }
            const uint32_t first = idx, count = code;
            push_children<false, true>(nodes, first, count, ri, stack, sp, overflow);
        }
        if (leaf == 0xFFFFFFFFu) break;
        double t;
        const uint32_t ti = me.tri_first + leaf;
        if (tri_intersect(sc.dtris[ti], mo, md, t)) {
            double wd = t / len;
            V3 ip = evaluate(o, d, wd);
            accept(wd, ip, ti, sp);
            if (any_ok && acc.found && acc.closest > PV_SHADOW_TOLERANCE && acc.closest < any_limit) return;
        }
    }
}

template <bool ANY_HIT>
__device__ inline void mesh_hits(const DScene& sc, uint32_t obj_index, const pvgpu_object& ob, const V3& o, const V3& d,
                                 HitAcc& acc, int32_t csg, TStack stack, int sp0, unsigned int* overflow, double any_limit = 0.0)
{
    mesh_hits_f<ANY_HIT>(sc, obj_index, ob, o, d, acc, stack, sp0, overflow, any_limit, [&](double wd, const V3& ip, uint32_t ti, int sp_now) {
        if (ob.clip_count == 0 || point_in_clip(sc, ob, ip, stack, sp_now)) consider(acc, wd, ip, obj_index, ti, csg);
    });
}

// Warp-synchronous form of mesh_hits for the hot path.  ALL 32 lanes of the warp call it together; lanes with
// `active` hold a mesh object, the others idle.  The loops have warp-uniform conditions (votes), so the warp
// provably reconverges between the two phases: every lane walks inner nodes until it holds a triangle (or runs out
// of work), then the lanes that found one run the FP64 triangle test together.
#define PV_FULL_MASK 0xffffffffu
// phase vote: leaves are served when (#lanes holding a leaf) * NUM > (#lanes wanting a node visit) * DEN
#ifndef PV_LEAF_BIAS_NUM
#define PV_LEAF_BIAS_NUM 2
#define PV_LEAF_BIAS_DEN 1
#endif
#define PV_NONE      0xFFFFFFFFu
// Scope of the phase votes.  Warp scope: the 32 lanes of a warp agree on the phase.  CTA scope (-DPV_CTA_SYNC): all warps of the
// thread block agree, so that the whole block runs the same few KB of code at a time - the heavy kernels are bound by
// instruction fetch (their hot path is several times the 32 KB L1.5 instruction cache when every warp of an SM sits in a
// different phase).  Every loop that contains a vote must then be uniform over the block.
#ifdef PV_CTA_SYNC
// NOT enabled in the shipped build (Makefile, CSG_FLAGS / QUARTIC_FLAGS): with it config 3 / 4 ran 5 % / 16 % faster, parity-green and
// memcheck-clean, but compute-sanitizer's synccheck reports "Divergent thread(s) in block" at these barriers on small frames although
// every loop around a vote is block-uniform at source level (ptxas peels the first turn of the vote loops, so the barriers exist twice
// in SASS).  The __syncwarp() in front of the warp-aligned barriers did not clear the report.  Open item - see profiles/README.md.
#ifdef PV_CTA_SYNC_UNALIGNED
// experiment: the barrier forms without .aligned, which PTX allows the lanes of a warp to reach one by one
__device__ __forceinline__ int vote_count(bool p)
{
    int r;
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %1, 0;\n\tbarrier.cta.red.popc.u32 %0, 0, q;\n\t}" : "=r"(r) : "r"(p ? 1u : 0u) : "memory");
    return r;
}
__device__ __forceinline__ bool vote_any(bool p)
{
    unsigned r;
    asm volatile("{\n\t.reg .pred q, o;\n\tsetp.ne.u32 q, %1, 0;\n\tbarrier.cta.red.or.pred o, 0, q;\n\tselp.u32 %0, 1, 0, o;\n\t}" : "=r"(r) : "r"(p ? 1u : 0u) : "memory");
    return r != 0u;
}
__device__ __forceinline__ void cta_barrier() { asm volatile("barrier.cta.sync 0;" ::: "memory"); }
#else
__device__ __forceinline__ int  vote_count(bool p) { __syncwarp(); return __syncthreads_count(p ? 1 : 0); }
__device__ __forceinline__ bool vote_any(bool p)   { __syncwarp(); return __syncthreads_or(p ? 1 : 0) != 0; }
__device__ __forceinline__ void cta_barrier() { __syncwarp(); __syncthreads(); }
#endif
#else
__device__ __forceinline__ int  vote_count(bool p) { return __popc(__ballot_sync(PV_FULL_MASK, p)); }
__device__ __forceinline__ bool vote_any(bool p)   { return __any_sync(PV_FULL_MASK, p); }
#endif
// The traversal kernels take their rays in chunks from a cursor in the wave's WaveCounts record: one chunk per warp or, with
// block-wide phase votes, one per thread block.  Rays of a chunk are neighbours in the queue (an 8 x 4 pixel block of the frame and
// what it spawned), whichever warp picks the chunk up; a warp whose rays end early simply fetches the next chunk instead of idling
// until the slowest warp of a static partition is done.
// A warp normally takes 32 rays.  When a wave is too small to give every warp of the grid a chunk (late waves, one GPU's share of
// a frame sharded over eight), chunks shrink to 16 or 8 rays and the other lanes idle: the time of such a wave is the time of its
// slowest warp, which is shorter the fewer (diverging) rays the warp has to serve turn by turn.
#ifndef PV_CHUNK_MIN
#define PV_CHUNK_MIN 8u
#endif
__device__ __forceinline__ uint32_t chunk_size(uint32_t n)
{
#ifdef PV_CTA_SYNC
    return blockDim.x;
#else
    const uint32_t warps = gridDim.x * (blockDim.x >> 5);
    uint32_t cs = 32u;
    while (cs > PV_CHUNK_MIN && (n + cs - 1u) / cs < warps) cs >>= 1;
    return cs;
#endif
}

// Next chunk: `i` = this lane's ray, or 0xFFFFFFFF for a lane without one.  Returns false when the wave is exhausted.
__device__ __forceinline__ bool next_chunk(unsigned int* cursor, uint32_t n, uint32_t cs, uint32_t& i)
{
#ifdef PV_CTA_SYNC
    __shared__ uint32_t s_base;
    cta_barrier();                                     // everybody has consumed the previous value
    if (threadIdx.x == 0) s_base = atomicAdd(cursor, cs);
    cta_barrier();
    const uint32_t base = s_base;
    i = (base + threadIdx.x < n) ? base + threadIdx.x : 0xFFFFFFFFu;
    return base < n;
#else
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t base = 0;
    if (lane == 0u) base = atomicAdd(cursor, cs);
    base = __shfl_sync(PV_FULL_MASK, base, 0);
    i = (lane < cs && base + lane < n) ? base + lane : 0xFFFFFFFFu;
    return base < n;
#endif
}

template <bool ANY_HIT>
__device__ inline void mesh_hits_sync(bool active, const DScene& sc, uint32_t obj_index, const V3& o, const V3& d,
                                      HitAcc& acc, TStack stack, int sp0, unsigned int* overflow, double any_limit, double limit0, TravCount& tc)
{
    V3 mo = o, md = d;
    double len = 1.0;
    RayInfo ri;
    const DNode* __restrict__ nodes = sc.dmnodes;
    uint32_t tri_first = 0;
    bool any_ok = false, clipped = false;
    int sp = sp0;
    if (active) {
        const pvgpu_object& ob = sc.objs[obj_index];
        const DMesh& me = sc.meshes[ob.mesh];
#if PV_HEAVY
        if (me.node_count == 0) {
            // a mesh without its own tree (mesh.cpp:1490-1500): the generic, divergent form does it (lean scenes have none)
            mesh_hits<ANY_HIT>(sc, obj_index, ob, o, d, acc, -1, stack, sp0, overflow, any_limit);
            active = false;
        } else
#endif
        {
            if (ob.transform >= 0) {
                const pvgpu_transform& t = sc.xf[ob.transform];
                mo = inv_trans_point(t, o);
                md = inv_trans_direction(t, d);
                len = length(md);
                md = md / len;
            }
            ri = make_rayinfo(mo, md);
            nodes = sc.dmnodes + me.node_first;
            tri_first = me.tri_first;
            clipped = ob.clip_count != 0;
            any_ok = ANY_HIT && (ob.flags & PVGPU_OPAQUE_FLAG) && !clipped;
            const NodeL root = load_node(nodes);
            float dmin;
            if (slab_test(root.lo, root.hi, ri, dmin)) stack.set(sp++, make_uint2(__float_as_uint(dmin), root.code));
        }
    }
    // Every turn of the loop the warp votes: the lanes that want to walk an inner node and the lanes that hold a
    // triangle are counted, and the larger group is served (node visit or FP64 triangle test); the others wait for
    // their turn.  Both branches are warp-uniform, so the lanes served run converged.
    uint32_t tri = PV_NONE;
    for (;;) {
        const bool want = active && tri == PV_NONE && sp > sp0;
        const int want_n = vote_count(want), hold_n = vote_count(tri != PV_NONE);
        if ((want_n | hold_n) == 0) break;
        if (hold_n * PV_LEAF_BIAS_NUM > want_n * PV_LEAF_BIAS_DEN || want_n == 0) {
            if (tri != PV_NONE) {
                double t;
                const uint32_t ti = tri_first + tri;
                tri = PV_NONE;
                tc.prims++;
                if (tri_intersect(sc.dtris[ti], mo, md, t)) {
                    const double wd = t / len;
                    const V3 ip = evaluate(o, d, wd);
                    if (!clipped || point_in_clip(sc, sc.objs[obj_index], ip, stack, sp)) consider(acc, wd, ip, obj_index, ti, -1);
                    if (any_ok && acc.found && acc.closest > PV_SHADOW_TOLERANCE && acc.closest < any_limit) sp = sp0;   // blocked: drop the rest
                }
            }
        } else if (want) {
            const uint2 e = stack.get(--sp);
            // entries further than the closest accepted hit cannot improve it; compared in mesh space (t = world * len)
            const double lim = fmin(acc.closest, limit0) * len;
            if (!((double)__uint_as_float(e.x) > lim)) {
                const uint32_t code = e.y >> 28, idx = e.y & PV_CODE_INDEX;
                if (code == 0u) tri = idx;
                else tc.nodes += code, push_children<false, !ANY_HIT>(nodes, idx, code, ri, stack, sp, overflow, (lim < 3.0e38) ? __double2float_ru(lim) : 3.0e38f);
            }
        }
    }
}

// Mesh::Inside + inside_bbox_tree (mesh.cpp:197-262, 2266-2315): parity of crossings along Inside_Vect.
__device__ inline bool mesh_inside(const DScene& sc, const pvgpu_object& ob, const V3& p, TStack stack, int sp0)
{
    const DMesh& me = sc.meshes[ob.mesh];
    if (!me.has_inside_vector) return false;
    V3 mo = p, md = ld3(me.inside_vector);
    if (ob.transform >= 0) {
        const pvgpu_transform& t = sc.xf[ob.transform];
        mo = inv_trans_point(t, p);
        md = normalized(inv_trans_direction(t, md));
    }
    unsigned found = 0;
    if (me.node_count == 0) {
        for (uint32_t i = 0; i < me.tri_count; i++) {
            double t;
            if (tri_intersect(sc.dtris[me.tri_first + i], mo, md, t)) found++;
        }
    } else {
        const RayInfo ri = make_rayinfo(mo, md);
        const DNode* __restrict__ nodes = sc.dmnodes + me.node_first;
        int sp = sp0;
        float dmin;
        unsigned int ovf = 0;
        const NodeL root = load_node(nodes);
        if (slab_test(root.lo, root.hi, ri, dmin)) stack.set(sp++, make_uint2(0u, root.code));
        while (sp > sp0) {
            const uint2 e = stack.get(--sp);
            const uint32_t code = e.y >> 28, idx = e.y & PV_CODE_INDEX;
            if (code) {
                const uint32_t first = idx, count = code;
                push_children<false, false>(nodes, first, count, ri, stack, sp, &ovf);
            } else {
                double t;
                if (tri_intersect(sc.dtris[me.tri_first + idx], mo, md, t)) found++;
            }
        }
    }
    bool inside = (found & 1u) != 0;
    if (ob.flags & PVGPU_INVERTED_FLAG) inside = !inside;
    return inside;
}

// CSG*::All_Intersections (csg.cpp:128-375) without recursion: every primitive descendant contributes
// its hits; a hit survives iff, walking up to the CSG being tested, every intersection ancestor has the
// point inside all other children, every merge ancestor has it inside none of the other children, and
// every ancestor's clipped_by list contains it.  Intersection::Csg ends up as the outermost ancestor
// that sets it (unions without clipped_by do not, csg.cpp:137-150).
// Ray_In_Bound (object.cpp:385-400) for the bounded_by list of a CSG child: the ray must hit, or start inside, every bounding object.
// Bounding objects of CSG children are plain primitives (validated on the host), so no object graph is walked here.
#if PV_CLIPBOUND
static __device__ __noinline__ bool ray_in_prim_bounds(const DScene& sc, const pvgpu_object& ob, const V3& o, const V3& d, TStack stack, int sp0)
{
    for (uint32_t i = 0; i < ob.bound_count; i++) {
        const uint32_t bi = sc.index_list[ob.bound_first + i];
        const pvgpu_object& b = sc.objs[bi];
        bool hit = false;
        if (!type_uses_bbox_test(b.type) || object_bbox_test(b.bbox, o, d, (float)PV_HUGE_VAL)) {
            PrimHits h;
            prim_hits(sc, b, o, d, h);
            for (int k = 0; k < h.n; k++) if (h.depth[k] >= PV_MIN_ISECT_DEPTH) hit = true;       // Find_Intersection(isect, object, ray): any hit counts
        }
        if (!hit && !inside_object(sc, bi, o, stack, sp0)) return false;
    }
    return true;
}
#endif

// `limit`: hits at or beyond this depth cannot win (the caller only takes a hit nearer than its best so far), so their Inside
// tests - the expensive part - are skipped, as are those of hits consider() would drop anyway (too near, behind the post-condition,
// farther than the object's closest accepted hit so far).  The tests are pure functions of the point: skipping them changes nothing.
#define PV_CSG_HIT_CAN_WIN(depth) ((depth) <= acc.closest && (depth) < limit && (depth) >= PV_MIN_ISECT_DEPTH && (depth) > acc.post_min)
__device__ inline void csg_hits(const DScene& sc, uint32_t top, const V3& o, const V3& d, uint32_t rflags, bool shadow_ray,
                                HitAcc& acc, TStack stack, int sp0, unsigned int* overflow, double limit)
{
    uint2 range = sc.csg_leaf_range[top];
    if (range.y & 0x80000000u) {
        // fast path for the common shape: one CSG level whose children are simple primitives without clip lists, so
        // Inside_Object(sibling) is the primitive's own Inside() and needs no walk over an object graph
        const pvgpu_object& po = sc.objs[top];
        const uint32_t n = range.y & 0x7FFFFFFFu;
        for (uint32_t li = 0; li < n; li++) {
            const uint32_t leaf = sc.csg_leaves[range.x + li];
            const pvgpu_object& lo = sc.objs[leaf];
            if (po.type == PVGPU_OBJ_CSG_UNION && !test_ray_flags(lo.flags, rflags, shadow_ray, false)) continue;
            if (po.type == PVGPU_OBJ_CSG_MERGE && !test_ray_flags(lo.flags, rflags, shadow_ray, true)) continue;
            PrimHits h;
            prim_hits(sc, lo, o, d, h);
            for (int i = 0; i < h.n; i++) {
                if (!PV_CSG_HIT_CAN_WIN(h.depth[i])) continue;
                const V3 ip = h.ip[i];
                bool keep = true;
                int32_t csg = (int32_t)top;
                if (po.type == PVGPU_OBJ_CSG_INTERSECTION) {
                    for (uint32_t k = 0; k < n && keep; k++) {
                        const uint32_t sib = sc.csg_leaves[range.x + k];
                        if (sib != leaf && !simple_inside(sc, sc.objs[sib], ip)) keep = false;
                    }
                    if (keep && po.clip_count && !point_in_clip(sc, po, ip, stack, sp0)) keep = false;
                } else if (po.type == PVGPU_OBJ_CSG_MERGE) {
                    if (po.clip_count && !point_in_clip(sc, po, ip, stack, sp0)) keep = false;
                    for (uint32_t k = 0; k < n && keep; k++) {
                        const uint32_t sib = sc.csg_leaves[range.x + k];
                        if (sib != leaf && test_ray_flags(sc.objs[sib].flags, rflags, shadow_ray, true) && simple_inside(sc, sc.objs[sib], ip)) keep = false;
                    }
                } else {   // union: Intersection::Csg is only set when the union clips (csg.cpp:137-150)
                    if (po.clip_count) { if (!point_in_clip(sc, po, ip, stack, sp0)) keep = false; }
                    else csg = -1;
                }
                if (keep) consider(acc, h.depth[i], ip, leaf, h.aux[i], csg);
            }
        }
        return;
    }
    // The general shape.  `survives`: a hit of `leaf` at `ip` passes every ancestor up to `top`; csg = what Intersection::Csg becomes.
    auto survives = [&](uint32_t leaf, const V3& ip, int sp_now, int32_t& csg) -> bool {
        bool keep = true;
        csg = -1;
        uint32_t child = leaf;
        while (child != top && keep) {
            const uint32_t par = (uint32_t)sc.objs[child].parent;
            const pvgpu_object& po = sc.objs[par];
            if (po.type == PVGPU_OBJ_CSG_INTERSECTION) {
                for (uint32_t k = 0; k < po.child_count && keep; k++) {
                    const uint32_t sib = sc.index_list[po.child_first + k];
                    if (sib != child && !inside_object(sc, sib, ip, stack, sp_now)) keep = false;
                }
                if (keep && po.clip_count && !point_in_clip(sc, po, ip, stack, sp_now)) keep = false;
                if (keep) csg = (int32_t)par;
            } else if (po.type == PVGPU_OBJ_CSG_MERGE) {
                if (po.clip_count && !point_in_clip(sc, po, ip, stack, sp_now)) keep = false;
                for (uint32_t k = 0; k < po.child_count && keep; k++) {
                    const uint32_t sib = sc.index_list[po.child_first + k];
                    if (sib != child && test_ray_flags(sc.objs[sib].flags, rflags, shadow_ray, true) &&
                        inside_object(sc, sib, ip, stack, sp_now)) keep = false;
                }
                if (keep) csg = (int32_t)par;
            } else {   // union
                if (po.clip_count) {
                    if (!point_in_clip(sc, po, ip, stack, sp_now)) keep = false;
                    else csg = (int32_t)par;
                }
            }
            child = par;
        }
        return keep;
    };
    for (uint32_t li = 0; li < range.y; li++) {
        const uint32_t leaf = sc.csg_leaves[range.x + li];
        const pvgpu_object& lo = sc.objs[leaf];
        // ray-kind visibility of the leaf and of every ancestor below `top` (children of unions / merges only), and their
        // bounded_by lists: a child is only intersected when the ray is in its bounds (Ray_In_Bound, csg.cpp:163-166, 226-229, 322-325)
        bool visible = true;
        for (uint32_t c = leaf; c != top && visible; c = (uint32_t)sc.objs[c].parent) {
            const uint32_t ptype = sc.objs[sc.objs[c].parent].type;
            if (ptype == PVGPU_OBJ_CSG_UNION) visible = test_ray_flags(sc.objs[c].flags, rflags, shadow_ray, false);
            else if (ptype == PVGPU_OBJ_CSG_MERGE) visible = test_ray_flags(sc.objs[c].flags, rflags, shadow_ray, true);
#if PV_CLIPBOUND
            if (visible && sc.objs[c].bound_count) visible = ray_in_prim_bounds(sc, sc.objs[c], o, d, stack, sp0);
#endif
        }
        if (!visible) continue;
        if (lo.type == PVGPU_OBJ_MESH) {
            // Mesh::All_Intersections of a CSG child (mesh.cpp:138-195): every triangle hit is a candidate
            if (PV_HAS(PVGPU_OBJ_MESH))
                mesh_hits_f<false>(sc, leaf, lo, o, d, acc, stack, sp0, overflow, 0.0, [&](double wd, const V3& ip, uint32_t ti, int sp_now) {
                    if (!PV_CSG_HIT_CAN_WIN(wd)) return;
                    if (lo.clip_count && !point_in_clip(sc, lo, ip, stack, sp_now)) return;
                    int32_t csg;
                    if (survives(leaf, ip, sp_now, csg)) consider(acc, wd, ip, leaf, ti, csg);
                });
            continue;
        }
        int resume = (lo.type == PVGPU_OBJ_BLOB || lo.type == PVGPU_OBJ_GLYPH || lo.type == PVGPU_OBJ_PRISM || lo.type == PVGPU_OBJ_SUPERELLIPSOID) ? 0 : -1;       // blob and glyph children report their hits in batches (blob_hits, glyph_hits)
        do {
            PrimHits h;
            prim_hits(sc, lo, o, d, h, overflow, (resume >= 0) ? &resume : nullptr);
            for (int i = 0; i < h.n; i++) {
                if (!PV_CSG_HIT_CAN_WIN(h.depth[i])) continue;
                const V3 ip = h.ip[i];
                if (lo.clip_count && !point_in_clip(sc, lo, ip, stack, sp0)) continue;
                int32_t csg;
                if (survives(leaf, ip, sp0, csg)) consider(acc, h.depth[i], ip, leaf, h.aux[i], csg);
            }
        } while (resume >= 0);
    }
}

static __device__ __noinline__ bool object_find_simple(const DScene& sc, uint32_t idx, const V3& o, const V3& d, uint32_t rflags,
                                                TStack stack, int sp0, unsigned int* overflow);

// Find_Intersection for one frame-level object (object.cpp:172-224 / trace.cpp:345-443).
template <bool ANY_OPAQUE>
__device__ inline bool object_find(const DScene& sc, uint32_t idx, const V3& o, const V3& d, uint32_t rflags, bool shadow_ray,
                                   double post_min, float bbox_maxd, Hit& out, TStack stack, int sp0, unsigned int* overflow,
                                   double opaque_limit = 0.0, double limit = PV_HUGE_VAL)
{
    const pvgpu_object& ob = sc.objs[idx];
    if (type_uses_bbox_test(ob.type) && !object_bbox_test(ob.bbox, o, d, bbox_maxd)) return false;
    // Ray_In_Bound (object.cpp:385-400)
#if PV_CLIPBOUND
    for (uint32_t i = 0; i < ob.bound_count; i++) {
        const uint32_t b = sc.index_list[ob.bound_first + i];
        if (!object_find_simple(sc, b, o, d, rflags, stack, sp0, overflow) && !inside_object(sc, b, o, stack, sp0)) return false;
    }
#endif
    HitAcc acc;
    acc.closest = PV_HUGE_VAL; acc.post_min = post_min; acc.found = false;
#if PV_HEAVY
    // (the lean variant walks meshes only through mesh_hits_sync and serves no CSG)
    if (PVGPU_IS_CSG(ob.type)) { if (PV_HAS(PVGPU_OBJ_CSG_UNION)) csg_hits(sc, idx, o, d, rflags, shadow_ray, acc, stack, sp0, overflow, limit); }
    else if (PV_HAS(PVGPU_OBJ_MESH) && ob.type == PVGPU_OBJ_MESH) mesh_hits<ANY_OPAQUE>(sc, idx, ob, o, d, acc, -1, stack, sp0, overflow, opaque_limit);
    else
#endif
    {
        int resume = (PV_HEAVY && PV_HAS(PVGPU_OBJ_GLYPH) && (ob.type == PVGPU_OBJ_GLYPH || ob.type == PVGPU_OBJ_PRISM || ob.type == PVGPU_OBJ_SUPERELLIPSOID)) ? 0 : -1;       // glyph / prism hits come in batches
        do {
            PrimHits h;
            prim_hits(sc, ob, o, d, h, overflow, (resume >= 0) ? &resume : nullptr);
            for (int i = 0; i < h.n; i++)
                if (ob.clip_count == 0 || point_in_clip(sc, ob, h.ip[i], stack, sp0)) consider(acc, h.depth[i], h.ip[i], idx, h.aux[i], -1);
        } while (resume >= 0);
    }
    if (acc.found) out = acc.best;
    return acc.found;
}

// The plain Find_Intersection(isect, object, ray) used for bounded_by objects (no post-condition,
// maxd = HUGE_VAL; nested bounded_by lists are rejected on the host).
static __device__ __noinline__ bool object_find_simple(const DScene& sc, uint32_t idx, const V3& o, const V3& d, uint32_t rflags,
                                                TStack stack, int sp0, unsigned int* overflow)
{
    const pvgpu_object& ob = sc.objs[idx];
    if (type_uses_bbox_test(ob.type) && !object_bbox_test(ob.bbox, o, d, (float)PV_HUGE_VAL)) return false;
    HitAcc acc;
    acc.closest = PV_HUGE_VAL; acc.post_min = -1.0; acc.found = false;
#if PV_HEAVY
    if (PVGPU_IS_CSG(ob.type)) { if (PV_HAS(PVGPU_OBJ_CSG_UNION)) csg_hits(sc, idx, o, d, rflags, false, acc, stack, sp0, overflow, PV_HUGE_VAL); }
    else
#endif
    if (PV_HAS(PVGPU_OBJ_MESH) && ob.type == PVGPU_OBJ_MESH) mesh_hits<false>(sc, idx, ob, o, d, acc, -1, stack, sp0, overflow);
    else {
        int resume = (PV_HEAVY && PV_HAS(PVGPU_OBJ_GLYPH) && (ob.type == PVGPU_OBJ_GLYPH || ob.type == PVGPU_OBJ_PRISM || ob.type == PVGPU_OBJ_SUPERELLIPSOID)) ? 0 : -1;       // glyph / prism hits come in batches
        do {
            PrimHits h;
            prim_hits(sc, ob, o, d, h, overflow, (resume >= 0) ? &resume : nullptr);
            for (int i = 0; i < h.n; i++)
                if (ob.clip_count == 0 || point_in_clip(sc, ob, h.ip[i], stack, sp0)) consider(acc, h.depth[i], h.ip[i], idx, h.aux[i], -1);
        } while (resume >= 0);
    }
    return acc.found;
}

// Pre-condition of the traversal (NoSomethingFlagRayObjectCondition trace.cpp:84-95 for TraceRay,
// NoShadowFlagRayObjectCondition trace.cpp:1943 for shadow rays).
__device__ __forceinline__ bool precondition(uint32_t oflags, uint32_t rflags, bool shadow_ray)
{
    if (shadow_ray) return !(oflags & PVGPU_NO_SHADOW_FLAG);
    const bool reflection = (rflags & PV_RAY_REFLECTION) != 0;
    const bool image = (rflags & PV_RAY_PRIMARY) || ((rflags & PV_RAY_REFRACTION) && !reflection);
    if (image && (oflags & PVGPU_NO_IMAGE_FLAG)) return false;
    if (reflection && (oflags & PVGPU_NO_REFLECTION_FLAG)) return false;
    return true;
}

// Trace::FindIntersection(bestisect, ray, precondition, postcondition) (trace.cpp:285-344).
//   best.depth must be preset to the search limit (HUGE_VAL, Max_Ray_Distance or the light distance).
//   ANY_OPAQUE: shadow-ray mode that returns as soon as a hit on an OPAQUE object satisfies the shadow
//   window (depth in (SHADOW_TOLERANCE, limit - SHADOW_TOLERANCE)); see trace_shadow in pv_shade.cuh.
template <bool ANY_OPAQUE>
__device__ inline bool find_intersection(const DScene& sc, const V3& o, const V3& d, uint32_t rflags, bool shadow_ray,
                                         double post_min, Hit& best, TStack stack, unsigned int* overflow,
                                         double opaque_limit = 0.0)
{
    bool found = false;
    if (!sc.use_tree) {
        // boundingMethod 0: linear loop over SceneData::objects (trace.cpp:321-340)
        for (uint32_t i = 0; i < sc.n_frame; i++) {
            const uint32_t idx = sc.frame[i];
            if (!precondition(sc.objs[idx].flags, rflags, shadow_ray)) continue;
            Hit h;
            if (object_find<ANY_OPAQUE>(sc, idx, o, d, rflags, shadow_ray, post_min, (float)PV_HUGE_VAL, h, stack, 0, overflow, opaque_limit, best.depth) && h.depth < best.depth) {
                best = h;
                found = true;
                if (ANY_OPAQUE && (sc.objs[h.obj].flags & PVGPU_OPAQUE_FLAG) && h.depth > PV_SHADOW_TOLERANCE && h.depth < opaque_limit) return true;
            }
        }
        return found;
    }
    const RayInfo ri = make_rayinfo(o, d);
    const DNode* __restrict__ nodes = sc.dnodes;
    int sp = 0;
    {
        const NodeL root = load_node(nodes);
        float dmin;
        if (root.code & PV_CODE_INFINITE) dmin = -PV_MAX_DISTANCE_F;
        else if (!slab_test(root.lo, root.hi, ri, dmin)) return false;
        stack.set(sp++, make_uint2(__float_as_uint(dmin), root.code));
    }
    for (;;) {
        uint32_t leaf = 0xFFFFFFFFu;
        while (sp > 0) {
            const uint2 e = stack.get(--sp);
            if ((double)__uint_as_float(e.x) > best.depth) continue;      // "Depth > Best_Intersection->Depth" (boundingbox.cpp:517)
            const uint32_t code = e.y >> 28, idx = e.y & PV_CODE_INDEX;
            if (code == 0u) { leaf = idx; break; }
            const uint32_t first = idx, count = code;
            push_children<true, true>(nodes, first, count, ri, stack, sp, overflow);
        }
        if (leaf == 0xFFFFFFFFu) break;
        if (!precondition(sc.objs[leaf].flags, rflags, shadow_ray)) continue;
        Hit h;
        if (object_find<ANY_OPAQUE>(sc, leaf, o, d, rflags, shadow_ray, post_min, (float)PV_HUGE_VAL, h, stack, sp, overflow, opaque_limit, best.depth) && h.depth < best.depth) {
            best = h;
            found = true;
            if (ANY_OPAQUE && (sc.objs[h.obj].flags & PVGPU_OPAQUE_FLAG) && h.depth > PV_SHADOW_TOLERANCE && h.depth < opaque_limit) return true;
        }
    }
    return found;
}

// Warp-synchronous form of find_intersection used by the wavefront kernels.  ALL 32 lanes of the warp call it
// together; lanes with `alive` carry a ray.  Phase A walks the scene tree until every lane holds a leaf object or has
// run out of work, phase B runs the per-object tests (meshes through mesh_hits_sync, so the nested traversal stays
// converged as well).  Same results as find_intersection(); only the scheduling of the work inside the warp differs.
template <bool ANY_OPAQUE>
__device__ inline bool find_intersection_sync(bool alive, const DScene& sc, const V3& o, const V3& d, uint32_t rflags, bool shadow_ray,
                                              double post_min, Hit& best, TStack stack, unsigned int* overflow, TravCount& tc,
                                              double opaque_limit = 0.0)
{
    bool found = false;
    int sp = 0;
    // phase B for the lanes that hold `leaf`; returns true when an ANY_OPAQUE search is satisfied
    auto leaf_phase = [&](bool has_leaf, uint32_t leaf) -> bool {
        bool is_mesh = false, done = false;
        HitAcc acc;
        acc.closest = PV_HUGE_VAL; acc.post_min = post_min; acc.found = false;
        if (has_leaf) {
            const pvgpu_object& ob = sc.objs[leaf];
            tc.prims++;
            if (PV_HAS(PVGPU_OBJ_MESH) && ob.type == PVGPU_OBJ_MESH) {
                // object_find's prelude: FP32 box test and Ray_In_Bound (object.cpp:186-193, 385-400)
                is_mesh = object_bbox_test(ob.bbox, o, d, (float)PV_HUGE_VAL);
#if PV_CLIPBOUND
                for (uint32_t i = 0; is_mesh && i < ob.bound_count; i++) {
                    const uint32_t b = sc.index_list[ob.bound_first + i];
                    if (!object_find_simple(sc, b, o, d, rflags, stack, sp, overflow) && !inside_object(sc, b, o, stack, sp)) is_mesh = false;
                }
#endif
            } else {
                Hit h;
                if (object_find<ANY_OPAQUE>(sc, leaf, o, d, rflags, shadow_ray, post_min, (float)PV_HUGE_VAL, h, stack, sp, overflow, opaque_limit, best.depth) && h.depth < best.depth) {
                    best = h;
                    found = true;
                    if (ANY_OPAQUE && (sc.objs[h.obj].flags & PVGPU_OPAQUE_FLAG) && h.depth > PV_SHADOW_TOLERANCE && h.depth < opaque_limit) done = true;
                }
            }
        }
        if (PV_HAS(PVGPU_OBJ_MESH) && vote_any(is_mesh)) {
            mesh_hits_sync<ANY_OPAQUE>(is_mesh, sc, leaf, o, d, acc, stack, sp, overflow, opaque_limit, best.depth, tc);
            if (is_mesh && acc.found && acc.best.depth < best.depth) {
                best = acc.best;
                found = true;
                if (ANY_OPAQUE && (sc.objs[best.obj].flags & PVGPU_OPAQUE_FLAG) && best.depth > PV_SHADOW_TOLERANCE && best.depth < opaque_limit) done = true;
            }
        }
        return done;
    };

    if (!sc.use_tree) {
        // boundingMethod 0: linear loop over SceneData::objects (trace.cpp:321-340); the trip count is uniform
        for (uint32_t i = 0; i < sc.n_frame; i++) {
            const uint32_t idx = sc.frame[i];
            const bool cand = alive && precondition(sc.objs[idx].flags, rflags, shadow_ray);
            if (leaf_phase(cand, idx)) alive = false;
        }
        return found;
    }
    RayInfo ri;
    const DNode* __restrict__ nodes = sc.dnodes;
    if (alive) {
        ri = make_rayinfo(o, d);
        const NodeL root = load_node(nodes);
        float dmin;
        bool ok = true;
        if (root.code & PV_CODE_INFINITE) dmin = -PV_MAX_DISTANCE_F;
        else ok = slab_test(root.lo, root.hi, ri, dmin);
        if (ok) stack.set(sp++, make_uint2(__float_as_uint(dmin), root.code));
    }
    uint32_t leaf = PV_NONE;
    for (;;) {
        const bool want = alive && leaf == PV_NONE && sp > 0;
        const int want_n = vote_count(want), hold_n = vote_count(leaf != PV_NONE);
        if ((want_n | hold_n) == 0) break;
        if (hold_n * PV_LEAF_BIAS_NUM > want_n * PV_LEAF_BIAS_DEN || want_n == 0) {
            const uint32_t cur = leaf;
            leaf = PV_NONE;
            if (leaf_phase(cur != PV_NONE, cur)) { alive = false; sp = 0; }
        } else if (want) {
            const uint2 e = stack.get(--sp);
            if (!((double)__uint_as_float(e.x) > best.depth)) {      // "Depth > Best_Intersection->Depth" (boundingbox.cpp:517)
                const uint32_t code = e.y >> 28, idx = e.y & PV_CODE_INDEX;
                if (code == 0u) { if (precondition(sc.objs[idx].flags, rflags, shadow_ray)) leaf = idx; }
                else tc.nodes += code, push_children<true, !ANY_OPAQUE>(nodes, idx, code, ri, stack, sp, overflow, (best.depth < 3.0e38) ? __double2float_ru(best.depth) : 3.0e38f);
            }
        }
    }
    return found;
}

}  // namespace pvgpu
