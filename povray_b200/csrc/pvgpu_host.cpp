// Host side of the pvgpu C ABI: scene container, table setters, validation, the bounding-tree builder
// (for scenes that do not come with a BBOX_TREE from POV-Ray's BoundingTask), the mesh2 helper and
// the flat-scene file format.  No tracing happens here; everything that computes pixels or
// intersections lives in pvgpu_device.cu and runs on the GPU only.
#include "pvgpu_scene.hpp"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <future>
#include <functional>
#include <limits>

namespace pvgpu {

// ------------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";

int fail(int code, const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
void clear_error() { g_err[0] = 0; }

// ------------------------------------------------------------------------------------------------
// Bounding tree construction.
//
// Follows Build_BBox_Tree / sort_and_split / find_axis / calc_bbox / build_area_table
// (source/core/bounding/boundingbox.cpp:262-323, 700-927) so that a scene assembled through the ABI
// gets the SAME tree POV-Ray's BoundingTask would hand us: leaves are bunched bottom-up into nodes of
// at most BUNCHING_FACTOR (4) entries unless splitting stops paying off, and the pass is repeated on
// the freshly made nodes until one node is left.  All box arithmetic is FP32 exactly as there.
// ------------------------------------------------------------------------------------------------
namespace {

constexpr int    kBunching  = 4;          // BUNCHING_FACTOR, boundingbox.cpp:68
constexpr double kBoundHuge = 2.0e10;     // BOUND_HUGE, configcore.h:184

struct BuildNode {
    float lo[3], size[3];
    std::vector<int64_t> kids;            // empty: leaf
    uint32_t payload = 0;
    bool infinite = false;
};

// One pass of the reference's sort_and_split over work[first, last) creates ONE level of new nodes (bunches of the current
// level's entries), left to right; the passes repeat on the new level until a single node is left.  The two halves of a split
// touch disjoint sub-ranges of `work`, so the upper levels of the recursion run as parallel tasks; each task collects its new
// nodes in its own list and the lists are concatenated left before right, which is the order the sequential recursion creates
// them in - node numbering, and with it the final table, is byte-identical to the reference's (tests/golden/make_golden_1080.py
// checks that against the parser's own tree).  The comparator reads precomputed FP32 keys instead of re-deriving them.
struct TreeBuilder {
    std::deque<BuildNode> pool;
    std::vector<int64_t>  work;           // the reference's `Finite` array (ids into pool)
    std::vector<float>    box;            // per pool id: lo[3], size[3]     (compact copy for the hot loops)
    std::vector<float>    key[3];         // per pool id: compboxes<>' sort key on each axis
    int64_t root = -1;

    static thread_local const float* sort_key;

    void note_box(const BuildNode& n)
    {
        for (int k = 0; k < 3; k++) box.push_back(n.lo[k]);
        for (int k = 0; k < 3; k++) box.push_back(n.size[k]);
        // compboxes<>: am = 2.0 * lowerLeft + size evaluated in double, stored as BBoxScalar
        for (int k = 0; k < 3; k++) key[k].push_back((float)(2.0 * n.lo[k] + n.size[k]));
    }

    static int cmp(const void* pa, const void* pb)
    {
        const float am = sort_key[(size_t)*(const int64_t*)pa], bm = sort_key[(size_t)*(const int64_t*)pb];
        if (am < bm) return -1;
        return (am == bm) ? 0 : 1;
    }

    int find_axis(ptrdiff_t first, ptrdiff_t last) const
    {
        float mins[3], maxs[3];
        for (int k = 0; k < 3; k++) { mins[k] = (float)kBoundHuge; maxs[k] = (float)-kBoundHuge; }
        for (ptrdiff_t i = first; i < last; i++) {
            const float* n = box.data() + 6 * (size_t)work[i];
            for (int k = 0; k < 3; k++) {
                if (n[k] < mins[k]) mins[k] = n[k];
                float hi = n[k] + n[3 + k];
                if (hi > maxs[k]) maxs[k] = hi;
            }
        }
        int which = 0;
        float d = (float)-kBoundHuge, e;
        e = maxs[0] - mins[0]; if (e > d) { d = e; which = 0; }
        e = maxs[1] - mins[1]; if (e > d) { d = e; which = 1; }
        e = maxs[2] - mins[2]; if (e > d) { which = 2; }
        return which;
    }

    // calc_bbox: double min/max of FP32 boxes, stored back as FP32 lowerLeft / size
    void calc_bbox(BuildNode& out, const std::vector<int64_t>& ids) const
    {
        double bmin[3] = { kBoundHuge, kBoundHuge, kBoundHuge };
        double bmax[3] = { -kBoundHuge, -kBoundHuge, -kBoundHuge };
        for (int64_t id : ids) {
            const float* n = box.data() + 6 * (size_t)id;
            for (int k = 0; k < 3; k++) {
                double tmin = n[k];
                double tmax = tmin + n[3 + k];
                if (tmin < bmin[k]) bmin[k] = tmin;
                if (tmax > bmax[k]) bmax[k] = tmax;
            }
        }
        for (int k = 0; k < 3; k++) { out.lo[k] = (float)bmin[k]; out.size[k] = (float)(bmax[k] - bmin[k]); }
    }

    void area_table(ptrdiff_t a, ptrdiff_t b, float* areas) const
    {
        ptrdiff_t imin = (a < b) ? a : b, dir = (a < b) ? 1 : -1;
        float bmin[3], bmax[3];
        for (int k = 0; k < 3; k++) { bmin[k] = (float)kBoundHuge; bmax[k] = (float)-kBoundHuge; }
        for (ptrdiff_t i = a; i != b + dir; i += dir) {
            const float* n = box.data() + 6 * (size_t)work[i];
            for (int k = 0; k < 3; k++) {
                float tmin = n[k], tmax = tmin + n[3 + k];
                if (tmin < bmin[k]) bmin[k] = tmin;
                if (tmax > bmax[k]) bmax[k] = tmax;
            }
            float lx = bmax[0] - bmin[0], ly = bmax[1] - bmin[1], lz = bmax[2] - bmin[2];
            areas[i - imin] = lx * (ly + lz) + ly * lz;
        }
    }

    // the new nodes of work[first, last), left to right, as lists of their entries; returns true when the range was split
    bool split(ptrdiff_t first, ptrdiff_t last, std::vector<std::vector<int64_t>>& made, int depth)
    {
        ptrdiff_t size = last - first, best_loc = -1;
        if (size <= 0) return false;
        if (size > kBunching) {
            const int axis = find_axis(first, last);
            sort_key = key[axis].data();
            std::qsort(work.data() + first, (size_t)size, sizeof(int64_t), cmp);   // same libc routine as the reference
            std::vector<float> area((size_t)(2 * size));
            float* area_left = area.data();
            float* area_right = area_left + size;
            area_table(first, last - 1, area_left);
            area_table(last - 1, first, area_right);
            float best_index = area_right[0] * float(size - 3);   // cost of not subdividing
            for (ptrdiff_t i = 1; i < size; i++) {
                float new_index = float(i) * area_left[i - 1] + float(size - i) * area_right[i];
                if (new_index < best_index) { best_index = new_index; best_loc = i + first; }
            }
        }
        if (best_loc < 0) {
            made.emplace_back(work.begin() + first, work.begin() + last);
            return false;
        }
        if (depth < 4 && size > 20000) {
            std::vector<std::vector<int64_t>> right;
            std::future<void> left = std::async(std::launch::async, [&] { split(first, best_loc, made, depth + 1); });
            split(best_loc, last, right, depth + 1);
            left.get();
            for (auto& r : right) made.push_back(std::move(r));
        } else {
            split(first, best_loc, made, depth + 1);
            split(best_loc, last, made, depth + 1);
        }
        return true;
    }

    bool sort_and_split(ptrdiff_t first, ptrdiff_t last)
    {
        std::vector<std::vector<int64_t>> made;
        const bool was_split = split(first, last, made, 0);
        for (auto& kids : made) {
            BuildNode n;
            n.kids = std::move(kids);
            calc_bbox(n, n.kids);
            note_box(n);
            pool.push_back(std::move(n));
            root = (int64_t)pool.size() - 1;
            work.push_back(root);
        }
        return was_split;
    }
};
thread_local const float* TreeBuilder::sort_key = nullptr;

}  // namespace

void build_bbox_tree(const std::vector<LeafBox>& finite, const std::vector<LeafBox>& infinite,
                     std::vector<pvgpu_node>& out)
{
    out.clear();
    TreeBuilder tb;
    auto add_leaf = [&](const LeafBox& b, bool inf) {
        BuildNode n;
        std::memcpy(n.lo, b.lo, sizeof n.lo);
        std::memcpy(n.size, b.size, sizeof n.size);
        n.payload = b.payload;
        n.infinite = inf;
        tb.note_box(n);
        tb.pool.push_back(std::move(n));
        return (int64_t)tb.pool.size() - 1;
    };
    for (const LeafBox& b : finite) tb.work.push_back(add_leaf(b, false));
    std::vector<int64_t> inf_ids;
    for (const LeafBox& b : infinite) inf_ids.push_back(add_leaf(b, true));

    if (!finite.empty()) {
        ptrdiff_t low = 0, high = (ptrdiff_t)finite.size();
        while (tb.sort_and_split(low, high)) { low = high; high = (ptrdiff_t)tb.work.size(); }
        if (!inf_ids.empty()) {
            // infinite objects go into a new first child of the root (boundingbox.cpp:288-306)
            BuildNode cd;
            cd.kids = inf_ids;
            tb.calc_bbox(cd, cd.kids);
            cd.infinite = true;
            tb.note_box(cd);
            tb.pool.push_back(std::move(cd));
            int64_t cd_id = (int64_t)tb.pool.size() - 1;
            BuildNode& r = tb.pool[(size_t)tb.root];
            r.kids.insert(r.kids.begin(), cd_id);
            tb.calc_bbox(r, r.kids);
            r.infinite = true;
        }
    } else if (!inf_ids.empty()) {
        BuildNode cd;
        cd.kids = inf_ids;
        tb.calc_bbox(cd, cd.kids);
        cd.infinite = true;
        tb.note_box(cd);
        tb.pool.push_back(std::move(cd));
        tb.root = (int64_t)tb.pool.size() - 1;
    } else {
        return;
    }

    // breadth-first layout: children of a node are contiguous, root = 0
    std::vector<int64_t> order{ tb.root };
    out.resize(1);
    for (size_t qi = 0; qi < order.size(); qi++) {
        const BuildNode& n = tb.pool[(size_t)order[qi]];
        pvgpu_node pn{};
        std::memcpy(pn.lo, n.lo, sizeof pn.lo);
        std::memcpy(pn.size, n.size, sizeof pn.size);
        pn.flags = n.infinite ? PVGPU_NODE_INFINITE : 0;
        if (n.kids.empty()) {
            pn.count = 0;
            pn.first = n.payload;
        } else {
            pn.count = (uint16_t)n.kids.size();
            pn.first = (uint32_t)order.size();
            for (int64_t k : n.kids) order.push_back(k);
            out.resize(order.size());
        }
        out[qi] = pn;
    }
}

// ------------------------------------------------------------------------------------------------
// validation
// ------------------------------------------------------------------------------------------------
static bool range_ok(uint32_t first, uint32_t count, size_t n) { return (size_t)first + count <= n; }

int validate_scene(Scene& s)
{
    const size_t no = s.objects.size();
    // (a scene without objects is legal: every ray ends in ComputeSky)
    for (uint32_t f : s.frame)
        if (f >= no) return fail(PVGPU_E_INVALID, "frame object index %u out of range", f);
    // Tables other tables index into are validated first (blend maps and their entries, warps), then the records that refer
    // to them, then the nesting walks: nothing is dereferenced before its range has been checked.
    for (size_t i = 0; i < s.blend_maps.size(); i++) {
        const pvgpu_blend_map& m = s.blend_maps[i];
        if (m.entry_count == 0 || !range_ok(m.entry_first, m.entry_count, s.blend_entries.size()))
            return fail(PVGPU_E_INVALID, "blend map %zu: bad entry range", i);
        if ((m.blend_mode & ~(PVGPU_BLEND_PIGMENT_MAP | PVGPU_BLEND_TEXTURE_MAP | PVGPU_BLEND_NORMAL_MAP)) != 0)
            return fail(PVGPU_E_UNSUPPORTED, "blend map %zu: blend_mode %d is outside the hot-path scope", i, m.blend_mode & ~(PVGPU_BLEND_PIGMENT_MAP | PVGPU_BLEND_TEXTURE_MAP | PVGPU_BLEND_NORMAL_MAP));
        // entries of pigment / texture / normal maps carry an index into the respective table in colour[0]
        const size_t limit = (m.blend_mode & PVGPU_BLEND_PIGMENT_MAP) ? s.pigments.size() : (m.blend_mode & PVGPU_BLEND_TEXTURE_MAP) ? s.textures.size()
                           : (m.blend_mode & PVGPU_BLEND_NORMAL_MAP) ? s.tnormals.size() : 0;
        if (m.blend_mode & (PVGPU_BLEND_PIGMENT_MAP | PVGPU_BLEND_TEXTURE_MAP | PVGPU_BLEND_NORMAL_MAP))
            for (uint32_t k = 0; k < m.entry_count; k++) {
                const float pi = s.blend_entries[m.entry_first + k].colour[0];
                if (!(pi >= 0.0f) || pi >= (float)limit || pi != std::floor(pi))
                    return fail(PVGPU_E_INVALID, "blend map %zu: entry %u is not a valid table index", i, k);
            }
    }
    for (size_t i = 0; i < s.warps.size(); i++) {
        const pvgpu_warp& w = s.warps[i];
        if (w.type < PVGPU_WARP_TRANSFORM || w.type > PVGPU_WARP_PLANAR)
            return fail(PVGPU_E_UNSUPPORTED, "warp %zu: type %u unsupported", i, w.type);
        if (w.type >= PVGPU_WARP_BLACK_HOLE) {
            static const uint32_t n_par[] = { 11, 8, 0, 4, 4, 5, 4 };
            if (w.transform < 0 || !range_ok((uint32_t)w.transform, n_par[w.type - PVGPU_WARP_BLACK_HOLE], s.shape_data.size()))
                return fail(PVGPU_E_INVALID, "warp %zu: parameters outside the shape-data table", i);
            if (w.type == PVGPU_WARP_REPEAT && !(s.shape_data[w.transform] >= 0.0 && s.shape_data[w.transform] <= 2.0))
                return fail(PVGPU_E_INVALID, "warp %zu: repeat axis out of range", i);
        }
        if (w.type == PVGPU_WARP_TRANSFORM && (w.transform < 0 || w.transform >= (int32_t)s.transforms.size()))
            return fail(PVGPU_E_INVALID, "warp %zu: bad transform", i);
    }
    for (size_t i = 0; i < no; i++) {
        const pvgpu_object& o = s.objects[i];
        if (o.type < PVGPU_OBJ_SPHERE || o.type > PVGPU_OBJ_LAST)
            return fail(PVGPU_E_UNSUPPORTED, "object %zu: primitive type %u is outside the hot-path scope", i, o.type);
        if (!range_ok(o.child_first, o.child_count, s.index_list.size()) ||
            !range_ok(o.clip_first, o.clip_count, s.index_list.size()) ||
            !range_ok(o.bound_first, o.bound_count, s.index_list.size()))
            return fail(PVGPU_E_INVALID, "object %zu: index-list range out of bounds", i);
        if (o.texture >= (int32_t)s.textures.size() || o.interior_texture >= (int32_t)s.textures.size())
            return fail(PVGPU_E_INVALID, "object %zu: texture index out of range", i);
        if (o.interior >= (int32_t)s.interiors.size())
            return fail(PVGPU_E_INVALID, "object %zu: interior index out of range", i);
        if (o.transform >= (int32_t)s.transforms.size())
            return fail(PVGPU_E_INVALID, "object %zu: transform index out of range", i);
        if (o.parent >= (int32_t)no) return fail(PVGPU_E_INVALID, "object %zu: parent out of range", i);
        if (o.type == PVGPU_OBJ_MESH && (o.mesh < 0 || o.mesh >= (int32_t)s.meshes.size()))
            return fail(PVGPU_E_INVALID, "object %zu: mesh index out of range", i);
        if (o.type == PVGPU_OBJ_BLOB) {
            if (o.mesh < 0 || o.mesh >= (int32_t)s.blobs.size()) return fail(PVGPU_E_INVALID, "object %zu: blob index out of range", i);
        }
        if ((o.type == PVGPU_OBJ_CONE || o.type == PVGPU_OBJ_DISC) && o.transform < 0)
            return fail(PVGPU_E_INVALID, "object %zu: cone / cylinder / disc without transform", i);
        if (o.type == PVGPU_OBJ_TRIANGLE &&
            (o.mesh < 0 || !range_ok((uint32_t)o.mesh, (o.aux & PVGPU_TRIANGLE_SMOOTH) ? 25u : 13u, s.shape_data.size()) || (o.aux & 3u) > 2u || ((o.aux & PVGPU_TRIANGLE_SMOOTH) && ((o.aux >> 2) & 3u) > 2u)))
            return fail(PVGPU_E_INVALID, "object %zu: triangle record outside the shape-data table", i);
        if (o.type == PVGPU_OBJ_POLYGON &&
            (o.transform < 0 || o.mesh < 0 || o.aux > (1u << 24) || !range_ok((uint32_t)o.mesh, 2u * o.aux, s.shape_data.size())))
            return fail(PVGPU_E_INVALID, "object %zu: polygon without transform or with points outside the shape-data table", i);
        if (o.type == PVGPU_OBJ_POLY) {
            if (o.aux < 1 || o.aux > 4) return fail(PVGPU_E_UNSUPPORTED, "object %zu: poly of order %u (the device solver handles order <= 4)", i, o.aux);
            if (o.transform < 0 || o.mesh < 0 || !range_ok((uint32_t)o.mesh, (o.aux + 1) * (o.aux + 2) * (o.aux + 3) / 6, s.shape_data.size()))
                return fail(PVGPU_E_INVALID, "object %zu: poly without transform or with coefficients outside the shape-data table", i);
        }
        if (o.type == PVGPU_OBJ_GLYPH) {
            if (o.transform < 0 || o.mesh < 0 || !range_ok((uint32_t)o.mesh, 1u, s.shape_data.size()))
                return fail(PVGPU_E_INVALID, "object %zu: glyph without transform or with its outline outside the shape-data table", i);
            const double nseg = s.shape_data[o.mesh];
            if (!(nseg >= 0.0 && nseg <= (double)(1u << 24)) || nseg != (double)(uint32_t)nseg || !range_ok((uint32_t)o.mesh + 1u, 7u * (uint32_t)nseg, s.shape_data.size()))
                return fail(PVGPU_E_INVALID, "object %zu: glyph outline outside the shape-data table", i);
        }
        if (o.type == PVGPU_OBJ_PRISM) {
            const uint32_t spline = o.aux & 15u, sweep = (o.aux >> 4) & 15u;
            if (o.transform < 0 || o.mesh < 0 || !range_ok((uint32_t)o.mesh, 1u, s.shape_data.size()) || spline < 1 || spline > 4 || sweep < 1 || sweep > 2 || (o.aux >> 8))
                return fail(PVGPU_E_INVALID, "object %zu: prism without transform, with its spline outside the shape-data table or with an unknown spline / sweep type", i);
            const double number = s.shape_data[o.mesh];
            if (!(number >= 0.0 && number <= (double)(1u << 24)) || number != (double)(uint32_t)number || !range_ok((uint32_t)o.mesh + 1u, 15u * (uint32_t)number, s.shape_data.size()))
                return fail(PVGPU_E_INVALID, "object %zu: prism spline outside the shape-data table", i);
        }
        if (o.type == PVGPU_OBJ_SUPERELLIPSOID) {
            if (o.transform < 0) return fail(PVGPU_E_INVALID, "object %zu: superellipsoid without transform", i);
            if (o.clip_count) return fail(PVGPU_E_UNSUPPORTED, "object %zu: clipped_by on a superellipsoid (the clip test steers the reference's hit walk) is outside the hot-path scope", i);
        }
        if (o.type == PVGPU_OBJ_TORUS && o.transform < 0)
            return fail(PVGPU_E_INVALID, "object %zu: torus without transform", i);
        if ((o.flags & PVGPU_UV_FLAG) && o.type == PVGPU_OBJ_CONE)
            return fail(PVGPU_E_UNSUPPORTED, "object %zu: uv_mapping on a cone / cylinder is outside the hot-path scope", i);
        if ((o.flags & PVGPU_CUTAWAY_TEXTURES_FLAG) && o.texture < 0)
            return fail(PVGPU_E_UNSUPPORTED, "object %zu: cutaway_textures is outside the hot-path scope", i);
        if (o.parent >= 0 && o.bound_count) {
            // bounded_by on a CSG child: the device tests such lists without walking an object graph (ray_in_prim_bounds)
            for (uint32_t k = 0; k < o.bound_count; k++) {
                const uint32_t bi = s.index_list[o.bound_first + k];
                if (bi >= no) return fail(PVGPU_E_INVALID, "object %zu: bounded_by index out of range", i);
                const pvgpu_object& b = s.objects[bi];
                if (PVGPU_IS_CSG(b.type) || b.type == PVGPU_OBJ_MESH || b.type == PVGPU_OBJ_BLOB || b.bound_count || b.clip_count)
                    return fail(PVGPU_E_UNSUPPORTED, "object %zu: bounded_by on a CSG child with a compound bounding object", i);
            }
        }
    }
    for (size_t i = 0; i < s.images.size(); i++) {
        const pvgpu_image& im = s.images[i];
        if (im.width == 0 || im.height == 0 || (size_t)im.data_first + (size_t)im.width * im.height > s.texels.size() / 5)
            return fail(PVGPU_E_INVALID, "image %zu: texels outside the texel table", i);
        if (!(im.map_type == 0 || im.map_type == 1 || im.map_type == 2 || im.map_type == 5 || im.map_type == 7))
            return fail(PVGPU_E_UNSUPPORTED, "image %zu: map_type %u", i, im.map_type);
        if (!(im.interpolation == 0 || im.interpolation == 2 || im.interpolation == 3 || im.interpolation == 4))
            return fail(PVGPU_E_UNSUPPORTED, "image %zu: interpolation %u", i, im.interpolation);
    }
    if (!s.tri_uv.empty()) {
        if (s.tri_uv.size() != 3 * s.triangles.size() || (s.mesh_uv.size() & 1)) return fail(PVGPU_E_INVALID, "mesh UV table: %zu indices for %zu triangles", s.tri_uv.size(), s.triangles.size());
        for (uint32_t u : s.tri_uv) if ((size_t)u >= s.mesh_uv.size() / 2) return fail(PVGPU_E_INVALID, "mesh UV table: index out of range");
    }
    if (!s.blob_textures.empty()) {
        if (s.blob_textures.size() != s.blob_elements.size()) return fail(PVGPU_E_INVALID, "blob texture table: %zu entries for %zu blob elements", s.blob_textures.size(), s.blob_elements.size());
        for (int32_t t : s.blob_textures) if (t >= (int32_t)s.textures.size()) return fail(PVGPU_E_INVALID, "blob texture table: texture index out of range");
    }
    for (size_t i = 0; i < s.blobs.size(); i++) {
        const pvgpu_blob& b = s.blobs[i];
        if (!range_ok(b.element_first, b.element_count, s.blob_elements.size()) || !range_ok(b.node_first, b.node_count, s.blob_nodes.size()) || b.element_count == 0)
            return fail(PVGPU_E_INVALID, "blob %zu: element / node range out of bounds", i);
        for (uint32_t k = 0; k < b.element_count; k++) {
            const pvgpu_blob_element& e = s.blob_elements[b.element_first + k];
            const bool known = e.type == PVGPU_BLOB_SPHERE || e.type == PVGPU_BLOB_CYLINDER || e.type == PVGPU_BLOB_ELLIPSOID ||
                               e.type == PVGPU_BLOB_BASE_HEMISPHERE || e.type == PVGPU_BLOB_APEX_HEMISPHERE;
            if (!known) return fail(PVGPU_E_UNSUPPORTED, "blob %zu: component type %u", i, e.type);
            if (e.transform >= (int32_t)s.transforms.size() || (e.type != PVGPU_BLOB_SPHERE && e.transform < 0))
                return fail(PVGPU_E_INVALID, "blob %zu: component transform out of range", i);
        }
        for (uint32_t k = 0; k < b.node_count; k++) {
            const pvgpu_blob_node& n = s.blob_nodes[b.node_first + k];
            if (n.count ? !range_ok(n.first, n.count, b.node_count) : n.first >= b.element_count)
                return fail(PVGPU_E_INVALID, "blob %zu: bounding-sphere node %u out of range", i, k);
        }
    }
    // index-list ranges are checked against the table they refer to: children / clipped_by / bounded_by lists -> objects,
    // mesh texture lists -> textures (below), the sky_sphere's pigment list -> pigments (below)
    for (size_t i = 0; i < no; i++) {
        const pvgpu_object& o = s.objects[i];
        for (uint32_t k = 0; k < o.child_count; k++) if (s.index_list[o.child_first + k] >= no) return fail(PVGPU_E_INVALID, "object %zu: child index out of range", i);
        for (uint32_t k = 0; k < o.clip_count; k++) if (s.index_list[o.clip_first + k] >= no) return fail(PVGPU_E_INVALID, "object %zu: clipped_by index out of range", i);
        for (uint32_t k = 0; k < o.bound_count; k++) if (s.index_list[o.bound_first + k] >= no) return fail(PVGPU_E_INVALID, "object %zu: bounded_by index out of range", i);
    }
    for (size_t i = 0; i < s.nodes.size(); i++) {
        const pvgpu_node& n = s.nodes[i];
        if (n.count ? !range_ok(n.first, n.count, s.nodes.size()) : n.first >= no)
            return fail(PVGPU_E_INVALID, "tree node %zu: bad child / object reference", i);
    }
    for (size_t m = 0; m < s.meshes.size(); m++) {
        const pvgpu_mesh& me = s.meshes[m];
        if (!range_ok(me.vertex_first, me.vertex_count, s.vertices.size() / 3) ||
            !range_ok(me.normal_first, me.normal_count, s.normals.size() / 3) ||
            !range_ok(me.triangle_first, me.triangle_count, s.triangles.size()) ||
            !range_ok(me.node_first, me.node_count, s.mesh_nodes.size()) ||
            !range_ok(me.texture_first, me.texture_count, s.index_list.size()))
            return fail(PVGPU_E_INVALID, "mesh %zu: range out of bounds", m);
        for (uint32_t k = 0; k < me.texture_count; k++)
            if (s.index_list[me.texture_first + k] >= s.textures.size()) return fail(PVGPU_E_INVALID, "mesh %zu: texture list entry %u out of range", m, k);
        for (uint32_t t = 0; t < me.triangle_count; t++) {
            const pvgpu_triangle& tr = s.triangles[me.triangle_first + t];
            if (tr.p1 < 0 || tr.p2 < 0 || tr.p3 < 0 || (uint32_t)tr.p1 >= me.vertex_count ||
                (uint32_t)tr.p2 >= me.vertex_count || (uint32_t)tr.p3 >= me.vertex_count ||
                tr.normal_ind < 0 || (uint32_t)tr.normal_ind >= me.normal_count || tr.dominant_axis > 2)
                return fail(PVGPU_E_INVALID, "mesh %zu triangle %u: bad index", m, t);
            if ((tr.flags & PVGPU_TRI_SMOOTH) &&
                (tr.n1 < 0 || tr.n2 < 0 || tr.n3 < 0 || (uint32_t)tr.n1 >= me.normal_count ||
                 (uint32_t)tr.n2 >= me.normal_count || (uint32_t)tr.n3 >= me.normal_count))
                return fail(PVGPU_E_INVALID, "mesh %zu triangle %u: bad normal index", m, t);
            if (tr.flags & PVGPU_TRI_THREETEX)
                return fail(PVGPU_E_UNSUPPORTED, "mesh %zu: per-vertex textures are outside the hot-path scope", m);
            if (tr.texture >= (int32_t)me.texture_count)
                return fail(PVGPU_E_INVALID, "mesh %zu triangle %u: bad texture index", m, t);
        }
        for (uint32_t k = 0; k < me.node_count; k++) {
            const pvgpu_node& n = s.mesh_nodes[me.node_first + k];
            if (n.count ? !range_ok(n.first, n.count, me.node_count) : n.first >= me.triangle_count)
                return fail(PVGPU_E_INVALID, "mesh %zu node %u: bad reference", m, k);
        }
    }
    for (size_t i = 0; i < s.textures.size(); i++) {
        const pvgpu_texture& t = s.textures[i];
        if (t.type != PVGPU_PAT_PLAIN) {
            // texture_map / average texture_map: pattern carrier + a blend map whose entries are texture indices
            if (t.type == PVGPU_PAT_PIGMENT || t.type == PVGPU_PAT_UV_MAP) return fail(PVGPU_E_UNSUPPORTED, "texture %zu: pigment_pattern / uv_mapping as the pattern of a texture_map is outside the hot-path scope", i);
            if (t.type > PVGPU_PAT_LAST || t.pigment < 0 || t.pigment >= (int32_t)s.pigments.size() || t.blend_map < 0 || t.blend_map >= (int32_t)s.blend_maps.size() ||
                !(s.blend_maps[t.blend_map].blend_mode & PVGPU_BLEND_TEXTURE_MAP))
                return fail(PVGPU_E_INVALID, "texture %zu: patterned texture without pattern carrier / texture map", i);
            const pvgpu_blend_map& m = s.blend_maps[t.blend_map];
            for (uint32_t k = 0; k < m.entry_count; k++) {
                const float ti = s.blend_entries[m.entry_first + k].colour[0];
                if (!(ti >= 0.0f) || ti >= (float)s.textures.size() || ti != std::floor(ti))
                    return fail(PVGPU_E_INVALID, "texture %zu: map entry %u is not a texture index", i, k);
            }
            continue;
        }
        if (t.pigment < 0 || t.pigment >= (int32_t)s.pigments.size() || t.finish < 0 ||
            t.finish >= (int32_t)s.finishes.size() || t.next >= (int32_t)s.textures.size())
            return fail(PVGPU_E_INVALID, "texture %zu: bad pigment / finish / next index", i);
        if (t.tnormal >= (int32_t)s.tnormals.size())
            return fail(PVGPU_E_INVALID, "texture %zu: bad tnormal index", i);
        const pvgpu_pigment& tp = s.pigments[t.pigment];
        if (tp.pattern != PVGPU_PAT_PLAIN && tp.pattern != PVGPU_PAT_IMAGE_MAP && tp.pattern != PVGPU_PAT_UV_MAP && (tp.blend_map < 0 || tp.blend_map >= (int32_t)s.blend_maps.size()))
            return fail(PVGPU_E_INVALID, "texture %zu: patterned pigment without blend map", i);
        int depth = 0;
        for (int32_t k = (int32_t)i; k >= 0; k = s.textures[k].next)
            if (++depth > 8) return fail(PVGPU_E_UNSUPPORTED, "texture %zu: more than 8 layers", i);
    }
    for (size_t i = 0; i < s.pigments.size(); i++) {
        const pvgpu_pigment& p = s.pigments[i];
        if (p.pattern < PVGPU_PAT_PLAIN || p.pattern > PVGPU_PAT_LAST)
            return fail(PVGPU_E_UNSUPPORTED, "pigment %zu: pattern %u unsupported", i, p.pattern);
        if (p.blend_map >= (int32_t)s.blend_maps.size())        // (-1 is legal for the pattern carrier of a tnormal)
            return fail(PVGPU_E_INVALID, "pigment %zu: bad blend map index", i);
        if (!range_ok(p.warp_first, p.warp_count, s.warps.size()))
            return fail(PVGPU_E_INVALID, "pigment %zu: warp range out of bounds", i);
        if (p.pattern == PVGPU_PAT_PIGMENT && p.data >= s.pigments.size())
            return fail(PVGPU_E_INVALID, "pigment %zu: pigment_pattern pigment index out of range", i);
        if (p.pattern == PVGPU_PAT_UV_MAP) {
            if (p.data >= s.pigments.size()) return fail(PVGPU_E_INVALID, "pigment %zu: uv_mapping pigment index out of range", i);
            const pvgpu_pigment& q = s.pigments[p.data];
            if (q.pattern != PVGPU_PAT_PLAIN && q.pattern != PVGPU_PAT_IMAGE_MAP && q.pattern != PVGPU_PAT_UV_MAP && q.blend_map < 0)
                return fail(PVGPU_E_INVALID, "pigment %zu: uv_mapping over a patterned pigment without blend map", i);
        }
        if (p.pattern == PVGPU_PAT_IMAGE_MAP && p.data >= s.images.size())
            return fail(PVGPU_E_INVALID, "pigment %zu: image index out of range", i);
        if (p.pattern == PVGPU_PAT_FRACTAL && (!range_ok(p.data, 8, s.shape_data.size()) || !(s.shape_data[p.data] >= 0.0 && s.shape_data[p.data] <= PVGPU_FRACTAL_MAGNET2J) ||
                                               !(s.shape_data[p.data + 1] >= 1.0 && s.shape_data[p.data + 1] <= 1.0e6) ||
                                               (s.shape_data[p.data + 2] == 7.0 && !(s.shape_data[p.data + 4] >= 1.0)) || (s.shape_data[p.data + 2] == 8.0 && !(s.shape_data[p.data + 4] >= 0.0))))
            return fail(PVGPU_E_INVALID, "pigment %zu: fractal pattern record out of range", i);
        if (p.pattern == PVGPU_PAT_CRACKLE && !range_ok(p.data, 9, s.shape_data.size()))
            return fail(PVGPU_E_INVALID, "pigment %zu: crackle parameters outside the shape-data table", i);
    }
    for (size_t i = 0; i < s.tnormals.size(); i++) {
        const pvgpu_tnormal& t = s.tnormals[i];
        if (t.type < PVGPU_NORM_BUMPS || t.type > PVGPU_NORM_AVERAGE)
            return fail(PVGPU_E_UNSUPPORTED, "tnormal %zu: type %u unsupported", i, t.type);
        if (t.normal_map) {
            if (t.normal_map > s.blend_maps.size() || !(s.blend_maps[t.normal_map - 1].blend_mode & PVGPU_BLEND_NORMAL_MAP) ||
                (t.type != PVGPU_NORM_PATTERN && t.type != PVGPU_NORM_AVERAGE))
                return fail(PVGPU_E_INVALID, "tnormal %zu: bad normal_map", i);
            const pvgpu_blend_map& m = s.blend_maps[t.normal_map - 1];
            for (uint32_t k = 0; k < m.entry_count; k++) {
                const float ni = s.blend_entries[m.entry_first + k].colour[0];
                if (!(ni >= 0.0f) || ni >= (float)s.tnormals.size() || ni != std::floor(ni))
                    return fail(PVGPU_E_INVALID, "tnormal %zu: normal_map entry %u is not a tnormal index", i, k);
            }
        } else if (t.type == PVGPU_NORM_AVERAGE) return fail(PVGPU_E_INVALID, "tnormal %zu: average without normal_map", i);
        if (t.pattern < 0 || t.pattern >= (int32_t)s.pigments.size() || !range_ok(t.slope_first, t.slope_count, s.slope_entries.size()))
            return fail(PVGPU_E_INVALID, "tnormal %zu: bad pattern carrier / slope map range", i);
        const pvgpu_pigment& c = s.pigments[t.pattern];
        // (a block pattern without a normal_map is sampled on the pyramid like any other pattern, normal.cpp:880-905)
        if (t.type == PVGPU_NORM_PATTERN && !t.normal_map && c.pattern < PVGPU_PAT_CHECKER)
            return fail(PVGPU_E_INVALID, "tnormal %zu: a pattern normal needs a pattern", i);
        for (uint32_t k = 0; k < c.warp_count; k++)
            if (s.warps[c.warp_first + k].type < PVGPU_WARP_TRANSFORM || s.warps[c.warp_first + k].type > PVGPU_WARP_PLANAR)
                return fail(PVGPU_E_UNSUPPORTED, "tnormal %zu: warp unsupported", i);
    }
    {   // pigment_map nesting: bounded depth, no cycles
        std::vector<int> depth(s.pigments.size(), -1);
        std::function<int(size_t, int)> walk = [&](size_t pi, int level) -> int {
            const pvgpu_pigment& p = s.pigments[pi];
            int worst = 0;
            if (p.pattern == PVGPU_PAT_UV_MAP || p.pattern == PVGPU_PAT_PIGMENT) {           // one level of nesting like a pigment_map entry
                if (level >= 6) return -1;
                const int dch = walk((size_t)p.data, level + 1);
                if (dch < 0) return -1;
                worst = dch + 1;
                if (p.pattern == PVGPU_PAT_UV_MAP) return worst;
            }
            if (p.blend_map < 0 || !(s.blend_maps[p.blend_map].blend_mode & PVGPU_BLEND_PIGMENT_MAP)) return worst;
            if (level >= 6) return -1;          // PV_PIGMENT_MAP_LEVELS of the device code
            const pvgpu_blend_map& m = s.blend_maps[p.blend_map];
            for (uint32_t k = 0; k < m.entry_count; k++) {
                int dch = walk((size_t)s.blend_entries[m.entry_first + k].colour[0], level + 1);
                if (dch < 0) return -1;
                worst = std::max(worst, dch + 1);
            }
            return worst;
        };
        for (size_t i = 0; i < s.pigments.size(); i++)
            if (walk(i, 0) < 0) return fail(PVGPU_E_UNSUPPORTED, "pigment %zu: pigment_map nested deeper than 6 levels", i);
    }
    {   // normal_map nesting: at most 3 levels (PV_NORMAL_MAP_LEVELS of the device code)
        std::function<int(size_t, int)> depth = [&](size_t ni, int level) -> int {
            const pvgpu_tnormal& t = s.tnormals[ni];
            if (!t.normal_map) return 0;
            if (level >= 3) return -1;
            const pvgpu_blend_map& m = s.blend_maps[t.normal_map - 1];
            for (uint32_t k = 0; k < m.entry_count; k++)
                if (depth((size_t)s.blend_entries[m.entry_first + k].colour[0], level + 1) < 0) return -1;
            return 1;
        };
        for (size_t i = 0; i < s.tnormals.size(); i++)
            if (depth(i, 0) < 0) return fail(PVGPU_E_UNSUPPORTED, "tnormal %zu: normal_map nested deeper than 3 levels", i);
    }
    {   // texture_map nesting: bounded depth (the device resolves a hit's texture tree into at most 16 weighted plain textures)
        std::function<int(size_t, int)> leaves = [&](size_t ti, int level) -> int {
            const pvgpu_texture& t = s.textures[ti];
            if (t.type == PVGPU_PAT_PLAIN) return 1;
            if (level >= 4) return -1;
            const pvgpu_blend_map& m = s.blend_maps[t.blend_map];
            int worst = 0, sum = 0;
            for (uint32_t k = 0; k < m.entry_count; k++) {
                int n = leaves((size_t)s.blend_entries[m.entry_first + k].colour[0], level + 1);
                if (n < 0) return -1;
                worst = std::max(worst, n); sum += n;
            }
            return (t.type == PVGPU_PAT_AVERAGE) ? sum : 2 * worst;      // a map blends two neighbouring entries
        };
        for (size_t i = 0; i < s.textures.size(); i++) {
            int n = leaves(i, 0);
            if (n < 0 || n > 16) return fail(PVGPU_E_UNSUPPORTED, "texture %zu: texture_map nested deeper than 4 levels / more than 16 blended textures", i);
        }
    }
    if (s.sky_spheres.size() > 1) return fail(PVGPU_E_INVALID, "more than one sky_sphere");
    for (const pvgpu_sky_sphere& k : s.sky_spheres) {
        if (!range_ok(k.pigment_first, k.pigment_count, s.index_list.size()) || k.transform >= (int32_t)s.transforms.size())
            return fail(PVGPU_E_INVALID, "sky_sphere: bad pigment range / transform");
        for (uint32_t i = 0; i < k.pigment_count; i++) {
            const uint32_t pi = s.index_list[k.pigment_first + i];
            if (pi >= s.pigments.size() || s.pigments[pi].pattern == PVGPU_PAT_UV_MAP || (s.pigments[pi].pattern != PVGPU_PAT_PLAIN && s.pigments[pi].pattern != PVGPU_PAT_IMAGE_MAP && s.pigments[pi].blend_map < 0))
                return fail(PVGPU_E_INVALID, "sky_sphere: bad pigment %u", pi);
        }
    }
    for (size_t i = 0; i < s.fogs.size(); i++) {
        const pvgpu_fog& f = s.fogs[i];
        if (f.type != PVGPU_FOG_CONSTANT && f.type != PVGPU_FOG_GROUND) return fail(PVGPU_E_UNSUPPORTED, "fog %zu: type %u unsupported", i, f.type);
        if (f.turbulence >= (int32_t)s.warps.size() || (f.turbulence >= 0 && s.warps[f.turbulence].type != PVGPU_WARP_TURBULENCE && s.warps[f.turbulence].type != PVGPU_WARP_CLASSIC_TURBULENCE))
            return fail(PVGPU_E_INVALID, "fog %zu: bad turbulence warp", i);
    }
    if (!s.fogs.empty())
        for (size_t i = 0; i < s.lights.size(); i++)
            if ((s.lights[i].flags & PVGPU_LIGHT_MEDIA_ATTEN) && (s.lights[i].flags & PVGPU_LIGHT_MEDIA_INTERACT))
                return fail(PVGPU_E_UNSUPPORTED, "light %zu: media_attenuation with fog (fog on shadow rays) is outside the hot-path scope", i);
    for (size_t i = 0; i < s.finishes.size(); i++) {
        const pvgpu_finish& f = s.finishes[i];
        // (a reflection exponent != 1 is non-linear in the child ray's colour: continuation records, see Cont in pv_common.cuh)
        if (f.irid > 0.0f && s.irid_wavelengths.size() != 3)
            return fail(PVGPU_E_INVALID, "finish %zu: iridescence needs pvgpu_scene_set_irid_wavelengths", i);
        if (f.crand > 0.0f) return fail(PVGPU_E_UNSUPPORTED, "finish %zu: crand is excluded from parity (per-thread RNG)", i);
        if (f.use_subsurface) return fail(PVGPU_E_UNSUPPORTED, "finish %zu: subsurface is outside the hot-path scope", i);
    }
    for (size_t i = 0; i < s.lights.size(); i++) {
        const pvgpu_light& l = s.lights[i];
        if (l.type < PVGPU_LIGHT_POINT || l.type > PVGPU_LIGHT_CYLINDER)
            return fail(PVGPU_E_INVALID, "light %zu: bad type", i);
        if (l.flags & PVGPU_LIGHT_AREA) {
            if (l.area_size1 < 1 || l.area_size2 < 1 || l.area_size1 > 2048 || l.area_size2 > 2048 || (long long)l.area_size1 * l.area_size2 > 4096)
                return fail(PVGPU_E_UNSUPPORTED, "light %zu: area light grid %d x %d (supported: up to 4096 samples)", i, l.area_size1, l.area_size2);
            if (l.flags & PVGPU_LIGHT_JITTER)
                return fail(PVGPU_E_UNSUPPORTED, "light %zu: area light jitter is excluded from parity (per-thread RNG)", i);
            if (l.flags & PVGPU_LIGHT_FULL_AREA)
                return fail(PVGPU_E_UNSUPPORTED, "light %zu: area_illumination is outside the hot-path scope", i);
            if ((l.flags & PVGPU_LIGHT_CIRCULAR) && (l.area_size1 < 2 || l.area_size2 < 2))
                return fail(PVGPU_E_INVALID, "light %zu: circular area light needs at least 2 x 2 samples", i);
        }
        if (l.projected_through >= 0)
            return fail(PVGPU_E_UNSUPPORTED, "light %zu: projected_through is outside the hot-path scope", i);
        if (l.flags & PVGPU_LIGHT_GROUP)
            return fail(PVGPU_E_UNSUPPORTED, "light %zu: light groups are outside the hot-path scope", i);
    }
    for (size_t i = 0; i < s.interiors.size(); i++)
        if (std::fabs((double)s.interiors[i].dispersion - 1.0) >= 1e-10)
            return fail(PVGPU_E_UNSUPPORTED, "interior %zu: dispersion is outside the hot-path scope", i);
    if (std::fabs((double)s.globals.atmosphere_dispersion - 1.0) >= 1e-10)
        return fail(PVGPU_E_UNSUPPORTED, "atmosphere dispersion is outside the hot-path scope");
    if (!s.have_camera) return fail(PVGPU_E_INVALID, "no camera set");
    if (s.camera.type < PVGPU_CAMERA_PERSPECTIVE || s.camera.type > PVGPU_CAMERA_SPHERICAL)
        return fail(PVGPU_E_UNSUPPORTED, "camera type %u is outside the hot-path scope", s.camera.type);
    if (s.camera.reserved) {
        if (s.camera.reserved > s.tnormals.size()) return fail(PVGPU_E_INVALID, "camera: normal index out of range");
        if (s.camera.type > PVGPU_CAMERA_ORTHOGRAPHIC) return fail(PVGPU_E_UNSUPPORTED, "camera: normal perturbation on camera type %u is outside the hot-path scope", s.camera.type);
    }
    if (s.camera.type > PVGPU_CAMERA_ORTHOGRAPHIC && s.camera_ext.size() != 3)
        return fail(PVGPU_E_INVALID, "camera type %u needs pvgpu_scene_set_camera_angles", s.camera.type);
    if (s.globals.bounding_method == 1 && s.nodes.empty())
        return fail(PVGPU_E_INVALID, "bounding_method 1 without a tree (call pvgpu_scene_set_tree or pvgpu_scene_build_tree)");

    // shadow rays may stop at the first opaque blocker only when every shadow caster is opaque;
    // otherwise the reference's closest-hit filter loop is followed (trace.cpp:2025-2075)
    s.all_shadow_casters_opaque = true;
    for (const pvgpu_object& o : s.objects)
        if (!PVGPU_IS_CSG(o.type) && !(o.flags & PVGPU_NO_SHADOW_FLAG) && !(o.flags & PVGPU_OPAQUE_FLAG))
            s.all_shadow_casters_opaque = false;
    return PVGPU_OK;
}

}  // namespace pvgpu

static const char kMagic[8] = { 'P', 'V', 'G', 'P', 'U', 'S', 'C', '1' };

template <class T> static bool put(FILE* f, const std::vector<T>& v)
{
    uint64_t n = v.size();
    return fwrite(&n, sizeof n, 1, f) == 1 && (n == 0 || fwrite(v.data(), sizeof(T), n, f) == n);
}
template <class T> static bool get(FILE* f, std::vector<T>& v)
{
    uint64_t n = 0;
    if (fread(&n, sizeof n, 1, f) != 1 || n > (1ull << 34) / sizeof(T)) return false;
    v.resize(n);
    return n == 0 || fread(v.data(), sizeof(T), n, f) == n;
}


// ================================================================================================
// C ABI
// ================================================================================================
using namespace pvgpu;

extern "C" {

int pvgpu_abi_version(void) { return PVGPU_ABI_VERSION; }
const char* pvgpu_last_error(void) { return g_err; }

int pvgpu_scene_create(pvgpu_scene** out, const pvgpu_globals* g)
{
    clear_error();
    if (!out || !g) return fail(PVGPU_E_INVALID, "pvgpu_scene_create: null argument");
    Scene* s = new Scene();
    s->globals = *g;
    *out = reinterpret_cast<pvgpu_scene*>(s);
    return PVGPU_OK;
}

void pvgpu_scene_destroy(pvgpu_scene* sc)
{
    if (!sc) return;
    Scene* s = reinterpret_cast<Scene*>(sc);
    device_release(*s);
    delete s;
}

#define SCENE_OR_FAIL(sc) \
    clear_error(); \
    if (!(sc)) return fail(PVGPU_E_INVALID, "%s: null scene", __func__); \
    Scene& s = *reinterpret_cast<Scene*>(sc); \
    if (s.dev) return fail(PVGPU_E_INVALID, "%s: scene already finalized", __func__)

int pvgpu_scene_set_objects(pvgpu_scene* sc, const pvgpu_object* objs, size_t n_objs,
                            const uint32_t* index_list, size_t n_index,
                            const uint32_t* frame, size_t n_frame)
{
    SCENE_OR_FAIL(sc);
    if ((!objs && n_objs) || (!index_list && n_index) || (!frame && n_frame))
        return fail(PVGPU_E_INVALID, "pvgpu_scene_set_objects: null array");
    s.objects.assign(objs, objs + n_objs);
    s.index_list.assign(index_list, index_list + n_index);
    s.frame.assign(frame, frame + n_frame);
    return PVGPU_OK;
}

int pvgpu_scene_set_transforms(pvgpu_scene* sc, const pvgpu_transform* t, size_t n)
{
    SCENE_OR_FAIL(sc);
    if (!t && n) return fail(PVGPU_E_INVALID, "pvgpu_scene_set_transforms: null array");
    s.transforms.assign(t, t + n);
    return PVGPU_OK;
}

int pvgpu_scene_set_tree(pvgpu_scene* sc, const pvgpu_node* nodes, size_t n)
{
    SCENE_OR_FAIL(sc);
    if (!nodes && n) return fail(PVGPU_E_INVALID, "pvgpu_scene_set_tree: null array");
    s.nodes.assign(nodes, nodes + n);
    return PVGPU_OK;
}

int pvgpu_scene_build_tree(pvgpu_scene* sc)
{
    SCENE_OR_FAIL(sc);
    // Build_Bounding_Slabs (boundingbox.cpp:325-430): frame-level objects split by INFINITE_FLAG
    std::vector<LeafBox> fin, inf;
    for (uint32_t f : s.frame) {
        if (f >= s.objects.size()) return fail(PVGPU_E_INVALID, "frame object index out of range");
        const pvgpu_object& o = s.objects[f];
        LeafBox b;
        std::memcpy(b.lo, o.bbox, sizeof b.lo);
        std::memcpy(b.size, o.bbox + 3, sizeof b.size);
        b.payload = f;
        ((o.flags & PVGPU_INFINITE_FLAG) ? inf : fin).push_back(b);
    }
    build_bbox_tree(fin, inf, s.nodes);
    return PVGPU_OK;
}

int pvgpu_scene_set_blobs(pvgpu_scene* sc, const pvgpu_blob* blobs, size_t n_blobs, const pvgpu_blob_element* elements, size_t n_elements,
                          const pvgpu_blob_node* nodes, size_t n_nodes)
{
    clear_error();
    if (!sc || (!blobs && n_blobs) || (!elements && n_elements) || (!nodes && n_nodes))
        return fail(PVGPU_E_INVALID, "pvgpu_scene_set_blobs: null array");
    Scene& s = *reinterpret_cast<Scene*>(sc);
    s.blobs.assign(blobs, blobs + n_blobs);
    s.blob_elements.assign(elements, elements + n_elements);
    s.blob_nodes.assign(nodes, nodes + n_nodes);
    return PVGPU_OK;
}

int pvgpu_scene_set_images(pvgpu_scene* sc, const pvgpu_image* images, size_t n_images, const float* texels, size_t n_texel_floats)
{
    SCENE_OR_FAIL(sc);
    if ((!images && n_images) || (!texels && n_texel_floats)) return fail(PVGPU_E_INVALID, "pvgpu_scene_set_images: null array");
    s.images.assign(images, images + n_images);
    s.texels.assign(texels, texels + n_texel_floats);
    return PVGPU_OK;
}

int pvgpu_scene_set_mesh_uv(pvgpu_scene* sc, const double* uv, size_t n_uv, const uint32_t* tri_uv, size_t n_tri)
{
    SCENE_OR_FAIL(sc);
    if ((!uv && n_uv) || (!tri_uv && n_tri)) return fail(PVGPU_E_INVALID, "pvgpu_scene_set_mesh_uv: null array");
    s.mesh_uv.assign(uv, uv + 2 * n_uv);
    s.tri_uv.assign(tri_uv, tri_uv + 3 * n_tri);
    return PVGPU_OK;
}

int pvgpu_scene_set_blob_textures(pvgpu_scene* sc, const int32_t* textures, size_t n)
{
    SCENE_OR_FAIL(sc);
    if (!textures && n) return fail(PVGPU_E_INVALID, "pvgpu_scene_set_blob_textures: null array");
    s.blob_textures.assign(textures, textures + n);
    return PVGPU_OK;
}

int pvgpu_scene_set_normals(pvgpu_scene* sc, const pvgpu_tnormal* tn, size_t n_tn, const pvgpu_slope_entry* slopes, size_t n_slopes)
{
    SCENE_OR_FAIL(sc);
    if ((!tn && n_tn) || (!slopes && n_slopes)) return fail(PVGPU_E_INVALID, "pvgpu_scene_set_normals: null array");
    s.tnormals.assign(tn, tn + n_tn);
    s.slope_entries.assign(slopes, slopes + n_slopes);
    return PVGPU_OK;
}

int pvgpu_scene_set_camera_angles(pvgpu_scene* sc, double angle, double h_angle, double v_angle)
{
    SCENE_OR_FAIL(sc);
    s.camera_ext = { angle, h_angle, v_angle };
    if (s.dev) return fail(PVGPU_E_INVALID, "pvgpu_scene_set_camera_angles: call before pvgpu_scene_finalize");
    return PVGPU_OK;
}

int pvgpu_scene_set_irid_wavelengths(pvgpu_scene* sc, const float wavelengths[3])
{
    SCENE_OR_FAIL(sc);
    if (!wavelengths) return fail(PVGPU_E_INVALID, "pvgpu_scene_set_irid_wavelengths: null argument");
    s.irid_wavelengths.assign(wavelengths, wavelengths + 3);
    return PVGPU_OK;
}

int pvgpu_scene_set_atmosphere(pvgpu_scene* sc, const pvgpu_sky_sphere* sky, const pvgpu_fog* fogs, size_t n_fogs)
{
    SCENE_OR_FAIL(sc);
    if (!fogs && n_fogs) return fail(PVGPU_E_INVALID, "pvgpu_scene_set_atmosphere: null array");
    s.sky_spheres.clear();
    if (sky) s.sky_spheres.push_back(*sky);
    s.fogs.assign(fogs, fogs + n_fogs);
    return PVGPU_OK;
}

int pvgpu_scene_set_shape_data(pvgpu_scene* sc, const double* data, size_t n)
{
    SCENE_OR_FAIL(sc);
    if (!data && n) return fail(PVGPU_E_INVALID, "pvgpu_scene_set_shape_data: null array");
    s.shape_data.assign(data, data + n);
    return PVGPU_OK;
}

int pvgpu_scene_set_meshes(pvgpu_scene* sc, const pvgpu_mesh* meshes, size_t n_meshes,
                           const float* vertices, size_t n_vertices,
                           const float* normals, size_t n_normals,
                           const pvgpu_triangle* tris, size_t n_tris,
                           const pvgpu_node* nodes, size_t n_nodes)
{
    SCENE_OR_FAIL(sc);
    if ((!meshes && n_meshes) || (!vertices && n_vertices) || (!normals && n_normals) || (!tris && n_tris) ||
        (!nodes && n_nodes))
        return fail(PVGPU_E_INVALID, "pvgpu_scene_set_meshes: null array");
    s.meshes.assign(meshes, meshes + n_meshes);
    s.vertices.assign(vertices, vertices + 3 * n_vertices);
    s.normals.assign(normals, normals + 3 * n_normals);
    s.triangles.assign(tris, tris + n_tris);
    s.mesh_nodes.assign(nodes, nodes + n_nodes);
    return PVGPU_OK;
}

int pvgpu_scene_set_lights(pvgpu_scene* sc, const pvgpu_light* l, size_t n)
{
    SCENE_OR_FAIL(sc);
    if (!l && n) return fail(PVGPU_E_INVALID, "pvgpu_scene_set_lights: null array");
    s.lights.assign(l, l + n);
    return PVGPU_OK;
}

int pvgpu_scene_set_materials(pvgpu_scene* sc,
                              const pvgpu_texture* tex, size_t n_tex,
                              const pvgpu_pigment* pig, size_t n_pig,
                              const pvgpu_finish* fin, size_t n_fin,
                              const pvgpu_blend_map* maps, size_t n_maps,
                              const pvgpu_blend_entry* entries, size_t n_entries,
                              const pvgpu_warp* warps, size_t n_warps,
                              const pvgpu_interior* interiors, size_t n_interiors)
{
    SCENE_OR_FAIL(sc);
    if ((!tex && n_tex) || (!pig && n_pig) || (!fin && n_fin) || (!maps && n_maps) || (!entries && n_entries) ||
        (!warps && n_warps) || (!interiors && n_interiors))
        return fail(PVGPU_E_INVALID, "pvgpu_scene_set_materials: null array");
    s.textures.assign(tex, tex + n_tex);
    s.pigments.assign(pig, pig + n_pig);
    s.finishes.assign(fin, fin + n_fin);
    s.blend_maps.assign(maps, maps + n_maps);
    s.blend_entries.assign(entries, entries + n_entries);
    s.warps.assign(warps, warps + n_warps);
    s.interiors.assign(interiors, interiors + n_interiors);
    return PVGPU_OK;
}

int pvgpu_scene_set_camera(pvgpu_scene* sc, const pvgpu_camera* cam)
{
    clear_error();
    if (!sc || !cam) return fail(PVGPU_E_INVALID, "pvgpu_scene_set_camera: null argument");
    Scene& s = *reinterpret_cast<Scene*>(sc);    // the camera may change between frames of a finalized scene
    s.camera = *cam;
    s.have_camera = true;
    return PVGPU_OK;
}

int pvgpu_scene_get_camera(const pvgpu_scene* sc, pvgpu_camera* cam)
{
    clear_error();
    if (!sc || !cam) return fail(PVGPU_E_INVALID, "pvgpu_scene_get_camera: null argument");
    const Scene& s = *reinterpret_cast<const Scene*>(sc);
    if (!s.have_camera) return fail(PVGPU_E_INVALID, "no camera set");
    *cam = s.camera;
    return PVGPU_OK;
}

// ------------------------------------------------------------------------------------------------
// mesh2 helper: what Parser::Parse_Mesh2 (parser.cpp:4079-4620) + Mesh::Compute_Mesh_Triangle
// (mesh.cpp:838-954) + Mesh::Build_Mesh_BBox_Tree (mesh.cpp:1376-1413) leave in MESH_DATA for a
// mesh2 { vertex_vectors, face_indices } without normals / uv / textures (flat triangles).
// ------------------------------------------------------------------------------------------------
static inline int max3_coordinate(double x, double y, double z)   // mesh.cpp:94
{
    return (x > y) ? ((x > z) ? 0 : 2) : ((y > z) ? 1 : 2);
}

int pvgpu_scene_add_mesh2(pvgpu_scene* sc, const double* vertices, size_t n_vertices,
                          const int32_t* indices, size_t n_faces, int32_t* out_mesh)
{
    SCENE_OR_FAIL(sc);
    if (!vertices || !indices || !n_vertices || !n_faces)
        return fail(PVGPU_E_INVALID, "pvgpu_scene_add_mesh2: empty mesh");
    pvgpu_mesh me{};
    me.vertex_first = (uint32_t)(s.vertices.size() / 3);
    me.vertex_count = (uint32_t)n_vertices;
    me.normal_first = (uint32_t)(s.normals.size() / 3);
    me.normal_count = (uint32_t)n_faces;
    me.triangle_first = (uint32_t)s.triangles.size();
    me.triangle_count = (uint32_t)n_faces;
    me.node_first = (uint32_t)s.mesh_nodes.size();

    // vertices are stored as MeshVector (FP32); every later computation starts from the rounded values
    const size_t v0 = s.vertices.size();
    s.vertices.resize(v0 + 3 * n_vertices);
    for (size_t i = 0; i < 3 * n_vertices; i++) s.vertices[v0 + i] = (float)vertices[i];
    const float* V = s.vertices.data() + v0;

    const size_t n0 = s.normals.size();
    s.normals.resize(n0 + 3 * n_faces);
    const size_t t0 = s.triangles.size();
    s.triangles.resize(t0 + n_faces);
    std::vector<LeafBox> leaves(n_faces);

    for (size_t f = 0; f < n_faces; f++) {
        int32_t a = indices[3 * f], b = indices[3 * f + 1], c = indices[3 * f + 2];
        if (a < 0 || b < 0 || c < 0 || (size_t)a >= n_vertices || (size_t)b >= n_vertices || (size_t)c >= n_vertices)
            return fail(PVGPU_E_INVALID, "mesh face %zu: index out of range", f);
        pvgpu_triangle tr{};
        tr.p1 = a; tr.p2 = b; tr.p3 = c;
        tr.n1 = tr.n2 = tr.n3 = -1;
        tr.texture = tr.texture2 = tr.texture3 = -1;
        double P1[3], P2[3], P3[3];
        for (int k = 0; k < 3; k++) { P1[k] = V[3 * a + k]; P2[k] = V[3 * b + k]; P3[k] = V[3 * c + k]; }
        double V1[3], V2[3], N[3];
        for (int k = 0; k < 3; k++) { V1[k] = P2[k] - P1[k]; V2[k] = P3[k] - P1[k]; }
        N[0] = V2[1] * V1[2] - V2[2] * V1[1];       // cross(V2, V1)
        N[1] = V2[2] * V1[0] - V2[0] * V1[2];
        N[2] = V2[0] * V1[1] - V2[1] * V1[0];
        double len = std::sqrt(N[0] * N[0] + N[1] * N[1] + N[2] * N[2]);
        if (len != 0.0) {
            for (int k = 0; k < 3; k++) N[k] /= len;
            double dist = N[0] * P1[0] + N[1] * P1[1] + N[2] * P1[2];
            dist *= -1.0;
            tr.distance = (float)dist;
            int dom = max3_coordinate(std::fabs(N[0]), std::fabs(N[1]), std::fabs(N[2]));
            tr.dominant_axis = (uint8_t)dom;
            const int ua = (dom == 0) ? 1 : 0, va = (dom == 2) ? 1 : 2;     // X:(Y,Z) Y:(X,Z) Z:(X,Y)
            bool swap = (P2[ua] - P3[ua]) * (P2[va] - P1[va]) < (P2[va] - P3[va]) * (P2[ua] - P1[ua]);
            const double* q1 = P1; const double* q2 = P2;
            if (swap) { std::swap(tr.p1, tr.p2); q1 = P2; q2 = P1; }
            // compute_smooth_triangle (mesh.cpp:984-1010): Perp / vAxis are filled for flat triangles too
            double d32[3] = { P3[0] - q2[0], P3[1] - q2[1], P3[2] - q2[2] };
            tr.v_axis = (uint8_t)max3_coordinate(std::fabs(d32[0]), std::fabs(d32[1]), std::fabs(d32[2]));
            double t1[3] = { q2[0] - P3[0], q2[1] - P3[1], q2[2] - P3[2] };
            double l1 = std::sqrt(t1[0] * t1[0] + t1[1] * t1[1] + t1[2] * t1[2]);
            if (l1 != 0) for (int k = 0; k < 3; k++) t1[k] /= l1;
            double t2[3] = { q1[0] - P3[0], q1[1] - P3[1], q1[2] - P3[2] };
            double proj = t2[0] * t1[0] + t2[1] * t1[1] + t2[2] * t1[2];
            for (int k = 0; k < 3; k++) t1[k] *= proj;
            double pp[3] = { t1[0] - t2[0], t1[1] - t2[1], t1[2] - t2[2] };
            double lp = std::sqrt(pp[0] * pp[0] + pp[1] * pp[1] + pp[2] * pp[2]);
            if (lp != 0) for (int k = 0; k < 3; k++) pp[k] /= lp;
            float perp[3] = { (float)pp[0], (float)pp[1], (float)pp[2] };
            double uden = -(t2[0] * (double)perp[0] + t2[1] * (double)perp[1] + t2[2] * (double)perp[2]);
            for (int k = 0; k < 3; k++) tr.perp[k] = perp[k] / (float)uden;      // MeshVector /= DBL divides in FP32
        }
        tr.normal_ind = (int32_t)f;
        for (int k = 0; k < 3; k++) s.normals[n0 + 3 * f + k] = (float)N[k];
        s.triangles[t0 + f] = tr;
        // get_triangle_bbox (mesh.cpp:1335): double min/max of the FP32 vertices
        LeafBox& lb = leaves[f];
        for (int k = 0; k < 3; k++) {
            double mn = std::min(P1[k], std::min(P2[k], P3[k]));
            double mx = std::max(P1[k], std::max(P2[k], P3[k]));
            lb.lo[k] = (float)mn;
            lb.size[k] = (float)(mx - mn);
        }
        lb.payload = (uint32_t)f;
    }
    std::vector<pvgpu_node> tree;
    build_bbox_tree(leaves, {}, tree);
    me.node_count = (uint32_t)tree.size();
    s.mesh_nodes.insert(s.mesh_nodes.end(), tree.begin(), tree.end());
    s.meshes.push_back(me);
    if (out_mesh) *out_mesh = (int32_t)s.meshes.size() - 1;
    return PVGPU_OK;
}

int pvgpu_scene_finalize(pvgpu_scene* sc, int device)
{
    SCENE_OR_FAIL(sc);
    int rc = validate_scene(s);
    if (rc != PVGPU_OK) return rc;
    return device_upload(s, &device, 1);
}

int pvgpu_scene_finalize_multi(pvgpu_scene* sc, const int* devices, int n_devices)
{
    SCENE_OR_FAIL(sc);
    int rc = validate_scene(s);
    if (rc != PVGPU_OK) return rc;
    return device_upload(s, devices, n_devices);
}

int pvgpu_scene_device_count(const pvgpu_scene* sc)
{
    return sc ? (int)reinterpret_cast<const Scene*>(sc)->devs.size() : 0;
}

size_t pvgpu_scene_device_bytes(const pvgpu_scene* sc)
{
    return sc ? reinterpret_cast<const Scene*>(sc)->device_bytes : 0;
}

// ------------------------------------------------------------------------------------------------
// flat-scene file: magic, ABI version, then every table as (u64 count, raw records)
// ------------------------------------------------------------------------------------------------
int pvgpu_scene_save(const pvgpu_scene* sc, const char* path)
{
    clear_error();
    if (!sc || !path) return fail(PVGPU_E_INVALID, "pvgpu_scene_save: null argument");
    const Scene& s = *reinterpret_cast<const Scene*>(sc);
    FILE* f = fopen(path, "wb");
    if (!f) return fail(PVGPU_E_IO, "cannot open %s for writing", path);
    uint32_t ver = PVGPU_FILE_VERSION, have_cam = s.have_camera;
    bool ok = fwrite(kMagic, 8, 1, f) == 1 && fwrite(&ver, 4, 1, f) == 1 && fwrite(&have_cam, 4, 1, f) == 1 &&
              fwrite(&s.globals, sizeof s.globals, 1, f) == 1 && fwrite(&s.camera, sizeof s.camera, 1, f) == 1 &&
              put(f, s.objects) && put(f, s.index_list) && put(f, s.frame) && put(f, s.transforms) &&
              put(f, s.nodes) && put(f, s.meshes) && put(f, s.vertices) && put(f, s.normals) &&
              put(f, s.triangles) && put(f, s.mesh_nodes) && put(f, s.lights) && put(f, s.textures) &&
              put(f, s.pigments) && put(f, s.finishes) && put(f, s.blend_maps) && put(f, s.blend_entries) &&
              put(f, s.warps) && put(f, s.interiors);
    // optional trailing sections in fixed order; a section is written when it or a later one holds data
    const bool sec9 = !s.tri_uv.empty();
    const bool sec8 = sec9 || !s.images.empty();
    const bool sec7 = sec8 || !s.blob_textures.empty();
    const bool sec6 = sec7 || !s.irid_wavelengths.empty();
    const bool sec5 = sec6 || !s.camera_ext.empty();
    const bool sec4 = sec5 || !s.sky_spheres.empty() || !s.fogs.empty();
    const bool sec3 = sec4 || !s.tnormals.empty();
    const bool sec2 = sec3 || !s.shape_data.empty();
    const bool sec1 = sec2 || !s.blobs.empty();
    if (ok && sec1) ok = put(f, s.blobs) && put(f, s.blob_elements) && put(f, s.blob_nodes);
    if (ok && sec2) ok = put(f, s.shape_data);
    if (ok && sec3) ok = put(f, s.tnormals) && put(f, s.slope_entries);
    if (ok && sec4) ok = put(f, s.sky_spheres) && put(f, s.fogs);
    if (ok && sec5) ok = put(f, s.camera_ext);
    if (ok && sec6) ok = put(f, s.irid_wavelengths);
    if (ok && sec7) ok = put(f, s.blob_textures);
    if (ok && sec8) ok = put(f, s.images) && put(f, s.texels);
    if (ok && sec9) ok = put(f, s.mesh_uv) && put(f, s.tri_uv);
    ok = (fclose(f) == 0) && ok;
    return ok ? PVGPU_OK : fail(PVGPU_E_IO, "short write to %s", path);
}

int pvgpu_scene_load(pvgpu_scene** out, const char* path)
{
    clear_error();
    if (!out || !path) return fail(PVGPU_E_INVALID, "pvgpu_scene_load: null argument");
    FILE* f = fopen(path, "rb");
    if (!f) return fail(PVGPU_E_IO, "cannot open %s", path);
    Scene* s = new Scene();
    char magic[8];
    uint32_t ver = 0, have_cam = 0;
    bool ok = fread(magic, 8, 1, f) == 1 && memcmp(magic, kMagic, 8) == 0 && fread(&ver, 4, 1, f) == 1 &&
              ver == PVGPU_FILE_VERSION && fread(&have_cam, 4, 1, f) == 1 &&
              fread(&s->globals, sizeof s->globals, 1, f) == 1 && fread(&s->camera, sizeof s->camera, 1, f) == 1 &&
              get(f, s->objects) && get(f, s->index_list) && get(f, s->frame) && get(f, s->transforms) &&
              get(f, s->nodes) && get(f, s->meshes) && get(f, s->vertices) && get(f, s->normals) &&
              get(f, s->triangles) && get(f, s->mesh_nodes) && get(f, s->lights) && get(f, s->textures) &&
              get(f, s->pigments) && get(f, s->finishes) && get(f, s->blend_maps) && get(f, s->blend_entries) &&
              get(f, s->warps) && get(f, s->interiors);
    if (ok) {
        const int c = fgetc(f);
        if (c != EOF) { ungetc(c, f); ok = get(f, s->blobs) && get(f, s->blob_elements) && get(f, s->blob_nodes); }
    }
    if (ok) {
        const int c = fgetc(f);
        if (c != EOF) { ungetc(c, f); ok = get(f, s->shape_data); }
    }
    if (ok) {
        const int c = fgetc(f);
        if (c != EOF) { ungetc(c, f); ok = get(f, s->tnormals) && get(f, s->slope_entries); }
    }
    if (ok) {
        const int c = fgetc(f);
        if (c != EOF) { ungetc(c, f); ok = get(f, s->sky_spheres) && get(f, s->fogs); }
    }
    if (ok) {
        const int c = fgetc(f);
        if (c != EOF) { ungetc(c, f); ok = get(f, s->camera_ext); }
    }
    if (ok) {
        const int c = fgetc(f);
        if (c != EOF) { ungetc(c, f); ok = get(f, s->irid_wavelengths); }
    }
    if (ok) {
        const int c = fgetc(f);
        if (c != EOF) { ungetc(c, f); ok = get(f, s->blob_textures); }
    }
    if (ok) {
        const int c = fgetc(f);
        if (c != EOF) { ungetc(c, f); ok = get(f, s->images) && get(f, s->texels); }
    }
    if (ok) {
        const int c = fgetc(f);
        if (c != EOF) { ungetc(c, f); ok = get(f, s->mesh_uv) && get(f, s->tri_uv); }
    }
    fclose(f);
    if (!ok) { delete s; return fail(PVGPU_E_IO, "%s is not a pvgpu scene file of version %d", path, PVGPU_FILE_VERSION); }
    s->have_camera = have_cam != 0;
    *out = reinterpret_cast<pvgpu_scene*>(s);
    return PVGPU_OK;
}

}  // extern "C"
