// Reference-side binding of the pvgpu trace path: a replacement for POV-Ray's
// source/backend/render/tracetask.cpp.  It is compiled AGAINST the reference headers (never copied
// into them) and linked with the reference's other objects instead of the stock tracetask.o, giving a
// `povray-gpu` binary whose parser, SDL, INI/CLI options, RenderBackend / View / vfe session code are
// the unmodified reference and whose TraceTask::Run() hands the tiles to the GPU:
//
//   View::StartRender (view.cpp:1186) -> TraceTask::Run()
//        flatten SceneData -> pvgpu_scene_*          (once per view, first task to arrive)
//        ViewData::GetNextRectangle (view.cpp:236)   drain the tile queue
//        pvgpu_render                                 one call for all drained tiles
//        ViewData::CompletedRectangle (view.cpp:405)  same rect / serial / row-major RGBT pixels
//
// Environment switches (used by the parity tests; all optional):
//   PVGPU_RENDER=gpu|stock   who computes the delivered pixels (default gpu).  `stock` runs the reference's
//                            own TracePixel per pixel centre exactly like SimpleSamplingM0 (tracetask.cpp:403).
//   PVGPU_DUMP_SCENE=<file>  write the flattened scene (pvgpu_scene_save)
//   PVGPU_DUMP_RAYS=<file>   per pixel: stock camera ray + stock Trace::FindIntersection result
//   PVGPU_DUMP_RGBT=<file>   per pixel float RGBT from the stock TracePixel
//   PVGPU_DUMP_GPU=<file>    per pixel float RGBT from pvgpu_render
//   PVGPU_PROBE_IN=<file> PVGPU_PROBE_OUT=<file>   known-answer vectors: the reference's own Solve_Polynomial / Noise / DNoise /
//                            Turbulence on the inputs of <in> (format: tests/golden/make_golden_probe.py), written once
// Dumps cover the whole render area and are written when the tile queue is empty (run with +WT1).
//
// Build recipe: oracle/Makefile (target adapter).  See INTEGRATION.md.

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <typeinfo>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

// The flattening needs a few members the reference keeps private/protected (Sphere::Do_Ellipsoid,
// SpindleTorus::mSpindleMode, TracePixel::CreateCameraRay).  A maintainer would add accessors or a
// friend declaration; the out-of-tree adapter opens the classes up instead.
#define private public
#define protected public
#include "backend/render/tracetask.h"
#include "backend/scene/backendscenedata.h"
#include "backend/scene/view.h"
#include "backend/scene/viewthreaddata.h"
#include "core/bounding/boundingbox.h"
#include "core/lighting/lightsource.h"
#include "core/material/blendmap.h"
#include "core/material/interior.h"
#include "core/support/imageutil.h"
#include "base/image/image.h"
#include "core/material/pattern.h"
#include "core/material/normal.h"
#include "core/material/pigment.h"
#include "core/material/texture.h"
#include "core/material/warp.h"
#include "core/math/matrix.h"
#include "core/math/polynomialsolver.h"
#include "core/material/noise.h"
#include "core/render/ray.h"
#include "core/scene/atmosphere.h"
#include "core/scene/object.h"
#include "core/scene/scenedata.h"
#include "core/scene/tracethreaddata.h"
#include "core/bounding/boundingsphere.h"
#include "core/shape/blob.h"
#include "core/shape/box.h"
#include "core/shape/cone.h"
#include "core/shape/disc.h"
#include "core/shape/csg.h"
#include "core/shape/mesh.h"
#include "core/shape/plane.h"
#include "core/shape/polygon.h"
#include "core/shape/polynomial.h"
#include "core/shape/prism.h"
#include "core/shape/quadric.h"
#include "core/shape/sphere.h"
#include "core/shape/superellipsoid.h"
#include "core/shape/torus.h"
#include "core/shape/triangle.h"
#include "core/shape/truetype.h"
#include "core/support/statistics.h"
#undef private
#undef protected

#include "pvgpu.h"

// this must be the last file included
#include "base/image/colourspace.h"
#include "base/povdebug.h"

namespace pov
{

using std::vector;

namespace
{

// ------------------------------------------------------------------------------------------------
// SceneData -> pvgpu tables
// ------------------------------------------------------------------------------------------------

// The outline of a glyph is private to truetype.cpp (GlyphHeader / Contour / GlyphStruct, truetype.cpp:231-271; TrueType::glyph is an
// opaque pointer in truetype.h).  These mirrors repeat the members in order, so that the flattener can read the contours of the glyphs
// the parser built; they are only valid next to reference objects built by the same compiler with the same flags (oracle/build_ref.sh).
namespace ttf_mirror {
struct GlyphHeader { POV_INT16 numContours, xMin, yMin, xMax, yMax; };
struct Contour { POV_UINT8 inside_flag; POV_UINT16 count; std::vector<POV_UINT8> flags; std::vector<DBL> x, y; };
struct GlyphStruct { GlyphHeader header; POV_UINT32 glyph_index; Contour* contours; POV_UINT16 unitsPerEm; POV_UINT32 myMetrics; };
}

struct Flattener
{
    vector<pvgpu_object> objects;
    vector<uint32_t> index_list, frame;
    vector<pvgpu_transform> transforms;
    vector<pvgpu_node> nodes, mesh_nodes;
    vector<pvgpu_mesh> meshes;
    vector<pvgpu_blob> blobs;
    vector<pvgpu_blob_element> blob_elements;
    vector<pvgpu_blob_node> blob_nodes;
    vector<float> vertices, normals;
    vector<double> shape_data;
    vector<pvgpu_triangle> triangles;
    vector<pvgpu_light> lights;
    vector<pvgpu_texture> textures;
    vector<pvgpu_pigment> pigments;
    vector<pvgpu_finish> finishes;
    vector<pvgpu_blend_map> maps;
    vector<pvgpu_blend_entry> entries;
    vector<pvgpu_warp> warps;
    vector<pvgpu_interior> interiors;
    vector<pvgpu_tnormal> tnormals;
    vector<pvgpu_slope_entry> slope_entries;
    std::map<const void*, int32_t> object_ids, texture_ids, interior_ids;
    std::map<const void*, uint32_t> mesh_tri_first;          // Mesh object -> first triangle of its copy in the triangle table (ray dumps)
    std::vector<int32_t> blob_textures;                      // per blob element (pvgpu_scene_set_blob_textures)
    std::vector<double> mesh_uv;                             // (u, v) pairs + three indices per triangle (pvgpu_scene_set_mesh_uv)
    std::vector<uint32_t> tri_uv;
    bool any_mesh_uv = false;
    std::vector<pvgpu_image> images;
    std::vector<float> texels;
    std::map<const void*, int32_t> image_ids;
    bool any_blob_texture = false;
    std::string error;

    void unsupported(const std::string& what) { if (error.empty()) error = what; }

    int32_t add_transform(const TRANSFORM* t)
    {
        if (t == nullptr) return -1;
        pvgpu_transform x;
        for (int r = 0; r < 4; r++)
            for (int c = 0; c < 4; c++) { x.matrix[4 * r + c] = t->matrix[r][c]; x.inverse[4 * r + c] = t->inverse[r][c]; }
        transforms.push_back(x);
        return (int32_t)transforms.size() - 1;
    }

    int32_t add_interior(const Interior* in)
    {
        if (in == nullptr) return -1;
        auto it = interior_ids.find(in);
        if (it != interior_ids.end()) return it->second;
        if (!in->media.empty()) unsupported("media inside an interior");
        pvgpu_interior p{};
        p.hollow = in->hollow; p.disp_nelems = in->Disp_NElems;
        p.ior = in->IOR; p.dispersion = in->Dispersion; p.caustics = in->Caustics; p.old_refract = in->Old_Refract;
        p.fade_distance = in->Fade_Distance; p.fade_power = in->Fade_Power;
        for (int k = 0; k < 3; k++) p.fade_colour[k] = in->Fade_Colour[k];
        interiors.push_back(p);
        return interior_ids[in] = (int32_t)interiors.size() - 1;
    }

    void add_warps(const WarpList& wl, uint32_t& first, uint32_t& count)
    {
        first = (uint32_t)warps.size();
        count = 0;
        for (WarpList::const_iterator i = wl.begin(); i != wl.end(); ++i) {
            pvgpu_warp w{};
            w.transform = -1;
            if (const TransformWarp* tw = dynamic_cast<const TransformWarp*>(*i)) {
                w.type = PVGPU_WARP_TRANSFORM;
                w.transform = add_transform(&tw->Trans);
            } else if (const GenericTurbulenceWarp* gt = dynamic_cast<const GenericTurbulenceWarp*>(*i)) {
                const ClassicTurbulence* ct = dynamic_cast<const ClassicTurbulence*>(*i);
                w.type = ct ? PVGPU_WARP_CLASSIC_TURBULENCE : PVGPU_WARP_TURBULENCE;
                for (int k = 0; k < 3; k++) w.turbulence[k] = gt->Turbulence[k];
                w.octaves = gt->Octaves; w.lambda = gt->Lambda; w.omega = gt->Omega;
                w.handled_by_pattern = ct ? ct->handledByPattern : 0;
            } else if (dynamic_cast<const IdentityWarp*>(*i)) {
                continue;                                   // maps every point onto itself
            } else {
                // the point-mapping warps: parameters as doubles in the shape-data table (pvgpu.h, PVGPU_WARP_*)
                w.transform = (int32_t)shape_data.size();
                if (const BlackHoleWarp* bh = dynamic_cast<const BlackHoleWarp*>(*i)) {
                    w.type = PVGPU_WARP_BLACK_HOLE;
                    if (bh->Uncertain) unsupported("black_hole warp with turbulence (WarpRands)");
                    for (double v : { (double)bh->Center[X], (double)bh->Center[Y], (double)bh->Center[Z], (double)bh->Repeat_Vector[X], (double)bh->Repeat_Vector[Y],
                                      (double)bh->Repeat_Vector[Z], (double)bh->Strength, (double)bh->Radius, (double)bh->Power,
                                      (double)((bh->Inverted ? 1 : 0) | (bh->Repeat ? 2 : 0)), (double)bh->Type }) shape_data.push_back(v);
                } else if (const RepeatWarp* rw = dynamic_cast<const RepeatWarp*>(*i)) {
                    w.type = PVGPU_WARP_REPEAT;
                    for (double v : { (double)rw->Axis, (double)rw->Width, (double)rw->Flip[X], (double)rw->Flip[Y], (double)rw->Flip[Z],
                                      (double)rw->Offset[X], (double)rw->Offset[Y], (double)rw->Offset[Z] }) shape_data.push_back(v);
                } else if (dynamic_cast<const CubicWarp*>(*i)) {
                    w.type = PVGPU_WARP_CUBIC;
                } else if (const CylindricalWarp* cw = dynamic_cast<const CylindricalWarp*>(*i)) {
                    w.type = PVGPU_WARP_CYLINDRICAL;
                    for (double v : { (double)cw->Orientation_Vector[X], (double)cw->Orientation_Vector[Y], (double)cw->Orientation_Vector[Z], (double)cw->DistExp }) shape_data.push_back(v);
                } else if (const SphericalWarp* sw = dynamic_cast<const SphericalWarp*>(*i)) {
                    w.type = PVGPU_WARP_SPHERICAL;
                    for (double v : { (double)sw->Orientation_Vector[X], (double)sw->Orientation_Vector[Y], (double)sw->Orientation_Vector[Z], (double)sw->DistExp }) shape_data.push_back(v);
                } else if (const ToroidalWarp* tw2 = dynamic_cast<const ToroidalWarp*>(*i)) {
                    w.type = PVGPU_WARP_TOROIDAL;
                    for (double v : { (double)tw2->Orientation_Vector[X], (double)tw2->Orientation_Vector[Y], (double)tw2->Orientation_Vector[Z], (double)tw2->DistExp, (double)tw2->MajorRadius }) shape_data.push_back(v);
                } else if (const PlanarWarp* pw = dynamic_cast<const PlanarWarp*>(*i)) {
                    w.type = PVGPU_WARP_PLANAR;
                    for (double v : { (double)pw->Orientation_Vector[X], (double)pw->Orientation_Vector[Y], (double)pw->Orientation_Vector[Z], (double)pw->OffSet }) shape_data.push_back(v);
                } else { unsupported("warp type outside the hot-path scope"); w.type = PVGPU_WARP_TRANSFORM; w.transform = -1; }
            }
            warps.push_back(w);
            count++;
        }
    }

    // the TPATTERN part PIGMENT and TNORMAL share: pattern kind + parameters, waveform, noise generator, warps
    void fill_pattern(const BasicPattern* bp, pvgpu_pigment& p, const char* user)
    {
        if (dynamic_cast<const CheckerPattern*>(bp)) p.pattern = PVGPU_PAT_CHECKER;
        else if (dynamic_cast<const BozoPattern*>(bp)) p.pattern = PVGPU_PAT_BOZO;
        else if (dynamic_cast<const SpottedPattern*>(bp)) p.pattern = PVGPU_PAT_SPOTTED;
        else if (dynamic_cast<const GranitePattern*>(bp)) p.pattern = PVGPU_PAT_GRANITE;
        else if (const GradientPattern* g = dynamic_cast<const GradientPattern*>(bp)) { p.pattern = PVGPU_PAT_GRADIENT; for (int k = 0; k < 3; k++) p.p[k] = g->gradient[k]; }
        else if (dynamic_cast<const MarblePattern*>(bp)) p.pattern = PVGPU_PAT_MARBLE;
        else if (dynamic_cast<const OnionPattern*>(bp)) p.pattern = PVGPU_PAT_ONION;
        else if (dynamic_cast<const WrinklesPattern*>(bp)) p.pattern = PVGPU_PAT_WRINKLES;
        else if (const AgatePattern* a = dynamic_cast<const AgatePattern*>(bp)) { p.pattern = PVGPU_PAT_AGATE; p.p[0] = a->agateTurbScale; }
        else if (const BrickPattern* b = dynamic_cast<const BrickPattern*>(bp)) { p.pattern = PVGPU_PAT_BRICK; for (int k = 0; k < 3; k++) p.p[k] = b->brickSize[k]; p.p[3] = b->mortar; }
        else if (dynamic_cast<const HexagonPattern*>(bp)) p.pattern = PVGPU_PAT_HEXAGON;
        else if (dynamic_cast<const WoodPattern*>(bp)) p.pattern = PVGPU_PAT_WOOD;
        else if (dynamic_cast<const LeopardPattern*>(bp)) p.pattern = PVGPU_PAT_LEOPARD;
        else if (dynamic_cast<const SphericalPattern*>(bp)) p.pattern = PVGPU_PAT_SPHERICAL;
        else if (dynamic_cast<const BoxedPattern*>(bp)) p.pattern = PVGPU_PAT_BOXED;
        else if (dynamic_cast<const RadialPattern*>(bp)) p.pattern = PVGPU_PAT_RADIAL;
        else if (dynamic_cast<const CylindricalPattern*>(bp)) p.pattern = PVGPU_PAT_CYLINDRICAL;
        else if (dynamic_cast<const PlanarPattern*>(bp)) p.pattern = PVGPU_PAT_PLANAR;
        else if (dynamic_cast<const DentsPattern*>(bp)) p.pattern = PVGPU_PAT_DENTS;
        else if (dynamic_cast<const RipplesPattern*>(bp)) p.pattern = PVGPU_PAT_RIPPLES;
        else if (dynamic_cast<const WavesPattern*>(bp)) p.pattern = PVGPU_PAT_WAVES;
        else if (const QuiltedPattern* q = dynamic_cast<const QuiltedPattern*>(bp)) { p.pattern = PVGPU_PAT_QUILTED; p.p[0] = q->Control0; p.p[1] = q->Control1; }
        else if (const CracklePattern* cr = dynamic_cast<const CracklePattern*>(bp)) {
            p.pattern = PVGPU_PAT_CRACKLE;
            p.data = (uint32_t)shape_data.size();
            for (int k = 0; k < 3; k++) shape_data.push_back(cr->crackleForm[k]);
            shape_data.push_back(cr->crackleMetric); shape_data.push_back(cr->crackleOffset); shape_data.push_back(cr->crackleIsSolid ? 1.0 : 0.0);
            shape_data.push_back(cr->repeat.x()); shape_data.push_back(cr->repeat.y()); shape_data.push_back(cr->repeat.z());
        }
        else if (dynamic_cast<const CellsPattern*>(bp)) p.pattern = PVGPU_PAT_CELLS;
        else if (const FractalPattern* fp = dynamic_cast<const FractalPattern*>(bp)) {
            int kind = -1;
            if (dynamic_cast<const Mandel2Pattern*>(bp)) kind = PVGPU_FRACTAL_MANDEL2;
            else if (dynamic_cast<const Mandel3Pattern*>(bp)) kind = PVGPU_FRACTAL_MANDEL3;
            else if (dynamic_cast<const Mandel4Pattern*>(bp)) kind = PVGPU_FRACTAL_MANDEL4;
            else if (dynamic_cast<const Magnet1MPattern*>(bp)) kind = PVGPU_FRACTAL_MAGNET1M;
            else if (dynamic_cast<const Magnet2MPattern*>(bp)) kind = PVGPU_FRACTAL_MAGNET2M;
            else if (dynamic_cast<const Magnet1JPattern*>(bp)) kind = PVGPU_FRACTAL_MAGNET1J;
            else if (dynamic_cast<const Magnet2JPattern*>(bp)) kind = PVGPU_FRACTAL_MAGNET2J;
            else if (dynamic_cast<const Julia3Pattern*>(bp)) kind = PVGPU_FRACTAL_JULIA3;
            else if (dynamic_cast<const Julia4Pattern*>(bp)) kind = PVGPU_FRACTAL_JULIA4;
            else if (dynamic_cast<const JuliaXPattern*>(bp) == nullptr && typeid(*bp) == typeid(JuliaPattern)) kind = PVGPU_FRACTAL_JULIA2;
            if (kind < 0 || fp->maxIterations < 1 || (fp->exteriorType == 7 && fp->exteriorFactor < 1.0) || (fp->exteriorType == 8 && fp->exteriorFactor < 0.0))
                unsupported(std::string(user) + " pattern outside the hot-path scope: " + typeid(*bp).name());
            else {
                p.pattern = PVGPU_PAT_FRACTAL;
                p.data = (uint32_t)shape_data.size();
                const JuliaPattern* jp = dynamic_cast<const JuliaPattern*>(bp);
                for (double v : { (double)kind, (double)fp->maxIterations, (double)fp->exteriorType, (double)fp->interiorType, (double)fp->exteriorFactor,
                                  (double)fp->interiorFactor, jp ? (double)jp->juliaCoord[U] : 0.0, jp ? (double)jp->juliaCoord[V] : 0.0 }) shape_data.push_back(v);
            }
        }
        else if (const SpiralPattern* sp = dynamic_cast<const SpiralPattern*>(bp)) {
            p.pattern = dynamic_cast<const Spiral1Pattern*>(bp) ? PVGPU_PAT_SPIRAL1 : PVGPU_PAT_SPIRAL2;
            p.p[0] = (double)sp->arms;
        }
        else if (const PigmentPattern* pp = dynamic_cast<const PigmentPattern*>(bp)) {
            if (pp->pPigment == nullptr) unsupported(std::string(user) + " pigment_pattern without a pigment");
            else { p.pattern = PVGPU_PAT_PIGMENT; p.data = (uint32_t)add_pigment(pp->pPigment); }
        }
        else if (dynamic_cast<const BumpsPattern*>(bp)) p.pattern = PVGPU_PAT_BOZO;      // BumpsPattern is a NoisePattern (pattern.h:989)
        else unsupported(std::string(user) + " pattern outside the hot-path scope: " + typeid(*bp).name());
        if (const ContinuousPattern* cp = dynamic_cast<const ContinuousPattern*>(bp)) {
            p.wave_type = cp->waveType; p.frequency = cp->waveFrequency; p.phase = cp->wavePhase; p.exponent = cp->waveExponent;
        }
        p.noise_generator = bp->noiseGenerator;
        add_warps(bp->warps, p.warp_first, p.warp_count);
    }

    // TNORMAL (normal.h:118-123)
    int32_t add_tnormal(const TNORMAL* tn)
    {
        pvgpu_tnormal t{};
        pvgpu_pigment carrier{};
        carrier.blend_map = -1;
        carrier.pattern = PVGPU_PAT_PLAIN;
        carrier.wave_type = PVGPU_WAVE_RAMP; carrier.frequency = 1.0f; carrier.exponent = 1.0f;
        const BasicPattern* bp = tn->pattern.get();
        t.flags = tn->Flags & PVGPU_DONT_SCALE_BUMPS_FLAG;
        t.amount = tn->Amount; t.delta = tn->Delta;
        const SlopeBlendMap* sm = dynamic_cast<const SlopeBlendMap*>(tn->Blend_Map.get());
        const NormalBlendMap* nm = dynamic_cast<const NormalBlendMap*>(tn->Blend_Map.get());
        if (tn->Blend_Map != nullptr && sm == nullptr && nm == nullptr) unsupported("normal blend map of unknown kind");
        if (nm != nullptr) {
            // normal_map: the entries are normals of their own (normal.cpp:824-848, 1033-1059)
            if (tn->Type == UV_MAP_PATTERN) unsupported("uv_mapping normal");
            vector<pvgpu_blend_entry> own;
            for (const auto& e : nm->Blend_Map_Entries) {
                pvgpu_blend_entry be{};
                be.value = e.value;
                be.colour[0] = (float)add_tnormal(e.Vals);
                own.push_back(be);
            }
            pvgpu_blend_map m{};
            m.entry_first = (uint32_t)entries.size();
            m.entry_count = (uint32_t)own.size();
            m.blend_mode = PVGPU_BLEND_NORMAL_MAP;
            entries.insert(entries.end(), own.begin(), own.end());
            maps.push_back(m);
            t.normal_map = (uint32_t)maps.size();
            if (tn->Type == AVERAGE_PATTERN) { t.type = PVGPU_NORM_AVERAGE; add_warps(bp->warps, carrier.warp_first, carrier.warp_count); }
            else { t.type = PVGPU_NORM_PATTERN; fill_pattern(bp, carrier, "normal"); }
            pigments.push_back(carrier);
            t.pattern = (int32_t)pigments.size() - 1;
            tnormals.push_back(t);
            return (int32_t)tnormals.size() - 1;
        }
        bool special = true;
        switch (tn->Type) {
            case BUMPS_PATTERN:    t.type = PVGPU_NORM_BUMPS; break;
            case DENTS_PATTERN:    t.type = PVGPU_NORM_DENTS; break;
            case RIPPLES_PATTERN:  t.type = PVGPU_NORM_RIPPLES; break;
            case WAVES_PATTERN:    t.type = PVGPU_NORM_WAVES; break;
            case WRINKLES_PATTERN: t.type = PVGPU_NORM_WRINKLES; break;
            case QUILTED_PATTERN:  t.type = PVGPU_NORM_QUILTED; break;
            default:
                special = false;
                if (tn->Type <= LAST_SPECIAL_NORM_PATTERN) unsupported("normal type outside the hot-path scope (bump_map, facets, average, uv_mapping)");
                else t.type = PVGPU_NORM_PATTERN;
        }
        if (special) {
            if (const ContinuousPattern* cp = dynamic_cast<const ContinuousPattern*>(bp)) {
                carrier.wave_type = cp->waveType; carrier.frequency = cp->waveFrequency; carrier.phase = cp->wavePhase; carrier.exponent = cp->waveExponent;
            }
            if (const QuiltedPattern* q = dynamic_cast<const QuiltedPattern*>(bp)) { carrier.p[0] = q->Control0; carrier.p[1] = q->Control1; }
            carrier.noise_generator = bp->noiseGenerator;
            add_warps(bp->warps, carrier.warp_first, carrier.warp_count);
        } else if (t.type == PVGPU_NORM_PATTERN) {
            fill_pattern(bp, carrier, "normal");     // (a block pattern without a normal_map is sampled like any other pattern, normal.cpp:880-905)
        }
        // WarpNormal exists for transform warps only (warp.cpp:582-640): any other warp leaves the normal as it is
        if (sm != nullptr) {
            t.slope_first = (uint32_t)slope_entries.size();
            t.slope_count = (uint32_t)sm->Blend_Map_Entries.size();
            for (const auto& e : sm->Blend_Map_Entries) {
                pvgpu_slope_entry se{};
                se.value = e.value; se.height = e.Vals[0]; se.slope = e.Vals[1];
                slope_entries.push_back(se);
            }
        }
        pigments.push_back(carrier);
        t.pattern = (int32_t)pigments.size() - 1;
        tnormals.push_back(t);
        return (int32_t)tnormals.size() - 1;
    }

    // PIGMENT::Blend_Map: a colour_map (ColourBlendMap) or a pigment_map (PigmentBlendMap, entries are pigments of their own)
    void add_pigment_map(const PIGMENT* pg, pvgpu_pigment& p)
    {
        const ColourBlendMap* cm = dynamic_cast<const ColourBlendMap*>(pg->Blend_Map.get());
        const PigmentBlendMap* pm = dynamic_cast<const PigmentBlendMap*>(pg->Blend_Map.get());
        const GenericPigmentBlendMap* gm = dynamic_cast<const GenericPigmentBlendMap*>(pg->Blend_Map.get());
        if (cm == nullptr && pm == nullptr) { unsupported("pigment without a colour / pigment blend map"); return; }
        if (gm != nullptr && gm->blendMode != 0) unsupported("colour_map blend_mode other than 0");
        vector<pvgpu_blend_entry> own;
        if (cm != nullptr)
            for (const auto& e : cm->Blend_Map_Entries) {
                pvgpu_blend_entry be;
                be.value = e.value;
                for (int k = 0; k < 3; k++) be.colour[k] = e.Vals.colour()[k];
                be.colour[3] = e.Vals.filter(); be.colour[4] = e.Vals.transm();
                own.push_back(be);
            }
        else
            for (const auto& e : pm->Blend_Map_Entries) {
                pvgpu_blend_entry be{};
                be.value = e.value;
                be.colour[0] = (float)add_pigment(e.Vals);      // nested pigments first: their own maps' entries stay contiguous
                own.push_back(be);
            }
        pvgpu_blend_map m{};
        m.entry_first = (uint32_t)entries.size();
        m.entry_count = (uint32_t)own.size();
        m.blend_mode = (pm != nullptr) ? PVGPU_BLEND_PIGMENT_MAP : 0;
        entries.insert(entries.end(), own.begin(), own.end());
        maps.push_back(m);
        p.blend_map = (int32_t)maps.size() - 1;
    }

    // ImageData of an image_map (imageutil.h:104-131).  The texels are decoded here through the reference's own Image interface
    // (file gamma, palettes, per-index filter / transmit), in the space image_colour_at(..., premul = false) interpolates in
    // (imageutil.cpp:400-409); the legacy "transmit / filter all" is added to every texel like no_interpolation does (:1042-1049).
    int32_t add_image(const ImageData* id)
    {
        auto it = image_ids.find(id);
        if (it != image_ids.end()) return it->second;
        const Image* img = id->data;
        pvgpu_image pi{};
        pi.width = (uint32_t)id->iwidth; pi.height = (uint32_t)id->iheight;
        pi.fwidth = id->width; pi.fheight = id->height;
        pi.map_type = (uint32_t)id->Map_Type; pi.interpolation = (uint32_t)id->Interpolation_Type;
        if (pi.interpolation == 1) pi.interpolation = 0;        // NEAREST_NEIGHBOR "would be essentially the same as NO_INTERPOLATION" (imageutil.h:90)
        pi.all_filter = id->AllFilter; pi.all_transmit = id->AllTransmit;
        for (int k = 0; k < 3; k++) pi.gradient[k] = id->Gradient[k];
        pi.offset[0] = id->Offset[U]; pi.offset[1] = id->Offset[V];
        const bool proper_all = img->HasTransparency() && !id->AllTransmitLegacyMode && !img->IsIndexed() && ((id->AllTransmit != 0.0) || (id->AllFilter != 0.0));
        const bool get_premul = proper_all ? false : img->IsPremultiplied();
        pi.flags = (id->Once_Flag ? PVGPU_IMAGE_ONCE : 0u) | (get_premul ? PVGPU_IMAGE_PREMULTIPLIED : 0u) | (proper_all ? PVGPU_IMAGE_TRANSMIT_ALL : 0u);
        pi.data_first = (uint32_t)(texels.size() / 5);
        const bool legacy_add = !img->IsIndexed() && id->AllTransmitLegacyMode;
        for (int y = 0; y < id->iheight; y++)
            for (int x = 0; x < id->iwidth; x++) {
                RGBFTColour c;
                img->GetRGBFTValue((unsigned int)x, (unsigned int)y, c, get_premul);
                if (legacy_add) { c.transm() += id->AllTransmit; c.filter() += id->AllFilter; }
                texels.push_back(c.red()); texels.push_back(c.green()); texels.push_back(c.blue()); texels.push_back(c.filter()); texels.push_back(c.transm());
            }
        images.push_back(pi);
        image_ids[id] = (int32_t)images.size() - 1;
        return (int32_t)images.size() - 1;
    }

    int32_t add_pigment(const PIGMENT* pg)
    {
        pvgpu_pigment p{};
        p.blend_map = -1;
        for (int k = 0; k < 3; k++) { p.colour[k] = pg->colour.colour()[k]; p.quick_colour[k] = pg->Quick_Colour.colour()[k]; }
        p.colour[3] = pg->colour.filter(); p.colour[4] = pg->colour.transm();
        p.quick_colour[3] = pg->Quick_Colour.filter(); p.quick_colour[4] = pg->Quick_Colour.transm();
        p.wave_type = PVGPU_WAVE_RAMP; p.frequency = 1.0f; p.exponent = 1.0f;
        const BasicPattern* bp = pg->pattern.get();
        if (pg->Type == PLAIN_PATTERN) p.pattern = PVGPU_PAT_PLAIN;
        else if (pg->Type == AVERAGE_PATTERN) {
            p.pattern = PVGPU_PAT_AVERAGE;
            add_warps(pg->pattern->warps, p.warp_first, p.warp_count);
            add_pigment_map(pg, p);
        }
        else if (pg->Type == IMAGE_MAP_PATTERN && dynamic_cast<const ColourImagePattern*>(bp) != nullptr &&
                 dynamic_cast<const ColourImagePattern*>(bp)->pImage != nullptr && dynamic_cast<const ColourImagePattern*>(bp)->pImage->data != nullptr) {
            p.pattern = PVGPU_PAT_IMAGE_MAP;
            add_warps(pg->pattern->warps, p.warp_first, p.warp_count);
            p.data = (uint32_t)add_image(dynamic_cast<const ColourImagePattern*>(bp)->pImage);
        }
        else if (pg->Type == UV_MAP_PATTERN && std::dynamic_pointer_cast<PigmentBlendMap>(pg->Blend_Map) != nullptr &&
                 !std::dynamic_pointer_cast<PigmentBlendMap>(pg->Blend_Map)->Blend_Map_Entries.empty()) {
            // pigment { uv_mapping <pigment> }: the one entry of the blend list is evaluated at the hit's (u, v, 0) (pigment.cpp:603-618)
            p.pattern = PVGPU_PAT_UV_MAP;
            p.data = (uint32_t)add_pigment(std::dynamic_pointer_cast<PigmentBlendMap>(pg->Blend_Map)->Blend_Map_Entries[0].Vals);
        }
        else if (pg->Type <= LAST_SPECIAL_PATTERN) unsupported("pigment type other than plain / pattern / average / image_map / uv_mapping");
        else {
            fill_pattern(bp, p, "pigment");
            add_pigment_map(pg, p);
        }
        pigments.push_back(p);
        return (int32_t)pigments.size() - 1;
    }

    int32_t add_finish(const FINISH* f)
    {
        pvgpu_finish p{};
        p.diffuse = f->Diffuse; p.diffuse_back = f->DiffuseBack; p.brilliance = f->Brilliance;
        p.brilliance_adjust = f->BrillianceAdjust; p.brilliance_adjust_rad = f->BrillianceAdjustRad;
        p.specular = f->Specular; p.roughness = f->Roughness; p.phong = f->Phong; p.phong_size = f->Phong_Size;
        p.irid = f->Irid; p.irid_film_thickness = f->Irid_Film_Thickness; p.irid_turb = f->Irid_Turb;
        p.reflect_exp = f->Reflect_Exp; p.crand = f->Crand; p.metallic = f->Metallic;
        for (int k = 0; k < 3; k++) {
            p.ambient[k] = f->Ambient[k]; p.emission[k] = f->Emission[k];
            p.reflection_max[k] = f->Reflection_Max[k]; p.reflection_min[k] = f->Reflection_Min[k];
        }
        p.reflection_falloff = f->Reflection_Falloff; p.fresnel = f->Fresnel; p.reflect_metallic = f->Reflect_Metallic;
        p.reflection_fresnel = f->Reflection_Fresnel; p.conserve_energy = f->Conserve_Energy;
        p.alpha_knockout = f->AlphaKnockout; p.use_subsurface = f->UseSubsurface;
        finishes.push_back(p);
        return (int32_t)finishes.size() - 1;
    }

    int32_t add_texture(const TEXTURE* t)
    {
        if (t == nullptr) return -1;
        auto it = texture_ids.find(t);
        if (it != texture_ids.end()) return it->second;
        // reserve the slots of the whole layer chain first so that `next` can be filled in
        vector<const TEXTURE*> chain;
        for (const TEXTURE* l = t; l != nullptr; l = l->Next) chain.push_back(l);
        int32_t first = (int32_t)textures.size();
        textures.resize(textures.size() + chain.size());
        for (size_t i = 0; i < chain.size(); i++) {
            const TEXTURE* l = chain[i];
            pvgpu_texture p{};
            p.next = (i + 1 < chain.size()) ? first + (int32_t)i + 1 : -1;
            p.tnormal = -1;
            if (l->Type != PLAIN_PATTERN) {
                // texture_map / average texture_map (texture.h:96-117): pattern carrier + map of whole textures
                const TextureBlendMap* tm = dynamic_cast<const TextureBlendMap*>(l->Blend_Map.get());
                if (l->Type == BITMAP_PATTERN || l->Type == UV_MAP_PATTERN || tm == nullptr) { unsupported("material_map / uv_mapping texture"); p.type = 0; }
                else if (l->Flags & PVGPU_DONT_SCALE_BUMPS_FLAG) { unsupported("no_bump_scale on a patterned texture"); p.type = 0; }
                else {
                    pvgpu_pigment carrier{};
                    carrier.blend_map = -1;
                    carrier.pattern = PVGPU_PAT_PLAIN;
                    carrier.wave_type = PVGPU_WAVE_RAMP; carrier.frequency = 1.0f; carrier.exponent = 1.0f;
                    if (l->Type == AVERAGE_PATTERN) { p.type = PVGPU_PAT_AVERAGE; add_warps(l->pattern->warps, carrier.warp_first, carrier.warp_count); }
                    else { fill_pattern(l->pattern.get(), carrier, "texture"); p.type = carrier.pattern; }
                    vector<pvgpu_blend_entry> own;
                    for (const auto& e : tm->Blend_Map_Entries) {
                        pvgpu_blend_entry be{};
                        be.value = e.value;
                        be.colour[0] = (float)add_texture(e.Vals);
                        own.push_back(be);
                    }
                    pvgpu_blend_map m{};
                    m.entry_first = (uint32_t)entries.size();
                    m.entry_count = (uint32_t)own.size();
                    m.blend_mode = PVGPU_BLEND_TEXTURE_MAP;
                    entries.insert(entries.end(), own.begin(), own.end());
                    maps.push_back(m);
                    p.blend_map = (int32_t)maps.size() - 1;
                    pigments.push_back(carrier);
                    p.pigment = (int32_t)pigments.size() - 1;
                    p.finish = -1;
                }
            }
            else {
                p.type = PVGPU_PAT_PLAIN;
                p.pigment = add_pigment(l->Pigment);
                p.finish = add_finish(l->Finish);
                if (l->Tnormal != nullptr) p.tnormal = add_tnormal(l->Tnormal);
            }
            textures[first + i] = p;
            texture_ids[l] = first + (int32_t)i;
        }
        return first;
    }

    void add_index_range(const vector<ObjectPtr>& v, uint32_t self, bool as_children, uint32_t& first, uint32_t& count)
    {
        // children are flattened first so that the range in the index list is contiguous
        vector<uint32_t> ids;
        for (ObjectPtr o : v) {
            if ((o->Type & LIGHT_SOURCE_OBJECT) != 0) {
                if (!(reinterpret_cast<LightSource*>(o))->children.empty()) unsupported("light source with looks_like inside a compound object");
                continue;      // a light without geometry takes no part in intersection / inside tests
            }
            ids.push_back(add_object(o, as_children ? (int32_t)self : -1));
        }
        first = (uint32_t)index_list.size();
        count = (uint32_t)ids.size();
        index_list.insert(index_list.end(), ids.begin(), ids.end());
    }

    int32_t add_mesh(Mesh* m)
    {
        pvgpu_mesh me{};
        const MESH_DATA* D = m->Data;
        me.vertex_first = (uint32_t)(vertices.size() / 3); me.vertex_count = D->Number_Of_Vertices;
        me.normal_first = (uint32_t)(normals.size() / 3);  me.normal_count = D->Number_Of_Normals;
        me.triangle_first = (uint32_t)triangles.size();    me.triangle_count = D->Number_Of_Triangles;
        mesh_tri_first[m] = me.triangle_first;
        for (int i = 0; i < D->Number_Of_Vertices; i++) for (int k = 0; k < 3; k++) vertices.push_back(D->Vertices[i][k]);
        for (int i = 0; i < D->Number_Of_Normals; i++) for (int k = 0; k < 3; k++) normals.push_back(D->Normals[i][k]);
        for (int i = 0; i < D->Number_Of_Triangles; i++) {
            const MESH_TRIANGLE& t = D->Triangles[i];
            pvgpu_triangle p{};
            for (int k = 0; k < 3; k++) p.perp[k] = t.Perp[k];
            p.distance = t.Distance; p.normal_ind = t.Normal_Ind;
            p.p1 = t.P1; p.p2 = t.P2; p.p3 = t.P3; p.n1 = t.N1; p.n2 = t.N2; p.n3 = t.N3;
            p.texture = t.Texture; p.texture2 = t.Texture2; p.texture3 = t.Texture3;
            p.flags = (t.Smooth ? PVGPU_TRI_SMOOTH : 0) | (t.ThreeTex ? PVGPU_TRI_THREETEX : 0);
            p.dominant_axis = t.Dominant_Axis; p.v_axis = t.vAxis;
            triangles.push_back(p);
            // UV1..UV3 as indices into the scene-wide UV table (Mesh::UVCoord, mesh.cpp:2328-2330)
            const uint32_t uv0 = (uint32_t)(mesh_uv.size() / 2);
            const bool has_uv = D->UVCoords != nullptr && D->Number_Of_UVCoords > 0;
            for (MeshIndex u : { t.UV1, t.UV2, t.UV3 }) tri_uv.push_back(has_uv ? uv0 + (uint32_t)u : 0u);
            any_mesh_uv = any_mesh_uv || has_uv;
        }
        if (D->UVCoords != nullptr) for (int i = 0; i < D->Number_Of_UVCoords; i++) { mesh_uv.push_back(D->UVCoords[i][U]); mesh_uv.push_back(D->UVCoords[i][V]); }
        if (mesh_uv.empty()) { mesh_uv.push_back(0.0); mesh_uv.push_back(0.0); }      // (entry 0: what a mesh without uv_vectors indexes)
        me.texture_first = (uint32_t)index_list.size();
        me.texture_count = (uint32_t)m->Number_Of_Textures;
        {
            vector<uint32_t> tids;
            for (int i = 0; i < m->Number_Of_Textures; i++) tids.push_back((uint32_t)add_texture(m->Textures[i]));
            me.texture_first = (uint32_t)index_list.size();
            index_list.insert(index_list.end(), tids.begin(), tids.end());
        }
        me.has_inside_vector = m->has_inside_vector;
        if (m->has_inside_vector) for (int k = 0; k < 3; k++) me.inside_vector[k] = D->Inside_Vect[k];
        me.node_first = (uint32_t)mesh_nodes.size();
        if (D->Tree != nullptr) {
            vector<pvgpu_node> tree;
            flatten_tree(D->Tree, tree, [&](const BBOX_TREE* leaf) { return (uint32_t)(reinterpret_cast<const MESH_TRIANGLE*>(leaf->Node) - D->Triangles); });
            me.node_count = (uint32_t)tree.size();
            mesh_nodes.insert(mesh_nodes.end(), tree.begin(), tree.end());
        }
        meshes.push_back(me);
        return (int32_t)meshes.size() - 1;
    }

    // Blob_Data as Make_Blob / build_bounding_hierarchy left it (blob.cpp:2516-2766)
    int32_t add_blob(Blob* b)
    {
        const Blob_Data* D = b->Data;
        pvgpu_blob pb{};
        pb.threshold = D->Threshold;
        pb.element_first = (uint32_t)blob_elements.size();
        pb.element_count = (uint32_t)D->Entry.size();
        for (const Blob_Element& e : D->Entry) {
            pvgpu_blob_element pe{};
            pe.type = (uint32_t)e.Type;
            pe.transform = add_transform(e.Trans);
            for (int k = 0; k < 3; k++) { pe.o[k] = e.O[k]; pe.c[k] = e.c[k]; }
            pe.len = e.len; pe.rad2 = e.rad2;
            // Blob::Element_Texture (blob.h:167): the texture Determine_Textures blends for this component, or the blob's own
            const size_t ei = blob_elements.size() - pb.element_first;
            TEXTURE* et = (ei < b->Element_Texture.size()) ? b->Element_Texture[ei] : nullptr;
            blob_textures.push_back(et != nullptr ? add_texture(et) : -1);
            if (et != nullptr) any_blob_texture = true;
            blob_elements.push_back(pe);
        }
        pb.node_first = (uint32_t)blob_nodes.size();
        if (D->Tree != nullptr) {
            // breadth-first so that the children of a node are contiguous; a leaf's Node points at its Blob_Element
            vector<const BSPHERE_TREE*> order{ D->Tree };
            vector<pvgpu_blob_node> tree(1);
            for (size_t qi = 0; qi < order.size(); qi++) {
                const BSPHERE_TREE* n = order[qi];
                pvgpu_blob_node pn{};
                for (int k = 0; k < 3; k++) pn.c[k] = n->C[k];
                pn.r2 = n->r2;
                if (n->Entries <= 0) {
                    pn.count = 0;
                    pn.first = (uint32_t)(reinterpret_cast<const Blob_Element*>(n->Node) - D->Entry.data());
                } else {
                    pn.count = (uint32_t)n->Entries;
                    pn.first = (uint32_t)order.size();
                    for (int i = 0; i < (int)n->Entries; i++) { order.push_back(n->Node[i]); tree.push_back(pvgpu_blob_node{}); }
                }
                tree[qi] = pn;
            }
            pb.node_count = (uint32_t)tree.size();
            blob_nodes.insert(blob_nodes.end(), tree.begin(), tree.end());
        }
        blobs.push_back(pb);
        return (int32_t)blobs.size() - 1;
    }

    // Outline of a glyph as the segment list GlyphIntersect / Inside_Glyph walk (truetype.cpp:2406-2560, 2790-2920): a point with
    // the on-curve flag ends a straight line, any other point is the control point of a parabola whose far end is the next point
    // (wrapping to the first), moved to the midpoint when that one is off-curve too.  Lines of zero length (the step onto the end point
    // of a parabola) are skipped by both walks (y0 == y1; |t0| < EPSILON) and left out.
    std::map<const void*, int32_t> glyph_ids;
    int32_t add_glyph(const TrueType* tt)
    {
        auto it = glyph_ids.find(tt->glyph);
        if (it != glyph_ids.end()) return it->second;
        const ttf_mirror::GlyphStruct* g = reinterpret_cast<const ttf_mirror::GlyphStruct*>(tt->glyph);
        const int32_t first = (int32_t)shape_data.size();
        glyph_ids[tt->glyph] = first;
        shape_data.push_back(0.0);
        if (g == nullptr || g->header.numContours < 0 || g->header.numContours > 4096) { unsupported("glyph outline not readable"); return first; }
        size_t nseg = 0;
        for (int i = 0; i < g->header.numContours; i++) {
            const ttf_mirror::Contour& c = g->contours[i];
            const size_t n1 = c.count;
            if (n1 == 0) continue;
            if (c.x.size() < n1 + 1 || c.y.size() < n1 + 1 || c.flags.size() < n1 + 1) { unsupported("glyph outline not readable"); return first; }
            double x0 = c.x[0], y0 = c.y[0];
            for (size_t j = 1; j <= n1; j++) {
                const double x1 = c.x[j], y1 = c.y[j];
                if (c.flags[j] & 0x01) {        // ONCURVE
                    if (!(x0 == x1 && y0 == y1)) { for (double v : { 0.0, x0, y0, x1, y1, 0.0, 0.0 }) shape_data.push_back(v); nseg++; }
                    x0 = x1; y0 = y1;
                } else {
                    double x2, y2;
                    if (j == n1) { x2 = c.x[0]; y2 = c.y[0]; }
                    else {
                        x2 = c.x[j + 1]; y2 = c.y[j + 1];
                        if (!(c.flags[j + 1] & 0x01)) { x2 = 0.5 * (x1 + x2); y2 = 0.5 * (y1 + y2); }
                    }
                    for (double v : { 1.0, x0, y0, x1, y1, x2, y2 }) shape_data.push_back(v);
                    nseg++;
                    x0 = x2; y0 = y2;
                }
            }
        }
        shape_data[first] = (double)nseg;
        return first;
    }

    int32_t add_object(ObjectPtr o, int32_t parent)
    {
        auto it = object_ids.find(o);
        if (it != object_ids.end()) return it->second;
        const uint32_t self = (uint32_t)objects.size();
        object_ids[o] = (int32_t)self;
        objects.push_back(pvgpu_object{});
        pvgpu_object p{};
        p.flags = o->Flags;
        p.parent = parent;
        p.mesh = -1;
        p.texture = add_texture(o->Texture);
        p.interior_texture = add_texture(o->Interior_Texture);
        p.interior = add_interior(o->interior.get());
        p.bbox[0] = o->BBox.lowerLeft[X]; p.bbox[1] = o->BBox.lowerLeft[Y]; p.bbox[2] = o->BBox.lowerLeft[Z];
        p.bbox[3] = o->BBox.size[X]; p.bbox[4] = o->BBox.size[Y]; p.bbox[5] = o->BBox.size[Z];
        p.transform = -1;
        if (!o->LLights.empty()) unsupported("light_group");
        if (Sphere* s = dynamic_cast<Sphere*>(o)) {
            p.type = PVGPU_OBJ_SPHERE;
            for (int k = 0; k < 3; k++) p.p[k] = s->Center[k];
            p.p[3] = s->Radius;
            p.aux = s->Do_Ellipsoid ? 1 : 0;
            if (s->Do_Ellipsoid) p.transform = add_transform(s->Trans);
            else if (s->Trans != nullptr) p.transform = add_transform(s->Trans);       // kept for Sphere::UVCoord only (sphere.cpp:699-704)
        } else if (Box* b = dynamic_cast<Box*>(o)) {
            p.type = PVGPU_OBJ_BOX;
            for (int k = 0; k < 3; k++) { p.p[k] = b->bounds[0][k]; p.p[3 + k] = b->bounds[1][k]; }
            p.transform = add_transform(b->Trans);
        } else if (Plane* pl = dynamic_cast<Plane*>(o)) {
            p.type = PVGPU_OBJ_PLANE;
            for (int k = 0; k < 3; k++) p.p[k] = pl->Normal_Vector[k];
            p.p[3] = pl->Distance;
            p.transform = add_transform(pl->Trans);
        } else if (Quadric* q = dynamic_cast<Quadric*>(o)) {
            p.type = PVGPU_OBJ_QUADRIC;
            for (int k = 0; k < 3; k++) { p.p[k] = q->Square_Terms[k]; p.p[3 + k] = q->Mixed_Terms[k]; p.p[6 + k] = q->Terms[k]; }
            p.p[9] = q->Constant;
        } else if (Torus* t = dynamic_cast<Torus*>(o)) {
            p.type = PVGPU_OBJ_TORUS;
            p.p[0] = t->MajorRadius; p.p[1] = t->MinorRadius;
            if (SpindleTorus* st = dynamic_cast<SpindleTorus*>(o)) { p.aux = (uint32_t)st->mSpindleMode; p.p[2] = st->mSpindleTipYSqr; }
            p.transform = add_transform(t->Trans);
        } else if (Mesh* m = dynamic_cast<Mesh*>(o)) {
            p.type = PVGPU_OBJ_MESH;
            p.mesh = add_mesh(m);
            p.transform = add_transform(m->Trans);
        } else if (Disc* dc = dynamic_cast<Disc*>(o)) {
            p.type = PVGPU_OBJ_DISC;
            for (int k = 0; k < 3; k++) p.p[k] = dc->normal[k];
            p.p[3] = dc->iradius2; p.p[4] = dc->oradius2;
            p.transform = add_transform(dc->Trans);
        } else if (Triangle* tr = dynamic_cast<Triangle*>(o)) {
            p.type = PVGPU_OBJ_TRIANGLE;
            p.mesh = (int32_t)shape_data.size();
            p.aux = tr->Dominant_Axis;                  // (vAxis is only initialised for smooth triangles)
            for (const Vector3d* v : { &tr->P1, &tr->P2, &tr->P3, &tr->Normal_Vector }) for (int k = 0; k < 3; k++) shape_data.push_back((*v)[k]);
            shape_data.push_back(tr->Distance);
            if (SmoothTriangle* st = dynamic_cast<SmoothTriangle*>(o)) {
                p.aux |= PVGPU_TRIANGLE_SMOOTH | (tr->vAxis << 2);
                for (const Vector3d* v : { &st->N1, &st->N2, &st->N3, &st->Perp }) for (int k = 0; k < 3; k++) shape_data.push_back((*v)[k]);
            }
        } else if (Poly* po = dynamic_cast<Poly*>(o)) {
            p.type = PVGPU_OBJ_POLY;
            p.aux = (uint32_t)po->Order;
            p.mesh = (int32_t)shape_data.size();
            if (po->Order < 1 || po->Order > 4) unsupported("poly of order > 4");
            else for (int k = 0; k < (po->Order + 1) * (po->Order + 2) * (po->Order + 3) / 6; k++) shape_data.push_back(po->Coeffs[k]);
            p.transform = add_transform(po->Trans);
        } else if (Polygon* pg = dynamic_cast<Polygon*>(o)) {
            p.type = PVGPU_OBJ_POLYGON;
            for (int k = 0; k < 3; k++) p.p[k] = pg->S_Normal[k];
            p.mesh = (int32_t)shape_data.size();
            p.aux = (uint32_t)pg->Data->Number;
            for (int i = 0; i < pg->Data->Number; i++) { shape_data.push_back(pg->Data->Points[i][X]); shape_data.push_back(pg->Data->Points[i][Y]); }
            p.transform = add_transform(pg->Trans);
        } else if (Cone* cn = dynamic_cast<Cone*>(o)) {
            p.type = PVGPU_OBJ_CONE;
            p.p[0] = cn->dist;
            p.transform = add_transform(cn->Trans);
        } else if (Superellipsoid* se = dynamic_cast<Superellipsoid*>(o)) {
            p.type = PVGPU_OBJ_SUPERELLIPSOID;
            for (int k = 0; k < 3; k++) p.p[k] = se->Power[k];
            p.aux = (se->Type & IS_CHILD_OBJECT) ? 1u : 0u;
            p.transform = add_transform(se->Trans);
        } else if (Prism* pr = dynamic_cast<Prism*>(o)) {
            p.type = PVGPU_OBJ_PRISM;
            p.p[0] = pr->Height1; p.p[1] = pr->Height2;
            p.p[2] = pr->x1; p.p[3] = pr->y1; p.p[4] = pr->x2; p.p[5] = pr->y2;
            p.p[6] = pr->u1; p.p[7] = pr->v1; p.p[8] = pr->u2; p.p[9] = pr->v2;
            p.aux = (uint32_t)pr->Spline_Type | ((uint32_t)pr->Sweep_Type << 4);
            p.mesh = (int32_t)shape_data.size();
            shape_data.push_back((double)pr->Number);
            for (int k = 0; k < pr->Number; k++) {
                const PRISM_SPLINE_ENTRY& e = pr->Spline->Entry[k];
                for (double v : { e.x1, e.y1, e.x2, e.y2, e.v1, e.u2, e.v2, e.A[X], e.A[Y], e.B[X], e.B[Y], e.C[X], e.C[Y], e.D[X], e.D[Y] }) shape_data.push_back(v);
            }
            p.transform = add_transform(pr->Trans);
        } else if (TrueType* tt = dynamic_cast<TrueType*>(o)) {
            p.type = PVGPU_OBJ_GLYPH;
            p.p[0] = tt->depth;
            p.mesh = add_glyph(tt);
            p.transform = add_transform(tt->Trans);
        } else if (Blob* bl = dynamic_cast<Blob*>(o)) {
            p.type = PVGPU_OBJ_BLOB;
            p.mesh = add_blob(bl);
            p.transform = add_transform(bl->Trans);
            p.aux = (bl->Type & IS_CHILD_OBJECT) ? 1u : 0u;
        } else if (CSG* c = dynamic_cast<CSG*>(o)) {
            if (dynamic_cast<CSGMerge*>(o)) p.type = PVGPU_OBJ_CSG_MERGE;
            else if (dynamic_cast<CSGUnion*>(o)) p.type = PVGPU_OBJ_CSG_UNION;
            else if (dynamic_cast<CSGIntersection*>(o)) p.type = PVGPU_OBJ_CSG_INTERSECTION;
            else unsupported("unknown CSG class");
            add_index_range(c->children, self, true, p.child_first, p.child_count);
        } else {
            unsupported(std::string("primitive outside the hot-path scope: ") + typeid(*o).name());
            p.type = 0;
        }
        add_index_range(o->Clip, self, false, p.clip_first, p.clip_count);
        // Bound and Clip may share their objects (bounded_by { clipped_by }); the memo keeps them shared here too
        add_index_range(o->Bound, self, false, p.bound_first, p.bound_count);
        objects[self] = p;
        return (int32_t)self;
    }

    template <typename LEAF>
    void flatten_tree(const BBOX_TREE* root, vector<pvgpu_node>& out, LEAF leaf_payload)
    {
        vector<const BBOX_TREE*> order{ root };
        out.assign(1, pvgpu_node{});
        for (size_t qi = 0; qi < order.size(); qi++) {
            const BBOX_TREE* n = order[qi];
            pvgpu_node pn{};
            pn.lo[0] = n->BBox.lowerLeft[X]; pn.lo[1] = n->BBox.lowerLeft[Y]; pn.lo[2] = n->BBox.lowerLeft[Z];
            pn.size[0] = n->BBox.size[X]; pn.size[1] = n->BBox.size[Y]; pn.size[2] = n->BBox.size[Z];
            pn.flags = n->Infinite ? PVGPU_NODE_INFINITE : 0;
            if (n->Entries == 0) { pn.count = 0; pn.first = leaf_payload(n); }
            else {
                pn.count = (uint16_t)n->Entries;
                pn.first = (uint32_t)order.size();
                for (int i = 0; i < n->Entries; i++) order.push_back(n->Node[i]);
                out.resize(order.size());
            }
            out[qi] = pn;
        }
    }

    void add_light(const LightSource* l)
    {
        pvgpu_light p{};
        p.type = l->Light_Type;
        p.flags = (l->Area_Light ? PVGPU_LIGHT_AREA : 0) | (l->Use_Full_Area_Lighting ? PVGPU_LIGHT_FULL_AREA : 0) |
                  (l->Jitter ? PVGPU_LIGHT_JITTER : 0) | (l->Orient ? PVGPU_LIGHT_ORIENT : 0) | (l->Circular ? PVGPU_LIGHT_CIRCULAR : 0) |
                  (l->Parallel ? PVGPU_LIGHT_PARALLEL : 0) | (l->Media_Attenuation ? PVGPU_LIGHT_MEDIA_ATTEN : 0) |
                  (l->Media_Interaction ? PVGPU_LIGHT_MEDIA_INTERACT : 0) | (l->lightGroupLight ? PVGPU_LIGHT_GROUP : 0);
        for (int k = 0; k < 3; k++) {
            p.colour[k] = l->colour[k];
            p.center[k] = l->Center[k]; p.direction[k] = l->Direction[k]; p.points_at[k] = l->Points_At[k];
            p.axis1[k] = l->Axis1[k]; p.axis2[k] = l->Axis2[k];
        }
        p.projected_through = l->Projected_Through_Object ? add_object(l->Projected_Through_Object, -1) : -1;
        p.coeff = l->Coeff; p.radius = l->Radius; p.falloff = l->Falloff;
        p.fade_distance = l->Fade_Distance; p.fade_power = l->Fade_Power;
        p.area_size1 = l->Area_Size1; p.area_size2 = l->Area_Size2; p.adaptive_level = l->Adaptive_Level;
        p.object_flags = l->Flags;
        lights.push_back(p);
    }
};

int quality_bits(const QualityFlags& q)
{
    return (q.ambientOnly ? PVGPU_Q_AMBIENT_ONLY : 0) | (q.quickColour ? PVGPU_Q_QUICK_COLOUR : 0) | (q.shadows ? PVGPU_Q_SHADOWS : 0) |
           (q.areaLights ? PVGPU_Q_AREA_LIGHTS : 0) | (q.refractions ? PVGPU_Q_REFRACTIONS : 0) | (q.reflections ? PVGPU_Q_REFLECTIONS : 0) |
           (q.normals ? PVGPU_Q_NORMALS : 0) | (q.media ? PVGPU_Q_MEDIA : 0);
}

// The flattened scene of one SceneData.  Its lifetime is tied to that SceneData: the cache entry keeps a weak_ptr, an entry whose
// SceneData is gone (next frame of an animation, or a new SceneData allocated at the same address) is dropped and its device
// scene destroyed; views of the same scene rendered with different quality settings get their own entry.
struct GpuView
{
    pvgpu_scene* scene = nullptr;
    std::map<const void*, int32_t> object_ids;
    std::map<const void*, uint32_t> mesh_tri_first;
    bool finalized = false;
    std::string error;
    std::weak_ptr<BackendSceneData> owner;
    int quality = 0;
    GpuView() = default;
    GpuView(const GpuView&) = delete;
    GpuView& operator=(const GpuView&) = delete;
    ~GpuView() { if (scene) pvgpu_scene_destroy(scene); }
};

std::mutex g_mutex;
std::mutex g_drain_mutex;

// The CUDA driver and context come up on a background thread while the parser runs (about a second that a one-shot render
// would otherwise spend between parsing and the first frame).
struct Prewarm
{
    Prewarm()
    {
        const char* mode = getenv("PVGPU_RENDER");
        if (mode != nullptr && strcmp(mode, "stock") == 0) return;
        const char* dev = getenv("PVGPU_DEVICE");
        pvgpu_prewarm(dev ? atoi(dev) : 0);
    }
} g_prewarm;
std::map<std::pair<const void*, int>, std::shared_ptr<GpuView>> g_views;     // keyed by (SceneData, quality flags), validated through GpuView::owner

void check(int rc, const char* what)
{
    if (rc != PVGPU_OK)
        throw POV_EXCEPTION_STRING((std::string("pvgpu: ") + what + ": " + pvgpu_last_error()).c_str());
}

void finalize_view(GpuView& gv)
{
    if (!gv.error.empty())
        throw POV_EXCEPTION_STRING((std::string("pvgpu: scene uses a feature outside the GPU trace path: ") + gv.error).c_str());
    // PVGPU_DEVICES = "all" | "<n>" | "i,j,k": the scene is replicated on these devices and every frame is sharded over them
    // behind pvgpu_render (one atomic tile counter, like the render threads' GetNextRectangle); PVGPU_DEVICE = one index
    const char* devs = getenv("PVGPU_DEVICES");
    const char* dev = getenv("PVGPU_DEVICE");
    int rc;
    if (devs && *devs) {
        std::vector<int> list;
        if (strchr(devs, ',')) { for (const char* p = devs; *p;) { list.push_back(atoi(p)); p = strchr(p, ','); if (!p) break; p++; } }
        if (!list.empty()) rc = pvgpu_scene_finalize_multi(gv.scene, list.data(), (int)list.size());
        else rc = pvgpu_scene_finalize_multi(gv.scene, nullptr, strcmp(devs, "all") == 0 ? 0 : atoi(devs));
    } else rc = pvgpu_scene_finalize(gv.scene, dev ? atoi(dev) : 0);
    if (rc != PVGPU_OK) throw POV_EXCEPTION_STRING((std::string("pvgpu: scene_finalize: ") + pvgpu_last_error()).c_str());
    gv.finalized = true;
}

std::shared_ptr<GpuView> flatten_scene_locked(ViewData* vd, bool need_device);

std::shared_ptr<GpuView> flatten_scene(ViewData* vd, bool need_device)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    // entries of scenes that no longer exist go first (and free their device memory)
    for (auto it = g_views.begin(); it != g_views.end();)
        if (it->second->owner.expired()) it = g_views.erase(it); else ++it;
    const std::pair<const void*, int> key(vd->GetSceneData().get(), quality_bits(vd->GetQualityFeatureFlags()));
    auto it = g_views.find(key);
    if (it != g_views.end()) {
        if (need_device && !it->second->finalized) {
            try { finalize_view(*it->second); }
            catch (...) { g_views.erase(it); throw; }
        }
        return it->second;
    }
    try {
        std::shared_ptr<GpuView> gv = flatten_scene_locked(vd, need_device);
        g_views[key] = gv;
        return gv;
    } catch (...) {
        g_views.erase(key);        // nothing half-built stays behind for later tasks to pick up
        throw;
    }
}

std::shared_ptr<GpuView> flatten_scene_locked(ViewData* vd, bool need_device)
{
    SceneData* sd = vd->GetSceneData().get();
    std::shared_ptr<GpuView> slot(new GpuView());
    GpuView& gv = *slot;
    gv.owner = vd->GetSceneData();
    gv.quality = quality_bits(vd->GetQualityFeatureFlags());

    if (sd->rainbow != nullptr || !sd->atmosphere.empty())
        gv.error = "rainbow / atmospheric media (SURVEY 8f 'next')";
    if (sd->radiositySettings.radiosityEnabled) gv.error = "radiosity";
    if (sd->photonSettings.photonsEnabled) gv.error = "photons";
    if (sd->boundingMethod == 2) gv.error = "BSP bounding (+BM2)";
    if (sd->useSubsurface) gv.error = "subsurface light transport";

    Flattener fl;
    for (ObjectPtr o : sd->objects) {
        if ((o->Type & LIGHT_SOURCE_OBJECT) != 0) {
            // looks_like.  With a slab tree the tree's leaf IS children[0] (Build_Bounding_Slabs, boundingbox.cpp:337-347, 390-396): the
            // child is a frame-level object like any other.  Without one (boundingMethod 0, or too few objects for a tree) the loop over
            // SceneData::objects meets the light source: its flags gate the ray kinds (trace.cpp:84-95, 1943), its All_Intersections hands
            // the ray to children[0] behind that child's bounded_by list (lightsource.cpp:83-95), and InitRayContainerState looks at the
            // light source's own (absent) interior (tracepixel.cpp:950-955).
            LightSource* ls = reinterpret_cast<LightSource*>(o);
            if (ls->children.empty()) continue;
            ObjectPtr child = ls->children[0];
            const int32_t id = fl.add_object(child, -1);
            if (!(sd->boundingMethod == 1 && sd->boundingSlabs != nullptr)) {
                if (child->interior != nullptr && child->Inside(vd->GetCamera().Location, nullptr))
                    fl.unsupported("camera inside the interior of a looks_like object, no bounding tree");
                const uint32_t ray_kind = PVGPU_NO_SHADOW_FLAG | PVGPU_NO_IMAGE_FLAG | PVGPU_NO_REFLECTION_FLAG;
                fl.objects[id].flags = (fl.objects[id].flags & ~ray_kind) | ((uint32_t)o->Flags & ray_kind);
            }
            fl.frame.push_back((uint32_t)id);
            continue;
        }
        fl.frame.push_back((uint32_t)fl.add_object(o, -1));
    }
    for (const LightSource* l : sd->lightSources) fl.add_light(l);
    if (!sd->lightGroupLightSources.empty()) fl.unsupported("light_group");
    // sky_sphere and fog (atmosphere.h:81-117)
    pvgpu_sky_sphere sky{};
    if (sd->skysphere != nullptr) {
        sky.pigment_first = (uint32_t)fl.index_list.size();
        sky.pigment_count = (uint32_t)sd->skysphere->Pigments.size();
        vector<uint32_t> ids;
        for (const PIGMENT* pg : sd->skysphere->Pigments) ids.push_back((uint32_t)fl.add_pigment(pg));
        sky.pigment_first = (uint32_t)fl.index_list.size();
        fl.index_list.insert(fl.index_list.end(), ids.begin(), ids.end());
        sky.transform = fl.add_transform(sd->skysphere->Trans);
        for (int k = 0; k < 3; k++) sky.emission[k] = sd->skysphere->Emission[k];
    }
    vector<pvgpu_fog> fogs;
    for (const FOG* fog = sd->fog; fog != nullptr; fog = fog->Next) {
        pvgpu_fog f{};
        f.type = (uint32_t)fog->Type;
        f.turbulence = -1;
        if (fog->Turb != nullptr) {
            pvgpu_warp w{};
            w.type = PVGPU_WARP_TURBULENCE; w.transform = -1;
            for (int k = 0; k < 3; k++) w.turbulence[k] = fog->Turb->Turbulence[k];
            w.octaves = fog->Turb->Octaves; w.lambda = fog->Turb->Lambda; w.omega = fog->Turb->Omega;
            fl.warps.push_back(w);
            f.turbulence = (int32_t)fl.warps.size() - 1;
        }
        f.distance = fog->Distance; f.alt = fog->Alt; f.offset = fog->Offset;
        for (int k = 0; k < 3; k++) { f.up[k] = fog->Up[k]; f.colour[k] = fog->colour.colour()[k]; }
        f.colour[3] = fog->colour.filter(); f.colour[4] = fog->colour.transm();
        f.turb_depth = fog->Turb_Depth;
        fogs.push_back(f);
    }
    if (sd->boundingSlabs != nullptr)
        fl.flatten_tree(sd->boundingSlabs, fl.nodes, [&](const BBOX_TREE* leaf) {
            // (a leaf the flattener did not take - e.g. the object of a light source with looks_like - makes the scene unsupported,
            //  it must not end the render with an out_of_range exception: 28 distribution scenes used to "fail" that way)
            auto it = fl.object_ids.find(reinterpret_cast<const void*>(leaf->Node));
            if (it == fl.object_ids.end()) { fl.unsupported("bounding-tree leaf that is not a flattened object (light source geometry, looks_like)"); return 0u; }
            return (uint32_t)it->second;
        });
    if (gv.error.empty()) gv.error = fl.error;
    gv.object_ids = fl.object_ids;
    gv.mesh_tri_first = fl.mesh_tri_first;

    pvgpu_globals g{};
    g.max_trace_level = sd->parsedMaxTraceLevel;
    g.language_version = sd->EffectiveLanguageVersion();
    g.noise_generator = sd->noiseGenerator;
    g.bounding_method = (sd->boundingMethod == 1 && sd->boundingSlabs != nullptr) ? 1 : 0;
    g.quality_flags = quality_bits(vd->GetQualityFeatureFlags());
    g.output_alpha = sd->outputAlpha;
    g.adc_bailout = sd->parsedAdcBailout;
    for (int k = 0; k < 3; k++) { g.ambient_light[k] = sd->ambientLight[k]; g.background[k] = sd->backgroundColour.colour()[k]; }
    g.background[3] = sd->backgroundColour.filter(); g.background[4] = sd->backgroundColour.transm();
    g.atmosphere_ior = sd->atmosphereIOR; g.atmosphere_dispersion = sd->atmosphereDispersion;
    g.number_of_waves = sd->numberOfWaves;

    const Camera& cam = vd->GetCamera();
    pvgpu_camera c{};
    c.type = cam.Type;
    for (int k = 0; k < 3; k++) { c.location[k] = cam.Location[k]; c.direction[k] = cam.Direction[k]; c.up[k] = cam.Up[k]; c.right[k] = cam.Right[k]; }
    c.max_ray_distance = cam.Max_Ray_Distance;
    if (cam.Tnormal != nullptr) {
        // camera { normal { ... } }: the ray direction is perturbed like a surface normal (tracepixel.cpp:917-924)
        if (cam.Type > ORTHOGRAPHIC_CAMERA) gv.error = "camera normal perturbation on a non-planar camera";
        else {
            c.reserved = (uint32_t)fl.add_tnormal(cam.Tnormal) + 1u;
            if (gv.error.empty()) gv.error = fl.error;
        }
    }
    if ((cam.Aperture != 0.0) && (cam.Blur_Samples > 0)) gv.error = "focal blur";
    if (cam.Rays_Per_Pixel != 1) gv.error = "mesh camera";

    check(pvgpu_scene_create(&gv.scene, &g), "scene_create");
    check(pvgpu_scene_set_objects(gv.scene, fl.objects.data(), fl.objects.size(), fl.index_list.data(), fl.index_list.size(),
                                  fl.frame.data(), fl.frame.size()), "set_objects");
    check(pvgpu_scene_set_transforms(gv.scene, fl.transforms.data(), fl.transforms.size()), "set_transforms");
    check(pvgpu_scene_set_tree(gv.scene, fl.nodes.data(), fl.nodes.size()), "set_tree");
    check(pvgpu_scene_set_blobs(gv.scene, fl.blobs.data(), fl.blobs.size(), fl.blob_elements.data(), fl.blob_elements.size(),
                                fl.blob_nodes.data(), fl.blob_nodes.size()), "set_blobs");
    if (!fl.images.empty()) check(pvgpu_scene_set_images(gv.scene, fl.images.data(), fl.images.size(), fl.texels.data(), fl.texels.size()), "set_images");
    if (fl.any_mesh_uv) check(pvgpu_scene_set_mesh_uv(gv.scene, fl.mesh_uv.data(), fl.mesh_uv.size() / 2, fl.tri_uv.data(), fl.tri_uv.size() / 3), "set_mesh_uv");
    if (fl.any_blob_texture) check(pvgpu_scene_set_blob_textures(gv.scene, fl.blob_textures.data(), fl.blob_textures.size()), "set_blob_textures");
    check(pvgpu_scene_set_shape_data(gv.scene, fl.shape_data.data(), fl.shape_data.size()), "set_shape_data");
    check(pvgpu_scene_set_meshes(gv.scene, fl.meshes.data(), fl.meshes.size(), fl.vertices.data(), fl.vertices.size() / 3,
                                 fl.normals.data(), fl.normals.size() / 3, fl.triangles.data(), fl.triangles.size(),
                                 fl.mesh_nodes.data(), fl.mesh_nodes.size()), "set_meshes");
    check(pvgpu_scene_set_lights(gv.scene, fl.lights.data(), fl.lights.size()), "set_lights");
    check(pvgpu_scene_set_materials(gv.scene, fl.textures.data(), fl.textures.size(), fl.pigments.data(), fl.pigments.size(),
                                    fl.finishes.data(), fl.finishes.size(), fl.maps.data(), fl.maps.size(),
                                    fl.entries.data(), fl.entries.size(), fl.warps.data(), fl.warps.size(),
                                    fl.interiors.data(), fl.interiors.size()), "set_materials");
    check(pvgpu_scene_set_normals(gv.scene, fl.tnormals.data(), fl.tnormals.size(), fl.slope_entries.data(), fl.slope_entries.size()), "set_normals");
    check(pvgpu_scene_set_atmosphere(gv.scene, sd->skysphere != nullptr ? &sky : nullptr, fogs.data(), fogs.size()), "set_atmosphere");
    {
        const float wl[3] = { (float)sd->iridWavelengths[0], (float)sd->iridWavelengths[1], (float)sd->iridWavelengths[2] };
        bool any_irid = false;
        for (const pvgpu_finish& f : fl.finishes) if (f.irid > 0.0f) any_irid = true;
        if (any_irid) check(pvgpu_scene_set_irid_wavelengths(gv.scene, wl), "set_irid_wavelengths");
    }
    check(pvgpu_scene_set_camera(gv.scene, &c), "set_camera");
    if (cam.Type > ORTHOGRAPHIC_CAMERA) check(pvgpu_scene_set_camera_angles(gv.scene, cam.Angle, cam.H_Angle, cam.V_Angle), "set_camera_angles");
    if (const char* path = getenv("PVGPU_DUMP_SCENE")) check(pvgpu_scene_save(gv.scene, path), "scene_save");
    if (!gv.error.empty()) fprintf(stderr, "pvgpu adapter: scene uses a feature outside the GPU trace path: %s\n", gv.error.c_str());
    if (need_device) finalize_view(gv);
    return slot;
}

void write_file(const char* path, const void* data, size_t bytes)
{
    FILE* f = fopen(path, "wb");
    if (f == nullptr || fwrite(data, 1, bytes, f) != bytes) throw POV_EXCEPTION_STRING("pvgpu adapter: cannot write dump file");
    fclose(f);
}

#pragma pack(push, 1)
struct RayDumpRecord { double org[3], dir[3], depth; int32_t object, aux; };
#pragma pack(pop)

}  // namespace

// ------------------------------------------------------------------------------------------------
// TraceTask: same class (tracetask.h), GPU-backed implementation
// ------------------------------------------------------------------------------------------------
TraceTask::TraceTask(ViewData *vd, unsigned int tm, DBL js, DBL aat, DBL aac, unsigned int aad, pov_base::GammaCurvePtr& aag,
                     unsigned int ps, bool psc, bool contributesToImage, bool hr, size_t seed) :
    RenderTask(vd, seed, "Trace"),
    tracingMethod(tm), jitterScale(js), aaThreshold(aat), aaConfidence(aac), aaDepth(aad),
    previewSize(ps), previewSkipCorner(psc), passContributesToImage(contributesToImage),
    passCompletesImage((ps == 0) || ((ps == 1) && contributesToImage)), highReproducibility(hr), aaGamma(aag),
    trace(vd->GetSceneData(), &vd->GetCamera(), GetViewDataPtr(), vd->GetSceneData()->parsedMaxTraceLevel,
          vd->GetSceneData()->parsedAdcBailout, vd->GetQualityFeatureFlags(), cooperate, media, radiosity),
    cooperate(*this),
    media(GetViewDataPtr(), &trace, &photonGatherer),
    radiosity(vd->GetSceneData(), GetViewDataPtr(), vd->GetSceneData()->radiositySettings, vd->GetRadiosityCache(),
              cooperate, true, vd->GetCamera().Location),
    photonGatherer(&vd->GetSceneData()->mediaPhotonMap, vd->GetSceneData()->photonSettings)
{
    GetViewDataPtr()->qualityFlags = vd->GetQualityFeatureFlags();
}

TraceTask::~TraceTask() {}

// Known-answer vectors straight from the reference's functions (tests/golden/make_golden_probe.py).
// in:  u32 n_poly, n_poly x { i32 degree, i32 sturm, f64 epsilon, f64 c[5] }, u32 n_pts, n_pts x { f64 x, y, z, i32 generator, i32 octaves }
// out: n_poly x { i32 count, f64 roots[4] }, n_pts x { f64 noise, f64 dnoise[3], f64 turbulence (lambda 2, omega 0.5) }
static void run_probe(const char* in_path, const char* out_path, TraceThreadData* td)
{
    static std::mutex m;
    static bool done = false;
    std::lock_guard<std::mutex> lock(m);
    if (done) return;
    done = true;
    FILE* f = fopen(in_path, "rb");
    FILE* g = fopen(out_path, "wb");
    if (!f || !g) throw POV_EXCEPTION_STRING("pvgpu adapter: cannot open the probe files");
    uint32_t n = 0;
    if (fread(&n, 4, 1, f) != 1) n = 0;
    for (uint32_t i = 0; i < n; i++) {
        int32_t hdr[2]; double eps, c[5], r[4] = { 0, 0, 0, 0 };
        if (fread(hdr, 4, 2, f) != 2 || fread(&eps, 8, 1, f) != 1 || fread(c, 8, 5, f) != 5) break;
        int32_t cnt = Solve_Polynomial(hdr[0], c + (4 - hdr[0]), r, hdr[1], eps, td->Stats());
        for (int k = cnt; k < 4; k++) r[k] = 0.0;
        fwrite(&cnt, 4, 1, g); fwrite(r, 8, 4, g);
    }
    if (fread(&n, 4, 1, f) != 1) n = 0;
    for (uint32_t i = 0; i < n; i++) {
        double p[3]; int32_t gen[2];
        if (fread(p, 8, 3, f) != 3 || fread(gen, 4, 2, f) != 2) break;
        Vector3d P(p[0], p[1], p[2]), D;
        double out[5];
        out[0] = Noise(P, gen[0]);
        DNoise(D, P);
        out[1] = D[X]; out[2] = D[Y]; out[3] = D[Z];
        ClassicTurbulence tw(false);
        tw.Octaves = gen[1]; tw.Lambda = 2.0; tw.Omega = 0.5;
        out[4] = Turbulence(P, &tw, gen[0]);
        fwrite(out, 8, 5, g);
    }
    fclose(f); fclose(g);
}

void TraceTask::Run()
{
    if (getenv("PVGPU_PROBE_IN") && getenv("PVGPU_PROBE_OUT")) run_probe(getenv("PVGPU_PROBE_IN"), getenv("PVGPU_PROBE_OUT"), GetViewDataPtr());
    ViewData* vd = GetViewData();
    const unsigned int width = vd->GetWidth(), height = vd->GetHeight();
    const char* mode_env = getenv("PVGPU_RENDER");
    const bool use_gpu = !(mode_env && strcmp(mode_env, "stock") == 0);
    const char* dump_rays = getenv("PVGPU_DUMP_RAYS");
    const char* dump_rgbt = getenv("PVGPU_DUMP_RGBT");
    const char* dump_gpu = getenv("PVGPU_DUMP_GPU");

    if (previewSize > 0) {
        // mosaic preview passes (SimpleSamplingM0P) carry no final pixels; the final pass renders everything
        POVRect r; unsigned int serial;
        while (vd->GetNextRectangle(r, serial)) { vector<RGBTColour> px(r.GetArea()); vd->CompletedRectangle(r, serial, px, 1, false, passCompletesImage); Cooperate(); }
        return;
    }
    // anti-aliasing: methods 1 and 2 run on the device; the aaGamma curve must be neutral or a power law
    double aa_decoding_gamma = 1.0;
    if (use_gpu && tracingMethod != 0) {
        if (tracingMethod > 2)
            throw POV_EXCEPTION_STRING("pvgpu: sampling method 3 (stochastic supersampling) is outside the GPU trace path; use +AM1 or +AM2");
        if (!GammaCurve::IsNeutral(aaGamma)) {
            if (dynamic_cast<PowerLawGammaCurve*>(aaGamma.get()) == nullptr)
                throw POV_EXCEPTION_STRING("pvgpu: Antialias_Gamma combined with a non-power-law working gamma is outside the GPU trace path");
            aa_decoding_gamma = aaGamma->ApproximateDecodingGamma();
        }
    }

    std::shared_ptr<GpuView> gv = flatten_scene(vd, use_gpu);

    // drain the tile queue
    vector<POVRect> rects;
    vector<unsigned int> serials;
    {
        // one task takes the whole queue (the device renders all tiles in one call); tasks arriving later find it empty
        std::lock_guard<std::mutex> lock(g_drain_mutex);
        POVRect r; unsigned int serial;
        while (vd->GetNextRectangle(r, serial)) { rects.push_back(r); serials.push_back(serial); }
    }
    if (rects.empty()) return;
    Cooperate();

    vector<float> frame_ref, frame_gpu;
    vector<RayDumpRecord> frame_rays;
    if (dump_rgbt) frame_ref.assign((size_t)width * height * 4, 0.0f);
    if (dump_gpu) frame_gpu.assign((size_t)width * height * 4, 0.0f);
    if (dump_rays) frame_rays.assign((size_t)width * height, RayDumpRecord{});

    // page-locked frame buffer from the library: pvgpu_render then needs no staging copy
    std::unique_ptr<float, void (*)(float*)> gpu_pixels(nullptr, [](float* p) { pvgpu_host_free(p); });
    if (use_gpu) {
        vector<pvgpu_rect> pr(rects.size());
        size_t total = 0;
        for (size_t i = 0; i < rects.size(); i++) {
            pr[i].left = (int32_t)rects[i].left; pr[i].top = (int32_t)rects[i].top;
            pr[i].right = (int32_t)rects[i].right; pr[i].bottom = (int32_t)rects[i].bottom;
            total += rects[i].GetArea();
        }
        gpu_pixels.reset(static_cast<float*>(pvgpu_host_alloc(total * 4 * sizeof(float))));
        if (!gpu_pixels) throw POV_EXCEPTION_STRING("pvgpu: out of page-locked host memory");
        pvgpu_aa aa{};
        aa.method = tracingMethod;
        aa.depth = aaDepth;
        aa.threshold = aaThreshold;
        aa.jitter_scale = jitterScale;
        aa.gamma = aa_decoding_gamma;
        pvgpu_stats st{};
        check(pvgpu_render(gv->scene, &aa, (int)width, (int)height, pr.data(), pr.size(), gpu_pixels.get(), &st, nullptr, nullptr), "render");
        // the stock statistics page keeps working: counters the device kept in the reference's units
        GetViewDataPtr()->Stats()[Number_Of_Rays] += st.rays;
        GetViewDataPtr()->Stats()[Shadow_Ray_Tests] += st.shadow_ray_tests;
        GetViewDataPtr()->Stats()[Reflected_Rays_Traced] += st.reflected_rays;
        GetViewDataPtr()->Stats()[Refracted_Rays_Traced] += st.refracted_rays;
        GetViewDataPtr()->Stats()[Transmitted_Rays_Traced] += st.transmitted_rays;
        GetViewDataPtr()->Stats()[ADC_Saves] += st.adc_saves;
        GetViewDataPtr()->Stats()[Number_Of_Samples] += st.samples;
        vd->SetHighestTraceLevel(st.max_trace_level);
    }

    size_t cursor = 0;
    for (size_t ri = 0; ri < rects.size(); ri++) {
        const POVRect& rect = rects[ri];
        vector<RGBTColour> pixels;
        pixels.reserve(rect.GetArea());
        for (unsigned int y = rect.top; y <= rect.bottom; y++) {
            for (unsigned int x = rect.left; x <= rect.right; x++, cursor++) {
                const size_t fi = (size_t)y * width + x;
                RGBTColour ref;
                if (!use_gpu || dump_rgbt) {
                    trace(DBL(x) + 0.5, DBL(y) + 0.5, width, height, ref);     // the reference's own TracePixel
                    if (dump_rgbt) { frame_ref[4 * fi] = ref.red(); frame_ref[4 * fi + 1] = ref.green(); frame_ref[4 * fi + 2] = ref.blue(); frame_ref[4 * fi + 3] = ref.transm(); }
                }
                if (dump_rays) {
                    TraceTicket ticket(GetSceneData()->parsedMaxTraceLevel, GetSceneData()->parsedAdcBailout, GetSceneData()->outputAlpha);
                    Ray ray(ticket);
                    RayDumpRecord& rec = frame_rays[fi];
                    rec.object = -1; rec.aux = 0; rec.depth = BOUND_HUGE;
                    if (trace.CreateCameraRay(ray, DBL(x) + 0.5, DBL(y) + 0.5, width, height, 0)) {
                        Intersection isect;
                        NoSomethingFlagRayObjectCondition precond;
                        TrueRayObjectCondition postcond;
                        for (int k = 0; k < 3; k++) { rec.org[k] = ray.Origin[k]; rec.dir[k] = ray.Direction[k]; }
                        if (trace.FindIntersection(isect, ray, precond, postcond)) {
                            auto it = gv->object_ids.find(isect.Object);
                            rec.object = (it == gv->object_ids.end()) ? -2 : it->second;
                            rec.depth = isect.Depth;
                            if (Mesh* m = dynamic_cast<Mesh*>(isect.Object)) {      // index into the flattened scene's triangle table
                                auto mf = gv->mesh_tri_first.find(isect.Object);
                                rec.aux = (int32_t)(reinterpret_cast<const MESH_TRIANGLE*>(isect.Pointer) - m->Data->Triangles) + (int32_t)(mf == gv->mesh_tri_first.end() ? 0u : mf->second);
                            }
                            else if (dynamic_cast<TrueType*>(isect.Object) || dynamic_cast<Prism*>(isect.Object)) rec.aux = -1;      // a glyph hit carries its normal, a prism hit a spline parameter: not comparable
                            else rec.aux = isect.i1;
                        }
                    }
                }
                if (use_gpu) {
                    const float* g = gpu_pixels.get() + 4 * cursor;
                    pixels.push_back(RGBTColour(g[0], g[1], g[2], g[3]));
                    if (dump_gpu) memcpy(&frame_gpu[4 * fi], g, 4 * sizeof(float));
                } else pixels.push_back(ref);
                GetViewDataPtr()->Stats()[Number_Of_Pixels]++;
            }
            Cooperate();
        }
        GetViewDataPtr()->AfterTile();
        vd->CompletedRectangle(rect, serials[ri], pixels, 1, passContributesToImage, passCompletesImage);
    }
    if (dump_rgbt) write_file(dump_rgbt, frame_ref.data(), frame_ref.size() * sizeof(float));
    if (dump_gpu) write_file(dump_gpu, frame_gpu.data(), frame_gpu.size() * sizeof(float));
    if (dump_rays) write_file(dump_rays, frame_rays.data(), frame_rays.size() * sizeof(RayDumpRecord));
    if (!use_gpu) vd->SetHighestTraceLevel(trace.GetHighestTraceLevel());
}

void TraceTask::Stopped() {}

void TraceTask::Finish()
{
    GetViewDataPtr()->timeType = TraceThreadData::kRenderTime;
    GetViewDataPtr()->realTime = ConsumedRealTime();
    GetViewDataPtr()->cpuTime = ConsumedCPUTime();
}

}  // namespace pov
