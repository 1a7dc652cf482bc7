// Shadow rays: Trace::TraceShadowRay -> TracePointLightShadowRay (trace.cpp:1892-2076), ComputeShadowColour /
// ComputeShadowTexture (trace.cpp:2274-2439, 1181-1262).  Needs both the traversal and the pigment evaluation.
#pragma once
#include "pv_traverse.cuh"
#include "pv_shade.cuh"

namespace pvgpu {

// ---- shadow rays ----------------------------------------------------------------------------------
// Trace::ComputeShadowTexture (trace.cpp:1181-1262) for one plain (layered) texture evaluated at `epoint`: filter colour x fade
__device__ inline void shadow_texture(const DScene& sc, const pvgpu_object& ob, const pvgpu_interior* in, const Hit& hit, const V3& dir, const V3& rawnormal,
                                      bool inside_now, int32_t tex0, const V3& epoint, const V3& uvp, const TexLeaf* leaf, float tc[3])
{
    float tmp[3] = { 1.0f, 1.0f, 1.0f };
    for (int32_t li = tex0; li >= 0; li = sc.textures[li].next) {
        float lc[5];
        if (compute_pigment(sc, sc.textures[li].pigment, epoint, lc, uvp)) {        // (not found: outside an image_map used `once`, trace.cpp:1198-1205)
            #pragma unroll
            for (int k = 0; k < 3; k++) tmp[k] *= (lc[k] * lc[3] + lc[4]);
        }
        if (in && in->caustics != 0.0f) {
            V3 layer_normal = rawnormal;
#if PV_FULL_MATERIALS
            if ((sc.g.quality_flags & PVGPU_Q_NORMALS) && sc.textures[li].tnormal >= 0) {       // trace.cpp:1208-1227
                layer_normal = warp_normal_chain(sc, leaf, layer_normal, false);
                layer_normal = perturb_normal(sc, sc.textures[li].tnormal, layer_normal, epoint);
                if (sc.tnormals[sc.textures[li].tnormal].flags & PVGPU_DONT_SCALE_BUMPS_FLAG) layer_normal = normalized(layer_normal);
                layer_normal = warp_normal_chain(sc, leaf, layer_normal, true);
            }
#endif
            double dotval = dot(layer_normal, dir);
            float kk = (float)(1.0 + pow(fabs(dotval), (double)in->caustics));
            tmp[0] *= kk; tmp[1] *= kk; tmp[2] *= kk;
        }
    }
    float refr[3] = { 1.0f, 1.0f, 1.0f };
    if (in && inside_now && in->fade_power > 0.0f && fabs((double)in->fade_distance) > PV_EPSILON) {
        if (in->fade_power >= 1000.0f) {
            #pragma unroll
            for (int k = 0; k < 3; k++) refr[k] *= expf((float)(-(1.0 - (double)in->fade_colour[k]) * (hit.depth / (double)in->fade_distance)));
        } else {
            double kk = 1.0 + pow(hit.depth / (double)in->fade_distance, (double)in->fade_power);
            #pragma unroll
            for (int k = 0; k < 3; k++) refr[k] *= (float)((double)in->fade_colour[k] + (1.0 - (double)in->fade_colour[k]) / kk);
        }
    }
    tc[0] = tmp[0] * refr[0]; tc[1] = tmp[1] * refr[1]; tc[2] = tmp[2] * refr[2];
}


// Trace::ComputeShadowTexture (trace.cpp:1181-1262) for a transparent blocker.
__device__ inline void shadow_filter(const DScene& sc, const Hit& hit, const V3& org, const V3& dir, const PRay* parent, bool inside_now, float f[3])
{
    const pvgpu_object& ob = sc.objs[hit.obj];
    V3 rawnormal = object_normal(sc, ob, hit, org, dir);
    if (ob.flags & PVGPU_INVERTED_FLAG) rawnormal = -rawnormal;
    const double nd = dot(rawnormal, dir);
    if (nd > 0.0) rawnormal = -rawnormal;
    float tc[3];
    const pvgpu_interior* in = (ob.interior >= 0) ? &sc.interiors[ob.interior] : nullptr;
    V3 tex_ip, uvp;
    texture_points(sc, ob, hit, tex_ip, uvp);       // (ComputeShadowColour switches to UV like ComputeTextureColour, trace.cpp:2351-2362)
#if PV_FULL_MATERIALS
    if (ob.type == PVGPU_OBJ_BLOB && (ob.flags & PVGPU_MULTITEXTURE_FLAG) && sc.blob_textures != nullptr) {
        // Blob::Determine_Textures: weighted sum of the components' filter colours (trace.cpp:2382, 2399-2417)
        int32_t tex[PV_MAX_TEX_LEAVES];
        float wt[PV_MAX_TEX_LEAVES];
        const int nt = blob_weighted_textures(sc, ob, hit.ip, tex, wt, nullptr);
        TexLeaf leaves[PV_MAX_TEX_LEAVES];
        int n = 0;
        for (int i = 0; i < nt; i++) if (tex[i] >= 0 && !((double)wt[i] < sc.g.adc_bailout)) n = resolve_texture(sc, tex[i], hit.ip, leaves, n, (double)wt[i]);
        tc[0] = tc[1] = tc[2] = 0.0f;
        for (int i = 0; i < n; i++) {
            float t1[3];
            shadow_texture(sc, ob, in, hit, dir, rawnormal, inside_now, leaves[i].tex, leaves[i].p, uvp, &leaves[i], t1);
            #pragma unroll
            for (int k = 0; k < 3; k++) tc[k] += (float)((double)t1[k] * leaves[i].w);
        }
        if (fabsf((fabsf(tc[0]) + fabsf(tc[1]) + fabsf(tc[2])) / 3.0f) < (float)sc.g.adc_bailout) { f[0] = f[1] = f[2] = 0.0f; return; }
        f[0] *= tc[0]; f[1] *= tc[1]; f[2] *= tc[2];
        return;
    }
#endif
    const int32_t tex0 = hit_texture(sc, ob, hit, nd > 0.0);
    if (tex0 < 0) return;       // texture list empty: colour unchanged (trace.cpp:2391-2397)
#if PV_FULL_MATERIALS
    if (sc.textures[tex0].type != PVGPU_PAT_PLAIN) {
        // texture_map: weighted sum of the leaves' filter colours (ComputeOneTextureColour with shadowflag, trace.cpp:671-692)
        TexLeaf leaves[PV_MAX_TEX_LEAVES];
        const int n = resolve_texture(sc, tex0, tex_ip, leaves);
        tc[0] = tc[1] = tc[2] = 0.0f;
        for (int i = 0; i < n; i++) {
            float t1[3];
            shadow_texture(sc, ob, in, hit, dir, rawnormal, inside_now, leaves[i].tex, leaves[i].p, uvp, &leaves[i], t1);
            #pragma unroll
            for (int k = 0; k < 3; k++) tc[k] += (float)((double)t1[k] * leaves[i].w);
        }
    } else
#endif
    shadow_texture(sc, ob, in, hit, dir, rawnormal, inside_now, tex0, tex_ip, uvp, nullptr, tc);
    // ComputeShadowColour: "close enough to full shadow" (trace.cpp:2419-2424)
    if (fabsf((fabsf(tc[0]) + fabsf(tc[1]) + fabsf(tc[2])) / 3.0f) < (float)sc.g.adc_bailout) { f[0] = f[1] = f[2] = 0.0f; return; }
    f[0] *= tc[0]; f[1] *= tc[1]; f[2] *= tc[2];
}

// Trace::TraceShadowRay -> TracePointLightShadowRay (trace.cpp:1892-2076) without the (result-neutral)
// shadow caches.  Returns the factor the light colour is multiplied with and adds the number of
// FindIntersection calls (Shadow_Ray_Tests) to `tests`.
//   ALL_OPAQUE: every shadow caster of the scene has OPAQUE_FLAG, so the first blocker inside the shadow
//   window ends the search (any-hit) and no filtering code is needed.
template <bool ALL_OPAQUE>
__device__ inline void trace_shadow(bool alive, const DScene& sc, V3 o, const V3& d, double depth, const PRay* wave, uint32_t parent,
                                    TStack stack, Counters* cnt, float f[3], unsigned long long& tests, TravCount& tc)
{
    // all 32 lanes of the warp are here together (the traversal is warp-synchronous); `alive` lanes carry a shadow ray
    f[0] = f[1] = f[2] = 1.0f;
    if (ALL_OPAQUE) {
        Hit best;
        best.depth = depth;
        best.obj = PV_NO_OBJECT;
        if (alive) tests++;
        const bool found = find_intersection_sync<true>(alive, sc, o, d, 0u, true, PV_SMALL_TOLERANCE, best, stack, &cnt->overflow, tc, depth - PV_SHADOW_TOLERANCE);
        if (found && (best.depth < depth - PV_SHADOW_TOLERANCE) && (depth - best.depth > 0.0) && (best.depth > PV_SHADOW_TOLERANCE))
            f[0] = f[1] = f[2] = 0.0f;     // ComputeShadowColour: full shadow (trace.cpp:2318-2323)
        return;
    }
    // interiors of the light ray start as a copy of the eye ray's (Ray lightsourceray(eye), trace.cpp:1642)
    PRay in_state;
    in_state.n_int = 0;
    bool have_state = false;
    for (int iter = 0; iter < 256; iter++) {
        if (!vote_any(alive)) break;
        Hit best;
        best.depth = depth;
        best.obj = PV_NO_OBJECT;
        if (alive) tests++;
        const bool found = find_intersection_sync<false>(alive, sc, o, d, 0u, true, PV_SMALL_TOLERANCE, best, stack, &cnt->overflow, tc);
        if (!alive) continue;
        if (!(found && (best.depth < depth - PV_SHADOW_TOLERANCE) && (depth - best.depth > 0.0) && (best.depth > PV_SHADOW_TOLERANCE))) { alive = false; continue; }
        const pvgpu_object& ob = sc.objs[best.obj];
        if (ob.flags & PVGPU_OPAQUE_FLAG) { f[0] = f[1] = f[2] = 0.0f; alive = false; continue; }     // ComputeShadowColour: full shadow (trace.cpp:2318-2323)
        if (!have_state) { in_state = wave[parent]; have_state = true; }
        shadow_filter(sc, best, o, d, &in_state, ob.interior >= 0 && ray_is_interior(in_state, ob.interior), f);
        // ComputeShadowMedia (trace.cpp:3046-3071) toggles the blocker's interior on the light ray
        if (ob.interior >= 0) {
            if (!ray_remove_interior(in_state, ob.interior)) ray_append_interior(in_state, ob.interior, &cnt->overflow);
        }
        if (f[0] == 0.0f && f[1] == 0.0f && f[2] == 0.0f) {
            // colour is black; the reference keeps looping only to find an opaque object for its cache
            alive = false;
            continue;
        }
        depth -= best.depth;
        o = best.ip;
    }
}

}  // namespace pvgpu
