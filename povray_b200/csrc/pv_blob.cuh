// Blob (source/core/shape/blob.cpp): intervals of influence of the components along the ray
// (determine_influences :1269-1377, intersect_* :716-1267, insert_hit :616-714), the summed quartic per
// interval with the Bezier-hull rejection and Solve_Polynomial (All_Intersections :239-610), the field value
// for Inside (:1379-1624) and the gradient normal (:1677-1934).  The component coefficients c[], transforms and the
// bounding-sphere tree come from the reference's parser (Blob::Make_Blob / build_bounding_hierarchy) through the
// tables of include/pvgpu.h; tree walks keep the reference's LIFO order so that sums and tie orders are the same.
//
// The per-ray scratch (interval list, per-component coefficients, tree queue) lives in the local frame of the
// __noinline__ entry points, so kernels only pay for it while a blob is being tested.
#pragma once
#include "pv_shapes.cuh"

namespace pvgpu {

#define PV_BLOB_DEPTH_TOL     1.0e-2    // DEPTH_TOLERANCE    blob.cpp:135
#define PV_BLOB_INSIDE_TOL    1.0e-6    // INSIDE_TOLERANCE   blob.cpp:138
#define PV_BLOB_MAX_ACTIVE    32        // components one ray can be influenced by
#define PV_BLOB_QUEUE         64        // bounding-sphere tree queue

struct BlobInterval { double bound; uint32_t type; uint32_t elem; };     // Blob_Interval_Struct (blob.h:126-131); type bit 0: 0 entering, 1 exiting

// intersect_sphere / _ellipsoid / _hemisphere / _cylinder (blob.cpp:716-1174) behind intersect_element (:1176-1267)
__device__ inline bool blob_intersect_element(const DScene& sc, const pvgpu_blob_element& e, const V3& P, const V3& D, double mindist,
                                              double& tmin, double& tmax)
{
    tmin = PV_BOUND_HUGE;
    tmax = -PV_BOUND_HUGE;
    if (e.type == PVGPU_BLOB_SPHERE) {
        const V3 v1 = P - ld3(e.o);
        const double b = dot(v1, D), t = length_sqr(v1);
        double d = b * b - t + e.rad2;
        if (d < PV_EPSILON) return false;
        d = sqrt(d);
        tmax = -b + d; if (tmax < mindist) tmax = 0.0;
        tmin = -b - d; if (tmin < mindist) tmin = 0.0;
        if (tmax == tmin) return false;
        if (tmax < tmin) { d = tmin; tmin = tmax; tmax = d; }
        return true;
    }
    const pvgpu_transform& tr = sc.xf[e.transform];
    V3 PP = inv_trans_point(tr, P);
    V3 DD = inv_trans_direction(tr, D);
    const double len = length(DD);
    DD = DD / len;
    if (e.type == PVGPU_BLOB_ELLIPSOID) {
        const V3 v1 = PP - ld3(e.o);
        const double b = dot(v1, DD), t = length_sqr(v1);
        double d = b * b - t + e.rad2;
        if (d < PV_EPSILON) return false;
        d = sqrt(d);
        tmax = (-b + d) / len; if (tmax < mindist) tmax = 0.0;
        tmin = (-b - d) / len; if (tmin < mindist) tmin = 0.0;
        if (tmax == tmin) return false;
        if (tmax < tmin) { d = tmin; tmin = tmax; tmax = d; }
        return true;
    }
    if (e.type == PVGPU_BLOB_BASE_HEMISPHERE || e.type == PVGPU_BLOB_APEX_HEMISPHERE) {
        const bool base = e.type == PVGPU_BLOB_BASE_HEMISPHERE;
        if (!base) PP.z -= e.len;
        const double b = dot(PP, DD);
        double t = length_sqr(PP);
        double d = b * b - t + e.rad2;
        if (d < PV_EPSILON) return false;
        d = sqrt(d);
        tmax = -b + d;
        tmin = -b - d;
        if (tmax < tmin) { d = tmin; tmin = tmax; tmax = d; }
        // cut the intersection at the plane
        const double z1 = PP.z + tmin * DD.z, z2 = PP.z + tmax * DD.z;
        if (base ? ((z1 >= 0.0) && (z2 >= 0.0)) : ((z1 <= 0.0) && (z2 <= 0.0))) return false;
        if (base ? ((z1 < 0.0) && (z2 < 0.0)) : ((z1 > 0.0) && (z2 > 0.0))) { tmin /= len; tmax /= len; return true; }
        t = -PP.z / DD.z;
        if (base ? (z1 >= 0.0) : (z1 <= 0.0)) tmin = (t < mindist) ? 0.0 : t;     // crossing the plane from inside to outside
        else tmax = (t < mindist) ? 0.0 : t;                                       // from outside to inside
        tmin /= len;
        tmax /= len;
        return true;
    }
    // cylinder
    const double a = DD.x * DD.x + DD.y * DD.y;
    if (a > PV_EPSILON) {
        const double b = PP.x * DD.x + PP.y * DD.y;
        const double c = PP.x * PP.x + PP.y * PP.y - e.rad2;
        double d = b * b - a * c;
        if (d > PV_EPSILON) {
            d = sqrt(d);
            double t = (-b + d) / a;
            double w = PP.z + t * DD.z;
            if ((w >= 0.0) && (w <= e.len)) { if (t < tmin) tmin = t; if (t > tmax) tmax = t; }
            t = (-b - d) / a;
            w = PP.z + t * DD.z;
            if ((w >= 0.0) && (w <= e.len)) { if (t < tmin) tmin = t; if (t > tmax) tmax = t; }
        }
    }
    if (fabs(DD.z) > PV_EPSILON) {
        double t = -PP.z / DD.z;
        double u = PP.x + t * DD.x, v = PP.y + t * DD.y;
        if ((u * u + v * v) <= e.rad2) { if (t < tmin) tmin = t; if (t > tmax) tmax = t; }
        t = (e.len - PP.z) / DD.z;
        u = PP.x + t * DD.x; v = PP.y + t * DD.y;
        if ((u * u + v * v) <= e.rad2) { if (t < tmin) tmin = t; if (t > tmax) tmax = t; }
    }
    tmin /= len;
    tmax /= len;
    if (tmin < mindist) tmin = 0.0;
    if (tmax < mindist) tmax = 0.0;
    return !(tmin >= tmax);
}

// insert_hit (blob.cpp:616-714): sorted insertion of the entering and the exiting bound; iv[cnt] serves as the sentinel
__device__ inline void blob_insert_hit(uint32_t elem, uint32_t type, double t0, double t1, BlobInterval* iv, uint32_t& cnt)
{
    uint32_t k;
    iv[cnt].type = type & ~1u; iv[cnt].bound = t0; iv[cnt].elem = elem;
    for (k = 0; t0 > iv[k].bound; k++);
    if (k < cnt) {
        for (uint32_t m = cnt; m > k; m--) iv[m] = iv[m - 1];
        iv[k].type = type & ~1u; iv[k].bound = t0; iv[k].elem = elem;
        cnt++;
        iv[cnt].type = type | 1u; iv[cnt].bound = t1; iv[cnt].elem = elem;
        for (k = k + 1; t1 > iv[k].bound; k++);
        if (k < cnt) {
            for (uint32_t m = cnt; m > k; m--) iv[m] = iv[m - 1];
            iv[k].type = type | 1u; iv[k].bound = t1; iv[k].elem = elem;
        }
        cnt++;
    } else {
        cnt++;
        iv[cnt].type = type | 1u; iv[cnt].bound = t1; iv[cnt].elem = elem;
        cnt++;
    }
}

// Blob::All_Intersections (blob.cpp:239-610).  A blob that is not a CSG child stops after the first interval that produced a
// hit (blob.cpp:596-599: every further intersection is further away).  A CSG child has to report them all: the caller passes
// `resume` (start with 0) and calls again while it comes back >= 0 - every call reports the hits of the next interval that has
// any (at most four, the roots of one quartic).  Returns false when a per-ray capacity was exceeded.
static __device__ __noinline__ bool blob_hits(const DScene& sc, const pvgpu_object& ob, const V3& o, const V3& d, PrimHits& h, int* resume = nullptr)
{
    h.n = 0;
    const uint32_t start = resume ? (uint32_t)*resume : 0u;
    if (resume) *resume = -1;
    const pvgpu_blob& bl = sc.blobs[ob.mesh];
    const pvgpu_blob_element* el = sc.blob_elements + bl.element_first;
    V3 P = o, D = d;
    double len = 1.0;
    if (ob.transform >= 0) {
        const pvgpu_transform& t = sc.xf[ob.transform];
        P = inv_trans_point(t, o);
        D = inv_trans_direction(t, d);
        len = length(D);
        D = D / len;
    }
    BlobInterval iv[2 * PV_BLOB_MAX_ACTIVE + 2];
    uint32_t cnt = 0;
    bool fits = true;
    // determine_influences (blob.cpp:1269-1377)
    if (bl.node_count == 0) {
        for (uint32_t i = 0; i < bl.element_count; i++) {
            double t0, t1;
            if (blob_intersect_element(sc, el[i], P, D, PV_BLOB_DEPTH_TOL, t0, t1)) {
                if (cnt + 2 > 2 * PV_BLOB_MAX_ACTIVE) { fits = false; break; }
                blob_insert_hit(i, el[i].type, t0, t1, iv, cnt);
            }
        }
    } else {
        const pvgpu_blob_node* nodes = sc.blob_nodes + bl.node_first;
        uint32_t queue[PV_BLOB_QUEUE];
        uint32_t size = 0;
        queue[size++] = 0u;
        while (size > 0 && fits) {
            const pvgpu_blob_node nd = nodes[queue[--size]];
            if (nd.count == 0) {
                double t0, t1;
                if (blob_intersect_element(sc, el[nd.first], P, D, PV_BLOB_DEPTH_TOL, t0, t1)) {
                    if (cnt + 2 > 2 * PV_BLOB_MAX_ACTIVE) { fits = false; break; }
                    blob_insert_hit(nd.first, el[nd.first].type, t0, t1, iv, cnt);
                }
            } else {
                for (uint32_t i = 0; i < nd.count; i++) {
                    const pvgpu_blob_node& ch = nodes[nd.first + i];
                    const V3 v1 = ld3(ch.c) - P;
                    const double b = dot(v1, D), t = length_sqr(v1);
                    if ((t - sqr(b)) <= ch.r2) {
                        if (size >= PV_BLOB_QUEUE) { fits = false; break; }
                        queue[size++] = nd.first + i;
                    }
                }
            }
        }
    }
    if (!fits) return false;
    if (cnt == 0) return true;

    // to avoid numerical problems start at the first interval and scale the direction (blob.cpp:283-316)
    double start_dist = iv[0].bound;
    if (start_dist < PV_SMALL_TOLERANCE) start_dist = 0.0;
    for (uint32_t i = 0; i < cnt; i++) iv[i].bound -= start_dist;
    P = P + start_dist * D;
    double max_bound = iv[0].bound;
    for (uint32_t i = 0; i < cnt; i++) if (iv[i].bound > max_bound) max_bound = iv[i].bound;
    if (max_bound != 0) {
        D = D * max_bound;
        for (uint32_t i = 0; i < cnt; i++) iv[i].bound /= max_bound;
    } else max_bound = 1;

    double coeffs[5] = { 0.0, 0.0, 0.0, 0.0, -bl.threshold };
    double fco[PV_BLOB_MAX_ACTIVE][5];      // Blob_Coefficients of the components met so far
    uint32_t fel[PV_BLOB_MAX_ACTIVE];
    uint32_t nf = 0;
    int in_flag = 0;
    bool found = false;
    for (uint32_t i = 0; i < cnt; i++) {
        if ((iv[i].type & 1u) == 0u) {
            in_flag++;
            const pvgpu_blob_element& e = el[iv[i].elem];
            double t0, t1, t2;
            if (e.type == PVGPU_BLOB_SPHERE) {
                const V3 v1 = P - ld3(e.o);
                t0 = length_sqr(v1); t1 = dot(v1, D); t2 = max_bound * max_bound;
            } else {
                const pvgpu_transform& tr = sc.xf[e.transform];
                V3 PP = inv_trans_point(tr, P);
                const V3 DD = inv_trans_direction(tr, D);
                if (e.type == PVGPU_BLOB_ELLIPSOID) {
                    const V3 v1 = PP - ld3(e.o);
                    t0 = length_sqr(v1); t1 = dot(v1, DD); t2 = length_sqr(DD);
                } else if (e.type == PVGPU_BLOB_CYLINDER) {
                    t0 = PP.x * PP.x + PP.y * PP.y; t1 = PP.x * DD.x + PP.y * DD.y; t2 = DD.x * DD.x + DD.y * DD.y;
                } else {
                    if (e.type == PVGPU_BLOB_APEX_HEMISPHERE) PP.z -= e.len;
                    t0 = length_sqr(PP); t1 = dot(PP, DD); t2 = length_sqr(DD);
                }
            }
            const double c0 = e.c[0], c1 = e.c[1], c2 = e.c[2];
            double* f = fco[nf];
            fel[nf] = iv[i].elem;
            nf++;
            f[0] = c0 * t2 * t2;
            f[1] = 4.0 * c0 * t1 * t2;
            f[2] = 2.0 * c0 * (2.0 * t1 * t1 + t0 * t2) + c1 * t2;
            f[3] = 2.0 * t1 * (2.0 * c0 * t0 + c1);
            f[4] = t0 * (c0 * t0 + c1) + c2;
            for (int j = 0; j < 5; j++) coeffs[j] += f[j];
        } else {
            uint32_t k = 0;
            while (k < nf && fel[k] != iv[i].elem) k++;
            for (int j = 0; j < 5; j++) coeffs[j] -= fco[k][j];
            if (--in_flag == 0) continue;
        }
        // next bound (almost) at the same place: add / subtract it first (blob.cpp:497-500)
        if ((i + 1 < cnt) && (fabs(iv[i].bound - iv[i + 1].bound) < PV_EPSILON)) continue;
        if (i < start) continue;          // intervals already reported to the caller (their coefficients are accounted for above)
        // move the interval to [0, 1] and test the convex hull of the Bezier form (blob.cpp:507-537)
        const double l = iv[i].bound, w = iv[i + 1].bound - l;
        double nc[5], dk[5];
        nc[0] = coeffs[0] * w * w * w * w;
        nc[1] = (coeffs[1] + 4.0 * coeffs[0] * l) * w * w * w;
        nc[2] = (3.0 * l * (2.0 * coeffs[0] * l + coeffs[1]) + coeffs[2]) * w * w;
        nc[3] = (2.0 * l * (2.0 * l * (coeffs[0] * l + 0.75 * coeffs[1]) + coeffs[2]) + coeffs[3]) * w;
        nc[4] = l * (l * (l * (coeffs[0] * l + coeffs[1]) + coeffs[2]) + coeffs[3]) + coeffs[4];
        dk[0] = nc[4];
        dk[1] = nc[4] + 0.25 * nc[3];
        dk[2] = nc[4] + 0.50 * (nc[3] + nc[2] / 3.0);
        dk[3] = nc[4] + 0.50 * (1.5 * nc[3] + nc[2] + 0.5 * nc[1]);
        dk[4] = nc[4] + nc[3] + nc[2] + nc[1] + nc[0];
        if (((dk[0] >= 0.0) && (dk[1] >= 0.0) && (dk[2] >= 0.0) && (dk[3] >= 0.0) && (dk[4] >= 0.0)) ||
            ((dk[0] <= 0.0) && (dk[1] <= 0.0) && (dk[2] <= 0.0) && (dk[3] <= 0.0) && (dk[4] <= 0.0))) continue;
        double roots[4];
        const int nr = solve_polynomial(4, coeffs, roots, (ob.flags & PVGPU_STURM_FLAG) ? 1 : 0, 1.0e-11);
        for (int j = 0; j < nr; j++) {
            double dist = roots[j];
            if ((dist >= iv[i].bound) && (dist <= iv[i + 1].bound)) {
                dist = (dist * max_bound + start_dist) / len;
                if ((dist > PV_BLOB_DEPTH_TOL) && (dist < PV_MAX_DISTANCE) && h.n < 4) {
                    h.depth[h.n] = dist; h.ip[h.n] = evaluate(o, d, dist); h.aux[h.n] = 0; h.n++;
                    found = true;
                }
            }
        }
        if (found) {
            if (resume) *resume = (int)i + 1;       // CSG child: the caller comes back for the intervals behind this one
            break;                                  // not a CSG child: every further intersection is further away (blob.cpp:596-599)
        }
    }
    return true;
}

// calculate_element_field (blob.cpp:1379-1500)
__device__ inline double blob_element_field(const DScene& sc, const pvgpu_blob_element& e, const V3& P)
{
    double rad2;
    if (e.type == PVGPU_BLOB_SPHERE) {
        rad2 = length_sqr(P - ld3(e.o));
        return (rad2 < e.rad2) ? rad2 * (rad2 * e.c[0] + e.c[1]) + e.c[2] : 0.0;
    }
    V3 PP = inv_trans_point(sc.xf[e.transform], P);
    if (e.type == PVGPU_BLOB_ELLIPSOID) {
        rad2 = length_sqr(PP - ld3(e.o));
        return (rad2 < e.rad2) ? rad2 * (rad2 * e.c[0] + e.c[1]) + e.c[2] : 0.0;
    }
    if (e.type == PVGPU_BLOB_BASE_HEMISPHERE) {
        if (!(PP.z <= 0.0)) return 0.0;
        rad2 = length_sqr(PP);
        return (rad2 <= e.rad2) ? rad2 * (rad2 * e.c[0] + e.c[1]) + e.c[2] : 0.0;
    }
    if (e.type == PVGPU_BLOB_APEX_HEMISPHERE) {
        PP.z -= e.len;
        if (!(PP.z >= 0.0)) return 0.0;
        rad2 = length_sqr(PP);
        return (rad2 <= e.rad2) ? rad2 * (rad2 * e.c[0] + e.c[1]) + e.c[2] : 0.0;
    }
    if ((PP.z >= 0.0) && (PP.z <= e.len)) {
        rad2 = sqr(PP.x) + sqr(PP.y);
        if (rad2 <= e.rad2) return rad2 * (rad2 * e.c[0] + e.c[1]) + e.c[2];
    }
    return 0.0;
}

// Blob::Inside via calculate_field_value (blob.cpp:1502-1624)
static __device__ __noinline__ bool blob_inside(const DScene& sc, const pvgpu_object& ob, const V3& p)
{
    const pvgpu_blob& bl = sc.blobs[ob.mesh];
    const pvgpu_blob_element* el = sc.blob_elements + bl.element_first;
    const V3 P = (ob.transform >= 0) ? inv_trans_point(sc.xf[ob.transform], p) : p;
    double density = 0.0;
    if (bl.node_count == 0) {
        for (uint32_t i = 0; i < bl.element_count; i++) density += blob_element_field(sc, el[i], P);
    } else {
        const pvgpu_blob_node* nodes = sc.blob_nodes + bl.node_first;
        uint32_t queue[PV_BLOB_QUEUE];
        uint32_t size = 0;
        queue[size++] = 0u;
        while (size > 0) {
            const pvgpu_blob_node nd = nodes[queue[--size]];
            if (nd.count == 0) density += blob_element_field(sc, el[nd.first], P);
            else for (uint32_t i = 0; i < nd.count; i++) {
                const pvgpu_blob_node& ch = nodes[nd.first + i];
                if (length_sqr(P - ld3(ch.c)) <= ch.r2 && size < PV_BLOB_QUEUE) queue[size++] = nd.first + i;
            }
        }
    }
    const bool inside = density > bl.threshold - PV_BLOB_INSIDE_TOL;
    return (ob.flags & PVGPU_INVERTED_FLAG) ? !inside : inside;
}

// element_normal (blob.cpp:1677-1813)
__device__ inline void blob_element_normal(const DScene& sc, const pvgpu_blob_element& e, const V3& P, V3& result)
{
    if (e.type == PVGPU_BLOB_SPHERE) {
        const V3 v1 = P - ld3(e.o);
        const double dist = length_sqr(v1);
        if (dist <= e.rad2) { const double val = -2.0 * e.c[0] * dist - e.c[1]; result = result + val * v1; }
        return;
    }
    const pvgpu_transform& tr = sc.xf[e.transform];
    V3 PP = inv_trans_point(tr, P);
    if (e.type == PVGPU_BLOB_ELLIPSOID) {
        V3 v1 = PP - ld3(e.o);
        const double dist = length_sqr(v1);
        if (dist <= e.rad2) { const double val = -2.0 * e.c[0] * dist - e.c[1]; v1 = trans_normal(tr, v1); result = result + val * v1; }
        return;
    }
    if (e.type == PVGPU_BLOB_BASE_HEMISPHERE || e.type == PVGPU_BLOB_APEX_HEMISPHERE) {
        if (e.type == PVGPU_BLOB_APEX_HEMISPHERE) { PP.z -= e.len; if (!(PP.z >= 0.0)) return; }
        else if (!(PP.z <= 0.0)) return;
        const double dist = length_sqr(PP);
        if (dist <= e.rad2) { const double val = -2.0 * e.c[0] * dist - e.c[1]; PP = trans_normal(tr, PP); result = result + val * PP; }
        return;
    }
    if ((PP.z >= 0.0) && (PP.z <= e.len)) {
        const double dist = sqr(PP.x) + sqr(PP.y);
        if (dist <= e.rad2) { const double val = -2.0 * e.c[0] * dist - e.c[1]; PP.z = 0.0; PP = trans_normal(tr, PP); result = result + val * PP; }
    }
}

// Blob::Normal (blob.cpp:1815-1933)
static __device__ __noinline__ V3 blob_normal(const DScene& sc, const pvgpu_object& ob, const V3& ip)
{
    const pvgpu_blob& bl = sc.blobs[ob.mesh];
    const pvgpu_blob_element* el = sc.blob_elements + bl.element_first;
    const V3 P = (ob.transform >= 0) ? inv_trans_point(sc.xf[ob.transform], ip) : ip;
    V3 result = mk(0.0, 0.0, 0.0);
    if (bl.node_count == 0) {
        for (uint32_t i = 0; i < bl.element_count; i++) blob_element_normal(sc, el[i], P, result);
    } else {
        const pvgpu_blob_node* nodes = sc.blob_nodes + bl.node_first;
        uint32_t queue[PV_BLOB_QUEUE];
        uint32_t size = 0;
        queue[size++] = 0u;
        while (size > 0) {
            const pvgpu_blob_node nd = nodes[queue[--size]];
            if (nd.count == 0) blob_element_normal(sc, el[nd.first], P, result);
            else for (uint32_t i = 0; i < nd.count; i++) {
                const pvgpu_blob_node& ch = nodes[nd.first + i];
                if (length_sqr(P - ld3(ch.c)) <= ch.r2 && size < PV_BLOB_QUEUE) queue[size++] = nd.first + i;
            }
        }
    }
    double val = length_sqr(result);
    if (val == 0.0) result = mk(1.0, 0.0, 0.0);
    else { val = 1.0 / sqrt(val); result = result * val; }
    if (ob.transform >= 0) result = normalized(trans_normal(sc.xf[ob.transform], result));
    return result;
}

}  // namespace pvgpu
