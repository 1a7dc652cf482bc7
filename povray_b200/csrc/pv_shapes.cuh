// Per-primitive FP64 intersection / inside / normal device functions.  Each follows the control flow,
// tolerances and floating-point operation order of the reference method it cites, so that depths agree
// to the last bit wherever only +,-,*,/ and sqrt are involved.
#pragma once
#include "pv_math.cuh"
#include "pv_solver.cuh"

namespace pvgpu {

#define PV_SPHERE_DEPTH_TOL   1.0e-6   // sphere.cpp:63
#define PV_BOX_DEPTH_TOL      1.0e-6   // box.cpp:67
#define PV_BOX_CLOSE_TOL      1.0e-6   // box.cpp:70
#define PV_PLANE_DEPTH_TOL    1.0e-6   // plane.cpp:62
#define PV_QUADRIC_DEPTH_TOL  1.0e-6   // quadric.cpp:94
#define PV_MESH_DEPTH_TOL     1.0e-6   // mesh.cpp:92
#define PV_TORUS_DEPTH_TOL    1.0e-4   // torus.cpp:78
#define PV_TORUS_ROOT_TOL     1.0e-4   // torus.cpp:81

// Candidate hits of one primitive, in the order the reference pushes them on the IStack.
#if PV_HEAVY
#define PV_MAX_PRIM_HITS 4          // torus / blob quartics
#else
#define PV_MAX_PRIM_HITS 2          // sphere, box, plane
#endif
struct PrimHits {
    int    n;
    double depth[PV_MAX_PRIM_HITS];
    V3     ip[PV_MAX_PRIM_HITS];
    uint32_t aux[PV_MAX_PRIM_HITS];
};

// ---- sphere -------------------------------------------------------------------------------------
// Sphere::Intersect (sphere.cpp:211-243)
__device__ __forceinline__ bool sphere_intersect(const V3& o, const V3& d, const V3& center, double radius2,
                                                 double& depth1, double& depth2)
{
    V3 oc = center - o;
    double oc2 = length_sqr(oc);
    double tca = dot(oc, d);
    if ((oc2 >= radius2) && (tca < PV_EPSILON)) return false;
    double thc2 = radius2 - oc2 + sqr(tca);
    if (thc2 > PV_EPSILON) {
        double hc = sqrt(thc2);
        depth1 = tca - hc;
        depth2 = tca + hc;
        return true;
    }
    return false;
}

// Sphere::All_Intersections (sphere.cpp:92-170)
__device__ inline void sphere_hits(const DScene& sc, const pvgpu_object& ob, const V3& o, const V3& d, PrimHits& h)
{
    h.n = 0;
    double d1, d2;
    if (ob.aux) {   // Do_Ellipsoid: unit sphere in object space
        const pvgpu_transform& t = sc.xf[ob.transform];
        V3 no = inv_trans_point(t, o);
        V3 nd = inv_trans_direction(t, d);
        double len = length(nd);
        nd = nd / len;
        if (sphere_intersect(no, nd, mk(0.0, 0.0, 0.0), 1.0, d1, d2)) {
            if ((d1 > PV_SPHERE_DEPTH_TOL) && (d1 < PV_MAX_DISTANCE)) {
                h.depth[h.n] = d1 / len; h.ip[h.n] = trans_point(t, evaluate(no, nd, d1)); h.aux[h.n] = 0; h.n++;
            }
            if ((d2 > PV_SPHERE_DEPTH_TOL) && (d2 < PV_MAX_DISTANCE)) {
                h.depth[h.n] = d2 / len; h.ip[h.n] = trans_point(t, evaluate(no, nd, d2)); h.aux[h.n] = 0; h.n++;
            }
        }
    } else {
        if (sphere_intersect(o, d, ld3(ob.p), sqr(ob.p[3]), d1, d2)) {
            if ((d1 > PV_SPHERE_DEPTH_TOL) && (d1 < PV_MAX_DISTANCE)) {
                h.depth[h.n] = d1; h.ip[h.n] = evaluate(o, d, d1); h.aux[h.n] = 0; h.n++;
            }
            if ((d2 > PV_SPHERE_DEPTH_TOL) && (d2 < PV_MAX_DISTANCE)) {
                h.depth[h.n] = d2; h.ip[h.n] = evaluate(o, d, d2); h.aux[h.n] = 0; h.n++;
            }
        }
    }
}

// Sphere::Inside (sphere.cpp:262-300)
__device__ inline bool sphere_inside(const DScene& sc, const pvgpu_object& ob, const V3& p)
{
    double oc2;
    if (ob.aux) oc2 = length_sqr(inv_trans_point(sc.xf[ob.transform], p));
    else        oc2 = length_sqr(ld3(ob.p) - p);
    return (ob.flags & PVGPU_INVERTED_FLAG) ? (oc2 > sqr(ob.p[3])) : (oc2 < sqr(ob.p[3]));
}

// Sphere::Normal (sphere.cpp:318-340)
__device__ inline V3 sphere_normal(const DScene& sc, const pvgpu_object& ob, const V3& ip)
{
    if (ob.aux) {
        const pvgpu_transform& t = sc.xf[ob.transform];
        return normalized(trans_normal(t, inv_trans_point(t, ip)));
    }
    return (ip - ld3(ob.p)) / ob.p[3];
}

// ---- plane --------------------------------------------------------------------------------------
// Plane::Intersect + All_Intersections (plane.cpp:92-190)
__device__ inline void plane_hits(const DScene& sc, const pvgpu_object& ob, const V3& o, const V3& d, PrimHits& h)
{
    h.n = 0;
    V3 n = ld3(ob.p);
    double ndd, ndo;
    if (ob.transform < 0) {
        ndd = dot(n, d);
        if (fabs(ndd) < PV_EPSILON) return;
        ndo = dot(n, o);
    } else {
        const pvgpu_transform& t = sc.xf[ob.transform];
        V3 P = inv_trans_point(t, o);
        V3 D = inv_trans_direction(t, d);
        ndd = dot(n, D);
        if (fabs(ndd) < PV_EPSILON) return;
        ndo = dot(n, P);
    }
    double depth = -(ndo + ob.p[3]) / ndd;
    if ((depth >= PV_PLANE_DEPTH_TOL) && (depth <= PV_MAX_DISTANCE)) {
        h.depth[0] = depth; h.ip[0] = evaluate(o, d, depth); h.aux[0] = 0; h.n = 1;
    }
}

// Plane::Inside (plane.cpp:208-225)
__device__ inline bool plane_inside(const DScene& sc, const pvgpu_object& ob, const V3& p)
{
    double temp;
    if (ob.transform < 0) temp = dot(p, ld3(ob.p));
    else                  temp = dot(inv_trans_point(sc.xf[ob.transform], p), ld3(ob.p));
    return (temp + ob.p[3]) < PV_EPSILON;
}

// Plane::Normal (plane.cpp:243-253)
__device__ inline V3 plane_normal(const DScene& sc, const pvgpu_object& ob)
{
    V3 n = ld3(ob.p);
    if (ob.transform >= 0) n = normalized(trans_normal(sc.xf[ob.transform], n));
    return n;
}

// ---- box ----------------------------------------------------------------------------------------
// Box::Intersect (box.cpp:167-520).  The reference unrolls the three axes; the near/far bookkeeping is
// identical per axis except that X uses no CLOSE_TOLERANCE and Y / Z break near-ties by comparing the
// direction component with that of the axis of the side currently recorded (Y always against X).
__device__ inline bool box_intersect(const V3& P, const V3& D, const double* c1, const double* c2,
                                     double& depth1, double& depth2, int& side1, int& side2)
{
    int smin = 0, smax = 0;
    double tmin = 0.0, tmax = PV_BOUND_HUGE, t;
    // X
    if (D.x < -PV_EPSILON) {
        t = (c1[0] - P.x) / D.x;
        if (t < tmin) return false;
        if (t <= tmax) { smax = 1; tmax = t; }
        t = (c2[0] - P.x) / D.x;
        if (t >= tmin) { if (t > tmax) return false; smin = 2; tmin = t; }
    } else if (D.x > PV_EPSILON) {
        t = (c2[0] - P.x) / D.x;
        if (t < tmin) return false;
        if (t <= tmax) { smax = 2; tmax = t; }
        t = (c1[0] - P.x) / D.x;
        if (t >= tmin) { if (t > tmax) return false; smin = 1; tmin = t; }
    } else if ((P.x < c1[0]) || (P.x > c2[0])) return false;

    // Y and Z share one shape
    #pragma unroll
    for (int ax = 1; ax < 3; ax++) {
        const double Da = (ax == 1) ? D.y : D.z, Pa = (ax == 1) ? P.y : P.z;
        const int s0 = 2 * ax + 1, s1 = 2 * ax + 2;       // kSideHit_{Y0,Y1} = 3,4 ; {Z0,Z1} = 5,6
        if (Da < -PV_EPSILON || Da > PV_EPSILON) {
            const bool neg = Da < -PV_EPSILON;
            const double far_c = neg ? c1[ax] : c2[ax], near_c = neg ? c2[ax] : c1[ax];
            const int far_s = neg ? s0 : s1, near_s = neg ? s1 : s0;
            const double mag = neg ? -Da : Da;
            t = (far_c - Pa) / Da;
            if (t < tmin) return false;
            if (t <= tmax - PV_BOX_CLOSE_TOL) { smax = far_s; tmax = t; }
            else if (t <= tmax + PV_BOX_CLOSE_TOL) {
                if (ax == 1) { if (mag > fabs(D.x)) smax = far_s; }
                else if (smax == 1 || smax == 2) { if (mag > fabs(D.x)) smax = far_s; }
                else if (smax == 3 || smax == 4) { if (mag > fabs(D.y)) smax = far_s; }
            }
            t = (near_c - Pa) / Da;
            if (t >= tmin + PV_BOX_CLOSE_TOL) { if (t > tmax) return false; smin = near_s; tmin = t; }
            else if (t >= tmin - PV_BOX_CLOSE_TOL) {
                if (ax == 1) { if (mag > fabs(D.x)) smin = near_s; }
                else if (smin == 1 || smin == 2) { if (mag > fabs(D.x)) smin = near_s; }
                else if (smin == 3 || smin == 4) { if (mag > fabs(D.y)) smin = near_s; }
            }
        } else if ((Pa < c1[ax]) || (Pa > c2[ax])) return false;
    }
    if (tmax < PV_BOX_DEPTH_TOL) return false;
    depth1 = tmin; depth2 = tmax; side1 = smin; side2 = smax;
    return true;
}

// Box::All_Intersections (box.cpp:100-149)
__device__ inline void box_hits(const DScene& sc, const pvgpu_object& ob, const V3& o, const V3& d, PrimHits& h)
{
    h.n = 0;
    V3 P = o, D = d;
    if (ob.transform >= 0) {
        const pvgpu_transform& t = sc.xf[ob.transform];
        P = inv_trans_point(t, o);
        D = inv_trans_direction(t, d);
    }
    double d1, d2; int s1, s2;
    if (box_intersect(P, D, ob.p, ob.p + 3, d1, d2, s1, s2)) {
        if (d1 > PV_BOX_DEPTH_TOL) { h.depth[h.n] = d1; h.ip[h.n] = evaluate(o, d, d1); h.aux[h.n] = (uint32_t)s1; h.n++; }
        h.depth[h.n] = d2; h.ip[h.n] = evaluate(o, d, d2); h.aux[h.n] = (uint32_t)s2; h.n++;
    }
}

// Box::Inside (box.cpp:538-575)
__device__ inline bool box_inside(const DScene& sc, const pvgpu_object& ob, const V3& p)
{
    V3 q = (ob.transform >= 0) ? inv_trans_point(sc.xf[ob.transform], p) : p;
    const bool inv = (ob.flags & PVGPU_INVERTED_FLAG) != 0;
    if ((q.x < ob.p[0]) || (q.x > ob.p[3])) return inv;
    if ((q.y < ob.p[1]) || (q.y > ob.p[4])) return inv;
    if ((q.z < ob.p[2]) || (q.z > ob.p[5])) return inv;
    return !inv;
}

// Box::Normal (box.cpp:600-620)
__device__ inline V3 box_normal(const DScene& sc, const pvgpu_object& ob, uint32_t side)
{
    V3 n = mk(0.0, 0.0, 0.0);
    switch (side) {
        case 1: n.x = -1.0; break;
        case 2: n.x = 1.0; break;
        case 3: n.y = -1.0; break;
        case 4: n.y = 1.0; break;
        case 5: n.z = -1.0; break;
        case 6: n.z = 1.0; break;
    }
    if (ob.transform >= 0) n = normalized(trans_normal(sc.xf[ob.transform], n));
    return n;
}

// ---- quadric ------------------------------------------------------------------------------------
// Quadric::Intersect + All_Intersections (quadric.cpp:123-230); coefficient names as in quadric.cpp:80-92
__device__ inline void quadric_hits(const pvgpu_object& ob, const V3& o, const V3& d, PrimHits& h)
{
    h.n = 0;
    const double QA = ob.p[0], QE = ob.p[1], QH = ob.p[2], QB = ob.p[3], QC = ob.p[4], QF = ob.p[5],
                 QD = ob.p[6], QG = ob.p[7], QI = ob.p[8], QJ = ob.p[9];
    const double Xo = o.x, Yo = o.y, Zo = o.z, Xd = d.x, Yd = d.y, Zd = d.z;
    double a = Xd * (QA * Xd + QB * Yd + QC * Zd) + Yd * (QE * Yd + QF * Zd) + Zd * QH * Zd;
    double b = Xd * (QA * Xo + 0.5 * (QB * Yo + QC * Zo + QD)) +
               Yd * (QE * Yo + 0.5 * (QB * Xo + QF * Zo + QG)) +
               Zd * (QH * Zo + 0.5 * (QC * Xo + QF * Yo + QI));
    double c = Xo * (QA * Xo + QB * Yo + QC * Zo + QD) + Yo * (QE * Yo + QF * Zo + QG) + Zo * (QH * Zo + QI) + QJ;
    double d1, d2;
    if (a != 0.0) {
        double disc = sqr(b) - a * c;
        if (disc <= 0.0) return;
        disc = sqrt(disc);
        d1 = (-b + disc) / a;
        d2 = (-b - disc) / a;
    } else {
        if (b == 0.0) return;
        d1 = -0.5 * c / b;
        d2 = PV_MAX_DISTANCE;
    }
    if ((d1 > PV_QUADRIC_DEPTH_TOL) && (d1 < PV_MAX_DISTANCE)) { h.depth[h.n] = d1; h.ip[h.n] = evaluate(o, d, d1); h.aux[h.n] = 0; h.n++; }
    if ((d2 > PV_QUADRIC_DEPTH_TOL) && (d2 < PV_MAX_DISTANCE)) { h.depth[h.n] = d2; h.ip[h.n] = evaluate(o, d, d2); h.aux[h.n] = 0; h.n++; }
}

// Quadric::Inside (quadric.cpp:248-255)
__device__ inline bool quadric_inside(const pvgpu_object& ob, const V3& p)
{
    const double QA = ob.p[0], QE = ob.p[1], QH = ob.p[2], QB = ob.p[3], QC = ob.p[4], QF = ob.p[5],
                 QD = ob.p[6], QG = ob.p[7], QI = ob.p[8], QJ = ob.p[9];
    return (p.x * (QA * p.x + QB * p.y + QD) + p.y * (QE * p.y + QF * p.z + QG) + p.z * (QH * p.z + QC * p.x + QI) + QJ) <= 0.0;
}

// Quadric::Normal (quadric.cpp:273-310)
__device__ inline V3 quadric_normal(const pvgpu_object& ob, const V3& ip)
{
    const double QA = ob.p[0], QE = ob.p[1], QH = ob.p[2], QB = ob.p[3], QC = ob.p[4], QF = ob.p[5],
                 QD = ob.p[6], QG = ob.p[7], QI = ob.p[8];
    V3 n = mk(2.0 * QA * ip.x + QB * ip.y + QC * ip.z + QD,
              QB * ip.x + 2.0 * QE * ip.y + QF * ip.z + QG,
              QC * ip.x + QF * ip.y + 2.0 * QH * ip.z + QI);
    double len = length(n);
    if (len == 0.0) return mk(1.0, 0.0, 0.0);
    return n / len;
}

// ---- disc ------------------------------------------------------------------------------------------
// Disc::Intersect / All_Intersections / Inside / Normal (disc.cpp:90-230): the plane z = 0 of the disc's space
__device__ inline void disc_hits(const DScene& sc, const pvgpu_object& ob, const V3& o, const V3& d, PrimHits& h)
{
    h.n = 0;
    const pvgpu_transform& tr = sc.xf[ob.transform];
    const V3 P = inv_trans_point(tr, o);
    V3 D = inv_trans_direction(tr, d);
    const double len = length(D);
    D = D / len;
    if (fabs(D.z) > PV_EPSILON) {
        const double t = -P.z / D.z;
        if (t >= 0.0) {
            const double u = P.x + t * D.x, v = P.y + t * D.y, r2 = sqr(u) + sqr(v);
            if ((r2 >= ob.p[3]) && (r2 <= ob.p[4])) {
                const double depth = t / len;
                if ((depth > 1.0e-6) && (depth < PV_MAX_DISTANCE)) { h.depth[0] = depth; h.ip[0] = evaluate(o, d, depth); h.aux[0] = 0; h.n = 1; }
            }
        }
    }
}
__device__ inline bool disc_inside(const DScene& sc, const pvgpu_object& ob, const V3& p)
{
    const bool inv = (ob.flags & PVGPU_INVERTED_FLAG) != 0;
    return (inv_trans_point(sc.xf[ob.transform], p).z >= 0.0) ? inv : !inv;
}

// ---- triangle / smooth_triangle ----------------------------------------------------------------------
// Triangle::Intersect / All_Intersections (triangle.cpp:447-590); record layout: see PVGPU_OBJ_TRIANGLE.
__device__ inline void triangle_hits(const DScene& sc, const pvgpu_object& ob, const V3& o, const V3& d, PrimHits& h)
{
    h.n = 0;
    if (ob.flags & PVGPU_DEGENERATE_FLAG) return;
    const double* T = sc.shape_data + ob.mesh;
    const V3 N = ld3(T + 9);
    const double ndd = dot(N, d);
    if (fabs(ndd) < PV_EPSILON) return;
    const double ndo = dot(N, o);
    const double depth = -(T[12] + ndo) / ndd;
    if ((depth < 1.0e-6) || (depth > PV_MAX_DISTANCE)) return;
    const int dom = (int)(ob.aux & 3u);
    const int a = (dom == 0) ? 1 : 0, b = (dom == 2) ? 1 : 2;       // X: (Y, Z)   Y: (X, Z)   Z: (X, Y)
    const double s = comp(o, a) + depth * comp(d, a);
    const double t = comp(o, b) + depth * comp(d, b);
    const double p1a = T[a], p1b = T[b], p2a = T[3 + a], p2b = T[3 + b], p3a = T[6 + a], p3b = T[6 + b];
    if ((p2a - s) * (p2b - p1b) < (p2b - t) * (p2a - p1a)) return;
    if ((p3a - s) * (p3b - p2b) < (p3b - t) * (p3a - p2a)) return;
    if ((p1a - s) * (p1b - p3b) < (p1b - t) * (p1a - p3a)) return;
    h.depth[0] = depth; h.ip[0] = evaluate(o, d, depth); h.aux[0] = 0; h.n = 1;
}
// Triangle::Normal / SmoothTriangle::Normal (triangle.cpp:640-700)
__device__ inline V3 triangle_normal(const DScene& sc, const pvgpu_object& ob, const V3& ip)
{
    const double* T = sc.shape_data + ob.mesh;
    if (!(ob.aux & PVGPU_TRIANGLE_SMOOTH)) return ld3(T + 9);
    const V3 P1 = ld3(T), N1 = ld3(T + 13), N2 = ld3(T + 16), N3 = ld3(T + 19), Perp = ld3(T + 22);
    const V3 pmp1 = ip - P1;
    const double u = dot(pmp1, Perp);
    if (u < PV_EPSILON) return N1;
    const int axis = (int)((ob.aux >> 2) & 3u);
    const double v = (comp(pmp1, axis) / u + T[axis] - T[3 + axis]) / (T[6 + axis] - T[3 + axis]);
    return normalized(N1 + u * (N2 - N1 + v * (N3 - N2)));
}

// ---- polygon -----------------------------------------------------------------------------------------
// Polygon::in_polygon (polygon.cpp:905-980): crossings test over the closed sub-polygons of the point list
__device__ inline bool in_polygon(int number, const double* pts, double tx, double ty)
{
    int v0 = 0, v1 = 1, first = 0;
    bool yflag0 = (pts[2 * v0 + 1] >= ty), inside_flag = false;
    for (int i = 1; i < number; ) {
        const bool yflag1 = (pts[2 * v1 + 1] >= ty);
        if (yflag0 != yflag1) {
            if (((pts[2 * v1 + 1] - ty) * (pts[2 * v0] - pts[2 * v1]) >= (pts[2 * v1] - tx) * (pts[2 * v0 + 1] - pts[2 * v1 + 1])) == yflag1)
                inside_flag = !inside_flag;
        }
        if ((i < number - 2) && (pts[2 * v1] == pts[2 * first]) && (pts[2 * v1 + 1] == pts[2 * first + 1])) {
            v0 = ++i; v1 = ++i;
            yflag0 = (pts[2 * v0 + 1] >= ty);
            first = v0;
        } else {
            v0 = v1; v1 = ++i;
            yflag0 = yflag1;
        }
    }
    return inside_flag;
}
// Polygon::Intersect / All_Intersections (polygon.cpp:131-260)
static __device__ __noinline__ void polygon_hits(const DScene& sc, const pvgpu_object& ob, const V3& o, const V3& d, PrimHits& h)
{
    h.n = 0;
    if (ob.flags & PVGPU_DEGENERATE_FLAG) return;
    const pvgpu_transform& tr = sc.xf[ob.transform];
    const V3 P = inv_trans_point(tr, o);
    V3 D = inv_trans_direction(tr, d);
    const double len = length(D);
    D = D / len;
    if (fabs(D.z) < 1.0e-10) return;                   // ZERO_TOLERANCE polygon.cpp:93
    double depth = -P.z / D.z;
    if ((depth < 1.0e-8) || (depth > PV_MAX_DISTANCE)) return;      // DEPTH_TOLERANCE polygon.cpp:90
    const double x = P.x + depth * D.x, y = P.y + depth * D.y;
    if (!in_polygon((int)ob.aux, sc.shape_data + ob.mesh, x, y)) return;
    depth /= len;
    h.depth[0] = depth; h.ip[0] = evaluate(o, d, depth); h.aux[0] = 0; h.n = 1;
}

// ---- poly / cubic / quartic (order <= 4) ------------------------------------------------------------
// Poly (polynomial.cpp:211-1245); coefficients in the shape-data table at ob.mesh, ob.aux = Order, transform required.
#define PV_POLY_DEPTH_TOLERANCE 1.0e-4      // DEPTH_TOLERANCE, INSIDE_TOLERANCE, ROOT_TOLERANCE  polynomial.cpp:96-98
__device__ inline int poly_quadratic(const double* a, const V3& O, const V3& D, double* depths)      // intersect_quadratic :852-947
{
    const double x = O.x, y = O.y, z = O.z, xx = D.x, yy = D.y, zz = D.z;
    const double x2 = x * x, y2 = y * y, z2 = z * z, xx2 = xx * xx, yy2 = yy * yy, zz2 = zz * zz;
    double ac = (a[0]*xx2 + a[1]*xx*yy + a[2]*xx*zz + a[4]*yy2 + a[5]*yy*zz + a[7]*zz2);
    double bc = (2*a[0]*x*xx + a[1]*(x*yy + xx*y) + a[2]*(x*zz + xx*z) +
                 a[3]*xx + 2*a[4]*y*yy + a[5]*(y*zz + yy*z) + a[6]*yy +
                 2*a[7]*z*zz + a[8]*zz);
    double cc = a[0]*x2 + a[1]*x*y + a[2]*x*z + a[3]*x + a[4]*y2 +
                a[5]*y*z + a[6]*y + a[7]*z2 + a[8]*z + a[9];
    if (fabs(ac) < 1.0e-20) {
        if (fabs(bc) < 1.0e-20) return 0;
        depths[0] = -cc / bc;
        return 1;
    }
    double dd = bc * bc - 4.0 * ac * cc;
    if (dd < 0.0) return 0;
    dd = sqrt(dd);
    bc = -bc;
    const double t = 2.0 * ac;
    depths[0] = (bc + dd) / t;
    depths[1] = (bc - dd) / t;
    return 2;
}
__device__ inline int poly_general(const double* a, int order, bool sturm, const V3& O, const V3& D, double* depths)
{
    // Poly::intersect (polynomial.cpp:656-800): substitute the ray into every term, collect powers of t
    double eqn_v[3][5], eqn_vt[3][5], eqn[5], tt[3][5];
    for (int i = 0; i < 3; i++) { eqn_v[i][0] = 1.0; eqn_vt[i][0] = 1.0; }
    eqn_v[0][1] = O.x; eqn_v[1][1] = O.y; eqn_v[2][1] = O.z;
    eqn_vt[0][1] = D.x; eqn_vt[1][1] = D.y; eqn_vt[2][1] = D.z;
    for (int i = 2; i <= order; i++)
        for (int j = 0; j < 3; j++) { eqn_v[j][i] = eqn_v[j][1] * eqn_v[j][i - 1]; eqn_vt[j][i] = eqn_vt[j][1] * eqn_vt[j][i - 1]; }
    for (int i = 0; i <= order; i++) eqn[i] = 0.0;
    const unsigned int binom[5][5] = { { 1, 0, 0, 0, 0 }, { 1, 1, 0, 0, 0 }, { 1, 2, 1, 0, 0 }, { 1, 3, 3, 1, 0 }, { 1, 4, 6, 4, 1 } };
    int term = 0;
    for (int i = order; i >= 0; i--) {
        for (int h = 0; h <= i; h++) tt[0][h] = binom[i][h] * eqn_vt[0][i - h] * eqn_v[0][h];
        for (int j = order - i; j >= 0; j--) {
            for (int h = 0; h <= j; h++) tt[1][h] = binom[j][h] * eqn_vt[1][j - h] * eqn_v[1][h];
            for (int k = order - (i + j); k >= 0; k--) {
                if (a[term] != 0) {
                    for (int h = 0; h <= k; h++) tt[2][h] = binom[k][h] * eqn_vt[2][k - h] * eqn_v[2][h];
                    const int offset = order - (i + j + k);
                    for (int i1 = 0; i1 <= i; i1++)
                        for (int j1 = 0; j1 <= j; j1++)
                            for (int k1 = 0; k1 <= k; k1++) {
                                double val = a[term];
                                val *= tt[0][i1];
                                val *= tt[1][j1];
                                val *= tt[2][k1];
                                eqn[offset + i1 + j1 + k1] += val;
                            }
                }
                term++;
            }
        }
    }
    int lead = 0, deg = order;
    for (; lead <= order; lead++) { if (eqn[lead] != 0.0) break; else deg--; }
    if (deg <= 1) return 0;
    return solve_polynomial(deg, &eqn[lead], depths, sturm ? 1 : 0, PV_POLY_DEPTH_TOLERANCE);
}
static __device__ __noinline__ void poly_hits(const DScene& sc, const pvgpu_object& ob, const V3& o, const V3& d, PrimHits& h)
{
    h.n = 0;
    const pvgpu_transform& tr = sc.xf[ob.transform];
    const double* a = sc.shape_data + ob.mesh;
    const int order = (int)ob.aux;
    const V3 O = inv_trans_point(tr, o);
    V3 D = inv_trans_direction(tr, d);
    const double len = length(D);
    D = D / len;
    double depths[4];
    int cnt;
    if (order == 1) {                                   // intersect_linear :804-850
        const double t0 = a[0] * O.x + a[1] * O.y + a[2] * O.z;
        const double t1 = a[0] * D.x + a[1] * D.y + a[2] * D.z;
        if (fabs(t1) < PV_EPSILON) return;
        depths[0] = -(a[3] + t0) / t1;
        cnt = 1;
    } else if (order == 2) cnt = poly_quadratic(a, O, D, depths);
    else cnt = poly_general(a, order, (ob.flags & PVGPU_STURM_FLAG) != 0, O, D, depths);
    for (int i = 0; i < cnt; i++) {
        if (!(depths[i] > PV_POLY_DEPTH_TOLERANCE)) continue;
        bool same_root = false;
        for (int j = 0; j < i; j++) if (depths[i] == depths[j]) { same_root = true; break; }
        if (same_root) continue;
        h.depth[h.n] = depths[i] / len;
        h.ip[h.n] = trans_point(tr, evaluate(O, D, depths[i]));
        h.aux[h.n] = 0;
        h.n++;
    }
}
static __device__ __noinline__ bool poly_inside(const DScene& sc, const pvgpu_object& ob, const V3& p)     // Poly::Inside + inside :590-654, 1131-1178
{
    const double* a = sc.shape_data + ob.mesh;
    const int order = (int)ob.aux;
    const V3 P = inv_trans_point(sc.xf[ob.transform], p);
    double xp[5], yp[5], zp[5];
    xp[0] = 1.0; yp[0] = 1.0; zp[0] = 1.0;
    xp[1] = P.x; yp[1] = P.y; zp[1] = P.z;
    for (int i = 2; i <= order; i++) { xp[i] = xp[1] * xp[i - 1]; yp[i] = yp[1] * yp[i - 1]; zp[i] = zp[1] * zp[i - 1]; }
    double result = 0.0;
    int term = 0;
    for (int i = order; i >= 0; i--)
        for (int j = order - i; j >= 0; j--)
            for (int k = order - (i + j); k >= 0; k--) {
                const double c = a[term];
                if (c != 0.0) result += c * xp[i] * yp[j] * zp[k];
                term++;
            }
    const bool inv = (ob.flags & PVGPU_INVERTED_FLAG) != 0;
    return (result < PV_POLY_DEPTH_TOLERANCE) ? !inv : inv;
}
static __device__ __noinline__ V3 poly_normal(const DScene& sc, const pvgpu_object& ob, const V3& ip)      // Poly::Normal + normal1 :1035-1129, 1180-1244
{
    const pvgpu_transform& tr = sc.xf[ob.transform];
    const double* a = sc.shape_data + ob.mesh;
    const int order = (int)ob.aux;
    const V3 P = inv_trans_point(tr, ip);
    const double x = P.x, y = P.y, z = P.z;
    double rx = 0.0, ry = 0.0, rz = 0.0;
    switch (order) {
        case 1: rx = a[0]; ry = a[1]; rz = a[2]; break;
        case 2:
            rx = 2*a[0]*x+a[1]*y+a[2]*z+a[3];
            ry = a[1]*x+2*a[4]*y+a[5]*z+a[6];
            rz = a[2]*x+a[5]*y+2*a[7]*z+a[8];
            break;
        case 3: {
            const double x2 = x * x, y2 = y * y, z2 = z * z;
            rx = 3*a[0]*x2 + 2*x*(a[1]*y + a[2]*z + a[3]) + a[4]*y2 + y*(a[5]*z + a[6]) + a[7]*z2 + a[8]*z + a[9];
            ry = a[1]*x2 + x*(2*a[4]*y + a[5]*z + a[6]) + 3*a[10]*y2 + 2*y*(a[11]*z + a[12]) + a[13]*z2 + a[14]*z + a[15];
            rz = a[2]*x2 + x*(a[5]*y + 2*a[7]*z + a[8]) + a[11]*y2 + y*(2*a[13]*z + a[14]) + 3*a[16]*z2 + 2*a[17]*z + a[18];
            break;
        }
        default: {
            const double x2 = x * x, y2 = y * y, z2 = z * z, x3 = x * x2, y3 = y * y2, z3 = z * z2;
            rx = 4*a[ 0]*x3+3*x2*(a[ 1]*y+a[ 2]*z+a[ 3])+
                 2*x*(a[ 4]*y2+y*(a[ 5]*z+a[ 6])+a[ 7]*z2+a[ 8]*z+a[ 9])+
                 a[10]*y3+y2*(a[11]*z+a[12])+y*(a[13]*z2+a[14]*z+a[15])+
                 a[16]*z3+a[17]*z2+a[18]*z+a[19];
            ry = a[ 1]*x3+x2*(2*a[ 4]*y+a[ 5]*z+a[ 6])+
                 x*(3*a[10]*y2+2*y*(a[11]*z+a[12])+a[13]*z2+a[14]*z+a[15])+
                 4*a[20]*y3+3*y2*(a[21]*z+a[22])+2*y*(a[23]*z2+a[24]*z+a[25])+
                 a[26]*z3+a[27]*z2+a[28]*z+a[29];
            rz = a[ 2]*x3+x2*(a[ 5]*y+2*a[ 7]*z+a[ 8])+
                 x*(a[11]*y2+y*(2*a[13]*z+a[14])+3*a[16]*z2+2*a[17]*z+a[18])+
                 a[21]*y3+y2*(2*a[23]*z+a[24])+y*(3*a[26]*z2+2*a[27]*z+a[28])+
                 4*a[30]*z3+3*a[31]*z2+2*a[32]*z+a[33];
        }
    }
    V3 r = trans_normal(tr, mk(rx, ry, rz));
    double val = dot(r, r);
    if (val > 0.0) { val = 1.0 / sqrt(val); return r * val; }
    return mk(1.0, 0.0, 0.0);
}

// ---- cone / cylinder ------------------------------------------------------------------------------
#define PV_CONE_TOLERANCE 1.0e-9      // Cone_Tolerance  cone.cpp:65
#define PV_CONE_BASE_HIT 1u           // cone.cpp:71-73
#define PV_CONE_CAP_HIT  2u
#define PV_CONE_SIDE_HIT 3u

// Cone::Intersect + All_Intersections (cone.cpp:103-330): canonical space, z in [dist, 1] (cylinder: [0, 1])
__device__ inline void cone_hits(const DScene& sc, const pvgpu_object& ob, const V3& o, const V3& d, PrimHits& h)
{
    h.n = 0;
    const pvgpu_transform& tr = sc.xf[ob.transform];
    const V3 P = inv_trans_point(tr, o);
    V3 D = inv_trans_direction(tr, d);
    const double len = length(D);
    D = D / len;
    const double dist = ob.p[0];
    const bool cyl = (ob.flags & PVGPU_CYLINDER_FLAG) != 0;
    auto push = [&](double t, uint32_t kind) { const double w = t / len; h.depth[h.n] = w; h.ip[h.n] = evaluate(o, d, w); h.aux[h.n] = kind; h.n++; };
    double a, b, c, dd, t1, t2, z;
    if (cyl) {
        a = D.x * D.x + D.y * D.y;
        if (a > PV_EPSILON) {
            b = P.x * D.x + P.y * D.y;
            c = P.x * P.x + P.y * P.y - 1.0;
            dd = b * b - a * c;
            if (dd >= 0.0) {
                dd = sqrt(dd);
                t1 = (-b + dd) / a;
                t2 = (-b - dd) / a;
                z = P.z + t1 * D.z;
                if ((t1 > PV_CONE_TOLERANCE) && (t1 < PV_MAX_DISTANCE) && (z >= 0.0) && (z <= 1.0)) push(t1, PV_CONE_SIDE_HIT);
                z = P.z + t2 * D.z;
                if ((t2 > PV_CONE_TOLERANCE) && (t2 < PV_MAX_DISTANCE) && (z >= 0.0) && (z <= 1.0)) push(t2, PV_CONE_SIDE_HIT);
            }
        }
    } else {
        a = D.x * D.x + D.y * D.y - D.z * D.z;
        b = D.x * P.x + D.y * P.y - D.z * P.z;
        c = P.x * P.x + P.y * P.y - P.z * P.z;
        if (fabs(a) < PV_EPSILON) {
            if (fabs(b) > PV_EPSILON) {
                t1 = -0.5 * c / b;
                z = P.z + t1 * D.z;
                if ((t1 > PV_CONE_TOLERANCE) && (t1 < PV_MAX_DISTANCE) && (z >= dist) && (z <= 1.0)) push(t1, PV_CONE_SIDE_HIT);
            }
        } else {
            dd = b * b - a * c;
            if (dd >= 0.0) {
                dd = sqrt(dd);
                t1 = (-b - dd) / a;
                t2 = (-b + dd) / a;
                z = P.z + t1 * D.z;
                if ((t1 > PV_CONE_TOLERANCE) && (t1 < PV_MAX_DISTANCE) && (z >= dist) && (z <= 1.0)) push(t1, PV_CONE_SIDE_HIT);
                z = P.z + t2 * D.z;
                if ((t2 > PV_CONE_TOLERANCE) && (t2 < PV_MAX_DISTANCE) && (z >= dist) && (z <= 1.0)) push(t2, PV_CONE_SIDE_HIT);
            }
        }
    }
    if ((ob.flags & PVGPU_CLOSED_FLAG) && (fabs(D.z) > PV_EPSILON)) {
        dd = (1.0 - P.z) / D.z;
        a = (P.x + dd * D.x);
        b = (P.y + dd * D.y);
        if (((sqr(a) + sqr(b)) <= 1.0) && (dd > PV_CONE_TOLERANCE) && (dd < PV_MAX_DISTANCE)) push(dd, PV_CONE_CAP_HIT);
        dd = (dist - P.z) / D.z;
        a = (P.x + dd * D.x);
        b = (P.y + dd * D.y);
        if ((sqr(a) + sqr(b)) <= (cyl ? 1.0 : sqr(dist)) && (dd > PV_CONE_TOLERANCE) && (dd < PV_MAX_DISTANCE)) push(dd, PV_CONE_BASE_HIT);
    }
}

// Cone::Inside (cone.cpp:333-390)
__device__ inline bool cone_inside(const DScene& sc, const pvgpu_object& ob, const V3& p)
{
    const double offset = (ob.flags & PVGPU_CLOSED_FLAG) ? -PV_EPSILON : PV_EPSILON;
    const V3 q = inv_trans_point(sc.xf[ob.transform], p);
    const double w2 = q.x * q.x + q.y * q.y;
    bool outside;
    if (ob.flags & PVGPU_CYLINDER_FLAG) outside = (w2 > 1.0 + offset) || (q.z < 0.0 - offset) || (q.z > 1.0 + offset);
    else outside = (w2 > q.z * q.z + offset) || (q.z < ob.p[0] - offset) || (q.z > 1.0 + offset);
    const bool inv = (ob.flags & PVGPU_INVERTED_FLAG) != 0;
    return outside ? inv : !inv;
}

// Cone::Normal (cone.cpp:408-445)
__device__ inline V3 cone_normal(const DScene& sc, const pvgpu_object& ob, const V3& ip, uint32_t kind)
{
    const pvgpu_transform& tr = sc.xf[ob.transform];
    V3 r = inv_trans_point(tr, ip);
    if (kind == PV_CONE_SIDE_HIT) { if (ob.flags & PVGPU_CYLINDER_FLAG) r.z = 0.0; else r.z = -r.z; }
    else if (kind == PV_CONE_BASE_HIT) r = mk(0.0, 0.0, -1.0);
    else if (kind == PV_CONE_CAP_HIT) r = mk(0.0, 0.0, 1.0);
    return normalized(trans_normal(tr, r));
}

// ---- torus --------------------------------------------------------------------------------------
// Torus::Test_Thick_Cylinder (torus.cpp:932-1059)
__device__ inline bool torus_thick_cylinder(const V3& P, const V3& D, double h1, double h2, double r1, double r2)
{
    double a, b, c, d, u, v, k, r, h;
    if (fabs(D.y) < PV_EPSILON) {
        if ((P.y < h1) || (P.y > h2)) return false;
    } else {
        k = (h2 - P.y) / D.y;
        u = P.x + k * D.x;
        v = P.z + k * D.z;
        if ((k > PV_EPSILON) && (k < PV_MAX_DISTANCE)) {
            r = u * u + v * v;
            if ((r >= r1) && (r <= r2)) return true;
        }
        k = (h1 - P.y) / D.y;
        u = P.x + k * D.x;
        v = P.z + k * D.z;
        if ((k > PV_EPSILON) && (k < PV_MAX_DISTANCE)) {
            r = u * u + v * v;
            if ((r >= r1) && (r <= r2)) return true;
        }
    }
    a = D.x * D.x + D.z * D.z;
    if (a > PV_EPSILON) {
        b = P.x * D.x + P.z * D.z;
        c = P.x * P.x + P.z * P.z - r2;
        d = b * b - a * c;
        if (d >= 0.0) {
            d = sqrt(d);
            k = (-b + d) / a;
            if ((k > PV_EPSILON) && (k < PV_MAX_DISTANCE)) { h = P.y + k * D.y; if ((h >= h1) && (h <= h2)) return true; }
            k = (-b - d) / a;
            if ((k > PV_EPSILON) && (k < PV_MAX_DISTANCE)) { h = P.y + k * D.y; if ((h >= h1) && (h <= h2)) return true; }
        }
        c = P.x * P.x + P.z * P.z - r1;
        d = b * b - a * c;
        if (d >= 0.0) {
            d = sqrt(d);
            k = (-b + d) / a;
            if ((k > PV_EPSILON) && (k < PV_MAX_DISTANCE)) { h = P.y + k * D.y; if ((h >= h1) && (h <= h2)) return true; }
            k = (-b - d) / a;
            if ((k > PV_EPSILON) && (k < PV_MAX_DISTANCE)) { h = P.y + k * D.y; if ((h >= h1) && (h <= h2)) return true; }
        }
    }
    return false;
}

// Torus::Intersect + All_Intersections (torus.cpp:133-160, 229-330); SpindleTorus filter (:162-210)
__device__ inline void torus_hits(const DScene& sc, const pvgpu_object& ob, const V3& o, const V3& d, PrimHits& h)
{
    h.n = 0;
    const pvgpu_transform& t = sc.xf[ob.transform];
    const double R = ob.p[0], rr = ob.p[1];
    V3 P = inv_trans_point(t, o);
    V3 D = inv_trans_direction(t, d);
    double len = length(D);
    D = D / len;
    double y1 = -rr, y2 = rr;
    double r1 = sqr(R - rr);
    if (R < rr) r1 = 0;
    double r2 = sqr(R + rr);
    if (!torus_thick_cylinder(P, D, y1, y2, r1, r2)) return;
    double bsr = R + rr + rr;
    double distP = length_sqr(P);
    double closer = 0.0;
    if (distP > sqr(bsr)) {
        distP = sqrt(distP);
        closer = distP - bsr;
        P = P + closer * D;
    }
    double R2 = sqr(R);
    r2 = sqr(rr);
    double Py2 = P.y * P.y, Dy2 = D.y * D.y, PDy2 = P.y * D.y;
    double k1 = P.x * P.x + P.z * P.z + Py2 - R2 - r2;
    double k2 = P.x * D.x + P.z * D.z + PDy2;
    double c[5], r[4];
    c[0] = 1.0;
    c[1] = 4.0 * k2;
    c[2] = 2.0 * (k1 + 2.0 * (k2 * k2 + R2 * Dy2));
    c[3] = 4.0 * (k2 * k1 + 2.0 * R2 * PDy2);
    c[4] = k1 * k1 + 4.0 * R2 * (Py2 - r2);
    int n = solve_polynomial(4, c, r, (ob.flags & PVGPU_STURM_FLAG) ? 1 : 0, PV_TORUS_ROOT_TOL);
    while (n--) {
        double depth = (r[n] + closer) / len;
        if ((depth > PV_TORUS_DEPTH_TOL) && (depth < PV_MAX_DISTANCE)) {
            V3 ip = evaluate(o, d, depth);
            uint32_t aux = 0;
            if (ob.aux) {                          // SpindleTorus: keep the hit only on the visible part
                V3 lp = inv_trans_point(t, ip);
                bool on_spindle = (length_sqr(lp) < ob.p[2]);
                bool valid = on_spindle ? (ob.aux & 0x01u) : (ob.aux & 0x02u);   // SpindleVisible / NonSpindleVisible
                if (!valid) continue;
                aux = on_spindle ? 1u : 0u;
            }
            h.depth[h.n] = depth; h.ip[h.n] = ip; h.aux[h.n] = aux; h.n++;
        }
    }
}

// Torus::Inside (torus.cpp:348-365); spindle variant (:367-400)
__device__ inline bool torus_inside(const DScene& sc, const pvgpu_object& ob, const V3& p)
{
    V3 P = inv_trans_point(sc.xf[ob.transform], p);
    double r = sqrt(sqr(P.x) + sqr(P.z));
    double r2 = sqr(P.y) + sqr(r - ob.p[0]);
    bool inside;
    if (r2 <= sqr(ob.p[1])) {
        inside = true;
        if (ob.aux & 0x20u) {                       // SpindleRelevantForInside (torus.h:110)
            bool in_spindle = (sqr(P.y) + sqr(r + ob.p[0]) <= sqr(ob.p[1]));
            inside = (ob.aux & 0x04u) ? in_spindle : !in_spindle;   // SpindleInside (torus.h:107)
        }
    } else inside = false;
    return inside ? !(ob.flags & PVGPU_INVERTED_FLAG) : (ob.flags & PVGPU_INVERTED_FLAG) != 0;
}

// Torus::Normal (torus.cpp:418-445)
__device__ inline V3 torus_normal(const DScene& sc, const pvgpu_object& ob, const V3& ip, uint32_t on_spindle)
{
    const pvgpu_transform& t = sc.xf[ob.transform];
    V3 P = inv_trans_point(t, ip);
    double dist = sqrt(P.x * P.x + P.z * P.z);
    V3 M = mk(0.0, 0.0, 0.0);
    if (dist > PV_EPSILON) { M.x = ob.p[0] * P.x / dist; M.z = ob.p[0] * P.z / dist; }
    if (on_spindle) M = -M;                         // SpindleTorus::Normal (torus.cpp:447-489)
    return normalized(trans_normal(t, P - M));
}

// ---- mesh triangle ------------------------------------------------------------------------------
// Mesh::intersect_mesh_triangle (mesh.cpp:1040-1127): plane hit, then three edge tests in the 2-D
// projection along the dominant axis.
__device__ __forceinline__ bool tri_intersect(const DTri& tr, const V3& o, const V3& d, double& depth)
{
    V3 n = ld3f(tr.n);
    double ndd = dot(n, d);
    if (fabs(ndd) < PV_EPSILON) return false;
    double ndo = dot(n, o);
    depth = -((double)tr.dist + ndo) / ndd;
    if ((depth < PV_MESH_DEPTH_TOL) || (depth > PV_MAX_DISTANCE)) return false;
    const int ua = (tr.dom == 0) ? 1 : 0, va = (tr.dom == 2) ? 1 : 2;
    double s = comp(o, ua) + depth * comp(d, ua);
    double t = comp(o, va) + depth * comp(d, va);
    double p1u = tr.p1[ua], p1v = tr.p1[va], p2u = tr.p2[ua], p2v = tr.p2[va], p3u = tr.p3[ua], p3v = tr.p3[va];
    if ((p2u - s) * (p2v - p1v) < (p2v - t) * (p2u - p1u)) return false;
    if ((p3u - s) * (p3v - p2v) < (p3v - t) * (p3u - p2u)) return false;
    if ((p1u - s) * (p1v - p3v) < (p1v - t) * (p1u - p3u)) return false;
    return true;
}

#if PV_HEAVY
// ---- TrueType glyph -------------------------------------------------------------------------------
#define PV_TTF_TOLERANCE   1.0e-6     // TTF_Tolerance  truetype.cpp:79
#define PV_TTF_COEFF_LIMIT 1.0e-20    // COEFF_LIMIT    truetype.cpp:82

// TrueType::solve_quad (truetype.cpp:2567-2610): roots of c0 s^2 + c1 s + c2 inside [lo, hi], the "+" root first
__device__ inline int glyph_solve_quad(double c0, double c1, double c2, double& r0, double& r1, double lo, double hi)
{
    const double a = c0, b = -c1, c = c2;
    if (fabs(a) < PV_TTF_COEFF_LIMIT) {
        if (fabs(b) < PV_TTF_COEFF_LIMIT) return 0;
        const double q = c / b;
        if (q >= lo && q <= hi) { r0 = q; return 1; }
        return 0;
    }
    double dd = b * b - 4.0 * a * c;
    if (dd < PV_EPSILON) return 0;
    dd = sqrt(dd);
    const double t = 2.0 * a;
    int n = 0;
    double q = (b + dd) / t;
    if (q >= lo && q <= hi) { r0 = q; n = 1; }
    q = (b - dd) / t;
    if (q >= lo && q <= hi) { if (n) r1 = q; else r0 = q; n++; }
    return n;
}

// TrueType::Inside_Glyph (truetype.cpp:2392-2565): crossings of the +x half line from (x, y) with the outline
static __device__ __noinline__ bool glyph_inside_2d(const double* g, double x, double y)
{
    const uint32_t nseg = (uint32_t)g[0];
    int crossings = 0;
    for (uint32_t s = 0; s < nseg; s++) {
        const double* e = g + 1 + 7 * s;
        const double x0 = e[1], y0 = e[2], x1 = e[3], y1 = e[4];
        if (e[0] == 0.0) {
            if (y0 == y1) continue;
            const bool below0 = y0 < y, below1 = y1 < y;
            if (below0 == below1) continue;
            const bool right0 = x0 > x, right1 = x1 > x;
            if (right0 && right1) { crossings++; continue; }
            if (!right0 && !right1) continue;
            const double m = (y1 - y0) / (x1 - x0);
            const double b = (y1 - y) - m * (x1 - x);
            if ((b / m) < PV_EPSILON) crossings++;
        } else {
            const double x2 = e[5], y2 = e[6];
            if (((y0 < y) && (y1 < y) && (y2 < y)) || ((y0 > y) && (y1 > y) && (y2 > y))) continue;
            double r0 = 0.0, r1 = 0.0;
            int k = glyph_solve_quad(y0 - 2.0 * y1 + y2, 2.0 * (y1 - y0), y0 - y, r0, r1, 0.0, 1.0);
            // roots at the ends of the curve only count when y really is in the range of that end (truetype.cpp:2508-2530)
            auto discard = [&](double r) {
                if (r <= PV_EPSILON) return ((y <= y0) && (y < y1)) || ((y >= y0) && (y > y1));
                if (r >= (1.0 - PV_EPSILON)) return ((y < y2) && (y < y1)) || ((y > y2) && (y > y1));
                return false;
            };
            if (k == 2 && discard(r1)) k = 1;
            if (k >= 1 && discard(r0)) { r0 = r1; k--; }
            if (k > 0) {
                const double xt0 = x0 - 2.0 * x1 + x2, xt1 = 2.0 * (x1 - x0), xt2 = x0;
                if ((xt0 * r0 + xt1) * r0 + xt2 > x) crossings++;
                if (k > 1 && (xt0 * r1 + xt1) * r1 + xt2 > x) crossings++;
            }
        }
    }
    return (crossings & 1) != 0;
}

// TrueType::Inside (truetype.cpp:2957-2970)
__device__ inline bool glyph_inside(const DScene& sc, const pvgpu_object& ob, const V3& p)
{
    const V3 q = inv_trans_point(sc.xf[ob.transform], p);
    const bool in = q.z >= 0.0 && q.z <= ob.p[0] && glyph_inside_2d(sc.shape_data + ob.mesh, q.x, q.y);
    return in != ((ob.flags & PVGPU_INVERTED_FLAG) != 0);
}

// TrueType::All_Intersections -> GlyphIntersect (truetype.cpp:2706-2955).  A glyph has as many candidate hits as the ray crosses
// outline segments, so the hits come in batches: the caller starts with *resume = 0 and calls again while it comes back >= 0; the
// order over all batches is the reference's push order (face z = 0, face z = depth, then the walls in outline order).
static __device__ __noinline__ void glyph_hits(const DScene& sc, const pvgpu_object& ob, const V3& o, const V3& d, PrimHits& h, int* resume)
{
    h.n = 0;
    const pvgpu_transform& tr = sc.xf[ob.transform];
    const V3 P = inv_trans_point(tr, o);
    const V3 D = inv_trans_direction(tr, d);              // not normalised: depths are the world ray's own parameter
    const double* g = sc.shape_data + ob.mesh;
    const double gdepth = ob.p[0];
    const uint32_t nseg = (uint32_t)g[0];
    uint32_t s = 0;
    if (resume && *resume > 0) s = (uint32_t)*resume - 1u;
    else if (fabs(D.z) >= PV_EPSILON) {                   // GetZeroOneHits (truetype.cpp:2630-2660)
        double t = -P.z / D.z;
        if (t > 0.0 && t > PV_TTF_TOLERANCE && glyph_inside_2d(g, P.x + t * D.x, P.y + t * D.y)) {
            h.depth[h.n] = t; h.ip[h.n] = evaluate(o, d, t); h.aux[h.n] = 0u; h.n++;
        }
        t += (gdepth / D.z);
        if (t > 0.0 && t > PV_TTF_TOLERANCE && glyph_inside_2d(g, P.x + t * D.x, P.y + t * D.y)) {
            h.depth[h.n] = t; h.ip[h.n] = evaluate(o, d, t); h.aux[h.n] = 1u; h.n++;
        }
    }
    if (resume) *resume = -1;
    int dirflag = 1;
    if (fabs(D.x) < PV_EPSILON) {
        if (fabs(D.y) < PV_EPSILON) return;               // parallel to the walls
        dirflag = 0;
    }
    const double a = D.y, b = -D.x, c = (P.y * D.x - P.x * D.y);
    for (; s < nseg; s++) {
        if (h.n + 2 > PV_MAX_PRIM_HITS) {                 // a segment reports up to two hits: come back for the rest
            if (resume) *resume = (int)s + 1;
            return;
        }
        const double* e = g + 1 + 7 * s;
        const double x0 = e[1], y0 = e[2], x1 = e[3], y1 = e[4];
        if (e[0] == 0.0) {
            const double d0 = (x1 - x0), d1 = (y1 - y0);
            const double t0 = d1 * D.x - d0 * D.y;
            if (fabs(t0) < PV_EPSILON) continue;
            double t = (D.x * (P.y - y0) - D.y * (P.x - x0)) / t0;
            if (t < 0.0 || t > 1.0) continue;
            if (dirflag) t = ((x0 + t * d0) - P.x) / D.x;
            else t = ((y0 + t * d1) - P.y) / D.y;
            const double z = P.z + t * D.z;
            if (z >= 0 && z <= gdepth && t > PV_TTF_TOLERANCE) { h.depth[h.n] = t; h.ip[h.n] = evaluate(o, d, t); h.aux[h.n] = 2u | (s << 3); h.n++; }
        } else {
            const double x2 = e[5], y2 = e[6];
            const double xt2 = x0 - 2.0 * x1 + x2, xt1 = 2.0 * (x1 - x0), xt0 = x0;
            const double yt2 = y0 - 2.0 * y1 + y2, yt1 = 2.0 * (y1 - y0), yt0 = y0;
            double S0 = 0.0, S1 = 0.0;
            const int k = glyph_solve_quad(a * xt2 + b * yt2, a * xt1 + b * yt1, a * xt0 + b * yt0 + c, S0, S1, 0.0, 1.0);
            for (int l = 0; l < k; l++) {
                const double S = l ? S1 : S0;
                double t;
                if (dirflag) t = ((S * S * xt2 + S * xt1 + xt0) - P.x) / D.x;
                else t = ((S * S * yt2 + S * yt1 + yt0) - P.y) / D.y;
                const double z = P.z + t * D.z;
                if (z >= 0 && z <= gdepth && t > PV_TTF_TOLERANCE) { h.depth[h.n] = t; h.ip[h.n] = evaluate(o, d, t); h.aux[h.n] = 2u | ((uint32_t)l << 2) | (s << 3); h.n++; }
            }
        }
    }
}

// The normal GlyphIntersect stores with each hit (truetype.cpp:2734-2737, 2749-2752, 2830-2833, 2905-2908), recomputed from the hit's
// aux and the ray that found it: the wall's curve parameter is solved again with the very same operands.
static __device__ __noinline__ V3 glyph_normal(const DScene& sc, const pvgpu_object& ob, uint32_t aux, const V3& ray_o, const V3& ray_d)
{
    const pvgpu_transform& tr = sc.xf[ob.transform];
    V3 N;
    const uint32_t kind = aux & 3u;
    if (kind == 0u) N = mk(0.0, 0.0, -1.0);
    else if (kind == 1u) N = mk(0.0, 0.0, 1.0);
    else {
        const double* e = sc.shape_data + ob.mesh + 1 + 7 * (aux >> 3);
        const double x0 = e[1], y0 = e[2], x1 = e[3], y1 = e[4];
        if (e[0] == 0.0) N = mk(-(y1 - y0), (x1 - x0), 0.0);
        else {
            const V3 P = inv_trans_point(tr, ray_o);
            const V3 D = inv_trans_direction(tr, ray_d);
            const double x2 = e[5], y2 = e[6];
            const double xt2 = x0 - 2.0 * x1 + x2, xt1 = 2.0 * (x1 - x0), xt0 = x0;
            const double yt2 = y0 - 2.0 * y1 + y2, yt1 = 2.0 * (y1 - y0), yt0 = y0;
            const double a = D.y, b = -D.x, c = (P.y * D.x - P.x * D.y);
            double S0 = 0.0, S1 = 0.0;
            glyph_solve_quad(a * xt2 + b * yt2, a * xt1 + b * yt1, a * xt0 + b * yt0 + c, S0, S1, 0.0, 1.0);
            const double S = (aux & 4u) ? S1 : S0;
            N = mk(-2.0 * yt2 * S - yt1, 2.0 * xt2 * S + xt1, 0.0);
        }
    }
    return normalized(trans_normal(tr, N));
}
#endif  // PV_HEAVY

#if PV_HEAVY
// ---- prism ------------------------------------------------------------------------------------------
#define PV_PRISM_DEPTH_TOL 1.0e-4     // DEPTH_TOLERANCE prism.cpp:150
#define PV_PRISM_ENTRY 15             // doubles per spline segment: x1 y1 x2 y2, v1 u2 v2, A B C D (x, y each)

// Prism::in_curve (prism.cpp:1154-1201): crossing number of the half line u' >= u at height v with the closed spline
static __device__ __noinline__ bool prism_in_curve(const pvgpu_object& ob, const double* sp, double u, double v)
{
    int nc = 0;
    if ((u >= ob.p[6]) && (u <= ob.p[8]) && (v >= ob.p[7]) && (v <= ob.p[9])) {
        const uint32_t number = (uint32_t)sp[0];
        for (uint32_t i = 0; i < number; i++) {
            const double* e = sp + 1 + PV_PRISM_ENTRY * i;
            if ((v >= e[4]) && (v <= e[6]) && (u <= e[5])) {
                double x[4] = { e[8], e[10], e[12], e[14] - v }, y[3];
                int n = solve_polynomial(3, x, y, (ob.flags & PVGPU_STURM_FLAG) ? 1 : 0, 0.0);
                while (n--) {
                    const double w = y[n];
                    if ((w >= 0.0) && (w <= 1.0)) {
                        const double k = w * (w * (w * e[7] + e[9]) + e[11]) + e[13] - u;
                        if (k >= 0.0) nc++;
                    }
                }
            }
        }
    }
    return (nc & 1) != 0;
}

// Prism::test_rectangle (prism.cpp:1233-1345)
__device__ inline bool prism_test_rectangle(const V3& P, const V3& D, double x1, double z1, double x2, double z2)
{
    double dmin, dmax, tmin, tmax;
    if (fabs(D.x) > PV_EPSILON) {
        if (D.x > 0.0) {
            dmin = (x1 - P.x) / D.x;
            dmax = (x2 - P.x) / D.x;
            if (dmax < PV_EPSILON) return false;
        } else {
            dmax = (x1 - P.x) / D.x;
            if (dmax < PV_EPSILON) return false;
            dmin = (x2 - P.x) / D.x;
        }
        if (dmin > dmax) return false;
    } else {
        if ((P.x < x1) || (P.x > x2)) return false;
        dmin = -PV_BOUND_HUGE;
        dmax = PV_BOUND_HUGE;
    }
    if (fabs(D.z) > PV_EPSILON) {
        if (D.z > 0.0) { tmin = (z1 - P.z) / D.z; tmax = (z2 - P.z) / D.z; }
        else { tmax = (z1 - P.z) / D.z; tmin = (z2 - P.z) / D.z; }
        if (tmax < dmax) {
            if (tmax < PV_EPSILON) return false;
            if (tmin > dmin) { if (tmin > tmax) return false; }
            else if (dmin > tmax) return false;
        } else if (tmin > dmin) {
            if (tmin > dmax) return false;
        }
    } else if ((P.z < z1) || (P.z > z2)) return false;
    return true;
}

// The spline parameters at which the ray meets segment `e` (prism.cpp:310-350 linear sweep, 480-510 conic sweep): the roots in y[],
// in the solver's order; k1..k3 are the conic sweep's ray constants.
__device__ inline int prism_segment_roots(const pvgpu_object& ob, const double* e, const V3& P, const V3& D, bool conic, double k1, double k2, double k3, double* y)
{
    const uint32_t spline = ob.aux & 15u;
    const double Ax = e[7], Ay = e[8], Bx = e[9], By = e[10], Cx = e[11], Cy = e[12], Dx = e[13], Dy = e[14];
    double x[4];
    int n = 0;
    if (!conic) {
        switch (spline) {
            case 1:
                x[0] = Cx * D.z - Cy * D.x;
                x[1] = D.z * (Dx - P.x) - D.x * (Dy - P.z);
                if (fabs(x[0]) > PV_EPSILON) y[n++] = -x[1] / x[0];
                break;
            case 2:
                x[0] = Bx * D.z - By * D.x;
                x[1] = Cx * D.z - Cy * D.x;
                x[2] = D.z * (Dx - P.x) - D.x * (Dy - P.z);
                n = solve_polynomial(2, x, y, 0, 0.0);
                break;
            default:
                if (prism_test_rectangle(P, D, e[0], e[1], e[2], e[3])) {
                    x[0] = Ax * D.z - Ay * D.x;
                    x[1] = Bx * D.z - By * D.x;
                    x[2] = Cx * D.z - Cy * D.x;
                    x[3] = D.z * (Dx - P.x) - D.x * (Dy - P.z);
                    n = solve_polynomial(3, x, y, (ob.flags & PVGPU_STURM_FLAG) ? 1 : 0, 0.0);
                }
                break;
        }
    } else {
        switch (spline) {
            case 1:
                x[0] = Cx * k1 + Cy * k2;
                x[1] = Dx * k1 + Dy * k2 + k3;
                if (fabs(x[0]) > PV_EPSILON) y[n++] = -x[1] / x[0];
                break;
            case 2:
                x[0] = Bx * k1 + By * k2;
                x[1] = Cx * k1 + Cy * k2;
                x[2] = Dx * k1 + Dy * k2 + k3;
                n = solve_polynomial(2, x, y, 0, 0.0);
                break;
            default:
                x[0] = Ax * k1 + Ay * k2;
                x[1] = Bx * k1 + By * k2;
                x[2] = Cx * k1 + Cy * k2;
                x[3] = Dx * k1 + Dy * k2 + k3;
                n = solve_polynomial(3, x, y, (ob.flags & PVGPU_STURM_FLAG) ? 1 : 0, 0.0);
                break;
        }
    }
    return n;
}

// Prism::All_Intersections (prism.cpp:194-596).  Hits come in batches like a glyph's (resume protocol); over all batches they follow
// the reference's push order: cap, base, then the segments in order, each segment's roots from the last to the first.
static __device__ __noinline__ void prism_hits(const DScene& sc, const pvgpu_object& ob, const V3& o, const V3& d, PrimHits& h, int* resume)
{
    h.n = 0;
    uint32_t j = 0;
    const bool first = !(resume && *resume > 0);
    if (!first) j = (uint32_t)*resume - 1u;
    if (resume) *resume = -1;
    if (ob.flags & PVGPU_DEGENERATE_FLAG) return;
    const pvgpu_transform& tr = sc.xf[ob.transform];
    const V3 P = inv_trans_point(tr, o);
    V3 D = inv_trans_direction(tr, d);
    const double len = length(D);
    D = D / len;
    if (((D.x >= 0.0) && (P.x > ob.p[4])) || ((D.x <= 0.0) && (P.x < ob.p[2])) || ((D.z >= 0.0) && (P.z > ob.p[5])) || ((D.z <= 0.0) && (P.z < ob.p[3]))) return;
    const double* sp = sc.shape_data + ob.mesh;
    const uint32_t number = (uint32_t)sp[0];
    const double h1 = ob.p[0], h2 = ob.p[1];
    const bool conic = ((ob.aux >> 4) & 15u) == 2u;
    if (fabs(D.y) < PV_EPSILON) {
        if ((P.y < h1) || (P.y > h2)) return;
    } else if (first && (ob.flags & PVGPU_CLOSED_FLAG)) {
        for (int cap = 1; cap >= 0; cap--) {               // cap plane (Height2) first, then the base plane
            const double hh = cap ? h2 : h1;
            if (conic && !(fabs(hh) > PV_EPSILON)) continue;
            const double k = (hh - P.y) / D.y;
            if ((k > PV_PRISM_DEPTH_TOL) && (k < PV_MAX_DISTANCE)) {
                double u = P.x + k * D.x, v = P.z + k * D.z;
                if (conic) { u = u / hh; v = v / hh; }
                if (prism_in_curve(ob, sp, u, v)) {
                    const double dist = k / len;
                    if ((dist > PV_PRISM_DEPTH_TOL) && (dist < PV_MAX_DISTANCE)) { h.depth[h.n] = dist; h.ip[h.n] = evaluate(o, d, dist); h.aux[h.n] = (uint32_t)cap; h.n++; }
                }
            }
        }
    }
    if (!conic && !((fabs(D.x) > PV_EPSILON) || (fabs(D.z) > PV_EPSILON))) return;          // parallel to all sides
    const double k1 = P.z * D.y - P.y * D.z, k2 = P.y * D.x - P.x * D.y, k3 = P.x * D.z - P.z * D.x;
    for (; j < number; j++) {
        if (h.n + 3 > PV_MAX_PRIM_HITS) { if (resume) *resume = (int)j + 1; return; }
        const double* e = sp + 1 + PV_PRISM_ENTRY * j;
        if (((D.x >= 0.0) && (P.x > e[2])) || ((D.x <= 0.0) && (P.x < e[0])) || ((D.z >= 0.0) && (P.z > e[3])) || ((D.z <= 0.0) && (P.z < e[1]))) continue;
        double y[3];
        int n = prism_segment_roots(ob, e, P, D, conic, k1, k2, k3, y);
        while (n--) {
            const double w = y[n];
            if (!((w >= 0.0) && (w <= 1.0))) continue;
            double k;
            if (!conic) {
                if (fabs(D.x) > PV_EPSILON) k = (w * (w * (w * e[7] + e[9]) + e[11]) + e[13] - P.x) / D.x;
                else k = (w * (w * (w * e[8] + e[10]) + e[12]) + e[14] - P.z) / D.z;
            } else {
                k = w * (w * (w * e[7] + e[9]) + e[11]) + e[13];
                double hh = D.x - k * D.y;
                if (fabs(hh) > PV_EPSILON) k = (k * P.y - P.x) / hh;
                else {
                    k = w * (w * (w * e[8] + e[10]) + e[12]) + e[14];
                    hh = D.z - k * D.y;
                    if (fabs(hh) > PV_EPSILON) k = (k * P.y - P.z) / hh;
                    else continue;
                }
            }
            const double hh = P.y + k * D.y;
            if ((hh >= h1) && (hh <= h2)) {
                const double dist = k / len;
                if ((dist > PV_PRISM_DEPTH_TOL) && (dist < PV_MAX_DISTANCE)) { h.depth[h.n] = dist; h.ip[h.n] = evaluate(o, d, dist); h.aux[h.n] = 2u | ((uint32_t)n << 2) | (j << 4); h.n++; }
            }
        }
    }
}

// Prism::Inside (prism.cpp:640-672)
__device__ inline bool prism_inside(const DScene& sc, const pvgpu_object& ob, const V3& p)
{
    V3 P = inv_trans_point(sc.xf[ob.transform], p);
    const bool inv = (ob.flags & PVGPU_INVERTED_FLAG) != 0;
    if ((P.y >= ob.p[0]) && (P.y < ob.p[1])) {
        if (((ob.aux >> 4) & 15u) == 2u) {
            if (fabs(P.y) > PV_EPSILON) { P.x = P.x / P.y; P.z = P.z / P.y; }
            else P.x = P.z = PV_HUGE_VAL;
        }
        if (prism_in_curve(ob, sc.shape_data + ob.mesh, P.x, P.z)) return !inv;
    }
    return inv;
}

// Prism::Normal (prism.cpp:710-770); the spline parameter the reference keeps with the hit (Intersection::d1) is solved for again
// with the ray that found the hit (same operands, same solver path).
static __device__ __noinline__ V3 prism_normal(const DScene& sc, const pvgpu_object& ob, const V3& ip, uint32_t aux, const V3& ray_o, const V3& ray_d)
{
    const pvgpu_transform& tr = sc.xf[ob.transform];
    V3 N = mk(0.0, 0.0, 0.0);
    const uint32_t kind = aux & 3u;
    if (kind == 0u) N = mk(0.0, -1.0, 0.0);
    else if (kind == 1u) N = mk(0.0, 1.0, 0.0);
    else {
        const double* e = sc.shape_data + ob.mesh + 1 + PV_PRISM_ENTRY * (aux >> 4);
        const V3 P = inv_trans_point(tr, ray_o);
        V3 D = inv_trans_direction(tr, ray_d);
        D = D / length(D);
        const bool conic = ((ob.aux >> 4) & 15u) == 2u;
        const double k1 = P.z * D.y - P.y * D.z, k2 = P.y * D.x - P.x * D.y, k3 = P.x * D.z - P.z * D.x;
        double y[3] = { 0.0, 0.0, 0.0 };
        prism_segment_roots(ob, e, P, D, conic, k1, k2, k3, y);
        const double w = y[(aux >> 2) & 3u];
        if (!conic) {
            N.x = w * (3.0 * e[8] * w + 2.0 * e[10]) + e[12];
            N.y = 0.0;
            N.z = -(w * (3.0 * e[7] * w + 2.0 * e[9]) + e[11]);
        } else {
            const V3 Q = inv_trans_point(tr, ip);
            if (fabs(Q.y) > PV_EPSILON) {
                N.x = w * (3.0 * e[8] * w + 2.0 * e[10]) + e[12];
                N.z = -(w * (3.0 * e[7] * w + 2.0 * e[9]) + e[11]);
                N.y = -(Q.x * N.x + Q.z * N.z) / Q.y;
            }
        }
    }
    return normalized(trans_normal(tr, N));
}
#endif  // PV_HEAVY

#if PV_HEAVY
// ---- superellipsoid ---------------------------------------------------------------------------------
#define PV_SE_DEPTH_TOL  1.0e-4      // DEPTH_TOLERANCE superellipsoid.cpp:97
#define PV_SE_ZERO_TOL   1.0e-10     // ZERO_TOLERANCE  superellipsoid.cpp:101
#define PV_SE_MIN       -1.01        // MIN_VALUE / MAX_VALUE superellipsoid.cpp:107-108
#define PV_SE_MAX        1.01
#define PV_SE_MAX_ITER   20          // MAX_ITERATIONS  superellipsoid.cpp:110
#define PV_SE_MAX_HITS   12          // two box planes + nine cutting planes: at most 11 sample points, a hit per point / interval

// Superellipsoid::power (superellipsoid.cpp:1097-1150): small integer exponents by multiplication
__device__ inline double se_power(double x, double e)
{
    const int i = (int)e;
    if (e == (double)i) {
        double b = x;
        switch (i) {
            case 0: return 1.0;
            case 1: return b;
            case 2: return b * b;
            case 3: return (b * b) * b;
            case 4: b *= b; return b * b;
            case 5: b *= b; return (b * b) * x;
            case 6: b *= b; return (b * b) * b;
            default: return pow(x, e);
        }
    }
    return pow(x, e);
}
// evaluate_g / evaluate_superellipsoid (superellipsoid.cpp:1003-1062)
__device__ inline double se_g(double x, double y, double e)
{
    double g = 0;
    if (x > y) {
        g = 1 + se_power(y / x, e);
        if (g != 1) g = se_power(g, 1 / e);
        g *= x;
    } else if (y != 0) {
        g = 1 + se_power(x / y, e);
        if (g != 1) g = se_power(g, 1 / e);
        g *= y;
    }
    return g;
}
static __device__ __noinline__ double se_value(const pvgpu_object& ob, const V3& P)
{
    return se_g(se_g(fabs(P.x), fabs(P.y), ob.p[0]), fabs(P.z), ob.p[2]) - 1;
}

// Superellipsoid::intersect_box (superellipsoid.cpp:832-1001): the ray against the cube [-1.01, 1.01]^3
__device__ inline bool se_intersect_box(const V3& P, const V3& D, double& dmin, double& dmax)
{
    double tmin = 0.0, tmax = 0.0;
    if (fabs(D.x) > PV_EPSILON) {
        if (D.x > PV_EPSILON) { dmin = (PV_SE_MIN - P.x) / D.x; dmax = (PV_SE_MAX - P.x) / D.x; if (dmax < PV_EPSILON) return false; }
        else { dmax = (PV_SE_MIN - P.x) / D.x; if (dmax < PV_EPSILON) return false; dmin = (PV_SE_MAX - P.x) / D.x; }
        if (dmin > dmax) return false;
    } else {
        if ((P.x < PV_SE_MIN) || (P.x > PV_SE_MAX)) return false;
        dmin = -PV_BOUND_HUGE; dmax = PV_BOUND_HUGE;
    }
    #pragma unroll
    for (int axis = 1; axis < 3; axis++) {
        const double Pa = axis == 1 ? P.y : P.z, Da = axis == 1 ? D.y : D.z;
        if (fabs(Da) > PV_EPSILON) {
            if (Da > PV_EPSILON) { tmin = (PV_SE_MIN - Pa) / Da; tmax = (PV_SE_MAX - Pa) / Da; }
            else { tmax = (PV_SE_MIN - Pa) / Da; tmin = (PV_SE_MAX - Pa) / Da; }
            if (tmax < dmax) {
                if (tmax < PV_EPSILON) return false;
                if (tmin > dmin) { if (tmin > tmax) return false; dmin = tmin; }
                else if (dmin > tmax) return false;
                dmax = tmax;
            } else if (tmin > dmin) {
                if (tmin > dmax) return false;
                dmin = tmin;
            }
        } else if ((Pa < PV_SE_MIN) || (Pa > PV_SE_MAX)) return false;
    }
    return true;
}

// Superellipsoid::solve_hit1 (superellipsoid.cpp:1344-1450): secant and bisection steps on a bracketed root
__device__ inline V3 se_solve_hit1(const pvgpu_object& ob, double v0, V3 P0, double v1, V3 P1)
{
    for (int i = 0; i < PV_SE_MAX_ITER; i++) {
        if (fabs(v0) < PV_SE_ZERO_TOL) return P0;
        if (fabs(v1) < PV_SE_ZERO_TOL) return P1;
        const double x = fabs(v0) / fabs(v1 - v0);
        V3 P2 = P1 - P0;
        P2 = P0 + x * P2;
        const double v2 = se_value(ob, P2);
        V3 P3 = P1 - P0;
        P3 = P0 + 0.5 * P3;
        const double v3 = se_value(ob, P3);
        if (v2 * v3 < 0.0) { v0 = v2; P0 = P2; v1 = v3; P1 = P3; }
        else if (fabs(v2) < fabs(v3)) {
            if (v0 * v2 < 0) { v1 = v2; P1 = P2; } else { v0 = v2; P0 = P2; }
        } else {
            if (v0 * v3 < 0) { v1 = v3; P1 = P3; } else { v0 = v3; P0 = P3; }
        }
    }
    return (fabs(v0) < fabs(v1)) ? P0 : P1;
}

// Superellipsoid::check_hit2 (superellipsoid.cpp:1485-1560): no sign change between two sample points - walk towards the surface
__device__ inline bool se_check_hit2(const pvgpu_object& ob, const V3& P, const V3& D, double t0, const V3& P0, double v0, double t1, double& t)
{
    double dt0 = t0, dt1 = t0 + 0.0001 * (t1 - t0);
    const double maxdelta = t1 - t0;
    for (int i = 0; (dt0 < t1) && (i < PV_SE_MAX_ITER); i++) {
        const V3 P1 = P + dt1 * D;
        const double v1 = se_value(ob, P1);
        double deltat;
        if (v0 * v1 < 0) {
            const V3 Q = se_solve_hit1(ob, v0, P0, v1, P1);
            t = length(Q - P);
            return true;
        }
        if (fabs(v1) < PV_SE_ZERO_TOL) { t = dt1; return true; }
        if (((v0 > 0.0) && (v1 > v0)) || ((v0 < 0.0) && (v1 < v0))) break;
        if (v1 == v0) break;
        deltat = v1 * (dt1 - dt0) / (v1 - v0);
        if (fabs(deltat) > maxdelta) break;
        v0 = v1;
        dt0 = dt1;
        dt1 -= deltat;
    }
    return false;
}

// Superellipsoid::Intersect (superellipsoid.cpp:213-394): depths of the hits in the order the reference pushes them.  A superellipsoid
// that is not a CSG child stops at its first hit (`first_only`).  insert_hit's clip test steers that walk in the reference, so
// superellipsoids with a clipped_by list are rejected at finalize instead of being served with a different hit order.
static __device__ __noinline__ int se_depths(const DScene& sc, const pvgpu_object& ob, const V3& o, const V3& d, bool first_only, double* out)
{
    const pvgpu_transform& tr = sc.xf[ob.transform];
    const V3 P = inv_trans_point(tr, o);
    V3 D = inv_trans_direction(tr, d);
    const double len = length(D);
    D = D / len;
    double t1, t2;
    if (!se_intersect_box(P, D, t1, t2)) return 0;
    if (t2 < PV_SE_DEPTH_TOL) return 0;
    double dists[11];
    int cnt = 0;
    if (t1 < PV_SE_DEPTH_TOL) t1 = PV_SE_DEPTH_TOL;
    dists[cnt++] = t1;
    dists[cnt++] = t2;
    {   // find_ray_plane_points (superellipsoid.cpp:1271-1310)
        const double planes[9][3] = { { 1, 1, 0 }, { 1, -1, 0 }, { 1, 0, 1 }, { 1, 0, -1 }, { 0, 1, 1 }, { 0, 1, -1 }, { 1, 0, 0 }, { 0, 1, 0 }, { 0, 0, 1 } };
        const double slack = PV_EPSILON * (t2 - t1);
        const double mind = t1 - slack, maxd = t2 + slack;
        for (int i = 0; i < 9; i++) {
            const double dd = (D.x * planes[i][0] + D.y * planes[i][1] + D.z * planes[i][2]);
            if (fabs(dd) < PV_EPSILON) continue;
            const double t = (0.0 - (P.x * planes[i][0] + P.y * planes[i][1] + P.z * planes[i][2])) / dd;
            if ((t >= mind) && (t <= maxd)) dists[cnt++] = t;
        }
        for (int i = 1; i < cnt; i++) {                    // (qsort by value)
            const double v = dists[i];
            int j = i - 1;
            while (j >= 0 && dists[j] > v) { dists[j + 1] = dists[j]; j--; }
            dists[j + 1] = v;
        }
    }
    if (cnt <= 1) return 0;
    int n = 0;
    // insert_hit (superellipsoid.cpp:1172-1190)
    auto insert = [&](double depth) -> bool {
        if (!((depth > PV_SE_DEPTH_TOL) && (depth < PV_MAX_DISTANCE))) return false;
        if (n < PV_SE_MAX_HITS) out[n++] = depth;
        return true;
    };
    V3 P0 = P + dists[0] * D;
    double v0 = se_value(ob, P0);
    if (fabs(v0) < PV_SE_ZERO_TOL) { if (insert(dists[0] / len) && first_only) return n; }
    for (int i = 1; i < cnt; i++) {
        const V3 P1 = P + dists[i] * D;
        const double v1 = se_value(ob, P1);
        if (fabs(v1) < PV_SE_ZERO_TOL) { if (insert(dists[i] / len) && first_only) return n; }
        else if (v0 * v1 < 0.0) {
            const V3 P2 = se_solve_hit1(ob, v0, P0, v1, P1);
            const double t = length(P2 - P);
            if (insert(t / len) && first_only) return n;
        } else {
            double t;
            if (se_check_hit2(ob, P, D, dists[i - 1], P0, v0, dists[i], t)) {
                if (insert(t / len)) { if (first_only) return n; }
                else break;
            }
        }
        v0 = v1;
        P0 = P1;
    }
    return n;
}

// Superellipsoid::All_Intersections: hits in batches (resume protocol of glyph_hits); aux bit 0 of the object = IS_CHILD_OBJECT
static __device__ __noinline__ void superellipsoid_hits(const DScene& sc, const pvgpu_object& ob, const V3& o, const V3& d, PrimHits& h, int* resume)
{
    h.n = 0;
    double depths[PV_SE_MAX_HITS];
    const int n = se_depths(sc, ob, o, d, (ob.aux & 1u) == 0u, depths);
    const int start = (resume && *resume > 0) ? *resume - 1 : 0;
    if (resume) *resume = -1;
    for (int i = start; i < n; i++) {
        if (h.n == PV_MAX_PRIM_HITS) { if (resume) *resume = i + 1; return; }
        h.depth[h.n] = depths[i]; h.ip[h.n] = evaluate(o, d, depths[i]); h.aux[h.n] = 0u; h.n++;
    }
}

// Superellipsoid::Inside (superellipsoid.cpp:396-420)
__device__ inline bool superellipsoid_inside(const DScene& sc, const pvgpu_object& ob, const V3& p)
{
    const double val = se_value(ob, inv_trans_point(sc.xf[ob.transform], p));
    return (val < PV_EPSILON) != ((ob.flags & PVGPU_INVERTED_FLAG) != 0);
}

// Superellipsoid::Normal (superellipsoid.cpp:451-495)
static __device__ __noinline__ V3 superellipsoid_normal(const DScene& sc, const pvgpu_object& ob, const V3& ip)
{
    const pvgpu_transform& tr = sc.xf[ob.transform];
    V3 P = inv_trans_point(tr, ip);
    const double Ex = ob.p[0], Ez = ob.p[2];
    double r = 0.0, z2n = 0;          // (r is uninitialised in the reference when P.x == P.y == 0; its product with P.z is then the only use)
    if (P.z != 0) { z2n = se_power(fabs(P.z), Ez); P.z = z2n / P.z; }
    if (fabs(P.x) > fabs(P.y)) {
        r = se_power(fabs(P.y / P.x), Ex);
        P.x = (1 - z2n) / P.x;
        P.y = (P.y != 0.0) ? (1 - z2n) * r / P.y : 0;
    } else if (P.y != 0) {
        r = se_power(fabs(P.x / P.y), Ex);
        P.x = (P.x != 0.0) ? (1 - z2n) * r / P.x : 0;
        P.y = (1 - z2n) / P.y;
    }
    if (P.z != 0.0) P.z *= (1 + r);
    return normalized(trans_normal(tr, P));
}
#endif  // PV_HEAVY

}  // namespace pvgpu
