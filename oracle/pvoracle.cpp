// pvoracle -- CPU restatement of POV-Ray's per-pixel trace path.  TEST INFRASTRUCTURE ONLY.
//
// This file is the checker, not the product: only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load it.  Nothing under povray_b200/ links or calls it.
//
// It restates, in plain recursive C++ that mirrors the reference's own control flow (priority-queue tree
// walk, recursive TraceRay returning colours), the functions SURVEY.md section 8a lists; every function
// cites the reference file:line it follows.  Its structure is deliberately different from the CUDA
// wavefront implementation (heap instead of stack, recursion instead of ray queues, colours returned
// instead of weighted accumulation), so agreement between the two is evidence, not tautology.
//
// Pinning: the reference holds no golden vectors for this path, so the oracle is
// pinned against outputs of the UNMODIFIED reference built into oracle/_ref (ray-level object id + depth,
// float RGBT per pixel) -- see tests/test_oracle_vs_reference.py and tests/golden/.
//
// Input: the flat tables of include/pvgpu.h (a .pvs file written by pvgpu_scene_save).
// Build: make -C oracle oracle   (g++ -O2 -fno-fast-math -ffp-contract=off)
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <random>
#include <vector>

#include "pvgpu.h"

namespace {

constexpr double EPSILON = 1.0e-10, HUGE_VALUE = 1.0e17, BOUND_HUGE = 2.0e10, SMALL_TOLERANCE = 1.0e-3,
                 MAX_DISTANCE = 1.0e7, MIN_ISECT_DEPTH = 1.0e-4, SHADOW_TOLERANCE = 1.0e-3, COORDINATE_LIMIT = 1.0e17;

struct V3 {
    double x, y, z;
    double operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
inline V3 v3(double x, double y, double z) { return V3{ x, y, z }; }
inline V3 v3(const double* p) { return V3{ p[0], p[1], p[2] }; }
inline V3 v3f(const float* p) { return V3{ (double)p[0], (double)p[1], (double)p[2] }; }
inline V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 operator-(V3 a) { return v3(-a.x, -a.y, -a.z); }
inline V3 operator*(V3 a, double s) { return v3(a.x * s, a.y * s, a.z * s); }
inline V3 operator*(double s, V3 a) { return v3(a.x * s, a.y * s, a.z * s); }
inline V3 operator/(V3 a, double s) { return v3(a.x / s, a.y / s, a.z / s); }
inline double dot(V3 a, V3 b) { return ((a.x * b.x) + (a.y * b.y)) + (a.z * b.z); }        // vector.h:568
inline double len2(V3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
inline double len(V3 a) { return std::sqrt(len2(a)); }
inline V3 cross(V3 a, V3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
inline V3 unit(V3 a) { double l = len(a); return l != 0.0 ? a / l : a; }                     // vector.h:531
inline double sqr(double x) { return x * x; }

struct Col { float r, g, b; };
inline Col operator+(Col a, Col b) { return Col{ a.r + b.r, a.g + b.g, a.b + b.b }; }
inline Col operator*(Col a, Col b) { return Col{ a.r * b.r, a.g * b.g, a.b * b.b }; }
inline Col operator*(Col a, float s) { return Col{ a.r * s, a.g * s, a.b * s }; }
inline float grey(Col c) { return (float)(0.297 * c.r + 0.589 * c.g + 0.114 * c.b); }        // colour.h:1366
inline bool near_zero(Col c, float e) { return std::fabs(c.r) < e && std::fabs(c.g) < e && std::fabs(c.b) < e; }

struct Scene {
    pvgpu_globals g{};
    pvgpu_camera cam{};
    std::vector<pvgpu_object> objects;
    std::vector<uint32_t> index_list, frame;
    std::vector<pvgpu_transform> xf;
    std::vector<pvgpu_node> nodes;
    std::vector<pvgpu_mesh> meshes;
    std::vector<float> verts, norms;
    std::vector<pvgpu_triangle> tris;
    std::vector<pvgpu_node> mnodes;
    std::vector<pvgpu_light> lights;
    std::vector<pvgpu_texture> textures;
    std::vector<pvgpu_pigment> pigments;
    std::vector<pvgpu_finish> finishes;
    std::vector<pvgpu_blend_map> maps;
    std::vector<pvgpu_blend_entry> entries;
    std::vector<pvgpu_warp> warps;
    std::vector<pvgpu_interior> interiors;
    std::vector<pvgpu_blob> blobs;
    std::vector<pvgpu_blob_element> blob_elements;
    std::vector<int32_t> blob_textures;          // per blob element: texture or -1 (Blob::Element_Texture)
    std::vector<double> mesh_uv;                 // MESH_DATA::UVCoords of all meshes, (u, v) pairs
    std::vector<uint32_t> tri_uv;                // per triangle: UV1 UV2 UV3 as indices into mesh_uv
    std::vector<pvgpu_image> images;             // image_map pigments
    std::vector<float> texels;                   // r g b filter transmit per texel, row 0 = top row
    std::vector<pvgpu_blob_node> blob_nodes;
    std::vector<double> shape_data;
    std::vector<pvgpu_tnormal> tnormals;
    std::vector<pvgpu_slope_entry> slope_entries;
    std::vector<pvgpu_sky_sphere> sky_spheres;
    std::vector<pvgpu_fog> fogs;
    std::vector<double> camera_ext;
    std::vector<float> irid_wavelengths;
    std::vector<V3> waveSources;                 // TraceThreadData::waveSources / waveFrequencies (tracethreaddata.cpp:110-111)
    std::vector<double> waveFrequencies;
    // noise tables
    std::vector<double> patternRands;            // gPatternRands (pattern.cpp:91): mt19937 / 2^32, 32768 entries
    std::vector<unsigned short> hashTable;
    std::vector<double> RTable;
    std::vector<int> NoisePermutation;
    std::vector<V3> NoiseGradients;
    bool use_tree = false;
};

// ---- transforms (matrix.cpp:415-507, matrix.h:97-105) ------------------------------------------------
inline V3 mpoint(const double* m, V3 v)
{
    return v3(v.x * m[0] + v.y * m[4] + v.z * m[8] + m[12], v.x * m[1] + v.y * m[5] + v.z * m[9] + m[13],
              v.x * m[2] + v.y * m[6] + v.z * m[10] + m[14]);
}
inline V3 mdir(const double* m, V3 v)
{
    return v3(v.x * m[0] + v.y * m[4] + v.z * m[8], v.x * m[1] + v.y * m[5] + v.z * m[9], v.x * m[2] + v.y * m[6] + v.z * m[10]);
}
inline V3 mtransposed(const double* m, V3 v)
{
    return v3(v.x * m[0] + v.y * m[1] + v.z * m[2], v.x * m[4] + v.y * m[5] + v.z * m[6], v.x * m[8] + v.y * m[9] + v.z * m[10]);
}
inline V3 MTransPoint(const pvgpu_transform& t, V3 v) { return mpoint(t.matrix, v); }
inline V3 MInvTransPoint(const pvgpu_transform& t, V3 v) { return mpoint(t.inverse, v); }
inline V3 MInvTransDirection(const pvgpu_transform& t, V3 v) { return mdir(t.inverse, v); }
inline V3 MTransNormal(const pvgpu_transform& t, V3 v) { return mtransposed(t.inverse, v); }

// ---- polynomial solver (polynomialsolver.cpp) --------------------------------------------------------
constexpr double SMALL_ENOUGH = 1.0e-10, RELERROR = 1.0e-12, FUDGE_FACTOR1 = 1.0e12;
constexpr int MAX_ITERATIONS = 50, MAX_ORDER = 4;
constexpr double TWO_M_PI_3 = 2.0943951023931954923084, FOUR_M_PI_3 = 4.1887902047863909846168;

int solve_quadratic(const double* x, double* y)          // :810-861
{
    double a = x[0], b = -x[1], c = x[2];
    if (a == 0.0) { if (b == 0.0) return 0; y[0] = c / b; return 1; }
    b /= a; c /= a; a = 1.0;
    double d = b * b - 4.0 * a * c;
    if ((d > -SMALL_ENOUGH) && (d < SMALL_ENOUGH)) { y[0] = 0.5 * b / a; return 1; }
    if (d < 0.0) return 0;
    d = std::sqrt(d);
    double t = 2.0 * a;
    y[0] = (b + d) / t; y[1] = (b - d) / t;
    return 2;
}
int solve_cubic(const double* x, double* y)              // :903-978
{
    double a0 = x[0], a1, a2, a3;
    if (a0 == 0.0) return solve_quadratic(&x[1], y);
    if (a0 != 1.0) { a1 = x[1] / a0; a2 = x[2] / a0; a3 = x[3] / a0; } else { a1 = x[1]; a2 = x[2]; a3 = x[3]; }
    double A2 = a1 * a1, Q = (A2 - 3.0 * a2) / 9.0, R = (a1 * (A2 - 4.5 * a2) + 13.5 * a3) / 27.0;
    double Q3 = Q * Q * Q, R2 = R * R, d = Q3 - R2, an = a1 / 3.0;
    if (d >= 0.0) {
        d = R / std::sqrt(Q3);
        double theta = std::acos(d) / 3.0, sQ = -2.0 * std::sqrt(Q);
        y[0] = sQ * std::cos(theta) - an; y[1] = sQ * std::cos(theta + TWO_M_PI_3) - an; y[2] = sQ * std::cos(theta + FOUR_M_PI_3) - an;
        return 3;
    }
    double sQ = std::pow(std::sqrt(R2 - Q3) + std::fabs(R), 1.0 / 3.0);
    y[0] = (R < 0) ? (sQ + Q / sQ) - an : -(sQ + Q / sQ) - an;
    return 1;
}
int solve_quartic(const double* x, double* results)      // :1325-1450
{
    double cubic[4], roots[3], c0 = x[0], c1, c2, c3, c4;
    if (c0 != 1.0) { c1 = x[1] / c0; c2 = x[2] / c0; c3 = x[3] / c0; c4 = x[4] / c0; } else { c1 = x[1]; c2 = x[2]; c3 = x[3]; c4 = x[4]; }
    double c12 = c1 * c1, p = -0.375 * c12 + c2, q = 0.125 * c12 * c1 - 0.5 * c1 * c2 + c3;
    double r = -0.01171875 * c12 * c12 + 0.0625 * c12 * c2 - 0.25 * c1 * c3 + c4;
    cubic[0] = 1.0; cubic[1] = -0.5 * p; cubic[2] = -r; cubic[3] = 0.5 * r * p - 0.125 * q * q;
    int i = solve_cubic(cubic, roots);
    if (i <= 0) return 0;
    double z = roots[0], d1 = 2.0 * z - p, d2;
    if (d1 < 0.0) { if (d1 > -SMALL_ENOUGH) d1 = 0.0; else return 0; }
    if (d1 < SMALL_ENOUGH) { d2 = z * z - r; if (d2 < 0.0) return 0; d2 = std::sqrt(d2); }
    else { d1 = std::sqrt(d1); d2 = 0.5 * q / d1; }
    double q1 = d1 * d1, q2 = -0.25 * c1;
    i = 0;
    p = q1 - 4.0 * (z - d2);
    if (p == 0) results[i++] = -0.5 * d1 - q2;
    else if (p > 0) { p = std::sqrt(p); results[i++] = -0.5 * (d1 + p) + q2; results[i++] = -0.5 * (d1 - p) + q2; }
    p = q1 - 4.0 * (z + d2);
    if (p == 0) results[i++] = 0.5 * d1 - q2;
    else if (p > 0) { p = std::sqrt(p); results[i++] = 0.5 * (d1 + p) + q2; results[i++] = 0.5 * (d1 - p) + q2; }
    return i;
}
int difficult_coeffs(int n, const double* x)             // :1075-1109 (variant without USE_NEW_DIFFICULT_COEFFS)
{
    double biggest = 0.0;
    for (int i = 0; i <= n; i++) if (std::fabs(x[i]) > biggest) biggest = x[i];
    if (biggest == 0.0) return 0;
    for (int i = 0; i <= n; i++) if (x[i] != 0.0 && std::fabs(biggest / x[i]) > FUDGE_FACTOR1) return 1;
    return 0;
}
struct polynomial { int ord; double coef[MAX_ORDER + 1]; };
double polyeval(double x, int n, const double* c) { double v = c[n]; for (int i = n - 1; i >= 0; i--) v = v * x + c[i]; return v; }
int modp(const polynomial* u, const polynomial* v, polynomial* r)          // :171-215
{
    *r = *u;
    if (v->coef[v->ord] < 0.0) {
        for (int k = u->ord - v->ord - 1; k >= 0; k -= 2) r->coef[k] = -r->coef[k];
        for (int k = u->ord - v->ord; k >= 0; k--)
            for (int j = v->ord + k - 1; j >= k; j--) r->coef[j] = -r->coef[j] - r->coef[v->ord + k] * v->coef[j - k];
    } else {
        for (int k = u->ord - v->ord; k >= 0; k--)
            for (int j = v->ord + k - 1; j >= k; j--) r->coef[j] -= r->coef[v->ord + k] * v->coef[j - k];
    }
    int k = v->ord - 1;
    while (k >= 0 && std::fabs(r->coef[k]) < SMALL_ENOUGH) { r->coef[k] = 0.0; k--; }
    r->ord = (k < 0) ? 0 : k;
    return r->ord;
}
int buildsturm(int ord, polynomial* sseq)                // :233-270
{
    sseq[0].ord = ord; sseq[1].ord = ord - 1;
    double f = std::fabs(sseq[0].coef[ord] * ord);
    double* fp = sseq[1].coef; double* fc = sseq[0].coef + 1;
    for (int i = 1; i <= ord; i++) *fp++ = *fc++ * i / f;
    polynomial* sp;
    for (sp = sseq + 2; modp(sp - 2, sp - 1, sp); sp++) {
        f = -std::fabs(sp->coef[sp->ord]);
        for (fp = &sp->coef[sp->ord]; fp >= sp->coef; fp--) *fp /= f;
    }
    sp->coef[0] = -sp->coef[0];
    return (int)(sp - sseq);
}
int numchanges(int np, const polynomial* sseq, double a) // :330-352
{
    int changes = 0;
    double lf = polyeval(a, sseq[0].ord, sseq[0].coef);
    for (const polynomial* s = sseq + 1; s <= sseq + np; s++) {
        double f = polyeval(a, s->ord, s->coef);
        if (lf == 0.0 || lf * f < 0) changes++;
        lf = f;
    }
    return changes;
}
int visible_roots(int np, const polynomial* sseq)        // :288-328
{
    int atposinf = 0, atzero = 0;
    double lf = sseq[0].coef[sseq[0].ord];
    for (const polynomial* s = sseq + 1; s <= sseq + np; s++) { double f = s->coef[s->ord]; if (lf == 0.0 || lf * f < 0) atposinf++; lf = f; }
    lf = sseq[0].coef[0];
    for (const polynomial* s = sseq + 1; s <= sseq + np; s++) { double f = s->coef[0]; if (lf == 0.0 || lf * f < 0) atzero++; lf = f; }
    return atzero - atposinf;
}
int regula_falsa(int order, const double* coef, double a, double b, double* val)     // :700-775
{
    double fa = polyeval(a, order, coef), fb = polyeval(b, order, coef);
    if (fa * fb > 0.0) return 0;
    if (std::fabs(fa) < SMALL_ENOUGH) { *val = a; return 1; }
    if (std::fabs(fb) < SMALL_ENOUGH) { *val = b; return 1; }
    double lfx = fa;
    for (int its = 0; its < MAX_ITERATIONS; its++) {
        double x = (fb * a - fa * b) / (fb - fa), fx = polyeval(x, order, coef);
        if (std::fabs(x) > RELERROR) { if (std::fabs(fx / x) < RELERROR) { *val = x; return 1; } }
        else if (std::fabs(fx) < RELERROR) { *val = x; return 1; }
        if (fa < 0) {
            if (fx < 0) { a = x; fa = fx; if ((lfx * fx) > 0) fb /= 2; } else { b = x; fb = fx; if ((lfx * fx) > 0) fa /= 2; }
        } else {
            if (fx < 0) { b = x; fb = fx; if ((lfx * fx) > 0) fa /= 2; } else { a = x; fa = fx; if ((lfx * fx) > 0) fb /= 2; }
        }
        if (std::fabs(b - a) < RELERROR) { *val = x; return 1; }
        lfx = fx;
    }
    return 0;
}
int sbisect(int np, const polynomial* sseq, double min_value, double max_value, int atmin, int atmax, double* roots)   // :370-480
{
    double mid = 0;
    int n1, n2, its, atmid;
    if ((atmin - atmax) == 1) {
        if (regula_falsa(sseq->ord, sseq->coef, min_value, max_value, roots)) return 1;
        for (its = 0; its < MAX_ITERATIONS; its++) {
            mid = (min_value + max_value) / 2;
            atmid = numchanges(np, sseq, mid);
            if ((atmid < atmax) || (atmid > atmin)) return 0;
            if (std::fabs(mid) > RELERROR) { if (std::fabs((max_value - min_value) / mid) < RELERROR) { roots[0] = mid; return 1; } }
            else if (std::fabs(max_value - min_value) < RELERROR) { roots[0] = mid; return 1; }
            if ((atmin - atmid) == 0) min_value = mid; else max_value = mid;
        }
        roots[0] = mid;
        return 1;
    }
    for (its = 0; its < MAX_ITERATIONS; its++) {
        mid = (min_value + max_value) / 2;
        atmid = numchanges(np, sseq, mid);
        if ((atmid < atmax) || (atmid > atmin)) return 0;
        if (std::fabs(mid) > RELERROR) { if (std::fabs((max_value - min_value) / mid) < RELERROR) { roots[0] = mid; return 1; } }
        else if (std::fabs(max_value - min_value) < RELERROR) { roots[0] = mid; return 1; }
        n1 = atmin - atmid; n2 = atmid - atmax;
        if ((n1 != 0) && (n2 != 0)) {
            n1 = sbisect(np, sseq, min_value, mid, atmin, atmid, roots);
            n2 = sbisect(np, sseq, mid, max_value, atmid, atmax, &roots[n1]);
            return n1 + n2;
        }
        if (n1 == 0) min_value = mid; else max_value = mid;
    }
    roots[0] = mid;
    return 1;
}
int polysolve(int order, const double* Coeffs, double* roots)             // :1481-1523
{
    polynomial sseq[MAX_ORDER + 1];
    for (int i = 0; i <= order; i++) sseq[0].coef[order - i] = Coeffs[i] / Coeffs[0];
    int np = buildsturm(order, &sseq[0]);
    if (visible_roots(np, sseq) == 0) return 0;
    int atmin = numchanges(np, sseq, 0.0), atmax = numchanges(np, sseq, MAX_DISTANCE);
    if (atmin - atmax == 0) return 0;
    return sbisect(np, sseq, 0.0, MAX_DISTANCE, atmin, atmax, roots);
}
int Solve_Polynomial(int n, const double* c0, double* r, int sturm, double epsilon)   // :1585-1729 (n <= 4)
{
    int roots = 0, i = 0;
    while ((i < n) && (std::fabs(c0[i]) < SMALL_ENOUGH)) i++;
    n -= i;
    const double* c = &c0[i];
    switch (n) {
        case 0: break;
        case 1: if (c[0] != 0.0) r[roots++] = -c[1] / c[0]; break;
        case 2: roots = solve_quadratic(c, r); break;
        case 3:
            if (epsilon > 0.0 && (c[2] != 0.0) && (std::fabs(c[3] / c[2]) < epsilon)) { roots = solve_quadratic(c, r); break; }
            roots = sturm ? polysolve(3, c, r) : solve_cubic(c, r);
            break;
        case 4:
            if (epsilon > 0.0 && (c[3] != 0.0) && (std::fabs(c[4] / c[3]) < epsilon)) { roots = sturm ? polysolve(3, c, r) : solve_cubic(c, r); break; }
            if (difficult_coeffs(4, c)) sturm = 1;
            roots = sturm ? polysolve(4, c, r) : solve_quartic(c, r);
            break;
    }
    return roots;
}

// ---- noise (noise.cpp, portablenoise.cpp) ------------------------------------------------------------
const double kRTableEven[267] = {
#include "pv_rtable.inc"
};
void init_noise(Scene& s)
{
    s.hashTable.assign(8192, 0);                          // InitTextureTable noise.cpp:231-255
    for (int i = 0; i < 4096; i++) s.hashTable[i] = (unsigned short)i;
    int next_rand = 0;
    for (int i = 4095; i >= 0; i--) {
        next_rand = (int)((long long)next_rand * 1812433253LL + 12345LL);
        unsigned short j = (unsigned short)(((int)(next_rand >> 16) & 0x7FFF) % 4096);
        std::swap(s.hashTable[i], s.hashTable[j]);
    }
    for (int i = 0; i < 4096; i++) s.hashTable[4096 + i] = s.hashTable[i];
    s.RTable.assign(534, 0.0);                            // Initialize_Noise noise.cpp:181-182
    for (int i = 0; i < 267; i++) { s.RTable[2 * i] = kRTableEven[i]; s.RTable[2 * i + 1] = kRTableEven[i] * 0.5; }
    const int NE = 2048;                                  // InitSolidNoise noise.cpp:306-348
    s.NoisePermutation.assign(2 * (NE + 1), 0);
    s.NoiseGradients.assign(2 * (NE + 1), v3(0, 0, 0));
    next_rand = 1;
    for (int i = 0; i < NE; i++) {
        double v[3], q;
        do {
            for (int j = 0; j < 3; j++) {
                next_rand = (int)((long long)next_rand * 1812433253LL + 12345LL);
                v[j] = (double)((((int)(next_rand >> 16) & 0x7FFF) % (NE << 1)) - NE) / (double)NE;
            }
            q = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
        } while ((q > 1.0) || (q < 1.0e-5));
        double l = std::sqrt(q);
        s.NoiseGradients[i] = v3(v[0] / l, v[1] / l, v[2] / l);
    }
    for (int i = 0; i < NE; i++) s.NoisePermutation[i] = i;
    for (int i = NE; i > 0; i -= 2) {
        int k = s.NoisePermutation[i];
        next_rand = (int)((long long)next_rand * 1812433253LL + 12345LL);
        int j = ((int)(next_rand >> 16) & 0x7FFF) % NE;
        s.NoisePermutation[i] = s.NoisePermutation[j];
        s.NoisePermutation[j] = k;
    }
    for (int i = 0; i < NE + 2; i++) { s.NoisePermutation[NE + i] = s.NoisePermutation[i]; s.NoiseGradients[NE + i] = s.NoiseGradients[i]; }
}
inline double SCURVE(double a) { return a * a * (3.0 - 2.0 * a); }
inline double Lerp(double t, double a, double b) { return a + t * (b - a); }
double SolidNoise(const Scene& s, V3 P)                   // noise.cpp:392-440
{
    const int NE = 2048;
    const double ROLLOVER = 10000000.023157213;
    int b0[3], b1[3]; double r0[3], r1[3];
    for (int i = 0; i < 3; i++) {
        double t = P[i] + ROLLOVER; int it = (int)std::floor(t);
        b0[i] = it & (NE - 1); b1[i] = (b0[i] + 1) & (NE - 1); r0[i] = t - it; r1[i] = r0[i] - 1.0;
    }
    const std::vector<int>& NP = s.NoisePermutation;
    int i = NP[b0[0]], j = NP[b1[0]];
    int b00 = NP[i + b0[1]], b10 = NP[j + b0[1]], b01 = NP[i + b1[1]], b11 = NP[j + b1[1]];
    double sx = SCURVE(r0[0]), sy = SCURVE(r0[1]), sz = SCURVE(r0[2]);
    auto at = [&](int idx, double rx, double ry, double rz) { const V3& q = s.NoiseGradients[idx]; return rx * q.x + ry * q.y + rz * q.z; };
    double u, v, a, b, c, d;
    u = at(b00 + b0[2], r0[0], r0[1], r0[2]); v = at(b10 + b0[2], r1[0], r0[1], r0[2]); a = Lerp(sx, u, v);
    u = at(b01 + b0[2], r0[0], r1[1], r0[2]); v = at(b11 + b0[2], r1[0], r1[1], r0[2]); b = Lerp(sx, u, v);
    c = Lerp(sy, a, b);
    u = at(b00 + b1[2], r0[0], r0[1], r1[2]); v = at(b10 + b1[2], r1[0], r0[1], r1[2]); a = Lerp(sx, u, v);
    u = at(b01 + b1[2], r0[0], r1[1], r1[2]); v = at(b11 + b1[2], r1[0], r1[1], r1[2]); b = Lerp(sx, u, v);
    d = Lerp(sy, a, b);
    return Lerp(sz, c, d);
}
struct Lattice { int ix, iy, iz; double x_ix, y_iy, z_iz; };
inline void lattice_axis(double x, int& i, double& f)     // portablenoise.cpp:128-138
{
    int tmp = (x >= 0) ? (int)x : (int)(x - (1 - EPSILON));
    i = (int)((tmp - (-10000)) & 0xFFF);
    f = x - tmp;
}
#define HASH2D(a, b) (hT[(int)(hT[(int)(a)] ^ (b))])
#define RIDX(a, b) ((hT[(int)(a) ^ (b)] & 0xFF) * 2)
#define INCRSUMP(mp, s, x, y, z) ((s) * ((mp)[1] + (mp)[2] * (x) + (mp)[4] * (y) + (mp)[6] * (z)))
double Noise(const Scene& s, V3 P, int gen)               // PortableNoise portablenoise.cpp:103-225
{
    if (gen == 3) {
        double sum = 0.5 * (1.59 * SolidNoise(s, P) + 0.985);
        return std::min(std::max(sum, 0.0), 1.0);
    }
    const unsigned short* hT = s.hashTable.data();
    const double* RT = s.RTable.data();
    int ix, iy, iz; double x_ix, y_iy, z_iz;
    lattice_axis(P.x, ix, x_ix); lattice_axis(P.y, iy, y_iy); lattice_axis(P.z, iz, z_iz);
    double x_jx = x_ix - 1, y_jy = y_iy - 1, z_jz = z_iz - 1;
    double sx = SCURVE(x_ix), sy = SCURVE(y_iy), sz = SCURVE(z_iz), tx = 1 - sx, ty = 1 - sy, tz = 1 - sz;
    double txty = tx * ty, sxty = sx * ty, txsy = tx * sy, sxsy = sx * sy;
    int ixiy = HASH2D(ix, iy), jxiy = HASH2D(ix + 1, iy), ixjy = HASH2D(ix, iy + 1), jxjy = HASH2D(ix + 1, iy + 1);
    const double* mp; double sum;
    mp = RT + RIDX(ixiy, iz);     sum  = INCRSUMP(mp, (txty * tz), x_ix, y_iy, z_iz);
    mp = RT + RIDX(jxiy, iz);     sum += INCRSUMP(mp, (sxty * tz), x_jx, y_iy, z_iz);
    mp = RT + RIDX(ixjy, iz);     sum += INCRSUMP(mp, (txsy * tz), x_ix, y_jy, z_iz);
    mp = RT + RIDX(jxjy, iz);     sum += INCRSUMP(mp, (sxsy * tz), x_jx, y_jy, z_iz);
    mp = RT + RIDX(ixiy, iz + 1); sum += INCRSUMP(mp, (txty * sz), x_ix, y_iy, z_jz);
    mp = RT + RIDX(jxiy, iz + 1); sum += INCRSUMP(mp, (sxty * sz), x_jx, y_iy, z_jz);
    mp = RT + RIDX(ixjy, iz + 1); sum += INCRSUMP(mp, (txsy * sz), x_ix, y_jy, z_jz);
    mp = RT + RIDX(jxjy, iz + 1); sum += INCRSUMP(mp, (sxsy * sz), x_jx, y_jy, z_jz);
    if (gen == 2) { sum += 1.05242; sum *= 0.48985582; } else sum = sum + 0.5;
    if (sum < 0.0) sum = 0.0;
    if (sum > 1.0) sum = 1.0;
    return sum;
}
V3 DNoise(const Scene& s, V3 P)                           // PortableDNoise portablenoise.cpp:262-378
{
    const unsigned short* hT = s.hashTable.data();
    const double* RT = s.RTable.data();
    int ix, iy, iz; double x_ix, y_iy, z_iz;
    lattice_axis(P.x, ix, x_ix); lattice_axis(P.y, iy, y_iy); lattice_axis(P.z, iz, z_iz);
    double x_jx = x_ix - 1, y_jy = y_iy - 1, z_jz = z_iz - 1;
    double sx = SCURVE(x_ix), sy = SCURVE(y_iy), sz = SCURVE(z_iz), tx = 1 - sx, ty = 1 - sy, tz = 1 - sz;
    double txty = tx * ty, sxty = sx * ty, txsy = tx * sy, sxsy = sx * sy;
    int ixiy = HASH2D(ix, iy), jxiy = HASH2D(ix + 1, iy), ixjy = HASH2D(ix, iy + 1), jxjy = HASH2D(ix + 1, iy + 1);
    double r[3] = { 0, 0, 0 };
    bool first = true;
    auto corner = [&](int hash, int z, double sw, double fx, double fy, double fz) {
        const double* mp = RT + RIDX(hash, z);
        for (int k = 0; k < 3; k++, mp += 8) { double v = INCRSUMP(mp, sw, fx, fy, fz); if (first) r[k] = v; else r[k] += v; }
        first = false;
    };
    corner(ixiy, iz, txty * tz, x_ix, y_iy, z_iz); corner(jxiy, iz, sxty * tz, x_jx, y_iy, z_iz);
    corner(jxjy, iz, sxsy * tz, x_jx, y_jy, z_iz); corner(ixjy, iz, txsy * tz, x_ix, y_jy, z_iz);
    corner(ixjy, iz + 1, txsy * sz, x_ix, y_jy, z_jz); corner(jxjy, iz + 1, sxsy * sz, x_jx, y_jy, z_jz);
    corner(jxiy, iz + 1, sxty * sz, x_jx, y_iy, z_jz); corner(ixiy, iz + 1, txty * sz, x_ix, y_iy, z_jz);
    return v3(r[0], r[1], r[2]);
}
double Turbulence(const Scene& s, V3 P, const pvgpu_warp& T, int gen)      // noise.cpp:500-560
{
    double value;
    if (gen <= 1) value = Noise(s, P, gen);
    else { value = 2.0 * Noise(s, P, gen) - 0.5; value = std::min(std::max(value, 0.0), 1.0); }
    double Lambda = T.lambda, Omega = T.omega, l = Lambda, o = Omega;
    for (int i = 2; i <= T.octaves; i++) {
        V3 temp = P * l;
        if (gen <= 1) value += o * Noise(s, temp, gen); else value += o * (2.0 * Noise(s, temp, gen) - 0.5);
        if (i < T.octaves) { l *= Lambda; o *= Omega; }
    }
    return value;
}
V3 DTurbulence(const Scene& s, V3 P, const pvgpu_warp& T)                  // noise.cpp:580-610
{
    V3 result = DNoise(s, P);
    double Lambda = T.lambda, Omega = T.omega, l = Lambda, o = Omega;
    for (int i = 2; i <= T.octaves; i++) {
        V3 value = DNoise(s, P * l);
        result = result + o * value;
        if (i < T.octaves) { l *= Lambda; o *= Omega; }
    }
    return result;
}

// ---- rays, intersections ------------------------------------------------------------------------------
enum { RAY_PRIMARY = 1, RAY_REFLECTION = 2, RAY_REFRACTION = 4 };
struct Ray {
    V3 Origin, Direction;
    unsigned flags = RAY_PRIMARY;
    bool shadowTest = false;
    std::vector<int> interiors;
    bool IsImageRay() const { return (flags & RAY_PRIMARY) || ((flags & RAY_REFRACTION) && !(flags & RAY_REFLECTION)); }
    bool IsInterior(int i) const { return std::find(interiors.begin(), interiors.end(), i) != interiors.end(); }
    bool RemoveInterior(int i) { auto it = std::find(interiors.begin(), interiors.end(), i); if (it == interiors.end()) return false; interiors.erase(it); return true; }
    V3 Evaluate(double t) const { return v3(Origin.x + Direction.x * t, Origin.y + Direction.y * t, Origin.z + Direction.z * t); }
    bool IsHollowRay(const Scene& S) const { for (int i : interiors) if (!S.interiors[i].hollow) return false; return true; }   // ray.cpp:59-115
};
struct Ticket { unsigned traceLevel = 0, maxAllowedTraceLevel; double adcBailout; bool alphaBackground; unsigned maxFound = 0; };
struct Intersection { double Depth = BOUND_HUGE; V3 IPoint{ 0, 0, 0 }; int Object = -1, Csg = -1; uint32_t aux = 0; V3 INormal{ 0, 0, 0 }; double d1 = 0.0; };   // INormal: glyph hits (truetype.cpp stores it with the hit); d1: prism hits (spline parameter)
typedef std::vector<Intersection> IStack;

struct Stats { unsigned long long rays = 0, shadow_tests = 0; unsigned max_level = 0; };

// ---- TrueType glyph (truetype.cpp) ----------------------------------------------------------------------
// outline record in the shape-data table: segment count, then kind (0 line / 1 curve), x0 y0, x1 y1, x2 y2 per segment
static int ttf_solve_quad(const double* x, double* y, double mindist, double maxdist)                      // truetype.cpp:2567-2610
{
    const double COEFF_LIMIT = 1.0e-20;
    double a = x[0], b = -x[1], c = x[2];
    if (std::fabs(a) < COEFF_LIMIT) {
        if (std::fabs(b) < COEFF_LIMIT) return 0;
        double q = c / b;
        if (q >= mindist && q <= maxdist) { y[0] = q; return 1; }
        return 0;
    }
    double d = b * b - 4.0 * a * c;
    if (d < EPSILON) return 0;
    d = std::sqrt(d);
    double t = 2.0 * a, q = (b + d) / t;
    if (q >= mindist && q <= maxdist) {
        y[0] = q;
        q = (b - d) / t;
        if (q >= mindist && q <= maxdist) { y[1] = q; return 2; }
        return 1;
    }
    q = (b - d) / t;
    if (q >= mindist && q <= maxdist) { y[0] = q; return 1; }
    return 0;
}
static bool ttf_inside_glyph(const double* g, double x, double y)                                          // truetype.cpp:2392-2565
{
    int crossings = 0;
    const int n = (int)g[0];
    for (int j = 0; j < n; j++) {
        const double* e = g + 1 + 7 * j;
        double x0 = e[1], y0 = e[2], x1 = e[3], y1 = e[4];
        if (e[0] == 0.0) {                                                          // straight line: the crossing test
            if (y0 == y1) continue;
            int qi = (y0 < y) ? 1 : 0, qj = (y1 < y) ? 1 : 0;
            if (qi == qj) continue;
            int ri = (x0 > x) ? 1 : 0, rj = (x1 > x) ? 1 : 0;
            if (ri & rj) { crossings++; continue; }
            if ((ri | rj) == 0) continue;
            double m = (y1 - y0) / (x1 - x0), b = (y1 - y) - m * (x1 - x);
            if ((b / m) < EPSILON) crossings++;
        } else {
            double x2 = e[5], y2 = e[6], yt[3], xt[3], roots[2];
            if (((y0 < y) && (y1 < y) && (y2 < y)) || ((y0 > y) && (y1 > y) && (y2 > y))) continue;
            yt[0] = y0 - 2.0 * y1 + y2; yt[1] = 2.0 * (y1 - y0); yt[2] = y0 - y;
            int k = ttf_solve_quad(yt, roots, 0.0, 1.0);
            for (int ri = 0; ri < k;) {
                if (roots[ri] <= EPSILON) {
                    if (((y <= y0) && (y < y1)) || ((y >= y0) && (y > y1))) { k--; if (k > ri) roots[ri] = roots[ri + 1]; continue; }
                } else if (roots[ri] >= (1.0 - EPSILON)) {
                    if (((y < y2) && (y < y1)) || ((y > y2) && (y > y1))) { k--; if (k > ri) roots[ri] = roots[ri + 1]; continue; }
                }
                ri++;
            }
            if (k > 0) {
                xt[0] = x0 - 2.0 * x1 + x2; xt[1] = 2.0 * (x1 - x0); xt[2] = x0;
                double t = roots[0];
                if ((xt[0] * t + xt[1]) * t + xt[2] > x) crossings++;
                if (k > 1) { t = roots[1]; if ((xt[0] * t + xt[1]) * t + xt[2] > x) crossings++; }
            }
        }
    }
    return (crossings & 1) != 0;
}

// ---- prism (prism.cpp) ---------------------------------------------------------------------------------
// spline record in the shape-data table: Number, then per PRISM_SPLINE_ENTRY x1 y1 x2 y2, v1 u2 v2, A B C D (x y each)
struct PrismEntry { double x1, y1, x2, y2, v1, u2, v2, A[2], B[2], C[2], D[2]; };
static_assert(sizeof(PrismEntry) == 15 * sizeof(double), "PrismEntry mirrors 15 doubles of the shape-data table");
static int prism_in_curve(const pvgpu_object& ob, const double* sp, double u, double v)                   // prism.cpp:1154-1201
{
    int NC = 0;
    const double u1 = ob.p[6], v1 = ob.p[7], u2 = ob.p[8], v2 = ob.p[9];
    if ((u >= u1) && (u <= u2) && (v >= v1) && (v <= v2)) {
        const int Number = (int)sp[0];
        const PrismEntry* Entry = reinterpret_cast<const PrismEntry*>(sp + 1);
        for (int i = 0; i < Number; i++) {
            const PrismEntry& E = Entry[i];
            if ((v >= E.v1) && (v <= E.v2) && (u <= E.u2)) {
                double x[4] = { E.A[1], E.B[1], E.C[1], E.D[1] - v }, y[3];
                int n = Solve_Polynomial(3, x, y, (ob.flags & PVGPU_STURM_FLAG) ? 1 : 0, 0.0);
                while (n--) {
                    double w = y[n];
                    if ((w >= 0.0) && (w <= 1.0)) {
                        double k = w * (w * (w * E.A[0] + E.B[0]) + E.C[0]) + E.D[0] - u;
                        if (k >= 0.0) NC++;
                    }
                }
            }
        }
    }
    return NC & 1;
}
static bool prism_test_rectangle(V3 P, V3 D, double x1, double z1, double x2, double z2)                   // prism.cpp:1233-1345
{
    double dmin, dmax, tmin, tmax;
    if (std::fabs(D.x) > EPSILON) {
        if (D.x > 0.0) { dmin = (x1 - P.x) / D.x; dmax = (x2 - P.x) / D.x; if (dmax < EPSILON) return false; }
        else { dmax = (x1 - P.x) / D.x; if (dmax < EPSILON) return false; dmin = (x2 - P.x) / D.x; }
        if (dmin > dmax) return false;
    } else {
        if ((P.x < x1) || (P.x > x2)) return false;
        dmin = -BOUND_HUGE; dmax = BOUND_HUGE;
    }
    if (std::fabs(D.z) > EPSILON) {
        if (D.z > 0.0) { tmin = (z1 - P.z) / D.z; tmax = (z2 - P.z) / D.z; }
        else { tmax = (z1 - P.z) / D.z; tmin = (z2 - P.z) / D.z; }
        if (tmax < dmax) {
            if (tmax < EPSILON) return false;
            if (tmin > dmin) { if (tmin > tmax) return false; dmin = tmin; }
            else { if (dmin > tmax) return false; }
        } else {
            if (tmin > dmin) { if (tmin > dmax) return false; }
        }
    } else {
        if ((P.z < z1) || (P.z > z2)) return false;
    }
    return true;
}

// ---- superellipsoid (superellipsoid.cpp) ------------------------------------------------------------------
namespace superq {
const double DEPTH_TOLERANCE = 1.0e-4, ZERO_TOLERANCE = 1.0e-10, MIN_VALUE = -1.01, MAX_VALUE = 1.01;
const int MAX_ITERATIONS = 20, PLANECOUNT = 9;
const double planes[PLANECOUNT][4] = { { 1, 1, 0, 0 }, { 1, -1, 0, 0 }, { 1, 0, 1, 0 }, { 1, 0, -1, 0 }, { 0, 1, 1, 0 }, { 0, 1, -1, 0 }, { 1, 0, 0, 0 }, { 0, 1, 0, 0 }, { 0, 0, 1, 0 } };
static double power(double x, double e)                                                                    // :1097-1150
{
    double b = x;
    int i = (int)e;
    if (e == (double)i) {
        switch (i) {
            case 0: return 1.0;
            case 1: return b;
            case 2: return sqr(b);
            case 3: return sqr(b) * b;
            case 4: b *= b; return sqr(b);
            case 5: b *= b; return sqr(b) * x;
            case 6: b *= b; return sqr(b) * b;
            default: return std::pow(x, e);
        }
    }
    return std::pow(x, e);
}
static double evaluate_g(double x, double y, double e)                                                     // :1003-1021
{
    double g = 0;
    if (x > y) { g = 1 + power(y / x, e); if (g != 1) g = power(g, 1 / e); g *= x; }
    else if (y != 0) { g = 1 + power(x / y, e); if (g != 1) g = power(g, 1 / e); g *= y; }
    return g;
}
static double evaluate(const double* Power, V3 P) { return evaluate_g(evaluate_g(std::fabs(P.x), std::fabs(P.y), Power[0]), std::fabs(P.z), Power[2]) - 1; }   // :1059-1062
static bool intersect_box(V3 P, V3 D, double* dmin, double* dmax)                                          // :832-1001
{
    double tmin = 0.0, tmax = 0.0;
    if (std::fabs(D.x) > EPSILON) {
        if (D.x > EPSILON) { *dmin = (MIN_VALUE - P.x) / D.x; *dmax = (MAX_VALUE - P.x) / D.x; if (*dmax < EPSILON) return false; }
        else { *dmax = (MIN_VALUE - P.x) / D.x; if (*dmax < EPSILON) return false; *dmin = (MAX_VALUE - P.x) / D.x; }
        if (*dmin > *dmax) return false;
    } else {
        if ((P.x < MIN_VALUE) || (P.x > MAX_VALUE)) return false;
        *dmin = -BOUND_HUGE; *dmax = BOUND_HUGE;
    }
    const double Ps[2] = { P.y, P.z }, Ds[2] = { D.y, D.z };
    for (int a = 0; a < 2; a++) {                                                      // top / bottom, front / back
        if (std::fabs(Ds[a]) > EPSILON) {
            if (Ds[a] > EPSILON) { tmin = (MIN_VALUE - Ps[a]) / Ds[a]; tmax = (MAX_VALUE - Ps[a]) / Ds[a]; }
            else { tmax = (MIN_VALUE - Ps[a]) / Ds[a]; tmin = (MAX_VALUE - Ps[a]) / Ds[a]; }
            if (tmax < *dmax) {
                if (tmax < EPSILON) return false;
                if (tmin > *dmin) { if (tmin > tmax) return false; *dmin = tmin; }
                else { if (*dmin > tmax) return false; }
                *dmax = tmax;
            } else {
                if (tmin > *dmin) { if (tmin > *dmax) return false; *dmin = tmin; }
            }
        } else {
            if ((Ps[a] < MIN_VALUE) || (Ps[a] > MAX_VALUE)) return false;
        }
    }
    return true;
}
static void solve_hit1(const double* Power, double v0, V3 tP0, double v1, V3 tP1, V3& P)                    // :1344-1450
{
    V3 P0 = tP0, P1 = tP1, P2, P3;
    int i;
    for (i = 0; i < MAX_ITERATIONS; i++) {
        if (std::fabs(v0) < ZERO_TOLERANCE) { P = P0; break; }
        if (std::fabs(v1) < ZERO_TOLERANCE) { P = P1; break; }
        double x = std::fabs(v0) / std::fabs(v1 - v0);
        P2 = P1 - P0; P2 = P0 + P2 * x;
        double v2 = evaluate(Power, P2);
        P3 = P1 - P0; P3 = P0 + P3 * 0.5;
        double v3 = evaluate(Power, P3);
        if (v2 * v3 < 0.0) { v0 = v2; P0 = P2; v1 = v3; P1 = P3; }
        else if (std::fabs(v2) < std::fabs(v3)) { if (v0 * v2 < 0) { v1 = v2; P1 = P2; } else { v0 = v2; P0 = P2; } }
        else { if (v0 * v3 < 0) { v1 = v3; P1 = P3; } else { v0 = v3; P0 = P3; } }
    }
    if (i == MAX_ITERATIONS) P = (std::fabs(v0) < std::fabs(v1)) ? P0 : P1;
}
static bool check_hit2(const double* Power, V3 P, V3 D, double t0, V3& P0, double v0, double t1, double* t, V3& Q)   // :1485-1560
{
    double dt0 = t0, dt1 = t0 + 0.0001 * (t1 - t0), v1, deltat, maxdelta = t1 - t0;
    for (int i = 0; (dt0 < t1) && (i < MAX_ITERATIONS); i++) {
        V3 P1 = P + D * dt1;
        v1 = evaluate(Power, P1);
        if (v0 * v1 < 0) { solve_hit1(Power, v0, P0, v1, P1, Q); P0 = Q - P; *t = len(P0); return true; }
        if (std::fabs(v1) < ZERO_TOLERANCE) { Q = P + D * dt1; *t = dt1; return true; }
        if (((v0 > 0.0) && (v1 > v0)) || ((v0 < 0.0) && (v1 < v0))) break;
        if (v1 == v0) break;
        deltat = v1 * (dt1 - dt0) / (v1 - v0);
        if (std::fabs(deltat) > maxdelta) break;
        v0 = v1; dt0 = dt1; dt1 -= deltat;
    }
    return false;
}
}  // namespace superq

class Tracer {
public:
    const Scene& S;
    Stats st;
    explicit Tracer(const Scene& s) : S(s) {}

    // ---- primitives ----
    static bool sphere_intersect(V3 o, V3 d, V3 Center, double Radius2, double* Depth1, double* Depth2)   // sphere.cpp:211-243
    {
        V3 oc = Center - o;
        double OCSquared = len2(oc), t_ca = dot(oc, d);
        if ((OCSquared >= Radius2) && (t_ca < EPSILON)) return false;
        double thc2 = Radius2 - OCSquared + sqr(t_ca);
        if (thc2 > EPSILON) { double hc = std::sqrt(thc2); *Depth1 = t_ca - hc; *Depth2 = t_ca + hc; return true; }
        return false;
    }
    bool box_intersect(V3 P, V3 D, const double* c1, const double* c2, double* Depth1, double* Depth2, int* Side1, int* Side2) const;
    bool clip_ok(const pvgpu_object& o, V3 ip) const { return o.clip_count == 0 || Point_In_Clip(ip, o); }
    bool Point_In_Clip(V3 ip, const pvgpu_object& o) const                                               // object.cpp:430-443
    {
        for (uint32_t i = 0; i < o.clip_count; i++) if (!Inside_Object(ip, S.index_list[o.clip_first + i])) return false;
        return true;
    }
    bool Inside_Object(V3 ip, uint32_t idx) const                                                         // object.cpp:346-355
    {
        const pvgpu_object& o = S.objects[idx];
        for (uint32_t i = 0; i < o.clip_count; i++) if (!Inside_Object(ip, S.index_list[o.clip_first + i])) return false;
        return Inside(ip, idx);
    }
    bool Inside(V3 p, uint32_t idx) const;
    bool All_Intersections(uint32_t idx, const Ray& ray, IStack& stack) const;
    bool Find_Intersection(Intersection* isect, uint32_t idx, const Ray& ray, double post_min) const;
    bool Ray_In_Bound(const Ray& ray, const pvgpu_object& o) const                                        // object.cpp:385-400
    {
        for (uint32_t i = 0; i < o.bound_count; i++) {
            Intersection local;
            uint32_t b = S.index_list[o.bound_first + i];
            if (!Find_Intersection(&local, b, ray, -1.0) && !Inside_Object(ray.Origin, b)) return false;
        }
        return true;
    }
    bool tri_intersect(const pvgpu_mesh& me, const pvgpu_triangle& tr, V3 o, V3 d, double* Depth) const;
    bool mesh_intersect(uint32_t idx, const Ray& ray, IStack& stack) const;
    bool mesh_inside(const pvgpu_object& ob, V3 p) const;
    bool blob_intersect(uint32_t idx, const Ray& ray, IStack& stack) const;
    bool blob_element_hit(const pvgpu_blob_element& e, V3 P, V3 D, double mindist, double* tmin, double* tmax) const;
    double blob_element_field(const pvgpu_blob_element& e, V3 P) const;
    void blob_element_normal(const pvgpu_blob_element& e, V3 P, V3& Result) const;
    template <class F> void blob_walk_point(const pvgpu_blob& bl, V3 P, F&& leaf) const;
    V3 Normal(const Intersection& isect) const;

    // ---- tree ----
    struct Rayinfo { float origin[3], invDirection[3]; bool nonzero[3], positive[3]; };
    static Rayinfo make_rayinfo(V3 o, V3 d)                                                               // boundingbox.h:182-216
    {
        Rayinfo ri;
        for (int k = 0; k < 3; k++) {
            ri.origin[k] = (float)o[k];
            double t = d[k];
            ri.nonzero[k] = (t != 0.0);
            ri.invDirection[k] = ri.nonzero[k] ? (float)(1.0 / t) : 0.0f;
            ri.positive[k] = (t > 0.0);
        }
        return ri;
    }
    struct Qelem { double depth; uint32_t node; };
    static void heap_insert(std::vector<Qelem>& q, double depth, uint32_t node)                           // boundingbox.cpp:88-103
    {
        size_t i = q.size();
        q.push_back(Qelem{});
        while ((i > 1) && (depth < q[i / 2].depth)) { q[i] = q[i / 2]; i /= 2; }
        q[i].depth = depth; q[i].node = node;
    }
    static bool heap_remove_min(std::vector<Qelem>& q, double& depth, uint32_t& node)                     // boundingbox.cpp:105-137
    {
        size_t size = q.size() - 1;
        if (size == 0) return false;
        depth = q[1].depth; node = q[1].node;
        size_t i = 1, j;
        while (i <= size / 2) {
            if ((2 * i == size) || (q[2 * i].depth < q[2 * i + 1].depth)) j = 2 * i; else j = 2 * i + 1;
            if (q[size].depth <= q[j].depth) break;
            q[i] = q[j];
            i = j;
        }
        if (i != size) q[i] = q[size];
        q.pop_back();
        return true;
    }
    static void Check_And_Enqueue(std::vector<Qelem>& q, const pvgpu_node* nodes, uint32_t ni, const Rayinfo& ri)   // boundingbox.cpp:541-648
    {
        const pvgpu_node& n = nodes[ni];
        double dmin, dmax;
        if (!(n.flags & PVGPU_NODE_INFINITE)) {
            dmin = -BOUND_HUGE; dmax = BOUND_HUGE;
            for (int dim = 0; dim < 3; dim++) {
                const float lo = n.lo[dim], hi = n.lo[dim] + n.size[dim];
                if (ri.nonzero[dim]) {
                    double tmin, tmax;
                    if (ri.positive[dim]) {
                        tmax = (hi - ri.origin[dim]) * ri.invDirection[dim];
                        if (tmax < EPSILON) return;
                        tmin = (lo - ri.origin[dim]) * ri.invDirection[dim];
                    } else {
                        tmax = (lo - ri.origin[dim]) * ri.invDirection[dim];
                        if (tmax < EPSILON) return;
                        tmin = (hi - ri.origin[dim]) * ri.invDirection[dim];
                    }
                    if (tmax < dmax) {
                        if (tmin > dmin) { if (tmin > tmax) return; dmin = tmin; }
                        else if (dmin > tmax) return;
                        dmax = tmax;
                    } else if (tmin > dmin) { if (tmin > dmax) return; dmin = tmin; }
                } else if (!((lo <= ri.origin[dim]) && (ri.origin[dim] <= hi))) return;
            }
        } else dmin = -MAX_DISTANCE;
        heap_insert(q, dmin, ni);
    }
    bool precondition(const Ray& ray, const pvgpu_object& o) const
    {
        if (ray.shadowTest) return !(o.flags & PVGPU_NO_SHADOW_FLAG);                                     // trace.cpp:1943
        if (ray.IsImageRay() && (o.flags & PVGPU_NO_IMAGE_FLAG)) return false;                            // trace.cpp:84-95
        if ((ray.flags & RAY_REFLECTION) && (o.flags & PVGPU_NO_REFLECTION_FLAG)) return false;
        return true;
    }
    bool FindIntersection(Intersection& best, const Ray& ray, double post_min) const;                     // trace.cpp:285-344

    // ---- shading ----
    double TraceRay(Ray& ray, Ticket& tk, Col& colour, float& transm, float weight, bool continuedRay, double maxDepth = 0.0);
    void ComputeTextureColour(Intersection& isect, Col& colour, float& transm, Ray& ray, Ticket& tk, float weight);
    void ComputeLightedTexture(Col& resultColour, float& resultTransm, int texture, V3 ipoint, V3 rawnormal, Ray& ray, Ticket& tk, float weight, Intersection& isect,
                               const std::vector<int>& warps = std::vector<int>());
    void ComputeOneTextureColour(Col& resultColour, float& resultTransm, int texture, std::vector<int>& warps, V3 ipoint, V3 rawnormal, Ray& ray, Ticket& tk,
                                 float weight, Intersection& isect, bool shadowflag);
    void ComputeShadowTexture(Col& filtercolour, int texture, const std::vector<int>& warps, V3 ipoint, V3 rawnormal, Ray& lray, Intersection& isect);
    V3 Warp_Normal_Chain(V3 n, const std::vector<int>& warps, bool unwarp) const;
    void ComputeSky(const Ray& ray, const Ticket& tk, Col& colour, float& transm) const;
    void ComputeFog(const Ray& ray, double Depth, Col& colour, float& transm) const;
    bool Compute_Pigment(float col[5], int pigment, V3 EPoint) const;
    // the intersection whose textures are being evaluated (the reference hands `Intersect` down to Compute_Pigment for uv_mapping)
    mutable const Intersection* cur_isect = nullptr;
    V3 UVCoord(const Intersection& isect) const;
    // camera { normal { ... } }: the tail of TracePixel::CreateCameraRay (tracepixel.cpp:917-924), perspective / orthographic cameras
    void camera_normal(double x, double y, double width, double height, V3& Direction) const
    {
        if (!S.cam.reserved) return;
        const double x0 = x / width - 0.5, y0 = 0.5 - y / height;
        // (camera_ray has normalised the direction once, like the reference does before Perturb_Normal)
        Direction = unit(Perturb_Normal(Direction, (int)S.cam.reserved - 1, v3(x0, y0, 0.0)));
    }
    bool image_map_colour(const pvgpu_image& im, V3 p, float col[5]) const;
    V3 Warp_EPoint(const pvgpu_pigment& pg, V3 EPoint) const;
    double Evaluate_TPat(const pvgpu_pigment& pg, V3 p) const;
    V3 Perturb_Normal(V3 Layer_Normal, int tnormal, V3 EPoint) const;
    double relative_ior(const Ray& ray, int interior) const;
    void ComputeReflection(V3 ipoint, Ray& ray, Ticket& tk, V3 normal, V3 rawnormal, Col& colour, float weight, int finish = -1);
    void ComputeIridColour(const pvgpu_finish& fn, V3 lightDirection, V3 eyeDirection, V3 layer_normal, V3 ipoint, Col& colour) const;
    bool ComputeRefraction(int interior, V3 ipoint, Ray& ray, Ticket& tk, V3 normal, V3 rawnormal, Col& colour, float& transm, float weight, int finish = -1);
    void ComputeOneDiffuseLight(const pvgpu_light& L, const pvgpu_finish& fn, V3 ipoint, const Ray& eye, Ticket& tk, V3 layer_normal,
                                Col layer_pigment_colour, Col& colour, double attenuation, const pvgpu_object& object, double relativeIor,
                                std::pair<bool, Col>* light_cache);
    void TraceShadowRay(const pvgpu_light& L, double depth, Ray& lightsourceray, Ticket& tk, Col& colour);
    void TracePointLightShadowRay(double& lightsourcedepth, Ray& newray, Ticket& tk, Col& colour);
    void TraceAreaLightShadowRay(const pvgpu_light& L, double& lightsourcedepth, Ray& lightsourceray, V3 ipoint, Ticket& tk, Col& lightcolour);
    void TraceAreaLightSubsetShadowRay(const pvgpu_light& L, double& lightsourcedepth, Ray& lightsourceray, V3 ipoint, Ticket& tk, Col& lightcolour,
                                       int u1, int v1, int u2, int v2, int level, V3 axis1, V3 axis2, std::vector<Col>& lightGrid);
    void ComputeOneWhiteLightRay(const pvgpu_light& L, double& depth, Ray& lray, V3 ipoint, V3 jitter) const;
    void ComputeShadowColour(Intersection& isect, Ray& lightsourceray, const Ticket& tk, Col& colour);
    int hit_texture(const pvgpu_object& ob, const Intersection& isect, bool backside) const;
    std::vector<std::pair<float, int>> Determine_Textures(const pvgpu_object& ob, const Intersection& isect, bool backside) const;
};

// Box::Intersect (box.cpp:167-520)
bool Tracer::box_intersect(V3 P, V3 D, const double* c1, const double* c2, double* Depth1, double* Depth2, int* Side1, int* Side2) const
{
    const double CLOSE_TOLERANCE = 1.0e-6, DEPTH_TOLERANCE = 1.0e-6;
    int smin = 0, smax = 0;
    double t, tmin = 0.0, tmax = BOUND_HUGE;
    const double Pv[3] = { P.x, P.y, P.z }, Dv[3] = { D.x, D.y, D.z };
    for (int ax = 0; ax < 3; ax++) {
        const double close = (ax == 0) ? 0.0 : CLOSE_TOLERANCE;
        const int s0 = 2 * ax + 1, s1 = 2 * ax + 2;
        auto prefer = [&](int current) {
            // which axis is the currently recorded side on?  Y compares against X unconditionally, Z against the recorded one
            if (ax == 1) return std::fabs(Dv[1]) > std::fabs(Dv[0]);
            if (current == 1 || current == 2) return std::fabs(Dv[2]) > std::fabs(Dv[0]);
            if (current == 3 || current == 4) return std::fabs(Dv[2]) > std::fabs(Dv[1]);
            return false;
        };
        if (Dv[ax] < -EPSILON || Dv[ax] > EPSILON) {
            const bool neg = Dv[ax] < -EPSILON;
            const double farc = neg ? c1[ax] : c2[ax], nearc = neg ? c2[ax] : c1[ax];
            const int fars = neg ? s0 : s1, nears = neg ? s1 : s0;
            t = (farc - Pv[ax]) / Dv[ax];
            if (t < tmin) return false;
            if (ax == 0) { if (t <= tmax) { smax = fars; tmax = t; } }
            else if (t <= tmax - close) { smax = fars; tmax = t; }
            else if (t <= tmax + close) { if (prefer(smax)) smax = fars; }
            t = (nearc - Pv[ax]) / Dv[ax];
            if (ax == 0) { if (t >= tmin) { if (t > tmax) return false; smin = nears; tmin = t; } }
            else if (t >= tmin + close) { if (t > tmax) return false; smin = nears; tmin = t; }
            else if (t >= tmin - close) { if (prefer(smin)) smin = nears; }
        } else if ((Pv[ax] < c1[ax]) || (Pv[ax] > c2[ax])) return false;
    }
    if (tmax < DEPTH_TOLERANCE) return false;
    *Depth1 = tmin; *Depth2 = tmax; *Side1 = smin; *Side2 = smax;
    return true;
}

// Mesh::intersect_mesh_triangle (mesh.cpp:1040-1127)
bool Tracer::tri_intersect(const pvgpu_mesh& me, const pvgpu_triangle& tr, V3 o, V3 d, double* Depth) const
{
    const float* N = S.norms.data() + 3 * (size_t)me.normal_first;
    const float* V = S.verts.data() + 3 * (size_t)me.vertex_first;
    V3 n = v3f(N + 3 * tr.normal_ind);
    double ndd = dot(n, d);
    if (std::fabs(ndd) < EPSILON) return false;
    double ndo = dot(n, o);
    *Depth = -((double)tr.distance + ndo) / ndd;
    if ((*Depth < 1.0e-6) || (*Depth > MAX_DISTANCE)) return false;
    V3 P1 = v3f(V + 3 * tr.p1), P2 = v3f(V + 3 * tr.p2), P3 = v3f(V + 3 * tr.p3);
    const int a = (tr.dominant_axis == 0) ? 1 : 0, b = (tr.dominant_axis == 2) ? 1 : 2;
    double s = o[a] + *Depth * d[a], t = o[b] + *Depth * d[b];
    if ((P2[a] - s) * (P2[b] - P1[b]) < (P2[b] - t) * (P2[a] - P1[a])) return false;
    if ((P3[a] - s) * (P3[b] - P2[b]) < (P3[b] - t) * (P3[a] - P2[a])) return false;
    if ((P1[a] - s) * (P1[b] - P3[b]) < (P1[b] - t) * (P1[a] - P3[a])) return false;
    return true;
}

// Mesh::Intersect + intersect_bbox_tree + test_hit (mesh.cpp:155-195, 1452-1528, 1208-1243)
bool Tracer::mesh_intersect(uint32_t idx, const Ray& ray, IStack& stack) const
{
    const pvgpu_object& ob = S.objects[idx];
    const pvgpu_mesh& me = S.meshes[ob.mesh];
    V3 mo = ray.Origin, md = ray.Direction;
    double len_ = 1.0;
    if (ob.transform >= 0) {
        const pvgpu_transform& t = S.xf[ob.transform];
        mo = MInvTransPoint(t, ray.Origin); md = MInvTransDirection(t, ray.Direction);
        len_ = len(md); md = md / len_;
    }
    bool found = false;
    auto test_hit = [&](uint32_t ti, double Depth) {
        double world_dist = Depth / len_;
        V3 ip = ray.Evaluate(world_dist);
        if (clip_ok(ob, ip)) { Intersection is; is.Depth = world_dist; is.IPoint = ip; is.Object = (int)idx; is.aux = me.triangle_first + ti; stack.push_back(is); return true; }
        return false;
    };
    if (me.node_count == 0) {
        for (uint32_t i = 0; i < me.triangle_count; i++) {
            double t;
            if (tri_intersect(me, S.tris[me.triangle_first + i], mo, md, &t) && test_hit(i, t)) found = true;
        }
        return found;
    }
    const pvgpu_node* nodes = S.mnodes.data() + me.node_first;
    Rayinfo ri = make_rayinfo(mo, md);
    std::vector<Qelem> q(1);
    double Best = BOUND_HUGE, Depth;
    const bool OldStyle = me.has_inside_vector != 0;
    Check_And_Enqueue(q, nodes, 0, ri);
    uint32_t ni;
    while (heap_remove_min(q, Depth, ni)) {
        if (!OldStyle && Depth > Best) break;
        const pvgpu_node& n = nodes[ni];
        if (n.count) { for (uint32_t i = 0; i < n.count; i++) Check_And_Enqueue(q, nodes, n.first + i, ri); }
        else if (tri_intersect(me, S.tris[me.triangle_first + n.first], mo, md, &Depth) && test_hit(n.first, Depth)) { found = true; Best = Depth; }
    }
    return found;
}

// Mesh::Inside (mesh.cpp:197-262)
bool Tracer::mesh_inside(const pvgpu_object& ob, V3 p) const
{
    const pvgpu_mesh& me = S.meshes[ob.mesh];
    if (!me.has_inside_vector) return false;
    V3 mo = p, md = v3(me.inside_vector);
    if (ob.transform >= 0) { const pvgpu_transform& t = S.xf[ob.transform]; mo = MInvTransPoint(t, p); md = unit(MInvTransDirection(t, md)); }
    unsigned found = 0;
    // (the tree variant counts the same triangles; the linear form is enough for a checker)
    for (uint32_t i = 0; i < me.triangle_count; i++) { double t; if (tri_intersect(me, S.tris[me.triangle_first + i], mo, md, &t)) found++; }
    bool inside = (found & 1) != 0;
    if (ob.flags & PVGPU_INVERTED_FLAG) inside = !inside;
    return inside;
}

bool Tracer::Inside(V3 p, uint32_t idx) const
{
    const pvgpu_object& ob = S.objects[idx];
    const bool inv = (ob.flags & PVGPU_INVERTED_FLAG) != 0;
    switch (ob.type) {
        case PVGPU_OBJ_SPHERE: {                                                                          // sphere.cpp:262-300
            double oc2 = ob.aux ? len2(MInvTransPoint(S.xf[ob.transform], p)) : len2(v3(ob.p) - p);
            return inv ? (oc2 > sqr(ob.p[3])) : (oc2 < sqr(ob.p[3]));
        }
        case PVGPU_OBJ_BOX: {                                                                             // box.cpp:538-575
            V3 q = (ob.transform >= 0) ? MInvTransPoint(S.xf[ob.transform], p) : p;
            if ((q.x < ob.p[0]) || (q.x > ob.p[3])) return inv;
            if ((q.y < ob.p[1]) || (q.y > ob.p[4])) return inv;
            if ((q.z < ob.p[2]) || (q.z > ob.p[5])) return inv;
            return !inv;
        }
        case PVGPU_OBJ_PLANE: {                                                                           // plane.cpp:208-225
            double temp = (ob.transform < 0) ? dot(p, v3(ob.p)) : dot(MInvTransPoint(S.xf[ob.transform], p), v3(ob.p));
            return (temp + ob.p[3]) < EPSILON;
        }
        case PVGPU_OBJ_QUADRIC: {                                                                         // quadric.cpp:248-255
            const double* c = ob.p;
            return (p.x * (c[0] * p.x + c[3] * p.y + c[6]) + p.y * (c[1] * p.y + c[5] * p.z + c[7]) + p.z * (c[2] * p.z + c[4] * p.x + c[8]) + c[9]) <= 0.0;
        }
        case PVGPU_OBJ_TORUS: {                                                                           // torus.cpp:348-400
            V3 P = MInvTransPoint(S.xf[ob.transform], p);
            double r = std::sqrt(sqr(P.x) + sqr(P.z)), r2 = sqr(P.y) + sqr(r - ob.p[0]);
            bool inside = false;
            if (r2 <= sqr(ob.p[1])) {
                inside = true;
                if (ob.aux & 0x20u) { bool insp = (sqr(P.y) + sqr(r + ob.p[0]) <= sqr(ob.p[1])); inside = (ob.aux & 0x04u) ? insp : !insp; }
            }
            return inside ? !inv : inv;
        }
        case PVGPU_OBJ_POLY: {                                                                            // polynomial.cpp:590-654, 1131-1178
            const double* a = S.shape_data.data() + ob.mesh;
            const int order = (int)ob.aux;
            const V3 P = MInvTransPoint(S.xf[ob.transform], p);
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
            return (result < 1.0e-4) ? !inv : inv;
        }
        case PVGPU_OBJ_SUPERELLIPSOID:                                                                    // superellipsoid.cpp:396-420
            return (superq::evaluate(ob.p, MInvTransPoint(S.xf[ob.transform], p)) < EPSILON) ? !inv : inv;
        case PVGPU_OBJ_PRISM: {                                                                           // prism.cpp:640-672
            V3 P = MInvTransPoint(S.xf[ob.transform], p);
            if ((P.y >= ob.p[0]) && (P.y < ob.p[1])) {
                if (((ob.aux >> 4) & 15u) == 2u) {
                    if (std::fabs(P.y) > EPSILON) { P.x /= P.y; P.z /= P.y; } else P.x = P.z = HUGE_VALUE;
                }
                if (prism_in_curve(ob, S.shape_data.data() + ob.mesh, P.x, P.z)) return !inv;
            }
            return inv;
        }
        case PVGPU_OBJ_GLYPH: {                                                                           // truetype.cpp:2957-2970
            V3 q = MInvTransPoint(S.xf[ob.transform], p);
            return (q.z >= 0.0 && q.z <= ob.p[0] && ttf_inside_glyph(S.shape_data.data() + ob.mesh, q.x, q.y)) ? !inv : inv;
        }
        case PVGPU_OBJ_DISC:                                                                              // disc.cpp:200-224
            return (MInvTransPoint(S.xf[ob.transform], p).z >= 0.0) ? inv : !inv;
        case PVGPU_OBJ_CONE: {                                                                            // cone.cpp:333-390
            const double offset = (ob.flags & PVGPU_CLOSED_FLAG) ? -EPSILON : EPSILON;
            V3 q = MInvTransPoint(S.xf[ob.transform], p);
            double w2 = q.x * q.x + q.y * q.y;
            bool outside = (ob.flags & PVGPU_CYLINDER_FLAG) ? ((w2 > 1.0 + offset) || (q.z < 0.0 - offset) || (q.z > 1.0 + offset))
                                                            : ((w2 > q.z * q.z + offset) || (q.z < ob.p[0] - offset) || (q.z > 1.0 + offset));
            return outside ? inv : !inv;
        }
        case PVGPU_OBJ_MESH: return mesh_inside(ob, p);
        case PVGPU_OBJ_BLOB: {                                                                            // blob.cpp:1502-1624
            const pvgpu_blob& bl = S.blobs[ob.mesh];
            V3 P = (ob.transform >= 0) ? MInvTransPoint(S.xf[ob.transform], p) : p;
            double density = 0.0;
            blob_walk_point(bl, P, [&](const pvgpu_blob_element& e) { density += blob_element_field(e, P); });
            return (density > bl.threshold - 1.0e-6) ? !inv : inv;
        }
        case PVGPU_OBJ_CSG_UNION:
        case PVGPU_OBJ_CSG_MERGE:                                                                         // csg.cpp:393-410
            for (uint32_t i = 0; i < ob.child_count; i++) if (Inside_Object(p, S.index_list[ob.child_first + i])) return true;
            return false;
        case PVGPU_OBJ_CSG_INTERSECTION:                                                                  // csg.cpp:428-440
            for (uint32_t i = 0; i < ob.child_count; i++) if (!Inside_Object(p, S.index_list[ob.child_first + i])) return false;
            return true;
    }
    return false;
}

static bool test_ray_flags(const Ray& ray, uint32_t oflags, bool shadow_variant)                          // csg.cpp:80-104
{
    const bool image = !ray.shadowTest && ray.IsImageRay(), primary = !ray.shadowTest && (ray.flags & RAY_PRIMARY), refl = !ray.shadowTest && (ray.flags & RAY_REFLECTION);
    bool ok = (!(oflags & PVGPU_NO_IMAGE_FLAG) || !image || (!shadow_variant && primary)) && (!(oflags & PVGPU_NO_REFLECTION_FLAG) || !refl);
    if (shadow_variant && ray.shadowTest && !(oflags & PVGPU_NO_SHADOW_FLAG)) ok = true;
    return ok;
}

bool Tracer::All_Intersections(uint32_t idx, const Ray& ray, IStack& Depth_Stack) const
{
    const pvgpu_object& ob = S.objects[idx];
    const V3 o = ray.Origin, d = ray.Direction;
    auto push = [&](double depth, V3 ip, uint32_t aux) {
        if (clip_ok(ob, ip)) { Intersection is; is.Depth = depth; is.IPoint = ip; is.Object = (int)idx; is.aux = aux; Depth_Stack.push_back(is); return true; }
        return false;
    };
    bool found = false;
    switch (ob.type) {
        case PVGPU_OBJ_SPHERE: {                                                                          // sphere.cpp:92-170
            double d1, d2;
            if (ob.aux) {
                const pvgpu_transform& t = S.xf[ob.transform];
                V3 no = MInvTransPoint(t, o), nd = MInvTransDirection(t, d);
                double l = len(nd); nd = nd / l;
                if (sphere_intersect(no, nd, v3(0, 0, 0), 1.0, &d1, &d2)) {
                    if ((d1 > 1.0e-6) && (d1 < MAX_DISTANCE)) found |= push(d1 / l, MTransPoint(t, v3(no.x + nd.x * d1, no.y + nd.y * d1, no.z + nd.z * d1)), 0);
                    if ((d2 > 1.0e-6) && (d2 < MAX_DISTANCE)) found |= push(d2 / l, MTransPoint(t, v3(no.x + nd.x * d2, no.y + nd.y * d2, no.z + nd.z * d2)), 0);
                }
            } else if (sphere_intersect(o, d, v3(ob.p), sqr(ob.p[3]), &d1, &d2)) {
                if ((d1 > 1.0e-6) && (d1 < MAX_DISTANCE)) found |= push(d1, ray.Evaluate(d1), 0);
                if ((d2 > 1.0e-6) && (d2 < MAX_DISTANCE)) found |= push(d2, ray.Evaluate(d2), 0);
            }
            return found;
        }
        case PVGPU_OBJ_BOX: {                                                                             // box.cpp:100-149
            V3 P = o, D = d;
            if (ob.transform >= 0) { P = MInvTransPoint(S.xf[ob.transform], o); D = MInvTransDirection(S.xf[ob.transform], d); }
            double d1, d2; int s1, s2;
            if (box_intersect(P, D, ob.p, ob.p + 3, &d1, &d2, &s1, &s2)) {
                if (d1 > 1.0e-6) found |= push(d1, ray.Evaluate(d1), (uint32_t)s1);
                found |= push(d2, ray.Evaluate(d2), (uint32_t)s2);
            }
            return found;
        }
        case PVGPU_OBJ_PLANE: {                                                                           // plane.cpp:92-190
            V3 n = v3(ob.p);
            double ndd, ndo;
            if (ob.transform < 0) { ndd = dot(n, d); if (std::fabs(ndd) < EPSILON) return false; ndo = dot(n, o); }
            else {
                V3 P = MInvTransPoint(S.xf[ob.transform], o), D = MInvTransDirection(S.xf[ob.transform], d);
                ndd = dot(n, D); if (std::fabs(ndd) < EPSILON) return false; ndo = dot(n, P);
            }
            double depth = -(ndo + ob.p[3]) / ndd;
            if ((depth >= 1.0e-6) && (depth <= MAX_DISTANCE)) return push(depth, ray.Evaluate(depth), 0);
            return false;
        }
        case PVGPU_OBJ_QUADRIC: {                                                                         // quadric.cpp:123-230
            const double QA = ob.p[0], QE = ob.p[1], QH = ob.p[2], QB = ob.p[3], QC = ob.p[4], QF = ob.p[5], QD = ob.p[6], QG = ob.p[7], QI = ob.p[8], QJ = ob.p[9];
            const double Xo = o.x, Yo = o.y, Zo = o.z, Xd = d.x, Yd = d.y, Zd = d.z;
            double a = Xd * (QA * Xd + QB * Yd + QC * Zd) + Yd * (QE * Yd + QF * Zd) + Zd * QH * Zd;
            double b = Xd * (QA * Xo + 0.5 * (QB * Yo + QC * Zo + QD)) + Yd * (QE * Yo + 0.5 * (QB * Xo + QF * Zo + QG)) + Zd * (QH * Zo + 0.5 * (QC * Xo + QF * Yo + QI));
            double c = Xo * (QA * Xo + QB * Yo + QC * Zo + QD) + Yo * (QE * Yo + QF * Zo + QG) + Zo * (QH * Zo + QI) + QJ;
            double d1, d2;
            if (a != 0.0) { double dd = sqr(b) - a * c; if (dd <= 0.0) return false; dd = std::sqrt(dd); d1 = (-b + dd) / a; d2 = (-b - dd) / a; }
            else { if (b == 0.0) return false; d1 = -0.5 * c / b; d2 = MAX_DISTANCE; }
            if ((d1 > 1.0e-6) && (d1 < MAX_DISTANCE)) found |= push(d1, ray.Evaluate(d1), 0);
            if ((d2 > 1.0e-6) && (d2 < MAX_DISTANCE)) found |= push(d2, ray.Evaluate(d2), 0);
            return found;
        }
        case PVGPU_OBJ_TORUS: {                                                                           // torus.cpp:133-330, 932-1059
            const pvgpu_transform& t = S.xf[ob.transform];
            const double R = ob.p[0], r = ob.p[1];
            V3 P = MInvTransPoint(t, o), D = MInvTransDirection(t, d);
            double l = len(D); D = D / l;
            double y1 = -r, y2 = r, r1 = sqr(R - r); if (R < r) r1 = 0; double r2 = sqr(R + r);
            auto thick = [&]() {
                double a, b, c, dd, u, v, k, rr, h;
                if (std::fabs(D.y) < EPSILON) { if ((P.y < y1) || (P.y > y2)) return false; }
                else {
                    k = (y2 - P.y) / D.y; u = P.x + k * D.x; v = P.z + k * D.z;
                    if ((k > EPSILON) && (k < MAX_DISTANCE)) { rr = u * u + v * v; if ((rr >= r1) && (rr <= r2)) return true; }
                    k = (y1 - P.y) / D.y; u = P.x + k * D.x; v = P.z + k * D.z;
                    if ((k > EPSILON) && (k < MAX_DISTANCE)) { rr = u * u + v * v; if ((rr >= r1) && (rr <= r2)) return true; }
                }
                a = D.x * D.x + D.z * D.z;
                if (a > EPSILON) {
                    b = P.x * D.x + P.z * D.z;
                    for (int pass = 0; pass < 2; pass++) {
                        c = P.x * P.x + P.z * P.z - (pass == 0 ? r2 : r1);
                        dd = b * b - a * c;
                        if (dd >= 0.0) {
                            dd = std::sqrt(dd);
                            k = (-b + dd) / a; if ((k > EPSILON) && (k < MAX_DISTANCE)) { h = P.y + k * D.y; if ((h >= y1) && (h <= y2)) return true; }
                            k = (-b - dd) / a; if ((k > EPSILON) && (k < MAX_DISTANCE)) { h = P.y + k * D.y; if ((h >= y1) && (h <= y2)) return true; }
                        }
                    }
                }
                return false;
            };
            if (!thick()) return false;
            double bsr = R + r + r, DistanceP = len2(P), Closer = 0.0;
            if (DistanceP > sqr(bsr)) { DistanceP = std::sqrt(DistanceP); Closer = DistanceP - bsr; P = P + Closer * D; }
            double R2 = sqr(R); r2 = sqr(r);
            double Py2 = P.y * P.y, Dy2 = D.y * D.y, PDy2 = P.y * D.y;
            double k1 = P.x * P.x + P.z * P.z + Py2 - R2 - r2, k2 = P.x * D.x + P.z * D.z + PDy2;
            double c[5], rt[4];
            c[0] = 1.0; c[1] = 4.0 * k2; c[2] = 2.0 * (k1 + 2.0 * (k2 * k2 + R2 * Dy2)); c[3] = 4.0 * (k2 * k1 + 2.0 * R2 * PDy2); c[4] = k1 * k1 + 4.0 * R2 * (Py2 - r2);
            int n = Solve_Polynomial(4, c, rt, (ob.flags & PVGPU_STURM_FLAG) ? 1 : 0, 1.0e-4);
            while (n--) {
                double depth = (rt[n] + Closer) / l;
                if ((depth > 1.0e-4) && (depth < MAX_DISTANCE)) {
                    V3 ip = ray.Evaluate(depth);
                    uint32_t aux = 0;
                    if (ob.aux) {
                        if (!clip_ok(ob, ip)) continue;
                        bool on = len2(MInvTransPoint(t, ip)) < ob.p[2];
                        if (!(on ? (ob.aux & 1u) : (ob.aux & 2u))) continue;
                        aux = on ? 1u : 0u;
                    }
                    found |= push(depth, ip, aux);
                }
            }
            return found;
        }
        case PVGPU_OBJ_DISC: {                                                                            // disc.cpp:90-180
            const pvgpu_transform& t = S.xf[ob.transform];
            V3 P = MInvTransPoint(t, o), D = MInvTransDirection(t, d);
            const double length = len(D);
            D = D / length;
            if (std::fabs(D.z) > EPSILON) {
                double tt = -P.z / D.z;
                if (tt >= 0.0) {
                    double u = P.x + tt * D.x, v = P.y + tt * D.y, r2 = sqr(u) + sqr(v);
                    if ((r2 >= ob.p[3]) && (r2 <= ob.p[4])) {
                        double Depth = tt / length;
                        if ((Depth > 1.0e-6) && (Depth < MAX_DISTANCE)) found |= push(Depth, ray.Evaluate(Depth), 0);
                    }
                }
            }
            return found;
        }
        case PVGPU_OBJ_POLY: {                                                                            // polynomial.cpp:211-290, 656-947
            const pvgpu_transform& t = S.xf[ob.transform];
            const double* a = S.shape_data.data() + ob.mesh;
            const int order = (int)ob.aux;
            V3 O = MInvTransPoint(t, o), D = MInvTransDirection(t, d);
            const double length = len(D);
            D = D / length;
            double depths[4];
            int cnt = 0;
            if (order == 1) {
                double t0 = a[0] * O.x + a[1] * O.y + a[2] * O.z;
                double t1 = a[0] * D.x + a[1] * D.y + a[2] * D.z;
                if (std::fabs(t1) < EPSILON) return false;
                depths[0] = -(a[3] + t0) / t1;
                cnt = 1;
            } else if (order == 2) cnt = [&]() -> int {
    const double x = O.x, y = O.y, z = O.z, xx = D.x, yy = D.y, zz = D.z;
    const double x2 = x * x, y2 = y * y, z2 = z * z, xx2 = xx * xx, yy2 = yy * yy, zz2 = zz * zz;
    double ac = (a[0]*xx2 + a[1]*xx*yy + a[2]*xx*zz + a[4]*yy2 + a[5]*yy*zz + a[7]*zz2);
    double bc = (2*a[0]*x*xx + a[1]*(x*yy + xx*y) + a[2]*(x*zz + xx*z) +
                 a[3]*xx + 2*a[4]*y*yy + a[5]*(y*zz + yy*z) + a[6]*yy +
                 2*a[7]*z*zz + a[8]*zz);
    double cc = a[0]*x2 + a[1]*x*y + a[2]*x*z + a[3]*x + a[4]*y2 +
                a[5]*y*z + a[6]*y + a[7]*z2 + a[8]*z + a[9];
    if (std::fabs(ac) < 1.0e-20) {
        if (std::fabs(bc) < 1.0e-20) return 0;
        depths[0] = -cc / bc;
        return 1;
    }
    double dd = bc * bc - 4.0 * ac * cc;
    if (dd < 0.0) return 0;
    dd = std::sqrt(dd);
    bc = -bc;
    const double t = 2.0 * ac;
    depths[0] = (bc + dd) / t;
    depths[1] = (bc - dd) / t;
    return 2;
            }();
            else cnt = [&]() -> int {
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
                return Solve_Polynomial(deg, &eqn[lead], depths, (ob.flags & PVGPU_STURM_FLAG) ? 1 : 0, 1.0e-4);
            }();
            for (int i = 0; i < cnt; i++) {
                if (!(depths[i] > 1.0e-4)) continue;
                bool same_root = false;
                for (int j = 0; j < i; j++) if (depths[i] == depths[j]) { same_root = true; break; }
                if (same_root) continue;
                V3 IPoint = MTransPoint(t, v3(O.x + D.x * depths[i], O.y + D.y * depths[i], O.z + D.z * depths[i]));
                found |= push(depths[i] / length, IPoint, 0);
            }
            return found;
        }
        case PVGPU_OBJ_TRIANGLE: {                                                                        // triangle.cpp:447-590
            if (ob.flags & PVGPU_DEGENERATE_FLAG) return false;
            const double* T = S.shape_data.data() + ob.mesh;
            const V3 Normal_Vector = v3(T + 9), P1 = v3(T), P2 = v3(T + 3), P3 = v3(T + 6);
            double NormalDotDirection = dot(Normal_Vector, d);
            if (std::fabs(NormalDotDirection) < EPSILON) return false;
            double NormalDotOrigin = dot(Normal_Vector, o);
            double Depth = -(T[12] + NormalDotOrigin) / NormalDotDirection;
            if ((Depth < 1.0e-6) || (Depth > MAX_DISTANCE)) return false;
            const int dom = ob.aux & 3;
            const int A = (dom == 0) ? 1 : 0, B = (dom == 2) ? 1 : 2;
            double ss = o[A] + Depth * d[A], tt = o[B] + Depth * d[B];
            if ((P2[A] - ss) * (P2[B] - P1[B]) < (P2[B] - tt) * (P2[A] - P1[A])) return false;
            if ((P3[A] - ss) * (P3[B] - P2[B]) < (P3[B] - tt) * (P3[A] - P2[A])) return false;
            if ((P1[A] - ss) * (P1[B] - P3[B]) < (P1[B] - tt) * (P1[A] - P3[A])) return false;
            return push(Depth, ray.Evaluate(Depth), 0);
        }
        case PVGPU_OBJ_POLYGON: {                                                                         // polygon.cpp:131-260, 905-980
            if (ob.flags & PVGPU_DEGENERATE_FLAG) return false;
            const pvgpu_transform& t = S.xf[ob.transform];
            V3 P = MInvTransPoint(t, o), D = MInvTransDirection(t, d);
            const double length = len(D);
            D = D / length;
            if (std::fabs(D.z) < 1.0e-10) return false;
            double Depth = -P.z / D.z;
            if ((Depth < 1.0e-8) || (Depth > MAX_DISTANCE)) return false;
            const double tx = P.x + Depth * D.x, ty = P.y + Depth * D.y;
            const int number = (int)ob.aux;
            const double (*points)[2] = reinterpret_cast<const double (*)[2]>(S.shape_data.data() + ob.mesh);
            const double *vtx0 = points[0], *vtx1 = points[1], *first = vtx0;
            int yflag0 = (vtx0[1] >= ty), yflag1;
            bool inside_flag = false;
            for (int i = 1; i < number; ) {
                yflag1 = (vtx1[1] >= ty);
                if (yflag0 != yflag1)
                    if (((vtx1[1] - ty) * (vtx0[0] - vtx1[0]) >= (vtx1[0] - tx) * (vtx0[1] - vtx1[1])) == yflag1) inside_flag = !inside_flag;
                if ((i < number - 2) && (vtx1[0] == first[0]) && (vtx1[1] == first[1])) {
                    vtx0 = points[++i]; vtx1 = points[++i];
                    yflag0 = (vtx0[1] >= ty);
                    first = vtx0;
                } else {
                    vtx0 = vtx1; vtx1 = points[++i];
                    yflag0 = yflag1;
                }
            }
            if (!inside_flag) return false;
            Depth /= length;
            return push(Depth, ray.Evaluate(Depth), 0);
        }
        case PVGPU_OBJ_CONE: {                                                                            // cone.cpp:103-330
            const pvgpu_transform& t = S.xf[ob.transform];
            V3 P = MInvTransPoint(t, o), D = MInvTransDirection(t, d);
            const double length = len(D), dist = ob.p[0], tol = 1.0e-9;
            D = D / length;
            const bool cyl = (ob.flags & PVGPU_CYLINDER_FLAG) != 0;
            const double zlo = cyl ? 0.0 : dist;
            struct { double d; uint32_t t; } I[4]; int n = 0;
            auto side = [&](double tt) { double z = P.z + tt * D.z; if ((tt > tol) && (tt < MAX_DISTANCE) && (z >= zlo) && (z <= 1.0)) { I[n].d = tt / length; I[n++].t = 3; } };
            if (cyl) {
                double a = D.x * D.x + D.y * D.y;
                if (a > EPSILON) {
                    double b = P.x * D.x + P.y * D.y, c = P.x * P.x + P.y * P.y - 1.0, dd = b * b - a * c;
                    if (dd >= 0.0) { dd = std::sqrt(dd); side((-b + dd) / a); side((-b - dd) / a); }
                }
            } else {
                double a = D.x * D.x + D.y * D.y - D.z * D.z, b = D.x * P.x + D.y * P.y - D.z * P.z, c = P.x * P.x + P.y * P.y - P.z * P.z;
                if (std::fabs(a) < EPSILON) { if (std::fabs(b) > EPSILON) side(-0.5 * c / b); }
                else { double dd = b * b - a * c; if (dd >= 0.0) { dd = std::sqrt(dd); side((-b - dd) / a); side((-b + dd) / a); } }
            }
            if ((ob.flags & PVGPU_CLOSED_FLAG) && (std::fabs(D.z) > EPSILON)) {
                double dd = (1.0 - P.z) / D.z, a = P.x + dd * D.x, b = P.y + dd * D.y;
                if (((sqr(a) + sqr(b)) <= 1.0) && (dd > tol) && (dd < MAX_DISTANCE)) { I[n].d = dd / length; I[n++].t = 2; }
                dd = (dist - P.z) / D.z; a = P.x + dd * D.x; b = P.y + dd * D.y;
                if ((sqr(a) + sqr(b)) <= (cyl ? 1.0 : sqr(dist)) && (dd > tol) && (dd < MAX_DISTANCE)) { I[n].d = dd / length; I[n++].t = 1; }
            }
            for (int i = 0; i < n; i++) found |= push(I[i].d, ray.Evaluate(I[i].d), I[i].t);
            return found;
        }
        case PVGPU_OBJ_GLYPH: {                                                                           // truetype.cpp:2706-2955
            const double TTF_Tolerance = 1.0e-6;
            const pvgpu_transform& tr = S.xf[ob.transform];
            const V3 P = MInvTransPoint(tr, o), D = MInvTransDirection(tr, d);
            const double* g = S.shape_data.data() + ob.mesh;
            const double glyph_depth = ob.p[0];
            auto hit = [&](double Depth, V3 N, uint32_t aux) {
                V3 IPoint = ray.Evaluate(Depth);
                if (Depth > TTF_Tolerance && clip_ok(ob, IPoint)) {
                    Intersection is; is.Depth = Depth; is.IPoint = IPoint; is.Object = (int)idx; is.aux = aux; is.INormal = unit(MTransNormal(tr, N));
                    Depth_Stack.push_back(is); found = true;
                }
            };
            double t0 = -1.0, t1 = -1.0;                                                                   // GetZeroOneHits, :2630-2660
            if (!(std::fabs(D.z) < EPSILON)) {
                double t = -P.z / D.z;
                if (ttf_inside_glyph(g, P.x + t * D.x, P.y + t * D.y)) t0 = t;
                t += (glyph_depth / D.z);
                if (ttf_inside_glyph(g, P.x + t * D.x, P.y + t * D.y)) t1 = t;
            }
            if (t0 > 0.0) hit(t0, v3(0.0, 0.0, -1.0), 0u);
            if (t1 > 0.0) hit(t1, v3(0.0, 0.0, 1.0), 1u);
            int dirflag;
            if (std::fabs(D.x) < EPSILON) { if (std::fabs(D.y) < EPSILON) return found; dirflag = 0; } else dirflag = 1;
            const double a = D.y, b = -D.x, c = (P.y * D.x - P.x * D.y);
            const int n = (int)g[0];
            for (int j = 0; j < n; j++) {
                const double* e = g + 1 + 7 * j;
                const double x0 = e[1], y0 = e[2], x1 = e[3], y1 = e[4];
                if (e[0] == 0.0) {
                    double d0 = (x1 - x0), d1 = (y1 - y0);
                    t0 = d1 * D.x - d0 * D.y;
                    if (std::fabs(t0) < EPSILON) continue;
                    double t = (D.x * (P.y - y0) - D.y * (P.x - x0)) / t0;
                    if (t < 0.0 || t > 1.0) continue;
                    if (dirflag) t = ((x0 + t * d0) - P.x) / D.x; else t = ((y0 + t * d1) - P.y) / D.y;
                    double z = P.z + t * D.z;
                    if (z >= 0 && z <= glyph_depth && t > TTF_Tolerance) hit(t, v3(-d1, d0, 0.0), 2u | ((uint32_t)j << 3));
                } else {
                    const double x2 = e[5], y2 = e[6];
                    double xt2 = x0 - 2.0 * x1 + x2, xt1 = 2.0 * (x1 - x0), xt0 = x0, yt2 = y0 - 2.0 * y1 + y2, yt1 = 2.0 * (y1 - y0), yt0 = y0;
                    double C[3] = { a * xt2 + b * yt2, a * xt1 + b * yt1, a * xt0 + b * yt0 + c }, Sr[2];
                    int k = ttf_solve_quad(C, Sr, 0.0, 1.0);
                    for (int l = 0; l < k; l++) {
                        double t;
                        if (dirflag) t = ((Sr[l] * Sr[l] * xt2 + Sr[l] * xt1 + xt0) - P.x) / D.x;
                        else t = ((Sr[l] * Sr[l] * yt2 + Sr[l] * yt1 + yt0) - P.y) / D.y;
                        double z = P.z + t * D.z;
                        if (z >= 0 && z <= glyph_depth && t > TTF_Tolerance) hit(t, v3(-2.0 * yt2 * Sr[l] - yt1, 2.0 * xt2 * Sr[l] + xt1, 0.0), 2u | ((uint32_t)l << 2) | ((uint32_t)j << 3));
                    }
                }
            }
            return found;
        }
        case PVGPU_OBJ_PRISM: {                                                                           // prism.cpp:194-596
            if (ob.flags & PVGPU_DEGENERATE_FLAG) return false;
            const double DEPTH_TOLERANCE = 1.0e-4;
            const pvgpu_transform& tr = S.xf[ob.transform];
            V3 P = MInvTransPoint(tr, o), D = MInvTransDirection(tr, d);
            const double length = len(D);
            D = D / length;
            const double Height1 = ob.p[0], Height2 = ob.p[1], x1 = ob.p[2], y1 = ob.p[3], x2 = ob.p[4], y2 = ob.p[5];
            if (((D.x >= 0.0) && (P.x > x2)) || ((D.x <= 0.0) && (P.x < x1)) || ((D.z >= 0.0) && (P.z > y2)) || ((D.z <= 0.0) && (P.z < y1))) return false;
            const double* sp = S.shape_data.data() + ob.mesh;
            const int Number = (int)sp[0];
            const PrismEntry* Entry = reinterpret_cast<const PrismEntry*>(sp + 1);
            const uint32_t Spline_Type = ob.aux & 15u, Sweep_Type = (ob.aux >> 4) & 15u;
            const int sturm = (ob.flags & PVGPU_STURM_FLAG) ? 1 : 0;
            auto hit = [&](double k, uint32_t aux, double w) {
                double distance = k / length;
                if ((distance > DEPTH_TOLERANCE) && (distance < MAX_DISTANCE)) {
                    V3 IPoint = ray.Evaluate(distance);
                    if (clip_ok(ob, IPoint)) { Intersection is; is.Depth = distance; is.IPoint = IPoint; is.Object = (int)idx; is.aux = aux; is.d1 = w; Depth_Stack.push_back(is); found = true; }
                }
            };
            auto plane = [&](double Height, uint32_t aux) {                                                // cap / base plane
                if (Sweep_Type == 2u && !(std::fabs(Height) > EPSILON)) return;
                double k = (Height - P.y) / D.y;
                if ((k > DEPTH_TOLERANCE) && (k < MAX_DISTANCE)) {
                    double u = P.x + k * D.x, v = P.z + k * D.z;
                    if (Sweep_Type == 2u) { u = u / Height; v = v / Height; }
                    if (prism_in_curve(ob, sp, u, v)) hit(k, aux, 0.0);
                }
            };
            if (std::fabs(D.y) < EPSILON) { if ((P.y < Height1) || (P.y > Height2)) return false; }
            else if (ob.flags & PVGPU_CLOSED_FLAG) { plane(Height2, 1u); plane(Height1, 0u); }
            const double k1 = P.z * D.y - P.y * D.z, k2 = P.y * D.x - P.x * D.y, k3 = P.x * D.z - P.z * D.x;
            if (Sweep_Type == 1u && !((std::fabs(D.x) > EPSILON) || (std::fabs(D.z) > EPSILON))) return found;
            for (int j = 0; j < Number; j++) {
                const PrismEntry& E = Entry[j];
                if (((D.x >= 0.0) && (P.x > E.x2)) || ((D.x <= 0.0) && (P.x < E.x1)) || ((D.z >= 0.0) && (P.z > E.y2)) || ((D.z <= 0.0) && (P.z < E.y1))) continue;
                int n = 0;
                double x[4], y[3];
                // linear sweep: coefficient pairs (X, Y) weighted with (D.z, -D.x); conic sweep: with (k1, k2) and the constant k3
                auto coef = [&](const double* c) { return (Sweep_Type == 1u) ? c[0] * D.z - c[1] * D.x : c[0] * k1 + c[1] * k2; };
                const double last = (Sweep_Type == 1u) ? D.z * (E.D[0] - P.x) - D.x * (E.D[1] - P.z) : E.D[0] * k1 + E.D[1] * k2 + k3;
                switch (Spline_Type) {
                    case 1: x[0] = coef(E.C); x[1] = last; if (std::fabs(x[0]) > EPSILON) y[n++] = -x[1] / x[0]; break;
                    case 2: x[0] = coef(E.B); x[1] = coef(E.C); x[2] = last; n = Solve_Polynomial(2, x, y, 0, 0.0); break;
                    default:
                        if (Sweep_Type == 2u || prism_test_rectangle(P, D, E.x1, E.y1, E.x2, E.y2)) {
                            x[0] = coef(E.A); x[1] = coef(E.B); x[2] = coef(E.C); x[3] = last;
                            n = Solve_Polynomial(3, x, y, sturm, 0.0);
                        }
                        break;
                }
                while (n--) {
                    const double w = y[n];
                    if ((w >= 0.0) && (w <= 1.0)) {
                        double k, h;
                        if (Sweep_Type == 1u) {
                            if (std::fabs(D.x) > EPSILON) k = (w * (w * (w * E.A[0] + E.B[0]) + E.C[0]) + E.D[0] - P.x) / D.x;
                            else k = (w * (w * (w * E.A[1] + E.B[1]) + E.C[1]) + E.D[1] - P.z) / D.z;
                        } else {
                            k = w * (w * (w * E.A[0] + E.B[0]) + E.C[0]) + E.D[0];
                            h = D.x - k * D.y;
                            if (std::fabs(h) > EPSILON) k = (k * P.y - P.x) / h;
                            else {
                                k = w * (w * (w * E.A[1] + E.B[1]) + E.C[1]) + E.D[1];
                                h = D.z - k * D.y;
                                if (std::fabs(h) > EPSILON) k = (k * P.y - P.z) / h; else continue;
                            }
                        }
                        h = P.y + k * D.y;
                        if ((h >= Height1) && (h <= Height2)) hit(k, 2u | ((uint32_t)n << 2) | ((uint32_t)j << 4), w);
                    }
                }
            }
            return found;
        }
        case PVGPU_OBJ_SUPERELLIPSOID: {                                                                  // superellipsoid.cpp:213-394
            using namespace superq;
            const pvgpu_transform& tr = S.xf[ob.transform];
            V3 P = MInvTransPoint(tr, o), D = MInvTransDirection(tr, d), P0, P1, P2, P3;
            const double length = len(D);
            D = D / length;
            double t, t1, t2, v0, v1, dists[PLANECOUNT + 2];
            if (!intersect_box(P, D, &t1, &t2)) return false;
            if (t2 < DEPTH_TOLERANCE) return false;
            int cnt = 0;
            if (t1 < DEPTH_TOLERANCE) t1 = DEPTH_TOLERANCE;
            dists[cnt++] = t1; dists[cnt++] = t2;
            {                                                                                              // find_ray_plane_points :1271-1310
                double mindist = t1, maxdist = t2;
                t = EPSILON * (maxdist - mindist);
                mindist -= t; maxdist += t;
                for (int i = 0; i < PLANECOUNT; i++) {
                    double dd = (D.x * planes[i][0] + D.y * planes[i][1] + D.z * planes[i][2]);
                    if (std::fabs(dd) < EPSILON) continue;
                    t = (planes[i][3] - (P.x * planes[i][0] + P.y * planes[i][1] + P.z * planes[i][2])) / dd;
                    if ((t >= mindist) && (t <= maxdist)) dists[cnt++] = t;
                }
                std::sort(dists, dists + cnt);
            }
            if (cnt <= 1) return false;
            const bool child = (ob.aux & 1u) != 0;
            auto insert_hit = [&](double Depth) {                                                          // :1172-1190
                if ((Depth > DEPTH_TOLERANCE) && (Depth < MAX_DISTANCE)) return push(Depth, ray.Evaluate(Depth), 0);
                return false;
            };
            P0 = P + D * dists[0];
            v0 = evaluate(ob.p, P0);
            if (std::fabs(v0) < ZERO_TOLERANCE) { if (insert_hit(dists[0] / length)) { if (child) found = true; else return true; } }
            for (int i = 1; i < cnt; i++) {
                P1 = P + D * dists[i];
                v1 = evaluate(ob.p, P1);
                if (std::fabs(v1) < ZERO_TOLERANCE) { if (insert_hit(dists[i] / length)) { if (child) found = true; else return true; } }
                else if (v0 * v1 < 0.0) {
                    solve_hit1(ob.p, v0, P0, v1, P1, P2);
                    P3 = P2 - P;
                    t = len(P3);
                    if (insert_hit(t / length)) { if (child) found = true; else return true; }
                } else if (check_hit2(ob.p, P, D, dists[i - 1], P0, v0, dists[i], &t, P2)) {
                    if (insert_hit(t / length)) { if (child) found = true; else return true; }
                    else break;
                }
                v0 = v1; P0 = P1;
            }
            return found;
        }
        case PVGPU_OBJ_MESH: return mesh_intersect(idx, ray, Depth_Stack);
        case PVGPU_OBJ_BLOB: return blob_intersect(idx, ray, Depth_Stack);
        case PVGPU_OBJ_CSG_UNION: {                                                                       // csg.cpp:128-189
            for (uint32_t i = 0; i < ob.child_count; i++) {
                uint32_t ch = S.index_list[ob.child_first + i];
                const pvgpu_object& co = S.objects[ch];
                if (!test_ray_flags(ray, co.flags, false)) continue;
                if (co.bound_count && !Ray_In_Bound(ray, co)) continue;
                if (ob.clip_count == 0) { if (All_Intersections(ch, ray, Depth_Stack)) found = true; }
                else {
                    IStack local;
                    if (All_Intersections(ch, ray, local))
                        while (!local.empty()) {
                            if (Point_In_Clip(local.back().IPoint, ob)) { local.back().Csg = (int)idx; Depth_Stack.push_back(local.back()); found = true; }
                            local.pop_back();
                        }
                }
            }
            return found;
        }
        case PVGPU_OBJ_CSG_INTERSECTION: {                                                                // csg.cpp:219-277
            for (uint32_t i = 0; i < ob.child_count; i++) {
                uint32_t ch = S.index_list[ob.child_first + i];
                const pvgpu_object& co = S.objects[ch];
                if (co.bound_count && !Ray_In_Bound(ray, co)) continue;
                IStack local;
                if (All_Intersections(ch, ray, local))
                    while (!local.empty()) {
                        bool maybe = true;
                        for (uint32_t k = 0; k < ob.child_count; k++) {
                            uint32_t sib = S.index_list[ob.child_first + k];
                            if (sib != ch && !Inside_Object(local.back().IPoint, sib)) { maybe = false; break; }
                        }
                        if (maybe && (ob.clip_count == 0 || Point_In_Clip(local.back().IPoint, ob))) { local.back().Csg = (int)idx; Depth_Stack.push_back(local.back()); found = true; }
                        local.pop_back();
                    }
            }
            return found;
        }
        case PVGPU_OBJ_CSG_MERGE: {                                                                       // csg.cpp:307-375
            for (uint32_t i = 0; i < ob.child_count; i++) {
                uint32_t ch = S.index_list[ob.child_first + i];
                const pvgpu_object& co = S.objects[ch];
                if (!test_ray_flags(ray, co.flags, true)) continue;
                if (co.bound_count && !Ray_In_Bound(ray, co)) continue;
                IStack local;
                if (All_Intersections(ch, ray, local))
                    while (!local.empty()) {
                        if (ob.clip_count == 0 || Point_In_Clip(local.back().IPoint, ob)) {
                            bool inside_flag = true;
                            for (uint32_t k = 0; k < ob.child_count && inside_flag; k++) {
                                uint32_t sib = S.index_list[ob.child_first + k];
                                if (sib != ch && test_ray_flags(ray, S.objects[sib].flags, true) && Inside_Object(local.back().IPoint, sib)) inside_flag = false;
                            }
                            if (inside_flag) { local.back().Csg = (int)idx; found = true; Depth_Stack.push_back(local.back()); }
                        }
                        local.pop_back();
                    }
            }
            return found;
        }
    }
    return false;
}

// ObjectBase::Intersect_BBox / Intersect_BBox_Dir (object.cpp:917-941, 1074-1109)
static bool Intersect_BBox(const pvgpu_object& ob, V3 o, V3 d, float maxd)
{
    if (ob.type < PVGPU_OBJ_QUADRIC) return true;          // Sphere / Box / Plane override it (sphere.cpp:753, box.cpp:1079, plane.cpp:629)
    if (ob.type == PVGPU_OBJ_TRIANGLE || ob.type == PVGPU_OBJ_POLY) return true;   // and so do Triangle (triangle.cpp:1419) and Poly (polynomial.cpp:1510)
    float origin[3] = { (float)o.x, (float)o.y, (float)o.z };
    float invdir[3] = { (float)(1.0 / d.x), (float)(1.0 / d.y), (float)(1.0 / d.z) };
    float b[2][3] = { { ob.bbox[0], ob.bbox[1], ob.bbox[2] }, { ob.bbox[0] + ob.bbox[3], ob.bbox[1] + ob.bbox[4], ob.bbox[2] + ob.bbox[5] } };
    const int BX = invdir[0] < 0.0f, BY = invdir[1] < 0.0f, BZ = invdir[2] < 0.0f;
    float tmin = (b[BX][0] - origin[0]) * invdir[0], tmax = (b[1 - BX][0] - origin[0]) * invdir[0];
    float tymin = (b[BY][1] - origin[1]) * invdir[1], tymax = (b[1 - BY][1] - origin[1]) * invdir[1];
    if ((tmin > tymax) || (tymin > tmax)) return false;
    if (tymin > tmin) tmin = tymin;
    if (tymax < tmax) tmax = tymax;
    float tzmin = (b[BZ][2] - origin[2]) * invdir[2], tzmax = (b[1 - BZ][2] - origin[2]) * invdir[2];
    if ((tmin > tzmax) || (tzmin > tmax)) return false;
    if (tzmin > tmin) tmin = tzmin;
    if (tzmax < tmax) tmax = tzmax;
    return (tmin < maxd) && (tmax > (float)MIN_ISECT_DEPTH);
}

// Find_Intersection (object.cpp:118-224); post_min < 0: no post-condition, else SmallToleranceRayObjectCondition
bool Tracer::Find_Intersection(Intersection* isect, uint32_t idx, const Ray& ray, double post_min) const
{
    const pvgpu_object& ob = S.objects[idx];
    double closest = HUGE_VALUE;
    if (!Intersect_BBox(ob, ray.Origin, ray.Direction, (float)closest)) return false;
    if (ob.bound_count && !Ray_In_Bound(ray, ob)) return false;
    IStack depthstack;
    if (All_Intersections(idx, ray, depthstack)) {
        bool found = false;
        while (!depthstack.empty()) {
            double tmpDepth = depthstack.back().Depth;
            if (tmpDepth < closest && tmpDepth >= MIN_ISECT_DEPTH && tmpDepth > post_min) { *isect = depthstack.back(); closest = tmpDepth; found = true; }
            depthstack.pop_back();
        }
        return found;
    }
    return false;
}

// Trace::FindIntersection + Intersect_BBox_Tree (trace.cpp:285-344, boundingbox.cpp:485-539)
bool Tracer::FindIntersection(Intersection& best, const Ray& ray, double post_min) const
{
    bool found = false;
    if (!S.use_tree) {
        for (uint32_t f : S.frame) {
            if (!precondition(ray, S.objects[f])) continue;
            Intersection isect;
            if (Find_Intersection(&isect, f, ray, post_min) && (isect.Depth < best.Depth)) { best = isect; found = true; }
        }
        return found;
    }
    Rayinfo ri = make_rayinfo(ray.Origin, ray.Direction);
    std::vector<Qelem> q(1);
    Check_And_Enqueue(q, S.nodes.data(), 0, ri);
    double Depth; uint32_t ni;
    while (heap_remove_min(q, Depth, ni)) {
        if (Depth > best.Depth) break;
        const pvgpu_node& n = S.nodes[ni];
        if (n.count) { for (uint32_t i = 0; i < n.count; i++) Check_And_Enqueue(q, S.nodes.data(), n.first + i, ri); }
        else if (precondition(ray, S.objects[n.first])) {
            Intersection isect;
            if (Find_Intersection(&isect, n.first, ray, post_min) && isect.Depth < best.Depth) { best = isect; found = true; }
        }
    }
    return found;
}

// ---- blob (blob.cpp) -------------------------------------------------------------------------------------
// intersect_element and the four component intersectors (blob.cpp:716-1267)
bool Tracer::blob_element_hit(const pvgpu_blob_element& e, V3 P, V3 D, double mindist, double* tmin, double* tmax) const
{
    *tmin = BOUND_HUGE; *tmax = -BOUND_HUGE;
    double b, d, t, length = 1.0;
    V3 PP = P, DD = D;
    if (e.type != PVGPU_BLOB_SPHERE) {
        const pvgpu_transform& tr = S.xf[e.transform];
        PP = MInvTransPoint(tr, P); DD = MInvTransDirection(tr, D);
        length = len(DD); DD = DD / length;
    }
    switch (e.type) {
        case PVGPU_BLOB_SPHERE:
        case PVGPU_BLOB_ELLIPSOID: {
            V3 V1 = PP - v3(e.o);
            b = dot(V1, DD); t = len2(V1); d = b * b - t + e.rad2;
            if (d < EPSILON) return false;
            d = std::sqrt(d);
            if (e.type == PVGPU_BLOB_SPHERE) { *tmax = -b + d; *tmin = -b - d; } else { *tmax = (-b + d) / length; *tmin = (-b - d) / length; }
            if (*tmax < mindist) *tmax = 0.0;
            if (*tmin < mindist) *tmin = 0.0;
            if (*tmax == *tmin) return false;
            if (*tmax < *tmin) std::swap(*tmin, *tmax);
            return true;
        }
        case PVGPU_BLOB_BASE_HEMISPHERE:
        case PVGPU_BLOB_APEX_HEMISPHERE: {
            const bool base = e.type == PVGPU_BLOB_BASE_HEMISPHERE;
            if (!base) PP.z -= e.len;
            b = dot(PP, DD); t = len2(PP); d = b * b - t + e.rad2;
            if (d < EPSILON) return false;
            d = std::sqrt(d);
            *tmax = -b + d; *tmin = -b - d;
            if (*tmax < *tmin) std::swap(*tmin, *tmax);
            double z1 = PP.z + *tmin * DD.z, z2 = PP.z + *tmax * DD.z;
            bool in1 = base ? (z1 >= 0.0) : (z1 <= 0.0), in2 = base ? (z2 >= 0.0) : (z2 <= 0.0);      // "inside" = beyond the cutting plane
            bool out1 = base ? (z1 < 0.0) : (z1 > 0.0), out2 = base ? (z2 < 0.0) : (z2 > 0.0);
            if (in1 && in2) return false;
            if (out1 && out2) { *tmin /= length; *tmax /= length; return true; }
            t = -PP.z / DD.z;
            if (in1) *tmin = (t < mindist) ? 0.0 : t; else *tmax = (t < mindist) ? 0.0 : t;
            *tmin /= length; *tmax /= length;
            return true;
        }
        default: {      // cylinder
            double a = DD.x * DD.x + DD.y * DD.y, u, v, w;
            auto take = [&](double tt) { if (tt < *tmin) *tmin = tt; if (tt > *tmax) *tmax = tt; };
            if (a > EPSILON) {
                b = PP.x * DD.x + PP.y * DD.y;
                double c = PP.x * PP.x + PP.y * PP.y - e.rad2;
                d = b * b - a * c;
                if (d > EPSILON) {
                    d = std::sqrt(d);
                    t = (-b + d) / a; w = PP.z + t * DD.z; if ((w >= 0.0) && (w <= e.len)) take(t);
                    t = (-b - d) / a; w = PP.z + t * DD.z; if ((w >= 0.0) && (w <= e.len)) take(t);
                }
            }
            if (std::fabs(DD.z) > EPSILON) {
                t = -PP.z / DD.z; u = PP.x + t * DD.x; v = PP.y + t * DD.y; if ((u * u + v * v) <= e.rad2) take(t);
                t = (e.len - PP.z) / DD.z; u = PP.x + t * DD.x; v = PP.y + t * DD.y; if ((u * u + v * v) <= e.rad2) take(t);
            }
            *tmin /= length; *tmax /= length;
            if (*tmin < mindist) *tmin = 0.0;
            if (*tmax < mindist) *tmax = 0.0;
            return !(*tmin >= *tmax);
        }
    }
}

// calculate_element_field (blob.cpp:1379-1500)
double Tracer::blob_element_field(const pvgpu_blob_element& e, V3 P) const
{
    auto f = [&](double rad2) { return rad2 * (rad2 * e.c[0] + e.c[1]) + e.c[2]; };
    if (e.type == PVGPU_BLOB_SPHERE) { double r2 = len2(P - v3(e.o)); return (r2 < e.rad2) ? f(r2) : 0.0; }
    V3 PP = MInvTransPoint(S.xf[e.transform], P);
    switch (e.type) {
        case PVGPU_BLOB_ELLIPSOID: { double r2 = len2(PP - v3(e.o)); return (r2 < e.rad2) ? f(r2) : 0.0; }
        case PVGPU_BLOB_BASE_HEMISPHERE: if (PP.z <= 0.0) { double r2 = len2(PP); if (r2 <= e.rad2) return f(r2); } return 0.0;
        case PVGPU_BLOB_APEX_HEMISPHERE: PP.z -= e.len; if (PP.z >= 0.0) { double r2 = len2(PP); if (r2 <= e.rad2) return f(r2); } return 0.0;
        default: if ((PP.z >= 0.0) && (PP.z <= e.len)) { double r2 = sqr(PP.x) + sqr(PP.y); if (r2 <= e.rad2) return f(r2); } return 0.0;
    }
}

// element_normal (blob.cpp:1677-1813)
void Tracer::blob_element_normal(const pvgpu_blob_element& e, V3 P, V3& Result) const
{
    auto val = [&](double dist) { return -2.0 * e.c[0] * dist - e.c[1]; };
    if (e.type == PVGPU_BLOB_SPHERE) { V3 V1 = P - v3(e.o); double dist = len2(V1); if (dist <= e.rad2) Result = Result + V1 * val(dist); return; }
    const pvgpu_transform& tr = S.xf[e.transform];
    V3 PP = MInvTransPoint(tr, P);
    switch (e.type) {
        case PVGPU_BLOB_ELLIPSOID: { V3 V1 = PP - v3(e.o); double dist = len2(V1); if (dist <= e.rad2) Result = Result + MTransNormal(tr, V1) * val(dist); return; }
        case PVGPU_BLOB_BASE_HEMISPHERE: if (PP.z <= 0.0) { double dist = len2(PP); if (dist <= e.rad2) Result = Result + MTransNormal(tr, PP) * val(dist); } return;
        case PVGPU_BLOB_APEX_HEMISPHERE: PP.z -= e.len; if (PP.z >= 0.0) { double dist = len2(PP); if (dist <= e.rad2) Result = Result + MTransNormal(tr, PP) * val(dist); } return;
        default:
            if ((PP.z >= 0.0) && (PP.z <= e.len)) {
                double dist = sqr(PP.x) + sqr(PP.y);
                if (dist <= e.rad2) { double vv = val(dist); PP.z = 0.0; Result = Result + MTransNormal(tr, PP) * vv; }
            }
    }
}

// the point walks of calculate_field_value / Normal (blob.cpp:1502-1594, 1815-1905): all components, or the leaves of the
// bounding-sphere tree whose spheres contain the point, in the reference's LIFO order
template <class F> void Tracer::blob_walk_point(const pvgpu_blob& bl, V3 P, F&& leaf) const
{
    const pvgpu_blob_element* el = S.blob_elements.data() + bl.element_first;
    if (bl.node_count == 0) { for (uint32_t i = 0; i < bl.element_count; i++) leaf(el[i]); return; }
    const pvgpu_blob_node* nodes = S.blob_nodes.data() + bl.node_first;
    std::vector<uint32_t> queue{ 0u };
    while (!queue.empty()) {
        const pvgpu_blob_node& nd = nodes[queue.back()]; queue.pop_back();
        if (nd.count == 0) leaf(el[nd.first]);
        else for (uint32_t i = 0; i < nd.count; i++) if (len2(P - v3(nodes[nd.first + i].c)) <= nodes[nd.first + i].r2) queue.push_back(nd.first + i);
    }
}

// Blob::All_Intersections (blob.cpp:239-610) with determine_influences (:1269-1377) and insert_hit (:616-714)
bool Tracer::blob_intersect(uint32_t idx, const Ray& ray, IStack& Depth_Stack) const
{
    const pvgpu_object& ob = S.objects[idx];
    const pvgpu_blob& bl = S.blobs[ob.mesh];
    const pvgpu_blob_element* el = S.blob_elements.data() + bl.element_first;
    const double depthTolerance = 1.0e-2;
    V3 P = ray.Origin, D = ray.Direction;
    double length = 1.0;
    if (ob.transform >= 0) { const pvgpu_transform& t = S.xf[ob.transform]; P = MInvTransPoint(t, ray.Origin); D = MInvTransDirection(t, ray.Direction); length = len(D); D = D / length; }
    struct Interval { int type; double bound; uint32_t elem; };
    std::vector<Interval> iv;
    auto insert_hit = [&](uint32_t e, double t0, double t1) {
        // sorted insertion; a bound goes in front of the first stored bound that is not smaller
        size_t k = 0;
        while (k < iv.size() && t0 > iv[k].bound) k++;
        const bool appended = (k == iv.size());
        iv.insert(iv.begin() + k, Interval{ (int)el[e].type | 0, t0, e });
        if (appended) { iv.push_back(Interval{ (int)el[e].type | 1, t1, e }); return; }
        k++;
        while (k < iv.size() && t1 > iv[k].bound) k++;
        iv.insert(iv.begin() + k, Interval{ (int)el[e].type | 1, t1, e });
    };
    if (bl.node_count == 0) {
        for (uint32_t i = 0; i < bl.element_count; i++) { double t0, t1; if (blob_element_hit(el[i], P, D, depthTolerance, &t0, &t1)) insert_hit(i, t0, t1); }
    } else {
        const pvgpu_blob_node* nodes = S.blob_nodes.data() + bl.node_first;
        std::vector<uint32_t> queue{ 0u };
        while (!queue.empty()) {
            const pvgpu_blob_node& nd = nodes[queue.back()]; queue.pop_back();
            if (nd.count == 0) { double t0, t1; if (blob_element_hit(el[nd.first], P, D, depthTolerance, &t0, &t1)) insert_hit(nd.first, t0, t1); }
            else for (uint32_t i = 0; i < nd.count; i++) {
                V3 V1 = v3(nodes[nd.first + i].c) - P;
                double b = dot(V1, D), t = len2(V1);
                if ((t - sqr(b)) <= nodes[nd.first + i].r2) queue.push_back(nd.first + i);
            }
        }
    }
    const int cnt = (int)iv.size();
    if (cnt == 0) return false;
    double start_dist = iv[0].bound;
    if (start_dist < SMALL_TOLERANCE) start_dist = 0.0;
    for (auto& x : iv) x.bound -= start_dist;
    P = P + D * start_dist;
    double max_bound = iv[0].bound;
    for (auto& x : iv) if (x.bound > max_bound) max_bound = x.bound;
    if (max_bound != 0) { D = D * max_bound; for (auto& x : iv) x.bound /= max_bound; } else max_bound = 1;
    double coeffs[5] = { 0.0, 0.0, 0.0, 0.0, -bl.threshold };
    std::vector<double> fcoeffs((size_t)bl.element_count * 5, 0.0);
    bool found = false;
    int in_flag = 0;
    for (int i = 0; i < cnt; i++) {
        double* f = &fcoeffs[(size_t)iv[i].elem * 5];
        if ((iv[i].type & 1) == 0) {
            in_flag++;
            const pvgpu_blob_element& e = el[iv[i].elem];
            double t0, t1, t2;
            if (e.type == PVGPU_BLOB_SPHERE) { V3 V1 = P - v3(e.o); t0 = len2(V1); t1 = dot(V1, D); t2 = max_bound * max_bound; }
            else {
                const pvgpu_transform& tr = S.xf[e.transform];
                V3 PP = MInvTransPoint(tr, P), DD = MInvTransDirection(tr, D);
                if (e.type == PVGPU_BLOB_ELLIPSOID) { V3 V1 = PP - v3(e.o); t0 = len2(V1); t1 = dot(V1, DD); t2 = len2(DD); }
                else if (e.type == PVGPU_BLOB_CYLINDER) { t0 = PP.x * PP.x + PP.y * PP.y; t1 = PP.x * DD.x + PP.y * DD.y; t2 = DD.x * DD.x + DD.y * DD.y; }
                else { if (e.type == PVGPU_BLOB_APEX_HEMISPHERE) PP.z -= e.len; t0 = len2(PP); t1 = dot(PP, DD); t2 = len2(DD); }
            }
            const double c0 = e.c[0], c1 = e.c[1], c2 = e.c[2];
            f[0] = c0 * t2 * t2;
            f[1] = 4.0 * c0 * t1 * t2;
            f[2] = 2.0 * c0 * (2.0 * t1 * t1 + t0 * t2) + c1 * t2;
            f[3] = 2.0 * t1 * (2.0 * c0 * t0 + c1);
            f[4] = t0 * (c0 * t0 + c1) + c2;
            for (int j = 0; j < 5; j++) coeffs[j] += f[j];
        } else {
            for (int j = 0; j < 5; j++) coeffs[j] -= f[j];
            if (--in_flag == 0) continue;
        }
        if ((i + 1 < cnt) && (std::fabs(iv[i].bound - iv[i + 1].bound) < EPSILON)) continue;
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
        bool all_pos = true, all_neg = true;
        for (int j = 0; j < 5; j++) { all_pos = all_pos && (dk[j] >= 0.0); all_neg = all_neg && (dk[j] <= 0.0); }
        if (all_pos || all_neg) continue;
        double roots[4];
        int root_count = Solve_Polynomial(4, coeffs, roots, (ob.flags & PVGPU_STURM_FLAG) ? 1 : 0, 1.0e-11);
        for (int j = 0; j < root_count; j++) {
            double dist = roots[j];
            if ((dist >= iv[i].bound) && (dist <= iv[i + 1].bound)) {
                dist = (dist * max_bound + start_dist) / length;
                if ((dist > depthTolerance) && (dist < MAX_DISTANCE)) {
                    V3 ip = ray.Evaluate(dist);
                    if (clip_ok(ob, ip)) { Intersection is; is.Depth = dist; is.IPoint = ip; is.Object = (int)idx; is.aux = 0; Depth_Stack.push_back(is); found = true; }
                }
            }
        }
        if (!(ob.aux & 1u) && found) break;       // not a CSG child
    }
    return found;
}

// <Primitive>::Normal
V3 Tracer::Normal(const Intersection& isect) const
{
    const pvgpu_object& ob = S.objects[isect.Object];
    switch (ob.type) {
        case PVGPU_OBJ_SPHERE:                                                                            // sphere.cpp:318-340
            if (ob.aux) { const pvgpu_transform& t = S.xf[ob.transform]; return unit(MTransNormal(t, MInvTransPoint(t, isect.IPoint))); }
            return (isect.IPoint - v3(ob.p)) / ob.p[3];
        case PVGPU_OBJ_BOX: {                                                                             // box.cpp:600-620
            V3 n = v3(0, 0, 0);
            switch (isect.aux) { case 1: n.x = -1; break; case 2: n.x = 1; break; case 3: n.y = -1; break; case 4: n.y = 1; break; case 5: n.z = -1; break; case 6: n.z = 1; break; }
            if (ob.transform >= 0) n = unit(MTransNormal(S.xf[ob.transform], n));
            return n;
        }
        case PVGPU_OBJ_PLANE: {                                                                           // plane.cpp:243-253
            V3 n = v3(ob.p);
            if (ob.transform >= 0) n = unit(MTransNormal(S.xf[ob.transform], n));
            return n;
        }
        case PVGPU_OBJ_QUADRIC: {                                                                         // quadric.cpp:273-310
            const double* c = ob.p; V3 ip = isect.IPoint;
            V3 n = v3(2.0 * c[0] * ip.x + c[3] * ip.y + c[4] * ip.z + c[6], c[3] * ip.x + 2.0 * c[1] * ip.y + c[5] * ip.z + c[7], c[4] * ip.x + c[5] * ip.y + 2.0 * c[2] * ip.z + c[8]);
            double l = len(n);
            return l == 0.0 ? v3(1, 0, 0) : n / l;
        }
        case PVGPU_OBJ_TORUS: {                                                                           // torus.cpp:418-520
            const pvgpu_transform& t = S.xf[ob.transform];
            V3 P = MInvTransPoint(t, isect.IPoint);
            double dist = std::sqrt(P.x * P.x + P.z * P.z);
            V3 M = v3(0, 0, 0);
            if (dist > EPSILON) { M.x = ob.p[0] * P.x / dist; M.z = ob.p[0] * P.z / dist; }
            V3 N = isect.aux ? P + M : P - M;
            return unit(MTransNormal(t, N));
        }
        case PVGPU_OBJ_GLYPH: return isect.INormal;                                                       // truetype.cpp:2972-2976
        case PVGPU_OBJ_SUPERELLIPSOID: {                                                                  // superellipsoid.cpp:451-495
            const pvgpu_transform& t = S.xf[ob.transform];
            const double* E = ob.p;
            V3 P = MInvTransPoint(t, isect.IPoint);
            double r = 0.0, z2n = 0;
            if (P.z != 0) { z2n = superq::power(std::fabs(P.z), E[2]); P.z = z2n / P.z; }
            if (std::fabs(P.x) > std::fabs(P.y)) {
                r = superq::power(std::fabs(P.y / P.x), E[0]);
                P.x = (1 - z2n) / P.x;
                P.y = P.y ? (1 - z2n) * r / P.y : 0;
            } else if (P.y != 0) {
                r = superq::power(std::fabs(P.x / P.y), E[0]);
                P.x = P.x ? (1 - z2n) * r / P.x : 0;
                P.y = (1 - z2n) / P.y;
            }
            if (P.z) P.z *= (1 + r);
            return unit(MTransNormal(t, P));
        }
        case PVGPU_OBJ_PRISM: {                                                                           // prism.cpp:710-770
            const pvgpu_transform& t = S.xf[ob.transform];
            V3 N = v3(0, 0, 0);
            const uint32_t side = isect.aux & 3u;
            if (side == 0u) N = v3(0.0, -1.0, 0.0);
            else if (side == 1u) N = v3(0.0, 1.0, 0.0);
            else {
                const PrismEntry& E = reinterpret_cast<const PrismEntry*>(S.shape_data.data() + ob.mesh + 1)[isect.aux >> 4];
                const double d1 = isect.d1;
                if (((ob.aux >> 4) & 15u) == 1u) {
                    N.x = d1 * (3.0 * E.A[1] * d1 + 2.0 * E.B[1]) + E.C[1];
                    N.y = 0.0;
                    N.z = -(d1 * (3.0 * E.A[0] * d1 + 2.0 * E.B[0]) + E.C[0]);
                } else {
                    V3 P = MInvTransPoint(t, isect.IPoint);
                    if (std::fabs(P.y) > EPSILON) {
                        N.x = d1 * (3.0 * E.A[1] * d1 + 2.0 * E.B[1]) + E.C[1];
                        N.z = -(d1 * (3.0 * E.A[0] * d1 + 2.0 * E.B[0]) + E.C[0]);
                        N.y = -(P.x * N.x + P.z * N.z) / P.y;
                    }
                }
            }
            return unit(MTransNormal(t, N));
        }
        case PVGPU_OBJ_DISC: return v3(ob.p);                                                             // disc.cpp:226-229
        case PVGPU_OBJ_POLYGON: return v3(ob.p);                                                          // polygon.cpp:308-311
        case PVGPU_OBJ_POLY: {                                                                            // polynomial.cpp:1035-1129, 1180-1244
            const pvgpu_transform& t = S.xf[ob.transform];
            const double* a = S.shape_data.data() + ob.mesh;
            const int order = (int)ob.aux;
            const V3 P = MInvTransPoint(t, isect.IPoint);
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
            V3 r = MTransNormal(t, v3(rx, ry, rz));
            double val = len2(r);
            if (val > 0.0) { val = 1.0 / std::sqrt(val); return r * val; }
            return v3(1.0, 0.0, 0.0);
        }
        case PVGPU_OBJ_TRIANGLE: {                                                                        // triangle.cpp:640-700
            const double* T = S.shape_data.data() + ob.mesh;
            if (!(ob.aux & PVGPU_TRIANGLE_SMOOTH)) return v3(T + 9);
            V3 PIMinusP1 = isect.IPoint - v3(T);
            double u = dot(PIMinusP1, v3(T + 22));
            if (u < EPSILON) return v3(T + 13);
            int Axis = (ob.aux >> 2) & 3;
            double v = (PIMinusP1[Axis] / u + T[Axis] - T[3 + Axis]) / (T[6 + Axis] - T[3 + Axis]);
            return unit(v3(T + 13) + u * (v3(T + 16) - v3(T + 13) + v * (v3(T + 19) - v3(T + 16))));
        }
        case PVGPU_OBJ_CONE: {                                                                            // cone.cpp:408-445
            const pvgpu_transform& t = S.xf[ob.transform];
            V3 r = MInvTransPoint(t, isect.IPoint);
            if (isect.aux == 3) { if (ob.flags & PVGPU_CYLINDER_FLAG) r.z = 0.0; else r.z = -r.z; }
            else if (isect.aux == 1) r = v3(0, 0, -1);
            else if (isect.aux == 2) r = v3(0, 0, 1);
            return unit(MTransNormal(t, r));
        }
        case PVGPU_OBJ_BLOB: {                                                                            // blob.cpp:1815-1933
            const pvgpu_blob& bl = S.blobs[ob.mesh];
            V3 P = (ob.transform >= 0) ? MInvTransPoint(S.xf[ob.transform], isect.IPoint) : isect.IPoint;
            V3 Result = v3(0, 0, 0);
            blob_walk_point(bl, P, [&](const pvgpu_blob_element& e) { blob_element_normal(e, P, Result); });
            double val = len2(Result);
            if (val == 0.0) Result = v3(1, 0, 0);
            else { val = 1.0 / std::sqrt(val); Result = Result * val; }
            if (ob.transform >= 0) Result = unit(MTransNormal(S.xf[ob.transform], Result));
            return Result;
        }
        case PVGPU_OBJ_MESH: {                                                                            // mesh.cpp:283-375
            const pvgpu_mesh& me = S.meshes[ob.mesh];
            const pvgpu_triangle& tr = S.tris[isect.aux];
            const float* N = S.norms.data() + 3 * (size_t)me.normal_first;
            const float* V = S.verts.data() + 3 * (size_t)me.vertex_first;
            V3 result;
            if (tr.flags & PVGPU_TRI_SMOOTH) {
                V3 ip = (ob.transform >= 0) ? MInvTransPoint(S.xf[ob.transform], isect.IPoint) : isect.IPoint;
                V3 N1 = v3f(N + 3 * tr.n1), N2 = v3f(N + 3 * tr.n2), N3 = v3f(N + 3 * tr.n3);
                V3 PIMinusP1 = ip - v3f(V + 3 * tr.p1);
                double u = dot(PIMinusP1, v3f(tr.perp));
                if (u < EPSILON) result = N1;
                else {
                    int axis = tr.v_axis;
                    double k1 = V[3 * tr.p1 + axis], k2 = V[3 * tr.p2 + axis], k3 = V[3 * tr.p3 + axis];
                    double v = (PIMinusP1[axis] / u + k1 - k2) / (k3 - k2);
                    result = N1 + u * (N2 - N1 + v * (N3 - N2));
                }
                if (ob.transform >= 0) result = MTransNormal(S.xf[ob.transform], result);
                return unit(result);
            }
            result = v3f(N + 3 * tr.normal_ind);
            if (ob.transform >= 0) result = unit(MTransNormal(S.xf[ob.transform], result));
            return result;
        }
    }
    return v3(0, 1, 0);
}

// ---- materials -------------------------------------------------------------------------------------------
static double cycloidal(double value)                                                                     // texture.cpp:98-110
{
    const double TWO_M_PI = 6.283185307179586476925286766560;
    if (value >= 0.0) return std::sin(((value - std::floor(value)) * 50000.0) / 50000.0 * TWO_M_PI);
    return 0.0 - std::sin(((0.0 - (value + std::floor(0.0 - value))) * 50000.0) / 50000.0 * TWO_M_PI);
}
static double Triangle_Wave(double value)                                                                 // texture.cpp:128-150
{
    double offset = (value >= 0.0) ? value - std::floor(value) : value + 1.0 + std::floor(std::fabs(value));
    return (offset >= 0.5) ? 2.0 * (1.0 - offset) : 2.0 * offset;
}

static double FLOOR(double x) { return x >= 0.0 ? std::floor(x) : (0.0 - std::floor(0.0 - x) - 1.0); }    // texture.h:73

// BlackHoleWarp / RepeatWarp / CubicWarp / CylindricalWarp / SphericalWarp / ToroidalWarp / PlanarWarp::WarpPoint (warp.cpp:124-545);
// q = the warp's parameters in the shape-data table (pvgpu.h, PVGPU_WARP_*)
static bool WarpPoint_other(uint32_t type, const double* q, V3& TPoint)
{
    const double M_PI_ = 3.1415926535897932384626, TWO_M_PI = 6.283185307179586476925286766560;
    auto orientation = [&](const double* Orientation_Vector, double x, double y, double z) {
        if ((Orientation_Vector[0] == 0.0) && (Orientation_Vector[1] == 0.0) && (Orientation_Vector[2] == 1.0)) { TPoint = v3(x, y, z); return; }
        TPoint = v3((Orientation_Vector[0] * z) + (Orientation_Vector[1] * x) + (Orientation_Vector[2] * x),
                    (Orientation_Vector[0] * y) + (Orientation_Vector[1] * -z) + (Orientation_Vector[2] * y),
                    (Orientation_Vector[0] * -x) + (Orientation_Vector[1] * y) + (Orientation_Vector[2] * z));
    };
    double x = TPoint.x, y = TPoint.y, z = TPoint.z, len, theta, phi;
    switch (type) {
        case PVGPU_WARP_BLACK_HOLE: {                                                                     // :124-205
            V3 C = v3(q[0], q[1], q[2]);
            const uint32_t flags = (uint32_t)q[9];
            if (flags & 2u) {
                int blockX = 0, blockY = 0, blockZ = 0;
                if (q[3] >= EPSILON) blockX = (int)std::floor(TPoint.x / q[3]);
                if (q[4] >= EPSILON) blockY = (int)std::floor(TPoint.y / q[4]);
                if (q[5] >= EPSILON) blockZ = (int)std::floor(TPoint.z / q[5]);
                C.x += q[3] * blockX; C.y += q[4] * blockY; C.z += q[5] * blockZ;
            }
            V3 Delta = TPoint - C;
            double Length = ::len(Delta);
            if (Length >= q[7]) return true;
            if ((int)q[10] == 0) {
                Length = (q[7] - Length) / q[7];
                double Sv = std::pow(Length, q[8]) * q[6];
                if (Sv > 1.0) Sv = 1.0;
                Delta = Delta * ((flags & 1u) ? -Sv : Sv);
                TPoint = TPoint + Delta;
            }
            return true;
        }
        case PVGPU_WARP_REPEAT: {                                                                         // :354-366
            const int Axis = (int)q[0];
            const float Width = (float)q[1];
            double* T[3] = { &TPoint.x, &TPoint.y, &TPoint.z };
            float BlkNum = (float)std::floor(*T[Axis] / Width);
            *T[Axis] -= BlkNum * Width;
            if (((int)BlkNum) & 1) {
                TPoint = v3(TPoint.x * q[2], TPoint.y * q[3], TPoint.z * q[4]);
                if (q[2 + Axis] < 0) *T[Axis] += Width;
            }
            TPoint = TPoint + v3(q[5], q[6], q[7]) * (double)BlkNum;
            return true;
        }
        case PVGPU_WARP_CUBIC: {                                                                          // :207-253
            const double ax = std::fabs(x), ay = std::fabs(y), az = std::fabs(z);
            if (x >= 0 && x >= ay && x >= az) TPoint = v3(0.75 - 0.25 * (z / x + 1.0) / 2.0, 1.0 / 3.0 + (1.0 / 3.0) * (y / x + 1.0) / 2.0, x);
            else if (y >= 0 && y >= ax && y >= az) TPoint = v3(0.25 + 0.25 * (x / y + 1.0) / 2.0, 1.0 - (1.0 / 3.0) * (z / y + 1.0) / 2.0, y);
            else if (z >= 0 && z >= ax && z >= ay) TPoint = v3(0.25 + 0.25 * (x / z + 1.0) / 2.0, 1.0 / 3.0 + (1.0 / 3.0) * (y / z + 1.0) / 2.0, z);
            else if (x < 0 && x <= -ay && x <= -az) { x = -x; TPoint = v3(0.25 * (z / x + 1.0) / 2.0, 1.0 / 3.0 + (1.0 / 3.0) * (y / x + 1.0) / 2.0, x); }
            else if (y < 0 && y <= -ax && y <= -az) { y = -y; TPoint = v3(0.25 + 0.25 * (x / y + 1.0) / 2.0, (1.0 / 3.0) * (z / y + 1.0) / 2.0, y); }
            else { z = -z; TPoint = v3(1.0 - 0.25 * (x / z + 1.0) / 2.0, 1.0 / 3.0 + (1.0 / 3.0) * (y / z + 1.0) / 2.0, z); }
            return true;
        }
        case PVGPU_WARP_CYLINDRICAL:                                                                      // :255-310
            len = std::sqrt(x * x + z * z);
            if (len == 0.0) return false;
            if (z == 0.0) { if (x > 0) theta = 0.0; else theta = M_PI_; }
            else { theta = std::acos(x / len); if (z < 0.0) theta = TWO_M_PI - theta; }
            theta /= TWO_M_PI;
            if (q[3] == 1.0) theta *= len; else if (q[3] != 0.0) theta *= std::pow(len, q[3]);
            orientation(q, theta, y, len);
            return true;
        case PVGPU_WARP_SPHERICAL: {                                                                      // :368-452
            const double dist = std::sqrt(x * x + y * y + z * z);
            if (dist == 0.0) return false;
            x /= dist; y /= dist; z /= dist;
            phi = 0.5 + std::asin(y) / M_PI_;
            len = std::sqrt(x * x + z * z);
            if (len == 0.0) theta = 0;
            else {
                if (z == 0.0) { if (x > 0) theta = 0.0; else theta = M_PI_; }
                else { theta = std::acos(x / len); if (z < 0.0) theta = TWO_M_PI - theta; }
                theta /= TWO_M_PI;
            }
            if (q[3] == 1.0) { theta *= dist; phi *= dist; }
            else if (q[3] != 0.0) { theta *= std::pow(dist, q[3]); phi *= std::pow(dist, q[3]); }
            orientation(q, theta, phi, dist);
            return true;
        }
        case PVGPU_WARP_TOROIDAL:                                                                         // :454-545
            len = std::sqrt(x * x + z * z);
            if (len == 0.0) return false;
            if (z == 0.0) { if (x > 0) theta = 0.0; else theta = M_PI_; }
            else { theta = std::acos(x / len); if (z < 0.0) theta = TWO_M_PI - theta; }
            theta = 0.0 - theta;
            x = len - q[4];
            len = std::sqrt(x * x + y * y);
            phi = std::acos(-x / len);
            if (y > 0.0) phi = TWO_M_PI - phi;
            theta /= (-TWO_M_PI);
            phi /= TWO_M_PI;
            if (q[3] == 1.0) { theta *= len; phi *= len; }
            else if (q[3] != 0.0) { theta *= std::pow(len, q[3]); phi *= std::pow(len, q[3]); }
            orientation(q, theta, phi, len);
            return true;
        case PVGPU_WARP_PLANAR:                                                                           // :318-352
            orientation(q, x, y, q[3]);
            return true;
    }
    return false;
}

V3 Tracer::Warp_EPoint(const pvgpu_pigment& pg, V3 EPoint) const                                          // warp.cpp:103-122
{
    V3 p = EPoint;
    for (int i = (int)pg.warp_count - 1; i >= 0; i--) {
        const pvgpu_warp& w = S.warps[pg.warp_first + i];
        if (w.type == PVGPU_WARP_TRANSFORM) p = MInvTransPoint(S.xf[w.transform], p);
        else if (w.type > PVGPU_WARP_CLASSIC_TURBULENCE) WarpPoint_other(w.type, S.shape_data.data() + w.transform, p);
        else { V3 t = DTurbulence(S, p, w); p = v3(p.x + t.x * w.turbulence[0], p.y + t.y * w.turbulence[1], p.z + t.z * w.turbulence[2]); }
    }
    auto clampc = [](double& c) { if (c > COORDINATE_LIMIT) c = COORDINATE_LIMIT; else if (c < -COORDINATE_LIMIT) c = -COORDINATE_LIMIT; };
    clampc(p.x); clampc(p.y); clampc(p.z);
    return p;
}

// FractalPattern family: JuliaPattern .. Julia4Pattern (pattern.cpp:6895-7098), Magnet1M .. Magnet2J (7228-7550), Mandel2 .. Mandel4
// (7551-7751), ExteriorColour / InteriorColour (8990-9056).  rec = kind, maxIterations, exteriorType, interiorType, exteriorFactor,
// interiorFactor, juliaCoord.  Every kind is the same loop around its own iteration step.
static double fractal_pattern(const double* rec, V3 EPoint)
{
    const int kind = (int)rec[0], it_max = (int)rec[1], exteriorType = (int)rec[2], interiorType = (int)rec[3];
    const double exteriorFactor = rec[4], interiorFactor = rec[5];
    double x = EPoint.x, y = EPoint.y;                 // the constant c of the iteration: the point (M kinds) or juliaCoord (J kinds)
    double a, b, a2, b2, mindist2;
    bool magnet = false;
    switch (kind) {
        case PVGPU_FRACTAL_JULIA2: case PVGPU_FRACTAL_JULIA3: case PVGPU_FRACTAL_JULIA4:
            a = EPoint.x; b = EPoint.y; x = rec[6]; y = rec[7]; a2 = sqr(a); b2 = sqr(b); mindist2 = a2 + b2; break;
        case PVGPU_FRACTAL_MAGNET1J: case PVGPU_FRACTAL_MAGNET2J:
            magnet = true; a = EPoint.x; b = EPoint.y; x = rec[6]; y = rec[7]; a2 = sqr(a); b2 = sqr(b); mindist2 = a2 + b2; break;
        case PVGPU_FRACTAL_MAGNET1M: case PVGPU_FRACTAL_MAGNET2M:
            magnet = true; a = a2 = 0; b = b2 = 0; mindist2 = 10000; break;
        default:
            a = x; b = y; a2 = sqr(a); b2 = sqr(b); mindist2 = a2 + b2; break;
    }
    const double c1r = x - 1, c2r = x - 2, c1c2r = c1r * c2r - y * y, c1c2i = (c1r + c2r) * y;
    int col;
    double cf = 0.0;
    for (col = 0; col < it_max; col++) {
        double tmp, tmp1r, tmp1i, tmp2r, tmp2i;
        switch (kind) {
            case PVGPU_FRACTAL_MANDEL2: case PVGPU_FRACTAL_JULIA2: b = 2.0 * a * b + y; a = a2 - b2 + x; break;
            case PVGPU_FRACTAL_MANDEL3: case PVGPU_FRACTAL_JULIA3: b = 3.0 * a2 * b - b2 * b + y; a = a2 * a - 3.0 * a * b2 + x; break;
            case PVGPU_FRACTAL_MANDEL4: case PVGPU_FRACTAL_JULIA4: b = 4.0 * (a2 * a * b - a * b2 * b) + y; a = a2 * a2 - 6.0 * a2 * b2 + b2 * b2 + x; break;
            case PVGPU_FRACTAL_MAGNET1M: case PVGPU_FRACTAL_MAGNET1J:
                tmp1r = a2 - b2 + x - 1; tmp1i = 2 * a * b + y; tmp2r = 2 * a + x - 2; tmp2i = 2 * b + y;
                tmp = tmp2r * tmp2r + tmp2i * tmp2i;
                a = (tmp1r * tmp2r + tmp1i * tmp2i) / tmp; b = (tmp1i * tmp2r - tmp1r * tmp2i) / tmp;
                b2 = b * b; b = 2 * a * b; a = a * a - b2;
                break;
            default:
                tmp1r = a2 * a - 3 * a * b2 + 3 * (a * c1r - b * y) + c1c2r; tmp1i = 3 * a2 * b - b2 * b + 3 * (a * y + b * c1r) + c1c2i;
                tmp2r = 3 * (a2 - b2) + 3 * (a * c2r - b * y) + c1c2r + 1; tmp2i = 6 * a * b + 3 * (a * y + b * c2r) + c1c2i;
                tmp = tmp2r * tmp2r + tmp2i * tmp2i;
                a = (tmp1r * tmp2r + tmp1i * tmp2i) / tmp; b = (tmp1i * tmp2r - tmp1r * tmp2i) / tmp;
                b2 = b * b; b = 2 * a * b; a = a * a - b2;
                break;
        }
        a2 = sqr(a); b2 = sqr(b);
        const double dist2 = a2 + b2;
        if (dist2 < mindist2) mindist2 = dist2;
        const bool escaped = magnet ? (dist2 > 10000.0 || (a - 1) * (a - 1) + b2 < 1 / 10000.0) : (dist2 > 4.0);
        if (escaped) {
            switch (exteriorType) {                                                                           // pattern.cpp:8990-9015
                case 0: cf = exteriorFactor; break;
                case 1: cf = (double)col / (double)it_max; break;
                case 2: cf = a * exteriorFactor; break;
                case 3: cf = b * exteriorFactor; break;
                case 4: cf = a * a * exteriorFactor; break;
                case 5: cf = b * b * exteriorFactor; break;
                case 6: cf = std::sqrt(a * a + b * b) * exteriorFactor; break;
                case 7: cf = (double)(col % (unsigned int)exteriorFactor) / exteriorFactor; break;
                case 8: cf = (double)(col % (unsigned int)(1 + exteriorFactor)) / exteriorFactor; break;
            }
            break;
        }
    }
    if (col == it_max)
        switch (interiorType) {                                                                               // pattern.cpp:9036-9056
            case 0: cf = interiorFactor; break;
            case 1: cf = std::sqrt(mindist2) * interiorFactor; break;
            case 2: cf = a * interiorFactor; break;
            case 3: cf = b * interiorFactor; break;
            case 4: cf = a * a * interiorFactor; break;
            case 5: cf = b * b * interiorFactor; break;
            case 6: cf = a * a + b * b * interiorFactor; break;
        }
    return cf;
}

static double crackle_pattern(const Scene& S, const pvgpu_pigment& pg, V3 ep, int gen)                      // pattern.cpp:5760-5987 (no cell cache)
{
    const double* cp = S.shape_data.data() + pg.data;
    const double form_x = cp[0], form_y = cp[1], form_z = cp[2], metric = cp[3], offset = cp[4];
    const bool is_solid = cp[5] != 0.0;
    const int rep[3] = { (int)cp[6], (int)cp[7], (int)cp[8] };
    const bool use_square = (metric == 2), use_unity = (metric == 1);
    auto wrap = [](double val, double upper) {                       // wrap() mathutil.h:102-121
        double t = std::fmod(val, upper);
        if (t < 0.0) t += upper;
        if (t >= upper) t = 0.0;
        return t;
    };
    V3 tp = ep;
    if (rep[0]) tp.x = wrap(tp.x, (double)rep[0]);
    if (rep[1]) tp.y = wrap(tp.y, (double)rep[1]);
    if (rep[2]) tp.z = wrap(tp.z, (double)rep[2]);
    const int flo[3] = { (int)std::floor(tp.x - EPSILON), (int)std::floor(tp.y - EPSILON), (int)std::floor(tp.z - EPSILON) };
    // nucleus of cube `index` of gaCrackleCubeTable (pattern.cpp:9349-9374) around the point: IntPickInCube (pattern.cpp:8808-8819)
    auto nucleus = [&](int ax, int ay, int az) {
        int c[3] = { flo[0] + ax, flo[1] + ay, flo[2] + az };
        double woff[3] = { 0.0, 0.0, 0.0 };
        for (int k = 0; k < 3; k++)
            if (rep[k]) { int w = c[k] % rep[k]; if (w < 0) w += rep[k]; woff[k] += (c[k] - w); c[k] = w; }     // wrapInt
        const unsigned seed = S.hashTable[S.hashTable[S.hashTable[c[0] & 0xfff] ^ (c[1] & 0xfff)] ^ (c[2] & 0xfff)];                  // Hash3d texture.h:75
        double nx = c[0] + S.patternRands[seed % 32768u], ny = c[1] + S.patternRands[(seed + 1u) % 32768u], nz = c[2] + S.patternRands[(seed + 2u) % 32768u];
        nx += woff[0]; ny += woff[1]; nz += woff[2];
        return v3(nx, ny, nz);
    };
    auto dist = [&](const V3& n) {
        const double dx = n.x - tp.x, dy = n.y - tp.y, dz = n.z - tp.z;
        if (use_square) return dx * dx + dy * dy + dz * dz;
        if (use_unity) return std::fabs(dx) + std::fabs(dy) + std::fabs(dz);
        return std::pow(std::fabs(dx), metric) + std::pow(std::fabs(dy), metric) + std::pow(std::fabs(dz), metric);
    };
    double minsum = 0.0, minsum2 = 0.0, minsum3 = 0.0, tf;
    int min_idx = 0, i = 0;
    for (int ax = -2; ax <= 2; ax++)
        for (int ay = -2; ay <= 2; ay++)
            for (int az = -2; az <= 2; az++) {
                if ((std::abs(ax) == 2) + (std::abs(ay) == 2) + (std::abs(az) == 2) > 1) continue;
                const double sum = dist(nucleus(ax, ay, az));
                if (i == 0) minsum = sum;
                else if (i == 1) minsum2 = sum;
                else if (i == 2) {
                    minsum3 = sum;
                    if (minsum2 < minsum) { tf = minsum; minsum = minsum2; minsum2 = tf; min_idx = 1; }
                    if (minsum3 < minsum) { tf = minsum; minsum = minsum3; minsum3 = tf; min_idx = 2; }
                    if (minsum3 < minsum2) { tf = minsum2; minsum2 = minsum3; minsum3 = tf; }
                } else {
                    if (sum < minsum) { minsum3 = minsum2; minsum2 = minsum; minsum = sum; min_idx = i; }
                    else if (sum < minsum2) { minsum3 = minsum2; minsum2 = sum; }
                    else if (sum < minsum3) { minsum3 = sum; }
                }
                i++;
            }
    if (offset != 0.0) {
        if (use_square) { minsum += offset * offset; minsum2 += offset * offset; minsum3 += offset * offset; }
        else if (use_unity) { minsum += offset; minsum2 += offset; minsum3 += offset; }
        else { minsum += std::pow(offset, metric); minsum2 += std::pow(offset, metric); minsum3 += std::pow(offset, metric); }
    }
    if (is_solid) {
        V3 minvec = v3(0.0, 0.0, 0.0);
        i = 0;
        for (int ax = -2; ax <= 2; ax++)
            for (int ay = -2; ay <= 2; ay++)
                for (int az = -2; az <= 2; az++) {
                    if ((std::abs(ax) == 2) + (std::abs(ay) == 2) + (std::abs(az) == 2) > 1) continue;
                    if (i == min_idx) minvec = nucleus(ax, ay, az);
                    i++;
                }
        tf = Noise(S, minvec, gen);
    }
    else if (use_square) tf = form_x * std::sqrt(minsum) + form_y * std::sqrt(minsum2) + form_z * std::sqrt(minsum3);
    else if (use_unity) tf = form_x * minsum + form_y * minsum2 + form_z * minsum3;
    else tf = form_x * std::pow(minsum, 1.0 / metric) + form_y * std::pow(minsum2, 1.0 / metric) + form_z * std::pow(minsum3, 1.0 / metric);
    return std::max(std::min(tf, 1.), 0.);
}

double Tracer::Evaluate_TPat(const pvgpu_pigment& pg, V3 p) const                                         // pattern.cpp:354-392 + EvaluateRaw
{
    const int gen = pg.noise_generator ? pg.noise_generator : S.g.noise_generator;
    const pvgpu_warp* turb = (pg.warp_count && S.warps[pg.warp_first].type == PVGPU_WARP_CLASSIC_TURBULENCE) ? &S.warps[pg.warp_first] : nullptr;
    double value = 0.0;
    bool discrete = false;
    switch (pg.pattern) {
        case PVGPU_PAT_CHECKER: {                                                                         // pattern.cpp:5691-5707
            int v = (int)(std::floor(p.x + EPSILON) + std::floor(p.y + EPSILON) + std::floor(p.z + EPSILON));
            value = (v & 1) ? 1.0 : 0.0; discrete = true; break;
        }
        case PVGPU_PAT_BOZO: case PVGPU_PAT_SPOTTED: value = Noise(S, p, gen); break;                     // pattern.cpp:7858
        case PVGPU_PAT_GRANITE: {                                                                         // pattern.cpp:6429-6465
            double noise = 0.0, freq = 1.0; V3 tv1 = p * 4.0;
            for (int i = 0; i < 6; freq *= 2.0, i++) {
                V3 tv2 = tv1 * freq; double temp;
                if (gen <= 1) temp = std::fabs(0.5 - Noise(S, tv2, gen));
                else { temp = std::fabs(1.0 - 2.0 * Noise(S, tv2, gen)); if (temp > 0.5) temp = 0.5; }
                noise += temp / freq;
            }
            value = noise; break;
        }
        case PVGPU_PAT_GRADIENT: { double r = dot(p, v3(pg.p)); value = (r > 1.0) ? std::fmod(r, 1.0) : r; break; }   // pattern.cpp:6386
        case PVGPU_PAT_MARBLE: value = p.x + (turb ? turb->turbulence[0] * Turbulence(S, p, *turb, gen) : 0.0); break; // pattern.cpp:7831
        case PVGPU_PAT_ONION: value = std::fmod(len(p), 1.0); break;                                      // pattern.cpp:7934
        case PVGPU_PAT_WRINKLES: {                                                                        // pattern.cpp:8720-8775
            double lambda = 2.0, omega = 0.5;
            auto n1 = [&](V3 q) { return gen <= 1 ? Noise(S, q, gen) : std::min(std::max(Noise(S, q, gen) * 2.0 - 0.5, 0.0), 1.0); };
            value = n1(p);
            for (int i = 1; i < 10; i++) { value += omega * n1(p * lambda); lambda *= 2.0; omega *= 0.5; }
            value = value / 2.0; break;
        }
        case PVGPU_PAT_AGATE: {                                                                           // pattern.cpp:5396-5421
            double tv = turb ? pg.p[0] * Turbulence(S, p, *turb, gen) : 0.0;
            double noise = 0.5 * (cycloidal(1.3 * tv + 1.1 * p.z) + 1.0);
            if (noise < 0.0) noise = 0.0; else { noise = std::min(1.0, noise); noise = std::pow(noise, 0.77); }
            value = noise; break;
        }
        case PVGPU_PAT_BRICK: {     // BrickPattern::Evaluate (pattern.cpp:5495-5608): discrete
            const double mortar = pg.p[3], fudgit = EPSILON + mortar;
            const double x = p.x + fudgit, y = p.y + fudgit, z = p.z + fudgit;
            const double bw = pg.p[0], bh = pg.p[1], bd = pg.p[2];
            const double mw = mortar / bw, mh = mortar / bh, md = mortar / bd;
            double by = y / bh; by -= (double)(int)by; if (by < 0.0) by += 1.0;
            if (by <= mh) { value = 0.0; discrete = true; break; }
            by = (y / bh) * 0.5; by -= (double)(int)by; if (by < 0.0) by += 1.0;
            double bx = x / bw; bx -= (double)(int)bx; if (bx < 0.0) bx += 1.0;
            if ((bx <= mw) && (by <= 0.5)) { value = 0.0; discrete = true; break; }
            bx = (x / bw) + 0.5; bx -= (double)(int)bx; if (bx < 0.0) bx += 1.0;
            if ((bx <= mw) && (by > 0.5)) { value = 0.0; discrete = true; break; }
            double bz = z / bd; bz -= (double)(int)bz; if (bz < 0.0) bz += 1.0;
            if ((bz <= md) && (by > 0.5)) { value = 0.0; discrete = true; break; }
            bz = (z / bd) + 0.5; bz -= (double)(int)bz; if (bz < 0.0) bz += 1.0;
            if ((bz <= md) && (by <= 0.5)) { value = 0.0; discrete = true; break; }
            value = 1.0; discrete = true; break;
        }
        case PVGPU_PAT_HEXAGON: {   // HexagonPattern::Evaluate (pattern.cpp:6512-6655): discrete 0 / 1 / 2
            double x = fabs(p.x), z = (p.z < 0.0) ? 5.196152424 - fabs(p.z) : p.z;
            double xs = x / 0.5, zs = z / 0.866025404;
            xs -= floor(xs / 6.0) * 6.0;
            zs -= floor(zs / 6.0) * 6.0;
            const int xm = (int)FLOOR(xs) % 6, zm = (int)FLOOR(zs) % 6;
            int v = 0;
            if (xm == 0 || xm == 5) v = (zm == 0 || zm == 5) ? 0 : ((zm == 1 || zm == 2) ? 1 : 2);
            else if (xm == 2 || xm == 3) v = (zm == 0 || zm == 1) ? 2 : ((zm == 2 || zm == 3) ? 0 : 1);
            else {
                double xl = xs - xm, zl = zs - zm;
                if (((xm + zm) % 2) == 1) xl = 1.0 - xl;
                if (xl == 0.0) xl = 0.0001;
                const bool brk = (zl / xl) < 1.0;
                const int zc = zm % 3;                 // (0,3) (1,4) (2,5)
                if (brk) v = (zc == 0) ? 0 : ((zc == 2) ? 1 : 2);
                else     v = (zc == 0) ? 2 : ((zc == 2) ? 0 : 1);
            }
            value = std::fmod((double)v, 3.0); discrete = true; break;
        }
        case PVGPU_PAT_WOOD: {      // WoodPattern::EvaluateRaw (pattern.cpp:8651-8683)
            double px = 0.0, py = 0.0;
            if (turb) {
                const V3 wt = DTurbulence(S, p, *turb);
                px = cycloidal((p.x + wt.x) * turb->turbulence[0]);
                py = cycloidal((p.y + wt.y) * turb->turbulence[1]);
            }
            px += p.x; py += p.y;
            value = len(v3(px, py, 0.0));
            break;
        }
        case PVGPU_PAT_LEOPARD:     // LeopardPattern::EvaluateRaw (pattern.cpp:7179-7193)
            value = sqr((sin(p.x) + sin(p.y) + sin(p.z)) / 3.0);
            break;
        case PVGPU_PAT_SPHERICAL:   // SphericalPattern / BoxedPattern / CylindricalPattern / PlanarPattern + CLIP_DENSITY (pattern.cpp:85)
        case PVGPU_PAT_BOXED:
        case PVGPU_PAT_CYLINDRICAL:
        case PVGPU_PAT_PLANAR:
            if (pg.pattern == PVGPU_PAT_SPHERICAL) value = len(p);
            else if (pg.pattern == PVGPU_PAT_BOXED) value = fmax(fabs(p.x), fmax(fabs(p.y), fabs(p.z)));
            else if (pg.pattern == PVGPU_PAT_CYLINDRICAL) value = sqrt(sqr(p.x) + sqr(p.z));
            else value = fabs(p.y);
            if (value < 0.0) value = 1.0; else if (value > 1.0) value = 0.0; else value = 1.0 - value;
            break;
        case PVGPU_PAT_RADIAL:      // RadialPattern::EvaluateRaw (pattern.cpp:8115-8129)
            if ((fabs(p.x) < 0.001) && (fabs(p.z) < 0.001)) value = 0.25;
            else value = 0.25 + (atan2(p.x, p.z) + 3.1415926535897932384626) / 6.283185307179586476925286766560;
            break;
        case PVGPU_PAT_DENTS: {     // DentsPattern::EvaluateRaw (pattern.cpp:6307-6313)
            const double n = Noise(S, p, gen);
            value = n * n * n;
            break;
        }
        case PVGPU_PAT_RIPPLES:     // RipplesPattern::EvaluateRaw (pattern.cpp:8163-8186)
        case PVGPU_PAT_WAVES: {     // WavesPattern::EvaluateRaw (pattern.cpp:8593-8618)
            const uint32_t nw = S.g.number_of_waves;
            double scalar = 0.0;
            for (uint32_t i = 0; i < nw; i++) {
                double length = len(p - S.waveSources[i]);
                if (length == 0.0) length = 1.0;
                if (pg.pattern == PVGPU_PAT_RIPPLES) scalar += cycloidal(length * (double)pg.frequency + (double)pg.phase);
                else { const double f = S.waveFrequencies[i]; scalar += cycloidal(length * (double)pg.frequency * f + (double)pg.phase) / f; }
            }
            value = (pg.pattern == PVGPU_PAT_RIPPLES) ? 0.5 * (1.0 + (scalar / (double)nw)) : 0.2 * (2.5 + (scalar / (double)nw));
            break;
        }
        case PVGPU_PAT_QUILTED: {   // QuiltedPattern::EvaluateRaw (pattern.cpp:8067-8083)
            V3 v = v3(p.x - FLOOR(p.x) - 0.5, p.y - FLOOR(p.y) - 0.5, p.z - FLOOR(p.z) - 0.5);
            double t = len(v);
            const double it = 1 - t, itsqrd = it * it, tsqrd = t * t, tcubed = t * tsqrd;
            t = (tcubed + 3.0 * t * itsqrd * pg.p[0] + 3.0 * tsqrd * it * pg.p[1]) * 1.154700538;
            v = v * t;
            value = (fabs(v.x) + fabs(v.y) + fabs(v.z)) / 3.0;
            break;
        }
    }
    if (pg.pattern == PVGPU_PAT_CRACKLE) value = crackle_pattern(S, pg, p, gen);
    else if (pg.pattern == PVGPU_PAT_PIGMENT) {                                                           // PigmentPattern::EvaluateRaw, pattern.cpp:7974-7990
        float Col[5];
        const bool colour_found = Compute_Pigment(Col, (int)pg.data, p);
        value = colour_found ? (double)(float)(0.297 * Col[0] + 0.589 * Col[1] + 0.114 * Col[2]) : 0.0;      // TransColour::Greyscale (colour.h:232-237)
    }
    else if (pg.pattern == PVGPU_PAT_FRACTAL) value = fractal_pattern(&S.shape_data[pg.data], p);
    else if (pg.pattern == PVGPU_PAT_SPIRAL1 || pg.pattern == PVGPU_PAT_SPIRAL2) {                        // pattern.cpp:8396-8437, 8473-8517
        const double M_PI_2_ = 1.57079632679489661923, TWO_M_PI = 6.283185307179586476925286766560;
        const double turb_val = turb ? turb->turbulence[0] * Turbulence(S, p, *turb, gen) : 0.0;
        const double rad = std::sqrt(p.x * p.x + p.y * p.y);
        double phi;
        if (rad == 0.0) phi = 0.0;
        else if (p.x < 0.0) phi = 3.0 * M_PI_2_ - std::asin(p.y / rad);
        else phi = M_PI_2_ + std::asin(p.y / rad);
        const double spiral = p.z + rad + pg.p[0] * phi / TWO_M_PI + turb_val;
        value = (pg.pattern == PVGPU_PAT_SPIRAL1) ? spiral : Triangle_Wave(rad) + Triangle_Wave(spiral);
    }
    else if (pg.pattern == PVGPU_PAT_CELLS)                                                               // pattern.cpp:5652-5660
        value = std::min(S.patternRands[S.hashTable[S.hashTable[S.hashTable[(int)std::floor(p.x + EPSILON) & 0xfff] ^ ((int)std::floor(p.y + EPSILON) & 0xfff)] ^
                                                    ((int)std::floor(p.z + EPSILON) & 0xfff)] % 32768u], 1.0);
    if (!discrete && pg.wave_type != PVGPU_WAVE_RAW) {                                                    // pattern.cpp:354-392
        if (pg.frequency != 0.0f) value = std::fmod(value * (double)pg.frequency + (double)pg.phase, 1.00001);
        if (value < 0.0) value -= std::floor(value);
        switch (pg.wave_type) {
            case PVGPU_WAVE_SINE: value = (1.0 + cycloidal(value)) * 0.5; break;
            case PVGPU_WAVE_TRIANGLE: value = Triangle_Wave(value); break;
            case PVGPU_WAVE_SCALLOP: value = std::fabs(cycloidal(value * 0.5)); break;
            case PVGPU_WAVE_CUBIC: value = sqr(value) * ((-2.0 * value) + 3.0); break;
            case PVGPU_WAVE_POLY: value = std::pow(value, (double)pg.exponent); break;
        }
    }
    return value;
}

V3 Tracer::Perturb_Normal(V3 Layer_Normal, int tnormal, V3 EPoint) const                                  // normal.cpp:784-927
{
    const pvgpu_tnormal& tn = S.tnormals[tnormal];
    const pvgpu_pigment& pat = S.pigments[tn.pattern];
    const bool DontScaleBumps = (tn.flags & PVGPU_DONT_SCALE_BUMPS_FLAG) != 0;
    // Warp_Normal (warp.cpp:563-580)
    if (!DontScaleBumps) Layer_Normal = unit(Layer_Normal);
    for (int i = (int)pat.warp_count - 1; i >= 0; i--) {
        const pvgpu_warp& w = S.warps[pat.warp_first + i];
        if (w.type == PVGPU_WARP_TRANSFORM) Layer_Normal = mtransposed(S.xf[w.transform].matrix, Layer_Normal);      // MInvTransNormal
    }
    if (!DontScaleBumps) Layer_Normal = unit(Layer_Normal);
    if (tn.normal_map) {
        const pvgpu_blend_map& m = S.maps[tn.normal_map - 1];
        const pvgpu_blend_entry* e = S.entries.data() + m.entry_first;
        auto warpn = [&](V3 v) {                                                                          // Warp_Normal
            if (!DontScaleBumps) v = unit(v);
            for (int i = (int)pat.warp_count - 1; i >= 0; i--) { const pvgpu_warp& w = S.warps[pat.warp_first + i]; if (w.type == PVGPU_WARP_TRANSFORM) v = mtransposed(S.xf[w.transform].matrix, v); }
            if (!DontScaleBumps) v = unit(v);
            return v;
        };
        auto unwarpn = [&](V3 v) {                                                                        // UnWarp_Normal
            if (!DontScaleBumps) v = unit(v);
            for (uint32_t i = 0; i < pat.warp_count; i++) { const pvgpu_warp& w = S.warps[pat.warp_first + i]; if (w.type == PVGPU_WARP_TRANSFORM) v = MTransNormal(S.xf[w.transform], v); }
            if (!DontScaleBumps) v = unit(v);
            return v;
        };
        // (Layer_Normal was already put through Warp_Normal above)
        if (tn.type == PVGPU_NORM_AVERAGE) {                                                              // normal.cpp:861-883, 1033-1059
            const V3 TPointA = Warp_EPoint(pat, EPoint);
            V3 V1 = v3(0.0, 0.0, 0.0);
            float Total = 0.0f;
            for (uint32_t i = 0; i < m.entry_count; i++) {
                V3 V2 = Perturb_Normal(Layer_Normal, (int)e[i].colour[0], TPointA);
                V1 = V1 + V2 * (double)e[i].value;
                Total += e[i].value;
            }
            return unwarpn(V1 / (double)Total);
        }
        const V3 TPointM = Warp_EPoint(pat, EPoint);                                                      // normal.cpp:824-848
        const double value1 = Evaluate_TPat(pat, TPointM);
        const uint32_t Max_Ent = m.entry_count - 1;
        uint32_t iP, iN; double prevW = 0.0, curW = 1.0;
        if (value1 >= e[Max_Ent].value) iP = iN = Max_Ent;
        else {
            iP = iN = 0;
            while (value1 > e[iN].value) { iP = iN; iN++; }
            if ((value1 == e[iN].value) || (iP == iN)) iP = iN;
            else { prevW = (e[iN].value - value1) / (e[iN].value - e[iP].value); curW = 1.0 - prevW; }
        }
        V3 P1 = Layer_Normal;
        Layer_Normal = Perturb_Normal(Layer_Normal, (int)e[iN].colour[0], TPointM);
        if (iP != iN) { P1 = Perturb_Normal(P1, (int)e[iP].colour[0], TPointM); Layer_Normal = prevW * P1 + curW * Layer_Normal; }
        (void)warpn;
        return unit(unwarpn(Layer_Normal));
    }
    const V3 TPoint = Warp_EPoint(pat, EPoint);
    const double Amount = (double)tn.amount;
    switch (tn.type) {
        case PVGPU_NORM_BUMPS: Layer_Normal = Layer_Normal + DNoise(S, TPoint) * Amount; break;             // normal.cpp:235-246
        case PVGPU_NORM_DENTS: {                                                                          // normal.cpp:272-288
            double noise = Noise(S, TPoint, pat.noise_generator ? pat.noise_generator : S.g.noise_generator);
            noise = noise * noise * noise * tn.amount;
            Layer_Normal = Layer_Normal + DNoise(S, TPoint) * noise;
            break;
        }
        case PVGPU_NORM_RIPPLES:                                                                          // normal.cpp:130-155
            for (unsigned i = 0; i < S.g.number_of_waves; i++) {
                V3 point = TPoint - S.waveSources[i];
                double length = len(point);
                if (length == 0.0) length = 1.0;
                double index = length * pat.frequency + pat.phase;
                double scalar = cycloidal(index) * tn.amount;
                Layer_Normal = Layer_Normal + point * (scalar / (length * (double)S.g.number_of_waves));
            }
            break;
        case PVGPU_NORM_WAVES:                                                                            // normal.cpp:180-209
            for (unsigned i = 0; i < S.g.number_of_waves; i++) {
                V3 point = TPoint - S.waveSources[i];
                double length = len(point);
                if (length == 0.0) length = 1.0;
                double index = length * pat.frequency * S.waveFrequencies[i] + pat.phase;
                double sinValue = cycloidal(index);
                double scalar = sinValue * tn.amount / S.waveFrequencies[i];
                Layer_Normal = Layer_Normal + point * (scalar / (length * (double)S.g.number_of_waves));
            }
            break;
        case PVGPU_NORM_WRINKLES: {                                                                       // normal.cpp:325-347
            double scale = 1.0;
            V3 result = v3(0.0, 0.0, 0.0);
            for (int i = 0; i < 10; scale *= 2.0, i++) {
                V3 value = DNoise(S, TPoint * scale);
                result = v3(result.x + std::fabs(value.x / scale), result.y + std::fabs(value.y / scale), result.z + std::fabs(value.z / scale));
            }
            Layer_Normal = Layer_Normal + result * Amount;
            break;
        }
        case PVGPU_NORM_QUILTED: {                                                                        // normal.cpp:371-391, pattern.cpp:8949-8969
            V3 value = v3(TPoint.x - FLOOR(TPoint.x) - 0.5, TPoint.y - FLOOR(TPoint.y) - 0.5, TPoint.z - FLOOR(TPoint.z) - 0.5);
            double t = len(value);
            const double p1 = pat.p[0], p2 = pat.p[1];
            double it = (1 - t), itsqrd = it * it, tsqrd = t * t, tcubed = t * tsqrd;
            t = (tcubed + 3.0 * t * itsqrd * p1 + 3.0 * tsqrd * it * p2) * 1.154700538;
            value = value * t;
            Layer_Normal = Layer_Normal + value * Amount;
            break;
        }
        default: {                                                                                        // normal.cpp:893-918
            static const V3 Pyramid_Vect[4] = { { 0.942809041, -0.333333333, 0.0 }, { -0.471404521, -0.333333333, 0.816496581 },
                                                { -0.471404521, -0.333333333, -0.816496581 }, { 0.0, 1.0, 0.0 } };
            double Amt = tn.amount * -5.0;
            Amt *= 0.02 / tn.delta;
            for (int i = 0; i <= 3; i++) {
                V3 P1 = TPoint + Pyramid_Vect[i] * (double)tn.delta;
                double value1 = Evaluate_TPat(pat, P1);
                if (tn.slope_count) {                                                                     // Do_Slope_Map / Hermite_Cubic normal.cpp:929-1001
                    const pvgpu_slope_entry* e = S.slope_entries.data() + tn.slope_first;
                    const uint32_t Max_Ent = tn.slope_count - 1;
                    uint32_t iP, iN; double prevW = 0.0, curW = 1.0;
                    if (value1 >= e[Max_Ent].value) iP = iN = Max_Ent;
                    else {
                        iP = iN = 0;
                        while (value1 > e[iN].value) { iP = iN; iN++; }
                        if ((value1 == e[iN].value) || (iP == iN)) iP = iN;
                        else { prevW = (e[iN].value - value1) / (e[iN].value - e[iP].value); curW = 1.0 - prevW; }
                    }
                    if (iP == iN) value1 = e[iN].height;
                    else {
                        const double T1 = curW, TT = T1 * T1, TTT = TT * T1;
                        double rv = TTT * (e[iP].slope + e[iN].slope + 2.0 * (e[iP].height - e[iN].height));
                        rv += -TT * (2.0 * e[iP].slope + e[iN].slope + 3.0 * (e[iP].height - e[iN].height));
                        rv += T1 * e[iP].slope + e[iP].height;
                        value1 = rv;
                    }
                }
                Layer_Normal = Layer_Normal + Pyramid_Vect[i] * (value1 * Amt);
            }
            break;
        }
    }
    // UnWarp_Normal (warp.cpp:603-620)
    if (!DontScaleBumps) Layer_Normal = unit(Layer_Normal);
    for (uint32_t i = 0; i < pat.warp_count; i++) {
        const pvgpu_warp& w = S.warps[pat.warp_first + i];
        if (w.type == PVGPU_WARP_TRANSFORM) Layer_Normal = MTransNormal(S.xf[w.transform], Layer_Normal);
    }
    if (!DontScaleBumps) Layer_Normal = unit(Layer_Normal);
    return Layer_Normal;
}

// map_pos (imageutil.cpp:932-999) with its mappers (:557-930), then image_colour_at(..., premul = false) (:396-466) with
// no_interpolation / Interp / InterpolateBicubic (:1001-1203) on the decoded texels.  false = outside the map (once).
static double img_wrap(double val, double upperLimit)                                                     // mathutil.h:102-127
{
    double t = std::fmod(val, upperLimit);
    if (t < 0.0) t += upperLimit;
    if (t >= upperLimit) t = 0.0;
    return t;
}
static double img_theta(double x, double z, double len)
{
    if (z == 0.0) return (x > 0) ? 0.0 : M_PI;
    double theta = std::acos(x / len);
    if (z < 0.0) theta = 2.0 * M_PI - theta;
    return theta;
}
bool Tracer::image_map_colour(const pvgpu_image& im, V3 p, float col[5]) const
{
    const bool once = (im.flags & PVGPU_IMAGE_ONCE) != 0;
    const double W = im.fwidth, H = im.fheight;
    double x = p.x, y = p.y, z = p.z, u = 0.0, v = 0.0, len;
    bool inside = true;
    switch (im.map_type) {
        case 1:                                                                                           // spherical_image_map
            len = std::sqrt(x * x + y * y + z * z);
            if (len == 0.0) { inside = false; break; }
            x /= len; y /= len; z /= len;
            v = (0.5 + std::asin(y) / M_PI) * H;
            len = std::sqrt(x * x + z * z);
            u = ((len == 0.0) ? 0.0 : img_theta(x, z, len) / (2.0 * M_PI)) * W;
            break;
        case 2:                                                                                           // cylindrical_image_map
            if (once && ((y < 0.0) || (y > 1.0))) { inside = false; break; }
            v = std::fmod(y * H, H);
            len = std::sqrt(x * x + y * y + z * z);
            if (len == 0.0) { inside = false; break; }
            x /= len; z /= len;
            len = std::sqrt(x * x + z * z);
            if (len == 0.0) { inside = false; break; }
            u = (img_theta(x, z, len) / (2.0 * M_PI)) * W;
            break;
        case 5: {                                                                                         // torus_image_map
            const double r0 = im.gradient[0];
            len = std::sqrt(x * x + z * z);
            if (len == 0.0) { inside = false; break; }
            double theta = 0.0 - img_theta(x, z, len);
            x = len - r0;
            len = std::sqrt(x * x + y * y);
            double phi = std::acos(-x / len);
            if (y > 0.0) phi = 2.0 * M_PI - phi;
            theta /= 2.0 * M_PI; phi /= 2.0 * M_PI;
            u = -theta * W; v = phi * H;
            break;
        }
        case 7: {                                                                                         // angular_image_map
            len = std::sqrt(x * x + y * y + z * z);
            if (len == 0.0) { inside = false; break; }
            x /= len; y /= len; z /= len;
            const double r = ((x == 0) && (y == 0)) ? 0.0 : (1 / M_PI) * std::acos(z) / std::sqrt(x * x + y * y);
            u = (x * r + 1) / 2 * W; v = (y * r + 1) / 2 * H;
            break;
        }
        default: {                                                                                        // planar_image_map
            const double c[3] = { x, y, z };
            for (int k = 0; k < 3 && inside; k++)
                if (im.gradient[k] != 0.0) {
                    if (once && ((c[k] < 0.0) || (c[k] > 1.0))) { inside = false; break; }
                    if (im.gradient[k] > 0) u = std::fmod(c[k] * W, W); else v = std::fmod(c[k] * H, H);
                }
        }
    }
    if (inside) {
        u += im.offset[0] + EPSILON; v += im.offset[1] + EPSILON;
        if (once && ((u >= (double)im.width) || (v >= (double)im.height) || (u < 0.0) || (v < 0.0))) inside = false;
    }
    if (!inside) { col[0] = col[1] = col[2] = 1.0f; col[3] = 0.0f; col[4] = 1.0f; return false; }       // pattern.cpp:502-506
    double xcoor = img_wrap(u, (double)im.width), ycoor = img_wrap(-v, (double)im.height);
    auto texel = [&](double xc, double yc) -> const float* {                                              // no_interpolation
        int ix, iy;
        if (once) {
            ix = (xc < 0.0) ? 0 : (xc >= (double)im.width) ? (int)im.width - 1 : (int)xc;
            iy = (yc < 0.0) ? 0 : (yc >= (double)im.height) ? (int)im.height - 1 : (int)yc;
        } else { ix = (int)img_wrap(xc, (double)im.width); iy = (int)img_wrap(yc, (double)im.height); }
        return S.texels.data() + 5 * ((size_t)im.data_first + (size_t)iy * im.width + (size_t)ix);
    };
    if (im.interpolation == 0) { const float* t = texel(xcoor, ycoor); for (int k = 0; k < 5; k++) col[k] = t[k]; }
    else {
        double acc[5] = { 0, 0, 0, 0, 0 };
        xcoor += 0.5; ycoor += 0.5;
        const int iy = (int)ycoor, ix = (int)xcoor;
        if (im.interpolation == 3) {                                                                      // InterpolateBicubic
            auto cubic = [](double* f, double xx) { double pp = xx - (int)xx, q = 1 - pp; f[0] = -0.5 * pp * q * q; f[1] = 0.5 * q * (q * (3 * pp + 1) + 1); f[2] = 0.5 * pp * (pp * (3 * q + 1) + 1); f[3] = -0.5 * q * pp * pp; };
            double fx[4], fy[4];
            cubic(fx, xcoor); cubic(fy, ycoor);
            for (int i = 0; i < 4; i++)
                for (int j = 0; j < 4; j++) {
                    const float* t = texel((double)ix + i - 2, (double)iy + j - 2);
                    const double f = fx[i] * fy[j];
                    for (int k = 0; k < 5; k++) acc[k] += (double)t[k] * f;
                }
        } else {                                                                                          // Interp: bilinear / norm_dist
            const double pp = xcoor - (int)xcoor, q = ycoor - (int)ycoor;
            double f[4];
            if (im.interpolation == 2) { f[0] = pp * q; f[1] = (1 - pp) * q; f[2] = pp * (1 - q); f[3] = (1 - pp) * (1 - q); }
            else {
                double w[4] = { 1.0 / ((1 - pp) * (1 - pp) + (1 - q) * (1 - q)), 1.0 / (pp * pp + (1 - q) * (1 - q)), 1.0 / ((1 - pp) * (1 - pp) + q * q), 1.0 / (pp * pp + q * q) };
                double sum = 0.0;
                for (int i = 0; i < 4; i++) sum += w[i];
                for (int i = 0; i < 4; i++) f[i] = w[i] / sum;
            }
            const double cx[4] = { (double)ix, (double)ix - 1, (double)ix, (double)ix - 1 }, cy[4] = { (double)iy, (double)iy, (double)iy - 1, (double)iy - 1 };
            for (int i = 0; i < 4; i++) { const float* t = texel(cx[i], cy[i]); for (int k = 0; k < 5; k++) acc[k] += (double)t[k] * f[i]; }
        }
        for (int k = 0; k < 5; k++) col[k] = (float)acc[k];
    }
    if (im.flags & PVGPU_IMAGE_PREMULTIPLIED) { float a = 1.0f - col[4]; if (a == 0) a = 1.0e-6f; col[0] /= a; col[1] /= a; col[2] /= a; }      // AlphaUnPremultiply
    if (im.flags & PVGPU_IMAGE_TRANSMIT_ALL) { const float alpha = 1.0f - col[4]; if (alpha != 0.0f) { col[4] += im.all_transmit * alpha; col[3] += im.all_filter * alpha; } }
    return true;
}

// <Object>::UVCoord as the point (u, v, 0): sphere.cpp:688-752, box.cpp:1028-1077, torus.cpp:1118-1147, mesh.cpp:2256-2332, object.cpp:882-886
V3 Tracer::UVCoord(const Intersection& isect) const
{
    const pvgpu_object& ob = S.objects[isect.Object];
    const double M_PI_ = 3.1415926535897932384626, TWO_M_PI = 6.283185307179586476925286766560;
    switch (ob.type) {
        case PVGPU_OBJ_SPHERE: {
            V3 New_Point;
            if (ob.aux) New_Point = MInvTransPoint(S.xf[ob.transform], isect.IPoint);
            else { New_Point = isect.IPoint - v3(ob.p); if (ob.transform >= 0) New_Point = MInvTransPoint(S.xf[ob.transform], New_Point); }
            double x = New_Point.x, y = New_Point.y, z = New_Point.z, phi, theta;
            double l = std::sqrt(x * x + y * y + z * z);
            if (l == 0.0) return v3(0, 0, 0);
            x /= l; y /= l; z /= l;
            phi = 0.5 + std::asin(y) / M_PI_;
            l = x * x + z * z;
            if (l > EPSILON) {
                l = std::sqrt(l);
                if (z == 0.0) { if (x > 0) theta = 0.0; else theta = M_PI_; }
                else { theta = std::acos(x / l); if (z < 0.0) theta = TWO_M_PI - theta; }
                theta /= TWO_M_PI;
            } else theta = 0;
            return v3(theta, phi, 0.0);
        }
        case PVGPU_OBJ_BOX: {
            V3 P = (ob.transform >= 0) ? MInvTransPoint(S.xf[ob.transform], isect.IPoint) : isect.IPoint;
            V3 Box_Diff = v3(ob.p + 3) - v3(ob.p);
            P = P - v3(ob.p);
            P = v3(P.x / Box_Diff.x, P.y / Box_Diff.y, P.z / Box_Diff.z);
            switch (isect.aux) {
                case 1: return v3((P.z / 4.0), (1.0 / 3.0) + (P.y / 3.0), 0.0);
                case 2: return v3((3.0 / 4.0) - (P.z / 4.0), (1.0 / 3.0) + (P.y / 3.0), 0.0);
                case 3: return v3((1.0 / 4.0) + (P.x / 4.0), (P.z / 3.0), 0.0);
                case 4: return v3((1.0 / 4.0) + (P.x / 4.0), (3.0 / 3.0) - (P.z / 3.0), 0.0);
                case 5: return v3(1.0 - (P.x / 4.0), (1.0 / 3.0) + (P.y / 3.0), 0.0);
                default: return v3((1.0 / 4.0) + (P.x / 4.0), (1.0 / 3.0) + (P.y / 3.0), 0.0);
            }
        }
        case PVGPU_OBJ_TORUS: {
            V3 P = MInvTransPoint(S.xf[ob.transform], isect.IPoint);
            double x = P.x, y = P.y, z = P.z;
            double u = (1.0 - (std::atan2(z, x) + M_PI_) / TWO_M_PI);
            double l = std::sqrt(x * x + z * z);
            x = l - ob.p[0];
            double v = (std::atan2(y, x) + M_PI_) / TWO_M_PI;
            return v3(u, v, 0.0);
        }
        case PVGPU_OBJ_MESH: {
            if (S.tri_uv.empty()) return v3(0, 0, 0);
            const pvgpu_mesh& me = S.meshes[ob.mesh];
            V3 P = (ob.transform >= 0) ? MInvTransPoint(S.xf[ob.transform], isect.IPoint) : isect.IPoint;
            const pvgpu_triangle& T = S.tris[isect.aux];
            const float* V = S.verts.data() + 3 * (size_t)me.vertex_first;
            auto vert = [&](int i) { return V + 3 * (size_t)i; };
            auto sngl_diff = [](const float* a, const float* b) { return v3((double)(float)(a[0] - b[0]), (double)(float)(a[1] - b[1]), (double)(float)(a[2] - b[2])); };   // SNGL vectors
            auto weight = [&](const float* far_a, const float* far_b, const float* own) {
                V3 Side1 = sngl_diff(far_a, far_b), Side2 = sngl_diff(far_a, own);
                V3 vA = P - v3((double)own[0], (double)own[1], (double)own[2]);
                double t1 = dot(Side2, Side1), t2 = dot(Side1, Side1);
                V3 vB = Side1 * (t1 / t2) - Side2;
                t1 = dot(vA, vB); t2 = dot(vB, vB);
                return 1 + t1 / t2;
            };
            const double w1 = weight(vert(T.p3), vert(T.p2), vert(T.p1)), w2 = weight(vert(T.p3), vert(T.p1), vert(T.p2)), w3 = weight(vert(T.p2), vert(T.p1), vert(T.p3));
            const uint32_t* iu = S.tri_uv.data() + 3 * (size_t)isect.aux;
            const double* uv = S.mesh_uv.data();
            return v3((w1 * uv[2 * iu[0]] + w2 * uv[2 * iu[1]]) + w3 * uv[2 * iu[2]], (w1 * uv[2 * iu[0] + 1] + w2 * uv[2 * iu[1] + 1]) + w3 * uv[2 * iu[2] + 1], 0.0);
        }
        default: return v3(isect.IPoint.x, isect.IPoint.y, 0.0);
    }
}

bool Tracer::Compute_Pigment(float col[5], int pigment, V3 EPoint) const                                  // pigment.cpp:395-466
{
    const pvgpu_pigment& pg = S.pigments[pigment];
    if (pg.pattern == PVGPU_PAT_PLAIN) { for (int k = 0; k < 5; k++) col[k] = pg.colour[k]; return true; }
    if (pg.pattern == PVGPU_PAT_UV_MAP) {                                                                 // PigmentBlendMap::ComputeUVMapped, pigment.cpp:603-618
        if (cur_isect == nullptr) { for (int k = 0; k < 5; k++) col[k] = 0.0f; return false; }       // (the reference throws: no uv_mapping outside a hit)
        return Compute_Pigment(col, (int)pg.data, UVCoord(*cur_isect));
    }
    const V3 TPoint = Warp_EPoint(pg, EPoint);
    if (pg.pattern == PVGPU_PAT_IMAGE_MAP) return image_map_colour(S.images[pg.data], TPoint, col);      // ColourImagePattern::Evaluate
    const pvgpu_blend_map& m = S.maps[pg.blend_map];
    const pvgpu_blend_entry* e = S.entries.data() + m.entry_first;
    const bool pmap = (m.blend_mode & PVGPU_BLEND_PIGMENT_MAP) != 0;
    bool found = !pmap;
    auto entry_colour = [&](uint32_t i, float out[5]) {                    // BlendMapEntry<TransColour> or BlendMapEntry<PIGMENT*>
        if (pmap) { if (Compute_Pigment(out, (int)e[i].colour[0], TPoint)) found = true; }
        else for (int k = 0; k < 5; k++) out[k] = e[i].colour[k];
    };
    if (pg.pattern == PVGPU_PAT_AVERAGE) {                                                                // pigment.cpp:566-596
        float Total = 0.0f;
        for (int k = 0; k < 5; k++) col[k] = 0.0f;
        for (uint32_t i = 0; i < m.entry_count; i++) {
            float t[5];
            entry_colour(i, t);
            for (int k = 0; k < 5; k++) col[k] += (float)(t[k] * (double)e[i].value);
            Total += e[i].value;
        }
        for (int k = 0; k < 5; k++) col[k] = (float)(col[k] / (double)Total);
        return true;
    }
    const double value = Evaluate_TPat(pg, TPoint);
    // BlendMap::Search + ColourBlendMap / PigmentBlendMap::Compute (pattern.cpp:1068-1112, pigment.cpp:513-564)
    const uint32_t Max_Ent = m.entry_count - 1;
    uint32_t iP, iN; double prevW = 0.0, nextW = 1.0;
    if (value >= e[Max_Ent].value) iP = iN = Max_Ent;
    else {
        iP = iN = 0;
        while (value > e[iN].value) { iP = iN; iN++; }
        if ((value == e[iN].value) || (iP == iN)) iP = iN;
        else { prevW = (e[iN].value - value) / (e[iN].value - e[iP].value); nextW = 1.0 - prevW; }
    }
    entry_colour(iN, col);
    if (iP != iN) {
        float t[5];
        entry_colour(iP, t);
        for (int k = 0; k < 5; k++) col[k] = (float)(t[k] * prevW) + (float)(col[k] * nextW);
    }
    return found;
}

static double FresnelR(double cosTi, double n)                                                            // trace.cpp:2680-2708
{
    double sqrg = sqr(n) + sqr(cosTi) - 1.0;
    if (sqrg <= 0.0) return 1.0;
    double g = std::sqrt(sqrg), quot1 = (g - cosTi) / (g + cosTi), quot2 = (cosTi * (g + cosTi) - 1.0) / (cosTi * (g - cosTi) + 1.0);
    double f = 0.5 * sqr(quot1) * (1.0 + sqr(quot2));
    return std::min(std::max(f, 0.0), 1.0);
}
static void ComputeMetallic(Col& c, double metallic, Col mc, double cosAngle)                             // trace.cpp:2656-2669
{
    if (metallic != 0.0) {
        double x = std::fabs(std::acos(cosAngle)) / 1.57079632679489661923;
        double F = 0.014567225 / sqr(x - 1.12) - 0.011612903;
        F = std::min(1.0, std::max(0.0, F));
        c.r *= (float)(1.0 + (metallic * (1.0 - F)) * (mc.r - 1.0)); c.g *= (float)(1.0 + (metallic * (1.0 - F)) * (mc.g - 1.0)); c.b *= (float)(1.0 + (metallic * (1.0 - F)) * (mc.b - 1.0));
    }
}
static double cubic_spline(double low, double high, double pos)                                           // lightsource.cpp:501-517
{
    if (pos < low) return 0.0;
    if (pos >= high) return 1.0;
    pos = (pos - low) / (high - low);
    return (3 - 2 * pos) * pos * pos;
}
static double Attenuate_Light(const pvgpu_light& L, const Ray& ray, double Distance)                      // lightsource.cpp:548-633
{
    double Attenuation = 1.0;
    if (L.type == PVGPU_LIGHT_SPOT) {
        double costheta = dot(ray.Direction, v3(L.direction));
        if (Distance > 0.0) costheta = -costheta;
        if (costheta > 0.0) { Attenuation = std::pow(costheta, L.coeff); if (L.radius > 0.0 && costheta < L.radius) Attenuation *= cubic_spline(L.falloff, L.radius, costheta); }
        else return 0.0;
    } else if (L.type == PVGPU_LIGHT_CYLINDER) {
        V3 V1 = ray.Origin - v3(L.center);
        double k = dot(V1, v3(L.direction));
        if (k > 0.0) {
            V3 P = V1 - k * v3(L.direction);
            double l = len(P);
            if (l < L.falloff) { double dist = 1.0 - l / L.falloff; Attenuation = std::pow(dist, L.coeff); if (L.radius > 0.0 && l > L.radius) Attenuation *= cubic_spline(0.0, 1.0 - L.radius / L.falloff, dist); }
            else return 0.0;
        } else return 0.0;
    }
    if (Attenuation > 0.0 && L.fade_power > 0.0) {
        if (std::fabs(L.fade_distance) >= EPSILON) Attenuation *= 2.0 / (1.0 + std::pow(Distance / L.fade_distance, L.fade_power));
        else Attenuation *= std::pow(Distance, -L.fade_power);
    }
    return Attenuation;
}

double Tracer::relative_ior(const Ray& ray, int interior) const                                           // trace.cpp:2595-2625
{
    if (interior < 0) return 1.0;
    // SceneData::atmosphereIOR is DBL, Interior::IOR is SNGL: atmosphere ratios divide in FP64, object / object ratios in FP32
    const float ior = S.interiors[interior].ior;
    const double atmosphereIOR = S.g.atmosphere_ior;
    if (ray.interiors.empty()) return ior / atmosphereIOR;
    if (ray.IsInterior(interior)) {
        if (ray.interiors.size() == 1) return atmosphereIOR / ior;
        return S.interiors[ray.interiors.back()].ior / ior;
    }
    return ior / S.interiors[ray.interiors.back()].ior;
}

int Tracer::hit_texture(const pvgpu_object& ob, const Intersection& isect, bool backside) const
{
    if (ob.type == PVGPU_OBJ_MESH && (ob.flags & PVGPU_MULTITEXTURE_FLAG)) {                              // mesh.cpp:2421-2457
        if (backside && ob.interior_texture >= 0) return ob.interior_texture;
        const pvgpu_triangle& tr = S.tris[isect.aux];
        if (tr.texture >= 0) return (int)S.index_list[S.meshes[ob.mesh].texture_first + tr.texture];
        return ob.texture;
    }
    if (ob.texture < 0) return -1;
    return (backside && ob.interior_texture >= 0) ? ob.interior_texture : ob.texture;                     // trace.cpp:513-530
}

// The weighted texture list of a hit (trace.cpp:513-530): one texture of weight 1, or Blob::Determine_Textures (blob.cpp:2768-2880)
// for a blob with per-component textures: every component with a non-zero field at the point, weight |field|, normalised.
std::vector<std::pair<float, int>> Tracer::Determine_Textures(const pvgpu_object& ob, const Intersection& isect, bool backside) const
{
    std::vector<std::pair<float, int>> out;
    if (ob.type == PVGPU_OBJ_BLOB && (ob.flags & PVGPU_MULTITEXTURE_FLAG) && !S.blob_textures.empty()) {
        const pvgpu_blob& bl = S.blobs[ob.mesh];
        const pvgpu_blob_element* el = S.blob_elements.data() + bl.element_first;
        const V3 P = (ob.transform >= 0) ? MInvTransPoint(S.xf[ob.transform], isect.IPoint) : isect.IPoint;
        auto add = [&](uint32_t ei) {
            const double density = std::fabs(blob_element_field(el[ei], P));
            if (density > 0.0) { const int t = S.blob_textures[bl.element_first + ei]; out.push_back({ (float)density, t >= 0 ? t : ob.texture }); }
        };
        if (bl.node_count == 0) for (uint32_t i = 0; i < bl.element_count; i++) add(i);
        else {
            const pvgpu_blob_node* nodes = S.blob_nodes.data() + bl.node_first;
            std::vector<uint32_t> queue{ 0u };
            while (!queue.empty()) {
                const pvgpu_blob_node nd = nodes[queue.back()]; queue.pop_back();
                if (nd.count == 0) { add(nd.first); continue; }
                for (uint32_t i = 0; i < nd.count; i++) { const pvgpu_blob_node& ch = nodes[nd.first + i]; if (len2(P - v3(ch.c)) <= ch.r2) queue.push_back(nd.first + i); }
            }
        }
        if (!out.empty()) { float sum = 0.0f; for (auto& e : out) sum += e.first; sum = 1.0f / sum; for (auto& e : out) e.first *= sum; }
        return out;
    }
    const int tex = hit_texture(ob, isect, backside);
    if (tex >= 0) out.push_back({ 1.0f, tex });
    return out;
}

// GenericColour::operator*=(double) rounds to FP32 after the FP64 product (colour.h:1681)
static inline Col cmul(Col a, double b) { return Col{ (float)(a.r * b), (float)(a.g * b), (float)(a.b * b) }; }
static inline Col cadd(Col a, double b) { return Col{ (float)(a.r + b), (float)(a.g + b), (float)(a.b + b) }; }

void Tracer::ComputeSky(const Ray& ray, const Ticket& tk, Col& colour, float& transm) const               // trace.cpp:2769-2890
{
    const float* bg = S.g.background;
    const pvgpu_sky_sphere* sky = S.sky_spheres.empty() ? nullptr : &S.sky_spheres[0];
    V3 p = ray.Direction;
    if (sky && sky->transform >= 0) p = MInvTransPoint(S.xf[sky->transform], ray.Direction);
    if (S.g.language_version < 370) {
        if (tk.alphaBackground) { colour = Col{ 0, 0, 0 }; transm = 1.0f; return; }
        colour = Col{ bg[0], bg[1], bg[2] }; transm = bg[4];
        if (!sky) return;
        Col col{ 0, 0, 0 }, filterc_colour{ 1, 1, 1 };
        float filterc_filter = 1.0f, filterc_transm = 1.0f;
        double trans = 1.0;
        for (int i = (int)sky->pigment_count - 1; i >= 0; i--) {
            float t[5];
            Compute_Pigment(t, (int)S.index_list[sky->pigment_first + i], p);
            double att = trans * (float)(1.0 - t[3] - t[4]);
            col = col + cmul(Col{ t[0], t[1], t[2] }, att);
            filterc_colour = filterc_colour * Col{ t[0], t[1], t[2] };
            filterc_filter *= t[3]; filterc_transm *= t[4];
            trans = std::fabs(filterc_filter) + std::fabs(filterc_transm);
        }
        col = col * Col{ sky->emission[0], sky->emission[1], sky->emission[2] };
        Col transColour = cadd(cmul(filterc_colour, filterc_filter), filterc_transm);
        colour = colour * transColour + col;
        transm *= filterc_transm;
        return;
    }
    Col filCol{ 1, 1, 1 }, col{ 0, 0, 0 };
    if (sky) {
        const Col Emission{ sky->emission[0], sky->emission[1], sky->emission[2] };
        for (int i = (int)sky->pigment_count - 1; i >= 0; i--) {
            float t[5];
            Compute_Pigment(t, (int)S.index_list[sky->pigment_first + i], p);
            double att = (float)(1.0 - t[3] - t[4]);
            col = col + cmul(Col{ t[0], t[1], t[2] }, att) * filCol * Emission;
            filCol = filCol * cadd(cmul(Col{ t[0], t[1], t[2] }, t[3]), t[4]);
        }
    }
    float f = tk.alphaBackground ? bg[3] : 0.0f, t = tk.alphaBackground ? bg[4] : 0.0f;
    double att = (float)(1.0 - f - t);
    col = col + cmul(Col{ bg[0], bg[1], bg[2] }, att) * filCol;
    filCol = filCol * cadd(cmul(Col{ bg[0], bg[1], bg[2] }, f), t);
    colour = col;
    transm = std::min(1.0f, std::fabs(grey(filCol)));
}

void Tracer::ComputeFog(const Ray& ray, double Depth, Col& colour, float& transm) const                   // trace.cpp:2892-3044
{
    Col sum_att{ 1, 1, 1 }, sum_col{ 0, 0, 0 };
    for (const pvgpu_fog& fog : S.fogs) {
        if (!(std::fabs(fog.distance) > EPSILON)) continue;
        double width = Depth, att;
        const pvgpu_warp* Turb = fog.turbulence >= 0 ? &S.warps[fog.turbulence] : nullptr;
        if (fog.type == PVGPU_FOG_GROUND) {                                                               // ComputeGroundFogDepth
            V3 p1 = ray.Evaluate(0.0), p2 = p1 + ray.Direction * width;
            double y1 = dot(p1, v3(fog.up)), y2 = dot(p2, v3(fog.up));
            double start = (y1 - fog.offset) / fog.alt, end = (y2 - fog.offset) / fog.alt, fog_density;
            if (start <= 0.0) {
                if (end <= 0.0) fog_density = 1.0;
                else fog_density = (std::atan(end) - start) / (end - start);
            } else {
                if (end <= 0.0) fog_density = (std::atan(start) - end) / (start - end);
                else {
                    double delta = start - end;
                    if (std::fabs(delta) > EPSILON) fog_density = (std::atan(start) - std::atan(end)) / delta;
                    else fog_density = 1.0 / (sqr(start) + 1.0);
                }
            }
            if (Turb) {
                V3 p = (p1 + p2) * 0.5;
                p = v3(p.x * Turb->turbulence[0], p.y * Turb->turbulence[1], p.z * Turb->turbulence[2]);
                double k = std::exp(-width / fog.distance);
                width *= (1.0 - k * std::min(1.0, Turbulence(S, p, *Turb, S.g.noise_generator) * fog.turb_depth));
            }
            att = std::exp(-width * fog_density / fog.distance);
        } else {                                                                                          // ComputeConstantFogDepth
            if (Turb) {
                V3 p = ray.Evaluate(width / 2.0);
                p = v3(p.x * Turb->turbulence[0], p.y * Turb->turbulence[1], p.z * Turb->turbulence[2]);
                double k = std::exp(-width / fog.distance);
                width *= (1.0 - k * std::min(1.0, Turbulence(S, p, *Turb, S.g.noise_generator) * fog.turb_depth));
            }
            att = std::exp(-width / fog.distance);
        }
        const Col col_fog{ fog.colour[0], fog.colour[1], fog.colour[2] };
        const float filter_fog = fog.colour[3], transm_fog = fog.colour[4];
        if (att < transm_fog) att = transm_fog;
        sum_att = sum_att * cmul(cadd(cmul(col_fog, filter_fog), 1.0 - filter_fog), att);
        sum_col = sum_col + cmul(col_fog, 1.0 - att);
    }
    colour = sum_col + sum_att * colour;
    transm *= grey(sum_att);
}

double Tracer::TraceRay(Ray& ray, Ticket& tk, Col& colour, float& transm, float weight, bool continuedRay, double maxDepth)   // trace.cpp:135-228
{
    st.rays++;
    if ((tk.traceLevel >= tk.maxAllowedTraceLevel) || (weight < tk.adcBailout)) { colour = Col{ 0, 0, 0 }; transm = 0.0f; return HUGE_VALUE; }
    Intersection bestisect;
    if (maxDepth >= EPSILON) bestisect.Depth = maxDepth;
    bool found = FindIntersection(bestisect, ray, -1.0);
    const bool inc = !continuedRay;
    if (inc) { tk.traceLevel++; tk.maxFound = std::max(tk.maxFound, tk.traceLevel); }
    if (found) ComputeTextureColour(bestisect, colour, transm, ray, tk, weight);
    else ComputeSky(ray, tk, colour, transm);
    if ((S.g.quality_flags & PVGPU_Q_MEDIA) && !S.fogs.empty() && ray.IsHollowRay(S)) ComputeFog(ray, bestisect.Depth, colour, transm);   // trace.cpp:207-216
    if (inc) tk.traceLevel--;
    st.max_level = std::max(st.max_level, tk.maxFound);
    return found ? bestisect.Depth : HUGE_VALUE;
}

void Tracer::ComputeTextureColour(Intersection& isect, Col& colour, float& transm, Ray& ray, Ticket& tk, float weight)       // trace.cpp:457-586
{
    const pvgpu_object& ob = S.objects[isect.Object];
    V3 rawnormal = Normal(isect);
    if (ob.flags & PVGPU_INVERTED_FLAG) rawnormal = -rawnormal;
    double normaldirection = dot(rawnormal, ray.Direction);
    if (normaldirection > 0.0) rawnormal = -rawnormal;
    const std::vector<std::pair<float, int>> wtextures = Determine_Textures(ob, isect, normaldirection > 0.0);
    if (wtextures.empty()) return;
    Col tmpCol{ 0, 0, 0 }; float tmpTransm = 0.0f;
    for (const auto& wt : wtextures) {
        if ((wt.first < tk.adcBailout) || wt.second < 0) continue;                                         // trace.cpp:541
        Col c1{ 0, 0, 0 }; float t1 = 0.0f;
        std::vector<int> warps;
        cur_isect = &isect;
        const V3 ipoint = (ob.flags & PVGPU_UV_FLAG) ? UVCoord(isect) : isect.IPoint;                      // trace.cpp:500-512
        ComputeOneTextureColour(c1, t1, wt.second, warps, ipoint, rawnormal, ray, tk, weight, isect, false);
        tmpCol = tmpCol + c1 * wt.first; tmpTransm += wt.first * t1;
    }
    colour = colour + tmpCol; transm += tmpTransm;
}

// Warp_Normal / UnWarp_Normal through the patterned textures that enclose a layer (trace.cpp:816-827): transform warps only
V3 Tracer::Warp_Normal_Chain(V3 n, const std::vector<int>& warps, bool unwarp) const
{
    auto one = [&](int tex, V3 v) {
        const pvgpu_pigment& pat = S.pigments[S.textures[tex].pigment];
        v = unit(v);
        if (!unwarp) { for (int i = (int)pat.warp_count - 1; i >= 0; i--) { const pvgpu_warp& w = S.warps[pat.warp_first + i]; if (w.type == PVGPU_WARP_TRANSFORM) v = mtransposed(S.xf[w.transform].matrix, v); } }
        else { for (uint32_t i = 0; i < pat.warp_count; i++) { const pvgpu_warp& w = S.warps[pat.warp_first + i]; if (w.type == PVGPU_WARP_TRANSFORM) v = MTransNormal(S.xf[w.transform], v); } }
        return unit(v);
    };
    if (!unwarp) for (size_t i = 0; i < warps.size(); i++) n = one(warps[i], n);
    else for (size_t i = warps.size(); i-- > 0;) n = one(warps[i], n);
    return n;
}

void Tracer::ComputeOneTextureColour(Col& resultColour, float& resultTransm, int texture, std::vector<int>& warps, V3 ipoint, V3 rawnormal, Ray& ray, Ticket& tk,
                                     float weight, Intersection& isect, bool shadowflag)                  // trace.cpp:588-694
{
    const pvgpu_texture& tx = S.textures[texture];
    if (tx.type == PVGPU_PAT_PLAIN) {
        if (shadowflag) ComputeShadowTexture(resultColour, texture, warps, ipoint, rawnormal, ray, isect);
        else ComputeLightedTexture(resultColour, resultTransm, texture, ipoint, rawnormal, ray, tk, weight, isect, warps);
        return;
    }
    warps.push_back(texture);
    const pvgpu_pigment& pat = S.pigments[tx.pigment];
    const pvgpu_blend_map& m = S.maps[tx.blend_map];
    const pvgpu_blend_entry* e = S.entries.data() + m.entry_first;
    const V3 tpoint = Warp_EPoint(pat, ipoint);
    if (tx.type == PVGPU_PAT_AVERAGE) {                                                                   // ComputeAverageTextureColours trace.cpp:697-737
        float total = 0.0f;
        resultColour = Col{ 0, 0, 0 }; resultTransm = 0.0f;
        for (uint32_t i = 0; i < m.entry_count; i++) {
            Col lc{ 0, 0, 0 }; float lt = 0.0f;
            const float val = e[i].value;
            std::vector<int> w2(warps);
            ComputeOneTextureColour(lc, lt, (int)e[i].colour[0], w2, tpoint, rawnormal, ray, tk, weight, isect, shadowflag);
            resultColour = resultColour + cmul(lc, val); resultTransm += (float)(lt * (double)val);
            total += val;
        }
        resultColour = Col{ (float)(resultColour.r / (double)total), (float)(resultColour.g / (double)total), (float)(resultColour.b / (double)total) };
        resultTransm = (float)(resultTransm / (double)total);
        return;
    }
    const double value1 = Evaluate_TPat(pat, tpoint);
    const uint32_t Max_Ent = m.entry_count - 1;
    uint32_t iP, iN; double prevW = 0.0, curW = 1.0;
    if (value1 >= e[Max_Ent].value) iP = iN = Max_Ent;
    else {
        iP = iN = 0;
        while (value1 > e[iN].value) { iP = iN; iN++; }
        if ((value1 == e[iN].value) || (iP == iN)) iP = iN;
        else { prevW = (e[iN].value - value1) / (e[iN].value - e[iP].value); curW = 1.0 - prevW; }
    }
    {
        std::vector<int> w2(warps);
        ComputeOneTextureColour(resultColour, resultTransm, (int)e[iN].colour[0], w2, tpoint, rawnormal, ray, tk, weight, isect, shadowflag);
    }
    if (iP != iN) {
        Col c2{ 0, 0, 0 }; float t2 = 0.0f;
        std::vector<int> w2(warps);
        ComputeOneTextureColour(c2, t2, (int)e[iP].colour[0], w2, tpoint, rawnormal, ray, tk, weight, isect, shadowflag);
        resultColour = cmul(resultColour, curW) + cmul(c2, prevW);
        resultTransm = (float)(curW * resultTransm + prevW * t2);
    }
}

void Tracer::ComputeLightedTexture(Col& resultColour, float& resultTransm, int texture, V3 ipoint, V3 rawnormal, Ray& ray, Ticket& tk, float weight, Intersection& isect,
                                   const std::vector<int>& warps)   // trace.cpp:739-1179
{
    const pvgpu_object& ob = S.objects[isect.Object];
    const double relativeIor = relative_ior(ray, ob.interior);
    struct WNRX { double weight; V3 normal; Col reflec; float reflex; int finish; };
    std::vector<WNRX> listWNRX;
    resultColour = Col{ 0, 0, 0 }; resultTransm = 0.0f;
    Col filCol{ 1, 1, 1 };
    double trans = 1.0;
    bool one_colour_found = false;
    V3 topNormal = rawnormal;
    int layer_number = 0;
    std::vector<std::pair<bool, Col>> light_cache(S.lights.size(), std::make_pair(false, Col{ 0, 0, 0 }));
    for (int layer = texture; (layer >= 0) && (trans > tk.adcBailout); layer_number++, layer = S.textures[layer].next) {
        const pvgpu_finish& fn = S.finishes[S.textures[layer].finish];
        V3 layNormal = rawnormal;
        if ((S.g.quality_flags & PVGPU_Q_NORMALS) && S.textures[layer].tnormal >= 0) {                    // trace.cpp:814-828
            layNormal = Warp_Normal_Chain(layNormal, warps, false);
            layNormal = Perturb_Normal(layNormal, S.textures[layer].tnormal, ipoint);
            if (S.tnormals[S.textures[layer].tnormal].flags & PVGPU_DONT_SCALE_BUMPS_FLAG) layNormal = unit(layNormal);
            layNormal = Warp_Normal_Chain(layNormal, warps, true);
        }
        if (layer_number == 0) topNormal = layNormal;
        double new_Weight = weight * trans;
        float lc[5];
        const bool colour_found = Compute_Pigment(lc, S.textures[layer].pigment, ipoint);
        one_colour_found = one_colour_found || colour_found;                                              // trace.cpp:838-841
        Col layCol{ lc[0], lc[1], lc[2] };
        listWNRX.push_back(WNRX{ new_Weight, layNormal, Col{ 0, 0, 0 }, fn.reflect_exp, S.textures[layer].finish });
        double cos_Angle_Incidence = -dot(ray.Direction, layNormal);
        // ComputeReflectivity (trace.cpp:2627-2654)
        WNRX& W = listWNRX.back();
        if (!fn.reflection_fresnel) {
            double wmax = std::max(std::max(fn.reflection_max[0], fn.reflection_max[1]), fn.reflection_max[2]);
            double wmin = std::max(std::max(fn.reflection_min[0], fn.reflection_min[1]), fn.reflection_min[2]);
            W.weight = W.weight * std::max(wmax, wmin);
            double frac = (std::fabs(fn.reflection_falloff - 1.0) > EPSILON) ? std::pow(1.0 - cos_Angle_Incidence, (double)fn.reflection_falloff) : 1.0 - cos_Angle_Incidence;
            float* o3 = &W.reflec.r;
            for (int k = 0; k < 3; k++) {
                if (std::fabs(frac) < EPSILON) o3[k] = fn.reflection_min[k];
                else if (std::fabs(frac - 1.0) < EPSILON) o3[k] = fn.reflection_max[k];
                else o3[k] = (float)(frac * fn.reflection_max[k]) + (float)((1.0 - frac) * fn.reflection_min[k]);
            }
        } else {
            double f = FresnelR(cos_Angle_Incidence, relativeIor);
            float* o3 = &W.reflec.r;
            for (int k = 0; k < 3; k++) o3[k] = (float)(f * fn.reflection_max[k]) + (float)((1.0 - f) * fn.reflection_min[k]);
            W.weight = W.weight * std::max(std::max(W.reflec.r, W.reflec.g), W.reflec.b);
        }
        ComputeMetallic(W.reflec, fn.reflect_metallic, layCol, cos_Angle_Incidence);
        double att = (S.g.language_version < 370) ? (float)(1.0 - ((double)(lc[3] * std::max(std::max(lc[0], lc[1]), lc[2])) + (double)lc[4])) : (float)(1.0 - (double)lc[3] - (double)lc[4]);
        if (fn.alpha_knockout) W.reflec = W.reflec * (float)att;
        Col tmpCol{ 0, 0, 0 };
        Col emission{ fn.emission[0] + fn.ambient[0] * S.g.ambient_light[0], fn.emission[1] + fn.ambient[1] * S.g.ambient_light[1], fn.emission[2] + fn.ambient[2] * S.g.ambient_light[2] };
        if (fn.fresnel != 0.0f) emission = emission * (float)(1.0 - (double)fn.fresnel * FresnelR(cos_Angle_Incidence, relativeIor));
        tmpCol = tmpCol + (layCol * emission) * (float)att;
        if (((fn.diffuse != 0.0f) || (fn.diffuse_back != 0.0f) || (fn.specular != 0.0f) || (fn.phong != 0.0f)) && ((!fn.alpha_knockout) || (att != 0.0))) {
            Col classic{ 0, 0, 0 };
            if (!(ob.flags & PVGPU_NO_GLOBAL_LIGHTS_FLAG))                                               // ComputeDiffuseLight trace.cpp:1488-1510
                for (size_t li = 0; li < S.lights.size(); li++)
                    ComputeOneDiffuseLight(S.lights[li], fn, isect.IPoint, ray, tk, layNormal, layCol, classic, att, ob, relativeIor, &light_cache[li]);
            tmpCol = tmpCol + classic;
        }
        tmpCol = tmpCol * filCol;
        resultColour = resultColour + tmpCol;
        if (colour_found) {                                                                               // trace.cpp:1066-1075
            Col tc{ lc[0] * lc[3] + lc[4], lc[1] * lc[3] + lc[4], lc[2] * lc[3] + lc[4] };
            filCol = filCol * tc;
            if (fn.conserve_energy != 0) filCol = filCol * Col{ std::min(1.0f - W.reflec.r, 1.0f), std::min(1.0f - W.reflec.g, 1.0f), std::min(1.0f - W.reflec.b, 1.0f) };
        }
        trans = std::min(1.0, (double)std::fabs(grey(filCol)));
    }
    bool tir_occured = false;
    if ((ob.interior >= 0) && (trans > tk.adcBailout) && (S.g.quality_flags & PVGPU_Q_REFRACTIONS)) {     // trace.cpp:1085-1145
        const pvgpu_interior& in = S.interiors[ob.interior];
        double w1 = std::max(std::max((double)std::fabs(filCol.r), (double)std::fabs(filCol.g)), (double)std::fabs(filCol.b));
        double new_Weight = weight * w1;
        Col rfrCol{ 0, 0, 0 }; float rfrTransm = 0.0f;
        tir_occured = ComputeRefraction(ob.interior, isect.IPoint, ray, tk, topNormal, rawnormal, rfrCol, rfrTransm, (float)new_Weight, S.textures[texture].finish);
        Col attCol{ in.old_refract, in.old_refract, in.old_refract };
        if (ray.IsInterior(ob.interior) && std::fabs(in.fade_distance) > EPSILON) {
            if (in.fade_power >= 1000) {
                double depth = isect.Depth / in.fade_distance;
                attCol = attCol * Col{ std::exp((float)(-(1.0 - in.fade_colour[0]) * depth)), std::exp((float)(-(1.0 - in.fade_colour[1]) * depth)), std::exp((float)(-(1.0 - in.fade_colour[2]) * depth)) };
            } else {
                double a = 1.0 + std::pow(isect.Depth / in.fade_distance, (double)in.fade_power);
                attCol = attCol * Col{ (float)(in.fade_colour[0] + (1.0 - in.fade_colour[0]) / a), (float)(in.fade_colour[1] + (1.0 - in.fade_colour[1]) / a), (float)(in.fade_colour[2] + (1.0 - in.fade_colour[2]) / a) };
            }
        }
        if (tir_occured) resultColour = resultColour + attCol * rfrCol;
        else if (one_colour_found) { resultColour = resultColour + attCol * rfrCol * filCol; resultTransm = grey(attCol) * rfrTransm * (float)trans; }
        else { resultColour = resultColour + attCol * rfrCol; resultTransm = grey(attCol) * rfrTransm; }  // trace.cpp:1137-1142
    }
    if (S.g.quality_flags & PVGPU_Q_REFLECTIONS) {                                                        // trace.cpp:1151-1178
        for (int i = 0; i < layer_number; i++) {
            const WNRX& W = listWNRX[i];
            if ((!tir_occured) || (std::fabs(topNormal.x - W.normal.x) > EPSILON) || (std::fabs(topNormal.y - W.normal.y) > EPSILON) || (std::fabs(topNormal.z - W.normal.z) > EPSILON)) {
                if (!(W.reflec.r == 0.0f && W.reflec.g == 0.0f && W.reflec.b == 0.0f)) {
                    Col rflCol{ 0, 0, 0 };
                    ComputeReflection(isect.IPoint, ray, tk, W.normal, rawnormal, rflCol, (float)W.weight, W.finish);
                    if (W.reflex != 1.0f) resultColour = resultColour + W.reflec * Col{ std::pow(rflCol.r, W.reflex), std::pow(rflCol.g, W.reflex), std::pow(rflCol.b, W.reflex) };
                    else resultColour = resultColour + W.reflec * rflCol;
                }
            }
        }
    }
}

void Tracer::ComputeIridColour(const pvgpu_finish& fn, V3 lightDirection, V3 eyeDirection, V3 layer_normal, V3 ipoint, Col& colour) const   // trace.cpp:2486-2518
{
    double film_thickness = fn.irid_film_thickness;
    if (fn.irid_turb != 0) {
        pvgpu_warp turb{};
        turb.omega = 0.5f; turb.lambda = 2.0f; turb.octaves = 5;
        double noise = Turbulence(S, ipoint, turb, S.g.noise_generator);
        noise = 2.0 * noise - 1.0;
        noise = 1.0 + noise * fn.irid_turb;
        film_thickness *= noise;
    }
    double cl = std::fabs(dot(layer_normal, lightDirection)), ce = std::fabs(dot(layer_normal, eyeDirection));
    double interference = 2.0 * 3.1415926535897932384626 * film_thickness * (cl + ce);
    const float w[3] = { S.irid_wavelengths.size() == 3 ? S.irid_wavelengths[0] : 1.0f, S.irid_wavelengths.size() == 3 ? S.irid_wavelengths[1] : 1.0f, S.irid_wavelengths.size() == 3 ? S.irid_wavelengths[2] : 1.0f };
    auto factor = [&](float wl) { float q = (float)interference / wl; float cs = (float)std::cos((double)q); return (float)((double)(float)((double)cs * (double)fn.irid) + 1.0); };
    colour = Col{ colour.r * factor(w[0]), colour.g * factor(w[1]), colour.b * factor(w[2]) };
}

void Tracer::ComputeReflection(V3 ipoint, Ray& ray, Ticket& tk, V3 normal, V3 rawnormal, Col& colour, float weight, int finish)           // trace.cpp:1264-1321
{
    Ray nray(ray);
    nray.flags = RAY_REFLECTION | (ray.flags & RAY_REFRACTION);
    double n = -2.0 * dot(ray.Direction, normal);
    nray.Direction = ray.Direction + n * normal;
    n = dot(nray.Direction, rawnormal);
    if (n < 0.0) {
        double n2 = dot(nray.Direction, normal);
        if (n2 < 0.0) { n = -2.0 * dot(ray.Direction, rawnormal); nray.Direction = ray.Direction + n * rawnormal; }
        else { n *= -2.0; nray.Direction = nray.Direction + n * rawnormal; }
    }
    nray.Direction = unit(nray.Direction);
    nray.Origin = ipoint;
    bool alphaBackground = tk.alphaBackground;
    tk.alphaBackground = false;
    float dummyTransm = 0.0f;
    Col c{ 0, 0, 0 };
    TraceRay(nray, tk, c, dummyTransm, weight, false);
    if (finish >= 0 && S.finishes[finish].irid > 0.0f) ComputeIridColour(S.finishes[finish], nray.Direction, ray.Direction, normal, ipoint, c);   // trace.cpp:1306-1314
    colour = colour + c;
    tk.alphaBackground = alphaBackground;
}

bool Tracer::ComputeRefraction(int interior, V3 ipoint, Ray& ray, Ticket& tk, V3 normal, V3 rawnormal, Col& colour, float& transm, float weight, int finish)   // trace.cpp:1323-1485
{
    const pvgpu_interior& in = S.interiors[interior];
    Ray nray(ray);
    nray.flags = RAY_REFRACTION | (ray.flags & RAY_REFLECTION);
    nray.Origin = ipoint;
    double ior;
    const double atmosphereIOR = S.g.atmosphere_ior;      // DBL in the reference (scenedata.h:103); Interior::IOR is SNGL
    if (nray.interiors.empty()) { nray.interiors.push_back(interior); ior = atmosphereIOR / in.ior; }
    else if (interior == nray.interiors.back()) {
        nray.RemoveInterior(interior);
        if (nray.interiors.empty()) ior = in.ior / atmosphereIOR;
        else ior = in.ior / S.interiors[nray.interiors.back()].ior;
    } else if (nray.RemoveInterior(interior)) ior = 1.0;
    else { ior = S.interiors[nray.interiors.back()].ior / in.ior; nray.interiors.push_back(interior); }
    if (std::fabs(ior - 1.0) < EPSILON) {
        nray.Direction = ray.Direction;
        colour = Col{ 0, 0, 0 }; transm = 0.0f;
        TraceRay(nray, tk, colour, transm, weight, true);
        return false;
    }
    double n = dot(ray.Direction, normal);
    V3 localnormal;
    if (n <= 0.0) { localnormal = normal; n = -n; } else localnormal = -normal;
    double t = 1.0 + sqr(ior) * (sqr(n) - 1.0);                                                           // TraceRefractionRay trace.cpp:1456-1485
    if (t < 0.0) {
        Col tempcolour{ 0, 0, 0 };
        ComputeReflection(ipoint, ray, tk, normal, rawnormal, tempcolour, weight, finish);
        colour = colour + tempcolour;
        return true;
    }
    t = ior * n - std::sqrt(t);
    nray.Direction = ior * ray.Direction + t * localnormal;
    colour = Col{ 0, 0, 0 }; transm = 0.0f;
    TraceRay(nray, tk, colour, transm, weight, false);
    return false;
}

void Tracer::ComputeOneWhiteLightRay(const pvgpu_light& L, double& depth, Ray& lray, V3 ipoint, V3 jitter) const             // trace.cpp:2710-2767
{
    V3 center = v3(L.center) + jitter;
    lray.Origin = ipoint;
    if (L.type == PVGPU_LIGHT_CYLINDER) {
        lray.Direction = center - v3(L.points_at);
        V3 toLightCtr = center - ipoint;
        double distToPointsAt = len(lray.Direction);
        depth = dot(toLightCtr, lray.Direction);
        depth /= distToPointsAt;
        lray.Direction = unit(lray.Direction);
    } else { lray.Direction = center - ipoint; depth = len(lray.Direction); lray.Direction = lray.Direction / depth; }
    if (L.flags & PVGPU_LIGHT_PARALLEL) {
        if (L.flags & PVGPU_LIGHT_AREA) {
            V3 v1 = unit(center - v3(L.points_at));
            double a = dot(v1, lray.Direction);
            depth *= a;
            lray.Direction = v1;
        } else { double a = dot(v3(L.direction), lray.Direction); depth *= (-a); lray.Direction = -v3(L.direction); }
    }
}

void Tracer::ComputeOneDiffuseLight(const pvgpu_light& L, const pvgpu_finish& fn, V3 ipoint, const Ray& eye, Ticket& tk, V3 layer_normal,
                                    Col pig, Col& colour, double attenuation, const pvgpu_object& object, double relativeIor,
                                    std::pair<bool, Col>* light_cache)     // trace.cpp:1637-1728
{
    Ray lray(eye);
    double depth;
    ComputeOneWhiteLightRay(L, depth, lray, ipoint, v3(0.0, 0.0, 0.0));
    double latt = Attenuate_Light(L, lray, depth);
    Col lightcolour{ (float)(L.colour[0] * latt), (float)(L.colour[1] * latt), (float)(L.colour[2] * latt) };
    if (near_zero(lightcolour, (float)EPSILON)) return;
    bool backside = false;
    if (!(object.flags & PVGPU_DOUBLE_ILLUMINATE_FLAG)) {
        double cos_shadow_angle = dot(layer_normal, lray.Direction);
        if (cos_shadow_angle < EPSILON) { if (fn.diffuse_back != 0.0f) backside = true; else return; }
    }
    if ((S.g.quality_flags & PVGPU_Q_SHADOWS) && (L.type != PVGPU_LIGHT_FILL)) {
        // lightColorCache (trace.cpp:1671-1682): one shadow test per light for all layers of a layered texture
        if (!light_cache->first) { TraceShadowRay(L, depth, lray, tk, lightcolour); light_cache->first = true; light_cache->second = lightcolour; }
        else lightcolour = light_cache->second;
    }
    Col tmpCol{ 0, 0, 0 };
    if (!near_zero(lightcolour, (float)EPSILON)) {
        // ComputeDiffuseColour (trace.cpp:2441-2484)
        double diffuse = (double)((backside ? fn.diffuse_back : fn.diffuse) * fn.brilliance_adjust);
        if (diffuse > 0.0) {
            double cai = dot(layer_normal, lray.Direction);
            double intensity = (fn.brilliance != 1.0f) ? std::pow(std::fabs(cai), (double)fn.brilliance) : std::fabs(cai);
            intensity *= diffuse * attenuation;
            double ff = 1.0;
            if (fn.fresnel != 0.0f) {
                double f1 = fn.fresnel * FresnelR(cai, relativeIor), f2 = fn.fresnel * FresnelR(-dot(layer_normal, eye.Direction), relativeIor);
                ff = (1.0 - f1) * (1.0 - f2);
            }
            tmpCol = tmpCol + (pig * lightcolour) * (float)(intensity * ff);
        }
        Col tempLight = fn.alpha_knockout ? lightcolour * (float)attenuation : lightcolour;
        if ((L.type != PVGPU_LIGHT_FILL) && !backside) {
            if (fn.phong > 0.0f) {                                                                        // ComputePhongColour trace.cpp:2518-2555
                double c = -2.0 * dot(eye.Direction, layer_normal);
                V3 rd = eye.Direction + c * layer_normal;
                c = dot(rd, lray.Direction);
                if (c > 0.0 && ((fn.phong_size < 60) || (c > 0.0008))) {
                    double intensity = fn.phong * std::pow(c, (double)fn.phong_size);
                    Col cs{ 1, 1, 1 };
                    if ((fn.fresnel != 0.0f) || (fn.metallic != 0.0f)) {
                        double ndotl = dot(layer_normal, lray.Direction);
                        if (fn.fresnel != 0.0f) cs = cs * (float)(fn.fresnel * FresnelR(ndotl, relativeIor));
                        ComputeMetallic(cs, fn.metallic, pig, ndotl);
                    }
                    tmpCol = tmpCol + (tempLight * cs) * (float)intensity;
                }
            }
            if (fn.specular > 0.0f) {                                                                     // ComputeSpecularColour trace.cpp:2557-2593
                V3 halfway = ((-eye.Direction) + lray.Direction) * 0.5;
                double hl = len(halfway);
                if (hl > 0.0) {
                    double c = dot(halfway, layer_normal) / hl;
                    if (c > 0.0) {
                        double intensity = fn.specular * std::pow(c, (double)fn.roughness);
                        Col cs{ 1, 1, 1 };
                        if ((fn.fresnel != 0.0f) || (fn.metallic != 0.0f)) {
                            double ndotl = dot(halfway, lray.Direction) / hl;
                            if (fn.fresnel != 0.0f) cs = cs * (float)(fn.fresnel * FresnelR(ndotl, relativeIor));
                            ComputeMetallic(cs, fn.metallic, pig, ndotl);
                        }
                        tmpCol = tmpCol + (tempLight * cs) * (float)intensity;
                    }
                }
            }
        }
    }
    if (fn.irid > 0.0f) ComputeIridColour(fn, lray.Direction, eye.Direction, layer_normal, ipoint, tmpCol);        // trace.cpp:1723-1724
    colour = colour + tmpCol;
}

void Tracer::TraceShadowRay(const pvgpu_light& L, double depth, Ray& lightsourceray, Ticket& tk, Col& colour)                // trace.cpp:1892-1933
{
    if (tk.traceLevel > tk.maxAllowedTraceLevel) { colour = Col{ 0, 0, 0 }; return; }
    tk.maxFound = std::max(tk.maxFound, tk.traceLevel);
    tk.traceLevel++;
    Ray newray(lightsourceray);
    newray.flags = 0; newray.shadowTest = true;
    double newdepth = depth;
    if ((L.flags & PVGPU_LIGHT_AREA) && (S.g.quality_flags & PVGPU_Q_AREA_LIGHTS)) TraceAreaLightShadowRay(L, newdepth, newray, lightsourceray.Origin, tk, colour);
    else TracePointLightShadowRay(newdepth, newray, tk, colour);
    tk.traceLevel--;
    st.max_level = std::max(st.max_level, tk.maxFound);
}

void Tracer::TracePointLightShadowRay(double& lightsourcedepth, Ray& newray, Ticket& tk, Col& colour)                         // trace.cpp:1946-2076 (no caches)
{
    while (true) {
        Intersection bi;
        bi.Depth = lightsourcedepth;
        st.shadow_tests++;
        bool found = FindIntersection(bi, newray, SMALL_TOLERANCE);
        if (found && (bi.Depth < lightsourcedepth - SHADOW_TOLERANCE) && (lightsourcedepth - bi.Depth > 0.0) && (bi.Depth > SHADOW_TOLERANCE)) {
            ComputeShadowColour(bi, newray, tk, colour);
            int testObject = bi.Csg >= 0 ? bi.Csg : bi.Object;
            if (near_zero(colour, (float)EPSILON) && (S.objects[testObject].flags & PVGPU_OPAQUE_FLAG)) break;
            if (near_zero(colour, (float)EPSILON)) break;       // black already: further segments cannot change the result
            lightsourcedepth -= bi.Depth;
            newray.Origin = bi.IPoint;
        } else break;
    }
}

void Tracer::TraceAreaLightShadowRay(const pvgpu_light& L, double& lightsourcedepth, Ray& lightsourceray, V3 ipoint, Ticket& tk, Col& lightcolour)   // trace.cpp:2078-2131
{
    std::vector<Col> lightGrid((size_t)L.area_size1 * L.area_size2, Col{ std::nanf(""), 0.0f, 0.0f });       // Invalidate(): NaN in the first channel (colour.h:71-73)
    V3 axis1Temp = v3(L.axis1), axis2Temp = v3(L.axis2);
    if (L.flags & PVGPU_LIGHT_ORIENT) {
        ComputeOneWhiteLightRay(L, lightsourcedepth, lightsourceray, ipoint, v3(0.0, 0.0, 0.0));
        double axis1_Length = len(axis1Temp);
        V3 temp = (std::fabs(std::fabs(lightsourceray.Direction.z) - 1.0) < 0.01) ? v3(0.0, 1.0, 0.0) : v3(0.0, 0.0, 1.0);
        axis1Temp = unit(cross(lightsourceray.Direction, temp));
        axis2Temp = unit(cross(lightsourceray.Direction, axis1Temp));
        axis1Temp = axis1Temp * axis1_Length;
        axis2Temp = axis2Temp * axis1_Length;
    }
    TraceAreaLightSubsetShadowRay(L, lightsourcedepth, lightsourceray, ipoint, tk, lightcolour, 0, 0, L.area_size1 - 1, L.area_size2 - 1, 0, axis1Temp, axis2Temp, lightGrid);
}

void Tracer::TraceAreaLightSubsetShadowRay(const pvgpu_light& L, double& lightsourcedepth, Ray& lightsourceray, V3 ipoint, Ticket& tk, Col& lightcolour,
                                           int u1, int v1, int u2, int v2, int level, V3 axis1, V3 axis2, std::vector<Col>& lightGrid)   // trace.cpp:2133-2271
{
    Col sample_Colour[4];
    auto ColourDistance = [](Col a, Col b) { return std::fabs(a.r - b.r) + std::fabs(a.g - b.g) + std::fabs(a.b - b.b); };
    for (int i = 0; i < 4; i++) {
        Ray lsr(lightsourceray);
        const int u = (i == 1 || i == 3) ? u2 : u1, v = (i >= 2) ? v2 : v1;
        Col& cell = lightGrid[(size_t)u * L.area_size2 + v];
        if (!std::isnan(cell.r)) sample_Colour[i] = cell;
        else {
            V3 jitterAxis1, jitterAxis2;
            double jitter_u = (double)u, jitter_v = (double)v, scaleFactor;
            // (jitter draws from the thread's random number generator: not reproducible, rejected at scene validation)
            if (L.flags & PVGPU_LIGHT_CIRCULAR) {
                jitter_u = jitter_u / (L.area_size1 - 1) - 0.5 + 0.001;
                jitter_v = jitter_v / (L.area_size2 - 1) - 0.5 + 0.001;
                scaleFactor = ((std::fabs(jitter_u) > std::fabs(jitter_v)) ? std::fabs(jitter_u) : std::fabs(jitter_v));
                scaleFactor /= std::sqrt(jitter_u * jitter_u + jitter_v * jitter_v);
                jitter_u *= scaleFactor;
                jitter_v *= scaleFactor;
                jitterAxis1 = axis1 * jitter_u;
                jitterAxis2 = axis2 * jitter_v;
            } else {
                if (L.area_size1 > 1) { scaleFactor = jitter_u / (double)(L.area_size1 - 1) - 0.5; jitterAxis1 = axis1 * scaleFactor; }
                else jitterAxis1 = v3(0.0, 0.0, 0.0);
                if (L.area_size2 > 1) { scaleFactor = jitter_v / (double)(L.area_size2 - 1) - 0.5; jitterAxis2 = axis2 * scaleFactor; }
                else jitterAxis2 = v3(0.0, 0.0, 0.0);
            }
            ComputeOneWhiteLightRay(L, lightsourcedepth, lsr, ipoint, jitterAxis1 + jitterAxis2);
            sample_Colour[i] = lightcolour;
            TracePointLightShadowRay(lightsourcedepth, lsr, tk, sample_Colour[i]);
            cell = sample_Colour[i];
        }
    }
    if ((u2 - u1 > 1) || (v2 - v1 > 1)) {
        if ((level < L.adaptive_level) || (ColourDistance(sample_Colour[0], sample_Colour[1]) > 0.1) || (ColourDistance(sample_Colour[1], sample_Colour[3]) > 0.1) ||
            (ColourDistance(sample_Colour[3], sample_Colour[2]) > 0.1) || (ColourDistance(sample_Colour[2], sample_Colour[0]) > 0.1)) {
            for (int i = 0; i < 4; i++) {
                int new_u1, new_v1, new_u2, new_v2;
                switch (i) {
                    case 0: new_u1 = u1; new_v1 = v1; new_u2 = (int)std::floor((u1 + u2) / 2.0); new_v2 = (int)std::floor((v1 + v2) / 2.0); break;
                    case 1: new_u1 = (int)std::ceil((u1 + u2) / 2.0); new_v1 = v1; new_u2 = u2; new_v2 = (int)std::floor((v1 + v2) / 2.0); break;
                    case 2: new_u1 = u1; new_v1 = (int)std::ceil((v1 + v2) / 2.0); new_u2 = (int)std::floor((u1 + u2) / 2.0); new_v2 = v2; break;
                    default: new_u1 = (int)std::ceil((u1 + u2) / 2.0); new_v1 = (int)std::ceil((v1 + v2) / 2.0); new_u2 = u2; new_v2 = v2; break;
                }
                sample_Colour[i] = lightcolour;
                TraceAreaLightSubsetShadowRay(L, lightsourcedepth, lightsourceray, ipoint, tk, sample_Colour[i], new_u1, new_v1, new_u2, new_v2, level + 1, axis1, axis2, lightGrid);
            }
        }
    }
    lightcolour = (((sample_Colour[0] + sample_Colour[1]) + sample_Colour[2]) + sample_Colour[3]) * 0.25f;
}

void Tracer::ComputeShadowTexture(Col& filtercolour, int tex, const std::vector<int>& warps, V3 ipoint, V3 raw, Ray& lray, Intersection& isect)   // trace.cpp:1181-1262
{
    const pvgpu_object& ob = S.objects[isect.Object];
    Col tmpCol{ 1, 1, 1 };
    const pvgpu_interior* in = ob.interior >= 0 ? &S.interiors[ob.interior] : nullptr;
    for (int layer = tex; layer >= 0; layer = S.textures[layer].next) {
        float lc[5];
        if (Compute_Pigment(lc, S.textures[layer].pigment, ipoint))                                       // trace.cpp:1198-1205
            tmpCol = tmpCol * Col{ lc[0] * lc[3] + lc[4], lc[1] * lc[3] + lc[4], lc[2] * lc[3] + lc[4] };
        if (in && in->caustics != 0.0f) {                                                                 // trace.cpp:1208-1234
            V3 layer_Normal = raw;
            if ((S.g.quality_flags & PVGPU_Q_NORMALS) && S.textures[layer].tnormal >= 0) {
                layer_Normal = Warp_Normal_Chain(layer_Normal, warps, false);
                layer_Normal = Perturb_Normal(layer_Normal, S.textures[layer].tnormal, ipoint);
                if (S.tnormals[S.textures[layer].tnormal].flags & PVGPU_DONT_SCALE_BUMPS_FLAG) layer_Normal = unit(layer_Normal);
                layer_Normal = Warp_Normal_Chain(layer_Normal, warps, true);
            }
            double k = 1.0 + std::pow(std::fabs(dot(layer_Normal, lray.Direction)), (double)in->caustics);
            tmpCol = tmpCol * (float)k;
        }
    }
    Col refraction{ 1, 1, 1 };
    if (in && lray.IsInterior(ob.interior) && (in->fade_power > 0.0f) && (std::fabs(in->fade_distance) > EPSILON)) {
        if (in->fade_power >= 1000) {
            double dd = isect.Depth / in->fade_distance;
            refraction = refraction * Col{ std::exp((float)(-(1.0 - in->fade_colour[0]) * dd)), std::exp((float)(-(1.0 - in->fade_colour[1]) * dd)), std::exp((float)(-(1.0 - in->fade_colour[2]) * dd)) };
        } else {
            double k = 1.0 + std::pow(isect.Depth / in->fade_distance, (double)in->fade_power);
            refraction = refraction * Col{ (float)(in->fade_colour[0] + (1.0 - in->fade_colour[0]) / k), (float)(in->fade_colour[1] + (1.0 - in->fade_colour[1]) / k), (float)(in->fade_colour[2] + (1.0 - in->fade_colour[2]) / k) };
        }
    }
    filtercolour = tmpCol * refraction;
}

void Tracer::ComputeShadowColour(Intersection& isect, Ray& lray, const Ticket& tk, Col& colour)           // trace.cpp:2274-2439 + ComputeShadowTexture :1181-1262
{
    const pvgpu_object& ob = S.objects[isect.Object];
    if (!(S.g.quality_flags & PVGPU_Q_SHADOWS)) return;
    if (ob.flags & PVGPU_OPAQUE_FLAG) { colour = Col{ 0, 0, 0 }; return; }
    V3 raw = Normal(isect);
    if (ob.flags & PVGPU_INVERTED_FLAG) raw = -raw;
    double nd = dot(raw, lray.Direction);
    if (nd > 0.0) raw = -raw;
    const std::vector<std::pair<float, int>> wtextures = Determine_Textures(ob, isect, nd > 0.0);
    if (wtextures.empty() && !(ob.type == PVGPU_OBJ_BLOB && (ob.flags & PVGPU_MULTITEXTURE_FLAG))) return;
    // ComputeOneTextureColour(..., shadowflag = true) per weighted texture (trace.cpp:2399-2417)
    Col temp{ 0, 0, 0 };
    Ticket tk2 = tk;
    for (const auto& wt : wtextures) {
        if ((wt.first < tk.adcBailout) || wt.second < 0) continue;
        Col fc1{ 0, 0, 0 }; float dummy = 0.0f;
        std::vector<int> warps;
        cur_isect = &isect;
        const V3 ipoint = (S.objects[isect.Object].flags & PVGPU_UV_FLAG) ? UVCoord(isect) : isect.IPoint;   // trace.cpp:2351-2362
        ComputeOneTextureColour(fc1, dummy, wt.second, warps, ipoint, raw, lray, tk2, 0.0f, isect, true);
        temp = temp + fc1 * wt.first;
    }
    if (std::fabs((std::fabs(temp.r) + std::fabs(temp.g) + std::fabs(temp.b)) / 3.0f) < tk.adcBailout) { colour = Col{ 0, 0, 0 }; return; }
    colour = colour * temp;
    // ComputeShadowMedia (trace.cpp:3046-3071): toggle the blocker's interior on the light ray
    if (!near_zero(colour, (float)EPSILON) && ob.interior >= 0)
        if (lray.interiors.empty() || !lray.RemoveInterior(ob.interior)) lray.interiors.push_back(ob.interior);
}

// TracePixel::CreateCameraRay (tracepixel.cpp:341-391, 917-927)
bool camera_ray(const Scene& S, double x, double y, double width, double height, V3& o, V3& d)
{
    const pvgpu_camera& cam = S.cam;
    V3 loc = v3(cam.location), dir = v3(cam.direction), right = v3(cam.right), up = v3(cam.up);
    o = loc;
    if (cam.type <= PVGPU_CAMERA_ORTHOGRAPHIC) {
        double x0 = x / width - 0.5, y0 = 0.5 - y / height;
        if (cam.type == PVGPU_CAMERA_ORTHOGRAPHIC) { d = dir; o = (loc + x0 * right) + y0 * up; }
        else d = (dir + x0 * right) + y0 * up;
        d = unit(d);
        return true;
    }
    // SetupCamera (tracepixel.cpp:235-309)
    const double cameraLengthRight = len(right), cameraLengthUp = len(up);
    double aspectRatio;
    switch (cam.type) {
        case PVGPU_CAMERA_CYL_1: case PVGPU_CAMERA_CYL_3: aspectRatio = cameraLengthUp; break;
        case PVGPU_CAMERA_CYL_2: case PVGPU_CAMERA_CYL_4: aspectRatio = cameraLengthRight; break;
        case PVGPU_CAMERA_ULTRA_WIDE_ANGLE: aspectRatio = cameraLengthUp / cameraLengthRight; break;
        default: aspectRatio = cameraLengthRight / cameraLengthUp; break;
    }
    if (cam.type != PVGPU_CAMERA_PANORAMIC && cam.type != PVGPU_CAMERA_SPHERICAL) { right = unit(right); up = unit(up); dir = unit(dir); }
    const double Angle = S.camera_ext.size() == 3 ? S.camera_ext[0] : 0.0, H_Angle = S.camera_ext.size() == 3 ? S.camera_ext[1] : 0.0,
                 V_Angle = S.camera_ext.size() == 3 ? S.camera_ext[2] : 0.0;
    const double M_PI_ = 3.1415926535897932384626, M_PI_180_ = 0.01745329251994329576, M_PI_360_ = 0.00872664625997164788, M_PI_2_ = 1.57079632679489661923,
                 TWO_M_PI_ = 6.283185307179586476925286766560;
    double x0, y0, cx, sx, cy, sy, rad, phi, ty;
    switch (cam.type) {
        case PVGPU_CAMERA_FISHEYE:                                                                        // tracepixel.cpp:394-436
            x0 = 2.0 * (x / width - 0.5); y0 = 2.0 * (0.5 - y / height);
            x0 *= cameraLengthRight; y0 *= cameraLengthUp;
            rad = std::sqrt(x0 * x0 + y0 * y0);
            if (rad > 1.0) return false;
            if (rad == 0.0) phi = 0.0; else if (x0 < 0.0) phi = M_PI_ - std::asin(y0 / rad); else phi = std::asin(y0 / rad);
            x0 = phi; y0 = rad * Angle * M_PI_360_;
            cx = std::cos(x0); sx = std::sin(x0); cy = std::cos(y0); sy = std::sin(y0);
            d = ((cx * sy) * right + (sx * sy) * up) + cy * dir;
            break;
        case PVGPU_CAMERA_OMNIMAX:                                                                        // :438-492
            x0 = 2.0 * (x / width - 0.5); y0 = 2.0 * (0.5 - y / height);
            if (aspectRatio > 1.0) {
                if (aspectRatio > 1.283458) { x0 *= aspectRatio / 1.283458; y0 = (y0 - 1.0) / 1.283458 + 1.0; }
                else y0 = (y0 - 1.0) / aspectRatio + 1.0;
            } else y0 /= aspectRatio;
            rad = std::sqrt(x0 * x0 + y0 * y0);
            if (rad > 1.0) return false;
            if (rad == 0.0) phi = 0.0; else if (x0 < 0.0) phi = M_PI_ - std::asin(y0 / rad); else phi = std::asin(y0 / rad);
            x0 = phi;
            y0 = 1.411269 * rad - 0.09439 * rad * rad * rad + 0.25674 * rad * rad * rad * rad * rad;
            cx = std::cos(x0); sx = std::sin(x0); cy = std::cos(y0); sy = std::sin(y0);
            if (sx * sy < std::tan(135.0 * M_PI_180_) * cy) return false;
            d = ((cx * sy) * right + (sx * sy) * up) + cy * dir;
            break;
        case PVGPU_CAMERA_PANORAMIC:                                                                      // :494-526
            x0 = x / width; y0 = 2.0 * (0.5 - y / height);
            x0 = (1.0 - x0) * M_PI_; y0 = M_PI_2_ * y0;
            cx = std::cos(x0); sx = std::sin(x0);
            if (std::fabs(M_PI_2_ - std::fabs(y0)) < EPSILON) ty = (y0 > 0.0) ? BOUND_HUGE : -BOUND_HUGE; else ty = std::tan(y0);
            d = (cx * right + ty * up) + sx * dir;
            break;
        case PVGPU_CAMERA_ULTRA_WIDE_ANGLE:                                                               // :528-551
            x0 = x / width - 0.5; y0 = 0.5 - y / height;
            x0 *= Angle * M_PI_180_; y0 *= Angle * aspectRatio * M_PI_180_;
            cx = std::cos(x0); sx = std::sin(x0); cy = std::cos(y0); sy = std::sin(y0);
            d = (sx * right + sy * up) + (cx * cy) * dir;
            break;
        case PVGPU_CAMERA_CYL_1: case PVGPU_CAMERA_CYL_3:                                                 // :553-574, 598-621
            x0 = x / width - 0.5; y0 = 0.5 - y / height;
            x0 *= Angle * M_PI_180_; y0 *= aspectRatio;
            cx = std::cos(x0); sx = std::sin(x0);
            if (cam.type == PVGPU_CAMERA_CYL_1) d = (sx * right + y0 * up) + cx * dir;
            else { d = sx * right + cx * dir; o = loc + y0 * up; }
            break;
        case PVGPU_CAMERA_CYL_2: case PVGPU_CAMERA_CYL_4:                                                 // :576-596, 623-646
            x0 = x / width - 0.5; y0 = 0.5 - y / height;
            y0 *= Angle * M_PI_180_; x0 *= aspectRatio;
            cy = std::cos(y0); sy = std::sin(y0);
            if (cam.type == PVGPU_CAMERA_CYL_2) d = (x0 * right + sy * up) + cy * dir;
            else { d = sy * up + cy * dir; o = loc + x0 * right; }
            break;
        default: {                                                                                        // spherical :648-673
            x0 = x / width - 0.5; y0 = 0.5 - y / height;
            y0 *= (V_Angle / 360) * TWO_M_PI_; x0 *= (H_Angle / 360) * TWO_M_PI_;
            auto rotate = [](V3 axis, double angle, V3 p) {                                               // matrix.cpp:825-850 + MTransPoint
                V3 a = unit(axis);
                double cosx = std::cos(angle), sinx = std::sin(angle);
                double m00 = a.x * a.x + cosx * (1.0 - a.x * a.x), m01 = a.x * a.y * (1.0 - cosx) + a.z * sinx, m02 = a.x * a.z * (1.0 - cosx) - a.y * sinx;
                double m10 = a.x * a.y * (1.0 - cosx) - a.z * sinx, m11 = a.y * a.y + cosx * (1.0 - a.y * a.y), m12 = a.y * a.z * (1.0 - cosx) + a.x * sinx;
                double m20 = a.x * a.z * (1.0 - cosx) + a.y * sinx, m21 = a.y * a.z * (1.0 - cosx) - a.x * sinx, m22 = a.z * a.z + cosx * (1.0 - a.z * a.z);
                return v3(p.x * m00 + p.y * m10 + p.z * m20 + 0.0, p.x * m01 + p.y * m11 + p.z * m21 + 0.0, p.x * m02 + p.y * m12 + p.z * m22 + 0.0);
            };
            V3 V1 = rotate(right, -y0, dir);
            d = rotate(up, x0, V1);
            break;
        }
    }
    d = unit(d);
    return true;
}

// TracePixel::InitRayContainerState (tracepixel.cpp:929-1006)
void container_state(const Scene& S, const Tracer& T, V3 p, std::vector<int>& out)
{
    out.clear();
    auto inside_bbox = [&](const float* lo, const float* size) {
        if (p.x < (double)lo[0] || p.y < (double)lo[1] || p.z < (double)lo[2]) return false;
        if (p.x > (double)lo[0] + (double)size[0] || p.y > (double)lo[1] + (double)size[1] || p.z > (double)lo[2] + (double)size[2]) return false;
        return true;
    };
    if (!S.use_tree) {
        for (uint32_t f : S.frame) { const pvgpu_object& o = S.objects[f]; if (o.interior >= 0 && inside_bbox(o.bbox, o.bbox + 3) && T.Inside(p, f)) out.push_back(o.interior); }
        return;
    }
    std::vector<uint32_t> st{ 0 };
    while (!st.empty()) {
        uint32_t ni = st.back(); st.pop_back();
        const pvgpu_node& n = S.nodes[ni];
        if (!inside_bbox(n.lo, n.size)) continue;
        if (n.count == 0) { const pvgpu_object& o = S.objects[n.first]; if (o.interior >= 0 && T.Inside(p, n.first)) out.push_back(o.interior); }
        else for (uint32_t c = n.count; c-- > 0;) st.push_back(n.first + c);
    }
}

template <class T> bool get(FILE* f, std::vector<T>& v)
{
    uint64_t n = 0;
    if (fread(&n, sizeof n, 1, f) != 1 || n > (1ull << 34) / sizeof(T)) return false;
    v.resize(n);
    return n == 0 || fread(v.data(), sizeof(T), n, f) == n;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------
// C interface (ctypes)
// ---------------------------------------------------------------------------------------------------------
extern "C" {

void* pvo_scene_load(const char* path)
{
    FILE* f = fopen(path, "rb");
    if (!f) return nullptr;
    Scene* s = new Scene();
    char magic[8]; uint32_t ver = 0, have_cam = 0;
    bool ok = fread(magic, 8, 1, f) == 1 && memcmp(magic, "PVGPUSC1", 8) == 0 && fread(&ver, 4, 1, f) == 1 && fread(&have_cam, 4, 1, f) == 1 &&
              fread(&s->g, sizeof s->g, 1, f) == 1 && fread(&s->cam, sizeof s->cam, 1, f) == 1 &&
              get(f, s->objects) && get(f, s->index_list) && get(f, s->frame) && get(f, s->xf) && get(f, s->nodes) && get(f, s->meshes) &&
              get(f, s->verts) && get(f, s->norms) && get(f, s->tris) && get(f, s->mnodes) && get(f, s->lights) && get(f, s->textures) &&
              get(f, s->pigments) && get(f, s->finishes) && get(f, s->maps) && get(f, s->entries) && get(f, s->warps) && get(f, s->interiors);
    if (ok) { int c = fgetc(f); if (c != EOF) { ungetc(c, f); ok = get(f, s->blobs) && get(f, s->blob_elements) && get(f, s->blob_nodes); } }
    if (ok) { int c = fgetc(f); if (c != EOF) { ungetc(c, f); ok = get(f, s->shape_data); } }
    if (ok) { int c = fgetc(f); if (c != EOF) { ungetc(c, f); ok = get(f, s->tnormals) && get(f, s->slope_entries); } }
    if (ok) { int c = fgetc(f); if (c != EOF) { ungetc(c, f); ok = get(f, s->sky_spheres) && get(f, s->fogs); } }
    if (ok) { int c = fgetc(f); if (c != EOF) { ungetc(c, f); ok = get(f, s->camera_ext); } }
    if (ok) { int c = fgetc(f); if (c != EOF) { ungetc(c, f); ok = get(f, s->irid_wavelengths); } }
    if (ok) { int c = fgetc(f); if (c != EOF) { ungetc(c, f); ok = get(f, s->blob_textures); } }
    if (ok) { int c = fgetc(f); if (c != EOF) { ungetc(c, f); ok = get(f, s->images) && get(f, s->texels); } }
    if (ok) { int c = fgetc(f); if (c != EOF) { ungetc(c, f); ok = get(f, s->mesh_uv) && get(f, s->tri_uv); } }
    fclose(f);
    if (!ok) { delete s; return nullptr; }
    s->use_tree = (s->g.bounding_method == 1 && !s->nodes.empty());
    init_noise(*s);
    { std::mt19937 gen; s->patternRands.resize(32768); for (double& v : s->patternRands) v = gen() / 4294967296.0; }     // RandomDoubles (randomsequence.cpp:138-149)
    // Initialize_Waves (noise.cpp:189-205)
    for (int i = 0, next_rand = -560851967; i < (int)s->g.number_of_waves; i++) {
        s->waveSources.push_back(unit(DNoise(*s, v3((double)i, 0.0, 0.0))));
        next_rand = (int)((long long)next_rand * 1812433253LL + 12345LL);
        s->waveFrequencies.push_back((double((int)(next_rand >> 16) & 0x7FFF) * 0.000030518509476) + 0.01);
    }
    return s;
}

void pvo_scene_destroy(void* sc) { delete reinterpret_cast<Scene*>(sc); }

// Renders pixel centres of the rectangle [left..right] x [top..bottom] (inclusive) of a width x height image,
// row-major RGBT, with `threads` worker threads pulling rows.  stats[0] = rays, stats[1] = shadow ray tests,
// stats[2] = highest trace level.
int pvo_render(void* sc, int width, int height, int left, int top, int right, int bottom, float* rgbt, int threads, unsigned long long* stats)
{
    const Scene& S = *reinterpret_cast<Scene*>(sc);
    if (threads < 1) threads = 1;
    std::atomic<int> next_row(top);
    std::vector<Stats> tstats(threads);
    auto worker = [&](int ti) {
        Tracer T(S);
        std::vector<int> cam_interiors;
        if (!(S.cam.type == PVGPU_CAMERA_ORTHOGRAPHIC || S.cam.type == PVGPU_CAMERA_CYL_3 || S.cam.type == PVGPU_CAMERA_CYL_4)) container_state(S, T, v3(S.cam.location), cam_interiors);
        const int w = right - left + 1;
        for (;;) {
            int y = next_row.fetch_add(1);
            if (y > bottom) break;
            for (int x = left; x <= right; x++) {
                Ticket tk; tk.maxAllowedTraceLevel = S.g.max_trace_level; tk.adcBailout = S.g.adc_bailout; tk.alphaBackground = S.g.output_alpha != 0;
                Ray ray;
                Col col{ 0, 0, 0 }; float transm = 0.0f;
                if (camera_ray(S, x + 0.5, y + 0.5, width, height, ray.Origin, ray.Direction)) {
                    T.camera_normal(x + 0.5, y + 0.5, width, height, ray.Direction);
                    if (S.cam.type == PVGPU_CAMERA_ORTHOGRAPHIC || S.cam.type == PVGPU_CAMERA_CYL_3 || S.cam.type == PVGPU_CAMERA_CYL_4) container_state(S, T, ray.Origin, cam_interiors);
                    ray.interiors = cam_interiors;
                    T.TraceRay(ray, tk, col, transm, 1.0f, false, S.cam.max_ray_distance);
                } else transm = 1.0f;                     // numTraced == 0 (tracepixel.cpp:332-335)
                float* o = rgbt + 4 * ((size_t)(y - top) * w + (x - left));
                o[0] = col.r; o[1] = col.g; o[2] = col.b; o[3] = transm;
            }
        }
        tstats[ti] = T.st;
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; t++) pool.emplace_back(worker, t);
    worker(0);
    for (auto& th : pool) th.join();
    if (stats) {
        stats[0] = stats[1] = stats[2] = 0;
        for (const Stats& s : tstats) { stats[0] += s.rays; stats[1] += s.shadow_tests; stats[2] = std::max<unsigned long long>(stats[2], s.max_level); }
    }
    return 0;
}

// ---- anti-aliasing (tracetask.cpp:521-657, 838-1074) -------------------------------------------------
static const float JitterTable[256] = {
#include "pv_jitter.inc"
};

// Jitter2d(DBL x, DBL y, DBL& jx, DBL& jy) (jitter.h:92-96); hashTable is the noise hash table
static void Jitter2d(const Scene& S, double x, double y, double& jx, double& jy)
{
    const unsigned short* h = S.hashTable.data();
    jx = JitterTable[int(h[int(h[(int(x * 1021.0) & 0xfff)] ^ int(y * 1019.0)) & 0xfff]) & 0xff];
    jy = JitterTable[int(h[int(h[(int(x * 1019.0) & 0xfff)] ^ int(y * 1021.0)) & 0xfff]) & 0xff];
}

struct Px { float r, g, b, t; };
static inline Px px_add(Px a, Px b) { return Px{ a.r + b.r, a.g + b.g, a.b + b.b, a.t + b.t }; }
static inline Px px_div(Px a, double d) { return Px{ (float)(a.r / d), (float)(a.g / d), (float)(a.b / d), (float)(a.t / d) }; }   // colour.h:531-536,1167
// GammaCurve::Encode(aaGamma, RGBTColour) with a power-law curve (colourspace.h:163-169, colourspace.cpp:306-313): transm is not encoded
static inline Px px_encode(Px a, float enc_gamma, bool neutral)
{
    if (neutral) return a;
    return Px{ std::pow(std::max(a.r, 0.0f), enc_gamma), std::pow(std::max(a.g, 0.0f), enc_gamma), std::pow(std::max(a.b, 0.0f), enc_gamma), a.t };
}
// ColourDistanceRGBT (colour.h:616-621, 1221-1224)
static inline float px_dist(Px a, Px b) { return std::fabs(a.r - b.r) + std::fabs(a.g - b.g) + std::fabs(a.b - b.b) + std::fabs(a.t - b.t); }

struct AAParams { int method; int depth; double threshold; double jitter_scale; double gamma; };

struct AATracer {
    const Scene& S; Tracer& T; int width, height; std::vector<int>& cam_interiors;
    float enc_gamma; bool neutral; AAParams aa; double jitterScale;
    unsigned long long samples = 0;
    Px trace(double x, double y)                          // TracePixel::operator() (tracepixel.cpp:311-339)
    {
        Ticket tk; tk.maxAllowedTraceLevel = S.g.max_trace_level; tk.adcBailout = S.g.adc_bailout; tk.alphaBackground = S.g.output_alpha != 0;
        Ray ray;
        Col col{ 0, 0, 0 }; float transm = 0.0f;
        if (!camera_ray(S, x, y, width, height, ray.Origin, ray.Direction)) return Px{ 0.0f, 0.0f, 0.0f, 1.0f };
        T.camera_normal(x, y, width, height, ray.Direction);
        if (S.cam.type == PVGPU_CAMERA_ORTHOGRAPHIC || S.cam.type == PVGPU_CAMERA_CYL_3 || S.cam.type == PVGPU_CAMERA_CYL_4) container_state(S, T, ray.Origin, cam_interiors);
        ray.interiors = cam_interiors;
        T.TraceRay(ray, tk, col, transm, 1.0f, false, S.cam.max_ray_distance);
        return Px{ col.r, col.g, col.b, transm };
    }
    bool differs(Px a, Px b) const { return px_dist(px_encode(a, enc_gamma, neutral), px_encode(b, enc_gamma, neutral)) >= aa.threshold; }

    // SupersampleOnePixel (tracetask.cpp:860-890)
    void supersample(double x, double y, Px& col)
    {
        const double step = 1.0 / double(aa.depth), range = 0.5 - (step * 0.5);
        for (double yy = -range; yy <= (range + EPSILON); yy += step)
            for (double xx = -range; xx <= (range + EPSILON); xx += step) {
                Px t;
                if (jitterScale > 0.0) {
                    double rx, ry;
                    Jitter2d(S, x + xx, y + yy, rx, ry);
                    t = trace(x + 0.5 + xx + (rx * jitterScale), y + 0.5 + yy + (ry * jitterScale));
                } else t = trace(x + 0.5 + xx, y + 0.5 + yy);
                col = px_add(col, t);
                samples++;
            }
        col = px_div(col, double(aa.depth * aa.depth + 1));
    }

    // NonAdaptiveSupersamplingM1 for one rectangle (tracetask.cpp:521-602, 838-858)
    void method1(int left, int top, int right, int bottom, float* out)
    {
        const int w = right - left + 1, h = bottom - top + 1, W1 = w + 1;
        std::vector<Px> px((size_t)W1 * (h + 1));
        std::vector<char> flag((size_t)W1 * (h + 1), 0);
        auto at = [&](int x, int y) -> size_t { return (size_t)(y - (top - 1)) * W1 + (x - (left - 1)); };
        for (int x = left; x <= right; x++) { px[at(x, top - 1)] = trace(x + 0.5, top - 0.5); flag[at(x, top - 1)] = 1; }
        for (int y = top; y <= bottom; y++) {
            px[at(left - 1, y)] = trace(left - 0.5, y + 0.5); flag[at(left - 1, y)] = 1;
            for (int x = left; x <= right; x++) {
                Px& cur = px[at(x, y)]; Px& lft = px[at(x - 1, y)]; Px& tp = px[at(x, y - 1)];
                cur = trace(x + 0.5, y + 0.5);
                bool sampleleft = !flag[at(x - 1, y)], sampletop = !flag[at(x, y - 1)];
                const bool leftdiff = differs(lft, cur), topdiff = differs(tp, cur);
                sampleleft = sampleleft && leftdiff;
                sampletop = sampletop && topdiff;
                const bool samplecurrent = leftdiff || topdiff;
                if (sampleleft) { supersample(x - 1.0, y, lft); flag[at(x - 1, y)] = 1; }
                if (sampletop) { supersample(x, y - 1.0, tp); flag[at(x, y - 1)] = 1; }
                if (samplecurrent) { supersample(x, y, cur); flag[at(x, y)] = 1; }
            }
        }
        for (int y = top; y <= bottom; y++)
            for (int x = left; x <= right; x++) { Px p = px[at(x, y)]; float* o = out + 4 * ((size_t)(y - top) * w + (x - left)); o[0] = p.r; o[1] = p.g; o[2] = p.b; o[3] = p.t; }
    }

    // SubdivideOnePixel (tracetask.cpp:892-1074): buf is the (subsize+1)^2 sample buffer of one pixel
    struct SubBuf { int n; std::vector<Px> v; std::vector<char> have; };
    Px subdivide(double x, double y, double d, int bx, int by, int bstep, SubBuf& buf, int level)
    {
        auto idx = [&](int i, int j) { return (size_t)j * buf.n + i; };
        const Px c00 = buf.v[idx(bx, by)], c02 = buf.v[idx(bx, by + bstep)], c20 = buf.v[idx(bx + bstep, by)], c22 = buf.v[idx(bx + bstep, by + bstep)];
        const int half = bstep / 2;
        if ((level > 0) && (differs(c00, c02) || differs(c00, c20) || differs(c00, c22) || differs(c02, c20) || differs(c02, c22) || differs(c20, c22))) {
            auto need = [&](int i, int j, double ox, double oy) {
                if (buf.have[idx(i, j)]) return;
                Px t;
                if (jitterScale > 0.0) {
                    double rx, ry;
                    Jitter2d(S, x + ox, y + oy, rx, ry);
                    t = trace(x + 0.5 + ox + (rx * jitterScale), y + 0.5 + oy + (ry * jitterScale));
                } else t = trace(x + 0.5 + ox, y + 0.5 + oy);
                buf.v[idx(i, j)] = t; buf.have[idx(i, j)] = 1;
                samples++;
            };
            need(bx, by + half, -d, 0.0);
            need(bx + half, by, 0.0, -d);
            need(bx + bstep, by + half, d, 0.0);
            need(bx + half, by + bstep, 0.0, d);
            need(bx + half, by + half, 0.0, 0.0);
            const double d2 = d * 0.5;
            const Px r00 = subdivide(x - d2, y - d2, d2, bx, by, half, buf, level - 1);
            const Px r01 = subdivide(x - d2, y + d2, d2, bx, by + half, half, buf, level - 1);
            const Px r10 = subdivide(x + d2, y - d2, d2, bx + half, by, half, buf, level - 1);
            const Px r11 = subdivide(x + d2, y + d2, d2, bx + half, by + half, half, buf, level - 1);
            return px_div(px_add(px_add(px_add(r00, r01), r10), r11), 4.0);
        }
        return px_div(px_add(px_add(px_add(c00, c02), c20), c22), 4.0);
    }

    // AdaptiveSupersamplingM2 for one rectangle (tracetask.cpp:604-657)
    void method2(int left, int top, int right, int bottom, float* out)
    {
        const int w = right - left + 1, h = bottom - top + 1, W1 = w + 1, subsize = 1 << aa.depth;
        std::vector<Px> corner((size_t)W1 * (h + 1));
        for (int y = top; y <= bottom + 1; y++)
            for (int x = left; x <= right + 1; x++) corner[(size_t)(y - top) * W1 + (x - left)] = trace(x, y);
        SubBuf buf; buf.n = subsize + 1; buf.v.resize((size_t)buf.n * buf.n); buf.have.resize(buf.v.size());
        for (int y = top; y <= bottom; y++)
            for (int x = left; x <= right; x++) {
                std::fill(buf.have.begin(), buf.have.end(), 0);
                auto set = [&](int i, int j, Px p) { buf.v[(size_t)j * buf.n + i] = p; buf.have[(size_t)j * buf.n + i] = 1; };
                set(0, 0, corner[(size_t)(y - top) * W1 + (x - left)]);
                set(0, subsize, corner[(size_t)(y + 1 - top) * W1 + (x - left)]);
                set(subsize, 0, corner[(size_t)(y - top) * W1 + (x + 1 - left)]);
                set(subsize, subsize, corner[(size_t)(y + 1 - top) * W1 + (x + 1 - left)]);
                Px p = subdivide(double(x), double(y), 0.5, 0, 0, subsize, buf, aa.depth - 1);
                float* o = out + 4 * ((size_t)(y - top) * w + (x - left)); o[0] = p.r; o[1] = p.g; o[2] = p.b; o[3] = p.t;
            }
    }
};

// Anti-aliased render of a list of rectangles (left, top, right, bottom each), rect-major output like
// ViewData::CompletedRectangle.  method 1 = NonAdaptiveSupersamplingM1, 2 = AdaptiveSupersamplingM2, 0 = SimpleSamplingM0.
// jitter_scale is TraceTask's constructor argument (0 = no jitter); gamma = decoding gamma of the AA curve (<= 0 or 1: neutral).
// stats: [0] rays, [1] shadow tests, [2] max level, [3] samples (Number_Of_Samples).
int pvo_render_aa(void* sc, int width, int height, const int* rects, int n_rects, int method, int depth, double threshold,
                  double jitter_scale, double gamma, float* rgbt, int threads, unsigned long long* stats)
{
    const Scene& S = *reinterpret_cast<Scene*>(sc);
    if (threads < 1) threads = 1;
    std::vector<size_t> off(n_rects + 1, 0);
    for (int i = 0; i < n_rects; i++) off[i + 1] = off[i] + (size_t)(rects[4 * i + 2] - rects[4 * i] + 1) * (rects[4 * i + 3] - rects[4 * i + 1] + 1);
    std::atomic<int> next(0);
    std::vector<Stats> tstats(threads);
    std::vector<unsigned long long> tsamples(threads, 0);
    auto worker = [&](int ti) {
        Tracer T(S);
        std::vector<int> cam_interiors;
        if (!(S.cam.type == PVGPU_CAMERA_ORTHOGRAPHIC || S.cam.type == PVGPU_CAMERA_CYL_3 || S.cam.type == PVGPU_CAMERA_CYL_4)) container_state(S, T, v3(S.cam.location), cam_interiors);
        AATracer A{ S, T, width, height, cam_interiors, 1.0f, true, AAParams{ method, depth, threshold, jitter_scale, gamma }, 0.0 };
        if (gamma > 0.0 && gamma != 1.0) { A.enc_gamma = 1.0f / (float)gamma; A.neutral = false; }
        // jitterScale = jitterScale / aaDepth (M1, tracetask.cpp:526) or / ((1 << aaDepth) + 1) (M2, tracetask.cpp:611)
        A.jitterScale = (method == 1) ? jitter_scale / double(depth) : jitter_scale / double((1 << depth) + 1);
        for (;;) {
            int i = next.fetch_add(1);
            if (i >= n_rects) break;
            const int l = rects[4 * i], t = rects[4 * i + 1], r = rects[4 * i + 2], b = rects[4 * i + 3];
            float* out = rgbt + 4 * off[i];
            if (method == 1) A.method1(l, t, r, b, out);
            else if (method == 2) A.method2(l, t, r, b, out);
            else {
                const int w = r - l + 1;
                for (int y = t; y <= b; y++)
                    for (int x = l; x <= r; x++) { Px p = A.trace(x + 0.5, y + 0.5); float* o = out + 4 * ((size_t)(y - t) * w + (x - l)); o[0] = p.r; o[1] = p.g; o[2] = p.b; o[3] = p.t; }
            }
        }
        tstats[ti] = T.st;
        tsamples[ti] = A.samples;
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; t++) pool.emplace_back(worker, t);
    worker(0);
    for (auto& th : pool) th.join();
    if (stats) {
        stats[0] = stats[1] = stats[2] = stats[3] = 0;
        for (int t = 0; t < threads; t++) { stats[0] += tstats[t].rays; stats[1] += tstats[t].shadow_tests; stats[2] = std::max<unsigned long long>(stats[2], tstats[t].max_level); stats[3] += tsamples[t]; }
    }
    return 0;
}

// Trace::FindIntersection under primary-ray conditions for explicit rays (6 doubles each).
int pvo_trace_rays(void* sc, const double* org_dir, size_t n, int32_t* obj, double* depth, uint32_t* aux)
{
    const Scene& S = *reinterpret_cast<Scene*>(sc);
    Tracer T(S);
    for (size_t i = 0; i < n; i++) {
        Ray ray; ray.Origin = v3(org_dir + 6 * i); ray.Direction = v3(org_dir + 6 * i + 3);
        Intersection best;
        bool found = T.FindIntersection(best, ray, -1.0);
        obj[i] = found ? best.Object : -1; depth[i] = found ? best.Depth : BOUND_HUGE;
        if (aux) aux[i] = found ? best.aux : 0;
    }
    return 0;
}

int pvo_camera_rays(void* sc, int width, int height, const double* xy, size_t n, double* org_dir)
{
    const Scene& S = *reinterpret_cast<Scene*>(sc);
    for (size_t i = 0; i < n; i++) {
        V3 o, d;
        if (!camera_ray(S, xy[2 * i], xy[2 * i + 1], width, height, o, d)) { o = v3(0.0, 0.0, 0.0); d = v3(0.0, 0.0, 0.0); }
        else Tracer(S).camera_normal(xy[2 * i], xy[2 * i + 1], width, height, d);
        double* r = org_dir + 6 * i;
        r[0] = o.x; r[1] = o.y; r[2] = o.z; r[3] = d.x; r[4] = d.y; r[5] = d.z;
    }
    return 0;
}

// Solve_Polynomial / Noise / DNoise probes for unit tests
int pvo_solve_polynomial(int n, const double* c, double* r, int sturm, double epsilon) { return Solve_Polynomial(n, c, r, sturm, epsilon); }
double pvo_noise(void* sc, double x, double y, double z, int gen) { return Noise(*reinterpret_cast<Scene*>(sc), v3(x, y, z), gen); }
double pvo_turbulence(void* sc, double x, double y, double z, int gen, int octaves)
{
    pvgpu_warp w{}; w.octaves = octaves; w.lambda = 2.0f; w.omega = 0.5f;
    return Turbulence(*reinterpret_cast<Scene*>(sc), v3(x, y, z), w, gen);
}
void pvo_dnoise(void* sc, double x, double y, double z, double* out) { V3 r = DNoise(*reinterpret_cast<Scene*>(sc), v3(x, y, z)); out[0] = r.x; out[1] = r.y; out[2] = r.z; }

}  // extern "C"
