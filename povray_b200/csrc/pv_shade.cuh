// Surface shading in wavefront form: Trace::ComputeTextureColour / ComputeLightedTexture
// (source/core/render/trace.cpp:457-1179), the light loop (trace.cpp:1488-1728, 1872-1889,
// 2710-2767; lightsource.cpp:548-633), the finish models (trace.cpp:2441-2708), reflection and
// refraction ray set-up (trace.cpp:1264-1485, 2595-2625), ComputeSky (trace.cpp:2769-2890) and the
// pigment / pattern evaluation they call (pigment.cpp:395-545, pattern.cpp, warp.cpp:103-122).
//
// The reference recurses: a child's colour is multiplied by a factor and added to the parent's.
// Every such factor on this path is linear in the child colour, so a ray carries the product of the
// factors (`w`, `wt`) and adds its own local terms straight into the sample's accumulator.
#pragma once
#include "pv_shapes.cuh"
#if PV_HEAVY
#include "pv_blob.cuh"
#endif
#include "pv_noise.cuh"
#if PV_FULL_MATERIALS
#include "pv_image.cuh"
#endif
#include "pv_kernels.hpp"

namespace pvgpu {

__device__ __forceinline__ void accum_add(float4* accum, uint32_t sample, float r, float g, float b, float t)
{
    float* a = reinterpret_cast<float*>(accum + sample);
    if (r != 0.0f) atomicAdd(a + 0, r);
    if (g != 0.0f) atomicAdd(a + 1, g);
    if (b != 0.0f) atomicAdd(a + 2, b);
    if (t != 0.0f) atomicAdd(a + 3, t);
}

// ---- pattern evaluation --------------------------------------------------------------------------
// cycloidal (texture.cpp:98-110)
__device__ inline double cycloidal(double value)
{
    const double two_pi = 6.283185307179586476925286766560;
    if (value >= 0.0) return sin(((value - floor(value)) * 50000.0) / 50000.0 * two_pi);
    return 0.0 - sin(((0.0 - (value + floor(0.0 - value))) * 50000.0) / 50000.0 * two_pi);
}
// RGBColour::Greyscale (colour.h:232-237)
__device__ __forceinline__ float greyscale_rgb(const float* c) { return (float)(0.297 * c[0] + 0.589 * c[1] + 0.114 * c[2]); }

// Triangle_Wave (texture.cpp:128-150)
__device__ inline double triangle_wave(double value)
{
    double offset = (value >= 0.0) ? value - floor(value) : value + 1.0 + floor(fabs(value));
    return (offset >= 0.5) ? 2.0 * (1.0 - offset) : 2.0 * offset;
}

#if PV_FULL_MATERIALS
// The point-mapping warps: BlackHoleWarp, RepeatWarp, CubicWarp, CylindricalWarp, SphericalWarp, ToroidalWarp, PlanarWarp::WarpPoint
// (warp.cpp:124-545).  A warp that cannot map the point (a cylindrical warp on its axis ...) leaves it as it is.
static __device__ __noinline__ V3 warp_point_other(const DScene& sc, const pvgpu_warp& w, const V3& p)
{
    const double pi = 3.1415926535897932384626, two_pi = 6.283185307179586476925286766560;
    const double* q = sc.shape_data + w.transform;
    // the orientation step the cylindrical / spherical / toroidal / planar warps end with
    auto orient = [](const double* O, double x, double y, double z) {
        if ((O[0] == 0.0) && (O[1] == 0.0) && (O[2] == 1.0)) return mk(x, y, z);
        return mk((O[0] * z) + (O[1] * x) + (O[2] * x), (O[0] * y) + (O[1] * -z) + (O[2] * y), (O[0] * -x) + (O[1] * y) + (O[2] * z));
    };
    // angle of (x, z) from (1, 0) in the x-z plane, 0 .. 2 pi
    auto azimuth = [&](double x, double z, double len) {
        if (z == 0.0) return (x > 0) ? 0.0 : pi;
        const double t = acos(x / len);
        return (z < 0.0) ? two_pi - t : t;
    };
    double x = p.x, y = p.y, z = p.z;
    switch (w.type) {
        case PVGPU_WARP_BLACK_HOLE: {
            V3 C = mk(q[0], q[1], q[2]);
            const uint32_t flags = (uint32_t)q[9];
            if (flags & 2u) {
                int bx = 0, by = 0, bz = 0;
                if (q[3] >= PV_EPSILON) bx = (int)floor(p.x / q[3]);
                if (q[4] >= PV_EPSILON) by = (int)floor(p.y / q[4]);
                if (q[5] >= PV_EPSILON) bz = (int)floor(p.z / q[5]);
                C.x += q[3] * bx; C.y += q[4] * by; C.z += q[5] * bz;
            }
            V3 delta = p - C;
            double len = length(delta);
            if (len >= q[7]) return p;
            if ((int)q[10] == 0) {
                len = (q[7] - len) / q[7];
                double S = pow(len, q[8]) * q[6];
                if (S > 1.0) S = 1.0;
                delta = delta * ((flags & 1u) ? -S : S);
                return p + delta;
            }
            return p;
        }
        case PVGPU_WARP_REPEAT: {
            const int axis = (int)q[0];
            const float width = (float)q[1];
            V3 t = p;
            double ta = comp(t, axis);
            const float blk = (float)floor(ta / (double)width);
            ta -= (double)(blk * width);
            if (axis == 0) t.x = ta; else if (axis == 1) t.y = ta; else t.z = ta;
            if (((int)blk) & 1) {
                t = mk(t.x * q[2], t.y * q[3], t.z * q[4]);
                if (q[2 + axis] < 0) { if (axis == 0) t.x += (double)width; else if (axis == 1) t.y += (double)width; else t.z += (double)width; }
            }
            return t + (double)blk * mk(q[5], q[6], q[7]);
        }
        case PVGPU_WARP_CUBIC: {
            const double ax = fabs(x), ay = fabs(y), az = fabs(z);
            if (x >= 0 && x >= ay && x >= az) return mk(0.75 - 0.25 * (z / x + 1.0) / 2.0, 1.0 / 3.0 + (1.0 / 3.0) * (y / x + 1.0) / 2.0, x);
            if (y >= 0 && y >= ax && y >= az) return mk(0.25 + 0.25 * (x / y + 1.0) / 2.0, 1.0 - (1.0 / 3.0) * (z / y + 1.0) / 2.0, y);
            if (z >= 0 && z >= ax && z >= ay) return mk(0.25 + 0.25 * (x / z + 1.0) / 2.0, 1.0 / 3.0 + (1.0 / 3.0) * (y / z + 1.0) / 2.0, z);
            if (x < 0 && x <= -ay && x <= -az) { x = -x; return mk(0.25 * (z / x + 1.0) / 2.0, 1.0 / 3.0 + (1.0 / 3.0) * (y / x + 1.0) / 2.0, x); }
            if (y < 0 && y <= -ax && y <= -az) { y = -y; return mk(0.25 + 0.25 * (x / y + 1.0) / 2.0, (1.0 / 3.0) * (z / y + 1.0) / 2.0, y); }
            z = -z;
            return mk(1.0 - 0.25 * (x / z + 1.0) / 2.0, 1.0 / 3.0 + (1.0 / 3.0) * (y / z + 1.0) / 2.0, z);
        }
        case PVGPU_WARP_CYLINDRICAL: {
            const double len = sqrt(x * x + z * z);
            if (len == 0.0) return p;
            double theta = azimuth(x, z, len) / two_pi;
            if (q[3] == 1.0) theta *= len; else if (q[3] != 0.0) theta *= pow(len, q[3]);
            return orient(q, theta, y, len);
        }
        case PVGPU_WARP_SPHERICAL: {
            const double dist = sqrt(x * x + y * y + z * z);
            if (dist == 0.0) return p;
            x /= dist; y /= dist; z /= dist;
            double phi = 0.5 + asin(y) / pi, theta;
            double len = sqrt(x * x + z * z);
            if (len == 0.0) theta = 0;            // at a pole: "any value of theta will do"
            else theta = azimuth(x, z, len) / two_pi;
            if (q[3] == 1.0) { theta *= dist; phi *= dist; }
            else if (q[3] != 0.0) { theta *= pow(dist, q[3]); phi *= pow(dist, q[3]); }
            return orient(q, theta, phi, dist);
        }
        case PVGPU_WARP_TOROIDAL: {
            double len = sqrt(x * x + z * z);
            if (len == 0.0) return p;
            double theta = 0.0 - azimuth(x, z, len);
            x = len - q[4];
            len = sqrt(x * x + y * y);
            double phi = acos(-x / len);
            if (y > 0.0) phi = two_pi - phi;
            theta /= (-two_pi);
            phi /= two_pi;
            if (q[3] == 1.0) { theta *= len; phi *= len; }
            else if (q[3] != 0.0) { theta *= pow(len, q[3]); phi *= pow(len, q[3]); }
            return orient(q, theta, phi, len);
        }
        case PVGPU_WARP_PLANAR: return orient(q, x, y, q[3]);
        default: return p;
    }
}
#endif

// Warp_EPoint (warp.cpp:103-122): warps applied last-to-first, then clamped to COORDINATE_LIMIT.
__device__ inline V3 warp_epoint(const DScene& sc, const pvgpu_pigment& pg, const V3& ep)
{
    V3 p = ep;
    for (int i = (int)pg.warp_count - 1; i >= 0; i--) {
        const pvgpu_warp& w = sc.warps[pg.warp_first + i];
        if (w.type == PVGPU_WARP_TRANSFORM) p = inv_trans_point(sc.xf[w.transform], p);
#if PV_FULL_MATERIALS
        else if (w.type > PVGPU_WARP_CLASSIC_TURBULENCE) p = warp_point_other(sc, w, p);
#endif
#if !PV_BASIC_PATTERNS
        else {   // GenericTurbulenceWarp::WarpPoint (warp.cpp:553-559)
            V3 t = dturbulence(sc.noise, p, w.octaves, (double)w.lambda, (double)w.omega);
            p = mk(p.x + t.x * w.turbulence[0], p.y + t.y * w.turbulence[1], p.z + t.z * w.turbulence[2]);
        }
#endif
    }
    if (p.x > PV_COORDINATE_LIMIT) p.x = PV_COORDINATE_LIMIT; else if (p.x < -PV_COORDINATE_LIMIT) p.x = -PV_COORDINATE_LIMIT;
    if (p.y > PV_COORDINATE_LIMIT) p.y = PV_COORDINATE_LIMIT; else if (p.y < -PV_COORDINATE_LIMIT) p.y = -PV_COORDINATE_LIMIT;
    if (p.z > PV_COORDINATE_LIMIT) p.z = PV_COORDINATE_LIMIT; else if (p.z < -PV_COORDINATE_LIMIT) p.z = -PV_COORDINATE_LIMIT;
    return p;
}

#if PV_FULL_MATERIALS
// The FractalPattern family (pattern.cpp:6895-7098 julia, 7228-7550 magnet, 7551-7751 mandel; ExteriorColour / InteriorColour 8990-9056):
// one orbit loop, the iteration formula picked by the record's kind.  z starts at the point (julia kinds: c is the record's
// constant; mandel kinds: c is the point) or at 0 (magnet "m" kinds, which also start mindist2 at 10000).
static __device__ __noinline__ double fractal_pattern(const double* fd, const V3& p)
{
    const int kind = (int)fd[0], it_max = (int)fd[1], ext_type = (int)fd[2], int_type = (int)fd[3];
    const double ext_f = fd[4], int_f = fd[5];
    const bool julia = (kind == PVGPU_FRACTAL_JULIA2 || kind == PVGPU_FRACTAL_JULIA3 || kind == PVGPU_FRACTAL_JULIA4 ||
                        kind == PVGPU_FRACTAL_MAGNET1J || kind == PVGPU_FRACTAL_MAGNET2J);
    const bool magnet = kind >= PVGPU_FRACTAL_MAGNET1M;
    const double cr = julia ? fd[6] : p.x, ci = julia ? fd[7] : p.y;
    double a = p.x, b = p.y;
    if (magnet && !julia) a = b = 0.0;
    double a2 = a * a, b2 = b * b;
    double mindist2 = (magnet && !julia) ? 10000.0 : a2 + b2;
    const double c1r = cr - 1, c2r = cr - 2;
    const double c1c2r = c1r * c2r - ci * ci, c1c2i = (c1r + c2r) * ci;           // magnet 2 only
    int col;
    for (col = 0; col < it_max; col++) {
        if (!magnet) {
            switch (kind) {
                case PVGPU_FRACTAL_MANDEL2: case PVGPU_FRACTAL_JULIA2:
                    b = 2.0 * a * b + ci;
                    a = a2 - b2 + cr;
                    break;
                case PVGPU_FRACTAL_MANDEL3: case PVGPU_FRACTAL_JULIA3:
                    b = 3.0 * a2 * b - b2 * b + ci;
                    a = a2 * a - 3.0 * a * b2 + cr;
                    break;
                default:
                    b = 4.0 * (a2 * a * b - a * b2 * b) + ci;
                    a = a2 * a2 - 6.0 * a2 * b2 + b2 * b2 + cr;
                    break;
            }
        } else {
            double t1r, t1i, t2r, t2i;
            if (kind <= PVGPU_FRACTAL_MAGNET1J) {
                t1r = a2 - b2 + cr - 1;
                t1i = 2 * a * b + ci;
                t2r = 2 * a + cr - 2;
                t2i = 2 * b + ci;
            } else {
                t1r = a2 * a - 3 * a * b2 + 3 * (a * c1r - b * ci) + c1c2r;
                t1i = 3 * a2 * b - b2 * b + 3 * (a * ci + b * c1r) + c1c2i;
                t2r = 3 * (a2 - b2) + 3 * (a * c2r - b * ci) + c1c2r + 1;
                t2i = 6 * a * b + 3 * (a * ci + b * c2r) + c1c2i;
            }
            const double den = t2r * t2r + t2i * t2i;
            a = (t1r * t2r + t1i * t2i) / den;
            b = (t1i * t2r - t1r * t2i) / den;
            b2 = b * b;
            b = 2 * a * b;
            a = a * a - b2;
        }
        a2 = a * a;
        b2 = b * b;
        const double dist2 = a2 + b2;
        if (dist2 < mindist2) mindist2 = dist2;
        bool out;
        if (magnet) { const double am1 = a - 1; out = dist2 > 10000.0 || am1 * am1 + b2 < 1 / 10000.0; }
        else out = dist2 > 4.0;
        if (out) {
            switch (ext_type) {
                case 0: return ext_f;
                case 1: return (double)col / (double)it_max;
                case 2: return a * ext_f;
                case 3: return b * ext_f;
                case 4: return a * a * ext_f;
                case 5: return b * b * ext_f;
                case 6: return sqrt(a * a + b * b) * ext_f;
                case 7: return (double)((unsigned int)col % (unsigned int)ext_f) / ext_f;
                case 8: return (double)((unsigned int)col % (unsigned int)(1 + ext_f)) / ext_f;
                default: return 0.0;
            }
        }
    }
    switch (int_type) {
        case 0: return int_f;
        case 1: return sqrt(mindist2) * int_f;
        case 2: return a * int_f;
        case 3: return b * int_f;
        case 4: return a * a * int_f;
        case 5: return b * b * int_f;
        case 6: return a * a + b * b * int_f;
        default: return 0.0;
    }
}

// Spiral1Pattern / Spiral2Pattern::EvaluateRaw (pattern.cpp:8396-8437, 8473-8517); tv = the classic-turbulence term
__device__ inline double spiral_pattern(const V3& p, double arms, double tv, bool second)
{
    const double rad = sqrt(p.x * p.x + p.y * p.y);
    double phi = 0.0;
    if (rad != 0.0) phi = (p.x < 0.0) ? 3.0 * 1.57079632679489661923 - asin(p.y / rad) : 1.57079632679489661923 + asin(p.y / rad);
    const double s = p.z + rad + arms * phi / 6.283185307179586476925286766560 + tv;
    return second ? triangle_wave(rad) + triangle_wave(s) : s;
}

// CracklePattern::EvaluateRaw (pattern.cpp:5760-5987) without the (result-neutral) per-thread cell cache: the 81 nuclei of the
// cubes around the point come straight from IntPickInCube = Hash3d + three entries of gPatternRands (mt19937 / 2^32).
static __device__ __noinline__ double crackle_pattern(const DScene& sc, const pvgpu_pigment& pg, const V3& ep, int gen)
{
    const double* cp = sc.shape_data + pg.data;
    const double form_x = cp[0], form_y = cp[1], form_z = cp[2], metric = cp[3], offset = cp[4];
    const bool is_solid = cp[5] != 0.0;
    const int rep[3] = { (int)cp[6], (int)cp[7], (int)cp[8] };
    const bool use_square = (metric == 2), use_unity = (metric == 1);
    auto wrap = [](double val, double upper) {                       // wrap() mathutil.h:102-121
        double t = fmod(val, upper);
        if (t < 0.0) t += upper;
        if (t >= upper) t = 0.0;
        return t;
    };
    V3 tp = ep;
    if (rep[0]) tp.x = wrap(tp.x, (double)rep[0]);
    if (rep[1]) tp.y = wrap(tp.y, (double)rep[1]);
    if (rep[2]) tp.z = wrap(tp.z, (double)rep[2]);
    const int flo[3] = { (int)floor(tp.x - PV_EPSILON), (int)floor(tp.y - PV_EPSILON), (int)floor(tp.z - PV_EPSILON) };
    // nucleus of cube `index` of gaCrackleCubeTable (pattern.cpp:9349-9374) around the point: IntPickInCube (pattern.cpp:8808-8819)
    auto nucleus = [&](int ax, int ay, int az) {
        int c[3] = { flo[0] + ax, flo[1] + ay, flo[2] + az };
        double woff[3] = { 0.0, 0.0, 0.0 };
        for (int k = 0; k < 3; k++)
            if (rep[k]) { int w = c[k] % rep[k]; if (w < 0) w += rep[k]; woff[k] += (c[k] - w); c[k] = w; }     // wrapInt
        const unsigned seed = sc.noise.hash[sc.noise.hash[sc.noise.hash[c[0] & 0xfff] ^ (c[1] & 0xfff)] ^ (c[2] & 0xfff)];                  // Hash3d texture.h:75
        double nx = c[0] + sc.pattern_rands[seed % 32768u], ny = c[1] + sc.pattern_rands[(seed + 1u) % 32768u], nz = c[2] + sc.pattern_rands[(seed + 2u) % 32768u];
        nx += woff[0]; ny += woff[1]; nz += woff[2];
        return mk(nx, ny, nz);
    };
    auto dist = [&](const V3& n) {
        const double dx = n.x - tp.x, dy = n.y - tp.y, dz = n.z - tp.z;
        if (use_square) return dx * dx + dy * dy + dz * dz;
        if (use_unity) return fabs(dx) + fabs(dy) + fabs(dz);
        return pow(fabs(dx), metric) + pow(fabs(dy), metric) + pow(fabs(dz), metric);
    };
    double minsum = 0.0, minsum2 = 0.0, minsum3 = 0.0, tf;
    int min_idx = 0, i = 0;
    for (int ax = -2; ax <= 2; ax++)
        for (int ay = -2; ay <= 2; ay++)
            for (int az = -2; az <= 2; az++) {
                if ((abs(ax) == 2) + (abs(ay) == 2) + (abs(az) == 2) > 1) continue;
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
        else { minsum += pow(offset, metric); minsum2 += pow(offset, metric); minsum3 += pow(offset, metric); }
    }
    if (is_solid) {
        V3 minvec = mk(0.0, 0.0, 0.0);
        i = 0;
        for (int ax = -2; ax <= 2; ax++)
            for (int ay = -2; ay <= 2; ay++)
                for (int az = -2; az <= 2; az++) {
                    if ((abs(ax) == 2) + (abs(ay) == 2) + (abs(az) == 2) > 1) continue;
                    if (i == min_idx) minvec = nucleus(ax, ay, az);
                    i++;
                }
        tf = noise3(sc.noise, minvec, gen);
    }
    else if (use_square) tf = form_x * sqrt(minsum) + form_y * sqrt(minsum2) + form_z * sqrt(minsum3);
    else if (use_unity) tf = form_x * minsum + form_y * minsum2 + form_z * minsum3;
    else tf = form_x * pow(minsum, 1.0 / metric) + form_y * pow(minsum2, 1.0 / metric) + form_z * pow(minsum3, 1.0 / metric);
    return fmax(fmin(tf, 1.), 0.);
}
#endif

// Pattern value for a warped point: Evaluate_TPat -> <Pattern>::Evaluate.
// (out of line in the heavy variants: the pattern switch is large and only patterned pigments / normals come here; the lean
//  variant serves scenes whose pigments use the first pattern set only - device_upload - and keeps it inline)
#if PV_HEAVY
static __device__ __noinline__ double evaluate_pattern(const DScene& sc, const pvgpu_pigment& pg, const V3& p)
#else
__device__ inline double evaluate_pattern(const DScene& sc, const pvgpu_pigment& pg, const V3& p)
#endif
{
    const int gen = pg.noise_generator ? pg.noise_generator : sc.g.noise_generator;   // BasicPattern::GetNoiseGen
    // first warp is a ClassicTurbulence? (GetTurb, pattern.cpp:1135)
    const pvgpu_warp* turb = (pg.warp_count && sc.warps[pg.warp_first].type == PVGPU_WARP_CLASSIC_TURBULENCE) ? &sc.warps[pg.warp_first] : nullptr;
    double value;
    switch (pg.pattern) {
        case PVGPU_PAT_CHECKER: {   // CheckerPattern::Evaluate (pattern.cpp:5691-5707): discrete, no wave processing
            int v = (int)(floor(p.x + PV_EPSILON) + floor(p.y + PV_EPSILON) + floor(p.z + PV_EPSILON));
            return (v & 1) ? 1.0 : 0.0;
        }
#if !PV_BASIC_PATTERNS
        case PVGPU_PAT_BOZO:
        case PVGPU_PAT_SPOTTED:     // NoisePattern::EvaluateRaw (pattern.cpp:7858)
            value = noise3(sc.noise, p, gen);
            break;
        case PVGPU_PAT_GRANITE: {   // GranitePattern::EvaluateRaw (pattern.cpp:6429-6465)
            double noise = 0.0, freq = 1.0;
            V3 tv1 = p * 4.0;
            for (int i = 0; i < 6; freq *= 2.0, i++) {
                V3 tv2 = tv1 * freq;
                double temp;
                if (gen <= 1) temp = fabs(0.5 - noise3(sc.noise, tv2, gen));
                else { temp = fabs(1.0 - 2.0 * noise3(sc.noise, tv2, gen)); if (temp > 0.5) temp = 0.5; }
                noise += temp / freq;
            }
            value = noise;
            break;
        }
        case PVGPU_PAT_GRADIENT: {  // GradientPattern::EvaluateRaw (pattern.cpp:6386-6393)
            double r = dot(p, ld3(pg.p));
            value = (r > 1.0) ? fmod(r, 1.0) : r;
            break;
        }
        case PVGPU_PAT_MARBLE: {    // MarblePattern::EvaluateRaw (pattern.cpp:7831-7847)
            double tv = 0.0;
            if (turb) tv = turb->turbulence[0] * turbulence(sc.noise, p, turb->octaves, (double)turb->lambda, (double)turb->omega, gen);
            value = p.x + tv;
            break;
        }
        case PVGPU_PAT_ONION:       // OnionPattern::EvaluateRaw (pattern.cpp:7934-7953)
            value = fmod(length(p), 1.0);
            break;
        case PVGPU_PAT_WRINKLES: {  // WrinklesPattern::EvaluateRaw (pattern.cpp:8720-8775)
            double lambda = 2.0, omega = 0.5;
            if (gen <= 1) value = noise3(sc.noise, p, gen);
            else value = fmin(fmax(noise3(sc.noise, p, gen) * 2.0 - 0.5, 0.0), 1.0);
            for (int i = 1; i < 10; i++) {
                V3 temp = p * lambda;
                if (gen <= 1) value += omega * noise3(sc.noise, temp, gen);
                else value += omega * fmin(fmax(noise3(sc.noise, temp, gen) * 2.0 - 0.5, 0.0), 1.0);
                lambda *= 2.0;
                omega *= 0.5;
            }
            value = value / 2.0;
            break;
        }
        case PVGPU_PAT_AGATE: {     // AgatePattern::EvaluateRaw (pattern.cpp:5396-5421)
            double tv = 0.0;
            if (turb) tv = pg.p[0] * turbulence(sc.noise, p, turb->octaves, (double)turb->lambda, (double)turb->omega, gen);
            double noise = 0.5 * (cycloidal(1.3 * tv + 1.1 * p.z) + 1.0);
            if (noise < 0.0) noise = 0.0;
            else { noise = fmin(1.0, noise); noise = pow(noise, 0.77); }
            value = noise;
            break;
        }
#if PV_HEAVY
        case PVGPU_PAT_BRICK: {     // BrickPattern::Evaluate (pattern.cpp:5495-5608): discrete
            const double mortar = pg.p[3], fudgit = PV_EPSILON + mortar;
            const double x = p.x + fudgit, y = p.y + fudgit, z = p.z + fudgit;
            const double bw = pg.p[0], bh = pg.p[1], bd = pg.p[2];
            const double mw = mortar / bw, mh = mortar / bh, md = mortar / bd;
            double by = y / bh; by -= (double)(int)by; if (by < 0.0) by += 1.0;
            if (by <= mh) return 0.0;
            by = (y / bh) * 0.5; by -= (double)(int)by; if (by < 0.0) by += 1.0;
            double bx = x / bw; bx -= (double)(int)bx; if (bx < 0.0) bx += 1.0;
            if ((bx <= mw) && (by <= 0.5)) return 0.0;
            bx = (x / bw) + 0.5; bx -= (double)(int)bx; if (bx < 0.0) bx += 1.0;
            if ((bx <= mw) && (by > 0.5)) return 0.0;
            double bz = z / bd; bz -= (double)(int)bz; if (bz < 0.0) bz += 1.0;
            if ((bz <= md) && (by > 0.5)) return 0.0;
            bz = (z / bd) + 0.5; bz -= (double)(int)bz; if (bz < 0.0) bz += 1.0;
            if ((bz <= md) && (by <= 0.5)) return 0.0;
            return 1.0;
        }
        case PVGPU_PAT_HEXAGON: {   // HexagonPattern::Evaluate (pattern.cpp:6512-6655): discrete 0 / 1 / 2
            double x = fabs(p.x), z = (p.z < 0.0) ? 5.196152424 - fabs(p.z) : p.z;
            double xs = x / 0.5, zs = z / 0.866025404;
            xs -= floor(xs / 6.0) * 6.0;
            zs -= floor(zs / 6.0) * 6.0;
            const int xm = (int)pv_floor(xs) % 6, zm = (int)pv_floor(zs) % 6;
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
            return fmod((double)v, 3.0);
        }
        case PVGPU_PAT_WOOD: {      // WoodPattern::EvaluateRaw (pattern.cpp:8651-8683)
            double px = 0.0, py = 0.0;
            if (turb) {
                const V3 wt = dturbulence(sc.noise, p, turb->octaves, (double)turb->lambda, (double)turb->omega);
                px = cycloidal((p.x + wt.x) * turb->turbulence[0]);
                py = cycloidal((p.y + wt.y) * turb->turbulence[1]);
            }
            px += p.x; py += p.y;
            value = length(mk(px, py, 0.0));
            break;
        }
        case PVGPU_PAT_LEOPARD:     // LeopardPattern::EvaluateRaw (pattern.cpp:7179-7193)
            value = sqr((sin(p.x) + sin(p.y) + sin(p.z)) / 3.0);
            break;
        case PVGPU_PAT_SPHERICAL:   // SphericalPattern / BoxedPattern / CylindricalPattern / PlanarPattern + CLIP_DENSITY (pattern.cpp:85)
        case PVGPU_PAT_BOXED:
        case PVGPU_PAT_CYLINDRICAL:
        case PVGPU_PAT_PLANAR:
            if (pg.pattern == PVGPU_PAT_SPHERICAL) value = length(p);
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
            const double n = noise3(sc.noise, p, gen);
            value = n * n * n;
            break;
        }
        case PVGPU_PAT_RIPPLES:     // RipplesPattern::EvaluateRaw (pattern.cpp:8163-8186)
        case PVGPU_PAT_WAVES: {     // WavesPattern::EvaluateRaw (pattern.cpp:8593-8618)
            const uint32_t nw = sc.g.number_of_waves;
            double scalar = 0.0;
            for (uint32_t i = 0; i < nw; i++) {
                double len = length(p - ld3(sc.wave_sources + 3 * i));
                if (len == 0.0) len = 1.0;
                if (pg.pattern == PVGPU_PAT_RIPPLES) scalar += cycloidal(len * (double)pg.frequency + (double)pg.phase);
                else { const double f = sc.wave_freqs[i]; scalar += cycloidal(len * (double)pg.frequency * f + (double)pg.phase) / f; }
            }
            value = (pg.pattern == PVGPU_PAT_RIPPLES) ? 0.5 * (1.0 + (scalar / (double)nw)) : 0.2 * (2.5 + (scalar / (double)nw));
            break;
        }
        case PVGPU_PAT_QUILTED: {   // QuiltedPattern::EvaluateRaw (pattern.cpp:8067-8083)
            V3 v = mk(p.x - pv_floor(p.x) - 0.5, p.y - pv_floor(p.y) - 0.5, p.z - pv_floor(p.z) - 0.5);
            double t = length(v);
            const double it = 1 - t, itsqrd = it * it, tsqrd = t * t, tcubed = t * tsqrd;
            t = (tcubed + 3.0 * t * itsqrd * pg.p[0] + 3.0 * tsqrd * it * pg.p[1]) * 1.154700538;
            v = v * t;
            value = (fabs(v.x) + fabs(v.y) + fabs(v.z)) / 3.0;
            break;
        }
#endif
#endif  // !PV_BASIC_PATTERNS
#if PV_FULL_MATERIALS
        case PVGPU_PAT_CRACKLE:
            value = crackle_pattern(sc, pg, p, gen);
            break;
        case PVGPU_PAT_FRACTAL:
            value = fractal_pattern(sc.shape_data + pg.data, p);
            break;
        case PVGPU_PAT_SPIRAL1:
        case PVGPU_PAT_SPIRAL2: {
            double tv = 0.0;
            if (turb) tv = turb->turbulence[0] * turbulence(sc.noise, p, turb->octaves, (double)turb->lambda, (double)turb->omega, gen);
            value = spiral_pattern(p, pg.p[0], tv, pg.pattern == PVGPU_PAT_SPIRAL2);
            break;
        }
        case PVGPU_PAT_CELLS:       // CellsPattern::EvaluateRaw (pattern.cpp:5652-5660)
            value = fmin(sc.pattern_rands[sc.noise.hash[sc.noise.hash[sc.noise.hash[(int)floor(p.x + PV_EPSILON) & 0xfff] ^ ((int)floor(p.y + PV_EPSILON) & 0xfff)] ^
                                                        ((int)floor(p.z + PV_EPSILON) & 0xfff)] % 32768u], 1.0);
            break;
#endif
        default:
            value = 0.0;
            break;
    }
    // ContinuousPattern::Evaluate (pattern.cpp:354-392)
    if (pg.wave_type == PVGPU_WAVE_RAW) return value;
    if (pg.frequency != 0.0f) value = fmod(value * (double)pg.frequency + (double)pg.phase, 1.00001);
    if (value < 0.0) value -= floor(value);
    switch (pg.wave_type) {
        case PVGPU_WAVE_SINE:     value = (1.0 + cycloidal(value)) * 0.5; break;
        case PVGPU_WAVE_TRIANGLE: value = triangle_wave(value); break;
        case PVGPU_WAVE_SCALLOP:  value = fabs(cycloidal(value * 0.5)); break;
        case PVGPU_WAVE_CUBIC:    value = sqr(value) * ((-2.0 * value) + 3.0); break;
        case PVGPU_WAVE_POLY:     value = pow(value, (double)pg.exponent); break;
        default: break;
    }
    return value;
}

// BlendMap::Search (pattern.cpp:1068-1112): previous / next entry and the previous entry's weight (next = 1 - wp)
__device__ __forceinline__ void blend_search(const pvgpu_blend_entry* e, uint32_t count, double value, uint32_t& ip, uint32_t& in, double& wp)
{
    const uint32_t last = count - 1;
    ip = in = last;
    wp = 0.0;
    if (!(value >= (double)e[last].value)) {
        ip = in = 0;
        while (value > (double)e[in].value) { ip = in; in++; }
        if ((value == (double)e[in].value) || (ip == in)) { ip = in; }
        else wp = ((double)e[in].value - value) / (double)__fsub_rn(e[in].value, e[ip].value);     // SNGL - SNGL rounds to FP32 (pattern.cpp:1105)
    }
}

#if PV_FULL_MATERIALS
// ContinuousPattern::Evaluate (pattern.cpp:354-392) on a raw value computed elsewhere (pigment_pattern: the value is a pigment's greyscale)
__device__ inline double pattern_waveform(const pvgpu_pigment& pg, double value)
{
    if (pg.wave_type == PVGPU_WAVE_RAW) return value;
    if (pg.frequency != 0.0f) value = fmod(value * (double)pg.frequency + (double)pg.phase, 1.00001);
    if (value < 0.0) value -= floor(value);
    switch (pg.wave_type) {
        case PVGPU_WAVE_SINE:     value = (1.0 + cycloidal(value)) * 0.5; break;
        case PVGPU_WAVE_TRIANGLE: value = triangle_wave(value); break;
        case PVGPU_WAVE_SCALLOP:  value = fabs(cycloidal(value * 0.5)); break;
        case PVGPU_WAVE_CUBIC:    value = sqr(value) * ((-2.0 * value) + 3.0); break;
        case PVGPU_WAVE_POLY:     value = pow(value, (double)pg.exponent); break;
        default: break;
    }
    return value;
}
#endif

#if PV_FULL_MATERIALS
// pigment_map / average pigments: Compute_Pigment recursing through PigmentBlendMap::Compute / ComputeAverage
// (pigment.cpp:395-466, 546-596).  The recursion is unrolled over LEVEL (maps nest at most 6 deep: validated on the host), so the
// call graph stays acyclic and ptxas sizes the stack statically.
#define PV_PIGMENT_MAP_LEVELS 6
template <int LEVEL>
static __device__ __noinline__ bool compute_pigment_rec(const DScene& sc, int32_t pig_index, const V3& ep, float col[5], const V3& uvp)
{
    // (the return value is Compute_Pigment's Colour_Found: false only where an image_map used `once` does not cover the point)
    bool found = false;
    auto child = [&](int32_t idx, const V3& p, float out[5]) {
        if constexpr (LEVEL > 0) { if (compute_pigment_rec<LEVEL - 1>(sc, idx, p, out, uvp)) found = true; }
        else { for (int k = 0; k < 5; k++) out[k] = sc.pigments[idx].colour[k]; found = true; }     // unreachable: nesting depth is validated
    };
    const pvgpu_pigment& pg = sc.pigments[pig_index];
    if ((sc.g.quality_flags & PVGPU_Q_QUICK_COLOUR) && pg.quick_colour[0] == pg.quick_colour[0]) {     // pigment.cpp:401-405
        for (int k = 0; k < 5; k++) col[k] = pg.quick_colour[k];
        return true;
    }
    if (pg.pattern == PVGPU_PAT_PLAIN) {
        for (int k = 0; k < 5; k++) col[k] = pg.colour[k];
        return true;
    }
    if (pg.pattern == PVGPU_PAT_UV_MAP) {          // PigmentBlendMap::ComputeUVMapped (pigment.cpp:603-618): no warps, the hit's (u, v, 0)
        child((int32_t)pg.data, uvp, col);
        return found;
    }
    const V3 tp = warp_epoint(sc, pg, ep);
    if (pg.pattern == PVGPU_PAT_IMAGE_MAP) return image_map_colour(sc, sc.images[pg.data], tp, col);
    const pvgpu_blend_map& m = sc.maps[pg.blend_map];
    const pvgpu_blend_entry* e = sc.entries + m.entry_first;
    const bool pmap = (m.blend_mode & PVGPU_BLEND_PIGMENT_MAP) != 0;
    if (pg.pattern == PVGPU_PAT_AVERAGE) {          // ColourBlendMap / PigmentBlendMap::ComputeAverage
        float total = 0.0f;
        for (int k = 0; k < 5; k++) col[k] = 0.0f;
        for (uint32_t i = 0; i < m.entry_count; i++) {
            float t[5];
            if (pmap) child((int32_t)e[i].colour[0], tp, t);
            else for (int k = 0; k < 5; k++) t[k] = e[i].colour[k];
            for (int k = 0; k < 5; k++) col[k] += (float)((double)t[k] * (double)e[i].value);
            total += e[i].value;
        }
        for (int k = 0; k < 5; k++) col[k] = (float)((double)col[k] / (double)total);
        return true;           // Do_Average_Pigments does not report Colour_Found (pigment.cpp:427-433)
    }
    double value;
    if (pg.pattern == PVGPU_PAT_PIGMENT) {          // PigmentPattern::EvaluateRaw (pattern.cpp:7974-7990): greyscale of a pigment at the warped point
        float pc[5] = { 0.0f, 0.0f, 0.0f, 0.0f, 0.0f };
        bool pf = false;
        if constexpr (LEVEL > 0) pf = compute_pigment_rec<LEVEL - 1>(sc, (int32_t)pg.data, tp, pc, uvp);
        value = pattern_waveform(pg, pf ? (double)greyscale_rgb(pc) : 0.0);
    } else value = evaluate_pattern(sc, pg, tp);
    uint32_t ip, in;
    double wp;
    blend_search(e, m.entry_count, value, ip, in, wp);
    if (pmap) child((int32_t)e[in].colour[0], tp, col);
    else for (int k = 0; k < 5; k++) col[k] = e[in].colour[k];
    if (ip != in) {
        float t[5];
        if (pmap) child((int32_t)e[ip].colour[0], tp, t);
        else for (int k = 0; k < 5; k++) t[k] = e[ip].colour[k];
        const double wn = 1.0 - wp;
        for (int k = 0; k < 5; k++) col[k] = (float)((double)t[k] * wp) + (float)((double)col[k] * wn);
    }
    return pmap ? found : true;
}
#endif

// Compute_Pigment (pigment.cpp:395-466) + ColourBlendMap::Compute (pigment.cpp:513-530).  col = rgb, filter, transmit.
// (uvp: the hit's UVCoord as a point, for uv_mapping pigments - hit_uv; anything where there is no hit)
__device__ inline bool compute_pigment(const DScene& sc, int32_t pig_index, const V3& ep, float col[5], const V3& uvp)
{
    const pvgpu_pigment& pg = sc.pigments[pig_index];
    // quickColour (+Q5 and below): Quick_Colour replaces the pigment where the scene gives one (pigment.cpp:401-405; NaN red = none)
    if ((sc.g.quality_flags & PVGPU_Q_QUICK_COLOUR) && pg.quick_colour[0] == pg.quick_colour[0]) {
        #pragma unroll
        for (int k = 0; k < 5; k++) col[k] = pg.quick_colour[k];
        return true;
    }
    if (pg.pattern == PVGPU_PAT_PLAIN) {
        #pragma unroll
        for (int k = 0; k < 5; k++) col[k] = pg.colour[k];
        return true;
    }
#if PV_FULL_MATERIALS
    if (pg.pattern == PVGPU_PAT_IMAGE_MAP) return image_map_colour(sc, sc.images[pg.data], warp_epoint(sc, pg, ep), col);
    if (pg.pattern == PVGPU_PAT_UV_MAP || pg.pattern == PVGPU_PAT_PIGMENT) return compute_pigment_rec<PV_PIGMENT_MAP_LEVELS>(sc, pig_index, ep, col, uvp);
#endif
    const pvgpu_blend_map& m = sc.maps[pg.blend_map];
#if PV_FULL_MATERIALS
    if ((m.blend_mode & PVGPU_BLEND_PIGMENT_MAP) || pg.pattern == PVGPU_PAT_AVERAGE) return compute_pigment_rec<PV_PIGMENT_MAP_LEVELS>(sc, pig_index, ep, col, uvp);
#endif
    const V3 tp = warp_epoint(sc, pg, ep);
    const double value = evaluate_pattern(sc, pg, tp);
    const pvgpu_blend_entry* e = sc.entries + m.entry_first;
    uint32_t ip, in;
    double wp;
    blend_search(e, m.entry_count, value, ip, in, wp);
    if (ip == in) {
        #pragma unroll
        for (int k = 0; k < 5; k++) col[k] = e[in].colour[k];
    } else {
        // GenericPigmentBlendMap::Blend, default blend mode (pigment.cpp:468-511): colour1*w1 + colour2*w2
        const double wn = 1.0 - wp;
        #pragma unroll
        for (int k = 0; k < 5; k++) col[k] = (float)(e[ip].colour[k] * wp) + (float)(e[in].colour[k] * wn);
    }
    return true;
}

// ---- normal perturbation: Perturb_Normal (normal.cpp:784-927) ---------------------------------------
#if PV_FULL_MATERIALS
// Evaluate_TPat of a normal's pattern carrier: pigment_pattern values come from Compute_Pigment, everything else from evaluate_pattern
__device__ inline double evaluate_normal_pattern(const DScene& sc, const pvgpu_pigment& c, const V3& p)
{
    if (c.pattern == PVGPU_PAT_PIGMENT) {
        float pc[5];
        const bool pf = compute_pigment_rec<PV_PIGMENT_MAP_LEVELS>(sc, (int32_t)c.data, p, pc, p);
        return pattern_waveform(c, pf ? (double)greyscale_rgb(pc) : 0.0);
    }
    return evaluate_pattern(sc, c, p);
}
// Warp_Normal / UnWarp_Normal (warp.cpp:563-640): only transform warps act on normals
__device__ inline V3 warp_normal(const DScene& sc, const pvgpu_pigment& c, V3 n, bool dont_scale)
{
    if (!dont_scale) n = normalized(n);
    for (int i = (int)c.warp_count - 1; i >= 0; i--) {
        const pvgpu_warp& w = sc.warps[c.warp_first + i];
        if (w.type == PVGPU_WARP_TRANSFORM) n = inv_trans_normal(sc.xf[w.transform], n);
    }
    if (!dont_scale) n = normalized(n);
    return n;
}
__device__ inline V3 unwarp_normal(const DScene& sc, const pvgpu_pigment& c, V3 n, bool dont_scale)
{
    if (!dont_scale) n = normalized(n);
    for (uint32_t i = 0; i < c.warp_count; i++) {
        const pvgpu_warp& w = sc.warps[c.warp_first + i];
        if (w.type == PVGPU_WARP_TRANSFORM) n = trans_normal(sc.xf[w.transform], n);
    }
    if (!dont_scale) n = normalized(n);
    return n;
}
// Do_Slope_Map + Hermite_Cubic (normal.cpp:929-1001), BlendMap::Search (pattern.cpp:1068-1112)
__device__ inline double do_slope_map(const DScene& sc, const pvgpu_tnormal& tn, double value)
{
    if (tn.slope_count == 0) return value;
    const pvgpu_slope_entry* e = sc.slopes + tn.slope_first;
    const uint32_t last = tn.slope_count - 1;
    if (value >= (double)e[last].value) return e[last].height;
    uint32_t ip = 0, in = 0;
    while (value > (double)e[in].value) { ip = in; in++; }
    if ((value == (double)e[in].value) || (ip == in)) return e[in].height;
    const double wp = ((double)e[in].value - value) / (double)__fsub_rn(e[in].value, e[ip].value);
    const double t1 = 1.0 - wp;
    const double tt = t1 * t1, ttt = tt * t1;
    double rv = ttt * (e[ip].slope + e[in].slope + 2.0 * (e[ip].height - e[in].height));
    rv += -tt * (2.0 * e[ip].slope + e[in].slope + 3.0 * (e[ip].height - e[in].height));
    rv += t1 * e[ip].slope + e[ip].height;
    return rv;
}
// (LEVEL unrolls the recursion through normal_maps, nesting depth validated on the host: the call graph stays acyclic)
#define PV_NORMAL_MAP_LEVELS 3
template <int LEVEL>
static __device__ __noinline__ V3 perturb_normal_t(const DScene& sc, int32_t tn_index, V3 n, const V3& epoint)
{
    const pvgpu_tnormal& tn = sc.tnormals[tn_index];
    const pvgpu_pigment& c = sc.pigments[tn.pattern];
    const bool dont_scale = (tn.flags & PVGPU_DONT_SCALE_BUMPS_FLAG) != 0;
    const double amount = (double)tn.amount;
    if (tn.normal_map) {
        auto child = [&](int32_t idx, const V3& v, const V3& p) -> V3 {
            if constexpr (LEVEL > 0) return perturb_normal_t<LEVEL - 1>(sc, idx, v, p);
            else return v;          // unreachable: nesting depth is validated
        };
        const pvgpu_blend_map& m = sc.maps[tn.normal_map - 1];
        const pvgpu_blend_entry* e = sc.entries + m.entry_first;
        if (tn.type == PVGPU_NORM_AVERAGE) {            // normal.cpp:861-883 (special branch) + NormalBlendMap::ComputeAverage :1033-1059
            n = warp_normal(sc, c, n, dont_scale);
            const V3 tpa = warp_epoint(sc, c, epoint);
            V3 v1 = mk(0.0, 0.0, 0.0);
            float total = 0.0f;
            for (uint32_t i = 0; i < m.entry_count; i++) {
                const V3 v2 = child((int32_t)e[i].colour[0], n, tpa);
                v1 = v1 + (double)e[i].value * v2;
                total += e[i].value;
            }
            n = v1 / (double)total;
            return unwarp_normal(sc, c, n, dont_scale);
        }
        // normal_map selected by a pattern (normal.cpp:824-848)
        const V3 tpm = warp_epoint(sc, c, epoint);
        const double value1 = evaluate_normal_pattern(sc, c, tpm);
        uint32_t ip, in;
        double wp;
        blend_search(e, m.entry_count, value1, ip, in, wp);
        n = warp_normal(sc, c, n, dont_scale);
        const V3 p1 = n;
        n = child((int32_t)e[in].colour[0], n, tpm);
        if (ip != in) {
            const V3 q = child((int32_t)e[ip].colour[0], p1, tpm);
            n = wp * q + (1.0 - wp) * n;
        }
        n = unwarp_normal(sc, c, n, dont_scale);
        return normalized(n);
    }
    n = warp_normal(sc, c, n, dont_scale);
    const V3 tp = warp_epoint(sc, c, epoint);
    switch (tn.type) {
        case PVGPU_NORM_BUMPS:                         // normal.cpp:235-246
            n = n + amount * dnoise3(sc.noise, tp);
            break;
        case PVGPU_NORM_DENTS: {                       // normal.cpp:272-288
            const int gen = c.noise_generator ? c.noise_generator : sc.g.noise_generator;
            double noise = noise3(sc.noise, tp, gen);
            noise = noise * noise * noise * amount;
            n = n + noise * dnoise3(sc.noise, tp);
            break;
        }
        case PVGPU_NORM_RIPPLES:                       // normal.cpp:130-155
        case PVGPU_NORM_WAVES: {                       // normal.cpp:180-209
            const uint32_t nw = sc.g.number_of_waves;
            for (uint32_t i = 0; i < nw; i++) {
                const V3 point = tp - ld3(sc.wave_sources + 3 * i);
                double len = length(point);
                if (len == 0.0) len = 1.0;
                double scalar;
                if (tn.type == PVGPU_NORM_RIPPLES) {
                    const double index = len * (double)c.frequency + (double)c.phase;
                    scalar = cycloidal(index) * amount;
                } else {
                    const double f = sc.wave_freqs[i];
                    const double index = len * (double)c.frequency * f + (double)c.phase;
                    scalar = cycloidal(index) * amount / f;
                }
                n = n + (scalar / (len * (double)nw)) * point;
            }
            break;
        }
        case PVGPU_NORM_WRINKLES: {                    // normal.cpp:325-347
            double scale = 1.0;
            V3 result = mk(0.0, 0.0, 0.0);
            for (int i = 0; i < 10; scale *= 2.0, i++) {
                const V3 value = dnoise3(sc.noise, tp * scale);
                result.x += fabs(value.x / scale); result.y += fabs(value.y / scale); result.z += fabs(value.z / scale);
            }
            n = n + amount * result;
            break;
        }
        case PVGPU_NORM_QUILTED: {                     // normal.cpp:371-391, quilt_cubic pattern.cpp:8949
            V3 value = mk(tp.x - pv_floor(tp.x) - 0.5, tp.y - pv_floor(tp.y) - 0.5, tp.z - pv_floor(tp.z) - 0.5);
            double t = length(value);
            const double p1 = (double)(float)c.p[0], p2 = (double)(float)c.p[1];
            const double it = 1 - t, itsqrd = it * it, tsqrd = t * t, tcubed = t * tsqrd;
            t = (tcubed + 3.0 * t * itsqrd * p1 + 3.0 * tsqrd * it * p2) * 1.154700538;
            n = n + amount * (value * t);
            break;
        }
        default: {                                     // PVGPU_NORM_PATTERN: normal.cpp:893-918
            const double pyr[4][3] = { { 0.942809041, -0.333333333, 0.0 }, { -0.471404521, -0.333333333, 0.816496581 },
                                       { -0.471404521, -0.333333333, -0.816496581 }, { 0.0, 1.0, 0.0 } };
            double am = amount * -5.0;
            am *= 0.02 / (double)tn.delta;
            for (int i = 0; i <= 3; i++) {
                const V3 pv = mk(pyr[i][0], pyr[i][1], pyr[i][2]);
                const V3 p1 = tp + (double)tn.delta * pv;
                const double value1 = do_slope_map(sc, tn, evaluate_normal_pattern(sc, c, p1));
                n = n + (value1 * am) * pv;
            }
            break;
        }
    }
    return unwarp_normal(sc, c, n, dont_scale);
}
__device__ __forceinline__ V3 perturb_normal(const DScene& sc, int32_t tn_index, const V3& n, const V3& epoint)
{
    return perturb_normal_t<PV_NORMAL_MAP_LEVELS>(sc, tn_index, n, epoint);
}
#endif

// ---- finish helpers ------------------------------------------------------------------------------
__device__ __forceinline__ float greyscale(const float* c) { return (float)(0.297 * c[0] + 0.589 * c[1] + 0.114 * c[2]); }   // colour.h:1366

// Trace::FresnelR (trace.cpp:2680-2708)
__device__ inline double fresnel_r(double cosTi, double n)
{
    double sqrg = sqr(n) + sqr(cosTi) - 1.0;
    if (sqrg <= 0.0) return 1.0;
    double g = sqrt(sqrg);
    double quot1 = (g - cosTi) / (g + cosTi);
    double quot2 = (cosTi * (g + cosTi) - 1.0) / (cosTi * (g - cosTi) + 1.0);
    double f = 0.5 * sqr(quot1) * (1.0 + sqr(quot2));
    return fmin(fmax(f, 0.0), 1.0);
}

// Trace::ComputeMetallic (trace.cpp:2656-2669)
__device__ inline void compute_metallic(float c[3], double metallic, const float* mcol, double cos_angle)
{
    if (metallic != 0.0) {
        double x = fabs(acos(cos_angle)) / 1.57079632679489661923;
        double F = 0.014567225 / sqr(x - 1.12) - 0.011612903;
        F = fmin(1.0, fmax(0.0, F));
        #pragma unroll
        for (int k = 0; k < 3; k++) c[k] *= (float)(1.0 + (metallic * (1.0 - F)) * ((double)mcol[k] - 1.0));
    }
}

// Trace::ComputeReflectivity (trace.cpp:2627-2654) + ComputeFresnel (:2671-2675)
__device__ inline void compute_reflectivity(double& weight, float refl[3], const pvgpu_finish& f, double cos_angle, double rel_ior)
{
    if (!f.reflection_fresnel) {
        double wmax = fmax(fmax((double)f.reflection_max[0], (double)f.reflection_max[1]), (double)f.reflection_max[2]);
        double wmin = fmax(fmax((double)f.reflection_min[0], (double)f.reflection_min[1]), (double)f.reflection_min[2]);
        weight = weight * fmax(wmax, wmin);
        double frac;
        if (fabs((double)f.reflection_falloff - 1.0) > PV_EPSILON) frac = pow(1.0 - cos_angle, (double)f.reflection_falloff);
        else frac = 1.0 - cos_angle;
        #pragma unroll
        for (int k = 0; k < 3; k++) {
            if (fabs(frac) < PV_EPSILON) refl[k] = f.reflection_min[k];
            else if (fabs(frac - 1.0) < PV_EPSILON) refl[k] = f.reflection_max[k];
            else refl[k] = (float)(frac * f.reflection_max[k]) + (float)((1.0 - frac) * f.reflection_min[k]);
        }
    } else {
        double fr = fresnel_r(cos_angle, rel_ior);
        #pragma unroll
        for (int k = 0; k < 3; k++) refl[k] = (float)(fr * f.reflection_max[k]) + (float)((1.0 - fr) * f.reflection_min[k]);
        weight = weight * fmax(fmax((double)refl[0], (double)refl[1]), (double)refl[2]);
    }
}

#if PV_FULL_MATERIALS
// Trace::ComputeIridColour (trace.cpp:2486-2518): per-channel interference factor of a thin film; c[] is multiplied in place.
// MathColour arithmetic: FP32 channels, every colour-times-double product rounds to FP32 (colour.h:1681).
__device__ inline void irid_colour(const DScene& sc, const pvgpu_finish& fn, const V3& light_dir, const V3& eye_dir, const V3& layer_normal, const V3& ipoint, float c[3])
{
    double film_thickness = (double)fn.irid_film_thickness;
    if (fn.irid_turb != 0.0f) {
        double noise = turbulence(sc.noise, ipoint, 5, 2.0, 0.5, sc.g.noise_generator);
        noise = 2.0 * noise - 1.0;
        noise = 1.0 + noise * (double)fn.irid_turb;
        film_thickness *= noise;
    }
    const double cl = fabs(dot(layer_normal, light_dir)), ce = fabs(dot(layer_normal, eye_dir));
    const double interference = 2.0 * 3.1415926535897932384626 * film_thickness * (cl + ce);
    #pragma unroll
    for (int k = 0; k < 3; k++) {
        const float q = __fdiv_rn((float)interference, sc.irid_wavelengths[k]);       // GenericColour(a) / b: FP32 division
        const float cs = (float)cos((double)q);
        const float f = (float)((double)(float)((double)cs * (double)fn.irid) + 1.0);
        c[k] *= f;
    }
}
#endif

// cubic_spline (lightsource.cpp:501-517)
__device__ inline double cubic_spline(double low, double high, double pos)
{
    if (pos < low) return 0.0;
    if (pos >= high) return 1.0;
    pos = (pos - low) / (high - low);
    return (3 - 2 * pos) * pos * pos;
}

// Attenuate_Light (lightsource.cpp:548-633)
__device__ inline double attenuate_light(const pvgpu_light& L, const V3& ray_o, const V3& ray_d, double distance)
{
    double att = 1.0;
    if (L.type == PVGPU_LIGHT_SPOT) {
        double costheta = dot(ray_d, ld3(L.direction));
        if (distance > 0.0) costheta = -costheta;
        if (costheta > 0.0) {
            att = pow(costheta, L.coeff);
            if (L.radius > 0.0 && costheta < L.radius) att *= cubic_spline(L.falloff, L.radius, costheta);
        } else return 0.0;
    } else if (L.type == PVGPU_LIGHT_CYLINDER) {
        V3 v1 = ray_o - ld3(L.center);
        double k = dot(v1, ld3(L.direction));
        if (k > 0.0) {
            V3 p = v1 - k * ld3(L.direction);
            double len = length(p);
            if (len < L.falloff) {
                double dist = 1.0 - len / L.falloff;
                att = pow(dist, L.coeff);
                if (L.radius > 0.0 && len > L.radius) att *= cubic_spline(0.0, 1.0 - L.radius / L.falloff, dist);
            } else return 0.0;
        } else return 0.0;
    }
    if (att > 0.0 && L.fade_power > 0.0) {
        if (fabs(L.fade_distance) >= PV_EPSILON) att *= 2.0 / (1.0 + pow(distance / L.fade_distance, L.fade_power));
        else att *= pow(distance, -L.fade_power);
    }
    return att;
}

// ComputeOneWhiteLightRay (trace.cpp:2710-2767); `jitter` = offset of the sampled point of an area light from its centre
__device__ inline void light_ray(const pvgpu_light& L, const V3& ipoint, V3& dir, double& depth, const V3& jitter)
{
    V3 center = ld3(L.center) + jitter;
    if (L.type == PVGPU_LIGHT_CYLINDER) {
        dir = center - ld3(L.points_at);
        V3 to_ctr = center - ipoint;
        double dist_pa = length(dir);
        depth = dot(to_ctr, dir);
        depth /= dist_pa;
        dir = normalized(dir);
    } else {
        dir = center - ipoint;
        depth = length(dir);
        dir = dir / depth;
    }
    if (L.flags & PVGPU_LIGHT_PARALLEL) {
        if (L.flags & PVGPU_LIGHT_AREA) {
            const V3 v1 = normalized(center - ld3(L.points_at));
            depth *= dot(v1, dir);
            dir = v1;
        } else {
            double a = dot(ld3(L.direction), dir);
            depth *= (-a);
            dir = -ld3(L.direction);
        }
    }
}
__device__ inline void light_ray(const pvgpu_light& L, const V3& ipoint, V3& dir, double& depth)
{
    light_ray(L, ipoint, dir, depth, mk(0.0, 0.0, 0.0));
}

// ---- interiors -----------------------------------------------------------------------------------
__device__ __forceinline__ bool ray_is_interior(const PRay& r, int32_t interior)
{
    for (int i = 0; i < r.n_int; i++) if (r.interiors[i] == (uint16_t)interior) return true;
    return false;
}
__device__ __forceinline__ bool ray_remove_interior(PRay& r, int32_t interior)
{
    for (int i = 0; i < r.n_int; i++)
        if (r.interiors[i] == (uint16_t)interior) {
            for (int j = i + 1; j < r.n_int; j++) r.interiors[j - 1] = r.interiors[j];
            r.n_int--;
            return true;
        }
    return false;
}
__device__ __forceinline__ void ray_append_interior(PRay& r, int32_t interior, unsigned int* overflow)
{
    if (r.n_int < PV_MAX_INTERIORS) r.interiors[r.n_int++] = (uint16_t)interior;
    else atomicOr(overflow, 4u);
}

// Trace::ComputeRelativeIOR (trace.cpp:2595-2625)
__device__ inline double relative_ior(const DScene& sc, const PRay& ray, int32_t interior)
{
    if (interior < 0) return 1.0;
    // SceneData::atmosphereIOR is DBL, Interior::IOR is SNGL: atmosphere ratios divide in FP64, object / object ratios in FP32
    const float iorf = sc.interiors[interior].ior;
    const double ior = iorf;
    if (ray.n_int == 0) return ior / (double)sc.g.atmosphere_ior;
    if (ray_is_interior(ray, interior)) {
        if (ray.n_int == 1) return (double)sc.g.atmosphere_ior / ior;
        return (double)__fdiv_rn(sc.interiors[ray.interiors[ray.n_int - 1]].ior, iorf);
    }
    return (double)__fdiv_rn(iorf, sc.interiors[ray.interiors[ray.n_int - 1]].ior);
}

// ---- queues --------------------------------------------------------------------------------------
// Queue slots are taken warp-aggregated: the lanes that arrive at a push together take one atomic and consecutive slots in lane
// order, so the rays of a warp's 8 x 4 pixel block stay neighbours in the next wave (coherent node fetches in the traversal
// kernels that follow) whatever order the warps of the grid finish in.
__device__ __forceinline__ unsigned int take_slot(unsigned int* counter)
{
#ifdef PV_PLAIN_PUSH
    return atomicAdd(counter, 1u);
#else
    const unsigned mask = __activemask();
    const unsigned lane = threadIdx.x & 31u;
    const int leader = __ffs(mask) - 1;
    unsigned int base = 0;
    if ((int)lane == leader) base = atomicAdd(counter, (unsigned int)__popc(mask));
    base = __shfl_sync(mask, base, leader);
    return base + (unsigned int)__popc(mask & ((1u << lane) - 1u));
#endif
}
__device__ __forceinline__ void push_ray(WaveCtx& ctx, const PRay& r)
{
    unsigned int slot = take_slot(ctx.n_next);
    if (slot < ctx.next_cap) store_cs(ctx.next + slot, r);
    else atomicOr(&ctx.cnt->overflow, 8u);
}
__device__ __forceinline__ void push_shadow(WaveCtx& ctx, const SRay& r)
{
    unsigned int slot = take_slot(ctx.n_shadow);
    if (slot < ctx.shadow_cap) store_cs(ctx.shadow + slot, r);
    else atomicOr(&ctx.cnt->overflow, 16u);
}

// ---- object normal / texture selection -----------------------------------------------------------
// Mesh::Normal + Smooth_Mesh_Normal (mesh.cpp:283-375)
__device__ inline V3 mesh_normal(const DScene& sc, const pvgpu_object& ob, const Hit& hit)
{
    const DMesh& me = sc.meshes[ob.mesh];
    const pvgpu_triangle& tr = sc.tris[hit.aux];
    const float* N = sc.norms + 3 * (size_t)me.normal_first;
    const float* V = sc.verts + 3 * (size_t)me.vertex_first;
    V3 result;
    if (tr.flags & PVGPU_TRI_SMOOTH) {
        V3 ip = (ob.transform >= 0) ? inv_trans_point(sc.xf[ob.transform], hit.ip) : hit.ip;
        V3 n1 = ld3f(N + 3 * tr.n1), n2 = ld3f(N + 3 * tr.n2), n3 = ld3f(N + 3 * tr.n3);
        V3 pmp1 = ip - ld3f(V + 3 * tr.p1);
        double u = dot(pmp1, ld3f(tr.perp));
        if (u < PV_EPSILON) result = n1;
        else {
            const int axis = tr.v_axis;
            double k1 = V[3 * tr.p1 + axis], k2 = V[3 * tr.p2 + axis], k3 = V[3 * tr.p3 + axis];
            double v = (comp(pmp1, axis) / u + k1 - k2) / (k3 - k2);
            result = n1 + u * (n2 - n1 + v * (n3 - n2));
        }
        if (ob.transform >= 0) result = trans_normal(sc.xf[ob.transform], result);
        result = normalized(result);
    } else {
        result = ld3f(N + 3 * tr.normal_ind);
        if (ob.transform >= 0) result = normalized(trans_normal(sc.xf[ob.transform], result));
    }
    return result;
}

// (ray_o, ray_d: the ray that found the hit - a glyph's wall normal is a function of the curve parameter the ray solved for)
__device__ inline V3 object_normal(const DScene& sc, const pvgpu_object& ob, const Hit& hit, const V3& ray_o, const V3& ray_d)
{
    switch (ob.type) {
        case PVGPU_OBJ_SPHERE: if (PV_HAS(PVGPU_OBJ_SPHERE)) return sphere_normal(sc, ob, hit.ip); break;
        case PVGPU_OBJ_BOX: if (PV_HAS(PVGPU_OBJ_BOX)) return box_normal(sc, ob, hit.aux); break;
        case PVGPU_OBJ_PLANE: if (PV_HAS(PVGPU_OBJ_PLANE)) return plane_normal(sc, ob); break;
        case PVGPU_OBJ_MESH: if (PV_HAS(PVGPU_OBJ_MESH)) return mesh_normal(sc, ob, hit); break;
#if PV_HEAVY
        case PVGPU_OBJ_QUADRIC: if (PV_HAS(PVGPU_OBJ_QUADRIC)) return quadric_normal(ob, hit.ip); break;
        case PVGPU_OBJ_TORUS: if (PV_HAS(PVGPU_OBJ_TORUS)) return torus_normal(sc, ob, hit.ip, hit.aux); break;
        case PVGPU_OBJ_BLOB: if (PV_HAS(PVGPU_OBJ_BLOB)) return blob_normal(sc, ob, hit.ip); break;
        case PVGPU_OBJ_CONE: if (PV_HAS(PVGPU_OBJ_CONE)) return cone_normal(sc, ob, hit.ip, hit.aux); break;
        case PVGPU_OBJ_DISC: if (PV_HAS(PVGPU_OBJ_DISC)) return ld3(ob.p); break;       // Disc::Normal (disc.cpp:226-229)
        case PVGPU_OBJ_TRIANGLE: if (PV_HAS(PVGPU_OBJ_TRIANGLE)) return triangle_normal(sc, ob, hit.ip); break;
        case PVGPU_OBJ_POLYGON: if (PV_HAS(PVGPU_OBJ_POLYGON)) return ld3(ob.p); break;       // Polygon::Normal (polygon.cpp:308-311)
        case PVGPU_OBJ_POLY: if (PV_HAS(PVGPU_OBJ_POLY)) return poly_normal(sc, ob, hit.ip); break;
#endif
#if PV_FULL_MATERIALS
        // (full variant only - device_upload picks it for scenes with these primitives: 50 KB of solver code behind these two calls
        //  cost config 3 13 % when the plain heavy k_shade carried them)
        case PVGPU_OBJ_GLYPH: if (PV_HAS(PVGPU_OBJ_GLYPH)) return glyph_normal(sc, ob, hit.aux, ray_o, ray_d); break;
        case PVGPU_OBJ_PRISM: if (PV_HAS(PVGPU_OBJ_PRISM)) return prism_normal(sc, ob, hit.ip, hit.aux, ray_o, ray_d); break;
        case PVGPU_OBJ_SUPERELLIPSOID: if (PV_HAS(PVGPU_OBJ_SUPERELLIPSOID)) return superellipsoid_normal(sc, ob, hit.ip); break;
#endif
    }
    return mk(0.0, 1.0, 0.0);
}

#if PV_FULL_MATERIALS
// <Object>::UVCoord of a hit as the point (u, v, 0) texture evaluation uses for it (trace.cpp:500-512, pigment.cpp:603-618):
// Sphere (sphere.cpp:688-752), Box (box.cpp:1028-1077), Torus (torus.cpp:1118-1147), Mesh (mesh.cpp:2256-2332), every other
// primitive ObjectBase::UVCoord (object.cpp:882-886).
static __device__ __noinline__ V3 hit_uv(const DScene& sc, const pvgpu_object& ob, const Hit& hit)
{
    const double pi = 3.1415926535897932384626, two_pi = 6.283185307179586476925286766560;
    switch (ob.type) {
        case PVGPU_OBJ_SPHERE: {
            V3 p;
            if (ob.aux) p = inv_trans_point(sc.xf[ob.transform], hit.ip);
            else { p = hit.ip - ld3(ob.p); if (ob.transform >= 0) p = inv_trans_point(sc.xf[ob.transform], p); }
            double x = p.x, y = p.y, z = p.z;
            double len = sqrt(x * x + y * y + z * z);
            if (len == 0.0) return mk(0.0, 0.0, 0.0);
            x /= len; y /= len; z /= len;
            const double phi = 0.5 + asin(y) / pi;
            double theta;
            len = x * x + z * z;
            if (len > PV_EPSILON) {
                len = sqrt(len);
                if (z == 0.0) theta = (x > 0) ? 0.0 : pi;
                else { theta = acos(x / len); if (z < 0.0) theta = two_pi - theta; }
                theta /= two_pi;
            } else theta = 0;
            return mk(theta, phi, 0.0);
        }
        case PVGPU_OBJ_BOX: {
            V3 P = (ob.transform >= 0) ? inv_trans_point(sc.xf[ob.transform], hit.ip) : hit.ip;
            const V3 b0 = ld3(ob.p), diff = ld3(ob.p + 3) - b0;
            P = P - b0;
            P = mk(P.x / diff.x, P.y / diff.y, P.z / diff.z);
            double u = 0.0, v = 0.0;
            switch (hit.aux) {                                  // kSideHit_X0 .. kSideHit_Z1 = 1 .. 6
                case 1: u = (P.z / 4.0);               v = (1.0 / 3.0) + (P.y / 3.0); break;
                case 2: u = (3.0 / 4.0) - (P.z / 4.0); v = (1.0 / 3.0) + (P.y / 3.0); break;
                case 3: u = (1.0 / 4.0) + (P.x / 4.0); v = (P.z / 3.0); break;
                case 4: u = (1.0 / 4.0) + (P.x / 4.0); v = (3.0 / 3.0) - (P.z / 3.0); break;
                case 5: u = 1.0 - (P.x / 4.0);         v = (1.0 / 3.0) + (P.y / 3.0); break;
                default: u = (1.0 / 4.0) + (P.x / 4.0); v = (1.0 / 3.0) + (P.y / 3.0); break;
            }
            return mk(u, v, 0.0);
        }
        case PVGPU_OBJ_TORUS: {
            const V3 P = inv_trans_point(sc.xf[ob.transform], hit.ip);
            const double u = (1.0 - (atan2(P.z, P.x) + pi) / two_pi);
            const double len = sqrt(P.x * P.x + P.z * P.z);
            const double v = (atan2(P.y, len - ob.p[0]) + pi) / two_pi;
            return mk(u, v, 0.0);
        }
        case PVGPU_OBJ_MESH: {
            if (sc.tri_uv == nullptr) return mk(0.0, 0.0, 0.0);
            const V3 P = (ob.transform >= 0) ? inv_trans_point(sc.xf[ob.transform], hit.ip) : hit.ip;
            const pvgpu_triangle& tr = sc.tris[hit.aux];
            const float* V = sc.verts + 3 * (size_t)sc.meshes[ob.mesh].vertex_first;
            // the vertices are SNGL vectors: their differences are taken in FP32 and then widened (mesh.cpp:2273-2314)
            auto vert = [&](uint32_t i, float out[3]) { out[0] = V[3 * i]; out[1] = V[3 * i + 1]; out[2] = V[3 * i + 2]; };
            float p1[3], p2[3], p3[3];
            vert(tr.p1, p1); vert(tr.p2, p2); vert(tr.p3, p3);
            auto diff = [](const float a[3], const float b[3]) { return mk((double)__fsub_rn(a[0], b[0]), (double)__fsub_rn(a[1], b[1]), (double)__fsub_rn(a[2], b[2])); };
            auto weight = [&](const float far_a[3], const float far_b[3], const float own[3]) {
                // Side1 = far_a - far_b is the opposite side, Side2 = far_a - own an adjacent one; vB = the part of Side1 scaled to reach it
                const V3 side1 = diff(far_a, far_b), side2 = diff(far_a, own);
                const V3 vA = P - mk((double)own[0], (double)own[1], (double)own[2]);
                double t1 = dot(side2, side1), t2 = dot(side1, side1);
                const V3 vB = side1 * (t1 / t2) - side2;
                t1 = dot(vA, vB); t2 = dot(vB, vB);
                return 1 + t1 / t2;
            };
            const double w1 = weight(p3, p2, p1), w2 = weight(p3, p1, p2), w3 = weight(p2, p1, p3);
            const uint32_t* iu = sc.tri_uv + 3 * hit.aux;
            const double* uv = sc.mesh_uv;
            const double u = (w1 * uv[2 * iu[0]] + w2 * uv[2 * iu[1]]) + w3 * uv[2 * iu[2]];
            const double v = (w1 * uv[2 * iu[0] + 1] + w2 * uv[2 * iu[1] + 1]) + w3 * uv[2 * iu[2] + 1];
            return mk(u, v, 0.0);
        }
        default: return mk(hit.ip.x, hit.ip.y, 0.0);
    }
}
#endif

// Where the textures of a hit are evaluated (the intersection point, or (u, v, 0) for an object with `uv_mapping`: trace.cpp:500-512,
// 2351-2362) and the hit's UVCoord for uv_mapping pigments.
__device__ __forceinline__ void texture_points(const DScene& sc, const pvgpu_object& ob, const Hit& hit, V3& tex_ip, V3& uvp)
{
    tex_ip = hit.ip;
    uvp = hit.ip;
#if PV_FULL_MATERIALS
    if (sc.has_uv) {
        uvp = hit_uv(sc, ob, hit);
        if (ob.flags & PVGPU_UV_FLAG) tex_ip = uvp;
    }
#endif
}

// Texture of the hit: ObjectBase::Texture / Interior_Texture (trace.cpp:513-530) or, for a
// multi-textured mesh, the triangle's texture (Mesh::Determine_Textures, mesh.cpp:2421-2457).
__device__ inline int32_t hit_texture(const DScene& sc, const pvgpu_object& ob, const Hit& hit, bool backside)
{
    if (ob.type == PVGPU_OBJ_MESH && (ob.flags & PVGPU_MULTITEXTURE_FLAG)) {
        // Mesh::Determine_Textures (mesh.cpp:2421-2457); ThreeTex triangles are rejected at finalize
        if (backside && ob.interior_texture >= 0) return ob.interior_texture;
        const pvgpu_triangle& tr = sc.tris[hit.aux];
        if (tr.texture >= 0) return (int32_t)sc.index_list[sc.meshes[ob.mesh].texture_first + tr.texture];
        return ob.texture;
    }
    if (ob.texture < 0) return -1;
    return (backside && ob.interior_texture >= 0) ? ob.interior_texture : ob.texture;
}

// Trace::ComputeSky (trace.cpp:2769-2890): background and sky_sphere pigments, both language-version branches.
// Colour * double rounds to FP32 after every product (GenericColour::operator*=(double), colour.h:1681).
__device__ inline void compute_sky(const DScene& sc, const PRay& ray, float col[3], float& transm)
{
    const bool alpha_bg = (ray.flags & PV_RAY_ALPHA_BG) != 0;
    const float* bg = sc.g.background;
    if (sc.g.language_version < 370) {
        if (alpha_bg) { col[0] = col[1] = col[2] = 0.0f; transm = 1.0f; return; }
        col[0] = bg[0]; col[1] = bg[1]; col[2] = bg[2];
        transm = bg[4];
#if PV_FULL_MATERIALS
        if (sc.has_sky) {
            float c[3] = { 0.0f, 0.0f, 0.0f }, fc[3] = { 1.0f, 1.0f, 1.0f }, ff = 1.0f, ft = 1.0f;
            double trans = 1.0;
            V3 p = ld3(ray.d);
            if (sc.sky.transform >= 0) p = inv_trans_point(sc.xf[sc.sky.transform], p);
            for (int i = (int)sc.sky.pigment_count - 1; i >= 0; i--) {
                float t[5];
                compute_pigment(sc, (int32_t)sc.index_list[sc.sky.pigment_first + i], p, t, p);
                const double att = trans * (double)(float)(1.0 - (double)t[3] - (double)t[4]);
                #pragma unroll
                for (int k = 0; k < 3; k++) { c[k] += (float)((double)t[k] * att); fc[k] *= t[k]; }
                ff *= t[3]; ft *= t[4];
                trans = fabs((double)ff) + fabs((double)ft);
            }
            #pragma unroll
            for (int k = 0; k < 3; k++) {
                c[k] *= sc.sky.emission[k];
                const float tc = (float)((double)(float)((double)fc[k] * (double)ff) + (double)ft);
                col[k] = col[k] * tc + c[k];
            }
            transm *= ft;
        }
#endif
        return;
    }
    float c[3] = { 0.0f, 0.0f, 0.0f }, fil[3] = { 1.0f, 1.0f, 1.0f };
#if PV_FULL_MATERIALS
    if (sc.has_sky) {
        V3 p = ld3(ray.d);
        if (sc.sky.transform >= 0) p = inv_trans_point(sc.xf[sc.sky.transform], p);
        for (int i = (int)sc.sky.pigment_count - 1; i >= 0; i--) {
            float t[5];
            compute_pigment(sc, (int32_t)sc.index_list[sc.sky.pigment_first + i], p, t, p);
            const double att = (double)(float)(1.0 - (double)t[3] - (double)t[4]);      // TransColour::Opacity
            #pragma unroll
            for (int k = 0; k < 3; k++) {
                c[k] += ((float)((double)t[k] * att) * fil[k]) * sc.sky.emission[k];
                fil[k] *= (float)((double)(float)((double)t[k] * (double)t[3]) + (double)t[4]);   // TransmittedColour
            }
        }
    }
#endif
    const float f = alpha_bg ? bg[3] : 0.0f, t = alpha_bg ? bg[4] : 0.0f;
    const double att = (double)(float)(1.0 - (double)f - (double)t);
    #pragma unroll
    for (int k = 0; k < 3; k++) {
        c[k] += (float)((double)bg[k] * att) * fil[k];
        fil[k] *= (float)((double)(float)((double)bg[k] * (double)f) + (double)t);
        col[k] = c[k];
    }
    transm = fminf(1.0f, fabsf(greyscale(fil)));
}

#if PV_FULL_MATERIALS
// Ray::IsHollowRay (ray.cpp:59-115): every interior the ray is inside of is hollow
__device__ __forceinline__ bool ray_is_hollow(const DScene& sc, const PRay& r)
{
    for (int i = 0; i < r.n_int; i++) if (!sc.interiors[r.interiors[i]].hollow) return false;
    return true;
}
// Trace::ComputeFog + ComputeConstantFogDepth + ComputeGroundFogDepth (trace.cpp:2892-3044) for a non-shadow ray:
// the caller's colour becomes sum_col + sum_att * colour, its transmittance is scaled by greyscale(sum_att).
static __device__ __noinline__ void compute_fog(const DScene& sc, const V3& o, const V3& d, double depth, float sum_att[3], float sum_col[3])
{
    sum_att[0] = sum_att[1] = sum_att[2] = 1.0f;
    sum_col[0] = sum_col[1] = sum_col[2] = 0.0f;
    for (uint32_t i = 0; i < sc.n_fogs; i++) {
        const pvgpu_fog& fog = sc.fogs[i];
        if (!(fabs(fog.distance) > PV_EPSILON)) continue;
        double width = depth, att;
        const pvgpu_warp* turb = (fog.turbulence >= 0) ? &sc.warps[fog.turbulence] : nullptr;
        if (fog.type == PVGPU_FOG_GROUND) {
            const V3 p1 = o;                                   // ray.Evaluate(0.0)
            const V3 p2 = p1 + d * width;
            const double y1 = dot(p1, ld3(fog.up)), y2 = dot(p2, ld3(fog.up));
            const double start = (y1 - fog.offset) / fog.alt, end = (y2 - fog.offset) / fog.alt;
            double fog_density;
            if (start <= 0.0) {
                if (end <= 0.0) fog_density = 1.0;
                else fog_density = (atan(end) - start) / (end - start);
            } else {
                if (end <= 0.0) fog_density = (atan(start) - end) / (start - end);
                else {
                    const double delta = start - end;
                    if (fabs(delta) > PV_EPSILON) fog_density = (atan(start) - atan(end)) / delta;
                    else fog_density = 1.0 / (sqr(start) + 1.0);
                }
            }
            if (turb) {
                V3 p = (p1 + p2) * 0.5;
                p = mk(p.x * turb->turbulence[0], p.y * turb->turbulence[1], p.z * turb->turbulence[2]);
                const double k = exp(-width / fog.distance);
                width *= (1.0 - k * fmin(1.0, turbulence(sc.noise, p, turb->octaves, (double)turb->lambda, (double)turb->omega, sc.g.noise_generator) * (double)fog.turb_depth));
            }
            att = exp(-width * fog_density / fog.distance);
        } else {
            if (turb) {
                V3 p = evaluate(o, d, width / 2.0);
                p = mk(p.x * turb->turbulence[0], p.y * turb->turbulence[1], p.z * turb->turbulence[2]);
                const double k = exp(-width / fog.distance);
                width *= (1.0 - k * fmin(1.0, turbulence(sc.noise, p, turb->octaves, (double)turb->lambda, (double)turb->omega, sc.g.noise_generator) * (double)fog.turb_depth));
            }
            att = exp(-width / fog.distance);
        }
        const float filter_fog = fog.colour[3], transm_fog = fog.colour[4];
        if (att < (double)transm_fog) att = (double)transm_fog;
        #pragma unroll
        for (int k = 0; k < 3; k++) {
            float t = (float)((double)fog.colour[k] * (double)filter_fog);              // filter_fog * col_fog
            t = (float)((double)t + (1.0 - (double)filter_fog));                        // (1.0 - filter_fog) + ...
            t = (float)((double)t * att);                                               // att * (...)
            sum_att[k] *= t;
            sum_col[k] += (float)((double)fog.colour[k] * (1.0 - att));
        }
    }
}
#endif

// ComputeReflection's direction rule (trace.cpp:1264-1300)
__device__ inline V3 reflect_direction(const V3& dir, const V3& normal, const V3& rawnormal)
{
    double n = -2.0 * dot(dir, normal);
    V3 nd = dir + n * normal;
    n = dot(nd, rawnormal);
    if (n < 0.0) {
        double n2 = dot(nd, normal);
        if (n2 < 0.0) {
            n = -2.0 * dot(dir, rawnormal);
            nd = dir + n * rawnormal;
        } else {
            n *= -2.0;
            nd = nd + n * rawnormal;
        }
    }
    return normalized(nd);
}

// ---- the shading of one closest hit ---------------------------------------------------------------
// ray:      the TraceRay call being served (already past the level / ADC test)
// ray_slot: its index in the current wave (shadow records point back at it)
// One plain (layered) texture a hit's texture tree resolves to: ComputeOneTextureColour (trace.cpp:588-694) walks texture_map /
// average textures down to plain textures, each evaluated at the point its enclosing patterns warped and weighted by the product
// of the blend weights on the way; `chain` = the enclosing patterned textures (their transform warps act on the layer normals).
#define PV_MAX_TEX_LEAVES 16
#define PV_MAX_TEX_DEPTH  4
struct TexLeaf { int32_t tex; int32_t n_chain; double w; V3 p; int32_t chain[PV_MAX_TEX_DEPTH]; };

#if PV_FULL_MATERIALS
// Resolves texture `tex0` at `ipoint` into plain leaves (explicit stack; depth and leaf count are validated on the host).
//   n0 / w0: leaves already in the list and the weight of this texture (a list of weighted textures: Blob::Determine_Textures)
static __device__ __noinline__ int resolve_texture(const DScene& sc, int32_t tex0, const V3& ipoint, TexLeaf* leaves, int n0 = 0, double w0 = 1.0)
{
    TexLeaf st[PV_MAX_TEX_LEAVES];
    int sp = 0, n = n0;
    st[sp].tex = tex0; st[sp].n_chain = 0; st[sp].w = w0; st[sp].p = ipoint; sp++;
    while (sp > 0) {
        const TexLeaf cur = st[--sp];
        const pvgpu_texture& tx = sc.textures[cur.tex];
        if (tx.type == PVGPU_PAT_PLAIN) { if (n < PV_MAX_TEX_LEAVES) leaves[n++] = cur; continue; }
        const pvgpu_pigment& pat = sc.pigments[tx.pigment];
        const pvgpu_blend_map& m = sc.maps[tx.blend_map];
        const pvgpu_blend_entry* e = sc.entries + m.entry_first;
        TexLeaf child = cur;
        if (child.n_chain < PV_MAX_TEX_DEPTH) child.chain[child.n_chain++] = tx.pigment;
        child.p = warp_epoint(sc, pat, cur.p);
        if (tx.type == PVGPU_PAT_AVERAGE) {                      // ComputeAverageTextureColours (trace.cpp:697-737)
            float total = 0.0f;
            for (uint32_t i = 0; i < m.entry_count; i++) total += e[i].value;
            for (uint32_t i = 0; i < m.entry_count && sp < PV_MAX_TEX_LEAVES; i++) {
                child.tex = (int32_t)e[i].colour[0];
                child.w = cur.w * ((double)e[i].value / (double)total);
                st[sp++] = child;
            }
            continue;
        }
        const double value = evaluate_pattern(sc, pat, child.p);
        uint32_t ip, in;
        double wp;
        blend_search(e, m.entry_count, value, ip, in, wp);
        if (sp < PV_MAX_TEX_LEAVES) { child.tex = (int32_t)e[in].colour[0]; child.w = (ip != in) ? cur.w * (1.0 - wp) : cur.w; st[sp++] = child; }
        if (ip != in && sp < PV_MAX_TEX_LEAVES) { child.tex = (int32_t)e[ip].colour[0]; child.w = cur.w * wp; st[sp++] = child; }
    }
    return n;
}
// Warp_Normal / UnWarp_Normal through the enclosing patterned textures (trace.cpp:816-827)
__device__ inline V3 warp_normal_chain(const DScene& sc, const TexLeaf* leaf, V3 n, bool unwarp)
{
    if (!leaf) return n;
    if (!unwarp) for (int i = 0; i < leaf->n_chain; i++) n = warp_normal(sc, sc.pigments[leaf->chain[i]], n, false);
    else for (int i = leaf->n_chain - 1; i >= 0; i--) n = unwarp_normal(sc, sc.pigments[leaf->chain[i]], n, false);
    return n;
}
// Blob::Determine_Textures (blob.cpp:2768-2880): every component whose field is non-zero at the hit point contributes its texture
// (or the blob's own) with weight |field|, normalised to sum 1.  Returns the number of (texture, weight) pairs, at most PV_MAX_TEX_LEAVES.
static __device__ __noinline__ int blob_weighted_textures(const DScene& sc, const pvgpu_object& ob, const V3& ip, int32_t* tex, float* w, unsigned int* overflow)
{
    const pvgpu_blob& bl = sc.blobs[ob.mesh];
    const pvgpu_blob_element* el = sc.blob_elements + bl.element_first;
    const V3 P = (ob.transform >= 0) ? inv_trans_point(sc.xf[ob.transform], ip) : ip;
    int n = 0;
    auto add = [&](uint32_t ei) {
        const double density = fabs(blob_element_field(sc, el[ei], P));
        if (density > 0.0) {
            if (n < PV_MAX_TEX_LEAVES) { const int32_t t = sc.blob_textures[bl.element_first + ei]; tex[n] = (t >= 0) ? t : ob.texture; w[n] = (float)density; n++; }
            else if (overflow) atomicOr(overflow, 64u);
        }
    };
    if (bl.node_count == 0) { for (uint32_t i = 0; i < bl.element_count; i++) add(i); }
    else {
        const pvgpu_blob_node* nodes = sc.blob_nodes + bl.node_first;
        uint32_t queue[PV_BLOB_QUEUE];
        uint32_t size = 0;
        queue[size++] = 0u;
        while (size > 0) {
            const pvgpu_blob_node nd = nodes[queue[--size]];
            if (nd.count == 0) { add(nd.first); continue; }
            for (uint32_t i = 0; i < nd.count; i++) {
                const pvgpu_blob_node& ch = nodes[nd.first + i];
                if (length_sqr(P - ld3(ch.c)) <= ch.r2 && size < PV_BLOB_QUEUE) queue[size++] = nd.first + i;
            }
        }
    }
    if (n > 0) {
        float sum = 0.0f;
        for (int i = 0; i < n; i++) sum += w[i];
        sum = 1.0f / sum;
        for (int i = 0; i < n; i++) w[i] *= sum;
    }
    return n;
}
template <bool LEAF> __device__ inline void shade_hit_impl(const DScene& sc, const PRay& ray, uint32_t ray_slot, const Hit& hit, WaveCtx& ctx, const TexLeaf* leaf);
static __device__ __noinline__ void shade_texture_map(const DScene& sc, const PRay& ray, uint32_t ray_slot, const Hit& hit, WaveCtx& ctx, int32_t tex0);
static __device__ __noinline__ void shade_blob_textures(const DScene& sc, const PRay& ray, uint32_t ray_slot, const Hit& hit, WaveCtx& ctx);
#endif

template <bool LEAF>
__device__ inline void shade_hit_impl(const DScene& sc, const PRay& ray, uint32_t ray_slot, const Hit& hit, WaveCtx& ctx, const TexLeaf* leaf)
{
    const pvgpu_object& ob = sc.objs[hit.obj];
    const V3 dir = ld3(ray.d);
    const V3 ipoint = hit.ip;
    const uint8_t child_level = (ray.flags & PV_RAY_CONTINUED) ? ray.level : (uint8_t)(ray.level + 1);
    const double adc = sc.g.adc_bailout;
    const double weight = (double)ray.adc;

    V3 rawnormal = object_normal(sc, ob, hit, ld3(ray.o), dir);
    if (ob.flags & PVGPU_INVERTED_FLAG) rawnormal = -rawnormal;
    const double normaldirection = dot(rawnormal, dir);
    if (normaldirection > 0.0) rawnormal = -rawnormal;

#if PV_FULL_MATERIALS
    if (!LEAF && ob.type == PVGPU_OBJ_BLOB && (ob.flags & PVGPU_MULTITEXTURE_FLAG) && sc.blob_textures != nullptr) { shade_blob_textures(sc, ray, ray_slot, hit, ctx); return; }
#endif
    const int32_t tex0 = (LEAF && leaf) ? leaf->tex : hit_texture(sc, ob, hit, normaldirection > 0.0);
    if (tex0 < 0) return;
    // a single WeightedTexture of weight 1.0: skipped if 1.0 < adcBailout (trace.cpp:541)
    if (1.0 < adc) return;
#if PV_FULL_MATERIALS
    if (!LEAF && sc.textures[tex0].type != PVGPU_PAT_PLAIN) { shade_texture_map(sc, ray, ray_slot, hit, ctx, tex0); return; }
#endif
    V3 tex_ip, uvp;
    texture_points(sc, ob, hit, tex_ip, uvp);
    const V3 epoint = (LEAF && leaf) ? leaf->p : tex_ip;       // where pigments and normals are evaluated (trace.cpp:588-694)

    const double rel_ior = relative_ior(sc, ray, ob.interior);

    struct Layer {
        float col[3]; float fil[3]; float refl[3]; double att; double rweight; int32_t finish;
#if PV_FULL_MATERIALS
        V3 n;                 // layNormal of the layer (trace.cpp:812-828); the lean variant serves scenes without normal{}
#endif
    };
    Layer layers[PV_MAX_LAYERS];
    int nlayers = 0;
    float fil[3] = { 1.0f, 1.0f, 1.0f };
    double trans = 1.0;
    float amb[3] = { 0.0f, 0.0f, 0.0f };
    bool one_colour_found = false;       // some layer's pigment returned a colour (false only outside an image_map used `once`, trace.cpp:838-841)
    V3 top_normal = rawnormal;
#if PV_FULL_MATERIALS
    const bool has_tn = (sc.has_tnormals != 0u) && (sc.g.quality_flags & PVGPU_Q_NORMALS);
    #define LAYER_NORMAL(L) (has_tn ? (L).n : rawnormal)
#else
    #define LAYER_NORMAL(L) rawnormal
#endif

    for (int32_t li = tex0; li >= 0 && trans > adc && nlayers < PV_MAX_LAYERS; li = sc.textures[li].next) {
        const pvgpu_texture& tx = sc.textures[li];
        const pvgpu_finish& fn = sc.finishes[tx.finish];
        Layer& L = layers[nlayers];
#if PV_FULL_MATERIALS
        if (has_tn) {
            L.n = rawnormal;
            if (tx.tnormal >= 0) {                                                    // trace.cpp:814-828
                L.n = warp_normal_chain(sc, LEAF ? leaf : nullptr, L.n, false);
                L.n = perturb_normal(sc, tx.tnormal, L.n, epoint);
                if (sc.tnormals[tx.tnormal].flags & PVGPU_DONT_SCALE_BUMPS_FLAG) L.n = normalized(L.n);
                L.n = warp_normal_chain(sc, LEAF ? leaf : nullptr, L.n, true);
            }
            if (nlayers == 0) top_normal = L.n;
        }
#endif
        const V3 lay_normal = LAYER_NORMAL(L);
        const double cos_inc = -dot(dir, lay_normal);
        float lc[5];
        const bool colour_found = compute_pigment(sc, tx.pigment, epoint, lc, uvp);
        one_colour_found = one_colour_found || colour_found;
        if (sc.g.quality_flags & PVGPU_Q_AMBIENT_ONLY) {
            // +Q0 / +Q1 (trace.cpp:848-853): the result IS the layer's pigment colour (the last layer reached wins), no transparency,
            // no lights, no secondary rays; the filter colour still decides whether the next layer is looked at (trace.cpp:1059-1076)
            amb[0] = lc[0]; amb[1] = lc[1]; amb[2] = lc[2];
            if (colour_found) {
                #pragma unroll
                for (int k = 0; k < 3; k++) fil[k] *= (lc[k] * lc[3] + lc[4]);
            }
            trans = fmin(1.0, (double)fabsf(greyscale(fil)));
            continue;
        }
        L.col[0] = lc[0]; L.col[1] = lc[1]; L.col[2] = lc[2];
        L.fil[0] = fil[0]; L.fil[1] = fil[1]; L.fil[2] = fil[2];
        L.finish = tx.finish;
        L.rweight = weight * trans;
        compute_reflectivity(L.rweight, L.refl, fn, cos_inc, rel_ior);
        compute_metallic(L.refl, (double)fn.reflect_metallic, L.col, cos_inc);
        double att;
        if (sc.g.language_version < 370) att = (float)(1.0 - ((double)(lc[3] * fmaxf(fmaxf(lc[0], lc[1]), lc[2])) + (double)lc[4]));   // LegacyOpacity
        else att = (float)(1.0 - (double)lc[3] - (double)lc[4]);                                                                   // Opacity
        L.att = att;
        if (fn.alpha_knockout) { L.refl[0] *= (float)att; L.refl[1] *= (float)att; L.refl[2] *= (float)att; }
        // emission + classic ambient (trace.cpp:978-1004); radiosity is outside this path
        float em[3];
        #pragma unroll
        for (int k = 0; k < 3; k++) em[k] = fn.emission[k] + fn.ambient[k] * sc.g.ambient_light[k];
        if (fn.fresnel != 0.0f) {
            float f1 = (float)(1.0 - (double)fn.fresnel * fresnel_r(cos_inc, rel_ior));
            em[0] *= f1; em[1] *= f1; em[2] *= f1;
        }
        #pragma unroll
        for (int k = 0; k < 3; k++) amb[k] += (L.col[k] * em[k] * (float)att) * fil[k];
        nlayers++;
        // new filter colour and remaining translucency (trace.cpp:1059-1076)
        if (colour_found) {
            #pragma unroll
            for (int k = 0; k < 3; k++) fil[k] *= (lc[k] * lc[3] + lc[4]);
            if (fn.conserve_energy != 0) {
                #pragma unroll
                for (int k = 0; k < 3; k++) fil[k] *= fminf(1.0f - L.refl[k], 1.0f);
            }
        }
        trans = fmin(1.0, (double)fabsf(greyscale(fil)));
    }

    // local (non-recursive) term
    accum_add(ctx.accum, ray.sample, ray.w[0] * amb[0], ray.w[1] * amb[1], ray.w[2] * amb[2], 0.0f);
    if (sc.g.quality_flags & PVGPU_Q_AMBIENT_ONLY) return;       // resultTransm = 0, nothing else is computed

    // ---- classic lights: ComputeDiffuseLight / ComputeOneDiffuseLight (trace.cpp:1488-1510, 1637-1728)
    if (!(ob.flags & PVGPU_NO_GLOBAL_LIGHTS_FLAG)) {
        for (uint32_t l = 0; l < sc.n_lights; l++) {
            const pvgpu_light& Lt = sc.lights[l];
            V3 ldir; double ldepth;
            light_ray(Lt, ipoint, ldir, ldepth);
            const double latt = attenuate_light(Lt, ipoint, ldir, ldepth);
            float lcol[3] = { (float)(Lt.colour[0] * latt), (float)(Lt.colour[1] * latt), (float)(Lt.colour[2] * latt) };
            if (fabsf(lcol[0]) < (float)PV_EPSILON && fabsf(lcol[1]) < (float)PV_EPSILON && fabsf(lcol[2]) < (float)PV_EPSILON) continue;
            float K[3] = { 0.0f, 0.0f, 0.0f };
            bool lit = false;           // some layer got as far as the reference's TraceShadowRay call (trace.cpp:1668-1673)
            for (int i = 0; i < nlayers; i++) {
                const Layer& L = layers[i];
                const pvgpu_finish& fn = sc.finishes[L.finish];
                if (!((fn.diffuse != 0.0f) || (fn.diffuse_back != 0.0f) || (fn.specular != 0.0f) || (fn.phong != 0.0f))) continue;
                if (fn.alpha_knockout && L.att == 0.0) continue;
                const V3 lay_normal = LAYER_NORMAL(L);
                bool backside = false;
                if (!(ob.flags & PVGPU_DOUBLE_ILLUMINATE_FLAG)) {
                    double cos_shadow = dot(lay_normal, ldir);
                    if (cos_shadow < PV_EPSILON) {
                        if (fn.diffuse_back != 0.0f) backside = true;
                        else continue;
                    }
                }
                lit = true;
                float k3[3] = { 0.0f, 0.0f, 0.0f };
                // ComputeDiffuseColour (trace.cpp:2441-2484)
                {
                    double diffuse = (double)((backside ? fn.diffuse_back : fn.diffuse) * fn.brilliance_adjust);
                    if (diffuse > 0.0) {
                        double cai = dot(lay_normal, ldir);
                        double intensity = (fn.brilliance != 1.0f) ? pow(fabs(cai), (double)fn.brilliance) : fabs(cai);
                        intensity *= diffuse * L.att;
                        double ff = 1.0;
                        if (fn.fresnel != 0.0f) {
                            double f1 = (double)fn.fresnel * fresnel_r(cai, rel_ior);
                            double f2 = (double)fn.fresnel * fresnel_r(-dot(lay_normal, dir), rel_ior);
                            ff = (1.0 - f1) * (1.0 - f2);
                        }
                        #pragma unroll
                        for (int k = 0; k < 3; k++) k3[k] += (float)(intensity * ff) * L.col[k];
                    }
                }
                const float hl_att = fn.alpha_knockout ? (float)L.att : 1.0f;     // tempLightColour (trace.cpp:1698)
                if (Lt.type != PVGPU_LIGHT_FILL && !backside) {
                    // ComputePhongColour (trace.cpp:2518-2555)
                    if (fn.phong > 0.0f) {
                        double c = -2.0 * dot(dir, lay_normal);
                        V3 rd = dir + c * lay_normal;
                        c = dot(rd, ldir);
                        if (c > 0.0 && ((fn.phong_size < 60.0f) || (c > 0.0008))) {
                            double intensity = (double)fn.phong * pow(c, (double)fn.phong_size);
                            float cs[3] = { 1.0f, 1.0f, 1.0f };
                            if ((fn.fresnel != 0.0f) || (fn.metallic != 0.0f)) {
                                double ndotl = dot(lay_normal, ldir);
                                if (fn.fresnel != 0.0f) { float fr = (float)((double)fn.fresnel * fresnel_r(ndotl, rel_ior)); cs[0] *= fr; cs[1] *= fr; cs[2] *= fr; }
                                compute_metallic(cs, (double)fn.metallic, L.col, ndotl);
                            }
                            #pragma unroll
                            for (int k = 0; k < 3; k++) k3[k] += (float)intensity * cs[k] * hl_att;
                        }
                    }
                    // ComputeSpecularColour (trace.cpp:2557-2593)
                    if (fn.specular > 0.0f) {
                        V3 halfway = ((-dir) + ldir) * 0.5;
                        double hl = length(halfway);
                        if (hl > 0.0) {
                            double c = dot(halfway, lay_normal) / hl;
                            if (c > 0.0) {
                                double intensity = (double)fn.specular * pow(c, (double)fn.roughness);
                                float cs[3] = { 1.0f, 1.0f, 1.0f };
                                if ((fn.fresnel != 0.0f) || (fn.metallic != 0.0f)) {
                                    double ndotl = dot(halfway, ldir) / hl;
                                    if (fn.fresnel != 0.0f) { float fr = (float)((double)fn.fresnel * fresnel_r(ndotl, rel_ior)); cs[0] *= fr; cs[1] *= fr; cs[2] *= fr; }
                                    compute_metallic(cs, (double)fn.metallic, L.col, ndotl);
                                }
                                #pragma unroll
                                for (int k = 0; k < 3; k++) k3[k] += (float)intensity * cs[k] * hl_att;
                            }
                        }
                    }
                }
#if PV_FULL_MATERIALS
                if (fn.irid > 0.0f) irid_colour(sc, fn, ldir, dir, lay_normal, ipoint, k3);      // trace.cpp:1723-1724
#endif
                #pragma unroll
                for (int k = 0; k < 3; k++) K[k] += L.fil[k] * k3[k];
            }
            float a[3] = { ray.w[0] * lcol[0] * K[0], ray.w[1] * lcol[1] * K[1], ray.w[2] * lcol[2] * K[2] };
            // a light that contributes nothing is still shadow-tested by the reference once a layer asked for it (a fully
            // transparent texel of an image_map, say): the ray is traced so that Shadow_Ray_Tests counts the same
            if (a[0] == 0.0f && a[1] == 0.0f && a[2] == 0.0f && !lit) continue;
            const bool shadowed = (sc.g.quality_flags & PVGPU_Q_SHADOWS) && (Lt.type != PVGPU_LIGHT_FILL);
            if (!shadowed) { accum_add(ctx.accum, ray.sample, a[0], a[1], a[2], 0.0f); continue; }
#if PV_FULL_MATERIALS
            // area light: k_shadow_area averages the light colour over the sampled grid and multiplies it in (trace.cpp:2078-2271)
            if ((Lt.flags & PVGPU_LIGHT_AREA) && (sc.g.quality_flags & PVGPU_Q_AREA_LIGHTS)) { a[0] = ray.w[0] * K[0]; a[1] = ray.w[1] * K[1]; a[2] = ray.w[2] * K[2]; }
#endif
            SRay s;
            s.o[0] = ipoint.x; s.o[1] = ipoint.y; s.o[2] = ipoint.z;
            s.d[0] = ldir.x; s.d[1] = ldir.y; s.d[2] = ldir.z;
            s.depth = ldepth;
            s.a[0] = a[0]; s.a[1] = a[1]; s.a[2] = a[2];
            s.sample = ray.sample; s.parent = ray_slot; s.light = l;
            s.pad[0] = s.pad[1] = 0;
            push_shadow(ctx, s);
        }
    }

    // ---- transmitted component: ComputeRefraction / TraceRefractionRay (trace.cpp:1085-1145, 1323-1485)
    bool tir = false;
    if (ob.interior >= 0 && trans > adc && (sc.g.quality_flags & PVGPU_Q_REFRACTIONS)) {
        const pvgpu_interior& in = sc.interiors[ob.interior];
        const double w1 = fmax(fmax((double)fabsf(fil[0]), (double)fabsf(fil[1])), (double)fabsf(fil[2]));   // WeightMaxAbs
        const double new_weight = weight * w1;
        // distance based attenuation (trace.cpp:1098-1121); uses the INCOMING ray's interior list
        float attc[3] = { in.old_refract, in.old_refract, in.old_refract };
        if (ray_is_interior(ray, ob.interior) && fabs((double)in.fade_distance) > PV_EPSILON) {
            if (in.fade_power >= 1000.0f) {
                double depth = hit.depth / (double)in.fade_distance;
                #pragma unroll
                for (int k = 0; k < 3; k++) attc[k] *= expf((float)(-(1.0 - (double)in.fade_colour[k]) * depth));
            } else {
                double a = 1.0 + pow(hit.depth / (double)in.fade_distance, (double)in.fade_power);
                #pragma unroll
                for (int k = 0; k < 3; k++) attc[k] *= (float)((double)in.fade_colour[k] + (1.0 - (double)in.fade_colour[k]) / a);
            }
        }
        PRay nr = ray;
        nr.flags = (uint8_t)((ray.flags & (PV_RAY_REFLECTION | PV_RAY_ALPHA_BG)) | PV_RAY_REFRACTION);
        nr.o[0] = ipoint.x; nr.o[1] = ipoint.y; nr.o[2] = ipoint.z;
        nr.level = child_level;
        nr.adc = (float)new_weight;
        double ior;
        if (nr.n_int == 0) {
            ray_append_interior(nr, ob.interior, &ctx.cnt->overflow);
            ior = (double)sc.g.atmosphere_ior / (double)in.ior;
        } else if ((uint16_t)ob.interior == nr.interiors[nr.n_int - 1]) {
            ray_remove_interior(nr, ob.interior);
            if (nr.n_int == 0) ior = (double)in.ior / (double)sc.g.atmosphere_ior;
            else ior = (double)__fdiv_rn(in.ior, sc.interiors[nr.interiors[nr.n_int - 1]].ior);     // SNGL / SNGL (trace.cpp:1371)
        } else if (ray_remove_interior(nr, ob.interior)) {
            ior = 1.0;
        } else {
            ior = (double)__fdiv_rn(sc.interiors[nr.interiors[nr.n_int - 1]].ior, in.ior);         // SNGL / SNGL (trace.cpp:1388)
            ray_append_interior(nr, ob.interior, &ctx.cnt->overflow);
        }
        bool spawn = true;
        if (fabs(ior - 1.0) < PV_EPSILON) {
            nr.flags |= PV_RAY_CONTINUED;     // TraceRay(nray, ..., continuedRay = true)
            atomicAdd(&ctx.cnt->transmitted, 1ull);
        } else {
            double n = dot(dir, top_normal);
            V3 localnormal;
            if (n <= 0.0) { localnormal = top_normal; n = -n; }
            else localnormal = -top_normal;
            double t = 1.0 + sqr(ior) * (sqr(n) - 1.0);
            if (t < 0.0) {
                // total internal reflection: ComputeReflection with the ORIGINAL ray's interiors (trace.cpp:1462-1470)
                tir = true;
                spawn = false;
                atomicAdd(&ctx.cnt->tir, 1ull);
                atomicAdd(&ctx.cnt->reflected, 1ull);
                PRay rr = ray;
                V3 rd = reflect_direction(dir, top_normal, rawnormal);
                rr.o[0] = ipoint.x; rr.o[1] = ipoint.y; rr.o[2] = ipoint.z;
                rr.d[0] = rd.x; rr.d[1] = rd.y; rr.d[2] = rd.z;
                rr.flags = (uint8_t)((ray.flags & PV_RAY_REFRACTION) | PV_RAY_REFLECTION);
                rr.level = child_level;
                rr.adc = (float)new_weight;
                rr.w[0] = ray.w[0] * attc[0]; rr.w[1] = ray.w[1] * attc[1]; rr.w[2] = ray.w[2] * attc[2];
#if PV_FULL_MATERIALS
                // ComputeReflection(texture->Finish, ...) of the top layer (trace.cpp:1099, 1462-1470)
                if (nlayers > 0 && sc.finishes[layers[0].finish].irid > 0.0f) irid_colour(sc, sc.finishes[layers[0].finish], rd, dir, top_normal, ipoint, rr.w);
#endif
                rr.wt = 0.0f;
                push_ray(ctx, rr);
            } else {
                t = ior * n - sqrt(t);
                V3 nd = ior * dir + t * localnormal;
                nr.d[0] = nd.x; nr.d[1] = nd.y; nr.d[2] = nd.z;
                atomicAdd(&ctx.cnt->refracted, 1ull);
            }
        }
        if (spawn) {
            if (one_colour_found) {      // trace.cpp:1131-1143
                nr.w[0] = ray.w[0] * attc[0] * fil[0]; nr.w[1] = ray.w[1] * attc[1] * fil[1]; nr.w[2] = ray.w[2] * attc[2] * fil[2];
                nr.wt = ray.wt * (float)((double)greyscale(attc) * trans);
            } else {
                nr.w[0] = ray.w[0] * attc[0]; nr.w[1] = ray.w[1] * attc[1]; nr.w[2] = ray.w[2] * attc[2];
                nr.wt = ray.wt * greyscale(attc);
            }
            push_ray(ctx, nr);
        }
    }

    // ---- reflected component (trace.cpp:1151-1178)
    if (sc.g.quality_flags & PVGPU_Q_REFLECTIONS) {
        for (int i = 0; i < nlayers; i++) {
            const Layer& L = layers[i];
            const V3 lay_normal = LAYER_NORMAL(L);
            // after total internal reflection the reflections that use topNormal are skipped (trace.cpp:1155-1159)
            if (tir && !((fabs(top_normal.x - lay_normal.x) > PV_EPSILON) || (fabs(top_normal.y - lay_normal.y) > PV_EPSILON) ||
                         (fabs(top_normal.z - lay_normal.z) > PV_EPSILON))) continue;
            if (L.refl[0] == 0.0f && L.refl[1] == 0.0f && L.refl[2] == 0.0f) continue;
            PRay rr = ray;
            V3 rd = reflect_direction(dir, lay_normal, rawnormal);
            rr.o[0] = ipoint.x; rr.o[1] = ipoint.y; rr.o[2] = ipoint.z;
            rr.d[0] = rd.x; rr.d[1] = rd.y; rr.d[2] = rd.z;
            rr.flags = (uint8_t)((ray.flags & PV_RAY_REFRACTION) | PV_RAY_REFLECTION);   // Ray::SetFlags(ReflectionRay, ray); alphaBackground off
            rr.level = child_level;
            rr.adc = (float)L.rweight;
            rr.w[0] = ray.w[0] * L.refl[0]; rr.w[1] = ray.w[1] * L.refl[1]; rr.w[2] = ray.w[2] * L.refl[2];
#if PV_FULL_MATERIALS
            const float rexp = sc.finishes[L.finish].reflect_exp;
            if (rexp != 1.0f && ctx.conts != nullptr) {
                // resultColour += reflec * Pow(rflCol, Reflect_Exp) (trace.cpp:1166-1168): the reflected subtree is gathered in a slot
                // of its own and enters this ray's slot through a continuation record when the batch's waves are done
                const unsigned int ci = atomicAdd(&ctx.cnt->n_cont, 1u);
                if (ci >= ctx.cont_cap) { atomicOr(&ctx.cnt->overflow, 8u); continue; }
                Cont cr;
                cr.parent = ray.sample; cr.wave = ctx.wave; cr.exponent = rexp;
                cr.w[0] = rr.w[0]; cr.w[1] = rr.w[1]; cr.w[2] = rr.w[2];
                ctx.conts[ci] = cr;
                rr.sample = ctx.cont_base + ci;
                rr.w[0] = rr.w[1] = rr.w[2] = 1.0f;
            }
            if (sc.finishes[L.finish].irid > 0.0f) irid_colour(sc, sc.finishes[L.finish], rd, dir, lay_normal, ipoint, rr.w);    // trace.cpp:1306-1314
#endif
            rr.wt = 0.0f;
            atomicAdd(&ctx.cnt->reflected, 1ull);
            push_ray(ctx, rr);
        }
    }
    #undef LAYER_NORMAL
}

#if PV_FULL_MATERIALS
// texture_map / average texture: every resolved plain texture is shaded like a texture of its own with the ray's weights scaled
// by its blend weight (the reference blends the finished colours, trace.cpp:671-692; everything downstream is linear in them)
static __device__ __noinline__ void shade_texture_map(const DScene& sc, const PRay& ray, uint32_t ray_slot, const Hit& hit, WaveCtx& ctx, int32_t tex0)
{
    TexLeaf leaves[PV_MAX_TEX_LEAVES];
    V3 tex_ip, uvp;
    texture_points(sc, sc.objs[hit.obj], hit, tex_ip, uvp);
    const int n = resolve_texture(sc, tex0, tex_ip, leaves);
    for (int i = 0; i < n; i++) {
        PRay sr = ray;
        const float w = (float)leaves[i].w;
        sr.w[0] *= w; sr.w[1] *= w; sr.w[2] *= w; sr.wt *= w;
        shade_hit_impl<true>(sc, sr, ray_slot, hit, ctx, &leaves[i]);
    }
}
#endif

#if PV_FULL_MATERIALS
// A blob with per-component textures: the weighted texture list of Blob::Determine_Textures, every texture resolved into plain leaves
// and shaded like shade_texture_map does (textures whose weight is below the ADC bailout are skipped, trace.cpp:541)
static __device__ __noinline__ void shade_blob_textures(const DScene& sc, const PRay& ray, uint32_t ray_slot, const Hit& hit, WaveCtx& ctx)
{
    int32_t tex[PV_MAX_TEX_LEAVES];
    float wt[PV_MAX_TEX_LEAVES];
    const int nt = blob_weighted_textures(sc, sc.objs[hit.obj], hit.ip, tex, wt, &ctx.cnt->overflow);
    TexLeaf leaves[PV_MAX_TEX_LEAVES];
    int n = 0;
    for (int i = 0; i < nt; i++) {
        if (tex[i] < 0 || (double)wt[i] < sc.g.adc_bailout) continue;
        n = resolve_texture(sc, tex[i], hit.ip, leaves, n, (double)wt[i]);
    }
    for (int i = 0; i < n; i++) {
        PRay sr = ray;
        const float w = (float)leaves[i].w;
        sr.w[0] *= w; sr.w[1] *= w; sr.w[2] *= w; sr.wt *= w;
        shade_hit_impl<true>(sc, sr, ray_slot, hit, ctx, &leaves[i]);
    }
}
#endif

__device__ inline void shade_hit(const DScene& sc, const PRay& ray, uint32_t ray_slot, const Hit& hit, WaveCtx& ctx)
{
    shade_hit_impl<false>(sc, ray, ray_slot, hit, ctx, nullptr);
}

}  // namespace pvgpu
