// Lattice noise, vector noise and turbulence on the device: PortableNoise / PortableDNoise
// (source/core/material/portablenoise.cpp:103-378), SolidNoise (noise.cpp:392-440), Turbulence and
// DTurbulence (noise.cpp:500-630).  The tables (hashTable, RTable, Perlin permutation + gradients) are
// built on the host with the reference's LCG (noise.cpp:231-255, 306-348) and live in global memory.
#pragma once
#include "pv_math.cuh"

namespace pvgpu {

#define PV_NOISE_MIN   (-10000)           // NOISE_MINX/Y/Z  noise.h:88-90
#define PV_NOISE_ENTRIES 2048             // NoiseEntries    noise.cpp:300
#define PV_ROLLOVER    10000000.023157213 // ROLLOVER        noise.cpp:304

__device__ __forceinline__ double scurve(double a) { return a * a * (3.0 - 2.0 * a); }

struct NoiseCell {
    int ix, iy, iz;
    double x_ix, y_iy, z_iz;
};

__device__ __forceinline__ void noise_cell_axis(double x, int& i, double& f)
{
    int tmp = (x >= 0) ? (int)x : (int)(x - (1 - PV_EPSILON));
    i = (int)((tmp - PV_NOISE_MIN) & 0xFFF);
    f = x - tmp;
}

#define PV_HASH2D(h, a, b) (h[(int)(h[(int)(a)] ^ (b))])
#define PV_RIDX(h, a, b)   ((h[(int)(a) ^ (b)] & 0xFF) * 2)
#define PV_INCRSUMP(mp, s, x, y, z) ((s) * ((mp)[1] + (mp)[2] * (x) + (mp)[4] * (y) + (mp)[6] * (z)))

__device__ inline double solid_noise(const NoiseTables& nt, const V3& P)
{
    int b0[3], b1[3];
    double r0[3], r1[3];
    const double pp[3] = { P.x, P.y, P.z };
    #pragma unroll
    for (int k = 0; k < 3; k++) {
        double t = pp[k] + PV_ROLLOVER;
        int it = (int)floor(t);
        b0[k] = it & (PV_NOISE_ENTRIES - 1);
        b1[k] = (b0[k] + 1) & (PV_NOISE_ENTRIES - 1);
        r0[k] = t - it;
        r1[k] = r0[k] - 1.0;
    }
    int i = nt.perm[b0[0]], j = nt.perm[b1[0]];
    int b00 = nt.perm[i + b0[1]], b10 = nt.perm[j + b0[1]], b01 = nt.perm[i + b1[1]], b11 = nt.perm[j + b1[1]];
    double sx = scurve(r0[0]), sy = scurve(r0[1]), sz = scurve(r0[2]);
    auto at = [&](int idx, double rx, double ry, double rz) {
        const double* q = nt.grad + 3 * idx;
        return rx * q[0] + ry * q[1] + rz * q[2];
    };
    auto lerp = [](double t, double a, double b) { return a + t * (b - a); };
    double u, v, a, b, c, d;
    u = at(b00 + b0[2], r0[0], r0[1], r0[2]); v = at(b10 + b0[2], r1[0], r0[1], r0[2]); a = lerp(sx, u, v);
    u = at(b01 + b0[2], r0[0], r1[1], r0[2]); v = at(b11 + b0[2], r1[0], r1[1], r0[2]); b = lerp(sx, u, v);
    c = lerp(sy, a, b);
    u = at(b00 + b1[2], r0[0], r0[1], r1[2]); v = at(b10 + b1[2], r1[0], r0[1], r1[2]); a = lerp(sx, u, v);
    u = at(b01 + b1[2], r0[0], r1[1], r1[2]); v = at(b11 + b1[2], r1[0], r1[1], r1[2]); b = lerp(sx, u, v);
    d = lerp(sy, a, b);
    return lerp(sz, c, d);
}

// PortableNoise(EPoint, noise_generator): generators 0/1 original, 2 range-corrected, 3 Perlin.
__device__ inline double noise3(const NoiseTables& nt, const V3& P, int gen)
{
    if (gen == 3) {
        double sum = 0.5 * (1.59 * solid_noise(nt, P) + 0.985);
        if (sum < 0.0) sum = 0.0;
        if (sum > 1.0) sum = 1.0;
        return sum;
    }
    int ix, iy, iz;
    double x_ix, y_iy, z_iz;
    noise_cell_axis(P.x, ix, x_ix);
    noise_cell_axis(P.y, iy, y_iy);
    noise_cell_axis(P.z, iz, z_iz);
    double x_jx = x_ix - 1, y_jy = y_iy - 1, z_jz = z_iz - 1;
    double sx = scurve(x_ix), sy = scurve(y_iy), sz = scurve(z_iz);
    double tx = 1 - sx, ty = 1 - sy, tz = 1 - sz;
    double txty = tx * ty, sxty = sx * ty, txsy = tx * sy, sxsy = sx * sy;
    const uint16_t* h = nt.hash;
    int ixiy = PV_HASH2D(h, ix, iy), jxiy = PV_HASH2D(h, ix + 1, iy), ixjy = PV_HASH2D(h, ix, iy + 1), jxjy = PV_HASH2D(h, ix + 1, iy + 1);
    const double* mp;
    double sum;
    mp = nt.rtable + PV_RIDX(h, ixiy, iz);     sum  = PV_INCRSUMP(mp, (txty * tz), x_ix, y_iy, z_iz);
    mp = nt.rtable + PV_RIDX(h, jxiy, iz);     sum += PV_INCRSUMP(mp, (sxty * tz), x_jx, y_iy, z_iz);
    mp = nt.rtable + PV_RIDX(h, ixjy, iz);     sum += PV_INCRSUMP(mp, (txsy * tz), x_ix, y_jy, z_iz);
    mp = nt.rtable + PV_RIDX(h, jxjy, iz);     sum += PV_INCRSUMP(mp, (sxsy * tz), x_jx, y_jy, z_iz);
    mp = nt.rtable + PV_RIDX(h, ixiy, iz + 1); sum += PV_INCRSUMP(mp, (txty * sz), x_ix, y_iy, z_jz);
    mp = nt.rtable + PV_RIDX(h, jxiy, iz + 1); sum += PV_INCRSUMP(mp, (sxty * sz), x_jx, y_iy, z_jz);
    mp = nt.rtable + PV_RIDX(h, ixjy, iz + 1); sum += PV_INCRSUMP(mp, (txsy * sz), x_ix, y_jy, z_jz);
    mp = nt.rtable + PV_RIDX(h, jxjy, iz + 1); sum += PV_INCRSUMP(mp, (sxsy * sz), x_jx, y_jy, z_jz);
    if (gen == 2) {
        sum += 1.05242;
        sum *= 0.48985582;
    } else {
        sum = sum + 0.5;
    }
    if (sum < 0.0) sum = 0.0;
    if (sum > 1.0) sum = 1.0;
    return sum;
}

// PortableDNoise(result, EPoint)
__device__ inline V3 dnoise3(const NoiseTables& nt, const V3& P)
{
    int ix, iy, iz;
    double x_ix, y_iy, z_iz;
    noise_cell_axis(P.x, ix, x_ix);
    noise_cell_axis(P.y, iy, y_iy);
    noise_cell_axis(P.z, iz, z_iz);
    double x_jx = x_ix - 1, y_jy = y_iy - 1, z_jz = z_iz - 1;
    double sx = scurve(x_ix), sy = scurve(y_iy), sz = scurve(z_iz);
    double tx = 1 - sx, ty = 1 - sy, tz = 1 - sz;
    double txty = tx * ty, sxty = sx * ty, txsy = tx * sy, sxsy = sx * sy;
    const uint16_t* h = nt.hash;
    int ixiy = PV_HASH2D(h, ix, iy), jxiy = PV_HASH2D(h, ix + 1, iy), ixjy = PV_HASH2D(h, ix, iy + 1), jxjy = PV_HASH2D(h, ix + 1, iy + 1);
    V3 r = mk(0.0, 0.0, 0.0);
    const double* mp;
    double s;
    bool first = true;
    auto corner = [&](int hash, int z, double sw, double fx, double fy, double fz) {
        mp = nt.rtable + PV_RIDX(h, hash, z);
        s = sw;
        if (first) {
            r.x = PV_INCRSUMP(mp, s, fx, fy, fz); mp += 8;
            r.y = PV_INCRSUMP(mp, s, fx, fy, fz); mp += 8;
            r.z = PV_INCRSUMP(mp, s, fx, fy, fz);
            first = false;
        } else {
            r.x += PV_INCRSUMP(mp, s, fx, fy, fz); mp += 8;
            r.y += PV_INCRSUMP(mp, s, fx, fy, fz); mp += 8;
            r.z += PV_INCRSUMP(mp, s, fx, fy, fz);
        }
    };
    // corner order of portablenoise.cpp:318-376
    corner(ixiy, iz,     txty * tz, x_ix, y_iy, z_iz);
    corner(jxiy, iz,     sxty * tz, x_jx, y_iy, z_iz);
    corner(jxjy, iz,     sxsy * tz, x_jx, y_jy, z_iz);
    corner(ixjy, iz,     txsy * tz, x_ix, y_jy, z_iz);
    corner(ixjy, iz + 1, txsy * sz, x_ix, y_jy, z_jz);
    corner(jxjy, iz + 1, sxsy * sz, x_jx, y_jy, z_jz);
    corner(jxiy, iz + 1, sxty * sz, x_jx, y_iy, z_jz);
    corner(ixiy, iz + 1, txty * sz, x_ix, y_iy, z_jz);
    return r;
}

// Turbulence(EPoint, Turb, noise_generator)                                   noise.cpp:500-560
__device__ inline double turbulence(const NoiseTables& nt, const V3& P, int octaves, double lambda, double omega, int gen)
{
    double value;
    if (gen <= 1) value = noise3(nt, P, gen);
    else { value = 2.0 * noise3(nt, P, gen) - 0.5; value = fmin(fmax(value, 0.0), 1.0); }
    double l = lambda, o = omega;
    for (int i = 2; i <= octaves; i++) {
        V3 temp = P * l;
        if (gen <= 1) value += o * noise3(nt, temp, gen);
        else value += o * (2.0 * noise3(nt, temp, gen) - 0.5);
        if (i < octaves) { l *= lambda; o *= omega; }
    }
    return value;
}

// DTurbulence(result, EPoint, Turb)                                           noise.cpp:580-610
__device__ inline V3 dturbulence(const NoiseTables& nt, const V3& P, int octaves, double lambda, double omega)
{
    V3 result = dnoise3(nt, P);
    double l = lambda, o = omega;
    for (int i = 2; i <= octaves; i++) {
        V3 temp = P * l;
        V3 value = dnoise3(nt, temp);
        result = result + o * value;
        if (i < octaves) { l *= lambda; o *= omega; }
    }
    return result;
}

}  // namespace pvgpu
