// image_map pigments: map_pos + the five mappers (imageutil.cpp:557-999), image_colour_at with its interpolation kinds
// (imageutil.cpp:396-466, 1001-1297) on texels the host has decoded to float RGBFT (Image::GetRGBFTValue), row 0 = top row.
#pragma once
#include "pv_math.cuh"

namespace pvgpu {

#define PV_TWO_PI 6.283185307179586476925286766560
#define PV_PI     3.1415926535897932384626
#define PV_ALPHA_EPSILON 1.0e-6f          // ALPHA_EPSILON (base/image/encoding.cpp)

// wrap (base/mathutil.h:102-127): into [0, limit)
__device__ __forceinline__ double img_wrap(double v, double limit)
{
    double t = fmod(v, limit);
    if (t < 0.0) t += limit;
    if (t >= limit) t = 0.0;
    return t;
}

// angle of (x, z) from the +x axis in the x-z plane, as the mappers compute it (acos form, imageutil.cpp:598-612)
__device__ __forceinline__ double img_theta(double x, double z, double len)
{
    if (z == 0.0) return (x > 0.0) ? 0.0 : PV_PI;
    double theta = acos(x / len);
    if (z < 0.0) theta = PV_TWO_PI - theta;
    return theta;
}

// map_pos (imageutil.cpp:932-999): false = the point is outside the image (no colour)
__device__ inline bool image_map_pos(const pvgpu_image& im, const V3& p, double& xc, double& yc)
{
    const bool once = (im.flags & PVGPU_IMAGE_ONCE) != 0;
    double x = p.x, y = p.y, z = p.z, u = 0.0, v = 0.0, len;
    switch (im.map_type) {
        case 1: {      // spherical_image_map
            len = sqrt(x * x + y * y + z * z);
            if (len == 0.0) return false;
            x /= len; y /= len; z /= len;
            const double phi = 0.5 + asin(y) / PV_PI;
            len = sqrt(x * x + z * z);
            const double theta = (len == 0.0) ? 0.0 : img_theta(x, z, len) / PV_TWO_PI;
            u = theta * (double)im.fwidth;
            v = phi * (double)im.fheight;
            break;
        }
        case 2: {      // cylindrical_image_map
            if (once && ((y < 0.0) || (y > 1.0))) return false;
            v = fmod(y * (double)im.fheight, (double)im.fheight);
            len = sqrt(x * x + y * y + z * z);
            if (len == 0.0) return false;
            x /= len; z /= len;
            len = sqrt(x * x + z * z);
            if (len == 0.0) return false;
            u = (img_theta(x, z, len) / PV_TWO_PI) * (double)im.fwidth;
            break;
        }
        case 5: {      // torus_image_map
            const double r0 = im.gradient[0];
            len = sqrt(x * x + z * z);
            if (len == 0.0) return false;
            double theta = 0.0 - img_theta(x, z, len);
            x = len - r0;
            len = sqrt(x * x + y * y);
            double phi = acos(-x / len);
            if (y > 0.0) phi = PV_TWO_PI - phi;
            theta /= PV_TWO_PI;
            phi /= PV_TWO_PI;
            u = -theta * (double)im.fwidth;
            v = phi * (double)im.fheight;
            break;
        }
        case 7: {      // angular_image_map
            len = sqrt(x * x + y * y + z * z);
            if (len == 0.0) return false;
            x /= len; y /= len; z /= len;
            const double r = ((x == 0.0) && (y == 0.0)) ? 0.0 : (1.0 / PV_PI) * acos(z) / sqrt(x * x + y * y);
            u = (x * r + 1.0) / 2.0 * (double)im.fwidth;
            v = (y * r + 1.0) / 2.0 * (double)im.fheight;
            break;
        }
        default: {     // planar_image_map
            const double c[3] = { x, y, z };
            #pragma unroll
            for (int k = 0; k < 3; k++) {
                if (im.gradient[k] != 0.0) {
                    if (once && ((c[k] < 0.0) || (c[k] > 1.0))) return false;
                    if (im.gradient[k] > 0.0) u = fmod(c[k] * (double)im.fwidth, (double)im.fwidth);
                    else v = fmod(c[k] * (double)im.fheight, (double)im.fheight);
                }
            }
            break;
        }
    }
    u += im.offset[0] + PV_EPSILON;
    v += im.offset[1] + PV_EPSILON;
    if (once && ((u >= (double)im.width) || (v >= (double)im.height) || (u < 0.0) || (v < 0.0))) return false;
    xc = img_wrap(u, (double)im.width);
    yc = img_wrap(-v, (double)im.height);      // image rows run top to bottom
    return true;
}

// no_interpolation (imageutil.cpp:1001-1055): the texel under (x, y), clamped (once) or wrapped (repeat)
__device__ __forceinline__ const float* image_texel(const DScene& sc, const pvgpu_image& im, double xc, double yc)
{
    int ix, iy;
    if (im.flags & PVGPU_IMAGE_ONCE) {
        ix = (xc < 0.0) ? 0 : (xc >= (double)im.width) ? (int)im.width - 1 : (int)xc;
        iy = (yc < 0.0) ? 0 : (yc >= (double)im.height) ? (int)im.height - 1 : (int)yc;
    } else {
        ix = (int)img_wrap(xc, (double)im.width);
        iy = (int)img_wrap(yc, (double)im.height);
    }
    return sc.texels + 5 * ((size_t)im.data_first + (size_t)iy * im.width + (size_t)ix);
}

// ColourImagePattern::Evaluate (pattern.cpp:493-514) = map_pos + image_colour_at(..., premul = false).  Returns false (and the
// clear colour) outside the map.  col = r g b filter transmit.
__device__ inline bool image_map_colour(const DScene& sc, const pvgpu_image& im, const V3& p, float col[5])
{
    double xc, yc;
    if (!image_map_pos(im, p, xc, yc)) { col[0] = col[1] = col[2] = 1.0f; col[3] = 0.0f; col[4] = 1.0f; return false; }
    if (im.interpolation == 0) {
        const float* t = image_texel(sc, im, xc, yc);
        #pragma unroll
        for (int k = 0; k < 5; k++) col[k] = t[k];
    } else {
        // Interp / InterpolateBicubic (imageutil.cpp:1074-1203): weights in FP64, the sum rounded to FP32 once
        double acc[5] = { 0.0, 0.0, 0.0, 0.0, 0.0 };
        xc += 0.5; yc += 0.5;
        const int ix = (int)xc, iy = (int)yc;
        if (im.interpolation == 3) {
            double fx[4], fy[4];
            { const double pp = xc - (double)(int)xc, q = 1.0 - pp; fx[0] = -0.5 * pp * q * q; fx[1] = 0.5 * q * (q * (3.0 * pp + 1.0) + 1.0); fx[2] = 0.5 * pp * (pp * (3.0 * q + 1.0) + 1.0); fx[3] = -0.5 * q * pp * pp; }
            { const double pp = yc - (double)(int)yc, q = 1.0 - pp; fy[0] = -0.5 * pp * q * q; fy[1] = 0.5 * q * (q * (3.0 * pp + 1.0) + 1.0); fy[2] = 0.5 * pp * (pp * (3.0 * q + 1.0) + 1.0); fy[3] = -0.5 * q * pp * pp; }
            for (int i = 0; i < 4; i++)
                for (int j = 0; j < 4; j++) {
                    const float* t = image_texel(sc, im, (double)ix + i - 2, (double)iy + j - 2);
                    const double f = fx[i] * fy[j];
                    #pragma unroll
                    for (int k = 0; k < 5; k++) acc[k] += (double)t[k] * f;
                }
        } else {
            const double pp = xc - (double)(int)xc, q = yc - (double)(int)yc;
            double f[4];
            if (im.interpolation == 2) { f[0] = pp * q; f[1] = (1.0 - pp) * q; f[2] = pp * (1.0 - q); f[3] = (1.0 - pp) * (1.0 - q); }      // bilinear
            else {                                                                                                                       // norm_dist
                double w[4] = { 1.0 / ((1.0 - pp) * (1.0 - pp) + (1.0 - q) * (1.0 - q)), 1.0 / (pp * pp + (1.0 - q) * (1.0 - q)),
                                1.0 / ((1.0 - pp) * (1.0 - pp) + q * q), 1.0 / (pp * pp + q * q) };
                double sum = 0.0;
                for (int i = 0; i < 4; i++) sum += w[i];
                for (int i = 0; i < 4; i++) f[i] = w[i] / sum;
            }
            const double cx[4] = { (double)ix, (double)ix - 1.0, (double)ix, (double)ix - 1.0 };
            const double cy[4] = { (double)iy, (double)iy, (double)iy - 1.0, (double)iy - 1.0 };
            for (int i = 0; i < 4; i++) {
                const float* t = image_texel(sc, im, cx[i], cy[i]);
                #pragma unroll
                for (int k = 0; k < 5; k++) acc[k] += (double)t[k] * f[i];
            }
        }
        #pragma unroll
        for (int k = 0; k < 5; k++) col[k] = (float)acc[k];
    }
    if (im.flags & PVGPU_IMAGE_PREMULTIPLIED) {          // texels were fetched premultiplied, the pigment wants them straight (imageutil.cpp:427-433)
        float a = 1.0f - col[4];
        if (a == 0.0f) a = PV_ALPHA_EPSILON;
        col[0] /= a; col[1] /= a; col[2] /= a;
    }
    if (im.flags & PVGPU_IMAGE_TRANSMIT_ALL) {           // "transmit / filter all" scaled by the image's own alpha (imageutil.cpp:435-457)
        const float alpha = 1.0f - col[4];
        if (alpha != 0.0f) { col[4] += im.all_transmit * alpha; col[3] += im.all_filter * alpha; }
    }
    return true;
}

}  // namespace pvgpu
