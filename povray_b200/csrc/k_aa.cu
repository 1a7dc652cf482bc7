// Anti-aliasing bookkeeping kernels: sampling method 1 (NonAdaptiveSupersamplingM1, tracetask.cpp:521-602, 838-890)
// and method 2 (AdaptiveSupersamplingM2 / SubdivideOnePixel, tracetask.cpp:604-657, 892-1074).
//
// The reference supersamples while it walks a tile pixel by pixel.  Here the tracing itself is always done by the
// wavefront kernels on explicit lists of image-plane coordinates; the kernels of this file only decide WHICH
// coordinates are needed and combine the traced samples in the reference's order:
//   method 1: c0 = pixel centres + the frame column / row left of and above every rectangle; candidates for
//             supersampling = every pixel whose c0-based threshold test against one of its 4 neighbours fires (a
//             superset of what the sequential walk can request from c0 values); their aaDepth^2 jittered samples are
//             traced into one sum slot each; then one thread per rectangle replays the reference's sequential walk
//             exactly; a pixel the walk wants that has no sum yet (possible when a neighbour's supersampled colour
//             fires a test its c0 did not) is queued, traced, and the walk is replayed.
//   method 2: pixel-corner samples; per pixel a (2^aaDepth + 1)^2 sample buffer is filled level by level (one
//             tracing round per subdivision level), then the recursion is replayed to combine the samples.
#include "pv_common.cuh"
#include "pv_kernels.hpp"

namespace pvgpu {

__device__ const float kJitterTable[256] = {
#include "pv_jitter.inc"
};

// Jitter2d(DBL x, DBL y, DBL& jx, DBL& jy) (jitter.h:92-96); hashTable is the noise hash table (noise.cpp:231-255)
__device__ __forceinline__ void jitter2d(const uint16_t* h, double x, double y, double& jx, double& jy)
{
    jx = (double)kJitterTable[int(h[int(h[(int(x * 1021.0) & 0xfff)] ^ int(y * 1019.0)) & 0xfff]) & 0xff];
    jy = (double)kJitterTable[int(h[int(h[(int(x * 1019.0) & 0xfff)] ^ int(y * 1021.0)) & 0xfff]) & 0xff];
}

__device__ __forceinline__ float4 px_add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
// RGBTColour / double: every channel is divided in double and narrowed (colour.h:531-536, 1167-1171)
__device__ __forceinline__ float4 px_div(float4 a, double d)
{
    return make_float4((float)((double)a.x / d), (float)((double)a.y / d), (float)((double)a.z / d), (float)((double)a.w / d));
}
// GammaCurve::Encode(aaGamma, RGBTColour) for a power-law curve: transm is not encoded (colourspace.h:163-169, colourspace.cpp:310)
__device__ __forceinline__ float4 px_encode(float4 a, const AAParams& aa)
{
    if (aa.neutral) return a;
    return make_float4(powf(fmaxf(a.x, 0.0f), aa.enc_gamma), powf(fmaxf(a.y, 0.0f), aa.enc_gamma), powf(fmaxf(a.z, 0.0f), aa.enc_gamma), a.w);
}
// ColourDistanceRGBT (colour.h:616-621, 1221-1224) >= aaThreshold on gamma-encoded colours
__device__ __forceinline__ bool px_differs(float4 a, float4 b, const AAParams& aa)
{
    const float4 ea = px_encode(a, aa), eb = px_encode(b, aa);
    const float dist = fabsf(ea.x - eb.x) + fabsf(ea.y - eb.y) + fabsf(ea.z - eb.z) + fabsf(ea.w - eb.w);
    return (double)dist >= aa.threshold;
}

__device__ __forceinline__ uint32_t find_rect(const uint32_t* off, uint32_t n_rects, uint32_t i)
{
    uint32_t lo = 0, hi = n_rects;
    while (hi - lo > 1) { uint32_t mid = (lo + hi) >> 1; if (off[mid] <= i) lo = mid; else hi = mid; }
    return lo;
}

// ---- method 1 -----------------------------------------------------------------------------------------------
// frame sample j of rectangle r: the w pixels above the rectangle, then the h pixels left of it (tracetask.cpp:534-566)
__global__ void k_aa1_frame_coords(AALayout L, double2* coords)
{
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < L.n_frame; j += gridDim.x * blockDim.x) {
        const uint32_t r = find_rect(L.frame_off, L.n_rects, j);
        const pvgpu_rect rc = L.rects[r];
        const uint32_t k = j - L.frame_off[r], w = (uint32_t)(rc.right - rc.left + 1);
        if (k < w) coords[j] = make_double2((double)(rc.left + (int)k) + 0.5, (double)rc.top - 0.5);
        else coords[j] = make_double2((double)rc.left - 0.5, (double)(rc.top + (int)(k - w)) + 0.5);
    }
}

// c0-based candidate test: any of the four threshold tests the sequential walk can apply to this pixel with
// un-supersampled colours on both sides
__global__ void k_aa1_candidates(AALayout L, AAParams aa, const float4* accum, int32_t* s_slot, uint32_t* cand_list, unsigned int* n_cand)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < L.n_px; i += gridDim.x * blockDim.x) {
        const uint32_t r = find_rect(L.rect_off, L.n_rects, i);
        const pvgpu_rect rc = L.rects[r];
        const uint32_t k = i - L.rect_off[r], w = (uint32_t)(rc.right - rc.left + 1), h = (uint32_t)(rc.bottom - rc.top + 1);
        const uint32_t x = k % w, y = k / w;
        const uint32_t fbase = L.n_px + L.frame_off[r];
        const float4 c = accum[i];
        const float4 left = x ? accum[i - 1] : accum[fbase + w + y];
        const float4 top = y ? accum[i - w] : accum[fbase + x];
        bool cand = px_differs(left, c, aa) || px_differs(top, c, aa);
        if (!cand && x + 1 < w) cand = px_differs(c, accum[i + 1], aa);
        if (!cand && y + 1 < h) cand = px_differs(c, accum[i + w], aa);
        int32_t slot = -1;
        if (cand) {
            const unsigned int idx = atomicAdd(n_cand, 1u);
            cand_list[idx] = i;
            slot = (int32_t)(L.s_base + idx);
        }
        s_slot[i] = slot;
    }
}

// SupersampleOnePixel's sample positions (tracetask.cpp:860-885) for candidates [first, first + n): the (xx, yy)
// offsets come from the host, which ran the reference's floating-point loop once.
__global__ void k_aa1_sample_coords(AALayout L, AAParams aa, const uint16_t* hash, const uint32_t* cand_list, uint32_t first, uint32_t n,
                                    const double2* offsets, uint32_t n_off, double2* coords, uint32_t* slots)
{
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < n * n_off; t += gridDim.x * blockDim.x) {
        const uint32_t c = first + t / n_off, k = t % n_off;
        const uint32_t i = cand_list[c];
        const uint32_t r = find_rect(L.rect_off, L.n_rects, i);
        const pvgpu_rect rc = L.rects[r];
        const uint32_t kk = i - L.rect_off[r], w = (uint32_t)(rc.right - rc.left + 1);
        const double x = (double)(rc.left + (int)(kk % w)), y = (double)(rc.top + (int)(kk / w));
        const double xx = offsets[k].x, yy = offsets[k].y;
        double sx = x + 0.5 + xx, sy = y + 0.5 + yy;
        if (aa.jitter_scale > 0.0) {
            double rx, ry;
            jitter2d(hash, x + xx, y + yy, rx, ry);
            sx = x + 0.5 + xx + (rx * aa.jitter_scale);
            sy = y + 0.5 + yy + (ry * aa.jitter_scale);
        }
        coords[t] = make_double2(sx, sy);
        slots[t] = L.s_base + c;
    }
}

// The reference's sequential walk over one rectangle (tracetask.cpp:568-592 + NonAdaptiveSupersamplingForOnePixel
// :838-858), one thread per rectangle.  out[i] doubles as the walk's pixel state, flag[i] as SmartBlock's flags.
__global__ void k_aa1_decide(AALayout L, AAParams aa, const float4* accum, int32_t* s_slot, uint32_t* cand_list, unsigned int* n_cand,
                             float4* out, uint8_t* flag, unsigned int* n_supersampled)
{
    const double denom = (double)(aa.depth * aa.depth + 1);
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < L.n_rects; r += gridDim.x * blockDim.x) {
        const pvgpu_rect rc = L.rects[r];
        const uint32_t w = (uint32_t)(rc.right - rc.left + 1), h = (uint32_t)(rc.bottom - rc.top + 1);
        const uint32_t base = L.rect_off[r], fbase = L.n_px + L.frame_off[r];
        unsigned int n_ss = 0;
        // S(p) = (c0 + sum of the aaDepth^2 samples) / (aaDepth^2 + 1); queues p when its samples have not been traced yet
        auto supersampled = [&](uint32_t p, float4 c0) -> float4 {
            int32_t slot = s_slot[p];
            if (slot < 0) {
                const unsigned int idx = atomicAdd(n_cand, 1u);
                cand_list[idx] = p;
                s_slot[p] = (int32_t)(L.s_base + idx);
                return c0;                                    // stand-in; the walk is replayed once the samples exist
            }
            n_ss++;
            return px_div(px_add(c0, accum[slot]), denom);
        };
        for (uint32_t y = 0; y < h; y++) {
            for (uint32_t x = 0; x < w; x++) {
                const uint32_t i = base + y * w + x;
                float4 cur = accum[i];
                const float4 leftcol = x ? out[i - 1] : accum[fbase + w + y];
                const float4 topcol = y ? out[i - w] : accum[fbase + x];
                const bool leftdiff = px_differs(leftcol, cur, aa), topdiff = px_differs(topcol, cur, aa);
                const bool sampleleft = x && !flag[i - 1] && leftdiff;       // frame pixels are flagged "already supersampled"
                const bool sampletop = y && !flag[i - w] && topdiff;
                const bool samplecurrent = leftdiff || topdiff;
                if (sampleleft) { out[i - 1] = supersampled(i - 1, leftcol); flag[i - 1] = 1; }
                if (sampletop) { out[i - w] = supersampled(i - w, topcol); flag[i - w] = 1; }
                if (samplecurrent) cur = supersampled(i, cur);
                out[i] = cur;
                flag[i] = samplecurrent ? 1 : 0;
            }
        }
        if (n_ss) atomicAdd(n_supersampled, n_ss);
    }
}

// ---- method 2 -----------------------------------------------------------------------------------------------
// corner sample j of rectangle r: the (w + 1) x (h + 1) integer pixel corners (tracetask.cpp:619-629)
__global__ void k_aa2_corner_coords(AALayout L, double2* coords)
{
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < L.n_corner; j += gridDim.x * blockDim.x) {
        const uint32_t r = find_rect(L.corner_off, L.n_rects, j);
        const pvgpu_rect rc = L.rects[r];
        const uint32_t k = j - L.corner_off[r], w1 = (uint32_t)(rc.right - rc.left + 2);
        coords[j] = make_double2((double)(rc.left + (int)(k % w1)), (double)(rc.top + (int)(k / w1)));
    }
}

struct PixelCorners { uint32_t c00, c20, c02, c22; double x, y; };
__device__ __forceinline__ PixelCorners pixel_corners(const AALayout& L, uint32_t i)
{
    const uint32_t r = find_rect(L.rect_off, L.n_rects, i);
    const pvgpu_rect rc = L.rects[r];
    const uint32_t k = i - L.rect_off[r], w = (uint32_t)(rc.right - rc.left + 1);
    const uint32_t x = k % w, y = k / w, cb = L.corner_off[r] + y * (w + 1) + x;
    PixelCorners pc;
    pc.c00 = cb; pc.c20 = cb + 1; pc.c02 = cb + (w + 1); pc.c22 = cb + (w + 1) + 1;
    pc.x = (double)(rc.left + (int)x); pc.y = (double)(rc.top + (int)y);
    return pc;
}

__device__ __forceinline__ bool square_fires(float4 a, float4 b, float4 c, float4 d, const AAParams& aa)
{
    // the six pairwise tests of SubdivideOnePixel (tracetask.cpp:912-918), on gamma-encoded colours
    const float4 ea = px_encode(a, aa), eb = px_encode(b, aa), ec = px_encode(c, aa), ed = px_encode(d, aa);
    auto dist = [](float4 p, float4 q) { return (double)(fabsf(p.x - q.x) + fabsf(p.y - q.y) + fabsf(p.z - q.z) + fabsf(p.w - q.w)); };
    const double t = aa.threshold;
    return dist(ea, eb) >= t || dist(ea, ec) >= t || dist(ea, ed) >= t || dist(eb, ec) >= t || dist(eb, ed) >= t || dist(ec, ed) >= t;
}

// pixels whose top-level square fires get a sample buffer (tracetask.cpp:912 with level = aaDepth - 1)
__global__ void k_aa2_mark(AALayout L, AAParams aa, const float4* accum, int32_t* act_idx, uint32_t* act_list, unsigned int* n_active)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < L.n_px; i += gridDim.x * blockDim.x) {
        const PixelCorners pc = pixel_corners(L, i);
        int32_t a = -1;
        if (aa.depth > 1 && square_fires(accum[pc.c00], accum[pc.c02], accum[pc.c20], accum[pc.c22], aa)) {
            a = (int32_t)atomicAdd(n_active, 1u);
            act_list[a] = i;
        }
        act_idx[i] = a;
    }
}

// sample (ix, iy) of the buffer of active pixel a; the four buffer corners are the pixel-corner samples
struct SubBuf {
    const float4* accum; uint32_t base; uint32_t n1, S; PixelCorners pc;
    __device__ __forceinline__ float4 get(uint32_t ix, uint32_t iy) const
    {
        if ((ix == 0 || ix == S) && (iy == 0 || iy == S)) return accum[ix == 0 ? (iy == 0 ? pc.c00 : pc.c02) : (iy == 0 ? pc.c20 : pc.c22)];
        return accum[base + iy * n1 + ix];
    }
};

struct Square { double x, y, d; uint32_t bx, by, bs; int level; };
__device__ __forceinline__ Square child_square(const Square& q, int k)
{
    // order of the reference's four recursive calls: (x-,y-), (x-,y+), (x+,y-), (x+,y+)   (tracetask.cpp:1036-1066)
    const double d2 = q.d * 0.5;
    const uint32_t half = q.bs / 2;
    Square c;
    c.d = d2; c.bs = half; c.level = q.level - 1;
    c.x = (k & 2) ? q.x + d2 : q.x - d2;
    c.y = (k & 1) ? q.y + d2 : q.y - d2;
    c.bx = (k & 2) ? q.bx + half : q.bx;
    c.by = (k & 1) ? q.by + half : q.by;
    return c;
}

// round `target`: squares `target` levels below the pixel that fire request their five new samples
// (tracetask.cpp:933-1028); one thread per active pixel
__global__ void k_aa2_expand(AALayout L, AAParams aa, const uint16_t* hash, const float4* accum, const uint32_t* act_list, uint32_t n_active,
                             int target, uint32_t* sampled, double2* coords, uint32_t* slots, unsigned int* n_samples, uint32_t cap)
{
    const uint32_t S = 1u << aa.depth, n1 = S + 1, words = (n1 * n1 + 31) / 32;
    for (uint32_t a = blockIdx.x * blockDim.x + threadIdx.x; a < n_active; a += gridDim.x * blockDim.x) {
        SubBuf buf;
        buf.accum = accum; buf.base = L.s_base + a * n1 * n1; buf.n1 = n1; buf.S = S;
        buf.pc = pixel_corners(L, act_list[a]);
        uint32_t* bits = sampled + (size_t)a * words;
        Square st[48];
        int depth_of[48];
        // the samples this pixel requests are collected first and appended as ONE contiguous run, so that the camera rays
        // of a pixel (and of the neighbouring pixels of its warp) stay together in the wave
        double2 lc[80];
        uint32_t ls[80];
        int ln = 0;
        int sp = 0;
        st[0].x = buf.pc.x; st[0].y = buf.pc.y; st[0].d = 0.5; st[0].bx = 0; st[0].by = 0; st[0].bs = S; st[0].level = aa.depth - 1;
        depth_of[0] = 0;
        sp = 1;
        while (sp > 0) {
            const Square q = st[--sp];
            const int dep = depth_of[sp];
            if (q.level <= 0) continue;
            if (!square_fires(buf.get(q.bx, q.by), buf.get(q.bx, q.by + q.bs), buf.get(q.bx + q.bs, q.by), buf.get(q.bx + q.bs, q.by + q.bs), aa)) continue;
            if (dep < target) {
                for (int k = 0; k < 4; k++) { st[sp] = child_square(q, k); depth_of[sp] = dep + 1; sp++; }
                continue;
            }
            const uint32_t half = q.bs / 2;
            const uint32_t ix[5] = { q.bx, q.bx + half, q.bx + q.bs, q.bx + half, q.bx + half };
            const uint32_t iy[5] = { q.by + half, q.by, q.by + half, q.by + q.bs, q.by + half };
            const double ox[5] = { -q.d, 0.0, q.d, 0.0, 0.0 };
            const double oy[5] = { 0.0, -q.d, 0.0, q.d, 0.0 };
            for (int k = 0; k < 5; k++) {
                const uint32_t bit = iy[k] * n1 + ix[k];
                if (bits[bit >> 5] & (1u << (bit & 31))) continue;          // buffer.Sampled(...)
                bits[bit >> 5] |= 1u << (bit & 31);
                double sx = q.x + 0.5 + ox[k], sy = q.y + 0.5 + oy[k];
                if (aa.jitter_scale > 0.0) {
                    double rx, ry;
                    jitter2d(hash, q.x + ox[k], q.y + oy[k], rx, ry);
                    sx = q.x + 0.5 + ox[k] + (rx * aa.jitter_scale);
                    sy = q.y + 0.5 + oy[k] + (ry * aa.jitter_scale);
                }
                if (ln == 80) {       // flush a full run (only deep levels can exceed it)
                    const unsigned int base = atomicAdd(n_samples, 80u);
                    for (int m = 0; m < 80; m++) if (base + m < cap) { coords[base + m] = lc[m]; slots[base + m] = ls[m]; }
                    ln = 0;
                }
                lc[ln] = make_double2(sx, sy);
                ls[ln] = buf.base + bit;
                ln++;
            }
        }
        if (ln) {
            const unsigned int base = atomicAdd(n_samples, (unsigned int)ln);
            for (int m = 0; m < ln; m++) if (base + m < cap) { coords[base + m] = lc[m]; slots[base + m] = ls[m]; }
        }
    }
}

// SubdivideOnePixel's result for every pixel (tracetask.cpp:892-1074), iterative post-order over the sample buffer
//   group == nullptr: every pixel; the sample buffer of subdividing pixel number a starts at s_base + a * n1^2, or - skip_active - only
//   the pixels that do not subdivide.   group != nullptr: the n_group subdividing pixels listed there, buffers numbered within the group
//   (deep subdivision levels: the buffers of all subdividing pixels do not fit at once, so the pixels are worked off in groups)
__global__ void k_aa2_resolve(AALayout L, AAParams aa, const float4* accum, const int32_t* act_idx, float4* out,
                              const uint32_t* group, uint32_t n_group, int skip_active)
{
    const uint32_t S = 1u << aa.depth, n1 = S + 1;
    const uint32_t n_items = group ? n_group : L.n_px;
    for (uint32_t it = blockIdx.x * blockDim.x + threadIdx.x; it < n_items; it += gridDim.x * blockDim.x) {
        const uint32_t i = group ? group[it] : it;
        SubBuf buf;
        buf.accum = accum; buf.n1 = n1; buf.S = S;
        buf.pc = pixel_corners(L, i);
        const int32_t a = group ? (int32_t)it : act_idx[i];
        if (a < 0) {
            out[i] = px_div(px_add(px_add(px_add(accum[buf.pc.c00], accum[buf.pc.c02]), accum[buf.pc.c20]), accum[buf.pc.c22]), 4.0);
            continue;
        }
        if (skip_active) continue;
        buf.base = L.s_base + (uint32_t)a * n1 * n1;
        struct Frame { Square q; int next; float4 acc; };
        Frame fr[12];
        int sp = 0;
        fr[0].q.x = buf.pc.x; fr[0].q.y = buf.pc.y; fr[0].q.d = 0.5; fr[0].q.bx = 0; fr[0].q.by = 0; fr[0].q.bs = S; fr[0].q.level = aa.depth - 1;
        fr[0].next = -1;
        float4 ret = make_float4(0.f, 0.f, 0.f, 0.f);
        for (;;) {
            Frame& f = fr[sp];
            if (f.next < 0) {
                const Square& q = f.q;
                const float4 c00 = buf.get(q.bx, q.by), c02 = buf.get(q.bx, q.by + q.bs), c20 = buf.get(q.bx + q.bs, q.by), c22 = buf.get(q.bx + q.bs, q.by + q.bs);
                if (!(q.level > 0 && square_fires(c00, c02, c20, c22, aa))) {
                    ret = px_div(px_add(px_add(px_add(c00, c02), c20), c22), 4.0);
                    if (--sp < 0) break;
                    continue;
                }
                f.next = 0;
            } else {
                f.acc = (f.next == 1) ? ret : px_add(f.acc, ret);          // ((r00 + r01) + r10) + r11
                if (f.next == 4) {
                    ret = px_div(f.acc, 4.0);
                    if (--sp < 0) break;
                    continue;
                }
            }
            const int k = f.next++;
            fr[sp + 1].q = child_square(f.q, k);
            fr[sp + 1].next = -1;
            sp++;
        }
        out[i] = ret;
    }
}

// copies method-0-style results: out[i] = accum[i]
__global__ void k_copy_pixels(const float4* accum, uint32_t n, float4* out)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = accum[i];
}

// ---- launchers -------------------------------------------------------------------------------------------------
void launch_aa1_frame_coords(const AALayout& L, double2* coords, cudaStream_t st)
{ if (L.n_frame) k_aa1_frame_coords<<<grid_for(L.n_frame, 256, 8), 256, 0, st>>>(L, coords); }
void launch_aa1_candidates(const AALayout& L, const AAParams& aa, const float4* accum, int32_t* s_slot, uint32_t* cand_list, unsigned int* n_cand, cudaStream_t st)
{ k_aa1_candidates<<<grid_for(L.n_px, 256, 8), 256, 0, st>>>(L, aa, accum, s_slot, cand_list, n_cand); }
void launch_aa1_sample_coords(const AALayout& L, const AAParams& aa, const uint16_t* hash, const uint32_t* cand_list, uint32_t first, uint32_t n,
                              const double2* offsets, uint32_t n_off, double2* coords, uint32_t* slots, cudaStream_t st)
{ if (n) k_aa1_sample_coords<<<grid_for(n * n_off, 256, 8), 256, 0, st>>>(L, aa, hash, cand_list, first, n, offsets, n_off, coords, slots); }
void launch_aa1_decide(const AALayout& L, const AAParams& aa, const float4* accum, int32_t* s_slot, uint32_t* cand_list, unsigned int* n_cand,
                       float4* out, uint8_t* flag, unsigned int* n_supersampled, cudaStream_t st)
{ k_aa1_decide<<<grid_for(L.n_rects, 32, 16), 32, 0, st>>>(L, aa, accum, s_slot, cand_list, n_cand, out, flag, n_supersampled); }
void launch_aa2_corner_coords(const AALayout& L, double2* coords, cudaStream_t st)
{ k_aa2_corner_coords<<<grid_for(L.n_corner, 256, 8), 256, 0, st>>>(L, coords); }
void launch_aa2_mark(const AALayout& L, const AAParams& aa, const float4* accum, int32_t* act_idx, uint32_t* act_list, unsigned int* n_active, cudaStream_t st)
{ k_aa2_mark<<<grid_for(L.n_px, 256, 8), 256, 0, st>>>(L, aa, accum, act_idx, act_list, n_active); }
void launch_aa2_expand(const AALayout& L, const AAParams& aa, const uint16_t* hash, const float4* accum, const uint32_t* act_list, uint32_t n_active,
                       int target, uint32_t* sampled, double2* coords, uint32_t* slots, unsigned int* n_samples, uint32_t cap, cudaStream_t st)
{ if (n_active) k_aa2_expand<<<grid_for(n_active, 64, 16), 64, 0, st>>>(L, aa, hash, accum, act_list, n_active, target, sampled, coords, slots, n_samples, cap); }
void launch_aa2_resolve(const AALayout& L, const AAParams& aa, const float4* accum, const int32_t* act_idx, float4* out, cudaStream_t st,
                        const uint32_t* group, uint32_t n_group, int skip_active)
{ k_aa2_resolve<<<grid_for(group ? n_group : L.n_px, 128, 8), 128, 0, st>>>(L, aa, accum, act_idx, out, group, n_group, skip_active); }

}  // namespace pvgpu
