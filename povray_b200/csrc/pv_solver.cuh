// Device root solver for polynomials of degree <= 4: the behaviour of Solve_Polynomial
// (source/core/math/polynomialsolver.cpp:1585-1729) with its closed-form quadratic (:810-861),
// trigonometric/Cardano cubic (:903-978), Vieta quartic (:1325-1450), the `difficult_coeffs` switch
// (:1075-1109, the variant compiled without USE_NEW_DIFFICULT_COEFFS) and the Sturm-sequence
// bisection / regula-falsi path (:171-775, 1481-1523).  Recursion in sbisect is replaced by an
// explicit interval stack that visits sub-intervals in the same (left first) order.
#pragma once
#include "pv_common.cuh"

namespace pvgpu {

#define PV_FUDGE_FACTOR1   1.0e12
#define PV_SMALL_ENOUGH    1.0e-10
#define PV_RELERROR        1.0e-12
#define PV_MAX_ITERATIONS  50
#define PV_TWO_M_PI_3      2.0943951023931954923084
#define PV_FOUR_M_PI_3     4.1887902047863909846168
#define PV_POLY_MAX_ORDER  4

__device__ inline int solve_quadratic(const double* x, double* y)
{
    double a = x[0], b = -x[1], c = x[2];
    if (a == 0.0) {
        if (b == 0.0) return 0;
        y[0] = c / b;
        return 1;
    }
    b /= a;
    c /= a;
    a = 1.0;
    double d = b * b - 4.0 * a * c;
    if ((d > -PV_SMALL_ENOUGH) && (d < PV_SMALL_ENOUGH)) {
        y[0] = 0.5 * b / a;
        return 1;
    }
    if (d < 0.0) return 0;
    d = sqrt(d);
    double t = 2.0 * a;
    y[0] = (b + d) / t;
    y[1] = (b - d) / t;
    return 2;
}

// (the solver is one of the largest pieces of the heavy kernels' hot path, which is bound by instruction fetch: its building blocks are
//  single out-of-line copies with rolled loops - measured on config 4, profiles/README.md - instead of inlined, unrolled ones)
static __device__ __noinline__ int solve_cubic(const double* x, double* y)
{
    double a0 = x[0], a1, a2, a3;
    if (a0 == 0.0) return solve_quadratic(&x[1], y);
    if (a0 != 1.0) { a1 = x[1] / a0; a2 = x[2] / a0; a3 = x[3] / a0; }
    else { a1 = x[1]; a2 = x[2]; a3 = x[3]; }
    double A2 = a1 * a1;
    double Q = (A2 - 3.0 * a2) / 9.0;
    double R = (a1 * (A2 - 4.5 * a2) + 13.5 * a3) / 27.0;
    double Q3 = Q * Q * Q;
    double R2 = R * R;
    double d = Q3 - R2;
    double an = a1 / 3.0;
    if (d >= 0.0) {
        d = R / sqrt(Q3);
        double theta = acos(d) / 3.0;
        double sQ = -2.0 * sqrt(Q);
        y[0] = sQ * cos(theta) - an;
        y[1] = sQ * cos(theta + PV_TWO_M_PI_3) - an;
        y[2] = sQ * cos(theta + PV_FOUR_M_PI_3) - an;
        return 3;
    }
    double sQ = pow(sqrt(R2 - Q3) + fabs(R), 1.0 / 3.0);
    if (R < 0) y[0] = (sQ + Q / sQ) - an;
    else       y[0] = -(sQ + Q / sQ) - an;
    return 1;
}

__device__ inline int solve_quartic(const double* x, double* results)
{
    double cubic[4], roots[3];
    double c0 = x[0], c1, c2, c3, c4;
    if (c0 != 1.0) { c1 = x[1] / c0; c2 = x[2] / c0; c3 = x[3] / c0; c4 = x[4] / c0; }
    else { c1 = x[1]; c2 = x[2]; c3 = x[3]; c4 = x[4]; }
    double c12 = c1 * c1;
    double p = -0.375 * c12 + c2;
    double q = 0.125 * c12 * c1 - 0.5 * c1 * c2 + c3;
    double r = -0.01171875 * c12 * c12 + 0.0625 * c12 * c2 - 0.25 * c1 * c3 + c4;
    cubic[0] = 1.0;
    cubic[1] = -0.5 * p;
    cubic[2] = -r;
    cubic[3] = 0.5 * r * p - 0.125 * q * q;
    int i = solve_cubic(cubic, roots);
    if (i <= 0) return 0;
    double z = roots[0];
    double d1 = 2.0 * z - p, d2;
    if (d1 < 0.0) {
        if (d1 > -PV_SMALL_ENOUGH) d1 = 0.0;
        else return 0;
    }
    if (d1 < PV_SMALL_ENOUGH) {
        d2 = z * z - r;
        if (d2 < 0.0) return 0;
        d2 = sqrt(d2);
    } else {
        d1 = sqrt(d1);
        d2 = 0.5 * q / d1;
    }
    double q1 = d1 * d1;
    double q2 = -0.25 * c1;
    i = 0;
    p = q1 - 4.0 * (z - d2);
    if (p == 0) results[i++] = -0.5 * d1 - q2;
    else if (p > 0) {
        p = sqrt(p);
        results[i++] = -0.5 * (d1 + p) + q2;
        results[i++] = -0.5 * (d1 - p) + q2;
    }
    p = q1 - 4.0 * (z + d2);
    if (p == 0) results[i++] = 0.5 * d1 - q2;
    else if (p > 0) {
        p = sqrt(p);
        results[i++] = 0.5 * (d1 + p) + q2;
        results[i++] = 0.5 * (d1 - p) + q2;
    }
    return i;
}

// difficult_coeffs, as compiled (no USE_NEW_DIFFICULT_COEFFS): note `biggest` keeps the SIGNED value.
__device__ inline int difficult_coeffs(int n, const double* x)
{
    double biggest = 0.0;
    #pragma unroll 1
    for (int i = 0; i <= n; i++)
        if (fabs(x[i]) > biggest) biggest = x[i];
    if (biggest == 0.0) return 0;
    #pragma unroll 1
    for (int i = 0; i <= n; i++)
        if (x[i] != 0.0)
            if (fabs(biggest / x[i]) > PV_FUDGE_FACTOR1) return 1;
    return 0;
}

struct Poly { int ord; double coef[PV_POLY_MAX_ORDER + 1]; };

__device__ __forceinline__ double polyeval(double x, int n, const double* c)
{
    double val = c[n];
    #pragma unroll 1
    for (int i = n - 1; i >= 0; i--) val = val * x + c[i];
    return val;
}

__device__ inline int modp(const Poly* u, const Poly* v, Poly* r)
{
    *r = *u;
    if (v->coef[v->ord] < 0.0) {
        #pragma unroll 1
        for (int k = u->ord - v->ord - 1; k >= 0; k -= 2) r->coef[k] = -r->coef[k];
        #pragma unroll 1
        for (int k = u->ord - v->ord; k >= 0; k--) {
            #pragma unroll 1
            for (int j = v->ord + k - 1; j >= k; j--)
                r->coef[j] = -r->coef[j] - r->coef[v->ord + k] * v->coef[j - k];
        }
    } else {
        #pragma unroll 1
        for (int k = u->ord - v->ord; k >= 0; k--) {
            #pragma unroll 1
            for (int j = v->ord + k - 1; j >= k; j--)
                r->coef[j] -= r->coef[v->ord + k] * v->coef[j - k];
        }
    }
    int k = v->ord - 1;
    while (k >= 0 && fabs(r->coef[k]) < PV_SMALL_ENOUGH) { r->coef[k] = 0.0; k--; }
    r->ord = (k < 0) ? 0 : k;
    return r->ord;
}

__device__ inline int buildsturm(int ord, Poly* sseq)
{
    sseq[0].ord = ord;
    sseq[1].ord = ord - 1;
    double f = fabs(sseq[0].coef[ord] * ord);
    #pragma unroll 1
    for (int i = 1; i <= ord; i++) sseq[1].coef[i - 1] = sseq[0].coef[i] * i / f;
    int sp = 2;
    #pragma unroll 1
    for (; modp(&sseq[sp - 2], &sseq[sp - 1], &sseq[sp]); sp++) {
        f = -fabs(sseq[sp].coef[sseq[sp].ord]);
        #pragma unroll 1
        for (int k = sseq[sp].ord; k >= 0; k--) sseq[sp].coef[k] /= f;
    }
    sseq[sp].coef[0] = -sseq[sp].coef[0];
    return sp;
}

__device__ inline int visible_roots(int np, const Poly* sseq)
{
    int atposinf = 0, atzero = 0;
    double lf = sseq[0].coef[sseq[0].ord];
    #pragma unroll 1
    for (int s = 1; s <= np; s++) {
        double f = sseq[s].coef[sseq[s].ord];
        if (lf == 0.0 || lf * f < 0) atposinf++;
        lf = f;
    }
    lf = sseq[0].coef[0];
    #pragma unroll 1
    for (int s = 1; s <= np; s++) {
        double f = sseq[s].coef[0];
        if (lf == 0.0 || lf * f < 0) atzero++;
        lf = f;
    }
    return atzero - atposinf;
}

static __device__ __noinline__ int numchanges(int np, const Poly* sseq, double a)
{
    int changes = 0;
    double lf = polyeval(a, sseq[0].ord, sseq[0].coef);
    #pragma unroll 1
    for (int s = 1; s <= np; s++) {
        double f = polyeval(a, sseq[s].ord, sseq[s].coef);
        if (lf == 0.0 || lf * f < 0) changes++;
        lf = f;
    }
    return changes;
}

static __device__ __noinline__ int regula_falsa(int order, const double* coef, double a, double b, double* val)
{
    double fa = polyeval(a, order, coef), fb = polyeval(b, order, coef);
    if (fa * fb > 0.0) return 0;
    if (fabs(fa) < PV_SMALL_ENOUGH) { *val = a; return 1; }
    if (fabs(fb) < PV_SMALL_ENOUGH) { *val = b; return 1; }
    double lfx = fa;
    #pragma unroll 1
    for (int its = 0; its < PV_MAX_ITERATIONS; its++) {
        double x = (fb * a - fa * b) / (fb - fa);
        double fx = polyeval(x, order, coef);
        if (fabs(x) > PV_RELERROR) {
            if (fabs(fx / x) < PV_RELERROR) { *val = x; return 1; }
        } else if (fabs(fx) < PV_RELERROR) { *val = x; return 1; }
        if (fa < 0) {
            if (fx < 0) { a = x; fa = fx; if ((lfx * fx) > 0) fb /= 2; }
            else        { b = x; fb = fx; if ((lfx * fx) > 0) fa /= 2; }
        } else {
            if (fx < 0) { b = x; fb = fx; if ((lfx * fx) > 0) fa /= 2; }
            else        { a = x; fa = fx; if ((lfx * fx) > 0) fb /= 2; }
        }
        if (fabs(b - a) < PV_RELERROR) { *val = x; return 1; }
        lfx = fx;
    }
    return 0;
}

// sbisect with an explicit stack.  The reference recursion returns the roots of the left half before
// those of the right half; pushing the right half first reproduces that order in `roots`.
__device__ inline int sbisect(int np, const Poly* sseq, double min0, double max0, int atmin0, int atmax0, double* roots)
{
    struct Iv { double lo, hi; int atlo, athi; };
    Iv st[PV_POLY_MAX_ORDER + 2];
    int sp = 0, nroots = 0;
    st[sp++] = { min0, max0, atmin0, atmax0 };
    while (sp > 0) {
        Iv iv = st[--sp];
        double min_value = iv.lo, max_value = iv.hi, mid = 0.0;
        int atmin = iv.atlo, atmax = iv.athi;
        if ((atmin - atmax) == 1) {
            if (regula_falsa(sseq[0].ord, sseq[0].coef, min_value, max_value, &roots[nroots])) { nroots++; continue; }
            bool done = false;
            #pragma unroll 1
            for (int its = 0; its < PV_MAX_ITERATIONS; its++) {
                mid = (min_value + max_value) / 2;
                int atmid = numchanges(np, sseq, mid);
                if ((atmid < atmax) || (atmid > atmin)) { done = true; break; }          // returns 0 roots
                if (fabs(mid) > PV_RELERROR) {
                    if (fabs((max_value - min_value) / mid) < PV_RELERROR) { roots[nroots++] = mid; done = true; break; }
                } else if (fabs(max_value - min_value) < PV_RELERROR) { roots[nroots++] = mid; done = true; break; }
                if ((atmin - atmid) == 0) min_value = mid; else max_value = mid;
            }
            if (!done) roots[nroots++] = mid;
            continue;
        }
        bool done = false;
        #pragma unroll 1
        for (int its = 0; its < PV_MAX_ITERATIONS; its++) {
            mid = (min_value + max_value) / 2;
            int atmid = numchanges(np, sseq, mid);
            if ((atmid < atmax) || (atmid > atmin)) { done = true; break; }
            if (fabs(mid) > PV_RELERROR) {
                if (fabs((max_value - min_value) / mid) < PV_RELERROR) { roots[nroots++] = mid; done = true; break; }
            } else if (fabs(max_value - min_value) < PV_RELERROR) { roots[nroots++] = mid; done = true; break; }
            int n1 = atmin - atmid, n2 = atmid - atmax;
            if ((n1 != 0) && (n2 != 0)) {
                st[sp++] = { mid, max_value, atmid, atmax };     // right half: visited second
                st[sp++] = { min_value, mid, atmin, atmid };     // left half: visited first
                done = true;
                break;
            }
            if (n1 == 0) min_value = mid; else max_value = mid;
        }
        if (!done) roots[nroots++] = mid;
    }
    return nroots;
}

static __device__ __noinline__ int polysolve(int order, const double* coeffs, double* roots)
{
    Poly sseq[PV_POLY_MAX_ORDER + 1];
    #pragma unroll 1
    for (int i = 0; i <= order; i++) sseq[0].coef[order - i] = coeffs[i] / coeffs[0];
    int np = buildsturm(order, sseq);
    if (visible_roots(np, sseq) == 0) return 0;
    double min_value = 0.0, max_value = PV_MAX_DISTANCE;
    int atmin = numchanges(np, sseq, min_value);
    int atmax = numchanges(np, sseq, max_value);
    if (atmin - atmax == 0) return 0;
    return sbisect(np, sseq, min_value, max_value, atmin, atmax, roots);
}

// Solve_Polynomial for n <= 4.
// (out of line: torus, poly and blob all come here; three inlined copies were 40 % of the heavy kernels' code)
static __device__ __noinline__ int solve_polynomial(int n, const double* c0, double* r, int sturm, double epsilon)
{
    int roots = 0, i = 0;
    while ((i < n) && (fabs(c0[i]) < PV_SMALL_ENOUGH)) i++;
    n -= i;
    const double* c = &c0[i];
    switch (n) {
        case 0: break;
        case 1: if (c[0] != 0.0) r[roots++] = -c[1] / c[0]; break;
        case 2: roots = solve_quadratic(c, r); break;
        case 3:
            if (epsilon > 0.0 && (c[2] != 0.0) && (fabs(c[3] / c[2]) < epsilon)) { roots = solve_quadratic(c, r); break; }
            roots = sturm ? polysolve(3, c, r) : solve_cubic(c, r);
            break;
        case 4:
            if (epsilon > 0.0 && (c[3] != 0.0) && (fabs(c[4] / c[3]) < epsilon)) {
                roots = sturm ? polysolve(3, c, r) : solve_cubic(c, r);
                break;
            }
            if (difficult_coeffs(4, c)) sturm = 1;
            roots = sturm ? polysolve(4, c, r) : solve_quartic(c, r);
            break;
    }
    return roots;
}

}  // namespace pvgpu
