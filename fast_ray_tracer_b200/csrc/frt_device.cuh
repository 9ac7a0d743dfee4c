/*
 * frt_device.cuh -- device-side scene model and ray/shape arithmetic of the B200 render core.
 *
 * Everything here is FP64 like the reference (typedef double Point[4], linalg.h:31-33): the parity gate is
 * "within 1 LSB of sRGB-8 on >= 99.9 % of pixels" against a reference whose epsilon logic (EPSILON 1e-5,
 * t > 0 tests) flips on silhouettes in FP32.  B200 issues FP64 FMAs at half the FP32 rate, so the exact
 * formulation is affordable; see DESIGN.md "precision".
 *
 * Each device function cites the reference function whose arithmetic it restates.
 */
#pragma once

#include <cuda_runtime.h>
#include <math_constants.h>
#include <math.h>
#include <stdint.h>

#include "frt_b200.h"

#define FRT_EPS 0.00001 /* EPSILON, linalg.h:7 */
#define FRT_CSG_CAP 32  /* per-ray CSG interval stack (entries) */

struct DScene {
    const int4 *nodes;        /* 2 x int4 per node: {type, skip, xform, material} {param, csg_op, right, parent} */
    const double *bbox;       /* 6 per node */
    const double *xinv;       /* 12 per xform */
    const double *params;
    const frt_material *mats;
    const frt_pattern *pats;
    const frt_texture *texs;
    const float4 *texels;     /* one per texel: canvas_pixel_at evaluated at upload, linear FP32 RGB (frt_patterns.cuh) */
    const frt_light *lights;
    const double *lpoints;
    const int *roots;
    int n_roots, n_nodes, n_lights, pad;
};

struct Ray {
    double ox, oy, oz, dx, dy, dz;
};

struct NodeA {
    int type, skip, xform, material;
};
struct NodeB {
    int param, csg_op, right, parent;
};

__device__ __forceinline__ NodeA
load_node_a(const DScene &S, int i)
{
    int4 v = __ldg(S.nodes + 2 * i);
    return NodeA{ v.x, v.y, v.z, v.w };
}

__device__ __forceinline__ NodeB
load_node_b(const DScene &S, int i)
{
    int4 v = __ldg(S.nodes + 2 * i + 1);
    return NodeB{ v.x, v.y, v.z, v.w };
}

/* ray_transform (ray.c:14-18) with the composite world->local matrix of the node */
__device__ __forceinline__ Ray
ray_to_local(const DScene &S, int xf, const Ray &w)
{
    if (xf == 0) {
        return w;
    }
    const double *m = S.xinv + 12 * xf;
    Ray r;
    double m0 = __ldg(m + 0), m1 = __ldg(m + 1), m2 = __ldg(m + 2), m3 = __ldg(m + 3);
    r.ox = m0 * w.ox + m1 * w.oy + m2 * w.oz + m3;
    r.dx = m0 * w.dx + m1 * w.dy + m2 * w.dz;
    m0 = __ldg(m + 4), m1 = __ldg(m + 5), m2 = __ldg(m + 6), m3 = __ldg(m + 7);
    r.oy = m0 * w.ox + m1 * w.oy + m2 * w.oz + m3;
    r.dy = m0 * w.dx + m1 * w.dy + m2 * w.dz;
    m0 = __ldg(m + 8), m1 = __ldg(m + 9), m2 = __ldg(m + 10), m3 = __ldg(m + 11);
    r.oz = m0 * w.ox + m1 * w.oy + m2 * w.oz + m3;
    r.dz = m0 * w.dx + m1 * w.dy + m2 * w.dz;
    return r;
}

__device__ __forceinline__ void
point_to_local(const DScene &S, int xf, const double p[3], double out[3])
{
    if (xf == 0) {
        out[0] = p[0];
        out[1] = p[1];
        out[2] = p[2];
        return;
    }
    const double *m = S.xinv + 12 * xf;
    out[0] = __ldg(m + 0) * p[0] + __ldg(m + 1) * p[1] + __ldg(m + 2) * p[2] + __ldg(m + 3);
    out[1] = __ldg(m + 4) * p[0] + __ldg(m + 5) * p[1] + __ldg(m + 6) * p[2] + __ldg(m + 7);
    out[2] = __ldg(m + 8) * p[0] + __ldg(m + 9) * p[1] + __ldg(m + 10) * p[2] + __ldg(m + 11);
}

/* check_axis (cube.c:16-54) / bbox_check_axis (bounding_box.c:124-162): slab interval on one axis */
__device__ __forceinline__ void
slab_axis(double origin, double direction, double lo, double hi, double &t0, double &t1)
{
    double n0 = lo - origin;
    double n1 = hi - origin;
    double a, b;
    if (fabs(direction) >= FRT_EPS) {
        a = n0 / direction;
        b = n1 / direction;
    } else {
        /* numerator * INFINITY; NaN (0 * inf) becomes +inf, or -inf when the numerator is negative (it is not) */
        a = n0 * CUDART_INF;
        if (isnan(a)) {
            a = CUDART_INF;
        }
        b = n1 * CUDART_INF;
        if (isnan(b)) {
            b = CUDART_INF;
        }
    }
    if (a > b) {
        t0 = b;
        t1 = a;
    } else {
        t0 = a;
        t1 = b;
    }
}

/* bounding_box_intersects, bounding_box.c:165-175 */
__device__ __forceinline__ bool
bbox_hit(const DScene &S, int node, const Ray &r)
{
    const double *b = S.bbox + 6 * node;
    double x0, x1, y0, y1, z0, z1;
    slab_axis(r.ox, r.dx, __ldg(b + 0), __ldg(b + 3), x0, x1);
    slab_axis(r.oy, r.dy, __ldg(b + 1), __ldg(b + 4), y0, y1);
    slab_axis(r.oz, r.dz, __ldg(b + 2), __ldg(b + 5), z0, z1);
    double tmin = fmax(fmax(x0, y0), z0);
    double tmax = fmin(fmin(x1, y1), z1);
    return tmin <= tmax;
}

/* ---- fast FP64 reciprocal / reciprocal square root: MUFU seed + Newton steps, no slow-path branch.  Results are
 *      within 1 ulp of the correctly rounded value, which only moves a t value by 1 ulp (never a decision that is
 *      not already a tie).  Used by the shadow traversal, where every shadow ray would otherwise pay for a dozen
 *      IEEE divisions per node. */
__device__ __forceinline__ double
rcp_fast(double x)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    return r;
}

__device__ __forceinline__ double
rsqrt_fast(double x)
{
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double h = 0.5 * x;
    r = r * fma(-h * r, r, 1.5);
    r = r * fma(-h * r, r, 1.5);
    return r;
}

/* 1 / direction per axis for the slab tests; +inf where |direction| < EPSILON, which reproduces the
 * "numerator * INFINITY" branch of check_axis / bbox_check_axis (cube.c:27-33, bounding_box.c:135-141) */
struct InvDir {
    double x, y, z;
};

__device__ __forceinline__ InvDir
inv_dir(const Ray &r)
{
    InvDir v;
    v.x = fabs(r.dx) >= FRT_EPS ? rcp_fast(r.dx) : CUDART_INF;
    v.y = fabs(r.dy) >= FRT_EPS ? rcp_fast(r.dy) : CUDART_INF;
    v.z = fabs(r.dz) >= FRT_EPS ? rcp_fast(r.dz) : CUDART_INF;
    return v;
}

__device__ __forceinline__ void
slab_axis_inv(double origin, double inv, double lo, double hi, double &t0, double &t1)
{
    double a = (lo - origin) * inv;
    double b = (hi - origin) * inv;
    if (isnan(a)) a = CUDART_INF; /* 0 * inf */
    if (isnan(b)) b = CUDART_INF;
    t0 = fmin(a, b);
    t1 = fmax(a, b);
}

/* the slab interval [tmin, tmax] of the node's box (bounding_box_intersects, bounding_box.c:165-175, hits iff tmin <= tmax) */
__device__ __forceinline__ void
bbox_range_inv(const DScene &S, int node, const Ray &r, const InvDir &inv, double &tmin, double &tmax)
{
    const double2 *b = reinterpret_cast<const double2 *>(S.bbox + 6 * node);
    double2 b0 = __ldg(b), b1 = __ldg(b + 1), b2 = __ldg(b + 2); /* min.x min.y | min.z max.x | max.y max.z */
    double x0, x1, y0, y1, z0, z1;
    slab_axis_inv(r.ox, inv.x, b0.x, b1.y, x0, x1);
    slab_axis_inv(r.oy, inv.y, b0.y, b2.x, y0, y1);
    slab_axis_inv(r.oz, inv.z, b1.x, b2.y, z0, z1);
    tmin = fmax(fmax(x0, y0), z0);
    tmax = fmin(fmin(x1, y1), z1);
}

__device__ __forceinline__ bool
bbox_hit_inv(const DScene &S, int node, const Ray &r, const InvDir &inv)
{
    const double2 *b = reinterpret_cast<const double2 *>(S.bbox + 6 * node);
    double2 b0 = __ldg(b), b1 = __ldg(b + 1), b2 = __ldg(b + 2); /* min.x min.y | min.z max.x | max.y max.z */
    double x0, x1, y0, y1, z0, z1;
    slab_axis_inv(r.ox, inv.x, b0.x, b1.y, x0, x1);
    slab_axis_inv(r.oy, inv.y, b0.y, b2.x, y0, y1);
    slab_axis_inv(r.oz, inv.z, b1.x, b2.y, z0, z1);
    return fmax(fmax(x0, y0), z0) <= fmin(fmin(x1, y1), z1);
}

/* ---- Schwarze's quadric / cubic / quartic solver (Graphics Gems; reference src/libs/quartic/Roots3And4.c) */

#define FRT_EQN_EPS 1e-9
__device__ __forceinline__ bool
eqn_is_zero(double x)
{
    return x > -FRT_EQN_EPS && x < FRT_EQN_EPS;
}

/* Roots3And4.c:44-75 */
__device__ __noinline__ int
solve_quadric(const double c[3], double s[2])
{
    double p = c[1] / (2 * c[2]);
    double q = c[0] / c[2];
    double D = p * p - q;
    if (eqn_is_zero(D)) {
        s[0] = -p;
        return 1;
    } else if (D < 0) {
        return 0;
    }
    double sqrt_D = sqrt(D);
    s[0] = sqrt_D - p;
    s[1] = -sqrt_D - p;
    return 2;
}

/* Roots3And4.c:77-149 */
__device__ __noinline__ int
solve_cubic(const double c[4], double s[3])
{
    int num;
    double A = c[2] / c[3];
    double B = c[1] / c[3];
    double C = c[0] / c[3];
    double sq_A = A * A;
    double p = 1.0 / 3 * (-1.0 / 3 * sq_A + B);
    double q = 1.0 / 2 * (2.0 / 27 * A * sq_A - 1.0 / 3 * A * B + C);
    double cb_p = p * p * p;
    double D = q * q + cb_p;

    if (eqn_is_zero(D)) {
        if (eqn_is_zero(q)) {
            s[0] = 0;
            num = 1;
        } else {
            double u = cbrt(-q);
            s[0] = 2 * u;
            s[1] = -u;
            num = 2;
        }
    } else if (D < 0) {
        double phi = 1.0 / 3 * acos(-q / sqrt(-cb_p));
        double t = 2 * sqrt(-p);
        s[0] = t * cos(phi);
        s[1] = -t * cos(phi + M_PI / 3);
        s[2] = -t * cos(phi - M_PI / 3);
        num = 3;
    } else {
        double sqrt_D = sqrt(D);
        double u = cbrt(sqrt_D - q);
        double v = -cbrt(sqrt_D + q);
        s[0] = u + v;
        num = 1;
    }
    double sub = 1.0 / 3 * A;
    for (int i = 0; i < num; ++i) {
        s[i] -= sub;
    }
    return num;
}

/* Roots3And4.c:151-245 */
__device__ __noinline__ int
solve_quartic(const double c[5], double s[4])
{
    double coeffs[4];
    int num;
    double A = c[3] / c[4];
    double B = c[2] / c[4];
    double C = c[1] / c[4];
    double D = c[0] / c[4];
    double sq_A = A * A;
    double p = -3.0 / 8 * sq_A + B;
    double q = 1.0 / 8 * sq_A * A - 1.0 / 2 * A * B + C;
    double r = -3.0 / 256 * sq_A * sq_A + 1.0 / 16 * sq_A * B - 1.0 / 4 * A * C + D;

    if (eqn_is_zero(r)) {
        coeffs[0] = q;
        coeffs[1] = p;
        coeffs[2] = 0;
        coeffs[3] = 1;
        num = solve_cubic(coeffs, s);
        s[num++] = 0;
    } else {
        coeffs[0] = 1.0 / 2 * r * p - 1.0 / 8 * q * q;
        coeffs[1] = -r;
        coeffs[2] = -1.0 / 2 * p;
        coeffs[3] = 1;
        (void)solve_cubic(coeffs, s);
        double z = s[0];
        double u = z * z - r;
        double v = 2 * z - p;
        if (eqn_is_zero(u)) {
            u = 0;
        } else if (u > 0) {
            u = sqrt(u);
        } else {
            return 0;
        }
        if (eqn_is_zero(v)) {
            v = 0;
        } else if (v > 0) {
            v = sqrt(v);
        } else {
            return 0;
        }
        coeffs[0] = z - u;
        coeffs[1] = q < 0 ? -v : v;
        coeffs[2] = 1;
        num = solve_quadric(coeffs, s);
        coeffs[0] = z + u;
        coeffs[1] = q < 0 ? v : -v;
        coeffs[2] = 1;
        num += solve_quadric(coeffs, s + num);
    }
    double sub = 1.0 / 4 * A;
    for (int i = 0; i < num; ++i) {
        s[i] -= sub;
    }
    return num;
}

/* ---- local_intersect of the eight primitives.  Returns the number of t values written (unsorted, negative
 *      values kept, exactly as the reference's per-shape xs lists); uv is written for (smooth) triangles. */

/* toroid_local_intersect, toroid.c:15-53 */
__device__ __noinline__ int
toroid_intersect(const double *prm, const Ray &r, double t[4])
{
    double r1 = __ldg(prm + 0), r2 = __ldg(prm + 1);
    double sum_d_sq = r.dx * r.dx + r.dy * r.dy + r.dz * r.dz;
    double e = r.ox * r.ox + r.oy * r.oy + r.oz * r.oz - r1 * r1 - r2 * r2;
    double f = r.ox * r.dx + r.oy * r.dy + r.oz * r.dz;
    double four_a_sq = 4.0 * r1 * r1;
    double coeffs[5] = {
        e * e - four_a_sq * (r2 * r2 - r.oy * r.oy),
        4.0 * f * e + 2.0 * four_a_sq * r.oy * r.dy,
        2.0 * sum_d_sq * e + 4.0 * f * f + four_a_sq * r.dy * r.dy,
        4.0 * sum_d_sq * f,
        sum_d_sq * sum_d_sq
    };
    double sol[4];
    int n = solve_quartic(coeffs, sol);
    /* the reference stores solutions[n-1] first (toroid.c:47-49) */
    for (int k = 0; k < n; ++k) {
        t[k] = sol[n - 1 - k];
    }
    return n;
}

/*
 * PRIMS: bit t set = a leaf of type t can occur.  A kernel instantiated for the types a scene really contains drops the
 * code (and the registers) of the others -- the quartic solver of the torus alone is a third of the general body.
 */
#define FRT_PRIMS_ALL 0x3ff
#define FRT_PRIMS_BOXES_AND_BALLS ((1 << FRT_CUBE) | (1 << FRT_SPHERE) | (1 << FRT_PLANE) | (1 << FRT_CSG) | (1 << FRT_GROUP))
template <int PRIMS = FRT_PRIMS_ALL>
__device__ __forceinline__ int
prim_intersect(int type, const double *prm, const Ray &r, double t[4], double uv[2])
{
    if (PRIMS != FRT_PRIMS_ALL) { /* the host checked that the scene holds no other type */
        if (type == FRT_SPHERE) {
            if (!((PRIMS >> FRT_SPHERE) & 1)) return 0;
        } else if (type == FRT_PLANE) {
            if (!((PRIMS >> FRT_PLANE) & 1)) return 0;
        } else if (type == FRT_CUBE) {
            if (!((PRIMS >> FRT_CUBE) & 1)) return 0;
        } else if (!((PRIMS >> FRT_CYLINDER) & 1) && !((PRIMS >> FRT_CONE) & 1) && !((PRIMS >> FRT_TOROID) & 1) &&
                   !((PRIMS >> FRT_TRIANGLE) & 1) && !((PRIMS >> FRT_SMOOTH_TRIANGLE) & 1)) {
            return 0;
        }
    }
    switch (type) {
    case FRT_SPHERE: { /* sphere_local_intersect, sphere.c:14-40 */
        double a = r.dx * r.dx + r.dy * r.dy + r.dz * r.dz;
        double b = 2 * (r.dx * r.ox + r.dy * r.oy + r.dz * r.oz);
        double c = (r.ox * r.ox + r.oy * r.oy + r.oz * r.oz) - 1.0;
        double disc = b * b - 4 * a * c;
        if (disc < 0) {
            return 0;
        }
        disc = sqrt(disc);
        a = 1.0 / (2 * a);
        t[0] = (-b - disc) * a;
        t[1] = (-b + disc) * a;
        return 2;
    }
    case FRT_PLANE: /* plane_local_intersect, plane.c:11-25 */
        if (fabs(r.dy) < FRT_EPS) {
            return 0;
        }
        t[0] = -r.oy / r.dy;
        return 1;
    case FRT_CUBE: { /* cube_local_intersect, cube.c:56-78 */
        double x0, x1, y0, y1, z0, z1;
        slab_axis(r.ox, r.dx, -1.0, 1.0, x0, x1);
        slab_axis(r.oy, r.dy, -1.0, 1.0, y0, y1);
        slab_axis(r.oz, r.dz, -1.0, 1.0, z0, z1);
        double tmin = fmax(fmax(x0, y0), z0);
        double tmax = fmin(fmin(x1, y1), z1);
        if (tmin > tmax) {
            return 0;
        }
        t[0] = tmin;
        t[1] = tmax;
        return 2;
    }
    case FRT_CYLINDER: { /* cylinder_local_intersect, cylinder.c:43-87 (+ caps :21-41) */
        double mn = __ldg(prm + 0), mx = __ldg(prm + 1);
        bool closed = __ldg(prm + 2) != 0.0;
        int n = 0;
        double a = r.dx * r.dx + r.dz * r.dz;
        double b = 2 * (r.ox * r.dx + r.oz * r.dz);
        double c = r.ox * r.ox + r.oz * r.oz - 1;
        if (!(fabs(a) < FRT_EPS)) {
            double disc = b * b - 4 * a * c;
            if (disc < 0) {
                return 0;
            }
            double sq = sqrt(disc);
            double t0 = (-b - sq) / (2 * a);
            double t1 = (-b + sq) / (2 * a);
            if (t0 > t1) {
                double tmp = t0;
                t0 = t1;
                t1 = tmp;
            }
            double y0 = r.oy + t0 * r.dy;
            if (mn <= y0 && y0 <= mx) {
                t[n++] = t0;
            }
            double y1 = r.oy + t1 * r.dy;
            if (mn <= y1 && y1 <= mx) {
                t[n++] = t1;
            }
        }
        if (closed && !(fabs(r.dy) < FRT_EPS)) {
            double ta = (mn - r.oy) / r.dy;
            double tb = (mx - r.oy) / r.dy;
            double x = r.ox + ta * r.dx, z = r.oz + ta * r.dz;
            if (x * x + z * z <= 1) {
                t[n++] = ta;
            }
            x = r.ox + tb * r.dx;
            z = r.oz + tb * r.dz;
            if (x * x + z * z <= 1) {
                t[n++] = tb;
            }
        }
        return n;
    }
    case FRT_CONE: { /* cone_local_intersect, cone.c:43-96 (+ caps :12-41; cap radius test is sm <= |y| as written) */
        double mn = __ldg(prm + 0), mx = __ldg(prm + 1);
        bool closed = __ldg(prm + 2) != 0.0;
        int n = 0;
        double a = r.dx * r.dx + r.dz * r.dz - r.dy * r.dy;
        double b = 2 * (r.ox * r.dx + r.oz * r.dz - r.oy * r.dy);
        double c = r.ox * r.ox + r.oz * r.oz - r.oy * r.oy;
        if (fabs(a) < FRT_EPS) {
            if (!(fabs(b) < FRT_EPS)) {
                t[n++] = -c / (2 * b);
            }
        } else {
            double disc = b * b - 4 * a * c;
            if (disc < 0) {
                return 0;
            }
            double sq = sqrt(disc);
            double t0 = (-b - sq) / (2 * a);
            double t1 = (-b + sq) / (2 * a);
            if (t0 > t1) {
                double tmp = t0;
                t0 = t1;
                t1 = tmp;
            }
            double y0 = r.oy + t0 * r.dy;
            if (mn < y0 && y0 < mx) {
                t[n++] = t0;
            }
            double y1 = r.oy + t1 * r.dy;
            if (mn < y1 && y1 < mx) {
                t[n++] = t1;
            }
        }
        if (closed && !(fabs(r.dy) < FRT_EPS)) {
            double ta = (mn - r.oy) / r.dy;
            double x = r.ox + ta * r.dx, z = r.oz + ta * r.dz;
            if (x * x + z * z <= fabs(mn)) {
                t[n++] = ta;
            }
            double tb = (mx - r.oy) / r.dy;
            x = r.ox + tb * r.dx;
            z = r.oz + tb * r.dz;
            if (x * x + z * z <= fabs(mx)) {
                t[n++] = tb;
            }
        }
        return n;
    }
    case FRT_TOROID:
        return toroid_intersect(prm, r, t);
    case FRT_TRIANGLE:
    case FRT_SMOOTH_TRIANGLE: { /* triangle_local_intersect, triangle.c:11-45 / :122-156 (Moller-Trumbore) */
        double p1x = __ldg(prm + 0), p1y = __ldg(prm + 1), p1z = __ldg(prm + 2);
        double e1x = __ldg(prm + 9), e1y = __ldg(prm + 10), e1z = __ldg(prm + 11);
        double e2x = __ldg(prm + 12), e2y = __ldg(prm + 13), e2z = __ldg(prm + 14);
        double cx = r.dy * e2z - r.dz * e2y;
        double cy = r.dz * e2x - r.dx * e2z;
        double cz = r.dx * e2y - r.dy * e2x;
        double det = e1x * cx + e1y * cy + e1z * cz;
        if (fabs(det) < FRT_EPS) {
            return 0;
        }
        double f = 1.0 / det;
        double sx = r.ox - p1x, sy = r.oy - p1y, sz = r.oz - p1z;
        double u = f * (sx * cx + sy * cy + sz * cz);
        if (u < 0 || u > 1) {
            return 0;
        }
        double qx = sy * e1z - sz * e1y;
        double qy = sz * e1x - sx * e1z;
        double qz = sx * e1y - sy * e1x;
        double v = f * (r.dx * qx + r.dy * qy + r.dz * qz);
        if (v < 0 || (u + v) > 1) {
            return 0;
        }
        t[0] = f * (e2x * qx + e2y * qy + e2z * qz);
        uv[0] = u;
        uv[1] = v;
        return 1;
    }
    default:
        return 0;
    }
}

/* Algorithmic flop per event, the table frozen in BASELINE.md section 4 (used only by the counting build of the
 * kernels, FRT_FLAG_COUNT_RAYS, to state the roofline numerator). */
#define FRT_COST_XFORM 33
#define FRT_COST_BBOX 16
#define FRT_COST_LIGHT_SAMPLE 100
__device__ __forceinline__ unsigned int
prim_cost(int type)
{
    switch (type) {
    case FRT_SPHERE: return 28;
    case FRT_CUBE: return 16;
    case FRT_PLANE: return 2;
    case FRT_CYLINDER: return 42;
    case FRT_CONE: return 50;
    case FRT_TOROID: return 200;
    default: return 45; /* triangles */
    }
}

/* ---- CSG: per-ray interval stack ------------------------------------------------------------------- */

struct CsgHit {
    double t;
    int leaf;
};

/* intersection_allowed, csg.c:28-40 */
__device__ __forceinline__ bool
csg_allowed(int op, bool lhit, bool inl, bool inr)
{
    if (op == FRT_CSG_UNION) {
        return (lhit && !inr) || (!lhit && !inl);
    } else if (op == FRT_CSG_INTERSECT) {
        return (lhit && inr) || (!lhit && inl);
    }
    return (lhit && !inr) || (!lhit && inl);
}

/*
 * csg_local_intersect (csg.c:74-125) for the whole CSG subtree rooted at `root`, evaluated bottom-up over the
 * pre-order node array.  buf[0..return) receives the filtered crossings (t, leaf node), sorted by t when both
 * operands contributed (csg.c:104-118).  Groups inside a CSG are transparent containers here: their own sort is
 * unobservable (the enclosing CSG sorts or keeps/drops the list wholesale) -- the one reference behaviour not
 * reproduced is the shadow early-out of a group nested INSIDE a CSG operand (group.c:105-123).
 * `cur_xf`/`lr` cache the ray in the current node space.  Sets *overflow when the interval stack is too small.
 */
template <bool COUNT, int PRIMS = FRT_PRIMS_ALL>
__device__ __noinline__ int
csg_eval_t(const DScene &S, int root, const Ray &wr, CsgHit *buf, int *overflow, unsigned long long *flops)
{
    struct Frame {
        int node, right, skip, start, mid, op;
    };
    Frame st[8];
    int sp = 0;
    int n = 0;
    int i = root;
    int end = load_node_a(S, root).skip;
    int cur_xf = -1;
    Ray lr = wr;

    while (i < end) {
        NodeA a = load_node_a(S, i);
        if (a.xform != cur_xf) {
            cur_xf = a.xform;
            lr = ray_to_local(S, cur_xf, wr);
            if (COUNT && cur_xf != 0) *flops += FRT_COST_XFORM;
        }
        if (COUNT) *flops += (a.type >= FRT_CSG) ? FRT_COST_BBOX : prim_cost(a.type);
        if (a.type == FRT_CSG) {
            if (!bbox_hit(S, i, lr)) {
                i = a.skip;
            } else {
                NodeB b = load_node_b(S, i);
                if (sp == 8) {
                    *overflow = 1;
                    return 0;
                }
                st[sp++] = Frame{ i, b.right, a.skip, n, -1, b.csg_op };
                i = i + 1;
            }
        } else if (a.type == FRT_GROUP) {
            i = bbox_hit(S, i, lr) ? i + 1 : a.skip;
        } else {
            NodeB b = load_node_b(S, i);
            double t[4], uv[2];
            int k = prim_intersect<PRIMS>(a.type, S.params + (b.param < 0 ? 0 : b.param), lr, t, uv);
            for (int j = 0; j < k; ++j) {
                if (n == FRT_CSG_CAP) {
                    *overflow = 1;
                    return 0;
                }
                buf[n].t = t[j];
                buf[n].leaf = i;
                ++n;
            }
            i = i + 1;
        }
        /* close every frame whose left / right operand just ended */
        while (sp > 0) {
            Frame &f = st[sp - 1];
            if (f.mid < 0 && i >= f.right) {
                f.mid = n;
            }
            if (i < f.skip) {
                break;
            }
            int nl = f.mid - f.start, nr = n - f.mid;
            if (nl > 0 && nr > 0) {
                /* intersections_sort over both operands (insertion sort; lists are a handful of entries) */
                for (int x = f.start + 1; x < n; ++x) {
                    CsgHit h = buf[x];
                    int y = x - 1;
                    while (y >= f.start && buf[y].t > h.t) {
                        buf[y + 1] = buf[y];
                        --y;
                    }
                    buf[y + 1] = h;
                }
            }
            /* csg_filter_intersections, csg.c:43-71; lhit = leaf lies in the left operand's subtree */
            bool inl = false, inr = false;
            int out = f.start;
            for (int x = f.start; x < n; ++x) {
                bool lhit = buf[x].leaf < f.right;
                if (csg_allowed(f.op, lhit, inl, inr)) {
                    buf[out++] = buf[x];
                }
                if (lhit) {
                    inl = !inl;
                } else {
                    inr = !inr;
                }
            }
            n = out;
            --sp;
        }
    }
    return n;
}

template <int PRIMS = FRT_PRIMS_ALL>
__device__ __forceinline__ int
csg_eval(const DScene &S, int root, const Ray &wr, CsgHit *buf, int *overflow)
{
    return csg_eval_t<false, PRIMS>(S, root, wr, buf, overflow, nullptr);
}

/* ---- traversal -------------------------------------------------------------------------------------- */

struct Hit {
    double t, u, v;
    int leaf;
};

/*
 * intersect_world(w, r, false) + hit(xs, false) (world.c:164-197, intersection.c:42-54): the smallest t > 0 over
 * every leaf whose ancestors' bounding boxes the ray meets.  Order of visit is irrelevant for the minimum; ties
 * at equal t (glibc qsort order in the reference) resolve to the first leaf in tree order.
 */
template <bool CASTERS>
__device__ __forceinline__ Hit
trace_closest_t(const DScene &S, const Ray &wr, int *overflow)
{
    Hit best;
    best.t = CUDART_INF;
    best.u = best.v = -1.0;
    best.leaf = -1;
    CsgHit buf[FRT_CSG_CAP];
    const InvDir winv = inv_dir(wr);
    for (int rt = 0; rt < S.n_roots; ++rt) {
        int i = __ldg(S.roots + rt);
        int end = load_node_a(S, i).skip;
        int cur_xf = 0;
        Ray lr = wr;
        InvDir inv = winv;
        while (i < end) {
            NodeA a = load_node_a(S, i);
            if (a.xform != cur_xf) {
                cur_xf = a.xform;
                if (cur_xf == 0) {
                    lr = wr;
                    inv = winv;
                } else {
                    lr = ray_to_local(S, cur_xf, wr);
                    inv = inv_dir(lr);
                }
            }
            if (a.type >= FRT_CSG) {
                /* The minimum over the leaves does not depend on the order or on which empty subtrees are skipped: besides
                 * the reference's own cull (tmin > tmax) skip boxes wholly behind the origin (no t > 0 inside) and boxes
                 * wholly beyond the best hit so far (t is the same parameter in every node's frame). */
                double tmin, tmax;
                bbox_range_inv(S, i, lr, inv, tmin, tmax);
                const bool miss = !(tmin <= tmax) || tmax < 0.0 || tmin > best.t;
                if (a.type == FRT_GROUP) {
                    i = miss ? a.skip : i + 1;
                } else {
                    if (!miss) {
                        int n = csg_eval(S, i, wr, buf, overflow);
                        for (int k = 0; k < n; ++k) {
                            if (buf[k].t > 0 && buf[k].t < best.t &&
                                (!CASTERS || S.mats[load_node_a(S, buf[k].leaf).material].casts_shadow)) {
                                best.t = buf[k].t;
                                best.leaf = buf[k].leaf;
                                best.u = best.v = -1.0;
                            }
                        }
                    }
                    i = a.skip;
                }
            } else {
                NodeB b = load_node_b(S, i);
                double t[4], uv[2];
                uv[0] = uv[1] = -1.0;
                int k = prim_intersect(a.type, S.params + (b.param < 0 ? 0 : b.param), lr, t, uv);
                if (CASTERS && !S.mats[a.material].casts_shadow) {
                    k = 0; /* hit(xs, true) skips objects that do not cast shadows (intersection.c:42-54) */
                }
                for (int j = 0; j < k; ++j) {
                    if (t[j] > 0 && t[j] < best.t) {
                        best.t = t[j];
                        best.leaf = i;
                        best.u = uv[0];
                        best.v = uv[1];
                    }
                }
                i = i + 1;
            }
        }
    }
    return best;
}

__device__ __forceinline__ Hit
trace_closest(const DScene &S, const Ray &wr, int *overflow)
{
    return trace_closest_t<false>(S, wr, overflow);
}

/*
 * Shadow-ray summary of one primitive: `stop` = its crossing list holds a t that is not <= 0 (the reference's
 * search ends here, group.c:105-123), `tmin` = its smallest t > 0.  Sphere and cube -- the two types a Cornell
 * shadow ray meets -- are answered from their two roots without materialising the list; the other types go
 * through prim_intersect.
 */
__device__ __forceinline__ int
prim_intersect_inv(int type, const double *prm, const Ray &r, const InvDir &inv, double t[4], double uv[2])
{
    if (type == FRT_CUBE) { /* cube_local_intersect, cube.c:56-78 */
        double x0, x1, y0, y1, z0, z1;
        slab_axis_inv(r.ox, inv.x, -1.0, 1.0, x0, x1);
        slab_axis_inv(r.oy, inv.y, -1.0, 1.0, y0, y1);
        slab_axis_inv(r.oz, inv.z, -1.0, 1.0, z0, z1);
        double tmin = fmax(fmax(x0, y0), z0);
        double tmax = fmin(fmin(x1, y1), z1);
        if (tmin > tmax) {
            return 0;
        }
        t[0] = tmin;
        t[1] = tmax;
        return 2;
    }
    if (type == FRT_SPHERE) { /* sphere_local_intersect, sphere.c:14-40 */
        double a = r.dx * r.dx + r.dy * r.dy + r.dz * r.dz;
        double b = 2 * (r.dx * r.ox + r.dy * r.oy + r.dz * r.oz);
        double c = (r.ox * r.ox + r.oy * r.oy + r.oz * r.oz) - 1.0;
        double disc = b * b - 4 * a * c;
        if (disc < 0) {
            return 0;
        }
        disc = disc > 0 ? disc * rsqrt_fast(disc) : 0.0;
        a = rcp_fast(2 * a);
        t[0] = (-b - disc) * a;
        t[1] = (-b + disc) * a;
        return 2;
    }
    return prim_intersect(type, prm, r, t, uv);
}

/*
 * is_shadowed (renderer.c:73-93) = intersect_world(w, r, true) + hit(xs, true) with the reference's
 * order-dependent early-out (group.c:105-123): walk the divided tree in the reference's child
 * order; the search ENDS at the first leaf (or CSG) whose crossing list holds any t that is not <= 0, whether or
 * not that crossing is nearer than the light or casts a shadow.  The point is shadowed iff that leaf has a
 * positive crossing on a casts_shadow material nearer than `distance`.
 *
 * One loop serves plain leaves and CSG subtrees (csg_local_intersect, csg.c:74-125): inside a CSG the leaves'
 * crossings are pushed on the per-ray interval stack `buf`, and when a CSG node's subtree ends its operands are
 * merged, sorted and filtered in place (csg_filter_intersections, csg.c:43-71); the outermost CSG is then judged
 * like a leaf.  Keeping a single primitive-intersection site keeps the kernel's instruction footprint small: the
 * first version of this stage spent half of its issue slots waiting on instruction fetch (profiles/).
 */
#define FRT_CSG_DEPTH 8
template <bool COUNT>
__device__ __forceinline__ bool
trace_shadow(const DScene &S, const Ray &wr, double distance, int *overflow, unsigned long long *nodes_visited,
             unsigned long long *flops)
{
    struct Frame {
        int right, skip, start, mid, op;
    };
    CsgHit buf[FRT_CSG_CAP];
    Frame st[FRT_CSG_DEPTH];
    int sp = 0, n = 0;
    unsigned int visited = 0, cost = 0;
    bool result = false;
    const InvDir winv = inv_dir(wr);

    for (int rt = 0; rt < S.n_roots; ++rt) {
        int i = __ldg(S.roots + rt);
        const int end = load_node_a(S, i).skip;
        int cur_xf = 0;
        Ray lr = wr;
        InvDir inv = winv;
        bool any = false; /* world.c:189-191: stop after the first top-level shape that returned anything */
        bool done = false;
        while (i < end) {
            const NodeA a = load_node_a(S, i);
            if (COUNT) ++visited;
            if (a.xform != cur_xf) {
                cur_xf = a.xform;
                if (cur_xf == 0) {
                    lr = wr;
                    inv = winv;
                } else {
                    lr = ray_to_local(S, cur_xf, wr);
                    inv = inv_dir(lr);
                    if (COUNT) cost += FRT_COST_XFORM;
                }
            }
            if (COUNT) cost += (a.type >= FRT_CSG) ? FRT_COST_BBOX : prim_cost(a.type);
            if (a.type >= FRT_CSG) { /* CSG or group: cull by the node's own bounds */
                double bt0, bt1;
                bbox_range_inv(S, i, lr, inv, bt0, bt1);
                /* the reference's cull (tmin > tmax) plus boxes wholly behind the origin: every crossing inside has
                 * t <= 0, which neither stops the search nor shadows (group.c:105-123) -- top-level subtrees only, a CSG
                 * operand's negative crossings still toggle the filter state */
                if (!(bt0 <= bt1) || (sp == 0 && bt1 < 0.0)) {
                    i = a.skip;
                } else {
                    if (a.type == FRT_CSG) {
                        if (sp == FRT_CSG_DEPTH) {
                            *overflow = 1;
                            return false;
                        }
                        const NodeB b = load_node_b(S, i);
                        st[sp++] = Frame{ b.right, a.skip, n, -1, b.csg_op };
                    }
                    i = i + 1;
                }
            } else {
                const NodeB b = load_node_b(S, i);
                double t[4], uv[2];
                const int k = prim_intersect_inv(a.type, S.params + (b.param < 0 ? 0 : b.param), lr, inv, t, uv);
                if (sp == 0) {
                    bool stop = false;
                    double tmin = CUDART_INF;
                    for (int j = 0; j < k; ++j) {
                        any = true;
                        stop = stop || !(t[j] <= 0);
                        if (t[j] > 0 && t[j] < tmin) {
                            tmin = t[j];
                        }
                    }
                    if (stop) {
                        result = S.mats[a.material].casts_shadow && tmin < distance;
                        done = true;
                        break;
                    }
                } else {
                    for (int j = 0; j < k; ++j) {
                        if (n == FRT_CSG_CAP) {
                            *overflow = 1;
                            return false;
                        }
                        buf[n].t = t[j];
                        buf[n].leaf = i;
                        ++n;
                    }
                }
                i = i + 1;
            }
            /* close every CSG whose left / right operand just ended */
            while (sp > 0) {
                Frame &f = st[sp - 1];
                if (f.mid < 0 && i >= f.right) {
                    f.mid = n;
                }
                if (i < f.skip) {
                    break;
                }
                if (f.mid - f.start > 0 && n - f.mid > 0) { /* intersections_sort over both operands */
                    for (int x = f.start + 1; x < n; ++x) {
                        CsgHit h = buf[x];
                        int y = x - 1;
                        while (y >= f.start && buf[y].t > h.t) {
                            buf[y + 1] = buf[y];
                            --y;
                        }
                        buf[y + 1] = h;
                    }
                }
                bool inl = false, inr = false;
                int out = f.start;
                for (int x = f.start; x < n; ++x) {
                    const bool lhit = buf[x].leaf < f.right;
                    if (csg_allowed(f.op, lhit, inl, inr)) {
                        buf[out++] = buf[x];
                    }
                    if (lhit) {
                        inl = !inl;
                    } else {
                        inr = !inr;
                    }
                }
                n = out;
                --sp;
                if (sp == 0) { /* the outermost CSG is judged like a leaf */
                    bool stop = false;
                    double tmin = CUDART_INF;
                    for (int x = 0; x < n; ++x) {
                        any = true;
                        stop = stop || !(buf[x].t <= 0);
                        if (buf[x].t > 0 && buf[x].t < tmin && S.mats[load_node_a(S, buf[x].leaf).material].casts_shadow) {
                            tmin = buf[x].t;
                        }
                    }
                    n = 0;
                    if (stop) {
                        result = tmin < distance;
                        done = true;
                    }
                }
            }
            if (done) {
                break;
            }
        }
        if (done || any) {
            break;
        }
    }
    if (COUNT) {
        *nodes_visited += visited;
        *flops += cost;
    }
    return result;
}

/*
 * The refraction containers of prepare_computations (renderer.c:403-447), without the list: every crossing that
 * sorts before the hit has t <= 0 (the hit is the first t > 0), so a leaf is "open" iff it has an odd number of
 * crossings with t <= 0, and the container order is the order of each open leaf's LAST such crossing.
 *   n1 = Ni of the open leaf whose last crossing is largest (1.0 when none is open);
 *   n2 = the hit leaf was open -> it is removed: Ni of the best other open leaf (1.0 if none),
 *        else it is appended: Ni of the hit leaf.
 */
__device__ __noinline__ void
trace_containers(const DScene &S, const Ray &wr, int hit_leaf, double &n1, double &n2, int *overflow)
{
    CsgHit buf[FRT_CSG_CAP];
    double best_t = -CUDART_INF, best_other_t = -CUDART_INF;
    int best_leaf = -1, best_other_leaf = -1;
    bool hit_open = false;

    for (int rt = 0; rt < S.n_roots; ++rt) {
        int i = __ldg(S.roots + rt);
        int end = load_node_a(S, i).skip;
        int cur_xf = 0;
        Ray lr = wr;
        InvDir inv = inv_dir(wr);
        while (i < end) {
            NodeA a = load_node_a(S, i);
            if (a.xform != cur_xf) {
                cur_xf = a.xform;
                lr = ray_to_local(S, cur_xf, wr);
                inv = inv_dir(lr);
            }
            int n = 0;
            if (a.type >= FRT_CSG) {
                /* only crossings at t <= 0 count (see above): a box the ray enters at t > 0 holds none of them.  A primary
                 * ray has the whole scene in front of it, so this walk ends at the first level of groups. */
                double tmin, tmax;
                bbox_range_inv(S, i, lr, inv, tmin, tmax);
                const bool miss = !(tmin <= tmax) || tmin > 0.0;
                if (a.type == FRT_GROUP) {
                    i = miss ? a.skip : i + 1;
                    continue;
                }
                if (!miss) {
                    n = csg_eval(S, i, wr, buf, overflow);
                }
                i = a.skip;
            } else {
                NodeB b = load_node_b(S, i);
                double t[4], uv[2];
                n = prim_intersect(a.type, S.params + (b.param < 0 ? 0 : b.param), lr, t, uv);
                for (int j = 0; j < n; ++j) {
                    buf[j].t = t[j];
                    buf[j].leaf = i;
                }
                i = i + 1;
            }
            /* per leaf present in buf: parity and last crossing among t <= 0 */
            for (int x = 0; x < n; ++x) {
                int leaf = buf[x].leaf;
                bool first = true;
                for (int y = 0; y < x; ++y) {
                    if (buf[y].leaf == leaf) {
                        first = false;
                    }
                }
                if (!first) {
                    continue;
                }
                int cnt = 0;
                double last = -CUDART_INF;
                for (int y = x; y < n; ++y) {
                    if (buf[y].leaf == leaf && !(buf[y].t > 0)) {
                        ++cnt;
                        last = fmax(last, buf[y].t);
                    }
                }
                if (cnt & 1) {
                    if (last > best_t || best_leaf < 0) {
                        best_t = last;
                        best_leaf = leaf;
                    }
                    if (leaf == hit_leaf) {
                        hit_open = true;
                    } else if (last > best_other_t || best_other_leaf < 0) {
                        best_other_t = last;
                        best_other_leaf = leaf;
                    }
                }
            }
        }
    }
    n1 = best_leaf >= 0 ? S.mats[load_node_a(S, best_leaf).material].Ni : 1.0;
    if (hit_open) {
        n2 = best_other_leaf >= 0 ? S.mats[load_node_a(S, best_other_leaf).material].Ni : 1.0;
    } else {
        n2 = S.mats[load_node_a(S, hit_leaf).material].Ni;
    }
}

/* ---- normals ---------------------------------------------------------------------------------------- */

/* the per-shape local_normal_at functions: sphere.c:42, plane.c:27, cube.c:80, cylinder.c:90, cone.c:99,
 * toroid.c:55, triangle.c:47, triangle.c:158 */
__device__ __forceinline__ void
local_normal(int type, const double *prm, const double lp[3], double u, double v, double n[3])
{
    n[0] = n[1] = n[2] = 0.0;
    switch (type) {
    case FRT_SPHERE:
        n[0] = lp[0];
        n[1] = lp[1];
        n[2] = lp[2];
        break;
    case FRT_PLANE:
        n[1] = 1.0;
        break;
    case FRT_CUBE: {
        double ax = fabs(lp[0]), ay = fabs(lp[1]), az = fabs(lp[2]);
        double maxc = fmax(fmax(ax, ay), az);
        if (fabs(maxc - ax) < FRT_EPS) {
            n[0] = lp[0];
        } else if (fabs(maxc - ay) < FRT_EPS) {
            n[1] = lp[1];
        } else {
            n[2] = lp[2];
        }
        break;
    }
    case FRT_CYLINDER: {
        double dist = lp[0] * lp[0] + lp[2] * lp[2];
        double mn = __ldg(prm + 0), mx = __ldg(prm + 1);
        if (dist < 1 && (mx - FRT_EPS) <= lp[1]) {
            n[1] = 1;
        } else if (dist < 1 && (mn + FRT_EPS) >= lp[1]) {
            n[1] = -1;
        } else {
            n[0] = lp[0];
            n[2] = lp[2];
        }
        break;
    }
    case FRT_CONE: {
        double dist = lp[0] * lp[0] + lp[2] * lp[2];
        double mn = __ldg(prm + 0), mx = __ldg(prm + 1);
        if (dist < 1 && (mx - FRT_EPS) <= lp[1]) {
            n[1] = 1;
        } else if (dist < 1 && (mn + FRT_EPS) >= lp[1]) {
            n[1] = -1;
        } else {
            double y = sqrt(dist);
            if (lp[1] > 0) {
                y = -y;
            }
            n[0] = lp[0];
            n[1] = y;
            n[2] = lp[2];
        }
        break;
    }
    case FRT_TOROID: {
        double r1 = __ldg(prm + 0), r2 = __ldg(prm + 1);
        double p_sq = r1 * r1 + r2 * r2;
        double mag = lp[0] * lp[0] + lp[1] * lp[1] + lp[2] * lp[2];
        double x = 4.0 * lp[0] * (mag - p_sq);
        double y = 4.0 * lp[1] * (mag - p_sq + 2.0 * r1 * r1);
        double z = 4.0 * lp[2] * (mag - p_sq);
        double inv = 1.0 / sqrt(x * x + y * y + z * z);
        n[0] = x * inv;
        n[1] = y * inv;
        n[2] = z * inv;
        break;
    }
    case FRT_TRIANGLE:
        n[0] = __ldg(prm + 15);
        n[1] = __ldg(prm + 16);
        n[2] = __ldg(prm + 17);
        break;
    case FRT_SMOOTH_TRIANGLE: {
        double w = 1.0 - u - v;
        for (int k = 0; k < 3; ++k) {
            double a = __ldg(prm + 15 + k) * w;
            double b = __ldg(prm + 18 + k) * u;
            double c = __ldg(prm + 21 + k) * v;
            n[k] = a + (b + c);
        }
        break;
    }
    default:
        break;
    }
}

/*
 * shape_normal_to_world (shapes.c:92-114) through the whole parent chain.  Each non-identity level multiplies by
 * the transpose of its inverse and normalises, so the direction is composite^T * n and the length is 1 as soon as
 * one level is non-identity (xf != 0); with an all-identity chain the local normal passes through unscaled, which
 * matters only to the bump-map sum in shape_normal_at (shapes.c:76-86).
 */
__device__ __forceinline__ void
normal_to_world(const DScene &S, int xf, const double ln[3], double wn[3])
{
    if (xf == 0) {
        wn[0] = ln[0];
        wn[1] = ln[1];
        wn[2] = ln[2];
        return;
    }
    const double *m = S.xinv + 12 * xf;
    double x = __ldg(m + 0) * ln[0] + __ldg(m + 4) * ln[1] + __ldg(m + 8) * ln[2];
    double y = __ldg(m + 1) * ln[0] + __ldg(m + 5) * ln[1] + __ldg(m + 9) * ln[2];
    double z = __ldg(m + 2) * ln[0] + __ldg(m + 6) * ln[1] + __ldg(m + 10) * ln[2];
    double inv = 1.0 / sqrt(x * x + y * y + z * z);
    wn[0] = x * inv;
    wn[1] = y * inv;
    wn[2] = z * inv;
}
