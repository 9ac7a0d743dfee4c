/*
 * frt_shadow_f32.cuh -- FP32 filtered shadow-ray traversal.
 *
 * is_shadowed (renderer.c:73-93) is 98.7 % of the reference's rays on the Cornell scene and
 * its answer is one bit.  That bit is the outcome of a chain of comparisons (slab order, t against 0, t against the
 * light distance, order of CSG crossings).  This file walks the same tree in the same order as the FP64 traversal
 * (trace_shadow in frt_device.cuh) but in FP32 interval arithmetic: every t value is an interval [lo, hi] that
 * surely contains the value the FP64 traversal computes.
 *
 *   - a comparison of two disjoint intervals has the same outcome as in FP64: keep going;
 *   - a group / CSG bounding-box test is only a cull, so it is resolved conservatively (when in doubt, descend:
 *     the children are inside the box, so descending where the reference culls changes nothing);
 *   - anything else -- overlapping intervals, a primitive type or a CSG shape this filter has no fast form for --
 *     makes the ray UNDECIDED.  The caller appends it to a queue (warp-aggregated) and a second kernel re-traces
 *     the queue with the general FP64 traversal.  On the Cornell frame that is ~2 % of the rays.
 *
 * The filter is therefore free to be narrow: it only has to be right when it answers.  FRT_FLAG_VERIFY_F32 checks
 * that claim ray by ray on the device (every decided ray is also traced in FP64; tests assert zero disagreements).
 *
 * What it has fast forms for
 *   - cube leaves and group / CSG bounds whose composite transform is axis-aligned (scale, translate, quarter
 *     turns: every wall of the Cornell box): slab test against the WORLD-space box with the world ray -- no
 *     transform at all.  t is the same parameter in world and local space, so the value is the reference's up to
 *     FP64 rounding (1e-16), far inside the FP32 intervals;
 *   - other cubes, spheres, planes: transform to the leaf's frame (FP32 copy of the matrix), then slab / quadratic;
 *   - CSG (csg_local_intersect + csg_filter_intersections, csg.c:43-125) over operands that each cross the ray as
 *     ONE interval [enter, exit]: union / intersection / difference of two intervals by the reference's in/out
 *     toggling rules, worked out per ordering (see csg_combine).  A result that would be two intervals is left to
 *     FP64.
 *
 * Error model.  u = 2^-24 (FP32 unit roundoff), G = 2^-21 = 8u bounds the rounding of an expression of up to four
 * products / sums relative to the sum of the magnitudes of its terms.  World frame: origin rounded from FP64 and box
 * bounds rounded to nearest, eo = 2u (|o|max + Bmax); direction normalised in FP32 from the rounded difference of
 * the two world points, ed = G + 2u (|p|max + |o|max) / |p - o|.  A transform M (rows m_k, translation T_k) maps
 * these to  eo_k = R_k eo + G (R_k |o|max + |T_k|),  ed_k = R_k (ed + G),  R_k = sum_j |m_kj|.  A slab value
 * t = (b - o_k) / d_k with |d_k| >= EPSILON + 2 ed_k lies within  |1/d_k| (eo_k + 2 |t| ed_k) + G |t|  of its
 * FP32 evaluation; axes with a smaller |d_k| (the reference's "* INFINITY" branch, cube.c:27-33) are left open for
 * culls and make a leaf undecided.
 */
#pragma once

#include "frt_device.cuh"

#define FRT_F32_G 4.76837158203125e-07f   /* 2^-21 */
#define FRT_F32_U 5.9604644775390625e-08f /* 2^-24 */
#define FRT_EPS_F 0.00001f

enum { FRT_SH_LIT = 0, FRT_SH_SHADOWED = 1, FRT_SH_UNDECIDED = 2 };

/* FP32 mirror of one tree node, 3 x float4 */
enum {
    FRT_FN_TYPE_MASK = 15, /* enum frt_node_type */
    FRT_FN_WORLD = 16,     /* lo / hi are WORLD-space bounds: test them with the world ray (no transform) */
    FRT_FN_CASTS = 32,     /* leaf: its material casts shadows */
    FRT_FN_FAST = 64,      /* leaf: the filter has a fast form for this type (cube, sphere, plane);
                              CSG: it has a postfix program that fits the three-register span stack */
    FRT_FN_NOCULL = 128,   /* group: descend without testing its bounds (culling is optional; the root's box is always hit) */
    FRT_FN_OP_SHIFT = 8,   /* CSG: enum frt_csg_op in bits 8..9 */
    FRT_FN_LEAFBOX = 1024  /* triangle leaf: lo / hi are the bounds of its vertices (frame q0.z, like a group's): cull before the FP64 test */
};

struct DSceneF {
    const float4 *fnodes; /* per node: {flags, skip, xform, right (CSG)} as int bits, {lo.xyz, w}, {hi.xyz, w};
                             outermost CSG nodes: lo.w / hi.w = start / length of the node's postfix program */
    const float4 *fx;     /* per xform: rows 0..2 of the world->local matrix, then {R_0, R_1, R_2, 0} */
    const float *lpoints; /* FP32 copy of the light sample points, 3 per point */
    const float4 *wbox;   /* per node: WORLD-space bounding box {min.xyz, 0} {max.xyz, 0}, rounded outward (shaft culling) */
    const float4 *wsphere;/* per node (first 32): {centre xyz, radius} of a sphere leaf whose transform is a similarity (a ball in
                             WORLD space), radius 0 otherwise -- the shaft tests of k_light_pre / trace_shadow_bulk */
    const float4 *shaft;  /* per light: 4 corners of a parallelogram that contains every surface sample of the light */
    const int *csg_prog;  /* postfix programs of the outermost CSG nodes: node index of a leaf, or -(op + 1) */
    const double *lbox;   /* per light: 5 axis-aligned boxes {min xyz, max xyz} of its sample points over every cached set,
                             box 0 = all samples, boxes 1..4 = the quadrants of the sample grid (trace_shadow_bulk) */
    const int4 *lquad;    /* per light: {samples per pending entry, half usteps (0: the grid is not split), half vsteps, usteps} */
    float bmax;           /* largest finite |bound| of a WORLD node */
    float smin;           /* smallest |axis scale| of an axis-aligned world->local transform (<= 1) */
    float ealign;         /* 2 x the largest off-axis / on-axis ratio of a transform treated as axis-aligned (<= 2e-9):
                             a WORLD box test sees the world point up to ealign |o|max, the direction up to ealign, off */
    int n_nodes;
    int n_csg_prog;          /* words in csg_prog */
    unsigned int entry_fast; /* bit i: node i (< 32) is a WORLD cube leaf, a CSG over WORLD cube leaves or a WORLD ball (entry_node_span) */
};

/*
 * Shaft culling.  All shadow rays of one hit start at the same point o and aim at points of the light's parallelogram
 * P, so they lie in the pyramid { o + t (p - o) : p in P, t > 0 }.  A node whose world-space box is separated from
 * that pyramid by a plane cannot be crossed at t > 0 by any of them, and a subtree without a positive crossing neither
 * stops the reference's search nor shadows (group.c:105-123): it is skipped for every ray of the hit.  The test is
 * one-sided (a box that is not provably outside stays), per hit, ~60 flop per node; it runs once per hit in
 * k_light_pre and leaves a bit mask over the first 32 nodes.  Separating planes tried: the pyramid's four sides, the
 * plane through o facing the light (when every corner is in front of it), and the box's own six faces.
 */
struct ShaftF {
    float d[4][3]; /* corner - o */
    float n[4][3]; /* inward side normals: n_i = +-(d_i x d_{i+1}) */
    float ax[3];   /* sum of the corner directions (the axis) */
    bool axis_ok;  /* every corner direction has a positive component along the axis */
    float dmin[3], dmax[3];
    float scale;   /* magnitude for the tolerance */
};

__device__ __forceinline__ void
shaft_setup(ShaftF &s, const float4 *corners, float ox, float oy, float oz)
{
    float m = 0.f;
    for (int i = 0; i < 4; ++i) {
        const float4 c = __ldg(corners + i);
        s.d[i][0] = c.x - ox;
        s.d[i][1] = c.y - oy;
        s.d[i][2] = c.z - oz;
        m = fmaxf(m, fmaxf(fabsf(s.d[i][0]), fmaxf(fabsf(s.d[i][1]), fabsf(s.d[i][2]))));
    }
    for (int k = 0; k < 3; ++k) {
        s.ax[k] = s.d[0][k] + s.d[1][k] + s.d[2][k] + s.d[3][k];
        s.dmin[k] = fminf(fminf(s.d[0][k], s.d[1][k]), fminf(s.d[2][k], s.d[3][k]));
        s.dmax[k] = fmaxf(fmaxf(s.d[0][k], s.d[1][k]), fmaxf(s.d[2][k], s.d[3][k]));
    }
    s.axis_ok = true;
    for (int i = 0; i < 4; ++i) {
        const float *a = s.d[i], *b = s.d[(i + 1) & 3];
        float nx = a[1] * b[2] - a[2] * b[1], ny = a[2] * b[0] - a[0] * b[2], nz = a[0] * b[1] - a[1] * b[0];
        /* orient towards the inside: the opposite corner lies inside */
        const float *c = s.d[(i + 2) & 3];
        const float side = nx * c[0] + ny * c[1] + nz * c[2];
        /* an origin (almost) in the light's plane leaves the orientation undecided: drop the plane */
        const float nn = fabsf(nx) + fabsf(ny) + fabsf(nz), cc = fabsf(c[0]) + fabsf(c[1]) + fabsf(c[2]);
        const float sg = fabsf(side) <= 1e-4f * nn * cc ? 0.f : (side < 0.f ? -1.f : 1.f);
        s.n[i][0] = nx * sg;
        s.n[i][1] = ny * sg;
        s.n[i][2] = nz * sg;
        s.axis_ok = s.axis_ok && (a[0] * s.ax[0] + a[1] * s.ax[1] + a[2] * s.ax[2] > 0.f);
    }
    s.scale = m;
}

/* true when the box [lo, hi] (world space) is provably outside the pyramid */
__device__ __forceinline__ bool
shaft_misses_box(const ShaftF &s, const float4 lo, const float4 hi, float ox, float oy, float oz)
{
    const float l[3] = { lo.x - ox, lo.y - oy, lo.z - oz }, h[3] = { hi.x - ox, hi.y - oy, hi.z - oz };
    float ext = 0.f;
    for (int k = 0; k < 3; ++k) {
        ext = fmaxf(ext, fmaxf(fabsf(l[k]), fabsf(h[k])));
    }
    if (!(ext < 3.0e38f)) {
        return false; /* unbounded box */
    }
    /* tolerance: FP32 evaluation of products of magnitude scale^2 * ext, and the corners carry the jitter's rounding */
    const float tol1 = 1e-5f * s.scale * s.scale * ext + 1e-30f;
    for (int i = 0; i < 4; ++i) { /* farthest box corner along the inward normal still outside? */
        const float v = fmaxf(s.n[i][0] * l[0], s.n[i][0] * h[0]) + fmaxf(s.n[i][1] * l[1], s.n[i][1] * h[1]) +
                        fmaxf(s.n[i][2] * l[2], s.n[i][2] * h[2]);
        if (v < -tol1) {
            return true;
        }
    }
    if (s.axis_ok) {
        const float v = fmaxf(s.ax[0] * l[0], s.ax[0] * h[0]) + fmaxf(s.ax[1] * l[1], s.ax[1] * h[1]) + fmaxf(s.ax[2] * l[2], s.ax[2] * h[2]);
        if (v < -1e-5f * s.scale * ext) {
            return true; /* wholly behind the origin */
        }
    }
    const float tol2 = 1e-6f * (ext + s.scale);
    for (int k = 0; k < 3; ++k) { /* the box's own faces: origin beyond a face and no ray heading back towards it */
        if ((l[k] > tol2 && s.dmax[k] <= 0.f) || (h[k] < -tol2 && s.dmin[k] >= 0.f)) {
            return true;
        }
    }
    return false;
}

/* true when the WORLD-space ball sp = {centre, radius} is provably outside the pyramid: its centre lies farther than the
 * (inflated) radius beyond one of the pyramid's sides, or beyond the plane through o that every ray leaves */
__device__ __forceinline__ bool
shaft_misses_sphere(const ShaftF &s, const float4 sp, float ox, float oy, float oz)
{
    const float mx = sp.x - ox, my = sp.y - oy, mz = sp.z - oz;
    const float mm = fmaxf(fabsf(mx), fmaxf(fabsf(my), fabsf(mz)));
    /* FP32 cross products of corner directions of magnitude `scale`: the plane distances carry ~1e-6 |m| / sin(light's
     * angular size); the inflation is two orders above that for any light wider than a few milliradians */
    const float r = fmaf(sp.w, 1.001f, 1e-4f * (mm + s.scale));
    const float r2 = r * r;
    for (int i = 0; i < 4; ++i) {
        const float dist = s.n[i][0] * mx + s.n[i][1] * my + s.n[i][2] * mz;
        const float nn = s.n[i][0] * s.n[i][0] + s.n[i][1] * s.n[i][1] + s.n[i][2] * s.n[i][2];
        if (dist < 0.f && dist * dist > r2 * nn * 1.0001f) { /* a dropped plane has n = 0: never true */
            return true;
        }
    }
    if (s.axis_ok) {
        const float dist = s.ax[0] * mx + s.ax[1] * my + s.ax[2] * mz;
        const float aa = s.ax[0] * s.ax[0] + s.ax[1] * s.ax[1] + s.ax[2] * s.ax[2];
        if (dist < 0.f && dist * dist > r2 * aa * 1.0001f) {
            return true;
        }
    }
    return false;
}

/* a ray in some frame, ready for slab tests */
struct FrameF {
    float ox, oy, oz, dx, dy, dz;
    float ix, iy, iz;    /* 1 / d_k, or 0 on an axis left open */
    float c1x, c1y, c1z; /* |1/d_k| eo_k, or +inf on an axis left open */
    float c2x, c2y, c2z; /* 2 |1/d_k| ed_k + G */
    float eo, ed;        /* max_k eo_k, max_k ed_k (sphere) */
};

__device__ __forceinline__ float
rcpf_fast(float x)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r; /* <= 1 ulp: covered by G */
}

__device__ __forceinline__ void
frame_axis(float d, float eo, float ed, float &inv, float &c1, float &c2)
{
    const bool ok = fabsf(d) >= FRT_EPS_F + 2.0f * ed;
    const float r = rcpf_fast(d);
    inv = ok ? r : 0.0f;
    c1 = ok ? fabsf(r) * eo : CUDART_INF_F;
    c2 = ok ? fmaf(2.0f * fabsf(r), ed, FRT_F32_G) : FRT_F32_G;
}

__device__ __forceinline__ void
frame_finish(FrameF &f, float eox, float eoy, float eoz, float edx, float edy, float edz)
{
    frame_axis(f.dx, eox, edx, f.ix, f.c1x, f.c2x);
    frame_axis(f.dy, eoy, edy, f.iy, f.c1y, f.c2y);
    frame_axis(f.dz, eoz, edz, f.iz, f.c1z, f.c2z);
    f.eo = fmaxf(fmaxf(eox, eoy), eoz);
    f.ed = fmaxf(fmaxf(edx, edy), edz);
}

/* the world ray in the frame of transform xf (xf != 0) */
__device__ __forceinline__ void
frame_local(FrameF &f, const DSceneF &SF, int xf, const FrameF &w, float omax, float eo_w, float ed_w)
{
    const float4 m0 = __ldg(SF.fx + 4 * xf), m1 = __ldg(SF.fx + 4 * xf + 1), m2 = __ldg(SF.fx + 4 * xf + 2);
    const float4 R = __ldg(SF.fx + 4 * xf + 3);
    f.ox = fmaf(m0.x, w.ox, fmaf(m0.y, w.oy, fmaf(m0.z, w.oz, m0.w)));
    f.oy = fmaf(m1.x, w.ox, fmaf(m1.y, w.oy, fmaf(m1.z, w.oz, m1.w)));
    f.oz = fmaf(m2.x, w.ox, fmaf(m2.y, w.oy, fmaf(m2.z, w.oz, m2.w)));
    f.dx = fmaf(m0.x, w.dx, fmaf(m0.y, w.dy, m0.z * w.dz));
    f.dy = fmaf(m1.x, w.dx, fmaf(m1.y, w.dy, m1.z * w.dz));
    f.dz = fmaf(m2.x, w.dx, fmaf(m2.y, w.dy, m2.z * w.dz));
    const float go = fmaf(FRT_F32_G, omax, eo_w), gd = ed_w + FRT_F32_G;
    frame_finish(f, fmaf(R.x, go, FRT_F32_G * fabsf(m0.w)), fmaf(R.y, go, FRT_F32_G * fabsf(m1.w)),
                 fmaf(R.z, go, FRT_F32_G * fabsf(m2.w)), R.x * gd, R.y * gd, R.z * gd);
}

/* entry / exit of the ray through the box [lo, hi] as intervals: tn in [tn_lo, tn_hi], tf in [tf_lo, tf_hi] */
__device__ __forceinline__ void
box_f(const FrameF &f, const float4 lo, const float4 hi, float &tn_lo, float &tn_hi, float &tf_lo, float &tf_hi)
{
    float a = (lo.x - f.ox) * f.ix, b = (hi.x - f.ox) * f.ix;
    float mn = fminf(a, b), mx = fmaxf(a, b);
    float E = fmaf(fmaxf(fabsf(a), fabsf(b)), f.c2x, f.c1x);
    tn_lo = mn - E;
    tn_hi = mn + E;
    tf_lo = mx - E;
    tf_hi = mx + E;
    a = (lo.y - f.oy) * f.iy;
    b = (hi.y - f.oy) * f.iy;
    mn = fminf(a, b);
    mx = fmaxf(a, b);
    E = fmaf(fmaxf(fabsf(a), fabsf(b)), f.c2y, f.c1y);
    tn_lo = fmaxf(tn_lo, mn - E);
    tn_hi = fmaxf(tn_hi, mn + E);
    tf_lo = fminf(tf_lo, mx - E);
    tf_hi = fminf(tf_hi, mx + E);
    a = (lo.z - f.oz) * f.iz;
    b = (hi.z - f.oz) * f.iz;
    mn = fminf(a, b);
    mx = fmaxf(a, b);
    E = fmaf(fmaxf(fabsf(a), fabsf(b)), f.c2z, f.c1z);
    tn_lo = fmaxf(tn_lo, mn - E);
    tn_hi = fmaxf(tn_hi, mn + E);
    tf_lo = fminf(tf_lo, mx - E);
    tf_hi = fminf(tf_hi, mx + E);
    /* an unbounded side gives NaN (inf - inf); fminf / fmaxf drop NaN operands, which leaves that side open */
}

/* The crossings of one operand (a leaf, or a closed CSG) as ONE interval: enter in [a_lo, a_hi], exit in [b_lo, b_hi],
 * enter < exit for sure.  flags: 1 = present, 2 / 4 = the enter / exit surface casts shadows. */
template <typename T>
struct SpanT {
    T a_lo, a_hi, b_lo, b_hi;
    int flags;
};
typedef SpanT<float> SpanF;   /* one ray, FP32 rounding intervals */
typedef SpanT<double> SpanD;  /* every shadow ray of a hit at once (trace_shadow_bulk) */

/*
 * csg_filter_intersections (csg.c:43-71) for two operands that are one interval each.  Walking the merged, sorted
 * crossings with the reference's inl / inr toggles gives, per ordering of the four ends:
 *     union         disjoint -> both intervals (2 spans: undecided here)     otherwise -> [first enter, last exit]
 *     intersection  disjoint -> nothing                                      otherwise -> [later enter, earlier exit]
 *     difference    disjoint -> L     L inside R -> nothing     R inside L -> 2 spans (undecided)
 *                   L enters first -> [L enter, R enter]        R enters first -> [R exit, L exit]
 * Each end keeps the casts_shadow bit of the leaf it came from.  Returns false when an ordering is not decided or the
 * result is two intervals.
 */
template <typename SP>
__device__ __forceinline__ bool
csg_combine(int op, const SP &L, const SP &R, SP &out)
{
    out.flags = 0;
    if (!(L.flags & 1)) {
        if (op == FRT_CSG_UNION) {
            out = R;
        }
        return true; /* intersection, difference: nothing survives without the left operand */
    }
    if (!(R.flags & 1)) {
        if (op != FRT_CSG_INTERSECT) {
            out = L;
        }
        return true;
    }
    if (L.b_hi < R.a_lo || R.b_hi < L.a_lo) { /* disjoint for sure */
        if (op == FRT_CSG_DIFFERENCE) {
            out = L;
            return true;
        }
        return op == FRT_CSG_INTERSECT; /* intersection: empty; union: two spans */
    }
    if (!(L.a_hi < R.b_lo) || !(R.a_hi < L.b_lo)) {
        return false; /* neither surely disjoint nor surely overlapping */
    }
    bool l_enters_first, l_exits_last;
    if (L.a_hi < R.a_lo) {
        l_enters_first = true;
    } else if (R.a_hi < L.a_lo) {
        l_enters_first = false;
    } else {
        return false;
    }
    if (R.b_hi < L.b_lo) {
        l_exits_last = true;
    } else if (L.b_hi < R.b_lo) {
        l_exits_last = false;
    } else {
        return false;
    }
    const SP &first = l_enters_first ? L : R, &second = l_enters_first ? R : L;
    const SP &last = l_exits_last ? L : R, &inner = l_exits_last ? R : L;
    if (op == FRT_CSG_UNION) {
        out.a_lo = first.a_lo;
        out.a_hi = first.a_hi;
        out.b_lo = last.b_lo;
        out.b_hi = last.b_hi;
        out.flags = 1 | (first.flags & 2) | (last.flags & 4);
        return true;
    }
    if (op == FRT_CSG_INTERSECT) {
        out.a_lo = second.a_lo;
        out.a_hi = second.a_hi;
        out.b_lo = inner.b_lo;
        out.b_hi = inner.b_hi;
        out.flags = 1 | (second.flags & 2) | (inner.flags & 4);
        return true;
    }
    /* difference L - R */
    if (l_enters_first && l_exits_last) {
        return false; /* R strictly inside L: two spans */
    }
    if (!l_enters_first && !l_exits_last) {
        return true; /* L inside R: nothing */
    }
    if (l_enters_first) { /* [L enter, R enter]: the exit surface is R's entry face */
        out.a_lo = L.a_lo;
        out.a_hi = L.a_hi;
        out.b_lo = R.a_lo;
        out.b_hi = R.a_hi;
        out.flags = 1 | (L.flags & 2) | ((R.flags & 2) << 1);
    } else { /* [R exit, L exit] */
        out.a_lo = R.b_lo;
        out.a_hi = R.b_hi;
        out.b_lo = L.b_lo;
        out.b_hi = L.b_hi;
        out.flags = 1 | ((R.flags & 4) >> 1) | (L.flags & 4);
    }
    return true;
}

/*
 * The reference's verdict on a crossing list that is one span (group.c:105-123, renderer.c:87-90):
 *   the search stops here iff some t is not <= 0, i.e. iff exit > 0;
 *   the point is shadowed iff some t > 0 on a casts_shadow surface is < distance.
 * Returns 0 = does not stop (keep walking), 1 = stops, lit, 2 = stops, shadowed, 3 = undecided.
 */
template <typename T>
__device__ __forceinline__ int
judge_span(const SpanT<T> &s, T D_lo, T D_hi)
{
    if (s.b_hi <= (T)0) {
        return 0;
    }
    if (!(s.b_lo > (T)0)) {
        return 3;
    }
    /* per end: 2 = surely a positive casting crossing nearer than the light, 0 = surely not, 1 = cannot tell */
    int ea, eb;
    if (!(s.flags & 2) || s.a_hi <= (T)0 || s.a_lo >= D_hi) {
        ea = 0;
    } else {
        ea = (s.a_lo > (T)0 && s.a_hi < D_lo) ? 2 : 1;
    }
    if (!(s.flags & 4) || s.b_lo >= D_hi) {
        eb = 0;
    } else {
        eb = (s.b_hi < D_lo) ? 2 : 1;
    }
    if (ea == 2 || eb == 2) {
        return 2;
    }
    return (ea | eb) ? 3 : 1;
}

__device__ __forceinline__ bool sphere_span(const FrameF &f, float r2, float dr2, SpanF &s);

/* one cube or sphere leaf as a span; false = undecided */
__device__ __forceinline__ bool
leaf_span(int type, const float4 lo, const float4 hi, const FrameF &f, SpanF &s)
{
    s.flags = 0;
    if (type == FRT_CUBE) { /* cube_local_intersect, cube.c:56-78 */
        box_f(f, lo, hi, s.a_lo, s.a_hi, s.b_lo, s.b_hi);
        if (s.a_lo > s.b_hi) {
            return true; /* tmin > tmax for sure: no crossings */
        }
        if (!(s.a_hi < s.b_lo)) {
            return false;
        }
        s.flags = 1;
        return true;
    }
    return sphere_span(f, 1.0f, 0.0f, s);
}

/*
 * sphere_local_intersect (sphere.c:14-40) for the ray of frame f against the ball |x|^2 = r2 (r2 known up to dr2): the unit
 * sphere of a leaf's own frame (r2 = 1), or a WORLD-space ball with f = the world ray moved to its centre -- a similarity
 * transform scales a, b and c of the quadratic alike, the roots are the reference's.  false = undecided.
 */
__device__ __forceinline__ bool
sphere_span(const FrameF &f, float r2, float dr2, SpanF &s)
{
    s.flags = 0;
    const float So = fabsf(f.ox) + fabsf(f.oy) + fabsf(f.oz), Sd = fabsf(f.dx) + fabsf(f.dy) + fabsf(f.dz);
    const float a = fmaf(f.dx, f.dx, fmaf(f.dy, f.dy, f.dz * f.dz));
    const float hb = fmaf(f.dx, f.ox, fmaf(f.dy, f.oy, f.dz * f.oz));
    const float habs = fmaf(fabsf(f.dx), fabsf(f.ox), fmaf(fabsf(f.dy), fabsf(f.oy), fabsf(f.dz * f.oz)));
    const float oo = fmaf(f.ox, f.ox, fmaf(f.oy, f.oy, f.oz * f.oz));
    const float b = 2.0f * hb, c = oo - r2;
    const float da = fmaf(2.0f * Sd + 3.0f * f.ed, f.ed, FRT_F32_G * a);
    const float db = 2.0f * (fmaf(Sd, f.eo, fmaf(So, f.ed, 3.0f * f.eo * f.ed)) + FRT_F32_G * habs);
    const float dc = fmaf(2.0f * So + 3.0f * f.eo, f.eo, FRT_F32_G * (oo + r2)) + dr2;
    const float disc = fmaf(b, b, -4.0f * a * c);
    const float dD = fmaf(2.0f * fabsf(b) + db, db, 4.0f * (fmaf(a, dc, fmaf(fabsf(c), da, da * dc)))) +
                     FRT_F32_G * fmaf(b, b, 4.0f * a * fabsf(c));
    if (disc + dD < 0.0f) {
        return true;
    }
    if (!(disc > 4.0f * dD) || !(a > 4.0f * da)) {
        return false;
    }
    const float sq = sqrtf(disc);
    const float ds = dD / sq + FRT_F32_G * sq;
    /* Cancellation-free roots: q = -(b + sign(b) sqrt(disc)) / 2, roots q / a and c / q.  (-b + sqrt(disc)) loses every
     * digit for a ray that starts on the sphere (c ~ 1e-5), which is every shadow ray of a hit on the sphere itself. */
    const float q = -0.5f * (b + copysignf(sq, b));
    const float dq = 0.5f * (db + ds) + FRT_F32_G * fabsf(q);
    const float ia = rcpf_fast(a), iq = rcpf_fast(q);
    const float tb = q * ia, ts = c * iq;                 /* the root far from / near zero */
    /* first-order bounds, times 1.5 for the second-order terms (a > 4 da and |q| > 4 dq are required) */
    const float eb = 1.5f * fmaf(fabsf(tb), da * ia + 2.0f * FRT_F32_G, dq * ia);
    const float es = 1.5f * fmaf(fabsf(ts), dq * fabsf(iq) + 2.0f * FRT_F32_G, dc * fabsf(iq));
    if (!(fabsf(q) > 4.0f * dq)) {
        return false;
    }
    const bool big_first = tb < ts;
    const float t0 = big_first ? tb : ts, e0 = big_first ? eb : es;
    const float t1 = big_first ? ts : tb, e1 = big_first ? es : eb;
    s.a_lo = t0 - e0;
    s.a_hi = t0 + e0;
    s.b_lo = t1 - e1;
    s.b_hi = t1 + e1;
    if (!(s.a_hi < s.b_lo)) {
        return false;
    }
    s.flags = 1;
    return true;
}

/* plane_local_intersect (plane.c:11-25) as a degenerate span: one crossing */
__device__ __forceinline__ void
plane_span(const FrameF &f, SpanF &s)
{
    const float tt = -f.oy * f.iy;
    const float E = fmaf(fabsf(tt), f.c2y, f.c1y);
    s.a_lo = s.b_lo = tt - E;
    s.a_hi = s.b_hi = tt + E;
    s.flags = 1;
}

/*
 * One leaf of the tree as a span; false = undecided.  WORLD leaves (xf = 0) are tested with the world frame `w`, the
 * others with `lf`, the frame of transform cur_xf, which is rebuilt only when the leaf's transform differs from it --
 * two call sites instead of one frame that is copied back and forth.
 */
template <bool COUNT>
__device__ __forceinline__ bool
node_span(const DSceneF &SF, const float4 q0, const float4 lo, const float4 hi, const FrameF &w, FrameF &lf, int &cur_xf, float omax,
          float eo_w, float ed_w, bool standalone, SpanF &s, unsigned int &cost)
{
    const int flags = __float_as_int(q0.x), xf = __float_as_int(q0.z);
    const int type = flags & FRT_FN_TYPE_MASK;
    s.flags = 0;
    if (!(flags & FRT_FN_FAST) || (type == FRT_PLANE && !standalone)) {
        return false; /* no fast form; a plane (one crossing) inside a CSG */
    }
    if (xf == 0) {
        if (type == FRT_PLANE) {
            plane_span(w, s);
        } else if (!leaf_span(type, lo, hi, w, s)) {
            return false;
        }
    } else {
        if (xf != cur_xf) {
            cur_xf = xf;
            frame_local(lf, SF, xf, w, omax, eo_w, ed_w);
            if (COUNT) cost += FRT_COST_XFORM;
        }
        if (type == FRT_PLANE) {
            plane_span(lf, s);
        } else if (!leaf_span(type, lo, hi, lf, s)) {
            return false;
        }
    }
    if (s.flags && (flags & FRT_FN_CASTS)) {
        s.flags |= 6;
    }
    return true;
}

/*
 * FP32 filter twin of trace_shadow (frt_device.cuh).  `w` is the world frame of the ray (frame_finish'ed by the
 * caller), omax / eo_w / ed_w its error terms, [D_lo, D_hi] the interval of the light distance; `fnodes` is the
 * node mirror (in shared memory when the tree is small), `relevant` the hit's shaft-culling mask over nodes 0..31.
 *
 * An outermost CSG node is evaluated from its postfix program (SF.csg_prog, built at upload).  Only LEFT-DEEP trees
 * have one -- ((A op B) op C) op D ..., which is how the scenes build windows, lenses and gears -- so the program is
 * "leaf, then (leaf, operator) pairs" and one accumulator span suffices; anything else is left to FP64.  The walk
 * itself therefore never nests: every step is a group (cull), a CSG (cull, then the program) or a leaf.  The
 * reference's culls of the CSG nodes INSIDE another CSG (csg.c:82-86) are skipped: a nested CSG whose box is missed
 * contributes no crossings either way.
 * Returns FRT_SH_LIT, FRT_SH_SHADOWED or FRT_SH_UNDECIDED (the latter with a reason code in bits 4.. for the counting
 * build's histogram; callers mask with 15).
 */
template <bool COUNT>
__device__ __forceinline__ int
trace_shadow_f32(const DSceneF &SF, const float4 *fnodes, int root, int start, int tail, unsigned int relevant, const FrameF &w, float omax,
                 float eo_w,
                 float ed_w, float D_lo, float D_hi, unsigned long long *nodes_visited, unsigned long long *flops)
{
    unsigned int visited = 0, cost = 0;
    /* start: the root, or the node X at which the shaft walk of this hit / quadrant got stuck; tail: what that walk found
     * for the rays X does not stop (trace_shadow_bulk): 0 nothing, 1 lit, 2 shadowed */
    int i = start;
    const int end = __float_as_int(fnodes[3 * root].y);
    int cur_xf = 0;
    FrameF lf = w; /* the ray in the frame of transform cur_xf, once cur_xf != 0 */
    int verdict = FRT_SH_LIT;

    while (i < end) {
        if (i < 32 && !((relevant >> i) & 1u)) {
            /* shaft culling (see ShaftF): no ray of this hit can cross this subtree at t > 0 */
            i = __float_as_int(fnodes[3 * i].y);
            continue;
        }
        const float4 q0 = fnodes[3 * i], lo = fnodes[3 * i + 1], hi = fnodes[3 * i + 2];
        const int flags = __float_as_int(q0.x), skip = __float_as_int(q0.y);
        const int type = flags & FRT_FN_TYPE_MASK;
        const bool at_start = i == start;
        if (COUNT) {
            ++visited;
            cost += (type >= FRT_CSG) ? FRT_COST_BBOX : prim_cost(type);
        }
        SpanF s;
        if (type >= FRT_CSG) { /* group or CSG: conservative cull by its bounds */
            if (!(flags & FRT_FN_NOCULL)) {
                const int xf = __float_as_int(q0.z);
                float tn_lo, tn_hi, tf_lo, tf_hi;
                if (xf == 0) {
                    box_f(w, lo, hi, tn_lo, tn_hi, tf_lo, tf_hi);
                } else {
                    if (xf != cur_xf) {
                        cur_xf = xf;
                        frame_local(lf, SF, xf, w, omax, eo_w, ed_w);
                        if (COUNT) cost += FRT_COST_XFORM;
                    }
                    box_f(lf, lo, hi, tn_lo, tn_hi, tf_lo, tf_hi);
                }
                /* surely missed, or surely wholly behind the origin (only t <= 0 crossings inside) */
                if (tn_lo > tf_hi || tf_hi < 0.0f) {
                    if (tail && i == start) {
                        verdict = tail == 2 ? FRT_SH_SHADOWED : FRT_SH_LIT;
                        break;
                    }
                    i = skip;
                    continue;
                }
            }
            if (type == FRT_GROUP) {
                i = i + 1;
                continue;
            }
            /* CSG: run its postfix program */
            if (!(flags & FRT_FN_FAST)) {
                return FRT_SH_UNDECIDED | (1 << 4); /* not left-deep, or a group inside an operand */
            }
            int pc = __float_as_int(lo.w);
            const int pc1 = pc + __float_as_int(hi.w);
            {
                const int code = __ldg(SF.csg_prog + pc);
                if (COUNT) {
                    ++visited;
                    cost += prim_cost(__float_as_int(fnodes[3 * code].x) & FRT_FN_TYPE_MASK);
                }
                if (!node_span<COUNT>(SF, fnodes[3 * code], fnodes[3 * code + 1], fnodes[3 * code + 2], w, lf, cur_xf, omax, eo_w, ed_w,
                                      false, s, cost)) {
                    return FRT_SH_UNDECIDED | (5 << 4) | ((code & 31) << 8);
                }
            }
            for (pc += 1; pc < pc1; pc += 2) {
                const int code = __ldg(SF.csg_prog + pc), op = -__ldg(SF.csg_prog + pc + 1) - 1;
                if (COUNT) {
                    ++visited;
                    cost += prim_cost(__float_as_int(fnodes[3 * code].x) & FRT_FN_TYPE_MASK);
                }
                SpanF t, r;
                if (!node_span<COUNT>(SF, fnodes[3 * code], fnodes[3 * code + 1], fnodes[3 * code + 2], w, lf, cur_xf, omax, eo_w, ed_w,
                                      false, t, cost)) {
                    return FRT_SH_UNDECIDED | (5 << 4) | ((code & 31) << 8);
                }
                r.a_lo = r.a_hi = r.b_lo = r.b_hi = 0.0f;
                if (!csg_combine(op, s, t, r)) {
                    return FRT_SH_UNDECIDED | (8 << 4);
                }
                s = r;
            }
            i = skip;
        } else {
            if (!node_span<COUNT>(SF, q0, lo, hi, w, lf, cur_xf, omax, eo_w, ed_w, true, s, cost)) {
                return FRT_SH_UNDECIDED | (5 << 4) | ((i & 31) << 8);
            }
            i = i + 1;
        }
        if (s.flags) { /* a leaf's or an outermost CSG's crossing list */
            const int v = judge_span(s, D_lo, D_hi);
            if (v == 3) {
                return FRT_SH_UNDECIDED | (6 << 4) | (((i - 1) & 31) << 8);
            }
            if (v != 0) {
                verdict = (v == 2) ? FRT_SH_SHADOWED : FRT_SH_LIT;
                break;
            }
        }
        if (tail && at_start) { /* X did not stop this ray: the shaft walk knows the rest */
            verdict = tail == 2 ? FRT_SH_SHADOWED : FRT_SH_LIT;
            break;
        }
    }
    if (COUNT) {
        *nodes_visited += visited;
        *flops += cost;
    }
    return verdict;
}

/*
 * csg_combine without a data-dependent branch: inside k_shadow_entry every lane of a warp evaluates the SAME operand pair
 * of the same CSG program for a different ray, so the operator is warp-uniform and only the orderings differ from lane
 * to lane -- they become selects.  Same case analysis, same results as csg_combine (the templates are compared on every
 * fixture through FRT_FLAG_VERIFY_F32).
 */
__device__ __forceinline__ bool
csg_combine_sel(int op, const SpanF &L, const SpanF &R, SpanF &out)
{
    const bool hasL = (L.flags & 1) != 0, hasR = (R.flags & 1) != 0;
    const bool disj = L.b_hi < R.a_lo || R.b_hi < L.a_lo;         /* disjoint for sure */
    const bool over = L.a_hi < R.b_lo && R.a_hi < L.b_lo;         /* overlapping for sure */
    const bool lf = L.a_hi < R.a_lo, rf = R.a_hi < L.a_lo;        /* who enters first */
    const bool ll = R.b_hi < L.b_lo, rl = L.b_hi < R.b_lo;        /* who exits last */
    const bool ordered = over && (lf || rf) && (ll || rl);
    bool ok;
    /* the result's ends: enter from span `ea` end `ea_b` (false: its a-end, true: its b-end), likewise the exit */
    bool present, a_from_L, a_is_b, b_from_L, b_is_a;
    if (op == FRT_CSG_UNION) {
        /* both: [first enter, last exit]; one absent: the other */
        ok = !(hasL && hasR) || (!disj && ordered);
        present = hasL || hasR;
        a_from_L = hasL && (!hasR || lf);
        b_from_L = hasL && (!hasR || ll);
        a_is_b = false;
        b_is_a = false;
    } else if (op == FRT_CSG_INTERSECT) {
        /* both and overlapping: [later enter, earlier exit]; anything else: nothing */
        ok = !(hasL && hasR) || disj || ordered;
        present = hasL && hasR && !disj;
        a_from_L = !lf;
        b_from_L = !ll;
        a_is_b = false;
        b_is_a = false;
    } else { /* difference L - R */
        const bool both = hasL && hasR && !disj;
        ok = !both || (ordered && !(lf && ll));                    /* R strictly inside L: two spans */
        present = hasL && !(both && rf && rl);                     /* L inside R: nothing */
        /* L alone, or disjoint: L.  L enters first: [L enter, R enter].  R enters first: [R exit, L exit]. */
        a_from_L = !both || lf;
        a_is_b = both && !lf;      /* R's exit */
        b_from_L = !both || !lf;
        b_is_a = both && lf;       /* R's entry */
    }
    const SpanF &A = a_from_L ? L : R, &B = b_from_L ? L : R;
    out.a_lo = a_is_b ? A.b_lo : A.a_lo;
    out.a_hi = a_is_b ? A.b_hi : A.a_hi;
    out.b_lo = b_is_a ? B.a_lo : B.b_lo;
    out.b_hi = b_is_a ? B.a_hi : B.b_hi;
    const int fa = a_is_b ? ((A.flags & 4) >> 1) : (A.flags & 2);
    const int fb = b_is_a ? ((B.flags & 2) << 1) : (B.flags & 4);
    out.flags = present ? (1 | fa | fb) : 0;
    return ok;
}

/*
 * What the shaft walk of a (hit, quadrant) leaves to its rays: a PROGRAM of up to FRT_PROG_MAX nodes X1 < X2 < X3 at which
 * the walk could not tell (in the reference's order; every node between and before them was passed with a decided "ends
 * no search" for every ray of the shaft) and the verdict the walk reached for the rays none of them stops (trace_shadow_bulk):
 *     bits 0..1  tail: 1 lit, 2 shadowed, 0 = the walk got stuck more often than the program holds (general walk from X1)
 *     bits 2..3  number of nodes (1..3)
 *     bits 4..8, 9..13, 14..18  X1, X2, X3
 *     bit 19     the rays can run the program without the tree walk (entry_program_is_fast, decided once per entry)
 * A ray evaluates X1, X2, ... in turn and takes the first one's verdict that stops it, else the tail (trace_entry_program).
 * Nodes it can evaluate without the tree walk (DSceneF::entry_fast): WORLD-space cube leaves, outermost CSGs whose program
 * runs over WORLD-space cube leaves (every wall, box and window of the Cornell scene: X's own bounds are not tested -- a
 * ray that misses them misses every operand inside) and WORLD-space balls (DSceneF::wsphere).
 */
#define FRT_PROG_MAX 3
#define FRT_PROG_FAST (1u << 19) /* every node of the program is in DSceneF::entry_fast and the tail verdict is known: set by the shaft kernels */
#define FRT_PROG_ROOT(root) ((1u << 2) | ((unsigned int)(root) << 4)) /* one node, no tail: the general walk from `root` */
#define FRT_PROG_TAIL(p) ((int)((p) & 3u))
#define FRT_PROG_COUNT(p) ((int)(((p) >> 2) & 3u))
#define FRT_PROG_NODE(p, k) ((int)(((p) >> (4 + 5 * (k))) & 31u))

/* one program node as a span (flags 0: the ray does not cross it); false = undecided */
template <bool COUNT>
__device__ __forceinline__ bool
entry_node_span(const DSceneF &SF, const float4 *fnodes, const int *cprog, int node, const FrameF &w, float omax, float eo_o, SpanF &s,
                unsigned int &visited, unsigned int &cost)
{
    const float4 q0 = fnodes[3 * node];
    const int flags = __float_as_int(q0.x), type = flags & FRT_FN_TYPE_MASK;
    bool ok = true;
    if (type == FRT_CSG) {
        int pc = __float_as_int(fnodes[3 * node + 1].w);
        const int pc1 = pc + __float_as_int(fnodes[3 * node + 2].w);
        {
            const int code = cprog[pc];
            const int lf = __float_as_int(fnodes[3 * code].x);
            box_f(w, fnodes[3 * code + 1], fnodes[3 * code + 2], s.a_lo, s.a_hi, s.b_lo, s.b_hi);
            const bool miss = s.a_lo > s.b_hi;
            ok = miss || s.a_hi < s.b_lo;
            s.flags = miss ? 0 : ((lf & FRT_FN_CASTS) ? 7 : 1);
        }
        for (pc += 1; pc < pc1; pc += 2) {
            const int code = cprog[pc], op = -cprog[pc + 1] - 1;
            const int lf = __float_as_int(fnodes[3 * code].x);
            SpanF t, r;
            box_f(w, fnodes[3 * code + 1], fnodes[3 * code + 2], t.a_lo, t.a_hi, t.b_lo, t.b_hi);
            const bool miss = t.a_lo > t.b_hi;
            ok = ok && (miss || t.a_hi < t.b_lo);
            t.flags = miss ? 0 : ((lf & FRT_FN_CASTS) ? 7 : 1);
            ok = csg_combine_sel(op, s, t, r) && ok;
            s = r;
        }
        if (COUNT) {
            const int n_ops = (pc1 - __float_as_int(fnodes[3 * node + 1].w) + 1) / 2;
            visited += 1 + n_ops;
            cost += FRT_COST_BBOX + n_ops * prim_cost(FRT_CUBE);
        }
    } else if (type == FRT_SPHERE) {
        /* the world ray moved to the ball's centre: o - c carries the rounding of o (eo_o), of c (u |c|) and of the difference */
        const float4 sp = __ldg(SF.wsphere + node);
        FrameF f;
        f.ox = w.ox - sp.x;
        f.oy = w.oy - sp.y;
        f.oz = w.oz - sp.z;
        f.dx = w.dx;
        f.dy = w.dy;
        f.dz = w.dz;
        const float cmax = fmaxf(fmaxf(fabsf(sp.x), fabsf(sp.y)), fabsf(sp.z));
        f.eo = eo_o + 2.0f * FRT_F32_U * (omax + cmax);
        f.ed = w.ed;
        const float r2 = sp.w * sp.w;
        ok = sphere_span(f, r2, 4.0f * FRT_F32_U * r2, s);
        if (s.flags && (flags & FRT_FN_CASTS)) {
            s.flags |= 6;
        }
        if (COUNT) {
            visited += 1;
            cost += prim_cost(FRT_SPHERE);
        }
    } else {
        box_f(w, fnodes[3 * node + 1], fnodes[3 * node + 2], s.a_lo, s.a_hi, s.b_lo, s.b_hi);
        const bool miss = s.a_lo > s.b_hi;
        ok = miss || s.a_hi < s.b_lo;
        s.flags = miss ? 0 : ((flags & FRT_FN_CASTS) ? 7 : 1);
        if (COUNT) {
            visited += 1;
            cost += prim_cost(FRT_CUBE);
        }
    }
    return ok;
}

/* the program of a pending entry for one ray; the caller checked that every node is in DSceneF::entry_fast and tail != 0.
 * Returns FRT_SH_* like trace_shadow_f32 (reason / node of an undecided ray in bits 4.. for the counting build). */
template <bool COUNT>
__device__ __forceinline__ int
trace_entry_program(const DSceneF &SF, const float4 *fnodes, const int *cprog, unsigned int prog, const FrameF &w, float omax, float eo_o,
                    float D_lo, float D_hi, unsigned long long *nodes_visited, unsigned long long *flops)
{
    const int n = FRT_PROG_COUNT(prog);
    unsigned int visited = 0, cost = 0;
    int res = FRT_PROG_TAIL(prog) == 2 ? FRT_SH_SHADOWED : FRT_SH_LIT; /* the rays no node stops */
    for (int k = 0; k < n; ++k) {
        const int node = FRT_PROG_NODE(prog, k);
        SpanF s;
        if (!entry_node_span<COUNT>(SF, fnodes, cprog, node, w, omax, eo_o, s, visited, cost)) {
            res = FRT_SH_UNDECIDED | (5 << 4) | ((node & 31) << 8);
            break;
        }
        if (s.flags) {
            const int v = judge_span(s, D_lo, D_hi);
            if (v == 3) {
                res = FRT_SH_UNDECIDED | (6 << 4) | ((node & 31) << 8);
                break;
            }
            if (v != 0) {
                res = v == 2 ? FRT_SH_SHADOWED : FRT_SH_LIT;
                break;
            }
        }
    }
    if (COUNT) {
        *nodes_visited += visited;
        *flops += cost;
    }
    return res;
}

/* are all nodes of the program evaluable by entry_node_span, and does it carry a tail verdict? */
__device__ __forceinline__ bool
entry_program_is_fast(unsigned int prog, unsigned int entry_fast)
{
    const int n = FRT_PROG_COUNT(prog);
    bool ok = FRT_PROG_TAIL(prog) != 0 && n >= 1;
    for (int k = 0; k < FRT_PROG_MAX; ++k) {
        ok = ok && (k >= n || ((entry_fast >> FRT_PROG_NODE(prog, k)) & 1u));
    }
    return ok;
}

/* can node `i` be evaluated by entry_node_span?  (decided once per scene on the host: bit i of DSceneF::entry_fast) */
static inline bool
node_is_entry_fast(const float4 *fn, const int *prog, const float4 *wsph, int i)
{
    auto as_int = [](float f) { int v; memcpy(&v, &f, sizeof(v)); return v; };
    const int flags = as_int(fn[3 * i].x), type = flags & FRT_FN_TYPE_MASK;
    auto world_cube = [&](int k) {
        const int f = as_int(fn[3 * k].x);
        return (f & FRT_FN_TYPE_MASK) == FRT_CUBE && (f & FRT_FN_FAST) && (f & FRT_FN_WORLD) && as_int(fn[3 * k].z) == 0;
    };
    if (type == FRT_CUBE) {
        return world_cube(i);
    }
    if (type == FRT_SPHERE) {
        return wsph[i].w > 0.f;
    }
    if (type != FRT_CSG || !(flags & FRT_FN_FAST)) {
        return false;
    }
    const int pc = as_int(fn[3 * i + 1].w), len = as_int(fn[3 * i + 2].w);
    for (int k = pc; k < pc + len; ++k) {
        if (prog[k] >= 0 && !world_cube(prog[k])) {
            return false;
        }
    }
    return true;
}

/*
 * All shadow rays of ONE hit at once.  They share the origin o and aim at points of the light's parallelogram, so in
 * the parametrisation  o + t (p - o)  (t = 1 is the light; t against 0, t against the light distance and the order of
 * two crossings are invariant under the per-ray scale |p - o|) every direction component d_k lies in
 * [min, max] over the four corners.  The slab values of a WORLD-space box are then intervals over the whole family:
 *   d_k of one sign on every ray:   (b - o_k) / d_k is monotone in d_k, the ends are attained at the interval's ends;
 *   d_k changes sign (or gets small enough for the reference's "* INFINITY" branch, cube.c:27-33) and the origin is
 *   strictly inside the slab:       near_k <= max(lo/d+, hi/d-) < 0  and  far_k >= min(hi/d+, lo/d-) > 0;
 *   otherwise the axis says nothing.
 * The same walk as trace_shadow_f32 is done ONCE with these intervals; when every comparison on the way separates,
 * all rays of the hit take the same branches and get the same verdict (csg_combine / judge_span are shared with the
 * per-ray filter), and none of them has to be traced: on the Cornell frame that is every hit in the umbra of the
 * window wall, four fifths of the frame.  Anything else -- a sphere, a local frame, an overlap -- leaves the hit to
 * the per-ray kernels.  Arithmetic is FP64 on the FP64 over-point; the only error of note is the FP32 rounding of the
 * mirror's box bounds (en), the rest is covered by relative slacks far above FP64 rounding and far below any feature.
 * Checked like the per-ray filter: FRT_FLAG_VERIFY_F32 traces every ray of a bulk-decided hit in FP64 as well.
 */
/*
 * The shaft arithmetic exists in two precisions.  double: the first version (every quantity exact to 1e-12, the only
 * error of note the FP32 rounding of the mirror's box bounds).  float (the default): on sm_100 an FP64 fmin / fmax /
 * select expands to several integer instructions and the walk is made of them (ncu, profiles/r1j_k_shadow_bulk.txt), so
 * the same interval walk in FP32 with every bound pushed outward by the rounding it can have picked up -- relative 1e-6
 * on a quotient (three roundings of 6e-8 and the reciprocal's), absolute 4e-7 (|bounds| + |origin|) on a slab numerator.
 * A shaft only decides when its intervals separate, so wider intervals cost decisions within 1e-6 of a boundary, never
 * correctness; FRT_FLAG_VERIFY_F32 checks every decided ray against the FP64 walk either way.
 */
template <typename T> struct ShaftEps;
template <> struct ShaftEps<double> {
    static __device__ __forceinline__ double rel() { return 1e-12; }   /* relative slack of a quotient */
    static __device__ __forceinline__ double arith() { return 1e-12; } /* relative slack of a difference of coordinates */
    static __device__ __forceinline__ double round32() { return 2.4e-7; } /* mirror bounds rounded to FP32: 2^-22 Bmax */
    static __device__ __forceinline__ double dist() { return 1e-9; }   /* the light sits at t = 1 +- dist */
    static __device__ __forceinline__ double tiny() { return 1e-300; }
    static __device__ __forceinline__ double inf() { return CUDART_INF; }
};
template <> struct ShaftEps<float> {
    static __device__ __forceinline__ float rel() { return 1e-6f; }
    static __device__ __forceinline__ float arith() { return 2e-7f; }
    static __device__ __forceinline__ float round32() { return 4e-7f; }
    static __device__ __forceinline__ float dist() { return 2e-6f; }
    static __device__ __forceinline__ float tiny() { return 1e-30f; }
    static __device__ __forceinline__ float inf() { return CUDART_INF_F; }
};
__device__ __forceinline__ float shaft_max(float a, float b) { return fmaxf(a, b); }
__device__ __forceinline__ double shaft_max(double a, double b) { return fmax(a, b); }
__device__ __forceinline__ float shaft_min(float a, float b) { return fminf(a, b); }
__device__ __forceinline__ double shaft_min(double a, double b) { return fmin(a, b); }
__device__ __forceinline__ float shaft_abs(float a) { return fabsf(a); }
__device__ __forceinline__ double shaft_abs(double a) { return fabs(a); }
__device__ __forceinline__ float shaft_sqrt(float a) { return sqrtf(a); }
__device__ __forceinline__ double shaft_sqrt(double a) { return sqrt(a); }

template <typename T>
struct ShaftT {
    T o[3];
    T ia[3], ib[3]; /* sgn != 0: 1 / min |d_k|, 1 / max |d_k|;  sgn == 0: 1 / max(d_k, tiny), 1 / min(d_k, -tiny) */
    int sgn[3]; /* +1 / -1: d_k has that sign on every ray and the reference divides; 0: see above */
    T en;       /* bound on the error of a slab numerator (b - o_k) */
    float dl[3], dh[3]; /* the box of (unnormalised) ray directions p - o, slack included (shaft_sphere) */
};
typedef ShaftT<double> ShaftD;

template <typename T>
__device__ __forceinline__ void
shaft_d_setup(ShaftT<T> &s, const double *box, const T *over, T orel, float bmax, float smin, float ealign)
{
    /* over: the rays' common origin, known up to orel * |over| per component (the FP32 copy of the over-point in LightTmp:
     * 2^-24; reading it instead of the FP64 record saves the per-hit kernels a 152-byte-strided load) */
    /* box: axis-aligned bounds {min xyz, max xyz} of the light points the rays aim at (the whole light, or one quadrant
     * of its sample grid), measured over every cached sample set at upload and inflated there */
    typedef ShaftEps<T> E;
    T cmax = (T)0, len2max = (T)0, dlo[3], dhi[3];
    for (int k = 0; k < 3; ++k) {
        const T lo = (T)__ldg(box + k), hi = (T)__ldg(box + 3 + k);
        s.o[k] = over[k];
        dlo[k] = lo - over[k];
        dhi[k] = hi - over[k];
        cmax = shaft_max(cmax, shaft_max(shaft_abs(over[k]), shaft_max(shaft_abs(lo), shaft_abs(hi))));
        const T m = shaft_max(shaft_abs(dlo[k]), shaft_abs(dhi[k]));
        len2max += m * m;
    }
    const T sd = (E::arith() + (T)ealign + orel) * (T)2 * cmax + E::tiny();
    /* the reference divides by the LOCAL normalised component when it is >= EPSILON: local = scale * world */
    const T thr = (T)2 * (T)FRT_EPS * shaft_sqrt(len2max) / (T)smin;
    for (int k = 0; k < 3; ++k) {
        dlo[k] -= sd;
        dhi[k] += sd;
        s.dl[k] = (float)dlo[k];
        s.dh[k] = (float)dhi[k];
        s.sgn[k] = dlo[k] > thr ? 1 : (dhi[k] < -thr ? -1 : 0);
        /* the slab quotients become products with these reciprocals (their rounding sits in shaft_box_d's relative slack) */
        if (s.sgn[k] > 0) {
            s.ia[k] = (T)1 / dlo[k];
            s.ib[k] = (T)1 / dhi[k];
        } else if (s.sgn[k] < 0) {
            s.ia[k] = (T)-1 / dhi[k];
            s.ib[k] = (T)-1 / dlo[k];
        } else {
            s.ia[k] = (T)1 / shaft_max(dhi[k], E::tiny());
            s.ib[k] = (T)1 / shaft_min(dlo[k], -E::tiny());
        }
    }
    s.en = E::round32() * ((T)bmax + (sizeof(T) == 4 ? cmax : (T)0)) + (E::arith() + (T)ealign + orel) * cmax + E::tiny();
}

/* quotient range of n in [n_lo, n_hi] over e in [e_lo, e_hi], e_lo > 0, given ia = 1 / e_lo and ib = 1 / e_hi */
template <typename T>
__device__ __forceinline__ void
shaft_div(T n_lo, T n_hi, T ia, T ib, T &q_lo, T &q_hi)
{
    q_lo = n_lo * (n_lo >= (T)0 ? ib : ia);
    q_hi = n_hi * (n_hi >= (T)0 ? ia : ib);
}

/* entry / exit of every ray of the shaft through the world box [lo, hi], as intervals */
template <typename T>
__device__ __forceinline__ void
shaft_box_d(const ShaftT<T> &s, const float4 lo, const float4 hi, T &tn_lo, T &tn_hi, T &tf_lo, T &tf_hi)
{
    typedef ShaftEps<T> E;
    const float l[3] = { lo.x, lo.y, lo.z }, h[3] = { hi.x, hi.y, hi.z };
    tn_lo = tn_hi = -E::inf();
    tf_lo = tf_hi = E::inf();
    for (int k = 0; k < 3; ++k) {
        const T nl = (T)l[k] - s.o[k], nh = (T)h[k] - s.o[k];
        T a_lo = -E::inf(), a_hi = E::inf(), b_lo = -E::inf(), b_hi = E::inf();
        if (s.sgn[k] > 0) {
            shaft_div(nl - s.en, nl + s.en, s.ia[k], s.ib[k], a_lo, a_hi);
            shaft_div(nh - s.en, nh + s.en, s.ia[k], s.ib[k], b_lo, b_hi);
        } else if (s.sgn[k] < 0) { /* near = hi / d = (-hi) / (-d) */
            shaft_div(-nh - s.en, -nh + s.en, s.ia[k], s.ib[k], a_lo, a_hi);
            shaft_div(-nl - s.en, -nl + s.en, s.ia[k], s.ib[k], b_lo, b_hi);
        } else if (nl + s.en < (T)0 && nh - s.en > (T)0) { /* ia = 1 / (largest positive d), ib = 1 / (most negative d) */
            a_hi = shaft_max((nl + s.en) * s.ia[k], (nh - s.en) * s.ib[k]);
            b_lo = shaft_min((nh - s.en) * s.ia[k], (nl + s.en) * s.ib[k]);
        }
        tn_lo = shaft_max(tn_lo, a_lo);
        tn_hi = shaft_max(tn_hi, a_hi);
        tf_lo = shaft_min(tf_lo, b_lo);
        tf_hi = shaft_min(tf_hi, b_hi);
    }
    /* the reference evaluates the same quotients in the leaf's frame (FP64 rounding of either side); in FP32 the products
     * and reciprocals above carry their own roundings */
    tn_lo -= E::rel() * shaft_abs(tn_lo);
    tn_hi += E::rel() * shaft_abs(tn_hi);
    tf_lo -= E::rel() * shaft_abs(tf_lo);
    tf_hi += E::rel() * shaft_abs(tf_hi);
}

/* a WORLD cube leaf over the shaft; false = undecided */
template <typename T>
__device__ __forceinline__ bool
shaft_leaf_span(const ShaftT<T> &sh, const float4 q0, const float4 lo, const float4 hi, SpanT<T> &s)
{
    const int flags = __float_as_int(q0.x);
    s.flags = 0;
    s.a_lo = s.a_hi = s.b_lo = s.b_hi = (T)0;
    if ((flags & FRT_FN_TYPE_MASK) != FRT_CUBE || !(flags & FRT_FN_FAST) || !(flags & FRT_FN_WORLD) || __float_as_int(q0.z) != 0) {
        return false;
    }
    shaft_box_d(sh, lo, hi, s.a_lo, s.a_hi, s.b_lo, s.b_hi);
    if (s.a_lo > s.b_hi) {
        return true; /* missed by every ray */
    }
    if (!(s.a_hi < s.b_lo)) {
        return false;
    }
    s.flags = 1 | ((flags & FRT_FN_CASTS) ? 6 : 0);
    return true;
}

/*
 * A WORLD-space ball over the shaft (sphere_local_intersect, sphere.c:14-40, for every ray o + t (p - o), p in the box of
 * light points).  Returns 0 = no ray has a crossing at t > 0, 1 / 2 = every ray ends its search here, lit / shadowed,
 * 3 = cannot tell.  FP32 with margins of 2e-5 on every cosine: the quantities are O(1) geometry, FP32 evaluation and the
 * FP32 over-point move them by ~1e-6, the reference's FP64 evaluation by 1e-15.
 *   (B) the origin is outside the ball and every direction points away from the centre ((p - o) . (o - c) > 0): the
 *       quadratic has b > 0, c > 0 -- no positive root.  Every shadow ray of a hit ON the ball.
 *   (C) every corner direction of the box lies inside the ball's tangent cone (the set of hitting directions is a convex
 *       cone, the box's directions are in the conical hull of its corners) and the far side of the ball is nearer than
 *       the nearest light point: every ray has a positive crossing nearer than the light.
 *   (A) a circular cone around the box's axis that contains every corner direction (hence every direction: (u . w) / |w|
 *       is quasi-concave where positive) lies outside the tangent cone: angle(axis, centre) > cone angle + tangent angle.
 */
template <typename T>
__device__ __forceinline__ int
shaft_sphere(const ShaftT<T> &sh, const float4 sp, bool casts)
{
    const float r = sp.w;
    const float mx = sp.x - (float)sh.o[0], my = sp.y - (float)sh.o[1], mz = sp.z - (float)sh.o[2];
    const float L2 = fmaf(mx, mx, fmaf(my, my, mz * mz)), r2 = r * r;
    const float wm_max = fmaxf(sh.dl[0] * mx, sh.dh[0] * mx) + fmaxf(sh.dl[1] * my, sh.dh[1] * my) + fmaxf(sh.dl[2] * mz, sh.dh[2] * mz);
    float w2max = 0.f, w2min = 0.f;
    for (int k = 0; k < 3; ++k) {
        const float a = fabsf(sh.dl[k]), b = fabsf(sh.dh[k]);
        w2max = fmaf(fmaxf(a, b), fmaxf(a, b), w2max);
        const float lo = (sh.dl[k] > 0.f || sh.dh[k] < 0.f) ? fminf(a, b) : 0.f;
        w2min = fmaf(lo, lo, w2min);
    }
    if (L2 > r2 * (1.0f + 8e-6f) && wm_max < -2e-5f * sqrtf(w2max * L2)) {
        return 0; /* (B) */
    }
    if (!(L2 > 1.1025f * r2)) {
        return 3; /* origin within 5 % of the surface: the tangent angle is ill-conditioned */
    }
    const float invL = rsqrtf(L2);
    const float sina = r * invL, cosa = sqrtf(fmaxf(1.0f - sina * sina, 0.f));
    float ux = 0.5f * (sh.dl[0] + sh.dh[0]), uy = 0.5f * (sh.dl[1] + sh.dh[1]), uz = 0.5f * (sh.dl[2] + sh.dh[2]);
    const float u2 = fmaf(ux, ux, fmaf(uy, uy, uz * uz));
    if (!(u2 > 1e-20f)) {
        return 3;
    }
    const float iu = rsqrtf(u2);
    ux *= iu;
    uy *= iu;
    uz *= iu;
    float cmin_u = 1.0f, cmin_m = 1.0f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const float wx = (c & 1) ? sh.dh[0] : sh.dl[0], wy = (c & 2) ? sh.dh[1] : sh.dl[1], wz = (c & 4) ? sh.dh[2] : sh.dl[2];
        const float iw = rsqrtf(fmaxf(fmaf(wx, wx, fmaf(wy, wy, wz * wz)), 1e-30f));
        cmin_u = fminf(cmin_u, fmaf(ux, wx, fmaf(uy, wy, uz * wz)) * iw);
        cmin_m = fminf(cmin_m, fmaf(mx, wx, fmaf(my, wy, mz * wz)) * iw * invL);
    }
    if (cmin_m > cosa + 2e-5f && (sqrtf(L2) + r) * 1.00002f < sqrtf(w2min)) {
        return casts ? 2 : 1; /* (C) */
    }
    if (!(cmin_u > 0.1f)) {
        return 3;
    }
    const float sint = sqrtf(fmaxf(1.0f - cmin_u * cmin_u, 0.f));
    const float cphi = fmaf(ux, mx, fmaf(uy, my, uz * mz)) * invL;
    if (cphi < fmaf(cmin_u, cosa, -sint * sina) - 2e-5f) {
        return 0; /* (A) */
    }
    return 3;
}

/*
 * One node of the shaft walk.  Returns 0 = no ray of the shaft ends its search here (go on at *next), 1 / 2 = every ray
 * ends it here, lit / shadowed, 3 = cannot tell (a leaf or a CSG node; *next is the node after it).
 */
template <typename T>
__device__ __forceinline__ int
shaft_node(const DSceneF &SF, const float4 *fnodes, unsigned int relevant, const ShaftT<T> &sh, int i, int *next)
{
    const T D_lo = (T)1 - ShaftEps<T>::dist(), D_hi = (T)1 + ShaftEps<T>::dist();
    const float4 q0 = __ldg(fnodes + 3 * i);
    const int flags = __float_as_int(q0.x), skip = __float_as_int(q0.y);
    if (!((relevant >> i) & 1u)) {
        *next = skip;
        return 0;
    }
    const float4 lo = __ldg(fnodes + 3 * i + 1), hi = __ldg(fnodes + 3 * i + 2);
    const int type = flags & FRT_FN_TYPE_MASK;
    SpanT<T> s;
    if (type >= FRT_CSG) {
        *next = skip;
        if (!(flags & FRT_FN_NOCULL) && (flags & FRT_FN_WORLD) && __float_as_int(q0.z) == 0) {
            T tn_lo, tn_hi, tf_lo, tf_hi;
            shaft_box_d(sh, lo, hi, tn_lo, tn_hi, tf_lo, tf_hi);
            if (tn_lo > tf_hi || tf_hi < (T)0) {
                return 0;
            }
        }
        if (type == FRT_GROUP) {
            *next = i + 1;
            return 0;
        }
        if (!(flags & FRT_FN_FAST)) {
            return 3;
        }
        int pc = __float_as_int(lo.w);
        const int pc1 = pc + __float_as_int(hi.w);
        s.flags = 0;
        s.a_lo = s.a_hi = s.b_lo = s.b_hi = (T)0;
        for (bool first = true; pc < pc1; first = false) {
            const int code = __ldg(SF.csg_prog + pc);
            SpanT<T> t;
            t.flags = 0;
            t.a_lo = t.a_hi = t.b_lo = t.b_hi = (T)0;
            /* an operand outside the shaft has no crossing at t > 0: for the crossings at t > 0 it is absent */
            if (((relevant >> code) & 1u) &&
                !shaft_leaf_span(sh, __ldg(fnodes + 3 * code), __ldg(fnodes + 3 * code + 1), __ldg(fnodes + 3 * code + 2), t)) {
                return 3;
            }
            if (first) {
                s = t;
                pc += 1;
            } else {
                const int op = -__ldg(SF.csg_prog + pc + 1) - 1;
                SpanT<T> r;
                r.a_lo = r.a_hi = r.b_lo = r.b_hi = (T)0;
                if (!csg_combine(op, s, t, r)) {
                    return 3;
                }
                s = r;
                pc += 2;
            }
        }
    } else {
        *next = i + 1;
        if (type == FRT_SPHERE) {
            const float4 sp = __ldg(SF.wsphere + i);
            return sp.w > 0.f ? shaft_sphere(sh, sp, (flags & FRT_FN_CASTS) != 0) : 3;
        }
        if (!shaft_leaf_span(sh, q0, lo, hi, s)) {
            return 3;
        }
    }
    if (s.flags) {
        const int v = judge_span(s, D_lo, D_hi);
        if (v == 3) {
            return 3;
        }
        if (v != 0) {
            return v == 2 ? 2 : 1;
        }
    }
    return 0;
}

/*
 * FRT_SH_LIT / FRT_SH_SHADOWED: the verdict of every shadow ray of the hit; FRT_SH_UNDECIDED: trace them one by one.
 * Trees of more than 32 nodes are not tried (`relevant` covers nodes 0..31).
 *
 * When the walk gets stuck at a node X it does not give up: X goes into the entry's program and the walk goes on AS IF X
 * ended no search.  When the rest is decided for the whole shaft -- a later node ends every search with one verdict, or
 * the tree ends (lit) -- that verdict is the answer of every ray the program's nodes do not stop: *prog (layout above
 * trace_entry_program).  A ray of a Cornell penumbra hit then evaluates the window wall's CSG program and nothing else.
 * Stuck more than FRT_PROG_MAX times: tail 0, the rays walk the tree from X1 on.
 */
#define FRT_RESUME_NODE_MASK 31
template <typename T>
__device__ __forceinline__ int
trace_shadow_bulk(const DSceneF &SF, int root, unsigned int relevant, const ShaftT<T> &sh, unsigned int *prog)
{
    const float4 *fnodes = SF.fnodes;
    int i = root, n_stuck = 0;
    unsigned int list = 0u;
    const int end = __float_as_int(__ldg(fnodes + 3 * i).y);
    while (i < end) {
        int next = i + 1;
        const int code = shaft_node(SF, fnodes, relevant, sh, i, &next);
        if (code == 3) {
            if (n_stuck == FRT_PROG_MAX) {
                *prog = list | ((unsigned int)n_stuck << 2);
                return FRT_SH_UNDECIDED;
            }
            list |= (unsigned int)i << (4 + 5 * n_stuck);
            ++n_stuck;
        } else if (code != 0) {
            if (n_stuck == 0) {
                return code == 2 ? FRT_SH_SHADOWED : FRT_SH_LIT;
            }
            *prog = list | ((unsigned int)n_stuck << 2) | (unsigned int)code;
            return FRT_SH_UNDECIDED;
        }
        i = next;
    }
    if (n_stuck == 0) {
        return FRT_SH_LIT;
    }
    *prog = list | ((unsigned int)n_stuck << 2) | 1u;
    return FRT_SH_UNDECIDED;
}

/*
 * The walks of the mesh path keep ONE current frame -- the world ray, or the ray in the frame of the transform a node names
 * -- and replace it when a node names another one.  (Choosing between a world and a local frame per node made the compiler
 * keep both in local memory behind a pointer: 14 % of k_shadow_mesh's instructions were local loads.)
 */
__device__ __forceinline__ void
current_frame(FrameF &cf, int &cur_xf, const DSceneF &SF, int xf, const FrameF &w, float omax, float eo, float ed_w)
{
    if (xf != cur_xf) {
        cur_xf = xf;
        if (xf == 0) {
            cf = w;
        } else {
            frame_local(cf, SF, xf, w, omax, eo, ed_w);
        }
    }
}

/*
 * A triangle leaf's mirror record holds the bounds of its vertices (FRT_FN_LEAFBOX): true when the ray surely misses
 * them, surely has them behind its origin (`behind`: no crossing at t <= 0 is wanted), or surely enters them beyond
 * `t_best`.  The reference tests no box in front of a triangle (triangle.c:11-45); a ray that misses the box of the
 * vertices by more than the box's padding gets no crossing from its Moeller-Trumbore test either.
 */
__device__ __forceinline__ bool
leaf_box_missed(const DSceneF &SF, const float4 q0, const float4 lo, const float4 hi, const FrameF &w, FrameF &lf, int &cur_xf_f,
                float omax, float eo, float ed_w, bool behind, double t_best)
{
    const int xf = __float_as_int(q0.z);
    float tn_lo, tn_hi, tf_lo, tf_hi;
    current_frame(lf, cur_xf_f, SF, xf, w, omax, eo, ed_w);
    box_f(lf, lo, hi, tn_lo, tn_hi, tf_lo, tf_hi);
    return tn_lo > tf_hi || (behind && tf_hi < 0.0f) || (double)tn_lo > t_best;
}

/*
 * The FP64 traversal (trace_shadow, frt_device.cuh) with its CULLS taken from the FP32 mirror: the walk, the group /
 * CSG box tests and the behind-the-origin test use the conservative FP32 slabs of this file (one 48-byte node record,
 * three loads issued together, instead of a header load followed by a dependent bounding-box load), every LEAF is still
 * intersected in FP64 on the exact ray and every verdict is taken in FP64.  Descending where the reference culls
 * changes nothing (the children lie inside the box), so the answer is trace_shadow's.  This is what re-traces the
 * rays the filter deferred -- on mesh scenes (triangles have no fast form) that is every shadow ray, and the
 * reference's divided tree makes them long: the 6-dragon scene visits 357 nodes per shadow ray, the longest rays tens
 * of thousands, and the kernel's duration is the latency of the longest ray.
 */
template <bool COUNT>
__device__ __forceinline__ bool
trace_shadow_mixed(const DScene &S, const DSceneF &SF, const Ray &wr, double distance, const FrameF &w, float omax, float eo_w,
                   float ed_w, int *overflow, unsigned long long *nodes_visited, unsigned long long *flops)
{
    struct Frame {
        int right, skip, start, mid, op;
    };
    CsgHit buf[FRT_CSG_CAP];
    Frame st[FRT_CSG_DEPTH];
    int sp = 0, n = 0;
    unsigned int visited = 0, cost = 0;
    bool result = false;
    const float4 *fnodes = SF.fnodes;
    int i = __ldg(S.roots);
    const int end = __float_as_int(__ldg(fnodes + 3 * i).y);
    int cur_xf_f = 0, cur_xf_d = 0;
    FrameF lf = w;
    Ray lr = wr;
    InvDir inv = inv_dir(wr);

    while (i < end) {
        const float4 q0 = __ldg(fnodes + 3 * i), lo = __ldg(fnodes + 3 * i + 1), hi = __ldg(fnodes + 3 * i + 2);
        const int flags = __float_as_int(q0.x), skip = __float_as_int(q0.y);
        const int type = flags & FRT_FN_TYPE_MASK;
        if (COUNT) {
            ++visited;
            cost += (type >= FRT_CSG) ? FRT_COST_BBOX : prim_cost(type);
        }
        if (type >= FRT_CSG) {
            bool miss = false;
            if (!(flags & FRT_FN_NOCULL)) {
                const int xf = __float_as_int(q0.z);
                float tn_lo, tn_hi, tf_lo, tf_hi;
                current_frame(lf, cur_xf_f, SF, xf, w, omax, eo_w, ed_w);
                box_f(lf, lo, hi, tn_lo, tn_hi, tf_lo, tf_hi);
                miss = tn_lo > tf_hi || (sp == 0 && tf_hi < 0.0f);
            }
            if (miss) {
                i = skip;
            } else {
                if (type == FRT_CSG) {
                    if (sp == FRT_CSG_DEPTH) {
                        *overflow = 1;
                        return false;
                    }
                    st[sp++] = Frame{ __float_as_int(q0.w), skip, n, -1, (flags >> FRT_FN_OP_SHIFT) & 3 };
                }
                i = i + 1;
            }
        } else if ((flags & FRT_FN_LEAFBOX) && leaf_box_missed(SF, q0, lo, hi, w, lf, cur_xf_f, omax, eo_w, ed_w, sp == 0, CUDART_INF)) {
            i = i + 1; /* the ray misses the bounds of the triangle's vertices */
        } else {
            const NodeA a = load_node_a(S, i);
            const NodeB b = load_node_b(S, i);
            if (a.xform != cur_xf_d) {
                cur_xf_d = a.xform;
                if (cur_xf_d == 0) {
                    lr = wr;
                } else {
                    lr = ray_to_local(S, cur_xf_d, wr);
                    if (COUNT) cost += FRT_COST_XFORM;
                }
                inv = inv_dir(lr);
            }
            double t[4], uv[2];
            const int k = prim_intersect_inv(a.type, S.params + (b.param < 0 ? 0 : b.param), lr, inv, t, uv);
            if (sp == 0) {
                bool stop = false;
                double tmin = CUDART_INF;
                for (int j = 0; j < k; ++j) {
                    stop = stop || !(t[j] <= 0);
                    if (t[j] > 0 && t[j] < tmin) {
                        tmin = t[j];
                    }
                }
                if (stop) {
                    result = S.mats[a.material].casts_shadow && tmin < distance;
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
        bool done = false;
        while (sp > 0) { /* close every CSG whose left / right operand just ended (csg.c:104-118, :43-71) */
            Frame &f = st[sp - 1];
            if (f.mid < 0 && i >= f.right) {
                f.mid = n;
            }
            if (i < f.skip) {
                break;
            }
            if (f.mid - f.start > 0 && n - f.mid > 0) {
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
            if (sp == 0) {
                bool stop = false;
                double tmin = CUDART_INF;
                for (int x = 0; x < n; ++x) {
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
    if (COUNT) {
        *nodes_visited += visited;
        *flops += cost;
    }
    return result;
}

/*
 * trace_closest_t (frt_device.cuh) with its CULLS taken from the FP32 mirror, like trace_shadow_mixed: one 48-byte
 * record per node instead of a header plus a dependent FP64 box, conservative FP32 slabs (a box is skipped only when it
 * is surely missed, surely behind the origin, or surely beyond the best hit so far), every leaf intersected in FP64 on
 * the exact ray, and the leaf's parameter offset / transform / casts-shadow bit read from its mirror record.  The
 * minimum over the leaves does not depend on which empty subtrees are skipped.
 */
template <bool CASTERS, int PRIMS = FRT_PRIMS_ALL>
__device__ __forceinline__ Hit
trace_closest_mixed(const DScene &S, const DSceneF &SF, const Ray &wr, int *overflow)
{
    Hit best;
    best.t = CUDART_INF;
    best.u = best.v = -1.0;
    best.leaf = -1;
    CsgHit buf[FRT_CSG_CAP];
    const float4 *fnodes = SF.fnodes;
    FrameF w;
    w.ox = (float)wr.ox;
    w.oy = (float)wr.oy;
    w.oz = (float)wr.oz;
    w.dx = (float)wr.dx;
    w.dy = (float)wr.dy;
    w.dz = (float)wr.dz;
    const float omax = fmaxf(fmaxf(fabsf(w.ox), fabsf(w.oy)), fabsf(w.oz));
    const float dmax = fmaxf(1.0f, fmaxf(fmaxf(fabsf(w.dx), fabsf(w.dy)), fabsf(w.dz))); /* refracted rays are not renormalised */
    const float eo_o = 2.0f * FRT_F32_U * omax;
    const float eo_w = fmaf(2.0f * FRT_F32_U, SF.bmax, fmaf(SF.ealign, omax, eo_o));
    const float ed_w = (FRT_F32_G + SF.ealign) * dmax;
    frame_finish(w, eo_w, eo_w, eo_w, ed_w, ed_w, ed_w);
    const InvDir winv = inv_dir(wr);
    for (int rt = 0; rt < S.n_roots; ++rt) {
        int i = __ldg(S.roots + rt);
        const int end = __float_as_int(__ldg(fnodes + 3 * i).y);
        int cur_xf_f = 0, cur_xf_d = 0;
        FrameF lf = w;
        Ray lr = wr;
        InvDir inv = winv;
        while (i < end) {
            const float4 q0 = __ldg(fnodes + 3 * i), lo = __ldg(fnodes + 3 * i + 1);
            const int flags = __float_as_int(q0.x), skip = __float_as_int(q0.y);
            const int type = flags & FRT_FN_TYPE_MASK;
            if (type >= FRT_CSG) {
                const float4 hi = __ldg(fnodes + 3 * i + 2);
                const int xf = __float_as_int(q0.z);
                float tn_lo, tn_hi, tf_lo, tf_hi;
                current_frame(lf, cur_xf_f, SF, xf, w, omax, eo_o, ed_w);
                box_f(lf, lo, hi, tn_lo, tn_hi, tf_lo, tf_hi);
                const bool miss = tn_lo > tf_hi || tf_hi < 0.0f || (double)tn_lo > best.t;
                if (type == FRT_GROUP) {
                    i = miss ? skip : i + 1;
                } else {
                    if (!miss) {
                        const int n = csg_eval<PRIMS>(S, i, wr, buf, overflow);
                        for (int k = 0; k < n; ++k) {
                            if (buf[k].t > 0 && buf[k].t < best.t &&
                                (!CASTERS || (__float_as_int(__ldg(fnodes + 3 * buf[k].leaf).x) & FRT_FN_CASTS))) {
                                best.t = buf[k].t;
                                best.leaf = buf[k].leaf;
                                best.u = best.v = -1.0;
                            }
                        }
                    }
                    i = skip;
                }
            } else if ((flags & FRT_FN_LEAFBOX) &&
                       leaf_box_missed(SF, q0, lo, __ldg(fnodes + 3 * i + 2), w, lf, cur_xf_f, omax, eo_o, ed_w, true, best.t)) {
                i = i + 1; /* the ray misses the bounds of the triangle's vertices, or meets them beyond the best hit */
            } else {
                const int xform = __float_as_int(lo.w), param = __float_as_int(q0.w);
                if (xform != cur_xf_d) {
                    cur_xf_d = xform;
                    if (xform == 0) {
                        lr = wr;
                        inv = winv;
                    } else {
                        lr = ray_to_local(S, xform, wr);
                        inv = inv_dir(lr);
                    }
                }
                if (!CASTERS || (flags & FRT_FN_CASTS)) { /* hit(xs, true) skips objects that do not cast shadows */
                    double t[4], uv[2];
                    uv[0] = uv[1] = -1.0;
                    const int k = prim_intersect<PRIMS>(type, S.params + (param < 0 ? 0 : param), lr, t, uv);
                    for (int j = 0; j < k; ++j) {
                        if (t[j] > 0 && t[j] < best.t) {
                            best.t = t[j];
                            best.leaf = i;
                            best.u = uv[0];
                            best.v = uv[1];
                        }
                    }
                }
                i = i + 1;
            }
        }
    }
    return best;
}
