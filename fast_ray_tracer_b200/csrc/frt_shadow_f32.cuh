/*
 * frt_shadow_f32.cuh -- FP32 filtered shadow-ray traversal.
 *
 * is_shadowed (renderer.c:73-93) is 98.7 % of the reference's rays on the Cornell scene and
 * its answer is one bit.  That bit is the outcome of a chain of comparisons (slab order, t against 0, t against the
 * light distance, order of CSG crossings).  This file walks the same tree in the same order as the FP64 traversal
 * (trace_shadow in frt_device.cuh) but in FP32, carrying a conservative absolute error bound next to every t value:
 *
 *   - a comparison whose operands are further apart than their bounds has the same outcome as in FP64: keep going;
 *   - a group / CSG bounding-box test is only a cull, so it is resolved conservatively (when in doubt, descend:
 *     the children are inside the box, so descending where the reference culls changes nothing);
 *   - any other comparison that the bounds cannot decide makes the whole ray UNDECIDED.  The caller appends the
 *     ray to a queue (warp-aggregated) and a second kernel re-traces the queue in FP64.  On the Cornell frame that
 *     is a few percent of the rays (silhouettes, grazing starts, exact CSG ties).
 *
 * The bit that comes out is therefore the FP64 traversal's bit; FRT_FLAG_VERIFY_F32 checks that claim ray by ray on
 * the device (every decided ray is also traced in FP64 and disagreements are counted; tests assert zero).
 *
 * Error model.  u = 2^-24 is the FP32 unit roundoff.  G = 2^-21 = 8u bounds the rounding of any expression of up
 * to four products / sums relative to the sum of the magnitudes of its terms.  The world ray is the FP64 ray
 * rounded to FP32 (origin: relative u per component; unit direction and distance: a few u after the FP32
 * normalisation, taken as G).  A transform M (rows m_k, translation T_k) maps these to local bounds
 *      eo_k = R_k * (u + G) * |o|max + G * |T_k|,     ed_k = R_k * 2G,          R_k = sum_j |m_kj|
 * and a slab value t = (b - o_k) / d_k with |d_k| >= 2 ed_k carries
 *      E_t = |1/d_k| * (eo_k + 2 |t| ed_k) + G |t|.
 */
#pragma once

#include "frt_device.cuh"

#define FRT_F32_G 4.76837158203125e-07f /* 2^-21 */
#define FRT_F32_U 5.9604644775390625e-08f /* 2^-24 */
#define FRT_EPS_F 0.00001f

enum { FRT_SH_LIT = 0, FRT_SH_SHADOWED = 1, FRT_SH_UNDECIDED = 2 };

struct DSceneF { /* FP32 mirror of the tree, built at upload */
    const float4 *fx;    /* 4 x float4 per xform: rows 0..2 of the world->local matrix, then {R_0, R_1, R_2, 0} */
    const float4 *fbbox; /* 2 x float4 per node: {min.xyz, 0} {max.xyz, 0}, rounded outward */
};

struct RayF {
    float ox, oy, oz, dx, dy, dz;
};

struct LocalF {
    RayF r;
    float ix, iy, iz;    /* 1 / d_k */
    float eox, eoy, eoz; /* bounds on the local origin */
    float edx, edy, edz; /* bounds on the local direction */
};

__device__ __forceinline__ float
rcpf_fast(float x)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r; /* <= 1 ulp: covered by G */
}

__device__ __forceinline__ void
local_setup(LocalF &L, const DSceneF &SF, int xf, const RayF &w, float omax)
{
    const float c_o = (FRT_F32_U + FRT_F32_G) * omax;
    if (xf == 0) {
        L.r = w;
        L.eox = L.eoy = L.eoz = c_o;
        L.edx = L.edy = L.edz = FRT_F32_G;
    } else {
        const float4 m0 = __ldg(SF.fx + 4 * xf), m1 = __ldg(SF.fx + 4 * xf + 1), m2 = __ldg(SF.fx + 4 * xf + 2);
        const float4 R = __ldg(SF.fx + 4 * xf + 3);
        L.r.ox = fmaf(m0.x, w.ox, fmaf(m0.y, w.oy, fmaf(m0.z, w.oz, m0.w)));
        L.r.oy = fmaf(m1.x, w.ox, fmaf(m1.y, w.oy, fmaf(m1.z, w.oz, m1.w)));
        L.r.oz = fmaf(m2.x, w.ox, fmaf(m2.y, w.oy, fmaf(m2.z, w.oz, m2.w)));
        L.r.dx = fmaf(m0.x, w.dx, fmaf(m0.y, w.dy, m0.z * w.dz));
        L.r.dy = fmaf(m1.x, w.dx, fmaf(m1.y, w.dy, m1.z * w.dz));
        L.r.dz = fmaf(m2.x, w.dx, fmaf(m2.y, w.dy, m2.z * w.dz));
        L.eox = fmaf(R.x, c_o, FRT_F32_G * fabsf(m0.w));
        L.eoy = fmaf(R.y, c_o, FRT_F32_G * fabsf(m1.w));
        L.eoz = fmaf(R.z, c_o, FRT_F32_G * fabsf(m2.w));
        L.edx = R.x * (2.0f * FRT_F32_G);
        L.edy = R.y * (2.0f * FRT_F32_G);
        L.edz = R.z * (2.0f * FRT_F32_G);
    }
    L.ix = rcpf_fast(L.r.dx);
    L.iy = rcpf_fast(L.r.dy);
    L.iz = rcpf_fast(L.r.dz);
}

/* one slab axis: interval [lo, hi] of t and the bound E on both ends; valid only when |d| >= EPS + 2 ed */
__device__ __forceinline__ void
slab_f(float o, float inv, float eo, float ed, float blo, float bhi, float &lo, float &hi, float &E)
{
    const float a = (blo - o) * inv;
    const float b = (bhi - o) * inv;
    lo = fminf(a, b);
    hi = fmaxf(a, b);
    const float tabs = fmaxf(fabsf(a), fabsf(b));
    E = fmaf(fabsf(inv), fmaf(2.0f * tabs, ed, eo), FRT_F32_G * tabs);
}

/* conservative bounding_box_intersects (bounding_box.c:165-175): false only when the FP64 test is surely false */
__device__ __forceinline__ bool
bbox_maybe_f(const DSceneF &SF, int node, const LocalF &L)
{
    const float4 bmin = __ldg(SF.fbbox + 2 * node), bmax = __ldg(SF.fbbox + 2 * node + 1);
    float tn = -CUDART_INF_F, tf = CUDART_INF_F;
    float lo, hi, E;
    if (fabsf(L.r.dx) >= FRT_EPS_F + 2.0f * L.edx) {
        slab_f(L.r.ox, L.ix, L.eox, L.edx, bmin.x, bmax.x, lo, hi, E);
        tn = fmaxf(tn, lo - E); /* a NaN (unbounded box side) drops out of fmaxf / fminf: the axis stays open */
        tf = fminf(tf, hi + E);
    }
    if (fabsf(L.r.dy) >= FRT_EPS_F + 2.0f * L.edy) {
        slab_f(L.r.oy, L.iy, L.eoy, L.edy, bmin.y, bmax.y, lo, hi, E);
        tn = fmaxf(tn, lo - E);
        tf = fminf(tf, hi + E);
    }
    if (fabsf(L.r.dz) >= FRT_EPS_F + 2.0f * L.edz) {
        slab_f(L.r.oz, L.iz, L.eoz, L.edz, bmin.z, bmax.z, lo, hi, E);
        tn = fmaxf(tn, lo - E);
        tf = fminf(tf, hi + E);
    }
    return !(tn > tf);
}

/*
 * Crossings of one leaf in FP32 as intervals [tlo_j, thi_j] that surely contain the FP64 values: returns the count
 * (like prim_intersect), -1 when the count itself is undecided.  Cube, sphere and plane are evaluated in FP32;
 * every other type is evaluated in FP64 on the exact ray (`wr`), so its values only carry the final rounding.
 */
__device__ __forceinline__ int
prim_f(const DScene &S, int type, int xf, int param, const LocalF &L, const Ray &wr, float tlo[4], float thi[4])
{
    if (type == FRT_CUBE) { /* cube_local_intersect, cube.c:56-78 */
        if (!(fabsf(L.r.dx) >= FRT_EPS_F + 2.0f * L.edx) || !(fabsf(L.r.dy) >= FRT_EPS_F + 2.0f * L.edy) ||
            !(fabsf(L.r.dz) >= FRT_EPS_F + 2.0f * L.edz)) {
            return -1; /* an axis-parallel ray takes the reference's INFINITY branch: leave it to FP64 */
        }
        float x0, x1, y0, y1, z0, z1, ex, ey, ez;
        slab_f(L.r.ox, L.ix, L.eox, L.edx, -1.0f, 1.0f, x0, x1, ex);
        slab_f(L.r.oy, L.iy, L.eoy, L.edy, -1.0f, 1.0f, y0, y1, ey);
        slab_f(L.r.oz, L.iz, L.eoz, L.edz, -1.0f, 1.0f, z0, z1, ez);
        /* tmin = max of the entries, tmax = min of the exits, as intervals */
        const float tn_lo = fmaxf(fmaxf(x0 - ex, y0 - ey), z0 - ez), tn_hi = fmaxf(fmaxf(x0 + ex, y0 + ey), z0 + ez);
        const float tf_lo = fminf(fminf(x1 - ex, y1 - ey), z1 - ez), tf_hi = fminf(fminf(x1 + ex, y1 + ey), z1 + ez);
        if (tn_lo > tf_hi) {
            return 0; /* tmin > tmax for sure */
        }
        if (!(tn_hi < tf_lo)) {
            return -1;
        }
        tlo[0] = tn_lo;
        thi[0] = tn_hi;
        tlo[1] = tf_lo;
        thi[1] = tf_hi;
        return 2;
    }
    if (type == FRT_PLANE) { /* plane_local_intersect, plane.c:11-25 */
        if (fabsf(L.r.dy) + L.edy < FRT_EPS_F) {
            return 0;
        }
        if (!(fabsf(L.r.dy) >= FRT_EPS_F + 2.0f * L.edy)) {
            return -1;
        }
        const float tt = -L.r.oy * L.iy;
        const float E = fmaf(fabsf(L.iy), fmaf(2.0f * fabsf(tt), L.edy, L.eoy), FRT_F32_G * fabsf(tt));
        tlo[0] = tt - E;
        thi[0] = tt + E;
        return 1;
    }
    if (type == FRT_SPHERE) { /* sphere_local_intersect, sphere.c:14-40 */
        const RayF &r = L.r;
        const float eo = fmaxf(fmaxf(L.eox, L.eoy), L.eoz), ed = fmaxf(fmaxf(L.edx, L.edy), L.edz);
        const float So = fabsf(r.ox) + fabsf(r.oy) + fabsf(r.oz), Sd = fabsf(r.dx) + fabsf(r.dy) + fabsf(r.dz);
        const float a = fmaf(r.dx, r.dx, fmaf(r.dy, r.dy, r.dz * r.dz));
        const float hb = fmaf(r.dx, r.ox, fmaf(r.dy, r.oy, r.dz * r.oz));
        const float habs = fmaf(fabsf(r.dx), fabsf(r.ox), fmaf(fabsf(r.dy), fabsf(r.oy), fabsf(r.dz * r.oz)));
        const float oo = fmaf(r.ox, r.ox, fmaf(r.oy, r.oy, r.oz * r.oz));
        const float b = 2.0f * hb, c = oo - 1.0f;
        const float da = fmaf(2.0f * Sd + 3.0f * ed, ed, FRT_F32_G * a);
        const float db = 2.0f * (fmaf(Sd, eo, fmaf(So, ed, 3.0f * eo * ed)) + FRT_F32_G * habs);
        const float dc = fmaf(2.0f * So + 3.0f * eo, eo, FRT_F32_G * (oo + 1.0f));
        const float disc = fmaf(b, b, -4.0f * a * c);
        const float dD = fmaf(2.0f * fabsf(b) + db, db, 4.0f * (fmaf(a, dc, fmaf(fabsf(c), da, da * dc)))) +
                         FRT_F32_G * fmaf(b, b, 4.0f * a * fabsf(c));
        if (disc + dD < 0.0f) {
            return 0;
        }
        if (!(disc > 4.0f * dD) || !(a > 4.0f * da)) {
            return -1;
        }
        const float s = sqrtf(disc);
        const float ds = dD / s + FRT_F32_G * s;
        const float i2a = rcpf_fast(2.0f * a);
        const float ra = 2.0f * (da / a) + 2.0f * FRT_F32_G;
        const float t0 = (-b - s) * i2a, t1 = (-b + s) * i2a;
        const float E0 = (db + ds) * 2.0f * i2a;
        const float e0 = fmaf(fabsf(t0), ra, E0), e1 = fmaf(fabsf(t1), ra, E0);
        tlo[0] = t0 - e0;
        thi[0] = t0 + e0;
        tlo[1] = t1 - e1;
        thi[1] = t1 + e1;
        return 2;
    }
    /* cylinder, cone, torus, triangles: FP64 on the exact ray (`wr` arrives with its direction not yet normalised) */
    Ray er = wr;
    {
        const double inv = rsqrt_fast(er.dx * er.dx + er.dy * er.dy + er.dz * er.dz);
        er.dx *= inv;
        er.dy *= inv;
        er.dz *= inv;
    }
    const Ray lr = ray_to_local(S, xf, er);
    double td[4], uv[2];
    const int k = prim_intersect(type, S.params + (param < 0 ? 0 : param), lr, td, uv);
    for (int j = 0; j < k; ++j) {
        const float tt = (float)td[j];
        const float E = 2.0f * FRT_F32_U * fabsf(tt) + 1e-37f;
        tlo[j] = tt - E;
        thi[j] = tt + E;
    }
    return k;
}

struct CsgHitF {
    float lo, hi; /* the crossing's t lies in [lo, hi] */
    int leaf;
};

/*
 * FP32 twin of trace_shadow (frt_device.cuh).  `wr` is the exact FP64 ray with an unnormalised direction (only read
 * for the rare FP64 leaf types), `w` / `Df` its normalised FP32 image; returns FRT_SH_LIT, FRT_SH_SHADOWED or FRT_SH_UNDECIDED.
 */
template <bool COUNT>
__device__ __forceinline__ int
trace_shadow_f32(const DScene &S, const DSceneF &SF, const Ray &wr, const RayF &w, float Df, int *overflow,
                 unsigned long long *nodes_visited, unsigned long long *flops)
{
    struct Frame {
        int right, skip, start, mid, op;
    };
    CsgHitF buf[FRT_CSG_CAP];
    Frame st[FRT_CSG_DEPTH];
    int sp = 0, n = 0;
    unsigned int visited = 0, cost = 0;
    const float omax = fmaxf(fmaxf(fabsf(w.ox), fabsf(w.oy)), fabsf(w.oz));
    const float D_lo = Df - FRT_F32_G * Df, D_hi = Df + FRT_F32_G * Df; /* the light distance lies in [D_lo, D_hi] */
    int result = FRT_SH_LIT;

    for (int rt = 0; rt < S.n_roots; ++rt) {
        int i = __ldg(S.roots + rt);
        const int end = load_node_a(S, i).skip;
        int cur_xf = 0;
        LocalF L;
        local_setup(L, SF, 0, w, omax);
        bool any = false, done = false;
        while (i < end) {
            const NodeA a = load_node_a(S, i);
            if (COUNT) ++visited;
            if (a.xform != cur_xf) {
                cur_xf = a.xform;
                local_setup(L, SF, cur_xf, w, omax);
                if (COUNT && cur_xf != 0) cost += FRT_COST_XFORM;
            }
            if (COUNT) cost += (a.type >= FRT_CSG) ? FRT_COST_BBOX : prim_cost(a.type);
            if (a.type >= FRT_CSG) {
                if (!bbox_maybe_f(SF, i, L)) {
                    i = a.skip;
                } else {
                    if (a.type == FRT_CSG) {
                        if (sp == FRT_CSG_DEPTH) {
                            *overflow = 1;
                            return FRT_SH_LIT;
                        }
                        const NodeB b = load_node_b(S, i);
                        st[sp++] = Frame{ b.right, a.skip, n, -1, b.csg_op };
                    }
                    i = i + 1;
                }
            } else {
                float tlo[4], thi[4];
                const int k = prim_f(S, a.type, a.xform, load_node_b(S, i).param, L, wr, tlo, thi);
                if (k < 0) {
                    return FRT_SH_UNDECIDED;
                }
                if (sp == 0) {
                    if (k > 0) {
                        any = true;
                        /* the search stops here iff some t is not <= 0 (group.c:105-123); shadowed iff the leaf casts
                         * shadows and its smallest t > 0 is nearer than the light (renderer.c:87-90) */
                        bool stop = false, amb = false, near_sure = false, far_amb = false;
                        for (int j = 0; j < k; ++j) {
                            if (tlo[j] > 0.0f) {
                                stop = true;
                                if (thi[j] < D_lo) {
                                    near_sure = true;
                                } else if (!(tlo[j] >= D_hi)) {
                                    far_amb = true;
                                }
                            } else if (!(thi[j] <= 0.0f)) {
                                amb = true; /* sign not decided (or NaN) */
                            }
                        }
                        if (stop) {
                            if (!S.mats[a.material].casts_shadow) {
                                result = FRT_SH_LIT;
                            } else if (near_sure) {
                                result = FRT_SH_SHADOWED;
                            } else if (!amb && !far_amb) {
                                result = FRT_SH_LIT;
                            } else {
                                result = FRT_SH_UNDECIDED;
                            }
                            done = true;
                        } else if (amb) {
                            return FRT_SH_UNDECIDED;
                        }
                    }
                } else {
                    for (int j = 0; j < k; ++j) {
                        if (n == FRT_CSG_CAP) {
                            *overflow = 1;
                            return FRT_SH_LIT;
                        }
                        buf[n].lo = tlo[j];
                        buf[n].hi = thi[j];
                        buf[n].leaf = i;
                        ++n;
                    }
                }
                i = i + 1;
            }
            /* close every CSG whose left / right operand just ended (csg.c:104-118, :43-71) */
            while (sp > 0) {
                Frame &f = st[sp - 1];
                if (f.mid < 0 && i >= f.right) {
                    f.mid = n;
                }
                if (i < f.skip) {
                    break;
                }
                if (f.mid - f.start > 0 && n - f.mid > 0) {
                    for (int x = f.start + 1; x < n; ++x) {
                        const CsgHitF h = buf[x];
                        int y = x - 1;
                        while (y >= f.start) {
                            if (buf[y].hi < h.lo) {
                                break; /* surely in order */
                            }
                            if (!(buf[y].lo > h.hi)) {
                                return FRT_SH_UNDECIDED; /* order of two crossings not decided in FP32 */
                            }
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
                if (sp == 0) { /* the outermost CSG is judged like a leaf, with a material per crossing */
                    bool stop = false, amb = false, near_sure = false, far_amb = false;
                    for (int x = 0; x < n; ++x) {
                        any = true;
                        if (buf[x].lo > 0.0f) {
                            stop = true;
                            if (S.mats[load_node_a(S, buf[x].leaf).material].casts_shadow) {
                                if (buf[x].hi < D_lo) {
                                    near_sure = true;
                                } else if (!(buf[x].lo >= D_hi)) {
                                    far_amb = true;
                                }
                            }
                        } else if (!(buf[x].hi <= 0.0f)) {
                            amb = true;
                        }
                    }
                    n = 0;
                    if (stop) {
                        if (near_sure) {
                            result = FRT_SH_SHADOWED;
                        } else if (!amb && !far_amb) {
                            result = FRT_SH_LIT;
                        } else {
                            result = FRT_SH_UNDECIDED;
                        }
                        done = true;
                    } else if (amb) {
                        return FRT_SH_UNDECIDED;
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
