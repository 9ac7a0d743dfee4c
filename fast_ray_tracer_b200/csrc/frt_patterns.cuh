/*
 * frt_patterns.cuh -- device evaluators for src/pattern/pattern.c, src/libs/perlin/perlin.c and the texel
 * fetch of src/libs/canvas/canvas.c:115-148.  Textures stay resident in HBM and are read through the
 * read-only (__ldg) path.
 */
#pragma once

#include "frt_device.cuh"

/* ---- perlin.c: integer-hash value noise with cosine interpolation (signed wrap-around made explicit) ---- */

/* rawnoise, perlin.c:9-12: n = (n << 13) ^ n; 1 - ((n*(n*n*15731 + 789221) + 1376312589) & 0x7fffffff) / 1073741824 */
__device__ __forceinline__ double
perlin_raw(int n)
{
    unsigned int u = (unsigned int)n;
    u = (u << 13) ^ u;
    unsigned int v = u * (u * u * 15731u + 789221u) + 1376312589u;
    return 1.0 - (double)(v & 0x7fffffffu) / 1073741824.0;
}

/* noise3d, perlin.c:22-24 */
__device__ __forceinline__ double
perlin_noise3d(int x, int y, int z, int octave, int seed)
{
    return perlin_raw(x * 1919 + y * 31337 + z * 7669 + octave * 3463 + seed * 13397);
}

/* interpolate, perlin.c:26-30 (cosine) */
__device__ __forceinline__ double
perlin_lerp(double a, double b, double x)
{
    double f = (1.0 - cos(x * M_PI)) * 0.5;
    return a * (1.0 - f) + b * f;
}

/* smooth3d, perlin.c:59-87: the lattice cell is (int)|x|, the fraction x - cell (negative for x < 0) */
__device__ __noinline__ double
perlin_smooth3d(double x, double y, double z, int octave, int seed)
{
    int ix = (int)(x < 0 ? -x : x), iy = (int)(y < 0 ? -y : y), iz = (int)(z < 0 ? -z : z);
    double fx = x - ix, fy = y - iy, fz = z - iz;
    double v1 = perlin_noise3d(ix, iy, iz, octave, seed);
    double v2 = perlin_noise3d(ix + 1, iy, iz, octave, seed);
    double v3 = perlin_noise3d(ix, iy + 1, iz, octave, seed);
    double v4 = perlin_noise3d(ix + 1, iy + 1, iz, octave, seed);
    double v5 = perlin_noise3d(ix, iy, iz + 1, octave, seed);
    double v6 = perlin_noise3d(ix + 1, iy, iz + 1, octave, seed);
    double v7 = perlin_noise3d(ix, iy + 1, iz + 1, octave, seed);
    double v8 = perlin_noise3d(ix + 1, iy + 1, iz + 1, octave, seed);
    double i1 = perlin_lerp(v1, v2, fx);
    double i2 = perlin_lerp(v3, v4, fx);
    double i3 = perlin_lerp(v5, v6, fx);
    double i4 = perlin_lerp(v7, v8, fx);
    double j1 = perlin_lerp(i1, i2, fy);
    double j2 = perlin_lerp(i3, i4, fy);
    return perlin_lerp(j1, j2, fz);
}

/* pnoise3d, perlin.c:118-131 */
__device__ __noinline__ double
perlin_pnoise3d(double x, double y, double z, double persistence, double frequency, int octaves, int seed)
{
    double total = 0.0;
    double amplitude = 1.0;
    for (int i = 0; i < octaves; ++i) {
        total += perlin_smooth3d(x * frequency, y * frequency, z * frequency, i, seed) * amplitude;
        frequency /= 2.0;
        amplitude *= persistence;
    }
    return total;
}

/* ---- textures ---------------------------------------------------------------------------------------- */

/* srgb_to_rgb, src/color/srgb.c:15-24 */
__device__ __forceinline__ double
srgb_to_linear(double c)
{
    return c <= 0.04045 ? c / 12.92 : pow((c + 0.055) / 1.055, 2.4);
}

/*
 * Texture ingest.  The reference keeps an image as a Canvas of FP64 RGB (32 B per texel on the host)
 * and evaluates canvas_pixel_at (canvas.c:115-148) at EVERY fetch: a 3 x 3 wrap-around box average when the canvas is
 * super-sampled, then color_space_fn -- two pow() per channel for sRGB sources.  Both depend on the texel only, so the
 * upload evaluates them once per texel, in FP64 and in the reference's order of operations, and stores linear FP32 RGBA
 * (16 B per texel, one 128-bit load through the read-only path per fetch); a fetch returns exactly what canvas_pixel_at
 * returns, rounded to FP32 (6e-8 relative, four orders below one sRGB-8 LSB; bump maps perturb the normal by the same).
 */
__global__ void
k_texture_ingest(const double *__restrict__ raw, int width, int height, int super_sample, int color_fn, float4 *__restrict__ out)
{
    const long n = (long)width * height;
    for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long)gridDim.x * blockDim.x) {
        const long row = t / width, col = t - row * width;
        double c[3] = { 0.0, 0.0, 0.0 };
        if (super_sample) {
            for (int j = -1; j <= 1; ++j) { /* columns outside, rows inside: the reference's order of accumulation */
                const long cc = (col + j + width) % width;
                for (int i = -1; i <= 1; ++i) {
                    const long rr = (row + i + height) % height;
                    const double *p = raw + 3 * (rr * width + cc);
                    c[0] += p[0];
                    c[1] += p[1];
                    c[2] += p[2];
                }
            }
            c[0] *= 1.0 / 9.0;
            c[1] *= 1.0 / 9.0;
            c[2] *= 1.0 / 9.0;
        } else {
            const double *p = raw + 3 * t;
            c[0] = p[0];
            c[1] = p[1];
            c[2] = p[2];
        }
        if (color_fn == FRT_COLOR_SRGB_TO_RGB) {
            c[0] = srgb_to_linear(c[0]);
            c[1] = srgb_to_linear(c[1]);
            c[2] = srgb_to_linear(c[2]);
        }
        out[t] = make_float4((float)c[0], (float)c[1], (float)c[2], 0.f);
    }
}

/* canvas_pixel_at (canvas.c:115-148) of the ingested texture: one texel, already filtered and in linear RGB */
__device__ __forceinline__ void
texture_fetch(const DScene &S, int tex, long col, long row, double out[3])
{
    const frt_texture T = S.texs[tex];
    if (col < 0) col = 0;
    if (row < 0) row = 0;
    if (col >= T.width) col = T.width - 1;
    if (row >= T.height) row = T.height - 1;
    const float4 t = __ldg(S.texels + T.texel_offset + row * T.width + col);
    out[0] = (double)t.x;
    out[1] = (double)t.y;
    out[2] = (double)t.z;
}

/* ---- uv maps, pattern.c:310-488 ---------------------------------------------------------------------- */

__device__ __noinline__ void
uv_map_eval(int map_type, int leaf_type, const double *prm, const double pt[3], int &face, double &u, double &v)
{
    face = 0;
    switch (map_type) {
    case FRT_UV_CUBE: { /* cube_uv_map, pattern.c:310-358 */
        double coord = fmax(fmax(fabs(pt[0]), fabs(pt[1])), fabs(pt[2]));
        face = fabs(coord - pt[0]) < FRT_EPS ? 0
             : fabs(coord + pt[0]) < FRT_EPS ? 1
             : fabs(coord - pt[1]) < FRT_EPS ? 2
             : fabs(coord + pt[1]) < FRT_EPS ? 3
             : fabs(coord - pt[2]) < FRT_EPS ? 4 : 5;
        switch (face) {
        case 0:
            u = fmod(1.0 - pt[2], 2.0) / 2.0;
            v = fmod(pt[1] + 1.0, 2.0) / 2.0;
            break;
        case 1:
            u = fmod(pt[2] + 1.0, 2.0) / 2.0;
            v = fmod(pt[1] + 1.0, 2.0) / 2.0;
            break;
        case 2:
            u = fmod(pt[0] + 1.0, 2.0) / 2.0;
            v = fmod(1.0 - pt[2], 2.0) / 2.0;
            break;
        case 3:
            u = fmod(pt[0] + 1.0, 2.0) / 2.0;
            v = fmod(pt[2] + 1.0, 2.0) / 2.0;
            break;
        case 4:
            u = fmod(pt[0] + 1.0, 2.0) / 2.0;
            v = fmod(pt[1] + 1.0, 2.0) / 2.0;
            break;
        default:
            u = fmod(1.0 - pt[0], 2.0) / 2.0;
            v = fmod(pt[1] + 1.0, 2.0) / 2.0;
            break;
        }
        break;
    }
    case FRT_UV_CYLINDER: { /* cylinder_uv_map, pattern.c:360-391; reads the shape's cylinder min/max */
        double mn = 0.0, mx = 0.0;
        if (leaf_type == FRT_CYLINDER || leaf_type == FRT_CONE) {
            mn = __ldg(prm + 0);
            mx = __ldg(prm + 1);
        }
        face = (mx - FRT_EPS) <= pt[1] ? 1 : (mn + FRT_EPS) >= pt[1] ? 2 : 0;
        if (face == 0) {
            double theta = atan2(pt[0], pt[2]);
            double raw_u = theta / (2.0 * M_PI);
            u = 1.0 - (raw_u + 0.5);
            v = fmod(pt[1], 1.0);
        } else if (face == 1) {
            u = fmod(pt[0] + 1.0, 2.0) / 2.0;
            v = fmod(1.0 - pt[2], 2.0) / 2.0;
        } else {
            u = fmod(pt[0] + 1.0, 2.0) / 2.0;
            v = fmod(pt[2] + 1.0, 2.0) / 2.0;
        }
        break;
    }
    case FRT_UV_PLANE: /* plane_uv_map, pattern.c:442-457 */
        u = fmod(pt[0], 1.0);
        v = fmod(pt[2], 1.0);
        if (u < 0) u += 1.0;
        if (v < 0) v += 1.0;
        break;
    case FRT_UV_SPHERE: { /* sphere_uv_map, pattern.c:459-475 */
        double theta = atan2(pt[0], pt[2]);
        double radius = sqrt(pt[0] * pt[0] + pt[1] * pt[1] + pt[2] * pt[2]);
        double phi = acos(pt[1] / radius);
        double raw_u = theta / (2 * M_PI);
        u = 1 - (raw_u + 0.5);
        v = 1 - phi / M_PI;
        break;
    }
    case FRT_UV_TOROID: { /* toroid_uv_map, pattern.c:477-488 */
        double r1 = leaf_type == FRT_TOROID ? __ldg(prm + 0) : 0.0;
        u = (1.0 - (atan2(pt[2], pt[0]) + M_PI) / (2 * M_PI));
        double len = sqrt(pt[0] * pt[0] + pt[2] * pt[2]);
        double x = len - r1;
        v = (atan2(pt[1], x) + M_PI) / (2 * M_PI);
        break;
    }
    default: { /* triangle_uv_map, pattern.c:393-440 (barycentrics recomputed from the point, fmod'ed) */
        double p1[3], e1[3], e2[3];
        for (int k = 0; k < 3; ++k) {
            p1[k] = __ldg(prm + k);
            e1[k] = __ldg(prm + 9 + k);
            e2[k] = __ldg(prm + 12 + k);
        }
        double v2x = pt[0] - p1[0], v2y = pt[1] - p1[1], v2z = pt[2] - p1[2];
        double d00 = e1[0] * e1[0] + e1[1] * e1[1] + e1[2] * e1[2];
        double d01 = e1[0] * e2[0] + e1[1] * e2[1] + e1[2] * e2[2];
        double d11 = e2[0] * e2[0] + e2[1] * e2[1] + e2[2] * e2[2];
        double d20 = v2x * e1[0] + v2y * e1[1] + v2z * e1[2];
        double d21 = v2x * e2[0] + v2y * e2[1] + v2z * e2[2];
        double denom = 1.0 / (d00 * d11 - d01 * d01);
        double bv = fmod((d11 * d20 - d01 * d21) * denom, 1.0);
        double bw = fmod((d00 * d21 - d01 * d20) * denom, 1.0);
        double bu = 1.0 - bv - bw;
        if (__ldg(prm + 33) != 0.0) {
            double s0 = __ldg(prm + 24) * bu + (__ldg(prm + 27) * bv + __ldg(prm + 30) * (1.0 - bu - bv));
            double s1 = __ldg(prm + 25) * bu + (__ldg(prm + 28) * bv + __ldg(prm + 31) * (1.0 - bu - bv));
            u = fmod(s0, 1.0);
            v = fmod(s1, 1.0);
        } else {
            u = bu;
            v = bv;
        }
        if (u < 0) u += 1.0;
        if (v < 0) v += 1.0;
        break;
    }
    }
}

/* the concrete pattern_at functions, pattern.c:126-195; a/b may be overridden by a nested parent (pattern.c:43-77) */
__device__ __forceinline__ void
concrete_pattern_at(int type, const double *a, const double *b, const double pt[3], double out[3])
{
    const double *sel = a;
    switch (type) {
    case FRT_PAT_CHECKER: {
        int t = (int)floor(pt[0]) + (int)floor(pt[1]) + (int)floor(pt[2]);
        sel = (t % 2 == 0) ? a : b;
        break;
    }
    case FRT_PAT_RING: {
        int t = (int)floor(sqrt(pt[0] * pt[0] + pt[2] * pt[2]));
        sel = (t % 2 == 0) ? a : b;
        break;
    }
    case FRT_PAT_STRIPE: {
        int t = (int)floor(pt[0]);
        sel = (t % 2 == 0) ? a : b;
        break;
    }
    case FRT_PAT_GRADIENT:
    case FRT_PAT_UV_GRADIENT: {
        double fr = pt[0] - floor(pt[0]);
        for (int k = 0; k < 3; ++k) {
            out[k] = a[k] + (b[k] - a[k]) * fr;
        }
        return;
    }
    case FRT_PAT_RADIAL_GRADIENT:
    case FRT_PAT_UV_RADIAL_GRADIENT: {
        double mag = sqrt(pt[0] * pt[0] + pt[2] * pt[2]);
        double fr = mag - floor(mag);
        for (int k = 0; k < 3; ++k) {
            out[k] = a[k] + (b[k] - a[k]) * fr;
        }
        return;
    }
    default:
        break;
    }
    out[0] = sel[0];
    out[1] = sel[1];
    out[2] = sel[2];
}

/* the uv_pattern_at functions, pattern.c:225-300 */
__device__ __noinline__ void
uv_pattern_at(const DScene &S, const frt_pattern &P, double u, double v, double out[3])
{
    switch (P.type) {
    case FRT_PAT_UV_ALIGN_CHECKER: {
        const double *sel = P.c;
        if (v > 0.8) {
            if (u < 0.2) sel = P.c + 3;
            else if (u > 0.8) sel = P.c + 6;
        } else if (v < 0.2) {
            if (u < 0.2) sel = P.c + 9;
            else if (u > 0.8) sel = P.c + 12;
        }
        out[0] = sel[0];
        out[1] = sel[1];
        out[2] = sel[2];
        break;
    }
    case FRT_PAT_UV_CHECKER: {
        int u2 = (int)floor(u * (double)P.i[0]);
        int v2 = (int)floor(v * (double)P.i[1]);
        const double *sel = ((u2 + v2) % 2 == 0) ? P.c : P.c + 3;
        out[0] = sel[0];
        out[1] = sel[1];
        out[2] = sel[2];
        break;
    }
    case FRT_PAT_UV_GRADIENT:
    case FRT_PAT_UV_RADIAL_GRADIENT: {
        double pt[3] = { u, v, 0.0 };
        concrete_pattern_at(P.type, P.c, P.c + 3, pt, out);
        break;
    }
    case FRT_PAT_UV_TEXTURE: { /* uv_texture_uv_pattern_at, pattern.c:287-300 */
        const frt_texture T = S.texs[P.i[0]];
        double vv = 1 - v;
        long col = (long)round(u * (double)(T.width - 1));
        long row = (long)round(vv * (double)(T.height - 1));
        texture_fetch(S, P.i[0], col, row, out);
        break;
    }
    default: /* base_uv_pattern_at, pattern.c:217-223 */
        out[0] = u;
        out[1] = v;
        out[2] = 0;
        break;
    }
}

/*
 * pattern->pattern_at_shape(pattern, shape, world_point) for every pattern kind: base (pattern.c:10-29),
 * blended (:31-40), nested (:43-77), perturbed (:79-117).  `ov` (6 doubles or NULL) carries the colours a nested
 * parent writes into a concrete child (the reference mutates the shared child, pattern.c:56-57; here it is a
 * per-thread override, which is what a race-free run of the reference computes).
 */
#define FRT_PATTERN_DEPTH 3 /* abstract patterns (blended / nested / perturbed) may nest this deep */
template <int D>
__device__ __noinline__ void
pattern_at_shape_d(const DScene &S, int pat, int leaf, const double wp[3], const double *ov, double out[3])
{
    const frt_pattern &P = S.pats[pat];
    if (P.type == FRT_PAT_BLENDED || P.type == FRT_PAT_NESTED || P.type == FRT_PAT_PERTURBED) {
        if constexpr (D > 0) {
            if (P.type == FRT_PAT_BLENDED) {
                double c1[3], c2[3];
                pattern_at_shape_d<D - 1>(S, P.i[0], leaf, wp, NULL, c1);
                pattern_at_shape_d<D - 1>(S, P.i[1], leaf, wp, NULL, c2);
                for (int k = 0; k < 3; ++k) {
                    out[k] = (c1[k] + c2[k]) / 2.0;
                }
            } else if (P.type == FRT_PAT_NESTED) {
                double ab[6];
                pattern_at_shape_d<D - 1>(S, P.i[1], leaf, wp, NULL, ab);
                pattern_at_shape_d<D - 1>(S, P.i[2], leaf, wp, NULL, ab + 3);
                int ct = S.pats[P.i[0]].type;
                bool concrete = ct <= FRT_PAT_STRIPE;
                pattern_at_shape_d<D - 1>(S, P.i[0], leaf, wp, concrete ? ab : NULL, out);
            } else {
                double x = wp[0], y = wp[1], z = wp[2];
                double q[3];
                q[0] = wp[0] + P.f[1] * perlin_pnoise3d(x, y, z, P.f[2], P.f[0], P.i[1], P.i[2]);
                z = z < 0 ? z - 1.0 : z + 1.0;
                q[1] = wp[1] + P.f[1] * perlin_pnoise3d(x, y, z, P.f[2], P.f[0], P.i[1], P.i[2]);
                z = z < 0 ? z - 1.0 : z + 1.0;
                q[2] = wp[2] + P.f[1] * perlin_pnoise3d(x, y, z, P.f[2], P.f[0], P.i[1], P.i[2]);
                pattern_at_shape_d<D - 1>(S, P.i[0], leaf, q, NULL, out);
            }
        } else {
            out[0] = out[1] = out[2] = 0.0; /* nesting deeper than FRT_PATTERN_DEPTH (rejected at scene creation) */
        }
        return;
    }

    /* base_pattern_at_shape: world -> object (whole parent chain) -> pattern space */
    NodeA a = load_node_a(S, leaf);
    NodeB b = load_node_b(S, leaf);
    double op[3], pp[3];
    point_to_local(S, a.xform, wp, op);
    if (P.identity) {
        pp[0] = op[0];
        pp[1] = op[1];
        pp[2] = op[2];
    } else {
        for (int k = 0; k < 3; ++k) {
            pp[k] = P.inv[4 * k + 0] * op[0] + P.inv[4 * k + 1] * op[1] + P.inv[4 * k + 2] * op[2] + P.inv[4 * k + 3];
        }
    }
    if (P.type <= FRT_PAT_STRIPE) {
        concrete_pattern_at(P.type, ov ? ov : P.c, ov ? ov + 3 : P.c + 3, pp, out);
    } else if (P.type >= FRT_PAT_CUBE_MAP) { /* texture_map_pattern_at, pattern.c:198-215 */
        const double *prm = S.params + (b.param < 0 ? 0 : b.param);
        int face;
        double u, v;
        uv_map_eval(P.i[0], a.type, prm, pp, face, u, v);
        const frt_pattern &F = S.pats[P.i[1] + face];
        double fp[3];
        if (F.identity) {
            fp[0] = pp[0];
            fp[1] = pp[1];
            fp[2] = pp[2];
        } else {
            for (int k = 0; k < 3; ++k) {
                fp[k] = F.inv[4 * k + 0] * pp[0] + F.inv[4 * k + 1] * pp[1] + F.inv[4 * k + 2] * pp[2] + F.inv[4 * k + 3];
            }
        }
        int face2;
        uv_map_eval(P.i[0], a.type, prm, fp, face2, u, v);
        uv_pattern_at(S, F, u, v, out);
    } else { /* base_pattern_at, pattern.c:119-124: a uv pattern used without a map returns the point */
        out[0] = pp[0];
        out[1] = pp[1];
        out[2] = pp[2];
    }
}

__device__ __forceinline__ void
pattern_at_shape(const DScene &S, int pat, int leaf, const double wp[3], const double *ov, double out[3], int)
{
    pattern_at_shape_d<FRT_PATTERN_DEPTH>(S, pat, leaf, wp, ov, out);
}
