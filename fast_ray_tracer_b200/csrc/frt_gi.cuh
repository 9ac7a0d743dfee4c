/*
 * frt_gi.cuh -- photon pass, photon map and the global-illumination terms of shade_hit.
 * Included by frt_core.cu after the light stage (uses its RNG, queues and LightRec).
 *
 *   k_photon_trace     trace_photons' emission loop + power_at / photon_hit with     photon_tracer.c:114-257,
 *                      Russian roulette, storing into the caustic / global map        light.c:14-98, pm.c:261
 *   k_pm_bounds/_count/_scan/_scatter   replace pm_balance (pm.c:329): photons are binned into a uniform grid
 *                      of cell size = half the estimate radius, sorted by cell
 *   k_fg_trace         final_gather + color_at_gi: one thread per (hit, CMJ cell)    renderer.c:647, :319
 *   k_gi_points        lighting_caustics / lighting_gi (visualize) query per hit     renderer.c:829, :862
 *   k_knn              pm_irradiance_estimate + pm_locate_photons: one WARP per      pm.c:91-252
 *                      query gathers the photons of the cell rows its search sphere reaches, selects the n nearest
 *                      by a three-pass radix select over the squared distances held in shared memory, sums them
 *   k_gi_resolve       the ambient-slot sum and its sqrt(3) clamp                    renderer.c:755-770
 *
 * Why a grid and not Jensen's left-balanced kd-tree: the estimate depends only on the SET of the n nearest
 * photons within max_dist (and the distance of the farthest one), not on the structure that finds it.  With a fixed
 * search radius a grid of half that cell size touches at most 5 x 5 rows of cells; a row is one contiguous photon
 * range, so a warp streams it with coalesced 16-byte loads instead of chasing a tree one photon at a time.
 *
 * Photon record (32 bytes, two float4): {x, y, z, theta | phi << 8} {power r, g, b, 0} -- theta / phi are Jensen's
 * 8-bit direction (pm.c:286-300).
 */
#pragma once

#include <cuda_fp16.h>

#define FRT_KNN_WARPS 4
#define FRT_KNN_CAP 1024

struct PMView { /* one photon map as the kernels see it */
    const float4 *a, *b;            /* sorted by cell: {position, packed direction} {power} */
    const float4 *c;                /* sorted by cell: the direction pm_photon_dir (pm.c:80-86) reads from the tables, {x, y, z, 0} */
    const unsigned int *cell_start; /* n_cells + 1 */
    float gx, gy, gz, inv_cell;
    int nx, ny, nz;
    unsigned int count;
    const float *dir_tab;           /* sintheta[256] costheta[256] cosphi[256] sinphi[256]: init_Photon_map's tables (pm.c:54-60) */
};

struct GQuery { /* one radiance-estimate request */
    float x, y, z;       /* position */
    float ex, ey, ez;    /* the vector the reference passes as "normal" (the eye vector, renderer.c:842, :875) */
    float wr, wg, wb;    /* weight applied to the scaled estimate */
    unsigned int target; /* hit index; bit 31: accumulate into the final-gather sum, else into the ambient sum; bit 30: caustic map */
};

struct GIParams {
    int usteps, vsteps;      /* final gather grid */
    int n_photons;           /* irradiance_estimate_num */
    float radius, cone_k;
    int visualize;
    int use_caustics, use_final_gather;
    unsigned long long seed;
};

/* ------------------------------------------------------------------------------------------------ sampling */

/* create_coordinate_system, sampler.c:67-85 */
__device__ __forceinline__ void
coordinate_system(const double n[3], double nt[3], double nb[3])
{
    double t0, t1, t2;
    if (fabs(n[0]) > fabs(n[1])) {
        double s = sqrt(n[0] * n[0] + n[2] * n[2]);
        t0 = n[2] * s;
        t1 = 0.0;
        t2 = -n[0] * s;
    } else {
        double s = sqrt(n[1] * n[1] + n[2] * n[2]);
        t0 = 0.0;
        t1 = -n[2] * s;
        t2 = n[1] * s;
    }
    double inv = -1.0 / sqrt(t0 * t0 + t1 * t1 + t2 * t2);
    nt[0] = t0 * inv;
    nt[1] = t1 * inv;
    nt[2] = t2 * inv;
    nb[0] = n[1] * nt[2] - n[2] * nt[1];
    nb[1] = n[2] * nt[0] - n[0] * nt[2];
    nb[2] = n[0] * nt[1] - n[1] * nt[0];
}

/* sampler_hemisphere with cosine weighting (sampler.c:40-64, :88-114) */
__device__ __forceinline__ void
cosine_hemisphere(const double n[3], double r1, double r2, double out[3])
{
    double nt[3], nb[3];
    coordinate_system(n, nt, nb);
    double r = sqrt(r2);
    double sn, cs;
    sincospi(2.0 * r1, &sn, &cs);
    double v0 = r * cs, v2 = r * sn, v1 = sqrt(fmax(0.0, 1.0 - r2));
    double inv = 1.0 / sqrt(v0 * v0 + v1 * v1 + v2 * v2);
    v0 *= inv;
    v1 *= inv;
    v2 *= inv;
    double x = v0 * nb[0] + v1 * n[0] + v2 * nt[0];
    double y = v0 * nb[1] + v1 * n[1] + v2 * nt[1];
    double z = v0 * nb[2] + v1 * n[2] + v2 * nt[2];
    inv = 1.0 / sqrt(x * x + y * y + z * z);
    out[0] = x * inv;
    out[1] = y * inv;
    out[2] = z * inv;
}

/*
 * Entry [qu, qv] of a jittered s x s correlated-multi-jitter table (sampler_reset_2d, sampler.c:415-470) without
 * building the table: the two shuffles are row / column permutations, replayed on a packed index array.
 * Draw numbering follows the reference's order: 2 s^2 canonical jitters, s row picks, s column picks.
 */
__device__ __forceinline__ void
cmj_entry_square(int s, unsigned long long key, int qu, int qv, double &x, double &y)
{
    const unsigned int base = 2u * s * s;
    unsigned long long perm = 0xfedcba9876543210ull; /* nibble v = row now sitting at position v */
    for (int j = 0; j < s; ++j) {
        int k = (int)(j + u01(mix64(key + base + j)) * (s - j));
        unsigned long long a = (perm >> (4 * j)) & 15ull, b = (perm >> (4 * k)) & 15ull;
        perm &= ~((15ull << (4 * j)) | (15ull << (4 * k)));
        perm |= (b << (4 * j)) | (a << (4 * k));
    }
    const int rp = (int)((perm >> (4 * qv)) & 15ull);
    perm = 0xfedcba9876543210ull;
    for (int i = 0; i < s; ++i) {
        int k = (int)(i + u01(mix64(key + base + s + i)) * (s - i));
        unsigned long long a = (perm >> (4 * i)) & 15ull, b = (perm >> (4 * k)) & 15ull;
        perm &= ~((15ull << (4 * i)) | (15ull << (4 * k)));
        perm |= (b << (4 * i)) | (a << (4 * k));
    }
    const int cp = (int)((perm >> (4 * qu)) & 15ull);
    /* canonical entries: arr[2 (j s + i)] = (i + (j + xi) / s) / s, arr[.. + 1] = (j + (i + xi) / s) / s */
    const double xi_x = u01(mix64(key + 2u * (rp * s + qu)));
    const double xi_y = u01(mix64(key + 2u * (qv * s + cp) + 1u));
    x = (qu + (rp + xi_x) / (double)s) / (double)s;
    y = (qv + (cp + xi_y) / (double)s) / (double)s;
}

/* ------------------------------------------------------------------------------------------------ surface */

struct SurfaceG { /* the part of prepare_computations (renderer.c:368-495) the GI paths need */
    double p[3], n[3], eye[3], over[3], under[3], reflv[3];
    double Kd[3];
    int material;
};

template <bool MAPS>
__device__ __forceinline__ void
surface_at(const DScene &S, const Ray &r, const Hit &h, SurfaceG &g)
{
    const NodeA a = load_node_a(S, h.leaf);
    const NodeB b = load_node_b(S, h.leaf);
    const frt_material &M = S.mats[a.material];
    const double *prm = S.params + (b.param < 0 ? 0 : b.param);
    g.material = a.material;
    g.p[0] = r.ox + r.dx * h.t;
    g.p[1] = r.oy + r.dy * h.t;
    g.p[2] = r.oz + r.dz * h.t;
    double lp[3], ln[3];
    point_to_local(S, a.xform, g.p, lp);
    local_normal(a.type, prm, lp, h.u, h.v, ln);
    normal_to_world(S, a.xform, ln, g.n);
    if (MAPS && M.map_bump >= 0) { /* MAPS = false: the scene binds no pattern to any material, the interpreter is left out */
        double tex[3];
        pattern_at_shape(S, M.map_bump, h.leaf, g.p, NULL, tex, 0);
        g.n[0] += 2.0 * tex[0] - 1.0;
        g.n[1] += 2.0 * tex[1] - 1.0;
        g.n[2] += 2.0 * tex[2] - 1.0;
    }
    double inv = 1.0 / sqrt(g.n[0] * g.n[0] + g.n[1] * g.n[1] + g.n[2] * g.n[2]);
    g.eye[0] = -r.dx;
    g.eye[1] = -r.dy;
    g.eye[2] = -r.dz;
    double s = (g.n[0] * g.eye[0] + g.n[1] * g.eye[1] + g.n[2] * g.eye[2]) < 0 ? -inv : inv; /* inside flip */
    double ddot = 0.0;
    for (int k = 0; k < 3; ++k) {
        g.n[k] *= s;
    }
    ddot = 2 * (r.dx * g.n[0] + r.dy * g.n[1] + r.dz * g.n[2]);
    g.reflv[0] = r.dx - g.n[0] * ddot;
    g.reflv[1] = r.dy - g.n[1] * ddot;
    g.reflv[2] = r.dz - g.n[2] * ddot;
    for (int k = 0; k < 3; ++k) {
        g.over[k] = g.p[k] + g.n[k] * FRT_EPS;
        g.under[k] = g.p[k] - g.n[k] * FRT_EPS;
    }
    material_color(S, MAPS ? M.map_Kd : -1, M.Kd, h.leaf, g.over, g.Kd);
}

/* ------------------------------------------------------------------------------------------------ photon pass */

struct PhotonParams {
    int light, map_type; /* 0 = caustic, 1 = global */
    int path_length;
    int rank, world;
    unsigned long long first, count; /* this launch traces photon indices first + i, i < count, of the rank's shard */
    unsigned long long seed;
};

/* pm_store's direction quantisation, pm.c:286-300 */
__device__ __forceinline__ unsigned int
pack_photon_dir(const double d[3])
{
    int theta = (int)(acos(d[2]) * (256.0 / M_PI));
    theta = theta > 255 ? 255 : (theta < 0 ? 0 : theta);
    int phi = (int)(atan2(d[1], d[0]) * (256.0 / (2.0 * M_PI)));
    if (phi > 255) {
        phi = 255;
    } else if (phi < 0) {
        phi = (phi + 256) & 255;
    }
    return (unsigned int)theta | ((unsigned int)phi << 8);
}

template <int PRIMS, bool MAPS>
__global__ void __launch_bounds__(128)
k_photon_trace(DScene S, DSceneF SF, PhotonParams P, float4 *__restrict__ pa, float4 *__restrict__ pb, unsigned int *stored, unsigned int cap,
               Counters *cnt)
{
    const frt_light L = S.lights[P.light];
    int overflow = 0;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < P.count;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long index = (P.first + i) * (unsigned long long)P.world + (unsigned long long)P.rank;
        unsigned long long key = mix64(P.seed ^ ((unsigned long long)(P.map_type + 1) << 56) ^ ((unsigned long long)P.light << 44) ^
                                       (index * 0x9E3779B97F4A7C15ull));
        unsigned int ctr = 0;
#define DRAW() u01(mix64(key + (ctr++)))
        /* ---- emit_photon, light.c:14-98 */
        Ray r;
        if (L.type == 3 /* POINT_LIGHT: a direction in the unit ball, NOT normalised (light.c:82-97) */) {
            double dx, dy, dz;
            int guard = 0;
            do {
                dx = 2 * DRAW() - 1;
                dy = 2 * DRAW() - 1;
                dz = 2 * DRAW() - 1;
            } while (dx * dx + dy * dy + dz * dz > 1 && ++guard < 64);
            r = Ray{ L.position[0], L.position[1], L.position[2], dx, dy, dz };
        } else {
            double o[3] = { L.position[0], L.position[1], L.position[2] };
            if (L.type == 0 || L.type == 1) { /* area / circle: a random point of a random cached sample set */
                unsigned long long set = (unsigned long long)(DRAW() * L.cache_len);
                unsigned long long pt = (unsigned long long)(DRAW() * L.num_samples);
                set = set >= (unsigned long long)L.cache_len ? L.cache_len - 1 : set;
                pt = pt >= (unsigned long long)L.num_samples ? L.num_samples - 1 : pt;
                const double *p = S.lpoints + 3 * (L.point_offset + set * L.num_samples + pt);
                o[0] = __ldg(p);
                o[1] = __ldg(p + 1);
                o[2] = __ldg(p + 2);
            }
            const double r1 = DRAW(), r2 = DRAW(); /* sampler_2d(true, 1, 1): two jitters ... */
            ctr += 2;                              /* ... and two shuffle picks that swap an entry with itself */
            double d[3];
            cosine_hemisphere(L.normal, r1, r2, d);
            r = Ray{ o[0], o[1], o[2], d[0], d[1], d[2] };
        }
        double power[3] = { L.intensity[0], L.intensity[1], L.intensity[2] };
        bool had_diffuse = false, had_specular = false;

        /* ---- power_at / photon_hit, photon_tracer.c:114-201, as a loop */
        for (int remaining = P.path_length; remaining > 0; --remaining) {
            const Hit h = trace_closest_mixed<true, PRIMS>(S, SF, r, &overflow); /* hit(xs, true) */
            if (h.leaf < 0) {
                break;
            }
            if (power[0] <= 0 && power[1] <= 0 && power[2] <= 0) {
                break;
            }
            SurfaceG g;
            surface_at<MAPS>(S, r, h, g);
            const frt_material &M = S.mats[g.material];
            double refl[3];
            material_color(S, MAPS ? M.map_refl : -1, M.refl, h.leaf, g.over, refl);
            const double avg_d = (g.Kd[0] + g.Kd[1] + g.Kd[2]) / 3.0;
            if (g.Kd[0] > 0 || g.Kd[1] > 0 || g.Kd[2] > 0) {
                const bool store = (P.map_type == 0) ? had_specular : had_diffuse;
                if (store) {
                    const unsigned int slot = atomicAdd(stored, 1u);
                    if (slot < cap) {
                        const double dir[3] = { r.dx, r.dy, r.dz };
                        pa[slot] = make_float4((float)g.p[0], (float)g.p[1], (float)g.p[2], __uint_as_float(pack_photon_dir(dir)));
                        pb[slot] = make_float4((float)(g.Kd[0] * power[0]), (float)(g.Kd[1] * power[1]), (float)(g.Kd[2] * power[2]), 0.f);
                    }
                    if (P.map_type == 0) {
                        break; /* a caustic photon is stored once */
                    }
                }
            }
            /* Russian roulette, photon_tracer.c:156-180 */
            const double rr = DRAW();
            const double avg_s = (refl[0] + refl[1] + refl[2]) / 3.0;
            const double avg_t = (M.Tf[0] + M.Tf[1] + M.Tf[2]) / 3.0;
            int action; /* 0 diffuse, 1 specular, 2 refract, 3 absorbed */
            if (P.map_type == 1) {
                const double tot = avg_d + avg_s + avg_t;
                action = (rr * tot < avg_d) ? 0 : (rr * tot < avg_d + avg_s) ? 1 : (rr * tot < avg_d + avg_s + avg_t) ? 2 : 3;
            } else {
                const double tot = avg_s + avg_t;
                action = (rr * tot < avg_s) ? 1 : (rr * tot < avg_s + avg_t) ? 2 : 3;
            }
            if (action == 0) { /* reflect_photon_diffuse :33-62 */
                for (int k = 0; k < 3; ++k) {
                    power[k] *= g.Kd[k];
                }
                const double r1 = DRAW(), r2 = DRAW();
                ctr += 2;
                double d[3];
                cosine_hemisphere(g.n, r1, r2, d);
                r = Ray{ g.over[0], g.over[1], g.over[2], d[0], d[1], d[2] };
                had_diffuse = true;
            } else if (action == 1) { /* reflect_photon_specular :64-78 */
                if (!M.reflective) {
                    break;
                }
                for (int k = 0; k < 3; ++k) {
                    power[k] *= 1.0 / avg_s;
                }
                r = Ray{ g.over[0], g.over[1], g.over[2], g.reflv[0], g.reflv[1], g.reflv[2] };
                had_specular = true;
            } else if (action == 2) { /* refract_photon :80-112 */
                if (fabs(M.Tr) < FRT_EPS) {
                    break;
                }
                double n1 = 1.0, n2 = 1.0;
                trace_containers(S, r, h.leaf, n1, n2, &overflow);
                const double n_ratio = n1 / n2;
                const double cos_i = g.eye[0] * g.n[0] + g.eye[1] * g.n[1] + g.eye[2] * g.n[2];
                const double sin2_t = n_ratio * n_ratio * (1.0 - cos_i * cos_i);
                if (sin2_t > 1.0) {
                    break;
                }
                const double sc = n_ratio * cos_i - sqrt(1.0 - sin2_t);
                for (int k = 0; k < 3; ++k) {
                    power[k] *= 1.0 / avg_t;
                }
                r = Ray{ g.under[0], g.under[1], g.under[2], g.n[0] * sc - g.eye[0] * n_ratio, g.n[1] * sc - g.eye[1] * n_ratio,
                         g.n[2] * sc - g.eye[2] * n_ratio };
                had_specular = true;
            } else {
                break;
            }
        }
#undef DRAW
    }
    if (overflow) {
        atomicOr(&cnt->overflow_csg, 1u);
    }
}

/* ------------------------------------------------------------------------------------------------ grid build */

__device__ __forceinline__ unsigned int
float_order(float f) /* monotone map float -> uint */
{
    unsigned int u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__host__ __device__ __forceinline__ float
float_unorder(unsigned int u)
{
    unsigned int v = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
    float f;
#ifdef __CUDA_ARCH__
    f = __uint_as_float(v);
#else
    memcpy(&f, &v, sizeof(f));
#endif
    return f;
}

__global__ void
k_pm_scale_bounds(float4 *__restrict__ a, float4 *__restrict__ b, unsigned int n, float scale, unsigned int *bounds /* min xyz, max xyz */)
{
    unsigned int mn[3] = { 0xffffffffu, 0xffffffffu, 0xffffffffu }, mx[3] = { 0u, 0u, 0u };
    for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 p = a[i];
        float4 w = b[i];
        w.x *= scale; /* pm_scale_photon_power(map, 1 / photon_count), photon_tracer.c:251-253 */
        w.y *= scale;
        w.z *= scale;
        b[i] = w;
        const unsigned int o[3] = { float_order(p.x), float_order(p.y), float_order(p.z) };
        for (int k = 0; k < 3; ++k) {
            mn[k] = min(mn[k], o[k]);
            mx[k] = max(mx[k], o[k]);
        }
    }
    for (int k = 0; k < 3; ++k) {
        mn[k] = __reduce_min_sync(0xffffffffu, mn[k]);
        mx[k] = __reduce_max_sync(0xffffffffu, mx[k]);
    }
    if ((threadIdx.x & 31) == 0) {
        for (int k = 0; k < 3; ++k) {
            atomicMin(bounds + k, mn[k]);
            atomicMax(bounds + 3 + k, mx[k]);
        }
    }
}

__device__ __forceinline__ int
pm_cell_of(const PMView &M, float x, float y, float z)
{
    int cx = min(max((int)floorf((x - M.gx) * M.inv_cell), 0), M.nx - 1);
    int cy = min(max((int)floorf((y - M.gy) * M.inv_cell), 0), M.ny - 1);
    int cz = min(max((int)floorf((z - M.gz) * M.inv_cell), 0), M.nz - 1);
    return (cz * M.ny + cy) * M.nx + cx;
}

__global__ void
k_pm_count(PMView M, const float4 *__restrict__ a, unsigned int n, unsigned int *counts)
{
    for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 p = a[i];
        atomicAdd(counts + pm_cell_of(M, p.x, p.y, p.z), 1u);
    }
}

/* exclusive scan of counts[0..n) into start[0..n], single block */
__global__ void __launch_bounds__(1024)
k_pm_scan(const unsigned int *__restrict__ counts, unsigned int *__restrict__ start, unsigned int n)
{
    __shared__ unsigned int s_part[1024];
    const unsigned int per = (n + 1023u) / 1024u;
    const unsigned int lo = min(threadIdx.x * per, n), hi = min(lo + per, n);
    unsigned int sum = 0;
    for (unsigned int i = lo; i < hi; ++i) {
        sum += counts[i];
    }
    s_part[threadIdx.x] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int run = 0;
        for (int k = 0; k < 1024; ++k) {
            unsigned int v = s_part[k];
            s_part[k] = run;
            run += v;
        }
        start[n] = run;
    }
    __syncthreads();
    unsigned int run = s_part[threadIdx.x];
    for (unsigned int i = lo; i < hi; ++i) {
        start[i] = run;
        run += counts[i];
    }
}

__global__ void
k_pm_scatter(PMView M, const float4 *__restrict__ a, const float4 *__restrict__ b, unsigned int n, unsigned int *cursor,
             float4 *__restrict__ sa, float4 *__restrict__ sb)
{
    for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 p = a[i];
        const unsigned int slot = atomicAdd(cursor + pm_cell_of(M, p.x, p.y, p.z), 1u);
        sa[slot] = p;
        sb[slot] = b[i]; /* the 8-bit direction stays in sa[slot].w; k_knn looks it up in the reference's tables */
    }
}

/* pm_photon_dir (pm.c:80-86) for every photon, once: the products the facing test of the radiance estimate multiplies out */
__global__ void
k_pm_dirs(const float4 *__restrict__ sa, unsigned int n, const float *__restrict__ tab, float4 *__restrict__ sc)
{
    for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const unsigned int dbits = __float_as_uint(sa[i].w);
        const unsigned int theta = dbits & 255u, phi = (dbits >> 8) & 255u;
        const float st = __ldg(tab + theta);
        sc[i] = make_float4(st * __ldg(tab + 512 + phi), st * __ldg(tab + 768 + phi), __ldg(tab + 256 + theta), 0.f);
    }
}

/* ------------------------------------------------------------------------------------------------ queries */

/*
 * final_gather (renderer.c:647-687) + color_at_gi (:319-345): one thread per (hit, CMJ cell).  The gather ray's first
 * hit, if diffuse, becomes a radiance-estimate request weighted by  pi * Kd * (eye . n) * rands[0]  (shade_hit_gi
 * :626-645, lighting_gi :862-892, the "scale by theta" of :672).
 */
template <int PRIMS, bool MAPS>
__global__ void __launch_bounds__(128)
k_fg_trace(DScene S, DSceneF SF, FrameParams F, GIParams G, const LightRec *__restrict__ recs, unsigned int first_hit, unsigned int n_hits_batch,
           GQuery *__restrict__ queries, unsigned int *n_queries, unsigned int qcap, Counters *cnt, int level)
{
    const unsigned int cells = (unsigned int)(G.usteps * G.vsteps);
    const unsigned long long total = (unsigned long long)n_hits_batch * cells;
    int overflow = 0;
    unsigned long long n_rays = 0;
    for (unsigned long long base = (unsigned long long)blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < total;
         base += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long item = base + (threadIdx.x & 31);
        bool want = false;
        GQuery q;
        if (item < total) {
            const unsigned int h = first_hit + (unsigned int)(item / cells);
            const int c = (int)(item % cells);
            const LightRec *R = recs + h;
            if (R->Kd[0] > 0 || R->Kd[1] > 0 || R->Kd[2] > 0) { /* renderer.c:739 */
                const int qu = c % G.usteps, qv = c / G.usteps;
                const unsigned long long key = mix64(G.seed ^ 0x66676174686572ull ^ ((unsigned long long)R->rng << 24) ^ (unsigned long long)level);
                double r1, r2;
                if (G.usteps == G.vsteps && G.usteps <= 16) {
                    cmj_entry_square(G.usteps, key, qu, qv, r1, r2);
                } else {
                    cmj_jittered(G.usteps, G.vsteps, key, qu, qv, r1, r2);
                }
                const double n[3] = { R->n[0], R->n[1], R->n[2] };
                double d[3];
                cosine_hemisphere(n, r1, r2, d);
                const Ray r{ R->over[0], R->over[1], R->over[2], d[0], d[1], d[2] };
                const Hit hit = trace_closest_mixed<false, PRIMS>(S, SF, r, &overflow); /* instantiated for the primitive types the scene holds, like k_extend */
                ++n_rays;
                if (hit.leaf >= 0) {
                    /* color_at_gi tests the diffuse colour at the hit point itself (:331-337) ... */
                    const frt_material &M = S.mats[load_node_a(S, hit.leaf).material];
                    const double p[3] = { r.ox + r.dx * hit.t, r.oy + r.dy * hit.t, r.oz + r.dz * hit.t };
                    double kd0[3];
                    material_color(S, MAPS ? M.map_Kd : -1, M.Kd, hit.leaf, p, kd0);
                    if (kd0[0] > 0 || kd0[1] > 0 || kd0[2] > 0) {
                        SurfaceG g; /* ... and lighting_gi uses over_Kd, sampled at over_point (:862-866) */
                        surface_at<MAPS>(S, r, hit, g);
                        if (g.Kd[0] > 0.0 || g.Kd[1] > 0.0 || g.Kd[2] > 0.0) {
                            const double edn = g.eye[0] * g.n[0] + g.eye[1] * g.n[1] + g.eye[2] * g.n[2];
                            const double base_w = M_PI * r1;
                            q.x = (float)g.over[0];
                            q.y = (float)g.over[1];
                            q.z = (float)g.over[2];
                            q.ex = (float)g.eye[0];
                            q.ey = (float)g.eye[1];
                            q.ez = (float)g.eye[2];
                            if (G.visualize) {
                                q.wr = q.wg = q.wb = (float)base_w;
                            } else {
                                q.wr = (float)(base_w * g.Kd[0] * edn);
                                q.wg = (float)(base_w * g.Kd[1] * edn);
                                q.wb = (float)(base_w * g.Kd[2] * edn);
                            }
                            q.target = h | 0x80000000u;
                            want = true;
                        }
                    }
                }
            }
        }
        const unsigned int slot = warp_append(n_queries, want);
        if (want) {
            if (slot < qcap) {
                queries[slot] = q;
            } else {
                atomicOr(&cnt->overflow_queue, 1u);
            }
        }
    }
    if (overflow) {
        atomicOr(&cnt->overflow_csg, 1u);
    }
    for (int o = 16; o > 0; o >>= 1) {
        n_rays += __shfl_down_sync(0xffffffffu, n_rays, o);
    }
    if ((threadIdx.x & 31) == 0 && n_rays) {
        atomicAdd(&cnt->rays_gather, n_rays);
    }
}

/* lighting_caustics (renderer.c:829-860) and the visualize_photon_map call of lighting_gi (:743-751): one request per hit */
__global__ void __launch_bounds__(256)
k_gi_points(FrameParams F, GIParams G, const LightRec *__restrict__ recs, unsigned int first_hit, unsigned int n_hits_batch,
            GQuery *__restrict__ queries, unsigned int *n_queries, unsigned int qcap, Counters *cnt, int want_caustic, int want_global)
{
    for (unsigned int base = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < n_hits_batch; base += gridDim.x * blockDim.x) {
        const unsigned int i = base + (threadIdx.x & 31);
        bool live = false;
        GQuery q;
        unsigned int h = 0;
        if (i < n_hits_batch) {
            h = first_hit + i;
            const LightRec *R = recs + h;
            if (R->Kd[0] > 0 || R->Kd[1] > 0 || R->Kd[2] > 0) {
                live = true;
                const double edn = R->eye[0] * R->n[0] + R->eye[1] * R->n[1] + R->eye[2] * R->n[2];
                q.x = (float)R->over[0];
                q.y = (float)R->over[1];
                q.z = (float)R->over[2];
                q.ex = (float)R->eye[0];
                q.ey = (float)R->eye[1];
                q.ez = (float)R->eye[2];
                if (G.visualize) {
                    q.wr = q.wg = q.wb = 1.0f;
                } else {
                    q.wr = (float)(R->Kd[0] * edn);
                    q.wg = (float)(R->Kd[1] * edn);
                    q.wb = (float)(R->Kd[2] * edn);
                }
            }
        }
        if (want_caustic) {
            const unsigned int slot = warp_append(n_queries, live);
            if (live) {
                q.target = h | 0x40000000u;
                if (slot < qcap) queries[slot] = q; else atomicOr(&cnt->overflow_queue, 1u);
            }
        }
        if (want_global) {
            const unsigned int slot = warp_append(n_queries, live);
            if (live) {
                q.target = h;
                if (slot < qcap) queries[slot] = q; else atomicOr(&cnt->overflow_queue, 1u);
            }
        }
    }
}

/*
 * Requests sorted by grid cell before k_knn.  The gather rays of a hit scatter over the whole scene, so consecutive
 * requests of the queue land in unrelated cells and every warp of k_knn pulls its candidate photons from L2 on its own
 * (ncu: L1 hit rate 4 %, long_scoreboard 4.4 warps per issue).  A counting sort by the cell of the request's position
 * (the photon grid's own cells, x fastest) makes neighbours in the queue neighbours in space: the four warps of a block
 * and the blocks of an SM then read the same rows of cells.  Cost: two passes over 40-byte records.
 */
__global__ void
k_gq_count(PMView M, const GQuery *__restrict__ q, const unsigned int *n_queries, unsigned int qcap, unsigned int *__restrict__ counts)
{
    const unsigned int n = min(*n_queries, qcap);
    for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        atomicAdd(counts + pm_cell_of(M, q[i].x, q[i].y, q[i].z), 1u);
    }
}

__global__ void
k_gq_scatter(PMView M, const GQuery *__restrict__ q, const unsigned int *n_queries, unsigned int qcap, unsigned int *__restrict__ cursor,
             GQuery *__restrict__ out)
{
    const unsigned int n = min(*n_queries, qcap);
    for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const GQuery g = q[i];
        out[atomicAdd(cursor + pm_cell_of(M, g.x, g.y, g.z), 1u)] = g;
    }
}

/*
 * Warp-cooperative selection: the 24-bit key of the want-th smallest of d2[0..count) (count > want), keys being
 * d2 * scale truncated to 24 bits.  Three 8-bit radix passes, each a shared-memory histogram of the candidates that
 * still match the digits found so far and a warp scan over its 256 bins -- three passes over the list instead of
 * the 31 of a bit-wise bisection.
 */
__device__ __forceinline__ unsigned int
knn_key(float v, float scale)
{
    return min((unsigned int)(v * scale), 0xffffffu);
}

__device__ __forceinline__ unsigned int
warp_select_key(const float *d2, unsigned int count, unsigned int want, float scale, unsigned int *hist, int lane,
                unsigned int *take_from_bin = nullptr, unsigned int *bin_size = nullptr)
{
    unsigned int prefix = 0, remaining = want; /* remaining: rank of the target among the keys that match `prefix` */
    for (int shift = 16; shift >= 0; shift -= 8) {
        for (int b = lane; b < 256; b += 32) {
            hist[b] = 0;
        }
        __syncwarp();
        const unsigned int hi_mask = shift == 16 ? 0u : (0xffffffu >> (shift + 8)) << (shift + 8);
        for (unsigned int k = lane; k < count; k += 32) {
            const unsigned int key = knn_key(d2[k], scale);
            if ((key & hi_mask) == prefix) {
                atomicAdd(&hist[(key >> shift) & 255u], 1u);
            }
        }
        __syncwarp();
        /* lane l owns bins 8l .. 8l+7 */
        unsigned int mine[8], sum = 0;
        for (int b = 0; b < 8; ++b) {
            mine[b] = hist[8 * lane + b];
            sum += mine[b];
        }
        unsigned int incl = sum;
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned int up = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) {
                incl += up;
            }
        }
        const unsigned int excl = incl - sum;
        /* the lane whose range [excl, incl) contains rank `remaining` (1-based) */
        const bool has = remaining > excl && remaining <= incl;
        const unsigned int owner_mask = __ballot_sync(0xffffffffu, has);
        const int owner = __ffs(owner_mask) - 1;
        unsigned int digit = 0, before = 0;
        if (lane == owner) {
            unsigned int run = excl;
            for (int b = 0; b < 8; ++b) {
                if (remaining <= run + mine[b]) {
                    digit = 8 * lane + b;
                    before = run;
                    break;
                }
                run += mine[b];
            }
        }
        digit = __shfl_sync(0xffffffffu, digit, owner < 0 ? 0 : owner);
        before = __shfl_sync(0xffffffffu, before, owner < 0 ? 0 : owner);
        prefix |= digit << shift;
        remaining -= before;
        if (shift == 0 && bin_size != nullptr) {
            *bin_size = hist[digit]; /* candidates that share the selected 24-bit key */
        }
        __syncwarp();
    }
    if (take_from_bin != nullptr) {
        *take_from_bin = remaining; /* how many of them belong to the `want` smallest */
    }
    return prefix;
}

/* Exact membership among the candidates that share the selected key (a few, and rarely more than are wanted): candidate k
 * belongs to the n nearest iff fewer than `take` of them precede it in (distance, list position) order. */
__device__ __forceinline__ bool
knn_tie_kept(const float *d2, unsigned int count, unsigned int k, unsigned int T, float scale, unsigned int take)
{
    const float v = d2[k];
    unsigned int rank = 0;
    for (unsigned int j = 0; j < count; ++j) {
        const float u = d2[j];
        if (knn_key(u, scale) == T && (u < v || (u == v && j < k))) {
            ++rank;
        }
    }
    return rank < take;
}

/*
 * pm_irradiance_estimate (pm.c:91-156): the n nearest photons within max_dist of the request, cone-filtered sum of
 * those whose direction faces the "normal", density from the distance of the farthest one; fewer than 8 photons give
 * nothing (:121).  The callers' rescaling (100 / found for the caustic map, 10 n / found for the global map,
 * renderer.c:845, :878) and the request's weight are applied here and the result is added to the hit's sum.
 *
 * One warp per request.  The grid's cells are half the search radius wide; the warp walks the rows of cells that the
 * search sphere can reach (a row = cells consecutive in x = one contiguous photon range), clipped to the sphere, and
 * keeps the squared distances < r^2 with their photon indices in shared memory.
 */
__global__ void __launch_bounds__(FRT_KNN_WARPS * 32)
k_knn_list(PMView MC, PMView MG, GIParams G, const GQuery *__restrict__ queries, const unsigned int *n_queries, unsigned int qcap,
      double *__restrict__ acc_amb, double *__restrict__ acc_fg, int *__restrict__ found_out)
{
    __shared__ float s_d2[FRT_KNN_WARPS][FRT_KNN_CAP];
    __shared__ unsigned int s_idx[FRT_KNN_WARPS][FRT_KNN_CAP];
    __shared__ unsigned int s_hist[FRT_KNN_WARPS][256];
    __shared__ float s_dir[1024]; /* pm_photon_dir's tables (pm.c:80-86): the facing test must flip where the reference's does */
    {
        const float *tab = MG.dir_tab != nullptr ? MG.dir_tab : MC.dir_tab;
        for (int k = threadIdx.x; k < 1024; k += blockDim.x) {
            s_dir[k] = tab != nullptr ? __ldg(tab + k) : 0.f;
        }
        __syncthreads();
    }
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned int nq = min(*n_queries, qcap);
    const unsigned int lt = (1u << lane) - 1u;
    float *d2 = s_d2[wib];
    unsigned int *idx = s_idx[wib];
    unsigned int *hist = s_hist[wib];
    const float R2 = G.radius * G.radius;
    const unsigned int want_n = (unsigned int)G.n_photons;

    for (unsigned int qi = blockIdx.x * FRT_KNN_WARPS + wib; qi < nq; qi += gridDim.x * FRT_KNN_WARPS) {
        const GQuery q = queries[qi];
        const bool caustic = (q.target & 0x40000000u) != 0;
        const PMView &M = caustic ? MC : MG;
        if (M.count == 0) {
            continue;
        }
        unsigned int count = 0;
        float r2cur = R2;
        const float cell = 1.0f / M.inv_cell;
        const float fy = (q.y - M.gy) * M.inv_cell, fz = (q.z - M.gz) * M.inv_cell;
        const int cy = (int)floorf(fy), cz = (int)floorf(fz);
        const int reach = (int)ceilf(G.radius * M.inv_cell); /* <= 2: the cells are at least half a radius wide */
        const int side = 2 * reach + 1;
        /* lane r owns row r of the (2 reach + 1)^2 rows of cells around the request: clip it to the search sphere and fetch
         * its photon range -- all rows at once, so the dependent loads (cell_start, then photons) are paid once per request */
        unsigned int row_s = 0, row_len = 0;
        if (lane < side * side) {
            const int z = cz - reach + lane / side, y = cy - reach + lane % side;
            if (z >= 0 && z < M.nz && y >= 0 && y < M.ny) {
                const float dzc = fmaxf(fmaxf((float)z - fz, fz - (float)(z + 1)), 0.0f) * cell;
                const float dyc = fmaxf(fmaxf((float)y - fy, fy - (float)(y + 1)), 0.0f) * cell;
                const float rem = R2 - dzc * dzc - dyc * dyc;
                if (rem > 0.0f) {
                    const float dxm = sqrtf(rem) * 1.0001f + 1e-7f;
                    const int x0 = max((int)floorf((q.x - dxm - M.gx) * M.inv_cell), 0);
                    const int x1 = min((int)floorf((q.x + dxm - M.gx) * M.inv_cell), M.nx - 1);
                    if (x0 <= x1) { /* cells of one x-row are contiguous in memory: one photon range */
                        const unsigned int base = (unsigned int)((z * M.ny + y) * M.nx);
                        row_s = __ldg(M.cell_start + base + x0);
                        row_len = __ldg(M.cell_start + base + x1 + 1) - row_s;
                    }
                }
            }
        }
        unsigned int incl = row_len;
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned int up = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) {
                incl += up;
            }
        }
        const unsigned int total = __shfl_sync(0xffffffffu, incl, 31);
        /* four chunks of 32 candidates in flight per step: the photon loads are L2 hits of a few hundred cycles each and the
         * warp has nothing else to overlap them with */
        for (unsigned int j0 = 0; j0 < total; j0 += 128) {
            unsigned int pp[4];
            float4 aa[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const unsigned int j = j0 + 32 * u + lane;
                /* which row holds candidate j: the first lane whose inclusive prefix exceeds j (binary search by shuffles) */
                int lo_r = 0;
                for (int step = 16; step > 0; step >>= 1) {
                    const unsigned int v = __shfl_sync(0xffffffffu, incl, min(lo_r + step - 1, 31));
                    if (v <= j) {
                        lo_r += step;
                    }
                }
                const int r = min(lo_r, 31);
                const unsigned int r_incl = __shfl_sync(0xffffffffu, incl, r), r_len = __shfl_sync(0xffffffffu, row_len, r);
                const unsigned int r_s = __shfl_sync(0xffffffffu, row_s, r);
                pp[u] = r_s + (j - (r_incl - r_len));
                aa[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (j < total) {
                    aa[u] = __ldg(M.a + pp[u]);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
            const unsigned int j = j0 + 32 * u + lane;
            const unsigned int p = pp[u];
            bool hit = false;
            float dd = 0.f;
            if (j < total) {
                const float dx = aa[u].x - q.x, dy = aa[u].y - q.y, dz = aa[u].z - q.z;
                dd = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
                hit = dd < r2cur;
            }
            const unsigned int mask = __ballot_sync(0xffffffffu, hit);
            if (hit) {
                const unsigned int pos = count + __popc(mask & lt);
                d2[pos] = dd;
                idx[pos] = p;
            }
            count += __popc(mask);
            __syncwarp();
            if (count > FRT_KNN_CAP - 32) {
                /* list nearly full: keep only the want_n nearest seen so far and tighten the radius */
                const float scale = 16777216.0f / r2cur;
                const unsigned int T = warp_select_key(d2, count, want_n, scale, hist, lane);
                unsigned int kept = 0;
                float mx = 0.f;
                for (unsigned int k0 = 0; k0 < count; k0 += 32) {
                    const unsigned int k = k0 + lane;
                    const float v = (k < count) ? d2[k] : 0.f;
                    const unsigned int id = (k < count) ? idx[k] : 0u;
                    const bool keep = (k < count) && knn_key(v, scale) <= T;
                    const unsigned int m2 = __ballot_sync(0xffffffffu, keep);
                    __syncwarp();
                    if (keep) {
                        const unsigned int pos = kept + __popc(m2 & lt);
                        d2[pos] = v;
                        idx[pos] = id;
                        mx = fmaxf(mx, v);
                    }
                    kept += __popc(m2);
                    __syncwarp();
                }
                for (int o = 16; o > 0; o >>= 1) {
                    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                }
                count = kept;
                r2cur = nextafterf(mx, 3.0e38f); /* candidates must now beat the current n-th distance */
            }
            }
        }
        __syncwarp();
        /* the n nearest: entries whose key is <= the key of the n-th smallest squared distance */
        const float scale = 16777216.0f / r2cur;
        unsigned int T = 0xffffffu;
        unsigned int found = count;
        const bool select = count > want_n;
        unsigned int take = 0, bin = 0;
        if (select) {
            T = warp_select_key(d2, count, want_n, scale, hist, lane, &take, &bin);
            found = want_n;
        }
        const bool ties = select && bin > take; /* more candidates share the n-th key than fit: order them exactly */
        float sr = 0.f, sg = 0.f, sb = 0.f, far2 = 0.f;
        if (found >= 8) {
            const float inv_kr = 1.0f / (G.cone_k * G.radius);
            for (unsigned int k = lane; k < count; k += 32) {
                const float v = d2[k];
                const unsigned int key = knn_key(v, scale);
                if (key < T || (key == T && (!ties || knn_tie_kept(d2, count, k, T, scale, take)))) {
                    far2 = fmaxf(far2, v);
                    const unsigned int id = idx[k];
                    const float4 pw = __ldg(M.b + id);
                    const unsigned int dbits = __float_as_uint(__ldg(M.a + id).w);
                    const unsigned int theta = dbits & 255u, phi = (dbits >> 8) & 255u;
                    const float st = s_dir[theta];
                    const float dot = fmaf(st * s_dir[512 + phi], q.ex, fmaf(st * s_dir[768 + phi], q.ey, s_dir[256 + theta] * q.ez));
                    if (dot < 0.0f) {
                        const float w = 1.0f - sqrtf(v) * inv_kr;
                        sr = fmaf(pw.x, w, sr);
                        sg = fmaf(pw.y, w, sg);
                        sb = fmaf(pw.z, w, sb);
                    }
                }
            }
            for (int o = 16; o > 0; o >>= 1) {
                sr += __shfl_xor_sync(0xffffffffu, sr, o);
                sg += __shfl_xor_sync(0xffffffffu, sg, o);
                sb += __shfl_xor_sync(0xffffffffu, sb, o);
                far2 = fmaxf(far2, __shfl_xor_sync(0xffffffffu, far2, o));
            }
        }
        if (lane == 0 && found_out != nullptr) {
            found_out[q.target & 0x3fffffffu] = (int)found; /* frt_photons_estimate: pm_irradiance_estimate's return value */
        }
        if (lane == 0 && found >= 8) {
            /* np.dist2[0] (pm.c:147): the search radius^2 until the heap of n photons is full, then the n-th distance^2 */
            const double r2_density = (select || r2cur < R2) ? (double)far2 : (double)R2;
            const double density = 1.0 / ((1.0 - 2.0 / (3.0 * (double)G.cone_k)) * (M_PI * r2_density));
            /* the callers' rescale (renderer.c:845, :878); frt_photons_estimate returns the estimate as pm.c does */
            const double rescale = found_out != nullptr ? 1.0 : (caustic ? 100.0 / (double)found : 10.0 * (double)G.n_photons / (double)found);
            const double f = density * rescale;
            double *acc = ((q.target & 0x80000000u) ? acc_fg : acc_amb) + 3 * (size_t)(q.target & 0x3fffffffu);
            const double vr = f * (double)sr * (double)q.wr, vg = f * (double)sg * (double)q.wg, vb = f * (double)sb * (double)q.wb;
            if (vr != 0.0) atomicAdd(acc, vr);
            if (vg != 0.0) atomicAdd(acc + 1, vg);
            if (vb != 0.0) atomicAdd(acc + 2, vb);
        }
        __syncwarp();
    }
}

/*
 * The same estimate WITHOUT a candidate list.  k_knn_list keeps every photon inside the search sphere (distance and index,
 * 8 KB of shared memory per warp: 24 resident warps per SM) and then radix-selects the n nearest in three passes over that
 * list; ncu (profiles/r1b_k_knn_radix.txt): 41 % issue utilisation at 27 % occupancy, ~2 100 warp-instructions per request.
 * Here the candidates are STREAMED twice from the photon grid (the second time out of L1 / L2):
 *   pass 1   histogram of the squared distances over 256 equal bins of [0, r^2)           (1 KB per warp)
 *            -> the bin B that holds the n-th nearest photon, and how many of its photons belong to the n nearest;
 *            a bin with more than FRT_KNN_TIES photons is split once more into 256 sub-bins by a second histogram pass
 *   pass 2   photons in bins below B are summed on the fly; those of bin B go to a short list, from which the missing
 *            ones are taken in exact (distance, position) order -- the result is the set pm_locate_photons keeps.
 * 2 KB of shared memory per warp, so residency is bounded by registers, not shared memory.
 */
#define FRT_KNN_TIES 128

/* which candidates the n nearest are: key16 < t_lo taken, t_lo <= key16 <= t_hi tie class, above: not */
__device__ __forceinline__ unsigned int
knn_key16(float dd, float kscale)
{
    return min((unsigned int)(dd * kscale), 65535u);
}

/* the bin of hist[0..256) in which the cumulative count reaches `want` (1-based), the count before it and its own count */
__device__ __forceinline__ void
knn_find_bin(const unsigned int *hist, unsigned int want, int lane, unsigned int &bin, unsigned int &before, unsigned int &inside,
             unsigned int &total)
{
    unsigned int mine[8], sum = 0;
    for (int b = 0; b < 8; ++b) {
        mine[b] = hist[8 * lane + b];
        sum += mine[b];
    }
    unsigned int incl = sum;
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned int up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) {
            incl += up;
        }
    }
    total = __shfl_sync(0xffffffffu, incl, 31);
    const unsigned int excl = incl - sum;
    const bool has = want > excl && want <= incl;
    const int owner = __ffs(__ballot_sync(0xffffffffu, has)) - 1;
    unsigned int b_ = 255u, bef = 0u, ins = 0u;
    if (lane == owner) {
        unsigned int run = excl;
        for (int b = 0; b < 8; ++b) {
            if (want <= run + mine[b]) {
                b_ = 8u * lane + b;
                bef = run;
                ins = mine[b];
                break;
            }
            run += mine[b];
        }
    }
    const int src = owner < 0 ? 0 : owner;
    bin = __shfl_sync(0xffffffffu, b_, src);
    before = __shfl_sync(0xffffffffu, bef, src);
    inside = __shfl_sync(0xffffffffu, ins, src);
}

#ifndef FRT_KNN_MINB
#define FRT_KNN_MINB 8
#endif
__global__ void __launch_bounds__(FRT_KNN_WARPS * 32, FRT_KNN_MINB)
k_knn(PMView MC, PMView MG, GIParams G, const GQuery *__restrict__ queries, const unsigned int *n_queries, unsigned int qcap,
      double *__restrict__ acc_amb, double *__restrict__ acc_fg, int *__restrict__ found_out)
{
    __shared__ unsigned int s_hist[FRT_KNN_WARPS][256];
    __shared__ float s_td2[FRT_KNN_WARPS][FRT_KNN_TIES];
    __shared__ unsigned int s_tid[FRT_KNN_WARPS][FRT_KNN_TIES];
    __shared__ float s_dir[1024]; /* pm_photon_dir's tables (pm.c:80-86): the facing test must flip where the reference's does */
    {
        const float *tab = MG.dir_tab != nullptr ? MG.dir_tab : MC.dir_tab;
        for (int k = threadIdx.x; k < 1024; k += blockDim.x) {
            s_dir[k] = tab != nullptr ? __ldg(tab + k) : 0.f;
        }
        __syncthreads();
    }
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned int nq = min(*n_queries, qcap);
    const unsigned int lt = (1u << lane) - 1u;
    unsigned int *hist = s_hist[wib];
    float *td2 = s_td2[wib];
    unsigned int *tid = s_tid[wib];
    const float R2 = G.radius * G.radius;
    const float kscale = 65536.0f / R2;
    const unsigned int want_n = (unsigned int)G.n_photons;
    const float inv_kr = 1.0f / (G.cone_k * G.radius);

    for (unsigned int qi = blockIdx.x * FRT_KNN_WARPS + wib; qi < nq; qi += gridDim.x * FRT_KNN_WARPS) {
        const GQuery q = queries[qi];
        const bool caustic = (q.target & 0x40000000u) != 0;
        const PMView &M = caustic ? MC : MG;
        if (M.count == 0) {
            continue;
        }
        const float cell = 1.0f / M.inv_cell;
        const float fy = (q.y - M.gy) * M.inv_cell, fz = (q.z - M.gz) * M.inv_cell;
        const int cy = (int)floorf(fy), cz = (int)floorf(fz);
        const int reach = (int)ceilf(G.radius * M.inv_cell); /* <= 2: the cells are at least half a radius wide */
        const int side = 2 * reach + 1;
        /* lane r owns row r of the (2 reach + 1)^2 rows of cells around the request, clipped to the search sphere */
        unsigned int row_s = 0, row_len = 0;
        if (lane < side * side) {
            const int z = cz - reach + lane / side, y = cy - reach + lane % side;
            if (z >= 0 && z < M.nz && y >= 0 && y < M.ny) {
                const float dzc = fmaxf(fmaxf((float)z - fz, fz - (float)(z + 1)), 0.0f) * cell;
                const float dyc = fmaxf(fmaxf((float)y - fy, fy - (float)(y + 1)), 0.0f) * cell;
                const float rem = R2 - dzc * dzc - dyc * dyc;
                if (rem > 0.0f) {
                    const float dxm = sqrtf(rem) * 1.0001f + 1e-7f;
                    const int x0 = max((int)floorf((q.x - dxm - M.gx) * M.inv_cell), 0);
                    const int x1 = min((int)floorf((q.x + dxm - M.gx) * M.inv_cell), M.nx - 1);
                    if (x0 <= x1) {
                        const unsigned int base = (unsigned int)((z * M.ny + y) * M.nx);
                        row_s = __ldg(M.cell_start + base + x0);
                        row_len = __ldg(M.cell_start + base + x1 + 1) - row_s;
                    }
                }
            }
        }
        unsigned int incl = row_len;
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned int up = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) {
                incl += up;
            }
        }
        const unsigned int total = __shfl_sync(0xffffffffu, incl, 31);
        const unsigned int rows_mask = __ballot_sync(0xffffffffu, row_len > 0u);
        if (total == 0u) {
            if (found_out != nullptr && lane == 0) {
                found_out[q.target & 0x3fffffffu] = 0;
            }
            continue;
        }

        /* mode 0: histogram of the high key byte; mode 1: of the low byte inside bin B (only when B is crowded);
         * mode 2: sum the photons below the tie class, list the tie class */
        unsigned int t_lo = 65536u, t_hi = 65536u; /* tie class in key16 space; everything below t_lo is taken */
        unsigned int need = 0, found = 0, n_ties = 0, B = 0;
        bool select = false;
        float sr = 0.f, sg = 0.f, sb = 0.f, far2 = 0.f;
        for (int mode = 0; mode < 3; ++mode) {
            if (mode < 2) {
                for (int b = lane; b < 256; b += 32) {
                    hist[b] = 0;
                }
                __syncwarp();
            }
            /* four rows of cells at a time: lane k takes photon k of each (a row is one contiguous photon range, ~25 photons
             * at 1 M photons), so four independent loads are in flight and nobody searches for the row of a candidate --
             * the flattened candidate list this replaces spent 30 % of the kernel's instructions on that search */
            for (unsigned int rm = rows_mask; rm != 0u;) {
                unsigned int rs[4], rl[4], lmax = 0;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int r = rm ? __ffs(rm) - 1 : 0;
                    const bool live = rm != 0u;
                    rm &= rm - 1u;
                    rs[u] = __shfl_sync(0xffffffffu, row_s, r);
                    rl[u] = live ? __shfl_sync(0xffffffffu, row_len, r) : 0u;
                    lmax = max(lmax, rl[u]);
                }
                for (unsigned int k0 = 0; k0 < lmax; k0 += 32) {
                unsigned int pp[4];
                float4 aa[4];
                bool valid[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const unsigned int k = k0 + lane;
                    valid[u] = k < rl[u];
                    pp[u] = rs[u] + k;
                    aa[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (valid[u]) {
                        aa[u] = __ldg(M.a + pp[u]);
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float dx = aa[u].x - q.x, dy = aa[u].y - q.y, dz = aa[u].z - q.z;
                    const float dd = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
                    const bool in = valid[u] && dd < R2;
                    const unsigned int key = knn_key16(dd, kscale);
                    if (mode == 0) {
                        if (in) {
                            atomicAdd(&hist[key >> 8], 1u);
                        }
                    } else if (mode == 1) {
                        if (in && (key >> 8) == B) {
                            atomicAdd(&hist[key & 255u], 1u);
                        }
                    } else {
                        const bool take = in && key < t_lo;
                        const bool tie = in && key >= t_lo && key <= t_hi;
                        if (take) {
                            far2 = fmaxf(far2, dd);
                            const float4 pw = __ldg(M.b + pp[u]);
                            const unsigned int dbits = __float_as_uint(aa[u].w);
                            const unsigned int theta = dbits & 255u, phi = (dbits >> 8) & 255u;
                            const float st = s_dir[theta];
                            const float dot = fmaf(st * s_dir[512 + phi], q.ex, fmaf(st * s_dir[768 + phi], q.ey, s_dir[256 + theta] * q.ez));
                            if (dot < 0.0f) {
                                const float w = 1.0f - sqrtf(dd) * inv_kr;
                                sr = fmaf(pw.x, w, sr);
                                sg = fmaf(pw.y, w, sg);
                                sb = fmaf(pw.z, w, sb);
                            }
                        }
                        const unsigned int tm = __ballot_sync(0xffffffffu, tie);
                        if (tie) {
                            const unsigned int pos = n_ties + __popc(tm & lt);
                            if (pos < FRT_KNN_TIES) {
                                td2[pos] = dd;
                                tid[pos] = pp[u];
                            }
                        }
                        n_ties += __popc(tm);
                    }
                }
                }
            }
            __syncwarp();
            if (mode == 0) {
                unsigned int before, inside, count;
                knn_find_bin(hist, want_n, lane, B, before, inside, count);
                found = min(count, want_n);
                select = count > want_n;
                if (found < 8) {
                    break; /* pm.c:121: fewer than 8 photons give nothing */
                }
                if (!select) {
                    mode = 1; /* every photon inside the sphere is taken: straight to the sums */
                    continue;
                }
                need = want_n - before;
                t_lo = B << 8;
                t_hi = (B << 8) | 255u;
                if (inside <= FRT_KNN_TIES) {
                    mode = 1; /* the tie class fits the list */
                    continue;
                }
            } else if (mode == 1) {
                unsigned int B2, before2, inside2, count2;
                knn_find_bin(hist, need, lane, B2, before2, inside2, count2);
                need -= before2;
                t_lo = (B << 8) | B2;
                t_hi = t_lo;
            }
        }
        if (found_out != nullptr && lane == 0) {
            found_out[q.target & 0x3fffffffu] = (int)found;
        }
        if (found < 8) {
            continue;
        }
        if (select) {
            /* the tie class: `need` of its n_ties photons belong to the n nearest, in (distance, position) order.  (More than
             * FRT_KNN_TIES photons within r^2 / 65536 of one another: the first ones in stream order stand for the class.) */
            const unsigned int m = min(n_ties, (unsigned int)FRT_KNN_TIES);
            for (unsigned int k = lane; k < m; k += 32) {
                const float v = td2[k];
                unsigned int rank = 0;
                for (unsigned int j = 0; j < m; ++j) {
                    const float u = td2[j];
                    rank += (u < v || (u == v && j < k)) ? 1u : 0u;
                }
                if (rank < need) {
                    far2 = fmaxf(far2, v);
                    const unsigned int id = tid[k];
                    const float4 pw = __ldg(M.b + id);
                    const unsigned int dbits = __float_as_uint(__ldg(M.a + id).w);
                    const unsigned int theta = dbits & 255u, phi = (dbits >> 8) & 255u;
                    const float st = s_dir[theta];
                    const float dot = fmaf(st * s_dir[512 + phi], q.ex, fmaf(st * s_dir[768 + phi], q.ey, s_dir[256 + theta] * q.ez));
                    if (dot < 0.0f) {
                        const float w = 1.0f - sqrtf(v) * inv_kr;
                        sr = fmaf(pw.x, w, sr);
                        sg = fmaf(pw.y, w, sg);
                        sb = fmaf(pw.z, w, sb);
                    }
                }
            }
        }
        for (int o = 16; o > 0; o >>= 1) {
            sr += __shfl_xor_sync(0xffffffffu, sr, o);
            sg += __shfl_xor_sync(0xffffffffu, sg, o);
            sb += __shfl_xor_sync(0xffffffffu, sb, o);
            far2 = fmaxf(far2, __shfl_xor_sync(0xffffffffu, far2, o));
        }
        if (lane == 0) {
            /* np.dist2[0] (pm.c:147): the search radius^2 until the heap of n photons is full, then the n-th distance^2 */
            const double r2_density = select ? (double)far2 : (double)R2;
            const double density = 1.0 / ((1.0 - 2.0 / (3.0 * (double)G.cone_k)) * (M_PI * r2_density));
            /* the callers' rescale (renderer.c:845, :878); frt_photons_estimate returns the estimate as pm.c does */
            const double rescale = found_out != nullptr ? 1.0 : (caustic ? 100.0 / (double)found : 10.0 * (double)G.n_photons / (double)found);
            const double f = density * rescale;
            double *acc = ((q.target & 0x80000000u) ? acc_fg : acc_amb) + 3 * (size_t)(q.target & 0x3fffffffu);
            const double vr = f * (double)sr * (double)q.wr, vg = f * (double)sg * (double)q.wg, vb = f * (double)sb * (double)q.wb;
            if (vr != 0.0) atomicAdd(acc, vr);
            if (vg != 0.0) atomicAdd(acc + 1, vg);
            if (vb != 0.0) atomicAdd(acc + 2, vb);
        }
        __syncwarp();
    }
}

/*
 * The same estimate with ONE LANE PER REQUEST and the candidates shared by a block (production path).
 *
 * k_knn gives every request a warp that pulls its own ~700 candidate photons from L2 twice (ncu: L1 hit rate 8 %, L2 hit
 * rate 99 %, issue 74 %, L1TEX 66 %): 3 800 warp-instructions and 30 KB of L2 traffic per request, and a final-gather
 * frame has 655 M of them against 1 M photons.  Here a batch of requests is first sorted by the photon grid's cell
 * (k_gq_count / scan / k_gq_scatter); a block takes 128 consecutive requests, and for each cell among them streams the
 * photons of the (2 reach + 1)^3 cells around it -- every photon a request of that cell can reach -- through shared
 * memory in tiles.  Every lane runs its own request over the tile: the candidate is a broadcast read, the distance is
 * 6 FP32 operations, and what a candidate costs besides (addresses, global loads, the direction looked up in the
 * reference's tables, pm.c:54-60, :80-86) is paid once per block, not once per request.
 *
 * Selection of the n nearest, per lane: pass 1 counts the photons inside the sphere into the lane's private column of a
 * shared-memory histogram (128 bins x 16 bit over the squared distance); a sphere with no more than n photons takes them
 * all (pm.c:147), otherwise the bin B of the n-th nearest follows from the column.  Pass 2 sums the photons of the bins
 * below B and LISTS the photons of bin B (a handful: 1 / 128 of the sphere), of which the nearest `need` are then taken
 * in exact (distance, position) order -- the same set k_knn takes.  A request whose bin B holds more than FRT_KC_TIES
 * photons, or that asks for the other map, is appended to `fallback` and goes through k_knn afterwards.
 */
#define FRT_KC_T 128      /* requests (= threads) per block */
#define FRT_KC_BINS 128
#define FRT_KC_TIES 16
#define FRT_KC_BUF_BYTES 11264                 /* one of the two tile buffers */
#define FRT_KC_TILE1 (FRT_KC_BUF_BYTES / 16)   /* pass 1: position only */
#define FRT_KC_TILE2 224                       /* pass 2: position, power, direction (48 bytes) */
#define FRT_KC_SMEM (2 * FRT_KC_BUF_BYTES + FRT_KC_BINS * FRT_KC_T * 2 + FRT_KC_TIES * FRT_KC_T * 8 + 65 * 4)

__device__ __forceinline__ void
cp_async16(void *smem, const void *gmem)
{
    const unsigned int s = (unsigned int)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem) : "memory");
}

__device__ __forceinline__ void
cp_async_commit()
{
    asm volatile("cp.async.commit_group;\n" ::: "memory");
}

template <int N>
__device__ __forceinline__ void
cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

__global__ void __launch_bounds__(FRT_KC_T, 3)
k_knn_cell(PMView M, int caustic_map, GIParams G, const GQuery *__restrict__ queries, const unsigned int *n_queries, unsigned int qcap,
           unsigned int *work /* [0] next chunk of requests, [1] requests in `fallback` */, double *__restrict__ acc_amb,
           double *__restrict__ acc_fg, int *__restrict__ found_out, GQuery *__restrict__ fallback, unsigned int fb_cap)
{
    extern __shared__ __align__(16) unsigned char kc_raw[];
    unsigned short *hist = reinterpret_cast<unsigned short *>(kc_raw + 2 * FRT_KC_BUF_BYTES);
    float *tdd = reinterpret_cast<float *>(hist + FRT_KC_BINS * FRT_KC_T);
    unsigned int *tix = reinterpret_cast<unsigned int *>(tdd + FRT_KC_TIES * FRT_KC_T);
    unsigned int *row_s = tix + FRT_KC_TIES * FRT_KC_T;
    unsigned int *row_pre = row_s + 32; /* 33 entries: candidates in front of row r */
    __shared__ unsigned int s_first, s_cur;

    const int tid = threadIdx.x, lane = tid & 31;
    const unsigned int nq = min(*n_queries, qcap);
    const float R2 = G.radius * G.radius;
    const float kscale = (float)FRT_KC_BINS / R2;
    const unsigned int want_n = (unsigned int)G.n_photons;
    const float inv_kr = 1.0f / (G.cone_k * G.radius);
    const int reach = (int)ceilf(G.radius * M.inv_cell);
    const int side = 2 * reach + 1;
    unsigned int *h32 = reinterpret_cast<unsigned int *>(hist) + tid; /* this lane's column: bins 2w and 2w + 1 in the halves of h32[w * FRT_KC_T] */
    auto bin_count = [&](int b) { return (h32[(b >> 1) * FRT_KC_T] >> ((b & 1) * 16)) & 0xffffu; };

    auto to_fallback = [&](const GQuery &q) {
        const unsigned int slot = atomicAdd(&work[1], 1u);
        if (slot < fb_cap) {
            fallback[slot] = q;
        }
    };
    /* the photon behind candidate c of the current cell's list */
    auto photon_of = [&](unsigned int c) {
        int r = 0;
#pragma unroll
        for (int step = 16; step > 0; step >>= 1) {
            r += row_pre[r + step] <= c ? step : 0; /* the last row whose first candidate is <= c */
        }
        return row_s[r] + (c - row_pre[r]);
    };
    /* tiles travel global -> shared with cp.async into one buffer while the lanes work on the other */
    auto buffer = [&](unsigned int k) { return reinterpret_cast<float4 *>(kc_raw + (k & 1u) * FRT_KC_BUF_BYTES); };
    auto fetch1 = [&](unsigned int t0, unsigned int n, float4 *dst) {
        for (unsigned int c = tid; c < n; c += FRT_KC_T) {
            cp_async16(dst + c, M.a + photon_of(t0 + c));
        }
        cp_async_commit();
    };
    auto fetch2 = [&](unsigned int t0, unsigned int n, float4 *dst) {
        for (unsigned int c = tid; c < n; c += FRT_KC_T) {
            const unsigned int p = photon_of(t0 + c);
            cp_async16(dst + c, M.a + p);
            cp_async16(dst + FRT_KC_TILE2 + c, M.b + p);
            cp_async16(dst + 2 * FRT_KC_TILE2 + c, M.c + p);
        }
        cp_async_commit();
    };

    for (;;) {
        __syncthreads();
        if (tid == 0) {
            s_first = atomicAdd(&work[0], (unsigned int)FRT_KC_T);
        }
        __syncthreads();
        const unsigned int first = s_first;
        if (first >= nq) {
            break;
        }
        const unsigned int qi = first + tid;
        GQuery q{};
        bool pending = qi < nq;
        unsigned int my_cell = 0xffffffffu;
        if (pending) {
            q = queries[qi];
            if (((q.target & 0x40000000u) != 0) != (caustic_map != 0) || side * side > 32) {
                to_fallback(q);
                pending = false;
            } else {
                my_cell = (unsigned int)pm_cell_of(M, q.x, q.y, q.z);
            }
        }
        for (;;) {
            __syncthreads();
            if (tid == 0) {
                s_cur = 0xffffffffu;
            }
            __syncthreads();
            if (pending) {
                atomicMin(&s_cur, my_cell);
            }
            __syncthreads();
            const unsigned int cur = s_cur;
            if (cur == 0xffffffffu) {
                break;
            }
            bool act = pending && my_cell == cur;
            /* the rows of cells (consecutive in x = one contiguous photon range) within `reach` cells of the cell */
            if (tid < 32) {
                const int cx = (int)(cur % (unsigned int)M.nx), cy = (int)((cur / (unsigned int)M.nx) % (unsigned int)M.ny);
                const int cz = (int)(cur / ((unsigned int)M.nx * (unsigned int)M.ny));
                unsigned int rs = 0, rl = 0;
                if (tid < side * side) {
                    const int z = cz - reach + tid / side, y = cy - reach + tid % side;
                    if (z >= 0 && z < M.nz && y >= 0 && y < M.ny) {
                        const int x0 = max(cx - reach, 0), x1 = min(cx + reach, M.nx - 1);
                        const unsigned int base = (unsigned int)((z * M.ny + y) * M.nx);
                        rs = __ldg(M.cell_start + base + x0);
                        rl = __ldg(M.cell_start + base + x1 + 1) - rs;
                    }
                }
                unsigned int incl = rl;
                for (int o = 1; o < 32; o <<= 1) {
                    const unsigned int up = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) {
                        incl += up;
                    }
                }
                row_s[tid] = rs;
                row_pre[tid + 1] = incl;
                if (tid == 0) {
                    row_pre[0] = 0;
                }
            }
            __syncthreads();
            const unsigned int C = row_pre[32];
            if (act) {
                pending = false;
                if (C > 65535u) { /* a histogram column counts in 16 bits */
                    to_fallback(q);
                    act = false;
                }
            }
            const bool warp_act = __any_sync(0xffffffffu, act);

            /* pass 1: photons inside the sphere, histogram of their squared distances */
            unsigned int cnt = 0;
            if (act) {
                for (int b = 0; b < FRT_KC_BINS / 2; ++b) {
                    h32[b * FRT_KC_T] = 0;
                }
            }
            if (C > 0) {
                fetch1(0, min(C, (unsigned int)FRT_KC_TILE1), buffer(0));
            }
            for (unsigned int t0 = 0, k = 0; t0 < C; t0 += FRT_KC_TILE1, ++k) {
                const unsigned int nt0 = min(C - t0, (unsigned int)FRT_KC_TILE1);
                if (t0 + FRT_KC_TILE1 < C) {
                    fetch1(t0 + FRT_KC_TILE1, min(C - t0 - FRT_KC_TILE1, (unsigned int)FRT_KC_TILE1), buffer(k + 1));
                    cp_async_wait<1>();
                } else {
                    cp_async_wait<0>();
                }
                __syncthreads();
                if (warp_act) {
                    const float4 *sP = buffer(k);
                    auto count = [&](float dd) {
                        /* two 16-bit bins to a word, bumped with a shared-memory reduction: nothing waits for the old value, and
                         * nothing branches -- a photon outside the sphere adds 0 to the last bin */
                        const bool in = act && dd < R2;
                        const unsigned int key = min((unsigned int)(dd * kscale), (unsigned int)FRT_KC_BINS - 1u);
                        cnt += in ? 1u : 0u;
                        atomicAdd(h32 + (key >> 1) * FRT_KC_T, in ? 1u << ((key & 1u) * 16u) : 0u);
                    };
                    unsigned int c = 0;
                    for (; c + 4 <= nt0; c += 4) {
                        const float4 p0 = sP[c], p1 = sP[c + 1], p2 = sP[c + 2], p3 = sP[c + 3];
                        float dx = p0.x - q.x, dy = p0.y - q.y, dz = p0.z - q.z;
                        const float d0 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
                        dx = p1.x - q.x, dy = p1.y - q.y, dz = p1.z - q.z;
                        const float d1 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
                        dx = p2.x - q.x, dy = p2.y - q.y, dz = p2.z - q.z;
                        const float d2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
                        dx = p3.x - q.x, dy = p3.y - q.y, dz = p3.z - q.z;
                        const float d3 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
                        count(d0);
                        count(d1);
                        count(d2);
                        count(d3);
                    }
                    for (; c < nt0; ++c) {
                        const float4 p0 = sP[c];
                        const float dx = p0.x - q.x, dy = p0.y - q.y, dz = p0.z - q.z;
                        count(fmaf(dx, dx, fmaf(dy, dy, dz * dz)));
                    }
                }
                __syncthreads(); /* the buffer is refilled two tiles on */
            }
            const bool select = act && cnt > want_n;
            unsigned int B = 0xffffffffu, need = 0; /* bins below B are taken whole; `need` photons of bin B */
            if (select) {
                unsigned int run = 0;
                int b = 0;
                for (; b < FRT_KC_BINS - 1; ++b) {
                    const unsigned int hv = bin_count(b);
                    if (run + hv >= want_n) {
                        break;
                    }
                    run += hv;
                }
                B = (unsigned int)b;
                need = want_n - run;
            }

            /* pass 2: sums over the bins below B; the photons of bin B are listed */
            float sr = 0.f, sg = 0.f, sb = 0.f, far2 = 0.f;
            unsigned int nt = 0;
            if (C > 0) {
                fetch2(0, min(C, (unsigned int)FRT_KC_TILE2), buffer(0));
            }
            for (unsigned int t0 = 0, k = 0; t0 < C; t0 += FRT_KC_TILE2, ++k) {
                const unsigned int nt0 = min(C - t0, (unsigned int)FRT_KC_TILE2);
                if (t0 + FRT_KC_TILE2 < C) {
                    fetch2(t0 + FRT_KC_TILE2, min(C - t0 - FRT_KC_TILE2, (unsigned int)FRT_KC_TILE2), buffer(k + 1));
                    cp_async_wait<1>();
                } else {
                    cp_async_wait<0>();
                }
                __syncthreads();
                if (warp_act) {
                    const float4 *sP = buffer(k), *sQ = sP + FRT_KC_TILE2, *sD = sP + 2 * FRT_KC_TILE2;
                    auto visit = [&](unsigned int c, float dd) {
                        if (act && dd < R2) {
                            const unsigned int key = min((unsigned int)(dd * kscale), (unsigned int)FRT_KC_BINS - 1u);
                            if (key < B) {
                                const float4 pw = sQ[c], dir = sD[c];
                                far2 = fmaxf(far2, dd);
                                const float dot = fmaf(dir.x, q.ex, fmaf(dir.y, q.ey, dir.z * q.ez));
                                if (dot < 0.0f) {
                                    const float w = 1.0f - sqrtf(dd) * inv_kr;
                                    sr = fmaf(pw.x, w, sr);
                                    sg = fmaf(pw.y, w, sg);
                                    sb = fmaf(pw.z, w, sb);
                                }
                            } else if (key == B) {
                                if (nt < FRT_KC_TIES) {
                                    tdd[nt * FRT_KC_T + tid] = dd;
                                    tix[nt * FRT_KC_T + tid] = t0 + c;
                                }
                                ++nt;
                            }
                        }
                    };
                    unsigned int c = 0;
                    for (; c + 4 <= nt0; c += 4) {
                        const float4 p0 = sP[c], p1 = sP[c + 1], p2 = sP[c + 2], p3 = sP[c + 3];
                        float dx = p0.x - q.x, dy = p0.y - q.y, dz = p0.z - q.z;
                        const float d0 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
                        dx = p1.x - q.x, dy = p1.y - q.y, dz = p1.z - q.z;
                        const float d1 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
                        dx = p2.x - q.x, dy = p2.y - q.y, dz = p2.z - q.z;
                        const float d2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
                        dx = p3.x - q.x, dy = p3.y - q.y, dz = p3.z - q.z;
                        const float d3 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
                        visit(c, d0);
                        visit(c + 1, d1);
                        visit(c + 2, d2);
                        visit(c + 3, d3);
                    }
                    for (; c < nt0; ++c) {
                        const float4 p0 = sP[c];
                        const float dx = p0.x - q.x, dy = p0.y - q.y, dz = p0.z - q.z;
                        visit(c, fmaf(dx, dx, fmaf(dy, dy, dz * dz)));
                    }
                }
                __syncthreads();
            }
            if (act) {
                if (nt > FRT_KC_TIES) {
                    to_fallback(q); /* a crowded bin: k_knn splits it once more */
                } else {
                    /* bin B: the nearest `need` in (distance, position) order */
                    for (unsigned int k = 0; k < nt; ++k) {
                        const float v = tdd[k * FRT_KC_T + tid];
                        unsigned int rank = 0;
                        for (unsigned int j = 0; j < nt; ++j) {
                            const float u = tdd[j * FRT_KC_T + tid];
                            rank += (u < v || (u == v && j < k)) ? 1u : 0u;
                        }
                        if (rank < need) {
                            const unsigned int p = photon_of(tix[k * FRT_KC_T + tid]);
                            const float4 pw = __ldg(M.b + p), dir = __ldg(M.c + p);
                            far2 = fmaxf(far2, v);
                            const float dot = fmaf(dir.x, q.ex, fmaf(dir.y, q.ey, dir.z * q.ez));
                            if (dot < 0.0f) {
                                const float w = 1.0f - sqrtf(v) * inv_kr;
                                sr = fmaf(pw.x, w, sr);
                                sg = fmaf(pw.y, w, sg);
                                sb = fmaf(pw.z, w, sb);
                            }
                        }
                    }
                    const unsigned int found = min(cnt, want_n);
                    if (found_out != nullptr) {
                        found_out[q.target & 0x3fffffffu] = (int)found;
                    }
                    if (found >= 8) { /* pm.c:121: fewer than 8 photons give nothing */
                        /* np.dist2[0] (pm.c:147): the search radius^2 until the heap of n photons is full, then the n-th distance^2 */
                        const double r2_density = select ? (double)far2 : (double)R2;
                        const double density = 1.0 / ((1.0 - 2.0 / (3.0 * (double)G.cone_k)) * (M_PI * r2_density));
                        const double rescale = found_out != nullptr ? 1.0
                                                                    : (caustic_map ? 100.0 / (double)found : 10.0 * (double)G.n_photons / (double)found);
                        const double f = density * rescale;
                        double *acc = ((q.target & 0x80000000u) ? acc_fg : acc_amb) + 3 * (size_t)(q.target & 0x3fffffffu);
                        const double vr = f * (double)sr * (double)q.wr, vg = f * (double)sg * (double)q.wg, vb = f * (double)sb * (double)q.wb;
                        if (vr != 0.0) atomicAdd(acc, vr);
                        if (vg != 0.0) atomicAdd(acc + 1, vg);
                        if (vb != 0.0) atomicAdd(acc + 2, vb);
                    }
                }
            }
        }
    }
}

/* exclusive scan of a large array in three launches: per-block sums, k_pm_scan over them, per-block scans */
#define FRT_SCAN_PER_THREAD 8
#define FRT_SCAN_CHUNK (1024 * FRT_SCAN_PER_THREAD)
__global__ void __launch_bounds__(1024)
k_scan_partial(const unsigned int *__restrict__ counts, unsigned int n, unsigned int *__restrict__ partial)
{
    __shared__ unsigned int s_w[32];
    const size_t base = (size_t)blockIdx.x * FRT_SCAN_CHUNK;
    unsigned int sum = 0;
    for (int k = 0; k < FRT_SCAN_PER_THREAD; ++k) {
        const size_t i = base + (size_t)k * 1024 + threadIdx.x;
        sum += i < n ? counts[i] : 0u;
    }
    sum = __reduce_add_sync(0xffffffffu, sum);
    if ((threadIdx.x & 31) == 0) {
        s_w[threadIdx.x >> 5] = sum;
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        const unsigned int v = __reduce_add_sync(0xffffffffu, s_w[threadIdx.x]);
        if (threadIdx.x == 0) {
            partial[blockIdx.x] = v;
        }
    }
}

__global__ void __launch_bounds__(1024)
k_scan_apply(const unsigned int *__restrict__ counts, const unsigned int *__restrict__ block_excl, unsigned int n, unsigned int *__restrict__ start)
{
    __shared__ unsigned int s_w[32];
    const size_t base = (size_t)blockIdx.x * FRT_SCAN_CHUNK + (size_t)threadIdx.x * FRT_SCAN_PER_THREAD;
    unsigned int v[FRT_SCAN_PER_THREAD], sum = 0;
    for (int k = 0; k < FRT_SCAN_PER_THREAD; ++k) {
        v[k] = base + k < n ? counts[base + k] : 0u;
        sum += v[k];
    }
    unsigned int incl = sum;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned int up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) {
            incl += up;
        }
    }
    if (lane == 31) {
        s_w[w] = incl;
    }
    __syncthreads();
    if (w == 0) {
        unsigned int x = s_w[lane], xi = x;
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned int up = __shfl_up_sync(0xffffffffu, xi, o);
            if (lane >= o) {
                xi += up;
            }
        }
        s_w[lane] = xi - x;
    }
    __syncthreads();
    unsigned int run = block_excl[blockIdx.x] + s_w[w] + (incl - sum);
    for (int k = 0; k < FRT_SCAN_PER_THREAD; ++k) {
        if (base + k < n) {
            start[base + k] = run;
        }
        run += v[k];
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 1023) {
        start[n] = run;
    }
}

/* the ambient slot of shade_hit (renderer.c:737-770): direct ambient + indirect + final gather + caustics, clamped to a
 * sum of sqrt(3) on diffuse surfaces, weighted into the pixel */
__global__ void __launch_bounds__(256)
k_gi_resolve(FrameParams F, GIParams G, const LightRec *__restrict__ recs, const double *__restrict__ acc_amb,
             const double *__restrict__ acc_fg, double *__restrict__ canvas, const Counters *cnt, int level)
{
    const unsigned int n = min(cnt->n_hits[level], F.capacity);
    const double fg_scale = 2.0 * M_PI / (double)(G.usteps * G.vsteps);
    for (unsigned int h = blockIdx.x * blockDim.x + threadIdx.x; h < n; h += gridDim.x * blockDim.x) {
        const LightRec *R = recs + h;
        double amb[3];
        const bool diffuse = R->Kd[0] > 0 || R->Kd[1] > 0 || R->Kd[2] > 0;
        for (int k = 0; k < 3; ++k) {
            amb[k] = acc_amb[3 * (size_t)h + k];
            if (diffuse && G.use_final_gather) {
                amb[k] += acc_fg[3 * (size_t)h + k] * fg_scale * R->Kd[k];
            }
        }
        if (diffuse) {
            const double len = amb[0] + amb[1] + amb[2];
            if (len > 1.7320508075688772) {
                for (int k = 0; k < 3; ++k) {
                    amb[k] = amb[k] * (1.0 / len) * 1.7320508075688772;
                }
            }
        }
        double *px = canvas + 4 * (size_t)R->pixel;
        for (int k = 0; k < 3; ++k) {
            const double v = R->w[k] * amb[k];
            if (v != 0.0) {
                atomicAdd(px + k, v);
            }
        }
    }
}
