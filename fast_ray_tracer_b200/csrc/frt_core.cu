/*
 * frt_core.cu -- wavefront render kernels of the B200-native core and the C ABI over them.
 *
 * Replaces, for one frame, the reference's pthread row pool + recursion (renderer.c:194-281, :347-366):
 *
 *   k_raygen   pixel_multi_sample + ray_for_pixel + sample_aperture      renderer.c:131, :95; camera.c:85
 *   k_extend   intersect_world(false) + hit(false)                       world.c:164, intersection.c:42
 *   k_shade    prepare_computations + the reflect/refract split of       renderer.c:368, :497, :534, :607, :689
 *              shade_hit, with per-ray RGB throughput instead of recursion
 *   k_light    light->intensity_at (shadow rays) + lighting_microfacet   light.c:229/:245, renderer.c:73, :894
 *
 * Rays live in SoA queues in HBM; every stage is a persistent grid-stride kernel that reads its item count
 * from device memory, so a frame needs no host round trip between stages.  Queue appends are warp-aggregated
 * (one atomic per warp).  Pixels accumulate in the Canvas layout itself (double[4] per pixel) with FP64 atomics.
 */
#include <cuda_runtime.h>
#include <math_constants.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "frt_b200.h"
#include "frt_internal.h"
#include "frt_device.cuh"
#include "frt_patterns.cuh"
#include "frt_shadow_f32.cuh"
#include "frt_lightgen.cuh"
#include "frt_leafruns.h"

/* ------------------------------------------------------------------------------------------------ errors */

static thread_local char g_err[512] = "";

extern "C" int
frt_set_error(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

extern "C" const char *
frt_last_error(void)
{
    return g_err;
}

extern "C" int
frt_abi_version(void)
{
    return FRT_ABI_VERSION;
}

extern "C" int
frt_abi_sizeof(const char *name)
{
    if (name == nullptr) return -1;
#define SZ(T) if (strcmp(name, #T) == 0) return (int)sizeof(T)
    SZ(frt_node); SZ(frt_xform); SZ(frt_material); SZ(frt_pattern); SZ(frt_texture); SZ(frt_light);
    SZ(frt_camera); SZ(frt_config); SZ(frt_scene_desc); SZ(frt_render_cfg); SZ(frt_stats); SZ(frt_photon_cfg);
#undef SZ
    return -1;
}

#define CK(call)                                                                                             \
    do {                                                                                                     \
        cudaError_t e_ = (call);                                                                             \
        if (e_ != cudaSuccess) {                                                                             \
            return frt_set_error(FRT_ERR_CUDA, "%s:%d: %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
        }                                                                                                    \
    } while (0)

extern "C" int
frt_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        return 0;
    }
    return n;
}

/* ------------------------------------------------------------------------------------------------ device data */

struct RayQ { /* SoA ray queue */
    double *ox, *oy, *oz, *dx, *dy, *dz; /* ray */
    double *wr, *wg, *wb;                /* RGB throughput of the path so far */
    int *pixel;                          /* canvas pixel index (row * hsize + col) */
    unsigned int *rng;                   /* path id for the counter-based RNG */
};

struct HitQ { /* result of k_extend, indexed like the ray queue */
    double *t, *u, *v;
    int *leaf;
};

/*
 * One shaded hit handed to the light stage (AoS: the lanes of a hit read the same record).  112 bytes: the over-point
 * stays FP64 (it is the origin of the hit's FP64 shadow rays), everything the light stage only feeds into its FP32
 * sums -- normal, eye vector, material colours, the path's weight -- is FP32.  The all-FP64 record was 184 bytes and
 * k_shade, which writes one per hit, ran at 41 % of the HBM peak with 19 % issue utilisation (profiles/r2_k_shade.txt:
 * 3.9 GB written per launch); k_light_pre and k_light_final read them back.
 */
struct alignas(16) LightRec {
    double over[3];
    float n[3], eye[3];
    float Ka[3], Kd[3], Ks[3];
    float Ns;
    float w[3];
    int pixel;
    unsigned int rng;
};

struct Counters {
    unsigned int n_rays[8];   /* rays queued for level l */
    unsigned int n_hits[8];   /* light records of level l */
    unsigned int overflow_queue, overflow_csg;
    unsigned int n_deferred;  /* shadow rays the FP32 pass left undecided in the current k_shadow_f32 launch */
    unsigned int n_pending;   /* hits whose shadow rays are traced one by one in the current light launch */
    unsigned long long rays_secondary, rays_shadow, hits_shaded, shadow_nodes, light_flops;
    unsigned long long deferred_total, f32_mismatch, rays_gather;
    unsigned long long rays_bulk; /* shadow rays decided per hit by k_shadow_bulk */
    unsigned long long mesh_next; /* k_shadow_mesh: next (hit, sample) item to hand out */
    unsigned long long rays_per_ray; /* rays handed to the per-ray shadow kernel (k_shadow_f32 / k_shadow_mesh / k_shadow_exact<ALL>) */
    unsigned long long undecided_node[32]; /* debug: node at which a leaf verdict was undecided */
    unsigned long long undecided_reason[10]; /* counting build: why the FP32 filter deferred a ray (codes in frt_shadow_f32.cuh) */
    unsigned long long entry_node[3][32];    /* counting build: pending entries by class (0 straight-line, 1 tail verdict but general walk,
                                                2 no tail verdict) and start node */
};

struct DCamera {
    int hsize, vsize, usteps, vsteps;
    double half_width, half_height, pixel_size, canvas_distance;
    double inv[12];
    int aperture_type, jitter;
    double aperture_size;
    double aperture_args[4];
    const double *samples; /* 2*usteps*vsteps doubles: the xi = 0.5 table when jitter is off */
};

struct FrameParams {
    int rank, world, rows_per_block, n_owned_rows;
    int flags;
    int path_length;
    int include_direct, use_ambient, use_diffuse, use_spec_highlight, include_specular;
    int use_gi; /* shade_hit's global-illumination block is live (renderer.c:737): ambient terms are resolved per hit */
    unsigned long long seed;
    unsigned int capacity; /* ray queue capacity */
};

/* ------------------------------------------------------------------------------------------------ RNG */

__host__ __device__ __forceinline__ unsigned long long
mix64(unsigned long long z)
{
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

__host__ __device__ __forceinline__ unsigned int
hash32(unsigned int x)
{
    x ^= x >> 16;
    x *= 0x7feb352du;
    x ^= x >> 15;
    x *= 0x846ca68bu;
    x ^= x >> 16;
    return x;
}

__host__ __device__ __forceinline__ double
u01(unsigned long long x)
{
    return (double)(x >> 11) * (1.0 / 9007199254740992.0);
}

/* ------------------------------------------------------------------------------------------------ helpers */

/* warp-aggregated append: one atomicAdd per warp, returns this lane's slot (or 0xffffffff when want == false) */
__device__ __forceinline__ unsigned int
warp_append(unsigned int *counter, bool want)
{
    unsigned int mask = __ballot_sync(__activemask(), want);
    if (!want) {
        return 0xffffffffu;
    }
    int lane = threadIdx.x & 31;
    int leader = __ffs(mask) - 1;
    unsigned int base = 0;
    if (lane == leader) {
        base = atomicAdd(counter, (unsigned int)__popc(mask));
    }
    base = __shfl_sync(mask, base, leader);
    return base + __popc(mask & ((1u << lane) - 1u));
}

__device__ __forceinline__ int
owned_row(const FrameParams &F, int j)
{
    int b = j / F.rows_per_block;
    return (b * F.world + F.rank) * F.rows_per_block + j % F.rows_per_block;
}

/* ------------------------------------------------------------------------------------------------ raygen */

/* the aperture rejection samplers, camera.c:12-82 (drand48 replaced by the counter-based stream) */
__device__ __noinline__ void
aperture_sample(const DCamera &C, unsigned long long key, double &ax, double &ay)
{
    if (C.aperture_type == 6 /* POINT_APERTURE */ || C.aperture_type == 4 || C.aperture_type == 5 || C.aperture_type == 8) {
        ax = 0.5;
        ay = 0.5;
        return;
    }
    for (unsigned int it = 0; it < 4096; ++it) {
        double x = u01(mix64(key + 2 * it));
        double y = u01(mix64(key + 2 * it + 1));
        double u = 2 * x - 1, v = 2 * y - 1;
        bool ok;
        switch (C.aperture_type) {
        case 0: /* CIRCULAR: compares u^2+v^2 with r1, not r1^2 (camera.c:20) */
            ok = !(u * u + v * v > C.aperture_args[0]);
            break;
        case 1: /* CROSS */
            ok = ((u > C.aperture_args[0]) && (u <= C.aperture_args[1])) || ((v > C.aperture_args[2]) && (v <= C.aperture_args[3]));
            break;
        case 2: /* DIAMOND */
            ok = (u <= 0) ? ((-u + C.aperture_args[0] <= v) && (v < u + C.aperture_args[1]))
                          : ((u + C.aperture_args[2] <= v) && (v < -u + C.aperture_args[3]));
            break;
        case 3: { /* DOUGHNUT */
            double mag = u * u + v * v;
            ok = !(mag > C.aperture_args[0] || mag < C.aperture_args[1]);
            break;
        }
        default: /* SQUARE */
            ok = true;
            break;
        }
        if (ok) {
            ax = x;
            ay = y;
            return;
        }
    }
    ax = 0.5;
    ay = 0.5;
}

/*
 * One CMJ table entry for a jittered pixel: sampler_reset_canonical_2d + sampler_shuffle_2d (sampler.c:415-461)
 * replayed with the pixel's RNG stream; returns entry [u, v] like sampler_get_point_2d (:472-477).
 */
#define FRT_MAX_SPP 64
__device__ __noinline__ void
cmj_jittered(int s0, int s1, unsigned long long key, int qu, int qv, double &jx, double &jy)
{
    double arr[2 * FRT_MAX_SPP];
    unsigned int ctr = 0;
    int n = s0, m = s1; /* canonical pass binds n = steps[0], m = steps[1] (sampler.c:417-418) */
    for (int j = 0; j < n; ++j) {
        for (int i = 0; i < m; ++i) {
            int idx = 2 * (j * m + i);
            arr[idx] = (i + (j + u01(mix64(key + ctr++))) / (double)n) / (double)m;
            arr[idx + 1] = (j + (i + u01(mix64(key + ctr++))) / (double)m) / (double)n;
        }
    }
    m = s0;
    n = s1; /* shuffle binds m = steps[0], n = steps[1] (sampler.c:435-436) */
    for (int j = 0; j < n; ++j) {
        int k = (int)(j + u01(mix64(key + ctr++)) * (n - j));
        for (int i = 0; i < m; ++i) {
            double tmp = arr[2 * (j * m + i)];
            arr[2 * (j * m + i)] = arr[2 * (k * m + i)];
            arr[2 * (k * m + i)] = tmp;
        }
    }
    for (int i = 0; i < m; ++i) {
        int k = (int)(i + u01(mix64(key + ctr++)) * (m - i));
        for (int j = 0; j < n; ++j) {
            double tmp = arr[2 * (j * m + i) + 1];
            arr[2 * (j * m + i) + 1] = arr[2 * (j * m + k) + 1];
            arr[2 * (j * m + k) + 1] = tmp;
        }
    }
    jx = arr[2 * (qv * s0 + qu)];
    jy = arr[2 * (qv * s0 + qu) + 1];
}

__global__ void __launch_bounds__(256)
k_raygen(DCamera C, FrameParams F, RayQ q, Counters *cnt, unsigned int first_sample, unsigned int n_samples)
{
    const int spp = C.usteps * C.vsteps;
    const double w0 = 1.0 / (3.0 * (double)spp); /* (A + D + S) / 3 of the per-pixel mean, renderer.c:174-176, :227-230 */
    for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_samples; i += gridDim.x * blockDim.x) {
        unsigned int s = first_sample + i;
        unsigned int pix_local = s / spp;
        int sub = (int)(s - pix_local * spp);
        int j = (int)(pix_local / C.hsize);
        int px = (int)(pix_local - (unsigned int)j * C.hsize);
        int py = owned_row(F, j);
        int su = sub % C.usteps, sv = sub / C.usteps;
        unsigned int pixel = (unsigned int)py * C.hsize + px;

        double jx, jy;
        if (C.jitter) {
            cmj_jittered(C.usteps, C.vsteps, mix64(F.seed ^ (0x5bd1e995ULL * (pixel + 1))) , su, sv, jx, jy);
        } else {
            jx = __ldg(C.samples + 2 * (sv * C.usteps + su));
            jy = __ldg(C.samples + 2 * (sv * C.usteps + su) + 1);
        }
        /* ray_for_pixel, renderer.c:95-129 */
        double xoffset = ((double)px + jx) * C.pixel_size;
        double yoffset = ((double)py + jy) * C.pixel_size;
        double wx = C.half_width - xoffset;
        double wy = C.half_height - yoffset;
        double wz = -C.canvas_distance;
        double pxl[3], org[3];
        for (int k = 0; k < 3; ++k) {
            pxl[k] = C.inv[4 * k] * wx + C.inv[4 * k + 1] * wy + C.inv[4 * k + 2] * wz + C.inv[4 * k + 3];
        }
        double ax, ay;
        /* every random stream of a path is keyed on the sample's GLOBAL id (pixel, sub-sample), never on its position in
         * this rank's queue: the frame does not depend on how its rows are partitioned */
        const unsigned int gid = pixel * (unsigned int)spp + (unsigned int)sub;
        aperture_sample(C, mix64(F.seed ^ (0x7f4a7c15ULL * ((unsigned long long)gid + 1ull))), ax, ay);
        ax = (ax - 0.5) * C.aperture_size;
        ay = (ay - 0.5) * C.aperture_size;
        for (int k = 0; k < 3; ++k) {
            org[k] = C.inv[4 * k] * ax + C.inv[4 * k + 1] * ay + C.inv[4 * k + 2] * 0.0 + C.inv[4 * k + 3];
        }
        double vx = pxl[0] - org[0], vy = pxl[1] - org[1], vz = pxl[2] - org[2];
        double inv = 1.0 / sqrt(vx * vx + vy * vy + vz * vz);
        q.ox[i] = org[0];
        q.oy[i] = org[1];
        q.oz[i] = org[2];
        q.dx[i] = vx * inv;
        q.dy[i] = vy * inv;
        q.dz[i] = vz * inv;
        q.wr[i] = w0;
        q.wg[i] = w0;
        q.wb[i] = w0;
        q.pixel[i] = (int)pixel;
        q.rng[i] = hash32(gid);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        cnt->n_rays[0] = n_samples;
    }
}

/* ------------------------------------------------------------------------------------------------ extend */

#ifndef FRT_EXTEND_MINB
#define FRT_EXTEND_MINB 2
#endif
template <int PRIMS>
__global__ void __launch_bounds__(256, FRT_EXTEND_MINB)
k_extend(DScene S, DSceneF SF, RayQ q, HitQ h, Counters *cnt, int level, unsigned int capacity)
{
    const unsigned int n = min(cnt->n_rays[level], capacity); /* an overflowed level is clamped; the frame is re-run */
    int overflow = 0;
    for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        Ray r{ q.ox[i], q.oy[i], q.oz[i], q.dx[i], q.dy[i], q.dz[i] };
        Hit best = trace_closest_mixed<false, PRIMS>(S, SF, r, &overflow);
        h.t[i] = best.t;
        h.u[i] = best.u;
        h.v[i] = best.v;
        h.leaf[i] = best.leaf;
    }
    if (overflow) {
        atomicOr(&cnt->overflow_csg, 1u);
    }
}

/* ------------------------------------------------------------------------------------------------ shade */

__device__ __forceinline__ void
material_color(const DScene &S, int map, const double *flat, int leaf, const double pt[3], double out[3])
{
    if (map >= 0) {
        pattern_at_shape(S, map, leaf, pt, NULL, out, 0);
    } else {
        out[0] = flat[0];
        out[1] = flat[1];
        out[2] = flat[2];
    }
}

/*
 * MAPS: some material of the scene has a pattern / texture / bump map; REFRACT: some material lets light through (the
 * n1 / n2 containers are walked).  A scene without either (the Cornell box: flat materials, one mirror) runs an
 * instantiation without the pattern interpreter and the container walk: the general one is 14 K SASS instructions with a
 * 2 KB local frame and stalls on it (profiles/r2_k_shade.txt: 15 % of its instructions are local loads / stores,
 * long_scoreboard 14 warps per issue).
 */
template <bool MAPS, bool REFRACT>
__global__ void __launch_bounds__(128, (MAPS || REFRACT) ? 4 : 5)
k_shade(DScene S, FrameParams F, RayQ q, HitQ h, RayQ qn, LightRec *recs, Counters *cnt, int level, int rec_level)
{
    static_assert(sizeof(LightRec) % 16 == 0, "LightRec is copied as 16-byte words");
    constexpr int FRT_REC_WORDS = (int)(sizeof(LightRec) / 16);
    __shared__ uint4 s_stage[4][32 * FRT_REC_WORDS];
    const unsigned int n = min(cnt->n_rays[level], F.capacity);
    const int remaining = F.path_length - level;
    int overflow = 0;
    unsigned long long n_secondary = 0, n_shaded = 0;

    for (unsigned int i0 = blockIdx.x * blockDim.x; i0 < n; i0 += gridDim.x * blockDim.x) {
        unsigned int i = i0 + threadIdx.x;
        bool live = i < n;
        int leaf = live ? h.leaf[i] : -1;
        live = live && leaf >= 0;

        bool want_rec = false, want_refl = false, want_refr = false;
        LightRec rec; /* built in registers / local memory (building it in shared memory measured slower: 1.75 vs 1.46 ms) */
        Ray rfl{}, rfr{};
        double w_refl[3] = { 0, 0, 0 }, w_refr[3] = { 0, 0, 0 };
        int pixel = 0;
        unsigned int rng = 0;

        if (live) {
            /* ---- prepare_computations, renderer.c:368-495 */
            Ray r{ q.ox[i], q.oy[i], q.oz[i], q.dx[i], q.dy[i], q.dz[i] };
            double w[3] = { q.wr[i], q.wg[i], q.wb[i] };
            pixel = q.pixel[i];
            rng = q.rng[i];
            double t = h.t[i];
            NodeA a = load_node_a(S, leaf);
            NodeB b = load_node_b(S, leaf);
            const frt_material &M = S.mats[a.material];
            const double *prm = S.params + (b.param < 0 ? 0 : b.param);

            double p[3] = { r.ox + r.dx * t, r.oy + r.dy * t, r.oz + r.dz * t };
            double lp[3], ln[3], nrm[3];
            point_to_local(S, a.xform, p, lp);
            local_normal(a.type, prm, lp, h.u[i], h.v[i], ln);
            normal_to_world(S, a.xform, ln, nrm);
            if (MAPS && M.map_bump >= 0) { /* shape_normal_at, shapes.c:76-86: n += 2*texel - 1, sampled at the hit point */
                double tex[3];
                pattern_at_shape(S, M.map_bump, leaf, p, NULL, tex, 0);
                nrm[0] += 2.0 * tex[0] - 1.0;
                nrm[1] += 2.0 * tex[1] - 1.0;
                nrm[2] += 2.0 * tex[2] - 1.0;
            }
            {
                double inv = 1.0 / sqrt(nrm[0] * nrm[0] + nrm[1] * nrm[1] + nrm[2] * nrm[2]);
                nrm[0] *= inv;
                nrm[1] *= inv;
                nrm[2] *= inv;
            }
            double eye[3] = { -r.dx, -r.dy, -r.dz };
            if (nrm[0] * eye[0] + nrm[1] * eye[1] + nrm[2] * eye[2] < 0) { /* inside */
                nrm[0] = -nrm[0];
                nrm[1] = -nrm[1];
                nrm[2] = -nrm[2];
            }
            double ddot = 2 * (r.dx * nrm[0] + r.dy * nrm[1] + r.dz * nrm[2]);
            double reflv[3] = { r.dx - nrm[0] * ddot, r.dy - nrm[1] * ddot, r.dz - nrm[2] * ddot };
            double over[3] = { p[0] + nrm[0] * FRT_EPS, p[1] + nrm[1] * FRT_EPS, p[2] + nrm[2] * FRT_EPS };
            double under[3] = { p[0] - nrm[0] * FRT_EPS, p[1] - nrm[1] * FRT_EPS, p[2] - nrm[2] * FRT_EPS };

            double Ka[3], Kd[3], Ks[3], refl[3];
            material_color(S, MAPS ? M.map_Ka : -1, M.Ka, leaf, over, Ka);
            material_color(S, MAPS ? M.map_Kd : -1, M.Kd, leaf, over, Kd);
            material_color(S, MAPS ? M.map_Ks : -1, M.Ks, leaf, over, Ks);
            material_color(S, MAPS ? M.map_refl : -1, M.refl, leaf, over, refl);
            double Ns = M.Ns;
            if (MAPS && M.map_Ns >= 0) {
                double tmp[3];
                pattern_at_shape(S, M.map_Ns, leaf, over, NULL, tmp, 0);
                Ns = tmp[0];
            }
            double over_d = 1.0 - M.Tr;
            if (MAPS && M.map_d >= 0) {
                double tmp[3];
                pattern_at_shape(S, M.map_d, leaf, over, NULL, tmp, 0);
                over_d = tmp[0];
            }

            /* ---- the specular split of shade_hit, renderer.c:772-822, as throughput weights */
            const bool no_prune = (F.flags & FRT_FLAG_NO_PRUNE) != 0;
            const bool spawn = F.include_specular && remaining > 0;
            bool do_refl = spawn && M.reflective;                       /* reflected_color :497-505 */
            bool do_refr = spawn && !(over_d <= 0.0);                   /* refracted_color :534-542 */
            bool tf_zero = M.Tf[0] == 0.0 && M.Tf[1] == 0.0 && M.Tf[2] == 0.0;
            bool refl_zero = refl[0] == 0.0 && refl[1] == 0.0 && refl[2] == 0.0;
            if (!no_prune) {
                /* a branch whose weight is exactly zero cannot change the pixel */
                do_refr = do_refr && !tf_zero;
                do_refl = do_refl && !refl_zero;
            }
            bool use_schlick = M.reflective && over_d < 1.0;            /* :788 */
            double n1 = 1.0, n2 = 1.0;
            if (REFRACT && (do_refr || (use_schlick && (do_refl || do_refr)))) {
                trace_containers(S, r, leaf, n1, n2, &overflow);
            }
            double cos_i = eye[0] * nrm[0] + eye[1] * nrm[1] + eye[2] * nrm[2];
            double n_ratio = n1 / n2;
            double sin2_t = n_ratio * n_ratio * (1.0 - cos_i * cos_i);
            if (do_refr && sin2_t > 1.0) { /* total internal reflection, :548-553 */
                do_refr = false;
            }
            double k_refl = 1.0, k_refr = 1.0;
            if (use_schlick) { /* schlick, :607-624 */
                double co = cos_i;
                double reflectance;
                bool tir = false;
                if (n1 > n2) {
                    double nn = n1 / n2;
                    double s2 = nn * nn * (1.0 - co * co);
                    if (s2 > 1.0) {
                        tir = true;
                    }
                    co = sqrt(1.0 - s2);
                }
                if (tir) {
                    reflectance = 1.0;
                } else {
                    double r0 = (n1 - n2) / (n1 + n2);
                    r0 = r0 * r0;
                    double m1 = 1.0 - co;
                    reflectance = r0 + (1.0 - r0) * m1 * m1 * m1 * m1 * m1;
                }
                k_refl = reflectance;
                k_refr = 1.0 - reflectance;
            }
            double dissolve = (M.Tr > 0.0 && over_d > 0.0) ? (1.0 - over_d) : 1.0; /* :804-817 */

            if (do_refl) {
                for (int k = 0; k < 3; ++k) {
                    w_refl[k] = w[k] * dissolve * k_refl * refl[k];
                }
                rfl = Ray{ over[0], over[1], over[2], reflv[0], reflv[1], reflv[2] };
                want_refl = no_prune || w_refl[0] != 0.0 || w_refl[1] != 0.0 || w_refl[2] != 0.0;
            }
            if (do_refr) {
                double cos_t = sqrt(1.0 - sin2_t);
                double s = n_ratio * cos_i - cos_t;
                rfr = Ray{ under[0], under[1], under[2],
                           nrm[0] * s - eye[0] * n_ratio, nrm[1] * s - eye[1] * n_ratio, nrm[2] * s - eye[2] * n_ratio };
                for (int k = 0; k < 3; ++k) {
                    w_refr[k] = w[k] * k_refr * M.Tf[k] * over_d;
                }
                want_refr = no_prune || w_refr[0] != 0.0 || w_refr[1] != 0.0 || w_refr[2] != 0.0;
            }

            /* ---- record for the light stage (direct illumination of this hit, weight = throughput * dissolve) */
            if ((F.include_direct && S.n_lights > 0) || F.use_gi) {
                want_rec = true;
                for (int k = 0; k < 3; ++k) {
                    rec.over[k] = over[k];
                    rec.n[k] = nrm[k];
                    rec.eye[k] = eye[k];
                    rec.Ka[k] = Ka[k];
                    rec.Kd[k] = Kd[k];
                    rec.Ks[k] = Ks[k];
                    rec.w[k] = w[k] * dissolve;
                }
                rec.Ns = Ns;
                rec.pixel = pixel;
                rec.rng = rng;
                if (!no_prune && rec.w[0] == 0.0 && rec.w[1] == 0.0 && rec.w[2] == 0.0) {
                    want_rec = false;
                }
            }
            ++n_shaded;
        }

        /* The warp's records take consecutive slots.  Written record by record, a warp's 19 stores per record each hit 32
         * different sectors, none of them whole (152-byte records): the L2 has to fetch every sector it is about to
         * overwrite.  Staged through shared memory the same bytes leave as 256-byte runs. */
        __syncwarp();
        unsigned int slot = warp_append(&cnt->n_hits[rec_level], want_rec); /* rec_level = 0 for every level: one list of hits, one light stage */
        {
            const unsigned int wmask = __ballot_sync(0xffffffffu, want_rec);
            if (wmask) {
                const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
                const int leader = __ffs(wmask) - 1;
                const unsigned int base = __shfl_sync(0xffffffffu, slot, leader);
                const unsigned int count = (unsigned int)__popc(wmask);
                constexpr int W = FRT_REC_WORDS;
                if (want_rec) {
                    const uint4 *src = reinterpret_cast<const uint4 *>(&rec);
                    uint4 *dst = s_stage[wib] + (slot - base) * W; /* rank among the warp's records */
#pragma unroll
                    for (int k = 0; k < W; ++k) {
                        dst[k] = src[k];
                    }
                }
                __syncwarp();
                if (base + count <= F.capacity) {
                    uint4 *g = reinterpret_cast<uint4 *>(recs + base);
                    for (unsigned int j = lane; j < count * W; j += 32) {
                        g[j] = s_stage[wib][j];
                    }
                } else {
                    if (want_rec && slot < F.capacity) {
                        recs[slot] = rec;
                    }
                    if (lane == leader) {
                        atomicOr(&cnt->overflow_queue, 1u);
                    }
                }
                __syncwarp();
            }
        }
        unsigned int s1 = warp_append(&cnt->n_rays[level + 1], want_refl);
        if (want_refl) {
            if (s1 < F.capacity) {
                qn.ox[s1] = rfl.ox; qn.oy[s1] = rfl.oy; qn.oz[s1] = rfl.oz;
                qn.dx[s1] = rfl.dx; qn.dy[s1] = rfl.dy; qn.dz[s1] = rfl.dz;
                qn.wr[s1] = w_refl[0]; qn.wg[s1] = w_refl[1]; qn.wb[s1] = w_refl[2];
                qn.pixel[s1] = pixel;
                qn.rng[s1] = hash32(rng + 0x9e3779b9u); /* child ids are hashed: no arithmetic relation between paths */
                ++n_secondary;
            } else {
                atomicOr(&cnt->overflow_queue, 1u);
            }
        }
        unsigned int s2 = warp_append(&cnt->n_rays[level + 1], want_refr);
        if (want_refr) {
            if (s2 < F.capacity) {
                qn.ox[s2] = rfr.ox; qn.oy[s2] = rfr.oy; qn.oz[s2] = rfr.oz;
                qn.dx[s2] = rfr.dx; qn.dy[s2] = rfr.dy; qn.dz[s2] = rfr.dz;
                qn.wr[s2] = w_refr[0]; qn.wg[s2] = w_refr[1]; qn.wb[s2] = w_refr[2];
                qn.pixel[s2] = pixel;
                qn.rng[s2] = hash32(rng + 0x3c6ef372u);
                ++n_secondary;
            } else {
                atomicOr(&cnt->overflow_queue, 1u);
            }
        }
    }
    if (overflow) {
        atomicOr(&cnt->overflow_csg, 1u);
    }
    /* block-level reduction of the counters: one atomic per warp */
    for (int o = 16; o > 0; o >>= 1) {
        n_secondary += __shfl_down_sync(0xffffffffu, n_secondary, o);
        n_shaded += __shfl_down_sync(0xffffffffu, n_shaded, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (n_secondary) atomicAdd(&cnt->rays_secondary, n_secondary);
        if (n_shaded) atomicAdd(&cnt->hits_shaded, n_shaded);
    }
}

/* ------------------------------------------------------------------------------------------------ light */

/*
 * Direct illumination of the shaded hits of one level by one light, in three kernels that each stay small enough
 * to live in the instruction cache (the first, fused version of this stage was 107 KB of SASS and spent half of
 * its issue slots waiting on instruction fetch; see profiles/):
 *
 *   k_light_pre      sample-set picks, shaft-culling mask, can-the-light-contribute test per hit
 *   k_shadow         light->intensity_at: one shadow ray per (hit, surface sample)   light.c:229-251, renderer.c:73
 *   k_light_final    lighting_microfacet for the hits that received light, weighting into the pixel  renderer.c:894-979
 *
 * The reference picks one of cache_len pre-computed sample sets with rand() once per pass (light.c:196); here the
 * pick is a hash of (seed, path id, light, pass).  With cache_len == 1 both passes use set 0 and the result is
 * exactly the reference's.
 */
struct LightTmp { /* per shaded hit, per light launch (32 bytes); the first 16 bytes are what k_shadow_f32 reads per ray */
    float ox, oy, oz;      /* over_point rounded to FP32 (origin of the hit's shadow rays in the FP32 filter) */
    int set_a;             /* sample set of the shadow pass; -1: the hit cannot receive light, skip its shadow rays */
    unsigned int relevant; /* shaft culling: bit i = node i may be crossed at t > 0 by a shadow ray of this hit */
    int set_b;             /* sample set of the lighting pass */
    int unshadowed;        /* shadow rays that reached the light */
    int contributes;       /* the light is not wholly behind the surface: the lighting sums can be non-zero */
};

/*
 * G lanes (a power of two <= 32) share a hit and stride over the samples; partial sums are combined with a
 * fixed-order butterfly, so the result does not depend on scheduling.  T = float is the production path: the sums
 * feed an 8-bit pixel through a continuous function, so FP32's 1e-7 relative error sits three orders of magnitude
 * under one sRGB LSB; T = double (FRT_FLAG_F64_SHADING) keeps the reference's arithmetic type for debugging.
 * Every geometric DECISION (hit / miss, shadowed / lit) stays in FP64 in k_extend / k_shadow.
 */
/* FP32 lighting arithmetic uses the SFU forms (relative error ~1e-6, far below one sRGB-8 LSB after the 100-sample
 * average); the FP64 instantiation keeps the IEEE forms */
__device__ __forceinline__ float sh_rsqrt(float x) { return rsqrtf(x); }
__device__ __forceinline__ double sh_rsqrt(double x) { return 1.0 / sqrt(x); }
__device__ __forceinline__ float sh_rcp(float x) { return __frcp_rn(x); }
__device__ __forceinline__ double sh_rcp(double x) { return 1.0 / x; }
__device__ __forceinline__ float sh_pow(float x, float y) { return x > 0.f ? exp2f(y * __log2f(x)) : (y == 0.f ? 1.f : 0.f); }
__device__ __forceinline__ double sh_pow(double x, double y) { return pow(x, y); }

/*
 * Per hit, before its shadow rays: the two sample-set picks, the shaft-culling mask (frt_shadow_f32.cuh; the G lanes of
 * a hit share the nodes) and a cheap test whether the light can contribute at all -- if every corner of the light's
 * parallelogram is behind the surface, every sample has N.L < 0, lighting_microfacet adds nothing (renderer.c:927-968)
 * and the visibility fraction cannot matter: no shadow rays.
 */
template <int G>
__global__ void __launch_bounds__(256)
k_light_pre(DScene S, FrameParams F, const LightRec *__restrict__ recs, LightTmp *__restrict__ tmp, const Counters *cnt, int level,
            int light_idx, DSceneF SF, int shaft_on)
{
    const unsigned int n = min(cnt->n_hits[level], F.capacity);
    const int cache_len = S.lights[light_idx].cache_len;
    const unsigned int lane_g = threadIdx.x % G;
    const unsigned int gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << ((threadIdx.x & 31) / G * G));
    const unsigned int groups = gridDim.x * blockDim.x / G;

    for (unsigned int hbase = (blockIdx.x * blockDim.x + threadIdx.x) / G;; hbase += groups) {
        unsigned int h0 = __shfl_sync(0xffffffffu, hbase, 0);
        if (h0 >= n) {
            break;
        }
        const bool live = hbase < n;
        const LightRec *R = recs + (live ? hbase : 0);
        const float ofx = (float)R->over[0], ofy = (float)R->over[1], ofz = (float)R->over[2];
        /* can the light contribute at all?  (every lane of the group evaluates it: the answer gates the shaft mask) */
        bool contributes = false;
        if (live && (F.use_diffuse || F.use_spec_highlight)) {
            for (int c = 0; c < 4; ++c) {
                const float4 p = __ldg(SF.shaft + 4 * light_idx + c);
                const double ndl = R->n[0] * ((double)p.x - R->over[0]) + R->n[1] * ((double)p.y - R->over[1]) +
                                   R->n[2] * ((double)p.z - R->over[2]);
                contributes = contributes || !(ndl < 0.0);
            }
        }
        const bool wants_rays = live && (contributes || (F.flags & FRT_FLAG_NO_PRUNE));
        unsigned int relevant = 0xffffffffu;
        if (shaft_on && __any_sync(gmask, wants_rays)) {
            relevant = 0u;
            if (wants_rays) { /* a hit without shadow rays needs no mask (38 % of the Cornell frame's hits) */
                ShaftF sh;
                shaft_setup(sh, SF.shaft + 4 * light_idx, ofx, ofy, ofz);
                const int nn = min(SF.n_nodes, 32);
                for (int k = lane_g; k < nn; k += G) {
                    bool miss = shaft_misses_box(sh, __ldg(SF.wbox + 2 * k), __ldg(SF.wbox + 2 * k + 1), ofx, ofy, ofz);
                    if (!miss) { /* a ball fills half of its box: test the ball itself */
                        const float4 sp = __ldg(SF.wsphere + k);
                        miss = sp.w > 0.f && shaft_misses_sphere(sh, sp, ofx, ofy, ofz);
                    }
                    if (!miss) {
                        relevant |= 1u << k;
                    }
                }
            }
            for (int o = G / 2; o > 0; o >>= 1) {
                relevant |= __shfl_xor_sync(gmask, relevant, o);
            }
            const int nn = min(SF.n_nodes, 32);
            if (nn < 32) {
                relevant |= ~((1u << nn) - 1u);
            }
        }
        if (live && lane_g == 0) {
            int set_a = 0, set_b = 0;
            if (cache_len > 1) {
                unsigned long long key = F.seed ^ ((unsigned long long)R->rng << 20) ^ ((unsigned long long)light_idx << 4);
                set_a = (int)(mix64(key) % (unsigned long long)cache_len);
                set_b = (int)(mix64(key + 1) % (unsigned long long)cache_len);
            }
            LightTmp t;
            t.ox = ofx;
            t.oy = ofy;
            t.oz = ofz;
            t.set_a = (contributes || (F.flags & FRT_FLAG_NO_PRUNE)) ? set_a : -1;
            t.relevant = relevant;
            t.set_b = set_b;
            t.unshadowed = 0;
            t.contributes = contributes ? 1 : 0;
            tmp[hbase] = t;
        }
    }
}

/*
 * Sample index of item k of quadrant q of the light's sample grid (the cache stores sample (u, v) at v * usteps + u,
 * light.c:172-186): quadrant q covers u in [(q & 1) hu, ...+hu), v in [(q >> 1) hv, ...+hv).  hu == 0: no split.
 */
__device__ __forceinline__ int
quadrant_sample(int4 lq, int q, int k)
{
    if (lq.y == 0) {
        return k;
    }
    const int lv = (int)(((float)k + 0.5f) * __frcp_rn((float)lq.y)); /* k / hu, exact for the few hundred samples of a light */
    const int lu = k - lv * lq.y;
    return ((q >> 1) * lq.z + lv) * lq.w + (q & 1) * lq.y + lu;
}

/* pending entry: hit in bits 0..27, quadrant in bits 28..29, bit 30 = decided by k_shadow_bulk, bit 31 = ... as lit */
#define FRT_BOX_CHUNKS 128
#define FRT_QS_MAX 256 /* samples per pending entry up to which k_shadow_f32 tabulates quadrant_sample in shared memory */
#define FRT_PEND_HIT_MASK 0x0fffffffu
#define FRT_PEND_BULK 0x40000000u
#define FRT_PEND_LIT 0x80000000u

/*
 * Per hit, after k_light_pre: try to decide ALL shadow rays of the hit at once (trace_shadow_bulk, frt_shadow_f32.cuh);
 * when the whole light is undecided (a penumbra hit) and the light's sample grid splits into quadrants, try each
 * quadrant of samples on its own -- an occluder's edge seldom crosses all four.  What is left is compacted into
 * `pending`, one entry per (hit, quadrant) whose rays the per-ray kernels must trace.  MODE as in k_shadow_f32: in
 * counting (1) and verifying (2) frames the decided entries stay in the list, flagged, so that k_shadow_f32 can count /
 * re-trace their rays; a verifying frame takes its visibility counts from those FP64 re-traces, like it does for the
 * per-ray filter.
 */
#ifndef FRT_SHAFT_MINB
#define FRT_SHAFT_MINB 4 /* blocks per SM the shaft kernels are compiled for (64 registers, 90 bytes of spills): measured 1.69 / 1.43 / 1.27 ms for bulk + quad at 2 / 3 / 4 (93 / 79 / 64 registers) */
#endif
template <int MODE, typename T>
__global__ void __launch_bounds__(256, FRT_SHAFT_MINB)
k_shadow_bulk(DScene S, DSceneF SF, FrameParams F, const LightRec *__restrict__ recs, LightTmp *__restrict__ tmp, Counters *cnt, int level,
              int light_idx, unsigned int *__restrict__ pending, unsigned int *__restrict__ pprog, unsigned int *__restrict__ retry,
              int bulk_on, int split_on)
{
    const unsigned int n = min(cnt->n_hits[level], F.capacity);
    const int4 lq = split_on ? __ldg(SF.lquad + light_idx) : make_int4(S.lights[light_idx].num_samples, 0, 0, 0);
    const int nq = lq.y ? 4 : 1;
    const int root = __ldg(S.roots);
    unsigned long long n_bulk = 0;
    for (unsigned int base = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < n; base += gridDim.x * blockDim.x) {
        const unsigned int h = base + (threadIdx.x & 31);
        /* 0 = nothing to do, 1 = undecided, 2 / 3 = decided shadowed / lit */
        int state = 0;
        unsigned int prog = FRT_PROG_ROOT(root);
        LightTmp t;
        t.set_a = -1;
        if (h < n) {
            t = tmp[h];
        }
        if (t.set_a >= 0) {
            state = 1;
            if (bulk_on) {
                ShaftT<T> sh;
                const T over[3] = { (T)t.ox, (T)t.oy, (T)t.oz }; /* FP32 over-point: within 2^-24 |o| of the FP64 one */
                shaft_d_setup(sh, SF.lbox + 30 * light_idx, over, (T)1.2e-7, SF.bmax, SF.smin, SF.ealign);
                const int res = trace_shadow_bulk(SF, root, t.relevant, sh, &prog);
                if (res != FRT_SH_UNDECIDED) {
                    state = res == FRT_SH_LIT ? 3 : 2;
                    n_bulk += (unsigned int)(nq * lq.x);
                    if (MODE != 2 && res == FRT_SH_LIT) {
                        tmp[h].unshadowed = nq * lq.x; /* k_light_pre left 0 */
                    }
                }
            }
        }
        /* undecided and the light's sample grid splits: k_shadow_quad tries the quadrants one by one */
        const unsigned int r = warp_append(&cnt->n_deferred, state == 1 && nq > 1);
        if (state == 1 && nq > 1) {
            retry[r] = h;
        }
        const bool keep = (state == 1 && nq == 1) || (MODE != 0 && state >= 2);
        for (int q = 0; q < nq; ++q) {
            const unsigned int slot = warp_append(&cnt->n_pending, keep);
            if (keep) {
                pending[slot] = h | ((unsigned int)q << 28) | (state >= 2 ? FRT_PEND_BULK : 0u) | (state == 3 ? FRT_PEND_LIT : 0u);
                pprog[slot] = state == 1 ? (prog | (entry_program_is_fast(prog, SF.entry_fast) ? FRT_PROG_FAST : 0u)) : FRT_PROG_ROOT(root);
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        n_bulk += __shfl_down_sync(0xffffffffu, n_bulk, o);
    }
    if ((threadIdx.x & 31) == 0 && n_bulk) {
        atomicAdd(&cnt->rays_bulk, n_bulk);
    }
}

/* one thread per (undecided hit, quadrant of the light's sample grid): the same walk against the quadrant's bounds */
template <int MODE, typename T>
__global__ void __launch_bounds__(256, FRT_SHAFT_MINB)
k_shadow_quad(DScene S, DSceneF SF, FrameParams F, const LightRec *__restrict__ recs, LightTmp *__restrict__ tmp, Counters *cnt, int light_idx,
              unsigned int *__restrict__ pending, unsigned int *__restrict__ pprog, const unsigned int *__restrict__ retry)
{
    const unsigned int n = min(cnt->n_deferred, F.capacity);
    const int4 lq = __ldg(SF.lquad + light_idx);
    const int root = __ldg(S.roots);
    unsigned long long n_bulk = 0;
    const unsigned int total = 4u * n; /* n <= 2^28 */
    for (unsigned int base = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < total; base += gridDim.x * blockDim.x) {
        const unsigned int item = base + (threadIdx.x & 31);
        int state = 0;
        unsigned int prog = FRT_PROG_ROOT(root);
        unsigned int h = 0;
        const int q = (int)(item & 3u);
        if (item < total) {
            h = __ldg(retry + (item >> 2));
            ShaftT<T> sh;
            const LightTmp t = tmp[h];
            const T over[3] = { (T)t.ox, (T)t.oy, (T)t.oz };
            shaft_d_setup(sh, SF.lbox + 30 * light_idx + 6 * (q + 1), over, (T)1.2e-7, SF.bmax, SF.smin, SF.ealign);
            const int res = trace_shadow_bulk(SF, root, t.relevant, sh, &prog);
            state = res == FRT_SH_UNDECIDED ? 1 : (res == FRT_SH_LIT ? 3 : 2);
            if (state >= 2) {
                n_bulk += (unsigned int)lq.x;
                if (MODE != 2 && state == 3) {
                    atomicAdd(&tmp[h].unshadowed, lq.x);
                }
            }
        }
        const bool keep = state == 1 || (MODE != 0 && state >= 2);
        const unsigned int slot = warp_append(&cnt->n_pending, keep);
        if (keep) {
            pending[slot] = h | ((unsigned int)q << 28) | (state >= 2 ? FRT_PEND_BULK : 0u) | (state == 3 ? FRT_PEND_LIT : 0u);
            pprog[slot] = state == 1 ? (prog | (entry_program_is_fast(prog, SF.entry_fast) ? FRT_PROG_FAST : 0u)) : FRT_PROG_ROOT(root);
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        n_bulk += __shfl_down_sync(0xffffffffu, n_bulk, o);
    }
    if ((threadIdx.x & 31) == 0 && n_bulk) {
        atomicAdd(&cnt->rays_bulk, n_bulk);
    }
}

/* axis-aligned bounds of the light points of quadrant blockIdx.y over the sample sets blockIdx.x, +gridDim.x, ... */
__global__ void __launch_bounds__(256)
k_light_boxes(const double *__restrict__ pts, int cache_len, int NS, int4 lq, double *__restrict__ partial, int chunk_stride)
{
    const int q = blockIdx.y;
    double lo[3] = { CUDART_INF, CUDART_INF, CUDART_INF }, hi[3] = { -CUDART_INF, -CUDART_INF, -CUDART_INF };
    /* the block's sets are blockIdx.x, + gridDim.x, ...; its threads share the (set, sample) pairs -- a quadrant holds 25 samples
     * of a 10 x 10 light, so a thread per sample of ONE set at a time left nine tenths of the block idle (0.34 -> 0.04 ms for the
     * 65 535 sets of the Cornell light: the frame's first light stage waits for these bounds) */
    const int sets_mine = blockIdx.x < cache_len ? (cache_len - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    for (int t = threadIdx.x; t < sets_mine * lq.x; t += blockDim.x) {
        const int j = t / lq.x, k = t - j * lq.x;
        const int set = blockIdx.x + j * gridDim.x;
        const double *p = pts + 3 * ((size_t)set * NS + quadrant_sample(lq, q, k));
        for (int c = 0; c < 3; ++c) {
            const double v = __ldg(p + c);
            lo[c] = fmin(lo[c], v);
            hi[c] = fmax(hi[c], v);
        }
    }
    __shared__ double sm[8][6];
    for (int c = 0; c < 3; ++c) {
        for (int o = 16; o > 0; o >>= 1) {
            lo[c] = fmin(lo[c], __shfl_down_sync(0xffffffffu, lo[c], o));
            hi[c] = fmax(hi[c], __shfl_down_sync(0xffffffffu, hi[c], o));
        }
    }
    if ((threadIdx.x & 31) == 0) {
        for (int c = 0; c < 3; ++c) {
            sm[threadIdx.x >> 5][c] = lo[c];
            sm[threadIdx.x >> 5][3 + c] = hi[c];
        }
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        double v = sm[0][threadIdx.x];
        for (int w = 1; w < 8; ++w) {
            v = threadIdx.x < 3 ? fmin(v, sm[w][threadIdx.x]) : fmax(v, sm[w][threadIdx.x]);
        }
        partial[((size_t)q * chunk_stride + blockIdx.x) * 6 + threadIdx.x] = v;
    }
}

/* second stage of k_light_boxes, one block per light: reduce the chunks, inflate, write {all, quadrant 0..3} */
__global__ void
k_light_boxes_finish(const double *__restrict__ partial, const int4 *__restrict__ lquad, const int *__restrict__ chunks, double *__restrict__ lbox)
{
    const int li = blockIdx.x, q = threadIdx.x;
    const int nq = lquad[li].y ? 4 : 1, nb = chunks[li];
    __shared__ double box[4][6];
    if (q < 4) {
        double b[6] = { CUDART_INF, CUDART_INF, CUDART_INF, -CUDART_INF, -CUDART_INF, -CUDART_INF };
        if (q < nq) {
            const double *pp = partial + ((size_t)li * 4 + q) * FRT_BOX_CHUNKS * 6;
            for (int c = 0; c < nb; ++c) {
                for (int k = 0; k < 3; ++k) {
                    b[k] = fmin(b[k], pp[6 * c + k]);
                    b[3 + k] = fmax(b[3 + k], pp[6 * c + 3 + k]);
                }
            }
            for (int k = 0; k < 3; ++k) { /* inflate: the points are exact, the slack covers nothing but habit */
                const double m = 1e-12 * (fmax(fabs(b[k]), fabs(b[3 + k])) + 1.0);
                b[k] -= m;
                b[3 + k] += m;
            }
        }
        for (int k = 0; k < 6; ++k) {
            box[q][k] = b[k];
        }
    }
    __syncthreads();
    if (q < 4) {
        double *out = lbox + (size_t)30 * li;
        for (int k = 0; k < 6; ++k) {
            double all = box[0][k];
            for (int j = 1; j < nq; ++j) {
                all = k < 3 ? fmin(all, box[j][k]) : fmax(all, box[j][k]);
            }
            if (q == 0) {
                out[k] = all;
            }
            out[6 * (q + 1) + k] = q < nq ? box[q][k] : all;
        }
    }
}

/*
 * One thread per (pending entry, sample): k_shadow_bulk / k_shadow_quad leave a list of (hit, quadrant of the light's
 * sample grid) entries whose rays still have to be traced; item = entry * samples-per-entry + k, so a warp holds
 * consecutive samples of one hit (or the tail of one entry and the head of the next) -- rays with a common origin and
 * nearly parallel directions, which walk the tree together.  Each entry also says at which node its rays start and,
 * when the shaft walk could tell, what becomes of the rays that node does not stop (trace_shadow_bulk).  Deferred
 * items keep the absolute encoding hit * num_samples + sample.  The unshadowed count of a hit is reduced inside the warp
 * (__match_any_sync on the hit index) and added with one integer atomic per (warp, hit): integer sums are
 * order-independent, so the frame is reproducible.
 *
 * k_shadow_f32 answers every ray the FP32 filtered traversal can decide (frt_shadow_f32.cuh) and appends the rest
 * to `queue` (one atomic per warp); k_shadow_exact then re-traces the queue in FP64.  MODE: 0 production,
 * 1 counting (FRT_FLAG_COUNT_RAYS), 2 verifying (FRT_FLAG_VERIFY_F32: every decided ray is also traced in FP64 and
 * disagreements are counted in Counters::f32_mismatch).
 */
__device__ __forceinline__ void
shadow_item(const DScene &S, const LightRec *__restrict__ recs, const LightTmp *__restrict__ tmp, const double *pts, int NS,
            unsigned long long item, bool small, unsigned int &h, int &set_a, Ray &sr, double &dist2)
{
    h = small ? (unsigned int)item / (unsigned int)NS : (unsigned int)(item / (unsigned int)NS);
    const int s = (int)(item - (unsigned long long)h * (unsigned int)NS);
    set_a = tmp[h].set_a;
    if (set_a >= 0) {
        const LightRec *R = recs + h;
        sr.ox = R->over[0];
        sr.oy = R->over[1];
        sr.oz = R->over[2];
        const double *pa = pts + 3 * ((size_t)set_a * NS + s);
        sr.dx = __ldg(pa) - sr.ox;
        sr.dy = __ldg(pa + 1) - sr.oy;
        sr.dz = __ldg(pa + 2) - sr.oz; /* not normalised yet */
        dist2 = sr.dx * sr.dx + sr.dy * sr.dy + sr.dz * sr.dz;
    }
}

__device__ __forceinline__ double
normalise_shadow_ray(Ray &sr, double dist2)
{
    const double inv = rsqrt_fast(dist2);
    sr.dx *= inv;
    sr.dy *= inv;
    sr.dz *= inv;
    return dist2 * inv;
}

/* the general walk as a call: the per-ray kernels' registers are sized for the straight-line path */
template <bool COUNT>
__device__ __noinline__ int
trace_shadow_f32_call(const DSceneF &SF, const float4 *fnodes, int root, int start, int tail, unsigned int relevant, const FrameF &w, float omax,
                      float eo_o, float ed_w, float D_lo, float D_hi, unsigned long long *nodes_visited, unsigned long long *flops)
{
    return trace_shadow_f32<COUNT>(SF, fnodes, root, start, tail, relevant, w, omax, eo_o, ed_w, D_lo, D_hi, nodes_visited, flops);
}

#ifndef FRT_SHADOW_MINB
#define FRT_SHADOW_MINB 3
#endif
template <int MODE>
__global__ void __launch_bounds__(256, FRT_SHADOW_MINB)
k_shadow_f32(DScene S, DSceneF SF, FrameParams F, const LightRec *__restrict__ recs, LightTmp *__restrict__ tmp, Counters *cnt,
             const unsigned int *__restrict__ pending, const unsigned int *__restrict__ pprog, int use_start, unsigned int pend_cap,
             int split_on, int light_idx,
             unsigned long long *__restrict__ queue, unsigned int qcap, int nodes_in_smem)
{
    constexpr bool COUNT = MODE != 0;
    extern __shared__ float4 s_nodes[];
    __shared__ unsigned short s_qs[4 * FRT_QS_MAX]; /* sample index of (quadrant, k): quadrant_sample, tabulated per block */
    const float4 *fnodes = SF.fnodes;
    const int *cprog = SF.csg_prog;
    const int NS = S.lights[light_idx].num_samples;
    const int4 lq = split_on ? __ldg(SF.lquad + light_idx) : make_int4(NS, 0, 0, 0);
    const int NSQ = lq.x; /* samples per entry */
    const bool qs_table = NSQ <= FRT_QS_MAX;
    if (nodes_in_smem) { /* small trees (every scene but the OBJ meshes) are walked out of shared memory, CSG programs included */
        for (int k = threadIdx.x; k < 3 * SF.n_nodes; k += blockDim.x) {
            s_nodes[k] = SF.fnodes[k];
        }
        int *s_prog = reinterpret_cast<int *>(s_nodes + 3 * SF.n_nodes);
        for (int k = threadIdx.x; k < SF.n_csg_prog; k += blockDim.x) {
            s_prog[k] = SF.csg_prog[k];
        }
        fnodes = s_nodes;
        cprog = s_prog;
    }
    if (qs_table) {
        for (int k = threadIdx.x; k < 4 * NSQ; k += blockDim.x) {
            s_qs[k] = (unsigned short)quadrant_sample(lq, k / NSQ, k % NSQ);
        }
    }
    __syncthreads();
    const unsigned int n = min(cnt->n_pending, pend_cap); /* (hit, quadrant) entries k_shadow_bulk left to be traced ray by ray */
    const float *fpts = SF.lpoints + 3 * S.lights[light_idx].point_offset;
    const double *pts = S.lpoints + 3 * S.lights[light_idx].point_offset;
    const unsigned long long total = (unsigned long long)n * (unsigned int)NSQ;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    const unsigned long long ns_magic = NSQ > 1 ? ~0ull / (unsigned int)NSQ + 1ull : 0ull; /* ceil(2^64 / NSQ) */
    const int root = __ldg(S.roots);
    int overflow = 0;
    unsigned long long n_shadow = 0, n_nodes = 0, n_flops = 0, n_mismatch = 0;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        atomicAdd(&cnt->rays_per_ray, total);
    }

    for (unsigned long long base = (unsigned long long)blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < total; base += stride) {
        const unsigned long long item = base + (threadIdx.x & 31);
        unsigned int h = 0;
        int s = 0;
        float4 head = make_float4(0.f, 0.f, 0.f, __int_as_float(-1));
        unsigned int entry = 0;
        unsigned int prog = FRT_PROG_ROOT(root);
        if (item < total) {
            const unsigned int idx = NSQ > 1 ? (unsigned int)__umul64hi(item, ns_magic) : (unsigned int)item; /* item / NSQ, exact while item * NSQ < 2^64 */
            entry = __ldg(pending + idx);
            prog = __ldg(pprog + idx);
            const int k = (int)(item - (unsigned long long)idx * (unsigned int)NSQ), q = (int)((entry >> 28) & 3u);
            s = qs_table ? (int)s_qs[q * NSQ + k] : quadrant_sample(lq, q, k);
            h = entry & FRT_PEND_HIT_MASK;
            head = *reinterpret_cast<const float4 *>(tmp + h);
        }
        const int set_a = __float_as_int(head.w);
        const int bulk = (MODE != 0 && set_a >= 0) ? (int)(entry >> 30) : 0; /* bit 0: decided by k_shadow_bulk, bit 1: as lit */
        int res = FRT_SH_SHADOWED;
        if (set_a >= 0) {
            if (bulk) {
                res = MODE == 2 && (bulk & 2) ? FRT_SH_LIT : FRT_SH_SHADOWED; /* counting frame: the hit's count is already set */
            } else if (S.n_roots != 1) {
                res = FRT_SH_UNDECIDED; /* several top-level shapes (world.c:189-191): never generated; FP64 handles it */
            } else {
                /* the FP32 ray: both world points rounded to FP32, difference and normalisation in FP32 */
                const float *pa = fpts + 3 * ((size_t)set_a * NS + s);
                const float px = __ldg(pa), py = __ldg(pa + 1), pz = __ldg(pa + 2);
                FrameF w;
                w.ox = head.x;
                w.oy = head.y;
                w.oz = head.z;
                const float vx = px - w.ox, vy = py - w.oy, vz = pz - w.oz;
                const float len2 = fmaf(vx, vx, fmaf(vy, vy, vz * vz));
                const float rinv = rsqrtf(len2);
                w.dx = vx * rinv;
                w.dy = vy * rinv;
                w.dz = vz * rinv;
                const float omax = fmaxf(fmaxf(fabsf(w.ox), fabsf(w.oy)), fabsf(w.oz));
                const float pmax = fmaxf(fmaxf(fabsf(px), fabsf(py)), fabsf(pz));
                /* origin rounded to FP32: 2u |o|; a WORLD box bound rounded to FP32 adds 2u Bmax to every slab numerator */
                const float eo_o = 2.0f * FRT_F32_U * omax;
                const float eo_w = fmaf(2.0f * FRT_F32_U, SF.bmax, fmaf(SF.ealign, omax, eo_o));
                const float ed_w = fmaf(2.0f * FRT_F32_U * (pmax + omax), rinv, FRT_F32_G + SF.ealign);
                frame_finish(w, eo_w, eo_w, eo_w, ed_w, ed_w, ed_w);
                const float Df = len2 * rinv;
                const float D_lo = Df - Df * ed_w, D_hi = Df + Df * ed_w;
                if (use_start && (prog & FRT_PROG_FAST)) {
                    res = trace_entry_program<COUNT>(SF, fnodes, cprog, prog, w, omax, eo_o, D_lo, D_hi, &n_nodes, &n_flops);
                } else {
                    /* the general walk from X1; a single-node program's tail verdict still ends it right after X1 */
                    const int tail = (use_start && FRT_PROG_COUNT(prog) == 1) ? FRT_PROG_TAIL(prog) : 0;
                    res = trace_shadow_f32_call<COUNT>(SF, fnodes, root, use_start ? FRT_PROG_NODE(prog, 0) : root, tail, tmp[h].relevant, w, omax,
                                                     eo_o, ed_w, D_lo, D_hi, &n_nodes, &n_flops);
                }
                if (COUNT && (res >> 4)) {
                    atomicAdd(&cnt->undecided_reason[min((res >> 4) & 15, 9)], 1ull);
                    atomicAdd(&cnt->undecided_node[(res >> 8) & 31], 1ull);
                }
                res &= 15;
            }
            if (COUNT) ++n_shadow;
            if (MODE == 2 && res != FRT_SH_UNDECIDED) {
                unsigned int h2;
                int sa2;
                Ray er;
                double dist2;
                shadow_item(S, recs, tmp, pts, NS, (unsigned long long)h * (unsigned int)NS + (unsigned int)s, false, h2, sa2, er, dist2);
                const double dist = normalise_shadow_ray(er, dist2);
                unsigned long long dn = 0, df = 0;
                const bool sh = trace_shadow<false>(S, er, dist, &overflow, &dn, &df);
                if (sh != (res == FRT_SH_SHADOWED)) {
                    ++n_mismatch;
                    res = sh ? FRT_SH_SHADOWED : FRT_SH_LIT;
                }
            }
        }
        const unsigned int slot = warp_append(&cnt->n_deferred, res == FRT_SH_UNDECIDED);
        if (res == FRT_SH_UNDECIDED) {
            if (slot < qcap) {
                queue[slot] = (unsigned long long)h * (unsigned int)NS + (unsigned int)s;
            } else {
                atomicOr(&cnt->overflow_queue, 1u);
            }
        }
        /* segmented count: lanes of the same hit are contiguous */
        const unsigned int active = __ballot_sync(0xffffffffu, set_a >= 0);
        const unsigned int lit_mask = __ballot_sync(0xffffffffu, res == FRT_SH_LIT);
        if (set_a >= 0) {
            const unsigned int peers = __match_any_sync(active, h);
            if ((threadIdx.x & 31) == (unsigned int)(__ffs(peers) - 1)) {
                int c = __popc(lit_mask & peers);
                if (c) {
                    atomicAdd(&tmp[h].unshadowed, c);
                }
            }
        }
    }
    if (overflow) {
        atomicOr(&cnt->overflow_csg, 1u);
    }
    if (COUNT) {
        for (int o = 16; o > 0; o >>= 1) {
            n_shadow += __shfl_down_sync(0xffffffffu, n_shadow, o);
            n_nodes += __shfl_down_sync(0xffffffffu, n_nodes, o);
            n_flops += __shfl_down_sync(0xffffffffu, n_flops, o);
            n_mismatch += __shfl_down_sync(0xffffffffu, n_mismatch, o);
        }
        if ((threadIdx.x & 31) == 0) {
            if (n_shadow) atomicAdd(&cnt->rays_shadow, n_shadow);
            if (n_nodes) atomicAdd(&cnt->shadow_nodes, n_nodes);
            if (n_flops) atomicAdd(&cnt->light_flops, n_flops);
            if (n_mismatch) atomicAdd(&cnt->f32_mismatch, n_mismatch);
        }
    }
}

/*
 * The same job with ONE WARP PER PENDING ENTRY: lanes are the samples of the entry's quadrant (25 of a 10 x 10 light;
 * larger entries take several passes).  Everything an entry carries -- hit, origin, sample set, start node, tail verdict,
 * shaft mask -- is the same in every lane, so the warp takes every branch of the walk together; what differs from lane to
 * lane is arithmetic (slab values, orderings), and in the common case (trace_entry_fast: the shaft walk left one WORLD
 * cube / CSG-of-cubes node and a tail verdict) it is straight-line code with selects.  The count of lit rays is one ballot
 * and one atomic per entry (no __match_any_sync), deferred rays are appended with one atomic per warp.  ncu on
 * k_shadow_f32 (profiles/r2a_k_shadow_f32_lines.txt): 1 030 warp-instructions per 32 rays at 23.3 active lanes, box tests
 * at 19.7 and the CSG combination at 15 lanes -- the warps there hold the tail of one entry and the head of the next.
 */
#ifndef FRT_ENTRY_MINB
#define FRT_ENTRY_MINB 6
#endif
template <int MODE>
__global__ void __launch_bounds__(128, FRT_ENTRY_MINB)
k_shadow_entry(DScene S, DSceneF SF, FrameParams F, const LightRec *__restrict__ recs, LightTmp *__restrict__ tmp, Counters *cnt,
               const unsigned int *__restrict__ pending, const unsigned int *__restrict__ pprog, int use_start, unsigned int pend_cap,
               int split_on, int light_idx, unsigned long long *__restrict__ queue, unsigned int qcap, int nodes_in_smem)
{
    constexpr bool COUNT = MODE != 0;
    extern __shared__ float4 s_nodes[];
    const float4 *fnodes = SF.fnodes;
    if (nodes_in_smem) {
        for (int k = threadIdx.x; k < 3 * SF.n_nodes; k += blockDim.x) {
            s_nodes[k] = SF.fnodes[k];
        }
        __syncthreads();
        fnodes = s_nodes;
    }
    const unsigned int n = min(cnt->n_pending, pend_cap);
    const int NS = S.lights[light_idx].num_samples;
    const int4 lq = split_on ? __ldg(SF.lquad + light_idx) : make_int4(NS, 0, 0, 0);
    const int NSQ = lq.x;
    const float *fpts = SF.lpoints + 3 * S.lights[light_idx].point_offset;
    const double *pts = S.lpoints + 3 * S.lights[light_idx].point_offset;
    const int root = __ldg(S.roots);
    const unsigned int lane = threadIdx.x & 31u;
    const unsigned int warps = gridDim.x * (blockDim.x >> 5), wid = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int overflow = 0;
    unsigned long long n_shadow = 0, n_nodes = 0, n_flops = 0, n_mismatch = 0;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        atomicAdd(&cnt->rays_per_ray, (unsigned long long)n * (unsigned int)NSQ);
    }
    const float bmax_term = 2.0f * FRT_F32_U * SF.bmax;

    for (unsigned int e = wid; e < n; e += warps) {
        const unsigned int entry = __ldg(pending + e);
        const unsigned int prog = use_start ? __ldg(pprog + e) : FRT_PROG_ROOT(root);
        const unsigned int h = entry & FRT_PEND_HIT_MASK;
        const int q = (int)((entry >> 28) & 3u);
        const float4 head = *reinterpret_cast<const float4 *>(tmp + h);
        const int set_a = __float_as_int(head.w);
        if (set_a < 0) {
            continue; /* the whole warp */
        }
        const unsigned int relevant = tmp[h].relevant;
        const int bulk = MODE != 0 ? (int)(entry >> 30) : 0; /* bit 0: decided by k_shadow_bulk, bit 1: as lit */
        const int node = FRT_PROG_NODE(prog, 0);
        const int tail = FRT_PROG_COUNT(prog) == 1 ? FRT_PROG_TAIL(prog) : 0; /* for the general walk from X1 */
        const bool fast = use_start && S.n_roots == 1 && (prog & FRT_PROG_FAST) != 0u;
        /* per entry: the origin's error terms */
        const float omax = fmaxf(fmaxf(fabsf(head.x), fabsf(head.y)), fabsf(head.z));
        const float eo_o = 2.0f * FRT_F32_U * omax;
        const float eo_w = bmax_term + fmaf(SF.ealign, omax, eo_o);
        const float *set_pts = fpts + 3 * (size_t)set_a * NS;
        if (COUNT && lane == 0 && !bulk) {
            atomicAdd(&cnt->entry_node[fast ? 0 : (FRT_PROG_TAIL(prog) != 0 ? 1 : 2)][(FRT_PROG_COUNT(prog) - 1) * 8 + min(node, 7)], 1ull);
        }
        int lit = 0;
        for (int k0 = 0; k0 < NSQ; k0 += 32) {
            const int k = k0 + (int)lane;
            const bool active = k < NSQ;
            const int s = active ? quadrant_sample(lq, q, k) : 0;
            int res = FRT_SH_SHADOWED;
            if (active) {
                if (bulk) {
                    res = MODE == 2 && (bulk & 2) ? FRT_SH_LIT : FRT_SH_SHADOWED; /* counting frame: the hit's count is already set */
                } else if (S.n_roots != 1) {
                    res = FRT_SH_UNDECIDED; /* several top-level shapes (world.c:189-191): never generated; FP64 handles it */
                } else {
                    /* the FP32 ray: both world points rounded to FP32, difference and normalisation in FP32 */
                    const float px = __ldg(set_pts + 3 * s), py = __ldg(set_pts + 3 * s + 1), pz = __ldg(set_pts + 3 * s + 2);
                    FrameF w;
                    w.ox = head.x;
                    w.oy = head.y;
                    w.oz = head.z;
                    const float vx = px - w.ox, vy = py - w.oy, vz = pz - w.oz;
                    const float len2 = fmaf(vx, vx, fmaf(vy, vy, vz * vz));
                    const float rinv = rsqrtf(len2);
                    w.dx = vx * rinv;
                    w.dy = vy * rinv;
                    w.dz = vz * rinv;
                    const float pmax = fmaxf(fmaxf(fabsf(px), fabsf(py)), fabsf(pz));
                    const float ed_w = fmaf(2.0f * FRT_F32_U * (pmax + omax), rinv, FRT_F32_G + SF.ealign);
                    frame_finish(w, eo_w, eo_w, eo_w, ed_w, ed_w, ed_w);
                    const float Df = len2 * rinv;
                    const float D_lo = Df - Df * ed_w, D_hi = Df + Df * ed_w;
                    if (fast) {
                        res = trace_entry_program<COUNT>(SF, fnodes, SF.csg_prog, prog, w, omax, eo_o, D_lo, D_hi, &n_nodes, &n_flops);
                    } else {
                        res = trace_shadow_f32_call<COUNT>(SF, fnodes, root, node, tail, relevant, w, omax, eo_o, ed_w, D_lo, D_hi, &n_nodes, &n_flops);
                    }
                    if (COUNT && (res >> 4)) {
                        atomicAdd(&cnt->undecided_reason[min((res >> 4) & 15, 9)], 1ull);
                        atomicAdd(&cnt->undecided_node[(res >> 8) & 31], 1ull);
                    }
                    res &= 15;
                }
                if (COUNT) ++n_shadow;
                if (MODE == 2 && res != FRT_SH_UNDECIDED) {
                    unsigned int h2;
                    int sa2;
                    Ray er;
                    double dist2;
                    shadow_item(S, recs, tmp, pts, NS, (unsigned long long)h * (unsigned int)NS + (unsigned int)s, false, h2, sa2, er, dist2);
                    const double dist = normalise_shadow_ray(er, dist2);
                    unsigned long long dn = 0, df = 0;
                    const bool sh = trace_shadow<false>(S, er, dist, &overflow, &dn, &df);
                    if (sh != (res == FRT_SH_SHADOWED)) {
                        ++n_mismatch;
                        res = sh ? FRT_SH_SHADOWED : FRT_SH_LIT;
                    }
                }
            }
            const bool defer = active && res == FRT_SH_UNDECIDED;
            const unsigned int dmask = __ballot_sync(0xffffffffu, defer);
            if (dmask) {
                unsigned int base = 0;
                if (lane == 0) {
                    base = atomicAdd(&cnt->n_deferred, (unsigned int)__popc(dmask));
                }
                base = __shfl_sync(0xffffffffu, base, 0);
                if (defer) {
                    const unsigned int slot = base + __popc(dmask & ((1u << lane) - 1u));
                    if (slot < qcap) {
                        queue[slot] = (unsigned long long)h * (unsigned int)NS + (unsigned int)s;
                    } else {
                        atomicOr(&cnt->overflow_queue, 1u);
                    }
                }
            }
            lit += __popc(__ballot_sync(0xffffffffu, active && res == FRT_SH_LIT));
        }
        if (lane == 0 && lit) {
            atomicAdd(&tmp[h].unshadowed, lit); /* integer sums: order-independent, the frame stays reproducible */
        }
    }
    if (overflow) {
        atomicOr(&cnt->overflow_csg, 1u);
    }
    if (COUNT) {
        for (int o = 16; o > 0; o >>= 1) {
            n_shadow += __shfl_down_sync(0xffffffffu, n_shadow, o);
            n_nodes += __shfl_down_sync(0xffffffffu, n_nodes, o);
            n_flops += __shfl_down_sync(0xffffffffu, n_flops, o);
            n_mismatch += __shfl_down_sync(0xffffffffu, n_mismatch, o);
        }
        if ((threadIdx.x & 31) == 0) {
            if (n_shadow) atomicAdd(&cnt->rays_shadow, n_shadow);
            if (n_nodes) atomicAdd(&cnt->shadow_nodes, n_nodes);
            if (n_flops) atomicAdd(&cnt->light_flops, n_flops);
            if (n_mismatch) atomicAdd(&cnt->f32_mismatch, n_mismatch);
        }
    }
}

__global__ void
k_to_float(const double *__restrict__ src, float *__restrict__ dst, size_t n)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        dst[i] = (float)src[i];
    }
}

/* the rays the FP32 pass could not decide, re-traced in FP64 (also the whole shadow pass under FRT_FLAG_F64_SHADOW) */
template <bool COUNT, bool ALL>
__global__ void __launch_bounds__(256)
k_shadow_exact(DScene S, DSceneF SF, FrameParams F, const LightRec *__restrict__ recs, LightTmp *__restrict__ tmp, Counters *cnt,
               int level, int light_idx, const unsigned long long *__restrict__ queue, unsigned int qcap)
{
    const unsigned int n = min(cnt->n_hits[level], F.capacity);
    const int NS = S.lights[light_idx].num_samples;
    const double *pts = S.lpoints + 3 * S.lights[light_idx].point_offset;
    const unsigned long long total = ALL ? (unsigned long long)n * (unsigned int)NS : (unsigned long long)min(cnt->n_deferred, qcap);
    const bool small = (unsigned long long)n * (unsigned int)NS <= 0xffffffffull;
    int overflow = 0;
    unsigned long long n_nodes = 0, n_flops = 0, n_shadow = 0;
    for (unsigned long long q = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; q < total;
         q += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long item = ALL ? q : queue[q];
        unsigned int h;
        int set_a;
        Ray sr;
        double dist2;
        shadow_item(S, recs, tmp, pts, NS, item, small, h, set_a, sr, dist2);
        if (set_a < 0) {
            continue;
        }
        const double dist = normalise_shadow_ray(sr, dist2);
        if (COUNT && ALL) ++n_shadow;
        bool shadowed;
        if (ALL || S.n_roots != 1) {
            shadowed = trace_shadow<COUNT>(S, sr, dist, &overflow, &n_nodes, &n_flops); /* pure FP64 (FRT_FLAG_F64_SHADOW) */
        } else {
            /* FP64 leaves and verdicts, FP32 conservative culls (trace_shadow_mixed) */
            FrameF w;
            w.ox = (float)sr.ox;
            w.oy = (float)sr.oy;
            w.oz = (float)sr.oz;
            w.dx = (float)sr.dx;
            w.dy = (float)sr.dy;
            w.dz = (float)sr.dz;
            const float omax = fmaxf(fmaxf(fabsf(w.ox), fabsf(w.oy)), fabsf(w.oz));
            const float eo_o = 2.0f * FRT_F32_U * omax;
            const float eo_w = fmaf(2.0f * FRT_F32_U, SF.bmax, fmaf(SF.ealign, omax, eo_o));
            const float ed_w = FRT_F32_G + SF.ealign; /* the FP64 unit direction rounded to FP32 */
            frame_finish(w, eo_w, eo_w, eo_w, ed_w, ed_w, ed_w);
            shadowed = trace_shadow_mixed<COUNT>(S, SF, sr, dist, w, omax, eo_o, ed_w, &overflow, &n_nodes, &n_flops);
            if (F.flags & FRT_FLAG_VERIFY_F32) { /* the deferred rays are checked against the pure FP64 walk too */
                unsigned long long dn = 0, df = 0;
                const bool ref = trace_shadow<false>(S, sr, dist, &overflow, &dn, &df);
                if (ref != shadowed) {
                    atomicAdd(&cnt->f32_mismatch, 1ull);
                    shadowed = ref;
                }
            }
        }
        if (!shadowed) {
            atomicAdd(&tmp[h].unshadowed, 1);
        }
    }
    if (overflow) {
        atomicOr(&cnt->overflow_csg, 1u);
    }
    if (COUNT) {
        if (n_shadow) atomicAdd(&cnt->rays_shadow, n_shadow);
        if (n_nodes) atomicAdd(&cnt->shadow_nodes, n_nodes);
        if (n_flops && ALL) atomicAdd(&cnt->light_flops, n_flops); /* light_flops describes the kernel that is timed */
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        atomicAdd(ALL ? &cnt->rays_per_ray : &cnt->deferred_total, total);
    }
}

/*
 * Shadow rays of MESH scenes (most leaves are triangles, which the FP32 filter has no fast form for, so every ray
 * would be deferred): all rays go straight to the mixed walk of trace_shadow_mixed -- FP32 conservative culls out of
 * the 48-byte node mirror, FP64 leaves and verdicts, the reference's order (H1) -- but run by PERSISTENT warps whose
 * lanes are REFILLED: the reference's divided tree makes ray lengths wildly uneven (100 to 10 000s of nodes), and
 * ncu on the queue version showed 14 of 32 lanes active and 30 % issue utilisation, the kernel's duration being the
 * latency of its longest rays.  Here a lane that finishes its ray takes the next (hit, sample) item from a global
 * counter as soon as FRT_MESH_REFILL lanes of its warp are idle; items are handed out in order, so the rays a warp
 * picks up together still share an origin.  The walk is "while-while": each lane first walks inner nodes (culls)
 * until it stands on a leaf, then the warp intersects its leaves together, and a leaf's mirror record carries the
 * parameter offset, transform and casts-shadow bit, so a leaf costs one dependent load (its FP64 vertices), not three.
 * C4 stand-in 800x1000, 61 M rays: 163 ms (deferred queue + one ray per thread) -> 100 ms.
 */
#define FRT_MESH_REFILL 24 /* idle lanes of a warp before it takes new items (measured: 4 -> 140 ms, 24 -> 104 ms, 32 = never -> 110 ms on the C4 stand-in) */
#define FRT_MESH_INNER 32 /* inner nodes a lane may walk before the warp looks at its leaves */
#define FRT_MESH_LIGHTS 8 /* lights whose shadow rays share one launch (point lights: few, long rays per launch otherwise) */
#ifndef FRT_MESH_MINB
#define FRT_MESH_MINB 6
#endif
/* triangle_local_intersect (triangle.c:11-45 / :122-156), the crossing only: prim_intersect's arithmetic, operation for operation */
__device__ __forceinline__ bool
triangle_crossing(const double *prm, const Ray &r, double &t)
{
    const double p1x = __ldg(prm + 0), p1y = __ldg(prm + 1), p1z = __ldg(prm + 2);
    const double e1x = __ldg(prm + 9), e1y = __ldg(prm + 10), e1z = __ldg(prm + 11);
    const double e2x = __ldg(prm + 12), e2y = __ldg(prm + 13), e2z = __ldg(prm + 14);
    const double cx = r.dy * e2z - r.dz * e2y;
    const double cy = r.dz * e2x - r.dx * e2z;
    const double cz = r.dx * e2y - r.dy * e2x;
    const double det = e1x * cx + e1y * cy + e1z * cz;
    if (fabs(det) < FRT_EPS) {
        return false;
    }
    const double f = 1.0 / det;
    const double sx = r.ox - p1x, sy = r.oy - p1y, sz = r.oz - p1z;
    const double u = f * (sx * cx + sy * cy + sz * cz);
    if (u < 0 || u > 1) {
        return false;
    }
    const double qx = sy * e1z - sz * e1y;
    const double qy = sz * e1x - sx * e1z;
    const double qz = sx * e1y - sy * e1x;
    const double v = f * (r.dx * qx + r.dy * qy + r.dz * qz);
    if (v < 0 || (u + v) > 1) {
        return false;
    }
    t = f * (e2x * qx + e2y * qy + e2z * qz);
    return true;
}

__device__ __noinline__ int
prim_intersect_inv_call(int type, const double *prm, const Ray &r, double t[4], double uv[2])
{
    return prim_intersect_inv(type, prm, r, inv_dir(r), t, uv);
}

template <bool COUNT, bool HAS_CSG>
__global__ void __launch_bounds__(128, FRT_MESH_MINB)
k_shadow_mesh(DScene S, DSceneF SF, FrameParams F, const LightRec *__restrict__ recs, LightTmp *__restrict__ tmp_base, size_t tmp_stride,
              Counters *cnt, int level, int first_light, int n_lights, int inner_budget, int refill_min)
{
    struct Frame {
        int right, skip, start, mid, op;
    };
    /* the lights first_light .. first_light + n_lights - 1 (<= FRT_MESH_LIGHTS) share the launch: items are light-major,
     * light k's (hit, sample) pairs follow light k - 1's; its per-hit records are tmp_base + k * tmp_stride */
    const unsigned int nh = min(cnt->n_hits[level], F.capacity);
    /* items in front of light k's: recomputed where it is needed (refill, a leaf under another transform) instead of nine
     * 64-bit values kept alive through the walk -- the walk's own state has to fit the registers */
    auto items_before = [&](int k) {
        unsigned long long c = 0;
        for (int j = 0; j < k && j < n_lights; ++j) {
            c += (unsigned long long)nh * (unsigned int)S.lights[first_light + j].num_samples;
        }
        return c;
    };
    auto light_of = [&](unsigned long long item, unsigned long long &before) {
        int lk = 0;
        before = 0;
        for (int j = 0; j + 1 < n_lights; ++j) {
            const unsigned long long c = (unsigned long long)nh * (unsigned int)S.lights[first_light + j].num_samples;
            if (item < before + c) {
                break;
            }
            before += c;
            lk = j + 1;
        }
        return lk;
    };
    const unsigned long long total = items_before(FRT_MESH_LIGHTS);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        atomicAdd(&cnt->rays_per_ray, total);
    }
    LightTmp *tmp = tmp_base;
    const float4 *fnodes = SF.fnodes;
    const int root = __ldg(S.roots);
    const int end = __float_as_int(__ldg(fnodes + 3 * root).y);
    const unsigned int lane = threadIdx.x & 31;
    const bool verify = (F.flags & FRT_FLAG_VERIFY_F32) != 0;

    bool active = false, exhausted = false;
    unsigned int h = 0;
    Ray lr{}; /* the ray in the frame of transform cur_xf_d; the world ray is re-derived from its item where it is needed again */
    unsigned long long my_item = 0;
    FrameF lf{}; /* the ONE current frame: the world ray, or the ray in the frame of transform cur_xf_f */
    float wox = 0.f, woy = 0.f, woz = 0.f, wdx = 0.f, wdy = 0.f, wdz = 0.f; /* the world ray in FP32: frames are rebuilt from it */
    double dist = 0.0;
    float omax = 0.f, eo_o = 0.f, ed_w = 0.f;
    auto world_frame = [&](FrameF &f) {
        f.ox = wox;
        f.oy = woy;
        f.oz = woz;
        f.dx = wdx;
        f.dy = wdy;
        f.dz = wdz;
        const float eo_w = fmaf(2.0f * FRT_F32_U, SF.bmax, fmaf(SF.ealign, omax, eo_o));
        frame_finish(f, eo_w, eo_w, eo_w, ed_w, ed_w, ed_w);
    };
    int cur_xf_f = 0, cur_xf_d = 0, i = 0, sp = 0, n = 0;
    CsgHit buf[HAS_CSG ? FRT_CSG_CAP : 1]; /* a tree without CSG nodes (OBJ meshes) needs neither list nor stack */
    Frame st[HAS_CSG ? FRT_CSG_DEPTH : 1];
    int overflow = 0;
    unsigned long long n_shadow = 0, n_nodes = 0, n_flops = 0, n_mismatch = 0;

    for (;;) {
        const unsigned int idle = __ballot_sync(0xffffffffu, !active);
        if (idle == 0xffffffffu && exhausted) {
            break;
        }
        if (!exhausted && (idle == 0xffffffffu || __popc(idle) >= refill_min)) {
            const int want = __popc(idle);
            unsigned long long base = 0;
            if (lane == 0) {
                base = atomicAdd(&cnt->mesh_next, (unsigned long long)want);
            }
            base = __shfl_sync(0xffffffffu, base, 0);
            exhausted = base + (unsigned long long)want >= total;
            const unsigned long long item = base + (unsigned int)__popc(idle & ((1u << lane) - 1u));
            if (!active && item < total) {
                int set_a;
                unsigned long long before;
                const int lk = light_of(item, before);
                const frt_light &L = S.lights[first_light + lk];
                tmp = tmp_base + (size_t)lk * tmp_stride;
                double dist2;
                Ray wr;
                shadow_item(S, recs, tmp, S.lpoints + 3 * L.point_offset, L.num_samples, item - before, false, h, set_a, wr, dist2);
                if (set_a >= 0) {
                    my_item = item;
                    dist = normalise_shadow_ray(wr, dist2);
                    wox = (float)wr.ox;
                    woy = (float)wr.oy;
                    woz = (float)wr.oz;
                    wdx = (float)wr.dx;
                    wdy = (float)wr.dy;
                    wdz = (float)wr.dz;
                    omax = fmaxf(fmaxf(fabsf(wox), fabsf(woy)), fabsf(woz));
                    eo_o = 2.0f * FRT_F32_U * omax;
                    ed_w = FRT_F32_G + SF.ealign; /* the FP64 unit direction rounded to FP32 */
                    world_frame(lf);
                    lr = wr;
                    cur_xf_f = cur_xf_d = 0;
                    i = root;
                    sp = 0;
                    n = 0;
                    active = true;
                    if (COUNT) ++n_shadow;
                }
            }
        }
        if (!active) {
            continue;
        }
        bool done = false, result = false;
        /* the world ray of this lane's item again (a leaf under another transform, the verifying build): same loads, same
         * operations as at refill */
        auto world_ray = [&]() {
            int set_a;
            unsigned long long before;
            const int lk = light_of(my_item, before);
            const frt_light &L = S.lights[first_light + lk];
            Ray wr;
            double dist2;
            unsigned int hh;
            shadow_item(S, recs, tmp, S.lpoints + 3 * L.point_offset, L.num_samples, my_item - before, false, hh, set_a, wr, dist2);
            normalise_shadow_ray(wr, dist2);
            return wr;
        };
        /* close every CSG whose left / right operand just ended (csg.c:104-118, :43-71) */
        auto close_frames = [&]() {
            while (HAS_CSG && sp > 0 && !done) {
                Frame &f = st[sp - 1];
                if (f.mid < 0 && i >= f.right) {
                    f.mid = n;
                }
                if (i < f.skip) {
                    break;
                }
                if (f.mid - f.start > 0 && n - f.mid > 0) {
                    for (int x = f.start + 1; x < n; ++x) {
                        CsgHit hh = buf[x];
                        int y = x - 1;
                        while (y >= f.start && buf[y].t > hh.t) {
                            buf[y + 1] = buf[y];
                            --y;
                        }
                        buf[y + 1] = hh;
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
                        result = tmin < dist;
                        done = true;
                    }
                }
            }
        };
        /* while-while: every lane first walks inner nodes (culls) until it stands on a leaf, then the lanes that do
         * intersect their leaves together -- the two code paths are not interleaved lane by lane */
        float4 q0 = make_float4(0.f, 0.f, 0.f, 0.f), lo = q0, hi = q0;
        bool at_leaf = false;
        for (int budget = inner_budget; budget > 0 && !done; --budget) {
            if (i >= end) {
                done = true;
                break;
            }
            q0 = __ldg(fnodes + 3 * i);
            lo = __ldg(fnodes + 3 * i + 1);
            const int flags = __float_as_int(q0.x), skip = __float_as_int(q0.y);
            const int type = flags & FRT_FN_TYPE_MASK;
            const bool leaf = type < FRT_CSG;
            if (leaf && !(flags & FRT_FN_LEAFBOX)) {
                at_leaf = true;
                break;
            }
            hi = __ldg(fnodes + 3 * i + 2);
            if (COUNT) {
                ++n_nodes;
                n_flops += FRT_COST_BBOX;
            }
            bool miss = false; /* a triangle leaf is culled by the bounds of its vertices like a group by its box */
            if (!(flags & FRT_FN_NOCULL)) {
                const int xf = __float_as_int(q0.z);
                float tn_lo, tn_hi, tf_lo, tf_hi;
                /* ONE current frame, replaced when a node names another one: choosing between two frames per node kept both
                 * in local memory (a pointer select: 14 % of the kernel's instructions were local loads) */
                if (xf != cur_xf_f) {
                    cur_xf_f = xf;
                    if (xf == 0) {
                        world_frame(lf);
                    } else {
                        FrameF w;
                        world_frame(w);
                        frame_local(lf, SF, xf, w, omax, eo_o, ed_w);
                    }
                }
                box_f(lf, lo, hi, tn_lo, tn_hi, tf_lo, tf_hi);
                miss = tn_lo > tf_hi || ((!HAS_CSG || sp == 0) && tf_hi < 0.0f);
            }
            if (miss) {
                i = skip;
            } else if (leaf) {
                at_leaf = true;
                break;
            } else {
                if (HAS_CSG && type == FRT_CSG) {
                    if (sp == FRT_CSG_DEPTH) {
                        overflow = 1;
                        done = true;
                    } else {
                        st[sp++] = Frame{ __float_as_int(q0.w), skip, n, -1, (flags >> FRT_FN_OP_SHIFT) & 3 };
                    }
                }
                i = i + 1;
            }
            if (HAS_CSG && sp > 0) {
                close_frames();
            }
        }
        if (at_leaf && !done) {
            /* leaves carry what the walk needs in their mirror record: {flags (type, casts), skip, -, param} {-, -, -, xform} */
            const int flags = __float_as_int(q0.x);
            const int type = flags & FRT_FN_TYPE_MASK, param = __float_as_int(q0.w), xform = __float_as_int(lo.w);
            if (COUNT) {
                ++n_nodes;
                n_flops += prim_cost(type);
            }
            if (xform != cur_xf_d) {
                cur_xf_d = xform;
                const Ray wr = world_ray();
                if (cur_xf_d == 0) {
                    lr = wr;
                } else {
                    lr = ray_to_local(S, cur_xf_d, wr);
                    if (COUNT) n_flops += FRT_COST_XFORM;
                }
            }
            const bool tri = type == FRT_TRIANGLE || type == FRT_SMOOTH_TRIANGLE;
            double t[4], uv[2], t_tri = 0.0;
            int k;
            if (tri) {
                /* the leaf of a mesh: Moeller-Trumbore inline, one crossing at most; everything else is a call, so that the
                 * quartic solver and the quadric bodies do not size this kernel's registers */
                k = triangle_crossing(S.params + param, lr, t_tri) ? 1 : 0;
            } else {
                k = prim_intersect_inv_call(type, S.params + (param < 0 ? 0 : param), lr, t, uv);
            }
            if ((!HAS_CSG || sp == 0) && tri) {
                if (k && !(t_tri <= 0)) {
                    result = (flags & FRT_FN_CASTS) && t_tri > 0 && t_tri < dist;
                    done = true;
                }
            } else if (!HAS_CSG || sp == 0) {
                bool stop = false;
                double tmin = CUDART_INF;
                for (int j = 0; j < k; ++j) {
                    stop = stop || !(t[j] <= 0);
                    if (t[j] > 0 && t[j] < tmin) {
                        tmin = t[j];
                    }
                }
                if (stop) {
                    result = (flags & FRT_FN_CASTS) && tmin < dist;
                    done = true;
                }
            } else {
                if (tri) {
                    t[0] = t_tri;
                }
                for (int j = 0; j < k; ++j) {
                    if (n == FRT_CSG_CAP) {
                        overflow = 1;
                        done = true;
                        break;
                    }
                    buf[n].t = t[j];
                    buf[n].leaf = i;
                    ++n;
                }
            }
            i = i + 1;
            if (HAS_CSG && sp > 0) {
                close_frames();
            }
        }
        if (done) {
            if (verify) { /* the mixed walk is checked against the pure FP64 walk ray by ray */
                unsigned long long dn = 0, df = 0;
                const bool ref = trace_shadow<false>(S, world_ray(), dist, &overflow, &dn, &df);
                if (ref != result) {
                    ++n_mismatch;
                    result = ref;
                }
            }
            if (!result) {
                atomicAdd(&tmp[h].unshadowed, 1);
            }
            active = false;
        }
    }
    if (overflow) {
        atomicOr(&cnt->overflow_csg, 1u);
    }
    if (COUNT) {
        if (n_shadow) atomicAdd(&cnt->rays_shadow, n_shadow);
        if (n_nodes) atomicAdd(&cnt->shadow_nodes, n_nodes);
        if (n_flops) atomicAdd(&cnt->light_flops, n_flops);
        if (n_shadow) atomicAdd(&cnt->deferred_total, n_shadow); /* every ray of a mesh scene takes the FP64-leaf walk */
    }
    if (n_mismatch) atomicAdd(&cnt->f32_mismatch, n_mismatch);
}

/*
 * lighting_microfacet (renderer.c:894-979) for the hits that received light, and the weighting into the pixel.
 * G lanes (a power of two <= 32) share a hit and stride over the sample set; partial sums are combined with a
 * fixed-order butterfly, so the result does not depend on scheduling.  Hits whose shadow rays were all blocked
 * (equal(shade_intensity, 0), :904) only add their ambient term -- most of a frame lit through a window.
 * T = float is the production path: the sums feed an 8-bit pixel through a continuous function, so FP32's 1e-7
 * relative error sits three orders of magnitude under one sRGB LSB; T = double (FRT_FLAG_F64_SHADING) keeps the
 * reference's arithmetic type.  Every geometric DECISION (hit / miss, shadowed / lit) stays in FP64 in k_extend /
 * k_shadow_*.
 */
#ifndef FRT_FINAL_MINB
#define FRT_FINAL_MINB 5 /* 48 registers: measured 1.71 / 1.60 / 1.65 ms at 4 / 5 / 6 blocks per SM (the kernel waits on the sample sets it reads: long_scoreboard 7.6) */
#endif
template <typename T, int G>
__global__ void __launch_bounds__(256, FRT_FINAL_MINB)
k_light_final(DScene S, FrameParams F, const LightRec *__restrict__ recs, const LightTmp *__restrict__ tmp, double *__restrict__ canvas,
              const Counters *cnt, int level, int light_idx, const float *__restrict__ flpoints, double *__restrict__ acc_amb)
{
    const unsigned int n = min(cnt->n_hits[level], F.capacity);
    const frt_light L = S.lights[light_idx];
    const int NS = L.num_samples;
    const double *pts = S.lpoints + 3 * L.point_offset;
    const float *fpts = flpoints + 3 * L.point_offset;
    const unsigned int lane_g = threadIdx.x % G;
    const unsigned int gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << ((threadIdx.x & 31) / G * G));
    const unsigned int groups = gridDim.x * blockDim.x / G;

    for (unsigned int hbase = (blockIdx.x * blockDim.x + threadIdx.x) / G;; hbase += groups) {
        /* all lanes of a warp iterate together so that the group shuffles stay converged */
        unsigned int h0 = __shfl_sync(0xffffffffu, hbase, 0);
        if (h0 >= n) {
            break;
        }
        const bool live = hbase < n;
        const LightRec *R = recs + (live ? hbase : 0);
        const LightTmp t = tmp[live ? hbase : 0];
        const double intensity = (double)t.unshadowed / (double)NS;
        const bool lit = live && t.contributes && !(fabs(intensity) < FRT_EPS); /* equal(shade_intensity, 0.0), renderer.c:904 */

        T sum_ndl = 0, sum_b = 0, sum_fb = 0;
        if (lit) {
            const double over[3] = { R->over[0], R->over[1], R->over[2] };
            const T nrm[3] = { (T)R->n[0], (T)R->n[1], (T)R->n[2] };
            const T eye[3] = { (T)R->eye[0], (T)R->eye[1], (T)R->eye[2] };
            const T Ns = (T)R->Ns;
            const T ndote = nrm[0] * eye[0] + nrm[1] * eye[1] + nrm[2] * eye[2];
            const double *pb = pts + 3 * (size_t)t.set_b * NS;
            const float *pbf = fpts + 3 * (size_t)t.set_b * NS;
            for (int s = lane_g; s < NS; s += G) {
                T lx, ly, lz;
                if (sizeof(T) == sizeof(float)) { /* FP32 copy of the sample points: half the bytes of the 157 MB cache */
                    lx = (T)(__ldg(pbf + 3 * s) - t.ox);
                    ly = (T)(__ldg(pbf + 3 * s + 1) - t.oy);
                    lz = (T)(__ldg(pbf + 3 * s + 2) - t.oz);
                } else {
                    lx = (T)(__ldg(pb + 3 * s) - over[0]);
                    ly = (T)(__ldg(pb + 3 * s + 1) - over[1]);
                    lz = (T)(__ldg(pb + 3 * s + 2) - over[2]);
                }
                T inv = sh_rsqrt(lx * lx + ly * ly + lz * lz);
                lx *= inv;
                ly *= inv;
                lz *= inv;
                T ndl = lx * nrm[0] + ly * nrm[1] + lz * nrm[2];
                if (ndl >= 0) {
                    if (F.use_diffuse) {
                        sum_ndl += ndl;
                    }
                    if (F.use_spec_highlight) {
                        T hx = lx + eye[0], hy = ly + eye[1], hz = lz + eye[2];
                        T hinv = sh_rsqrt(hx * hx + hy * hy + hz * hz);
                        hx *= hinv;
                        hy *= hinv;
                        hz *= hinv;
                        T ndh = max((T)0, nrm[0] * hx + nrm[1] * hy + nrm[2] * hz);
                        T edh_inv = sh_rcp(max((T)0, eye[0] * hx + eye[1] * hy + eye[2] * hz));
                        T ldh = lx * hx + ly * hy + lz * hz;
                        T dist_term = (Ns + 2) * sh_pow(ndh, Ns) * (T)(0.5 * M_1_PI);
                        T gc = 2 * ndh * edh_inv;
                        T geom = min((T)1, min(gc * ndote, gc * ndl));
                        T m1 = 1 - ldh;
                        T m2 = m1 * m1;
                        T factor = m2 * m2 * m1; /* pow(1 - L.H, 5) */
                        T brdf = dist_term * geom * sh_rcp(4 * ndl * ndote);
                        sum_b += brdf;
                        sum_fb += factor * brdf;
                    }
                }
            }
        }
        for (int o = G / 2; o > 0; o >>= 1) {
            sum_ndl += __shfl_xor_sync(gmask, sum_ndl, o);
            sum_b += __shfl_xor_sync(gmask, sum_b, o);
            sum_fb += __shfl_xor_sync(gmask, sum_fb, o);
        }
        if (live && lane_g == 0) {
            double c[3] = { 0.0, 0.0, 0.0 };
            if (lit) {
                const double scaling = intensity / (double)NS;
                for (int k = 0; k < 3; ++k) {
                    double d = R->Kd[k] * L.intensity[k] * (double)sum_ndl;
                    double sp = L.intensity[k] * (R->Ks[k] * (double)sum_b + (1.0 - R->Ks[k]) * (double)sum_fb);
                    c[k] = (d + sp) * scaling;
                }
            }
            if (F.use_ambient) {
                for (int k = 0; k < 3; ++k) {
                    if (F.use_gi) {
                        acc_amb[3 * (size_t)hbase + k] += R->Ka[k] * L.intensity[k]; /* joins the GI terms before the clamp of renderer.c:765 */
                    } else {
                        c[k] += R->Ka[k] * L.intensity[k];
                    }
                }
            }
            double *px = canvas + 4 * (size_t)R->pixel;
            for (int k = 0; k < 3; ++k) {
                double v = R->w[k] * c[k];
                if (v != 0.0) {
                    atomicAdd(px + k, v);
                }
            }
        }
    }
}

#include "frt_gi.cuh"
#include "frt_encode.cuh"

/* ------------------------------------------------------------------------------------------------ FMA peak */

template <typename T>
__global__ void
k_fma_peak(T *out, int iters)
{
    T a0 = (T)threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const T m = (T)0.999, c = (T)0.001;
    for (int i = 0; i < iters; ++i) {
        a0 = a0 * m + c; a1 = a1 * m + c; a2 = a2 * m + c; a3 = a3 * m + c;
        a4 = a4 * m + c; a5 = a5 * m + c; a6 = a6 * m + c; a7 = a7 * m + c;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

template <typename T>
static int
measure_fma(double *tflops)
{
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    const int blocks = sms * 8, threads = 256, iters = 1 << 15;
    T *out = nullptr;
    CK(cudaMalloc(&out, sizeof(T) * blocks * threads));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    k_fma_peak<T><<<blocks, threads>>>(out, 1024);
    double best = 0.0;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0));
        k_fma_peak<T><<<blocks, threads>>>(out, iters);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        double flops = 2.0 * 8.0 * (double)iters * blocks * threads;
        best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    *tflops = best;
    return FRT_OK;
}

extern "C" int
frt_measure_fma_peak(int device, double *fp64_tflops, double *fp32_tflops)
{
    CK(cudaSetDevice(device));
    int rc = measure_fma<double>(fp64_tflops);
    if (rc != FRT_OK) {
        return rc;
    }
    return measure_fma<float>(fp32_tflops);
}

/* ------------------------------------------------------------------------------------------------ scene */

static inline float
__int_as_float_host(int v)
{
    float f;
    memcpy(&f, &v, sizeof(f));
    return f;
}

static inline int
__float_as_int_host(float f)
{
    int v;
    memcpy(&v, &f, sizeof(v));
    return v;
}


struct frt_scene {
    int device = 0;
    DScene S{};
    DSceneF SF{};
    DCamera C{};
    frt_config cfg{};
    std::vector<void *> allocs;
    std::vector<size_t> alloc_bytes; /* parallel to allocs */
    std::vector<int> light_gw, light_ns;
    cudaStream_t upload_stream = nullptr; /* the light-sample cache and what is derived from it travel here */
    cudaEvent_t upload_ev = nullptr, ready_ev = nullptr;
    int sm_count = 148;
    bool upload_pending = false;          /* the render stream has not waited for upload_ev yet */
    unsigned int *gen_mismatch = nullptr; /* device word: words in which a rebuilt light-sample set differs from the caller's */
    bool gen_check_pending = false;       /* nobody has looked at it yet */
    bool gen_failed = false;
    unsigned int *h_nrays = nullptr; /* pinned: the next level's ray count, read back behind k_shade without stalling the stream */
    cudaEvent_t nrays_ev = nullptr;
    LightTmp *ltmp_multi = nullptr; /* mesh mode with several lights: one LightTmp array per light of a shared launch */
    size_t ltmp_multi_cap = 0;
    bool boxes_and_balls = false; /* no leaf type beyond cube / sphere / plane (k_extend instantiation) */
    bool has_csg = false;         /* the tree holds a CSG node (k_shadow_mesh instantiation) */
    bool has_maps = false;       /* some material carries a pattern / texture / bump map (k_shade instantiation) */
    bool has_refraction = false; /* some material refracts, or is a dissolving mirror: n1 / n2 containers are needed */
    bool mesh_mode = false; /* most leaves have no FP32 fast form (OBJ meshes): shadow rays go straight to k_shadow_mesh */
    double *canvas = nullptr;     /* hsize*vsize*4 doubles */
    double *samples = nullptr;
    int samples_u = 0, samples_v = 0;
    /* frame buffers, sized lazily */
    unsigned int capacity = 0;
    unsigned int chunk_samples = 12u << 20, cap_factor = 2;
    RayQ q[2]{};
    HitQ hq{};
    LightRec *recs = nullptr;
    LightTmp *ltmp = nullptr;
    unsigned int *pending = nullptr;  /* hits whose shadow rays k_shadow_bulk left to the per-ray kernels */
    unsigned long long *dq = nullptr; /* (hit, sample) items the FP32 shadow pass deferred to the FP64 pass */
    unsigned int dq_cap = 0;
    Counters *cnt = nullptr;
    /* photon maps (0 = caustic, 1 = global) and the GI work buffers */
    struct PMap {
        float4 *ra = nullptr, *rb = nullptr; /* photons as stored / imported */
        unsigned int cap = 0, count = 0;
        float4 *sa = nullptr, *sb = nullptr, *sd = nullptr; /* sorted by grid cell: position + packed direction, power, direction */
        unsigned int *cell_start = nullptr;
        PMView view{};
        bool built = false;
        bool scaled = false; /* pm_scale_photon_power already applied to ra / rb */
    } pm[2];
    unsigned int *pm_stored = nullptr; /* device counters, one per map */
    float *pm_dir_tab = nullptr;       /* the photon maps' direction tables (pm.c:54-60), 4 x 256 floats */
    float4 *pm_merged[2] = { nullptr, nullptr }; /* frt_multi_photons: every device's shard, gathered here before the import */
    bool pm_ready = false;
    GQuery *gq = nullptr, *gq2 = nullptr;     /* radiance-estimate requests of a batch, as queued / sorted by grid cell */
    unsigned int *gq_counts = nullptr, *gq_starts = nullptr;
    size_t gq_cells = 0;
    size_t gq2_cap = 0;                    /* requests gq2 holds */
    unsigned int *gq_part = nullptr;       /* per-block sums of the cell scan, and their exclusive scan */
    size_t gq_part_cap = 0;
    unsigned int *gq_work = nullptr;       /* k_knn_cell: {next chunk of requests, requests left to k_knn} */
    unsigned int gq_cap = 0;
    unsigned int *gq_n = nullptr;
    double *acc_amb = nullptr, *acc_fg = nullptr;
    unsigned int acc_cap = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[4]{};
    std::vector<cudaEvent_t> light_ev; /* start/stop pairs around every k_light launch of a frame (no host sync) */
    std::vector<void *> frame_allocs;
};

/*
 * Frame buffers (ray queues, hit records, deferred-ray queue) do not depend on the scene, and allocating a few GB of
 * them costs more than a frame.  When a scene is destroyed its set is parked per device and adopted by the next
 * scene that needs no more capacity, so render_multi() called once per frame from a host loop pays for them once.
 * frt_trim() frees the parked set.  (One frame at a time per device, like the reference's render_multi.)
 */
struct FrameSet {
    unsigned int capacity = 0;
    RayQ q[2]{};
    HitQ hq{};
    LightRec *recs = nullptr;
    LightTmp *ltmp = nullptr;
    unsigned int *pending = nullptr;
    unsigned long long *dq = nullptr;
    unsigned int dq_cap = 0;
    Counters *cnt = nullptr;
    std::vector<void *> allocs;
};
static std::mutex g_park_mu;
static std::map<int, FrameSet> g_parked;

static void
frameset_free(FrameSet &f)
{
    for (void *p : f.allocs) {
        cudaFree(p);
    }
    f = FrameSet{};
}

static void
scene_take(frt_scene *sc, FrameSet &f)
{
    sc->capacity = f.capacity;
    sc->q[0] = f.q[0];
    sc->q[1] = f.q[1];
    sc->hq = f.hq;
    sc->recs = f.recs;
    sc->ltmp = f.ltmp;
    sc->pending = f.pending;
    sc->dq = f.dq;
    sc->dq_cap = f.dq_cap;
    sc->cnt = f.cnt;
    sc->frame_allocs = std::move(f.allocs);
    f = FrameSet{};
}

static void
scene_give(frt_scene *sc, FrameSet &f)
{
    f.capacity = sc->capacity;
    f.q[0] = sc->q[0];
    f.q[1] = sc->q[1];
    f.hq = sc->hq;
    f.recs = sc->recs;
    f.ltmp = sc->ltmp;
    f.pending = sc->pending;
    f.dq = sc->dq;
    f.dq_cap = sc->dq_cap;
    f.cnt = sc->cnt;
    f.allocs = std::move(sc->frame_allocs);
    sc->frame_allocs.clear();
    sc->capacity = 0;
}

/*
 * The large scene buffers (the 157 MB light-sample cache of the shipped Cornell scene, its FP32 copy, the canvas) are
 * kept by size when a scene is destroyed, up to FRT_BLOCK_CACHE_MAX per device, and handed to the next scene that asks
 * for the same size: a host loop that builds a scene per frame then makes no cudaMalloc / cudaFree driver call per
 * frame at all (each is a device-wide synchronisation point and takes the driver lock).  frt_trim() frees them.
 */
#define FRT_BLOCK_CACHE_MIN ((size_t)1 << 20)
#define FRT_BLOCK_CACHE_MAX ((size_t)2 << 30)
struct BlockCache {
    std::multimap<size_t, void *> blocks;
    size_t bytes = 0;
};
static std::map<int, BlockCache> g_blocks; /* under g_park_mu */

static cudaError_t
scene_alloc(frt_scene *sc, void **p, size_t bytes)
{
    bytes = std::max<size_t>(bytes, 1);
    if (bytes < FRT_BLOCK_CACHE_MIN) { /* small buffers (tree, materials, tables) share a few power-of-two sizes */
        size_t r = 512;
        while (r < bytes) {
            r <<= 1;
        }
        bytes = r;
    }
    *p = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_park_mu);
        BlockCache &c = g_blocks[sc->device];
        auto it = c.blocks.find(bytes);
        if (it != c.blocks.end()) {
            *p = it->second;
            c.bytes -= bytes;
            c.blocks.erase(it);
        }
    }
    if (*p == nullptr) {
        cudaError_t e = cudaMalloc(p, bytes);
        if (e != cudaSuccess) {
            return e;
        }
    }
    sc->allocs.push_back(*p);
    sc->alloc_bytes.push_back(bytes);
    return cudaSuccess;
}

static void
scene_release_allocs(frt_scene *sc)
{
    std::lock_guard<std::mutex> lk(g_park_mu);
    BlockCache &c = g_blocks[sc->device];
    for (size_t i = 0; i < sc->allocs.size(); ++i) {
        const size_t bytes = i < sc->alloc_bytes.size() ? sc->alloc_bytes[i] : 0;
        if (bytes > 0 && c.bytes + bytes <= FRT_BLOCK_CACHE_MAX) {
            c.blocks.emplace(bytes, sc->allocs[i]);
            c.bytes += bytes;
        } else {
            cudaFree(sc->allocs[i]);
        }
    }
    sc->allocs.clear();
    sc->alloc_bytes.clear();
}

/* 64-byte page-locked slots, one per live scene (so that two scenes rendered from two threads on one device never share
 * the word the frame loop polls), recycled because cudaHostAlloc costs more than a small frame */
static std::vector<unsigned int *> g_pinned_free; /* under g_park_mu */

static unsigned int *
pinned_slot_take()
{
    {
        std::lock_guard<std::mutex> lk(g_park_mu);
        if (!g_pinned_free.empty()) {
            unsigned int *p = g_pinned_free.back();
            g_pinned_free.pop_back();
            return p;
        }
    }
    unsigned int *p = nullptr;
    if (cudaHostAlloc((void **)&p, 64, cudaHostAllocPortable) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}

static void
pinned_slot_give(unsigned int *p)
{
    if (p != nullptr) {
        std::lock_guard<std::mutex> lk(g_park_mu);
        g_pinned_free.push_back(p);
    }
}

/* streams and events of destroyed scenes, per device (creating two streams and a dozen events costs a small frame) */
struct HandlePool {
    std::vector<cudaStream_t> streams;
    std::vector<cudaEvent_t> timed, untimed;
};
static std::map<int, HandlePool> g_handles; /* under g_park_mu */

static cudaError_t
pool_stream(int device, cudaStream_t *out)
{
    {
        std::lock_guard<std::mutex> lk(g_park_mu);
        HandlePool &h = g_handles[device];
        if (!h.streams.empty()) {
            *out = h.streams.back();
            h.streams.pop_back();
            return cudaSuccess;
        }
    }
    return cudaStreamCreateWithFlags(out, cudaStreamNonBlocking);
}

static cudaError_t
pool_event(int device, bool timed, cudaEvent_t *out)
{
    {
        std::lock_guard<std::mutex> lk(g_park_mu);
        HandlePool &h = g_handles[device];
        std::vector<cudaEvent_t> &v = timed ? h.timed : h.untimed;
        if (!v.empty()) {
            *out = v.back();
            v.pop_back();
            return cudaSuccess;
        }
    }
    return timed ? cudaEventCreate(out) : cudaEventCreateWithFlags(out, cudaEventDisableTiming);
}

static void
pool_give(int device, cudaStream_t s)
{
    if (s != nullptr) {
        std::lock_guard<std::mutex> lk(g_park_mu);
        g_handles[device].streams.push_back(s);
    }
}

static void
pool_give(int device, bool timed, cudaEvent_t e)
{
    if (e != nullptr) {
        std::lock_guard<std::mutex> lk(g_park_mu);
        HandlePool &h = g_handles[device];
        (timed ? h.timed : h.untimed).push_back(e);
    }
}

static std::map<const char *, size_t> g_registered; /* page-locked caller buffers (frt_host_register), under g_park_mu */

static bool
host_range_registered(const void *ptr, size_t bytes)
{
    std::lock_guard<std::mutex> lk(g_park_mu);
    const char *p = (const char *)ptr;
    for (const auto &r : g_registered) {
        if (p >= r.first && p + bytes <= r.first + r.second) {
            return true;
        }
    }
    return false;
}

extern "C" int
frt_host_register(void *ptr, size_t bytes)
{
    if (ptr == nullptr || bytes == 0) {
        return frt_set_error(FRT_ERR_ARG, "frt_host_register: empty buffer");
    }
    CK(cudaHostRegister(ptr, bytes, cudaHostRegisterPortable));
    std::lock_guard<std::mutex> lk(g_park_mu);
    g_registered[(const char *)ptr] = bytes;
    return FRT_OK;
}

extern "C" int
frt_host_unregister(void *ptr)
{
    if (ptr == nullptr) {
        return frt_set_error(FRT_ERR_ARG, "frt_host_unregister: null pointer");
    }
    {
        std::lock_guard<std::mutex> lk(g_park_mu);
        g_registered.erase((const char *)ptr);
    }
    CK(cudaHostUnregister(ptr));
    return FRT_OK;
}

extern "C" void
frt_trim(int device)
{
    std::lock_guard<std::mutex> lk(g_park_mu);
    cudaSetDevice(device);
    auto it = g_parked.find(device);
    if (it != g_parked.end()) {
        frameset_free(it->second);
        g_parked.erase(it);
    }
    auto ht = g_handles.find(device);
    if (ht != g_handles.end()) {
        for (cudaStream_t st : ht->second.streams) cudaStreamDestroy(st);
        for (cudaEvent_t e : ht->second.timed) cudaEventDestroy(e);
        for (cudaEvent_t e : ht->second.untimed) cudaEventDestroy(e);
        g_handles.erase(ht);
    }
    auto bt = g_blocks.find(device);
    if (bt != g_blocks.end()) {
        for (auto &b : bt->second.blocks) {
            cudaFree(b.second);
        }
        g_blocks.erase(bt);
    }
}

template <typename T>
static int
upload(frt_scene *sc, const T *src, size_t count, const T **dst)
{
    /* Every copy travels on the scene's upload stream, the stream whose kernels (FP32 conversion, light bounds) read the
     * data; the render stream waits for an event recorded behind them (frt_scene_create).  A copy from pageable memory
     * returns once the source has been staged, so the caller's buffer may go away afterwards. */
    T *d = nullptr;
    size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
    CK(scene_alloc(sc, (void **)&d, bytes));
    if (count) {
        CK(cudaMemcpyAsync(d, src, count * sizeof(T), cudaMemcpyHostToDevice, sc->upload_stream));
    } else {
        CK(cudaMemsetAsync(d, 0, bytes, sc->upload_stream));
    }
    *dst = d;
    return FRT_OK;
}

static int
validate_desc(const frt_scene_desc *d)
{
    if (d == nullptr) {
        return frt_set_error(FRT_ERR_ARG, "null scene description");
    }
    if (d->abi_version != FRT_ABI_VERSION) {
        return frt_set_error(FRT_ERR_ARG, "scene description has ABI %d, library has %d", d->abi_version, FRT_ABI_VERSION);
    }
    if (d->n_nodes <= 0 || d->n_roots <= 0 || d->n_xforms <= 0 || d->nodes == nullptr || d->roots == nullptr || d->xforms == nullptr) {
        return frt_set_error(FRT_ERR_ARG, "scene has no nodes / roots / transforms");
    }
    if (d->n_materials < 0 || d->n_patterns < 0 || d->n_textures < 0 || d->n_lights < 0 || d->n_prim_params < 0 || d->n_texels < 0 ||
        d->n_light_points < 0 || d->n_pixel_samples < 0) {
        return frt_set_error(FRT_ERR_ARG, "scene description has a negative count");
    }
    if ((d->n_materials > 0 && d->materials == nullptr) || (d->n_patterns > 0 && d->patterns == nullptr) ||
        (d->n_textures > 0 && d->textures == nullptr) || (d->n_lights > 0 && d->lights == nullptr) ||
        (d->n_prim_params > 0 && d->prim_params == nullptr) || (d->n_texels > 0 && d->texels == nullptr)) {
        return frt_set_error(FRT_ERR_ARG, "scene description has a positive count with a null array");
    }
    /* every index and offset the kernels dereference */
    for (int i = 0; i < d->n_materials; ++i) {
        const frt_material &m = d->materials[i];
        const int maps[7] = { m.map_Ka, m.map_Kd, m.map_Ks, m.map_Ns, m.map_d, m.map_bump, m.map_refl };
        for (int k = 0; k < 7; ++k) {
            if (maps[k] < -1 || maps[k] >= d->n_patterns) {
                return frt_set_error(FRT_ERR_ARG, "material %d: map %d is pattern %d of %d", i, k, maps[k], d->n_patterns);
            }
        }
    }
    for (int i = 0; i < d->n_textures; ++i) {
        const frt_texture &t = d->textures[i];
        if (t.width <= 0 || t.height <= 0 || t.texel_offset < 0 || t.texel_offset + (int64_t)t.width * t.height > d->n_texels) {
            return frt_set_error(FRT_ERR_ARG, "texture %d (%d x %d at %lld) does not fit the %lld texels", i, t.width, t.height,
                                 (long long)t.texel_offset, (long long)d->n_texels);
        }
    }
    for (int i = 0; i < d->n_patterns; ++i) {
        const frt_pattern &p = d->patterns[i];
        bool ok = p.type >= 0 && p.type <= FRT_PAT_TEXTURE_MAP;
        auto child = [&](int idx) { return idx >= 0 && idx < d->n_patterns && idx != i; };
        if (ok) {
            switch (p.type) {
            case FRT_PAT_UV_TEXTURE: ok = p.i[0] >= 0 && p.i[0] < d->n_textures; break;
            case FRT_PAT_BLENDED: ok = child(p.i[0]) && child(p.i[1]); break;
            case FRT_PAT_NESTED: ok = child(p.i[0]) && child(p.i[1]) && child(p.i[2]); break;
            case FRT_PAT_PERTURBED: ok = child(p.i[0]); break;
            case FRT_PAT_CUBE_MAP:
            case FRT_PAT_CYLINDER_MAP:
            case FRT_PAT_TEXTURE_MAP: {
                const int faces = p.i[0] == FRT_UV_CUBE ? 6 : (p.i[0] == FRT_UV_CYLINDER ? 3 : 1);
                ok = p.i[0] >= 0 && p.i[0] <= FRT_UV_TRIANGLE && p.i[1] >= 0 && p.i[1] + faces <= d->n_patterns;
                break;
            }
            default: break;
            }
        }
        if (!ok) {
            return frt_set_error(FRT_ERR_ARG, "pattern %d (type %d) refers outside the pattern / texture tables", i, p.type);
        }
    }
    for (int i = 0; i < d->n_nodes; ++i) {
        const frt_node &n = d->nodes[i];
        if (n.type < 0 || n.type > FRT_GROUP || n.skip <= i || n.skip > d->n_nodes || n.xform < 0 || n.xform >= d->n_xforms) {
            return frt_set_error(FRT_ERR_ARG, "node %d is malformed (type %d skip %d xform %d)", i, n.type, n.skip, n.xform);
        }
        if (n.type < FRT_CSG && (n.material < 0 || n.material >= d->n_materials)) {
            return frt_set_error(FRT_ERR_ARG, "leaf node %d has material %d of %d", i, n.material, d->n_materials);
        }
        if (n.type == FRT_CSG && (n.right <= i + 1 || n.right >= n.skip)) {
            return frt_set_error(FRT_ERR_ARG, "CSG node %d has right child %d outside (%d, %d)", i, n.right, i + 1, n.skip);
        }
        bool needs_param = n.type == FRT_CONE || n.type == FRT_CYLINDER || n.type == FRT_TOROID || n.type == FRT_TRIANGLE || n.type == FRT_SMOOTH_TRIANGLE;
        const int64_t param_len = (n.type == FRT_TRIANGLE || n.type == FRT_SMOOTH_TRIANGLE) ? FRT_TRI_PARAMS : (n.type == FRT_TOROID ? 2 : 3);
        if (needs_param && (n.param < 0 || (int64_t)n.param + param_len > d->n_prim_params)) {
            return frt_set_error(FRT_ERR_ARG, "leaf node %d has parameter offset %d of %lld", i, n.param, (long long)d->n_prim_params);
        }
    }
    for (int i = 0; i < d->n_roots; ++i) {
        if (d->roots[i] < 0 || d->roots[i] >= d->n_nodes) {
            return frt_set_error(FRT_ERR_ARG, "root %d out of range", i);
        }
    }
    for (int i = 0; i < d->n_lights; ++i) {
        const frt_light &l = d->lights[i];
        if (l.num_samples <= 0 || l.cache_len <= 0 || l.point_offset < 0 || l.type < 0 || l.type > 3 ||
            l.point_offset + (int64_t)l.num_samples * l.cache_len > d->n_light_points) {
            return frt_set_error(FRT_ERR_ARG, "light %d has an inconsistent sample cache", i);
        }
    }
    const frt_camera &c = d->camera;
    if (c.hsize <= 0 || c.vsize <= 0 || c.usteps <= 0 || c.vsteps <= 0) {
        return frt_set_error(FRT_ERR_ARG, "camera has a non-positive size");
    }
    if (d->config.di_path_length < 0 || d->config.di_path_length > 6) {
        return frt_set_error(FRT_ERR_ARG, "di.path_length %d is outside 0..6", d->config.di_path_length);
    }
    return FRT_OK;
}

/* the xi = 0.5 table: sampler_reset_canonical_2d + sampler_shuffle_2d with no_jitter (sampler.c:401-461) */
static void
cmj_table_no_jitter(int s0, int s1, std::vector<double> &arr)
{
    arr.assign((size_t)2 * s0 * s1, 0.0);
    int n = s0, m = s1;
    for (int j = 0; j < n; ++j) {
        for (int i = 0; i < m; ++i) {
            int idx = 2 * (j * m + i);
            arr[idx] = (i + (j + 0.5) / (double)n) / (double)m;
            arr[idx + 1] = (j + (i + 0.5) / (double)m) / (double)n;
        }
    }
    m = s0;
    n = s1;
    for (int j = 0; j < n; ++j) {
        int k = (int)(j + 0.5 * (n - j));
        for (int i = 0; i < m; ++i) {
            std::swap(arr[2 * (j * m + i)], arr[2 * (k * m + i)]);
        }
    }
    for (int i = 0; i < m; ++i) {
        int k = (int)(i + 0.5 * (m - i));
        for (int j = 0; j < n; ++j) {
            std::swap(arr[2 * (j * m + i) + 1], arr[2 * (j * m + k) + 1]);
        }
    }
}

static void pm_free(frt_scene *sc);
static int pick_group_width(int num_samples);
static int ensure_gi_buffers(frt_scene *sc);

extern "C" void
frt_scene_destroy(frt_scene *sc)
{
    if (sc == nullptr) {
        return;
    }
    cudaSetDevice(sc->device);
    if (sc->upload_stream) {
        cudaStreamSynchronize(sc->upload_stream); /* the caller's page-locked buffer may go away after this call */
    }
    if (sc->stream) {
        cudaStreamSynchronize(sc->stream);
    }
    pm_free(sc);
    scene_release_allocs(sc);
    if (sc->capacity > 0 && !sc->frame_allocs.empty()) {
        std::lock_guard<std::mutex> lk(g_park_mu);
        FrameSet &slot = g_parked[sc->device];
        if (slot.capacity >= sc->capacity) {
            for (void *p : sc->frame_allocs) {
                cudaFree(p);
            }
        } else {
            frameset_free(slot);
            scene_give(sc, slot);
        }
    } else {
        for (void *p : sc->frame_allocs) {
            cudaFree(p);
        }
    }
    /* both streams are idle here (synchronised above): they and the events go back to the device's pool */
    for (auto &e : sc->ev) {
        pool_give(sc->device, true, e);
    }
    pool_give(sc->device, false, sc->nrays_ev);
    pinned_slot_give(sc->h_nrays);
    pool_give(sc->device, false, sc->upload_ev);
    pool_give(sc->device, false, sc->ready_ev);
    pool_give(sc->device, sc->upload_stream);
    for (auto &e : sc->light_ev) {
        pool_give(sc->device, true, e);
    }
    pool_give(sc->device, sc->stream);
    delete sc;
}

/*
 * FP32 mirror for the filtered shadow traversal (frt_shadow_f32.cuh).  A node whose composite world->local matrix is
 * axis-aligned (one significant entry per row: scalings, translations, quarter turns) gets its bounds -- the unit
 * cube of a cube leaf, the bounding box of a group / CSG -- mapped to WORLD space here, once, in FP64:
 *      local_k = s_k * world_p(k) + T_k   =>   world_p(k) in [(lo_k - T_k) / s_k, (hi_k - T_k) / s_k]  (sorted)
 * Leaf bounds are rounded to nearest (their rounding is part of the traversal's error term), cull bounds outward.
 */
static int
build_f32_mirror(frt_scene *sc, const frt_scene_desc *d)
{
    std::vector<float4> fx((size_t)4 * d->n_xforms), fn((size_t)3 * d->n_nodes), wb((size_t)2 * d->n_nodes);
    std::vector<float4> shaft((size_t)4 * std::max(d->n_lights, 1));
    double smin = 1.0, tilt = 0.0;
    std::vector<int> aligned(d->n_xforms, 0), perm((size_t)3 * d->n_xforms, 0);
    for (int i = 0; i < d->n_xforms; ++i) {
        const double *m = d->xforms[i].inv;
        float R[3];
        double row_tilt[3] = { 0.0, 0.0, 0.0 };
        bool ok = true;
        bool used[3] = { false, false, false };
        for (int k = 0; k < 3; ++k) {
            fx[4 * i + k] = make_float4((float)m[4 * k], (float)m[4 * k + 1], (float)m[4 * k + 2], (float)m[4 * k + 3]);
            double r = fabs(m[4 * k]) + fabs(m[4 * k + 1]) + fabs(m[4 * k + 2]);
            R[k] = nextafterf((float)(r * (1.0 + 1e-6)), INFINITY);
            int big = 0;
            for (int j = 1; j < 3; ++j) {
                if (fabs(m[4 * k + j]) > fabs(m[4 * k + big])) big = j;
            }
            double off = 0.0;
            for (int j = 0; j < 3; ++j) {
                if (j != big) off += fabs(m[4 * k + j]);
            }
            /* quarter turns built from a rounded pi leave off-axis entries of ~5e-12 (every wall of the Cornell box):
             * still axis-aligned for the filter, the tilt goes into its error terms (DSceneF::ealign) */
            if (off > 1e-9 * fabs(m[4 * k + big])) ok = false;
            row_tilt[k] = off / std::max(fabs(m[4 * k + big]), 1e-300);
            if (m[4 * k + big] == 0.0 || used[big]) ok = false;
            used[big] = true;
            perm[3 * i + k] = big;
        }
        fx[4 * i + 3] = make_float4(R[0], R[1], R[2], 0.f);
        aligned[i] = ok ? 1 : 0;
        for (int k = 0; ok && k < 3; ++k) {
            smin = std::min(smin, fabs(m[4 * k + perm[3 * i + k]]));
            tilt = std::max(tilt, row_tilt[k]);
        }
    }
    auto down = [](double x) { float f = (float)x; return ((double)f > x) ? nextafterf(f, -INFINITY) : f; };
    auto up = [](double x) { float f = (float)x; return ((double)f < x) ? nextafterf(f, INFINITY) : f; };
    double bmax = 0.0;
    for (int i = 0; i < d->n_nodes; ++i) {
        const frt_node &n = d->nodes[i];
        const bool inner = n.type >= FRT_CSG;
        const bool tri = n.type == FRT_TRIANGLE || n.type == FRT_SMOOTH_TRIANGLE;
        int flags = n.type;
        double lo[3], hi[3];
        if (inner) {
            for (int k = 0; k < 3; ++k) {
                lo[k] = n.bbox_min[k];
                hi[k] = n.bbox_max[k];
            }
            if (n.type == FRT_CSG) flags |= (n.csg_op & 3) << FRT_FN_OP_SHIFT;
        } else {
            for (int k = 0; k < 3; ++k) {
                lo[k] = -1.0;
                hi[k] = 1.0;
            }
            if (d->materials[n.material].casts_shadow) flags |= FRT_FN_CASTS;
            if (n.type == FRT_CUBE || n.type == FRT_SPHERE || n.type == FRT_PLANE) flags |= FRT_FN_FAST;
            if (tri) {
                /* a triangle's record carries the (padded) bounds of its vertices: the walks cull it like a group before the
                 * FP64 test -- the reference has no such test, and none is needed for a ray that misses the box */
                const frt_leafruns::Box b = frt_leafruns::triangle_box(d->prim_params + n.param);
                for (int k = 0; k < 3; ++k) {
                    lo[k] = b.lo[k];
                    hi[k] = b.hi[k];
                }
                flags |= FRT_FN_LEAFBOX;
            }
        }
        bool world = n.xform == 0;
        if (!world && aligned[n.xform] && (inner || n.type == FRT_CUBE || tri)) {
            const double *m = d->xforms[n.xform].inv;
            double wlo[3], whi[3];
            for (int k = 0; k < 3; ++k) {
                const int p = perm[3 * n.xform + k];
                const double sk = m[4 * k + p], T = m[4 * k + 3];
                double a = (lo[k] - T) / sk, b = (hi[k] - T) / sk;
                wlo[p] = std::min(a, b);
                whi[p] = std::max(a, b);
            }
            for (int k = 0; k < 3; ++k) {
                lo[k] = wlo[k];
                hi[k] = whi[k];
            }
            world = true;
        }
        if (n.type == FRT_GROUP && n.parent < 0) {
            flags |= FRT_FN_NOCULL; /* shadow rays start inside the world group's box: the test never culls */
        }
        if (world) {
            if (!tri) flags |= FRT_FN_WORLD;
            if (inner || n.type == FRT_CUBE || tri) {
                for (int k = 0; k < 3; ++k) {
                    if (std::isfinite(lo[k])) bmax = std::max(bmax, fabs(lo[k]));
                    if (std::isfinite(hi[k])) bmax = std::max(bmax, fabs(hi[k]));
                }
            }
        }
        /* world-space box of the node for shaft culling: the node's local bounds (leaf: by type) through the forward matrix */
        {
            double blo[3], bhi[3];
            bool finite = true;
            if (inner) {
                for (int k = 0; k < 3; ++k) {
                    blo[k] = n.bbox_min[k];
                    bhi[k] = n.bbox_max[k];
                }
            } else {
                const double *prm = d->prim_params + (n.param < 0 ? 0 : n.param);
                for (int k = 0; k < 3; ++k) {
                    blo[k] = -1.0;
                    bhi[k] = 1.0;
                }
                switch (n.type) {
                case FRT_PLANE:
                    finite = false;
                    break;
                case FRT_CYLINDER:
                    blo[1] = prm[0];
                    bhi[1] = prm[1];
                    break;
                case FRT_CONE: {
                    blo[1] = prm[0];
                    bhi[1] = prm[1];
                    double r = std::max(fabs(prm[0]), fabs(prm[1]));
                    blo[0] = blo[2] = -r;
                    bhi[0] = bhi[2] = r;
                    break;
                }
                case FRT_TOROID:
                    blo[0] = blo[2] = -(fabs(prm[0]) + fabs(prm[1]));
                    bhi[0] = bhi[2] = fabs(prm[0]) + fabs(prm[1]);
                    blo[1] = -fabs(prm[1]);
                    bhi[1] = fabs(prm[1]);
                    break;
                case FRT_TRIANGLE:
                case FRT_SMOOTH_TRIANGLE:
                    for (int k = 0; k < 3; ++k) {
                        blo[k] = std::min(prm[k], std::min(prm[3 + k], prm[6 + k]));
                        bhi[k] = std::max(prm[k], std::max(prm[3 + k], prm[6 + k]));
                    }
                    break;
                default:
                    break;
                }
            }
            for (int k = 0; k < 3; ++k) {
                finite = finite && std::isfinite(blo[k]) && std::isfinite(bhi[k]);
            }
            double wlo[3] = { -INFINITY, -INFINITY, -INFINITY }, whi[3] = { INFINITY, INFINITY, INFINITY };
            if (finite) {
                const double *m = d->xforms[n.xform].inv;
                double a[9] = { m[0], m[1], m[2], m[4], m[5], m[6], m[8], m[9], m[10] };
                double det = a[0] * (a[4] * a[8] - a[5] * a[7]) - a[1] * (a[3] * a[8] - a[5] * a[6]) + a[2] * (a[3] * a[7] - a[4] * a[6]);
                if (std::isfinite(det) && fabs(det) > 1e-300) {
                    double fi[9] = { (a[4] * a[8] - a[5] * a[7]) / det, (a[2] * a[7] - a[1] * a[8]) / det, (a[1] * a[5] - a[2] * a[4]) / det,
                                     (a[5] * a[6] - a[3] * a[8]) / det, (a[0] * a[8] - a[2] * a[6]) / det, (a[2] * a[3] - a[0] * a[5]) / det,
                                     (a[3] * a[7] - a[4] * a[6]) / det, (a[1] * a[6] - a[0] * a[7]) / det, (a[0] * a[4] - a[1] * a[3]) / det };
                    /* local = M w + T  =>  w = M^-1 (local - T): centre and half extent of the box's image */
                    double cl[3], hl[3];
                    for (int k = 0; k < 3; ++k) {
                        cl[k] = 0.5 * (blo[k] + bhi[k]) - m[4 * k + 3];
                        hl[k] = 0.5 * (bhi[k] - blo[k]);
                    }
                    bool ok = true;
                    for (int k = 0; k < 3; ++k) {
                        double c = fi[3 * k] * cl[0] + fi[3 * k + 1] * cl[1] + fi[3 * k + 2] * cl[2];
                        double e = fabs(fi[3 * k]) * hl[0] + fabs(fi[3 * k + 1]) * hl[1] + fabs(fi[3 * k + 2]) * hl[2];
                        e = e * (1.0 + 1e-9) + 1e-9 * (fabs(c) + 1.0); /* slack for the FP64 rounding of this inverse */
                        wlo[k] = c - e;
                        whi[k] = c + e;
                        ok = ok && std::isfinite(wlo[k]) && std::isfinite(whi[k]);
                    }
                    if (!ok) {
                        for (int k = 0; k < 3; ++k) {
                            wlo[k] = -INFINITY;
                            whi[k] = INFINITY;
                        }
                    }
                }
            }
            wb[2 * i] = make_float4(down(wlo[0]), down(wlo[1]), down(wlo[2]), 0.f);
            wb[2 * i + 1] = make_float4(up(whi[0]), up(whi[1]), up(whi[2]), 0.f);
        }
        float4 q0;
        q0.x = __int_as_float_host(flags);
        q0.y = __int_as_float_host(n.skip);
        q0.z = __int_as_float_host(world ? 0 : n.xform);
        q0.w = __int_as_float_host(inner ? n.right : n.param); /* leaves: what the FP64 leaf test needs, without a second record */
        fn[3 * i] = q0;
        if (inner) {
            fn[3 * i + 1] = make_float4(down(lo[0]), down(lo[1]), down(lo[2]), 0.f);
            fn[3 * i + 2] = make_float4(up(hi[0]), up(hi[1]), up(hi[2]), 0.f);
        } else if (tri) {
            fn[3 * i + 1] = make_float4(down(lo[0]), down(lo[1]), down(lo[2]), __int_as_float_host(n.xform));
            fn[3 * i + 2] = make_float4(up(hi[0]), up(hi[1]), up(hi[2]), __int_as_float_host(n.material));
        } else {
            fn[3 * i + 1] = make_float4((float)lo[0], (float)lo[1], (float)lo[2], __int_as_float_host(n.xform));
            fn[3 * i + 2] = make_float4((float)hi[0], (float)hi[1], (float)hi[2], __int_as_float_host(n.material));
        }
    }
    /* postfix programs of the outermost CSG nodes (frt_shadow_f32.cuh): operands first, then the operator */
    std::vector<int> prog;
    for (int i = 0; i < d->n_nodes; ++i) {
        const frt_node &n = d->nodes[i];
        if (n.type != FRT_CSG) {
            continue;
        }
        bool outermost = true;
        for (int p = n.parent; p >= 0; p = d->nodes[p].parent) {
            if (d->nodes[p].type == FRT_CSG) outermost = false;
        }
        if (!outermost) {
            continue;
        }
        const int start = (int)prog.size();
        bool ok = true;
        int depth = 0, max_depth = 0;
        /* iterative post-order over the CSG subtree */
        struct It { int node, stage; };
        std::vector<It> stk{ { i, 0 } };
        while (!stk.empty() && ok) {
            It &t = stk.back();
            const frt_node &c = d->nodes[t.node];
            if (c.type == FRT_CSG) {
                if (t.stage == 0) {
                    t.stage = 1;
                    stk.push_back({ t.node + 1, 0 });
                } else if (t.stage == 1) {
                    t.stage = 2;
                    stk.push_back({ c.right, 0 });
                } else {
                    prog.push_back(-(c.csg_op + 1));
                    depth -= 1;
                    stk.pop_back();
                }
            } else if (c.type == FRT_GROUP) {
                ok = false; /* a group as an operand: several spans */
            } else {
                prog.push_back(t.node);
                depth += 1;
                max_depth = std::max(max_depth, depth);
                stk.pop_back();
            }
        }
        /* left-deep: a leaf, then (leaf, operator) pairs -- one accumulator span evaluates it */
        const int len = (int)prog.size() - start;
        ok = ok && len >= 3 && (len % 2) == 1 && prog[start] >= 0;
        for (int k = start + 1; ok && k < (int)prog.size(); k += 2) {
            ok = prog[k] >= 0 && prog[k + 1] < 0;
        }
        if (ok) {
            int fl = __float_as_int_host(fn[3 * i].x) | FRT_FN_FAST;
            fn[3 * i].x = __int_as_float_host(fl);
            fn[3 * i + 1].w = __int_as_float_host(start);
            fn[3 * i + 2].w = __int_as_float_host((int)prog.size() - start);
        } else {
            prog.resize(start);
        }
    }
    if (prog.empty()) {
        prog.push_back(0);
    }
    /* sphere leaves whose composite transform is a similarity (uniform scale, rotation, translation): balls in WORLD space */
    std::vector<float4> wsph(32, make_float4(0.f, 0.f, 0.f, 0.f));
    for (int i = 0; i < std::min(d->n_nodes, 32); ++i) {
        const frt_node &n = d->nodes[i];
        if (n.type != FRT_SPHERE) {
            continue;
        }
        const double *m = d->xforms[n.xform].inv; /* local = A w + T */
        double rows[3], dots[3];
        for (int k = 0; k < 3; ++k) {
            rows[k] = sqrt(m[4 * k] * m[4 * k] + m[4 * k + 1] * m[4 * k + 1] + m[4 * k + 2] * m[4 * k + 2]);
        }
        dots[0] = m[0] * m[4] + m[1] * m[5] + m[2] * m[6];
        dots[1] = m[0] * m[8] + m[1] * m[9] + m[2] * m[10];
        dots[2] = m[4] * m[8] + m[5] * m[9] + m[6] * m[10];
        const double s0 = rows[0];
        bool ok = std::isfinite(s0) && s0 > 1e-12;
        for (int k = 0; ok && k < 3; ++k) {
            ok = fabs(rows[k] - s0) <= 1e-12 * s0 && fabs(dots[k]) <= 1e-12 * s0 * s0;
        }
        if (!ok) {
            continue;
        }
        /* A = s0 Q with Q orthogonal: w = A^-1 (local - T) = Q^T (local - T) / s0; the centre is the image of local 0 */
        double c[3];
        for (int k = 0; k < 3; ++k) {
            c[k] = -(m[k] * m[3] + m[4 + k] * m[7] + m[8 + k] * m[11]) / (s0 * s0);
        }
        wsph[i] = make_float4((float)c[0], (float)c[1], (float)c[2], (float)(1.0 / s0));
    }
    {
        const int rc_ = upload(sc, wsph.data(), wsph.size(), &sc->SF.wsphere);
        if (rc_ != FRT_OK) return rc_;
    }
    sc->SF.entry_fast = 0u;
    for (int i = 0; i < std::min(d->n_nodes, 32); ++i) {
        if (node_is_entry_fast(fn.data(), prog.data(), wsph.data(), i)) {
            sc->SF.entry_fast |= 1u << i;
        }
    }
    /* per light: a parallelogram that contains every surface sample (light.c:100-191), slightly inflated */
    for (int li = 0; li < d->n_lights; ++li) {
        const frt_light &L = d->lights[li];
        double c[4][3];
        if (L.type == 0) { /* area light: corner + s uvec*usteps + t vvec*vsteps, s, t in [0, 1] */
            for (int k = 0; k < 3; ++k) {
                const double U = L.uvec[k] * L.usteps, V = L.vvec[k] * L.vsteps;
                const double o = L.position[k] - 1e-6 * (U + V);
                c[0][k] = o;
                c[1][k] = o + U * (1.0 + 2e-6);
                c[2][k] = o + (U + V) * (1.0 + 2e-6);
                c[3][k] = o + V * (1.0 + 2e-6);
            }
        } else if (L.type == 1) { /* circle light: the square around the disc, in any basis of its plane */
            double nrm[3] = { L.normal[0], L.normal[1], L.normal[2] };
            double len = sqrt(nrm[0] * nrm[0] + nrm[1] * nrm[1] + nrm[2] * nrm[2]);
            double t1[3] = { 1, 0, 0 };
            if (len > 0) {
                for (int k = 0; k < 3; ++k) nrm[k] /= len;
            }
            if (fabs(nrm[0]) > 0.9) {
                t1[0] = 0;
                t1[1] = 1;
            }
            double dt = t1[0] * nrm[0] + t1[1] * nrm[1] + t1[2] * nrm[2];
            for (int k = 0; k < 3; ++k) t1[k] -= dt * nrm[k];
            len = sqrt(t1[0] * t1[0] + t1[1] * t1[1] + t1[2] * t1[2]);
            for (int k = 0; k < 3; ++k) t1[k] /= len;
            double t2[3] = { nrm[1] * t1[2] - nrm[2] * t1[1], nrm[2] * t1[0] - nrm[0] * t1[2], nrm[0] * t1[1] - nrm[1] * t1[0] };
            const double r = fabs(L.radius) * (1.0 + 1e-6) * 1.0000001;
            const double sgn[4][2] = { { -1, -1 }, { 1, -1 }, { 1, 1 }, { -1, 1 } };
            for (int q = 0; q < 4; ++q) {
                for (int k = 0; k < 3; ++k) {
                    c[q][k] = L.position[k] + r * (sgn[q][0] * t1[k] + sgn[q][1] * t2[k]);
                }
            }
        } else { /* point / hemisphere light: a single point */
            for (int q = 0; q < 4; ++q) {
                for (int k = 0; k < 3; ++k) c[q][k] = L.position[k];
            }
        }
        for (int q = 0; q < 4; ++q) {
            shaft[4 * li + q] = make_float4((float)c[q][0], (float)c[q][1], (float)c[q][2], 0.f);
        }
    }
    int rc = upload(sc, fx.data(), fx.size(), &sc->SF.fx);
    if (rc != FRT_OK) return rc;
    rc = upload(sc, wb.data(), wb.size(), &sc->SF.wbox);
    if (rc != FRT_OK) return rc;
    rc = upload(sc, shaft.data(), shaft.size(), &sc->SF.shaft);
    if (rc != FRT_OK) return rc;
    rc = upload(sc, prog.data(), prog.size(), &sc->SF.csg_prog);
    if (rc != FRT_OK) return rc;
    sc->SF.n_csg_prog = (int)prog.size();
    sc->SF.smin = (float)(smin * (1.0 - 1e-6));
    sc->SF.ealign = (float)(2.0 * tilt * (1.0 + 1e-6));
    rc = upload(sc, fn.data(), fn.size(), &sc->SF.fnodes);
    if (rc != FRT_OK) return rc;
    sc->SF.bmax = nextafterf((float)(bmax * (1.0 + 1e-6)), INFINITY);
    sc->SF.n_nodes = d->n_nodes;
    return FRT_OK;
}

/*
 * What the upload stream derives from the light points once they are on the device (copied, or rebuilt by k_light_gen):
 * their FP32 copy (only for the lights that were copied -- k_light_gen writes both precisions) and, per light, the
 * bounds of its sample points (all of them, and per quadrant of the sample grid) over every cached set, reduced on the
 * device from the points as they are -- no assumption about where the sampler puts sample (u, v).  All of it runs on the
 * upload stream behind the copy / the generator, so none of it needs the host.
 */
static int
build_light_bounds(frt_scene *sc, const frt_scene_desc *d, const std::vector<int> &generated)
{
    cudaStream_t us = sc->upload_stream;
    float *fp = const_cast<float *>(sc->SF.lpoints);
    for (int li = 0; li < d->n_lights; ++li) {
        if (generated[li]) {
            continue;
        }
        const frt_light &L = d->lights[li];
        const size_t np = (size_t)3 * L.num_samples * L.cache_len;
        if (np) {
            const int blocks = (int)std::min<size_t>((np + 255) / 256, (size_t)sc->sm_count * 8);
            k_to_float<<<blocks, 256, 0, us>>>(sc->S.lpoints + 3 * L.point_offset, fp + 3 * L.point_offset, np);
            CK(cudaGetLastError());
        }
    }
    const int nl = std::max(d->n_lights, 1);
    std::vector<int4> lquad(nl, make_int4(1, 0, 0, 0));
    std::vector<int> chunks(nl, 1);
    double *partial = nullptr, *lbox = nullptr;
    CK(scene_alloc(sc, (void **)&partial, sizeof(double) * 6 * 4 * FRT_BOX_CHUNKS * nl));
    CK(scene_alloc(sc, (void **)&lbox, sizeof(double) * 30 * nl));
    for (int li = 0; li < d->n_lights; ++li) {
        const frt_light &L = d->lights[li];
        const int NS = L.num_samples;
        int4 lq = make_int4(NS, 0, 0, 0);
        if (NS >= 16 && L.usteps > 0 && L.vsteps > 0 && L.usteps % 2 == 0 && L.vsteps % 2 == 0 && L.usteps * L.vsteps == NS) {
            lq = make_int4(NS / 4, L.usteps / 2, L.vsteps / 2, L.usteps);
        }
        lquad[li] = lq;
        chunks[li] = std::max(1, std::min(FRT_BOX_CHUNKS, L.cache_len));
    }
    int rc = upload(sc, lquad.data(), lquad.size(), &sc->SF.lquad);
    if (rc != FRT_OK) return rc;
    const int *dchunks = nullptr;
    rc = upload(sc, chunks.data(), chunks.size(), &dchunks);
    if (rc != FRT_OK) return rc;
    for (int li = 0; li < d->n_lights; ++li) {
        const frt_light &L = d->lights[li];
        const int nq = lquad[li].y ? 4 : 1;
        k_light_boxes<<<dim3(chunks[li], nq), 256, 0, us>>>(sc->S.lpoints + 3 * L.point_offset, L.cache_len, L.num_samples, lquad[li],
                                                          partial + (size_t)li * 4 * FRT_BOX_CHUNKS * 6, FRT_BOX_CHUNKS);
        CK(cudaGetLastError());
    }
    if (d->n_lights > 0) {
        k_light_boxes_finish<<<d->n_lights, 32, 0, us>>>(partial, sc->SF.lquad, dchunks, lbox);
        CK(cudaGetLastError());
    }
    sc->SF.lbox = lbox;
    return FRT_OK;
}

/* k_light_gen for one listed light, then the bit-for-bit comparison of the caller's sets; *d_mismatch accumulates */
static int
generate_light_cache(frt_scene *sc, const frt_scene_desc *d, const frt_light_gen &g, double *pool64, float *pool32, unsigned int *d_mismatch)
{
    const frt_light &L = d->lights[g.light];
    LightGenParams P{};
    for (int k = 0; k < 3; ++k) {
        P.corner[k] = L.position[k];
        P.uvec[k] = L.uvec[k];
        P.vvec[k] = L.vvec[k];
    }
    P.usteps = L.usteps;
    P.vsteps = L.vsteps;
    P.cache_len = L.cache_len;
    P.x0 = g.drand48_state & FRT_LCG_MASK;
    const unsigned long long per_set = 2ull * L.usteps * L.vsteps + L.usteps + L.vsteps;
    LcgJump j = lcg_jump(per_set);
    for (int b = 0; b < 32; ++b) {
        P.pow2[b] = j;
        const unsigned long long a = j.a;
        j.a = (a * a) & FRT_LCG_MASK;
        j.c = (a * j.c + j.c) & FRT_LCG_MASK;
    }
    for (int b = 0; b < 6; ++b) {
        P.step2[b] = lcg_jump(1ull << b);
    }
    const size_t smem = (size_t)FRT_LGEN_WARPS * (per_set + 2ull * L.num_samples) * sizeof(double);
    if (smem > 48 * 1024) {
        CK(cudaFuncSetAttribute(k_light_gen, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    const int blocks = (L.cache_len + FRT_LGEN_WARPS - 1) / FRT_LGEN_WARPS;
    cudaStream_t us = sc->upload_stream;
    k_light_gen<<<blocks, FRT_LGEN_WARPS * 32, smem, us>>>(P, pool64 + 3 * L.point_offset, pool32 + 3 * L.point_offset);
    CK(cudaGetLastError());
    const int words = 3 * L.num_samples;
    for (int k = 0; k < g.n_verify; ++k) {
        double *expect = nullptr;
        CK(scene_alloc(sc, (void **)&expect, sizeof(double) * words));
        CK(cudaMemcpyAsync(expect, g.verify_points[k], sizeof(double) * words, cudaMemcpyHostToDevice, us));
        k_light_gen_verify<<<1, 256, 0, us>>>(pool64 + 3 * (L.point_offset + (size_t)g.verify_set[k] * L.num_samples), expect, words, d_mismatch);
        CK(cudaGetLastError());
    }
    return FRT_OK;
}

static int
validate_gens(const frt_scene_desc *d, const frt_light_gen *gens, int n_gens, std::vector<int> &generated)
{
    generated.assign(std::max(d->n_lights, 1), 0);
    if (n_gens < 0 || (n_gens > 0 && gens == nullptr)) {
        return frt_set_error(FRT_ERR_ARG, "frt_scene_create_gen: %d generators, null list", n_gens);
    }
    for (int k = 0; k < n_gens; ++k) {
        const frt_light_gen &g = gens[k];
        if (g.light < 0 || g.light >= d->n_lights || generated[g.light]) {
            return frt_set_error(FRT_ERR_ARG, "light generator %d names light %d of %d (or names it twice)", k, g.light, d->n_lights);
        }
        const frt_light &L = d->lights[g.light];
        if (L.type != 0 || !L.jitter || L.usteps <= 0 || L.vsteps <= 0 || L.usteps > FRT_LGEN_MAX_STEPS || L.vsteps > FRT_LGEN_MAX_STEPS ||
            L.usteps * L.vsteps != L.num_samples) {
            return frt_set_error(FRT_ERR_ARG, "light %d is not a jittered rectangular area light with at most %d x %d steps", g.light,
                                 FRT_LGEN_MAX_STEPS, FRT_LGEN_MAX_STEPS);
        }
        if (g.n_verify < 0 || g.n_verify > FRT_GEN_VERIFY_MAX) {
            return frt_set_error(FRT_ERR_ARG, "light generator %d: n_verify %d is outside 0..%d", k, g.n_verify, FRT_GEN_VERIFY_MAX);
        }
        for (int v = 0; v < g.n_verify; ++v) {
            if (g.verify_points[v] == nullptr || g.verify_set[v] < 0 || g.verify_set[v] >= L.cache_len) {
                return frt_set_error(FRT_ERR_ARG, "light generator %d: verify set %d is null or outside the cache", k, v);
            }
        }
        generated[g.light] = 1;
    }
    for (int li = 0; li < d->n_lights; ++li) {
        if (!generated[li] && d->light_points == nullptr && d->n_light_points > 0) {
            return frt_set_error(FRT_ERR_ARG, "light %d has no generator and the description carries no light points", li);
        }
    }
    return FRT_OK;
}

extern "C" uint64_t
frt_drand48_advance(uint64_t x, uint64_t draws)
{
    const LcgJump j = lcg_jump(draws);
    return (j.a * (x & FRT_LCG_MASK) + j.c) & FRT_LCG_MASK;
}

/* wrapping sum over the words of bits(word) * (2 * index + 1): order- and position-sensitive, associative */
__global__ void
k_points_checksum(const double *__restrict__ pts, unsigned long long first_word, unsigned long long n_words, unsigned long long *sum)
{
    unsigned long long acc = 0;
    for (unsigned long long k = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; k < n_words; k += (unsigned long long)gridDim.x * blockDim.x) {
        acc += (unsigned long long)__double_as_longlong(pts[first_word + k]) * (2ull * k + 1ull);
    }
    for (int o = 16; o > 0; o >>= 1) {
        acc += __shfl_down_sync(0xffffffffu, acc, o);
    }
    if ((threadIdx.x & 31) == 0 && acc) {
        atomicAdd(sum, acc);
    }
}

extern "C" uint64_t
frt_light_points_checksum_host(const double *points, int64_t first_point, int64_t n_points)
{
    uint64_t acc = 0;
    if (points == nullptr || first_point < 0 || n_points <= 0) {
        return 0;
    }
    const double *p = points + 3 * first_point;
    for (uint64_t k = 0; k < (uint64_t)n_points * 3u; ++k) {
        uint64_t bits;
        memcpy(&bits, p + k, sizeof(bits));
        acc += bits * (2u * k + 1u);
    }
    return acc;
}

static int scene_create(const frt_scene_desc *d, int device, const frt_light_gen *gens, int n_gens, frt_scene **out);

extern "C" int
frt_scene_create(const frt_scene_desc *d, int device, frt_scene **out)
{
    return scene_create(d, device, nullptr, 0, out);
}

extern "C" int
frt_scene_create_gen(const frt_scene_desc *d, int device, const frt_light_gen *gens, int n_gens, frt_scene **out)
{
    return scene_create(d, device, gens, n_gens, out);
}

static int
scene_create(const frt_scene_desc *d, int device, const frt_light_gen *gens, int n_gens, frt_scene **out)
{
    if (out == nullptr) {
        return frt_set_error(FRT_ERR_ARG, "frt_scene_create: null out pointer");
    }
    *out = nullptr;
    int rc = validate_desc(d);
    if (rc != FRT_OK) {
        return rc;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        return frt_set_error(FRT_ERR_CUDA, "no CUDA device is visible: the B200 core has no CPU fallback");
    }
    if (device < 0 || device >= ndev) {
        return frt_set_error(FRT_ERR_ARG, "device %d of %d", device, ndev);
    }
    CK(cudaSetDevice(device));
    std::vector<int> generated;
    rc = validate_gens(d, gens, n_gens, generated);
    if (rc != FRT_OK) {
        return rc;
    }
    const bool timing = getenv("FRT_DEBUG_TIMING") != nullptr;
    /* bounding groups over the long triangle runs the reference's group_divide leaves behind (frt_leafruns.h): from here
     * on `d` is the tree with those groups in it; node indices never leave the library */
    frt_scene_desc with_runs;
    std::vector<frt_node> run_nodes;
    std::vector<int32_t> run_roots;
    {
        const char *env = getenv("FRT_LEAF_RUNS");
        long inserted = 0;
        if ((env == nullptr || atoi(env) != 0) && frt_leafruns::augment(d, run_nodes, run_roots, &inserted)) {
            with_runs = *d;
            with_runs.nodes = run_nodes.data();
            with_runs.n_nodes = (int32_t)run_nodes.size();
            with_runs.roots = run_roots.data();
            if (timing) {
                fprintf(stderr, "[frt] scene_create: %ld groups inserted over triangle runs (%d -> %d nodes)\n", inserted, d->n_nodes,
                        with_runs.n_nodes);
            }
            d = &with_runs;
        }
    }

    frt_scene *sc = new frt_scene();
    sc->device = device;
    auto t_start = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) {
        if (timing) {
            auto t = std::chrono::steady_clock::now();
            fprintf(stderr, "[frt] scene_create %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t - t_start).count());
            t_start = t;
        }
    };
    {
        /* cudaGetDeviceProperties takes milliseconds; the attribute query does not */
        int sms = 0;
        sc->sm_count = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && sms > 0 ? sms : 148;
    }
    lap("device attribute");
#define UP(expr)                     \
    do {                             \
        int rc_ = (expr);            \
        if (rc_ != FRT_OK) {         \
            frt_scene_destroy(sc);   \
            return rc_;              \
        }                            \
    } while (0)
    /* two streams: everything the host hands over travels on the upload stream, the frame runs on the render stream.
     * The render stream waits for `ready_ev` (the small buffers: tree, materials, mirror) before its first kernel and
     * for `upload_ev` (the light points and what is derived from them) where its first light stage begins. */
    if (pool_stream(device, &sc->upload_stream) != cudaSuccess || pool_stream(device, &sc->stream) != cudaSuccess ||
        pool_event(device, false, &sc->upload_ev) != cudaSuccess || pool_event(device, false, &sc->ready_ev) != cudaSuccess) {
        frt_scene_destroy(sc);
        return frt_set_error(FRT_ERR_CUDA, "cudaStreamCreate failed");
    }

    /* nodes -> 2 x int4, bounding boxes -> 6 doubles */
    std::vector<int4> nodes((size_t)2 * d->n_nodes);
    std::vector<double> bbox((size_t)6 * d->n_nodes);
    for (int i = 0; i < d->n_nodes; ++i) {
        const frt_node &n = d->nodes[i];
        nodes[2 * i] = make_int4(n.type, n.skip, n.xform, n.material);
        nodes[2 * i + 1] = make_int4(n.param, n.csg_op, n.right, n.parent);
        for (int k = 0; k < 3; ++k) {
            bbox[6 * i + k] = n.bbox_min[k];
            bbox[6 * i + 3 + k] = n.bbox_max[k];
        }
    }
    DScene &S = sc->S;
    UP(upload(sc, nodes.data(), nodes.size(), &S.nodes));
    UP(upload(sc, bbox.data(), bbox.size(), &S.bbox));
    UP(upload(sc, (const double *)d->xforms, (size_t)12 * d->n_xforms, &S.xinv));
    UP(upload(sc, d->prim_params, (size_t)d->n_prim_params, &S.params));
    UP(upload(sc, d->materials, (size_t)d->n_materials, &S.mats));
    UP(upload(sc, d->patterns, (size_t)d->n_patterns, &S.pats));
    UP(upload(sc, d->textures, (size_t)d->n_textures, &S.texs));
    {
        /* textures: the raw FP64 texels travel once, k_texture_ingest turns them into what a fetch returns */
        const double *raw = nullptr;
        UP(upload(sc, d->texels, (size_t)3 * d->n_texels, &raw));
        float4 *tex = nullptr;
        if (scene_alloc(sc, (void **)&tex, std::max<size_t>((size_t)d->n_texels, 1) * sizeof(float4)) != cudaSuccess) {
            frt_scene_destroy(sc);
            return frt_set_error(FRT_ERR_CUDA, "cudaMalloc of the textures failed");
        }
        for (int i = 0; i < d->n_textures; ++i) {
            const frt_texture &t = d->textures[i];
            const size_t n = (size_t)t.width * t.height;
            const int blocks = (int)std::min<size_t>((n + 255) / 256, (size_t)sc->sm_count * 8);
            k_texture_ingest<<<blocks, 256, 0, sc->upload_stream>>>(raw + 3 * t.texel_offset, t.width, t.height, t.super_sample, t.color_fn,
                                                                   tex + t.texel_offset);
        }
        if (cudaGetLastError() != cudaSuccess) {
            frt_scene_destroy(sc);
            return frt_set_error(FRT_ERR_CUDA, "texture ingest failed");
        }
        S.texels = tex;
    }
    UP(upload(sc, d->lights, (size_t)d->n_lights, &S.lights));
    UP(upload(sc, d->roots, (size_t)d->n_roots, &S.roots));
    lap("scene buffers");
    UP(build_f32_mirror(sc, d));
    lap("fp32 mirror");
    const frt_camera &c = d->camera;
    {
        std::vector<double> table;
        if (d->pixel_samples != nullptr && d->n_pixel_samples == (int64_t)2 * c.usteps * c.vsteps) {
            table.assign(d->pixel_samples, d->pixel_samples + d->n_pixel_samples);
        } else {
            cmj_table_no_jitter(c.usteps, c.vsteps, table);
        }
        const double *dtab = nullptr;
        UP(upload(sc, table.data(), table.size(), &dtab));
        sc->samples = const_cast<double *>(dtab);
        sc->samples_u = c.usteps;
        sc->samples_v = c.vsteps;
        sc->C.samples = dtab;
    }
    if (cudaEventRecord(sc->ready_ev, sc->upload_stream) != cudaSuccess || cudaStreamWaitEvent(sc->stream, sc->ready_ev, 0) != cudaSuccess) {
        frt_scene_destroy(sc);
        return frt_set_error(FRT_ERR_CUDA, "cudaEventRecord (ready) failed");
    }

    /* The light-sample cache is the one large buffer of a scene (157 MB for the shipped Cornell box).  A light listed in
     * `gens` has its cache rebuilt on the device (frt_lightgen.cuh); the others are copied -- asynchronously when the
     * caller page-locked the pool (frt_host_register): the first frame then only waits for it where its light stage
     * begins, and the buffer must stay valid until the first frt_render (or frt_scene_destroy) returns; an unregistered
     * buffer is staged before this call returns, as ever. */
    const size_t lp_count = (size_t)3 * d->n_light_points;
    const size_t lp_bytes = lp_count * sizeof(double);
    double *pool64 = nullptr;
    float *pool32 = nullptr;
    if (scene_alloc(sc, (void **)&pool64, std::max<size_t>(lp_bytes, 8)) != cudaSuccess ||
        scene_alloc(sc, (void **)&pool32, std::max<size_t>(lp_count, 1) * sizeof(float)) != cudaSuccess) {
        frt_scene_destroy(sc);
        return frt_set_error(FRT_ERR_CUDA, "cudaMalloc of the light points failed");
    }
    S.lpoints = pool64;
    sc->SF.lpoints = pool32;
    lap("light pool allocation");
    bool async_points = false;
    unsigned int *d_mismatch = nullptr;
    if (n_gens == 0) {
        async_points = lp_bytes >= ((size_t)1 << 20) && host_range_registered(d->light_points, lp_bytes);
        if (lp_bytes && cudaMemcpyAsync(pool64, d->light_points, lp_bytes, cudaMemcpyHostToDevice, sc->upload_stream) != cudaSuccess) {
            frt_scene_destroy(sc);
            return frt_set_error(FRT_ERR_CUDA, "upload of the light points failed");
        }
    } else {
        if (scene_alloc(sc, (void **)&d_mismatch, sizeof(unsigned int)) != cudaSuccess ||
            cudaMemsetAsync(d_mismatch, 0, sizeof(unsigned int), sc->upload_stream) != cudaSuccess) {
            frt_scene_destroy(sc);
            return frt_set_error(FRT_ERR_CUDA, "cudaMalloc (light generator) failed");
        }
        for (int li = 0; li < d->n_lights; ++li) {
            if (generated[li]) {
                continue;
            }
            const frt_light &L = d->lights[li];
            const size_t n = (size_t)3 * L.num_samples * L.cache_len;
            if (n && cudaMemcpyAsync(pool64 + 3 * L.point_offset, d->light_points + 3 * L.point_offset, n * sizeof(double),
                                     cudaMemcpyHostToDevice, sc->upload_stream) != cudaSuccess) {
                frt_scene_destroy(sc);
                return frt_set_error(FRT_ERR_CUDA, "upload of the light points failed");
            }
        }
        for (int k = 0; k < n_gens; ++k) {
            UP(generate_light_cache(sc, d, gens[k], pool64, pool32, d_mismatch));
        }
    }
    lap("light points enqueued");
    UP(build_light_bounds(sc, d, generated));
    lap("light bounds enqueued");
    if (cudaEventRecord(sc->upload_ev, sc->upload_stream) != cudaSuccess) {
        frt_scene_destroy(sc);
        return frt_set_error(FRT_ERR_CUDA, "cudaEventRecord (upload) failed");
    }
    sc->upload_pending = true;
    if (d_mismatch != nullptr) {
        /* the comparison runs behind this call: frt_scene_gen_status waits for it; frt_render reads it with its first frame */
        sc->gen_mismatch = d_mismatch;
        sc->gen_check_pending = true;
    } else if (!async_points) {
        cudaStreamSynchronize(sc->upload_stream); /* nothing of the caller's is read after this call returns */
    }
    lap("upload stream drained");
    S.n_roots = d->n_roots;
    S.n_nodes = d->n_nodes;
    S.n_lights = d->n_lights;

    DCamera &C = sc->C;
    C.hsize = c.hsize;
    C.vsize = c.vsize;
    C.usteps = c.usteps;
    C.vsteps = c.vsteps;
    C.half_width = c.half_width;
    C.half_height = c.half_height;
    C.pixel_size = c.pixel_size;
    C.canvas_distance = c.canvas_distance;
    memcpy(C.inv, c.inv, sizeof(C.inv));
    C.aperture_type = c.aperture_type;
    C.jitter = c.aperture_jitter;
    C.aperture_size = c.aperture_size;
    memcpy(C.aperture_args, c.aperture_args, sizeof(C.aperture_args));

    sc->cfg = d->config;
    for (int i = 0; i < d->n_materials; ++i) {
        const frt_material &m = d->materials[i];
        if (m.map_Ka >= 0 || m.map_Kd >= 0 || m.map_Ks >= 0 || m.map_Ns >= 0 || m.map_d >= 0 || m.map_bump >= 0 || m.map_refl >= 0) {
            sc->has_maps = true;
        }
        /* the containers feed the refraction ratio and Schlick's reflectance (renderer.c:403-447, :607): needed when a ray can
         * refract (over_d > 0 with a non-zero Tf under FRT_FLAG_NO_PRUNE too) or when a mirror dissolves (over_d < 1) */
        if (m.Tr > 0.0 || m.Tf[0] != 0.0 || m.Tf[1] != 0.0 || m.Tf[2] != 0.0 || m.Ni != 1.0) {
            sc->has_refraction = true;
        }
    }
    {
        long leaves = 0, slow = 0;
        for (int i = 0; i < d->n_nodes; ++i) {
            const int t = d->nodes[i].type;
            if (t < FRT_CSG) {
                ++leaves;
                slow += !(t == FRT_CUBE || t == FRT_SPHERE || t == FRT_PLANE);
            }
        }
        sc->mesh_mode = slow * 2 > leaves && d->n_nodes > 64;
        sc->boxes_and_balls = slow == 0;
        for (int i = 0; i < d->n_nodes; ++i) {
            sc->has_csg = sc->has_csg || d->nodes[i].type == FRT_CSG;
        }
        const char *env = getenv("FRT_MESH_MODE");
        if (env != nullptr && *env) {
            sc->mesh_mode = atoi(env) != 0;
        }
    }
    sc->light_gw.resize(d->n_lights);
    for (int i = 0; i < d->n_lights; ++i) {
        sc->light_gw[i] = pick_group_width(d->lights[i].num_samples);
        sc->light_ns.push_back(d->lights[i].num_samples);
    }

    size_t cbytes = (size_t)c.hsize * c.vsize * 4 * sizeof(double);
    void *cv = nullptr;
    if (scene_alloc(sc, &cv, cbytes) != cudaSuccess) {
        frt_scene_destroy(sc);
        return frt_set_error(FRT_ERR_CUDA, "cudaMalloc of the canvas failed");
    }
    sc->canvas = (double *)cv;
    for (auto &e : sc->ev) {
        if (pool_event(device, true, &e) != cudaSuccess) {
            frt_scene_destroy(sc);
            return frt_set_error(FRT_ERR_CUDA, "cudaEventCreate failed");
        }
    }
    /* one pinned slot per scene (recycled through a per-process free list): the next level's ray count lands here */
    if (pool_event(device, false, &sc->nrays_ev) == cudaSuccess) {
        sc->h_nrays = pinned_slot_take();
    }
#undef UP
    lap("canvas, events");
    *out = sc;
    return FRT_OK;
}

/* ------------------------------------------------------------------------------------------------ frame */

static int
count_owned_rows(int vsize, int rank, int world, int rpb)
{
    int n = 0;
    for (int y = 0; y < vsize; ++y) {
        if ((y / rpb) % world == rank) {
            ++n;
        }
    }
    return n;
}

extern "C" int
frt_owned_rows(const frt_scene_desc *d, const frt_render_cfg *cfg, int32_t *rows, int cap)
{
    if (d == nullptr || cfg == nullptr) {
        return 0;
    }
    int world = cfg->world > 0 ? cfg->world : 1;
    int rpb = cfg->rows_per_block > 0 ? cfg->rows_per_block : 4;
    int n = 0;
    for (int y = 0; y < d->camera.vsize; ++y) {
        if ((y / rpb) % world == cfg->rank) {
            if (rows != nullptr && n < cap) {
                rows[n] = y;
            }
            ++n;
        }
    }
    return n;
}

/*
 * A device buffer the other processes of the node can write (CUDA IPC): one process per GPU gathers a frame on one device
 * without a collective -- every rank passes the opened pointer to frt_render as its canvas and its row blocks travel over
 * NVLink in one strided copy on its render stream, straight to where they belong.
 */
extern "C" int
frt_shared_buffer_create(int device, size_t bytes, void **device_ptr, void *handle64)
{
    if (device_ptr == nullptr || handle64 == nullptr || bytes == 0) {
        return frt_set_error(FRT_ERR_ARG, "frt_shared_buffer_create: bad argument");
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "the handle travels as 64 bytes");
    CK(cudaSetDevice(device));
    void *p = nullptr;
    CK(cudaMalloc(&p, bytes)); /* an allocation of its own: the handle names the allocation, not an address inside one */
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        return frt_set_error(FRT_ERR_CUDA, "cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
    }
    memcpy(handle64, &h, sizeof(h));
    *device_ptr = p;
    return FRT_OK;
}

extern "C" int
frt_shared_buffer_open(int device, const void *handle64, void **device_ptr)
{
    if (device_ptr == nullptr || handle64 == nullptr) {
        return frt_set_error(FRT_ERR_ARG, "frt_shared_buffer_open: bad argument");
    }
    CK(cudaSetDevice(device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    CK(cudaIpcOpenMemHandle(device_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return FRT_OK;
}

extern "C" int
frt_shared_buffer_close(void *device_ptr, int opened)
{
    if (device_ptr == nullptr) {
        return FRT_OK;
    }
    if (opened) {
        CK(cudaIpcCloseMemHandle(device_ptr));
    } else {
        CK(cudaFree(device_ptr));
    }
    return FRT_OK;
}

/* the tree a scene is uploaded with (frt_leafruns.h): host only, no device needed */
extern "C" int
frt_tree_with_runs(const frt_scene_desc *d, frt_node *nodes, int cap, int32_t *roots)
{
    if (validate_desc(d) != FRT_OK) {
        return -1;
    }
    std::vector<frt_node> out;
    std::vector<int32_t> r;
    if (!frt_leafruns::augment(d, out, r, nullptr)) {
        out.assign(d->nodes, d->nodes + d->n_nodes);
        r.assign(d->roots, d->roots + d->n_roots);
    }
    if (nodes != nullptr && (size_t)cap >= out.size()) {
        memcpy(nodes, out.data(), out.size() * sizeof(frt_node));
        if (roots != nullptr) {
            memcpy(roots, r.data(), r.size() * sizeof(int32_t));
        }
    }
    return (int)out.size();
}

template <typename T>
static int
frame_alloc(frt_scene *sc, T **p, size_t count)
{
    void *d = nullptr;
    CK(cudaMalloc(&d, std::max<size_t>(count, 1) * sizeof(T)));
    sc->frame_allocs.push_back(d);
    *p = (T *)d;
    return FRT_OK;
}

static int
ensure_frame_buffers(frt_scene *sc, unsigned int capacity)
{
    if (sc->capacity >= capacity) {
        return FRT_OK;
    }
    for (void *p : sc->frame_allocs) {
        cudaFree(p);
    }
    sc->frame_allocs.clear();
    sc->capacity = 0;
    {
        std::lock_guard<std::mutex> lk(g_park_mu);
        auto it = g_parked.find(sc->device);
        if (it != g_parked.end()) {
            if (it->second.capacity >= capacity) {
                scene_take(sc, it->second);
                g_parked.erase(it);
                return FRT_OK;
            }
            frameset_free(it->second);
            g_parked.erase(it);
        }
    }
#define FA(p) do { int rc_ = frame_alloc(sc, &(p), capacity); if (rc_ != FRT_OK) return rc_; } while (0)
    for (int k = 0; k < 2; ++k) {
        RayQ &q = sc->q[k];
        FA(q.ox); FA(q.oy); FA(q.oz); FA(q.dx); FA(q.dy); FA(q.dz);
        FA(q.wr); FA(q.wg); FA(q.wb); FA(q.pixel); FA(q.rng);
    }
    FA(sc->hq.t); FA(sc->hq.u); FA(sc->hq.v); FA(sc->hq.leaf);
    FA(sc->recs);
    FA(sc->ltmp);
#undef FA
    {
        /* one entry per (hit, quadrant of the light's sample grid), then one word per entry: its program (frt_shadow_f32.cuh) */
        int rc_ = frame_alloc(sc, &sc->pending, (size_t)capacity * 8);
        if (rc_ != FRT_OK) {
            return rc_;
        }
    }
    {
        unsigned long long want_q = std::min<unsigned long long>((unsigned long long)capacity * 4ull, 0x7fffffffull);
        int rc_ = frame_alloc(sc, &sc->dq, (size_t)want_q);
        if (rc_ != FRT_OK) {
            return rc_;
        }
        sc->dq_cap = (unsigned int)want_q;
    }
    int rc = frame_alloc(sc, &sc->cnt, 1);
    if (rc != FRT_OK) {
        return rc;
    }
    sc->capacity = capacity;
    return FRT_OK;
}

/* the light stage is the first consumer of the light points and of what the upload stream derives from them */
static int
wait_for_upload(frt_scene *sc, cudaStream_t s)
{
    if (sc->upload_pending) {
        CK(cudaStreamWaitEvent(s, sc->upload_ev, 0));
        sc->upload_pending = false;
    }
    return FRT_OK;
}

template <bool COUNT>
static void
launch_shadow_mesh(frt_scene *sc, int blocks, cudaStream_t s, const FrameParams &F, LightTmp *tmp, size_t tmp_stride, int level, int first_light,
                   int n_lights, int inner, int refill)
{
    if (sc->has_csg) {
        k_shadow_mesh<COUNT, true><<<blocks, 128, 0, s>>>(sc->S, sc->SF, F, sc->recs, tmp, tmp_stride, sc->cnt, level, first_light, n_lights, inner, refill);
    } else {
        k_shadow_mesh<COUNT, false><<<blocks, 128, 0, s>>>(sc->S, sc->SF, F, sc->recs, tmp, tmp_stride, sc->cnt, level, first_light, n_lights, inner, refill);
    }
}

static void
launch_light_pre(frt_scene *sc, const FrameParams &F, int blocks, int level, int light, int g, LightTmp *tmp)
{
    const int shaft_on = sc->S.n_roots == 1 && !(F.flags & (FRT_FLAG_F64_SHADOW | FRT_FLAG_NO_SHAFT));
    {
        /* the lanes of a hit share its nodes' box tests but each repeats the pyramid's set-up: few nodes, few lanes */
        /* measured on the Cornell frame (16 nodes): 1 lane per hit 0.89 ms, 2: 1.05, 4: 1.51, 8: 2.40 */
        static const int pre_group = []() { const char *e = getenv("FRT_PRE_GROUP"); return (e != nullptr && *e) ? atoi(e) : 1; }();
        g = (pre_group == 2 || pre_group == 4 || pre_group == 8 || pre_group == 16 || pre_group == 32) ? pre_group : 1;
    }
#define LP(G) k_light_pre<G><<<blocks, 256, 0, sc->stream>>>(sc->S, F, sc->recs, tmp, sc->cnt, level, light, sc->SF, shaft_on)
    switch (g) {
    case 1: LP(1); break;
    case 2: LP(2); break;
    case 4: LP(4); break;
    case 8: LP(8); break;
    case 16: LP(16); break;
    default: LP(32); break;
    }
#undef LP
}

template <typename T>
static void
launch_light_final(frt_scene *sc, const FrameParams &F, int blocks, int level, int light, int g, const LightTmp *tmp)
{
#define LF(G) k_light_final<T, G><<<blocks, 256, 0, sc->stream>>>(sc->S, F, sc->recs, tmp, sc->canvas, sc->cnt, level, light, sc->SF.lpoints, sc->acc_amb)
    switch (g) {
    case 1: LF(1); break;
    case 2: LF(2); break;
    case 4: LF(4); break;
    case 8: LF(8); break;
    case 16: LF(16); break;
    default: LF(32); break;
    }
#undef LF
}

static int
pick_group_width(int num_samples)
{
    /* lanes per hit: the largest power of two <= 32 that wastes the fewest lane-iterations on this sample count */
    const char *env = getenv("FRT_LIGHT_GROUP");
    if (env != nullptr && *env) {
        int g = atoi(env);
        if (g == 1 || g == 2 || g == 4 || g == 8 || g == 16 || g == 32) {
            return g;
        }
    }
    int best = 1;
    double best_cost = 1e30;
    for (int g = 1; g <= 32; g *= 2) {
        int iters = (num_samples + g - 1) / g;
        double waste = (double)(iters * g) / (double)num_samples;
        /* prefer wider groups (more coherent warps) when the waste is equal, but never beyond 8 lanes per hit */
        double cost = waste - 1e-3 * g;
        if (g <= 8 && cost < best_cost) {
            best_cost = cost;
            best = g;
        }
    }
    return best;
}

extern "C" int
frt_canvas_download(frt_scene *sc, double *canvas_rgba)
{
    if (sc == nullptr || canvas_rgba == nullptr) {
        return frt_set_error(FRT_ERR_ARG, "frt_canvas_download: null argument");
    }
    CK(cudaSetDevice(sc->device));
    size_t bytes = (size_t)sc->C.hsize * sc->C.vsize * 4 * sizeof(double);
    CK(cudaMemcpyAsync(canvas_rgba, sc->canvas, bytes, cudaMemcpyDeviceToHost, sc->stream));
    CK(cudaStreamSynchronize(sc->stream));
    return FRT_OK;
}

extern "C" int
frt_canvas_device_ptr(frt_scene *sc, void **device_ptr)
{
    if (sc == nullptr || device_ptr == nullptr) {
        return frt_set_error(FRT_ERR_ARG, "frt_canvas_device_ptr: null argument");
    }
    *device_ptr = sc->canvas;
    return FRT_OK;
}

extern "C" int
frt_light_points_checksum(frt_scene *sc, int64_t first_point, int64_t n_points, uint64_t *sum)
{
    if (sc == nullptr || sum == nullptr || first_point < 0 || n_points < 0) {
        return frt_set_error(FRT_ERR_ARG, "frt_light_points_checksum: bad argument");
    }
    CK(cudaSetDevice(sc->device));
    CK(cudaStreamSynchronize(sc->upload_stream));
    unsigned long long *d = nullptr, h = 0;
    CK(cudaMalloc(&d, sizeof(h)));
    cudaError_t e = cudaMemsetAsync(d, 0, sizeof(h), sc->stream);
    if (e == cudaSuccess && n_points) {
        k_points_checksum<<<sc->sm_count * 8, 256, 0, sc->stream>>>(sc->S.lpoints, 3ull * (unsigned long long)first_point, 3ull * (unsigned long long)n_points, d);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(&h, d, sizeof(h), cudaMemcpyDeviceToHost, sc->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(sc->stream);
    cudaFree(d);
    if (e != cudaSuccess) {
        return frt_set_error(FRT_ERR_CUDA, "frt_light_points_checksum: %s", cudaGetErrorString(e));
    }
    *sum = h;
    return FRT_OK;
}

/* ---- output encode: write_ppm_file / construct_ppm (canvas.c:150-328) on the device ------------------------------ */

static size_t
ppm16_header(char *buf, size_t cap, int width, int height)
{
    return (size_t)snprintf(buf, cap, "P6\n%zu %zu\n65535\n", (size_t)width, (size_t)height); /* canvas.c:167 */
}

extern "C" size_t
frt_ppm16_size(int width, int height)
{
    char hdr[32];
    if (width <= 0 || height <= 0) {
        return 0;
    }
    return ppm16_header(hdr, sizeof(hdr), width, height) + (size_t)width * height * 6 + 1;
}

/* dev_canvas: width * height * 4 doubles on the current device; out: host buffer of frt_ppm16_size bytes */
static int
encode_ppm16(const double *dev_canvas, int width, int height, int use_scaling, cudaStream_t s, unsigned char *out, size_t out_cap,
             size_t *out_len, double *encode_ms)
{
    const size_t need = frt_ppm16_size(width, height);
    if (out == nullptr || out_cap < need || need == 0) {
        return frt_set_error(FRT_ERR_ARG, "PPM buffer of %zu bytes, %zu needed", out_cap, need);
    }
    const size_t n = (size_t)width * height;
    double *maxes = nullptr;
    unsigned char *data = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    float ms = 0.f;
    char hdr[32];
    const size_t hl = ppm16_header(hdr, sizeof(hdr), width, height);
    int sms = 148, dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    const int blocks = (int)std::min<size_t>((n + 255) / 256, (size_t)sms * 8);
    cudaError_t e = cudaMalloc(&maxes, 6 * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&data, n * 6);
    if (e == cudaSuccess) e = cudaEventCreate(&e0);
    if (e == cudaSuccess) e = cudaEventCreate(&e1);
    if (e == cudaSuccess) e = cudaEventRecord(e0, s);
    if (e == cudaSuccess) e = cudaMemsetAsync(maxes, 0, 6 * sizeof(double), s);
    if (e == cudaSuccess) {
        k_ppm_rgb_max<<<blocks, 256, 0, s>>>(dev_canvas, n, maxes);
        k_ppm_srgb_max<<<blocks, 256, 0, s>>>(dev_canvas, n, maxes);
        k_ppm_encode<<<blocks, 256, 0, s>>>(dev_canvas, n, maxes, use_scaling, data);
        e = cudaEventRecord(e1, s);
    }
    if (e == cudaSuccess) {
        memcpy(out, hdr, hl);
        e = cudaMemcpyAsync(out + hl, data, n * 6, cudaMemcpyDeviceToHost, s);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e == cudaSuccess) {
        out[hl + n * 6] = '\n'; /* canvas.c:298 */
        cudaEventElapsedTime(&ms, e0, e1);
    }
    /* every path releases what it took */
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    cudaFree(maxes);
    cudaFree(data);
    if (e != cudaSuccess) {
        return frt_set_error(FRT_ERR_CUDA, "PPM encode: %s", cudaGetErrorString(e));
    }
    if (out_len) *out_len = need;
    if (encode_ms) *encode_ms = ms;
    return FRT_OK;
}

extern "C" int
frt_canvas_encode_ppm16(frt_scene *sc, int use_scaling, unsigned char *out, size_t out_cap, size_t *out_len, double *encode_ms)
{
    if (sc == nullptr) {
        return frt_set_error(FRT_ERR_ARG, "frt_canvas_encode_ppm16: null scene");
    }
    CK(cudaSetDevice(sc->device));
    return encode_ppm16(sc->canvas, sc->C.hsize, sc->C.vsize, use_scaling, sc->stream, out, out_cap, out_len, encode_ms);
}

extern "C" int
frt_encode_ppm16(const double *canvas_rgba, int width, int height, int use_scaling, int device, unsigned char *out, size_t out_cap,
                 size_t *out_len, double *encode_ms)
{
    if (canvas_rgba == nullptr || width <= 0 || height <= 0) {
        return frt_set_error(FRT_ERR_ARG, "frt_encode_ppm16: empty canvas");
    }
    CK(cudaSetDevice(device));
    double *dev = nullptr;
    const size_t bytes = (size_t)width * height * 4 * sizeof(double);
    CK(cudaMalloc(&dev, bytes));
    cudaError_t e = cudaMemcpy(dev, canvas_rgba, bytes, cudaMemcpyHostToDevice);
    int rc = FRT_OK;
    if (e != cudaSuccess) {
        rc = frt_set_error(FRT_ERR_CUDA, "frt_encode_ppm16: %s", cudaGetErrorString(e));
    } else {
        rc = encode_ppm16(dev, width, height, use_scaling, 0, out, out_cap, out_len, encode_ms);
    }
    cudaFree(dev);
    return rc;
}

static int
frt_env_int(const char *name, int dflt)
{
    const char *v = getenv(name);
    return (v != nullptr && *v) ? atoi(v) : dflt;
}

static int launch_knn(frt_scene *sc, cudaStream_t s, const GIParams &G, GQuery *q, unsigned int *q_n, unsigned int q_cap, double *acc_amb,
                      double *acc_fg, int *found, int mode, int *launches);

/* The rows this rank owns leave for the caller's canvas (host memory, device memory, a peer's shared buffer) behind the frame's
 * last kernel: bracketed by ev[2] / ev[3], not waited for here.  The caller's other rows stay untouched. */
static int
enqueue_owned_rows(frt_scene *sc, const frt_render_cfg *cfg, double *canvas_rgba)
{
    CK(cudaEventRecord(sc->ev[2], sc->stream));
    const int world = cfg->world > 0 ? cfg->world : 1;
    const int rpb = cfg->rows_per_block > 0 ? cfg->rows_per_block : 4;
    const size_t row_bytes = (size_t)sc->C.hsize * 4 * sizeof(double);
    if (world == 1) {
        CK(cudaMemcpyAsync(canvas_rgba, sc->canvas, row_bytes * sc->C.vsize, cudaMemcpyDefault, sc->stream));
    } else {
        /* the owned row blocks are equally spaced runs of the canvas: one strided copy for the whole blocks (a block of
         * rpb rows every world * rpb rows), one plain copy for a last, shorter block */
        const int first = cfg->rank * rpb;
        const int full_blocks = first < sc->C.vsize ? (sc->C.vsize - first) / (world * rpb) + (((sc->C.vsize - first) % (world * rpb)) >= rpb ? 1 : 0) : 0;
        if (full_blocks > 0) {
            CK(cudaMemcpy2DAsync((char *)canvas_rgba + row_bytes * first, row_bytes * world * rpb, (char *)sc->canvas + row_bytes * first,
                                 row_bytes * world * rpb, row_bytes * rpb, (size_t)full_blocks, cudaMemcpyDefault, sc->stream));
        }
        const int y0 = first + full_blocks * world * rpb;
        if (y0 < sc->C.vsize) {
            CK(cudaMemcpyAsync((char *)canvas_rgba + row_bytes * y0, (char *)sc->canvas + row_bytes * y0, row_bytes * (sc->C.vsize - y0),
                               cudaMemcpyDefault, sc->stream));
        }
    }
    CK(cudaEventRecord(sc->ev[3], sc->stream));
    return FRT_OK;
}

static int
render_once(frt_scene *sc, const frt_render_cfg *cfg, frt_stats *st, unsigned int chunk_samples, unsigned int cap_factor, double *canvas_rgba)
{
    const frt_config &g = sc->cfg;
    DCamera C = sc->C;
    if (cfg->usteps > 0 && cfg->vsteps > 0 && (cfg->usteps != C.usteps || cfg->vsteps != C.vsteps)) {
        if (cfg->usteps != sc->samples_u || cfg->vsteps != sc->samples_v) {
            std::vector<double> table;
            cmj_table_no_jitter(cfg->usteps, cfg->vsteps, table);
            double *dtab = nullptr;
            CK(scene_alloc(sc, (void **)&dtab, table.size() * sizeof(double)));
            CK(cudaMemcpyAsync(dtab, table.data(), table.size() * sizeof(double), cudaMemcpyHostToDevice, sc->stream)); /* pageable: staged before the call returns */
            sc->samples = dtab;
            sc->samples_u = cfg->usteps;
            sc->samples_v = cfg->vsteps;
        }
        C.usteps = cfg->usteps;
        C.vsteps = cfg->vsteps;
        C.samples = sc->samples;
    }
    if (cfg->jitter >= 0) {
        C.jitter = cfg->jitter;
    }
    if (C.jitter && C.usteps * C.vsteps > FRT_MAX_SPP) {
        return frt_set_error(FRT_ERR_ARG, "jittered supersampling supports at most %d samples per pixel", FRT_MAX_SPP);
    }

    FrameParams F{};
    F.world = cfg->world > 0 ? cfg->world : 1;
    F.rank = cfg->rank;
    if (F.rank < 0 || F.rank >= F.world) {
        return frt_set_error(FRT_ERR_ARG, "rank %d of world %d", F.rank, F.world);
    }
    F.rows_per_block = cfg->rows_per_block > 0 ? cfg->rows_per_block : 4;
    F.n_owned_rows = count_owned_rows(C.vsize, F.rank, F.world, F.rows_per_block);
    F.flags = cfg->flags;
    F.path_length = g.di_path_length;
    F.include_direct = g.include_direct;
    F.use_ambient = g.di_include_ambient;
    F.use_diffuse = g.di_include_diffuse;
    F.use_spec_highlight = g.di_include_specular_highlight;
    F.include_specular = g.di_include_specular;
    F.seed = mix64(cfg->seed);
    /* shade_hit's GI block (renderer.c:737): use_gi = include_global || visualize_photon_map (renderer.c:62) */
    F.use_gi = (g.include_global || g.visualize_photon_map) ? 1 : 0;
    GIParams G{};
    if (F.use_gi) {
        if (!sc->pm_ready) {
            if (g.gi_photon_count <= 0) {
                F.use_gi = 0; /* the generated main() skips photon tracing when photon_count is 0 (yaml_parser.py:201-215) */
            } else {
                return frt_set_error(FRT_ERR_ARG, "the scene asks for global illumination but no photon map was built: "
                                                  "call frt_photons_emit() and frt_photons_finish() first");
            }
        }
        G.usteps = g.gi_usteps > 0 ? g.gi_usteps : 1;
        G.vsteps = g.gi_vsteps > 0 ? g.gi_vsteps : 1;
        G.n_photons = g.gi_irradiance_estimate_num;
        G.radius = (float)g.gi_irradiance_estimate_radius;
        G.cone_k = (float)g.gi_irradiance_estimate_cone_filter_k;
        G.visualize = g.visualize_photon_map;
        G.use_caustics = g.gi_include_caustics;
        G.use_final_gather = g.gi_include_final_gather;
        G.seed = F.seed;
        if (G.usteps * G.vsteps > FRT_MAX_SPP && G.usteps != G.vsteps) {
            return frt_set_error(FRT_ERR_ARG, "a non-square final-gather grid supports at most %d cells", FRT_MAX_SPP);
        }
        if (G.n_photons <= 0 || G.n_photons > FRT_KNN_CAP / 2) {
            return frt_set_error(FRT_ERR_ARG, "irradiance-estimate-num %d is outside 1..%d", G.n_photons, FRT_KNN_CAP / 2);
        }
    }

    const unsigned int spp = (unsigned int)(C.usteps * C.vsteps);
    const unsigned long long total = (unsigned long long)F.n_owned_rows * C.hsize * spp;
    if (total >= 0xfffffff0ULL) {
        return frt_set_error(FRT_ERR_ARG, "frame has too many samples for 32-bit sample ids");
    }
    unsigned int chunk = (unsigned int)std::min<unsigned long long>(std::max<unsigned long long>(total, 1), chunk_samples);
    unsigned long long want = (unsigned long long)chunk * cap_factor;
    unsigned int capacity = (unsigned int)std::min<unsigned long long>(want, (unsigned long long)FRT_PEND_HIT_MASK + 1ull); /* hit ids share a word with flags */
    int rc = ensure_frame_buffers(sc, capacity);
    if (rc != FRT_OK) {
        return rc;
    }
    F.capacity = sc->capacity;
    if (F.use_gi) {
        rc = ensure_gi_buffers(sc);
        if (rc != FRT_OK) {
            return rc;
        }
    }

    cudaStream_t s = sc->stream;
    size_t cbytes = (size_t)C.hsize * C.vsize * 4 * sizeof(double);
    const int sm_blocks = sc->sm_count;
    unsigned long long launches = 0, light_launches = 0;
    Counters totals{};
    /* Per-stage device times: every launch of the dominant shadow-ray kernel (FRT_ST_SHADOW_RAY) and of the per-hit shaft
     * kernels is bracketed by a pair of events on the render stream, the other stages too under FRT_FLAG_STAGE_TIMES; the
     * pairs are read after the frame's last event -- nothing here waits for the device. */
    std::vector<int> span_stage;
    const bool all_stages = (F.flags & FRT_FLAG_STAGE_TIMES) != 0;
    bool tick_failed = false;
    auto tick = [&](int stage) -> int {
        if (!all_stages && stage != FRT_ST_SHADOW_RAY && stage != FRT_ST_SHADOW_SHAFT) {
            return -1;
        }
        const size_t i = span_stage.size();
        while (sc->light_ev.size() < 2 * (i + 1)) {
            cudaEvent_t e = nullptr;
            if (pool_event(sc->device, true, &e) != cudaSuccess) {
                tick_failed = true;
                return -1;
            }
            sc->light_ev.push_back(e);
        }
        span_stage.push_back(stage);
        if (cudaEventRecord(sc->light_ev[2 * i], s) != cudaSuccess) tick_failed = true;
        return (int)i;
    };
    auto tock = [&](int i) {
        if (i >= 0 && cudaEventRecord(sc->light_ev[2 * (size_t)i + 1], s) != cudaSuccess) tick_failed = true;
    };

    CK(cudaEventRecord(sc->ev[0], s));
    bool frame_end_recorded = false;
    CK(cudaMemsetAsync(sc->canvas, 0, cbytes, s));

    const std::vector<int> &gw = sc->light_gw; /* lanes per hit in k_light_pre / k_light_final, chosen per light at upload */
    /* launch-shape overrides for experiments, read once per frame */
    auto env_int = [](const char *name, int dflt) { const char *v = getenv(name); return (v != nullptr && *v) ? atoi(v) : dflt; };
    const int mblocks = env_int("FRT_MESH_BLOCKS", sm_blocks * 8);
    const int m_inner = env_int("FRT_MESH_INNER_STEPS", FRT_MESH_INNER);
    const int m_refill = env_int("FRT_MESH_REFILL_MIN", FRT_MESH_REFILL);
    const int sblocks = env_int("FRT_SHADOW_BLOCKS", sm_blocks * 48);
    const int eblocks = env_int("FRT_ENTRY_BLOCKS", sm_blocks * 64);
    const int entry_kernel = env_int("FRT_ENTRY_KERNEL", 0); /* 1: one warp per pending entry (k_shadow_entry, A/B measurements: 6.5 vs 5.9 ms) */
    const bool debug_nodes = getenv("FRT_DEBUG_NODES") != nullptr;
    const int knn_mode = env_int("FRT_KNN_MODE", 1); /* 1: requests sorted by cell, a lane per request (k_knn_cell); 0: a warp per request (k_knn); 2: k_knn_list */

    for (unsigned long long first = 0; first < total; first += chunk) {
        unsigned int n = (unsigned int)std::min<unsigned long long>(chunk, total - first);
        CK(cudaMemsetAsync(sc->cnt, 0, sizeof(Counters), s));
        int rg_blocks = (int)std::min<unsigned int>((n + 255) / 256, sm_blocks * 16);
        int tk = tick(FRT_ST_RAYGEN);
        k_raygen<<<rg_blocks, 256, 0, s>>>(C, F, sc->q[0], sc->cnt, (unsigned int)first, n);
        tock(tk);
        ++launches;
        /* The hits of EVERY level go into one list and meet the lights in one stage (merge_levels): the reflected /
         * refracted rays of a Cornell frame are 0.5 % of its hits, but lit on their own they cost six more kernels whose
         * duration is mostly launch, ramp and tail -- a sixth of the frame of one of eight GPUs.  Extend / shade run level by
         * level (each waits for the count the previous k_shade leaves), the light stage once at the end.  With global
         * illumination the stages stay per level (the final gather's random streams are keyed on the level). */
        const bool merge_levels = !F.use_gi && sc->h_nrays != nullptr && env_int("FRT_MERGE_LEVELS", 1) != 0;
        for (int level = 0; level <= F.path_length; ++level) {
            bool trace = true;
            if (level > 0 && sc->h_nrays != nullptr) {
                /* the count was copied right behind the previous k_shade; with a light stage per level the stream is still busy
                 * with it, so this wait costs nothing -- and the empty tail of the recursion (Cornell: levels 2..5, ~9 launches
                 * of full grids each) is never enqueued */
                CK(cudaEventSynchronize(sc->nrays_ev));
                if (sc->h_nrays[0] == 0) {
                    if (!merge_levels) {
                        break;
                    }
                    trace = false; /* nothing left to trace: on to the one light stage */
                }
            }
            RayQ &qi = sc->q[level & 1];
            RayQ &qo = sc->q[(level + 1) & 1];
            const int rec_level = merge_levels ? 0 : level;
            if (trace) {
            /* level 0 has n rays; deeper levels read their count on the device: size the grid for the worst case
             * the level can hold, but never more than a few waves */
            int ex_blocks = sm_blocks * 8;
            tk = tick(FRT_ST_EXTEND);
            if (sc->boxes_and_balls) { /* the scene holds cubes, spheres, planes, CSGs and groups only */
                k_extend<FRT_PRIMS_BOXES_AND_BALLS><<<ex_blocks, 256, 0, s>>>(sc->S, sc->SF, qi, sc->hq, sc->cnt, level, F.capacity);
            } else {
                k_extend<FRT_PRIMS_ALL><<<ex_blocks, 256, 0, s>>>(sc->S, sc->SF, qi, sc->hq, sc->cnt, level, F.capacity);
            }
            tock(tk);
            tk = tick(FRT_ST_SHADE);
            if (sc->has_maps) {
                k_shade<true, true><<<sm_blocks * 8, 128, 0, s>>>(sc->S, F, qi, sc->hq, qo, sc->recs, sc->cnt, level, rec_level);
            } else if (sc->has_refraction) {
                k_shade<false, true><<<sm_blocks * 8, 128, 0, s>>>(sc->S, F, qi, sc->hq, qo, sc->recs, sc->cnt, level, rec_level);
            } else {
                k_shade<false, false><<<sm_blocks * 8, 128, 0, s>>>(sc->S, F, qi, sc->hq, qo, sc->recs, sc->cnt, level, rec_level);
            }
            tock(tk);
            launches += 2;
            if (sc->h_nrays != nullptr && level < F.path_length) {
                CK(cudaMemcpyAsync(sc->h_nrays, &sc->cnt->n_rays[level + 1], sizeof(unsigned int), cudaMemcpyDeviceToHost, s));
                CK(cudaEventRecord(sc->nrays_ev, s));
            }
            }
            if (merge_levels && trace && level < F.path_length) {
                continue; /* the next level first */
            }
            const int hl = rec_level; /* the level whose hit list the light stage reads */
            if (F.use_gi) {
                CK(cudaMemsetAsync(sc->acc_amb, 0, sizeof(double) * 3 * (size_t)sc->acc_cap, s));
                CK(cudaMemsetAsync(sc->acc_fg, 0, sizeof(double) * 3 * (size_t)sc->acc_cap, s));
            }
            if (F.include_direct && sc->mesh_mode && sc->S.n_roots == 1 && sc->S.n_lights > 1 && !(F.flags & FRT_FLAG_F64_SHADOW)) {
                /* mesh scene, several lights: their shadow rays share launches of k_shadow_mesh (FRT_MESH_LIGHTS at a time) */
                const size_t need = (size_t)sc->capacity * std::min(sc->S.n_lights, FRT_MESH_LIGHTS);
                if (sc->ltmp_multi_cap < need) {
                    void *p = nullptr;
                    CK(scene_alloc(sc, &p, need * sizeof(LightTmp)));
                    sc->ltmp_multi = (LightTmp *)p;
                    sc->ltmp_multi_cap = need;
                }
                const int blocks = sm_blocks * 8;
                const bool count = (F.flags & FRT_FLAG_COUNT_RAYS) != 0;
                for (int l0 = 0; l0 < sc->S.n_lights; l0 += FRT_MESH_LIGHTS) {
                    const int nl = std::min(FRT_MESH_LIGHTS, sc->S.n_lights - l0);
                    tk = tick(FRT_ST_LIGHT_PRE);
                    for (int k = 0; k < nl; ++k) {
                        launch_light_pre(sc, F, blocks, hl, l0 + k, gw[l0 + k], sc->ltmp_multi + (size_t)k * sc->capacity);
                    }
                    tock(tk);
                    if (wait_for_upload(sc, s) != FRT_OK) return FRT_ERR_CUDA;
                    CK(cudaMemsetAsync(&sc->cnt->mesh_next, 0, sizeof(unsigned long long), s));
                    tk = tick(FRT_ST_SHADOW_RAY);
                    if (count) {
                        launch_shadow_mesh<true>(sc, mblocks, s, F, sc->ltmp_multi, sc->capacity, hl, l0, nl, m_inner, m_refill);
                    } else {
                        launch_shadow_mesh<false>(sc, mblocks, s, F, sc->ltmp_multi, sc->capacity, hl, l0, nl, m_inner, m_refill);
                    }
                    tock(tk);
                    tk = tick(FRT_ST_LIGHT_FINAL);
                    for (int k = 0; k < nl; ++k) {
                        if (F.flags & FRT_FLAG_F64_SHADING) {
                            launch_light_final<double>(sc, F, blocks, hl, l0 + k, gw[l0 + k], sc->ltmp_multi + (size_t)k * sc->capacity);
                        } else {
                            launch_light_final<float>(sc, F, blocks, hl, l0 + k, gw[l0 + k], sc->ltmp_multi + (size_t)k * sc->capacity);
                        }
                    }
                    tock(tk);
                    launches += 1 + 2 * nl;
                    ++light_launches;
                }
            } else if (F.include_direct) {
                for (int li = 0; li < sc->S.n_lights; ++li) {
                    const int blocks = sm_blocks * 8;
                    tk = tick(FRT_ST_LIGHT_PRE);
                    launch_light_pre(sc, F, blocks, hl, li, gw[li], sc->ltmp);
                    tock(tk);
                    if (wait_for_upload(sc, s) != FRT_OK) return FRT_ERR_CUDA;
                    const bool count = (F.flags & FRT_FLAG_COUNT_RAYS) != 0;
                    if (F.flags & FRT_FLAG_F64_SHADOW) {
                        tk = tick(FRT_ST_SHADOW_RAY);
                        if (count) {
                            k_shadow_exact<true, true><<<blocks, 256, 0, s>>>(sc->S, sc->SF, F, sc->recs, sc->ltmp, sc->cnt, hl, li, sc->dq, sc->dq_cap);
                        } else {
                            k_shadow_exact<false, true><<<blocks, 256, 0, s>>>(sc->S, sc->SF, F, sc->recs, sc->ltmp, sc->cnt, hl, li, sc->dq, sc->dq_cap);
                        }
                        tock(tk);
                        launches += 1;
                    } else if (sc->mesh_mode && sc->S.n_roots == 1) {
                        CK(cudaMemsetAsync(&sc->cnt->mesh_next, 0, sizeof(unsigned long long), s));
                        tk = tick(FRT_ST_SHADOW_RAY);
                        if (count) {
                            launch_shadow_mesh<true>(sc, mblocks, s, F, sc->ltmp, 0, hl, li, 1, m_inner, m_refill);
                        } else {
                            launch_shadow_mesh<false>(sc, mblocks, s, F, sc->ltmp, 0, hl, li, 1, m_inner, m_refill);
                        }
                        tock(tk);
                        launches += 1;
                    } else {
                        const size_t f32_smem = (size_t)sc->S.n_nodes * 48 + (size_t)sc->SF.n_csg_prog * 4 <= 32768
                                                    ? (size_t)sc->S.n_nodes * 48 + (size_t)sc->SF.n_csg_prog * 4 : 0;
                        /* many short grid-stride trips balance the uneven per-ray work better than 8 CTAs per SM (measured: 35.0 -> 33.5 ms) */
                        CK(cudaMemsetAsync(&sc->cnt->n_deferred, 0, 2 * sizeof(unsigned int), s)); /* n_deferred, n_pending */
                        /* per hit: every shadow ray at once where the shaft's intervals separate (small trees, area lights) */
                        const int bulk_on = sc->S.n_roots == 1 && sc->S.n_nodes <= 32 && sc->light_ns[li] >= 4 &&
                                            !(F.flags & (FRT_FLAG_NO_SHAFT | FRT_FLAG_NO_BULK));
                        const int split_on = bulk_on && !(F.flags & FRT_FLAG_NO_SPLIT);
                        const unsigned int pend_cap = sc->capacity * 4u;
                        unsigned int *retry = reinterpret_cast<unsigned int *>(sc->dq); /* free until k_shadow_f32 defers rays */
                        unsigned int *pprog = sc->pending + (size_t)sc->capacity * 4; /* one program word per pending entry */
#define FRT_SHADOW_STAGE(M)                                                                                                                       \
    do {                                                                                                                                          \
        tk = tick(FRT_ST_SHADOW_SHAFT);                                                                                                           \
        if (F.flags & FRT_FLAG_F64_SHAFT) {                                                                                                       \
            k_shadow_bulk<M, double><<<blocks, 256, 0, s>>>(sc->S, sc->SF, F, sc->recs, sc->ltmp, sc->cnt, hl, li, sc->pending, pprog, retry, bulk_on, split_on); \
        } else {                                                                                                                                  \
            k_shadow_bulk<M, float><<<blocks, 256, 0, s>>>(sc->S, sc->SF, F, sc->recs, sc->ltmp, sc->cnt, hl, li, sc->pending, pprog, retry, bulk_on, split_on); \
        }                                                                                                                                         \
        if (split_on) {                                                                                                                           \
            if (F.flags & FRT_FLAG_F64_SHAFT) {                                                                                                   \
                k_shadow_quad<M, double><<<blocks, 256, 0, s>>>(sc->S, sc->SF, F, sc->recs, sc->ltmp, sc->cnt, li, sc->pending, pprog, retry);    \
            } else {                                                                                                                              \
                k_shadow_quad<M, float><<<blocks, 256, 0, s>>>(sc->S, sc->SF, F, sc->recs, sc->ltmp, sc->cnt, li, sc->pending, pprog, retry);     \
            }                                                                                                                                     \
            CK(cudaMemsetAsync(&sc->cnt->n_deferred, 0, sizeof(unsigned int), s));                                                                \
            launches += 1;                                                                                                                        \
        }                                                                                                                                         \
        tock(tk);                                                                                                                                 \
        tk = tick(FRT_ST_SHADOW_RAY);                                                                                                             \
        if (entry_kernel) {                                                                                                                       \
            k_shadow_entry<M><<<eblocks, 128, f32_smem, s>>>(sc->S, sc->SF, F, sc->recs, sc->ltmp, sc->cnt, sc->pending, pprog, bulk_on, pend_cap, split_on, li, \
                                                             sc->dq, sc->dq_cap, f32_smem != 0);                                                  \
        } else {                                                                                                                                  \
            k_shadow_f32<M><<<sblocks, 256, f32_smem, s>>>(sc->S, sc->SF, F, sc->recs, sc->ltmp, sc->cnt, sc->pending, pprog, bulk_on, pend_cap, split_on, li, sc->dq, \
                                                           sc->dq_cap, f32_smem != 0);                                                            \
        }                                                                                                                                         \
        tock(tk);                                                                                                                                 \
    } while (0)
                        if (F.flags & FRT_FLAG_VERIFY_F32) {
                            FRT_SHADOW_STAGE(2);
                        } else if (count) {
                            FRT_SHADOW_STAGE(1);
                        } else {
                            FRT_SHADOW_STAGE(0);
                        }
#undef FRT_SHADOW_STAGE
                        tk = tick(FRT_ST_SHADOW_EXACT);
                        if (count) {
                            k_shadow_exact<true, false><<<blocks, 256, 0, s>>>(sc->S, sc->SF, F, sc->recs, sc->ltmp, sc->cnt, hl, li, sc->dq, sc->dq_cap);
                        } else {
                            k_shadow_exact<false, false><<<blocks, 256, 0, s>>>(sc->S, sc->SF, F, sc->recs, sc->ltmp, sc->cnt, hl, li, sc->dq, sc->dq_cap);
                        }
                        tock(tk);
                        launches += 3;
                    }
                    tk = tick(FRT_ST_LIGHT_FINAL);
                    if (F.flags & FRT_FLAG_F64_SHADING) {
                        launch_light_final<double>(sc, F, blocks, hl, li, gw[li], sc->ltmp);
                    } else {
                        launch_light_final<float>(sc, F, blocks, hl, li, gw[li], sc->ltmp);
                    }
                    tock(tk);
                    launches += 2;
                    ++light_launches;
                }
            }
            if (merge_levels) {
                break; /* every level has been traced and lit */
            }
            if (F.use_gi) {
                /* the GI block of shade_hit for the hits of this level, in batches that fit the request queue */
                unsigned int n_hits = 0;
                CK(cudaMemcpyAsync(&n_hits, &sc->cnt->n_hits[level], sizeof(unsigned int), cudaMemcpyDeviceToHost, s));
                CK(cudaStreamSynchronize(s));
                n_hits = std::min(n_hits, F.capacity);
                const unsigned int cells = (unsigned int)(G.usteps * G.vsteps);
                const unsigned int per_hit = (G.use_final_gather ? cells : 0u) + (G.use_caustics ? 1u : 0u) + (G.visualize ? 1u : 0u);
                const unsigned int batch = per_hit ? std::max(1u, sc->gq_cap / per_hit) : n_hits;
                for (unsigned int first = 0; first < n_hits && per_hit; first += batch) {
                    const unsigned int nb = std::min(batch, n_hits - first);
                    CK(cudaMemsetAsync(sc->gq_n, 0, sizeof(unsigned int), s));
                    tk = tick(FRT_ST_GI_TRACE);
                    if (G.use_caustics || G.visualize) {
                        k_gi_points<<<sm_blocks * 4, 256, 0, s>>>(F, G, sc->recs, first, nb, sc->gq, sc->gq_n, sc->gq_cap, sc->cnt,
                                                                  G.use_caustics, G.visualize);
                        ++launches;
                    }
                    if (G.use_final_gather) {
#define FRT_FG(PR, MP) k_fg_trace<PR, MP><<<sm_blocks * 16, 128, 0, s>>>(sc->S, sc->SF, F, G, sc->recs, first, nb, sc->gq, sc->gq_n, sc->gq_cap, sc->cnt, level)
                        if (sc->boxes_and_balls && !sc->has_maps) {
                            FRT_FG(FRT_PRIMS_BOXES_AND_BALLS, false);
                        } else if (sc->boxes_and_balls) {
                            FRT_FG(FRT_PRIMS_BOXES_AND_BALLS, true);
                        } else if (!sc->has_maps) {
                            FRT_FG(FRT_PRIMS_ALL, false);
                        } else {
                            FRT_FG(FRT_PRIMS_ALL, true);
                        }
#undef FRT_FG
                        ++launches;
                    }
                    tock(tk);
                    tk = tick(FRT_ST_KNN);
                    {
                        int knn_launches = 0;
                        const int rc_knn = launch_knn(sc, s, G, sc->gq, sc->gq_n, sc->gq_cap, sc->acc_amb, sc->acc_fg, nullptr, knn_mode, &knn_launches);
                        if (rc_knn != FRT_OK) {
                            return rc_knn;
                        }
                        launches += knn_launches;
                    }
                    tock(tk);
                }
                tk = tick(FRT_ST_GI_RESOLVE);
                k_gi_resolve<<<sm_blocks * 8, 256, 0, s>>>(F, G, sc->recs, sc->acc_amb, sc->acc_fg, sc->canvas, sc->cnt, level);
                tock(tk);
                ++launches;
            }
        }
        if (canvas_rgba != nullptr && first + chunk >= total) {
            /* the frame's rows travel behind its last kernel, in front of the one wait of the frame (the counters below): a
             * frame that overflowed its queues is re-run and copied again.  The frame's own time ends here. */
            CK(cudaEventRecord(sc->ev[1], s));
            frame_end_recorded = true;
            const int rc_rows = enqueue_owned_rows(sc, cfg, canvas_rgba);
            if (rc_rows != FRT_OK) {
                return rc_rows;
            }
        }
        Counters hc;
        CK(cudaMemcpyAsync(&hc, sc->cnt, sizeof(Counters), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        totals.overflow_queue |= hc.overflow_queue;
        totals.overflow_csg |= hc.overflow_csg;
        totals.rays_secondary += hc.rays_secondary;
        totals.rays_shadow += hc.rays_shadow;
        totals.hits_shaded += hc.hits_shaded;
        totals.shadow_nodes += hc.shadow_nodes;
        totals.light_flops += hc.light_flops;
        totals.deferred_total += hc.deferred_total;
        totals.f32_mismatch += hc.f32_mismatch;
        totals.rays_gather += hc.rays_gather;
        totals.rays_bulk += hc.rays_bulk;
        totals.rays_per_ray += hc.rays_per_ray;
        for (int k = 0; k < 10; ++k) {
            totals.undecided_reason[k] += hc.undecided_reason[k];
        }
        if (debug_nodes) {
            for (int k = 0; k < 32; ++k) {
                if (hc.undecided_node[k]) fprintf(stderr, "undecided at node %d: %llu\n", k, hc.undecided_node[k]);
            }
            const char *cls[3] = { "straight-line", "tail, general walk", "no tail" };
            for (int c = 0; c < 3; ++c) {
                for (int k = 0; k < 32; ++k) {
                    if (hc.entry_node[c][k]) fprintf(stderr, "pending entries (%s) starting at node %d: %llu\n", cls[c], k, hc.entry_node[c][k]);
                }
            }
        }
        if (hc.overflow_queue) {
            break;
        }
    }
    if (!frame_end_recorded) {
        CK(cudaEventRecord(sc->ev[1], s));
    }
    CK(cudaEventSynchronize(sc->ev[1]));
    CK(cudaGetLastError());
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, sc->ev[0], sc->ev[1]));
    if (tick_failed) {
        return frt_set_error(FRT_ERR_CUDA, "cudaEventCreate / cudaEventRecord of a stage timer failed");
    }
    double stage_ms[FRT_ST_COUNT] = { 0 };
    unsigned long long stage_launches[FRT_ST_COUNT] = { 0 };
    for (size_t k = 0; k < span_stage.size(); ++k) {
        float lms = 0.f;
        CK(cudaEventElapsedTime(&lms, sc->light_ev[2 * k], sc->light_ev[2 * k + 1]));
        stage_ms[span_stage[k]] += lms;
        stage_launches[span_stage[k]] += 1;
    }

    if (st != nullptr) {
        st->frame_ms = ms;
        st->light_ms = stage_ms[FRT_ST_SHADOW_RAY];
        for (int k = 0; k < FRT_ST_COUNT; ++k) {
            st->stage_ms[k] = stage_ms[k];
        }
        st->shadow_ray_launches = stage_launches[FRT_ST_SHADOW_RAY];
        st->shadow_rays_traced = totals.rays_per_ray;
        st->rays_primary = total;
        st->rays_secondary = totals.rays_secondary;
        st->rays_shadow = totals.rays_shadow;
        st->hits_shaded = totals.hits_shaded;
        st->shadow_nodes = totals.shadow_nodes;
        st->light_flops = totals.light_flops;
        st->shadow_deferred = totals.deferred_total;
        st->shadow_mismatch = totals.f32_mismatch;
        st->rays_gather = totals.rays_gather;
        for (int k = 0; k < 10; ++k) {
            st->shadow_reasons[k] = totals.undecided_reason[k];
        }
        st->shadow_reasons[0] = totals.rays_bulk; /* slot 0 is no deferral reason: shadow rays decided per hit (k_shadow_bulk) */
        st->kernel_launches = launches;
        st->light_launches = light_launches;
        st->rows_rendered = F.n_owned_rows;
    }
    if (totals.overflow_csg) {
        return frt_set_error(FRT_ERR_OVERFLOW, "a ray met more than %d CSG crossings (or 8 nested CSG levels)", FRT_CSG_CAP);
    }
    if (totals.overflow_queue) {
        return -1; /* caller retries with smaller chunks */
    }
    return FRT_OK;
}

static int
gen_check(frt_scene *sc, cudaStream_t s)
{
    if (!sc->gen_check_pending) {
        return FRT_OK;
    }
    unsigned int bad = 0;
    CK(cudaMemcpyAsync(&bad, sc->gen_mismatch, sizeof(bad), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    sc->gen_check_pending = false;
    sc->gen_failed = bad != 0;
    if (bad) {
        return frt_set_error(FRT_ERR_MISMATCH, "a light-sample set rebuilt on the device differs from the caller's in %u words "
                                               "(another drand48 state, or another sampler): upload the host cache instead", bad);
    }
    return FRT_OK;
}

extern "C" int
frt_scene_gen_status(frt_scene *sc)
{
    if (sc == nullptr) {
        return frt_set_error(FRT_ERR_ARG, "frt_scene_gen_status: null scene");
    }
    CK(cudaSetDevice(sc->device));
    if (sc->gen_failed) {
        return frt_set_error(FRT_ERR_MISMATCH, "a light-sample set rebuilt on the device differs from the caller's");
    }
    return gen_check(sc, sc->upload_stream);
}

extern "C" int
frt_render(frt_scene *sc, const frt_render_cfg *cfg, double *canvas_rgba, frt_stats *stats)
{
    if (sc == nullptr || cfg == nullptr) {
        return frt_set_error(FRT_ERR_ARG, "frt_render: null argument");
    }
    CK(cudaSetDevice(sc->device));
    if (sc->gen_failed) {
        return frt_set_error(FRT_ERR_MISMATCH, "a light-sample set rebuilt on the device differs from the caller's");
    }
    frt_stats st;
    memset(&st, 0, sizeof(st));

    /* Queue sizing: a level of the wavefront holds at most chunk * cap_factor rays.  Reflective + refractive
     * surfaces can double the ray count per level, so on overflow the frame is re-run with a larger factor (while
     * it fits in a quarter of the free HBM) and then with smaller chunks; the working pair is kept for next frame. */
    unsigned int chunk = sc->chunk_samples, factor = sc->cap_factor;
    const char *env = getenv("FRT_CHUNK_SAMPLES");
    if (env != nullptr && *env) {
        chunk = (unsigned int)std::max(1024L, atol(env));
    }
    int rc;
    for (;;) {
        rc = render_once(sc, cfg, &st, chunk, factor, canvas_rgba);
        if (rc != -1) {
            break;
        }
        st.overflow += 1;
        size_t free_b = 0, total_b = 0;
        cudaMemGetInfo(&free_b, &total_b);
        const unsigned long long frame_samples = std::max<unsigned long long>(st.rays_primary, 1);
        const unsigned long long eff_chunk = std::min<unsigned long long>(chunk, frame_samples);
        const unsigned long long slot_bytes = 2 * 84 + 28 + sizeof(LightRec) + sizeof(LightTmp) + 4 * sizeof(unsigned long long);
        if (factor < 64 && eff_chunk * factor * 2 * slot_bytes < (free_b + (unsigned long long)sc->capacity * slot_bytes) / 4) {
            factor *= 2;
        } else if (eff_chunk > 4096) {
            chunk = (unsigned int)(eff_chunk / 2);
        } else {
            return frt_set_error(FRT_ERR_OVERFLOW, "secondary-ray queue overflow even with %u-sample chunks", chunk);
        }
    }
    if (rc == FRT_OK && (env == nullptr || !*env)) {
        sc->chunk_samples = chunk;
        sc->cap_factor = factor;
    }
    if (rc != FRT_OK) {
        return rc;
    }
    if (sc->gen_check_pending) { /* the frame is done: the comparison (upload stream) finished long ago */
        rc = gen_check(sc, sc->upload_stream);
        if (rc != FRT_OK) {
            return rc;
        }
    }
    if (canvas_rgba != nullptr) {
        /* enqueued by render_once behind the frame's last kernel and complete by now (it waited for the stream) */
        CK(cudaEventSynchronize(sc->ev[3]));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, sc->ev[2], sc->ev[3]));
        st.download_ms = ms;
    }
    if (stats != nullptr) {
        *stats = st;
    }
    return FRT_OK;
}

/* ------------------------------------------------------------------------------------------------ photons */

static double
lab_lightness(const double rgb[3])
{
    /* rgb_to_lab's L* (src/color/rgb.c:58, xyz.c:31-56): Y row of the RGB->XYZ matrix over the D65 white's Y = 1 */
    double y = 0.212671 * rgb[0] + 0.715160 * rgb[1] + 0.072169 * rgb[2];
    return y > 0.008856 ? 116.0 * cbrt(y) - 16.0 : 903.3 * y;
}

static int
pm_reserve(frt_scene *sc, int map, unsigned int cap)
{
    frt_scene::PMap &m = sc->pm[map];
    if (m.cap >= cap) {
        return FRT_OK;
    }
    float4 *na = nullptr, *nb = nullptr;
    CK(cudaMalloc(&na, sizeof(float4) * (size_t)cap));
    CK(cudaMalloc(&nb, sizeof(float4) * (size_t)cap));
    if (m.count) {
        CK(cudaMemcpyAsync(na, m.ra, sizeof(float4) * (size_t)m.count, cudaMemcpyDeviceToDevice, sc->stream));
        CK(cudaMemcpyAsync(nb, m.rb, sizeof(float4) * (size_t)m.count, cudaMemcpyDeviceToDevice, sc->stream));
        CK(cudaStreamSynchronize(sc->stream));
    }
    cudaFree(m.ra);
    cudaFree(m.rb);
    m.ra = na;
    m.rb = nb;
    m.cap = cap;
    return FRT_OK;
}

static void
pm_free(frt_scene *sc)
{
    for (auto &m : sc->pm) {
        cudaFree(m.ra);
        cudaFree(m.rb);
        cudaFree(m.sa);
        cudaFree(m.sb);
        cudaFree(m.sd);
        cudaFree(m.cell_start);
        m = frt_scene::PMap{};
    }
    cudaFree(sc->pm_stored);
    cudaFree(sc->pm_dir_tab);
    sc->pm_dir_tab = nullptr;
    cudaFree(sc->gq);
    cudaFree(sc->gq2);
    cudaFree(sc->gq_counts);
    cudaFree(sc->gq_starts);
    cudaFree(sc->gq_part);
    cudaFree(sc->gq_work);
    sc->gq2 = nullptr;
    sc->gq_counts = sc->gq_starts = sc->gq_part = sc->gq_work = nullptr;
    sc->gq_cells = sc->gq2_cap = sc->gq_part_cap = 0;
    cudaFree(sc->gq_n);
    cudaFree(sc->acc_amb);
    cudaFree(sc->acc_fg);
    sc->pm_stored = nullptr;
    sc->gq = nullptr;
    sc->gq_n = nullptr;
    sc->acc_amb = sc->acc_fg = nullptr;
    sc->gq_cap = sc->acc_cap = 0;
    sc->pm_ready = false;
}

/*
 * The radiance estimates of one batch of requests (q[0 .. *q_n)): sorted by the photon grid's cell and handed to
 * k_knn_cell, a lane per request; what that kernel passes on is left in q and goes through k_knn, a warp per request.
 * q is overwritten.
 */
static int
launch_knn(frt_scene *sc, cudaStream_t s, const GIParams &G, GQuery *q, unsigned int *q_n, unsigned int q_cap, double *acc_amb, double *acc_fg,
           int *found, int mode, int *launches)
{
    const int sm_blocks = sc->sm_count;
    const bool global_map = sc->pm[1].view.count != 0;
    const PMView &M = global_map ? sc->pm[1].view : sc->pm[0].view;
    const size_t n_cells = (size_t)M.nx * M.ny * M.nz;
    if (mode == 2) {
        k_knn_list<<<sm_blocks * 8, FRT_KNN_WARPS * 32, 0, s>>>(sc->pm[0].view, sc->pm[1].view, G, q, q_n, q_cap, acc_amb, acc_fg, found);
        ++*launches;
        return FRT_OK;
    }
    if (mode == 0 || M.count == 0 || n_cells == 0 || n_cells >= 0xffffffffull) {
        k_knn<<<sm_blocks * 16, FRT_KNN_WARPS * 32, 0, s>>>(sc->pm[0].view, sc->pm[1].view, G, q, q_n, q_cap, acc_amb, acc_fg, found);
        ++*launches;
        return FRT_OK;
    }
    const unsigned int n_part = (unsigned int)((n_cells + FRT_SCAN_CHUNK - 1) / FRT_SCAN_CHUNK);
    if (sc->gq_cells < n_cells + 1) {
        cudaFree(sc->gq_counts);
        cudaFree(sc->gq_starts);
        sc->gq_counts = sc->gq_starts = nullptr;
        sc->gq_cells = 0;
        CK(cudaMalloc(&sc->gq_counts, sizeof(unsigned int) * (n_cells + 1)));
        CK(cudaMalloc(&sc->gq_starts, sizeof(unsigned int) * (n_cells + 1)));
        sc->gq_cells = n_cells + 1;
    }
    if (sc->gq_part_cap < (size_t)n_part + 1) {
        cudaFree(sc->gq_part);
        sc->gq_part = nullptr;
        sc->gq_part_cap = 0;
        CK(cudaMalloc(&sc->gq_part, sizeof(unsigned int) * 2 * ((size_t)n_part + 1)));
        sc->gq_part_cap = (size_t)n_part + 1;
    }
    if (sc->gq2_cap < q_cap) {
        cudaFree(sc->gq2);
        sc->gq2 = nullptr;
        sc->gq2_cap = 0;
        CK(cudaMalloc(&sc->gq2, sizeof(GQuery) * (size_t)q_cap));
        sc->gq2_cap = q_cap;
    }
    if (sc->gq_work == nullptr) {
        CK(cudaMalloc(&sc->gq_work, 2 * sizeof(unsigned int)));
        CK(cudaFuncSetAttribute(k_knn_cell, cudaFuncAttributeMaxDynamicSharedMemorySize, FRT_KC_SMEM));
    }
    unsigned int *part = sc->gq_part, *part_excl = sc->gq_part + n_part + 1;
    const bool debug = frt_env_int("FRT_KNN_DEBUG", 0) != 0; /* development aid: the three phases timed, what was passed on counted */
    cudaEvent_t ev[4] = { nullptr, nullptr, nullptr, nullptr };
    if (debug) {
        for (auto &e : ev) {
            CK(cudaEventCreate(&e));
        }
        CK(cudaEventRecord(ev[0], s));
    }
    CK(cudaMemsetAsync(sc->gq_counts, 0, sizeof(unsigned int) * n_cells, s));
    CK(cudaMemsetAsync(sc->gq_work, 0, 2 * sizeof(unsigned int), s));
    k_gq_count<<<sm_blocks * 8, 256, 0, s>>>(M, q, q_n, q_cap, sc->gq_counts);
    k_scan_partial<<<n_part, 1024, 0, s>>>(sc->gq_counts, (unsigned int)n_cells, part);
    k_pm_scan<<<1, 1024, 0, s>>>(part, part_excl, n_part);
    k_scan_apply<<<n_part, 1024, 0, s>>>(sc->gq_counts, part_excl, (unsigned int)n_cells, sc->gq_starts);
    k_gq_scatter<<<sm_blocks * 8, 256, 0, s>>>(M, q, q_n, q_cap, sc->gq_starts, sc->gq2);
    if (debug) CK(cudaEventRecord(ev[1], s));
    k_knn_cell<<<sm_blocks * 3, FRT_KC_T, FRT_KC_SMEM, s>>>(M, global_map ? 0 : 1, G, sc->gq2, q_n, q_cap, sc->gq_work, acc_amb, acc_fg, found, q, q_cap);
    if (debug) CK(cudaEventRecord(ev[2], s));
    k_knn<<<sm_blocks * 16, FRT_KNN_WARPS * 32, 0, s>>>(sc->pm[0].view, sc->pm[1].view, G, q, sc->gq_work + 1, q_cap, acc_amb, acc_fg, found);
    *launches += 7;
    CK(cudaGetLastError());
    if (debug) {
        CK(cudaEventRecord(ev[3], s));
        CK(cudaStreamSynchronize(s));
        unsigned int hw[2] = { 0, 0 }, hn = 0;
        CK(cudaMemcpy(hw, sc->gq_work, sizeof(hw), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(&hn, q_n, sizeof(hn), cudaMemcpyDeviceToHost));
        float t[3] = { 0.f, 0.f, 0.f };
        for (int k = 0; k < 3; ++k) {
            cudaEventElapsedTime(&t[k], ev[k], ev[k + 1]);
        }
        fprintf(stderr, "[frt] knn batch: %u requests, %zu cells, sort %.2f ms, k_knn_cell %.2f ms, %u passed on to k_knn %.2f ms\n", hn, n_cells, t[0],
                t[1], hw[1], t[2]);
        for (auto &e : ev) {
            cudaEventDestroy(e);
        }
    }
    return FRT_OK;
}

static int
ensure_gi_buffers(frt_scene *sc)
{
    if (sc->acc_cap < sc->capacity) {
        cudaFree(sc->acc_amb);
        cudaFree(sc->acc_fg);
        sc->acc_amb = sc->acc_fg = nullptr;
        CK(cudaMalloc(&sc->acc_amb, sizeof(double) * 3 * (size_t)sc->capacity));
        CK(cudaMalloc(&sc->acc_fg, sizeof(double) * 3 * (size_t)sc->capacity));
        sc->acc_cap = sc->capacity;
    }
    if (sc->gq == nullptr) {
        unsigned int cap = 32u << 20; /* 32 Mi requests of 40 bytes */
        const char *env = getenv("FRT_GI_QUEUE");
        if (env != nullptr && *env) {
            cap = (unsigned int)std::max(4096L, atol(env));
        }
        CK(cudaMalloc(&sc->gq, sizeof(GQuery) * (size_t)cap));
        CK(cudaMalloc(&sc->gq_n, sizeof(unsigned int)));
        sc->gq_cap = cap;
    }
    return FRT_OK;
}

/*
 * trace_photons (photon_tracer.c:203-257) for this rank's shard.  Per light, the reference emits photons until the
 * light's quota (photon_count * L*(intensity) / sum L*) has been STORED (j += hit, :231-247); here the quota is
 * divided by `world`, photons are emitted in rounds sized from the measured stores-per-photon, and the rank's photon
 * indices are i * world + rank so that the shards are disjoint streams.
 */
extern "C" int
frt_photons_emit(frt_scene *sc, const frt_photon_cfg *cfg, frt_stats *stats)
{
    if (sc == nullptr || cfg == nullptr) {
        return frt_set_error(FRT_ERR_ARG, "frt_photons_emit: null argument");
    }
    CK(cudaSetDevice(sc->device));
    if (sc->upload_pending) { /* the emitters read the lights; be conservative about what the upload stream still writes */
        CK(cudaStreamSynchronize(sc->upload_stream));
        sc->upload_pending = false;
    }
    const frt_config &g = sc->cfg;
    const int world = cfg->world > 0 ? cfg->world : 1;
    if (cfg->rank < 0 || cfg->rank >= world) {
        return frt_set_error(FRT_ERR_ARG, "frt_photons_emit: rank %d of world %d", cfg->rank, world);
    }
    if (g.gi_photon_count <= 0 || sc->S.n_lights <= 0) {
        return frt_set_error(FRT_ERR_ARG, "frt_photons_emit: photon_count is %lld and the scene has %d lights",
                             (long long)g.gi_photon_count, sc->S.n_lights);
    }
    if (g.gi_path_length < 0 || g.gi_path_length > 64) {
        return frt_set_error(FRT_ERR_ARG, "gi.path_length %d is outside 0..64", g.gi_path_length);
    }
    std::vector<frt_light> hl(sc->S.n_lights);
    CK(cudaMemcpyAsync(hl.data(), sc->S.lights, sizeof(frt_light) * hl.size(), cudaMemcpyDeviceToHost, sc->stream));
    CK(cudaStreamSynchronize(sc->stream));
    double total_l = 0.0;
    for (auto &l : hl) {
        total_l += lab_lightness(l.intensity);
    }
    if (sc->pm_stored == nullptr) {
        CK(cudaMalloc(&sc->pm_stored, 2 * sizeof(unsigned int)));
    }
    if (sc->cnt == nullptr) {
        int rc0 = ensure_frame_buffers(sc, 1024);
        if (rc0 != FRT_OK) return rc0;
    }
    CK(cudaMemsetAsync(sc->cnt, 0, sizeof(Counters), sc->stream));
    unsigned long long emitted_total = 0;
    for (int map = 0; map < 2; ++map) {
        const bool want = map == 0 ? cfg->populate_caustic != 0 : cfg->populate_global != 0;
        frt_scene::PMap &m = sc->pm[map];
        m.count = 0;
        m.built = false;
        m.scaled = false;
        if (!want) {
            continue;
        }
        /* capacity: this rank's share of every light's quota + the overshoot of one path */
        const unsigned long long share = ((unsigned long long)g.gi_photon_count + world - 1) / world;
        int rc = pm_reserve(sc, map, (unsigned int)std::min<unsigned long long>(share + 64 + (unsigned long long)g.gi_path_length, 0x7ffffff0ull));
        if (rc != FRT_OK) {
            return rc;
        }
        CK(cudaMemsetAsync(sc->pm_stored + map, 0, sizeof(unsigned int), sc->stream));
        unsigned long long target = 0; /* cumulative stored-photon target over the lights */
        for (int li = 0; li < sc->S.n_lights; ++li) {
            const unsigned long long quota = (unsigned long long)((double)g.gi_photon_count * lab_lightness(hl[li].intensity) / total_l);
            target += (quota + world - 1) / world;
            unsigned long long first = 0, emitted = 0;
            unsigned int stored = 0, before = 0;
            CK(cudaMemcpyAsync(&before, sc->pm_stored + map, sizeof(unsigned int), cudaMemcpyDeviceToHost, sc->stream));
            CK(cudaStreamSynchronize(sc->stream));
            stored = before;
            for (int round = 0; round < 256 && stored < target; ++round) {
                unsigned long long need = target - stored, n;
                if (emitted == 0) {
                    n = std::max<unsigned long long>(need / 8, 4096);
                } else if (stored == before) {
                    if (emitted > 64ull * (target + 4096)) {
                        break; /* nothing stores (e.g. a caustic map in a scene without specular surfaces): the reference
                                  would loop forever here (photon_tracer.c:231-238) */
                    }
                    n = emitted * 2;
                } else {
                    double yield = (double)(stored - before) / (double)emitted;
                    n = (unsigned long long)((double)need / yield * 1.02) + 256;
                }
                PhotonParams P{};
                P.light = li;
                P.map_type = map;
                P.path_length = g.gi_path_length;
                P.rank = cfg->rank;
                P.world = world;
                P.first = first;
                P.count = n;
                P.seed = mix64(cfg->seed ^ 0x70686f746f6e73ull);
                const int blocks = (int)std::min<unsigned long long>((n + 127) / 128, (unsigned long long)sc->sm_count * 16);
#define FRT_PT(PR, MP) k_photon_trace<PR, MP><<<blocks, 128, 0, sc->stream>>>(sc->S, sc->SF, P, m.ra, m.rb, sc->pm_stored + map, m.cap, sc->cnt)
                if (sc->boxes_and_balls && !sc->has_maps) {
                    FRT_PT(FRT_PRIMS_BOXES_AND_BALLS, false);
                } else if (sc->boxes_and_balls) {
                    FRT_PT(FRT_PRIMS_BOXES_AND_BALLS, true);
                } else if (!sc->has_maps) {
                    FRT_PT(FRT_PRIMS_ALL, false);
                } else {
                    FRT_PT(FRT_PRIMS_ALL, true);
                }
#undef FRT_PT
                CK(cudaGetLastError());
                CK(cudaMemcpyAsync(&stored, sc->pm_stored + map, sizeof(unsigned int), cudaMemcpyDeviceToHost, sc->stream));
                CK(cudaStreamSynchronize(sc->stream));
                first += n;
                emitted += n;
            }
            emitted_total += emitted;
            /* photons past the cumulative target are dropped (the reference overshoots by less than one path) */
            unsigned int keep = (unsigned int)std::min<unsigned long long>(std::min<unsigned long long>(stored, target + (unsigned long long)g.gi_path_length), m.cap);
            CK(cudaMemcpyAsync(sc->pm_stored + map, &keep, sizeof(unsigned int), cudaMemcpyHostToDevice, sc->stream)); /* pageable: staged before the call returns */
            m.count = keep;
        }
    }
    Counters hc;
    CK(cudaMemcpyAsync(&hc, sc->cnt, sizeof(Counters), cudaMemcpyDeviceToHost, sc->stream));
    CK(cudaStreamSynchronize(sc->stream));
    if (hc.overflow_csg) {
        return frt_set_error(FRT_ERR_OVERFLOW, "a photon met more than %d CSG crossings", FRT_CSG_CAP);
    }
    sc->pm_ready = false;
    if (stats != nullptr) {
        stats->rays_photon = emitted_total;
        stats->photons_stored[0] = sc->pm[0].count;
        stats->photons_stored[1] = sc->pm[1].count;
        stats->photons_stored[2] = 0;
    }
    return FRT_OK;
}

extern "C" int64_t
frt_photons_count(frt_scene *sc, int map)
{
    if (sc == nullptr || map < 0 || map > 1) {
        return 0;
    }
    return (int64_t)sc->pm[map].count;
}

/* Layout of an exported map: count records of {x, y, z, dir bits} followed by count records of {r, g, b, 0} (32 B per photon). */
extern "C" int
frt_photons_export(frt_scene *sc, int map, void *dst, int dst_is_device)
{
    if (sc == nullptr || map < 0 || map > 1 || dst == nullptr) {
        return frt_set_error(FRT_ERR_ARG, "frt_photons_export: bad argument");
    }
    CK(cudaSetDevice(sc->device));
    const frt_scene::PMap &m = sc->pm[map];
    const size_t bytes = sizeof(float4) * (size_t)m.count;
    const cudaMemcpyKind kind = dst_is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    if (m.count) {
        CK(cudaMemcpyAsync(dst, m.ra, bytes, kind, sc->stream));
        CK(cudaMemcpyAsync((char *)dst + bytes, m.rb, bytes, kind, sc->stream));
        CK(cudaStreamSynchronize(sc->stream));
    }
    return FRT_OK;
}

/* Replaces the map's photons by `count` records in the export layout (e.g. the all-gathered shards of every rank). */
extern "C" int
frt_photons_import(frt_scene *sc, int map, const void *src, int64_t count, int src_is_device)
{
    if (sc == nullptr || map < 0 || map > 1 || count < 0 || count > 0x7ffffff0ll || (count > 0 && src == nullptr)) {
        return frt_set_error(FRT_ERR_ARG, "frt_photons_import: bad argument");
    }
    CK(cudaSetDevice(sc->device));
    frt_scene::PMap &m = sc->pm[map];
    m.count = 0;
    m.built = false;
    m.scaled = false;
    sc->pm_ready = false;
    int rc = pm_reserve(sc, map, (unsigned int)std::max<int64_t>(count, 1));
    if (rc != FRT_OK) {
        return rc;
    }
    const size_t bytes = sizeof(float4) * (size_t)count;
    const cudaMemcpyKind kind = src_is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    if (count) {
        CK(cudaMemcpyAsync(m.ra, src, bytes, kind, sc->stream));
        CK(cudaMemcpyAsync(m.rb, (const char *)src + bytes, bytes, kind, sc->stream));
        CK(cudaStreamSynchronize(sc->stream)); /* the caller's buffer may go away (and a device source may be reused) after this call */
    }
    m.count = (unsigned int)count;
    return FRT_OK;
}

/* pm_scale_photon_power + pm_balance (photon_tracer.c:251-256): scale by 1 / photon_count, bin into the grid */
extern "C" int
frt_photons_finish(frt_scene *sc)
{
    if (sc == nullptr) {
        return frt_set_error(FRT_ERR_ARG, "frt_photons_finish: null scene");
    }
    CK(cudaSetDevice(sc->device));
    const frt_config &g = sc->cfg;
    const float scale = (float)(1.0 / (double)std::max<int64_t>(g.gi_photon_count, 1));
    const float radius = (float)g.gi_irradiance_estimate_radius;
    if (!(radius > 0.f)) {
        return frt_set_error(FRT_ERR_ARG, "irradiance-estimate-radius must be positive");
    }
    if (sc->pm_dir_tab == nullptr) {
        /* init_Photon_map's direction tables (pm.c:54-60), evaluated in FP64 like there, stored as FP32 */
        std::vector<float> tab(1024);
        for (int i = 0; i < 256; ++i) {
            const double angle = (double)i * (1.0 / 256.0) * M_PI;
            tab[i] = (float)sin(angle);
            tab[256 + i] = (float)cos(angle);
            tab[512 + i] = (float)cos(2.0 * angle);
            tab[768 + i] = (float)sin(2.0 * angle);
        }
        CK(cudaMalloc(&sc->pm_dir_tab, sizeof(float) * 1024));
        CK(cudaMemcpyAsync(sc->pm_dir_tab, tab.data(), sizeof(float) * 1024, cudaMemcpyHostToDevice, sc->stream));
        CK(cudaStreamSynchronize(sc->stream));
    }
    for (int map = 0; map < 2; ++map) {
        frt_scene::PMap &m = sc->pm[map];
        cudaFree(m.sa);
        cudaFree(m.sb);
        cudaFree(m.sd);
        cudaFree(m.cell_start);
        m.sa = m.sb = m.sd = nullptr;
        m.cell_start = nullptr;
        m.view = PMView{};
        m.view.dir_tab = sc->pm_dir_tab;
        m.built = true;
        if (m.count == 0) {
            continue;
        }
        unsigned int hb[6] = { 0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u };
        unsigned int *db = nullptr;
        CK(cudaMalloc(&db, sizeof(hb)));
        CK(cudaMemcpyAsync(db, hb, sizeof(hb), cudaMemcpyHostToDevice, sc->stream)); /* pageable: staged before the call returns */
        k_pm_scale_bounds<<<sc->sm_count * 4, 256, 0, sc->stream>>>(m.ra, m.rb, m.count, m.scaled ? 1.0f : scale, db);
        m.scaled = true;
        CK(cudaMemcpyAsync(hb, db, sizeof(hb), cudaMemcpyDeviceToHost, sc->stream));
        CK(cudaStreamSynchronize(sc->stream));
        cudaFree(db);
        float lo[3], hi[3], ext = 0.f;
        for (int k = 0; k < 3; ++k) {
            lo[k] = float_unorder(hb[k]);
            hi[k] = float_unorder(hb[3 + k]);
            ext = std::max(ext, hi[k] - lo[k]);
        }
        /* half the search radius: 5 x 5 rows of cells, each clipped to the sphere.  Coarser cells only when the grid would
         * not fit (a few photons that left the box through the window make the Cornell map 20 units wide: at ext / 256 the
         * cells were 0.078 instead of 0.05 and a request looked at 2.4 x the photons) */
        float cell = std::max(0.5f * radius, ext / 2048.0f);
        for (;;) {
            const double cells = (floor((hi[0] - lo[0]) / cell) + 1.0) * (floor((hi[1] - lo[1]) / cell) + 1.0) * (floor((hi[2] - lo[2]) / cell) + 1.0);
            if (cells <= (double)(1u << 26)) {
                break;
            }
            cell *= 1.25f;
        }
        PMView V{};
        V.gx = lo[0];
        V.gy = lo[1];
        V.gz = lo[2];
        V.inv_cell = 1.0f / cell;
        V.nx = std::max(1, (int)floorf((hi[0] - lo[0]) / cell) + 1);
        V.ny = std::max(1, (int)floorf((hi[1] - lo[1]) / cell) + 1);
        V.nz = std::max(1, (int)floorf((hi[2] - lo[2]) / cell) + 1);
        V.count = m.count;
        V.dir_tab = sc->pm_dir_tab;
        const size_t n_cells = (size_t)V.nx * V.ny * V.nz;
        unsigned int *counts = nullptr;
        CK(cudaMalloc(&counts, sizeof(unsigned int) * n_cells));
        CK(cudaMalloc(&m.cell_start, sizeof(unsigned int) * (n_cells + 1)));
        CK(cudaMalloc(&m.sa, sizeof(float4) * (size_t)m.count));
        CK(cudaMalloc(&m.sb, sizeof(float4) * (size_t)m.count));
        CK(cudaMalloc(&m.sd, sizeof(float4) * (size_t)m.count));
        CK(cudaMemsetAsync(counts, 0, sizeof(unsigned int) * n_cells, sc->stream));
        k_pm_count<<<sc->sm_count * 4, 256, 0, sc->stream>>>(V, m.ra, m.count, counts);
        {
            /* the scan of the (up to 2^26) cell counts in three launches */
            const unsigned int n_part = (unsigned int)((n_cells + FRT_SCAN_CHUNK - 1) / FRT_SCAN_CHUNK);
            unsigned int *part = nullptr;
            CK(cudaMalloc(&part, sizeof(unsigned int) * 2 * ((size_t)n_part + 1)));
            k_scan_partial<<<n_part, 1024, 0, sc->stream>>>(counts, (unsigned int)n_cells, part);
            k_pm_scan<<<1, 1024, 0, sc->stream>>>(part, part + n_part + 1, n_part);
            k_scan_apply<<<n_part, 1024, 0, sc->stream>>>(counts, part + n_part + 1, (unsigned int)n_cells, m.cell_start);
            CK(cudaStreamSynchronize(sc->stream));
            cudaFree(part);
        }
        /* counts becomes the per-cell write cursor */
        CK(cudaMemcpyAsync(counts, m.cell_start, sizeof(unsigned int) * n_cells, cudaMemcpyDeviceToDevice, sc->stream));
        k_pm_scatter<<<sc->sm_count * 4, 256, 0, sc->stream>>>(V, m.ra, m.rb, m.count, counts, m.sa, m.sb);
        k_pm_dirs<<<sc->sm_count * 4, 256, 0, sc->stream>>>(m.sa, m.count, sc->pm_dir_tab, m.sd);
        CK(cudaStreamSynchronize(sc->stream));
        CK(cudaGetLastError());
        cudaFree(counts);
        V.a = m.sa;
        V.b = m.sb;
        V.c = m.sd;
        V.cell_start = m.cell_start;
        m.view = V;
    }
    sc->pm_ready = true;
    return FRT_OK;
}

/* ------------------------------------------------------------------------------------------------ several GPUs, one process */

/*
 * render_multi()'s row fan-out (renderer.c:244-281: a pthread pool, one job per image row, every thread on its own deep
 * copy of the World) across the GPUs of one node, behind the same call: the scene is replicated -- uploaded, or its
 * light cache rebuilt, once per device, concurrently -- device k renders the row blocks b with b % n == k on its own
 * host thread and stream, and writes them straight into the caller's Canvas.arr (a block of 4 rows is one contiguous
 * run of the canvas, so the rows need no packing, no collective and no reorder).  The photon pass shards the emission
 * the same way and exchanges the stored photons device to device (cudaMemcpyPeerAsync over NVLink) before every device
 * bins the full set.  One process per GPU over NCCL (fast_ray_tracer_b200/dist.py) remains the launch shape of bench.py.
 */
struct frt_multi {
    std::vector<int> devices;
    std::vector<frt_scene *> scenes;
};

template <typename Fn>
static int
multi_for_each(frt_multi *m, Fn fn)
{
    const int n = (int)m->devices.size();
    std::vector<int> rc(n, FRT_OK);
    std::vector<std::string> msg(n);
    std::vector<std::thread> th;
    th.reserve(n);
    for (int k = 0; k < n; ++k) {
        th.emplace_back([&, k]() {
            rc[k] = fn(k);
            if (rc[k] != FRT_OK) {
                msg[k] = frt_last_error(); /* the worker's thread-local message */
            }
        });
    }
    for (auto &t : th) {
        t.join();
    }
    for (int k = 0; k < n; ++k) {
        if (rc[k] != FRT_OK) {
            return frt_set_error(rc[k], "device %d: %s", m->devices[k], msg[k].c_str());
        }
    }
    return FRT_OK;
}

extern "C" void
frt_multi_destroy(frt_multi *m)
{
    if (m == nullptr) {
        return;
    }
    for (frt_scene *s : m->scenes) {
        frt_scene_destroy(s);
    }
    delete m;
}

extern "C" int
frt_multi_create(const frt_scene_desc *desc, const int32_t *devices, int n_devices, const frt_light_gen *gens, int n_gens, frt_multi **out)
{
    if (out == nullptr) {
        return frt_set_error(FRT_ERR_ARG, "frt_multi_create: null out pointer");
    }
    *out = nullptr;
    int ndev = frt_device_count();
    if (ndev <= 0) {
        return frt_set_error(FRT_ERR_CUDA, "no CUDA device is visible: the B200 core has no CPU fallback");
    }
    frt_multi *m = new frt_multi();
    if (devices == nullptr || n_devices <= 0) { /* all of them */
        for (int k = 0; k < ndev; ++k) {
            m->devices.push_back(k);
        }
    } else {
        for (int k = 0; k < n_devices; ++k) {
            if (devices[k] < 0 || devices[k] >= ndev || std::count(m->devices.begin(), m->devices.end(), devices[k])) {
                delete m;
                return frt_set_error(FRT_ERR_ARG, "frt_multi_create: device %d of %d (or listed twice)", devices[k], ndev);
            }
            m->devices.push_back(devices[k]);
        }
    }
    m->scenes.assign(m->devices.size(), nullptr);
    for (int a : m->devices) { /* photon shards travel device to device; without peer access the copies stage through the host */
        for (int b : m->devices) {
            int can = 0;
            if (a != b && cudaDeviceCanAccessPeer(&can, a, b) == cudaSuccess && can && cudaSetDevice(a) == cudaSuccess) {
                if (cudaDeviceEnablePeerAccess(b, 0) != cudaSuccess) {
                    cudaGetLastError(); /* already enabled */
                }
            }
        }
    }
    int rc = multi_for_each(m, [&](int k) { return scene_create(desc, m->devices[k], gens, n_gens, &m->scenes[k]); });
    if (rc != FRT_OK) {
        frt_multi_destroy(m);
        return rc;
    }
    *out = m;
    return FRT_OK;
}

extern "C" int
frt_multi_device_count(const frt_multi *m)
{
    return m == nullptr ? 0 : (int)m->devices.size();
}

extern "C" frt_scene *
frt_multi_scene(frt_multi *m, int k)
{
    return (m == nullptr || k < 0 || k >= (int)m->scenes.size()) ? nullptr : m->scenes[k];
}

extern "C" int
frt_multi_render(frt_multi *m, const frt_render_cfg *cfg, double *canvas_rgba, frt_stats *stats)
{
    if (m == nullptr || cfg == nullptr) {
        return frt_set_error(FRT_ERR_ARG, "frt_multi_render: null argument");
    }
    const int n = (int)m->devices.size();
    std::vector<frt_stats> st(n);
    int rc = multi_for_each(m, [&](int k) {
        frt_render_cfg c = *cfg;
        c.device = m->devices[k];
        c.rank = k;
        c.world = n;
        return frt_render(m->scenes[k], &c, canvas_rgba, &st[k]);
    });
    if (rc != FRT_OK) {
        return rc;
    }
    if (stats != nullptr) {
        frt_stats t = st[0];
        for (int k = 1; k < n; ++k) {
            const frt_stats &s = st[k];
            t.frame_ms = std::max(t.frame_ms, s.frame_ms); /* the devices run side by side */
            t.light_ms = std::max(t.light_ms, s.light_ms);
            t.download_ms = std::max(t.download_ms, s.download_ms);
            t.rays_primary += s.rays_primary;
            t.rays_secondary += s.rays_secondary;
            t.rays_shadow += s.rays_shadow;
            t.rays_gather += s.rays_gather;
            t.hits_shaded += s.hits_shaded;
            t.light_launches += s.light_launches;
            t.kernel_launches += s.kernel_launches;
            t.shadow_nodes += s.shadow_nodes;
            t.overflow += s.overflow;
            t.light_flops += s.light_flops;
            t.shadow_deferred += s.shadow_deferred;
            t.shadow_mismatch += s.shadow_mismatch;
            for (int j = 0; j < 10; ++j) {
                t.shadow_reasons[j] += s.shadow_reasons[j];
            }
            t.rows_rendered += s.rows_rendered;
        }
        *stats = t;
    }
    return FRT_OK;
}

/* trace_photons (photon_tracer.c:203) on every device: shard k of n is emitted on device k, then every device receives
 * the other shards (peer copies) and builds the same lookup grid from the full set. */
extern "C" int
frt_multi_photons(frt_multi *m, const frt_photon_cfg *cfg, frt_stats *stats)
{
    if (m == nullptr || cfg == nullptr) {
        return frt_set_error(FRT_ERR_ARG, "frt_multi_photons: null argument");
    }
    const int n = (int)m->devices.size();
    std::vector<frt_stats> st(n);
    int rc = multi_for_each(m, [&](int k) {
        frt_photon_cfg c = *cfg;
        c.device = m->devices[k];
        c.rank = k;
        c.world = n;
        memset(&st[k], 0, sizeof(frt_stats));
        return frt_photons_emit(m->scenes[k], &c, &st[k]);
    });
    if (rc != FRT_OK) {
        return rc;
    }
    if (n > 1) {
        for (int map = 0; map < 2; ++map) {
            std::vector<size_t> cnt(n), off(n + 1, 0);
            for (int k = 0; k < n; ++k) {
                cnt[k] = m->scenes[k]->pm[map].count;
                off[k + 1] = off[k] + cnt[k];
            }
            const size_t total = off[n];
            if (total == 0) {
                continue;
            }
            rc = multi_for_each(m, [&](int d) {
                frt_scene *sc = m->scenes[d];
                CK(cudaSetDevice(sc->device));
                float4 *merged = nullptr;
                CK(cudaMalloc(&merged, sizeof(float4) * 2 * total));
                cudaError_t e = cudaSuccess;
                for (int k = 0; k < n && e == cudaSuccess; ++k) {
                    const frt_scene::PMap &src = m->scenes[k]->pm[map];
                    if (cnt[k] == 0) continue;
                    e = cudaMemcpyPeerAsync(merged + off[k], sc->device, src.ra, m->devices[k], sizeof(float4) * cnt[k], sc->stream);
                    if (e == cudaSuccess) {
                        e = cudaMemcpyPeerAsync(merged + total + off[k], sc->device, src.rb, m->devices[k], sizeof(float4) * cnt[k], sc->stream);
                    }
                }
                if (e == cudaSuccess) e = cudaStreamSynchronize(sc->stream);
                int r = FRT_OK;
                if (e != cudaSuccess) {
                    r = frt_set_error(FRT_ERR_CUDA, "photon exchange: %s", cudaGetErrorString(e));
                }
                sc->pm_merged[map] = merged;
                return r;
            });
            /* every device has read every shard: only now may the shards be replaced */
            int rc2 = multi_for_each(m, [&](int d) {
                frt_scene *sc = m->scenes[d];
                CK(cudaSetDevice(sc->device));
                int r = FRT_OK;
                if (rc == FRT_OK && sc->pm_merged[map] != nullptr) {
                    r = frt_photons_import(sc, map, sc->pm_merged[map], (int64_t)total, 1);
                }
                cudaFree(sc->pm_merged[map]);
                sc->pm_merged[map] = nullptr;
                return r;
            });
            if (rc != FRT_OK) return rc;
            if (rc2 != FRT_OK) return rc2;
        }
    }
    rc = multi_for_each(m, [&](int k) { return frt_photons_finish(m->scenes[k]); });
    if (rc != FRT_OK) {
        return rc;
    }
    if (stats != nullptr) {
        memset(stats, 0, sizeof(*stats));
        for (int k = 0; k < n; ++k) {
            stats->rays_photon += st[k].rays_photon;
        }
        stats->photons_stored[0] = m->scenes[0]->pm[0].count;
        stats->photons_stored[1] = m->scenes[0]->pm[1].count;
    }
    return FRT_OK;
}

/*
 * pm_irradiance_estimate (pm.c:91-156) for n positions against the scene's photon map `map` (0 caustic, 1 global), with
 * the radius / photon count / cone filter of the scene's configuration: the unit-level twin of what k_knn does inside a
 * frame, without the callers' rescaling (renderer.c:845, :878).  pos / normal: n x 3 doubles (positions are rounded to
 * FP32 like every request of a frame); irrad: n x 3 doubles; found: n ints, the function's return value per position.
 */
extern "C" int
frt_photons_estimate(frt_scene *sc, int map, int64_t n, const double *pos, const double *normal, double *irrad, int32_t *found)
{
    if (sc == nullptr || map < 0 || map > 1 || n < 0 || n > 0x3fffffff || (n > 0 && (pos == nullptr || normal == nullptr || irrad == nullptr))) {
        return frt_set_error(FRT_ERR_ARG, "frt_photons_estimate: bad argument");
    }
    if (!sc->pm_ready) {
        return frt_set_error(FRT_ERR_ARG, "frt_photons_estimate: no photon map was built (frt_photons_emit / _import, then frt_photons_finish)");
    }
    if (n == 0) {
        return FRT_OK;
    }
    CK(cudaSetDevice(sc->device));
    const frt_config &g = sc->cfg;
    GIParams G{};
    G.usteps = G.vsteps = 1;
    G.n_photons = g.gi_irradiance_estimate_num;
    G.radius = (float)g.gi_irradiance_estimate_radius;
    G.cone_k = (float)g.gi_irradiance_estimate_cone_filter_k;
    if (G.n_photons <= 0 || G.n_photons > FRT_KNN_CAP / 2) {
        return frt_set_error(FRT_ERR_ARG, "irradiance-estimate-num %d is outside 1..%d", G.n_photons, FRT_KNN_CAP / 2);
    }
    std::vector<GQuery> hq((size_t)n);
    for (int64_t i = 0; i < n; ++i) {
        GQuery &q = hq[(size_t)i];
        q.x = (float)pos[3 * i];
        q.y = (float)pos[3 * i + 1];
        q.z = (float)pos[3 * i + 2];
        q.ex = (float)normal[3 * i];
        q.ey = (float)normal[3 * i + 1];
        q.ez = (float)normal[3 * i + 2];
        q.wr = q.wg = q.wb = 1.0f;
        q.target = (unsigned int)i | (map == 0 ? 0x40000000u : 0u);
    }
    GQuery *dq = nullptr;
    unsigned int *dn = nullptr;
    double *dacc = nullptr;
    int *dfound = nullptr;
    const unsigned int un = (unsigned int)n;
    cudaError_t e = cudaMalloc(&dq, sizeof(GQuery) * (size_t)n);
    if (e == cudaSuccess) e = cudaMalloc(&dn, sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaMalloc(&dacc, sizeof(double) * 3 * (size_t)n);
    if (e == cudaSuccess) e = cudaMalloc(&dfound, sizeof(int) * (size_t)n);
    cudaStream_t s = sc->stream;
    if (e == cudaSuccess) e = cudaMemcpyAsync(dq, hq.data(), sizeof(GQuery) * (size_t)n, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dn, &un, sizeof(un), cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaMemsetAsync(dacc, 0, sizeof(double) * 3 * (size_t)n, s);
    if (e == cudaSuccess) e = cudaMemsetAsync(dfound, 0, sizeof(int) * (size_t)n, s);
    if (e == cudaSuccess) {
        int launches = 0;
        if (launch_knn(sc, s, G, dq, dn, un, dacc, dacc, dfound, frt_env_int("FRT_KNN_MODE", 1), &launches) != FRT_OK) {
            cudaStreamSynchronize(s);
            cudaFree(dq);
            cudaFree(dn);
            cudaFree(dacc);
            cudaFree(dfound);
            return FRT_ERR_CUDA;
        }
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(irrad, dacc, sizeof(double) * 3 * (size_t)n, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess && found != nullptr) e = cudaMemcpyAsync(found, dfound, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    cudaFree(dq);
    cudaFree(dn);
    cudaFree(dacc);
    cudaFree(dfound);
    if (e != cudaSuccess) {
        return frt_set_error(FRT_ERR_CUDA, "frt_photons_estimate: %s", cudaGetErrorString(e));
    }
    return FRT_OK;
}

/*
 * Texture ingest on its own: what frt_scene_create does to every image of a scene.  raw_rgb: the
 * Canvas.arr of an image as the reference's read_png / read_ppm leave it (width * height * 3 doubles, row-major); out:
 * width * height * 4 floats, linear RGB + 0 -- per texel what canvas_pixel_at (canvas.c:115-148) returns, rounded to FP32.
 */
extern "C" int
frt_texture_ingest(const double *raw_rgb, int width, int height, int super_sample, int color_fn, int device, float *out_rgba)
{
    if (raw_rgb == nullptr || out_rgba == nullptr || width <= 0 || height <= 0 || (color_fn != FRT_COLOR_RGB && color_fn != FRT_COLOR_SRGB_TO_RGB)) {
        return frt_set_error(FRT_ERR_ARG, "frt_texture_ingest: bad argument");
    }
    CK(cudaSetDevice(device));
    const size_t n = (size_t)width * height;
    double *raw = nullptr;
    float4 *tex = nullptr;
    cudaError_t e = cudaMalloc(&raw, n * 3 * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&tex, n * sizeof(float4));
    if (e == cudaSuccess) e = cudaMemcpy(raw, raw_rgb, n * 3 * sizeof(double), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
        k_texture_ingest<<<(int)std::min<size_t>((n + 255) / 256, (size_t)sms * 8), 256>>>(raw, width, height, super_sample ? 1 : 0, color_fn, tex);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(out_rgba, tex, n * sizeof(float4), cudaMemcpyDeviceToHost);
    cudaFree(raw);
    cudaFree(tex);
    if (e != cudaSuccess) {
        return frt_set_error(FRT_ERR_CUDA, "frt_texture_ingest: %s", cudaGetErrorString(e));
    }
    return FRT_OK;
}
