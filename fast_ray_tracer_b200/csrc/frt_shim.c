/*
 * frt_shim.c -- host side of the drop-in boundary, in the reference's own language (C11).
 *
 * This translation unit REPLACES src/renderer/renderer.c and src/renderer/photon_tracer.c of
 * gbordelon/fast_ray_tracer when the reference program is linked against libfrt_b200.so.  It is
 * compiled against the reference's own headers (struct layouts are the API: generated main.c pokes
 * struct fields directly), and exports exactly the symbols the rest of the reference program links
 * against (nm-verified):
 *     render_multi, render            renderer.h:46-47   (renderer.c:243, :283)
 *     trace_photons                   photon_tracer.h:4  (photon_tracer.c:203)
 *     array_of_photon_maps            photon_tracer.h:6  (photon_tracer.c:259)
 *     is_shadowed                     renderer.h:42      (referenced by light.c:236,247 only)
 * plus shade_hit / schlick / prepare_computations, which renderer.h declares but nothing references.
 *
 * What it does: walk the finished World / Camera once, flatten them into the SoA description of
 * include/frt_b200.h (pre-order node array that keeps the reference's divided group tree and child
 * order), hand that to the CUDA core, and return an ordinary Canvas.  No rendering happens on the
 * host: if the core fails, the process aborts with a non-zero exit status (no CPU fallback).
 *
 * Environment:
 *   FRT_DEVICES=auto|all|n|a,b,c   GPUs render_multi() splits the frame's row blocks over, inside the library (default auto:
 *                           all visible GPUs for frames of >= FRT_AUTO_SAMPLES primary samples or with a photon pass,
 *                           else one); FRT_DEVICE=n is the one-device spelling
 *   FRT_LIGHT_GEN=0         upload the reference's area-light sample caches instead of rebuilding them on the device
 *   FRT_WARM=0              do not start CUDA on a thread while main() builds the scene
 *   FRT_COUNT_RAYS=1        print the frame's ray counters
 *   FRT_SEED=n              seed of the device RNG (default 0)
 *   FRT_DUMP_SCENE=path     also write the flattened scene as a blob (frt_scene_save)
 *   FRT_DUMP_ONLY=1         write the blob and return a black canvas without touching CUDA
 *   FRT_NO_PRUNE=1          trace zero-weight branches like the reference (ray-count parity)
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdbool.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "src/libs/linalg/linalg.h"
#include "src/libs/photon_map/pm.h"
#include "src/libs/sampler/sampler.h"
#include "src/libs/canvas/canvas.h"
#include "src/color/rgb.h"
#include "src/color/srgb.h"
#include "src/pattern/pattern.h"
#include "src/material/material.h"
#include "src/shapes/shapes.h"
#include "src/light/light.h"
#include "src/renderer/renderer.h"
#include "src/renderer/photon_tracer.h"
#include "src/renderer/world.h"
#include "src/renderer/camera.h"
#include "src/renderer/config.h"

#include <time.h>

#include "frt_b200.h"

/* ------------------------------------------------------------------ growable arrays */

#define VEC(T) struct { T *p; size_t n, cap; }
#define VEC_PUSH(v, T) ((v).n == (v).cap ? ((v).cap = (v).cap ? (v).cap * 2 : 64, \
                        (v).p = (T *)xrealloc((v).p, (v).cap * sizeof(T))) : 0, &(v).p[(v).n++])

static void *
xrealloc(void *p, size_t n)
{
    void *q = realloc(p, n);
    if (q == NULL && n != 0) {
        fprintf(stderr, "frt_shim: out of memory\n");
        exit(70);
    }
    return q;
}

static void
die(const char *what)
{
    fprintf(stderr, "frt_shim: %s: %s\n", what, frt_last_error());
    exit(71);
}

/* pointer -> index map (open addressing), used to share materials / patterns / textures */
struct ptrmap {
    const void **keys;
    int32_t *vals;
    size_t cap, n;
};

static int32_t
ptrmap_get(struct ptrmap *m, const void *k)
{
    if (m->cap == 0) {
        return -1;
    }
    size_t h = ((uintptr_t)k >> 4) * 0x9E3779B97F4A7C15ULL;
    for (size_t i = h & (m->cap - 1);; i = (i + 1) & (m->cap - 1)) {
        if (m->keys[i] == NULL) {
            return -1;
        }
        if (m->keys[i] == k) {
            return m->vals[i];
        }
    }
}

static void
ptrmap_put(struct ptrmap *m, const void *k, int32_t v)
{
    if ((m->n + 1) * 2 > m->cap) {
        struct ptrmap old = *m;
        m->cap = old.cap ? old.cap * 2 : 256;
        m->keys = (const void **)calloc(m->cap, sizeof(void *));
        m->vals = (int32_t *)calloc(m->cap, sizeof(int32_t));
        m->n = 0;
        for (size_t i = 0; i < old.cap; ++i) {
            if (old.keys[i] != NULL) {
                ptrmap_put(m, old.keys[i], old.vals[i]);
            }
        }
        free(old.keys);
        free(old.vals);
    }
    size_t h = ((uintptr_t)k >> 4) * 0x9E3779B97F4A7C15ULL;
    size_t i = h & (m->cap - 1);
    while (m->keys[i] != NULL) {
        i = (i + 1) & (m->cap - 1);
    }
    m->keys[i] = k;
    m->vals[i] = v;
    m->n++;
}

/* ------------------------------------------------------------------ flattener state */

struct flat {
    VEC(frt_node) nodes;
    VEC(int32_t) roots;
    VEC(frt_xform) xforms;
    VEC(double) params;
    VEC(frt_material) materials;
    VEC(frt_pattern) patterns;
    VEC(frt_texture) textures;
    VEC(double) texels;
    VEC(frt_light) lights;
    VEC(double) light_points;
    VEC(double) pixel_samples;
    struct ptrmap mat_map, pat_map, tex_map;
    /* area lights whose sample cache is to be rebuilt on the device (frt_scene_create_gen): their region of light_points
     * is reserved but not written unless the rebuilt sets turn out to differ from the reference's */
    struct lazy_light {
        int32_t light;
        Light l;
        uint64_t drand48_state;
    } lazy[16];
    int n_lazy;
    uint64_t draws; /* drand48 draws the light constructors seen so far have made (light.c:100-191, sampler.c:510-523) */
    bool gen_enabled;
};

static void
copy3(double *dst, const double *src)
{
    dst[0] = src[0];
    dst[1] = src[1];
    dst[2] = src[2];
}

static void
rows3x4(double *dst, const Matrix m)
{
    memcpy(dst, m, 12 * sizeof(double));
}

static int32_t flatten_pattern(struct flat *f, Pattern p);

static int32_t
flatten_texture(struct flat *f, Canvas c)
{
    int32_t idx = ptrmap_get(&f->tex_map, c);
    if (idx >= 0) {
        return idx;
    }
    idx = (int32_t)f->textures.n;
    frt_texture *t = VEC_PUSH(f->textures, frt_texture);
    memset(t, 0, sizeof(*t));
    t->width = (int32_t)c->width;
    t->height = (int32_t)c->height;
    t->super_sample = c->super_sample ? 1 : 0;
    if (c->color_space_fn == rgb_to_rgb) {
        t->color_fn = FRT_COLOR_RGB;
    } else if (c->color_space_fn == srgb_to_rgb) {
        t->color_fn = FRT_COLOR_SRGB_TO_RGB;
    } else {
        fprintf(stderr, "frt_shim: texture colour space other than RGB/sRGB is not supported on the device\n");
        exit(72);
    }
    t->texel_offset = (int64_t)(f->texels.n / 3);
    size_t n = c->width * c->height;
    for (size_t i = 0; i < n; ++i) {
        for (int k = 0; k < 3; ++k) {
            *VEC_PUSH(f->texels, double) = c->arr[i][k];
        }
    }
    ptrmap_put(&f->tex_map, c, idx);
    return idx;
}

static int
uv_map_faces(enum uv_map_type t)
{
    switch (t) {
    case CUBE_UV_MAP:
        return 6;
    case CYLINDER_UV_MAP:
        return 3;
    default:
        return 1;
    }
}

static void
fill_pattern(struct flat *f, int32_t idx, Pattern p)
{
    frt_pattern q;
    memset(&q, 0, sizeof(q));
    q.type = (int32_t)p->type;
    q.identity = p->transform_identity ? 1 : 0;
    rows3x4(q.inv, p->transform_inverse);
    q.i[0] = q.i[1] = q.i[2] = q.i[3] = -1;
    switch (p->type) {
    case CHECKER_PATTERN:
    case GRADIENT_PATTERN:
    case RADIAL_GRADIENT_PATTERN:
    case RING_PATTERN:
    case STRIPE_PATTERN:
    case UV_GRADIENT_PATTERN:
    case UV_RADIAL_GRADIENT_PATTERN:
        copy3(q.c + 0, p->fields.concrete.a);
        copy3(q.c + 3, p->fields.concrete.b);
        break;
    case UV_ALIGN_CHECKER_PATTERN:
        copy3(q.c + 0, p->fields.uv_align_check.main);
        copy3(q.c + 3, p->fields.uv_align_check.ul);
        copy3(q.c + 6, p->fields.uv_align_check.ur);
        copy3(q.c + 9, p->fields.uv_align_check.bl);
        copy3(q.c + 12, p->fields.uv_align_check.br);
        break;
    case UV_CHECKER_PATTERN:
        copy3(q.c + 0, p->fields.uv_check.a);
        copy3(q.c + 3, p->fields.uv_check.b);
        q.i[0] = (int32_t)p->fields.uv_check.width;
        q.i[1] = (int32_t)p->fields.uv_check.height;
        break;
    case UV_TEXTURE_PATTERN:
        q.i[0] = flatten_texture(f, p->fields.uv_texture.canvas);
        break;
    case BLENDED_PATTERN:
        q.i[0] = flatten_pattern(f, p->fields.blended.pattern1);
        q.i[1] = flatten_pattern(f, p->fields.blended.pattern2);
        break;
    case NESTED_PATTERN:
        q.i[0] = flatten_pattern(f, p->fields.nested.pattern1);
        q.i[1] = flatten_pattern(f, p->fields.nested.pattern2);
        q.i[2] = flatten_pattern(f, p->fields.nested.pattern3);
        break;
    case PERTURBED_PATTERN:
        q.i[0] = flatten_pattern(f, p->fields.perturbed.pattern1);
        q.i[1] = (int32_t)p->fields.perturbed.octaves;
        q.i[2] = (int32_t)p->fields.perturbed.seed;
        q.f[0] = p->fields.perturbed.frequency;
        q.f[1] = p->fields.perturbed.scale_factor;
        q.f[2] = p->fields.perturbed.persistence;
        break;
    case CUBE_MAP_PATTERN:
    case CYLINDER_MAP_PATTERN:
    case TEXTURE_MAP_PATTERN: {
        int nf = uv_map_faces(p->fields.uv_map.type);
        q.i[0] = (int32_t)p->fields.uv_map.type;
        /* faces are a contiguous array in the reference (pattern.h:103-106): keep them contiguous */
        int32_t first = (int32_t)f->patterns.n;
        for (int k = 0; k < nf; ++k) {
            frt_pattern *slot = VEC_PUSH(f->patterns, frt_pattern);
            memset(slot, 0, sizeof(*slot));
        }
        for (int k = 0; k < nf; ++k) {
            fill_pattern(f, first + k, p->fields.uv_map.uv_faces + k);
        }
        q.i[1] = first;
        break;
    }
    default:
        fprintf(stderr, "frt_shim: unknown pattern type %d\n", (int)p->type);
        exit(72);
    }
    f->patterns.p[idx] = q;
}

static int32_t
flatten_pattern(struct flat *f, Pattern p)
{
    if (p == NULL) {
        return -1;
    }
    int32_t idx = ptrmap_get(&f->pat_map, p);
    if (idx >= 0) {
        return idx;
    }
    idx = (int32_t)f->patterns.n;
    frt_pattern *slot = VEC_PUSH(f->patterns, frt_pattern);
    memset(slot, 0, sizeof(*slot));
    ptrmap_put(&f->pat_map, p, idx);
    fill_pattern(f, idx, p);
    return idx;
}

static int32_t
flatten_material(struct flat *f, Material m)
{
    if (m == NULL) {
        return -1;
    }
    int32_t idx = ptrmap_get(&f->mat_map, m);
    if (idx >= 0) {
        return idx;
    }
    frt_material q;
    memset(&q, 0, sizeof(q));
    copy3(q.Ka, m->Ka);
    copy3(q.Kd, m->Kd);
    copy3(q.Ks, m->Ks);
    copy3(q.Tf, m->Tf);
    copy3(q.refl, m->refl);
    q.Ns = m->Ns;
    q.Ni = m->Ni;
    q.Tr = m->Tr;
    q.casts_shadow = m->casts_shadow ? 1 : 0;
    q.reflective = m->reflective ? 1 : 0;
    q.map_Ka = flatten_pattern(f, m->map_Ka);
    q.map_Kd = flatten_pattern(f, m->map_Kd);
    q.map_Ks = flatten_pattern(f, m->map_Ks);
    q.map_Ns = flatten_pattern(f, m->map_Ns);
    q.map_d = flatten_pattern(f, m->map_d);
    q.map_bump = flatten_pattern(f, m->map_bump);
    q.map_refl = flatten_pattern(f, m->map_refl);
    idx = (int32_t)f->materials.n;
    *VEC_PUSH(f->materials, frt_material) = q;
    ptrmap_put(&f->mat_map, m, idx);
    return idx;
}

/*
 * Pre-order walk.  `pinv` is the composite world->parent-local matrix (4x4) and `pxf` its index.
 * A node whose transform_identity flag is set inherits its parent's composite, exactly like
 * shape_intersect skips the ray transform (shapes.c:47) and shape_world_to_object /
 * shape_normal_to_world skip theirs (shapes.c:117-131, :92-114).
 */
static int32_t
flatten_shape(struct flat *f, Shape s, int32_t parent, const Matrix pinv, int32_t pxf)
{
    int32_t idx = (int32_t)f->nodes.n;
    frt_node *slot = VEC_PUSH(f->nodes, frt_node);
    memset(slot, 0, sizeof(*slot));

    frt_node n;
    memset(&n, 0, sizeof(n));
    n.type = (int32_t)s->type;
    n.parent = parent;
    n.material = -1;
    n.param = -1;
    n.right = -1;

    Matrix comp;
    int32_t xf = pxf;
    if (s->transform_identity) {
        matrix_copy(pinv, comp);
    } else {
        matrix_multiply(s->transform_inverse, pinv, comp);
        xf = (int32_t)f->xforms.n;
        frt_xform *x = VEC_PUSH(f->xforms, frt_xform);
        rows3x4(x->inv, comp);
    }
    n.xform = xf;

    switch (s->type) {
    case SHAPE_GROUP: {
        Bounding_box box;
        s->bounds(s, &box);
        copy3(n.bbox_min, box.min);
        copy3(n.bbox_max, box.max);
        for (size_t i = 0; i < s->fields.group.num_children; ++i) {
            flatten_shape(f, s->fields.group.children + i, idx, comp, xf);
        }
        break;
    }
    case SHAPE_CSG: {
        Bounding_box box;
        s->bounds(s, &box);
        copy3(n.bbox_min, box.min);
        copy3(n.bbox_max, box.max);
        n.csg_op = (int32_t)s->fields.csg.op;
        flatten_shape(f, s->fields.csg.left, idx, comp, xf);
        n.right = flatten_shape(f, s->fields.csg.right, idx, comp, xf);
        break;
    }
    case SHAPE_CYLINDER:
    case SHAPE_CONE:
        n.param = (int32_t)f->params.n;
        *VEC_PUSH(f->params, double) = s->fields.cylinder.minimum;
        *VEC_PUSH(f->params, double) = s->fields.cylinder.maximum;
        *VEC_PUSH(f->params, double) = s->fields.cylinder.closed ? 1.0 : 0.0;
        n.material = flatten_material(f, s->material);
        break;
    case SHAPE_TOROID:
        n.param = (int32_t)f->params.n;
        *VEC_PUSH(f->params, double) = s->fields.toroid.r1;
        *VEC_PUSH(f->params, double) = s->fields.toroid.r2;
        n.material = flatten_material(f, s->material);
        break;
    case SHAPE_TRIANGLE:
    case SHAPE_SMOOTH_TRIANGLE: {
        n.param = (int32_t)f->params.n;
        const double *src[11];
        src[0] = s->fields.triangle.p1;
        src[1] = s->fields.triangle.p2;
        src[2] = s->fields.triangle.p3;
        src[3] = s->fields.triangle.e1;
        src[4] = s->fields.triangle.e2;
        if (s->type == SHAPE_TRIANGLE) {
            src[5] = src[6] = src[7] = s->fields.triangle.u_normals.normal;
        } else {
            src[5] = s->fields.triangle.u_normals.s_normals.n1;
            src[6] = s->fields.triangle.u_normals.s_normals.n2;
            src[7] = s->fields.triangle.u_normals.s_normals.n3;
        }
        src[8] = s->fields.triangle.t1;
        src[9] = s->fields.triangle.t2;
        src[10] = s->fields.triangle.t3;
        for (int k = 0; k < 11; ++k) {
            /* t1..t3 are uninitialised in the reference unless use_textures is set */
            bool live = k < 8 || s->fields.triangle.use_textures;
            for (int c = 0; c < 3; ++c) {
                *VEC_PUSH(f->params, double) = live ? src[k][c] : 0.0;
            }
        }
        *VEC_PUSH(f->params, double) = s->fields.triangle.use_textures ? 1.0 : 0.0;
        n.material = flatten_material(f, s->material);
        break;
    }
    default: /* sphere, cube, plane */
        n.material = flatten_material(f, s->material);
        break;
    }

    n.skip = (int32_t)f->nodes.n;
    f->nodes.p[idx] = n;
    return idx;
}

struct copy_job {
    Light l;
    double *dst;
    size_t s0, s1;
};

static void *
copy_light_sets(void *arg)
{
    const struct copy_job *j = (const struct copy_job *)arg;
    const size_t per_set = 3 * (size_t)j->l->num_samples;
    for (size_t s = j->s0; s < j->s1; ++s) {
        const Points pts = j->l->surface_points_cache + s;
        double *d = j->dst + s * per_set;
        for (size_t k = 0; k < pts->points_num; ++k) {
            d[0] = pts->points[k][0];
            d[1] = pts->points[k][1];
            d[2] = pts->points[k][2];
            d += 3;
        }
    }
    return NULL;
}

static void
flatten_light(struct flat *f, Light l)
{
    frt_light q;
    memset(&q, 0, sizeof(q));
    q.type = (int32_t)l->type;
    q.num_samples = (int32_t)l->num_samples;
    q.cache_len = (int32_t)l->surface_points_cache_len;
    copy3(q.intensity, l->intensity);
    switch (l->type) {
    case AREA_LIGHT: {
        Vector tmp, nrm;
        copy3(q.position, l->u.area.corner);
        copy3(q.uvec, l->u.area.uvec);
        copy3(q.vvec, l->u.area.vvec);
        q.usteps = (int32_t)l->u.area.usteps;
        q.vsteps = (int32_t)l->u.area.vsteps;
        q.jitter = l->u.area.jitter ? 1 : 0;
        vector_cross(l->u.area.uvec, l->u.area.vvec, tmp); /* light.c:59-61 */
        vector_normalize(tmp, nrm);
        copy3(q.normal, nrm);
        break;
    }
    case CIRCLE_LIGHT:
        copy3(q.position, l->u.circle.origin);
        copy3(q.normal, l->u.circle.normal);
        q.radius = l->u.circle.radius;
        q.usteps = (int32_t)l->u.circle.usteps;
        q.vsteps = (int32_t)l->u.circle.vsteps;
        q.jitter = l->u.circle.jitter ? 1 : 0;
        break;
    case HEMISPHERE_LIGHT:
        copy3(q.position, l->u.hemi.position);
        copy3(q.normal, l->u.hemi.normal);
        break;
    default:
        copy3(q.position, l->u.point.position);
        break;
    }
    q.point_offset = (int64_t)(f->light_points.n / 3);
    /* the whole cache at once (157 MB for the shipped Cornell light): one reservation, no per-value growth check */
    const size_t total = 3 * (size_t)l->surface_points_cache_len * (size_t)l->num_samples;
    for (size_t s = 0; s < l->surface_points_cache_len; ++s) {
        if (l->surface_points_cache[s].points_num != l->num_samples) {
            fprintf(stderr, "frt_shim: light sample set %zu has %zu points, expected %zu\n", s, l->surface_points_cache[s].points_num,
                    l->num_samples);
            exit(72);
        }
    }
    /* The constructors of jittered area / circle lights are the program's drand48 consumers before render_multi()
     * (sampler_2d draws one table when it is created, then one per cached set: light.c:166-171, sampler.c:510-523), in
     * world order.  An area light's cache is a pure function of the generator state in front of its first set, so the
     * core rebuilds it on the device and compares a few sets with the reference's bit for bit (frt_scene_create_gen);
     * the 157 MB of the shipped Cornell light are then neither repacked here (115 ms) nor copied over PCIe. */
    bool lazy = false;
    if ((l->type == AREA_LIGHT && l->u.area.jitter) || (l->type == CIRCLE_LIGHT && l->u.circle.jitter)) {
        const uint64_t per_set = 2 * (uint64_t)q.usteps * q.vsteps + q.usteps + q.vsteps;
        if (l->type == AREA_LIGHT && f->gen_enabled && f->n_lazy < 16 && q.cache_len >= 8 && q.usteps <= 64 && q.vsteps <= 64 &&
            (size_t)q.usteps * q.vsteps == l->num_samples) {
            lazy = true;
            f->lazy[f->n_lazy].light = (int32_t)f->lights.n;
            f->lazy[f->n_lazy].l = l;
            f->lazy[f->n_lazy].drand48_state = frt_drand48_advance(0, f->draws + per_set);
            f->n_lazy += 1;
        }
        f->draws += per_set * ((uint64_t)q.cache_len + 1);
    }
    if (f->light_points.n + total > f->light_points.cap) {
        /* calloc: a large block comes straight from mmap, so the pages of a region that is never written (a lazy light's)
         * are never touched */
        double *np = (double *)calloc(f->light_points.n + total, sizeof(double));
        if (np == NULL) {
            fprintf(stderr, "frt_shim: out of memory\n");
            exit(70);
        }
        if (f->light_points.n) {
            memcpy(np, f->light_points.p, f->light_points.n * sizeof(double));
        }
        free(f->light_points.p);
        f->light_points.p = np;
        f->light_points.cap = f->light_points.n + total;
    }
    if (!lazy) {
        /* separately allocated sets of 32-byte points -> one array of 24-byte points */
        struct copy_job job = { l, f->light_points.p + f->light_points.n, 0, l->surface_points_cache_len };
        copy_light_sets(&job);
    }
    f->light_points.n += total;
    *VEC_PUSH(f->lights, frt_light) = q;
}

/* a lazy light's sets after all: the rebuilt cache differed from the reference's (FRT_ERR_MISMATCH) */
static void
materialize_lazy_lights(struct flat *f)
{
    for (int k = 0; k < f->n_lazy; ++k) {
        Light l = f->lazy[k].l;
        struct copy_job job = { l, f->light_points.p + 3 * (size_t)f->lights.p[f->lazy[k].light].point_offset, 0, l->surface_points_cache_len };
        copy_light_sets(&job);
    }
    f->n_lazy = 0;
}

static void
flatten_world(struct flat *f, Camera cam, World w, size_t usteps, size_t vsteps, bool jitter, frt_scene_desc *d)
{
    memset(f, 0, sizeof(*f));
    {
        const char *e = getenv("FRT_LIGHT_GEN");
        f->gen_enabled = !(e != NULL && e[0] == '0') && !(getenv("FRT_DUMP_SCENE") != NULL && *getenv("FRT_DUMP_SCENE"));
    }
    /* the generated main() builds the camera before the lights (yaml_parser.py:182-186); a jittered aperture's sampler
     * has drawn one table by then (camera.c:170-175) */
    if (cam->aperture.jitter && cam->aperture.sampler.steps_by_dimension != NULL) {
        const uint64_t au = cam->aperture.sampler.steps_by_dimension[0], av = cam->aperture.sampler.steps_by_dimension[1];
        f->draws = 2 * au * av + au + av;
    }
    frt_xform *ident = VEC_PUSH(f->xforms, frt_xform);
    rows3x4(ident->inv, MATRIX_IDENTITY);

    for (size_t i = 0; i < w->shapes_num; ++i) {
        *VEC_PUSH(f->roots, int32_t) = flatten_shape(f, w->shapes + i, -1, MATRIX_IDENTITY, 0);
    }
    for (size_t i = 0; i < w->lights_num; ++i) {
        flatten_light(f, w->lights + i);
    }

    memset(d, 0, sizeof(*d));
    d->abi_version = FRT_ABI_VERSION;

    frt_camera *c = &d->camera;
    c->hsize = (int32_t)cam->hsize;
    c->vsize = (int32_t)cam->vsize;
    c->usteps = (int32_t)usteps;
    c->vsteps = (int32_t)vsteps;
    c->half_width = cam->half_width;
    c->half_height = cam->half_height;
    c->pixel_size = cam->pixel_size;
    c->canvas_distance = cam->canvas_distance;
    memcpy(c->inv, cam->transform_inverse, sizeof(Matrix));
    c->aperture_type = (int32_t)cam->aperture.type;
    c->aperture_jitter = jitter ? 1 : 0;
    c->aperture_size = cam->aperture.size;
    memcpy(c->aperture_args, &cam->aperture.u, sizeof(cam->aperture.u) < sizeof(c->aperture_args) ? sizeof(cam->aperture.u) : sizeof(c->aperture_args));

    if (!jitter) {
        /* the per-pixel table is the same for every pixel when xi == 0.5 (sampler.c:401-461): build it with the
         * reference's own sampler so odd grids keep its index conventions */
        struct sampler sm;
        sampler_2d(false, usteps, vsteps, sampler_default_constraint, &sm);
        for (size_t k = 0; k < 2 * usteps * vsteps; ++k) {
            *VEC_PUSH(f->pixel_samples, double) = sm.arr[k];
        }
        sampler_free(&sm);
    }

    const struct global_config *g = w->global_config;
    frt_config *q = &d->config;
    q->include_direct = g->illumination.include_direct;
    q->include_global = g->illumination.include_global;
    q->visualize_photon_map = g->illumination.debug_visualize_photon_map;
    q->visualize_soft_indirect = g->illumination.debug_visualize_soft_indirect;
    q->di_include_ambient = g->illumination.di.include_ambient;
    q->di_include_diffuse = g->illumination.di.include_diffuse;
    q->di_include_specular_highlight = g->illumination.di.include_specular_highlight;
    q->di_include_specular = g->illumination.di.include_specular;
    q->di_path_length = (int32_t)g->illumination.di.path_length;
    q->gi_include_caustics = g->illumination.gi.include_caustics;
    q->gi_include_final_gather = g->illumination.gi.include_final_gather;
    q->gi_usteps = (int32_t)g->illumination.gi.usteps;
    q->gi_vsteps = (int32_t)g->illumination.gi.vsteps;
    q->gi_irradiance_estimate_num = (int32_t)g->illumination.gi.irradiance_estimate_num;
    q->gi_path_length = (int32_t)g->illumination.gi.path_length;
    q->gi_irradiance_estimate_radius = g->illumination.gi.irradiance_estimate_radius;
    q->gi_irradiance_estimate_cone_filter_k = g->illumination.gi.irradiance_estimate_cone_filter_k;
    q->gi_photon_count = (int64_t)g->illumination.gi.photon_count;

    d->n_nodes = (int32_t)f->nodes.n;
    d->n_roots = (int32_t)f->roots.n;
    d->n_xforms = (int32_t)f->xforms.n;
    d->n_materials = (int32_t)f->materials.n;
    d->n_patterns = (int32_t)f->patterns.n;
    d->n_textures = (int32_t)f->textures.n;
    d->n_lights = (int32_t)f->lights.n;
    d->n_prim_params = (int64_t)f->params.n;
    d->n_texels = (int64_t)(f->texels.n / 3);
    d->n_light_points = (int64_t)(f->light_points.n / 3);
    d->n_pixel_samples = (int64_t)f->pixel_samples.n;
    d->nodes = f->nodes.p;
    d->roots = f->roots.p;
    d->xforms = f->xforms.p;
    d->prim_params = f->params.p;
    d->materials = f->materials.p;
    d->patterns = f->patterns.p;
    d->textures = f->textures.p;
    d->texels = f->texels.p;
    d->lights = f->lights.p;
    d->light_points = f->light_points.p;
    d->pixel_samples = f->pixel_samples.n ? f->pixel_samples.p : NULL;
}

static void
flat_free(struct flat *f)
{
    free(f->nodes.p);
    free(f->roots.p);
    free(f->xforms.p);
    free(f->params.p);
    free(f->materials.p);
    free(f->patterns.p);
    free(f->textures.p);
    free(f->texels.p);
    free(f->lights.p);
    free(f->light_points.p);
    free(f->pixel_samples.p);
    free(f->mat_map.keys);
    free(f->mat_map.vals);
    free(f->pat_map.keys);
    free(f->pat_map.vals);
    free(f->tex_map.keys);
    free(f->tex_map.vals);
    memset(f, 0, sizeof(*f));
}

/* ------------------------------------------------------------------ the boundary */

static long
env_long(const char *name, long dflt)
{
    const char *s = getenv(name);
    return (s == NULL || *s == '\0') ? dflt : strtol(s, NULL, 10);
}

/* the scene(s) and the Canvas of the last frame: with one device its pixels are still there when main() asks for the PPM */
static frt_multi *g_last_multi;
static Canvas g_last_image;
static uint64_t g_last_digest;

static void
drop_last_scene(void)
{
    if (g_last_multi != NULL) {
        frt_multi_destroy(g_last_multi);
        g_last_multi = NULL;
        g_last_image = NULL;
    }
}

/* sampled digest of a canvas (every 16th pixel): frt_shim_write_ppm_file only encodes the device copy of the frame when
 * the host Canvas still holds what render_multi() returned */
static uint64_t
canvas_digest(Canvas c)
{
    uint64_t acc = 0x9E3779B97F4A7C15ULL;
    const size_t n = c->width * c->height;
    for (size_t i = 0; i < n; i += 16) {
        for (int k = 0; k < 3; ++k) {
            uint64_t bits;
            memcpy(&bits, &c->arr[i][k], sizeof(bits));
            acc = (acc ^ bits) * 0x100000001B3ULL + i;
        }
    }
    return acc;
}

/* a trace_photons() request waiting for render_multi() (see trace_photons below) */
static bool g_photons_pending, g_photons_caustic, g_photons_global;
static World g_photons_world;

/*
 * CUDA start-up behind the host's scene construction.  The generated main() spends its first half second building the
 * World (the reference's own light-cache constructor alone: 0.5 s for the shipped Cornell light) before it reaches
 * render_multi(); driver initialisation and the first device's context take as long or longer (1.3 s on a 1-GPU box,
 * 5 s on an 8-GPU box) and depend on nothing the host is doing.  A constructor of the shim starts them on a thread;
 * render_multi() joins it.  FRT_WARM=0 switches it off (FRT_DUMP_ONLY runs never touch CUDA).
 */
#include <pthread.h>
static pthread_t g_warm_thread;
static bool g_warm_started;

static void *
warm_main(void *arg)
{
    (void)arg;
    if (frt_device_count() > 0) {
        const char *one = getenv("FRT_DEVICE");
        frt_trim(one != NULL && *one != '\0' ? (int)strtol(one, NULL, 10) : 0); /* cudaSetDevice: creates that device's context */
    }
    return NULL;
}

__attribute__((constructor)) static void
warm_start(void)
{
    const char *w = getenv("FRT_WARM"), *dump = getenv("FRT_DUMP_ONLY");
    if ((w != NULL && w[0] == '0') || (dump != NULL && dump[0] != '\0' && dump[0] != '0')) {
        return;
    }
    g_warm_started = pthread_create(&g_warm_thread, NULL, warm_main, NULL) == 0;
}

static void
warm_join(void)
{
    if (g_warm_started) {
        pthread_join(g_warm_thread, NULL);
        g_warm_started = false;
    }
}

static double
ms_since(const struct timespec *t0)
{
    struct timespec t1;
    clock_gettime(CLOCK_MONOTONIC, &t1);
    return 1e3 * (double)(t1.tv_sec - t0->tv_sec) + 1e-6 * (double)(t1.tv_nsec - t0->tv_nsec);
}

/*
 * FRT_DEVICES: "auto" (default), "all", a count ("4"), or a list of ordinals ("0,2,3"); FRT_DEVICE=n is the one-device
 * spelling.  auto: every visible GPU when the frame is worth it, one otherwise -- a CUDA context costs ~0.25 s per device
 * (measured on an 8 x B200 box: scene creation 0.29 s on one device, 1.8 s on eight, for a Cornell frame of 15 ms), so a
 * one-shot program only gains from more devices when its frame takes longer than that: FRT_AUTO_SAMPLES (default 2^28)
 * primary samples, or a photon pass.
 */
static int
pick_devices(int32_t *out, int cap, double primary_samples, bool photon_pass)
{
    const int visible = frt_device_count();
    const char *e = getenv("FRT_DEVICES");
    int n = 0;
    if (e == NULL || *e == '\0') {
        const char *one = getenv("FRT_DEVICE");
        if (one != NULL && *one != '\0') {
            out[0] = (int32_t)strtol(one, NULL, 10);
            return 1;
        }
        e = "auto";
    }
    if (strcmp(e, "auto") == 0) {
        e = (photon_pass || primary_samples >= (double)env_long("FRT_AUTO_SAMPLES", 1L << 28)) ? "all" : "1";
    }
    if (strcmp(e, "all") == 0) {
        for (n = 0; n < visible && n < cap; ++n) out[n] = n;
    } else if (strchr(e, ',') != NULL) {
        const char *p = e;
        while (*p != '\0' && n < cap) {
            char *end;
            long v = strtol(p, &end, 10);
            if (end == p) break;
            out[n++] = (int32_t)v;
            p = (*end == ',') ? end + 1 : end;
        }
    } else {
        long want = strtol(e, NULL, 10);
        for (n = 0; n < want && n < visible && n < cap; ++n) out[n] = n;
    }
    if (n == 0) {
        out[0] = 0;
        n = 1;
    }
    return n;
}

static Canvas
render_on_device(Camera cam, World w, size_t usteps, size_t vsteps, bool jitter)
{
    struct flat f;
    frt_scene_desc d;
    struct timespec t0, t_all;
    clock_gettime(CLOCK_MONOTONIC, &t_all);
    t0 = t_all;
    flatten_world(&f, cam, w, usteps, vsteps, jitter, &d);
    /* the one walk over the finished World that replaces world_copy x threads */
    printf("FRT_B200_FLATTEN_MS %.3f (%d nodes, %lld light points, %d light caches left to the device)\n", ms_since(&t0), (int)d.n_nodes,
           (long long)d.n_light_points, f.n_lazy);

    Canvas image = canvas_alloc(cam->hsize, cam->vsize, false, NULL); /* renderer.c:250 */

    const char *dump = getenv("FRT_DUMP_SCENE");
    if (dump != NULL && *dump != '\0') {
        if (frt_scene_save(&d, dump) != FRT_OK) {
            die("frt_scene_save");
        }
    }
    if (env_long("FRT_DUMP_ONLY", 0)) {
        memset(image->arr, 0, cam->hsize * cam->vsize * sizeof(Color));
        flat_free(&f);
        return image;
    }

    int32_t devices[64];
    warm_join();
    const int n_dev = pick_devices(devices, 64, (double)cam->hsize * (double)cam->vsize * (double)usteps * (double)vsteps,
                                   g_photons_pending && g_photons_world == w);

    /* light caches rebuilt on the device, checked against a few of the sets the reference built (first, last, spread) */
    frt_light_gen gens[16];
    double *verify_buf = NULL;
    if (f.n_lazy > 0) {
        size_t words = 0;
        for (int k = 0; k < f.n_lazy; ++k) {
            words += (size_t)FRT_GEN_VERIFY_MAX * 3 * f.lazy[k].l->num_samples;
        }
        verify_buf = (double *)xrealloc(NULL, words * sizeof(double));
        double *vb = verify_buf;
        for (int k = 0; k < f.n_lazy; ++k) {
            Light l = f.lazy[k].l;
            frt_light_gen *g = &gens[k];
            memset(g, 0, sizeof(*g));
            g->light = f.lazy[k].light;
            g->drand48_state = f.lazy[k].drand48_state;
            const size_t len = l->surface_points_cache_len;
            g->n_verify = (int32_t)(len < FRT_GEN_VERIFY_MAX ? len : FRT_GEN_VERIFY_MAX);
            for (int v = 0; v < g->n_verify; ++v) {
                const size_t set = g->n_verify == 1 ? 0 : (size_t)v * (len - 1) / (size_t)(g->n_verify - 1);
                struct copy_job job = { l, vb - set * 3 * l->num_samples, set, set + 1 };
                copy_light_sets(&job);
                g->verify_set[v] = (int32_t)set;
                g->verify_points[v] = vb;
                vb += 3 * l->num_samples;
            }
        }
    }
    clock_gettime(CLOCK_MONOTONIC, &t0);
    frt_multi *multi = NULL;
    int rc = frt_multi_create(&d, devices, n_dev, f.n_lazy ? gens : NULL, f.n_lazy, &multi);
    for (int k = 0; rc == FRT_OK && f.n_lazy > 0 && k < frt_multi_device_count(multi); ++k) {
        rc = frt_scene_gen_status(frt_multi_scene(multi, k)); /* waits for the comparison on that device */
    }
    if (rc == FRT_ERR_MISMATCH) {
        frt_multi_destroy(multi);
        multi = NULL;
        /* another generator state than the constructors' order implies (or another sampler): take the reference's sets */
        printf("FRT_B200_LIGHT_GEN mismatch (%s): uploading the host caches\n", frt_last_error());
        materialize_lazy_lights(&f);
        rc = frt_multi_create(&d, devices, n_dev, NULL, 0, &multi);
    } else if (f.n_lazy > 0 && rc == FRT_OK) {
        printf("FRT_B200_LIGHT_GEN %d light caches rebuilt on the device, %d sets each compared bit for bit\n", f.n_lazy, (int)gens[0].n_verify);
    }
    free(verify_buf);
    if (rc != FRT_OK) {
        die("frt_multi_create");
    }
    printf("FRT_B200_CREATE_MS %.3f (%d devices)\n", ms_since(&t0), frt_multi_device_count(multi));

    if (g_photons_pending && g_photons_world == w) {
        frt_photon_cfg pc;
        memset(&pc, 0, sizeof(pc));
        pc.populate_caustic = g_photons_caustic ? 1 : 0;
        pc.populate_global = g_photons_global ? 1 : 0;
        pc.seed = (uint64_t)env_long("FRT_SEED", 0);
        frt_stats ps;
        memset(&ps, 0, sizeof(ps));
        clock_gettime(CLOCK_MONOTONIC, &t0);
        if (frt_multi_photons(multi, &pc, &ps) != FRT_OK) {
            die("frt_multi_photons");
        }
        printf("FRT_B200_PHOTONS emitted %llu stored caustic %llu global %llu in %.3f ms\n", (unsigned long long)ps.rays_photon,
               (unsigned long long)ps.photons_stored[0], (unsigned long long)ps.photons_stored[1], ms_since(&t0));
        g_photons_pending = false;
    }

    frt_render_cfg cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.usteps = (int32_t)usteps;
    cfg.vsteps = (int32_t)vsteps;
    cfg.jitter = jitter ? 1 : 0;
    cfg.seed = (uint64_t)env_long("FRT_SEED", 0);
    cfg.flags = (env_long("FRT_COUNT_RAYS", 0) ? FRT_FLAG_COUNT_RAYS : 0) | (env_long("FRT_NO_PRUNE", 0) ? FRT_FLAG_NO_PRUNE : 0);

    frt_stats st;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    if (frt_multi_render(multi, &cfg, (double *)image->arr, &st) != FRT_OK) {
        die("frt_multi_render");
    }
    printf("FRT_B200_FRAME_MS %.3f (render call %.3f ms, canvas download %.3f ms)\n", st.frame_ms, ms_since(&t0), st.download_ms);
    if (cfg.flags & FRT_FLAG_COUNT_RAYS) {
        printf("FRT_B200_RAYS primary %llu secondary %llu shadow %llu gather %llu\n",
               (unsigned long long)st.rays_primary, (unsigned long long)st.rays_secondary,
               (unsigned long long)st.rays_shadow, (unsigned long long)st.rays_gather);
    }

    /* keep the frame on the device until the next one (or exit): frt_shim_write_ppm_file encodes it there */
    static bool registered;
    drop_last_scene();
    g_last_multi = multi;
    g_last_image = image;
    g_last_digest = canvas_digest(image);
    if (!registered) {
        atexit(drop_last_scene);
        registered = true;
    }
    flat_free(&f);
    printf("FRT_B200_RENDER_MULTI_MS %.3f\n", ms_since(&t_all));
    fflush(stdout);
    return image;
}

/*
 * write_ppm_file (src/libs/canvas/canvas.c:305-328) for the Canvas render_multi() just returned: construct_ppm's three
 * host passes (two pow() per channel and pixel, 0.2 s for 800 x 800 on one thread) become three kernels over the frame
 * that is still on the device (frt_canvas_encode_ppm16); the host writes the same bytes to <file_path>.ppm.  Returns 0
 * when it wrote the file, -1 when `c` is not that canvas (the caller then runs the reference's own write_ppm_file).
 * The generated main() reaches it through a three-line change in canvas.c or `-Wl,--wrap=write_ppm_file`, see
 * INTEGRATION.md.
 */
int
frt_shim_write_ppm_file(Canvas c, const bool use_scaling, const char *file_path)
{
    if (c == NULL || c != g_last_image || g_last_multi == NULL || file_path == NULL || canvas_digest(c) != g_last_digest) {
        return -1; /* not the frame render_multi() returned, or the caller changed it since */
    }
    size_t cap = frt_ppm16_size((int)c->width, (int)c->height), len = 0;
    unsigned char *buf = (unsigned char *)malloc(cap);
    double ms = 0.0;
    int rc;
    if (buf == NULL) {
        die("out of memory");
    }
    if (frt_multi_device_count(g_last_multi) == 1) {
        rc = frt_canvas_encode_ppm16(frt_multi_scene(g_last_multi, 0), use_scaling ? 1 : 0, buf, cap, &len, &ms);
    } else { /* every device holds its own rows only: encode the assembled host canvas on device 0 */
        rc = frt_encode_ppm16((const double *)c->arr, (int)c->width, (int)c->height, use_scaling ? 1 : 0, 0, buf, cap, &len, &ms);
    }
    if (rc != FRT_OK) {
        free(buf);
        die("frt_canvas_encode_ppm16");
    }
    size_t n = strlen(file_path);
    char *full = (char *)malloc(n + 5);
    if (full == NULL) {
        die("out of memory");
    }
    memcpy(full, file_path, n);
    memcpy(full + n, ".ppm", 5);
    FILE *fp = fopen(full, "wb");
    if (fp == NULL) {
        fprintf(stderr, "frt_shim: cannot open %s\n", full);
        exit(3);
    }
    fwrite(buf, 1, len, fp);
    fclose(fp);
    printf("FRT_B200_PPM_MS %.3f\n", ms);
    free(full);
    free(buf);
    return 0;
}

Canvas
render_multi(Camera cam, World w, size_t usteps, size_t vsteps, bool jitter)
{
    return render_on_device(cam, w, usteps, vsteps, jitter);
}

Canvas
render(Camera cam, World w, size_t usteps, size_t vsteps, bool jitter)
{
    return render_on_device(cam, w, usteps, vsteps, jitter);
}

/* trace_photons() runs before the camera exists (yaml_parser.py:201-218), so the shim only records the request; the
 * photons are traced on the device right after the scene has been uploaded in render_multi(). */

void
trace_photons(const World w, size_t num_maps, bool populate_caustic_map, bool populate_global_map)
{
    (void)num_maps;
    g_photons_pending = true;
    g_photons_caustic = populate_caustic_map;
    g_photons_global = populate_global_map;
    g_photons_world = w;
}

PhotonMap *
array_of_photon_maps(size_t num)
{
    return (PhotonMap *)malloc(num * sizeof(PhotonMap)); /* photon_tracer.c:259-263 */
}

/* light.c:236,247 reference this symbol; on the device path shadows never run on the host. */
bool
is_shadowed(World w, Point light_position, Point pt)
{
    (void)w;
    (void)light_position;
    (void)pt;
    fprintf(stderr, "frt_shim: is_shadowed() called on the host -- the B200 core owns shadow rays\n");
    abort();
}

void
shade_hit(World w, Computations comps, size_t remaining, Color res)
{
    (void)w;
    (void)comps;
    (void)remaining;
    (void)res;
    fprintf(stderr, "frt_shim: shade_hit() is device-only\n");
    abort();
}

double
schlick(Computations comps)
{
    (void)comps;
    fprintf(stderr, "frt_shim: schlick() is device-only\n");
    abort();
}

void
prepare_computations(Intersection i, Ray r, Color photon_power, Intersections xs, Computations res, struct container *container)
{
    (void)i;
    (void)r;
    (void)photon_power;
    (void)xs;
    (void)res;
    (void)container;
    fprintf(stderr, "frt_shim: prepare_computations() is device-only\n");
    abort();
}
