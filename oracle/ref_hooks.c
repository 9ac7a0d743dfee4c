/*
 * ref_hooks.c -- link-time hooks around the reference program (TEST INFRASTRUCTURE).
 *
 * Linked into every binary under oracle/_ref/ with
 *   -Wl,--wrap=render_multi -Wl,--wrap=intersect_world
 *   -Wl,--wrap=write_ppm_file -Wl,--wrap=write_png -Wl,--wrap=trace_photons
 * so that the UNMODIFIED reference sources (compiled where they lie under
 * /root/reference) can be timed, ray-counted and dumped without source edits.
 * The same hooks wrap the B200 shim build of a scene, so both arms see the
 * same camera / thread overrides.
 *
 * Environment knobs (all optional):
 *   FRT_REF_THREADS=N      overwrite global_config.threading.num_threads
 *   FRT_REF_HSIZE/VSIZE    re-derive the camera for another resolution, same
 *                          field of view (formulae: reference camera.c:103-138)
 *   FRT_REF_USTEPS/VSTEPS  overwrite the per-pixel sample grid
 *   FRT_COUNT_RAYS=1       count intersect_world() calls (reference world.c:164)
 *   FRT_CANVAS_OUT=path    dump Canvas->arr as raw doubles instead of the PPM
 *                          (int64 width, int64 height, then w*h*3 float64 RGB)
 *   FRT_SKIP_PPM=1         do not run the reference's own write_ppm_file
 *   FRT_DEVICE_PPM=0       B200 drop-in build: run the reference's host write_ppm_file instead of the device encoder
 *   FRT_REF_SEED=n         srand48(n) / srand(n) before trace_photons() (the reference has no seed control of its own;
 *                          two seeds give two independent reference renders of a photon-mapped scene)
 * Output lines (stdout): "FRT_RENDER_SECONDS <s>", "FRT_RAYS <n>", "FRT_THREADS <n>".
 */
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <string.h>
#include <time.h>
#include <math.h>

#include "src/renderer/camera.h"
#include "src/renderer/world.h"
#include "src/renderer/renderer.h"
#include "src/renderer/photon_tracer.h"
#include "src/libs/canvas/canvas.h"

Canvas __real_render_multi(Camera cam, World w, size_t usteps, size_t vsteps, bool jitter);
Intersections __real_intersect_world(const World w, const Ray r, bool stop_after_first_hit);
int __real_write_ppm_file(Canvas c, const bool use_scaling, const char *file_name);
/* present in the B200 drop-in build only (fast_ray_tracer_b200/csrc/frt_shim.c): the PPM encoded on the device */
int frt_shim_write_ppm_file(Canvas c, const bool use_scaling, const char *file_name) __attribute__((weak));
void __real_trace_photons(const World w, const size_t num_maps, bool populate_caustic, bool populate_global);

/* intersect_world() calls, counted per thread in cache-line-padded slots so that counting does not serialise the
 * reference's worker threads on one shared atomic */
#define FRT_COUNTER_SLOTS 256
static struct { unsigned long long n; char pad[56]; } frt_ray_slots[FRT_COUNTER_SLOTS] __attribute__((aligned(64)));
static unsigned int frt_next_slot;
static __thread int frt_my_slot = -1;
static int frt_count_rays;

static unsigned long long
frt_ray_total(void)
{
    unsigned long long t = 0;
    for (int i = 0; i < FRT_COUNTER_SLOTS; ++i) {
        t += __atomic_load_n(&frt_ray_slots[i].n, __ATOMIC_RELAXED);
    }
    return t;
}

static long
env_long(const char *name, long dflt)
{
    const char *s = getenv(name);
    if (s == NULL || *s == '\0') {
        return dflt;
    }
    return strtol(s, NULL, 10);
}

static double
now_seconds(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

Intersections
__wrap_intersect_world(const World w, const Ray r, bool stop_after_first_hit)
{
    if (frt_count_rays) {
        if (frt_my_slot < 0) {
            frt_my_slot = (int)(__atomic_fetch_add(&frt_next_slot, 1u, __ATOMIC_RELAXED) % FRT_COUNTER_SLOTS);
        }
        __atomic_fetch_add(&frt_ray_slots[frt_my_slot].n, 1ULL, __ATOMIC_RELAXED);
    }
    return __real_intersect_world(w, r, stop_after_first_hit);
}

void
__wrap_trace_photons(const World w, const size_t num_maps, bool populate_caustic, bool populate_global)
{
    long seed = env_long("FRT_REF_SEED", -1);
    if (seed >= 0) {
        srand48(seed);
        srand((unsigned int)seed);
    }
    double t0 = now_seconds();
    __real_trace_photons(w, num_maps, populate_caustic, populate_global);
    printf("\nFRT_PHOTON_SECONDS %.6f\n", now_seconds() - t0); /* main() printed "Tracing photons..." without a newline */
    fflush(stdout);
}

Canvas
__wrap_render_multi(Camera cam, World w, size_t usteps, size_t vsteps, bool jitter)
{
    long threads = env_long("FRT_REF_THREADS", 0);
    long hsize = env_long("FRT_REF_HSIZE", 0);
    long vsize = env_long("FRT_REF_VSIZE", 0);
    long us = env_long("FRT_REF_USTEPS", 0);
    long vs = env_long("FRT_REF_VSTEPS", 0);
    frt_count_rays = (int)env_long("FRT_COUNT_RAYS", 0);

    if (threads > 0) {
        w->global_config->threading.num_threads = (size_t)threads;
    }
    if (hsize > 0 && vsize > 0) {
        /* same arithmetic as the reference's camera() constructor */
        double half_view = cam->canvas_distance * tan(cam->field_of_view * 0.5);
        double aspect = (double)hsize / (double)vsize;
        cam->hsize = (size_t)hsize;
        cam->vsize = (size_t)vsize;
        if (aspect >= 1.0) {
            cam->half_width = half_view;
            cam->half_height = half_view / aspect;
        } else {
            cam->half_width = half_view * aspect;
            cam->half_height = half_view;
        }
        cam->pixel_size = cam->half_width * 2.0 / (double)hsize;
    }
    if (us > 0 && vs > 0) {
        usteps = cam->usteps = (size_t)us;
        vsteps = cam->vsteps = (size_t)vs;
    }

    memset(frt_ray_slots, 0, sizeof(frt_ray_slots));
    if (env_long("FRT_REF_SEED", -1) >= 0) { /* per-pixel CMJ jitter, aperture sampling and sample-set picks draw from these */
        srand48(env_long("FRT_REF_SEED", 0) + 17);
        srand((unsigned int)env_long("FRT_REF_SEED", 0) + 17u);
    }
    double t0 = now_seconds();
    Canvas c = __real_render_multi(cam, w, usteps, vsteps, jitter);
    double t1 = now_seconds();

    printf("FRT_THREADS %lu\n", (unsigned long)w->global_config->threading.num_threads);
    printf("FRT_SIZE %lu %lu %lu %lu\n", (unsigned long)cam->hsize, (unsigned long)cam->vsize,
           (unsigned long)usteps, (unsigned long)vsteps);
    printf("FRT_RENDER_SECONDS %.6f\n", t1 - t0);
    if (frt_count_rays) {
        printf("FRT_RAYS %llu\n", frt_ray_total());
    }
    fflush(stdout);
    return c;
}

int
__wrap_write_ppm_file(Canvas c, const bool use_scaling, const char *file_name)
{
    const char *out = getenv("FRT_CANVAS_OUT");
    if (out != NULL && *out != '\0') {
        FILE *f = fopen(out, "wb");
        if (f == NULL) {
            fprintf(stderr, "ref_hooks: cannot open %s\n", out);
            exit(3);
        }
        int64_t wh[2] = { (int64_t)c->width, (int64_t)c->height };
        fwrite(wh, sizeof(int64_t), 2, f);
        size_t i, n = c->width * c->height;
        for (i = 0; i < n; ++i) {
            fwrite(c->arr[i], sizeof(double), 3, f);
        }
        fclose(f);
    }
    if (env_long("FRT_SKIP_PPM", 0)) {
        return 0;
    }
    double t0 = now_seconds();
    int rc = -1;
    if (frt_shim_write_ppm_file != NULL && env_long("FRT_DEVICE_PPM", 1)) {
        rc = frt_shim_write_ppm_file(c, use_scaling, file_name);
    }
    if (rc != 0) {
        rc = __real_write_ppm_file(c, use_scaling, file_name);
    }
    printf("FRT_PPM_SECONDS %.6f\n", now_seconds() - t0); /* encode + file write, either arm */
    fflush(stdout);
    return rc;
}

int
__wrap_write_png(Canvas c, const char *file_name)
{
    (void)c;
    (void)file_name;
    return 0; /* PNG output needs libpng, which the image does not have */
}
