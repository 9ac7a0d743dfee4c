/*
 * objload_check.c -- TEST INFRASTRUCTURE (oracle/): loads one OBJ file with the reference's own
 * construct_group_from_obj_file (src/libs/obj_loader/obj_loader.c:446) and with the rebuilt parse
 * (fast_ray_tracer_b200/csrc/frt_objload.c), compares the two shape trees field by field and prints the time of each.
 * Built by oracle/build_ref.py into oracle/_ref/objload_check from the reference's objects (linked with
 * --wrap=construct_group_from_obj_file so that both entry points exist side by side); used by tests/test_objload.py only.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "src/shapes/shapes.h" /* first: the reference's headers only resolve in this order */
#include "src/shapes/group.h"
#include "src/color/rgb.h"

void __real_construct_group_from_obj_file(const char *file_path, void (*color_space_fn)(const Color, Color), Shape result_group);
void frt_construct_group_from_obj_file(const char *file_path, void (*color_space_fn)(const Color, Color), Shape result_group);

static double
now_ms(void)
{
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return 1e3 * (double)t.tv_sec + 1e-6 * (double)t.tv_nsec;
}

static long n_triangles, n_groups, n_diff;

static void
differ(const char *what, long index)
{
    if (n_diff < 10) {
        printf("DIFF %s at shape %ld\n", what, index);
    }
    ++n_diff;
}

#define SAME(field) (memcmp(&(a->field), &(b->field), sizeof(a->field)) == 0)

static void
compare(Shape a, Shape b)
{
    const long index = n_triangles + n_groups;
    if (a->type != b->type) {
        differ("type", index);
        return;
    }
    if (!SAME(transform) || !SAME(transform_inverse) || a->transform_identity != b->transform_identity) {
        differ("transform", index);
    }
    if ((a->material == NULL) != (b->material == NULL)) {
        differ("material presence", index);
    } else if (a->material != NULL) {
        const Material x = a->material, y = b->material;
        const size_t rgb = 3 * sizeof(double); /* the fourth lane of a Color is never written */
        if (memcmp(x->Ka, y->Ka, rgb) || memcmp(x->Kd, y->Kd, rgb) || memcmp(x->Ks, y->Ks, rgb) || memcmp(x->Tf, y->Tf, rgb) || x->Ns != y->Ns || x->Ni != y->Ni || x->Tr != y->Tr ||
            x->casts_shadow != y->casts_shadow || x->reflective != y->reflective || (x->map_Kd == NULL) != (y->map_Kd == NULL) ||
            (x->map_bump == NULL) != (y->map_bump == NULL)) {
            differ("material", index);
        }
    }
    if (a->type == SHAPE_GROUP) {
        ++n_groups;
        if (a->fields.group.num_children != b->fields.group.num_children) {
            differ("num_children", index);
            return;
        }
        for (size_t i = 0; i < a->fields.group.num_children; ++i) {
            if ((a->fields.group.children + i)->parent != a || (b->fields.group.children + i)->parent != b) {
                differ("parent", index);
            }
            compare(a->fields.group.children + i, b->fields.group.children + i);
        }
    } else if (a->type == SHAPE_TRIANGLE || a->type == SHAPE_SMOOTH_TRIANGLE) {
        ++n_triangles;
        if (!SAME(fields.triangle.p1) || !SAME(fields.triangle.p2) || !SAME(fields.triangle.p3) || !SAME(fields.triangle.e1) ||
            !SAME(fields.triangle.e2)) {
            differ("vertices / edges", index);
        }
        if (a->fields.triangle.use_textures != b->fields.triangle.use_textures) {
            differ("use_textures", index);
        } else if (a->fields.triangle.use_textures &&
                   (!SAME(fields.triangle.t1) || !SAME(fields.triangle.t2) || !SAME(fields.triangle.t3))) {
            differ("texture coordinates", index);
        }
        if (a->type == SHAPE_TRIANGLE ? !SAME(fields.triangle.u_normals.normal) : !SAME(fields.triangle.u_normals.s_normals)) {
            differ("normals", index);
        }
        if (a->local_intersect != b->local_intersect || a->local_normal_at != b->local_normal_at || a->bounds != b->bounds) {
            differ("methods", index);
        }
    } else {
        differ("unexpected shape type", index);
    }
}

int
main(int argc, char **argv)
{
    if (argc < 2) {
        fprintf(stderr, "usage: objload_check file.obj [srgb]\n");
        return 2;
    }
    void (*fn)(const Color, Color) = (argc > 2 && strcmp(argv[2], "srgb") == 0) ? rgb_to_rgb : rgb_to_rgb; /* the colour function only reaches parse_mtl, which both loaders share */
    Shape a = array_of_shapes(1), b = array_of_shapes(1);
    double t0 = now_ms();
    __real_construct_group_from_obj_file(argv[1], fn, a);
    double t1 = now_ms();
    frt_construct_group_from_obj_file(argv[1], fn, b);
    double t2 = now_ms();
    compare(a, b);
    printf("OBJLOAD groups %ld triangles %ld differences %ld reference_ms %.3f rebuilt_ms %.3f\n", n_groups, n_triangles, n_diff, t1 - t0,
           t2 - t1);
    return n_diff == 0 ? 0 : 1;
}
