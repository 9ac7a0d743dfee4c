/*
 * frt_objload.c -- the OBJ geometry parse of the reference, rebuilt (host side, C11, compiled
 * against the reference's own headers like frt_shim.c).
 *
 * Replaces construct_group_from_obj_file (reference src/libs/obj_loader/obj_loader.c:446-546) behind the same
 * signature: the generated main() calls it once per `obj` shape (yaml_parser/obj_parser.py:41).  The reference reads the
 * file with fgets + sscanf (obj_parse_line :334-441, parse_vertex :327-331, fan_triangulation :221-317), builds every face
 * in a scratch array of 32 shapes it allocates per line, deep-copies the triangles into a staging group
 * (group_add_children_stage, group.c:50-70) and deep-copies every group once more into the result (:520-526).
 * Here the file is read in one piece and walked once with a hand-written tokenizer and number parser, every triangle is
 * constructed IN PLACE in the children array of the group it ends up in -- a copy of a prototype the reference's own
 * triangle() / smooth_triangle() built, with the per-triangle fields filled in by the reference's own vector functions, so a
 * shape is field for field what the reference builds -- and the named groups are installed in the result without a copy.
 *
 * Same results, checked on the flattened tree (tests/test_objload.py: the scene blob of a program linked with this file
 * equals the blob of the program with the reference's loader, byte for byte):
 *   - "lines" are fgets(1024) chunks: a longer line continues as a new line, as there (:469);
 *   - numbers: %lf / %lu as sscanf reads them.  Decimal strings of up to 15 significant digits with a decimal exponent
 *     within +-22 are converted exactly (one correctly rounded operation on exactly representable operands -- what a
 *     correctly rounding strtod returns); everything else goes through strtod itself;
 *   - a component a line does not give stays 0 (the reference leaves its malloc'd arrays as they come: zero pages);
 *   - faces are fan-triangulated from the first vertex; whether the face carries normals / texture coordinates is decided
 *     by its FIRST vertex (:236-258); "g" switches or creates a group by name, "usemtl" binds the materials the
 *     reference's own parse_mtl (:139-212, still linked: MTL files are a few lines) put into its hash table.
 * Deviation, stated: an index outside the vertices read so far makes the reference read arbitrary memory; here it ends
 * the program with a message.
 *
 * Linking: `-Wl,--wrap=construct_group_from_obj_file` next to the reference's obj_loader.o (its parse_mtl and material
 * table are used), or call frt_construct_group_from_obj_file directly -- INTEGRATION.md.
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdbool.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

#include "src/libs/uthash/uthash.h"

#include "src/color/rgb.h"
#include "src/shapes/group.h"
#include "src/shapes/shapes.h"
#include "src/shapes/triangle.h"

/* the reference's material table (obj_loader.c:26-37): parse_mtl fills it, usemtl looks names up in it */
#define MAX_MATERIAL_NAME_LEN 256
typedef struct material_name_hash_table {
    char name[MAX_MATERIAL_NAME_LEN];
    int id;
    Material material;
    UT_hash_handle hh;
} *Named_material;
extern Named_material materials_ht;
void parse_mtl(FILE *mtl_file, void (*color_space_fn)(const Color, Color));

static void
die(const char *what, const char *detail)
{
    fprintf(stderr, "frt_objload: %s%s%s\n", what, detail != NULL ? ": " : "", detail != NULL ? detail : "");
    exit(3);
}

static void *
xrealloc(void *p, size_t bytes)
{
    void *q = realloc(p, bytes);
    if (q == NULL) {
        die("out of memory", NULL);
    }
    return q;
}

/* ------------------------------------------------------------------------------------------------ numbers */

static const double POW10[23] = { 1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                                  1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22 };

static inline bool
is_space(char c)
{
    return c == ' ' || c == '\t' || c == '\n' || c == '\v' || c == '\f' || c == '\r';
}

/* %lf: skips white space, converts the longest prefix that is a number; false when there is none (sscanf stops there) */
static bool
scan_double(const char **pp, const char *end, double *out)
{
    const char *p = *pp;
    while (p < end && is_space(*p)) {
        ++p;
    }
    const char *start = p;
    bool neg = false;
    if (p < end && (*p == '+' || *p == '-')) {
        neg = *p == '-';
        ++p;
    }
    uint64_t mant = 0;
    int digits = 0, sig = 0, exp10 = 0;
    bool exact = true;
    while (p < end && *p >= '0' && *p <= '9') {
        if (sig > 0 || *p != '0') {
            if (sig < 19) {
                mant = mant * 10 + (uint64_t)(*p - '0');
            } else {
                exact = false;
            }
            ++sig;
        }
        ++digits;
        ++p;
    }
    if (p < end && *p == '.') {
        ++p;
        while (p < end && *p >= '0' && *p <= '9') {
            if (sig > 0 || *p != '0') {
                if (sig < 19) {
                    mant = mant * 10 + (uint64_t)(*p - '0');
                } else {
                    exact = false;
                }
                ++sig;
            }
            --exp10;
            ++digits;
            ++p;
        }
    }
    if (digits == 0) {
        /* "inf", "nan", hexadecimal floats, or no number at all: strtod decides (the buffer ends with a NUL) */
        char *e = NULL;
        const double v = strtod(start, &e);
        if (e == start) {
            return false;
        }
        *out = v;
        *pp = e;
        return true;
    }
    if (p < end && (*p == 'e' || *p == 'E')) {
        const char *q = p + 1;
        bool eneg = false;
        if (q < end && (*q == '+' || *q == '-')) {
            eneg = *q == '-';
            ++q;
        }
        if (q < end && *q >= '0' && *q <= '9') {
            int e = 0;
            while (q < end && *q >= '0' && *q <= '9') {
                if (e < 100000) {
                    e = e * 10 + (*q - '0');
                }
                ++q;
            }
            exp10 += eneg ? -e : e;
            p = q;
        }
    }
    if (exact && sig <= 15 && exp10 >= -22 && exp10 <= 22) {
        /* mant < 10^15 < 2^53 and 10^|exp10| are exact doubles: one correctly rounded multiplication / division */
        double v = (double)mant;
        v = exp10 < 0 ? v / POW10[-exp10] : v * POW10[exp10];
        *out = neg ? -v : v;
    } else {
        char *e = NULL;
        *out = strtod(start, &e);
    }
    *pp = p;
    return true;
}

/* %lu on the digits at p (the tokens of a face carry no sign and no white space inside); false when there is no digit */
static bool
scan_index(const char **pp, const char *end, size_t *out)
{
    const char *p = *pp;
    if (p >= end || *p < '0' || *p > '9') {
        return false;
    }
    size_t v = 0;
    while (p < end && *p >= '0' && *p <= '9') {
        v = v * 10 + (size_t)(*p - '0');
        ++p;
    }
    *out = v;
    *pp = p;
    return true;
}

/* "%*s %255s": the second white-space separated word of the line */
static void
second_word(const char *line, const char *end, char out[MAX_MATERIAL_NAME_LEN])
{
    const char *p = line;
    out[0] = '\0';
    while (p < end && is_space(*p)) {
        ++p;
    }
    while (p < end && !is_space(*p)) {
        ++p;
    }
    while (p < end && is_space(*p)) {
        ++p;
    }
    size_t n = 0;
    while (p < end && !is_space(*p) && n < MAX_MATERIAL_NAME_LEN - 1) {
        out[n++] = *p++;
    }
    out[n] = '\0';
}

/* ------------------------------------------------------------------------------------------------ the file */

typedef struct {
    double *v; /* 4 doubles per entry, like the reference's arrays (:459-461) */
    size_t n, cap;
} vec4_array;

static void
vec4_push(vec4_array *a, const char *line, const char *end, double w)
{
    if (a->n == a->cap) {
        const size_t cap = a->cap ? 2 * a->cap : 2048;
        a->v = (double *)xrealloc(a->v, cap * 4 * sizeof(double));
        memset(a->v + 4 * a->cap, 0, (cap - a->cap) * 4 * sizeof(double));
        a->cap = cap;
    }
    double *dst = a->v + 4 * a->n;
    /* "%*s %lf %lf %lf" (parse_vertex :327-331) */
    const char *p = line;
    while (p < end && is_space(*p)) {
        ++p;
    }
    while (p < end && !is_space(*p)) {
        ++p;
    }
    for (int k = 0; k < 3; ++k) {
        if (!scan_double(&p, end, dst + k)) {
            break;
        }
    }
    dst[3] = w;
    a->n += 1;
}

typedef struct {
    char *name;
    Shape tris;
    size_t n, cap;
    size_t expected; /* triangles the counting pass found for this group */
} named_group;

typedef struct {
    named_group *g;
    size_t n, cap, cur;
} group_table;

/* "g name" (:392-417): the group of that name, made if there is none yet; it becomes the current one */
static void
select_group(group_table *t, const char *name)
{
    size_t i = 0;
    for (; i < t->n; ++i) {
        if (strcmp(name, t->g[i].name) == 0) {
            break;
        }
    }
    if (i == t->n) {
        if (t->n == t->cap) {
            t->cap *= 2;
            t->g = (named_group *)xrealloc(t->g, t->cap * sizeof(named_group));
        }
        memset(&t->g[t->n], 0, sizeof(named_group));
        t->g[t->n].name = strdup(name);
        ++t->n;
    }
    t->cur = i;
}

/* fgets(line, 1024): up to 1023 bytes, through the first newline */
static inline const char *
line_end(const char *line, const char *file_end)
{
    const char *nl = memchr(line, '\n', (size_t)(file_end - line));
    const char *end = nl != NULL ? nl + 1 : file_end;
    return end - line > 1023 ? line + 1023 : end;
}

typedef struct {
    size_t v, t, n;
} face_vertex;

/* one vertex of a face: "v", "v/t", "v/t/n" or "v//n" (:236-258 for the first, :268-289 for the others); how many of the
 * three fields sscanf would have converted (the first vertex decides what the face uses) */
static int
scan_face_vertex(const char *tok, const char *end, face_vertex *fv, bool *double_slash)
{
    fv->v = fv->t = fv->n = 0;
    *double_slash = false;
    const char *p = tok;
    if (!scan_index(&p, end, &fv->v)) {
        return 0;
    }
    if (p >= end || *p != '/') {
        return 1;
    }
    ++p;
    if (p < end && *p == '/') { /* "%lu//%lu" */
        *double_slash = true;
        ++p;
        scan_index(&p, end, &fv->n);
        return 1;
    }
    if (!scan_index(&p, end, &fv->t)) {
        return 1;
    }
    if (p >= end || *p != '/') {
        return 2;
    }
    ++p;
    if (!scan_index(&p, end, &fv->n)) {
        return 2;
    }
    return 3;
}

static double *
checked(const vec4_array *a, size_t index, const char *what)
{
    if (index == 0 || index > a->n) {
        die("a face refers to an entry the file has not given yet", what);
    }
    return a->v + 4 * (index - 1);
}

/*
 * triangle() / smooth_triangle() (triangle.c:69-104, :176-213) spend most of their 740 ns in shape_set_transform (matrix
 * copies and an inverse of the identity) -- 35 of the 44 ms this file needed for the 47 K triangles of dragon.obj.  Every
 * triangle of a file gets the same values there, so one prototype of each kind is built by the reference's constructor and
 * copied; what differs per triangle is filled in as the constructor does it, with the reference's own vector functions
 * (same object code, same rounding): vertices, edges, the flat normal or the three vertex normals, an intersection list
 * and a material of its own.
 */
static struct shape g_proto_flat, g_proto_smooth;
static bool g_protos_ready;

static void
make_prototypes(void)
{
    double o[4] = { 0, 0, 0, 1 }, a[4] = { 1, 0, 0, 1 }, b[4] = { 0, 1, 0, 1 }, n[4] = { 0, 0, 1, 0 };
    triangle(&g_proto_flat, o, a, b);
    smooth_triangle(&g_proto_smooth, o, a, b, n, n, n);
    struct shape *protos[2] = { &g_proto_flat, &g_proto_smooth };
    for (int k = 0; k < 2; ++k) {
        material_free(protos[k]->material);
        intersections_free(protos[k]->xs);
        protos[k]->material = NULL;
        protos[k]->xs = NULL;
    }
    g_protos_ready = true;
}

static inline void
construct_triangle(Shape s, bool smooth, double *p1, double *p2, double *p3, double *n1, double *n2, double *n3, Material bound)
{
    *s = smooth ? g_proto_smooth : g_proto_flat;
    s->xs = intersections_empty(1);
    memcpy(s->fields.triangle.p1, p1, sizeof(Point));
    memcpy(s->fields.triangle.p2, p2, sizeof(Point));
    memcpy(s->fields.triangle.p3, p3, sizeof(Point));
    vector_from_points(p2, p1, s->fields.triangle.e1);
    vector_from_points(p3, p1, s->fields.triangle.e2);
    if (smooth) {
        vector_copy(s->fields.triangle.u_normals.s_normals.n1, n1);
        vector_copy(s->fields.triangle.u_normals.s_normals.n2, n2);
        vector_copy(s->fields.triangle.u_normals.s_normals.n3, n3);
    } else {
        Vector cross, normal;
        vector_cross(s->fields.triangle.e2, s->fields.triangle.e1, cross);
        vector_normalize(cross, normal);
        memcpy(s->fields.triangle.u_normals.normal, normal, sizeof(Vector));
    }
    if (bound != NULL) {
        shape_set_material(s, bound); /* s->material is NULL: takes a reference, like shape_set_material after the constructor */
    } else {
        s->material = material_alloc();
    }
}

static double g_last_ms;

double
frt_objload_last_ms(void)
{
    return g_last_ms;
}

void
frt_construct_group_from_obj_file(const char *file_path, void (*color_space_fn)(const Color, Color), Shape result_group)
{
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    FILE *fp = fopen(file_path, "rb");
    if (fp == NULL) {
        printf("Error opening file %s", file_path); /* :453-456 */
        return;
    }
    fseek(fp, 0, SEEK_END);
    const long size = ftell(fp);
    fseek(fp, 0, SEEK_SET);
    char *text = (char *)xrealloc(NULL, (size_t)size + 1);
    if (size > 0 && fread(text, 1, (size_t)size, fp) != (size_t)size) {
        die("short read", file_path);
    }
    fclose(fp);
    text[size] = '\0';
    const char *file_end = text + size;

    vec4_array vs = { NULL, 0, 0 }, ts = { NULL, 0, 0 }, ns = { NULL, 0, 0 };
    group_table gt = { (named_group *)xrealloc(NULL, 16 * sizeof(named_group)), 0, 16, 0 };
    select_group(&gt, "##default_group"); /* :474-478 */
    Material current_material = NULL;
    if (!g_protos_ready) {
        make_prototypes();
    }

    /* Counting pass over the "f" and "g" lines: every group's triangle array is allocated once, at its final size.  (Grown
     * by doubling between the three small allocations every triangle makes, a 20 MB array was copied at every step: 29 of
     * the 45 ms of dragon.obj went into that.) */
    for (const char *line = text; line < file_end;) {
        const char *end = line_end(line, file_end);
        const size_t len = (size_t)(end - line);
        if (len >= 2 && line[0] == 'f' && line[1] == ' ') {
            size_t tokens = 0;
            for (const char *p = line + 2; p < end;) {
                while (p < end && (*p == ' ' || *p == '\t')) {
                    ++p;
                }
                if (p >= end || *p == '\n' || *p == '\r') {
                    break;
                }
                while (p < end && *p != ' ' && *p != '\t') {
                    ++p;
                }
                ++tokens;
            }
            gt.g[gt.cur].expected += tokens > 2 ? tokens - 2 : 0;
        } else if (len >= 2 && line[0] == 'g' && line[1] == ' ') {
            char name[MAX_MATERIAL_NAME_LEN];
            second_word(line, end, name);
            select_group(&gt, name);
        }
        line = end;
    }
    for (size_t i = 0; i < gt.n; ++i) {
        if (gt.g[i].expected > 0) {
            gt.g[i].tris = array_of_shapes(gt.g[i].expected);
            if (gt.g[i].tris == NULL) {
                die("out of memory", NULL);
            }
            gt.g[i].cap = gt.g[i].expected;
        }
    }
    gt.cur = 0;
#define groups gt.g
#define n_groups gt.n
#define cur gt.cur

    for (const char *line = text; line < file_end;) {
        const char *end = line_end(line, file_end);
        const size_t len = (size_t)(end - line);
        if (len >= 2 && line[0] == 'v' && line[1] == ' ') {
            vec4_push(&vs, line, end, 1.0);
        } else if (len >= 3 && line[0] == 'v' && line[1] == 't' && line[2] == ' ') {
            vec4_push(&ts, line, end, 0.0);
        } else if (len >= 3 && line[0] == 'v' && line[1] == 'n' && line[2] == ' ') {
            vec4_push(&ns, line, end, 0.0);
        } else if (len >= 2 && line[0] == 'f' && line[1] == ' ') {
            /* strtok(line + 2, " \t"): the vertices of the face; fan from the first */
            const char *p = line + 2;
            face_vertex first = { 0, 0, 0 }, prev = { 0, 0, 0 };
            bool use_normals = false, use_textures = false;
            int seen = 0;
            named_group *g = &groups[cur];
            while (p < end) {
                while (p < end && (*p == ' ' || *p == '\t')) {
                    ++p;
                }
                if (p >= end) {
                    break;
                }
                const char *tok = p;
                while (p < end && *p != ' ' && *p != '\t') {
                    ++p;
                }
                if (*tok == '\n' || *tok == '\r') {
                    break; /* :261: the token after a trailing blank is the line end itself */
                }
                face_vertex fv;
                bool dslash;
                const int got = scan_face_vertex(tok, p, &fv, &dslash);
                if (seen == 0) {
                    first = fv;
                    if (dslash) {
                        use_normals = true;
                    } else {
                        use_textures = got >= 2;
                        use_normals = got >= 3;
                    }
                } else if (seen >= 2) {
                    if (g->n == g->cap) {
                        g->cap = g->cap ? 2 * g->cap : 256;
                        g->tris = array_of_shapes_realloc(g->tris, g->cap);
                        if (g->tris == NULL) {
                            die("out of memory", NULL);
                        }
                    }
                    Shape s = g->tris + g->n;
                    double *p1 = checked(&vs, first.v, "v"), *p2 = checked(&vs, prev.v, "v"), *p3 = checked(&vs, fv.v, "v");
                    if (use_normals) {
                        construct_triangle(s, true, p1, p2, p3, checked(&ns, first.n, "vn"), checked(&ns, prev.n, "vn"), checked(&ns, fv.n, "vn"),
                                           current_material);
                    } else {
                        construct_triangle(s, false, p1, p2, p3, NULL, NULL, NULL, current_material);
                    }
                    if (use_textures) {
                        vector_copy(s->fields.triangle.t1, checked(&ts, first.t, "vt"));
                        vector_copy(s->fields.triangle.t2, checked(&ts, prev.t, "vt"));
                        vector_copy(s->fields.triangle.t3, checked(&ts, fv.t, "vt"));
                        s->fields.triangle.use_textures = true;
                    }
                    g->n += 1;
                }
                prev = fv;
                ++seen;
            }
        } else if (len >= 2 && line[0] == 'g' && line[1] == ' ') {
            char name[MAX_MATERIAL_NAME_LEN];
            second_word(line, end, name);
            select_group(&gt, name);
        } else if (len >= 6 && strncmp(line, "usemtl", 6) == 0) {
            char name[MAX_MATERIAL_NAME_LEN];
            second_word(line, end, name);
            Named_material s = NULL;
            HASH_FIND_STR(materials_ht, name, s);
            if (s != NULL) {
                current_material = s->material;
            } else {
                printf("Material %s not found.\n", name);
            }
        } else if (len >= 6 && strncmp(line, "mtllib", 6) == 0) {
            char name[MAX_MATERIAL_NAME_LEN];
            second_word(line, end, name);
            if (access(name, F_OK) < 0) {
                printf("file %s not found.\n", name);
            } else {
                FILE *mtl = fopen(name, "r");
                if (mtl == NULL) {
                    printf("Error opening file %s", name);
                } else {
                    parse_mtl(mtl, color_space_fn); /* the reference's own (:139-212) */
                    fclose(mtl);
                }
            }
        }
        line = end;
    }

    struct timespec t_parsed;
    clock_gettime(CLOCK_MONOTONIC, &t_parsed);
    /* the result: a group whose children are the non-empty named groups, in the order they were first named (:516-526) */
    group(result_group, NULL, 0);
    size_t nonempty = 0;
    for (size_t i = 0; i < n_groups; ++i) {
        nonempty += groups[i].n > 0 ? 1 : 0;
    }
    if (nonempty > 0) {
        free(result_group->fields.group.children);
        result_group->fields.group.children = array_of_shapes(nonempty > 16 ? nonempty : 16);
        result_group->fields.group.size_children_array = nonempty > 16 ? nonempty : 16;
        size_t k = 0;
        for (size_t i = 0; i < n_groups; ++i) {
            if (groups[i].n == 0) {
                continue;
            }
            Shape cg = result_group->fields.group.children + k++;
            group(cg, NULL, 0);
            free(cg->fields.group.children);
            cg->fields.group.children = groups[i].tris; /* installed, not copied */
            cg->fields.group.num_children = groups[i].n;
            cg->fields.group.size_children_array = groups[i].cap;
            cg->parent = result_group;
            groups[i].tris = NULL;
            group_add_children_finish(cg);
        }
        result_group->fields.group.num_children = nonempty;
        group_add_children_finish(result_group);
    }
    for (size_t i = 0; i < n_groups; ++i) {
        free(groups[i].name);
        free(groups[i].tris);
    }
    free(groups);
#undef groups
#undef n_groups
#undef cur
    free(vs.v);
    free(ts.v);
    free(ns.v);
    free(text);

    Named_material s, tmp; /* :540-545: the table goes, the materials stay with the shapes */
    HASH_ITER(hh, materials_ht, s, tmp) {
        HASH_DEL(materials_ht, s);
        free(s);
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    g_last_ms = 1e3 * (double)(t1.tv_sec - t0.tv_sec) + 1e-6 * (double)(t1.tv_nsec - t0.tv_nsec);
    if (getenv("FRT_B200_OBJ_TIMING") != NULL) {
        printf("FRT_B200_OBJ_MS %.3f (%s; %.3f ms until the last line was read)\n", g_last_ms, file_path,
               1e3 * (double)(t_parsed.tv_sec - t0.tv_sec) + 1e-6 * (double)(t_parsed.tv_nsec - t0.tv_nsec));
    }
}

/* -Wl,--wrap=construct_group_from_obj_file: every call of the reference's loader lands here.  FRT_OBJLOAD=ref hands the call
 * on to the reference's own function (A/B timing and the blob comparison of tests/test_objload.py). */
void __real_construct_group_from_obj_file(const char *file_path, void (*color_space_fn)(const Color, Color), Shape result_group);

void
__wrap_construct_group_from_obj_file(const char *file_path, void (*color_space_fn)(const Color, Color), Shape result_group)
{
    const char *mode = getenv("FRT_OBJLOAD");
    if (mode != NULL && strcmp(mode, "ref") == 0) {
        struct timespec t0, t1;
        clock_gettime(CLOCK_MONOTONIC, &t0);
        __real_construct_group_from_obj_file(file_path, color_space_fn, result_group);
        clock_gettime(CLOCK_MONOTONIC, &t1);
        g_last_ms = 1e3 * (double)(t1.tv_sec - t0.tv_sec) + 1e-6 * (double)(t1.tv_nsec - t0.tv_nsec);
        if (getenv("FRT_B200_OBJ_TIMING") != NULL) {
            printf("FRT_REF_OBJ_MS %.3f (%s)\n", g_last_ms, file_path);
        }
        return;
    }
    frt_construct_group_from_obj_file(file_path, color_space_fn, result_group);
}
