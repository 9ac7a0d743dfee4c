/*
 * frt_blob.c -- flat little-endian (de)serialisation of frt_scene_desc.
 *
 * A blob is how a flattened reference scene (World + Camera, see frt_shim.c) travels to a process that
 * has no access to the reference's host data model: the Python parity tests and bench.py load blobs
 * through the C ABI and render them on the GPU.  Layout: frt_blob_header, then every array of the
 * description in declaration order, each padded to 8 bytes.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "frt_b200.h"
#include "frt_internal.h"

#define FRT_BLOB_MAGIC 0x4e43535f54524600ULL /* "\0FRT_SCN" */
#define FRT_BLOB_VERSION 3 /* layout of the scene description sections; independent of the call ABI */

typedef struct frt_blob_header {
    uint64_t magic;
    int32_t abi_version;
    int32_t n_nodes, n_roots, n_xforms, n_materials, n_patterns, n_textures, n_lights;
    int64_t n_prim_params, n_texels, n_light_points, n_pixel_samples;
    frt_camera camera;
    frt_config config;
} frt_blob_header;

static size_t
pad8(size_t n)
{
    return (n + 7u) & ~(size_t)7u;
}

struct section {
    const void *ptr;
    size_t bytes;
};

static void
sections_of(const frt_scene_desc *d, struct section s[11])
{
    s[0] = (struct section){ d->nodes, (size_t)d->n_nodes * sizeof(frt_node) };
    s[1] = (struct section){ d->roots, (size_t)d->n_roots * sizeof(int32_t) };
    s[2] = (struct section){ d->xforms, (size_t)d->n_xforms * sizeof(frt_xform) };
    s[3] = (struct section){ d->prim_params, (size_t)d->n_prim_params * sizeof(double) };
    s[4] = (struct section){ d->materials, (size_t)d->n_materials * sizeof(frt_material) };
    s[5] = (struct section){ d->patterns, (size_t)d->n_patterns * sizeof(frt_pattern) };
    s[6] = (struct section){ d->textures, (size_t)d->n_textures * sizeof(frt_texture) };
    s[7] = (struct section){ d->texels, (size_t)d->n_texels * 3 * sizeof(double) };
    s[8] = (struct section){ d->lights, (size_t)d->n_lights * sizeof(frt_light) };
    s[9] = (struct section){ d->light_points, (size_t)d->n_light_points * 3 * sizeof(double) };
    s[10] = (struct section){ d->pixel_samples, d->pixel_samples ? (size_t)d->n_pixel_samples * sizeof(double) : 0 };
}

int
frt_scene_save(const frt_scene_desc *d, const char *path)
{
    if (d == NULL || path == NULL) {
        return frt_set_error(FRT_ERR_ARG, "frt_scene_save: null argument");
    }
    FILE *f = fopen(path, "wb");
    if (f == NULL) {
        return frt_set_error(FRT_ERR_IO, "frt_scene_save: cannot open %s", path);
    }
    frt_blob_header h;
    memset(&h, 0, sizeof(h));
    h.magic = FRT_BLOB_MAGIC;
    h.abi_version = FRT_BLOB_VERSION;
    h.n_nodes = d->n_nodes;
    h.n_roots = d->n_roots;
    h.n_xforms = d->n_xforms;
    h.n_materials = d->n_materials;
    h.n_patterns = d->n_patterns;
    h.n_textures = d->n_textures;
    h.n_lights = d->n_lights;
    h.n_prim_params = d->n_prim_params;
    h.n_texels = d->n_texels;
    h.n_light_points = d->n_light_points;
    h.n_pixel_samples = d->pixel_samples ? d->n_pixel_samples : 0;
    h.camera = d->camera;
    h.config = d->config;
    static const char zeros[8] = { 0 };
    int ok = fwrite(&h, sizeof(h), 1, f) == 1;
    ok = ok && fwrite(zeros, 1, pad8(sizeof(h)) - sizeof(h), f) == pad8(sizeof(h)) - sizeof(h);
    struct section s[11];
    sections_of(d, s);
    for (int i = 0; ok && i < 11; ++i) {
        if (s[i].bytes) {
            ok = fwrite(s[i].ptr, 1, s[i].bytes, f) == s[i].bytes;
            ok = ok && fwrite(zeros, 1, pad8(s[i].bytes) - s[i].bytes, f) == pad8(s[i].bytes) - s[i].bytes;
        }
    }
    if (fclose(f) != 0 || !ok) {
        return frt_set_error(FRT_ERR_IO, "frt_scene_save: short write to %s", path);
    }
    return FRT_OK;
}

int
frt_scene_load(const char *path, frt_scene_desc **out)
{
    if (path == NULL || out == NULL) {
        return frt_set_error(FRT_ERR_ARG, "frt_scene_load: null argument");
    }
    *out = NULL;
    FILE *f = fopen(path, "rb");
    if (f == NULL) {
        return frt_set_error(FRT_ERR_IO, "frt_scene_load: cannot open %s", path);
    }
    fseek(f, 0, SEEK_END);
    long size = ftell(f);
    fseek(f, 0, SEEK_SET);
    if (size < (long)sizeof(frt_blob_header)) {
        fclose(f);
        return frt_set_error(FRT_ERR_IO, "frt_scene_load: %s is too short", path);
    }
    /* one allocation: [frt_scene_desc][file bytes]; the desc points into the file image */
    size_t head = pad8(sizeof(frt_scene_desc));
    char *block = (char *)malloc(head + (size_t)size);
    if (block == NULL) {
        fclose(f);
        return frt_set_error(FRT_ERR_IO, "frt_scene_load: out of memory");
    }
    if (fread(block + head, 1, (size_t)size, f) != (size_t)size) {
        fclose(f);
        free(block);
        return frt_set_error(FRT_ERR_IO, "frt_scene_load: short read from %s", path);
    }
    fclose(f);

    const frt_blob_header *h = (const frt_blob_header *)(block + head);
    if (h->magic != FRT_BLOB_MAGIC || h->abi_version != FRT_BLOB_VERSION) {
        free(block);
        return frt_set_error(FRT_ERR_IO, "frt_scene_load: %s is not a version-%d scene blob", path, FRT_BLOB_VERSION);
    }
    frt_scene_desc *d = (frt_scene_desc *)block;
    memset(d, 0, sizeof(*d));
    d->abi_version = FRT_ABI_VERSION;
    d->n_nodes = h->n_nodes;
    d->n_roots = h->n_roots;
    d->n_xforms = h->n_xforms;
    d->n_materials = h->n_materials;
    d->n_patterns = h->n_patterns;
    d->n_textures = h->n_textures;
    d->n_lights = h->n_lights;
    d->n_prim_params = h->n_prim_params;
    d->n_texels = h->n_texels;
    d->n_light_points = h->n_light_points;
    d->n_pixel_samples = h->n_pixel_samples;
    d->camera = h->camera;
    d->config = h->config;

    /* counts come from the file: negative or absurd ones must not wrap the offsets below */
    const long long lim = (long long)size;
    if (d->n_nodes < 0 || d->n_roots < 0 || d->n_xforms < 0 || d->n_materials < 0 || d->n_patterns < 0 || d->n_textures < 0 ||
        d->n_lights < 0 || d->n_prim_params < 0 || d->n_texels < 0 || d->n_light_points < 0 || d->n_pixel_samples < 0 ||
        d->n_nodes > lim || d->n_roots > lim || d->n_xforms > lim || d->n_materials > lim || d->n_patterns > lim || d->n_textures > lim ||
        d->n_lights > lim || d->n_prim_params > lim || d->n_texels > lim || d->n_light_points > lim || d->n_pixel_samples > lim) {
        free(block);
        return frt_set_error(FRT_ERR_IO, "frt_scene_load: %s has a negative or impossible section count", path);
    }
    struct section s[11];
    sections_of(d, s); /* sizes only; pointers are filled below */
    size_t off = head + pad8(sizeof(frt_blob_header));
    const void *ptrs[11];
    s[10].bytes = (size_t)d->n_pixel_samples * sizeof(double);
    for (int i = 0; i < 11; ++i) {
        ptrs[i] = s[i].bytes ? block + off : NULL;
        if (s[i].bytes > (size_t)size || off > head + (size_t)size) { /* every count is <= size, so no product above wraps */
            off = (size_t)-1;
            break;
        }
        off += pad8(s[i].bytes);
    }
    if (off > head + (size_t)size) {
        free(block);
        return frt_set_error(FRT_ERR_IO, "frt_scene_load: %s is truncated", path);
    }
    d->nodes = (const frt_node *)ptrs[0];
    d->roots = (const int32_t *)ptrs[1];
    d->xforms = (const frt_xform *)ptrs[2];
    d->prim_params = (const double *)ptrs[3];
    d->materials = (const frt_material *)ptrs[4];
    d->patterns = (const frt_pattern *)ptrs[5];
    d->textures = (const frt_texture *)ptrs[6];
    d->texels = (const double *)ptrs[7];
    d->lights = (const frt_light *)ptrs[8];
    d->light_points = (const double *)ptrs[9];
    d->pixel_samples = (const double *)ptrs[10];
    *out = d;
    return FRT_OK;
}

void
frt_scene_desc_free(frt_scene_desc *desc)
{
    free(desc); /* single block, see frt_scene_load */
}
