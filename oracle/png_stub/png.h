/*
 * png.h -- harness-side stand-in for libpng's header (TEST INFRASTRUCTURE).
 *
 * libpng headers are not installed in the build image, and the reference's
 * src/libs/canvas/canvas.c includes <png.h> unconditionally.  This header lets
 * that translation unit compile and gives read_png() (reference canvas.c:532-671)
 * a working decoder, so that image-textured scenes can be rendered by the
 * unmodified reference for the golden fixtures:
 *
 *   reading   a small PNG decoder over zlib's inflate (non-interlaced, 8/16 bit,
 *             gray / gray+alpha / RGB / RGBA / palette), enough for every PNG the
 *             reference ships; the png_set_* transformations the reference asks
 *             for (palette->RGB, gray->RGB, strip alpha) are always applied, so
 *             png_read_image() delivers RGB rows
 *   writing   png_create_write_struct() returns NULL, which makes the reference's
 *             write_png() bail out cleanly (canvas.c:404-408); the harness reads
 *             the raw canvas instead (oracle/ref_hooks.c)
 *
 * Link with -lz.
 */
#ifndef FRT_ORACLE_PNG_STUB_H
#define FRT_ORACLE_PNG_STUB_H

#include <setjmp.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

typedef struct frt_png_stub {
    jmp_buf jb;
    FILE *fp;
    unsigned width, height;
    int bit_depth, color_type;
    int has_srgb;
    unsigned char *rgb; /* decoded image: width * height * 3 samples of bit_depth bits (16 bit: big endian) */
} *png_structp;
typedef struct frt_png_info_stub { int unused; } *png_infop;
typedef unsigned char png_byte;
typedef png_byte *png_bytep;
typedef struct { int compression; char *key; char *text; size_t text_length; } png_text;

#define PNG_LIBPNG_VER_STRING "stub"
#define PNG_COLOR_TYPE_GRAY 0
#define PNG_COLOR_TYPE_RGB 2
#define PNG_COLOR_TYPE_PALETTE 3
#define PNG_COLOR_TYPE_GRAY_ALPHA 4
#define PNG_COLOR_MASK_ALPHA 4
#define PNG_INTERLACE_NONE 0
#define PNG_COMPRESSION_TYPE_BASE 0
#define PNG_FILTER_TYPE_BASE 0
#define PNG_sRGB_INTENT_ABSOLUTE 3
#define PNG_TEXT_COMPRESSION_NONE -1
#define PNG_FREE_ALL 0xffffU
#define PNG_INFO_sRGB 0x0800U
#define png_jmpbuf(p) ((p)->jb)

/* ---- writing: not available */
static inline png_structp png_create_write_struct(const char *v, void *a, void *b, void *c) { (void)v; (void)a; (void)b; (void)c; return NULL; }
static inline void png_set_IHDR(png_structp p, png_infop i, unsigned w, unsigned h, int d, int ct, int il, int cm, int fm) { (void)p; (void)i; (void)w; (void)h; (void)d; (void)ct; (void)il; (void)cm; (void)fm; }
static inline void png_set_sRGB(png_structp p, png_infop i, int intent) { (void)p; (void)i; (void)intent; }
static inline void png_set_text(png_structp p, png_infop i, png_text *t, int n) { (void)p; (void)i; (void)t; (void)n; }
static inline void png_write_info(png_structp p, png_infop i) { (void)p; (void)i; }
static inline void png_write_row(png_structp p, png_bytep r) { (void)p; (void)r; }
static inline void png_write_end(png_structp p, png_infop i) { (void)p; (void)i; }
static inline void png_destroy_write_struct(png_structp *p, png_infop *i) { (void)p; (void)i; }

/* ---- reading */
static inline png_structp
png_create_read_struct(const char *v, void *a, void *b, void *c)
{
    (void)v; (void)a; (void)b; (void)c;
    return (png_structp)calloc(1, sizeof(struct frt_png_stub));
}

static inline png_infop
png_create_info_struct(png_structp p)
{
    (void)p;
    return (png_infop)calloc(1, sizeof(struct frt_png_info_stub));
}

static inline void png_init_io(png_structp p, FILE *f) { p->fp = f; }

static inline uint32_t frt_png_be32(const unsigned char *b) { return ((uint32_t)b[0] << 24) | ((uint32_t)b[1] << 16) | ((uint32_t)b[2] << 8) | b[3]; }

static inline int
frt_png_paeth(int a, int b, int c)
{
    int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

/* parse the whole file: IHDR, PLTE, sRGB, IDAT* -> inflate -> unfilter -> RGB */
static inline void
png_read_info(png_structp p, png_infop info)
{
    (void)info;
    unsigned char sig[8], hdr[8], *idat = NULL, palette[768];
    size_t idat_len = 0, idat_cap = 0;
    int have_ihdr = 0;
    memset(palette, 0, sizeof(palette));
    if (fread(sig, 1, 8, p->fp) != 8 || memcmp(sig, "\x89PNG\r\n\x1a\n", 8) != 0) longjmp(p->jb, 1);
    for (;;) {
        if (fread(hdr, 1, 8, p->fp) != 8) break;
        uint32_t len = frt_png_be32(hdr);
        unsigned char *data = (unsigned char *)malloc(len ? len : 1);
        if (data == NULL || fread(data, 1, len, p->fp) != len) longjmp(p->jb, 1);
        unsigned char crc[4];
        if (fread(crc, 1, 4, p->fp) != 4) longjmp(p->jb, 1);
        if (memcmp(hdr + 4, "IHDR", 4) == 0 && len >= 13) {
            p->width = frt_png_be32(data);
            p->height = frt_png_be32(data + 4);
            p->bit_depth = data[8];
            p->color_type = data[9];
            if (data[12] != 0) longjmp(p->jb, 1); /* interlaced: not needed by any shipped PNG */
            have_ihdr = 1;
        } else if (memcmp(hdr + 4, "PLTE", 4) == 0) {
            memcpy(palette, data, len < 768 ? len : 768);
        } else if (memcmp(hdr + 4, "sRGB", 4) == 0) {
            p->has_srgb = 1;
        } else if (memcmp(hdr + 4, "IDAT", 4) == 0) {
            if (idat_len + len > idat_cap) {
                idat_cap = (idat_len + len) * 2;
                idat = (unsigned char *)realloc(idat, idat_cap);
                if (idat == NULL) longjmp(p->jb, 1);
            }
            memcpy(idat + idat_len, data, len);
            idat_len += len;
        } else if (memcmp(hdr + 4, "IEND", 4) == 0) {
            free(data);
            break;
        }
        free(data);
    }
    if (!have_ihdr || idat == NULL) longjmp(p->jb, 1);
    int channels = p->color_type == 0 ? 1 : p->color_type == 2 ? 3 : p->color_type == 3 ? 1 : p->color_type == 4 ? 2 : 4;
    size_t bpp_bits = (size_t)channels * p->bit_depth;
    size_t stride = (p->width * bpp_bits + 7) / 8, bpp = (bpp_bits + 7) / 8;
    uLongf raw_len = (uLongf)((stride + 1) * p->height);
    unsigned char *raw = (unsigned char *)malloc(raw_len ? raw_len : 1);
    if (raw == NULL || uncompress(raw, &raw_len, idat, (uLong)idat_len) != Z_OK) longjmp(p->jb, 1);
    free(idat);
    /* unfilter in place */
    for (unsigned y = 0; y < p->height; ++y) {
        unsigned char *row = raw + (stride + 1) * y, *cur = row + 1;
        const unsigned char *prev = y ? row - stride : NULL;
        int ft = row[0];
        for (size_t x = 0; x < stride; ++x) {
            int a = x >= bpp ? cur[x - bpp] : 0, b = prev ? prev[x] : 0, c = (prev && x >= bpp) ? prev[x - bpp] : 0;
            int v = cur[x];
            switch (ft) {
            case 1: v += a; break;
            case 2: v += b; break;
            case 3: v += (a + b) / 2; break;
            case 4: v += frt_png_paeth(a, b, c); break;
            default: break;
            }
            cur[x] = (unsigned char)v;
        }
    }
    /* to RGB, keeping the bit depth (8 or 16); sub-byte gray / palette indices are expanded to 8 bit */
    int out_depth = p->bit_depth == 16 ? 16 : 8;
    size_t sample = out_depth / 8;
    p->rgb = (unsigned char *)malloc((size_t)p->width * p->height * 3 * sample);
    if (p->rgb == NULL) longjmp(p->jb, 1);
    for (unsigned y = 0; y < p->height; ++y) {
        const unsigned char *cur = raw + (stride + 1) * y + 1;
        for (unsigned x = 0; x < p->width; ++x) {
            unsigned char *o = p->rgb + ((size_t)y * p->width + x) * 3 * sample;
            if (p->bit_depth < 8) {
                unsigned bit = x * p->bit_depth;
                unsigned v = (cur[bit / 8] >> (8 - p->bit_depth - bit % 8)) & ((1u << p->bit_depth) - 1u);
                if (p->color_type == 3) {
                    o[0] = palette[3 * v]; o[1] = palette[3 * v + 1]; o[2] = palette[3 * v + 2];
                } else {
                    o[0] = o[1] = o[2] = (unsigned char)(v * 255u / ((1u << p->bit_depth) - 1u));
                }
            } else {
                const unsigned char *s = cur + (size_t)x * bpp;
                if (p->color_type == 3) {
                    o[0] = palette[3 * s[0]]; o[1] = palette[3 * s[0] + 1]; o[2] = palette[3 * s[0] + 2];
                } else if (p->color_type == 0 || p->color_type == 4) {
                    for (int k = 0; k < 3; ++k) memcpy(o + k * sample, s, sample);
                } else {
                    memcpy(o, s, 3 * sample);
                }
            }
        }
    }
    free(raw);
    if (p->bit_depth < 8) p->bit_depth = 8;
}

static inline void png_read_update_info(png_structp p, png_infop i) { (void)p; (void)i; }

static inline void
png_read_image(png_structp p, png_bytep *rows)
{
    size_t row_bytes = (size_t)p->width * 3 * (p->bit_depth == 16 ? 2 : 1);
    for (unsigned y = 0; y < p->height; ++y) {
        memcpy(rows[y], p->rgb + row_bytes * y, row_bytes);
    }
}

static inline png_byte png_get_color_type(png_structp p, png_infop i) { (void)i; return (png_byte)p->color_type; }
static inline png_byte png_get_bit_depth(png_structp p, png_infop i) { (void)i; return (png_byte)p->bit_depth; }
static inline unsigned png_get_image_width(png_structp p, png_infop i) { (void)i; return p->width; }
static inline unsigned png_get_image_height(png_structp p, png_infop i) { (void)i; return p->height; }
static inline unsigned png_get_sRGB(png_structp p, png_infop i, int *intent) { (void)i; *intent = 0; return p->has_srgb ? PNG_INFO_sRGB : 0; }
static inline void png_set_palette_to_rgb(png_structp p) { (void)p; }
static inline void png_set_expand_gray_1_2_4_to_8(png_structp p) { (void)p; }
static inline void png_set_gray_to_rgb(png_structp p) { (void)p; }
static inline void png_set_strip_alpha(png_structp p) { (void)p; }
static inline void png_free_data(png_structp p, png_infop i, unsigned mask, int num) { (void)p; (void)i; (void)mask; (void)num; }

static inline void
png_destroy_read_struct(png_structp *p, png_infop *i, png_infop *e)
{
    (void)e;
    if (p != NULL && *p != NULL) {
        free((*p)->rgb);
        free(*p);
        *p = NULL;
    }
    if (i != NULL && *i != NULL) {
        free(*i);
        *i = NULL;
    }
}

#endif
