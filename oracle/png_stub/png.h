/*
 * png.h -- harness-side stand-in for libpng's header (TEST INFRASTRUCTURE).
 *
 * libpng headers are not installed in the build image, and the reference's
 * src/libs/canvas/canvas.c includes <png.h> unconditionally.  This stub only
 * lets that translation unit compile: both png_create_*_struct() return NULL,
 * which makes the reference's write_png()/read_png() bail out cleanly
 * (reference canvas.c:404-408 and :551-555).  PNG file I/O is outside the
 * render hot path.
 */
#ifndef FRT_ORACLE_PNG_STUB_H
#define FRT_ORACLE_PNG_STUB_H

#include <setjmp.h>
#include <stddef.h>
#include <stdio.h>

typedef struct frt_png_stub { jmp_buf jb; } *png_structp;
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

static inline png_structp png_create_write_struct(const char *v, void *a, void *b, void *c) { (void)v; (void)a; (void)b; (void)c; return NULL; }
static inline png_structp png_create_read_struct(const char *v, void *a, void *b, void *c) { (void)v; (void)a; (void)b; (void)c; return NULL; }
static inline png_infop png_create_info_struct(png_structp p) { (void)p; return NULL; }
static inline void png_init_io(png_structp p, FILE *f) { (void)p; (void)f; }
static inline void png_set_IHDR(png_structp p, png_infop i, unsigned w, unsigned h, int d, int ct, int il, int cm, int fm) { (void)p; (void)i; (void)w; (void)h; (void)d; (void)ct; (void)il; (void)cm; (void)fm; }
static inline void png_set_sRGB(png_structp p, png_infop i, int intent) { (void)p; (void)i; (void)intent; }
static inline void png_set_text(png_structp p, png_infop i, png_text *t, int n) { (void)p; (void)i; (void)t; (void)n; }
static inline void png_write_info(png_structp p, png_infop i) { (void)p; (void)i; }
static inline void png_write_row(png_structp p, png_bytep r) { (void)p; (void)r; }
static inline void png_write_end(png_structp p, png_infop i) { (void)p; (void)i; }
static inline void png_read_info(png_structp p, png_infop i) { (void)p; (void)i; }
static inline void png_read_update_info(png_structp p, png_infop i) { (void)p; (void)i; }
static inline void png_read_image(png_structp p, png_bytep *rows) { (void)p; (void)rows; }
static inline png_byte png_get_color_type(png_structp p, png_infop i) { (void)p; (void)i; return 0; }
static inline png_byte png_get_bit_depth(png_structp p, png_infop i) { (void)p; (void)i; return 8; }
static inline unsigned png_get_image_width(png_structp p, png_infop i) { (void)p; (void)i; return 0; }
static inline unsigned png_get_image_height(png_structp p, png_infop i) { (void)p; (void)i; return 0; }
static inline unsigned png_get_sRGB(png_structp p, png_infop i, int *intent) { (void)p; (void)i; (void)intent; return 0; }
static inline void png_set_palette_to_rgb(png_structp p) { (void)p; }
static inline void png_set_expand_gray_1_2_4_to_8(png_structp p) { (void)p; }
static inline void png_set_gray_to_rgb(png_structp p) { (void)p; }
static inline void png_set_strip_alpha(png_structp p) { (void)p; }
static inline void png_free_data(png_structp p, png_infop i, unsigned mask, int num) { (void)p; (void)i; (void)mask; (void)num; }
static inline void png_destroy_write_struct(png_structp *p, png_infop *i) { (void)p; (void)i; }
static inline void png_destroy_read_struct(png_structp *p, png_infop *i, png_infop *e) { (void)p; (void)i; (void)e; }

#endif
