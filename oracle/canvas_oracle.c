/*
 * canvas_oracle.c -- the reference's canvas_pixel_at as a callable (TEST INFRASTRUCTURE).
 * Linked with the UNMODIFIED /root/reference/src/libs/canvas/canvas.c and src/color/*.c into oracle/_ref/libcanvas_ref.so by
 * oracle/build_ref.py; only tests/ load it (tests/test_texture_ref.py pins oracle/texture_ref.py against it).
 */
#include <stdlib.h>
#include <string.h>

#include "src/libs/canvas/canvas.h"
#include "src/color/rgb.h"
#include "src/color/srgb.h"

/* raw: h * w * 3 doubles; out: h * w * 3 doubles = canvas_pixel_at(col, row) for every texel */
int
canvas_oracle_pixels(const double *raw, int width, int height, int super_sample, int srgb, double *out)
{
    Canvas c = canvas_alloc((size_t)width, (size_t)height, super_sample != 0, srgb ? srgb_to_rgb : rgb_to_rgb);
    if (c == NULL) {
        return 1;
    }
    for (size_t i = 0; i < (size_t)width * height; ++i) {
        c->arr[i][0] = raw[3 * i];
        c->arr[i][1] = raw[3 * i + 1];
        c->arr[i][2] = raw[3 * i + 2];
    }
    for (int row = 0; row < height; ++row) {
        for (int col = 0; col < width; ++col) {
            Color res;
            canvas_pixel_at(c, col, row, res);
            double *o = out + 3 * ((size_t)row * width + col);
            o[0] = res[0];
            o[1] = res[1];
            o[2] = res[2];
        }
    }
    canvas_free(c);
    return 0;
}
