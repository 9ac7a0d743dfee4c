/*
 * frt_encode.cuh -- the reference's PPM encoder on the device (output encode).
 *
 * construct_ppm (src/libs/canvas/canvas.c:150-301), called by the generated main() as write_ppm_file(c, true, path)
 * (yaml_parser/yaml_parser.py:220), makes three passes over the float64 canvas on one host thread, with two pow() per
 * channel -- 0.2 s for an 800 x 800 frame the device renders in 0.02 s.  Here the canvas never leaves the device: two
 * reduction kernels for the per-channel maxima, one kernel that writes the big-endian 16-bit samples, and the host only
 * copies 6 bytes per pixel.  Byte for byte the reference's output: every operation is the IEEE double operation the C
 * code performs, in its order, never fused (checked against files the reference wrote, tests/golden/ppm_*.npz).
 *
 *     rgb_max[c]  = max(0, max over pixels of rgb[c])                                   canvas.c:184-196
 *     srgb_max[c] = max(0, max over pixels of rgb_to_srgb(rgb / rgb_max)[c])            canvas.c:199-215
 *     per pixel: use_scaling and r + g + b > sqrt(3): rgb = (rgb * (1 / sum)) * sqrt(3)  canvas.c:236-241
 *                otherwise clamp to [0, 1]                                              canvas.c:247-263
 *                srgb = rgb_to_srgb(rgb)                                                rgb.c:66-77
 *                65535 if srgb > srgb_max, 0 if srgb < 0, else (uint16_t)floor(srgb * (65535 / srgb_max))
 */
#pragma once

__device__ __forceinline__ double
enc_rgb_to_srgb(double x)
{
    /* x < 0.0031308 ? x * 12.92 : 1.055 * pow(x, 1 / 2.4) - 0.055 -- a NaN takes the pow branch, like the C ternary */
    if (x < 0.0031308) {
        return __dmul_rn(x, 12.92);
    }
    return __dsub_rn(__dmul_rn(1.055, pow(x, 1.0 / 2.4)), 0.055);
}

/* maxima start at 0 (color_default = BLACK) and `v > max` is false for a NaN: only positive values can raise them,
 * and the bit patterns of positive doubles order like unsigned integers */
__device__ __forceinline__ void
enc_atomic_max(double *dst, double v)
{
    if (v > 0.0) {
        atomicMax(reinterpret_cast<unsigned long long *>(dst), (unsigned long long)__double_as_longlong(v));
    }
}

__device__ __forceinline__ void
enc_block_max3(double m0, double m1, double m2, double *dst)
{
    for (int o = 16; o > 0; o >>= 1) {
        m0 = fmax(m0, __shfl_xor_sync(0xffffffffu, m0, o));
        m1 = fmax(m1, __shfl_xor_sync(0xffffffffu, m1, o));
        m2 = fmax(m2, __shfl_xor_sync(0xffffffffu, m2, o));
    }
    if ((threadIdx.x & 31) == 0) {
        enc_atomic_max(dst, m0);
        enc_atomic_max(dst + 1, m1);
        enc_atomic_max(dst + 2, m2);
    }
}

/* maxes[0..2] = rgb_max */
__global__ void __launch_bounds__(256)
k_ppm_rgb_max(const double *__restrict__ canvas, size_t n_pixels, double *__restrict__ maxes)
{
    double m[3] = { 0.0, 0.0, 0.0 };
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_pixels; i += (size_t)gridDim.x * blockDim.x) {
        const double2 a = *reinterpret_cast<const double2 *>(canvas + 4 * i);
        const double b = canvas[4 * i + 2];
        if (a.x > m[0]) m[0] = a.x;
        if (a.y > m[1]) m[1] = a.y;
        if (b > m[2]) m[2] = b;
    }
    enc_block_max3(m[0], m[1], m[2], maxes);
}

/* maxes[3..5] = srgb_max, from maxes[0..2] */
__global__ void __launch_bounds__(256)
k_ppm_srgb_max(const double *__restrict__ canvas, size_t n_pixels, double *__restrict__ maxes)
{
    const double r0 = maxes[0], r1 = maxes[1], r2 = maxes[2];
    double m[3] = { 0.0, 0.0, 0.0 };
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_pixels; i += (size_t)gridDim.x * blockDim.x) {
        const double2 a = *reinterpret_cast<const double2 *>(canvas + 4 * i);
        const double b = canvas[4 * i + 2];
        const double s0 = enc_rgb_to_srgb(a.x / r0), s1 = enc_rgb_to_srgb(a.y / r1), s2 = enc_rgb_to_srgb(b / r2);
        if (s0 > m[0]) m[0] = s0;
        if (s1 > m[1]) m[1] = s1;
        if (s2 > m[2]) m[2] = s2;
    }
    enc_block_max3(m[0], m[1], m[2], maxes + 3);
}

__device__ __forceinline__ unsigned int
enc_sample(double srgb, double srgb_max, double inverse)
{
    if (srgb > srgb_max) {
        return 65535u;
    }
    if (srgb < 0.0) {
        return 0u;
    }
    const double v = floor(__dmul_rn(srgb, inverse));
    /* (uint16_t) of a NaN (0 * inf on an all-black channel) is 0 on the reference's target; in-range values convert as is */
    return (v >= 0.0 && v <= 65535.0) ? (unsigned int)v : (v > 65535.0 ? 65535u : 0u);
}

__global__ void __launch_bounds__(256)
k_ppm_encode(const double *__restrict__ canvas, size_t n_pixels, const double *__restrict__ maxes, int use_scaling,
             unsigned char *__restrict__ out)
{
    const double sm0 = maxes[3], sm1 = maxes[4], sm2 = maxes[5];
    const double i0 = 65535.0 / sm0, i1 = 65535.0 / sm1, i2 = 65535.0 / sm2;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_pixels; i += (size_t)gridDim.x * blockDim.x) {
        const double2 a = *reinterpret_cast<const double2 *>(canvas + 4 * i);
        double r = a.x, g = a.y, b = canvas[4 * i + 2];
        if (use_scaling) {
            const double len = __dadd_rn(__dadd_rn(r, g), b);
            if (len > 1.7320508075688772) { /* sqrt(3) as max magnitude */
                const double s = 1.0 / len;
                r = __dmul_rn(__dmul_rn(r, s), 1.7320508075688772);
                g = __dmul_rn(__dmul_rn(g, s), 1.7320508075688772);
                b = __dmul_rn(__dmul_rn(b, s), 1.7320508075688772);
            }
        } else {
            r = r > 1.0 ? 1.0 : (r < 0 ? 0.0 : r);
            g = g > 1.0 ? 1.0 : (g < 0 ? 0.0 : g);
            b = b > 1.0 ? 1.0 : (b < 0 ? 0.0 : b);
        }
        const unsigned int vr = enc_sample(enc_rgb_to_srgb(r), sm0, i0);
        const unsigned int vg = enc_sample(enc_rgb_to_srgb(g), sm1, i1);
        const unsigned int vb = enc_sample(enc_rgb_to_srgb(b), sm2, i2);
        /* 6 bytes per pixel, big-endian: three 16-bit stores (the data starts at an even offset of its own buffer) */
        unsigned short *o = reinterpret_cast<unsigned short *>(out + 6 * i);
        o[0] = (unsigned short)(((vr & 0xffu) << 8) | (vr >> 8));
        o[1] = (unsigned short)(((vg & 0xffu) << 8) | (vg >> 8));
        o[2] = (unsigned short)(((vb & 0xffu) << 8) | (vb >> 8));
    }
}
