/*
 * frt_lightgen.cuh -- the reference's area-light sample cache, rebuilt on the device.
 *
 * construct_area_light_surface_points_cache (light.c:155-191) fills cache_size correlated-multi-jitter sample sets at
 * scene-build time, single threaded, from the process's drand48 stream: per set one sampler.reset() =
 * sampler_reset_canonical_2d + sampler_shuffle_2d (sampler.c:415-461), 2 u v + u + v draws, then
 * area_light_point_on_light (light.c:138-153) per sample.  The shipped Cornell light has 65 535 sets of 100 points:
 * 210 MB on the host, 157 MB flattened, which the drop-in used to repack (115 ms) and push over PCIe (3 ms pinned,
 * 25 ms pageable) per scene -- per rank at N GPUs.  The cache is a pure function of (corner, uvec, vvec, usteps,
 * vsteps, cache_size) and the state of the 48-bit LCG before the first draw, so it is rebuilt here instead:
 *
 *   - drand48 is X' = A X + C mod 2^48, result X' / 2^48 (glibc, drand48-iter.c); set s starts at J^s(X0) with J the
 *     jump over one set's draws -- the host uploads J^(2^b) as (multiplier, increment) pairs and a warp composes the
 *     ones s has bits for;
 *   - one warp per set: the lanes produce the set's draws into shared memory (lane l the draws l, l + 32, ... by jumping
 *     ahead in the stream), all lanes evaluate the canonical pass, the row / column swaps are applied one by one exactly as
 *     written (lanes over the swapped row / column), then the points leave coalesced -- FP64 for the exact kernels,
 *     FP32 for the filter and the lighting sums (k_to_float is not needed for a generated light);
 *   - every operation is an explicitly rounded IEEE double operation (__dadd_rn ...: no contraction into FMAs), the
 *     same sequence the host executes, so the points are the reference's bit for bit.  That claim is CHECKED per scene:
 *     the caller hands a few sets the reference built (first, last, some in between) and k_light_gen_verify compares
 *     them bit by bit; on any difference frt_scene_create_gen fails with FRT_ERR_MISMATCH and the caller uploads the
 *     host cache as before.  tests/test_lightcache.py pins the same arithmetic (numpy restatement) against a cache the
 *     reference's own constructor produced; tests/test_gpu_lightgen.py compares all sets of that fixture.
 */
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#define FRT_LCG_A 0x5DEECE66DULL
#define FRT_LCG_C 0xBULL
#define FRT_LCG_MASK ((1ULL << 48) - 1ULL)
#define FRT_LGEN_WARPS 8      /* sets per block */
#define FRT_LGEN_MAX_STEPS 64 /* usteps, vsteps <= 64 (dynamic shared memory: (4 u v + u + v) doubles per warp) */

struct LcgJump {
    unsigned long long a, c; /* X -> a X + c mod 2^48 */
};

/* (a, c) of n steps of the generator */
static inline LcgJump
lcg_jump(unsigned long long n)
{
    LcgJump r{ 1ULL, 0ULL }, p{ FRT_LCG_A, FRT_LCG_C };
    while (n) {
        if (n & 1ULL) { /* r = p o r */
            r.a = (p.a * r.a) & FRT_LCG_MASK;
            r.c = (p.a * r.c + p.c) & FRT_LCG_MASK;
        }
        /* p = p o p */
        const unsigned long long pa = p.a;
        p.a = (pa * pa) & FRT_LCG_MASK;
        p.c = (pa * p.c + p.c) & FRT_LCG_MASK;
        n >>= 1;
    }
    return r;
}

struct LightGenParams {
    double corner[3], uvec[3], vvec[3];
    int usteps, vsteps, cache_len;
    unsigned long long x0; /* generator state before the first draw of set 0 */
    LcgJump pow2[32];      /* pow2[b] = jump over 2^b sets */
    LcgJump step2[6];      /* step2[b] = jump over 2^b draws: lane l starts l + 1 draws into its set and strides by 32 */
};

/*
 * grid: ceil(cache_len / FRT_LGEN_WARPS) blocks of FRT_LGEN_WARPS warps; dynamic shared memory:
 * FRT_LGEN_WARPS * (4 u v + u + v) doubles (the draws, then the table).
 */
__global__ void __launch_bounds__(FRT_LGEN_WARPS * 32)
k_light_gen(LightGenParams P, double *__restrict__ out64, float *__restrict__ out32)
{
    extern __shared__ double s_gen[];
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int s0 = P.usteps, s1 = P.vsteps, NS = s0 * s1;
    const int per_set = 2 * NS + s0 + s1;
    double *xi = s_gen + (size_t)wib * (per_set + 2 * NS);
    double *arr = xi + per_set;
    const int set = blockIdx.x * FRT_LGEN_WARPS + wib;
    if (set >= P.cache_len) {
        return; /* whole warps leave; nothing below synchronises across warps */
    }
    {
        /* draw k of the set is the state after k + 1 steps from the set's start: lane l jumps l + 1 steps ahead (composed from
         * the power-of-two jumps) and then strides by 32 -- the draws are the stream's, only the order of evaluation differs */
        unsigned long long x = P.x0;
        for (int b = 0; b < 32; ++b) {
            if ((set >> b) & 1) {
                x = (P.pow2[b].a * x + P.pow2[b].c) & FRT_LCG_MASK;
            }
        }
        const int ahead = lane + 1;
        for (int b = 0; b < 6; ++b) {
            if ((ahead >> b) & 1) {
                x = (P.step2[b].a * x + P.step2[b].c) & FRT_LCG_MASK;
            }
        }
        for (int k = lane; k < per_set; k += 32) {
            xi[k] = (double)x * (1.0 / 281474976710656.0); /* exact: x < 2^48, the factor is a power of two */
            x = (P.step2[5].a * x + P.step2[5].c) & FRT_LCG_MASK;
        }
    }
    __syncwarp();
    /* sampler_reset_canonical_2d (sampler.c:415-430): n = steps[0], m = steps[1]; entry (j, i) draws 2 (j m + i), +1 */
    {
        const int n = s0, m = s1;
        const double dn = (double)n, dm = (double)m;
        for (int e = lane; e < NS; e += 32) {
            const int j = e / m, i = e - j * m;
            arr[2 * e] = __ddiv_rn(__dadd_rn((double)i, __ddiv_rn(__dadd_rn((double)j, xi[2 * e]), dn)), dm);
            arr[2 * e + 1] = __ddiv_rn(__dadd_rn((double)j, __ddiv_rn(__dadd_rn((double)i, xi[2 * e + 1]), dm)), dn);
        }
    }
    __syncwarp();
    /* sampler_shuffle_2d (sampler.c:433-461): m = steps[0], n = steps[1] */
    {
        const int m = s0, n = s1;
        for (int j = 0; j < n; ++j) {
            const int k = (int)__dadd_rn((double)j, __dmul_rn(xi[2 * NS + j], (double)(n - j)));
            if (k != j) {
                for (int i = lane; i < m; i += 32) {
                    const double t = arr[2 * (j * m + i)];
                    arr[2 * (j * m + i)] = arr[2 * (k * m + i)];
                    arr[2 * (k * m + i)] = t;
                }
            }
            __syncwarp();
        }
        for (int i = 0; i < m; ++i) {
            const int k = (int)__dadd_rn((double)i, __dmul_rn(xi[2 * NS + n + i], (double)(m - i)));
            if (k != i) {
                for (int j = lane; j < n; j += 32) {
                    const double t = arr[2 * (j * m + i) + 1];
                    arr[2 * (j * m + i) + 1] = arr[2 * (j * m + k) + 1];
                    arr[2 * (j * m + k) + 1] = t;
                }
            }
            __syncwarp();
        }
    }
    /* sample (u, v) = table entry v * steps[0] + u (sampler_get_point_2d :472-477), stored at v * usteps + u (light.c:172-186);
     * jitter scaled by the step counts, then corner + uvec * j0 + vvec * j1 component by component (light.c:138-153) */
    double *o64 = out64 + (size_t)set * NS * 3;
    float *o32 = out32 + (size_t)set * NS * 3;
    for (int w = lane; w < 3 * NS; w += 32) {
        const int e = w / 3, c = w - 3 * e;
        const double j0 = __dmul_rn(arr[2 * e], (double)s0), j1 = __dmul_rn(arr[2 * e + 1], (double)s1);
        const double p = __dadd_rn(__dadd_rn(P.corner[c], __dmul_rn(P.uvec[c], j0)), __dmul_rn(P.vvec[c], j1));
        o64[w] = p;
        o32[w] = (float)p;
    }
}

/* bit-for-bit comparison of one generated set with the set the reference built; *mismatch counts differing words */
__global__ void
k_light_gen_verify(const double *__restrict__ generated, const double *__restrict__ expected, int n_words, unsigned int *mismatch)
{
    unsigned int bad = 0;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n_words; k += gridDim.x * blockDim.x) {
        bad += __double_as_longlong(generated[k]) != __double_as_longlong(expected[k]);
    }
    if (bad) {
        atomicAdd(mismatch, bad);
    }
}
