/*
 * frt_b200.h -- C ABI of libfrt_b200.so, the B200-native render core that
 * replaces the per-pixel hot path of gbordelon/fast_ray_tracer.
 *
 * Boundary: the reference's generated main() calls
 *     trace_photons(w, 3, caustics, final_gather)      (yaml_parser/yaml_parser.py:209)
 *     Canvas c = render_multi(cam, w, usteps, vsteps, jitter)   (yaml_parser.py:218)
 * declared in src/renderer/renderer.h:41-47 and src/renderer/photon_tracer.h:4-6.
 * The host shim (fast_ray_tracer_b200/csrc/frt_shim.c) keeps those symbols,
 * walks the finished World/Camera once, fills a frt_scene_desc (below) and
 * calls the entry points declared here.  Nothing in this header is a C++ or
 * torch type: plain structs, caller-owned host buffers, int return codes
 * (0 = ok, non-zero = error, text via frt_last_error()).
 *
 * There is NO CPU fallback behind this ABI: every compute entry point fails
 * with FRT_ERR_CUDA when no sm_100-class device is usable.
 */
#ifndef FRT_B200_H
#define FRT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FRT_ABI_VERSION 6 /* struct layouts (scene blobs carry it); entry points added since: see FRT_API_LEVEL */
#define FRT_API_LEVEL 3   /* 2: frt_scene_create_gen, frt_render_multi, frt_photons_estimate, frt_light_cache_checksum; 3: frt_tree_with_runs, frt_shared_buffer_* */

enum frt_status {
    FRT_OK = 0,
    FRT_ERR_ARG = 1,      /* malformed scene description / bad argument */
    FRT_ERR_CUDA = 2,     /* CUDA runtime error or no usable device */
    FRT_ERR_IO = 3,       /* scene blob read/write */
    FRT_ERR_OVERFLOW = 4, /* a per-ray bounded stack (CSG interval list) overflowed */
    FRT_ERR_MISMATCH = 5  /* frt_scene_create_gen: a light-sample set rebuilt on the device differs from the caller's */
};

/* Same numbering as the reference's enum shape_enum (src/shapes/shapes.h:16-27). */
enum frt_node_type {
    FRT_CONE = 0, FRT_CUBE = 1, FRT_CYLINDER = 2, FRT_PLANE = 3, FRT_SMOOTH_TRIANGLE = 4,
    FRT_SPHERE = 5, FRT_TOROID = 6, FRT_TRIANGLE = 7, FRT_CSG = 8, FRT_GROUP = 9
};

/* enum csg_ops_enum, src/shapes/shapes.h:29-33 */
enum frt_csg_op { FRT_CSG_UNION = 0, FRT_CSG_INTERSECT = 1, FRT_CSG_DIFFERENCE = 2 };

/*
 * One node of the reference's shape tree (struct shape, shapes.h:85-118) after
 * group_divide, flattened in DFS PRE-ORDER with the reference's child order kept
 * (shadow rays depend on it: group.c:105-123).  `skip` is the index of the first
 * node after this node's subtree, so a traversal needs no stack:
 *     enter a group/CSG -> i+1,   cull or leave a leaf -> skip.
 */
typedef struct frt_node {
    int32_t type;        /* enum frt_node_type */
    int32_t skip;
    int32_t parent;      /* -1 for a root */
    int32_t xform;       /* index into frt_scene_desc.xforms: world->node-local composite; 0 = identity */
    int32_t material;    /* leaves: index into materials; group/CSG: -1 */
    int32_t param;       /* leaves: offset (in doubles) into prim_params; -1 when the type has none */
    int32_t csg_op;      /* CSG only (enum frt_csg_op) */
    int32_t right;       /* CSG only: pre-order index of the right child (left child is self+1) */
    double bbox_min[3];  /* group/CSG: bounds in the node's own space (group_bounds group.c:373, csg_bounds csg.c:150) */
    double bbox_max[3];
} frt_node;

/*
 * prim_params layout, by node type (all doubles):
 *   cylinder / cone : [minimum, maximum, closed(0|1)]                 (shapes.h:54-58)
 *   toroid          : [r1, r2]                                        (shapes.h:60-63)
 *   triangle / smooth triangle, FRT_TRI_PARAMS doubles:               (shapes.h:65-82)
 *       p1[3] p2[3] p3[3] e1[3] e2[3] n1[3] n2[3] n3[3] t1[3] t2[3] t3[3] use_textures
 *       (flat triangle: n1 = n2 = n3 = normal)
 */
#define FRT_TRI_PARAMS 34

/* world->local composite: rows 0..2 of the 4x4 inverse, row-major [r0c0 r0c1 r0c2 r0c3 r1c0 ...] */
typedef struct frt_xform {
    double inv[12];
} frt_xform;

/* struct material, src/material/material.h:196-220 (only the fields the path reads) */
typedef struct frt_material {
    double Ka[3], Kd[3], Ks[3], Tf[3], refl[3];
    double Ns, Ni, Tr;
    int32_t casts_shadow;
    int32_t reflective;
    /* pattern indices or -1: map_Ka, map_Kd, map_Ks, map_Ns, map_d, map_bump, map_refl */
    int32_t map_Ka, map_Kd, map_Ks, map_Ns, map_d, map_bump, map_refl;
    int32_t pad;
} frt_material;

/* enum pattern_type, src/pattern/pattern.h:24-45 (same numbering) */
enum frt_pattern_type {
    FRT_PAT_CHECKER = 0, FRT_PAT_GRADIENT = 1, FRT_PAT_RADIAL_GRADIENT = 2, FRT_PAT_RING = 3,
    FRT_PAT_STRIPE = 4, FRT_PAT_UV_ALIGN_CHECKER = 5, FRT_PAT_UV_CHECKER = 6, FRT_PAT_UV_GRADIENT = 7,
    FRT_PAT_UV_RADIAL_GRADIENT = 8, FRT_PAT_UV_TEXTURE = 9, FRT_PAT_BLENDED = 10, FRT_PAT_NESTED = 11,
    FRT_PAT_PERTURBED = 12, FRT_PAT_CUBE_MAP = 13, FRT_PAT_CYLINDER_MAP = 14, FRT_PAT_TEXTURE_MAP = 15
};

/* enum uv_map_type, pattern.h:94-101 */
enum frt_uv_map { FRT_UV_CUBE = 0, FRT_UV_CYLINDER = 1, FRT_UV_PLANE = 2, FRT_UV_SPHERE = 3, FRT_UV_TOROID = 4, FRT_UV_TRIANGLE = 5 };

/*
 * struct pattern, pattern.h:119-142.
 *   concrete (checker..stripe, uv gradient/radial): c[0..2]=a c[3..5]=b
 *   uv align check : c[0..14] = main, ul, ur, bl, br
 *   uv checker     : c = a,b ; i[0]=width i[1]=height
 *   uv texture     : i[0] = texture index
 *   blended        : i[0], i[1] = child patterns
 *   nested         : i[0], i[1], i[2] = pattern1..3
 *   perturbed      : i[0] = child, i[1] = octaves, i[2] = seed ; f[0]=frequency f[1]=scale_factor f[2]=persistence
 *   texture map    : i[0] = enum frt_uv_map, i[1] = index of first face pattern (faces contiguous)
 */
typedef struct frt_pattern {
    int32_t type;
    int32_t identity;   /* transform_identity */
    double inv[12];     /* rows 0..2 of transform_inverse */
    double c[15];
    double f[4];
    int32_t i[4];
} frt_pattern;

/* Texture = struct canvas used as an image (canvas.h:10-17).  The description carries the texels RAW, as read_png left
 * them; frt_scene_create evaluates canvas_pixel_at (canvas.c:115-148: the 3x3 wrap-around box of a super-sampled canvas,
 * then colour_fn) once per texel and keeps linear FP32 RGBA on the device (texture ingest). */
enum frt_color_fn { FRT_COLOR_RGB = 0, FRT_COLOR_SRGB_TO_RGB = 1 };
typedef struct frt_texture {
    int32_t width, height;
    int32_t super_sample;
    int32_t color_fn;       /* enum frt_color_fn */
    int64_t texel_offset;   /* offset in texels (3 doubles each) into frt_scene_desc.texels */
} frt_texture;

/* struct light, src/light/light.h:57-76; point/hemisphere lights are num_samples=1, cache_len=1 */
typedef struct frt_light {
    int32_t type;           /* enum light_enum, light.h:14-20 */
    int32_t num_samples;
    int32_t cache_len;      /* number of pre-computed sample sets */
    int32_t usteps, vsteps, jitter;
    double intensity[3];
    double position[3];     /* point/hemi: position; area: corner; circle: origin */
    double normal[3];       /* hemi/circle: normal; area: normalize(uvec x vvec) */
    double uvec[3], vvec[3];/* area light cell vectors (already divided by steps) */
    double radius;
    int64_t point_offset;   /* offset in points (3 doubles each) into light_points; set s, sample k at point_offset + s*num_samples + k */
} frt_light;

/* struct camera + struct aperture, src/renderer/camera.h:44-73 */
typedef struct frt_camera {
    int32_t hsize, vsize, usteps, vsteps;
    double half_width, half_height, pixel_size, canvas_distance;
    double inv[16];              /* transform_inverse, full 4x4 row-major */
    int32_t aperture_type;       /* enum aperture_type, camera.h:9-19 */
    int32_t aperture_jitter;
    double aperture_size;
    double aperture_args[4];     /* union u: circle r1 | cross x1 x2 y1 y2 | diamond b1..b4 | doughnut r1 r2 */
} frt_camera;

/* struct global_config, src/renderer/config.h:4-62 (the fields setup_config reads, renderer.c:53-71) */
typedef struct frt_config {
    int32_t include_direct, include_global;
    int32_t visualize_photon_map, visualize_soft_indirect;
    int32_t di_include_ambient, di_include_diffuse, di_include_specular_highlight, di_include_specular;
    int32_t di_path_length;
    int32_t gi_include_caustics, gi_include_final_gather;
    int32_t gi_usteps, gi_vsteps;
    int32_t gi_irradiance_estimate_num;
    int32_t gi_path_length;
    int32_t pad;
    double gi_irradiance_estimate_radius, gi_irradiance_estimate_cone_filter_k;
    int64_t gi_photon_count;
} frt_config;

typedef struct frt_scene_desc {
    int32_t abi_version;         /* FRT_ABI_VERSION */
    int32_t n_nodes, n_roots, n_xforms, n_materials, n_patterns, n_textures, n_lights;
    int64_t n_prim_params, n_texels, n_light_points, n_pixel_samples;
    const frt_node *nodes;
    const int32_t *roots;        /* pre-order indices of World.shapes[0..shapes_num) (world.h:25-34) */
    const frt_xform *xforms;     /* xforms[0] must be the identity */
    const double *prim_params;
    const frt_material *materials;
    const frt_pattern *patterns;
    const frt_texture *textures;
    const double *texels;        /* n_texels * 3 */
    const frt_light *lights;
    const double *light_points;  /* n_light_points * 3 */
    const double *pixel_samples; /* optional host-built CMJ table (sampler_2d, sampler.c:510), 2*usteps*vsteps doubles,
                                    used when the primary jitter flag is false; NULL -> the core derives the xi=0.5 table */
    frt_camera camera;
    frt_config config;
} frt_scene_desc;

typedef struct frt_scene frt_scene;   /* opaque: device-resident scene */

enum frt_render_flags {
    FRT_FLAG_NO_PRUNE = 1,   /* also trace branches whose weight is exactly zero, like the reference does
                                (renderer.c:534-605 on opaque surfaces) -- for ray-count parity only */
    FRT_FLAG_COUNT_RAYS = 2, /* fill the ray counters in frt_stats */
    FRT_FLAG_F64_SHADOW = 8, /* trace every shadow ray in FP64 (no FP32 filter pass) */
    FRT_FLAG_VERIFY_F32 = 16,/* trace every shadow ray the FP32 filter decides in FP64 as well and count disagreements
                                in frt_stats.shadow_mismatch (must be 0); the frame itself uses the FP64 answers */
    FRT_FLAG_NO_SHAFT = 32,  /* switch the per-hit shaft culling of the shadow filter off (A/B measurements, tests) */
    FRT_FLAG_NO_BULK = 64,   /* switch the per-hit decision of all shadow rays at once (k_shadow_bulk) off */
    FRT_FLAG_NO_SPLIT = 128, /* ... keep it, but do not retry undecided hits per quadrant of the light's sample grid */
    FRT_FLAG_F64_SHAFT = 512, /* per-hit / per-quadrant shaft walk in FP64 (the first version) instead of FP32 with outward slack */
    FRT_FLAG_STAGE_TIMES = 256, /* bracket every kernel of the frame with events and fill frt_stats.stage_ms completely (the
                                shadow-ray and shaft stages are always timed) */
    FRT_FLAG_F64_SHADING = 4 /* evaluate the lighting sums (lighting_microfacet, renderer.c:894-979) in FP64 like the
                                reference instead of FP32; geometric decisions are FP64 either way */
};

typedef struct frt_render_cfg {
    int32_t device;          /* CUDA device ordinal */
    int32_t rank, world;     /* this call renders row blocks b with b % world == rank (world <= 0 -> 1) */
    int32_t rows_per_block;  /* rows in a block (<= 0 -> 4) */
    int32_t usteps, vsteps;  /* <= 0 -> camera's */
    int32_t jitter;          /* < 0 -> camera.aperture_jitter */
    int32_t flags;
    uint64_t seed;           /* counter-based RNG seed (sample-set picks, jitter, aperture, photons) */
} frt_render_cfg;

/* stages of a frame, index into frt_stats.stage_ms */
enum frt_stage {
    FRT_ST_RAYGEN = 0, FRT_ST_EXTEND = 1, FRT_ST_SHADE = 2, FRT_ST_LIGHT_PRE = 3,
    FRT_ST_SHADOW_SHAFT = 4,  /* k_shadow_bulk + k_shadow_quad: all shadow rays of a hit / a quadrant of the light at once */
    FRT_ST_SHADOW_RAY = 5,    /* the per-ray kernel: k_shadow_f32 (k_shadow_mesh on mesh scenes, k_shadow_exact under F64_SHADOW) */
    FRT_ST_SHADOW_EXACT = 6,  /* FP64 re-trace of the rays the FP32 filter deferred */
    FRT_ST_LIGHT_FINAL = 7, FRT_ST_GI_TRACE = 8, FRT_ST_KNN = 9, FRT_ST_GI_RESOLVE = 10, FRT_ST_COUNT = 12
};

typedef struct frt_stats {
    double frame_ms;         /* device time of the frame, CUDA events on the render stream */
    double light_ms;         /* summed device time of the launches of the dominant kernel alone (= stage_ms[FRT_ST_SHADOW_RAY]):
                                k_shadow_f32; k_shadow_mesh on mesh scenes; k_shadow_exact under FRT_FLAG_F64_SHADOW */
    double upload_ms, download_ms;
    uint64_t rays_primary, rays_secondary, rays_shadow, rays_gather, rays_photon;
    uint64_t hits_shaded;
    uint64_t light_launches; /* launches of the dominant kernel */
    uint64_t kernel_launches;/* all launches in the frame */
    uint64_t shadow_nodes;   /* tree nodes visited by shadow rays */
    uint64_t overflow;       /* non-zero: a bounded queue overflowed and the frame was re-run in smaller chunks */
    uint64_t photons_stored[3];
    uint64_t light_flops;    /* with FRT_FLAG_COUNT_RAYS: algorithmic flop of the dominant kernel, counted event by event
                                with the cost table of BASELINE.md section 4 (ray transform 33, bbox slab 16, sphere 28, ...) */
    uint64_t shadow_deferred;/* shadow rays the FP32 filter pass left undecided and the FP64 pass re-traced */
    uint64_t shadow_mismatch;/* FRT_FLAG_VERIFY_F32: FP32-decided rays whose FP64 answer differs */
    uint64_t shadow_reasons[10]; /* [0]: shadow rays decided per hit, all at once, by the shaft-interval pass (never traced).
                                FRT_FLAG_COUNT_RAYS: deferred rays by reason: 1 CSG depth, 2 group inside a CSG operand,
                                3 primitive type without a fast form, 4 plane inside a CSG, 5 leaf hit/miss not separable,
                                6 leaf verdict, 7 two spans in one operand, 8 CSG ordering / two-span result, 9 CSG verdict */
    int32_t rows_rendered;
    int32_t pad;
    double stage_ms[12];     /* summed device time per stage (enum frt_stage); [4] and [5] always, the rest under FRT_FLAG_STAGE_TIMES */
    uint64_t shadow_ray_launches; /* launches behind stage_ms[FRT_ST_SHADOW_RAY] */
    uint64_t shadow_rays_traced;  /* FRT_FLAG_COUNT_RAYS: rays the per-ray kernel was handed (pending entries x samples) */
} frt_stats;

typedef struct frt_photon_cfg {
    int32_t device;
    int32_t rank, world;     /* emission shard: this call emits photon indices i with i % world == rank */
    int32_t populate_caustic, populate_global;
    int32_t pad;
    uint64_t seed;
} frt_photon_cfg;

/* ---- entry points ---------------------------------------------------- */

int frt_abi_version(void);
const char *frt_last_error(void);
int frt_device_count(void);
/* layout check for foreign-language bindings: sizeof of a named ABI struct ("frt_node", "frt_light", ...), or -1 */
int frt_abi_sizeof(const char *struct_name);

/* Upload a flattened scene (host pointers in desc are read during the call only; the one exception is a light_points
 * buffer registered with frt_host_register, see there). */
int frt_scene_create(const frt_scene_desc *desc, int device, frt_scene **out);
/*
 * The same, with the sample caches of jittered rectangular area lights REBUILT ON THE DEVICE instead of uploaded.
 * construct_area_light_surface_points_cache (src/light/light.c:155-191) is a pure function of the light's corner /
 * uvec / vvec / steps, cache_size and the state of the process's drand48 generator before the constructor's first draw
 * of set 0 (i.e. after the sampler_2d() call at :166 consumed one table of draws); the shipped Cornell light's cache is
 * 157 MB flattened, and rebuilding it costs less than copying it once.  For every entry of `gens`:
 *   - desc->lights[gen.light] must be a jittered area light (type 0) with usteps * vsteps == num_samples;
 *   - the light's region of desc->light_points is NOT read (desc->light_points may be NULL when every light is listed);
 *   - gen.verify_points[k] (k < gen.n_verify <= FRT_GEN_VERIFY_MAX): set gen.verify_set[k] as the reference built it,
 *     num_samples * 3 doubles; each is compared bit for bit with the rebuilt set on the device.
 * The comparison runs on the device BEHIND this call (which only fails for malformed arguments): frt_scene_gen_status
 * waits for it and returns FRT_OK or FRT_ERR_MISMATCH; a caller that does not ask gets the same answer from the scene's
 * first frt_render (FRT_ERR_MISMATCH, nothing written to the canvas; every later call on the scene fails the same way).
 * On a mismatch the caller destroys the scene, flattens the host cache and calls frt_scene_create.
 */
#define FRT_GEN_VERIFY_MAX 8
typedef struct frt_light_gen {
    int32_t light;                 /* index into desc->lights */
    int32_t n_verify;
    uint64_t drand48_state;        /* the 48-bit state X before the first draw of set 0 */
    int32_t verify_set[FRT_GEN_VERIFY_MAX];
    const double *verify_points[FRT_GEN_VERIFY_MAX];
} frt_light_gen;
int frt_scene_create_gen(const frt_scene_desc *desc, int device, const frt_light_gen *gens, int n_gens, frt_scene **out);
int frt_scene_gen_status(frt_scene *scene);
/* drand48 state after `draws` draws from state `x` (X' = 0x5DEECE66D X + 0xB mod 2^48; glibc starts from X = 0 when
 * srand48 was never called): what a caller needs to fill frt_light_gen.drand48_state. */
uint64_t frt_drand48_advance(uint64_t x, uint64_t draws);
/* 64-bit checksum (wrapping sum of the words' bit patterns, each multiplied by an odd function of its index) of the
 * scene's FP64 light-point pool on the device, points [first, first + count): lets a caller compare a whole rebuilt
 * cache with the host's without moving either. frt_light_points_checksum_host computes the same over host points. */
int frt_light_points_checksum(frt_scene *scene, int64_t first_point, int64_t n_points, uint64_t *sum);
uint64_t frt_light_points_checksum_host(const double *points, int64_t first_point, int64_t n_points);
void frt_scene_destroy(frt_scene *scene);
/* frt_scene_destroy parks the scene-independent frame buffers (ray queues) and the large scene buffers (by size) for
 * the next scene on the same device; frt_trim frees them. */
void frt_trim(int device);
/*
 * Output encode: write_ppm_file -> construct_ppm (src/libs/canvas/canvas.c:150-328), the 16-bit P6 file
 * the generated main() writes after render_multi (yaml_parser/yaml_parser.py:220), produced on the device -- byte for
 * byte the reference's file contents for the same canvas.  frt_ppm16_size: bytes of the file for a frame size.
 * frt_canvas_encode_ppm16 encodes the device-resident canvas of the scene's last frt_render (the canvas never crosses
 * PCIe as doubles); frt_encode_ppm16 takes a host canvas in the Canvas.arr layout (width * height * 4 doubles).
 * encode_ms (optional): device time of the three kernels.
 */
size_t frt_ppm16_size(int width, int height);
int frt_canvas_encode_ppm16(frt_scene *scene, int use_scaling, unsigned char *out, size_t out_cap, size_t *out_len, double *encode_ms);
int frt_encode_ppm16(const double *canvas_rgba, int width, int height, int use_scaling, int device, unsigned char *out, size_t out_cap,
                     size_t *out_len, double *encode_ms);

/* Texture ingest on its own: raw_rgb = width * height * 3 doubles (an image's Canvas.arr without the 4th lane), out_rgba =
 * width * height * 4 floats: per texel what canvas_pixel_at returns for it, rounded to FP32. */
int frt_texture_ingest(const double *raw_rgb, int width, int height, int super_sample, int color_fn, int device, float *out_rgba);

/* Page-lock / release a caller-owned host buffer a scene description points at (typically light_points, the 157 MB
 * sample-set cache the reference builds in light.c:100-191): frt_scene_create then uploads it at PCIe speed and
 * ASYNCHRONOUSLY -- the first frame waits for it only where its light stage begins, so the copy hides behind ray
 * generation, the primary rays and their shading.  A registered light_points buffer must therefore stay valid and
 * unmodified until the first frt_render of the scene (or frt_scene_destroy) has returned; every other buffer, and an
 * unregistered light_points, is read during frt_scene_create only.  Worth it for a host loop that creates a scene per
 * frame; a one-shot program need not call it. */
int frt_host_register(void *ptr, size_t bytes);
int frt_host_unregister(void *ptr);

/*
 * Replaces render_multi()/render() (renderer.c:243/:283).  canvas_rgba has the layout of Canvas.arr
 * (canvas.h:10-17): hsize*vsize Color = double[4], row-major, linear RGB, 4th lane 0 as the reference leaves it.
 * Only the rows owned by (rank, world) are written.  canvas_rgba may be NULL (frame stays on device;
 * fetch it later with frt_canvas_download).
 */
int frt_render(frt_scene *scene, const frt_render_cfg *cfg, double *canvas_rgba, frt_stats *stats);
int frt_canvas_download(frt_scene *scene, double *canvas_rgba);
/* the frame where it lives: device pointer to hsize*vsize*4 doubles, valid until frt_scene_destroy (for NCCL gathers) */
int frt_canvas_device_ptr(frt_scene *scene, void **device_ptr);
/* rows owned by (rank, world): writes up to cap row indices, returns the count */
int frt_owned_rows(const frt_scene_desc *desc, const frt_render_cfg *cfg, int32_t *rows, int cap);
/*
 * A device buffer other processes of the node can write through CUDA IPC (one process per GPU: the frame is gathered on
 * one device without a collective).  The owner creates it and hands the 64-byte handle to its peers; a peer opens it and
 * passes the pointer to frt_render as `canvas_rgba` (any pointer cudaMemcpyDefault can write is accepted there: pageable
 * or page-locked host memory, device memory, an opened peer buffer) -- its row blocks then travel over NVLink in one
 * strided copy on its render stream.  close: opened != 0 for a pointer from _open, 0 for the owner's.
 */
int frt_shared_buffer_create(int device, size_t bytes, void **device_ptr, void *handle64);
int frt_shared_buffer_open(int device, const void *handle64, void **device_ptr);
int frt_shared_buffer_close(void *device_ptr, int opened);

/*
 * The tree frt_scene_create uploads: the reference's divided tree (group_divide, group.c:300-370, leaves its straddling
 * triangles as direct children without a box of their own) with bounding groups inserted over runs of consecutive
 * triangle children, leaf order untouched.  Returns the node count (-1: malformed description); fills `nodes` / `roots`
 * (n_roots entries) when cap is large enough.  Host only.
 */
int frt_tree_with_runs(const frt_scene_desc *desc, frt_node *nodes, int cap, int32_t *roots);

/*
 * Replaces trace_photons() (photon_tracer.c:203): emits this rank's shard of photons on the device.
 * frt_photons_export/import move the stored photons (32 B records: float3 pos+power packed)
 * so ranks can all-gather them; frt_photons_finish builds the device kd-tree (pm_balance, pm.c:329).
 */
int frt_photons_emit(frt_scene *scene, const frt_photon_cfg *cfg, frt_stats *stats);
int64_t frt_photons_count(frt_scene *scene, int map);
int frt_photons_export(frt_scene *scene, int map, void *host_or_device_dst, int dst_is_device);
int frt_photons_import(frt_scene *scene, int map, const void *src, int64_t count, int src_is_device);
int frt_photons_finish(frt_scene *scene);

/*
 * pm_irradiance_estimate (src/libs/photon_map/pm.c:91-156) for n positions against photon map `map` (0 caustic,
 * 1 global) with the scene's irradiance-estimate radius / num / cone-filter-k: the estimate as pm.c returns it (no
 * caller rescale) and, per position, the function's return value (photons used).  pos_xyz / normal_xyz: n x 3 doubles
 * (positions are rounded to FP32, as every request inside a frame is); irrad_rgb: n x 3 doubles; found: n ints or NULL.
 */
int frt_photons_estimate(frt_scene *scene, int map, int64_t n, const double *pos_xyz, const double *normal_xyz, double *irrad_rgb,
                         int32_t *found);

/*
 * Several GPUs in ONE process: replaces render_multi()'s pthread row pool (src/renderer/renderer.c:244-281) at node
 * scale.  frt_multi_create replicates the scene on every listed device (devices == NULL: all visible ones), each on its
 * own host thread; frt_multi_render lets device k render the row blocks b with b % n == k and write them straight into
 * canvas_rgba (the Canvas.arr layout; a block of rows is one contiguous run, so there is no packing, no collective and
 * no reorder); frt_multi_photons shards the photon emission the same way (photon_tracer.c:203) and exchanges the
 * stored photons device to device before every device bins the full set.  cfg->device / rank / world are ignored
 * (filled per device).  stats: counters summed over the devices, times = the slowest device's.
 */
typedef struct frt_multi frt_multi;
int frt_multi_create(const frt_scene_desc *desc, const int32_t *devices, int n_devices, const frt_light_gen *gens, int n_gens,
                     frt_multi **out);
void frt_multi_destroy(frt_multi *multi);
int frt_multi_device_count(const frt_multi *multi);
frt_scene *frt_multi_scene(frt_multi *multi, int k);
int frt_multi_render(frt_multi *multi, const frt_render_cfg *cfg, double *canvas_rgba, frt_stats *stats);
int frt_multi_photons(frt_multi *multi, const frt_photon_cfg *cfg, frt_stats *stats);

/*
 * Threading contract.  Every entry point may be called from any host thread; frt_last_error() is per thread.  One
 * frt_scene (and one frt_multi) is used by one thread at a time.  Different scenes may be rendered concurrently from
 * different threads, also on the same device (each scene has its own streams, frame buffers and page-locked slot; the
 * buffers parked by frt_scene_destroy are handed out under a lock).
 */

/* Peak FP64/FP32 FMA issue rate of the device, measured by a register-resident FMA loop (TFLOP/s). */
int frt_measure_fma_peak(int device, double *fp64_tflops, double *fp32_tflops);

/* Scene blobs (test fixtures / cross-process hand-off): flat little-endian dump of frt_scene_desc. */
int frt_scene_save(const frt_scene_desc *desc, const char *path);
int frt_scene_load(const char *path, frt_scene_desc **out);
void frt_scene_desc_free(frt_scene_desc *desc);

#ifdef __cplusplus
}
#endif
#endif /* FRT_B200_H */
