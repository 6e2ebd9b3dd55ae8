/*
 * prt.h -- C ABI of the B200 path-tracing core (libprt.so).
 *
 * This is the drop-in boundary for pyrenderer's render loop.  The reference
 * (sontung/pyrenderer) has NO FFI: its boundary is the Python object protocol
 * consumed by main.py / main_taichi.py.  Each entry point below names the
 * reference interface it replaces (paths relative to the reference root); the
 * Python side of the boundary (pyrenderer_b200/_abi.py + core/, io_utils/)
 * binds these with ctypes -- see INTEGRATION.md for the stub a maintainer of
 * the reference would add.
 *
 * Conventions
 *  - every function returns PRT_OK (0) or a negative prt_status; nothing
 *    throws or aborts across the ABI; prt_last_error() gives the text.
 *  - "_dev" pointers are device memory owned by the CALLER (e.g. a torch
 *    tensor's data_ptr()); the library owns the scene, the BVH and the
 *    wavefront queues.  "_host" pointers are ordinary host memory.
 *  - `stream` is a cudaStream_t passed as void* (NULL = default stream).
 *    Calls are stream-ordered and asynchronous unless stated otherwise.
 *  - one context per device; a context is not thread-safe, distinct contexts
 *    are independent (prt_last_error(NULL), the text of a failed prt_create, is
 *    kept per calling thread).  Every call makes the context's device current
 *    for its duration and restores the caller's current device on return.
 *  - there is no CPU fallback: without a CUDA device prt_create fails.
 */
#ifndef PRT_H
#define PRT_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PRT_ABI_VERSION 3

typedef struct prt_ctx prt_ctx;

typedef enum {
    PRT_OK = 0,
    PRT_ERR_INVALID = -1, /* bad argument */
    PRT_ERR_CUDA = -2,    /* CUDA runtime error (text in prt_last_error) */
    PRT_ERR_STATE = -3,   /* call order: e.g. trace before bvh build */
    PRT_ERR_NOMEM = -4
} prt_status;

/* material types: core/bsdf.py:18-65 (lambert, null) and
 * core/bsdf_taichi.py:45-86 (Metal -> mirror/conductor, Dielectric) */
enum { PRT_MAT_LAMBERT = 0, PRT_MAT_EMITTER = 1, PRT_MAT_MIRROR = 2, PRT_MAT_DIELECTRIC = 3,
       PRT_MAT_CONDUCTOR = 4 };

typedef struct {
    float albedo[3];    /* BSDF*.rho / Materials.colors */
    uint32_t type;      /* PRT_MAT_* */
    float ior;          /* dielectric */
    float roughness;    /* conductor fuzz, bsdf_taichi.py:49 */
    uint32_t two_sided; /* 1 == reference `sided == 0` (normal flips to face the ray) */
    uint32_t pad;
    float emission[3];  /* radiance of an emitter (Tungsten primitive "emission", e.g. scene.json
                           [17,12,4]); used by PRT_RENDER_PHYSICAL only -- the reference's own
                           estimator hard-codes light_color and reads the light's albedo */
    uint32_t pad2;
} prt_material; /* 48 bytes */

/* ray record: 2 x float4.  Replaces core/ray.py:5-17 (position, direction,
 * bounds[0], bounds[1]). */
typedef struct {
    float ox, oy, oz, tmin;
    float dx, dy, dz, tmax;
} prt_ray; /* 32 bytes */

/* hit record: replaces the dict {"hit","t","position",...} of
 * mathematics/intersection.py:106-116.  tri == -1 is a miss; tri is the
 * GLOBAL triangle id (core/scene.py:40-46 merged face order); (u,v) are the
 * Moller-Trumbore barycentrics of p1 and p2. */
typedef struct {
    float t, u, v;
    int32_t tri;
} prt_hit; /* 16 bytes */

/* camera: core/camera.py:14-25,41-72.  iview is row-major in the reference's
 * row-vector convention (world = [x y z 1] @ iview). */
typedef struct {
    double iview[16];
    double sensor_w; /* tan(radians(fov)/2) * focal * aspect */
    double sensor_h; /* tan(radians(fov)/2) * focal */
    double focal;
    uint32_t width, height;
    double aperture; /* core/camera.py:63-65: side of the square lens the ray origin is drawn from
                        (camera space x,y in [-aperture/2, aperture/2)); 0 = pinhole (the loader's value) */
} prt_camera;

typedef struct {
    uint64_t seed;        /* Philox key */
    uint32_t spp_begin;   /* samples [spp_begin, spp_end) of every pixel */
    uint32_t spp_end;
    uint32_t max_depth;   /* PathTracer(depth) core/tracing.py:48 */
    uint32_t rr_start;    /* first bounce with Russian roulette; 0xffffffff = off (reference) */
    float light_color[3]; /* core/tracing.py:120 */
    float tmin, tmax;     /* core/tracing.py:127 */
    uint32_t flags;       /* PRT_RENDER_* */
} prt_render_params;

/* bounce-0 closest hit runs in PRT_TRACE_EXACT mode: primary-hit ids bit-exact */
#define PRT_RENDER_EXACT_PRIMARY 1u
/* physically-based light transport instead of the reference's estimator:
 * emitters radiate material.emission from their front side, next-event estimation and BSDF
 * sampling are combined with the power heuristic -- the MIS scheme drafted in the reference's
 * sample_direct_lighting2 (core/tracing.py:56-90: mis_power_heuristic, compute_area_light_pdf,
 * compute_brdf_pdf) with the real light area; a path ends on an emitter.  Renders in this mode
 * agree with media/cornell-box/TungstenRender.exr (tests/test_physical.py). */
#define PRT_RENDER_PHYSICAL 2u
/* measurement: run the counter-instrumented twins of the traversal kernels, so that prt_get_counters
 * reports node_visits / tri_tests over the render's own rays (all bounces).  Same image; not timed. */
#define PRT_RENDER_COUNT 4u

typedef struct {
    uint32_t n_tris, n_nodes, depth, max_leaf_tris;
    uint32_t morton_sorted; /* 1 if the radix sort left the Morton keys in order (self-check) */
    float sah_cost;
    float ms_total, ms_morton, ms_sort, ms_hierarchy, ms_refit, ms_emit; /* device time per phase (events) */
    float ms_wall;        /* host wall clock of the whole prt_bvh_build call: triangle buffer in -> traversable BVH */
    uint32_t morton_bits; /* 30 or 63: key width this build used */
} prt_bvh_stats;

typedef struct {
    uint64_t rays_closest, rays_shadow; /* rays traced since the last reset */
    uint64_t node_visits, tri_tests;    /* only counted by PRT_TRACE_COUNT launches */
    uint64_t flagged_rays;              /* rays re-resolved in FP64 by PRT_TRACE_EXACT */
    uint64_t paths;
    /* SIMT utilisation of the persistent traversal loop (PRT_TRACE_COUNT launches only):
     * loop iterations per warp, lanes doing a record visit summed over iterations, leaf
     * phases and lanes intersecting in them */
    uint64_t warp_iters, node_lane_iters, leaf_phases, leaf_lane_phases;
    /* PRT_TRACE_EXACT | PRT_TRACE_COUNT: triangle tests whose FP32 decision was inside its error
     * bound and were decided in place with the FP64 formula (flagged_rays counts the rays that
     * still needed the FP64 replay afterwards) */
    uint64_t f64_decisions;
} prt_counters;

/* trace flags */
#define PRT_TRACE_EXACT 1u /* bit-exact ids: the FP32 traversal carries forward error bounds; a triangle
                              whose decision is inside its bound is decided with the reference's FP64
                              Moller-Trumbore in place, rays that still cannot be ordered are replayed
                              in FP64; (t,u,v) of every hit = the reference's FP64 values rounded to
                              f32.  Same persistent kernel as the plain mode (about 0.88x its rate). */
#define PRT_TRACE_COUNT 2u /* counter-instrumented twin kernel (roofline N_node / N_tri) */
#define PRT_TRACE_BRUTE 4u /* ignore the BVH: test every triangle (Aggregator semantics,
                              accelerators/aggregator.py:74-85) */
/* Ray binning: before the traversal the rays of the call are counting-sorted by the cell of their origin
 * (8 x 8 x 4 cells of the scene box, one 8-bit pass over ray INDICES; the rays themselves stay where they
 * are, results land at the rays' own indices).  Rays that start close together walk the same part of the tree,
 * so node and triangle fetches hit in L2 instead of HBM: +24 % on the 10M-triangle soup (0.7 GB of BVH, L2 hit
 * rate 32 %), nothing on a BVH that already lives in L2.  Default (neither flag): on when the BVH is larger
 * than 96 MB and the call has at least 2^20 rays. */
#define PRT_TRACE_BIN 8u
#define PRT_TRACE_NO_BIN 16u

/* BVH build options */
typedef struct {
    uint32_t max_leaf_tris; /* 1..7, default 4 */
    float cost_node;        /* SAH traversal cost, default 1.0 */
    float cost_tri;         /* SAH intersection cost, default 2.0 */
    uint32_t rotations;     /* number of bottom-up SAH rotation passes during refit (0 = none, default 1).
                               With treelets and <= 1 pass the rebuilt subtrees are refitted by the warp that
                               built them and rotations apply above them only; > 1 re-derives every pass from
                               the leaves */
    uint32_t treelets;      /* 1 (default): every maximal subtree of <= 128 triangles of the Morton
                               hierarchy is rebuilt with SAH (exact for <= 6 triangles, binned above) before
                               refit; 0 = plain LBVH */
    uint32_t morton_bits;   /* 30, 63, or 0 (default) = 30 unless more than 1/16 of the sorted neighbours
                               share a 30-bit cell (clustered geometry), then 63 (21 bits per axis) */
} prt_bvh_options;

int prt_abi_version(void);

/* lifetime.  Replaces: ti.init(arch=ti.gpu) main_taichi.py:12 + World() :41 */
int prt_create(int device, prt_ctx** out);
void prt_destroy(prt_ctx* ctx);
/* text of the last error on this context (ctx == NULL: last prt_create error) */
const char* prt_last_error(const prt_ctx* ctx);

/* scene upload (host pointers).  Replaces Scene.add_primitive core/scene.py:31-46
 * + World.add/commit mathematics/intersection_taichi.py:220-233.
 *   verts      [nt][3][3] f32, global triangle-id order
 *   normals    [nt][3] f32 geometric normals with the reference's sign convention
 *              (mathematics/shapes.py:43-47,172-176); NULL = +normalize(e1 x e2)
 *   tri_material [nt] index into mats; NULL = all 0
 *   light_tris [nl] triangle ids that NEE samples (Scene.lights) */
int prt_scene_set_triangles(prt_ctx* ctx, const float* verts_host, const float* normals_host,
                            uint32_t nt, const uint32_t* tri_material_host,
                            const prt_material* mats_host, uint32_t nm,
                            const uint32_t* light_tris_host, uint32_t nl);
/* same, geometry already on the device (large soups); one default Lambert material */
int prt_scene_set_triangles_dev(prt_ctx* ctx, const float* verts_dev, uint32_t nt, void* stream);

/* GPU LBVH build.  Replaces BVH.build accelerators/bvh.py:192-215 and
 * Aggregator.push/update accelerators/aggregator.py:25-72 and
 * accelerators/bvh_taichi.py:111-161.  Synchronous.  opts/stats may be NULL. */
int prt_bvh_build(prt_ctx* ctx, const prt_bvh_options* opts, prt_bvh_stats* stats);

/* camera.  Replaces Camera.__init__/convert_to_taichi_camera core/camera.py:14-36 */
int prt_camera_set(prt_ctx* ctx, const prt_camera* cam);

/* primary rays for samples [s0,s1) of every pixel, written as
 * rays_dev[(pixel*(s1-s0) + s-s0)]; jitter=0 -> pixel centres.
 * Replaces Camera.generate_ray core/camera.py:41-72 under main.py:31-33. */
int prt_generate_rays(prt_ctx* ctx, uint64_t seed, uint32_t s0, uint32_t s1, int jitter,
                      float tmin, float tmax, prt_ray* rays_dev, void* stream);

/* closest hit.  Replaces Scene.hit/hit_faster core/scene.py:59-73,
 * BVH.hit accelerators/bvh.py:218-237, World.hit_all
 * mathematics/intersection_taichi.py:238-291.
 * Accept rule = the reference's numba kernel (mathematics/intersection.py:42-65): tmin <= t <= bound,
 * u >= 0, v >= 0, u + v <= 1, all inclusive, no back-face culling; closest = min t, lowest global id on
 * exact ties (:106-116, core/scene.py:66-73).  The strict f32 rule of the Taichi draft
 * (mathematics/intersection_taichi.py:69-91: t0 < t < t1, 0 <= u <= 1) is NOT offered: the numba code is
 * the only intersection code of the reference that runs, and the oracle is pinned to it. */
int prt_trace_closest(prt_ctx* ctx, const prt_ray* rays_dev, uint64_t n, prt_hit* hits_dev,
                      uint32_t flags, void* stream);
/* any hit in [tmin,tmax] (shadow rays, core/tracing.py:101-102): occluded_dev[i] = 0/1 */
int prt_trace_any(prt_ctx* ctx, const prt_ray* rays_dev, uint64_t n, uint8_t* occluded_dev,
                  uint32_t flags, void* stream);
/* full hit set per ray: count and order-free checksum sum((id+1)*0x9E3779B97F4A7C15) */
int prt_trace_all(prt_ctx* ctx, const prt_ray* rays_dev, uint64_t n, uint32_t* counts_dev,
                  uint64_t* sums_dev, uint32_t flags, void* stream);
/* closest hit with HOST buffers (copies in and out inside the call; synchronous) */
int prt_trace_closest_host(prt_ctx* ctx, const prt_ray* rays_host, uint64_t n,
                           prt_hit* hits_host, uint32_t flags);

/* render.  Replaces the pixel x spp loop of main.py:28-55 around
 * path_tracing(), and render() main_taichi.py:80-99 + PathTracer.trace
 * core/tracing.py:116-155.  ADDS samples [spp_begin, spp_end) of every pixel to
 * accum_dev [h][w][4] f32 (r,g,b sums, sample count); row 0 is v = 0 (bottom).
 * prim_ids_dev (optional) [h][w][spp_end-spp_begin] i32 primary-hit triangle ids. */
int prt_render(prt_ctx* ctx, const prt_render_params* params, float* accum_dev,
               int32_t* prim_ids_dev, void* stream);
/* path tracing from caller-supplied rays: replaces `e, r = path_tracing(ray, a_scene)` under the
 * reference's own ray generation (main.py:21-23,33-35).  For every ray i, samples
 * [spp_begin, spp_end) are traced (Philox stream "pixel" = i) and ADDED to radiance_dev[i] =
 * (r, g, b sums, sample count).  prim_ids_dev optional [n][spp_end-spp_begin]. */
int prt_trace_paths(prt_ctx* ctx, const prt_ray* rays_dev, uint64_t n, const prt_render_params* params,
                    float* radiance_dev, int32_t* prim_ids_dev, void* stream);
/* Path-segment log (debugging aid).  Replaces debug/ray_logger.py RayLogger.add / add_line as
 * used by main.py:66-85 (`path_tracing(ray, a_scene, ray_logger)`).  While a log is set, every
 * prt_render / prt_trace_paths call appends one record per traced path segment:
 *   p0 = ray origin, p1 = hit point (or origin + 5 * direction for a miss, RayLogger.add's default
 *   t = 5), kind = bounce index for path segments, -1 for an unoccluded light connection,
 *   path = pixel (or ray index) * samples_in_call + sample.
 * segments_dev: caller-owned device array of `capacity` records; count_dev: device counter the
 * caller zeroes (it keeps counting past capacity: count > capacity means the log is truncated).
 * NULL segments_dev switches logging off. */
typedef struct {
    float p0[3];
    int32_t kind;
    float p1[3];
    uint32_t path;
} prt_segment; /* 32 bytes */
int prt_set_path_log(prt_ctx* ctx, prt_segment* segments_dev, uint64_t capacity, uint32_t* count_dev);
/* same with a HOST accumulation buffer (upload, render, download; synchronous) */
int prt_render_host(prt_ctx* ctx, const prt_render_params* params, float* accum_host);
/* Known-answer hook for the specular BSDF device functions that the shade kernel calls
 * (core/bsdf_taichi.py:6-22 reflectance / reflect / refract, :45-86 Metal.scatter / Dielectric.scatter,
 * mathematics/vec3_taichi.py:33-39 random_in_unit_sphere): for every query, the un-normalised
 * scattered direction and a validity flag, wi_valid_dev[i] = (wi.x, wi.y, wi.z, 1 or 0).
 * d = incoming direction (any length), ns = shading normal on the incoming side, type = PRT_MAT_MIRROR /
 * _DIELECTRIC / _CONDUCTOR, u = uniforms in the order the reference draws them. */
typedef struct {
    float d[3];
    uint32_t type;
    float ns[3];
    uint32_t front; /* dielectric: 1 = entering (ratio 1/ior) */
    float ior, roughness;
    float u[3];
    uint32_t pad[3];
} prt_bsdf_query; /* 64 bytes */
int prt_eval_specular(prt_ctx* ctx, const prt_bsdf_query* queries_dev, uint64_t n, float* wi_valid_dev, void* stream);

/* Device memory is grow-only: scene, BVH, build scratch, staging and wavefront buffers are kept and
 * reused by later calls (a rebuild or a new scene of the same size allocates nothing).  This frees
 * everything that is scratch (build arena, host-call staging, exact-mode flag lists, wavefront
 * state); the scene and its BVH stay usable. */
int prt_release_scratch(prt_ctx* ctx);
/* paths per wavefront wave (default 64 Mi = 8.5 GiB of path state, allocated only as far as a render needs it:
 * width * height * min(spp, wave / pixels) paths); 0 keeps the current value */
int prt_set_wave_paths(prt_ctx* ctx, uint64_t paths);

/* ---- multi-GPU: sample-sharded render over a replicated scene + BVH (one process per GPU) ----
 * The reference renders on one device (main.py:28-55, main_taichi.py:80-99); sharding by sample
 * index is this library's extension (BASELINE.json north_star): rank r of G traces samples
 * [r*S/G, (r+1)*S/G) of every pixel with its own Philox streams, and ONE all-reduce per frame sums
 * the fp32 accumulation buffers over NVLink.  NCCL is bound at run time (dlopen of libnccl.so.2 --
 * inside a PyTorch process that is the copy torch already loaded).
 *   prt_comm_unique_id   rank 0 makes the 128-byte NCCL id; the caller hands it to every rank
 *                        (MPI_Bcast, torch.distributed.broadcast, a file ...)
 *   prt_comm_init        collective: every rank calls it with the same id
 *   prt_comm_attach      instead: use an ncclComm_t the caller already owns (not destroyed here)
 *   prt_allreduce_sum    in-place fp32 sum over the ranks, stream-ordered (no-op for one rank)
 *   prt_render_sharded   one frame: this rank's shard of [spp_begin, spp_end) into a library-owned
 *                        zeroed buffer, one all-reduce, the sum ADDED to accum_dev (every rank ends
 *                        with the same buffer) */
#define PRT_COMM_ID_BYTES 128
int prt_comm_unique_id(void* id_out);
int prt_comm_init(prt_ctx* ctx, const void* id, int world, int rank);
int prt_comm_attach(prt_ctx* ctx, void* nccl_comm, int world, int rank);
int prt_comm_destroy(prt_ctx* ctx);
int prt_comm_info(const prt_ctx* ctx, int* world, int* rank, int* nccl_version);
int prt_allreduce_sum(prt_ctx* ctx, float* buf_dev, uint64_t n_floats, void* stream);
int prt_render_sharded(prt_ctx* ctx, const prt_render_params* params, float* accum_dev, void* stream);

/* ---- device time per kernel class (measurement aid; bench.py's roofline lines use it) ----
 * Between prt_profile_begin and prt_profile_end every kernel launch of this context is bracketed
 * by a cudaEvent pair on the launching stream; _end synchronises the device and returns the sums. */
enum { PRT_PROF_RAYGEN = 0, PRT_PROF_CLOSEST = 1, PRT_PROF_SHADE = 2, PRT_PROF_SHADOW = 3,
       PRT_PROF_EXACT_FIXUP = 4 /* finalize + FP64 replay of PRT_TRACE_EXACT */, PRT_PROF_OTHER = 5,
       PRT_PROF_ALLREDUCE = 6, PRT_PROF_CLASSES = 8 };
typedef struct {
    float ms[PRT_PROF_CLASSES];
    uint32_t launches[PRT_PROF_CLASSES];
} prt_kernel_times;
int prt_profile_begin(prt_ctx* ctx);
int prt_profile_end(prt_ctx* ctx, prt_kernel_times* out); /* synchronous */

int prt_get_counters(prt_ctx* ctx, prt_counters* out); /* synchronous */
int prt_reset_counters(prt_ctx* ctx);
int prt_synchronize(prt_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* PRT_H */
