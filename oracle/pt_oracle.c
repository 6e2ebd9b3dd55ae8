/*
 * pt_oracle.c -- CPU parity oracle for the path-tracing hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under pyrenderer_b200/ may import, link
 * or execute this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, as the checker / CPU baseline.
 *
 * It is a plain-C, double-precision restatement of the reference's CPU
 * algorithm (sontung/pyrenderer, all paths relative to /root/reference):
 *
 *   closest hit     mathematics/intersection.py:42-82 (grouped numba kernel,
 *                   op order of mathematics/fast_op.py:60-63,74-92), closest
 *                   selection intersection.py:106-116 + core/scene.py:66-73
 *   slab test       mathematics/bbox.py:6-26
 *   camera          core/camera.py:41-72, mathematics/vec3.py:5-23
 *   samplers/frame  mathematics/samplers_debug.py:8-88
 *   light point     mathematics/shapes2.py:72-79
 *   integrator      core/tracing.py:92-155 driven by main.py:28-37
 *   BSDF extras     core/bsdf_taichi.py:6-86
 *
 * Parity pin: tests/golden/ npz files were produced by tests/golden/make_golden.py,
 * which imports the reference's own modules from /root/reference; the tests
 * in tests/test_oracle_golden.py check every function here against them
 * (closest-hit ids and t bit-exact vs the numba kernels, slab, sampler, frame,
 * camera, transforms, the 9-bounce golden path of test.py:38-57).
 * RADIANCE is "parity unpinned": the reference holds no rendered image or
 * radiance value of its own estimator; orc_render restates
 * core/tracing.py:116-155 and is pinned only by the
 * deterministic known answer "a primary ray that hits the light returns
 * light_color" and by analytic properties.
 *
 * Compile: gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC (see Makefile).
 * -ffp-contract=off matters: numba does not fuse multiply-add, and the
 * bit-exact t comparison depends on it.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* mathematics/constants.py:9-16 */
#define ORC_EPS 1.1754943508222875e-38 /* np.finfo(float32).tiny */
#define ORC_MAX_F 3.4028234663852886e+38
#define ORC_MACHINE_EPS (1.1920928955078125e-07 * 0.5)
#define ORC_GAMMA2_3 ((3 * ORC_MACHINE_EPS) / (1 - 3 * ORC_MACHINE_EPS))
#define ORC_PI 3.14159265358979323846
#define ORC_INV_PI 0.31830988618379067154
#define ORC_PI_OVER2 1.57079632679489661923
#define ORC_PI_OVER4 0.78539816339744830961

/* ---- shared record layouts (identical to include/prt.h) ---------------- */
typedef struct {
    float albedo[3];
    uint32_t type; /* 0 lambert, 1 emitter("null"), 2 mirror, 3 dielectric, 4 conductor */
    float ior;
    float roughness;
    uint32_t two_sided; /* reference "sided == 0" => 1 here */
    uint32_t pad;
    float emission[3]; /* Tungsten primitive "emission" (scene.json:234-238); physical mode only */
    uint32_t pad2;
} orc_material;

typedef struct {
    double iview[16]; /* row-major, row-vector convention (core/camera.py:18-19) */
    double sensor_w;  /* tan(radians(fov)/2)*focal*aspect  (camera.py:48-50) */
    double sensor_h;
    double focal;
    uint32_t width, height;
    double aperture; /* camera.py:63-65; 0 = pinhole */
} orc_camera;

typedef struct {
    uint64_t seed;
    uint32_t spp_begin, spp_end;
    uint32_t max_depth;
    uint32_t rr_start; /* first bounce index at which Russian roulette applies; 0xffffffff = off */
    float light_color[3]; /* core/tracing.py:120 */
    float tmin, tmax;     /* core/tracing.py:127 : 1e-5, 99999.9 */
    uint32_t flags; /* bit 1 (2u) = physically-based estimator, see trace_path_physical */
} orc_render_params;
#define ORC_RENDER_PHYSICAL 2u

/* ---- Philox4x32-10 (Salmon et al. 2011, Random123) --------------------- */
static inline void philox_round(uint32_t c[4], const uint32_t k[2]) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    uint32_t n0 = hi1 ^ c[1] ^ k[0];
    uint32_t n1 = lo1;
    uint32_t n2 = hi0 ^ c[3] ^ k[1];
    uint32_t n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
    uint32_t k[2] = {key[0], key[1]};
    for (int r = 0; r < 10; ++r) {
        philox_round(c, k);
        k[0] += 0x9E3779B9u;
        k[1] += 0xBB67AE85u;
    }
    out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}

/* stream layout shared with the CUDA path (DESIGN.md "RNG streams"):
 * counter = (pixel, sample, bounce, block), key = (seed lo, seed hi).
 * block 0: camera jitter (bounce 0);  block 1: bsdf u1,u2, light tri, rr;
 * block 2: light point u,v, spare, spare. */
static inline void rng4(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t bounce,
                        uint32_t block, uint32_t out[4]) {
    uint32_t c[4] = {pixel, sample, bounce, block};
    uint32_t k[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    orc_philox4x32_10(c, k, out);
}
/* 24-bit uniform in [0,1): exactly representable in f32 and f64 */
static inline double u24(uint32_t k) { return (double)(k >> 8) * (1.0 / 16777216.0); }
static inline uint32_t rand_index(uint32_t k, uint32_t n) {
    return (uint32_t)(((uint64_t)k * (uint64_t)n) >> 32);
}

/* ---- fast_op.py restatements ------------------------------------------- */
static inline void sub3(const double* x, const double* y, double* s) { /* fast_op.py:74-77 */
    s[0] = x[0] - y[0]; s[1] = x[1] - y[1]; s[2] = x[2] - y[2];
}
static inline void cross3(const double* x, const double* y, double* r) { /* fast_op.py:85-92 */
    r[0] = x[1] * y[2] - x[2] * y[1];
    r[1] = x[2] * y[0] - x[0] * y[2];
    r[2] = x[0] * y[1] - x[1] * y[0];
}
static inline double dot3(const double* x, const double* y) { /* fast_op.py:60-63 */
    return x[0] * y[0] + x[1] * y[1] + x[2] * y[2];
}

/* One triangle of the grouped kernel: intersection.py:68-82 computes the
 * seven arrays, intersection.py:42-65 decides.  Returns 1 on accept and
 * writes t,u,v.  `bound_hi` is ray_bound[1] (shrinks on accept in the caller),
 * `t_lo` generalises the literal EPS lower bound (intersection.py:49). */
static inline int mt_grouped(const double p0[3], const double e1[3], const double e2[3],
                             const double o[3], const double d[3], double t_lo, double bound_hi,
                             double* t_out, double* u_out, double* v_out) {
    double s[3], q[3], r[3];
    sub3(o, p0, s);
    cross3(d, e2, q);
    cross3(s, e1, r);
    double a = dot3(e1, q);
    double e2r = dot3(e2, r);
    double sq = dot3(s, q);
    double rdr = dot3(d, r);
    if (-ORC_EPS < a && a < ORC_EPS) return 0;
    double f = 1.0 / a;
    double t = f * e2r;
    if (t > bound_hi || t < t_lo) return 0;
    double u = f * sq;
    if (u < 0.0) return 0;
    double v = f * rdr;
    if (v < 0.0 || u + v > 1.0) return 0;
    *t_out = t; *u_out = u; *v_out = v;
    return 1;
}

/* scalar formulation, intersection.py:7-39 (np.cross / np.dot; accept order
 * u, t, v).  Used only to cross-check decisions against mt_grouped. */
int orc_mt_scalar(const double v0[3], const double v1[3], const double v2[3], const double o[3],
                  const double d[3], double bound_hi, double* t_out) {
    double e1[3], e2[3], q[3], s[3], r[3];
    sub3(v1, v0, e1);
    sub3(v2, v0, e2);
    cross3(d, e2, q);
    double a = dot3(e1, q);
    if (fabs(a) < ORC_EPS) return 0;
    double f = 1.0 / a;
    sub3(o, v0, s);
    double u = f * dot3(s, q);
    if (u < 0.0) return 0;
    cross3(s, e1, r);
    double t = f * dot3(e2, r);
    if (t > bound_hi || t < ORC_EPS) return 0;
    double v = f * dot3(d, r);
    if (v < 0.0 || u + v > 1.0) return 0;
    *t_out = t;
    return 1;
}

typedef struct {
    uint32_t nt;
    double* p0; /* [nt][3] */
    double* e1;
    double* e2;
} soup_t;

static int soup_init(soup_t* s, const float* tris, uint32_t nt) {
    s->nt = nt;
    s->p0 = (double*)malloc(sizeof(double) * 3 * (size_t)(nt ? nt : 1));
    s->e1 = (double*)malloc(sizeof(double) * 3 * (size_t)(nt ? nt : 1));
    s->e2 = (double*)malloc(sizeof(double) * 3 * (size_t)(nt ? nt : 1));
    if (!s->p0 || !s->e1 || !s->e2) return -1;
    for (uint32_t i = 0; i < nt; ++i) {
        double a[3], b[3], c[3];
        for (int k = 0; k < 3; ++k) {
            a[k] = (double)tris[(size_t)i * 9 + k];
            b[k] = (double)tris[(size_t)i * 9 + 3 + k];
            c[k] = (double)tris[(size_t)i * 9 + 6 + k];
        }
        /* shapes2.py:43-46: e1 = v1 - v0, e2 = v2 - v0 */
        for (int k = 0; k < 3; ++k) {
            s->p0[(size_t)i * 3 + k] = a[k];
            s->e1[(size_t)i * 3 + k] = b[k] - a[k];
            s->e2[(size_t)i * 3 + k] = c[k] - a[k];
        }
    }
    return 0;
}
static void soup_free(soup_t* s) { free(s->p0); free(s->e1); free(s->e2); }

/* Closest hit of one ray over the soup in global-ID order.
 * Sequential accept with shrinking bound (intersection.py:81-82,65), strict
 * "<" improvement scan (intersection.py:109; scene.py:71) => min t, lowest
 * index on exact ties.  SURVEY App. A.1: hits below t_lo never shrink the
 * bound (they are rejected inside mt_grouped by t < t_lo). */
static inline int closest_one(const soup_t* s, const double o[3], const double d[3], double t_lo,
                              double t_hi, double* t_out, double* u_out, double* v_out) {
    int best = -1;
    double bound = t_hi, tbest = ORC_MAX_F, ub = 0, vb = 0;
    for (uint32_t i = 0; i < s->nt; ++i) {
        double t, u, v;
        if (mt_grouped(s->p0 + (size_t)i * 3, s->e1 + (size_t)i * 3, s->e2 + (size_t)i * 3, o, d,
                       t_lo, bound, &t, &u, &v)) {
            bound = t;
            if (t < tbest) { tbest = t; best = (int)i; ub = u; vb = v; }
        }
    }
    *t_out = tbest; *u_out = ub; *v_out = vb;
    return best;
}

static inline int any_one(const soup_t* s, const double o[3], const double d[3], double t_lo,
                          double t_hi) {
    for (uint32_t i = 0; i < s->nt; ++i) {
        double t, u, v;
        if (mt_grouped(s->p0 + (size_t)i * 3, s->e1 + (size_t)i * 3, s->e2 + (size_t)i * 3, o, d,
                       t_lo, t_hi, &t, &u, &v))
            return 1;
    }
    return 0;
}

/* rays: f32 [n][8] = ox oy oz tmin dx dy dz tmax (the C-ABI ray record).
 * ids: -1 = miss.  ts/us/vs may be NULL. */
int orc_closest_hit(const float* tris, uint32_t nt, const float* rays, uint64_t n, int32_t* ids,
                    double* ts, double* us, double* vs, int nthreads) {
    soup_t s;
    if (soup_init(&s, tris, nt)) return -1;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        const float* r = rays + (size_t)i * 8;
        double o[3] = {r[0], r[1], r[2]}, d[3] = {r[4], r[5], r[6]};
        double t, u, v;
        int id = closest_one(&s, o, d, (double)r[3], (double)r[7], &t, &u, &v);
        ids[i] = id;
        if (ts) ts[i] = id >= 0 ? t : 0.0;
        if (us) us[i] = id >= 0 ? u : 0.0;
        if (vs) vs[i] = id >= 0 ? v : 0.0;
    }
    soup_free(&s);
    return 0;
}

int orc_any_hit(const float* tris, uint32_t nt, const float* rays, uint64_t n, uint8_t* occluded,
                int nthreads) {
    soup_t s;
    if (soup_init(&s, tris, nt)) return -1;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        const float* r = rays + (size_t)i * 8;
        double o[3] = {r[0], r[1], r[2]}, d[3] = {r[4], r[5], r[6]};
        occluded[i] = (uint8_t)any_one(&s, o, d, (double)r[3], (double)r[7]);
    }
    soup_free(&s);
    return 0;
}

/* Full hit set of each ray (no bound shrinking): count and an order-free
 * 64-bit checksum sum((id+1)*0x9E3779B97F4A7C15).  Used for the "BVH
 * traversal hit sets" parity check. */
int orc_all_hits(const float* tris, uint32_t nt, const float* rays, uint64_t n, uint32_t* counts,
                 uint64_t* sums, int nthreads) {
    soup_t s;
    if (soup_init(&s, tris, nt)) return -1;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        const float* r = rays + (size_t)i * 8;
        double o[3] = {r[0], r[1], r[2]}, d[3] = {r[4], r[5], r[6]};
        uint32_t c = 0;
        uint64_t h = 0;
        for (uint32_t k = 0; k < s.nt; ++k) {
            double t, u, v;
            if (mt_grouped(s.p0 + (size_t)k * 3, s.e1 + (size_t)k * 3, s.e2 + (size_t)k * 3, o, d,
                           (double)r[3], (double)r[7], &t, &u, &v)) {
                ++c;
                h += (uint64_t)(k + 1) * 0x9E3779B97F4A7C15ull;
            }
        }
        counts[i] = c;
        sums[i] = h;
    }
    soup_free(&s);
    return 0;
}

/* mathematics/bbox.py:6-26.  Returns 1 on hit and writes t0. */
int orc_slab(double t0, double t1, const double pos[3], const double inv_dir[3],
             const double bmin[3], const double bmax[3], double* t0_out) {
    for (int i = 0; i < 3; ++i) {
        double inv = inv_dir[i];
        double t_near = (bmin[i] - pos[i]) * inv;
        double t_far = (bmax[i] - pos[i]) * inv;
        if (t_near > t_far) { double tmp = t_near; t_near = t_far; t_far = tmp; }
        t_far *= 1 + 2 * ORC_GAMMA2_3;
        if (t_near > t0) t0 = t_near;
        if (t_far < t1) t1 = t_far;
        if (t0 > t1) return 0;
    }
    *t0_out = t0;
    return 1;
}

/* ---- samplers (mathematics/samplers_debug.py) -------------------------- */
static inline double norm3(const double* v) { /* vec3.py:5-10 */
    return sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
}
static inline void normalize3(double* v) { /* vec3.py:13-17 */
    double n = norm3(v);
    v[0] /= n; v[1] /= n; v[2] /= n;
}

void orc_concentric_sample_disk(double u1, double u2, double out[2]) { /* samplers_debug.py:8-20 */
    double ox = 2.0 * u1 - 1.0, oy = 2.0 * u2 - 1.0;
    if (ox == 0 && oy == 0) { out[0] = 0.0; out[1] = 0.0; return; }
    double r, theta;
    if (fabs(ox) > fabs(oy)) {
        r = ox;
        theta = ORC_PI_OVER4 * (oy / ox);
    } else {
        r = oy;
        theta = ORC_PI_OVER2 - ORC_PI_OVER4 * (ox / oy);
    }
    out[0] = r * cos(theta);
    out[1] = r * sin(theta);
}

/* samplers_debug.py:44-80: frame rows (x, z, n) after rotate_z_to */
void orc_frame_z_to(const double n_in[3], double r1[3], double r2[3], double r3[3]) {
    double v[3] = {n_in[0], n_in[1], n_in[2]};
    normalize3(v);
    if (fabs(v[1] - 1.0) < ORC_EPS) {
        r1[0] = 1; r1[1] = 0; r1[2] = 0;
        r3[0] = 0; r3[1] = 1; r3[2] = 0; /* res2 */
        r2[0] = 0; r2[1] = 0; r2[2] = 1; /* res3 */
    } else if (fabs(v[1] + 1.0) < ORC_EPS) {
        r1[0] = 1; r1[1] = 0; r1[2] = 0;
        r3[0] = 0; r3[1] = -1; r3[2] = 0;
        r2[0] = 0; r2[1] = 0; r2[2] = 1;
    } else {
        double Y[3] = {0.0, 1.0, 0.0}, x[3], z[3];
        cross3(v, Y, x);
        normalize3(x);
        cross3(x, v, z);
        normalize3(z);
        memcpy(r1, x, sizeof x);
        memcpy(r2, z, sizeof z);
        memcpy(r3, v, sizeof v);
    }
}

/* samplers_debug.py:31-37 + 83-88 with explicit uniforms */
void orc_cosine_sample_hemisphere(const double n[3], double u1, double u2, double out[3]) {
    double d[2];
    orc_concentric_sample_disk(u1, u2, d);
    double zz = 1 - d[0] * d[0] - d[1] * d[1];
    double z = sqrt(zz > 0.0 ? zz : 0.0);
    double r1[3], r2[3], r3[3];
    orc_frame_z_to(n, r1, r2, r3);
    for (int k = 0; k < 3; ++k) out[k] = d[0] * r1[k] + d[1] * r2[k] + z * r3[k];
    normalize3(out);
}

/* ---- camera (core/camera.py:41-72) ------------------------------------- */
/* (lu, lv) in [0,1)^2: the two random() draws of camera.py:64-65 (only read when aperture > 0) */
void orc_generate_ray_lens(const orc_camera* cam, double u, double v, double lu, double lv, double o_out[3],
                           double d_out[3]) {
    double cs0 = u - 0.5, cs1 = v - 0.5;
    double rd[3] = {cs0 * cam->sensor_w / 0.5, cs1 * cam->sensor_h / 0.5, -cam->focal};
    /* to_homogeneous_vector rounds to float32 (vec3.py:20-23) */
    double h[4] = {(double)(float)rd[0], (double)(float)rd[1], (double)(float)rd[2], 1.0};
    const double* m = cam->iview;
    double dw[4], ow[4];
    double ax = 0.0, ay = 0.0;
    if (cam->aperture > 0.0) { /* camera.py:63-65, then to_homogeneous_vector's f32 rounding */
        ax = (double)(float)(cam->aperture * lu - cam->aperture / 2.0);
        ay = (double)(float)(cam->aperture * lv - cam->aperture / 2.0);
    }
    for (int j = 0; j < 4; ++j) {
        dw[j] = ((h[0] * m[0 * 4 + j] + h[1] * m[1 * 4 + j]) + h[2] * m[2 * 4 + j]) + h[3] * m[3 * 4 + j];
        ow[j] = cam->aperture > 0.0 ? (ax * m[0 * 4 + j] + ay * m[1 * 4 + j]) + 1.0 * m[3 * 4 + j]
                                    : ((0.0 * m[0 * 4 + j] + 0.0 * m[1 * 4 + j]) + 0.0 * m[2 * 4 + j]) + 1.0 * m[3 * 4 + j];
    }
    double f[3] = {dw[0] - ow[0], dw[1] - ow[1], dw[2] - ow[2]};
    normalize3(f);
    o_out[0] = ow[0]; o_out[1] = ow[1]; o_out[2] = ow[2];
    d_out[0] = f[0]; d_out[1] = f[1]; d_out[2] = f[2];
}
void orc_generate_ray(const orc_camera* cam, double u, double v, double o_out[3], double d_out[3]) {
    orc_generate_ray_lens(cam, u, v, 0.5, 0.5, o_out, d_out); /* lens centre: offset exactly 0 */
}

/* Primary rays for pixel (i,j) samples [s0,s1): main.py:31-33.  Output in the
 * f32 interchange record; jitter==0 uses the pixel centre (config C2). */
int orc_generate_rays(const orc_camera* cam, uint64_t seed, uint32_t s0, uint32_t s1, int jitter,
                      float tmin, float tmax, float* rays) {
    uint32_t W = cam->width, H = cam->height, ns = s1 - s0;
#pragma omp parallel for schedule(static)
    for (int64_t p = 0; p < (int64_t)W * H; ++p) {
        uint32_t i = (uint32_t)(p % W), j = (uint32_t)(p / W);
        for (uint32_t s = s0; s < s1; ++s) {
            double jx = 0.5, jy = 0.5, lu = 0.5, lv = 0.5;
            if (jitter) {
                uint32_t r[4];
                rng4(seed, (uint32_t)p, s, 0, 0, r);
                jx = u24(r[0]); jy = u24(r[1]); lu = u24(r[2]); lv = u24(r[3]);
            }
            double u = ((double)i + jx) / (double)W, v = ((double)j + jy) / (double)H;
            double o[3], d[3];
            orc_generate_ray_lens(cam, u, v, lu, lv, o, d);
            float* out = rays + ((size_t)p * ns + (s - s0)) * 8;
            out[0] = (float)o[0]; out[1] = (float)o[1]; out[2] = (float)o[2]; out[3] = tmin;
            out[4] = (float)d[0]; out[5] = (float)d[1]; out[6] = (float)d[2]; out[7] = tmax;
        }
    }
    return 0;
}

/* ---- BSDF extras (core/bsdf_taichi.py) --------------------------------- */
static inline double schlick(double cosine, double idx) { /* bsdf_taichi.py:6-9 */
    double r0 = (1.0 - idx) / (1.0 + idx);
    r0 = r0 * r0;
    double m = 1.0 - cosine;
    return r0 + (1.0 - r0) * (m * m * m * m * m);
}
static inline void reflect3(const double* v, const double* n, double* out) { /* :12-14 */
    double k = 2.0 * dot3(v, n);
    out[0] = v[0] - k * n[0]; out[1] = v[1] - k * n[1]; out[2] = v[2] - k * n[2];
}
static inline void refract3(const double* v, const double* n, double eta, double* out) { /* :17-22 */
    double c = -dot3(v, n);
    if (c > 1.0) c = 1.0;
    double perp[3] = {eta * (v[0] + c * n[0]), eta * (v[1] + c * n[1]), eta * (v[2] + c * n[2])};
    double k = -sqrt(fabs(1.0 - dot3(perp, perp)));
    out[0] = perp[0] + k * n[0]; out[1] = perp[1] + k * n[1]; out[2] = perp[2] + k * n[2];
}
/* vec3_taichi.py:33-39 with explicit uniforms */
static inline void in_unit_sphere(double ua, double ub, double uc, double* out) {
    double theta = ua * ORC_PI * 2.0;
    double phi = acos(2.0 * ub - 1.0);
    double r = cbrt(uc);
    out[0] = r * sin(phi) * cos(theta);
    out[1] = r * sin(phi) * sin(theta);
    out[2] = r * cos(phi);
}

/* Specular scattering, semantics of core/bsdf_taichi.py:45-86.  `d` = incoming direction (any
 * length), `ns` = shading normal facing the incoming side.  type 2 = mirror (Metal, roughness 0),
 * 4 = conductor (Metal.scatter :54-60: reflect(unit d, n) + roughness * random_in_unit_sphere(),
 * valid only if it leaves on the normal's side), 3 = dielectric (Dielectric.scatter :71-86: ratio =
 * 1/ior on the front side, else ior; total internal reflection or Schlick > u1 reflects, else
 * refracts).  (u1, u2, u3) = the uniforms in the order the reference draws them (conductor: theta,
 * v, r of the sphere sample; dielectric: the Fresnel draw).  wi_out is NOT normalised (the
 * reference returns out_direction as is); returns 1 if the sample is valid. */
int orc_scatter_specular(uint32_t type, const double d[3], const double ns[3], int front, double ior,
                         double roughness, double u1, double u2, double u3, double wi_out[3]) {
    double ud[3] = {d[0], d[1], d[2]};
    normalize3(ud);
    if (type == 2) {
        reflect3(ud, ns, wi_out);
        return 1;
    }
    if (type == 4) {
        double f[3];
        reflect3(ud, ns, wi_out);
        in_unit_sphere(u1, u2, u3, f);
        for (int k = 0; k < 3; ++k) wi_out[k] += roughness * f[k];
        return dot3(wi_out, ns) > 0.0; /* bsdf_taichi.py:58 */
    }
    double ratio = front ? 1.0 / ior : ior;
    double ct = -dot3(ud, ns);
    if (ct > 1.0) ct = 1.0;
    double st = sqrt(1.0 - ct * ct);
    if (ratio * st > 1.0 || schlick(ct, ratio) > u1) reflect3(ud, ns, wi_out);
    else refract3(ud, ns, ratio, wi_out);
    return 1;
}
/* exported for the known-answer tests against the reference's own functions (tests/golden/radiance_golden.npz) */
double orc_schlick(double cosine, double idx) { return schlick(cosine, idx); }
void orc_reflect(const double v[3], const double n[3], double out[3]) { reflect3(v, n, out); }
void orc_refract(const double v[3], const double n[3], double eta, double out[3]) { refract3(v, n, eta, out); }
void orc_in_unit_sphere(double ua, double ub, double uc, double out[3]) { in_unit_sphere(ua, ub, uc, out); }

/* ---- integrator (core/tracing.py:116-155) ------------- */
typedef struct {
    soup_t soup;
    const float* normals; /* [nt][3] geometric normals, reference sign convention */
    const uint32_t* tri_mat;
    const orc_material* mats;
    const uint32_t* light_tris;
    uint32_t nl;
    const float* tris;
} scene_t;

/* one path; returns radiance in L[3]; counts rays.  prim_id = primary hit. */
static void trace_path(const scene_t* sc, const orc_render_params* P, uint32_t pixel,
                       uint32_t sample, double o[3], double d[3], double L[3], int32_t* prim_id,
                       uint64_t* n_closest, uint64_t* n_shadow) {
    double beta[3] = {1.0, 1.0, 1.0};
    L[0] = L[1] = L[2] = 0.0;
    *prim_id = -1;
    for (uint32_t bounce = 0; bounce < P->max_depth; ++bounce) {
        double t, bu, bv;
        int id = closest_one(&sc->soup, o, d, (double)P->tmin, (double)P->tmax, &t, &bu, &bv);
        ++*n_closest;
        if (bounce == 0) *prim_id = id;
        if (id < 0) break; /* tracing.py:141-142 */
        const orc_material* m = &sc->mats[sc->tri_mat[id]];
        double n[3] = {sc->normals[id * 3], sc->normals[id * 3 + 1], sc->normals[id * 3 + 2]};
        double nd[3] = {-d[0], -d[1], -d[2]};
        if (m->type == 1) { /* tracing.py:129-139 */
            double d1 = dot3(nd, n);
            if (d1 > 0.0) {
                double w = bounce == 0 ? 1.0 : d1;
                for (int k = 0; k < 3; ++k) L[k] += (double)P->light_color[k] * beta[k] * w;
            }
            break;
        }
        /* hit position: fast_op.py:110-112 compute_pos = o + d*t */
        double p[3] = {o[0] + d[0] * t, o[1] + d[1] * t, o[2] + d[2] * t};
        int front = dot3(n, nd) >= 0.0;
        /* two-sided flip: shapes2.py:93-96 / shapes.py:99-102 */
        if (m->two_sided && !front) { n[0] = -n[0]; n[1] = -n[1]; n[2] = -n[2]; }

        uint32_t r1[4];
        rng4(P->seed, pixel, sample, bounce, 1, r1);
        double wi[3];
        int do_nee = 0;
        if (m->type == 0) {
            orc_cosine_sample_hemisphere(n, u24(r1[0]), u24(r1[1]), wi);
            /* tracing.py:145-149: beta *= albedo*max(0,n.wi)/pdf/pi, pdf=|n.wi|/pi
             * (shapes.py:108); NaN guard -> pdf = 1e-4 */
            double c = dot3(n, wi);
            double pdf = fabs(c) * ORC_INV_PI;
            double cz = c > 0.0 ? c : 0.0;
            for (int k = 0; k < 3; ++k) {
                double nb = (double)m->albedo[k] * cz / pdf * ORC_INV_PI;
                if (isnan(nb)) nb = (double)m->albedo[k] * cz / 1e-4 * ORC_INV_PI;
                beta[k] *= nb;
            }
            do_nee = 1;
        } else {
            /* specular family, semantics of core/bsdf_taichi.py:45-86; the
             * shading normal faces the incoming side */
            double ns[3] = {n[0], n[1], n[2]};
            if (!front && !m->two_sided) { ns[0] = -n[0]; ns[1] = -n[1]; ns[2] = -n[2]; }
            uint32_t r2[4];
            rng4(P->seed, pixel, sample, bounce, 2, r2);
            if (!orc_scatter_specular(m->type, d, ns, front, (double)m->ior, (double)m->roughness, u24(r1[0]),
                                      u24(r1[1]), u24(r2[2]), wi))
                break;
            normalize3(wi);
            for (int k = 0; k < 3; ++k) beta[k] *= (double)m->albedo[k];
        }

        if (do_nee && sc->nl > 0) { /* tracing.py:92-108, shapes.py:62-71 */
            uint32_t r2[4];
            rng4(P->seed, pixel, sample, bounce, 2, r2);
            uint32_t lt = sc->light_tris[rand_index(r1[2], sc->nl)];
            double su = sqrt(u24(r2[0])), sv = u24(r2[1]);
            double a = su * (1 - sv), b = su * sv, c = 1.0 - a - b;
            const float* tv = sc->tris + (size_t)lt * 9;
            double p2[3];
            for (int k = 0; k < 3; ++k)
                p2[k] = a * (double)tv[k] + b * (double)tv[3 + k] + c * (double)tv[6 + k];
            double n2[3] = {sc->normals[lt * 3], sc->normals[lt * 3 + 1], sc->normals[lt * 3 + 2]};
            double w[3] = {p2[0] - p[0], p2[1] - p[1], p2[2] - p[2]};
            double dist2 = dot3(w, w); /* sqrLength(p - p2) */
            double dist = sqrt(dist2);
            double w2[3] = {-w[0] / dist, -w[1] / dist, -w[2] / dist};
            w[0] /= dist; w[1] /= dist; w[2] /= dist;
            /* shadow t_max = |p2-p|*(1-1e-4) */
            double tl = dist * (1.0 - 1e-4);
            ++*n_shadow;
            if (!any_one(&sc->soup, p, w, (double)P->tmin, tl)) {
                double dot1 = dot3(n, w), dot2 = dot3(n2, w2);
                if (dot1 > 0.0 && dot2 > 0.0) {
                    /* emissive = BSDFLight.evaluate() = albedo of the light bsdf (bsdf.py:52-53) */
                    const orc_material* lm = &sc->mats[sc->tri_mat[lt]];
                    for (int k = 0; k < 3; ++k)
                        L[k] += beta[k] * (double)lm->albedo[k] * dot1 * dot2 / dist2;
                }
            }
        }

        /* Russian roulette (not in the reference; unbiased extension, off when
         * rr_start == 0xffffffff): survive with q = min(1, max(beta)) */
        if (bounce >= P->rr_start) {
            double q = beta[0] > beta[1] ? beta[0] : beta[1];
            if (beta[2] > q) q = beta[2];
            if (q < 1.0) {
                if (!(u24(r1[3]) < q)) break;
                beta[0] /= q; beta[1] /= q; beta[2] /= q;
            }
        }
        o[0] = p[0]; o[1] = p[1]; o[2] = p[2]; /* tracing.py:153-154, no offset */
        d[0] = wi[0]; d[1] = wi[1]; d[2] = wi[2];
    }
}

/* Physically-based estimator -- NOT the reference's trace(): it is what the
 * reference drafts in sample_direct_lighting2 (core/tracing.py:56-90) carried through: emitters
 * radiate material.emission from their front side; at a Lambert vertex one light sample (triangle
 * uniform among the nl light triangles, point uniform on it: shapes.py:62-71) and the cosine BSDF
 * sample (samplers.py) are combined with mis_power_heuristic (tracing.py:17-22) using
 * compute_area_light_pdf (:25-33, with the real area instead of 1.0) and compute_brdf_pdf (:36-38);
 * the path ends on an emitter.  Same Philox streams as trace_path.  Its external anchor is media/cornell-box/TungstenRender.exr. */
static double tri_area(const float* tv) {
    double e1[3] = {(double)tv[3] - tv[0], (double)tv[4] - tv[1], (double)tv[5] - tv[2]};
    double e2[3] = {(double)tv[6] - tv[0], (double)tv[7] - tv[1], (double)tv[8] - tv[2]};
    double c[3];
    cross3(e1, e2, c);
    return 0.5 * sqrt(dot3(c, c));
}

static void trace_path_physical(const scene_t* sc, const orc_render_params* P, uint32_t pixel,
                                uint32_t sample, double o[3], double d[3], double L[3], int32_t* prim_id,
                                uint64_t* n_closest, uint64_t* n_shadow) {
    double beta[3] = {1.0, 1.0, 1.0};
    double pdf_prev = -1.0; /* solid-angle pdf of the BSDF sample that produced this ray; < 0: none/delta */
    L[0] = L[1] = L[2] = 0.0;
    *prim_id = -1;
    for (uint32_t bounce = 0; bounce < P->max_depth; ++bounce) {
        double t, bu, bv;
        int id = closest_one(&sc->soup, o, d, (double)P->tmin, (double)P->tmax, &t, &bu, &bv);
        ++*n_closest;
        if (bounce == 0) *prim_id = id;
        if (id < 0) break;
        const orc_material* m = &sc->mats[sc->tri_mat[id]];
        double n[3] = {sc->normals[id * 3], sc->normals[id * 3 + 1], sc->normals[id * 3 + 2]};
        double nd[3] = {-d[0], -d[1], -d[2]};
        double p[3] = {o[0] + d[0] * t, o[1] + d[1] * t, o[2] + d[2] * t};
        if (m->type == 1) {
            double cl = dot3(nd, n) / norm3(d);
            if (cl > 0.0) { /* one-sided emitter */
                double w = 1.0;
                if (pdf_prev > 0.0) {
                    double dist2 = t * t * dot3(d, d);
                    double pl = dist2 / (cl * tri_area(sc->tris + (size_t)id * 9) * (double)sc->nl);
                    w = pdf_prev * pdf_prev / (pdf_prev * pdf_prev + pl * pl);
                }
                for (int k = 0; k < 3; ++k) L[k] += beta[k] * (double)m->emission[k] * w;
            }
            break; /* the path ends on the emitter: with this rule the render agrees with
                      TungstenRender.exr to 0.03 % in the mean (pass-through: +1.0 %) */
        }
        int front = dot3(n, nd) >= 0.0;
        if (m->two_sided && !front) { n[0] = -n[0]; n[1] = -n[1]; n[2] = -n[2]; }
        uint32_t r1[4];
        rng4(P->seed, pixel, sample, bounce, 1, r1);
        double wi[3];
        if (m->type == 0) {
            if (sc->nl > 0) {
                uint32_t r2[4];
                rng4(P->seed, pixel, sample, bounce, 2, r2);
                uint32_t lt = sc->light_tris[rand_index(r1[2], sc->nl)];
                double su = sqrt(u24(r2[0])), sv = u24(r2[1]);
                double a = su * (1 - sv), b = su * sv, c = 1.0 - a - b;
                const float* tv = sc->tris + (size_t)lt * 9;
                double p2[3];
                for (int k = 0; k < 3; ++k)
                    p2[k] = a * (double)tv[k] + b * (double)tv[3 + k] + c * (double)tv[6 + k];
                double n2[3] = {sc->normals[lt * 3], sc->normals[lt * 3 + 1], sc->normals[lt * 3 + 2]};
                double w[3] = {p2[0] - p[0], p2[1] - p[1], p2[2] - p[2]};
                double dist2 = dot3(w, w), dist = sqrt(dist2);
                w[0] /= dist; w[1] /= dist; w[2] /= dist;
                double cos1 = dot3(n, w), cos2 = -dot3(n2, w);
                if (cos1 > 0.0 && cos2 > 0.0) {
                    ++*n_shadow;
                    if (!any_one(&sc->soup, p, w, (double)P->tmin, dist * (1.0 - 1e-4))) {
                        const orc_material* lm = &sc->mats[sc->tri_mat[lt]];
                        double pl = dist2 / (cos2 * tri_area(tv) * (double)sc->nl);
                        double pb = cos1 * ORC_INV_PI;
                        double wgt = pl * pl / (pl * pl + pb * pb);
                        for (int k = 0; k < 3; ++k)
                            L[k] += beta[k] * (double)m->albedo[k] * ORC_INV_PI * (double)lm->emission[k] * cos1 * wgt / pl;
                    }
                }
            }
            orc_cosine_sample_hemisphere(n, u24(r1[0]), u24(r1[1]), wi);
            double c = dot3(n, wi);
            if (!(c > 0.0)) break;
            for (int k = 0; k < 3; ++k) beta[k] *= (double)m->albedo[k]; /* f cos / pdf = albedo */
            pdf_prev = c * ORC_INV_PI;
        } else {
            double ns[3] = {n[0], n[1], n[2]};
            if (!front && !m->two_sided) { ns[0] = -n[0]; ns[1] = -n[1]; ns[2] = -n[2]; }
            double ud[3] = {d[0], d[1], d[2]};
            normalize3(ud);
            uint32_t r2[4];
            rng4(P->seed, pixel, sample, bounce, 2, r2);
            int ok = 1;
            if (m->type == 2) {
                reflect3(ud, ns, wi);
            } else if (m->type == 4) {
                double f[3];
                reflect3(ud, ns, wi);
                in_unit_sphere(u24(r1[0]), u24(r1[1]), u24(r2[2]), f);
                for (int k = 0; k < 3; ++k) wi[k] += (double)m->roughness * f[k];
                ok = dot3(wi, ns) > 0.0;
            } else {
                double ratio = front ? 1.0 / (double)m->ior : (double)m->ior;
                double ct = -dot3(ud, ns);
                if (ct > 1.0) ct = 1.0;
                double st = sqrt(1.0 - ct * ct);
                if (ratio * st > 1.0 || schlick(ct, ratio) > u24(r1[0])) reflect3(ud, ns, wi);
                else refract3(ud, ns, ratio, wi);
            }
            if (!ok) break;
            normalize3(wi);
            for (int k = 0; k < 3; ++k) beta[k] *= (double)m->albedo[k];
            pdf_prev = -1.0;
        }
        if (bounce >= P->rr_start) {
            double q = beta[0] > beta[1] ? beta[0] : beta[1];
            if (beta[2] > q) q = beta[2];
            if (q < 1.0) {
                if (!(u24(r1[3]) < q)) break;
                beta[0] /= q; beta[1] /= q; beta[2] /= q;
            }
        }
        o[0] = p[0]; o[1] = p[1]; o[2] = p[2];
        d[0] = wi[0]; d[1] = wi[1]; d[2] = wi[2];
    }
}

/* Render samples [spp_begin, spp_end) of every pixel and ADD them to
 * accum[h][w][4] (r,g,b sums and sample count) -- main.py:28-37 sums colours
 * per pixel; division by the count is the resolve step.
 * prim_ids (optional) [h][w][ns] primary-hit triangle ids per sample.
 * stats (optional) [2] = closest rays, shadow rays.
 * row0/row1 restrict to image rows [row0,row1) (bounded bench samples). */
int orc_render(const float* tris, const float* normals, uint32_t nt, const uint32_t* tri_mat,
               const orc_material* mats, uint32_t nm, const uint32_t* light_tris, uint32_t nl,
               const orc_camera* cam, const orc_render_params* P, uint32_t row0, uint32_t row1,
               double* accum, int32_t* prim_ids, uint64_t* stats, int nthreads) {
    (void)nm;
    scene_t sc;
    if (soup_init(&sc.soup, tris, nt)) return -1;
    sc.normals = normals; sc.tri_mat = tri_mat; sc.mats = mats;
    sc.light_tris = light_tris; sc.nl = nl; sc.tris = tris;
    uint32_t W = cam->width, H = cam->height, ns = P->spp_end - P->spp_begin;
    if (row1 > H) row1 = H;
    uint64_t nc = 0, nsh = 0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : nc, nsh)
    for (int64_t p = (int64_t)row0 * W; p < (int64_t)row1 * W; ++p) {
        uint32_t i = (uint32_t)(p % W), j = (uint32_t)(p / W);
        double sum[3] = {0, 0, 0};
        for (uint32_t s = P->spp_begin; s < P->spp_end; ++s) {
            uint32_t r[4];
            rng4(P->seed, (uint32_t)p, s, 0, 0, r);
            double u = ((double)i + u24(r[0])) / (double)W;
            double v = ((double)j + u24(r[1])) / (double)H;
            double o[3], d[3], L[3];
            orc_generate_ray_lens(cam, u, v, u24(r[2]), u24(r[3]), o, d);
            /* rays cross the device boundary as f32 records */
            for (int k = 0; k < 3; ++k) { o[k] = (double)(float)o[k]; d[k] = (double)(float)d[k]; }
            int32_t pid;
            uint64_t c1 = 0, c2 = 0;
            if (P->flags & ORC_RENDER_PHYSICAL) trace_path_physical(&sc, P, (uint32_t)p, s, o, d, L, &pid, &c1, &c2);
            else trace_path(&sc, P, (uint32_t)p, s, o, d, L, &pid, &c1, &c2);
            nc += c1; nsh += c2;
            if (prim_ids) prim_ids[(size_t)p * ns + (s - P->spp_begin)] = pid;
            sum[0] += L[0]; sum[1] += L[1]; sum[2] += L[2];
        }
        accum[(size_t)p * 4 + 0] += sum[0];
        accum[(size_t)p * 4 + 1] += sum[1];
        accum[(size_t)p * 4 + 2] += sum[2];
        accum[(size_t)p * 4 + 3] += (double)ns;
    }
    if (stats) { stats[0] = nc; stats[1] = nsh; }
    soup_free(&sc.soup);
    return 0;
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
