// wavefront.cu -- the wavefront path-tracing integrator.
//
// The reference traces one pixel per thread in a megakernel
// (main_taichi.py:80-99 -> PathTracer.trace core/tracing.py:116-155), or one
// pixel per joblib task on the CPU (main.py:28-55).  Here a "wave" is up to
// `wave_paths` paths = k whole samples of every pixel; per bounce four kernels
// run over device-side queues, all launched with a fixed persistent grid so the
// host never reads a queue length back:
//
//   raygen_kernel    camera rays (core/camera.py:41-72, FP64, bit-equal to the
//                    oracle's f32 ray records) + path state init
//   closest_kernel   persistent warps pull 32 queue entries at a time and run the
//                    closest-hit traversal (traverse.cuh)
//   shade_kernel     emitter / two-sided flip / cosine or specular sample /
//                    throughput update / NEE shadow-ray emission / Russian
//                    roulette; survivors are compacted into the next queue with a
//                    warp ballot + prefix popcount and ONE atomic per warp
//   shadow_kernel    any-hit traversal of the shadow queue, adds the pending NEE
//                    contribution to the path's radiance if unoccluded
//   advance_kernel   rotates the queue counters (1 thread)
//   accumulate_kernel  sums the k samples of each pixel in a fixed order and adds
//                    them to the caller's accumulation buffer (deterministic)
//
// The estimator is the reference's, including its quirks:
// hard-coded light colour for directly seen emitters, cosine-at-light factor
// after the first bounce, NEE without 1/pi or pdf, one-sided emitter.
#include "context.cuh"
#include "shading.cuh"
#include "persist.cuh"
#include "traverse.cuh"

namespace prt {

struct WaveState {
    float4* rays = nullptr;     // [2*cap] by path slot
    float4* hits = nullptr;     // [cap]
    float4* beta = nullptr;     // [cap]
    float4* L = nullptr;        // [cap]
    float4* srays = nullptr;    // [2*cap] by shadow-queue position
    float4* scontrib = nullptr; // [cap] rgb + bits(path slot)
    uint32_t* queue[2] = {nullptr, nullptr};
    unsigned int* cnt = nullptr;  // [8]: cur, next, shadow, fetch_closest, fetch_shadow
    uint64_t cap = 0;
    int grid_trace = 0, grid_shade = 0;
};

struct CamDev {
    double m[16];
    double sw, sh, focal, aperture;
    uint32_t W, H;
};

struct WaveParams {
    unsigned long long seed;
    uint32_t s_begin;      // first sample index of this wave
    uint32_t ns_wave;      // samples of every pixel in this wave
    uint32_t npix;
    uint32_t spp_begin;    // of the whole render call (prim_ids layout)
    uint32_t ns_total;
    uint32_t rr_start;
    float3 light_color;
    float tmin, tmax;
    // path-segment log (prt_set_path_log): nullptr = off
    float4* log;
    uint32_t* log_count;
    uint32_t log_capacity;
};

// one record = 2 x float4: (p0, kind), (p1, path)
__device__ __forceinline__ void log_segment(const WaveParams& P, float3 p0, float3 p1, int kind, uint32_t path) {
    const uint32_t i = atomicAdd(P.log_count, 1u);
    if (i < P.log_capacity) {
        P.log[2 * (size_t)i] = make_float4(p0.x, p0.y, p0.z, __int_as_float(kind));
        P.log[2 * (size_t)i + 1] = make_float4(p1.x, p1.y, p1.z, __uint_as_float(path));
    }
}

// (lu, lv): lens sample in [0,1)^2, used only when cam.aperture > 0 (core/camera.py:63-65)
__device__ __forceinline__ void camera_ray(const CamDev& cam, double u, double v, double lu, double lv, float4& ro,
                                           float4& rd, float tmin, float tmax) {
    // core/camera.py:48-70 in the oracle's operation order (oracle/pt_oracle.c orc_generate_ray)
    double cs0 = __dsub_rn(u, 0.5), cs1 = __dsub_rn(v, 0.5);
    double r0 = __ddiv_rn(__dmul_rn(cs0, cam.sw), 0.5), r1 = __ddiv_rn(__dmul_rn(cs1, cam.sh), 0.5);
    double h0 = (double)__double2float_rn(r0), h1 = (double)__double2float_rn(r1),
           h2 = (double)__double2float_rn(-cam.focal);
    double f[3], o[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        double dw = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(h0, cam.m[j]), __dmul_rn(h1, cam.m[4 + j])),
                                        __dmul_rn(h2, cam.m[8 + j])), cam.m[12 + j]);
        o[j] = cam.m[12 + j];
        if (cam.aperture > 0.0) {  // origin (ax, ay, 0, 1) @ iview, f32-rounded like every homogeneous vector (vec3.py:20-23)
            const double ax = (double)__double2float_rn(__dsub_rn(__dmul_rn(cam.aperture, lu), __ddiv_rn(cam.aperture, 2.0)));
            const double ay = (double)__double2float_rn(__dsub_rn(__dmul_rn(cam.aperture, lv), __ddiv_rn(cam.aperture, 2.0)));
            o[j] = __dadd_rn(__dadd_rn(__dmul_rn(ax, cam.m[j]), __dmul_rn(ay, cam.m[4 + j])), cam.m[12 + j]);
        }
        f[j] = __dsub_rn(dw, o[j]);
    }
    double n = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(f[0], f[0]), __dmul_rn(f[1], f[1])), __dmul_rn(f[2], f[2])));
    ro = make_float4(__double2float_rn(o[0]), __double2float_rn(o[1]), __double2float_rn(o[2]), tmin);
    rd = make_float4(__double2float_rn(__ddiv_rn(f[0], n)), __double2float_rn(__ddiv_rn(f[1], n)),
                     __double2float_rn(__ddiv_rn(f[2], n)), tmax);
}

// API-level ray generation: rays[(pixel*ns + s)] for s in [s0,s1)
__global__ void generate_rays_kernel(CamDev cam, unsigned long long seed, uint32_t s0, uint32_t ns,
                                     int jitter, float tmin, float tmax, float4* rays) {
    uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t total = (uint64_t)cam.W * cam.H * ns;
    if (idx >= total) return;
    uint32_t pixel = (uint32_t)(idx / ns), s = s0 + (uint32_t)(idx % ns);
    uint32_t i = pixel % cam.W, j = pixel / cam.W;
    double jx = 0.5, jy = 0.5, lu = 0.5, lv = 0.5;
    if (jitter) {
        uint4 r = rng4(seed, pixel, s, 0, 0);
        jx = (double)u24(r.x); jy = (double)u24(r.y);
        lu = (double)u24(r.z); lv = (double)u24(r.w);
    }
    double u = __ddiv_rn(__dadd_rn((double)i, jx), (double)cam.W);
    double v = __ddiv_rn(__dadd_rn((double)j, jy), (double)cam.H);
    float4 ro, rd;
    camera_ray(cam, u, v, lu, lv, ro, rd, tmin, tmax);
    rays[2 * idx] = ro;
    rays[2 * idx + 1] = rd;
}

__global__ void raygen_kernel(CamDev cam, WaveParams P, float4* rays, float4* beta, float4* L,
                              uint32_t* queue, unsigned int* cnt) {
    uint32_t n_paths = P.npix * P.ns_wave;
    uint32_t pid = blockIdx.x * blockDim.x + threadIdx.x;
    if (pid == 0) { cnt[0] = n_paths; cnt[1] = 0; cnt[2] = 0; cnt[3] = 0; cnt[4] = 0; }
    if (pid >= n_paths) return;
    uint32_t pixel = pid % P.npix, s = P.s_begin + pid / P.npix;
    uint32_t i = pixel % cam.W, j = pixel / cam.W;
    uint4 r = rng4(P.seed, pixel, s, 0, 0);
    double u = __ddiv_rn(__dadd_rn((double)i, (double)u24(r.x)), (double)cam.W);
    double v = __ddiv_rn(__dadd_rn((double)j, (double)u24(r.y)), (double)cam.H);
    float4 ro, rd;
    camera_ray(cam, u, v, (double)u24(r.z), (double)u24(r.w), ro, rd, P.tmin, P.tmax);
    rays[2 * (size_t)pid] = ro;
    rays[2 * (size_t)pid + 1] = rd;
    beta[pid] = make_float4(1.f, 1.f, 1.f, -1.f);
    L[pid] = make_float4(0.f, 0.f, 0.f, 0.f);
    queue[pid] = pid;
}

// path start from caller-supplied rays (prt_trace_paths): "pixel" == ray index
__global__ void init_paths_kernel(const float4* __restrict__ user_rays, WaveParams P, float4* rays, float4* beta,
                                  float4* L, uint32_t* queue, unsigned int* cnt) {
    uint32_t n_paths = P.npix * P.ns_wave;
    uint32_t pid = blockIdx.x * blockDim.x + threadIdx.x;
    if (pid == 0) { cnt[0] = n_paths; cnt[1] = 0; cnt[2] = 0; cnt[3] = 0; cnt[4] = 0; }
    if (pid >= n_paths) return;
    uint32_t pixel = pid % P.npix;
    rays[2 * (size_t)pid] = user_rays[2 * (size_t)pixel];
    rays[2 * (size_t)pid + 1] = user_rays[2 * (size_t)pixel + 1];
    beta[pid] = make_float4(1.f, 1.f, 1.f, -1.f);
    L[pid] = make_float4(0.f, 0.f, 0.f, 0.f);
    queue[pid] = pid;
}

// persistent closest-hit over the current queue (persist.cuh: lane-level dynamic fetch)
struct QueueClosestIO {
    const float4* rays;
    float4* hits;
    const uint32_t* queue;
    __device__ __forceinline__ void load(unsigned k, float4& ro, float4& rd, uint32_t& tag) const {
        tag = queue[k];
        ldg256_cs(rays + 2 * (size_t)tag, ro, rd);  // wave buffers are cudaMalloc'd: 32-byte aligned
    }
    __device__ __forceinline__ void prefetch(unsigned) const {}
    __device__ __forceinline__ void store(uint32_t tag, float t, float u, float v, int gid) const {
        hits[tag] = make_float4(t, u, v, __int_as_float(gid));
    }
    __device__ __forceinline__ void flag(uint32_t) const {}  // (EXACT instantiations only)
    __device__ __forceinline__ void reload(uint32_t, float4&, float4&) const {}
};

// COUNT (PRT_RENDER_COUNT): the counter-instrumented twin -- node visits / triangle tests per ray of
// the render's own ray population (all bounces), for the render leg's roofline; never timed
template <bool COUNT>
__global__ void __launch_bounds__(kTraceThreads, PRT_MIN_BLOCKS)
closest_kernel(SceneDev sc, const float4* __restrict__ rays, float4* hits,
               const uint32_t* __restrict__ queue, unsigned int* cnt, Counters* ctr) {
    __shared__ uint2 s_stack[kPStack][kTraceThreads];
    QueueClosestIO io{rays, hits, queue};
    trace_persistent<MODE_CLOSEST, COUNT, false>(sc, io, cnt + 3, cnt[0], &s_stack[0][threadIdx.x], ctr);
}

// shadow rays: an unoccluded ray adds its pending NEE contribution to the path's radiance
template <bool LOG>
struct QueueShadowIO {
    const float4* srays;
    const float4* scontrib;
    float4* L;
    const WaveParams* P;  // kernel parameter space; LOG only
    __device__ __forceinline__ void load(unsigned k, float4& ro, float4& rd, uint32_t& tag) const {
        tag = k;
        ldg256_cs(srays + 2 * (size_t)k, ro, rd);
    }
    __device__ __forceinline__ void prefetch(unsigned k) const {
        asm volatile("prefetch.global.L2 [%0];" ::"l"(srays + 2 * (size_t)k));
    }
    __device__ __forceinline__ void store(uint32_t tag, float, float, float, int gid) const {
        if (gid < 0) {
            float4 c = scontrib[tag];
            uint32_t pid = __float_as_uint(c.w);
            float4 l = L[pid];
            L[pid] = make_float4(l.x + c.x, l.y + c.y, l.z + c.z, 0.f);
            if (LOG) {  // unoccluded light connection: RayLogger.add_line(p, p_light)
                const float4 ro = srays[2 * (size_t)tag], rd = srays[2 * (size_t)tag + 1];
                const uint32_t pixel = pid % P->npix, s = P->s_begin + pid / P->npix;
                log_segment(*P, xyz(ro), xyz(ro) + xyz(rd) * (rd.w / (1.0f - 1e-4f)), -1,
                            pixel * P->ns_total + (s - P->spp_begin));
            }
        }
    }
    __device__ __forceinline__ void flag(uint32_t) const {}
    __device__ __forceinline__ void reload(uint32_t, float4&, float4&) const {}
};

// LOG: the path-segment log (prt_set_path_log) is a separate instantiation, so the production
// kernels carry none of it (as a run-time flag it cost the Cornell render 9 %)
template <bool LOG, bool COUNT>
__global__ void __launch_bounds__(kTraceThreads, PRT_MIN_BLOCKS)
shadow_kernel(SceneDev sc, const __grid_constant__ WaveParams P, const float4* __restrict__ srays,
              const float4* __restrict__ scontrib, float4* L, unsigned int* cnt, Counters* ctr) {
    __shared__ uint2 s_stack[kPStack][kTraceThreads];
    QueueShadowIO<LOG> io{srays, scontrib, L, &P};
    trace_persistent<MODE_ANY, COUNT, false>(sc, io, cnt + 4, cnt[2], &s_stack[0][threadIdx.x], ctr);
}

__device__ __forceinline__ float guard_beta(float albedo, float cz, float pdf) {
    // core/tracing.py:145-149
    float nb = albedo * cz / pdf * kInvPi;
    if (isnan(nb)) nb = albedo * cz / 1e-4f * kInvPi;
    return nb;
}

__device__ __forceinline__ float tri_area_gid(const SceneDev& sc, uint32_t gid) {
    const float3 a = xyz(__ldg(sc.verts_gid + 3 * (size_t)gid)), b = xyz(__ldg(sc.verts_gid + 3 * (size_t)gid + 1)),
                 c = xyz(__ldg(sc.verts_gid + 3 * (size_t)gid + 2));
    const float3 x = cross(b - a, c - a);
    return 0.5f * sqrtf(dot(x, x));
}

// PHYS = PRT_RENDER_PHYSICAL: emitters radiate material.emission, next-event estimation and BSDF
// sampling are weighted with the power heuristic (the scheme of the reference's draft
// sample_direct_lighting2, core/tracing.py:56-90), the path ends on an emitter.  beta.w carries
// the solid-angle pdf of the BSDF sample that produced the current ray (< 0: camera / specular).
#ifndef PRT_SHADE_MIN_BLOCKS
#define PRT_SHADE_MIN_BLOCKS 4
#endif
#ifndef PRT_SHADE_PREFETCH
#define PRT_SHADE_PREFETCH 1
#endif
template <bool PHYS, bool LOG>
__global__ void __launch_bounds__(256, PRT_SHADE_MIN_BLOCKS)
shade_kernel(SceneDev sc, WaveParams P, uint32_t bounce, uint32_t max_depth, float4* rays,
             const float4* __restrict__ hits, float4* beta, float4* L, float4* srays,
             float4* scontrib, const uint32_t* __restrict__ queue_in, uint32_t* queue_out,
             unsigned int* cnt, int32_t* prim_ids) {
    const unsigned int n = cnt[0];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned int stride = gridDim.x * blockDim.x;
    // queue compaction is aggregated per BLOCK (ballot -> per-warp counts in shared memory -> one
    // atomicAdd per queue per block iteration): with one atomic per warp the two queue tails took
    // 0.5 M same-address atomics per launch and 56 % of this kernel's stall samples were lanes
    // waiting for them (profiles/r1_shade_kernel_ncu.txt).  Double-buffered by iteration parity.
    __shared__ unsigned int s_cnt[2][2][8], s_base[2][2];
    unsigned int parity = 0;
    // block-uniform trip count, so the barriers below are reached by every thread
    // The path state of the NEXT block iteration is pulled into L2 while this one is shaded: the queue entry is
    // loaded one iteration ahead and its ray / hit / throughput records are prefetched, so the dependent chain
    // queue -> state no longer waits on DRAM twice per iteration (the kernel sat at half the issue slots and
    // half the DRAM bandwidth, profiles/ncu_cornell_shade.json).
    uint32_t pid_next = 0;
#if PRT_SHADE_PREFETCH
    if (blockIdx.x * blockDim.x + threadIdx.x < n) pid_next = queue_in[blockIdx.x * blockDim.x + threadIdx.x];
#endif
    for (unsigned int kb = blockIdx.x * blockDim.x; kb < n; kb += stride, parity ^= 1u) {
        unsigned int k = kb + threadIdx.x;
        bool alive = false, want_shadow = false;
        uint32_t pid = 0;
        float4 sro, srd, sc4;
#if PRT_SHADE_PREFETCH
        const uint32_t pid_cur = pid_next;
        if (k + stride < n) {
            pid_next = queue_in[k + stride];
            asm volatile("prefetch.global.L2 [%0];" ::"l"(rays + 2 * (size_t)pid_next + 1));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(hits + pid_next));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(beta + pid_next));
        }
#endif
        if (k < n) {
#if PRT_SHADE_PREFETCH
            pid = pid_cur;
#else
            pid = queue_in[k];
#endif
            float4 ro = rays[2 * (size_t)pid], rd = rays[2 * (size_t)pid + 1];
            float4 h = hits[pid];
            int gid = __float_as_int(h.w);
            uint32_t pixel = pid % P.npix, s = P.s_begin + pid / P.npix;
            if (bounce == 0 && prim_ids)
                prim_ids[(size_t)pixel * P.ns_total + (s - P.spp_begin)] = gid;
            if (LOG) {  // debug/ray_logger.py: origin -> hit point, or 5 units along a ray that escapes
                const float3 lo3 = xyz(ro), ld3 = xyz(rd);
                const float tl = gid >= 0 ? h.x : 5.0f;
                log_segment(P, lo3, lo3 + ld3 * tl, (int)bounce, pixel * P.ns_total + (s - P.spp_begin));
            }
            if (gid >= 0) {
                float3 o = xyz(ro), d = xyz(rd);
                float4 sh = __ldg(sc.shade + gid);
                float3 n = xyz(sh);
                const prt_material m = sc.mats[__float_as_uint(sh.w)];
                const float4 b4 = beta[pid];
                float3 b = xyz(b4);
                float pdf_prev = b4.w;
                float3 nd = -d;
                if (PHYS && m.type == PRT_MAT_EMITTER) {
                    const float dd = dot(d, d);
                    const float cl = dot(nd, n) * rsqrtf(dd);
                    if (cl > 0.0f) {  // one-sided
                        float w = 1.0f;
                        if (pdf_prev > 0.0f) {
                            const float pl = h.x * h.x * dd / (cl * tri_area_gid(sc, (uint32_t)gid) * (float)sc.nl);
                            w = pdf_prev * pdf_prev / (pdf_prev * pdf_prev + pl * pl);
                        }
                        float4 l = L[pid];
                        L[pid] = make_float4(l.x + b.x * m.emission[0] * w, l.y + b.y * m.emission[1] * w,
                                             l.z + b.z * m.emission[2] * w, 0.f);
                    }
                } else if (m.type == PRT_MAT_EMITTER) {  // core/tracing.py:129-139
                    float d1 = dot(nd, n);
                    if (d1 > 0.0f) {
                        float w = bounce == 0 ? 1.0f : d1;
                        float4 l = L[pid];
                        L[pid] = make_float4(l.x + P.light_color.x * b.x * w, l.y + P.light_color.y * b.y * w,
                                             l.z + P.light_color.z * b.z * w, 0.f);
                    }
                } else {
                    // Hit point.  The reference evaluates o + d*t in f64 (fast_op.py:110-112); in
                    // FP32 that leaves the point ~1e-7 off the surface and grazing continuation
                    // rays then re-hit it beyond t_min = 1e-5.  The barycentric form lands on the
                    // triangle's plane (exactly, for axis-aligned faces) -- same point, real arithmetic.
                    float3 q0 = xyz(__ldg(sc.verts_gid + 3 * (size_t)gid)),
                           q1 = xyz(__ldg(sc.verts_gid + 3 * (size_t)gid + 1)),
                           q2 = xyz(__ldg(sc.verts_gid + 3 * (size_t)gid + 2));
                    float3 e1 = q1 - q0, e2 = q2 - q0;
                    float3 p = make_float3(fmaf(h.z, e2.x, fmaf(h.y, e1.x, q0.x)), fmaf(h.z, e2.y, fmaf(h.y, e1.y, q0.y)),
                                           fmaf(h.z, e2.z, fmaf(h.y, e1.z, q0.z)));
                    bool front = dot(n, nd) >= 0.0f;
                    if (m.two_sided && !front) n = -n;  // mathematics/shapes.py:99-102
                    uint4 r1 = rng4(P.seed, pixel, s, bounce, 1);
                    float3 wi;
                    bool ok = true, nee = false;
                    float3 alb = make_float3(m.albedo[0], m.albedo[1], m.albedo[2]);
                    const float3 b_in = b;  // PHYS: the light sample is weighted with the throughput BEFORE this bounce
                    if (m.type == PRT_MAT_LAMBERT) {
                        wi = cosine_sample_hemisphere(n, u24(r1.x), u24(r1.y));
                        float c = dot(n, wi);
                        if (PHYS) {
                            ok = c > 0.0f;
                            b = b * alb;  // f cos / pdf = albedo
                            pdf_prev = c * kInvPi;
                        } else {
                            float pdf = fabsf(c) * kInvPi;
                            float cz = fmaxf(c, 0.0f);
                            b = make_float3(b.x * guard_beta(alb.x, cz, pdf), b.y * guard_beta(alb.y, cz, pdf),
                                            b.z * guard_beta(alb.z, cz, pdf));
                        }
                        nee = true;
                    } else {
                        pdf_prev = -1.0f;
                        float3 ns = (!front && !m.two_sided) ? -n : n;
                        float u3 = 0.0f;
                        if (m.type == PRT_MAT_CONDUCTOR) u3 = u24(rng4(P.seed, pixel, s, bounce, 2).z);
                        ok = sample_specular(m.type, d, ns, front, m.ior, m.roughness, u24(r1.x), u24(r1.y), u3, wi);
                        if (ok) {
                            wi = normalize_fast(wi);
                            b = b * alb;
                        }
                    }
                    if (ok || PHYS) {
                        if (nee && sc.nl > 0) {  // core/tracing.py:92-108, shapes.py:62-71
                            uint4 r2 = rng4(P.seed, pixel, s, bounce, 2);
                            uint32_t lt = sc.light_tris[rand_index(r1.z, sc.nl)];
                            float su = sqrtf(u24(r2.x)), sv = u24(r2.y);
                            float a = su * (1.0f - sv), bb = su * sv, c = 1.0f - a - bb;
                            float3 v0 = xyz(__ldg(sc.verts_gid + 3 * (size_t)lt)),
                                   v1 = xyz(__ldg(sc.verts_gid + 3 * (size_t)lt + 1)),
                                   v2 = xyz(__ldg(sc.verts_gid + 3 * (size_t)lt + 2));
                            float3 p2 = v0 * a + v1 * bb + v2 * c;
                            float4 lsh = __ldg(sc.shade + lt);
                            float3 w = p2 - p;
                            float dist2 = dot(w, w);
                            float idist = rsqrtf(dist2), dist = dist2 * idist;
                            w = make_float3(w.x * idist, w.y * idist, w.z * idist);
                            float dot1 = dot(n, w), dot2 = -dot(xyz(lsh), w);
                            if (dot1 > 0.0f && dot2 > 0.0f) {
                                const prt_material lm = sc.mats[__float_as_uint(lsh.w)];
                                want_shadow = true;
                                sro = make_float4(p.x, p.y, p.z, P.tmin);
                                srd = make_float4(w.x, w.y, w.z, dist * (1.0f - 1e-4f));
                                if (PHYS) {
                                    const float3 x = cross(v1 - v0, v2 - v0);
                                    const float pl = dist2 / (dot2 * 0.5f * sqrtf(dot(x, x)) * (float)sc.nl);
                                    const float pb = dot1 * kInvPi;
                                    const float g = kInvPi * dot1 * (pl * pl / (pl * pl + pb * pb)) / pl;
                                    sc4 = make_float4(b_in.x * alb.x * lm.emission[0] * g, b_in.y * alb.y * lm.emission[1] * g,
                                                      b_in.z * alb.z * lm.emission[2] * g, __uint_as_float(pid));
                                } else {
                                    float g = dot1 * dot2 / dist2;
                                    sc4 = make_float4(b.x * lm.albedo[0] * g, b.y * lm.albedo[1] * g,
                                                      b.z * lm.albedo[2] * g, __uint_as_float(pid));
                                }
                            }
                        }
                        alive = ok && bounce + 1 < max_depth;
                        if (bounce >= P.rr_start) {
                            float q = fmaxf(b.x, fmaxf(b.y, b.z));
                            if (q < 1.0f) {
                                if (!(u24(r1.w) < q)) alive = false;
                                else b = make_float3(b.x / q, b.y / q, b.z / q);
                            }
                        }
                        if (alive) {
                            beta[pid] = make_float4(b.x, b.y, b.z, pdf_prev);
                            rays[2 * (size_t)pid] = make_float4(p.x, p.y, p.z, P.tmin);
                            rays[2 * (size_t)pid + 1] = make_float4(wi.x, wi.y, wi.z, P.tmax);
                        }
                    }
                }
            }
        }
        // block-aggregated compaction: ballot + prefix popcount, one atomic per block per queue
        unsigned int am = __ballot_sync(0xffffffffu, alive);
        unsigned int sm = __ballot_sync(0xffffffffu, want_shadow);
        if (lane == 0) { s_cnt[parity][0][warp] = __popc(am); s_cnt[parity][1][warp] = __popc(sm); }
        __syncthreads();
        if (threadIdx.x < 2) {  // thread q reserves queue q
            unsigned int tot = 0;
#pragma unroll
            for (int w8 = 0; w8 < 8; ++w8) tot += s_cnt[parity][threadIdx.x][w8];
            s_base[parity][threadIdx.x] = tot ? atomicAdd(cnt + 1 + threadIdx.x, tot) : 0u;
        }
        __syncthreads();
        unsigned int abase = s_base[parity][0], sbase = s_base[parity][1];
        for (int w8 = 0; w8 < warp; ++w8) { abase += s_cnt[parity][0][w8]; sbase += s_cnt[parity][1][w8]; }
        unsigned int lt_mask = (1u << lane) - 1u;
        if (alive) queue_out[abase + __popc(am & lt_mask)] = pid;
        if (want_shadow) {
            unsigned int sk = sbase + __popc(sm & lt_mask);
            srays[2 * (size_t)sk] = sro;
            srays[2 * (size_t)sk + 1] = srd;
            scontrib[sk] = sc4;
        }
    }
}

// (count_rays = false under PRT_RENDER_COUNT: the counted traversal kernels add the rays themselves)
__global__ void advance_kernel(unsigned int* cnt, Counters* ctr, bool count_rays) {
    if (count_rays) {
        atomicAdd(&ctr->rays_closest, (unsigned long long)cnt[0]);
        atomicAdd(&ctr->rays_shadow, (unsigned long long)cnt[2]);
    }
    cnt[0] = cnt[1];
    cnt[1] = 0; cnt[2] = 0; cnt[3] = 0; cnt[4] = 0;
}

__global__ void accumulate_kernel(const float4* __restrict__ L, uint32_t npix, uint32_t ns_wave,
                                  float4* accum, Counters* ctr) {
    uint32_t pixel = blockIdx.x * blockDim.x + threadIdx.x;
    if (pixel == 0) atomicAdd(&ctr->paths, (unsigned long long)npix * ns_wave);
    if (pixel >= npix) return;
    float3 s = make_float3(0.f, 0.f, 0.f);
    for (uint32_t k = 0; k < ns_wave; ++k) {
        float4 l = L[(size_t)k * npix + pixel];
        s.x += l.x; s.y += l.y; s.z += l.z;
    }
    float4 a = accum[pixel];
    accum[pixel] = make_float4(a.x + s.x, a.y + s.y, a.z + s.z, a.w + (float)ns_wave);
}

// known-answer hook: the production specular sampler on caller-supplied inputs
__global__ void eval_specular_kernel(const prt_bsdf_query* __restrict__ q, uint64_t n, float4* out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const prt_bsdf_query a = q[i];
    float3 wi = make_float3(0.f, 0.f, 0.f);
    const bool ok = sample_specular(a.type, make_float3(a.d[0], a.d[1], a.d[2]), make_float3(a.ns[0], a.ns[1], a.ns[2]),
                                    a.front != 0u, a.ior, a.roughness, a.u[0], a.u[1], a.u[2], wi);
    out[i] = make_float4(wi.x, wi.y, wi.z, ok ? 1.0f : 0.0f);
}

int eval_specular(prt_ctx* ctx, const prt_bsdf_query* q, uint64_t n, float* out, cudaStream_t stream) {
    if (n == 0) return PRT_OK;
    eval_specular_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(q, n, (float4*)out);
    PRT_CUDA_TRY(ctx, cudaGetLastError());
    return PRT_OK;
}

static CamDev cam_dev(const prt_camera& c) {
    CamDev d;
    for (int i = 0; i < 16; ++i) d.m[i] = c.iview[i];
    d.sw = c.sensor_w; d.sh = c.sensor_h; d.focal = c.focal; d.aperture = c.aperture;
    d.W = c.width; d.H = c.height;
    return d;
}

int generate_rays(prt_ctx* ctx, uint64_t seed, uint32_t s0, uint32_t s1, int jitter, float tmin,
                  float tmax, float4* rays, cudaStream_t stream) {
    if (!ctx->cam_set) { ctx->set_error("generate_rays: camera not set"); return PRT_ERR_STATE; }
    if (s1 <= s0) { ctx->set_error("generate_rays: empty sample range"); return PRT_ERR_INVALID; }
    uint64_t total = (uint64_t)ctx->cam.width * ctx->cam.height * (s1 - s0);
    if (total > (1ull << 31)) { ctx->set_error("generate_rays: too many rays for one call"); return PRT_ERR_INVALID; }
    generate_rays_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(cam_dev(ctx->cam), seed, s0, s1 - s0,
                                                                           jitter, tmin, tmax, rays);
    PRT_CUDA_TRY(ctx, cudaGetLastError());
    return PRT_OK;
}

void wavefront_free(prt_ctx* ctx) {
    WaveState* w = (WaveState*)ctx->wf;
    if (!w) return;
    cudaFree(w->rays); cudaFree(w->hits); cudaFree(w->beta); cudaFree(w->L); cudaFree(w->srays);
    cudaFree(w->scontrib); cudaFree(w->queue[0]); cudaFree(w->queue[1]); cudaFree(w->cnt);
    delete w;
    ctx->wf = nullptr;
}

static int wave_alloc(prt_ctx* ctx, uint64_t cap) {
    WaveState* w = (WaveState*)ctx->wf;
    if (w && w->cap >= cap) return PRT_OK;
    wavefront_free(ctx);
    w = new WaveState();
    ctx->wf = w;
    PRT_CUDA_TRY(ctx, cudaMalloc(&w->rays, sizeof(float4) * 2 * cap));
    PRT_CUDA_TRY(ctx, cudaMalloc(&w->hits, sizeof(float4) * cap));
    PRT_CUDA_TRY(ctx, cudaMalloc(&w->beta, sizeof(float4) * cap));
    PRT_CUDA_TRY(ctx, cudaMalloc(&w->L, sizeof(float4) * cap));
    PRT_CUDA_TRY(ctx, cudaMalloc(&w->srays, sizeof(float4) * 2 * cap));
    PRT_CUDA_TRY(ctx, cudaMalloc(&w->scontrib, sizeof(float4) * cap));
    PRT_CUDA_TRY(ctx, cudaMalloc(&w->queue[0], sizeof(uint32_t) * cap));
    PRT_CUDA_TRY(ctx, cudaMalloc(&w->queue[1], sizeof(uint32_t) * cap));
    PRT_CUDA_TRY(ctx, cudaMalloc(&w->cnt, sizeof(unsigned int) * 8));
    w->cap = cap;
    int bt = 0, bs = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bt, closest_kernel<false>, kTraceThreads, 0);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bs, shade_kernel<false, false>, 256, 0);
    w->grid_trace = ctx->num_sms * (bt > 0 ? bt : 4);
    w->grid_shade = ctx->num_sms * (bs > 0 ? bs : 4);
    return PRT_OK;
}

// user_rays == nullptr: camera rays for every pixel (prt_render); else one path set per supplied ray
int render(prt_ctx* ctx, const prt_render_params* p, float* accum, int32_t* prim_ids,
           cudaStream_t stream, const float4* user_rays, uint64_t n_user) {
    if (!ctx->scene_set || !ctx->bvh_built) { ctx->set_error("render: scene/BVH not ready"); return PRT_ERR_STATE; }
    if (!user_rays && !ctx->cam_set) { ctx->set_error("render: camera not set"); return PRT_ERR_STATE; }
    if (p->spp_end < p->spp_begin) { ctx->set_error("render: spp_end < spp_begin"); return PRT_ERR_INVALID; }
    if (p->max_depth == 0 || p->spp_end == p->spp_begin) return PRT_OK;
    const uint64_t npix = user_rays ? n_user : (uint64_t)ctx->cam.width * ctx->cam.height;
    if (npix == 0 || npix > (1ull << 30)) { ctx->set_error("render: bad resolution / ray count"); return PRT_ERR_INVALID; }
    uint64_t per_wave = ctx->wave_paths / npix;
    if (per_wave == 0) per_wave = 1;
    const uint32_t ns_total = p->spp_end - p->spp_begin;
    if (per_wave > ns_total) per_wave = ns_total;
    {   // equal waves: 16 spp under a cap of 7 run as 6 + 6 + 4, not 7 + 7 + 2 (a small last wave runs at a fraction of the rate)
        const uint64_t n_waves = (ns_total + per_wave - 1) / per_wave;
        per_wave = (ns_total + n_waves - 1) / n_waves;
    }
    int rc = wave_alloc(ctx, npix * per_wave);
    if (rc != PRT_OK) return rc;
    WaveState* w = (WaveState*)ctx->wf;
    SceneDev sc = ctx->scene_dev();
    CamDev cam = cam_dev(ctx->cam);  // unused with user rays
    WaveParams P;
    P.seed = p->seed; P.npix = (uint32_t)npix; P.spp_begin = p->spp_begin; P.ns_total = ns_total;
    P.rr_start = p->rr_start;
    P.light_color = make_float3(p->light_color[0], p->light_color[1], p->light_color[2]);
    P.tmin = p->tmin; P.tmax = p->tmax;
    P.log = ctx->log_segments; P.log_count = ctx->log_count; P.log_capacity = ctx->log_capacity;
    for (uint32_t s = p->spp_begin; s < p->spp_end; s += (uint32_t)per_wave) {
        P.s_begin = s;
        P.ns_wave = (uint32_t)((p->spp_end - s) < per_wave ? (p->spp_end - s) : per_wave);
        uint32_t n_paths = P.npix * P.ns_wave;
        prof_begin(ctx, PROF_RAYGEN, stream);
        if (user_rays)
            init_paths_kernel<<<(n_paths + 255) / 256, 256, 0, stream>>>(user_rays, P, w->rays, w->beta, w->L, w->queue[0], w->cnt);
        else
            raygen_kernel<<<(n_paths + 255) / 256, 256, 0, stream>>>(cam, P, w->rays, w->beta, w->L, w->queue[0], w->cnt);
        prof_end(ctx, stream);
        const bool counted = (p->flags & PRT_RENDER_COUNT) != 0;
        for (uint32_t b = 0; b < p->max_depth; ++b) {
            uint32_t* qin = w->queue[b & 1];
            uint32_t* qout = w->queue[(b & 1) ^ 1];
            if (b == 0 && (p->flags & PRT_RENDER_EXACT_PRIMARY)) {
                // bounce 0: queue == identity, rays contiguous -> the API-level exact trace applies
                rc = launch_trace(ctx, MODE_CLOSEST, w->rays, n_paths, w->hits, nullptr,
                                  PRT_TRACE_EXACT | PRT_TRACE_NO_BIN | (counted ? PRT_TRACE_COUNT : 0u), stream);  // (camera rays: already in pixel order)
                if (rc != PRT_OK) return rc;
            } else {
                prof_begin(ctx, PROF_CLOSEST, stream);
                if (counted) closest_kernel<true><<<w->grid_trace, kTraceThreads, 0, stream>>>(sc, w->rays, w->hits, qin, w->cnt, ctx->counters);
                else closest_kernel<false><<<w->grid_trace, kTraceThreads, 0, stream>>>(sc, w->rays, w->hits, qin, w->cnt, ctx->counters);
                prof_end(ctx, stream);
            }
            const bool phys = p->flags & PRT_RENDER_PHYSICAL, logging = P.log != nullptr;
#define PRT_SHADE(PH, LG) shade_kernel<PH, LG><<<w->grid_shade, 256, 0, stream>>>(sc, P, b, p->max_depth, w->rays, w->hits, \
                                                  w->beta, w->L, w->srays, w->scontrib, qin, qout, w->cnt, prim_ids)
            prof_begin(ctx, PROF_SHADE, stream);
            if (logging) { if (phys) PRT_SHADE(true, true); else PRT_SHADE(false, true); }
            else { if (phys) PRT_SHADE(true, false); else PRT_SHADE(false, false); }
#undef PRT_SHADE
            prof_end(ctx, stream);
            prof_begin(ctx, PROF_SHADOW, stream);
            if (logging) shadow_kernel<true, false><<<w->grid_trace, kTraceThreads, 0, stream>>>(sc, P, w->srays, w->scontrib, w->L, w->cnt, ctx->counters);
            else if (counted) shadow_kernel<false, true><<<w->grid_trace, kTraceThreads, 0, stream>>>(sc, P, w->srays, w->scontrib, w->L, w->cnt, ctx->counters);
            else shadow_kernel<false, false><<<w->grid_trace, kTraceThreads, 0, stream>>>(sc, P, w->srays, w->scontrib, w->L, w->cnt, ctx->counters);
            prof_end(ctx, stream);
            prof_begin(ctx, PROF_OTHER, stream);
            advance_kernel<<<1, 1, 0, stream>>>(w->cnt, ctx->counters, !counted);
            prof_end(ctx, stream);
        }
        prof_begin(ctx, PROF_OTHER, stream);
        accumulate_kernel<<<(P.npix + 255) / 256, 256, 0, stream>>>(w->L, P.npix, P.ns_wave, (float4*)accum, ctx->counters);
        prof_end(ctx, stream);
    }
    PRT_CUDA_TRY(ctx, cudaGetLastError());
    return PRT_OK;
}

}  // namespace prt
