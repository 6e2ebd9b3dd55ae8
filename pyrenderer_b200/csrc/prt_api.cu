// prt_api.cu -- extern "C" entry points declared in include/prt.h.
#include <stdlib.h>
#include <string.h>

#include <new>
#include <vector>

#include "context.cuh"

// text of the last prt_create failure, per calling thread (prt_last_error(NULL))
static thread_local char g_create_error[512] = "";

// Every entry point runs on the context's device and puts the caller's current device back on
// return (the caller may be torch, with another device current).
struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    cudaError_t enter(int dev) {
        cudaError_t e = cudaGetDevice(&prev);
        if (e != cudaSuccess) return e;
        if (prev == dev) return cudaSuccess;
        e = cudaSetDevice(dev);
        switched = e == cudaSuccess;
        return e;
    }
    ~DeviceGuard() { if (switched) cudaSetDevice(prev); }
};

namespace prt {

// pack host triangles into the device layout: 3 x float4 per triangle with
// w = bits(global id), bits(material), 0
__global__ void pack_tris_kernel(const float* __restrict__ v, const uint32_t* __restrict__ tri_mat,
                                 uint32_t nt, float4* out) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nt) return;
    const float* p = v + 9ull * t;
    uint32_t m = tri_mat ? tri_mat[t] : 0u;
    out[3ull * t] = make_float4(p[0], p[1], p[2], __uint_as_float(t));
    out[3ull * t + 1] = make_float4(p[3], p[4], p[5], __uint_as_float(m));
    out[3ull * t + 2] = make_float4(p[6], p[7], p[8], 0.0f);
}

__global__ void pack_shade_kernel(const float* __restrict__ v, const float* __restrict__ normals,
                                  const uint32_t* __restrict__ tri_mat, uint32_t nt, float4* shade) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nt) return;
    float3 n;
    if (normals) {
        n = make_float3(normals[3ull * t], normals[3ull * t + 1], normals[3ull * t + 2]);
    } else {
        const float* p = v + 9ull * t;
        float3 e1 = make_float3(p[3] - p[0], p[4] - p[1], p[5] - p[2]);
        float3 e2 = make_float3(p[6] - p[0], p[7] - p[1], p[8] - p[2]);
        n = normalize(cross(e1, e2));
    }
    shade[t] = make_float4(n.x, n.y, n.z, __uint_as_float(tri_mat ? tri_mat[t] : 0u));
}

// forget the scene; the device buffers stay (grow-only) for the next one
static void reset_scene(prt_ctx* c) {
    c->nt = c->nm = c->nl = c->n_nodes = 0;
    c->scene_set = false; c->bvh_built = false;
}

static void free_scene(prt_ctx* c) {
    reset_scene(c);
    cudaFree(c->verts_gid); cudaFree(c->shade); cudaFree(c->mats); cudaFree(c->light_tris);
    cudaFree(c->tri_a); cudaFree(c->tri_b); cudaFree(c->nodes); cudaFree(c->build_arena);
    c->verts_gid = nullptr; c->shade = nullptr; c->mats = nullptr; c->light_tris = nullptr;
    c->tri_a = nullptr; c->tri_b = nullptr; c->nodes = nullptr; c->build_arena = nullptr;
    c->verts_bytes = c->shade_bytes = c->mats_bytes = c->lights_bytes = 0;
    c->tri_a_bytes = c->tri_b_bytes = c->nodes_bytes = c->build_arena_bytes = 0;
}

template <class T>
static cudaError_t reserve(T*& ptr, size_t& cap_bytes, size_t bytes) {
    if (cap_bytes >= bytes && ptr) return cudaSuccess;
    cudaFree(ptr);
    ptr = nullptr; cap_bytes = 0;
    const cudaError_t e = cudaMalloc(&ptr, bytes);
    if (e == cudaSuccess) cap_bytes = bytes;
    return e;
}

static int set_scene_common(prt_ctx* ctx, const float* verts_dev, const float* normals_dev,
                            const uint32_t* tri_mat_dev, uint32_t nt, const prt_material* mats_host,
                            uint32_t nm, const uint32_t* light_host, uint32_t nl, cudaStream_t s) {
    reset_scene(ctx);
    prt_material def;
    memset(&def, 0, sizeof def);
    def.albedo[0] = def.albedo[1] = def.albedo[2] = 0.5f;
    def.type = PRT_MAT_LAMBERT; def.ior = 1.0f; def.two_sided = 1;
    if (!mats_host || nm == 0) { mats_host = &def; nm = 1; }
    size_t ntz = nt ? nt : 1;
    PRT_CUDA_TRY(ctx, cudaStreamSynchronize(s));  // (a buffer that has to grow is freed: nothing may still read it)
    PRT_CUDA_TRY(ctx, reserve(ctx->verts_gid, ctx->verts_bytes, sizeof(float4) * 3 * ntz));
    PRT_CUDA_TRY(ctx, reserve(ctx->shade, ctx->shade_bytes, sizeof(float4) * ntz));
    PRT_CUDA_TRY(ctx, reserve(ctx->mats, ctx->mats_bytes, sizeof(prt_material) * nm));
    PRT_CUDA_TRY(ctx, reserve(ctx->light_tris, ctx->lights_bytes, sizeof(uint32_t) * (nl ? nl : 1)));
    PRT_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->mats, mats_host, sizeof(prt_material) * nm, cudaMemcpyHostToDevice, s));
    if (nl) PRT_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->light_tris, light_host, sizeof(uint32_t) * nl, cudaMemcpyHostToDevice, s));
    if (nt) {
        unsigned g = (nt + 255) / 256;
        pack_tris_kernel<<<g, 256, 0, s>>>(verts_dev, tri_mat_dev, nt, ctx->verts_gid);
        pack_shade_kernel<<<g, 256, 0, s>>>(verts_dev, normals_dev, tri_mat_dev, nt, ctx->shade);
        PRT_CUDA_TRY(ctx, cudaGetLastError());
    }
    PRT_CUDA_TRY(ctx, cudaStreamSynchronize(s));
    ctx->nt = nt; ctx->nm = nm; ctx->nl = nl;
    ctx->scene_set = true;
    return PRT_OK;
}

void prof_begin(prt_ctx* ctx, int cls, cudaStream_t stream, int launches) {
    if (!ctx->prof_on) return;
    const size_t i = ctx->prof_class.size();
    while (ctx->prof_events.size() < 2 * (i + 1)) {
        cudaEvent_t e = nullptr;
        if (cudaEventCreate(&e) != cudaSuccess) { ctx->prof_on = false; return; }
        ctx->prof_events.push_back(e);
    }
    ctx->prof_class.push_back(cls);
    ctx->prof_launches.push_back(launches);
    cudaEventRecord(ctx->prof_events[2 * i], stream);
}

void prof_end(prt_ctx* ctx, cudaStream_t stream) {
    if (!ctx->prof_on || ctx->prof_class.empty()) return;
    cudaEventRecord(ctx->prof_events[2 * (ctx->prof_class.size() - 1) + 1], stream);
}

}  // namespace prt

using namespace prt;

#define CHECK_CTX(ctx) \
    do { if (!(ctx)) return PRT_ERR_INVALID; } while (0)
#define USE_DEVICE(ctx)  \
    DeviceGuard _guard; \
    PRT_CUDA_TRY(ctx, _guard.enter((ctx)->device))

extern "C" {

int prt_abi_version(void) { return PRT_ABI_VERSION; }

int prt_create(int device, prt_ctx** out) {
    if (!out) { snprintf(g_create_error, sizeof g_create_error, "prt_create: out == NULL"); return PRT_ERR_INVALID; }
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        snprintf(g_create_error, sizeof g_create_error,
                 "prt_create: no CUDA device (%s); this library has no CPU fallback",
                 e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
        cudaGetLastError();
        return PRT_ERR_CUDA;
    }
    if (device < 0 || device >= ndev) {
        snprintf(g_create_error, sizeof g_create_error, "prt_create: device %d out of range [0,%d)", device, ndev);
        return PRT_ERR_INVALID;
    }
    prt_ctx* c = new (std::nothrow) prt_ctx();
    if (!c) return PRT_ERR_NOMEM;
    c->device = device;
    // tuning knobs (profiles/sweep.py), clamped to what the persistent loop can make progress with:
    // refill_idle > 32 would never refill (the kernel would spin), leaf_batch > 32 only ever fires
    // through the "nobody walks records" rule, a fetch chunk of 0 reserves nothing
    auto knob = [](const char* name, int lo, int hi, int dflt) {
        const char* v = getenv(name);
        if (!v) return dflt;
        const int x = atoi(v);
        return x < lo ? dflt : (x > hi ? hi : x);
    };
    c->refill_idle = knob("PRT_REFILL_IDLE", 1, 32, c->refill_idle);
    c->leaf_batch = knob("PRT_LEAF_BATCH", 1, 32, c->leaf_batch);
    c->fetch_chunk = knob("PRT_FETCH_CHUNK", 1, 4096, c->fetch_chunk);
    DeviceGuard guard;
    e = guard.enter(device);
    cudaDeviceProp prop;
    if (e == cudaSuccess) e = cudaGetDeviceProperties(&prop, device);
    if (e == cudaSuccess) {
        c->num_sms = prop.multiProcessorCount;
        if (prop.major < 10) {
            snprintf(g_create_error, sizeof g_create_error,
                     "prt_create: device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
            delete c;
            return PRT_ERR_CUDA;
        }
    }
    if (e == cudaSuccess) e = cudaMalloc(&c->counters, sizeof(Counters));
    if (e == cudaSuccess) e = cudaMemset(c->counters, 0, sizeof(Counters));
    if (e == cudaSuccess) e = cudaMalloc(&c->flag_count, sizeof(unsigned int) * prt_ctx::kFlagRing);
    if (e == cudaSuccess) e = cudaMalloc(&c->fetch_counters, sizeof(unsigned int) * prt_ctx::kFetchRing);
    if (e != cudaSuccess) {
        snprintf(g_create_error, sizeof g_create_error, "prt_create: %s", cudaGetErrorString(e));
        delete c;
        return PRT_ERR_CUDA;
    }
    *out = c;
    return PRT_OK;
}

void prt_destroy(prt_ctx* ctx) {
    if (!ctx) return;
    DeviceGuard guard;
    guard.enter(ctx->device);
    prt_comm_destroy(ctx);
    wavefront_free(ctx);
    free_scene(ctx);
    cudaFree(ctx->shard_accum);
    for (auto& e : ctx->prof_events) cudaEventDestroy(e);
    for (auto& e : ctx->build_ev) if (e) cudaEventDestroy(e);
    for (auto& l : ctx->flag_list) cudaFree(l);
    for (auto& b : ctx->bin_scratch) cudaFree(b);
    cudaFree(ctx->counters); cudaFree(ctx->flag_count);
    cudaFree(ctx->stage[0]); cudaFree(ctx->stage[1]); cudaFree(ctx->fetch_counters);
    for (auto& s : ctx->copy_stream) if (s) cudaStreamDestroy(s);
    for (auto& e : ctx->copy_event) if (e) cudaEventDestroy(e);
    delete ctx;
}

const char* prt_last_error(const prt_ctx* ctx) { return ctx ? ctx->err : g_create_error; }

int prt_scene_set_triangles(prt_ctx* ctx, const float* verts_host, const float* normals_host,
                            uint32_t nt, const uint32_t* tri_material_host,
                            const prt_material* mats_host, uint32_t nm,
                            const uint32_t* light_tris_host, uint32_t nl) {
    CHECK_CTX(ctx);
    USE_DEVICE(ctx);
    if (nt && !verts_host) { ctx->set_error("scene: verts == NULL"); return PRT_ERR_INVALID; }
    if (nt >= (1u << 28)) { ctx->set_error("scene: at most 2^28-1 triangles"); return PRT_ERR_INVALID; }
    if (nl && !light_tris_host) { ctx->set_error("scene: light_tris == NULL"); return PRT_ERR_INVALID; }
    for (uint32_t i = 0; i < nl; ++i)
        if (light_tris_host[i] >= nt) { ctx->set_error("scene: light triangle %u out of range", light_tris_host[i]); return PRT_ERR_INVALID; }
    if (tri_material_host) {
        // without a material table the scene gets ONE default material: only index 0 exists then
        const uint32_t nm_eff = (mats_host && nm) ? nm : 1u;
        for (uint32_t i = 0; i < nt; ++i)
            if (tri_material_host[i] >= nm_eff) {
                ctx->set_error("scene: material index %u of triangle %u out of range (%u materials)", tri_material_host[i], i, nm_eff);
                return PRT_ERR_INVALID;
            }
    }
    float *dv = nullptr, *dn = nullptr;
    uint32_t* dm = nullptr;
    int rc = PRT_OK;
    cudaError_t e = cudaSuccess;
    if (nt) {
        e = cudaMalloc(&dv, sizeof(float) * 9 * (size_t)nt);
        if (e == cudaSuccess) e = cudaMemcpy(dv, verts_host, sizeof(float) * 9 * (size_t)nt, cudaMemcpyHostToDevice);
        if (e == cudaSuccess && normals_host) {
            e = cudaMalloc(&dn, sizeof(float) * 3 * (size_t)nt);
            if (e == cudaSuccess) e = cudaMemcpy(dn, normals_host, sizeof(float) * 3 * (size_t)nt, cudaMemcpyHostToDevice);
        }
        if (e == cudaSuccess && tri_material_host) {
            e = cudaMalloc(&dm, sizeof(uint32_t) * (size_t)nt);
            if (e == cudaSuccess) e = cudaMemcpy(dm, tri_material_host, sizeof(uint32_t) * (size_t)nt, cudaMemcpyHostToDevice);
        }
    }
    if (e != cudaSuccess) {
        ctx->set_error("scene upload: %s", cudaGetErrorString(e));
        rc = PRT_ERR_CUDA;
    } else {
        rc = set_scene_common(ctx, dv, dn, dm, nt, mats_host, nm, light_tris_host, nl, 0);
    }
    cudaFree(dv); cudaFree(dn); cudaFree(dm);
    return rc;
}

int prt_scene_set_triangles_dev(prt_ctx* ctx, const float* verts_dev, uint32_t nt, void* stream) {
    CHECK_CTX(ctx);
    USE_DEVICE(ctx);
    if (nt && !verts_dev) { ctx->set_error("scene: verts == NULL"); return PRT_ERR_INVALID; }
    if (nt >= (1u << 28)) { ctx->set_error("scene: at most 2^28-1 triangles"); return PRT_ERR_INVALID; }
    return set_scene_common(ctx, verts_dev, nullptr, nullptr, nt, nullptr, 0, nullptr, 0, (cudaStream_t)stream);
}

int prt_bvh_build(prt_ctx* ctx, const prt_bvh_options* opts, prt_bvh_stats* stats) {
    CHECK_CTX(ctx);
    USE_DEVICE(ctx);
    return build_bvh(ctx, opts, stats);
}

int prt_camera_set(prt_ctx* ctx, const prt_camera* cam) {
    CHECK_CTX(ctx);
    if (!cam || cam->width == 0 || cam->height == 0) { ctx->set_error("camera: bad argument"); return PRT_ERR_INVALID; }
    ctx->cam = *cam;
    ctx->cam_set = true;
    return PRT_OK;
}

int prt_generate_rays(prt_ctx* ctx, uint64_t seed, uint32_t s0, uint32_t s1, int jitter, float tmin,
                      float tmax, prt_ray* rays_dev, void* stream) {
    CHECK_CTX(ctx);
    USE_DEVICE(ctx);
    if (!rays_dev) { ctx->set_error("generate_rays: rays == NULL"); return PRT_ERR_INVALID; }
    return generate_rays(ctx, seed, s0, s1, jitter, tmin, tmax, (float4*)rays_dev, (cudaStream_t)stream);
}

static int trace_common(prt_ctx* ctx, int mode, const prt_ray* rays, uint64_t n, void* o0, void* o1,
                        uint32_t flags, void* stream) {
    CHECK_CTX(ctx);
    USE_DEVICE(ctx);
    if (n && (!rays || !o0 || (mode == 2 && !o1))) { ctx->set_error("trace: NULL buffer"); return PRT_ERR_INVALID; }
    return launch_trace(ctx, mode, (const float4*)rays, n, o0, o1, flags, (cudaStream_t)stream);
}

int prt_trace_closest(prt_ctx* ctx, const prt_ray* rays_dev, uint64_t n, prt_hit* hits_dev,
                      uint32_t flags, void* stream) {
    return trace_common(ctx, 0, rays_dev, n, hits_dev, nullptr, flags, stream);
}
int prt_trace_any(prt_ctx* ctx, const prt_ray* rays_dev, uint64_t n, uint8_t* occluded_dev,
                  uint32_t flags, void* stream) {
    return trace_common(ctx, 1, rays_dev, n, occluded_dev, nullptr, flags, stream);
}
int prt_trace_all(prt_ctx* ctx, const prt_ray* rays_dev, uint64_t n, uint32_t* counts_dev,
                  uint64_t* sums_dev, uint32_t flags, void* stream) {
    return trace_common(ctx, 2, rays_dev, n, counts_dev, sums_dev, flags, stream);
}

static int stage_reserve(prt_ctx* ctx, int which, size_t bytes) {
    if (ctx->stage_bytes[which] >= bytes) return PRT_OK;
    cudaFree(ctx->stage[which]);
    ctx->stage[which] = nullptr; ctx->stage_bytes[which] = 0;
    PRT_CUDA_TRY(ctx, cudaMalloc(&ctx->stage[which], bytes));
    ctx->stage_bytes[which] = bytes;
    return PRT_OK;
}

// Host-buffer closest hit: rays are cut into chunks and pipelined over two streams so that
// the H2D copy of chunk k+1 and the D2H copy of chunk k-1 overlap the traversal of chunk k.
int prt_trace_closest_host(prt_ctx* ctx, const prt_ray* rays_host, uint64_t n, prt_hit* hits_host,
                           uint32_t flags) {
    CHECK_CTX(ctx);
    USE_DEVICE(ctx);
    if (n == 0) return PRT_OK;
    if (!rays_host || !hits_host) { ctx->set_error("trace_host: NULL buffer"); return PRT_ERR_INVALID; }
    // Upload, download and two alternating trace streams over kHostSlots staging slots: the upload
    // of chunk i+1.., the trace of chunk i and the download of chunk i-1 overlap, each copy engine
    // sees a FIFO, and because consecutive traces sit on different streams the CTAs of chunk i+1
    // move into the SM slots that the draining tail of chunk i frees (a persistent launch ends with
    // a few long rays on mostly idle SMs).  EXACT launches take their flag lists from a ring of
    // prt_ctx::kFlagRing, so they pipeline the same way.
    constexpr int S = prt_ctx::kHostSlots;
    const uint64_t chunk = 1ull << 21;
    const uint64_t cap = n < S * chunk ? n : S * chunk;
    int rc = stage_reserve(ctx, 0, sizeof(prt_ray) * cap);
    if (rc == PRT_OK) rc = stage_reserve(ctx, 1, sizeof(prt_hit) * cap);
    if (rc != PRT_OK) return rc;
    for (auto& st : ctx->copy_stream)
        if (!st) PRT_CUDA_TRY(ctx, cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    for (auto& ev : ctx->copy_event)
        if (!ev) PRT_CUDA_TRY(ctx, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    prt_ray* dr = (prt_ray*)ctx->stage[0];
    prt_hit* dh = (prt_hit*)ctx->stage[1];
    cudaStream_t s_up = ctx->copy_stream[0], s_down = ctx->copy_stream[2];
    uint64_t done = 0;
    for (uint64_t i = 0; done < n; ++i) {
        const int slot = (int)(i % S);
        cudaStream_t s_tr = ctx->copy_stream[(i & 1) == 0 ? 1 : 3];
        cudaEvent_t uploaded = ctx->copy_event[3 * slot], traced = ctx->copy_event[3 * slot + 1],
                    downloaded = ctx->copy_event[3 * slot + 2];
        const uint64_t m = n - done < chunk ? n - done : chunk;
        const uint64_t off = (cap == n) ? done : (uint64_t)slot * chunk;
        if (i >= (uint64_t)S) PRT_CUDA_TRY(ctx, cudaStreamWaitEvent(s_up, traced, 0));  // ray slot read by its last trace
        PRT_CUDA_TRY(ctx, cudaMemcpyAsync(dr + off, rays_host + done, sizeof(prt_ray) * m, cudaMemcpyHostToDevice, s_up));
        PRT_CUDA_TRY(ctx, cudaEventRecord(uploaded, s_up));
        PRT_CUDA_TRY(ctx, cudaStreamWaitEvent(s_tr, uploaded, 0));
        if (i >= (uint64_t)S) PRT_CUDA_TRY(ctx, cudaStreamWaitEvent(s_tr, downloaded, 0));  // hit slot drained
        rc = launch_trace(ctx, 0, (const float4*)(dr + off), m, dh + off, nullptr, flags, s_tr, (unsigned)slot);
        if (rc != PRT_OK) break;
        PRT_CUDA_TRY(ctx, cudaEventRecord(traced, s_tr));
        PRT_CUDA_TRY(ctx, cudaStreamWaitEvent(s_down, traced, 0));
        PRT_CUDA_TRY(ctx, cudaMemcpyAsync(hits_host + done, dh + off, sizeof(prt_hit) * m, cudaMemcpyDeviceToHost, s_down));
        PRT_CUDA_TRY(ctx, cudaEventRecord(downloaded, s_down));
        done += m;
    }
    cudaError_t err = cudaSuccess;
    for (auto& st : ctx->copy_stream) {
        const cudaError_t e = cudaStreamSynchronize(st);
        if (err == cudaSuccess) err = e;
    }
    if (rc != PRT_OK) return rc;
    if (err != cudaSuccess) {
        ctx->set_error("trace_host: %s", cudaGetErrorString(err));
        return PRT_ERR_CUDA;
    }
    return PRT_OK;
}

int prt_render(prt_ctx* ctx, const prt_render_params* params, float* accum_dev, int32_t* prim_ids_dev,
               void* stream) {
    CHECK_CTX(ctx);
    USE_DEVICE(ctx);
    if (!params || !accum_dev) { ctx->set_error("render: NULL argument"); return PRT_ERR_INVALID; }
    return render(ctx, params, accum_dev, prim_ids_dev, (cudaStream_t)stream);
}

int prt_trace_paths(prt_ctx* ctx, const prt_ray* rays_dev, uint64_t n, const prt_render_params* params,
                    float* radiance_dev, int32_t* prim_ids_dev, void* stream) {
    CHECK_CTX(ctx);
    USE_DEVICE(ctx);
    if (n == 0) return PRT_OK;
    if (!params || !rays_dev || !radiance_dev) { ctx->set_error("trace_paths: NULL argument"); return PRT_ERR_INVALID; }
    return render(ctx, params, radiance_dev, prim_ids_dev, (cudaStream_t)stream, (const float4*)rays_dev, n);
}

int prt_render_host(prt_ctx* ctx, const prt_render_params* params, float* accum_host) {
    CHECK_CTX(ctx);
    USE_DEVICE(ctx);
    if (!params || !accum_host) { ctx->set_error("render: NULL argument"); return PRT_ERR_INVALID; }
    if (!ctx->cam_set) { ctx->set_error("render: camera not set"); return PRT_ERR_STATE; }
    const size_t bytes = sizeof(float) * 4 * (size_t)ctx->cam.width * ctx->cam.height;
    int rc = stage_reserve(ctx, 0, bytes);  // grow-only device staging, reused across calls
    if (rc != PRT_OK) return rc;
    if (!ctx->copy_stream[0]) PRT_CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream[0], cudaStreamNonBlocking));
    cudaStream_t s = ctx->copy_stream[0];
    float* d = (float*)ctx->stage[0];
    // synchronous entry point: earlier asynchronous work on other streams may still use the
    // context's wavefront buffers
    PRT_CUDA_TRY(ctx, cudaDeviceSynchronize());
    PRT_CUDA_TRY(ctx, cudaMemcpyAsync(d, accum_host, bytes, cudaMemcpyHostToDevice, s));
    rc = render(ctx, params, d, nullptr, s);
    if (rc != PRT_OK) return rc;
    PRT_CUDA_TRY(ctx, cudaMemcpyAsync(accum_host, d, bytes, cudaMemcpyDeviceToHost, s));
    PRT_CUDA_TRY(ctx, cudaStreamSynchronize(s));
    return PRT_OK;
}

int prt_set_path_log(prt_ctx* ctx, prt_segment* segments_dev, uint64_t capacity, uint32_t* count_dev) {
    CHECK_CTX(ctx);
    if (segments_dev && (!count_dev || capacity == 0 || capacity > 0xffffffffull)) {
        ctx->set_error("path log: need a counter and 0 < capacity < 2^32");
        return PRT_ERR_INVALID;
    }
    ctx->log_segments = (float4*)segments_dev;
    ctx->log_count = segments_dev ? count_dev : nullptr;
    ctx->log_capacity = segments_dev ? (uint32_t)capacity : 0u;
    return PRT_OK;
}

int prt_eval_specular(prt_ctx* ctx, const prt_bsdf_query* queries_dev, uint64_t n, float* wi_valid_dev, void* stream) {
    CHECK_CTX(ctx);
    USE_DEVICE(ctx);
    if (n && (!queries_dev || !wi_valid_dev)) { ctx->set_error("eval_specular: NULL buffer"); return PRT_ERR_INVALID; }
    return eval_specular(ctx, queries_dev, n, wi_valid_dev, (cudaStream_t)stream);
}

int prt_release_scratch(prt_ctx* ctx) {
    CHECK_CTX(ctx);
    USE_DEVICE(ctx);
    PRT_CUDA_TRY(ctx, cudaDeviceSynchronize());
    cudaFree(ctx->build_arena);
    ctx->build_arena = nullptr; ctx->build_arena_bytes = 0;
    for (int k = 0; k < 2; ++k) { cudaFree(ctx->stage[k]); ctx->stage[k] = nullptr; ctx->stage_bytes[k] = 0; }
    cudaFree(ctx->shard_accum);
    ctx->shard_accum = nullptr; ctx->shard_bytes = 0;
    for (unsigned k = 0; k < prt_ctx::kFlagRing; ++k) { cudaFree(ctx->flag_list[k]); ctx->flag_list[k] = nullptr; ctx->flag_cap[k] = 0; }
    for (int k = 0; k < 4; ++k) { cudaFree(ctx->bin_scratch[k]); ctx->bin_scratch[k] = nullptr; ctx->bin_scratch_bytes[k] = 0; }
    wavefront_free(ctx);
    return PRT_OK;
}

int prt_set_wave_paths(prt_ctx* ctx, uint64_t paths) {
    CHECK_CTX(ctx);
    if (paths) ctx->wave_paths = paths;
    return PRT_OK;
}

int prt_profile_begin(prt_ctx* ctx) {
    CHECK_CTX(ctx);
    ctx->prof_class.clear();
    ctx->prof_launches.clear();
    ctx->prof_on = true;
    return PRT_OK;
}

int prt_profile_end(prt_ctx* ctx, prt_kernel_times* out) {
    CHECK_CTX(ctx);
    USE_DEVICE(ctx);
    if (!out) { ctx->set_error("profile_end: out == NULL"); return PRT_ERR_INVALID; }
    ctx->prof_on = false;
    memset(out, 0, sizeof *out);
    PRT_CUDA_TRY(ctx, cudaDeviceSynchronize());
    for (size_t i = 0; i < ctx->prof_class.size(); ++i) {
        float ms = 0.f;
        PRT_CUDA_TRY(ctx, cudaEventElapsedTime(&ms, ctx->prof_events[2 * i], ctx->prof_events[2 * i + 1]));
        const int c = ctx->prof_class[i];
        out->ms[c] += ms;
        out->launches[c] += (uint32_t)ctx->prof_launches[i];
    }
    ctx->prof_class.clear();
    ctx->prof_launches.clear();
    return PRT_OK;
}

int prt_get_counters(prt_ctx* ctx, prt_counters* out) {
    CHECK_CTX(ctx);
    USE_DEVICE(ctx);
    if (!out) { ctx->set_error("counters: out == NULL"); return PRT_ERR_INVALID; }
    Counters c;
    PRT_CUDA_TRY(ctx, cudaDeviceSynchronize());
    PRT_CUDA_TRY(ctx, cudaMemcpy(&c, ctx->counters, sizeof c, cudaMemcpyDeviceToHost));
    out->rays_closest = c.rays_closest; out->rays_shadow = c.rays_shadow;
    out->node_visits = c.node_visits; out->tri_tests = c.tri_tests;
    out->flagged_rays = c.flagged_rays; out->paths = c.paths;
    out->warp_iters = c.warp_iters; out->node_lane_iters = c.node_lane_iters;
    out->leaf_phases = c.leaf_phases; out->leaf_lane_phases = c.leaf_lane_phases;
    out->f64_decisions = c.f64_decisions;
    return PRT_OK;
}

int prt_reset_counters(prt_ctx* ctx) {
    CHECK_CTX(ctx);
    USE_DEVICE(ctx);
    PRT_CUDA_TRY(ctx, cudaMemset(ctx->counters, 0, sizeof(Counters)));
    return PRT_OK;
}

int prt_synchronize(prt_ctx* ctx) {
    CHECK_CTX(ctx);
    USE_DEVICE(ctx);
    PRT_CUDA_TRY(ctx, cudaDeviceSynchronize());
    return PRT_OK;
}

}  // extern "C"
