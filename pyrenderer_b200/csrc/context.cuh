// context.cuh -- the library-owned state behind prt_ctx.
#pragma once
#include <stdarg.h>

#include <string>
#include <vector>

#include "common.cuh"

struct prt_ctx {
    int device = 0;
    int num_sms = 148;
    char err[512] = {0};

    // scene (device)
    uint32_t nt = 0, nm = 0, nl = 0;
    float4* verts_gid = nullptr;   // [nt*3] triangles by global id (w unused)
    float4* shade = nullptr;       // [nt] normal.xyz, bits(material)
    prt_material* mats = nullptr;  // [nm]
    uint32_t* light_tris = nullptr;
    bool scene_set = false;

    // capacities (bytes) of the grow-only device buffers: a new scene or a rebuild of the same size
    // reuses them (no cudaMalloc / cudaFree on the set-scene / build path after the first call)
    size_t verts_bytes = 0, shade_bytes = 0, mats_bytes = 0, lights_bytes = 0;
    size_t tri_a_bytes = 0, tri_b_bytes = 0, nodes_bytes = 0;
    void* build_arena = nullptr;   // scratch of bvh_build.cu
    size_t build_arena_bytes = 0;
    cudaEvent_t build_ev[7] = {};  // phase timing of the build
    int grid_emit = 0;             // resident CTAs of the cooperative emit kernel

    // BVH (device)
    uint4* tri_a = nullptr;        // [nt*2] leaf order: p0.xyz, p1.xyz, p2.xy
    float2* tri_b = nullptr;       // [nt]   leaf order: p2.z, bits(global id)
    prt::Node64* nodes = nullptr;
    uint32_t n_nodes = 0;
    bool bvh_built = false;
    prt_bvh_stats bvh_stats = {};

    // camera
    prt_camera cam = {};
    bool cam_set = false;

    // ray binning (traverse.cu): scene box from the last build, per-slot sort scratch (keys x2, vals x2, radix temp)
    float scene_lo[3] = {0.f, 0.f, 0.f}, scene_hi[3] = {1.f, 1.f, 1.f};
    void* bin_scratch[4] = {};
    size_t bin_scratch_bytes[4] = {};

    // counters + exact-mode scratch: one (flag list, flag count) pair per in-flight EXACT launch,
    // so that the host-buffer pipeline can keep two exact traces on two streams
    prt::Counters* counters = nullptr;  // device
    static constexpr unsigned kFlagRing = 4;
    uint32_t* flag_list[kFlagRing] = {};
    unsigned int* flag_count = nullptr;  // [kFlagRing]
    uint64_t flag_cap[kFlagRing] = {};
    static constexpr unsigned kFetchRing = 32;
    unsigned int* fetch_counters = nullptr;  // [kFetchRing] ray-fetch counters of persistent launches
    unsigned fetch_next = 0;
    int grid_persist = 0, grid_persist_exact = 0;  // resident CTAs of the plain / EXACT persistent kernels
    int refill_idle = 0, leaf_batch = 0;  // 0 = by scene size (profiles/r1_sweeps.txt, r2_sweeps.txt)
    int fetch_chunk = 0;                  // ray indices reserved per atomic; 0 = default (32)

    // device staging of the *_host entry points (grow-only, reused across calls)
    void* stage[2] = {nullptr, nullptr};
    size_t stage_bytes[2] = {0, 0};
    static constexpr int kHostSlots = 4;  // chunks in flight in prt_trace_closest_host
    cudaStream_t copy_stream[4] = {nullptr, nullptr, nullptr, nullptr};  // upload, trace A, download, trace B
    cudaEvent_t copy_event[3 * kHostSlots] = {};                // per slot: uploaded, traced, downloaded

    // wavefront state (wavefront.cu)
    void* wf = nullptr;
    uint64_t wave_paths = 64ull << 20;  // 8.5 GiB of path state; Cornell 1024^2 per 16 spp: 12.68 ms (16 Mi) / 12.31 (32 Mi) / 12.15 (64 Mi) / 12.06 (128 Mi)
    // multi-GPU (collective.cu): communicator, this rank, and the per-frame shard buffer
    void* comm = nullptr;  // ncclComm_t
    bool comm_owned = false;
    int comm_world = 1, comm_rank = 0;
    float* shard_accum = nullptr;
    size_t shard_bytes = 0;

    // per-kernel-class timing (prt_profile_begin / _end): event pairs on the launching stream
    bool prof_on = false;
    std::vector<cudaEvent_t> prof_events;  // pool; pair i = events 2i, 2i+1
    std::vector<int> prof_class;           // class of pair i
    std::vector<int> prof_launches;        // kernel launches inside pair i

    // path-segment log (prt_set_path_log); caller-owned device memory
    float4* log_segments = nullptr;
    uint32_t* log_count = nullptr;
    uint32_t log_capacity = 0;

    void set_error(const char* fmt, ...) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(err, sizeof err, fmt, ap);
        va_end(ap);
    }

    prt::SceneDev scene_dev() const {
        prt::SceneDev s;
        s.tris = verts_gid;
        s.tri_a = bvh_built ? tri_a : nullptr;
        s.tri_b = bvh_built ? tri_b : nullptr;
        s.nodes = nodes;
        s.shade = shade;
        s.mats = mats;
        s.light_tris = light_tris;
        s.verts_gid = verts_gid;
        s.nt = nt; s.n_nodes = n_nodes; s.nl = nl; s.nm = nm;
        // short traversals (tiny scenes) amortise the ray set-up over more lanes per refill and wait for more
        // parked lanes before a leaf phase (Cornell 16-spp wave: refill / leaf 16 / 8 -> 12 / 16: 12.68 -> 12.50 ms;
        // on the 1M soup a leaf batch of 16 costs 9 %)
        const bool tiny = n_nodes < 4096;
        s.refill_idle = refill_idle > 0 ? refill_idle : (tiny ? 12 : 6);
        s.leaf_batch = leaf_batch > 0 ? leaf_batch : (tiny ? 16 : 8);
        // one warp's worth per atomic: larger chunks lengthen the end-of-launch tail (measured:
        // soup-1M 16..64 equal, 128 -0.6 %; Cornell 8 spp 32: 7.49 ms, 256: 7.99 ms, 1024: 13.4 ms)
        s.fetch_chunk = fetch_chunk > 0 ? fetch_chunk : 32;
        return s;
    }
};

namespace prt {
// kernel classes of prt_kernel_times (include/prt.h PRT_PROF_*)
enum { PROF_RAYGEN = 0, PROF_CLOSEST = 1, PROF_SHADE = 2, PROF_SHADOW = 3, PROF_EXACT_FIXUP = 4, PROF_OTHER = 5,
       PROF_ALLREDUCE = 6, PROF_N = 8 };
// prt_api.cu: bracket the next `launches` kernel launches on `stream` (no-ops unless profiling is on)
void prof_begin(prt_ctx* ctx, int cls, cudaStream_t stream, int launches = 1);
void prof_end(prt_ctx* ctx, cudaStream_t stream);
// traverse.cu
// flag_slot: which (flag list, flag count) pair an EXACT launch uses; launches that may overlap on
// different streams (the host-buffer pipeline) take different slots
int launch_trace(prt_ctx* ctx, int mode, const float4* rays, uint64_t n, void* out0, void* out1,
                 uint32_t flags, cudaStream_t stream, unsigned flag_slot = 0);
// bvh_build.cu
int build_bvh(prt_ctx* ctx, const prt_bvh_options* opts, prt_bvh_stats* stats);
// wavefront.cu
int generate_rays(prt_ctx* ctx, uint64_t seed, uint32_t s0, uint32_t s1, int jitter, float tmin,
                  float tmax, float4* rays, cudaStream_t stream);
int render(prt_ctx* ctx, const prt_render_params* p, float* accum, int32_t* prim_ids,
           cudaStream_t stream, const float4* user_rays = nullptr, uint64_t n_user = 0);
void wavefront_free(prt_ctx* ctx);
int eval_specular(prt_ctx* ctx, const prt_bsdf_query* q, uint64_t n, float* out, cudaStream_t stream);
}  // namespace prt
