// collective.cu -- the one exchange step of the path: summing the fp32 accumulation buffers of
// the ranks that rendered disjoint sample ranges of the same frame (the reference has
// no multi-process path at all: main.py:28-55 / main_taichi.py:80-99 render on one device).
//
// NCCL is bound at RUN time (dlopen "libnccl.so.2"), not at link time: inside a PyTorch process
// the soname resolves to the copy torch has already loaded, so both sides share one NCCL; a plain
// C / ctypes host gets the system library.  A process that never calls prt_comm_* never touches it.
#include <dlfcn.h>
#include <string.h>

#include "context.cuh"

namespace {

// the handful of NCCL declarations this file needs (nccl.h 2.x: stable ABI)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { kNcclSuccess = 0, kNcclFloat32 = 7, kNcclSum = 0 };

struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(ncclUniqueId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    int (*GetVersion)(int*) = nullptr;
    char err[256] = "";
};

NcclApi g_nccl;

const char* nccl_load() {  // nullptr on success, else the reason
    if (g_nccl.handle) return nullptr;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
        snprintf(g_nccl.err, sizeof g_nccl.err, "cannot load libnccl.so.2: %s", dlerror());
        return g_nccl.err;
    }
#define PRT_SYM(field, name)                                                             \
    *(void**)(&g_nccl.field) = dlsym(h, name);                                           \
    if (!g_nccl.field) {                                                                 \
        snprintf(g_nccl.err, sizeof g_nccl.err, "libnccl.so.2 has no symbol %s", name);  \
        dlclose(h);                                                                      \
        return g_nccl.err;                                                               \
    }
    PRT_SYM(GetUniqueId, "ncclGetUniqueId")
    PRT_SYM(CommInitRank, "ncclCommInitRank")
    PRT_SYM(CommDestroy, "ncclCommDestroy")
    PRT_SYM(AllReduce, "ncclAllReduce")
    PRT_SYM(GetErrorString, "ncclGetErrorString")
    PRT_SYM(GetVersion, "ncclGetVersion")
#undef PRT_SYM
    g_nccl.handle = h;
    return nullptr;
}

struct DeviceScope {  // same contract as prt_api.cu: run on the context's device, restore the caller's
    int prev = -1;
    bool switched = false;
    explicit DeviceScope(int dev) {
        if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) switched = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceScope() { if (switched) cudaSetDevice(prev); }
};

__global__ void add_into_kernel(const float4* __restrict__ src, float4* dst, uint64_t n4) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (uint64_t)gridDim.x * blockDim.x) {
        const float4 a = src[i];
        float4 b = dst[i];
        dst[i] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
    }
}

}  // namespace

#define NCCL_TRY(ctx, expr)                                                                   \
    do {                                                                                      \
        int _r = (expr);                                                                      \
        if (_r != kNcclSuccess) {                                                             \
            (ctx)->set_error("%s:%d %s -> NCCL: %s", __FILE__, __LINE__, #expr, g_nccl.GetErrorString(_r)); \
            return PRT_ERR_CUDA;                                                              \
        }                                                                                     \
    } while (0)

extern "C" {

int prt_comm_unique_id(void* id_out) {
    if (!id_out) return PRT_ERR_INVALID;
    if (nccl_load()) return PRT_ERR_STATE;
    ncclUniqueId id;
    if (g_nccl.GetUniqueId(&id) != kNcclSuccess) return PRT_ERR_CUDA;
    memcpy(id_out, &id, sizeof id);
    return PRT_OK;
}

int prt_comm_init(prt_ctx* ctx, const void* id, int world, int rank) {
    if (!ctx) return PRT_ERR_INVALID;
    if (!id || world < 1 || rank < 0 || rank >= world) { ctx->set_error("comm_init: bad argument (world %d, rank %d)", world, rank); return PRT_ERR_INVALID; }
    if (ctx->comm) { ctx->set_error("comm_init: this context already has a communicator"); return PRT_ERR_STATE; }
    if (const char* why = nccl_load()) { ctx->set_error("comm_init: %s", why); return PRT_ERR_STATE; }
    DeviceScope scope(ctx->device);
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof uid);
    ncclComm_t comm = nullptr;
    NCCL_TRY(ctx, g_nccl.CommInitRank(&comm, world, uid, rank));
    ctx->comm = comm;
    ctx->comm_owned = true;
    ctx->comm_world = world;
    ctx->comm_rank = rank;
    return PRT_OK;
}

int prt_comm_attach(prt_ctx* ctx, void* nccl_comm, int world, int rank) {
    if (!ctx) return PRT_ERR_INVALID;
    if (!nccl_comm || world < 1 || rank < 0 || rank >= world) { ctx->set_error("comm_attach: bad argument"); return PRT_ERR_INVALID; }
    if (ctx->comm) { ctx->set_error("comm_attach: this context already has a communicator"); return PRT_ERR_STATE; }
    if (const char* why = nccl_load()) { ctx->set_error("comm_attach: %s", why); return PRT_ERR_STATE; }
    ctx->comm = nccl_comm;
    ctx->comm_owned = false;
    ctx->comm_world = world;
    ctx->comm_rank = rank;
    return PRT_OK;
}

int prt_comm_destroy(prt_ctx* ctx) {
    if (!ctx) return PRT_ERR_INVALID;
    if (ctx->comm && ctx->comm_owned) {
        DeviceScope scope(ctx->device);
        cudaDeviceSynchronize();
        g_nccl.CommDestroy((ncclComm_t)ctx->comm);
    }
    ctx->comm = nullptr;
    ctx->comm_owned = false;
    ctx->comm_world = 1;
    ctx->comm_rank = 0;
    return PRT_OK;
}

int prt_comm_info(const prt_ctx* ctx, int* world, int* rank, int* nccl_version) {
    if (!ctx) return PRT_ERR_INVALID;
    if (world) *world = ctx->comm_world;
    if (rank) *rank = ctx->comm_rank;
    if (nccl_version) {
        *nccl_version = 0;
        if (g_nccl.handle) g_nccl.GetVersion(nccl_version);
    }
    return PRT_OK;
}

int prt_allreduce_sum(prt_ctx* ctx, float* buf_dev, uint64_t n, void* stream) {
    if (!ctx) return PRT_ERR_INVALID;
    if (n == 0) return PRT_OK;
    if (!buf_dev) { ctx->set_error("allreduce: NULL buffer"); return PRT_ERR_INVALID; }
    if (!ctx->comm) {
        if (ctx->comm_world == 1) return PRT_OK;  // single rank: the sum over ranks is the buffer itself
        ctx->set_error("allreduce: no communicator (call prt_comm_init or prt_comm_attach)");
        return PRT_ERR_STATE;
    }
    DeviceScope scope(ctx->device);
    NCCL_TRY(ctx, g_nccl.AllReduce(buf_dev, buf_dev, (size_t)n, kNcclFloat32, kNcclSum, (ncclComm_t)ctx->comm, (cudaStream_t)stream));
    return PRT_OK;
}

// One frame of SURVEY 8e in one call: rank r of G renders samples [b + r*S/G, b + (r+1)*S/G) of
// every pixel into a library-owned zeroed buffer, ONE all-reduce sums the shards over NVLink, and
// the sum is added to the caller's accumulation buffer (so progressive calls compose).
int prt_render_sharded(prt_ctx* ctx, const prt_render_params* params, float* accum_dev, void* stream) {
    if (!ctx) return PRT_ERR_INVALID;
    if (!params || !accum_dev) { ctx->set_error("render_sharded: NULL argument"); return PRT_ERR_INVALID; }
    if (!ctx->cam_set) { ctx->set_error("render_sharded: camera not set"); return PRT_ERR_STATE; }
    if (params->spp_end < params->spp_begin) { ctx->set_error("render_sharded: spp_end < spp_begin"); return PRT_ERR_INVALID; }
    DeviceScope scope(ctx->device);
    cudaStream_t s = (cudaStream_t)stream;
    const uint64_t n = 4ull * ctx->cam.width * ctx->cam.height;
    if (ctx->shard_bytes < n * sizeof(float)) {
        cudaFree(ctx->shard_accum);
        ctx->shard_accum = nullptr; ctx->shard_bytes = 0;
        PRT_CUDA_TRY(ctx, cudaMalloc(&ctx->shard_accum, n * sizeof(float)));
        ctx->shard_bytes = n * sizeof(float);
    }
    const uint64_t S = params->spp_end - params->spp_begin, G = (uint64_t)ctx->comm_world, r = (uint64_t)ctx->comm_rank;
    prt_render_params p = *params;
    p.spp_begin = params->spp_begin + (uint32_t)(S * r / G);
    p.spp_end = params->spp_begin + (uint32_t)(S * (r + 1) / G);
    PRT_CUDA_TRY(ctx, cudaMemsetAsync(ctx->shard_accum, 0, n * sizeof(float), s));
    int rc = prt::render(ctx, &p, ctx->shard_accum, nullptr, s);
    if (rc != PRT_OK) return rc;
    prt::prof_begin(ctx, prt::PROF_ALLREDUCE, s);
    rc = prt_allreduce_sum(ctx, ctx->shard_accum, n, stream);
    prt::prof_end(ctx, s);
    if (rc != PRT_OK) return rc;
    add_into_kernel<<<ctx->num_sms * 4, 256, 0, s>>>((const float4*)ctx->shard_accum, (float4*)accum_dev, n / 4);
    PRT_CUDA_TRY(ctx, cudaGetLastError());
    return PRT_OK;
}

}  // extern "C"
