// traverse.cu -- API-level trace kernels (prt_trace_closest / _any / _all).
//
// Closest / any hit over the BVH run on the persistent-warp kernel (persist.cuh), in plain FP32 or
// -- PRT_TRACE_EXACT -- with error bounds: rays whose FP32 decisions were within their bound are
// appended to a flag list and re-traced by resolve_kernel in FP64 with the reference's operation
// order (traverse.cuh / intersect.cuh); finalize_kernel gives every unflagged winner the
// reference's own (t, u, v).  All-hits and brute-force modes use the one-ray-per-thread
// trace_kernel (parity paths, not timed).
#include "context.cuh"
#include "persist.cuh"
#include "traverse.cuh"

namespace prt {


template <int MODE, bool EXACT, bool COUNT, bool BRUTE>
__global__ void __launch_bounds__(kTraceThreads)
trace_kernel(SceneDev sc, const float4* __restrict__ rays, uint64_t n, void* out0, void* out1,
             uint32_t* flag_list, unsigned int* flag_count, Counters* ctr) {
    __shared__ uint2 s_stack[kPStack][kTraceThreads];
    uint64_t i = (uint64_t)blockIdx.x * kTraceThreads + threadIdx.x;
    TraceResult res;
    res.n_nodes = 0; res.n_tris = 0;
    bool active = i < n;
    if (active) {
        float4 ro = __ldg(rays + 2 * i), rd = __ldg(rays + 2 * i + 1);
        trace_one<MODE, EXACT, COUNT, BRUTE>(sc, ro, rd, &s_stack[0][threadIdx.x], res);
        if (EXACT && MODE == MODE_CLOSEST && !res.uncertain && res.gid >= 0) {
            // the winner is certain; report its (t,u,v) from the reference's FP64 formula so
            // that t is the oracle's t rounded to f32 (FP32 watertight t degrades when grazing)
            const float4* tp = sc.verts_gid + 3ull * res.gid;
            double o[3] = {(double)ro.x, (double)ro.y, (double)ro.z};
            double d[3] = {(double)rd.x, (double)rd.y, (double)rd.z};
            double t, u, v;
            if (mt_f64(xyz(__ldg(tp)), xyz(__ldg(tp + 1)), xyz(__ldg(tp + 2)), o, d, -1e300, 1e300, t, u, v)) {
                res.t = (float)t; res.u = (float)u; res.v = (float)v;
            } else {
                res.uncertain = true;
            }
        }
        if (EXACT && res.uncertain) {
            unsigned int slot = atomicAdd(flag_count, 1u);
            flag_list[slot] = (uint32_t)i;
        }
        if (MODE == MODE_CLOSEST) {
            prt_hit h;
            h.t = res.gid >= 0 ? res.t : 0.0f; h.u = res.u; h.v = res.v; h.tri = res.gid;
            reinterpret_cast<float4*>(out0)[i] =
                make_float4(h.t, h.u, h.v, __int_as_float(h.tri));
        } else if (MODE == MODE_ANY) {
            reinterpret_cast<uint8_t*>(out0)[i] = res.gid >= 0 ? 1 : 0;
        } else {
            reinterpret_cast<uint32_t*>(out0)[i] = res.count;
            reinterpret_cast<unsigned long long*>(out1)[i] = res.sum;
        }
    }
    if (COUNT) {
        if (!active) { res.n_nodes = 0; res.n_tris = 0; }
        TraceResult r2 = res;
        unsigned long long nn = r2.n_nodes, nt = r2.n_tris, one = active ? 1 : 0;
        for (int o = 16; o > 0; o >>= 1) {
            nn += __shfl_down_sync(0xffffffffu, nn, o);
            nt += __shfl_down_sync(0xffffffffu, nt, o);
            one += __shfl_down_sync(0xffffffffu, one, o);
        }
        if ((threadIdx.x & 31) == 0) {
            atomicAdd(&ctr->node_visits, nn);
            atomicAdd(&ctr->tri_tests, nt);
            atomicAdd(MODE == MODE_ANY ? &ctr->rays_shadow : &ctr->rays_closest, one);
        }
    }
}

// Throughput path (plain FP32, BVH): persistent warps with dynamic ray fetch (persist.cuh).
// Ray binning (PRT_TRACE_BIN): 8-bit key = cell of the ray origin in an 8 x 8 x 4 grid over the scene box
// (origins outside are clamped onto it), and ONE counting-sort pass over ray indices gives the order in which
// the persistent warps fetch the rays.  Three launches over tiles of 4096 rays:
//   bin_count_kernel    key per ray (stored as a byte), per-tile histogram in shared memory -> global bin counts
//   bin_scan_kernel     exclusive scan of the 256 counts -> bin cursors
//   bin_scatter_kernel  rank inside the tile by shared-memory atomics, one global reservation per bin per tile
// The order inside a bin is whatever the atomics give (it only changes which warp traces which ray).
constexpr int kBinTile = 4096, kBinThreads = 256;
__global__ void __launch_bounds__(kBinThreads)
bin_count_kernel(const float4* __restrict__ rays, unsigned int n, float3 lo, float3 scale, uint8_t* keys, unsigned int* hist) {
    __shared__ unsigned int cnt[256];
    cnt[threadIdx.x] = 0;
    __syncthreads();
    const unsigned int base = blockIdx.x * kBinTile;
#pragma unroll 4
    for (int r = 0; r < kBinTile / kBinThreads; ++r) {
        const unsigned int i = base + r * kBinThreads + threadIdx.x;
        if (i < n) {
            const float4 o = __ldg(rays + 2ull * i);
            const int x = min(max((int)((o.x - lo.x) * scale.x), 0), 7), y = min(max((int)((o.y - lo.y) * scale.y), 0), 7),
                      z = min(max((int)((o.z - lo.z) * scale.z), 0), 3);
            const unsigned int k = (unsigned int)((z << 6) | (y << 3) | x);
            keys[i] = (uint8_t)k;
            atomicAdd(&cnt[k], 1u);
        }
    }
    __syncthreads();
    if (cnt[threadIdx.x]) atomicAdd(&hist[threadIdx.x], cnt[threadIdx.x]);
}
__global__ void __launch_bounds__(256) bin_scan_kernel(unsigned int* hist_to_cursor) {
    __shared__ unsigned int s[256];
    const unsigned int v = hist_to_cursor[threadIdx.x];
    s[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 256; o <<= 1) {
        const unsigned int t = threadIdx.x >= o ? s[threadIdx.x - o] : 0u;
        __syncthreads();
        s[threadIdx.x] += t;
        __syncthreads();
    }
    hist_to_cursor[threadIdx.x] = s[threadIdx.x] - v;
}
__global__ void __launch_bounds__(kBinThreads)
bin_scatter_kernel(const uint8_t* __restrict__ keys, unsigned int n, unsigned int* cursor, uint32_t* perm) {
    __shared__ unsigned int cnt[256];
    cnt[threadIdx.x] = 0;
    __syncthreads();
    const unsigned int base = blockIdx.x * kBinTile;
    unsigned int key[kBinTile / kBinThreads], rank[kBinTile / kBinThreads];
#pragma unroll
    for (int r = 0; r < kBinTile / kBinThreads; ++r) {
        const unsigned int i = base + r * kBinThreads + threadIdx.x;
        key[r] = 256u;
        if (i < n) {
            key[r] = keys[i];
            rank[r] = atomicAdd(&cnt[key[r]], 1u);
        }
    }
    __syncthreads();
    const unsigned int c = cnt[threadIdx.x];
    __syncthreads();
    cnt[threadIdx.x] = c ? atomicAdd(&cursor[threadIdx.x], c) : 0u;  // this tile's range of the bin
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kBinTile / kBinThreads; ++r)
        if (key[r] < 256u) perm[cnt[key[r]] + rank[r]] = base + r * kBinThreads + threadIdx.x;
}

template <int MODE>
struct ApiIO {
    const float4* rays;
    void* out;
    const uint32_t* perm;      // ray order of this launch (binning), or nullptr = as given
    uint32_t* flag_list;       // EXACT only
    unsigned int* flag_count;
    bool aligned32;  // ray array on a 32-byte boundary (any cudaMalloc'd / torch buffer): one 256-bit load per ray
    __device__ __forceinline__ void load(unsigned k, float4& ro, float4& rd, uint32_t& tag) const {
        if (perm) k = __ldcs(perm + k);
        // rays and hits stream through once: evict-first, so they do not push the BVH out of L2
        if (aligned32) {
            ldg256_cs(rays + 2ull * k, ro, rd);
        } else {
            ro = __ldcs(rays + 2ull * k);
            rd = __ldcs(rays + 2ull * k + 1);
        }
        tag = k;
    }
    __device__ __forceinline__ void prefetch(unsigned k) const {
        if (!perm) asm volatile("prefetch.global.L2 [%0];" ::"l"(rays + 2ull * k));
    }
    __device__ __forceinline__ void store(uint32_t tag, float t, float u, float v, int gid) const {
        if (MODE == MODE_CLOSEST)
            __stcs(reinterpret_cast<float4*>(out) + tag, make_float4(gid >= 0 ? t : 0.0f, u, v, __int_as_float(gid)));
        else
            reinterpret_cast<uint8_t*>(out)[tag] = gid >= 0 ? 1 : 0;
    }
    __device__ __forceinline__ void reload(uint32_t tag, float4& ro, float4& rd) const {  // EXACT slow path
        ro = __ldg(rays + 2ull * tag);
        rd = __ldg(rays + 2ull * tag + 1);
    }
    // EXACT: the ray goes to the FP64 replay; until then its record reads "miss" (finalize_kernel skips it)
    __device__ __forceinline__ void flag(uint32_t tag) const {
        flag_list[atomicAdd(flag_count, 1u)] = tag;
        if (MODE == MODE_CLOSEST) reinterpret_cast<float4*>(out)[tag] = make_float4(0.f, 0.f, 0.f, __int_as_float(-1));
        else reinterpret_cast<uint8_t*>(out)[tag] = 0;
    }
};

// (gpurun r2-7, soup-1M exact: 8 / 7 / 6 blocks per SM = 64 / 72 / 80 registers -> 9.54 / 10.13 / 10.89 ms)
#ifndef PRT_MIN_BLOCKS_EXACT
#define PRT_MIN_BLOCKS_EXACT 8
#endif
template <int MODE, bool COUNT, bool EXACT>
__global__ void __launch_bounds__(kTraceThreads, EXACT ? PRT_MIN_BLOCKS_EXACT : PRT_MIN_BLOCKS)
trace_persistent_kernel(SceneDev sc, const float4* __restrict__ rays, unsigned int n, void* out, const uint32_t* perm,
                        unsigned int* fetch, uint32_t* flag_list, unsigned int* flag_count, Counters* ctr) {
    __shared__ uint2 s_stack[kPStack][kTraceThreads];
    ApiIO<MODE> io{rays, out, perm, flag_list, flag_count, (reinterpret_cast<uintptr_t>(rays) & 31u) == 0u};
    trace_persistent<MODE, COUNT, EXACT>(sc, io, fetch, n, &s_stack[0][threadIdx.x], ctr);
}

// EXACT closest hit, second pass (fully convergent, one ray per thread): the winner of an unflagged
// ray is certain, so its (t, u, v) are re-evaluated with the reference's FP64 formula -- t becomes
// the oracle's t rounded to f32 (the FP32 watertight t degrades for grazing rays).  A winner the
// FP64 formula rejects (cannot happen within the error bounds; kept as a guard) is flagged.
__global__ void __launch_bounds__(256)
finalize_kernel(SceneDev sc, const float4* __restrict__ rays, unsigned int n, float4* hits,
                uint32_t* flag_list, unsigned int* flag_count) {
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 h = hits[i];
    const int gid = __float_as_int(h.w);
    if (gid < 0) return;
    const float4 ro = __ldcs(rays + 2ull * i), rd = __ldcs(rays + 2ull * i + 1);
    const float4* tp = sc.verts_gid + 3ull * gid;
    const double o[3] = {(double)ro.x, (double)ro.y, (double)ro.z};
    const double d[3] = {(double)rd.x, (double)rd.y, (double)rd.z};
    double t, u, v;
    if (mt_f64(xyz(__ldg(tp)), xyz(__ldg(tp + 1)), xyz(__ldg(tp + 2)), o, d, -1e300, 1e300, t, u, v))
        hits[i] = make_float4((float)t, (float)u, (float)v, h.w);
    else
        flag_list[atomicAdd(flag_count, 1u)] = i;
}

template <int MODE, bool BRUTE>
__global__ void __launch_bounds__(kTraceThreads)
resolve_kernel(SceneDev sc, const float4* __restrict__ rays, void* out0, void* out1,
               const uint32_t* __restrict__ flag_list, const unsigned int* __restrict__ flag_count,
               Counters* ctr) {
    __shared__ uint2 s_stack[kPStack][kTraceThreads];
    unsigned int nf = *flag_count;
    for (unsigned int k = blockIdx.x * kTraceThreads + threadIdx.x; k < nf;
         k += gridDim.x * kTraceThreads) {
        uint32_t i = flag_list[k];
        float4 ro = __ldg(rays + 2ull * i), rd = __ldg(rays + 2ull * i + 1);
        TraceResult64 res;
        trace_one_f64<MODE, BRUTE>(sc, ro, rd, &s_stack[0][threadIdx.x], res);
        if (MODE == MODE_CLOSEST) {
            float t = res.gid >= 0 ? (float)res.t : 0.0f;
            reinterpret_cast<float4*>(out0)[i] =
                make_float4(t, (float)res.u, (float)res.v, __int_as_float(res.gid));
        } else if (MODE == MODE_ANY) {
            reinterpret_cast<uint8_t*>(out0)[i] = res.gid >= 0 ? 1 : 0;
        } else {
            reinterpret_cast<uint32_t*>(out0)[i] = res.count;
            reinterpret_cast<unsigned long long*>(out1)[i] = res.sum;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&ctr->flagged_rays, (unsigned long long)nf);
}

template <int MODE, bool EXACT, bool COUNT>
static void launch2(bool brute, dim3 grid, cudaStream_t s, SceneDev sc, const float4* rays,
                    uint64_t n, void* o0, void* o1, uint32_t* fl, unsigned int* fc, Counters* c) {
    if (brute)
        trace_kernel<MODE, EXACT, COUNT, true><<<grid, kTraceThreads, 0, s>>>(sc, rays, n, o0, o1, fl, fc, c);
    else
        trace_kernel<MODE, EXACT, COUNT, false><<<grid, kTraceThreads, 0, s>>>(sc, rays, n, o0, o1, fl, fc, c);
}

template <int MODE>
static void launch1(bool exact, bool count, bool brute, dim3 grid, cudaStream_t s, SceneDev sc,
                    const float4* rays, uint64_t n, void* o0, void* o1, uint32_t* fl,
                    unsigned int* fc, Counters* c) {
    if (exact) {
        if (count) launch2<MODE, true, true>(brute, grid, s, sc, rays, n, o0, o1, fl, fc, c);
        else launch2<MODE, true, false>(brute, grid, s, sc, rays, n, o0, o1, fl, fc, c);
    } else {
        if (count) launch2<MODE, false, true>(brute, grid, s, sc, rays, n, o0, o1, fl, fc, c);
        else launch2<MODE, false, false>(brute, grid, s, sc, rays, n, o0, o1, fl, fc, c);
    }
}

int launch_trace(prt_ctx* ctx, int mode, const float4* rays, uint64_t n, void* out0, void* out1,
                 uint32_t flags, cudaStream_t stream, unsigned flag_slot) {
    if (n == 0) return PRT_OK;
    if (n > (1ull << 31)) { ctx->set_error("trace: n=%llu exceeds 2^31 rays per call", (unsigned long long)n); return PRT_ERR_INVALID; }
    bool exact = flags & PRT_TRACE_EXACT, count = flags & PRT_TRACE_COUNT, brute = flags & PRT_TRACE_BRUTE;
    if (!ctx->scene_set) { ctx->set_error("trace: no scene (call prt_scene_set_triangles first)"); return PRT_ERR_STATE; }
    if (!brute && !ctx->bvh_built) { ctx->set_error("trace: BVH not built (call prt_bvh_build or pass PRT_TRACE_BRUTE)"); return PRT_ERR_STATE; }
    SceneDev sc = ctx->scene_dev();
    uint32_t* flag_list = nullptr;
    unsigned int* flag_count = nullptr;
    if (exact) {  // grow-only flag list of this launch's ring slot (worst case: every ray flagged)
        const unsigned slot = flag_slot % prt_ctx::kFlagRing;
        if (ctx->flag_cap[slot] < n) {
            PRT_CUDA_TRY(ctx, cudaDeviceSynchronize());  // an earlier launch (any stream) may still use the old list
            cudaFree(ctx->flag_list[slot]);
            ctx->flag_list[slot] = nullptr; ctx->flag_cap[slot] = 0;
            PRT_CUDA_TRY(ctx, cudaMalloc(&ctx->flag_list[slot], n * sizeof(uint32_t)));
            ctx->flag_cap[slot] = n;
        }
        flag_list = ctx->flag_list[slot];
        flag_count = ctx->flag_count + slot;
        PRT_CUDA_TRY(ctx, cudaMemsetAsync(flag_count, 0, sizeof(unsigned int), stream));
    }
    if (!brute && mode != MODE_ALL) {
        // one fetch counter per in-flight launch (host-buffer calls pipeline two streams)
        unsigned int* fetch = ctx->fetch_counters + (ctx->fetch_next++ % prt_ctx::kFetchRing);
        PRT_CUDA_TRY(ctx, cudaMemsetAsync(fetch, 0, sizeof(unsigned int), stream));
        if (ctx->grid_persist == 0) {
            int b = 0;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, trace_persistent_kernel<MODE_CLOSEST, false, false>, kTraceThreads, 0);
            ctx->grid_persist = ctx->num_sms * (b > 0 ? b : 8);
        }
        if (exact && ctx->grid_persist_exact == 0) {
            int b = 0;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, trace_persistent_kernel<MODE_CLOSEST, false, true>, kTraceThreads, 0);
            ctx->grid_persist_exact = ctx->num_sms * (b > 0 ? b : 6);
        }
        unsigned g = (unsigned)(exact ? ctx->grid_persist_exact : ctx->grid_persist);
        unsigned need = (unsigned)((n + kTraceThreads - 1) / kTraceThreads);
        if (need < g) g = need;
        // ray binning: explicit, or by default when the BVH cannot live in L2 and the batch is large
        const size_t bvh_bytes = (size_t)ctx->n_nodes * sizeof(Node64) + (size_t)ctx->nt * 40u;
        const bool bin = !(flags & PRT_TRACE_NO_BIN) && ((flags & PRT_TRACE_BIN) || (bvh_bytes > (96u << 20) && n >= (1u << 20)));
        const uint32_t* perm = nullptr;
        if (bin) {
            const unsigned slot = flag_slot % 4u;
            const size_t nz = ((size_t)n + 255) & ~(size_t)255, need = 5 * nz + 1024;  // perm (u32), keys (u8), 256 cursors
            if (ctx->bin_scratch_bytes[slot] < need) {
                PRT_CUDA_TRY(ctx, cudaDeviceSynchronize());
                cudaFree(ctx->bin_scratch[slot]);
                ctx->bin_scratch[slot] = nullptr; ctx->bin_scratch_bytes[slot] = 0;
                PRT_CUDA_TRY(ctx, cudaMalloc(&ctx->bin_scratch[slot], need));
                ctx->bin_scratch_bytes[slot] = need;
            }
            uint32_t* perm_buf = (uint32_t*)ctx->bin_scratch[slot];
            uint8_t* keys = (uint8_t*)(perm_buf + nz);
            unsigned int* cursor = (unsigned int*)(keys + nz);
            const float3 lo = make_float3(ctx->scene_lo[0], ctx->scene_lo[1], ctx->scene_lo[2]);
            const float ex = ctx->scene_hi[0] - lo.x, ey = ctx->scene_hi[1] - lo.y, ez = ctx->scene_hi[2] - lo.z;
            const float3 scale = make_float3(ex > 0.f ? 8.0f / ex : 0.f, ey > 0.f ? 8.0f / ey : 0.f, ez > 0.f ? 4.0f / ez : 0.f);
            const unsigned tiles = (unsigned)((n + kBinTile - 1) / kBinTile);
            PRT_CUDA_TRY(ctx, cudaMemsetAsync(cursor, 0, 256 * sizeof(unsigned int), stream));
            prof_begin(ctx, PROF_OTHER, stream, 3);
            bin_count_kernel<<<tiles, kBinThreads, 0, stream>>>(rays, (unsigned)n, lo, scale, keys, cursor);
            bin_scan_kernel<<<1, 256, 0, stream>>>(cursor);
            bin_scatter_kernel<<<tiles, kBinThreads, 0, stream>>>(keys, (unsigned)n, cursor, perm_buf);
            prof_end(ctx, stream);
            perm = perm_buf;
        }
        prof_begin(ctx, mode == MODE_CLOSEST ? PROF_CLOSEST : PROF_SHADOW, stream);
#define PRT_PERSIST(M, C, E) trace_persistent_kernel<M, C, E><<<g, kTraceThreads, 0, stream>>>( \
        sc, rays, (unsigned)n, out0, perm, fetch, flag_list, flag_count, ctx->counters)
        if (mode == MODE_CLOSEST) {
            if (exact) { if (count) PRT_PERSIST(MODE_CLOSEST, true, true); else PRT_PERSIST(MODE_CLOSEST, false, true); }
            else { if (count) PRT_PERSIST(MODE_CLOSEST, true, false); else PRT_PERSIST(MODE_CLOSEST, false, false); }
        } else {
            if (exact) { if (count) PRT_PERSIST(MODE_ANY, true, true); else PRT_PERSIST(MODE_ANY, false, true); }
            else { if (count) PRT_PERSIST(MODE_ANY, true, false); else PRT_PERSIST(MODE_ANY, false, false); }
        }
#undef PRT_PERSIST
        prof_end(ctx, stream);
        PRT_CUDA_TRY(ctx, cudaGetLastError());
        if (exact) {
            const unsigned g2 = (unsigned)(ctx->num_sms * 4);
            prof_begin(ctx, PROF_EXACT_FIXUP, stream, mode == MODE_CLOSEST ? 2 : 1);
            if (mode == MODE_CLOSEST) {
                finalize_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(sc, rays, (unsigned)n, (float4*)out0,
                                                                                flag_list, flag_count);
                resolve_kernel<MODE_CLOSEST, false><<<g2, kTraceThreads, 0, stream>>>(sc, rays, out0, out1, flag_list, flag_count, ctx->counters);
            } else {
                resolve_kernel<MODE_ANY, false><<<g2, kTraceThreads, 0, stream>>>(sc, rays, out0, out1, flag_list, flag_count, ctx->counters);
            }
            prof_end(ctx, stream);
            PRT_CUDA_TRY(ctx, cudaGetLastError());
        }
        return PRT_OK;
    }
    dim3 grid((unsigned)((n + kTraceThreads - 1) / kTraceThreads));
    switch (mode) {
        case MODE_CLOSEST: launch1<MODE_CLOSEST>(exact, count, brute, grid, stream, sc, rays, n, out0, out1, flag_list, flag_count, ctx->counters); break;
        case MODE_ANY: launch1<MODE_ANY>(exact, count, brute, grid, stream, sc, rays, n, out0, out1, flag_list, flag_count, ctx->counters); break;
        default: launch1<MODE_ALL>(exact, count, brute, grid, stream, sc, rays, n, out0, out1, flag_list, flag_count, ctx->counters); break;
    }
    PRT_CUDA_TRY(ctx, cudaGetLastError());
    if (exact) {
        dim3 g2((unsigned)(ctx->num_sms * 4));
        if (mode == MODE_CLOSEST) {
            if (brute) resolve_kernel<MODE_CLOSEST, true><<<g2, kTraceThreads, 0, stream>>>(sc, rays, out0, out1, flag_list, flag_count, ctx->counters);
            else resolve_kernel<MODE_CLOSEST, false><<<g2, kTraceThreads, 0, stream>>>(sc, rays, out0, out1, flag_list, flag_count, ctx->counters);
        } else if (mode == MODE_ANY) {
            if (brute) resolve_kernel<MODE_ANY, true><<<g2, kTraceThreads, 0, stream>>>(sc, rays, out0, out1, flag_list, flag_count, ctx->counters);
            else resolve_kernel<MODE_ANY, false><<<g2, kTraceThreads, 0, stream>>>(sc, rays, out0, out1, flag_list, flag_count, ctx->counters);
        } else {
            if (brute) resolve_kernel<MODE_ALL, true><<<g2, kTraceThreads, 0, stream>>>(sc, rays, out0, out1, flag_list, flag_count, ctx->counters);
            else resolve_kernel<MODE_ALL, false><<<g2, kTraceThreads, 0, stream>>>(sc, rays, out0, out1, flag_list, flag_count, ctx->counters);
        }
        PRT_CUDA_TRY(ctx, cudaGetLastError());
    }
    return PRT_OK;
}

}  // namespace prt
