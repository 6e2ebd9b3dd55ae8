// common.cuh -- shared types of the sm_100a path-tracing core.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/prt.h"

namespace prt {

constexpr int kTraceThreads = 128;      // CTA size of the traversal kernels
constexpr uint32_t kLeafFlag = 0x80000000u;
constexpr uint32_t kNoChild = 0xFFFFFFFFu;   // absent child slot; also the "traversal finished" marker

// ---- device scene ---------------------------------------------------------
// Traversal reads triangles in BVH leaf order (a leaf's triangles are contiguous) from two
// arrays: 32 B (p0.xyz, p1.xyz, p2.xy) + 8 B (p2.z, bits(global id)) -- 40 B and exactly two
// L1 sectors per test (the former 3 x float4 record cost three: the data pipe moves one
// 32-byte sector per cycle for divergent lanes, and that pipe is what bounds the traversal).
// The global-id-order copy (3 x float4: (p0, bits(gid)), (p1, bits(material)), (p2, 0)) serves
// brute force, light sampling and hit-point reconstruction.  Shading data (normal + material)
// is a separate float4 array indexed by the GLOBAL id, touched once per path vertex only.
// 4-wide BVH node, 64 B = four float4 = two 32-byte sectors: 16 B per child box + reference,
// the same bytes-per-child as a 32-byte 2-wide node (see DESIGN.md 2 for why 4-wide).
struct Node64 {
    float ox, oy, oz;   // frame origin (node box min)
    float sx, sy, sz;   // plane = o + q * s, s = 2^k per axis (stored as the float itself: no decode)
    uint32_t qx_lo, qx_hi, qy_lo, qy_hi, qz_lo, qz_hi;  // byte c = child c : quantised lo / hi planes per axis
    uint32_t ref[4];    // child refs: record index, or kLeafFlag | first_tri<<3 | count, or kNoChild
};
static_assert(sizeof(Node64) == 64, "node must be 64 bytes");

struct SceneDev {
    const float4* tris;      // [nt*3] GLOBAL-id order, 3 x float4 (brute-force modes only)
    const uint4* tri_a;      // [nt*2] leaf order, 32 B: p0.xyz, p1.xyz, p2.xy   (one LDG.256 = one sector)
    const float2* tri_b;     // [nt]   leaf order,  8 B: p2.z, bits(global id)
    const Node64* nodes;     // [n_nodes]
    const float4* shade;     // [nt] by global id: normal.xyz, bits(material)
    const prt_material* mats;
    const uint32_t* light_tris;  // global ids
    const float4* verts_gid;     // [nt*3] by global id (light sampling)
    uint32_t nt, n_nodes, nl, nm;
    int refill_idle, leaf_batch, fetch_chunk;  // persist.cuh scheduling knobs
};

struct Counters {
    unsigned long long rays_closest, rays_shadow, node_visits, tri_tests, flagged_rays, paths;
    unsigned long long warp_iters, node_lane_iters, leaf_phases, leaf_lane_phases;  // persist.cuh utilisation (COUNT)
    unsigned long long f64_decisions;  // EXACT + COUNT: triangle tests decided in place in FP64
};

// ---- small vector helpers ---------------------------------------------------
__device__ __forceinline__ float3 operator+(float3 a, float3 b) { return make_float3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ float3 operator-(float3 a, float3 b) { return make_float3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ float3 operator*(float3 a, float s) { return make_float3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ float3 operator*(float3 a, float3 b) { return make_float3(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ float3 operator-(float3 a) { return make_float3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ float dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ float3 cross(float3 a, float3 b) {
    return make_float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ float3 normalize(float3 a) {
    float n = sqrtf(a.x * a.x + a.y * a.y + a.z * a.z);
    return make_float3(a.x / n, a.y / n, a.z / n);
}
// shading-only variant: one MUFU.RSQ (<= 2 ulp) instead of sqrt + three IEEE divisions
__device__ __forceinline__ float3 normalize_fast(float3 a) {
    float inv = rsqrtf(a.x * a.x + a.y * a.y + a.z * a.z);
    return make_float3(a.x * inv, a.y * inv, a.z * inv);
}
__device__ __forceinline__ float3 xyz(float4 a) { return make_float3(a.x, a.y, a.z); }

}  // namespace prt

#define PRT_CUDA_TRY(ctx, expr)                                                          \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess) {                                                         \
            (ctx)->set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr,                \
                             cudaGetErrorString(_e));                                    \
            return PRT_ERR_CUDA;                                                         \
        }                                                                                \
    } while (0)
