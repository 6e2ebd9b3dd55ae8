// bvh.cuh -- the 32-byte quantised node and the traversal loop.
//
// Node32 (common.cuh) holds the boxes of its TWO children, quantised to 8 bits
// per plane in the node's own frame: plane = o + q * 2^(e-127).  The builder
// (bvh_build.cu) rounds lo down / hi up against this exact decode expression,
// so decoded boxes always contain the child's true FP32 box.  The slab test
// never decodes the box: per axis it forms a = scale/d and b = (o - ray.o)/d once
// and each of the 12 planes costs one cvt + one FMA (t = q*a + b).
//
// Records are stored in DFS pre-order, triangles in DFS leaf order, which lets
// one 32-bit `link` address both children:
//   both internal : child0 = self+1,            child1 = link
//   one leaf      : internal child = self+1,    leaf triangles start at link
//   both leaves   : leaf0 starts at link,       leaf1 at link + cnt0
// meta = cnt0 | cnt1<<4 (cnt == 0 means "internal"; 0xF means "absent").
//
// Replaces: BVH.hit_helper accelerators/bvh.py:218-231 (recursive, unordered),
// World.hit_all mathematics/intersection_taichi.py:238-291 (threaded "next"
// links) and the slab tests mathematics/bbox.py:6-26 /
// accelerators/bvh_taichi.py:168-190.
#pragma once
#include "common.cuh"
#include "intersect.cuh"

namespace prt {

constexpr int kMaxStack = 96;  // builder guarantees tree depth < kMaxStack

__device__ __forceinline__ float clamp_dir(float d) {
    return fabsf(d) < 1e-20f ? copysignf(1e-20f, d) : d;
}

struct RayBox {  // per-ray constants of the slab test
    float3 o, idir;
    uint32_t selx, sely, selz;  // PRMT selectors: bytes -> (near0, far0, near1, far1) for this ray's octant
};

__device__ __forceinline__ RayBox make_raybox(float3 o, float3 d) {
    RayBox r;
    r.o = o;
    r.idir = make_float3(__fdiv_rn(1.0f, clamp_dir(d.x)), __fdiv_rn(1.0f, clamp_dir(d.y)),
                         __fdiv_rn(1.0f, clamp_dir(d.z)));
    r.selx = d.x < 0.0f ? 0x2301u : 0x3210u;
    r.sely = d.y < 0.0f ? 0x2301u : 0x3210u;
    r.selz = d.z < 0.0f ? 0x2301u : 0x3210u;
    return r;
}

__device__ __forceinline__ float qf(uint32_t w, int byte) {
    return (float)((w >> (8 * byte)) & 0xffu);
}

// Slab test of both children.  Returns hit mask (bit0 child0, bit1 child1) and
// entry distances.  Each axis word holds (c0.lo, c0.hi, c1.lo, c1.hi); one PRMT with the
// ray's octant selector turns it into (near0, far0, near1, far1), so no per-plane min/max
// is needed: t = q * (scale/d) + (o - ray.o)/d is already the near / far distance.
// EXACT widens every interval by the forward error bound.
template <bool EXACT>
__device__ __forceinline__ int node_test(const float4 n0, const float4 n1, const RayBox& r,
                                         float tmin, float tmax, float& t0, float& t1) {
    const uint32_t em = __float_as_uint(n0.w);
    const uint32_t qx = __byte_perm(__float_as_uint(n1.x), 0u, r.selx);
    const uint32_t qy = __byte_perm(__float_as_uint(n1.y), 0u, r.sely);
    const uint32_t qz = __byte_perm(__float_as_uint(n1.z), 0u, r.selz);
    const float sx = __uint_as_float((em & 0xffu) << 23);
    const float sy = __uint_as_float(((em >> 8) & 0xffu) << 23);
    const float sz = __uint_as_float(((em >> 16) & 0xffu) << 23);
    const float ax = sx * r.idir.x, ay = sy * r.idir.y, az = sz * r.idir.z;
    const float bx = (n0.x - r.o.x) * r.idir.x, by = (n0.y - r.o.y) * r.idir.y,
                bz = (n0.z - r.o.z) * r.idir.z;
    float n0t = fmaxf(fmaxf(fmaf(qf(qx, 0), ax, bx), fmaf(qf(qy, 0), ay, by)), fmaxf(fmaf(qf(qz, 0), az, bz), tmin));
    float f0t = fminf(fminf(fmaf(qf(qx, 1), ax, bx), fmaf(qf(qy, 1), ay, by)), fminf(fmaf(qf(qz, 1), az, bz), tmax));
    float n1t = fmaxf(fmaxf(fmaf(qf(qx, 2), ax, bx), fmaf(qf(qy, 2), ay, by)), fmaxf(fmaf(qf(qz, 2), az, bz), tmin));
    float f1t = fminf(fminf(fmaf(qf(qx, 3), ax, bx), fmaf(qf(qy, 3), ay, by)), fminf(fmaf(qf(qz, 3), az, bz), tmax));
    if (EXACT) {
        float m = 8.0f * kUnit * (fmaxf(fmaxf(fabsf(bx) + 255.0f * fabsf(ax), fabsf(by) + 255.0f * fabsf(ay)),
                                        fabsf(bz) + 255.0f * fabsf(az)));
        n0t -= m; n1t -= m; f0t += m; f1t += m;
    }
    t0 = n0t; t1 = n1t;
    const uint32_t meta = em >> 24;
    const int h0 = (n0t <= f0t) && ((meta & 0xFu) != 0xFu);
    const int h1 = (n1t <= f1t) && ((meta >> 4) != 0xFu);
    return h0 | (h1 << 1);
}

__device__ __forceinline__ void node_refs(uint32_t self, uint32_t em, uint32_t link, uint32_t& r0,
                                          uint32_t& r1) {
    uint32_t c0 = (em >> 24) & 0xFu, c1 = em >> 28;
    if (c0 == 0xFu) c0 = 1;  // absent children never pass node_test; value irrelevant
    if (c1 == 0xFu) c1 = 1;
    if (c0 == 0) {
        r0 = self + 1;
        r1 = c1 == 0 ? link : (kLeafFlag | (link << 3) | c1);
    } else {
        r0 = kLeafFlag | (link << 3) | c0;
        r1 = c1 == 0 ? self + 1 : (kLeafFlag | ((link + c0) << 3) | c1);
    }
}

// Per-thread traversal stack: the first kSmemStack levels live in shared memory
// ([level][thread], conflict-free), deeper levels spill to a local array.
struct Stack {
    uint32_t* smem;  // &s_stack[0][threadIdx.x]
    uint32_t ovf[kMaxStack - kSmemStack];
    int sp;
    __device__ __forceinline__ void push(uint32_t v) {
        if (sp < kSmemStack) smem[sp * kTraceThreads] = v;
        else ovf[sp - kSmemStack] = v;
        ++sp;
    }
    __device__ __forceinline__ uint32_t pop() {
        --sp;
        return sp < kSmemStack ? smem[sp * kTraceThreads] : ovf[sp - kSmemStack];
    }
};

}  // namespace prt
