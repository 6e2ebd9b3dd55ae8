// bvh.cuh -- the quantised 4-wide node, its slab test and the per-lane stack.
//
// Node64 (common.cuh) holds the boxes of up to FOUR children, quantised to 8 bits per
// plane in the node's own frame: plane = o + q * 2^(e-127).  The builder (bvh_build.cu)
// rounds lo down / hi up against this exact decode expression, so a decoded box always
// contains the child's true FP32 box.  The slab test never decodes a box: per axis it
// forms a = scale/d and b = (o - ray.o)/d once and each of the 24 planes costs one
// I2F.U8 + one FMA (t = q*a + b).  Planes are stored as one word of four "lo" bytes and
// one word of four "hi" bytes per axis; the ray's octant picks which word is the near
// side, so no per-plane min/max is needed.
//
// Child references are explicit (record index, or kLeafFlag | first_tri << 3 | count for
// a leaf of 1..7 triangles stored contiguously, or kNoChild).
//
// Replaces: BVH.hit_helper accelerators/bvh.py:218-231 (recursive, unordered),
// World.hit_all mathematics/intersection_taichi.py:238-291 (threaded "next" links) and
// the slab tests mathematics/bbox.py:6-26 / accelerators/bvh_taichi.py:168-190.
#pragma once
#include "common.cuh"
#include "intersect.cuh"

namespace prt {

constexpr int kMaxStack = 128;  // the builder guarantees 3 * depth + 1 <= kMaxStack
constexpr int kPStack = 16;     // levels kept in shared memory; deeper ones in local memory
constexpr int kPStackOvf = kMaxStack - kPStack;
constexpr uint32_t kDone = kNoChild;  // has kLeafFlag set

__device__ __forceinline__ float clamp_dir(float d) {
    return fabsf(d) < 1e-20f ? copysignf(1e-20f, d) : d;
}

struct RayBox {  // per-ray constants of the slab test
    float3 o, idir;
    bool negx, negy, negz;
};

__device__ __forceinline__ RayBox make_raybox(float3 o, float3 d) {
    RayBox r;
    r.o = o;
    r.idir = make_float3(__fdiv_rn(1.0f, clamp_dir(d.x)), __fdiv_rn(1.0f, clamp_dir(d.y)),
                         __fdiv_rn(1.0f, clamp_dir(d.z)));
    r.negx = d.x < 0.0f; r.negy = d.y < 0.0f; r.negz = d.z < 0.0f;
    return r;
}

__device__ __forceinline__ float qf(uint32_t w, int byte) {
    return (float)((w >> (8 * byte)) & 0xffu);  // I2F.U8 with a static byte selector
}

struct NodeHits {
    float t[4];       // entry distance per child (valid where the mask bit is set)
    uint32_t ref[4];
};

// Slab test of the four children.  Returns the hit mask.  EXACT widens every interval by
// the forward error bound of t = q*a + b so that no box the exact ray touches is missed.
template <bool EXACT>
__device__ __forceinline__ int node_test4(const Node64* __restrict__ node, const RayBox& r,
                                          float tmin, float tmax, NodeHits& h) {
    const uint4* np = reinterpret_cast<const uint4*>(node);
    const uint4 w0 = __ldg(np), w1 = __ldg(np + 1), w2 = __ldg(np + 2), w3 = __ldg(np + 3);
    const float sx = __uint_as_float((w0.w & 0xffu) << 23);
    const float sy = __uint_as_float(((w0.w >> 8) & 0xffu) << 23);
    const float sz = __uint_as_float(((w0.w >> 16) & 0xffu) << 23);
    const float ax = sx * r.idir.x, ay = sy * r.idir.y, az = sz * r.idir.z;
    const float bx = (__uint_as_float(w0.x) - r.o.x) * r.idir.x, by = (__uint_as_float(w0.y) - r.o.y) * r.idir.y,
                bz = (__uint_as_float(w0.z) - r.o.z) * r.idir.z;
    const uint32_t nx = r.negx ? w1.y : w1.x, fx = r.negx ? w1.x : w1.y;
    const uint32_t ny = r.negy ? w1.w : w1.z, fy = r.negy ? w1.z : w1.w;
    const uint32_t nz = r.negz ? w2.y : w2.x, fz = r.negz ? w2.x : w2.y;
    h.ref[0] = w2.z; h.ref[1] = w2.w; h.ref[2] = w3.x; h.ref[3] = w3.y;
    float m = 0.0f;
    if (EXACT)
        m = 8.0f * kUnit * (fmaxf(fmaxf(fabsf(bx) + 255.0f * fabsf(ax), fabsf(by) + 255.0f * fabsf(ay)),
                                  fabsf(bz) + 255.0f * fabsf(az)));
    int mask = 0;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        float tn = fmaxf(fmaxf(fmaf(qf(nx, c), ax, bx), fmaf(qf(ny, c), ay, by)), fmaxf(fmaf(qf(nz, c), az, bz), tmin));
        float tf = fminf(fminf(fmaf(qf(fx, c), ax, bx), fmaf(qf(fy, c), ay, by)), fminf(fmaf(qf(fz, c), az, bz), tmax));
        if (EXACT) { tn -= m; tf += m; }
        h.t[c] = tn;
        if (tn <= tf && h.ref[c] != kNoChild) mask |= 1 << c;
    }
    return mask;
}

// ---- per-lane stack of (child reference, entry distance) pairs -----------------------
// kPStack levels in shared memory ([level][lane] of 8-byte slots, conflict-free), deeper
// levels in local memory.  Keeping the entry distance lets a pop discard, without touching
// memory, every subtree that a closer hit found in the meantime has made irrelevant.
__device__ __forceinline__ void sstack_push(uint32_t saddr, uint2* ovf, int& sp, uint32_t ref, float t) {
    if (sp < kPStack)
        asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(saddr + (uint32_t)sp * (kTraceThreads * 8u)), "r"(ref), "r"(__float_as_uint(t)) : "memory");
    else
        ovf[sp - kPStack] = make_uint2(ref, __float_as_uint(t));
    ++sp;
}
__device__ __forceinline__ uint32_t sstack_pop(uint32_t saddr, const uint2* ovf, int& sp, float& t) {
    --sp;
    uint32_t ref, tb;
    if (sp < kPStack) {
        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(ref), "=r"(tb) : "r"(saddr + (uint32_t)sp * (kTraceThreads * 8u)) : "memory");
    } else {
        uint2 e = ovf[sp - kPStack];
        ref = e.x; tb = e.y;
    }
    t = __uint_as_float(tb);
    return ref;
}
// pop until an entry that can still beat `bound` (or the stack is empty -> kDone)
__device__ __forceinline__ uint32_t sstack_pop_live(uint32_t saddr, const uint2* ovf, int& sp, float bound) {
    while (sp > 0) {
        float t;
        const uint32_t ref = sstack_pop(saddr, ovf, sp, t);
        if (t <= bound) return ref;
    }
    return kDone;
}

__device__ __forceinline__ void cswap(float& ta, uint32_t& ra, float& tb, uint32_t& rb) {
    const bool s = tb < ta;
    const float t0 = s ? tb : ta, t1 = s ? ta : tb;
    const uint32_t r0 = s ? rb : ra, r1 = s ? ra : rb;
    ta = t0; tb = t1; ra = r0; rb = r1;
}

// Continue after a node visit: the nearest hit child becomes the next reference, the others
// are pushed far-first (so the nearer ones pop first); no hit -> pop.  Written without
// data-dependent branches around the sorting network: in a warp the lanes have 0..4 hits
// each, and serialising three code paths costs more than sorting unconditionally.
__device__ __forceinline__ uint32_t descend(int mask, NodeHits& h, uint32_t saddr, uint2* ovf, int& sp,
                                            float bound) {
    const float inf = __int_as_float(0x7f800000);
#pragma unroll
    for (int c = 0; c < 4; ++c) h.t[c] = (mask & (1 << c)) ? h.t[c] : inf;
    cswap(h.t[0], h.ref[0], h.t[1], h.ref[1]);  // sorting network for 4; misses sort to the back
    cswap(h.t[2], h.ref[2], h.t[3], h.ref[3]);
    cswap(h.t[0], h.ref[0], h.t[2], h.ref[2]);
    cswap(h.t[1], h.ref[1], h.t[3], h.ref[3]);
    cswap(h.t[1], h.ref[1], h.t[2], h.ref[2]);
    if (h.t[3] < inf) sstack_push(saddr, ovf, sp, h.ref[3], h.t[3]);
    if (h.t[2] < inf) sstack_push(saddr, ovf, sp, h.ref[2], h.t[2]);
    if (h.t[1] < inf) sstack_push(saddr, ovf, sp, h.ref[1], h.t[1]);
    if (mask == 0) return sstack_pop_live(saddr, ovf, sp, bound);
    return h.ref[0];
}

}  // namespace prt
