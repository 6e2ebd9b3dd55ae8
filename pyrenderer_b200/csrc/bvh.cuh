// bvh.cuh -- the quantised 4-wide node, its slab test and the per-lane stack.
//
// Node64 (common.cuh) holds the boxes of up to FOUR children, quantised to 8 bits per
// plane in the node's own frame: plane = o + q * scale, scale a power of two per axis.
// The builder (bvh_build.cu) rounds lo down / hi up against this exact decode expression,
// so a decoded box always contains the child's true FP32 box.  The slab test never
// decodes a box: per axis it forms a = scale/d and b = (o - ray.o)/d once; a plane then
// costs one byte->float conversion (I2F.U8 on the XU pipe, or a folded PRMT on the ALU
// pipe, PlaneEval) and half a packed FFMA2 (t = q*a + b for two children at once).
// Planes are stored as one word of four "lo" bytes and one word of four "hi" bytes per
// axis; the ray's octant picks which word is the near side, so no per-plane min/max is
// needed.  The node is fetched with two 256-bit loads (ldg_node).
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
constexpr int kPStackOvf = kMaxStack + 1;  // local-memory spill area: [0].x = spilled count, entries from [1]
constexpr uint32_t kDone = kNoChild;  // has kLeafFlag set

// |d| < 1e-20 -> +-1e-20 with the sign that `d < 0` sees: -0.0 counts as positive, like the octant
// flags below (copysign would give -1e-20, a reciprocal of -1e20 under a "positive" octant: near and
// far planes swapped, and a ray travelling inside a slab would miss every box)
__device__ __forceinline__ float clamp_dir(float d) {
    return fabsf(d) < 1e-20f ? (d < 0.0f ? -1e-20f : 1e-20f) : d;
}

struct RayBox {  // per-ray constants of the slab test
    float3 o, idir;
    bool negx, negy, negz;  // octant (consistent with clamp_dir: d < 0)
};

__device__ __forceinline__ RayBox make_raybox(float3 o, float3 d) {
    RayBox r;
    r.o = o;
    r.idir = make_float3(__fdiv_rn(1.0f, clamp_dir(d.x)), __fdiv_rn(1.0f, clamp_dir(d.y)),
                         __fdiv_rn(1.0f, clamp_dir(d.z)));
    r.negx = d.x < 0.0f; r.negy = d.y < 0.0f; r.negz = d.z < 0.0f;
    return r;
}

__device__ __forceinline__ RayBox make_raybox_fast(float3 o, float3 d) {  // persist.cuh: see rcp_fast
    RayBox r;
    r.o = o;
    r.idir = make_float3(rcp_fast(clamp_dir(d.x)), rcp_fast(clamp_dir(d.y)), rcp_fast(clamp_dir(d.z)));
    r.negx = d.x < 0.0f; r.negy = d.y < 0.0f; r.negz = d.z < 0.0f;
    return r;
}

__device__ __forceinline__ float qf(uint32_t w, int byte) {
    return (float)((w >> (8 * byte)) & 0xffu);  // I2F.U8 with a static byte selector
}

// Node fetch.  (Tried in round 2: serving the first 21 / 85 records -- the top 3 / 4 levels, which
// every ray walks -- from a per-CTA copy in shared memory: 8.52 -> 8.80 ms on soup-1M; the mixed
// LDS / LDG paths cost more issue slots than the shorter latency gives back.)
// The traversal is bound by L1TEX wavefronts (one tag look-up per distinct
// 128-byte line per load instruction, profiles/r1_trace_persistent_bvh4_ncu.txt: L1/TEX
// throughput 89 %), so the 64-byte node is read with TWO 256-bit loads (sm_100
// LDG.E.256, PTX ld.global.v8.b32) instead of four 128-bit ones.
#ifndef PRT_LDG256
#define PRT_LDG256 1
#endif
__device__ __forceinline__ void ldg256(const void* p, uint4& a, uint4& b) {
    asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
                 : "l"(p));
}
// streaming (evict-first) 256-bit load of one 32-byte ray record: one sector, one request
__device__ __forceinline__ void ldg256_cs(const void* p, float4& a, float4& b) {
    asm volatile("ld.global.cs.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
                 : "l"(p));
}
__device__ __forceinline__ void ldg_node(const Node64* __restrict__ node, uint4& w0, uint4& w1, uint4& w2, uint4& w3) {
#if PRT_LDG256
    ldg256(node, w0, w1);
    ldg256(reinterpret_cast<const char*>(node) + 32, w2, w3);
#else
    const uint4* np = reinterpret_cast<const uint4*>(node);
    w0 = __ldg(np); w1 = __ldg(np + 1); w2 = __ldg(np + 2); w3 = __ldg(np + 3);
#endif
}

// triangle fetch: leaf order = one 256-bit + one 64-bit load; BRUTE = global-id order, 3 x float4
template <bool BRUTE>
__device__ __forceinline__ void load_tri(const SceneDev& sc, uint32_t i, float3& p0, float3& p1, float3& p2, int& gid) {
    if (BRUTE) {
        const float4* tp = sc.tris + 3ull * i;
        const float4 a = __ldg(tp), b = __ldg(tp + 1), c = __ldg(tp + 2);
        p0 = xyz(a); p1 = xyz(b); p2 = xyz(c); gid = __float_as_int(a.w);
    } else {
        uint4 a, b;
        ldg256(sc.tri_a + 2ull * i, a, b);
        const float2 c = __ldg(sc.tri_b + i);
        p0 = make_float3(__uint_as_float(a.x), __uint_as_float(a.y), __uint_as_float(a.z));
        p1 = make_float3(__uint_as_float(a.w), __uint_as_float(b.x), __uint_as_float(b.y));
        p2 = make_float3(__uint_as_float(b.z), __uint_as_float(b.w), c.x);
        gid = __float_as_int(c.y);
    }
}

struct NodeHits {
    float t[4];       // entry distance per child, +inf where the child is missed or absent
    uint32_t ref[4];
};

// byte c of a plane word -> float.  PRT_QCONV_AXES of the three axes use the "folded" form:
// PRMT builds the float 1 + q * 2^-15 (0x3F80qq00) on the ALU pipe and the slab becomes
// t = f * (a * 2^15) + (b - a * 2^15); the other axes use I2F.U8 (XU pipe, 1/4 rate).  Splitting
// the 24 conversions of a visit between the two pipes keeps either from limiting issue.
#ifndef PRT_FMA2
#define PRT_FMA2 1
#endif
#ifndef PRT_QCONV_AXES
#define PRT_QCONV_AXES 2  // profiles/r1_sweeps.txt: 0 -> 1666, 1 -> 1706, 2 -> 1724, 3 -> 1689 Mrays/s (soup-1M)
#endif
// Two planes per instruction: sm_100 packed FP32 (FFMA2, PTX fma.rn.f32x2) with the per-axis
// constants broadcast to both halves -- 12 FFMA2 per visit instead of 24 FFMA.
__device__ __forceinline__ void fma2_bcast(float q0, float q1, float a, float b, float& t0, float& t1) {
    unsigned long long q, d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(q) : "f"(q0), "f"(q1));
    asm("{\n\t.reg .b64 aa, bb;\n\tmov.b64 aa, {%2, %2};\n\tmov.b64 bb, {%3, %3};\n\tfma.rn.f32x2 %0, %1, aa, bb;\n\t}"
        : "=l"(d) : "l"(q), "f"(a), "f"(b));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(t0), "=f"(t1) : "l"(d));
}
template <bool FOLD>
struct PlaneEval {
    float a, b;
    __device__ __forceinline__ PlaneEval(float scale, float origin, float ro, float idir) {
        const float a0 = scale * idir, b0 = (origin - ro) * idir;
        if (FOLD) { a = a0 * 32768.0f; b = b0 - a; } else { a = a0; b = b0; }
    }
    // EXACT: bound of |computed - exact| for any plane of this axis, in units of u = 2^-24.
    // Unfolded (t = q*a + b, a = s*idir exact up to idir's error e_i <= 2u, b two roundings + e_i):
    //   <= 3u*255|a| + 5u|b|  ->  8u (|b| + 255|a|).
    // Folded (a' = 2^15 a exact, b' = fl(b - a'), t = fl(f*a' + b')): the extra rounding of b' adds
    //   u (|b| + |a'|), in the stored constants <= 8u (|b'| + |a'|)  ->  11u (|b'| + |a'|),
    //   i.e. 2 % of one quantisation step: boxes grow by nothing that shows in the visit count.
    __device__ __forceinline__ float err_bound() const {
        return FOLD ? 11.0f * kUnit * (fabsf(b) + fabsf(a)) : 8.0f * kUnit * (fabsf(b) + 255.0f * fabsf(a));
    }
    __device__ __forceinline__ float q(uint32_t w, int c) const {  // byte c -> the float the FMA consumes
        if (FOLD) return __uint_as_float(__byte_perm(w, 0x3F800000u, 0x7604u | (uint32_t)(c << 4)));
        return qf(w, c);
    }
    // the four children's planes of one word; `shift` (EXACT) moves them by the error bound
    template <bool SHIFT>
    __device__ __forceinline__ void eval4(uint32_t w, float t[4], float shift) const {
        const float bb = SHIFT ? b + shift : b;
#if PRT_FMA2
        fma2_bcast(q(w, 0), q(w, 1), a, bb, t[0], t[1]);
        fma2_bcast(q(w, 2), q(w, 3), a, bb, t[2], t[3]);
#else
#pragma unroll
        for (int c = 0; c < 4; ++c) t[c] = fmaf(q(w, c), a, bb);
#endif
    }
};

// Slab test of the four children: h.t[c] = entry distance, +inf for a miss.  EXACT moves every
// plane distance outwards by the forward error bound of t = q*a + b OF ITS OWN AXIS, so that no
// box the exact ray touches is missed.  (Per axis, not one bound for the whole test: for a ray
// nearly parallel to an axis that axis has |idir| ~ 1e5 and an error bound larger than the other
// two slabs; applied to all three it made 3 % of the rays visit far too much and a handful
// traverse the whole tree -- a 9 ms tail on a 10 ms launch.)
template <bool EXACT>
__device__ __forceinline__ void node_test4(const Node64* __restrict__ node, const RayBox& r,
                                           float tmin, float tmax, NodeHits& h) {
    uint4 w0, w1, w2, w3;
    ldg_node(node, w0, w1, w2, w3);
    constexpr int kFold = PRT_QCONV_AXES;
    const PlaneEval<(kFold > 2)> px(__uint_as_float(w0.w), __uint_as_float(w0.x), r.o.x, r.idir.x);
    const PlaneEval<(kFold > 1)> py(__uint_as_float(w1.x), __uint_as_float(w0.y), r.o.y, r.idir.y);
    const PlaneEval<(kFold > 0)> pz(__uint_as_float(w1.y), __uint_as_float(w0.z), r.o.z, r.idir.z);
    const uint32_t nx = r.negx ? w1.w : w1.z, fx = r.negx ? w1.z : w1.w;
    const uint32_t ny = r.negy ? w2.y : w2.x, fy = r.negy ? w2.x : w2.y;
    const uint32_t nz = r.negz ? w2.w : w2.z, fz = r.negz ? w2.z : w2.w;
    h.ref[0] = w3.x; h.ref[1] = w3.y; h.ref[2] = w3.z; h.ref[3] = w3.w;
    const float mx = EXACT ? px.err_bound() : 0.0f, my = EXACT ? py.err_bound() : 0.0f, mz = EXACT ? pz.err_bound() : 0.0f;
    const float inf = __int_as_float(0x7f800000);
    float xn[4], xf[4], yn[4], yf[4], zn[4], zf[4];
    px.template eval4<EXACT>(nx, xn, -mx); px.template eval4<EXACT>(fx, xf, mx);
    py.template eval4<EXACT>(ny, yn, -my); py.template eval4<EXACT>(fy, yf, my);
    pz.template eval4<EXACT>(nz, zn, -mz); pz.template eval4<EXACT>(fz, zf, mz);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        float tn = fmaxf(fmaxf(xn[c], yn[c]), fmaxf(zn[c], tmin));
        float tf = fminf(fminf(xf[c], yf[c]), fminf(zf[c], tmax));
        // absent children carry inverted planes AND kNoChild: the planes alone are not proof
        // (255*a + b can round to b for a tiny record far away)
        // (children 0 and 1 always exist: the one-triangle scene pads slot 1 with an empty leaf)
        h.t[c] = (tn <= tf && (c < 2 || h.ref[c] != kNoChild)) ? tn : inf;
    }
}

// ---- per-lane stack of (child reference, entry distance) pairs -----------------------
// The top kPStack levels live in shared memory ([level][lane] of 8-byte slots, conflict-free).
// Hot paths never test for overflow: a push that finds fewer than three free levels first
// SPILLS the bottom kSpill levels to a local-memory array (and shifts the rest down), a pop that
// finds the shared part empty UNSPILLS kSpill levels -- both rare, out-of-line in effect.  The
// spilled count lives in ovf[0].x (local memory, not a register).  With SENTINEL the bottom entry
// of the stack is (kDone, -inf), which ends the traversal through the ordinary pop path, so the
// count is only read after a real spill (reading it at the end of every ray is 1.5 .. 3.4 % of the
// stall samples, one lane at a time): measured +1.3 % on the EXACT kernel, -1.8 % on the plain one
// (profiles/r2_sweeps.txt), so only the EXACT kernel uses it.  Keeping the entry distance lets a
// pop discard, without touching memory, every subtree that a closer hit found in the meantime
// has made irrelevant.
constexpr int kSpill = kPStack / 2;
// (inlined: as real calls the two slow paths cost 4 % -- ABI constraints on the hot loop)
constexpr uint32_t kStackStride = kTraceThreads * 8u;
__device__ __forceinline__ void sstack_st(uint32_t saddr, int i, uint32_t ref, uint32_t tb) {
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(saddr + (uint32_t)i * kStackStride), "r"(ref), "r"(tb) : "memory");
}
__device__ __forceinline__ void sstack_ld(uint32_t saddr, int i, uint32_t& ref, uint32_t& tb) {
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(ref), "=r"(tb) : "r"(saddr + (uint32_t)i * kStackStride) : "memory");
}
template <bool SENTINEL = false>
__device__ __forceinline__ void sstack_reset(uint32_t saddr, uint2* ovf, int& sp) {
    sp = 0;
    if (SENTINEL) {
        sstack_st(saddr, 0, kDone, 0xff800000u);  // "finished", entry distance -inf (never culled)
        sp = 1;
    }
    ovf[0].x = 0u;
}
// (sp by value in and out: a reference parameter of a real call would pin sp to local memory)
static __device__ __forceinline__ int sstack_spill(uint32_t saddr, uint2* ovf, int sp) {
    const uint32_t n = ovf[0].x;
    for (int i = 0; i < kSpill; ++i) {
        uint32_t r, t;
        sstack_ld(saddr, i, r, t);
        ovf[1 + n + i] = make_uint2(r, t);
    }
    for (int i = kSpill; i < sp; ++i) {
        uint32_t r, t;
        sstack_ld(saddr, i, r, t);
        sstack_st(saddr, i - kSpill, r, t);
    }
    ovf[0].x = n + kSpill;
    return sp - kSpill;
}
static __device__ __forceinline__ int sstack_unspill(uint32_t saddr, uint2* ovf) {  // shared part empty; returns the new sp
    const uint32_t n = ovf[0].x;
    if (n == 0u) return 0;
    for (int i = 0; i < kSpill; ++i) {
        const uint2 e = ovf[1 + n - kSpill + i];
        sstack_st(saddr, i, e.x, e.y);
    }
    ovf[0].x = n - kSpill;
    return kSpill;
}
// pop until an entry that can still beat `bound` (the sentinel always does -> kDone)
// (tried: peeling the first pop and draining four entries per shared-memory round trip after a dead
// one -- the drain is 17 % of the stall samples at 3..6 lanes -- but the window's registers spill in a
// 64-register kernel: 8.4 -> 9.3 (window 2) .. 13.9 ms (window 8), profiles/r2_sweeps.txt)
__device__ __forceinline__ uint32_t sstack_pop_live(uint32_t saddr, uint2* ovf, int& sp, float bound) {
    while (true) {
        if (sp == 0) {
            sp = sstack_unspill(saddr, ovf);
            if (sp == 0) return kDone;
        }
        --sp;
        uint32_t ref, tb;
        sstack_ld(saddr, sp, ref, tb);
        if (__uint_as_float(tb) <= bound) return ref;
    }
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
__device__ __forceinline__ void sort_hits(NodeHits& h) {
    cswap(h.t[0], h.ref[0], h.t[1], h.ref[1]);  // sorting network for 4; misses (+inf) sort to the back
    cswap(h.t[2], h.ref[2], h.t[3], h.ref[3]);
    cswap(h.t[0], h.ref[0], h.t[2], h.ref[2]);
    cswap(h.t[1], h.ref[1], h.t[3], h.ref[3]);
    cswap(h.t[1], h.ref[1], h.t[2], h.ref[2]);  // (dropping this comparator costs 5 % more visits)
}

// push the (sorted) hits 3, 2, 1: three unconditional stores, the stack pointer advances only
// past the valid ones -- no branch per child.
__device__ __forceinline__ void push_far(const NodeHits& h, uint32_t saddr, uint2* ovf, int& sp) {
    const float inf = __int_as_float(0x7f800000);
    if (sp > kPStack - 3) sp = sstack_spill(saddr, ovf, sp);
#pragma unroll
    for (int c = 3; c >= 1; --c) {
        sstack_st(saddr, sp, h.ref[c], __float_as_uint(h.t[c]));
        sp += h.t[c] < inf ? 1 : 0;
    }
}

__device__ __forceinline__ uint32_t descend(NodeHits& h, uint32_t saddr, uint2* ovf, int& sp, float bound) {
    const float inf = __int_as_float(0x7f800000);
    sort_hits(h);
    push_far(h, saddr, ovf, sp);
    if (!(h.t[0] < inf)) return sstack_pop_live(saddr, ovf, sp, bound);
    return h.ref[0];
}

// Any-hit traversal needs no order: all four children are stored unconditionally (the stack
// pointer advances past the hit ones) and the top is popped back -- no sorting network, no
// +inf selects, no culling on pop (an entry's distance was <= tmax when it was pushed and tmax
// never shrinks).  37 instead of 66 instructions of bookkeeping per visit.
__device__ __forceinline__ uint32_t descend_any(const NodeHits& h, uint32_t saddr, uint2* ovf, int& sp) {
    const float inf = __int_as_float(0x7f800000);
    if (sp > kPStack - 4) sp = sstack_spill(saddr, ovf, sp);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        sstack_st(saddr, sp, h.ref[c], __float_as_uint(h.t[c]));
        sp += h.t[c] < inf ? 1 : 0;
    }
    if (sp == 0) {
        sp = sstack_unspill(saddr, ovf);
        if (sp == 0) return kDone;
    }
    --sp;
    uint32_t ref, tb;
    sstack_ld(saddr, sp, ref, tb);
    return ref;
}

}  // namespace prt
