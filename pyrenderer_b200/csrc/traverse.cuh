// traverse.cuh -- one-ray traversal device functions shared by the API trace
// kernels (traverse.cu) and the wavefront integrator (wavefront.cu).
#pragma once
#include "bvh.cuh"

namespace prt {

enum { MODE_CLOSEST = 0, MODE_ANY = 1, MODE_ALL = 2 };
constexpr unsigned long long kHashMul = 0x9E3779B97F4A7C15ull;

struct TraceResult {
    float t, u, v, dt;
    int gid;              // -1 = miss
    bool uncertain;       // EXACT: needs the FP64 replay
    uint32_t count;       // MODE_ALL
    unsigned long long sum;
    uint32_t n_nodes, n_tris;  // COUNT
};

template <int MODE, bool EXACT, bool COUNT, bool BRUTE>
__device__ __forceinline__ bool process_tris(const SceneDev& sc, const RayW& rw, uint32_t start,
                                             uint32_t cnt, float tmin, float tmax,
                                             TraceResult& res) {
    for (uint32_t k = 0; k < cnt; ++k) {
        float3 p0, p1, p2;
        int gid;
        load_tri<BRUTE>(sc, start + k, p0, p1, p2, gid);
        if (COUNT) ++res.n_tris;
        TriHit h;
        bool unc = false;
        float bound = MODE == MODE_CLOSEST ? res.t : tmax;
        float berr = MODE == MODE_CLOSEST ? res.dt : 0.0f;
        int hit = tri_watertight<EXACT>(rw, p0, p1, p2, tmin, bound, berr, h, unc);
        if (EXACT && unc) { res.uncertain = true; return true; }
        if (!hit) continue;
        if (MODE == MODE_CLOSEST) {
            if (h.t < res.t || gid < res.gid || res.gid < 0) {
                res.t = h.t; res.u = h.u; res.v = h.v; res.dt = h.dt; res.gid = gid;
            }
        } else if (MODE == MODE_ANY) {
            res.gid = gid; res.t = h.t;
            return true;
        } else {
            ++res.count;
            res.sum += (unsigned long long)(gid + 1) * kHashMul;
        }
    }
    return false;
}

// FP32 traversal, one ray per thread.  `stack_col` = &s_stack[0][threadIdx.x] of a
// __shared__ uint2 s_stack[kPStack][kTraceThreads].
template <int MODE, bool EXACT, bool COUNT, bool BRUTE>
__device__ __forceinline__ void trace_one(const SceneDev& sc, float4 ro, float4 rd,
                                          uint2* stack_col, TraceResult& res) {
    float3 o = xyz(ro), d = xyz(rd);
    float tmin = ro.w, tmax = rd.w;
    res.t = tmax; res.u = 0.f; res.v = 0.f; res.dt = 0.f; res.gid = -1;
    res.uncertain = false; res.count = 0; res.sum = 0ull; res.n_nodes = 0; res.n_tris = 0;
    RayW rw = make_rayw(o, d);
    if (BRUTE) {
        for (uint32_t s = 0; s < sc.nt; s += 4096u) {
            uint32_t c = min(4096u, sc.nt - s);
            if (process_tris<MODE, EXACT, COUNT, true>(sc, rw, s, c, tmin, tmax, res)) return;
        }
        return;
    }
    if (sc.n_nodes == 0) return;
    RayBox rb = make_raybox(o, d);
    const uint32_t saddr = (uint32_t)__cvta_generic_to_shared(stack_col);
    uint2 ovf[kPStackOvf];
    int sp;
    sstack_reset(saddr, ovf, sp);
    uint32_t cur = 0;
    while (cur != kDone) {
        // EXACT: a candidate closer than best + its error bound must still be visited
        const float bound = MODE == MODE_CLOSEST ? res.t + (EXACT ? res.dt : 0.0f) : tmax;
        if (!(cur & kLeafFlag)) {
            if (COUNT) ++res.n_nodes;
            NodeHits h;
            node_test4<EXACT>(sc.nodes + cur, rb, tmin, bound, h);
            cur = descend(h, saddr, ovf, sp, bound);
        } else {
            uint32_t start = (cur & ~kLeafFlag) >> 3, cnt = cur & 7u;
            if (process_tris<MODE, EXACT, COUNT, false>(sc, rw, start, cnt, tmin, tmax, res)) return;
            cur = sstack_pop_live(saddr, ovf, sp, MODE == MODE_CLOSEST ? res.t + (EXACT ? res.dt : 0.0f) : tmax);
        }
    }
}

// FP64 replay: the reference's Moller-Trumbore in double over the same BVH
// (box culling stays FP32 but is widened by its error bound, so it never
// rejects a triangle the reference would accept).  Result follows
// intersection.py:106-116 / scene.py:66-73: min t, lowest GLOBAL id on ties.
struct TraceResult64 {
    double t, u, v;
    int gid;
    uint32_t count;
    unsigned long long sum;
};

template <int MODE, bool BRUTE>
__device__ __forceinline__ bool process_tris64(const SceneDev& sc, const double* o, const double* d,
                                               uint32_t start, uint32_t cnt, double tmin,
                                               double tmax, TraceResult64& res) {
    for (uint32_t k = 0; k < cnt; ++k) {
        float3 p0, p1, p2;
        int gid;
        load_tri<BRUTE>(sc, start + k, p0, p1, p2, gid);
        double t, u, v;
        double bound = MODE == MODE_CLOSEST ? res.t : tmax;
        if (!mt_f64(p0, p1, p2, o, d, tmin, bound, t, u, v)) continue;
        if (MODE == MODE_CLOSEST) {
            if (t < res.t || gid < res.gid || res.gid < 0) { res.t = t; res.u = u; res.v = v; res.gid = gid; }
        } else if (MODE == MODE_ANY) {
            res.gid = gid; res.t = t;
            return true;
        } else {
            ++res.count;
            res.sum += (unsigned long long)(gid + 1) * kHashMul;
        }
    }
    return false;
}

template <int MODE, bool BRUTE>
__device__ __forceinline__ void trace_one_f64(const SceneDev& sc, float4 ro, float4 rd,
                                              uint2* stack_col, TraceResult64& res) {
    double o[3] = {(double)ro.x, (double)ro.y, (double)ro.z};
    double d[3] = {(double)rd.x, (double)rd.y, (double)rd.z};
    double tmin = (double)ro.w, tmax = (double)rd.w;
    res.t = tmax; res.u = 0.0; res.v = 0.0; res.gid = -1; res.count = 0; res.sum = 0ull;
    if (BRUTE) {
        for (uint32_t s = 0; s < sc.nt; s += 4096u) {
            uint32_t c = min(4096u, sc.nt - s);
            if (process_tris64<MODE, true>(sc, o, d, s, c, tmin, tmax, res)) return;
        }
        return;
    }
    if (sc.n_nodes == 0) return;
    RayBox rb = make_raybox(xyz(ro), xyz(rd));
    const uint32_t saddr = (uint32_t)__cvta_generic_to_shared(stack_col);
    uint2 ovf[kPStackOvf];
    int sp;
    sstack_reset(saddr, ovf, sp);
    uint32_t cur = 0;
    while (cur != kDone) {
        // box culling stays FP32 but conservative: bound rounded up, intervals widened (EXACT)
        const float bound = MODE == MODE_CLOSEST ? __double2float_ru(res.t) : rd.w;
        if (!(cur & kLeafFlag)) {
            NodeHits h;
            node_test4<true>(sc.nodes + cur, rb, ro.w, bound, h);
            cur = descend(h, saddr, ovf, sp, bound);
        } else {
            uint32_t start = (cur & ~kLeafFlag) >> 3, cnt = cur & 7u;
            if (process_tris64<MODE, false>(sc, o, d, start, cnt, tmin, tmax, res)) return;
            cur = sstack_pop_live(saddr, ovf, sp, MODE == MODE_CLOSEST ? __double2float_ru(res.t) : rd.w);
        }
    }
}

}  // namespace prt
