// persist.cuh -- persistent-warp FP32 traversal with lane-level dynamic ray fetch
// and deferred leaves.
//
// Incoherent rays have very different traversal lengths (soup-1M: mean 38 record
// visits of a 4-wide node, ~5 leaf visits, long tail).  With one ray per thread a warp
// runs until its LONGEST ray ends and every leaf visit stalls the 31 other lanes: ncu
// measured 6.7 of 32 lanes active (profiles/r1_trace_kernel_ncu.txt).  Here
//   * warps are persistent: when `refill_idle` or more lanes hold no ray, the idle
//     lanes take new ray indices from the warp's reservation (chunks of the global
//     counter, one atomic per chunk, requested one chunk ahead; ballot + prefix
//     popcount for the slot) -- Aila & Laine, "Understanding the efficiency of ray
//     traversal on GPUs" (HPG 2009);
//   * every loop iteration is PRT_VISITS_PER_ITER record visits for the lanes that
//     are at an internal record; a lane that reaches a leaf parks until `leaf_batch`
//     lanes are parked (or nobody has records left), then the parked lanes intersect
//     their leaves together.  Record visits are 60 % of the issued instructions and
//     run at 19..26 of 32 lanes; they are what is kept convergent.
//
// IO is a functor: load(k, ro, rd, tag) / store(tag, t, u, v, gid) / flag(tag); it binds this
// loop to the API ray arrays or to the wavefront queues.
//
// EXACT (PRT_TRACE_EXACT on this same loop): the triangle test carries forward error bounds
// (tri_watertight_fast_exact), the box test is widened by its own bound (node_test4<true>),
// culling keeps every candidate closer than best t + its error bound, and a ray whose answer
// could depend on an FP32 rounding is handed to io.flag() instead of io.store(); the caller
// re-traces flagged rays in FP64 (resolve_kernel).  Unflagged results are the exact-arithmetic
// winners, i.e. the reference's (mathematics/intersection.py:42-65,106-116).
#pragma once
#include "bvh.cuh"
#include "traverse.cuh"

#ifndef PRT_VISITS_PER_ITER
#define PRT_VISITS_PER_ITER 3  // record visits between two rounds of warp votes (profiles/r1_sweeps.txt)
#endif

#ifndef PRT_RAY_PREFETCH
#define PRT_RAY_PREFETCH 1  // L2 prefetch of a reserved chunk's ray records (contiguous IO only)
#endif

#ifndef PRT_MIN_BLOCKS
#define PRT_MIN_BLOCKS 8  // __launch_bounds__ min blocks/SM of the persistent traversal kernels: caps ptxas at 64 registers
#endif

namespace prt {

// scheduling knobs live in SceneDev (refill_idle: refill when at least this many lanes are idle;
// leaf_batch: intersect leaves when at least this many lanes are parked); defaults in context.cuh,
// overridable with PRT_REFILL_IDLE / PRT_LEAF_BATCH for tuning sweeps.

// EXACT slow path: one triangle whose FP32 test was inside its error bound, decided with the
// reference's FP64 formula (mt_f64) and ordered against the current best by the reference's
// (t, id) rule.  Returns 0 = not a hit / not better, 1 = `best` updated, 2 = cannot decide here
// (the current best is not reproduced in FP64): flag the ray.  Kept out of line so that its FP64
// temporaries do not count against the registers of the traversal loop.
struct ExactBest { float t, u, v, dt; int gid; };
#ifndef PRT_EXACT_INLINE
#define PRT_EXACT_INLINE 0
#endif
template <int MODE>
#if PRT_EXACT_INLINE
__device__ __forceinline__
#else
__device__ __noinline__
#endif
int exact_decide_tri(const SceneDev& sc, float3 p0, float3 p1, float3 p2, int gid, float4 ro, float4 rd, ExactBest& best) {
    const double o64[3] = {(double)ro.x, (double)ro.y, (double)ro.z};
    const double d64[3] = {(double)rd.x, (double)rd.y, (double)rd.z};
    double t64, u64, v64;
    if (!mt_f64(p0, p1, p2, o64, d64, (double)ro.w, (double)rd.w, t64, u64, v64)) return 0;
    if (MODE == MODE_ANY) { best.t = (float)t64; best.gid = gid; return 1; }
    bool take = best.gid < 0 || t64 < (double)best.t - (double)best.dt;
    int rc = 0;
    if (!take && !(t64 > (double)best.t + (double)best.dt)) {
        // near tie with the current best: order the two by the reference's (t, id)
        const float4* tp = sc.verts_gid + 3ull * (uint32_t)best.gid;
        double tb, ub, vb;
        if (!mt_f64(xyz(__ldg(tp)), xyz(__ldg(tp + 1)), xyz(__ldg(tp + 2)), o64, d64, (double)ro.w, (double)rd.w, tb, ub, vb))
            return 2;
        take = t64 < tb || (t64 == tb && gid < best.gid);
        if (!take) { best.t = (float)tb; best.u = (float)ub; best.v = (float)vb; best.dt = 2.0f * kUnit * fabsf(best.t); rc = 1; }
    }
    if (take) {
        best.t = (float)t64; best.u = (float)u64; best.v = (float)v64; best.gid = gid;
        best.dt = 2.0f * kUnit * fabsf(best.t);
        rc = 1;
    }
    return rc;
}

template <int MODE, bool COUNT, bool EXACT, class IO>
__device__ __forceinline__ void trace_persistent(const SceneDev& sc, IO io, unsigned int* fetch,
                                                 unsigned int n, uint2* stack_col, Counters* ctr) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const uint32_t saddr = (uint32_t)__cvta_generic_to_shared(stack_col);
    uint2 ovf[kPStackOvf];
    bool has_ray = false, exhausted = false;
    // per-ray state
    RayWF rw;
    RayBox rb;
    float tmin = 0.f, tmax = 0.f, bt = 0.f, bu = 0.f, bv = 0.f;
    float bdt = 0.f;       // EXACT: error bound of bt
    bool flagged = false;  // EXACT: this ray needs the FP64 replay
    int bgid = -1, sp = 0;
    uint32_t cur = kDone, tag = 0;
    unsigned long long c_nodes = 0, c_tris = 0, c_rays = 0, c_f64 = 0;
    unsigned long long c_iters = 0, c_nlanes = 0, c_lphases = 0, c_llanes = 0;  // lane 0 only

    if (sc.n_nodes == 0) {  // empty scene: everything misses
        for (unsigned k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
            float4 ro, rd;
            io.load(k, ro, rd, tag);
            io.store(tag, rd.w, 0.f, 0.f, -1);
        }
        return;
    }

    // Ray indices are reserved per warp in chunks of sc.fetch_chunk with ONE atomic per chunk, and
    // the next chunk is requested when the current one is entered, i.e. long before it is needed:
    // idle lanes never wait for an atomic round trip (one atomic per refill made the fetch counter
    // a same-address hot spot: 1-2 M atomics per launch on short traversals).
    const unsigned chunk = (unsigned)sc.fetch_chunk;
    unsigned r_next = 0, r_end = 0, p_base = 0;  // warp-uniform: current reservation; p_base (lane 0): prefetched chunk
    if (lane == 0) {
        r_next = atomicAdd(fetch, chunk);
        p_base = atomicAdd(fetch, chunk);
    }
    r_next = __shfl_sync(FULL, r_next, 0);
    r_end = r_next + chunk;

    while (true) {
        // ---- refill idle lanes
        const unsigned idle = __ballot_sync(FULL, !has_ray);
        if (!exhausted && (__popc(idle) >= sc.refill_idle)) {
            const unsigned nidle = __popc(idle), pre = __popc(idle & lt), rem = r_end - r_next;
            unsigned k = r_next + pre;
            if (nidle > rem) {  // warp-uniform: continue in the prefetched chunk, request the one after it
                const unsigned pb = __shfl_sync(FULL, p_base, 0);
                if (pre >= rem) k = pb + (pre - rem);
                r_next = pb + (nidle - rem);
                r_end = pb + chunk;
                if (lane == 0) p_base = atomicAdd(fetch, chunk);
#if PRT_RAY_PREFETCH
                if ((unsigned)lane < chunk && pb + lane < n) io.prefetch(pb + lane);  // the chunk just entered: its rays are fetched over the next refills
#endif
            } else {
                r_next += nidle;
            }
            if (!has_ray) {
                if (k < n) {
                    float4 ro, rd;
                    io.load(k, ro, rd, tag);
                    const float3 o = xyz(ro), d = xyz(rd);
                    rw = make_raywf(o, d);
                    rb = make_raybox_fast(o, d);
                    tmin = ro.w; tmax = rd.w;
                    bt = tmax; bu = 0.f; bv = 0.f; bgid = -1;
                    bdt = 0.f; flagged = false;
                    cur = 0;
                    sstack_reset<EXACT>(saddr, ovf, sp);
                    has_ray = true;
                    if (COUNT) ++c_rays;
                }
            }
            if (r_next >= n) exhausted = true;  // chunk bases only grow: nothing at or after r_next is left
        }
        if (exhausted && __ballot_sync(FULL, has_ray) == 0) break;  // (idle mask may be stale after a refill)

        if (COUNT) {
            ++c_iters;
            c_nlanes += __popc(__ballot_sync(FULL, has_ray && !(cur & kLeafFlag)));
        }
        // ---- one record visit for every lane that is at an internal record
#pragma unroll
        for (int rep = 0; rep < PRT_VISITS_PER_ITER; ++rep) {
            // EXACT: a candidate closer than best t + its error bound must still be visited
            const float bound = MODE == MODE_CLOSEST ? (EXACT ? bt + bdt : bt) : tmax;
            if (has_ray && !(cur & kLeafFlag)) {
                if (COUNT) ++c_nodes;
                NodeHits h;
                node_test4<EXACT>(sc.nodes + cur, rb, tmin, bound, h);
                cur = MODE == MODE_ANY ? descend_any(h, saddr, ovf, sp) : descend(h, saddr, ovf, sp, bound);
            }
        }

        // ---- leaves: parked lanes go together
        const bool at_leaf = has_ray && (cur & kLeafFlag) && cur != kDone;
        const unsigned leafm = __ballot_sync(FULL, at_leaf);
        // idle lanes keep cur == kDone, so "no leaf flag" == "still walking records"
        const unsigned nodem = ~__ballot_sync(FULL, (cur & kLeafFlag) != 0u);
        if (leafm && (__popc(leafm) >= sc.leaf_batch || nodem == 0)) {
            if (COUNT) { ++c_lphases; c_llanes += __popc(leafm); }
            if (at_leaf) {
                const uint32_t start = (cur & ~kLeafFlag) >> 3, cnt = cur & 7u;
                bool stop = false;
                for (uint32_t k = 0; k < cnt; ++k) {
                    float3 p0, p1, p2;
                    int gid;
                    load_tri<false>(sc, start + k, p0, p1, p2, gid);
                    if (COUNT) ++c_tris;
                    TriHit h;
                    int hit;
                    if constexpr (EXACT) {
                        bool unc = false;
                        hit = tri_watertight_fast_exact(rw, p0, p1, p2, tmin, MODE == MODE_CLOSEST ? bt : tmax,
                                                        MODE == MODE_CLOSEST ? bdt : 0.0f, h, unc);
                        if (unc) {
                            // The FP32 decision is inside its error bound: decide THIS triangle with the
                            // reference's own FP64 formula, in place (rare).  Only what still cannot be
                            // ordered goes to the FP64 replay.
                            float4 ro4, rd4;
                            io.reload(tag, ro4, rd4);
                            if (COUNT) ++c_f64;
                            ExactBest eb{bt, bu, bv, bdt, bgid};
                            const int rc = exact_decide_tri<MODE>(sc, p0, p1, p2, gid, ro4, rd4, eb);
                            bt = eb.t; bu = eb.u; bv = eb.v; bdt = eb.dt; bgid = eb.gid;
                            if (rc == 2) { flagged = true; stop = true; break; }
                            if (rc == 1 && MODE == MODE_ANY) { stop = true; break; }
                            continue;
                        }
                    } else {
                        hit = tri_watertight_fast(rw, p0, p1, p2, tmin, MODE == MODE_CLOSEST ? bt : tmax, h);
                    }
                    if (hit) {
                        if (MODE == MODE_CLOSEST) {
                            if (h.t < bt || bgid < 0 || gid < bgid) { bt = h.t; bu = h.u; bv = h.v; bgid = gid; bdt = h.dt; }
                        } else {
                            bt = h.t; bgid = gid; stop = true;
                            break;
                        }
                    }
                }
                cur = stop ? kDone : sstack_pop_live(saddr, ovf, sp, MODE == MODE_CLOSEST ? (EXACT ? bt + bdt : bt) : tmax);
            }
        }
        if (has_ray && cur == kDone) {
            if (EXACT && flagged) io.flag(tag);
            else io.store(tag, bt, bu, bv, bgid);
            has_ray = false;
        }
    }
    if (COUNT) {
        for (int o = 16; o > 0; o >>= 1) {
            c_nodes += __shfl_down_sync(FULL, c_nodes, o);
            c_tris += __shfl_down_sync(FULL, c_tris, o);
            c_rays += __shfl_down_sync(FULL, c_rays, o);
            c_f64 += __shfl_down_sync(FULL, c_f64, o);
        }
        if (lane == 0) {
            atomicAdd(&ctr->node_visits, c_nodes);
            atomicAdd(&ctr->tri_tests, c_tris);
            atomicAdd(MODE == MODE_ANY ? &ctr->rays_shadow : &ctr->rays_closest, c_rays);
            atomicAdd(&ctr->warp_iters, c_iters);
            atomicAdd(&ctr->node_lane_iters, c_nlanes);
            atomicAdd(&ctr->leaf_phases, c_lphases);
            atomicAdd(&ctr->leaf_lane_phases, c_llanes);
            if (EXACT) atomicAdd(&ctr->f64_decisions, c_f64);
        }
    }
}

}  // namespace prt
