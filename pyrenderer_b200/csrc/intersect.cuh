// intersect.cuh -- ray/triangle device functions.
//
//  * tri_watertight<EXACT>: FP32 watertight test (Woop, Benthin, Wald 2013; the
//    shape of reference mathematics/intersection_taichi.py:94-161
//    `ray_triangle_hit2`), accept rule of the reference's Moller-Trumbore
//    (mathematics/intersection.py:42-65): t_lo <= t <= bound, u,v >= 0,
//    u+v <= 1, all edges inclusive, no back-face culling.
//    With EXACT it also evaluates forward error bounds and reports when the FP32
//    decision could differ from the exact-arithmetic one.
//  * mt_f64: the reference's grouped kernel in double, same operation order
//    (mathematics/fast_op.py:60-63,74-92; intersection.py:42-82), built from
//    __dmul_rn/__dadd_rn so nothing is contracted into FMA -- numba does not fuse.
#pragma once
#include "common.cuh"

namespace prt {

constexpr double kRefEps = 1.1754943508222875e-38;  // mathematics/constants.py:14
constexpr float kUnit = 5.9604644775390625e-08f;    // 2^-24
constexpr float kErrEdge = 12.0f * kUnit;           // see DESIGN.md "exact mode"

struct RayW {  // per-ray setup of the watertight test
    float3 o;
    float Sx, Sy, Sz;
    int kz;
};

__device__ __forceinline__ float3 perm3(float3 v, int kz) {
    // (kx,ky,kz) cyclic with kz = dominant axis
    return kz == 2 ? v : (kz == 0 ? make_float3(v.y, v.z, v.x) : make_float3(v.z, v.x, v.y));
}

__device__ __forceinline__ RayW make_rayw(float3 o, float3 d) {
    RayW r;
    r.o = o;
    float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
    r.kz = (ax > ay) ? (ax > az ? 0 : 2) : (ay > az ? 1 : 2);
    float3 p = perm3(d, r.kz);
    r.Sx = __fdiv_rn(p.x, p.z);
    r.Sy = __fdiv_rn(p.y, p.z);
    r.Sz = __fdiv_rn(1.0f, p.z);
    return r;
}

struct TriHit {
    float t, u, v;   // u,v = Moller-Trumbore barycentrics (weights of p1, p2)
    float dt;        // EXACT: absolute error bound of t
};

// returns 1 on accept.  `uncertain` (EXACT only) is set when the decision, or
// the ordering against `bound`, is within the FP32 error bound.
template <bool EXACT>
__device__ __forceinline__ int tri_watertight(const RayW& r, float3 p0, float3 p1, float3 p2,
                                              float t_lo, float bound, float bound_err,
                                              TriHit& h, bool& uncertain) {
    float3 A = perm3(p0 - r.o, r.kz), B = perm3(p1 - r.o, r.kz), C = perm3(p2 - r.o, r.kz);
    float Ax = fmaf(-r.Sx, A.z, A.x), Ay = fmaf(-r.Sy, A.z, A.y);
    float Bx = fmaf(-r.Sx, B.z, B.x), By = fmaf(-r.Sy, B.z, B.y);
    float Cx = fmaf(-r.Sx, C.z, C.x), Cy = fmaf(-r.Sy, C.z, C.y);
    // two rounded products and one subtraction: exactly antisymmetric under a
    // swap of the edge's end points, which is what makes shared edges watertight
    float U = __fsub_rn(__fmul_rn(Cx, By), __fmul_rn(Cy, Bx));
    float V = __fsub_rn(__fmul_rn(Ax, Cy), __fmul_rn(Ay, Cx));
    float W = __fsub_rn(__fmul_rn(Bx, Ay), __fmul_rn(By, Ax));
    if (U == 0.0f || V == 0.0f || W == 0.0f) {
        U = (float)(__dmul_rn((double)Cx, (double)By) - __dmul_rn((double)Cy, (double)Bx));
        V = (float)(__dmul_rn((double)Ax, (double)Cy) - __dmul_rn((double)Ay, (double)Cx));
        W = (float)(__dmul_rn((double)Bx, (double)Ay) - __dmul_rn((double)By, (double)Ax));
    }
    float eU = 0.f, eV = 0.f, eW = 0.f;
    if (EXACT) {
        float mAx = fabsf(A.x) + fabsf(r.Sx * A.z), mAy = fabsf(A.y) + fabsf(r.Sy * A.z);
        float mBx = fabsf(B.x) + fabsf(r.Sx * B.z), mBy = fabsf(B.y) + fabsf(r.Sy * B.z);
        float mCx = fabsf(C.x) + fabsf(r.Sx * C.z), mCy = fabsf(C.y) + fabsf(r.Sy * C.z);
        eU = kErrEdge * (mCx * mBy + mCy * mBx);
        eV = kErrEdge * (mAx * mCy + mAy * mCx);
        eW = kErrEdge * (mBx * mAy + mBy * mAx);
        bool neg = (U < -eU) || (V < -eV) || (W < -eW);
        bool pos = (U > eU) || (V > eV) || (W > eW);
        if (neg && pos) return 0;  // certainly outside
    } else {
        if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return 0;
    }
    float det = U + V + W;
    float Az = r.Sz * A.z, Bz = r.Sz * B.z, Cz = r.Sz * C.z;
    float T = fmaf(U, Az, fmaf(V, Bz, W * Cz));
    if (EXACT) {
        bool edge_sure = (fabsf(U) > eU) && (fabsf(V) > eV) && (fabsf(W) > eW);
        float eDet = eU + eV + eW + 4.0f * kUnit * (fabsf(U) + fabsf(V) + fabsf(W));
        if (fabsf(det) <= eDet) { uncertain = true; return 0; }
        float t = __fdiv_rn(T, det);
        float eT = 6.0f * kUnit * (fabsf(U * Az) + fabsf(V * Bz) + fabsf(W * Cz)) +
                   eU * fabsf(Az) + eV * fabsf(Bz) + eW * fabsf(Cz);
        float dt = (eT + fabsf(t) * eDet) / fabsf(det) + 2.0f * kUnit * fabsf(t);
        // certainly out of range?
        if (t < t_lo - dt || t > bound + dt + bound_err) return 0;
        bool range_sure = (t >= t_lo + dt) && (t <= bound - dt - bound_err);
        bool mixed = (U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f);
        if (!edge_sure || !range_sure) uncertain = true;
        if (mixed || !(t >= t_lo && t <= bound)) return 0;
        float inv = __fdiv_rn(1.0f, det);
        h.t = t; h.u = V * inv; h.v = W * inv; h.dt = dt;
        return 1;
    } else {
        if (det == 0.0f) return 0;
        float t = __fdiv_rn(T, det);
        if (!(t >= t_lo && t <= bound)) return 0;
        float inv = __fdiv_rn(1.0f, det);
        h.t = t; h.u = V * inv; h.v = W * inv; h.dt = 0.f;
        return 1;
    }
}

// ---- throughput variant --------------------------------------------------------------------
// Same watertight construction, organised for SIMT: the axis permutation + shear is folded into
// three per-ray coefficient vectors (x' = A.cx, y' = A.cy, z' = A.cz; two of the three
// coefficients of cx, cy are 1 and 0, so the products are exact), which removes the 3-way branch
// on the dominant axis that the select form compiles into, and t uses one IEEE reciprocal
// instead of two divisions.  Every vertex still maps to the same sheared coordinates in both
// triangles that share it, and the edge functions are still two rounded products and one
// subtraction, so shared edges stay watertight.  No error bounds: EXACT mode uses the version above.
struct RayWF {
    float3 o, cx, cy, cz;
};

// 1/x for the per-ray constants of the throughput path: MUFU.RCP + one Newton step (<= 1 ulp),
// straight-line.  The IEEE division it replaces carries a slow-path call that the few lanes of a
// refill executed one after the other (5 % of all issued instructions on the 1M-triangle soup).
// Any consistent per-ray constants keep the test watertight and the quantised boxes conservative.
__device__ __forceinline__ float rcp_fast(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return fmaf(r, fmaf(-x, r, 1.0f), r);
}

// Shear constants + dominant axis; the nine coefficients of RayWF are derived from them.
// (Keeping only these four per ray and rebuilding RayWF at every leaf phase was tried to relieve the
// 64-register traversal loop: ptxas spilled more, not less -- plain +0.8 %, EXACT +10 % slower.)
struct RayWC {
    float Sx, Sy, Sz;
    int kz;
};
__device__ __forceinline__ RayWC make_raywc(float3 d) {
    RayWC r;
    const float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
    r.kz = (ax > ay) ? (ax > az ? 0 : 2) : (ay > az ? 1 : 2);
    const float3 p = perm3(d, r.kz);
    r.Sz = rcp_fast(p.z);
    r.Sx = p.x * r.Sz;
    r.Sy = p.y * r.Sz;
    return r;
}
__device__ __forceinline__ RayWF expand_raywf(float3 o, const RayWC& r);

__device__ __forceinline__ RayWF make_raywf(float3 o, float3 d) { return expand_raywf(o, make_raywc(d)); }

__device__ __forceinline__ RayWF expand_raywf(float3 o, const RayWC& r) {
    RayWF f;
    f.o = o;
    const float nsx = -r.Sx, nsy = -r.Sy;
    if (r.kz == 0) {         // kx = 1, ky = 2
        f.cx = make_float3(nsx, 1.f, 0.f); f.cy = make_float3(nsy, 0.f, 1.f); f.cz = make_float3(r.Sz, 0.f, 0.f);
    } else if (r.kz == 1) {  // kx = 2, ky = 0
        f.cx = make_float3(0.f, nsx, 1.f); f.cy = make_float3(1.f, nsy, 0.f); f.cz = make_float3(0.f, r.Sz, 0.f);
    } else {                 // kx = 0, ky = 1
        f.cx = make_float3(1.f, 0.f, nsx); f.cy = make_float3(0.f, 1.f, nsy); f.cz = make_float3(0.f, 0.f, r.Sz);
    }
    return f;
}

__device__ __forceinline__ float dot3f(float3 a, float3 c) { return fmaf(a.z, c.z, fmaf(a.y, c.y, a.x * c.x)); }

__device__ __forceinline__ int tri_watertight_fast(const RayWF& r, float3 p0, float3 p1, float3 p2,
                                                   float t_lo, float bound, TriHit& h) {
    const float3 A = p0 - r.o, B = p1 - r.o, C = p2 - r.o;
    const float Ax = dot3f(A, r.cx), Ay = dot3f(A, r.cy);
    const float Bx = dot3f(B, r.cx), By = dot3f(B, r.cy);
    const float Cx = dot3f(C, r.cx), Cy = dot3f(C, r.cy);
    float U = __fsub_rn(__fmul_rn(Cx, By), __fmul_rn(Cy, Bx));
    float V = __fsub_rn(__fmul_rn(Ax, Cy), __fmul_rn(Ay, Cx));
    float W = __fsub_rn(__fmul_rn(Bx, Ay), __fmul_rn(By, Ax));
    if (U == 0.0f || V == 0.0f || W == 0.0f) {
        U = (float)(__dmul_rn((double)Cx, (double)By) - __dmul_rn((double)Cy, (double)Bx));
        V = (float)(__dmul_rn((double)Ax, (double)Cy) - __dmul_rn((double)Ay, (double)Cx));
        W = (float)(__dmul_rn((double)Bx, (double)Ay) - __dmul_rn((double)By, (double)Ax));
    }
    if (fminf(U, fminf(V, W)) < 0.0f && fmaxf(U, fmaxf(V, W)) > 0.0f) return 0;
    const float det = U + V + W;
    if (det == 0.0f) return 0;
    const float T = fmaf(U, dot3f(A, r.cz), fmaf(V, dot3f(B, r.cz), W * dot3f(C, r.cz)));
    const float inv = __frcp_rn(det);
    const float t = T * inv;
    if (!(t >= t_lo && t <= bound)) return 0;
    h.t = t; h.u = V * inv; h.v = W * inv; h.dt = 0.f;
    return 1;
}

// ---- exact-mode variant of the throughput test ----------------------------------------------
// tri_watertight_fast with forward error bounds (u = 2^-24), used by the persistent EXACT
// traversal.  It reports `uncertain` whenever the decision -- inside / outside, t in range, t
// against the current best -- could differ from the exact-arithmetic one; the caller then decides
// that triangle with mt_f64.  The bounds are built so that they scale with the triangle's
// footprint AROUND THE RAY, not with its distance from the ray origin:
//   * sheared coordinate of vertex P: |dPx|, |dPy| <= 5u nP, nP = |P.x|+|P.y|+|P.z| (P = p - o; the
//     shear coefficients are <= 1 in magnitude, Sz <= 1 ulp, Sx, Sy <= 1.5 ulp from rcp_fast);
//     e_P = 12u nP also absorbs the roundings of the two products of an edge function;
//   * U = Cx By - Cy Bx:  |dU| <= rC e_B + e_C (rB + 2 e_B), rP = |Px| + |Py|  -- ~ distance x
//     footprint, where the round-1 bound (M M) was ~ distance^2 (60x larger on the 1M soup: it
//     flagged 1 % of the rays and left 2.7 % of slack on the culling distance);
//   * t = Az + (V (Bz-Az) + W (Cz-Az)) / det: the ray origin cancels in Bz - Az, so the error of
//     the quotient scales with the triangle's depth extent; dt ~ 25u |t| for an ordinary hit.
// A result with dt > 2^-10 |t| (grazing) is reported uncertain, so the culling slack stays small.
__device__ __forceinline__ int tri_watertight_fast_exact(const RayWF& r, float3 p0, float3 p1, float3 p2,
                                                         float t_lo, float bound, float bound_err,
                                                         TriHit& h, bool& uncertain) {
    const float3 A = p0 - r.o, B = p1 - r.o, C = p2 - r.o;
    const float Ax = dot3f(A, r.cx), Ay = dot3f(A, r.cy);
    const float Bx = dot3f(B, r.cx), By = dot3f(B, r.cy);
    const float Cx = dot3f(C, r.cx), Cy = dot3f(C, r.cy);
    float U = __fsub_rn(__fmul_rn(Cx, By), __fmul_rn(Cy, Bx));
    float V = __fsub_rn(__fmul_rn(Ax, Cy), __fmul_rn(Ay, Cx));
    float W = __fsub_rn(__fmul_rn(Bx, Ay), __fmul_rn(By, Ax));
    constexpr float kE = 12.0f * kUnit;
    const float eA = kE * (fabsf(A.x) + fabsf(A.y) + fabsf(A.z));
    const float eB = kE * (fabsf(B.x) + fabsf(B.y) + fabsf(B.z));
    const float eC = kE * (fabsf(C.x) + fabsf(C.y) + fabsf(C.z));
    const float rA = fabsf(Ax) + fabsf(Ay), rB = fabsf(Bx) + fabsf(By), rC = fabsf(Cx) + fabsf(Cy);
    const float eU = fmaf(rC, eB, eC * fmaf(2.0f, eB, rB));
    const float eV = fmaf(rA, eC, eA * fmaf(2.0f, eC, rC));
    const float eW = fmaf(rB, eA, eB * fmaf(2.0f, eA, rA));
    const bool neg = (U < -eU) || (V < -eV) || (W < -eW);
    const bool pos = (U > eU) || (V > eV) || (W > eW);
    if (neg && pos) return 0;  // certainly outside
    const float det = U + V + W;
    const float eDet = eU + eV + eW + 4.0f * kUnit * (fabsf(U) + fabsf(V) + fabsf(W));
    if (!(fabsf(det) > eDet)) { uncertain = true; return 0; }
    const float Az = dot3f(A, r.cz), Bz = dot3f(B, r.cz), Cz = dot3f(C, r.cz);
    const float dB = Bz - Az, dC = Cz - Az;
    const float Tp = fmaf(V, dB, W * dC);
    const float inv = rcp_fast(det);
    const float tp = Tp * inv;
    const float t = Az + tp;
    const float e_dz = 6.0f * kUnit * (fabsf(Az) + fabsf(Bz) + fabsf(Cz));
    const float eTp = (fabsf(V) + fabsf(W)) * e_dz + eV * fabsf(dB) + eW * fabsf(dC) +
                      2.0f * kUnit * (fabsf(V * dB) + fabsf(W * dC));
    const float dt = 1.001f * (eTp + fabsf(tp) * eDet) * fabsf(inv) + 6.0f * kUnit * (fabsf(Az) + fabsf(tp));
    if (t < t_lo - dt || t > bound + dt + bound_err) return 0;  // certainly out of range
    const bool edge_sure = (fabsf(U) > eU) && (fabsf(V) > eV) && (fabsf(W) > eW);
    const bool range_sure = (t >= t_lo + dt) && (t <= bound - dt - bound_err);
    const bool tight = dt <= 0.0009765625f * fabsf(t);
    if (!edge_sure || !range_sure || !tight) { uncertain = true; return 0; }
    h.t = t; h.u = V * inv; h.v = W * inv; h.dt = dt;
    return 1;  // (edge_sure and not certainly outside => all three edge functions share a sign)
}

// ---- FP64 replay of the reference kernel -------------------------------------
__device__ __forceinline__ double ddot3(const double* x, const double* y) {
    return __dadd_rn(__dadd_rn(__dmul_rn(x[0], y[0]), __dmul_rn(x[1], y[1])), __dmul_rn(x[2], y[2]));
}
__device__ __forceinline__ void dcross3(const double* x, const double* y, double* r) {
    r[0] = __dsub_rn(__dmul_rn(x[1], y[2]), __dmul_rn(x[2], y[1]));
    r[1] = __dsub_rn(__dmul_rn(x[2], y[0]), __dmul_rn(x[0], y[2]));
    r[2] = __dsub_rn(__dmul_rn(x[0], y[1]), __dmul_rn(x[1], y[0]));
}

__device__ __forceinline__ int mt_f64(float3 fp0, float3 fp1, float3 fp2, const double* o,
                                      const double* d, double t_lo, double bound, double& t_out,
                                      double& u_out, double& v_out) {
    double p0[3] = {(double)fp0.x, (double)fp0.y, (double)fp0.z};
    double e1[3] = {__dsub_rn((double)fp1.x, p0[0]), __dsub_rn((double)fp1.y, p0[1]),
                    __dsub_rn((double)fp1.z, p0[2])};
    double e2[3] = {__dsub_rn((double)fp2.x, p0[0]), __dsub_rn((double)fp2.y, p0[1]),
                    __dsub_rn((double)fp2.z, p0[2])};
    double s[3] = {__dsub_rn(o[0], p0[0]), __dsub_rn(o[1], p0[1]), __dsub_rn(o[2], p0[2])};
    double q[3], r[3];
    dcross3(d, e2, q);
    dcross3(s, e1, r);
    double a = ddot3(e1, q);
    double e2r = ddot3(e2, r);
    double sq = ddot3(s, q);
    double rdr = ddot3(d, r);
    if (-kRefEps < a && a < kRefEps) return 0;
    double f = __ddiv_rn(1.0, a);
    double t = __dmul_rn(f, e2r);
    if (t > bound || t < t_lo) return 0;
    double u = __dmul_rn(f, sq);
    if (u < 0.0) return 0;
    double v = __dmul_rn(f, rdr);
    if (v < 0.0 || __dadd_rn(u, v) > 1.0) return 0;
    t_out = t; u_out = u; v_out = v;
    return 1;
}

}  // namespace prt
