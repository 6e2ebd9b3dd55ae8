// shading.cuh -- Philox4x32-10, samplers and BSDF device functions.
//
// RNG: the reference draws from Python `random`, np.random and ti.random with no
// seeding (main.py:19-20, samplers_debug.py:24, core/bsdf.py:31).  Here every
// random number is a pure function of (seed, pixel, sample, bounce, block):
//   counter = (pixel, sample, bounce, block), key = (seed lo, seed hi)
//   block 0: camera jitter x,y          (bounce 0 only; main.py:31-32)
//   block 1: bsdf u1,u2 | light triangle | Russian roulette
//   block 2: light point u,v | conductor fuzz | spare
// so any GPU count / wave size renders the same image, and the CPU oracle can
// consume the identical streams.  Uniforms are 24-bit: k>>8 * 2^-24, exact in
// FP32 and FP64.
#pragma once
#include "common.cuh"

namespace prt {

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}
__device__ __forceinline__ uint4 rng4(unsigned long long seed, uint32_t pixel, uint32_t sample,
                                      uint32_t bounce, uint32_t block) {
    return philox4x32_10(make_uint4(pixel, sample, bounce, block),
                         make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
}
__device__ __forceinline__ float u24(uint32_t k) { return (float)(k >> 8) * (1.0f / 16777216.0f); }
__device__ __forceinline__ uint32_t rand_index(uint32_t k, uint32_t n) {
    return (uint32_t)(((unsigned long long)k * n) >> 32);
}

constexpr float kPiOver4 = 0.78539816339744830961f;
constexpr float kPiOver2 = 1.57079632679489661923f;
constexpr float kInvPi = 0.31830988618379067154f;
constexpr float kPi = 3.14159265358979323846f;

// Shirley concentric map: mathematics/samplers.py:9-25 / samplers_debug.py:8-20
__device__ __forceinline__ float2 concentric_sample_disk(float u1, float u2) {
    float ox = 2.0f * u1 - 1.0f, oy = 2.0f * u2 - 1.0f;
    if (ox == 0.0f && oy == 0.0f) return make_float2(0.f, 0.f);
    float r, theta;
    if (fabsf(ox) > fabsf(oy)) {
        r = ox;
        theta = kPiOver4 * (oy / ox);
    } else {
        r = oy;
        theta = kPiOver2 - kPiOver4 * (ox / oy);
    }
    float s, c;
    sincosf(theta, &s, &c);
    return make_float2(r * c, r * s);
}

// cosine-weighted direction about n: mathematics/samplers.py:28-47 with the frame
// of mathematics/mat4_taichi.py:9-60 (x = n x Y, z = x x n; special-cased only
// for n.y == +-1 exactly, because the reference's EPS is 1.18e-38).
__device__ __forceinline__ float3 cosine_sample_hemisphere(float3 n_in, float u1, float u2) {
    float2 d = concentric_sample_disk(u1, u2);
    float z = sqrtf(fmaxf(0.0f, 1.0f - d.x * d.x - d.y * d.y));
    float3 n = normalize(n_in);  // exact: the n.y == +-1 special case below must fire as in the reference
    float3 r1, r2;
    if (fabsf(n.y - 1.0f) < 1.17549435e-38f) {
        r1 = make_float3(1.f, 0.f, 0.f); r2 = make_float3(0.f, 0.f, 1.f); n = make_float3(0.f, 1.f, 0.f);
    } else if (fabsf(n.y + 1.0f) < 1.17549435e-38f) {
        r1 = make_float3(1.f, 0.f, 0.f); r2 = make_float3(0.f, 0.f, 1.f); n = make_float3(0.f, -1.f, 0.f);
    } else {
        r1 = normalize_fast(cross(n, make_float3(0.f, 1.f, 0.f)));
        r2 = normalize_fast(cross(r1, n));
    }
    float3 w = r1 * d.x + r2 * d.y + n * z;
    return normalize_fast(w);
}

// core/bsdf_taichi.py:6-22
__device__ __forceinline__ float schlick(float cosine, float idx) {
    float r0 = (1.0f - idx) / (1.0f + idx);
    r0 = r0 * r0;
    float m = 1.0f - cosine;
    return r0 + (1.0f - r0) * (m * m * m * m * m);
}
__device__ __forceinline__ float3 reflect(float3 v, float3 n) { return v - n * (2.0f * dot(v, n)); }
__device__ __forceinline__ float3 refract(float3 v, float3 n, float eta) {
    float c = fminf(-dot(v, n), 1.0f);
    float3 perp = (v + n * c) * eta;
    float k = -sqrtf(fabsf(1.0f - dot(perp, perp)));
    return perp + n * k;
}
// mathematics/vec3_taichi.py:33-39 with explicit uniforms
__device__ __forceinline__ float3 in_unit_sphere(float ua, float ub, float uc) {
    float theta = ua * kPi * 2.0f;
    float phi = acosf(2.0f * ub - 1.0f);
    float r = cbrtf(uc);
    float st, ct, sp, cp;
    sincosf(theta, &st, &ct);
    sincosf(phi, &sp, &cp);
    return make_float3(r * sp * ct, r * sp * st, r * cp);
}

// Specular scattering, core/bsdf_taichi.py:45-86 (the oracle's orc_scatter_specular in FP32): `d` =
// incoming direction (any length), `ns` = shading normal on the incoming side; PRT_MAT_MIRROR =
// Metal with roughness 0, PRT_MAT_CONDUCTOR = Metal.scatter (:54-60, valid only if the fuzzed
// direction leaves on the normal's side), PRT_MAT_DIELECTRIC = Dielectric.scatter (:71-86).  wi is
// NOT normalised.  (u1, u2, u3): conductor = theta, v, r of random_in_unit_sphere; dielectric = u1 is
// the Fresnel draw.  shade_kernel and the known-answer entry point prt_eval_specular both call this.
__device__ __forceinline__ bool sample_specular(uint32_t type, float3 d, float3 ns, bool front, float ior,
                                                float roughness, float u1, float u2, float u3, float3& wi) {
    const float3 ud = normalize_fast(d);
    if (type == PRT_MAT_MIRROR) {
        wi = reflect(ud, ns);
        return true;
    }
    if (type == PRT_MAT_CONDUCTOR) {
        wi = reflect(ud, ns) + in_unit_sphere(u1, u2, u3) * roughness;
        return dot(wi, ns) > 0.0f;  // core/bsdf_taichi.py:58
    }
    const float ratio = front ? 1.0f / ior : ior;
    const float ct = fminf(-dot(ud, ns), 1.0f);
    const float st = sqrtf(1.0f - ct * ct);
    if (ratio * st > 1.0f || schlick(ct, ratio) > u1) wi = reflect(ud, ns);
    else wi = refract(ud, ns, ratio);
    return true;
}

}  // namespace prt
