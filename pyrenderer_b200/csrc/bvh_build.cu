// bvh_build.cu -- GPU LBVH build emitting the quantised 4-wide node layout (Node64).
//
// Pipeline (all on the device; the host only reads a queue length per tree level):
//   1. bounds_kernel     scene AABB (warp reduce + ordered-int atomics)
//   2. morton_kernel     30-bit Morton code of each triangle's box centre; 63-bit (21 bits per
//                        axis) when the 30-bit grid leaves more than 1/16 of the sorted
//                        neighbours in the same cell (clustered scenes), or on request
//   3. radix sort        (key, triangle) pairs, 4 x 8-bit stable LSD passes (radix_sort.cuh);
//                        63-bit keys: low word first, then the high word (stable)
//   4. hierarchy_kernel  Karras 2012 "Maximizing parallelism in the construction
//                        of BVHs, octrees and k-d trees": one thread per internal node
//   5. treelet_sah_kernel one warp per maximal subtree of <= 128 triangles: top-down SAH rebuild in shared
//                        memory (exact enumeration for 3..6 triangles, 8 / 16 bins above), then the warp
//                        refits its own subtree (boxes, SAH cost, leaf collapse, counts)
//      refit_kernel      bottom-up with arrival flags from the subtree ROOTS (and the leaves no subtree
//                        covers): boxes, SAH cost, SAH tree rotations and SAH leaf collapse (<= 7 tris)
//   6. emit_levels_kernel breadth-first collapse of the binary tree into 4-wide records, ONE
//                        cooperative launch (grid-wide barrier between levels), one 64-bit atomic
//                        per warp for the record / triangle reservations
//                        (expand the child of largest surface area), conservative 8-bit
//                        quantisation of the child boxes in the record's frame, leaf
//                        triangles copied to contiguous ranges
//
// Replaces: BVH.build / build_helper / sah_heuristic accelerators/bvh.py:70-215
// (recursive binned SAH on the CPU), Aggregator.update accelerators/aggregator.py:25-55
// (flat soup for zero-thickness primitives: no special path here, flat boxes
// quantise fine) and BVH.build accelerators/bvh_taichi.py:126-161.
#include <cooperative_groups.h>

#include <chrono>

#include "context.cuh"
#include "bvh.cuh"
#include "radix_sort.cuh"

namespace prt {

namespace {

// Scratch of one build, carved out of the context's grow-only arena (prt_ctx::build_arena): no
// cudaMalloc / cudaFree per build -- two dozen of them cost 30 ms of wall time around a 2 ms build.
struct BuildBuffers {
    uint32_t* keys[2] = {nullptr, nullptr};
    uint32_t* vals[2] = {nullptr, nullptr};
    uint32_t* keys_hi = nullptr;  // 63-bit Morton: high words in sorted order
    void* sort_tmp = nullptr;
    int* left = nullptr;     // [N-1]
    int* right = nullptr;    // [N-1]
    int* parent = nullptr;   // [2N-1]
    float4* bmin = nullptr;  // [2N-1]  w = SAH cost of the subtree
    float4* bmax = nullptr;  // [2N-1]
    uint32_t* tcount = nullptr;  // [2N-1] triangles below
    uint32_t* icount = nullptr;  // [2N-1] surviving records below (incl. self)
    uint8_t* collapsed = nullptr;  // [2N-1]
    unsigned int* flags = nullptr;  // [N-1]
    int* scene_box = nullptr;       // 6 ordered ints
    unsigned int* sort_check = nullptr;  // [2] inversions, equal neighbours
    int* queue = nullptr;           // emit: binary node of each wide record
    unsigned int* tails = nullptr;  // emit: queue tail, triangle tail, level begin, level end, depth
    int2* range = nullptr;          // [N-1] sorted positions covered by each internal node (treelets)
    int* roots = nullptr;           // [N-1] treelet roots
    unsigned int* n_roots = nullptr;
};

__device__ __forceinline__ int f2ord(float f) {  // order-preserving float -> int
    int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__device__ __forceinline__ void tri_box(const float4* v, uint32_t t, float3& lo, float3& hi) {
    float4 a = v[3ull * t], b = v[3ull * t + 1], c = v[3ull * t + 2];
    lo = make_float3(fminf(a.x, fminf(b.x, c.x)), fminf(a.y, fminf(b.y, c.y)), fminf(a.z, fminf(b.z, c.z)));
    hi = make_float3(fmaxf(a.x, fmaxf(b.x, c.x)), fmaxf(a.y, fmaxf(b.y, c.y)), fmaxf(a.z, fmaxf(b.z, c.z)));
}

__global__ void init_box_kernel(int* box) {
    if (threadIdx.x < 3) box[threadIdx.x] = f2ord(3.4e38f);
    else if (threadIdx.x < 6) box[threadIdx.x] = f2ord(-3.4e38f);
}

__global__ void bounds_kernel(const float4* __restrict__ v, uint32_t nt, int* box) {
    float3 lo = make_float3(3.4e38f, 3.4e38f, 3.4e38f), hi = make_float3(-3.4e38f, -3.4e38f, -3.4e38f);
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < nt; t += gridDim.x * blockDim.x) {
        float3 a, b;
        tri_box(v, t, a, b);
        lo = make_float3(fminf(lo.x, a.x), fminf(lo.y, a.y), fminf(lo.z, a.z));
        hi = make_float3(fmaxf(hi.x, b.x), fmaxf(hi.y, b.y), fmaxf(hi.z, b.z));
    }
    for (int o = 16; o > 0; o >>= 1) {
        lo.x = fminf(lo.x, __shfl_xor_sync(~0u, lo.x, o)); lo.y = fminf(lo.y, __shfl_xor_sync(~0u, lo.y, o));
        lo.z = fminf(lo.z, __shfl_xor_sync(~0u, lo.z, o)); hi.x = fmaxf(hi.x, __shfl_xor_sync(~0u, hi.x, o));
        hi.y = fmaxf(hi.y, __shfl_xor_sync(~0u, hi.y, o)); hi.z = fmaxf(hi.z, __shfl_xor_sync(~0u, hi.z, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(box + 0, f2ord(lo.x)); atomicMin(box + 1, f2ord(lo.y)); atomicMin(box + 2, f2ord(lo.z));
        atomicMax(box + 3, f2ord(hi.x)); atomicMax(box + 4, f2ord(hi.y)); atomicMax(box + 5, f2ord(hi.z));
    }
}

__device__ __forceinline__ uint32_t expand10(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}

__global__ void morton_kernel(const float4* __restrict__ v, uint32_t nt, const int* __restrict__ box,
                              uint32_t* keys, uint32_t* vals) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nt) return;
    float3 smin = make_float3(ord2f(box[0]), ord2f(box[1]), ord2f(box[2]));
    float3 smax = make_float3(ord2f(box[3]), ord2f(box[4]), ord2f(box[5]));
    float3 lo, hi;
    tri_box(v, t, lo, hi);
    float ex = smax.x - smin.x, ey = smax.y - smin.y, ez = smax.z - smin.z;
    float cx = ex > 0.f ? (0.5f * (lo.x + hi.x) - smin.x) / ex : 0.f;
    float cy = ey > 0.f ? (0.5f * (lo.y + hi.y) - smin.y) / ey : 0.f;
    float cz = ez > 0.f ? (0.5f * (lo.z + hi.z) - smin.z) / ez : 0.f;
    uint32_t x = (uint32_t)fminf(fmaxf(cx * 1024.f, 0.f), 1023.f);
    uint32_t y = (uint32_t)fminf(fmaxf(cy * 1024.f, 0.f), 1023.f);
    uint32_t z = (uint32_t)fminf(fmaxf(cz * 1024.f, 0.f), 1023.f);
    keys[t] = (expand10(x) << 2) | (expand10(y) << 1) | expand10(z);
    vals[t] = t;
}

__device__ __forceinline__ unsigned long long expand21(unsigned long long x) {
    x &= 0x1fffffull;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

// 63-bit code, 21 bits per axis: low word -> keys (sorted first), high word -> hi_by_tri
__global__ void morton63_kernel(const float4* __restrict__ v, uint32_t nt, const int* __restrict__ box,
                                uint32_t* keys, uint32_t* vals, uint32_t* lo_by_tri, uint32_t* hi_by_tri) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nt) return;
    float3 smin = make_float3(ord2f(box[0]), ord2f(box[1]), ord2f(box[2]));
    float3 smax = make_float3(ord2f(box[3]), ord2f(box[4]), ord2f(box[5]));
    float3 lo, hi;
    tri_box(v, t, lo, hi);
    // f64: 21 bits per axis are more than an f32 quotient resolves near the far end of the scene box
    const double ex = (double)smax.x - smin.x, ey = (double)smax.y - smin.y, ez = (double)smax.z - smin.z;
    const double cx = ex > 0. ? (0.5 * ((double)lo.x + hi.x) - smin.x) / ex : 0.;
    const double cy = ey > 0. ? (0.5 * ((double)lo.y + hi.y) - smin.y) / ey : 0.;
    const double cz = ez > 0. ? (0.5 * ((double)lo.z + hi.z) - smin.z) / ez : 0.;
    const unsigned long long x = (unsigned long long)fmin(fmax(cx * 2097152., 0.), 2097151.);
    const unsigned long long y = (unsigned long long)fmin(fmax(cy * 2097152., 0.), 2097151.);
    const unsigned long long z = (unsigned long long)fmin(fmax(cz * 2097152., 0.), 2097151.);
    const unsigned long long k = (expand21(x) << 2) | (expand21(y) << 1) | expand21(z);
    keys[t] = (uint32_t)k;
    lo_by_tri[t] = (uint32_t)k;
    hi_by_tri[t] = (uint32_t)(k >> 32);
    vals[t] = t;
}

__global__ void gather_kernel(const uint32_t* __restrict__ src, const uint32_t* __restrict__ idx, uint32_t n, uint32_t* dst) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[idx[i]];
}

// keys_hi == nullptr: 30-bit keys.  Equal keys are told apart by their sorted position.
__device__ __forceinline__ int delta(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ keys_hi, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    if (keys_hi) {
        const uint32_t ha = keys_hi[i], hb = keys_hi[j];
        if (ha != hb) return __clz(ha ^ hb);
        const uint32_t a = keys[i], b = keys[j];
        if (a != b) return 32 + __clz(a ^ b);
        return 64 + __clz((uint32_t)i ^ (uint32_t)j);
    }
    uint32_t a = keys[i], b = keys[j];
    if (a == b) return 32 + __clz((uint32_t)i ^ (uint32_t)j);
    return __clz(a ^ b);
}

// self-check of the radix sort: out[0] = inversions (reported in prt_bvh_stats), out[1] = equal
// neighbours (how many triangles the key grid could not separate: decides 30 vs 63 bits)
__global__ void check_sorted_kernel(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ keys_hi, uint32_t n,
                                    unsigned int* out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    bool inv = false, eq = false;
    if (i + 1 < n) {
        const unsigned long long a = ((unsigned long long)(keys_hi ? keys_hi[i] : 0u) << 32) | keys[i];
        const unsigned long long b = ((unsigned long long)(keys_hi ? keys_hi[i + 1] : 0u) << 32) | keys[i + 1];
        inv = a > b; eq = a == b;
    }
    const unsigned mi = __ballot_sync(0xffffffffu, inv), me = __ballot_sync(0xffffffffu, eq);
    if ((threadIdx.x & 31) == 0) {
        if (mi) atomicAdd(out, (unsigned)__popc(mi));
        if (me) atomicAdd(out + 1, (unsigned)__popc(me));
    }
}

// node ids: internal i -> i (0..n-2), leaf j -> (n-1)+j
__global__ void hierarchy_kernel(const uint32_t* __restrict__ keys_lo, const uint32_t* __restrict__ keys_hi, int n,
                                 int* left, int* right, int* parent, int2* range) {
    struct { const uint32_t* lo; const uint32_t* hi; } keys = {keys_lo, keys_hi};
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    int d = (delta(keys.lo, keys.hi, n, i, i + 1) - delta(keys.lo, keys.hi, n, i, i - 1)) >= 0 ? 1 : -1;
    int dmin = delta(keys.lo, keys.hi, n, i, i - d);
    int lmax = 2;
    while (delta(keys.lo, keys.hi, n, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (delta(keys.lo, keys.hi, n, i, i + (l + t) * d) > dmin) l += t;
    int j = i + l * d;
    int dnode = delta(keys.lo, keys.hi, n, i, j);
    int s = 0, t = l;
    do {
        t = (t + 1) >> 1;
        if (delta(keys.lo, keys.hi, n, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    int gamma = i + s * d + min(d, 0);
    int lc = (min(i, j) == gamma) ? (n - 1) + gamma : gamma;
    int rc = (max(i, j) == gamma + 1) ? (n - 1) + gamma + 1 : gamma + 1;
    left[i] = lc;
    right[i] = rc;
    parent[lc] = i;
    parent[rc] = i;
    if (i == 0) parent[0] = -1;
    range[i] = make_int2(min(i, j), max(i, j));  // sorted positions covered by node i (contiguous)
}

__device__ __forceinline__ float box_area(float3 lo, float3 hi) {
    float ex = hi.x - lo.x, ey = hi.y - lo.y, ez = hi.z - lo.z;
    return 2.0f * (ex * ey + ey * ez + ez * ex);
}

struct RefitParams {
    int n;
    uint32_t max_leaf;
    float cn, ct;
    int rotations;
};

// ---- SAH treelets -------------------------------------------------------------------------
// The Morton hierarchy is a spatial-median tree: fine at the top, but inside a neighbourhood of a
// few dozen triangles a surface-area-heuristic builder separates them better.  Every MAXIMAL
// subtree with at most kTreelet (128) triangles is therefore rebuilt top-down with binned SAH (16 bins
// x 3 axes) by one warp, entirely in shared memory; it reuses the subtree's own node names and
// its own range of sorted positions, so nothing outside the subtree changes.  Measured on the
// 1M-triangle soup (profiles/bvh_quality.c): record visits per ray 39.5 -> 37.8 (64) .. 37.4 (256), most of what a
// full SAH build would give (37.2).  The kernel is instruction- and shared-memory-bound (profiles/r2_sweeps.txt):
// what made it faster were fewer instructions per split (ranges of <= 32 triangles kept in registers, ranges
// of <= 6 enumerated) and conflict-free bin rows, not more warps or shorter dependency chains.
#ifndef PRT_TREELET
#define PRT_TREELET 128
#endif
constexpr int kTreelet = PRT_TREELET;  // <= 255 (8-bit permutation)
constexpr int kTreeletWarps = 4;  // warps per block
#ifndef PRT_TREELET_ENUM
#define PRT_TREELET_ENUM 1  // ranges of 3 .. 6 triangles: exact SAH over all their two-way partitions
#endif
// lane -> its candidate left set among the 3 / 7 / 15 / 31 two-way partitions of 3 / 4 / 5 / 6 items (6 bits each,
// packed c = 3 first): every subset of at most half the items; of the halves, those that contain item 0
__device__ const uint32_t kEnumMasks[32] = {0x041041u, 0x082082u, 0x104104u, 0x208200u, 0x4100c0u, 0x803140u, 0x0c5240u, 0x149000u, 0x251000u, 0x446000u, 0x84a000u, 0x192000u, 0x28c000u, 0x494000u, 0x898000u, 0x300000u, 0x500000u, 0x900000u, 0x600000u, 0xa00000u, 0xc00000u, 0x1c0000u, 0x2c0000u, 0x4c0000u, 0x8c0000u, 0x340000u, 0x540000u, 0x940000u, 0x640000u, 0xa40000u, 0xc40000u, 0x000000u};
#ifndef PRT_TREELET_FINE
#define PRT_TREELET_FINE 32  // one-chunk ranges with more triangles than this use 16 bins instead of 8
#endif
constexpr int kTreeletStack = 16;  // the larger child is pushed first, so the stack holds <= log2(kTreelet) + 1 ranges
constexpr int kTreeletWords = (kTreelet + 31) / 32;

__global__ void treelet_roots_kernel(const int2* __restrict__ range, const int* __restrict__ parent, int n,
                                     int* roots, unsigned int* n_roots) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int2 r = range[i];
    const int size = r.y - r.x + 1;
    if (size < 3 || size > kTreelet) return;  // 2 triangles: nothing to choose
    if (i != 0) {
        const int2 pr = range[parent[i]];
        if (pr.y - pr.x + 1 <= kTreelet) return;  // not maximal
    }
    roots[atomicAdd(n_roots, 1u)] = i;
}

struct TreeletShared {
    float lo[kTreelet][3], hi[kTreelet][3];   // triangle boxes
    uint32_t tri[kTreelet];                   // triangle ids (vals entries)
    uint8_t order[kTreelet], tmp[kTreelet];   // permutation being partitioned
    uint32_t pk[kTreelet];                    // per position of the current range: bin per axis, 8 bits each
    int names[kTreelet];                      // internal node names available to this subtree; [0] = root
    uint32_t link[kTreelet];                  // per names index: children, 16 bits each (0x8000 | position = leaf, else names index)
    int stack[kTreeletStack][3];              // begin, end, names index
    int bins[48][7];                          // per (axis * NB + bin): box lo.xyz, hi.xyz (float bits), count; odd stride: the (axis, bin) lanes hit different banks
};

__device__ __forceinline__ float area3(const float lo[3], const float hi[3]) {
    const float x = hi[0] - lo[0], y = hi[1] - lo[1], z = hi[2] - lo[2];
    return 2.0f * (x * y + y * z + z * x);
}

// Best binned-SAH plane of order[b, e) with NB bins per axis.  The bin boxes are accumulated in shared
// memory, lane (axis, bin) then takes its bin's box into registers; prefix / suffix unions over the NB bins of an axis are shuffle
// scans inside an NB-lane group; candidate plane j = "bins <= j go left".  NB = 16: two rounds
// (axes x,y then z); NB = 8: the 24 (axis, bin) pairs fit one round.  Returns cost (inf: none).
template <int NB>
__device__ __forceinline__ void treelet_bins_clear(TreeletShared& S, int lane) {
    for (int k = lane; k < 3 * NB; k += 32) {
        int* bb = S.bins[k];
        bb[0] = bb[1] = bb[2] = 0x7f800000;              // +inf (coordinates are >= +0: float bits order like ints)
        bb[3] = bb[4] = bb[5] = 0;
        bb[6] = 0;
    }
}
// one triangle into its bin of every axis; returns the three bin indices, 8 bits each
template <int NB>
__device__ __forceinline__ uint32_t treelet_bin_add(TreeletShared& S, const float lo[3], const float hi[3],
                                                    const float cmin[3], const float scale[3]) {
    const int l0 = __float_as_int(lo[0]), l1 = __float_as_int(lo[1]), l2 = __float_as_int(lo[2]);
    const int h0 = __float_as_int(hi[0]), h1 = __float_as_int(hi[1]), h2 = __float_as_int(hi[2]);
    uint32_t pk = 0;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        int j = (int)((0.5f * (lo[a] + hi[a]) - cmin[a]) * scale[a]);
        j = j > NB - 1 ? NB - 1 : j;
        pk |= (uint32_t)j << (8 * a);
        if (scale[a] > 0.0f) {
            int* bb = S.bins[a * NB + j];
            atomicMin(bb + 0, l0); atomicMin(bb + 1, l1); atomicMin(bb + 2, l2);
            atomicMax(bb + 3, h0); atomicMax(bb + 4, h1); atomicMax(bb + 5, h2);
            atomicAdd(bb + 6, 1);
        }
    }
    return pk;
}
template <int NB>
__device__ __forceinline__ float treelet_pick(TreeletShared& S, const float scale[3], int lane, int& best_slot);

template <int NB>
__device__ __forceinline__ float treelet_best_split(TreeletShared& S, int b, int e, const float cmin[3],
                                                    const float scale[3], int lane, int& best_slot) {
    // Bin boxes by shared-memory atomics on order-preserving ints (min / max do not depend on the order
    // of arrival, so the result is the same as a scan): one triangle per lane and step.  (Round 1 had
    // every (axis, bin) lane scan the whole range for its members -- 32 x c loop trips for c useful
    // ones, the bulk of the 0.9 ms the treelets took at 1M triangles.)
    treelet_bins_clear<NB>(S, lane);
    __syncwarp();
    for (int k = b + lane; k < e; k += 32) {
        const int q = S.order[k];
        S.pk[k] = treelet_bin_add<NB>(S, S.lo[q], S.hi[q], cmin, scale);
    }
    __syncwarp();
    return treelet_pick<NB>(S, scale, lane, best_slot);
}

template <int NB>
__device__ __forceinline__ float treelet_pick(TreeletShared& S, const float scale[3], int lane, int& best_slot) {
    const unsigned FULL = 0xffffffffu;
    const float inf = __int_as_float(0x7f800000);
    constexpr int kAxesPerRound = 32 / NB;             // 2 or 4 (4th group idle)
    constexpr int kRounds = NB == 16 ? 2 : 1;
    float best = inf;
    best_slot = -1;
#pragma unroll
    for (int round = 0; round < kRounds; ++round) {
        const int a = lane / NB + kAxesPerRound * round, bin = lane % NB;
        const float sc_a = a == 0 ? scale[0] : (a == 1 ? scale[1] : (a == 2 ? scale[2] : 0.0f));
        float l0 = inf, l1 = inf, l2 = inf, h0 = -inf, h1 = -inf, h2 = -inf;
        int cnt = 0;
        if (sc_a > 0.0f) {
            const int* bb = S.bins[a * NB + bin];
            cnt = bb[6];
            if (cnt > 0) {
                l0 = __int_as_float(bb[0]); l1 = __int_as_float(bb[1]); l2 = __int_as_float(bb[2]);
                h0 = __int_as_float(bb[3]); h1 = __int_as_float(bb[4]); h2 = __int_as_float(bb[5]);
            }
        }
        // inclusive prefix (bins 0..bin) and suffix (bins bin..NB-1) within the NB-lane group
        float p0 = l0, p1 = l1, p2 = l2, q0 = h0, q1 = h1, q2 = h2;
        float s0 = l0, s1 = l1, s2 = l2, t0 = h0, t1 = h1, t2 = h2;
        int pc = cnt, sc2 = cnt;
#pragma unroll
        for (int o = 1; o < NB; o <<= 1) {
            const float a0 = __shfl_up_sync(FULL, p0, o, NB), a1 = __shfl_up_sync(FULL, p1, o, NB), a2 = __shfl_up_sync(FULL, p2, o, NB);
            const float b0 = __shfl_up_sync(FULL, q0, o, NB), b1 = __shfl_up_sync(FULL, q1, o, NB), b2 = __shfl_up_sync(FULL, q2, o, NB);
            const int ac = __shfl_up_sync(FULL, pc, o, NB);
            if (bin >= o) { p0 = fminf(p0, a0); p1 = fminf(p1, a1); p2 = fminf(p2, a2); q0 = fmaxf(q0, b0); q1 = fmaxf(q1, b1); q2 = fmaxf(q2, b2); pc += ac; }
            const float c0 = __shfl_down_sync(FULL, s0, o, NB), c1 = __shfl_down_sync(FULL, s1, o, NB), c2 = __shfl_down_sync(FULL, s2, o, NB);
            const float d0 = __shfl_down_sync(FULL, t0, o, NB), d1 = __shfl_down_sync(FULL, t1, o, NB), d2 = __shfl_down_sync(FULL, t2, o, NB);
            const int dc = __shfl_down_sync(FULL, sc2, o, NB);
            if (bin + o < NB) { s0 = fminf(s0, c0); s1 = fminf(s1, c1); s2 = fminf(s2, c2); t0 = fmaxf(t0, d0); t1 = fmaxf(t1, d1); t2 = fmaxf(t2, d2); sc2 += dc; }
        }
        // right side of candidate `bin` = suffix of bin + 1
        const float r0 = __shfl_down_sync(FULL, s0, 1, NB), r1 = __shfl_down_sync(FULL, s1, 1, NB), r2 = __shfl_down_sync(FULL, s2, 1, NB);
        const float u0 = __shfl_down_sync(FULL, t0, 1, NB), u1 = __shfl_down_sync(FULL, t1, 1, NB), u2 = __shfl_down_sync(FULL, t2, 1, NB);
        const int rc = __shfl_down_sync(FULL, sc2, 1, NB);
        if (a < 3 && bin < NB - 1 && pc > 0 && rc > 0) {
            const float pl[3] = {p0, p1, p2}, ph[3] = {q0, q1, q2}, rl[3] = {r0, r1, r2}, rh[3] = {u0, u1, u2};
            const float cost = area3(pl, ph) * (float)pc + area3(rl, rh) * (float)rc;
            if (cost < best) { best = cost; best_slot = a * NB + bin; }
        }
    }
    // lowest cost, lowest slot among equal costs: two warp reductions (costs are >= 0 or +inf, so their
    // bit patterns order like unsigned integers)
    const unsigned key = __float_as_uint(best);
    const unsigned kmin = __reduce_min_sync(FULL, key);
    const int smin = __reduce_min_sync(FULL, (key == kmin && best_slot >= 0) ? best_slot : 0x7fffffff);
    best_slot = smin == 0x7fffffff ? -1 : smin;
    return __uint_as_float(kmin);
}

// FUSED: the warp that built a subtree also refits it -- boxes, SAH cost, leaf collapse, triangle and
// record counts of every node of the subtree, exactly what refit_kernel derives bottom-up (update_node), and
// marks the subtree's leaves as covered (collapsed[leaf] = 2).  refit_kernel then starts at the subtree
// ROOTS: 1 % of the leaves' arrival-flag chains are left (1M triangles: refit 0.43 -> 0.1 ms).  SAH rotations
// are not applied inside a subtree the SAH builder has just produced.
struct TreeletOut {
    float4* bmin;
    float4* bmax;
    uint32_t* tcount;
    uint32_t* icount;
    uint8_t* collapsed;
};

#ifndef PRT_TREELET_MIN_BLOCKS
#define PRT_TREELET_MIN_BLOCKS 7
#endif
template <bool FUSED>
__global__ void __launch_bounds__(32 * kTreeletWarps, PRT_TREELET_MIN_BLOCKS)
treelet_sah_kernel(const float4* __restrict__ verts, uint32_t* vals, const int2* __restrict__ range,
                   int* left, int* right, int* parent, int n, const int* __restrict__ roots,
                   const unsigned int* __restrict__ n_roots, TreeletOut O, RefitParams P) {
    __shared__ TreeletShared sh_all[kTreeletWarps];
    TreeletShared& S = sh_all[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    const unsigned FULL = 0xffffffffu;
    const unsigned nr = *n_roots;
    const float inf = __int_as_float(0x7f800000);
    const uint32_t enum_masks = kEnumMasks[lane];
    for (unsigned ri = blockIdx.x * kTreeletWarps + (threadIdx.x >> 5); ri < nr; ri += gridDim.x * kTreeletWarps) {
        const int root = roots[ri];
        const int2 rg = range[root];
        const int first = rg.x, m = rg.y - rg.x + 1;
        __syncwarp();
        // triangles of the subtree
        // Boxes are kept RELATIVE to the subtree's min corner: every coordinate is >= +0, so float bits order like
        // ints and the shared-memory atomics / REDUX below need no order-preserving conversion (54 of ~400
        // instructions per split).  They only steer the SAH decisions; the boxes that are written out come from
        // the vertices again (FUSED pass).
        float tmin[3] = {inf, inf, inf};
        for (int k = lane; k < m; k += 32) {
            const uint32_t t = vals[first + k];
            float3 a, b;
            tri_box(verts, t, a, b);
            S.lo[k][0] = a.x; S.lo[k][1] = a.y; S.lo[k][2] = a.z;
            S.hi[k][0] = b.x; S.hi[k][1] = b.y; S.hi[k][2] = b.z;
            S.tri[k] = t;
            S.order[k] = (uint8_t)k;
            tmin[0] = fminf(tmin[0], a.x); tmin[1] = fminf(tmin[1], a.y); tmin[2] = fminf(tmin[2], a.z);
        }
#pragma unroll
        for (int a = 0; a < 3; ++a) tmin[a] = ord2f(__reduce_min_sync(FULL, f2ord(tmin[a])));
        for (int k = lane; k < m; k += 32) {  // (each lane rewrites what it wrote)
#pragma unroll
            for (int a = 0; a < 3; ++a) { S.lo[k][a] -= tmin[a]; S.hi[k][a] -= tmin[a]; }
        }
        // internal node names inside the subtree: node i (first <= i <= last) whose range lies inside
        int n_names = 1;
        if (lane == 0) S.names[0] = root;
        for (int k0 = 0; k0 < m; k0 += 32) {
            const int i = first + k0 + lane;
            bool in = false;
            if (k0 + lane < m && i < n - 1 && i != root) {
                const int2 r = range[i];
                in = r.x >= rg.x && r.y <= rg.y;
            }
            const unsigned bm = __ballot_sync(FULL, in);
            if (in) S.names[n_names + __popc(bm & ((1u << lane) - 1u))] = i;
            n_names += __popc(bm);
        }
        __syncwarp();
        // (n_names == m - 1 by construction)
        int next_name = 1, sp = 0;
        if (lane == 0) { S.stack[0][0] = 0; S.stack[0][1] = m; S.stack[0][2] = 0; }
        sp = 1;
        __syncwarp();
        while (sp > 0) {
            --sp;
            const int b = S.stack[sp][0], e = S.stack[sp][1], ni = S.stack[sp][2];
            const int name = S.names[ni];
            const int c = e - b;
            __syncwarp();
            int mid = b + 1;
#if PRT_TREELET_ENUM
            if (c <= 6 && c > 2) {
                // ---- 3 .. 6 triangles (40 % of all splits): every way to split the set in two -- 3, 7, 15 or 31 of them,
                // kEnumMasks -- is costed exactly, one per lane, instead of going through bins, atomics and scans
                const unsigned lmask = (enum_masks >> (6 * (c - 3))) & 63u;  // this lane's left set (0: no candidate)
                float ll[3] = {inf, inf, inf}, lh[3] = {-inf, -inf, -inf}, rl[3] = {inf, inf, inf}, rh[3] = {-inf, -inf, -inf};
#pragma unroll
                for (int i = 0; i < 6; ++i) {
                    if (i < c) {
                        const int qi = S.order[b + i];  // (broadcast loads)
                        const bool in = (lmask >> i) & 1u;
#pragma unroll
                        for (int a = 0; a < 3; ++a) {
                            const float l = S.lo[qi][a], h = S.hi[qi][a];
                            ll[a] = fminf(ll[a], in ? l : inf); lh[a] = fmaxf(lh[a], in ? h : -inf);
                            rl[a] = fminf(rl[a], in ? inf : l); rh[a] = fmaxf(rh[a], in ? -inf : h);
                        }
                    }
                }
                const int myq = S.order[b + (lane < c ? lane : 0)];
                const int nl = __popc(lmask);
                const float cost = lmask ? area3(ll, lh) * (float)nl + area3(rl, rh) * (float)(c - nl) : inf;
                const unsigned key = __float_as_uint(cost);  // (costs are >= 0: bit patterns order like unsigned integers)
                const unsigned kmin = __reduce_min_sync(FULL, key);
                const int win = __reduce_min_sync(FULL, (lmask && key == kmin) ? lane : 31);
                const unsigned wmask = __shfl_sync(FULL, lmask, win);
                const int wl = __popc(wmask);
                const unsigned below = (1u << lane) - 1u;
                const int pos = ((wmask >> lane) & 1u) ? __popc(wmask & below) : wl + __popc(~wmask & below);
                __syncwarp();
                if (lane < c) S.order[b + pos] = (uint8_t)myq;  // stable on both sides
                mid = b + wl;
            } else
#endif
            if (c > 2 && c <= 32) {
                // ---- one triangle per lane: box, centroid and bins stay in registers from the centroid bounds to the
                // partition (most splits of a subtree are of this size; same bins, same plane, same order as below)
                const bool has = lane < c;
                int q = 0;
                float lo[3] = {inf, inf, inf}, hi[3] = {-inf, -inf, -inf};
                if (has) {
                    q = S.order[b + lane];
#pragma unroll
                    for (int a = 0; a < 3; ++a) { lo[a] = S.lo[q][a]; hi[a] = S.hi[q][a]; }
                }
                float cmin[3], cmax[3], scale[3];
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    const float cc = 0.5f * (lo[a] + hi[a]);  // (idle lanes: inf + -inf = NaN, replaced below)
                    cmin[a] = __int_as_float(__reduce_min_sync(FULL, __float_as_int(has ? cc : inf)));
                    cmax[a] = __int_as_float(__reduce_max_sync(FULL, __float_as_int(has ? cc : 0.0f)));
                    scale[a] = cmax[a] > cmin[a] ? __fdividef(8.0f, cmax[a] - cmin[a]) : 0.0f;
                }
                uint32_t pk = 0;
                int best_slot, nbs = 3;  // log2(bins)
                if (c > PRT_TREELET_FINE) {
                    nbs = 4;
#pragma unroll
                    for (int a = 0; a < 3; ++a) scale[a] *= 2.0f;
                    treelet_bins_clear<16>(S, lane);
                    __syncwarp();
                    if (has) pk = treelet_bin_add<16>(S, lo, hi, cmin, scale);
                    __syncwarp();
                    treelet_pick<16>(S, scale, lane, best_slot);
                } else {
                    treelet_bins_clear<8>(S, lane);
                    __syncwarp();
                    if (has) pk = treelet_bin_add<8>(S, lo, hi, cmin, scale);
                    __syncwarp();
                    treelet_pick<8>(S, scale, lane, best_slot);
                }
                if (best_slot >= 0) {
                    const int a = best_slot >> nbs, j = best_slot & ((1 << nbs) - 1);
                    const bool go_left = has && (int)((pk >> (8 * a)) & 0xffu) <= j;
                    const unsigned lm = __ballot_sync(FULL, go_left), vm = c == 32 ? FULL : (1u << c) - 1u;
                    const unsigned lt = (1u << lane) - 1u;
                    const int nleft = __popc(lm);
                    const int pos = go_left ? __popc(lm & lt) : nleft + __popc(~lm & vm & lt);
                    if (has) S.order[b + pos] = (uint8_t)q;  // stable; every lane read its entry above
                    mid = b + nleft;
                } else {
                    mid = b + c / 2;  // coincident centroids: split the range in the middle
                }
            } else if (c > 2) {
                // centroid bounds
                float cmin[3] = {inf, inf, inf}, cmax[3] = {-inf, -inf, -inf};
                for (int k = b + lane; k < e; k += 32) {
                    const int q = S.order[k];
#pragma unroll
                    for (int a = 0; a < 3; ++a) {
                        const float cc = 0.5f * (S.lo[q][a] + S.hi[q][a]);
                        cmin[a] = fminf(cmin[a], cc); cmax[a] = fmaxf(cmax[a], cc);
                    }
                }
#pragma unroll
                for (int a = 0; a < 3; ++a) {  // warp min / max through order-preserving ints (REDUX)
                    cmin[a] = __int_as_float(__reduce_min_sync(FULL, __float_as_int(cmin[a])));
                    cmax[a] = __int_as_float(__reduce_max_sync(FULL, __float_as_int(cmax[a])));  // (-inf of an idle lane: a negative int)
                }
                // 16 bins for the large ranges near the treelet root (8, one round, in the one-chunk path above)
                constexpr int nb = 16;
                float scale[3];
#pragma unroll
                for (int a = 0; a < 3; ++a) scale[a] = cmax[a] > cmin[a] ? __fdividef((float)nb, cmax[a] - cmin[a]) : 0.0f;
                int best_slot;
                treelet_best_split<16>(S, b, e, cmin, scale, lane, best_slot);
                if (best_slot >= 0) {
                    const int a = best_slot / nb, j = best_slot % nb;
                    // stable partition of order[b, e) by bin <= j
                    int nleft = 0;
                    for (int k0 = b; k0 < e; k0 += 32) {
                        const int k = k0 + lane;
                        bool go_left = false;
                        int q = 0;
                        if (k < e) {
                            q = S.order[k];
                            go_left = (int)((S.pk[k] >> (8 * a)) & 0xffu) <= j;
                        }
                        const unsigned lm = __ballot_sync(FULL, go_left);
                        if (go_left) S.tmp[b + nleft + __popc(lm & ((1u << lane) - 1u))] = (uint8_t)q;
                        nleft += __popc(lm);
                    }
                    int nright = 0;
                    for (int k0 = b; k0 < e; k0 += 32) {
                        const int k = k0 + lane;
                        bool go_right = false;
                        int q = 0;
                        if (k < e) {
                            q = S.order[k];
                            go_right = (int)((S.pk[k] >> (8 * a)) & 0xffu) > j;
                        }
                        const unsigned rm = __ballot_sync(FULL, go_right);
                        if (go_right) S.tmp[b + nleft + nright + __popc(rm & ((1u << lane) - 1u))] = (uint8_t)q;
                        nright += __popc(rm);
                    }
                    __syncwarp();
                    for (int k = b + lane; k < e; k += 32) S.order[k] = S.tmp[k];
                    mid = b + nleft;
                } else {
                    mid = b + c / 2;  // coincident centroids: split the range in the middle
                }
            }
            __syncwarp();
            // children: a range of one triangle is the leaf named (n-1) + its final sorted position
            const int lsize = mid - b, rsize = e - mid;
            const int lidx = lsize > 1 ? next_name++ : -1, ridx = rsize > 1 ? next_name++ : -1;  // names indices
            if (lane == 0) {
                const int lname = lsize == 1 ? (n - 1) + first + b : S.names[lidx];
                const int rname = rsize == 1 ? (n - 1) + first + mid : S.names[ridx];
                left[name] = lname; right[name] = rname;
                parent[lname] = name; parent[rname] = name;
                S.link[ni] = (lsize == 1 ? 0x8000u | (uint32_t)b : (uint32_t)lidx) |
                             ((rsize == 1 ? 0x8000u | (uint32_t)mid : (uint32_t)ridx) << 16);
                // the larger range goes in first (the smaller one is split next): <= log2(m) + 1 entries
                const bool l_first = lsize >= rsize;
                int s2 = sp;
                if (lsize > 1 && (l_first || rsize <= 1)) { S.stack[s2][0] = b; S.stack[s2][1] = mid; S.stack[s2][2] = lidx; ++s2; }
                if (rsize > 1) { S.stack[s2][0] = mid; S.stack[s2][1] = e; S.stack[s2][2] = ridx; ++s2; }
                if (lsize > 1 && !(l_first || rsize <= 1)) { S.stack[s2][0] = b; S.stack[s2][1] = mid; S.stack[s2][2] = lidx; }
            }
            sp += (lsize > 1 ? 1 : 0) + (rsize > 1 ? 1 : 0);
            __syncwarp();
        }
        // new order of the subtree's triangles
        for (int k = lane; k < m; k += 32) vals[first + k] = S.tri[S.order[k]];
        if (FUSED) {
            // leaves of the subtree (what refit_kernel writes for a leaf), marked as covered
            for (int k = lane; k < m; k += 32) {
                const int node = (n - 1) + first + k;
                float3 lo, hi;
                tri_box(verts, S.tri[S.order[k]], lo, hi);
                const float sa = box_area(lo, hi);
                __stcg(O.bmin + node, make_float4(lo.x, lo.y, lo.z, P.ct * sa));
                __stcg(O.bmax + node, make_float4(hi.x, hi.y, hi.z, sa));
                __stcg(O.tcount + node, 1u);
                __stcg(O.icount + node, 0u);
                O.collapsed[node] = 2;
            }
            // internal nodes bottom-up: node i (names index) is ready when both children are done; every round each
            // lane takes one ready node of its own (i = lane + 32 j); done[j] is the ballot of the finished ones
            // (warp-uniform registers).  Children's results come back from global memory (.cg), ordered by __syncwarp.
            unsigned done[kTreeletWords];
#pragma unroll
            for (int j = 0; j < kTreeletWords; ++j) done[j] = 0u;
            const int n_int = m - 1;
            int n_done = 0;
            for (int round = 0; n_done < n_int && round < 2 * kTreelet; ++round) {  // (>= 1 node per round; the cap only guards against a corrupt link)
                __syncwarp();
                int pick = -1;
                uint32_t plink = 0;
#pragma unroll
                for (int j = 0; j < kTreeletWords; ++j) {
                    const int i = lane + 32 * j;
                    if (pick < 0 && i < n_int && !((done[j] >> lane) & 1u)) {
                        const uint32_t lk = S.link[i];
                        const uint32_t l = lk & 0xffffu, r = lk >> 16;
                        unsigned wl = 0u, wr = 0u;
#pragma unroll
                        for (int jj = 0; jj < kTreeletWords; ++jj) {
                            if ((int)((l & 0x7fffu) >> 5) == jj) wl = done[jj];
                            if ((int)((r & 0x7fffu) >> 5) == jj) wr = done[jj];
                        }
                        const bool lr = (l & 0x8000u) || ((wl >> (l & 31u)) & 1u);
                        const bool rr = (r & 0x8000u) || ((wr >> (r & 31u)) & 1u);
                        if (lr && rr) { pick = j; plink = lk; }
                    }
                }
                if (pick >= 0) {
                    const int x = S.names[lane + 32 * pick];
                    float4 clo[2], chi[2];
                    uint32_t ctc[2], cic[2];
#pragma unroll
                    for (int s2 = 0; s2 < 2; ++s2) {
                        const uint32_t cref = s2 ? plink >> 16 : plink & 0xffffu;
                        const bool is_leaf = (cref & 0x8000u) != 0u;
                        const int cn = is_leaf ? (n - 1) + first + (int)(cref & 0x7fffu) : S.names[cref];
                        clo[s2] = __ldcg(O.bmin + cn); chi[s2] = __ldcg(O.bmax + cn);  // (leaves: written above)
                        ctc[s2] = is_leaf ? 1u : __ldcg(O.tcount + cn);
                        cic[s2] = is_leaf ? 0u : __ldcg(O.icount + cn);
                    }
                    // (update_node, with the children in registers)
                    const float3 lo = make_float3(fminf(clo[0].x, clo[1].x), fminf(clo[0].y, clo[1].y), fminf(clo[0].z, clo[1].z));
                    const float3 hi = make_float3(fmaxf(chi[0].x, chi[1].x), fmaxf(chi[0].y, chi[1].y), fmaxf(chi[0].z, chi[1].z));
                    const float sa = box_area(lo, hi);
                    const uint32_t tc = ctc[0] + ctc[1];
                    const float c_int = P.cn * sa + clo[0].w + clo[1].w;
                    const float c_leaf = P.ct * (float)tc * sa;
                    const bool col = (x != 0) && tc <= P.max_leaf && c_leaf <= c_int;
                    __stcg(O.bmin + x, make_float4(lo.x, lo.y, lo.z, col ? c_leaf : c_int));
                    __stcg(O.bmax + x, make_float4(hi.x, hi.y, hi.z, sa));
                    __stcg(O.tcount + x, tc);
                    __stcg(O.icount + x, col ? 0u : 1u + cic[0] + cic[1]);
                    O.collapsed[x] = (uint8_t)(col ? 1 : 0);
                }
                __syncwarp();
#pragma unroll
                for (int j = 0; j < kTreeletWords; ++j) {
                    const unsigned nb = __ballot_sync(FULL, pick == j);
                    done[j] |= nb;
                    n_done += __popc(nb);
                }
            }
        }
        __syncwarp();
    }
}

__device__ __forceinline__ float union_area(float4 alo, float4 ahi, float4 blo, float4 bhi) {
    float3 lo = make_float3(fminf(alo.x, blo.x), fminf(alo.y, blo.y), fminf(alo.z, blo.z));
    float3 hi = make_float3(fmaxf(ahi.x, bhi.x), fmaxf(ahi.y, bhi.y), fmaxf(ahi.z, bhi.z));
    return box_area(lo, hi);
}

// Everything refit_kernel reads may have been written by another SM earlier in
// the same launch (ordered by the arrival flags), so all loads bypass L1 (.cg).
__device__ __forceinline__ void update_node(int x, const int* left, const int* right, float4* bmin,
                                            float4* bmax, uint32_t* tcount, uint32_t* icount,
                                            uint8_t* collapsed, const RefitParams& P) {
    int l = __ldcg(left + x), r = __ldcg(right + x);
    float4 llo = __ldcg(bmin + l), lhi = __ldcg(bmax + l), rlo = __ldcg(bmin + r), rhi = __ldcg(bmax + r);
    float3 lo = make_float3(fminf(llo.x, rlo.x), fminf(llo.y, rlo.y), fminf(llo.z, rlo.z));
    float3 hi = make_float3(fmaxf(lhi.x, rhi.x), fmaxf(lhi.y, rhi.y), fmaxf(lhi.z, rhi.z));
    float sa = box_area(lo, hi);
    uint32_t tc = __ldcg(tcount + l) + __ldcg(tcount + r);
    float c_int = P.cn * sa + llo.w + rlo.w;
    float c_leaf = P.ct * (float)tc * sa;
    bool col = (x != 0) && tc <= P.max_leaf && c_leaf <= c_int;
    __stcg(bmin + x, make_float4(lo.x, lo.y, lo.z, col ? c_leaf : c_int));
    __stcg(bmax + x, make_float4(hi.x, hi.y, hi.z, sa));
    __stcg(tcount + x, tc);
    __stcg(icount + x, col ? 0u : 1u + __ldcg(icount + l) + __ldcg(icount + r));
    __stcg(collapsed + x, (uint8_t)(col ? 1 : 0));
}

// Threads [0, n): one per leaf; a leaf covered by a fused treelet (collapsed == 2) has nothing to do.
// Threads [n, n + *n_roots) (roots != nullptr): one per fused treelet root, whose subtree is already refitted.
__global__ void refit_kernel(const float4* __restrict__ verts, const uint32_t* __restrict__ vals,
                             int* left, int* right, int* parent, float4* bmin, float4* bmax,
                             uint32_t* tcount, uint32_t* icount, uint8_t* collapsed,
                             unsigned int* flags, RefitParams P, const int* __restrict__ roots,
                             const unsigned int* __restrict__ n_roots) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    int n = P.n;
    int node;
    if (j < n) {
        node = (n - 1) + j;
        if (roots && collapsed[node] == 2) return;
        float3 lo, hi;
        tri_box(verts, vals[j], lo, hi);
        float sa = box_area(lo, hi);
        __stcg(bmin + node, make_float4(lo.x, lo.y, lo.z, P.ct * sa));
        __stcg(bmax + node, make_float4(hi.x, hi.y, hi.z, sa));
        __stcg(tcount + node, 1u);
        __stcg(icount + node, 0u);
        if (n == 1) return;
    } else {
        if (!roots || (unsigned)(j - n) >= *n_roots) return;
        node = roots[j - n];
        if (node == 0) return;  // the whole tree is one treelet
    }
    int p = __ldcg(parent + node);
    while (true) {
        __threadfence();
        unsigned int old = atomicAdd(flags + p, 1u);
        if (old == 0) return;  // the sibling subtree is not finished; its thread continues
        __threadfence();
        // this thread now owns the whole subtree of p
        if (P.rotations) {
            int l = __ldcg(left + p), r = __ldcg(right + p);
            float4 llo = __ldcg(bmin + l), lhi = __ldcg(bmax + l), rlo = __ldcg(bmin + r), rhi = __ldcg(bmax + r);
            float best = 0.0f;
            int which = -1;
            bool r_int = r < n - 1 && !__ldcg(collapsed + r);
            bool l_int = l < n - 1 && !__ldcg(collapsed + l);
            int rl = -1, rr = -1, ll = -1, lr = -1;
            if (r_int) {
                rl = __ldcg(left + r); rr = __ldcg(right + r);
                float g0 = rhi.w - union_area(llo, lhi, __ldcg(bmin + rr), __ldcg(bmax + rr));  // swap l <-> rl
                float g1 = rhi.w - union_area(llo, lhi, __ldcg(bmin + rl), __ldcg(bmax + rl));  // swap l <-> rr
                if (g0 > best) { best = g0; which = 0; }
                if (g1 > best) { best = g1; which = 1; }
            }
            if (l_int) {
                ll = __ldcg(left + l); lr = __ldcg(right + l);
                float g2 = lhi.w - union_area(rlo, rhi, __ldcg(bmin + lr), __ldcg(bmax + lr));  // swap r <-> ll
                float g3 = lhi.w - union_area(rlo, rhi, __ldcg(bmin + ll), __ldcg(bmax + ll));  // swap r <-> lr
                if (g2 > best) { best = g2; which = 2; }
                if (g3 > best) { best = g3; which = 3; }
            }
            if (which == 0) { __stcg(left + r, l); __stcg(left + p, rl); __stcg(parent + l, r); __stcg(parent + rl, p); }
            else if (which == 1) { __stcg(right + r, l); __stcg(left + p, rr); __stcg(parent + l, r); __stcg(parent + rr, p); }
            else if (which == 2) { __stcg(left + l, r); __stcg(right + p, ll); __stcg(parent + r, l); __stcg(parent + ll, p); }
            else if (which == 3) { __stcg(right + l, r); __stcg(right + p, lr); __stcg(parent + r, l); __stcg(parent + lr, p); }
            if (which == 0 || which == 1) update_node(r, left, right, bmin, bmax, tcount, icount, collapsed, P);
            if (which == 2 || which == 3) update_node(l, left, right, bmin, bmax, tcount, icount, collapsed, P);
        }
        update_node(p, left, right, bmin, bmax, tcount, icount, collapsed, P);
        if (p == 0) return;
        p = __ldcg(parent + p);
    }
}

__device__ __forceinline__ uint32_t quant_axis_exp(float extent) {
    if (!(extent > 0.0f)) return 1u;
    int k;
    frexpf(extent * (1.0f / 254.0f), &k);  // extent/254 = m * 2^k, m in [0.5,1)  =>  2^k >= extent/254
    int e = k + 127;
    return (uint32_t)min(max(e, 1), 254);
}

__device__ __forceinline__ bool quant_planes(float o, float scale, float lo, float hi, uint32_t& qlo,
                                             uint32_t& qhi) {
    float inv = 1.0f / scale;  // exact: power of two
    int a = (int)floorf((lo - o) * inv), b = (int)ceilf((hi - o) * inv);
    a = min(max(a, 0), 255);
    b = min(max(b, 0), 255);
    while (a > 0 && fmaf((float)a, scale, o) > lo) --a;
    while (b < 255 && fmaf((float)b, scale, o) < hi) ++b;
    qlo = (uint32_t)a;
    qhi = (uint32_t)b;
    return fmaf((float)a, scale, o) <= lo && fmaf((float)b, scale, o) >= hi;
}

// leaf-order triangle record (common.cuh): 32 B (p0.xyz, p1.xyz, p2.xy) + 8 B (p2.z, global id)
__device__ __forceinline__ void write_tri(const float4* __restrict__ verts, uint32_t t, uint32_t k, uint4* out_a,
                                          float2* out_b) {
    const float4 a = __ldg(verts + 3ull * t), b = __ldg(verts + 3ull * t + 1), c = __ldg(verts + 3ull * t + 2);
    out_a[2ull * k] = make_uint4(__float_as_uint(a.x), __float_as_uint(a.y), __float_as_uint(a.z), __float_as_uint(b.x));
    out_a[2ull * k + 1] = make_uint4(__float_as_uint(b.y), __float_as_uint(b.z), __float_as_uint(c.x), __float_as_uint(c.y));
    out_b[k] = make_float2(c.z, a.w);
}

__device__ __forceinline__ void write_leaf_tris(const float4* __restrict__ verts,
                                                const uint32_t* __restrict__ vals, const int* left,
                                                const int* right, int n, int x, uint32_t start,
                                                uint4* out_a, float2* out_b) {
    int stack[16];
    int sp = 0;
    uint32_t k = start;
    int cur = x;
    while (true) {
        if (cur >= n - 1) {
            uint32_t t = __ldg(vals + (cur - (n - 1)));
            write_tri(verts, t, k, out_a, out_b);
            ++k;
            if (sp == 0) break;
            cur = stack[--sp];
        } else {
            if (sp < 16) stack[sp++] = __ldg(right + cur);
            cur = __ldg(left + cur);
        }
    }
}

// ---- emit: collapse the binary tree into 4-wide records, breadth first ---------------------
// queue[i] = binary node that becomes wide record i.  One launch per level: a thread takes one
// record, widens {left, right} by (at most twice) replacing the internal child of largest
// surface area with its two children, quantises the child boxes in the record's frame, appends
// internal children to the queue (their queue position IS their record index) and copies the
// triangles of leaf children to a freshly reserved contiguous range.
struct EmitArgs {
    const float4* verts;
    const uint32_t* vals;
    const int* left;
    const int* right;
    const float4* bmin;
    const float4* bmax;
    const uint32_t* tcount;
    const uint8_t* collapsed;
    int n;
    int* queue;
    unsigned int* queue_tail;
    unsigned int* tri_tail;
    Node64* nodes;
    uint4* tri_a;
    float2* tri_b;
};

__device__ __forceinline__ void write_record(Node64* out, const float o[3], const float ext[3],
                                             const float clo[4][3], const float chi[4][3],
                                             const uint32_t ref[4], int nc) {
    uint32_t e[3], qlo[3] = {0, 0, 0}, qhi[3] = {0, 0, 0};
    for (int a = 0; a < 3; ++a) {
        e[a] = quant_axis_exp(ext[a]);
        while (true) {
            const float scale = __uint_as_float(e[a] << 23);
            bool ok = true;
            uint32_t wl = 0, wh = 0;
            for (int c = 0; c < 4; ++c) {
                uint32_t l = 255u, h = 0u;  // absent child: inverted planes (and ref == kNoChild)
                if (c < nc) ok = quant_planes(o[a], scale, clo[c][a], chi[c][a], l, h) && ok;
                wl |= l << (8 * c);
                wh |= h << (8 * c);
            }
            qlo[a] = wl; qhi[a] = wh;
            if (ok || e[a] >= 254) break;
            ++e[a];
        }
    }
    uint4* p = reinterpret_cast<uint4*>(out);
    p[0] = make_uint4(__float_as_uint(o[0]), __float_as_uint(o[1]), __float_as_uint(o[2]), e[0] << 23);
    p[1] = make_uint4(e[1] << 23, e[2] << 23, qlo[0], qhi[0]);
    p[2] = make_uint4(qlo[1], qhi[1], qlo[2], qhi[2]);
    p[3] = make_uint4(ref[0], ref[1], ref[2], ref[3]);
}

// Everything emit reads except the queue is constant during the launch: __ldg loads, issued in batches (all
// children at once), so that a record costs ~12 dependent memory round trips instead of ~30 (the stores to the
// node / triangle arrays may alias the inputs as far as the compiler knows, which serialised every load behind
// the store before it: emit was 0.38 ms of a 1.55 ms build at 1M triangles, 65 us per record per thread).
// Called by whole warps (`active`: this lane has a record): the queue and triangle ranges of the warp's records
// are reserved with ONE 64-bit atomic on the (queue tail, triangle tail) pair -- with one atomic pair per record
// the two counters were the limit of the launch (9.8 M same-address atomics in 2.3 ms at 10M triangles).
__device__ __forceinline__ void emit_record(const EmitArgs& A, unsigned int idx, bool active) {
    const int n = A.n;
    const int v = active ? A.queue[idx] : 0;
    int ch[4];
    int nc = 2;
    ch[0] = __ldg(A.left + v);
    ch[1] = __ldg(A.right + v);
    const float4 lo4 = __ldg(A.bmin + v), hi4 = __ldg(A.bmax + v);
    // per child: box, surface area (bmax.w), triangle count, "is a leaf of the wide tree"
    float4 cl[4], chh[4];
    uint32_t ctc[4];
    bool leaf[4];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        cl[k] = __ldg(A.bmin + ch[k]); chh[k] = __ldg(A.bmax + ch[k]);
        ctc[k] = __ldg(A.tcount + ch[k]);
        leaf[k] = ch[k] >= n - 1 || __ldg(A.collapsed + ch[k]) != 0;
    }
#pragma unroll
    for (int it = 0; it < 2; ++it) {
        int best = -1;
        float ba = -1.0f;
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (k < nc && !leaf[k] && chh[k].w > ba) { ba = chh[k].w; best = k; }
        if (best < 0) break;
        int c = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) if (k == best) c = ch[k];
        const int c0 = __ldg(A.left + c), c1 = __ldg(A.right + c);
        const float4 l0 = __ldg(A.bmin + c0), h0 = __ldg(A.bmax + c0), l1 = __ldg(A.bmin + c1), h1 = __ldg(A.bmax + c1);
        const uint32_t t0 = __ldg(A.tcount + c0), t1 = __ldg(A.tcount + c1);
        const bool f0 = c0 >= n - 1 || __ldg(A.collapsed + c0) != 0, f1 = c1 >= n - 1 || __ldg(A.collapsed + c1) != 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (k == best) { ch[k] = c0; cl[k] = l0; chh[k] = h0; ctc[k] = t0; leaf[k] = f0; }
            if (k == nc) { ch[k] = c1; cl[k] = l1; chh[k] = h1; ctc[k] = t1; leaf[k] = f1; }
        }
        ++nc;
    }
    const float o[3] = {lo4.x, lo4.y, lo4.z};
    const float ext[3] = {hi4.x - lo4.x, hi4.y - lo4.y, hi4.z - lo4.z};
    float clo[4][3], chi[4][3];
    uint32_t ref[4] = {kNoChild, kNoChild, kNoChild, kNoChild};
    // siblings get adjacent record slots and adjacent triangle ranges (one reservation each):
    // a ray that enters two children of this record finds the second one in the same lines
    uint32_t n_int = 0, n_tri = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (k < nc) {
            if (leaf[k]) n_tri += ctc[k];
            else ++n_int;
        }
    }
    // one-triangle leaves (the common case): triangle ids and vertices of all of them fetched together
    uint32_t tid1[4] = {0, 0, 0, 0};
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (k < nc && ch[k] >= n - 1) tid1[k] = __ldg(A.vals + (ch[k] - (n - 1)));
    float4 va[4], vb[4], vc[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (k < nc && ch[k] >= n - 1) {
            va[k] = __ldg(A.verts + 3ull * tid1[k]); vb[k] = __ldg(A.verts + 3ull * tid1[k] + 1); vc[k] = __ldg(A.verts + 3ull * tid1[k] + 2);
        }
    if (!active) { n_int = 0; n_tri = 0; nc = 0; }
    uint32_t pos, start;
    {
        const unsigned FULL = 0xffffffffu;
        const int lane = threadIdx.x & 31;
        const uint32_t mine = n_int | (n_tri << 16);  // <= 4 and <= 28 per record: the warp's sums fit 16 bits each
        uint32_t incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t up = __shfl_up_sync(FULL, incl, o);
            if (lane >= o) incl += up;
        }
        const uint32_t tot = __shfl_sync(FULL, incl, 31);
        unsigned long long base = 0ull;
        if (lane == 31 && tot)  // (queue_tail, tri_tail) are adjacent 32-bit counters, 8-byte aligned
            base = atomicAdd(reinterpret_cast<unsigned long long*>(A.queue_tail),
                             (unsigned long long)(tot & 0xffffu) | ((unsigned long long)(tot >> 16) << 32));
        base = __shfl_sync(FULL, base, 31);
        const uint32_t excl = incl - mine;
        pos = (uint32_t)base + (excl & 0xffffu);
        start = (uint32_t)(base >> 32) + (excl >> 16);
    }
    if (!active) return;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (k < nc) {
            const int c = ch[k];
            clo[k][0] = cl[k].x; clo[k][1] = cl[k].y; clo[k][2] = cl[k].z;
            chi[k][0] = chh[k].x; chi[k][1] = chh[k].y; chi[k][2] = chh[k].z;
            if (leaf[k]) {
                const uint32_t cnt = ctc[k];
                if (c >= n - 1) {
                    A.tri_a[2ull * start] = make_uint4(__float_as_uint(va[k].x), __float_as_uint(va[k].y), __float_as_uint(va[k].z), __float_as_uint(vb[k].x));
                    A.tri_a[2ull * start + 1] = make_uint4(__float_as_uint(vb[k].y), __float_as_uint(vb[k].z), __float_as_uint(vc[k].x), __float_as_uint(vc[k].y));
                    A.tri_b[start] = make_float2(vc[k].z, va[k].w);
                } else {
                    write_leaf_tris(A.verts, A.vals, A.left, A.right, n, c, start, A.tri_a, A.tri_b);
                }
                ref[k] = kLeafFlag | (start << 3) | cnt;
                start += cnt;
            } else {
                A.queue[pos] = c;
                ref[k] = pos++;
            }
        }
    }
    write_record(A.nodes + idx, o, ext, clo, chi, ref, nc);
}

// All levels in ONE cooperative launch (the grid is sized to be resident): a grid-wide barrier
// separates the levels, thread 0 publishes the next level's range in between.  Replaces one
// emit + one advance launch per level (38 launches, 0.2 ms of gaps at 1M triangles).
// level[0] = begin, level[1] = end of the current level in the queue, level[2] = depth so far.
#ifndef PRT_EMIT_MIN_BLOCKS
#define PRT_EMIT_MIN_BLOCKS 3
#endif
__global__ void __launch_bounds__(128, PRT_EMIT_MIN_BLOCKS)
emit_levels_kernel(EmitArgs A, unsigned int* level, int max_levels) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    const unsigned int tid = blockIdx.x * blockDim.x + threadIdx.x, nthreads = gridDim.x * blockDim.x;
    for (int lv = 0; lv < max_levels; ++lv) {
        const unsigned int begin = __ldcg(level), end = __ldcg(level + 1);
        if (begin >= end) break;  // (grid-uniform)
        for (unsigned int base = begin + (tid & ~31u); base < end; base += nthreads) {  // (warp-uniform trip count)
            const unsigned int idx = base + (tid & 31u);
            emit_record(A, idx, idx < end);
        }
        grid.sync();
        if (tid == 0) {
            const unsigned int tail = __ldcg(A.queue_tail);
            __stcg(level, end);
            __stcg(level + 1, tail);
            if (end < tail) __stcg(level + 2, __ldcg(level + 2) + 1u);
        }
        grid.sync();
    }
}

// n == 1: one record whose only child is the only triangle
__global__ void emit_single_kernel(const float4* __restrict__ verts, Node64* nodes, uint4* tri_a, float2* tri_b) {
    float3 lo, hi;
    tri_box(verts, 0, lo, hi);
    const float o[3] = {lo.x, lo.y, lo.z};
    const float ext[3] = {hi.x - lo.x, hi.y - lo.y, hi.z - lo.z};
    float clo[4][3], chi[4][3];
    // node_test4 does not test slots 0 and 1 for kNoChild: slot 1 is an EMPTY leaf (count 0) with the
    // same box, so a ray that "hits" it visits nothing
    for (int c = 0; c < 2; ++c) {
        clo[c][0] = lo.x; clo[c][1] = lo.y; clo[c][2] = lo.z;
        chi[c][0] = hi.x; chi[c][1] = hi.y; chi[c][2] = hi.z;
    }
    const uint32_t ref[4] = {kLeafFlag | 1u, kLeafFlag, kNoChild, kNoChild};
    write_record(nodes, o, ext, clo, chi, ref, 2);
    write_tri(verts, 0, 0, tri_a, tri_b);
}

}  // namespace

#define BUILD_TRY(expr)                                                                      \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess) {                                                             \
            ctx->set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return PRT_ERR_CUDA;                                                             \
        }                                                                                    \
    } while (0)

// grow-only device buffer of the context (never shrinks; prt_release_scratch / prt_destroy free it)
template <class T>
static cudaError_t reserve(T*& ptr, size_t& cap_bytes, size_t bytes) {
    if (cap_bytes >= bytes && ptr) return cudaSuccess;
    cudaFree(ptr);
    ptr = nullptr; cap_bytes = 0;
    const cudaError_t e = cudaMalloc(&ptr, bytes);
    if (e == cudaSuccess) cap_bytes = bytes;
    return e;
}

static int build_once(prt_ctx* ctx, const prt_bvh_options& opt, prt_bvh_stats* stats, bool* too_deep) {
    const auto wall0 = std::chrono::steady_clock::now();
    BuildBuffers B;
    const uint32_t nt = ctx->nt;
    const int n = (int)nt;
    *too_deep = false;
    ctx->n_nodes = 0;
    ctx->bvh_built = false;
    prt_bvh_stats st = {};
    st.n_tris = nt;
    st.max_leaf_tris = opt.max_leaf_tris;
    st.morton_bits = 30;
    if (nt == 0) {
        ctx->bvh_built = true;
        ctx->bvh_stats = st;
        if (stats) *stats = st;
        return PRT_OK;
    }
    // the build overwrites the context's node / triangle buffers in place: nothing launched earlier (on
    // any stream, including non-blocking ones the default stream does not wait for) may still traverse them
    BUILD_TRY(cudaDeviceSynchronize());
    cudaEvent_t* ev = ctx->build_ev;
    for (int k = 0; k < 7; ++k)
        if (!ev[k]) BUILD_TRY(cudaEventCreate(&ev[k]));
    const size_t nn = 2 * (size_t)nt - 1, ni = nt > 1 ? nt - 1 : 1;
    BUILD_TRY(reserve(ctx->tri_a, ctx->tri_a_bytes, sizeof(uint4) * 2 * (size_t)nt));
    BUILD_TRY(reserve(ctx->tri_b, ctx->tri_b_bytes, sizeof(float2) * (size_t)nt));
    // ---- carve the scratch out of the arena; everything that must start at zero comes first
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    const size_t o_flags = take(sizeof(unsigned int) * ni), o_coll = take(nn), o_small = take(256);
    const size_t zero_bytes = off;
    const size_t o_keys0 = take(4ull * nt), o_keys1 = take(4ull * nt), o_vals0 = take(4ull * nt), o_vals1 = take(4ull * nt);
    const size_t o_lo = take(4ull * nt), o_hi = take(4ull * nt), o_khi = take(4ull * nt);
    const size_t o_sort = take(radix_sort_temp_bytes(nt)), o_queue = take(4ull * nt);
    const size_t o_left = take(4ull * ni), o_right = take(4ull * ni), o_parent = take(4ull * nn);
    const size_t o_bmin = take(16ull * nn), o_bmax = take(16ull * nn), o_tc = take(4ull * nn), o_ic = take(4ull * nn);
    const size_t o_range = take(8ull * ni), o_roots = take(4ull * ni);
    BUILD_TRY(reserve(ctx->build_arena, ctx->build_arena_bytes, off));
    char* base = (char*)ctx->build_arena;
    B.flags = (unsigned int*)(base + o_flags); B.collapsed = (uint8_t*)(base + o_coll);
    unsigned int* small = (unsigned int*)(base + o_small);  // zeroed: [0..1] sort check, [2] n_roots, [8..12] emit tails, [16..21] scene box
    B.sort_check = small; B.n_roots = small + 2; B.tails = small + 8; B.scene_box = (int*)(small + 16);
    B.keys[0] = (uint32_t*)(base + o_keys0); B.keys[1] = (uint32_t*)(base + o_keys1);
    B.vals[0] = (uint32_t*)(base + o_vals0); B.vals[1] = (uint32_t*)(base + o_vals1);
    uint32_t* lo_by_tri = (uint32_t*)(base + o_lo); uint32_t* hi_by_tri = (uint32_t*)(base + o_hi);
    B.keys_hi = (uint32_t*)(base + o_khi);
    B.sort_tmp = base + o_sort; B.queue = (int*)(base + o_queue);
    B.left = (int*)(base + o_left); B.right = (int*)(base + o_right); B.parent = (int*)(base + o_parent);
    B.bmin = (float4*)(base + o_bmin); B.bmax = (float4*)(base + o_bmax);
    B.tcount = (uint32_t*)(base + o_tc); B.icount = (uint32_t*)(base + o_ic);
    B.range = (int2*)(base + o_range); B.roots = (int*)(base + o_roots);
    BUILD_TRY(cudaMemsetAsync(base, 0, zero_bytes));

    const int T = 256;
    const unsigned gN = (nt + T - 1) / T;
    cudaEventRecord(ev[0]);
    init_box_kernel<<<1, 32>>>(B.scene_box);
    bounds_kernel<<<min(gN, (unsigned)ctx->num_sms * 8u), T>>>(ctx->verts_gid, nt, B.scene_box);
    int sorted = 0;
    const uint32_t* keys_hi = nullptr;
    bool use63 = opt.morton_bits == 63;
    if (!use63) {
        morton_kernel<<<gN, T>>>(ctx->verts_gid, nt, B.scene_box, B.keys[0], B.vals[0]);
        cudaEventRecord(ev[1]);
        sorted = radix_sort_pairs(B.keys, B.vals, nt, 30, B.sort_tmp, 0);  // result in keys/vals[sorted]
        check_sorted_kernel<<<gN, T>>>(B.keys[sorted], nullptr, nt, B.sort_check);
        if (opt.morton_bits == 0 && nt >= 4096) {  // auto: did the 30-bit grid separate the triangles?
            unsigned int chk[2] = {0, 0};
            BUILD_TRY(cudaMemcpy(chk, B.sort_check, sizeof chk, cudaMemcpyDeviceToHost));
            use63 = chk[1] > nt / 16;
        }
    }
    if (use63) {
        BUILD_TRY(cudaMemsetAsync(B.sort_check, 0, 2 * sizeof(unsigned int)));
        morton63_kernel<<<gN, T>>>(ctx->verts_gid, nt, B.scene_box, B.keys[0], B.vals[0], lo_by_tri, hi_by_tri);
        if (opt.morton_bits == 63) cudaEventRecord(ev[1]);
        int cur = radix_sort_pairs(B.keys, B.vals, nt, 32, B.sort_tmp, 0);  // by the low word ...
        gather_kernel<<<gN, T>>>(hi_by_tri, B.vals[cur], nt, B.keys[cur]);
        uint32_t* k2[2] = {B.keys[cur], B.keys[cur ^ 1]};
        uint32_t* v2[2] = {B.vals[cur], B.vals[cur ^ 1]};
        const int c2 = radix_sort_pairs(k2, v2, nt, 32, B.sort_tmp, 0);     // ... then, stably, by the high word
        sorted = cur ^ c2;
        cudaMemcpyAsync(B.keys_hi, B.keys[sorted], 4ull * nt, cudaMemcpyDeviceToDevice);
        gather_kernel<<<gN, T>>>(lo_by_tri, B.vals[sorted], nt, B.keys[sorted]);
        keys_hi = B.keys_hi;
        check_sorted_kernel<<<gN, T>>>(B.keys[sorted], keys_hi, nt, B.sort_check);
        st.morton_bits = 63;
    }
    cudaEventRecord(ev[2]);
    if (n > 1) hierarchy_kernel<<<gN, T>>>(B.keys[sorted], keys_hi, n, B.left, B.right, B.parent, B.range);
    if (n > 2 && opt.treelets) {
        treelet_roots_kernel<<<gN, T>>>(B.range, B.parent, n, B.roots, B.n_roots);
    }
    RefitParams P;
    P.n = n; P.max_leaf = opt.max_leaf_tris; P.cn = opt.cost_node; P.ct = opt.cost_tri;
    P.rotations = (int)opt.rotations;
    // one refit pass: the treelet warps refit their own subtrees and refit_kernel starts at the subtree roots
    // (repeated rotation passes re-derive everything from the leaves, so they take the unfused kernel)
    const bool fused = n > 2 && opt.treelets && opt.rotations <= 1;
    if (n > 2 && opt.treelets) {
        const TreeletOut O{B.bmin, B.bmax, B.tcount, B.icount, B.collapsed};
        if (fused)
            treelet_sah_kernel<true><<<ctx->num_sms * PRT_TREELET_MIN_BLOCKS, 32 * kTreeletWarps>>>(ctx->verts_gid, B.vals[sorted], B.range, B.left, B.right,
                                                                             B.parent, n, B.roots, B.n_roots, O, P);
        else
            treelet_sah_kernel<false><<<ctx->num_sms * PRT_TREELET_MIN_BLOCKS, 32 * kTreeletWarps>>>(ctx->verts_gid, B.vals[sorted], B.range, B.left, B.right,
                                                                              B.parent, n, B.roots, B.n_roots, O, P);
    }
    cudaEventRecord(ev[3]);
    // rotations = k > 1 repeats the bottom-up pass: every pass re-derives boxes and costs from the
    // leaves and applies the best rotation per node again (the topology from the last pass is kept)
    for (int pass = 0; pass < (P.rotations > 1 ? P.rotations : 1); ++pass) {
        if (pass) cudaMemsetAsync(B.flags, 0, sizeof(unsigned int) * ni);
        // (a treelet root covers >= 3 triangles: at most n / 3 of them)
        const unsigned g_refit = fused ? (unsigned)(((size_t)nt + nt / 3 + 1 + T - 1) / T) : gN;
        refit_kernel<<<g_refit, T>>>(ctx->verts_gid, B.vals[sorted], B.left, B.right, B.parent, B.bmin, B.bmax,
                                     B.tcount, B.icount, B.collapsed, B.flags, P, fused ? B.roots : nullptr, B.n_roots);
    }
    cudaEventRecord(ev[4]);
    BUILD_TRY(cudaGetLastError());
    // the one mid-build read-back: number of surviving records (sizes the node array) + root box / cost
    uint32_t n_rec = 1;
    float4 root_lo = {}, root_hi = {};
    if (n > 1) {
        BUILD_TRY(cudaMemcpyAsync(&n_rec, B.icount, sizeof(uint32_t), cudaMemcpyDeviceToHost));
        BUILD_TRY(cudaMemcpyAsync(&root_lo, B.bmin, sizeof(float4), cudaMemcpyDeviceToHost));
        BUILD_TRY(cudaMemcpy(&root_hi, B.bmax, sizeof(float4), cudaMemcpyDeviceToHost));
    }
    BUILD_TRY(reserve(ctx->nodes, ctx->nodes_bytes, sizeof(Node64) * (size_t)(n_rec ? n_rec : 1)));
    unsigned int depth = 0, n_wide = 1;
    cudaEventRecord(ev[5]);
    unsigned int fin[5] = {1u, 0u, 0u, 0u, 1u};
    if (n > 1) {
        const unsigned int init[5] = {1u, 0u, /*level:*/ 0u, 1u, 1u};  // queue tail, tri tail, begin, end, depth
        BUILD_TRY(cudaMemsetAsync(B.queue, 0, sizeof(int)));  // record 0 = binary root (node 0)
        BUILD_TRY(cudaMemcpyAsync(B.tails, init, sizeof init, cudaMemcpyHostToDevice));
        EmitArgs A;
        A.verts = ctx->verts_gid; A.vals = B.vals[sorted]; A.left = B.left; A.right = B.right;
        A.bmin = B.bmin; A.bmax = B.bmax; A.tcount = B.tcount; A.collapsed = B.collapsed; A.n = n;
        A.queue = B.queue; A.queue_tail = B.tails; A.tri_tail = B.tails + 1; A.nodes = ctx->nodes;
        A.tri_a = ctx->tri_a; A.tri_b = ctx->tri_b;
        // the tree has at most kMaxStack/3 levels that traversal can use
        int max_levels = kMaxStack / 3 + 1;
        if (ctx->grid_emit == 0) {
            int b = 0;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, emit_levels_kernel, 128, 0);
            ctx->grid_emit = ctx->num_sms * (b > 0 ? b : 1);
        }
        unsigned int* level = B.tails + 2;
        void* args[] = {&A, &level, &max_levels};
        BUILD_TRY(cudaLaunchCooperativeKernel((void*)emit_levels_kernel, dim3(ctx->grid_emit), dim3(128), args, 0, 0));
        BUILD_TRY(cudaMemcpyAsync(fin, B.tails, sizeof fin, cudaMemcpyDeviceToHost));
    } else {
        emit_single_kernel<<<1, 1>>>(ctx->verts_gid, ctx->nodes, ctx->tri_a, ctx->tri_b);
    }
    cudaEventRecord(ev[6]);
    unsigned int chk[2] = {0, 0};
    BUILD_TRY(cudaMemcpyAsync(chk, B.sort_check, sizeof chk, cudaMemcpyDeviceToHost));
    BUILD_TRY(cudaDeviceSynchronize());
    if (n > 1) {
        n_wide = fin[0];
        depth = fin[4];
        if (fin[2] < fin[3]) depth = kMaxStack;  // levels left over: deeper than traversal supports
    } else {
        depth = 1;
    }
    n_rec = n_wide;
    float ms[6];
    for (int k = 0; k < 6; ++k) cudaEventElapsedTime(&ms[k], ev[k], ev[k + 1]);
    st.n_nodes = n_rec;
    st.sah_cost = (n > 1 && root_hi.w > 0.f) ? root_lo.w / root_hi.w : 0.f;
    if (n > 1) {
        ctx->scene_lo[0] = root_lo.x; ctx->scene_lo[1] = root_lo.y; ctx->scene_lo[2] = root_lo.z;
        ctx->scene_hi[0] = root_hi.x; ctx->scene_hi[1] = root_hi.y; ctx->scene_hi[2] = root_hi.z;
    }
    st.ms_morton = ms[0]; st.ms_sort = ms[1]; st.ms_hierarchy = ms[2]; st.ms_refit = ms[3];
    st.ms_emit = ms[5];
    st.ms_total = ms[0] + ms[1] + ms[2] + ms[3] + ms[4] + ms[5];
    st.depth = depth;
    st.morton_sorted = chk[0] == 0 ? 1u : 0u;
    st.ms_wall = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - wall0).count();
    if (3 * depth + 1 > (unsigned)kMaxStack) { *too_deep = true; return PRT_OK; }  // <= 3 pushes per level
    ctx->n_nodes = n_rec;
    ctx->bvh_built = true;
    ctx->bvh_stats = st;
    if (stats) *stats = st;
    return PRT_OK;
}

int build_bvh(prt_ctx* ctx, const prt_bvh_options* opts, prt_bvh_stats* stats) {
    prt_bvh_options o;
    // cost_tri 2: coplanar pairs (quads) still collapse into one leaf, random soups do not
    // (profiles/r1_sweeps.txt: soup-1M prefers 1-triangle leaves, Cornell 2..4)
    o.max_leaf_tris = 4; o.cost_node = 1.0f; o.cost_tri = 2.0f; o.rotations = 1; o.treelets = 1; o.morton_bits = 0;
    if (opts) o = *opts;
    if (o.morton_bits != 0 && o.morton_bits != 30 && o.morton_bits != 63) { ctx->set_error("bvh: morton_bits must be 0 (auto), 30 or 63"); return PRT_ERR_INVALID; }
    if (o.max_leaf_tris < 1 || o.max_leaf_tris > 7) { ctx->set_error("bvh: max_leaf_tris must be in 1..7"); return PRT_ERR_INVALID; }
    if (!ctx->scene_set) { ctx->set_error("bvh: no scene"); return PRT_ERR_STATE; }
    bool too_deep = false;
    int rc = build_once(ctx, o, stats, &too_deep);
    if (rc != PRT_OK) return rc;
    if (too_deep) {  // rotations / SAH treelets can deepen a degenerate tree; the Karras tree is bounded (<= 62)
        o.rotations = 0;
        o.treelets = 0;
        rc = build_once(ctx, o, stats, &too_deep);
        if (rc != PRT_OK) return rc;
        if (too_deep) { ctx->set_error("bvh: tree deeper than the traversal stack"); return PRT_ERR_STATE; }
    }
    return PRT_OK;
}

}  // namespace prt
