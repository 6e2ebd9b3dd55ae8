/* bvh_quality.c -- offline experiment (CPU, not product code): how many 4-wide record visits per
 * ray would the 1M-triangle soup need under different binary-tree builders?
 *   lbvh : 30-bit Morton, split at the highest differing bit   (what csrc/bvh_build.cu ships)
 *   sah  : top-down binned SAH, 16 bins x 3 axes               (quality ceiling)
 *   ploc : parallel locally-ordered clustering, radius R       (Meister & Bittner 2018)
 * All go through the same greedy 4-wide collapse (expand the child of largest area) and the same
 * ordered closest-hit traversal with exact FP32 boxes.
 * build: gcc -O2 -fopenmp -o bvh_quality bvh_quality.c -lm ; run: ./bvh_quality tris.bin rays.bin nrays */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float lo[3], hi[3]; } Box;
typedef struct { Box b; int left, right; int tri; } BNode; /* tri >= 0: leaf */

static int N;
static float (*T)[3][3];
static Box* TB;
static float (*C)[3];

static inline float area(const Box* b) {
    float x = b->hi[0] - b->lo[0], y = b->hi[1] - b->lo[1], z = b->hi[2] - b->lo[2];
    return 2.f * (x * y + y * z + z * x);
}
static inline void grow(Box* a, const Box* b) {
    for (int k = 0; k < 3; ++k) { if (b->lo[k] < a->lo[k]) a->lo[k] = b->lo[k]; if (b->hi[k] > a->hi[k]) a->hi[k] = b->hi[k]; }
}
static inline Box empty_box(void) { Box b = {{3e38f, 3e38f, 3e38f}, {-3e38f, -3e38f, -3e38f}}; return b; }

/* ---------------- binary tree storage */
static BNode* nodes; static int n_nodes;
static int new_leaf(int t) { nodes[n_nodes].b = TB[t]; nodes[n_nodes].tri = t; nodes[n_nodes].left = nodes[n_nodes].right = -1; return n_nodes++; }
static int new_inner(int l, int r) {
    BNode* n = &nodes[n_nodes]; n->b = nodes[l].b; grow(&n->b, &nodes[r].b); n->left = l; n->right = r; n->tri = -1; return n_nodes++;
}

/* ---------------- LBVH */
static uint32_t expand10(uint32_t v) { v = (v * 0x00010001u) & 0xFF0000FFu; v = (v * 0x00000101u) & 0x0F00F00Fu; v = (v * 0x00000011u) & 0xC30C30C3u; v = (v * 0x00000005u) & 0x49249249u; return v; }
static uint64_t* mk; /* key<<32 | tri */
static int cmp64(const void* a, const void* b) { uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b; return x < y ? -1 : x > y; }
static void morton_sort(void) {
    Box s = empty_box(); for (int i = 0; i < N; ++i) grow(&s, &TB[i]);
    mk = malloc(sizeof(uint64_t) * N);
    for (int i = 0; i < N; ++i) {
        uint32_t q[3];
        for (int k = 0; k < 3; ++k) { float c = (C[i][k] - s.lo[k]) / (s.hi[k] - s.lo[k]) * 1024.f; q[k] = (uint32_t)fminf(fmaxf(c, 0.f), 1023.f); }
        uint32_t key = (expand10(q[0]) << 2) | (expand10(q[1]) << 1) | expand10(q[2]);
        mk[i] = ((uint64_t)key << 32) | (uint32_t)i;
    }
    qsort(mk, N, sizeof(uint64_t), cmp64);
}
static int sah_rec(int a, int b);
static int* idx;
static int HYB = 0; /* hybrid: Morton ranges of <= HYB triangles are rebuilt by SAH */
static int lbvh_rec(int a, int b) { /* [a,b] inclusive */
    if (a == b) return new_leaf((int)(mk[a] & 0xffffffffu));
    if (b - a + 1 <= HYB) { for (int i = a; i <= b; ++i) idx[i] = (int)(mk[i] & 0xffffffffu); return sah_rec(a, b + 1); }
    uint32_t ka = mk[a] >> 32, kb = mk[b] >> 32; int split;
    if (ka == kb) split = (a + b) / 2;
    else { int p = __builtin_clz(ka ^ kb); int lo = a, hi = b; /* last index whose key shares > p bits with ka */
        while (lo < hi) { int m = (lo + hi + 1) / 2; uint32_t km = mk[m] >> 32; if (km == ka || __builtin_clz(ka ^ km) > p) lo = m; else hi = m - 1; }
        split = lo; }
    int l = lbvh_rec(a, split), r = lbvh_rec(split + 1, b);
    return new_inner(l, r);
}

/* ---------------- binned SAH */
static int sah_rec(int a, int b) { /* [a,b) */
    if (b - a == 1) return new_leaf(idx[a]);
    Box cb = empty_box();
    for (int i = a; i < b; ++i) for (int k = 0; k < 3; ++k) { float c = C[idx[i]][k]; if (c < cb.lo[k]) cb.lo[k] = c; if (c > cb.hi[k]) cb.hi[k] = c; }
    enum { NB = 16 };
    float best = 3e38f; int bax = -1, bsp = -1;
    for (int ax = 0; ax < 3; ++ax) {
        float ext = cb.hi[ax] - cb.lo[ax]; if (!(ext > 0.f)) continue;
        Box bb[NB]; int cnt[NB]; for (int j = 0; j < NB; ++j) { bb[j] = empty_box(); cnt[j] = 0; }
        for (int i = a; i < b; ++i) { int j = (int)((C[idx[i]][ax] - cb.lo[ax]) / ext * NB); if (j >= NB) j = NB - 1; grow(&bb[j], &TB[idx[i]]); cnt[j]++; }
        float ra[NB]; int rc[NB]; Box acc = empty_box(); int c = 0;
        for (int j = NB - 1; j > 0; --j) { grow(&acc, &bb[j]); c += cnt[j]; ra[j] = c ? area(&acc) : 0.f; rc[j] = c; }
        acc = empty_box(); c = 0;
        for (int j = 0; j < NB - 1; ++j) { grow(&acc, &bb[j]); c += cnt[j]; if (c == 0 || rc[j + 1] == 0) continue;
            float cost = area(&acc) * c + ra[j + 1] * rc[j + 1]; if (cost < best) { best = cost; bax = ax; bsp = j; } }
    }
    int mid;
    if (bax < 0) mid = (a + b) / 2;
    else { float ext = cb.hi[bax] - cb.lo[bax]; int i = a, j = b - 1;
        while (i <= j) { int q = (int)((C[idx[i]][bax] - cb.lo[bax]) / ext * NB); if (q >= NB) q = NB - 1;
            if (q <= bsp) ++i; else { int t = idx[i]; idx[i] = idx[j]; idx[j] = t; --j; } }
        mid = i; if (mid == a || mid == b) mid = (a + b) / 2; }
    int l = sah_rec(a, mid), r = sah_rec(mid, b);
    return new_inner(l, r);
}

/* ---------------- PLOC */
static int ploc_build(int R) {
    int* cl = malloc(sizeof(int) * N), *nn = malloc(sizeof(int) * N), *nx = malloc(sizeof(int) * N);
    int n = N;
    for (int i = 0; i < N; ++i) cl[i] = new_leaf((int)(mk[i] & 0xffffffffu));
    int iters = 0;
    while (n > 1) {
#pragma omp parallel for schedule(static)
        for (int i = 0; i < n; ++i) {
            float best = 3e38f; int bj = -1; int lo = i - R < 0 ? 0 : i - R, hi = i + R >= n ? n - 1 : i + R;
            for (int j = lo; j <= hi; ++j) if (j != i) { Box b = nodes[cl[i]].b; grow(&b, &nodes[cl[j]].b); float A = area(&b); if (A < best) { best = A; bj = j; } }
            nn[i] = bj;
        }
        int m = 0;
        for (int i = 0; i < n; ++i) {
            int j = nn[i];
            if (nn[j] == i) { if (i < j) nx[m++] = new_inner(cl[i], cl[j]); }
            else nx[m++] = cl[i];
        }
        int* t = cl; cl = nx; nx = t; n = m; ++iters;
    }
    fprintf(stderr, "ploc R=%d iterations %d\n", R, iters);
    int root = cl[0]; free(cl); free(nn); free(nx); return root;
}

/* ---------------- 4-wide collapse + trace */
typedef struct { Box cb[4]; int ref[4]; int nch; } WNode; /* ref >= 0: wide node, < 0: ~tri */
static WNode* wn; static int n_wide;
static int collapse(int root) { /* BFS */
    int* q = malloc(sizeof(int) * (N + 1)); int qh = 0, qt = 0; q[qt++] = root; n_wide = 0;
    /* first pass assigns indices in BFS order: wide node k <-> binary node q[k] */
    while (qh < qt) {
        int bn = q[qh]; WNode* w = &wn[qh]; ++qh;
        int ch[4]; int n = 2; ch[0] = nodes[bn].left; ch[1] = nodes[bn].right;
        while (n < 4) { int bi = -1; float ba = -1.f; for (int i = 0; i < n; ++i) if (nodes[ch[i]].tri < 0) { float A = area(&nodes[ch[i]].b); if (A > ba) { ba = A; bi = i; } }
            if (bi < 0) break; int c = ch[bi]; ch[bi] = nodes[c].left; ch[n++] = nodes[c].right; }
        w->nch = n;
        for (int i = 0; i < n; ++i) { w->cb[i] = nodes[ch[i]].b; if (nodes[ch[i]].tri >= 0) w->ref[i] = ~nodes[ch[i]].tri; else { w->ref[i] = qt; q[qt++] = ch[i]; } }
    }
    n_wide = qt; free(q); return 0;
}
/* SAH-optimal 4-wide collapse (dynamic programme of Ylitie, Karras, Laine 2017, sect. 3.1) */
static float (*Cst)[4]; static unsigned char (*Dec)[4]; /* Dec[n][k-1]: 0 = n stays one child; i>0 = dissolved, left gets i slots */
static float CT = 1.0f;
static void dp_rec(int n) {
    if (nodes[n].tri >= 0) { for (int k = 0; k < 4; ++k) { Cst[n][k] = area(&nodes[n].b) * CT; Dec[n][k] = 0; } return; }
    int l = nodes[n].left, r = nodes[n].right; dp_rec(l); dp_rec(r);
    float D[5]; unsigned char Di[5];
    for (int k = 2; k <= 4; ++k) { D[k] = 3e38f; Di[k] = 1; for (int i = 1; i < k; ++i) { float c = Cst[l][i - 1] + Cst[r][k - i - 1]; if (c < D[k]) { D[k] = c; Di[k] = (unsigned char)i; } } }
    Cst[n][0] = area(&nodes[n].b) + D[4]; Dec[n][0] = 0;
    for (int k = 2; k <= 4; ++k) { if (D[k] < Cst[n][0]) { Cst[n][k - 1] = D[k]; Dec[n][k - 1] = Di[k]; } else { Cst[n][k - 1] = Cst[n][0]; Dec[n][k - 1] = 0; } }
}
static int gather(int n, int k, int* out, int cnt) { /* children of a wide node: n may use k slots */
    if (nodes[n].tri >= 0 || Dec[n][k - 1] == 0) { out[cnt++] = n; return cnt; }
    int i = Dec[n][k - 1]; cnt = gather(nodes[n].left, i, out, cnt); return gather(nodes[n].right, k - i, out, cnt);
}
static void collapse_dp(int root) {
    Cst = malloc(sizeof(float) * 4 * n_nodes); Dec = malloc(4 * n_nodes); dp_rec(root);
    int* q = malloc(sizeof(int) * (N + 1)); int qh = 0, qt = 0; q[qt++] = root;
    while (qh < qt) {
        int bn = q[qh]; WNode* w = &wn[qh]; ++qh; int ch[4]; int n = 0;
        /* the node itself is dissolved with 4 slots: best split of D[4] */
        { int l = nodes[bn].left, r = nodes[bn].right; float best = 3e38f; int bi = 1; for (int i = 1; i < 4; ++i) { float c = Cst[l][i - 1] + Cst[r][4 - i - 1]; if (c < best) { best = c; bi = i; } }
          n = gather(l, bi, ch, n); n = gather(r, 4 - bi, ch, n); }
        w->nch = n;
        for (int i = 0; i < n; ++i) { w->cb[i] = nodes[ch[i]].b; if (nodes[ch[i]].tri >= 0) w->ref[i] = ~nodes[ch[i]].tri; else { w->ref[i] = qt; q[qt++] = ch[i]; } }
    }
    n_wide = qt; free(q); free(Cst); free(Dec);
}
static double sah_binary(int root) { double s = 0; float ra = area(&nodes[root].b); for (int i = 0; i < n_nodes; ++i) s += area(&nodes[i].b) / ra; return s; }
static double sah_wide(void) { double s = 1.0; float ra = 0; Box r = empty_box(); for (int i = 0; i < wn[0].nch; ++i) grow(&r, &wn[0].cb[i]); ra = area(&r);
    for (int k = 0; k < n_wide; ++k) for (int i = 0; i < wn[k].nch; ++i) if (wn[k].ref[i] >= 0) s += area(&wn[k].cb[i]) / ra; return s; }

static inline int tri_hit(const float* o, const float* d, int t, float tmax, float* tt) {
    const float* p0 = T[t][0]; float e1[3], e2[3], q[3], s[3], r[3];
    for (int k = 0; k < 3; ++k) { e1[k] = T[t][1][k] - p0[k]; e2[k] = T[t][2][k] - p0[k]; s[k] = o[k] - p0[k]; }
    q[0] = d[1] * e2[2] - d[2] * e2[1]; q[1] = d[2] * e2[0] - d[0] * e2[2]; q[2] = d[0] * e2[1] - d[1] * e2[0];
    float a = e1[0] * q[0] + e1[1] * q[1] + e1[2] * q[2]; if (fabsf(a) < 1e-12f) return 0; float f = 1.f / a;
    float u = f * (s[0] * q[0] + s[1] * q[1] + s[2] * q[2]); if (u < 0.f) return 0;
    r[0] = s[1] * e1[2] - s[2] * e1[1]; r[1] = s[2] * e1[0] - s[0] * e1[2]; r[2] = s[0] * e1[1] - s[1] * e1[0];
    float v = f * (d[0] * r[0] + d[1] * r[1] + d[2] * r[2]); if (v < 0.f || u + v > 1.f) return 0;
    float t_ = f * (e2[0] * r[0] + e2[1] * r[1] + e2[2] * r[2]); if (t_ < 1e-5f || t_ > tmax) return 0; *tt = t_; return 1;
}
static void trace(const float* rays, int nr, const char* name, double build_sah) {
    double vn = 0, vt = 0; long hits = 0;
#pragma omp parallel for reduction(+ : vn, vt, hits) schedule(dynamic, 256)
    for (int i = 0; i < nr; ++i) {
        const float* o = rays + 8 * i; const float* d = rays + 8 * i + 4; float id[3] = {1.f / d[0], 1.f / d[1], 1.f / d[2]};
        float best = 3.4e38f; int bt = -1; int st[256]; float stt[256]; int sp = 0; int cur = 0; float curt = 0;
        while (1) {
            if (cur >= 0) {
                vn += 1; WNode* w = &wn[cur]; float ct[4]; int cr[4]; int n = 0;
                for (int c = 0; c < w->nch; ++c) { float tn = 1e-5f, tf = best;
                    for (int k = 0; k < 3; ++k) { float a = (w->cb[c].lo[k] - o[k]) * id[k], b = (w->cb[c].hi[k] - o[k]) * id[k]; if (a > b) { float t = a; a = b; b = t; } if (a > tn) tn = a; if (b < tf) tf = b; }
                    if (tn <= tf) { int j = n++; while (j > 0 && ct[j - 1] < tn) { ct[j] = ct[j - 1]; cr[j] = cr[j - 1]; --j; } ct[j] = tn; cr[j] = w->ref[c]; } } /* descending: far first */
                for (int j = 0; j < n; ++j) { st[sp] = cr[j]; stt[sp] = ct[j]; ++sp; }
            } else { vt += 1; float t; if (tri_hit(o, d, ~cur, best, &t)) { best = t; bt = ~cur; } }
            int found = 0; while (sp > 0) { --sp; if (stt[sp] <= best) { cur = st[sp]; curt = stt[sp]; found = 1; break; } }
            if (!found) break;
        }
        (void)curt; hits += bt >= 0;
    }
    printf("%-10s binary SAH %.1f  wide nodes %d wide SAH %.1f  N_node %.2f  N_tri %.2f  hit %.3f\n", name, build_sah, n_wide, sah_wide(), vn / nr, vt / nr, (double)hits / nr);
    fflush(stdout);
}

int main(int argc, char** argv) {
    FILE* f = fopen(argv[1], "rb"); fseek(f, 0, SEEK_END); long sz = ftell(f); fseek(f, 0, SEEK_SET); N = (int)(sz / 36);
    T = malloc(sz); if (fread(T, 1, sz, f) != (size_t)sz) return 1; fclose(f);
    int nr = atoi(argv[3]); float* rays = malloc(32l * nr); f = fopen(argv[2], "rb"); if (fread(rays, 32, nr, f) != (size_t)nr) return 1; fclose(f);
    TB = malloc(sizeof(Box) * N); C = malloc(sizeof(float) * 3 * N);
    for (int i = 0; i < N; ++i) { TB[i] = empty_box(); for (int v = 0; v < 3; ++v) for (int k = 0; k < 3; ++k) { float x = T[i][v][k]; if (x < TB[i].lo[k]) TB[i].lo[k] = x; if (x > TB[i].hi[k]) TB[i].hi[k] = x; }
        for (int k = 0; k < 3; ++k) C[i][k] = 0.5f * (TB[i].lo[k] + TB[i].hi[k]); }
    nodes = malloc(sizeof(BNode) * 2 * N); wn = malloc(sizeof(WNode) * N);
    morton_sort();
    const char* which = argc > 4 ? argv[4] : "lbvh,sah,ploc8,ploc16";
    idx = malloc(sizeof(int) * N);
    if (strstr(which, "hyb")) for (HYB = 4; HYB <= 4096; HYB *= 4) { char nm[32]; sprintf(nm, "hybrid%d", HYB); n_nodes = 0; int root = lbvh_rec(0, N - 1); collapse(root); trace(rays, nr, nm, sah_binary(root)); }
    HYB = 0;
    if (strstr(which, "lbvh")) { n_nodes = 0; int root = lbvh_rec(0, N - 1); collapse(root); trace(rays, nr, "lbvh", sah_binary(root));
        for (CT = 0.5f; CT <= 4.0f; CT *= 2.f) { char nm[32]; sprintf(nm, "lbvh-dp%.1f", CT); collapse_dp(root); trace(rays, nr, nm, sah_binary(root)); } }
    if (strstr(which, "sah")) { n_nodes = 0; for (int i = 0; i < N; ++i) idx[i] = i; int root = sah_rec(0, N); collapse(root); trace(rays, nr, "sah", sah_binary(root));
        CT = 1.0f; collapse_dp(root); trace(rays, nr, "sah-dp1.0", sah_binary(root)); }
    for (int R = 4; R <= 64; R *= 2) { char nm[16]; sprintf(nm, "ploc%d", R); if (!strstr(which, nm)) continue; n_nodes = 0; int root = ploc_build(R); collapse(root); trace(rays, nr, nm, sah_binary(root)); }
    return 0;
}
