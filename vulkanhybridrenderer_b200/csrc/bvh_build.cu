// bvh_build.cu — GPU acceleration-structure build (sm_100a), in place of the driver's BLAS/TLAS build
// (/root/reference/src/rendering_backend/resource_manager.cpp:593-801 UpdateBLAS / UpdateTLAS).
//
//   1. extract   world-space triangle soup: Primitive.transform applied per geometry (resource_manager.cpp:608-617,
//                636-641), indices relative to vertex_offset, geometry index = flat primitive index
//   2. morton    63-bit Morton code of the triangle-box centre inside the scene box
//   3. sort      CUB radix sort of (code, triangle)
//   4. hierarchy binary tree over the sorted triangles: PLOC (parallel locally-ordered clustering, Meister & Bittner 2018; default:
//                3-12 % faster to trace) or the radix tree over the codes (Karras 2012; VHR_BVH_BUILDER=0)
//   5. refit     bottom-up AABBs + SAH cost; subtrees of <= 3 triangles collapse into leaves when SAH prefers it
//   5b. collapse which binary nodes become wide nodes / leaves: cost-optimal dynamic programme (default) or greedy (VHR_COLLAPSE=0)
//   6. widen     level-synchronous emission of the 8-wide nodes, slots sorted along the axis of largest spread,
//                child boxes quantised to 8 bits (conservative), leaf triangles rewritten contiguously per node
#include <stdlib.h>
#include <cub/cub.cuh>

#include <algorithm>
#include <vector>

#include "bvh.cuh"
#include "vhr_internal.h"

namespace vhr {

namespace {

__device__ __forceinline__ int float_to_ordered(float f) {
    int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__host__ __device__ __forceinline__ float ordered_to_float(int i) {
    int b = i >= 0 ? i : i ^ 0x7fffffff;
#ifdef __CUDA_ARCH__
    return __int_as_float(b);
#else
    float f;
    memcpy(&f, &b, 4);
    return f;
#endif
}

// ---- 1. extract ------------------------------------------------------------------------------------------------
__global__ void extract_triangles_kernel(const Vertex *__restrict__ verts, const uint32_t *__restrict__ indices,
                                         const Primitive *__restrict__ prims, const uint32_t *__restrict__ prefix,
                                         uint32_t n_prims, uint32_t n_tris, TriRef *__restrict__ out, int *scene_bounds) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    float mn[3] = {3.0e38f, 3.0e38f, 3.0e38f}, mx[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
    if (t < n_tris) {
        // largest g with prefix[g] <= t
        uint32_t lo = 0, hi = n_prims;
        while (hi - lo > 1) {
            uint32_t mid = (lo + hi) >> 1;
            if (prefix[mid] <= t) lo = mid; else hi = mid;
        }
        const uint32_t g = lo, k = t - prefix[g];
        const Primitive &p = prims[g];
        float3 v[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            uint32_t vi = p.vertex_offset + indices[p.index_offset + 3 * k + c];
            const Vertex &vx = verts[vi];
            v[c] = xform_point_rn(p.transform, vx.pos[0], vx.pos[1], vx.pos[2]);
        }
        TriRef r;
        r.v0 = make_float4(v[0].x, v[0].y, v[0].z, __uint_as_float(g));
        r.v1 = make_float4(v[1].x, v[1].y, v[1].z, __uint_as_float(k));
        // v2.w: 1 = a candidate hit on this triangle has to go through the any-hit stage (gbuf.frag's discards / shadow_anyhit.rahit):
        // an alpha-masked material, a base-colour texture index (its texel's alpha may be 0 or below the cutoff; the image may arrive
        // after this build) or a base colour with alpha 0. 0 = every any-hit stage of the path accepts it without looking the material up —
        // the closest-hit traversals test the flag in the register the vertex came in instead of chasing triangle -> primitive -> material
        // for every candidate.
        const bool any_hit_stage = p.material.alpha_mask == 1 || p.material.base_color_texture >= 0 || p.material.base_color[3] == 0.0f;
        r.v2 = make_float4(v[2].x, v[2].y, v[2].z, __uint_as_float(any_hit_stage ? 1u : 0u));
        out[t] = r;
        mn[0] = fminf(v[0].x, fminf(v[1].x, v[2].x)); mx[0] = fmaxf(v[0].x, fmaxf(v[1].x, v[2].x));
        mn[1] = fminf(v[0].y, fminf(v[1].y, v[2].y)); mx[1] = fmaxf(v[0].y, fmaxf(v[1].y, v[2].y));
        mn[2] = fminf(v[0].z, fminf(v[1].z, v[2].z)); mx[2] = fmaxf(v[0].z, fmaxf(v[1].z, v[2].z));
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        float lo = mn[a], hi = mx[a];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        }
        if ((threadIdx.x & 31) == 0 && lo <= hi) {
            atomicMin(&scene_bounds[a], float_to_ordered(lo));
            atomicMax(&scene_bounds[3 + a], float_to_ordered(hi));
        }
    }
}

// ---- 2. morton -------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t expand21(uint64_t x) {
    x &= 0x1fffffull;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}
// mode 1 (default): one scale for all axes, the largest extent — cubic grid cells; the short axes leave their top bits at zero,
// so the first splits of the radix tree only cut the long axes, which is what a surface-area heuristic does with an elongated
// scene box. mode 0 (VHR_MORTON_MODE=0, the first version): every axis normalised by its own extent, cells with the scene box's
// aspect ratio (40 x 10 x 16 for the hall). Measured at 1080p / 3 M triangles: SAH cost 35.4 -> 31.9, shadow rays 0.415 -> 0.334 ms,
// shadow + AO 0.881 -> 0.782 ms, shadow + 2 AO + reflection 2.41 -> 2.29 ms (the closest-hit rays alone lose 5 %).
__global__ void morton_kernel(const TriRef *__restrict__ tris, uint32_t n, const int *__restrict__ scene_bounds,
                              uint64_t *__restrict__ keys, uint32_t *__restrict__ vals, int mode) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    float smin[3], inv[3], ext[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        smin[a] = ordered_to_float(scene_bounds[a]);
        ext[a] = ordered_to_float(scene_bounds[3 + a]) - smin[a];
    }
    const float emax = fmaxf(ext[0], fmaxf(ext[1], ext[2]));
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float e = mode == 1 ? emax : ext[a];
        inv[a] = e > 0.0f ? 2097152.0f / e : 0.0f;
    }
    const TriRef r = tris[t];
    float c[3];
    c[0] = 0.5f * (fminf(r.v0.x, fminf(r.v1.x, r.v2.x)) + fmaxf(r.v0.x, fmaxf(r.v1.x, r.v2.x)));
    c[1] = 0.5f * (fminf(r.v0.y, fminf(r.v1.y, r.v2.y)) + fmaxf(r.v0.y, fmaxf(r.v1.y, r.v2.y)));
    c[2] = 0.5f * (fminf(r.v0.z, fminf(r.v1.z, r.v2.z)) + fmaxf(r.v0.z, fmaxf(r.v1.z, r.v2.z)));
    uint64_t q[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        float f = (c[a] - smin[a]) * inv[a];
        f = fminf(fmaxf(f, 0.0f), 2097151.0f);
        q[a] = (uint64_t)f;
    }
    keys[t] = (expand21(q[0]) << 2) | (expand21(q[1]) << 1) | expand21(q[2]);
    vals[t] = t;
}

// ---- 4. karras -------------------------------------------------------------------------------------------------
struct Tree2 {
    uint32_t n;              // number of leaves (triangles)
    uint32_t *child_l;       // [n-1] unified ids: < n-1 internal, >= n-1 leaf (id - (n-1) = sorted position)
    uint32_t *child_r;
    uint32_t *parent;        // [2n-1]
    uint32_t *range_first;   // [n-1] first sorted position covered
    float4 *bmin;            // [2n-1] xyz = box min, w = SAH cost of the subtree
    float4 *bmax;            // [2n-1] xyz = box max, w = triangle count (as uint bits)
    uint8_t *cluster;        // [n-1] 1: subtree is emitted as one leaf (<= kMaxLeafTris triangles)
    int *visit;              // [n-1]
    uint32_t *lcount;        // [2n-1] number of leaf clusters (future leaf slots) in the subtree
};

__device__ __forceinline__ int delta(const uint64_t *keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    uint64_t a = keys[i], b = keys[j];
    if (a == b) return 64 + __clz((uint32_t)i ^ (uint32_t)j);
    return __clzll((long long)(a ^ b));
}

__global__ void karras_kernel(const uint64_t *__restrict__ keys, Tree2 t) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int n = (int)t.n;
    if (i >= n - 1) return;
    int d = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    int dmin = delta(keys, n, i, i - d);
    int lmax = 2;
    while (delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int s = lmax >> 1; s >= 1; s >>= 1)
        if (delta(keys, n, i, i + (l + s) * d) > dmin) l += s;
    int j = i + l * d;
    int dnode = delta(keys, n, i, j);
    int s = 0;
    int tt = l;
    do {
        tt = (tt + 1) >> 1;
        if (delta(keys, n, i, i + (s + tt) * d) > dnode) s += tt;
    } while (tt > 1);
    int gamma = i + s * d + min(d, 0);
    int lo = min(i, j), hi = max(i, j);
    uint32_t left = (lo == gamma) ? (uint32_t)(n - 1 + gamma) : (uint32_t)gamma;
    uint32_t right = (hi == gamma + 1) ? (uint32_t)(n - 1 + gamma + 1) : (uint32_t)(gamma + 1);
    t.child_l[i] = left;
    t.child_r[i] = right;
    t.parent[left] = (uint32_t)i;
    t.parent[right] = (uint32_t)i;
    t.range_first[i] = (uint32_t)lo;
    if (i == 0) t.parent[0] = 0xffffffffu;
}

__device__ __forceinline__ float box_half_area(float3 mn, float3 mx) {
    float dx = mx.x - mn.x, dy = mx.y - mn.y, dz = mx.z - mn.z;
    return dx * dy + dy * dz + dz * dx;
}

// ---- 4b. PLOC ---------------------------------------------------------------------------------------------------
// Parallel locally-ordered clustering (Meister & Bittner 2018) as an alternative to the radix tree: bottom-up agglomerative
// clustering restricted to a window of +-R neighbours in Morton order. Every round each cluster finds the neighbour whose
// union with it has the smallest surface area; mutual nearest neighbours merge into a new internal node; the survivors are
// compacted. Same Tree2 conventions as the radix tree (internal ids < n-1, root = 0: ids are handed out from n-2 downwards,
// and the n-1-th merge is the root), so refit / SAH leaf decisions / widening run unchanged on either hierarchy.
constexpr int kPlocMaxRounds = 512;     // these scenes take 30-45

struct PlocBuf {
    uint32_t *id;      // unified node id of the cluster
    float4 *mn, *mx;   // its box
};

__global__ void ploc_init_kernel(const TriRef *__restrict__ tris, const uint32_t *__restrict__ order, uint32_t n, PlocBuf c) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const TriRef r = tris[order[j]];
    c.id[j] = n - 1 + j;
    c.mn[j] = make_float4(fminf(r.v0.x, fminf(r.v1.x, r.v2.x)), fminf(r.v0.y, fminf(r.v1.y, r.v2.y)), fminf(r.v0.z, fminf(r.v1.z, r.v2.z)), 0.0f);
    c.mx[j] = make_float4(fmaxf(r.v0.x, fmaxf(r.v1.x, r.v2.x)), fmaxf(r.v0.y, fmaxf(r.v1.y, r.v2.y)), fmaxf(r.v0.z, fmaxf(r.v1.z, r.v2.z)), 0.0f);
}

template <int R>
__global__ void __launch_bounds__(256) ploc_nn_kernel(PlocBuf c, uint32_t N, uint32_t *__restrict__ nn) {
    constexpr int B = 256;
    __shared__ float4 smn[B + 2 * R], smx[B + 2 * R];
    const int base = (int)(blockIdx.x * B) - R;
    for (int k = threadIdx.x; k < B + 2 * R; k += B) {
        const int g = base + k;
        if (g >= 0 && g < (int)N) { smn[k] = c.mn[g]; smx[k] = c.mx[g]; }
    }
    __syncthreads();
    const int i = (int)(blockIdx.x * B + threadIdx.x);
    if (i >= (int)N) return;
    const int li = (int)threadIdx.x + R;
    const float4 amn = smn[li], amx = smx[li];
    float best = __int_as_float(0x7f800000);
    int bj = -1;
#pragma unroll 4
    for (int d = -R; d <= R; ++d) {
        const int g = i + d;
        if (d == 0 || g < 0 || g >= (int)N) continue;
        const float4 bmn = smn[li + d], bmx = smx[li + d];
        const float ar = fminf(box_half_area(make_float3(fminf(amn.x, bmn.x), fminf(amn.y, bmn.y), fminf(amn.z, bmn.z)),
                                             make_float3(fmaxf(amx.x, bmx.x), fmaxf(amx.y, bmx.y), fmaxf(amx.z, bmx.z))), 3.0e38f);   // NaN / inf boxes tie at 3e38
        if (ar < best) { best = ar; bj = g; }       // ascending scan: ties keep the smaller index, so a closest pair is always mutual
    }
    nn[i] = (uint32_t)bj;
}

// A cluster disappears from the list when it is the larger index of a mutual pair. One extra element (valid[N] = 0) makes the
// exclusive scan's last entry the next round's cluster count.
__global__ void ploc_flag_kernel(uint32_t N, const uint32_t *__restrict__ nn, uint32_t *__restrict__ valid) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > N) return;
    if (i == N) { valid[i] = 0u; return; }
    const uint32_t j = nn[i];
    valid[i] = (nn[j] == i && i > j) ? 0u : 1u;
}

// Merge + compaction in one pass. The smaller index of a mutual pair carries the merged cluster. Node ids are handed out without
// atomics: round r has merged (n - N) pairs before it, and within the round a merge is numbered by the rank of its absorbed
// partner among the absorbed clusters (j - pos[j]), so the tree and its numbering are the same on every run and every GPU.
__global__ void ploc_merge_kernel(PlocBuf in, PlocBuf out, uint32_t N, const uint32_t *__restrict__ nn, const uint32_t *__restrict__ valid,
                                  const uint32_t *__restrict__ pos, Tree2 t) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N || !valid[i]) return;
    const uint32_t o = pos[i], j = nn[i];
    if (nn[j] != i) { out.id[o] = in.id[i]; out.mn[o] = in.mn[i]; out.mx[o] = in.mx[i]; return; }
    const uint32_t id = t.n - 2u - ((t.n - N) + (j - pos[j]));
    const uint32_t l = in.id[i], r = in.id[j];
    t.child_l[id] = l; t.child_r[id] = r;
    t.parent[l] = id; t.parent[r] = id;
    const float4 amn = in.mn[i], amx = in.mx[i], bmn = in.mn[j], bmx = in.mx[j];
    out.id[o] = id;
    out.mn[o] = make_float4(fminf(amn.x, bmn.x), fminf(amn.y, bmn.y), fminf(amn.z, bmn.z), 0.0f);
    out.mx[o] = make_float4(fmaxf(amx.x, bmx.x), fmaxf(amx.y, bmx.y), fmaxf(amx.z, bmx.z), 0.0f);
}

// ---- 5. refit + SAH --------------------------------------------------------------------------------------------

__global__ void refit_kernel(const TriRef *__restrict__ tris, const uint32_t *__restrict__ order, Tree2 t, float tri_cost, uint32_t max_leaf) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= t.n) return;
    const TriRef r = tris[order[j]];
    float3 mn = make_float3(fminf(r.v0.x, fminf(r.v1.x, r.v2.x)), fminf(r.v0.y, fminf(r.v1.y, r.v2.y)), fminf(r.v0.z, fminf(r.v1.z, r.v2.z)));
    float3 mx = make_float3(fmaxf(r.v0.x, fmaxf(r.v1.x, r.v2.x)), fmaxf(r.v0.y, fmaxf(r.v1.y, r.v2.y)), fmaxf(r.v0.z, fmaxf(r.v1.z, r.v2.z)));
    uint32_t id = t.n - 1 + j;
    t.bmin[id] = make_float4(mn.x, mn.y, mn.z, box_half_area(mn, mx) * tri_cost);      // leaf cost = area * 1 triangle
    t.bmax[id] = make_float4(mx.x, mx.y, mx.z, __uint_as_float(1u));
    t.lcount[id] = 1u;
    if (t.n == 1) return;
    __threadfence();
    uint32_t p = t.parent[id];
    while (p != 0xffffffffu) {
        if (atomicAdd(&t.visit[p], 1) == 0) return;   // the sibling subtree is not finished: its thread continues
        __threadfence();
        uint32_t l = t.child_l[p], rr = t.child_r[p];
        float4 lmn = __ldcg(&t.bmin[l]), lmx = __ldcg(&t.bmax[l]);
        float4 rmn = __ldcg(&t.bmin[rr]), rmx = __ldcg(&t.bmax[rr]);
        mn = make_float3(fminf(lmn.x, rmn.x), fminf(lmn.y, rmn.y), fminf(lmn.z, rmn.z));
        mx = make_float3(fmaxf(lmx.x, rmx.x), fmaxf(lmx.y, rmx.y), fmaxf(lmx.z, rmx.z));
        uint32_t cnt = __float_as_uint(lmx.w) + __float_as_uint(rmx.w);
        float area = box_half_area(mn, mx);
        float cost_inner = area * 1.0f + lmn.w + rmn.w;       // node cost 1, triangle cost 1
        float cost_leaf = area * (float)cnt * tri_cost;
        bool cl = cnt <= max_leaf && cost_leaf <= cost_inner;
        t.cluster[p] = cl ? 1 : 0;
        t.lcount[p] = cl ? 1u : __ldcg(&t.lcount[l]) + __ldcg(&t.lcount[rr]);
        t.bmin[p] = make_float4(mn.x, mn.y, mn.z, cl ? cost_leaf : cost_inner);
        t.bmax[p] = make_float4(mx.x, mx.y, mx.z, __uint_as_float(cnt));
        __threadfence();
        p = t.parent[p];
    }
}

// ---- 5b. cost-optimal collapse (default; VHR_COLLAPSE=0 = greedy) ----------------------------------------------------------------------
// The greedy collapse (widen_kernel) leaves 30 % of the child slots empty (vhr_bvh_stats.n_used_slots). This is the dynamic programme of
// Ylitie, Karras & Laine 2017 (section 3.1) instead: C(x, i) = the smallest SAH cost of subtree x represented by at most i roots
// (children of one wide node), i = 1..7,
//     C(x, 1) = min( area * triangles * tri_cost   [x becomes a leaf, <= kMaxLeafTris triangles],
//                    area * node_cost + D(x, 8)    [x becomes a wide node whose children are the best 8 roots below it] )
//     C(x, i) = min( D(x, i), C(x, i-1) ),   D(x, j) = min over 0 < k < j of C(left, k) + C(right, j - k),
// filled bottom-up with the decisions kept (one byte per (node, i): the left share k, 0 = "same as i-1" / "leaf" for i = 1);
// the widening pass then reads the children of a wide node off the decisions. A single triangle costs area * tri_cost for every i.
struct DpTables {
    float *cost;      // [n-1][7]
    uint8_t *dec;     // [n-1][8] (entry 7 unused)
};

__device__ __forceinline__ float dp_cost_of(const Tree2 &t, const DpTables &d, uint32_t id, int i) {    // i = 1..7
    return id >= t.n - 1 ? __ldcg(&t.bmin[id]).w : __ldcg(&d.cost[(size_t)id * 7 + (i - 1)]);
}

__global__ void collapse_dp_kernel(Tree2 t, DpTables d, float node_cost, float tri_cost, uint32_t max_leaf) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= t.n || t.n == 1) return;
    uint32_t p = t.parent[t.n - 1 + j];
    while (p != 0xffffffffu) {
        if (atomicAdd(&t.visit[p], 1) == 0) return;   // the sibling subtree is not finished: its thread continues
        __threadfence();
        const uint32_t l = t.child_l[p], r = t.child_r[p];
        float cl[7], cr[7];
#pragma unroll
        for (int i = 0; i < 7; ++i) { cl[i] = dp_cost_of(t, d, l, i + 1); cr[i] = dp_cost_of(t, d, r, i + 1); }
        const float4 mn = t.bmin[p], mx = t.bmax[p];
        const float area = box_half_area(make_float3(mn.x, mn.y, mn.z), make_float3(mx.x, mx.y, mx.z));
        const uint32_t cnt = __float_as_uint(mx.w);
        float D[9];
        uint8_t K[9];
#pragma unroll
        for (int jj = 2; jj <= 8; ++jj) {
            float best = 3.0e38f;
            uint8_t bk = 1;
#pragma unroll
            for (int k = 1; k < jj; ++k) {
                if (k > 7 || jj - k > 7) continue;
                const float c = cl[k - 1] + cr[jj - k - 1];
                if (c < best) { best = c; bk = (uint8_t)k; }
            }
            D[jj] = best; K[jj] = bk;
        }
        const float c_int = area * node_cost + D[8];
        const float c_leaf = cnt <= max_leaf ? area * (float)cnt * tri_cost : 3.0e38f;
        const bool leaf = c_leaf <= c_int;
        float prev = leaf ? c_leaf : c_int;
        d.cost[(size_t)p * 7] = prev;
        d.dec[(size_t)p * 8] = leaf ? (uint8_t)0 : K[8];
        t.cluster[p] = leaf ? 1 : 0;
#pragma unroll
        for (int i = 2; i <= 7; ++i) {
            uint8_t dd = 0;
            if (D[i] < prev) { prev = D[i]; dd = K[i]; }
            d.cost[(size_t)p * 7 + (i - 1)] = prev;
            d.dec[(size_t)p * 8 + (i - 1)] = dd;
        }
        __threadfence();
        p = t.parent[p];
    }
}

// ---- 6. widen --------------------------------------------------------------------------------------------------
struct WidenArgs {
    Tree2 t;
    const TriRef *tris;          // unsorted triangles
    const uint32_t *order;       // sorted position -> unsorted triangle
    WideNode *wide;
    TriRef *tris_out;            // final triangle array (leaf order)
    const uint2 *queue_in;       // (binary node, wide node index)
    uint2 *queue_out;
    uint32_t n_in;
    uint32_t *counters;          // [0] wide nodes allocated, [1] triangles emitted, [2] queue_out size, [3] child slots in use
    float *sah;                  // accumulated SAH numerator
    int child_sort;              // 0: slots in collapse order; 1: largest surface area first; 2: ascending along the axis of largest spread
    const uint8_t *dp_dec;       // cost-optimal collapse: decisions of collapse_dp_kernel ([n-1][8]); nullptr = greedy collapse
};

__device__ __forceinline__ bool leaf_like(const Tree2 &t, uint32_t id) { return id >= t.n - 1 || t.cluster[id]; }

__device__ void emit_wide_node(const WidenArgs &a, uint32_t wide_idx, const uint32_t *kids, int nk, float3 nmn, float3 nmx, uint32_t order_axis = 3u) {
    const Tree2 &t = a.t;
    int n_inner = 0, n_leaf_tris = 0;
    for (int c = 0; c < nk; ++c) {
        if (leaf_like(t, kids[c])) n_leaf_tris += (int)__float_as_uint(t.bmax[kids[c]].w);
        else n_inner++;
    }
    uint32_t child_base = n_inner ? atomicAdd(&a.counters[0], (uint32_t)n_inner) : 0u;
    uint32_t tri_base = n_leaf_tris ? atomicAdd(&a.counters[1], (uint32_t)n_leaf_tris) : 0u;
    uint32_t q_base = n_inner ? atomicAdd(&a.counters[2], (uint32_t)n_inner) : 0u;

    WideNode w;
    w.origin[0] = nmn.x; w.origin[1] = nmn.y; w.origin[2] = nmn.z;
    float ext[3] = {nmx.x - nmn.x, nmx.y - nmn.y, nmx.z - nmn.z};
    float scale[3];
    for (int ax = 0; ax < 3; ++ax) {
        // smallest power of two with ext / scale <= 254 (one step of slack for the conservative fix-ups below)
        int e = 1;
        if (ext[ax] > 0.0f) {
            int ex;
            float m = frexpf(ext[ax] / 254.0f, &ex);   // ext/254 = m * 2^ex, m in [0.5,1)
            (void)m;
            e = ex + 127;                              // 2^ex >= ext/254
            e = max(1, min(254, e));
        }
        w.e[ax] = (uint8_t)e;
        scale[ax] = __uint_as_float((uint32_t)e << 23);
    }
    w.imask = 0;
    w.child_base = child_base | (order_axis << 30);     // bits 30-31: axis the slots are sorted along (3 = not sorted), see bvh.cuh
    w.tri_base = tri_base;
    for (int s = 0; s < 8; ++s) {
        w.meta[s] = 0;
        for (int ax = 0; ax < 3; ++ax) { w.qlo[ax][s] = 255; w.qhi[ax][s] = 0; }
    }
    int k_inner = 0, tri_off = 0;
    float sah = box_half_area(nmn, nmx);
    for (int c = 0; c < nk; ++c) {
        uint32_t id = kids[c];
        float4 cmn = t.bmin[id], cmx = t.bmax[id];
        float lo[3] = {cmn.x, cmn.y, cmn.z}, hi[3] = {cmx.x, cmx.y, cmx.z};
        for (int ax = 0; ax < 3; ++ax) {
            float org = w.origin[ax];
            float ql = floorf((lo[ax] - org) / scale[ax]);
            float qh = ceilf((hi[ax] - org) / scale[ax]);
            ql = fminf(fmaxf(ql, 0.0f), 255.0f);
            qh = fminf(fmaxf(qh, 0.0f), 255.0f);
            // conservative fix-ups against rounding in the two lines above
            while (ql > 0.0f && __fmaf_rn(ql, scale[ax], org) > lo[ax]) ql -= 1.0f;
            while (qh < 255.0f && __fmaf_rn(qh, scale[ax], org) < hi[ax]) qh += 1.0f;
            w.qlo[ax][c] = (uint8_t)ql;
            w.qhi[ax][c] = (uint8_t)qh;
        }
        if (leaf_like(t, id)) {
            uint32_t cnt = __float_as_uint(cmx.w);
            w.meta[c] = (uint8_t)((cnt << 5) | (uint32_t)tri_off);
            // the triangles of a collapsed subtree (<= kMaxLeafTris leaves), left to right: sorted order for the radix tree; a PLOC
            // subtree is not a contiguous range of the sorted array, so the subtree is walked
            uint32_t st[4];
            int sp = 0;
            uint32_t k = 0;
            st[sp++] = id;
            while (sp) {
                const uint32_t x = st[--sp];
                if (x >= t.n - 1) {
                    if (k < cnt) a.tris_out[tri_base + tri_off + k] = a.tris[a.order[x - (t.n - 1)]];
                    ++k;
                } else if (sp + 2 <= 4) {
                    st[sp++] = t.child_r[x];
                    st[sp++] = t.child_l[x];
                }
            }
            tri_off += (int)cnt;
            sah += box_half_area(make_float3(cmn.x, cmn.y, cmn.z), make_float3(cmx.x, cmx.y, cmx.z)) * (float)cnt;
        } else {
            w.imask |= (uint8_t)(1u << c);
            w.meta[c] = (uint8_t)(0x80u | (uint32_t)k_inner);
            a.queue_out[q_base + k_inner] = make_uint2(id, child_base + k_inner);
            k_inner++;
        }
    }
    a.wide[wide_idx] = w;
    atomicAdd(a.sah, sah);
    atomicAdd(&a.counters[3], (uint32_t)nk);
}

// Slot order, internal-first partition and emission of one wide node whose children have been chosen.
__device__ void finish_wide_node(const WidenArgs &a, const uint2 item, uint32_t *kids, const int nk) {
    const Tree2 &t = a.t;
    // Slot order = visiting order (the traversal pops the lowest set bit first). a.child_sort = 1 (VHR_CHILD_SORT=1, an experiment):
    // largest surface area first within the internal children and within the leaves, on the idea that an any-hit ray stops at its
    // first occluder and the largest box is the most likely to hold one. Measured: shadow + AO 0.782 -> 0.800 ms, primary rays
    // 0.70 -> 0.78 ms, reflections 1.25 -> 1.17 ms — not used.
    if (a.child_sort == 1) {
        float area[8];
        for (int c = 0; c < nk; ++c) {
            float4 mn = t.bmin[kids[c]], mx = t.bmax[kids[c]];
            area[c] = box_half_area(make_float3(mn.x, mn.y, mn.z), make_float3(mx.x, mx.y, mx.z));
        }
        for (int i = 1; i < nk; ++i) {          // insertion sort, descending
            uint32_t kid = kids[i];
            float ar = area[i];
            int j = i - 1;
            while (j >= 0 && area[j] < ar) { kids[j + 1] = kids[j]; area[j + 1] = area[j]; --j; }
            kids[j + 1] = kid; area[j + 1] = ar;
        }
    }
    // a.child_sort = 2 (default): slots ascending by box centre along the axis on which the centres spread the most; the axis goes into
    // the node and a closest-hit ray pops the children low-to-high or high-to-low by the sign of its direction on that axis (near side
    // first, so tmax shrinks early). Measured against the collapse order at 1080p (profiles/traces/r01j_trace.log, r01k_trace.log):
    // reflection pass 1.13 -> 0.91 ms (3 M triangles), 0.83 -> 0.69 ms (260 k); primary rays 0.675 -> 0.583 ms; shadow + AO unchanged
    // (0.780 -> 0.770 ms) with the any-hit rays popping lowest-first regardless of direction.
    uint32_t order_axis = 3u;
    if (a.child_sort == 2) {
        float c3[8][3], lo[3] = {3e38f, 3e38f, 3e38f}, hi[3] = {-3e38f, -3e38f, -3e38f};
        for (int c = 0; c < nk; ++c) {
            float4 mn = t.bmin[kids[c]], mx = t.bmax[kids[c]];
            c3[c][0] = mn.x + mx.x; c3[c][1] = mn.y + mx.y; c3[c][2] = mn.z + mx.z;
            for (int ax = 0; ax < 3; ++ax) { lo[ax] = fminf(lo[ax], c3[c][ax]); hi[ax] = fmaxf(hi[ax], c3[c][ax]); }
        }
        order_axis = 0u;
        if (hi[1] - lo[1] > hi[order_axis] - lo[order_axis]) order_axis = 1u;
        if (hi[2] - lo[2] > hi[order_axis] - lo[order_axis]) order_axis = 2u;
        for (int i = 1; i < nk; ++i) {          // insertion sort, ascending, stable
            uint32_t kid = kids[i];
            float key = c3[i][order_axis];
            int j = i - 1;
            while (j >= 0 && c3[j][order_axis] > key) { kids[j + 1] = kids[j]; c3[j + 1][order_axis] = c3[j][order_axis]; --j; }
            kids[j + 1] = kid; c3[j + 1][order_axis] = key;
        }
    }
    // internal children first (stable): slot index == child ordinal, and the internal-child mask is a run of low bits
    {
        uint32_t tmp[8];
        int k = 0;
        for (int c = 0; c < nk; ++c) if (!leaf_like(t, kids[c])) tmp[k++] = kids[c];
        for (int c = 0; c < nk; ++c) if (leaf_like(t, kids[c])) tmp[k++] = kids[c];
        for (int c = 0; c < nk; ++c) kids[c] = tmp[c];
    }
    float4 nmn = t.bmin[item.x], nmx = t.bmax[item.x];
    emit_wide_node(a, item.y, kids, nk, make_float3(nmn.x, nmn.y, nmn.z), make_float3(nmx.x, nmx.y, nmx.z), order_axis);
}

__global__ void widen_kernel(const __grid_constant__ WidenArgs a) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n_in) return;
    const Tree2 &t = a.t;
    const uint2 item = a.queue_in[i];
    uint32_t kids[8];
    int nk = 2;
    kids[0] = t.child_l[item.x];
    kids[1] = t.child_r[item.x];
    // Collapse rule. Every slot of a wide node is slab-tested whether it is used or not, so filling slots is free:
    //   1. an internal child whose whole subtree fits into the free slots (lcount - 1 <= free) is absorbed first, best
    //      area-per-slot first — this removes the small, mostly empty nodes an LBVH leaves at the bottom;
    //   2. otherwise the child with the largest surface area is opened one level (the usual SAH-greedy rule).
    while (nk < 8) {
        const int free_slots = 8 - nk;
        int best = -1, best_abs = -1;
        float best_area = -1.0f, best_ratio = -1.0f;
        for (int c = 0; c < nk; ++c) {
            if (leaf_like(t, kids[c])) continue;
            float4 mn = t.bmin[kids[c]], mx = t.bmax[kids[c]];
            float ar = box_half_area(make_float3(mn.x, mn.y, mn.z), make_float3(mx.x, mx.y, mx.z));
            if (ar > best_area) { best_area = ar; best = c; }
            const int need = (int)t.lcount[kids[c]] - 1;
            if (need <= free_slots) {
                float ratio = ar / (float)max(need, 1);
                if (ratio > best_ratio) { best_ratio = ratio; best_abs = c; }
            }
        }
        if (best < 0) break;
        if (best_abs >= 0) best = best_abs;
        uint32_t id = kids[best];
        kids[best] = t.child_l[id];
        kids[nk++] = t.child_r[id];
    }
    finish_wide_node(a, item, kids, nk);
}

// Children of the wide node rooted at binary node item.x, read off the decisions of collapse_dp_kernel: (node, i) = "node is
// represented by at most i roots"; a pair splits into (left, k) + (right, i - k) or falls back to (node, i - 1), and ends at i = 1
// or at a single triangle. The shares of the pending pairs always add up to the free slots, so eight stack entries are enough.
__global__ void widen_dp_kernel(const __grid_constant__ WidenArgs a) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n_in) return;
    const Tree2 &t = a.t;
    const uint2 item = a.queue_in[i];
    uint32_t kids[8];
    int nk = 0;
    uint32_t st_id[8];
    int st_share[8];
    int sp = 0;
    {
        const int k8 = a.dp_dec[(size_t)item.x * 8];          // 1..7: this node was chosen as a wide node
        st_id[sp] = t.child_r[item.x]; st_share[sp++] = 8 - k8;
        st_id[sp] = t.child_l[item.x]; st_share[sp++] = k8;
    }
    while (sp > 0 && nk < 8) {
        const uint32_t id = st_id[--sp];
        const int share = st_share[sp];
        if (id >= t.n - 1 || share <= 1) { kids[nk++] = id; continue; }
        const int k = a.dp_dec[(size_t)id * 8 + (share - 1)];
        if (k == 0) { st_id[sp] = id; st_share[sp++] = share - 1; continue; }
        st_id[sp] = t.child_r[id]; st_share[sp++] = share - k;
        st_id[sp] = t.child_l[id]; st_share[sp++] = k;
    }
    finish_wide_node(a, item, kids, nk);
}

// the whole tree is a single leaf (<= kMaxLeafTris triangles, or the root collapsed)
__global__ void widen_single_leaf_kernel(const __grid_constant__ WidenArgs a, uint32_t root_id) {
    if (blockIdx.x || threadIdx.x) return;
    uint32_t kids[1] = {root_id};
    float4 nmn = a.t.bmin[root_id], nmx = a.t.bmax[root_id];
    emit_wide_node(a, 0, kids, 1, make_float3(nmn.x, nmn.y, nmn.z), make_float3(nmx.x, nmx.y, nmx.z));
}

template <typename T>
int dmalloc(T **p, size_t count) {
    VHR_CUDA_CHECK(cudaMalloc((void **)p, std::max<size_t>(count, 1) * sizeof(T)));
    return VHR_OK;
}

}  // namespace

void free_bvh(vhr_context *ctx) {
    Bvh &b = ctx->bvh;
    if (b.wide_nodes) cudaFree(b.wide_nodes);
    if (b.tri_verts) cudaFree(b.tri_verts);
    b = Bvh();
}

// One build with the given hierarchy builder (1 PLOC, 0 radix tree). *retry_radix is set when the failure is one the radix tree does
// not share: PLOC needing more rounds than allowed, or a tree deeper than the traversal stack.
static int build_bvh_with(vhr_context *ctx, const int builder, bool *retry_radix) {
    Bvh &bvh = ctx->bvh;
    bvh = Bvh();
    cudaStream_t st = ctx->stream;
    // per-primitive triangle prefix (host copy of the primitives is what the caller just handed us; re-read from device
    // would need a sync anyway)
    std::vector<Primitive> prims(ctx->n_primitives);
    if (ctx->n_primitives) {
        VHR_CUDA_CHECK(cudaMemcpyAsync(prims.data(), ctx->d_primitives, prims.size() * sizeof(Primitive), cudaMemcpyDeviceToHost, st));
        VHR_CUDA_CHECK(cudaStreamSynchronize(st));
    }
    std::vector<uint32_t> prefix(ctx->n_primitives + 1, 0);
    uint64_t total = 0;
    for (uint32_t g = 0; g < ctx->n_primitives; ++g) {
        prefix[g] = (uint32_t)total;
        total += prims[g].index_count / 3;
    }
    prefix[ctx->n_primitives] = (uint32_t)total;
    if (total >= 0x3fffffffull) return fail(VHR_ERR_INVALID, "too many triangles (%llu)", (unsigned long long)total);
    const uint32_t n = (uint32_t)total;
    bvh.n_tris = n;
    bvh.stats.n_triangles = n;
    bvh.stats.max_leaf_size = kMaxLeafTris;
    if (n == 0) return VHR_OK;

    cudaEvent_t ev0, ev1;
    VHR_CUDA_CHECK(cudaEventCreate(&ev0));
    VHR_CUDA_CHECK(cudaEventCreate(&ev1));
    VHR_CUDA_CHECK(cudaEventRecord(ev0, st));

    uint32_t *d_prefix = nullptr;
    TriRef *d_tris = nullptr;
    int *d_scene = nullptr;
    uint64_t *d_keys = nullptr, *d_keys2 = nullptr;
    uint32_t *d_vals = nullptr, *d_vals2 = nullptr;
    void *d_tmp = nullptr;
    Tree2 t = {};
    t.n = n;
    uint2 *d_q[2] = {nullptr, nullptr};
    uint32_t *d_counters = nullptr;
    float *d_sah = nullptr;
    WideNode *d_wide = nullptr;
    TriRef *d_tris_out = nullptr;
    int rc = VHR_OK;
    std::vector<void *> scratch;
    auto track = [&](void *p) { scratch.push_back(p); };
#define TRY(x) do { rc = (x); if (rc) goto done; } while (0)
#define TRYCUDA(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) { rc = fail(VHR_ERR_CUDA, "%s: %s", #x, cudaGetErrorString(e__)); goto done; } } while (0)

    {
        TRY(dmalloc(&d_prefix, prefix.size())); track(d_prefix);
        TRY(dmalloc(&d_tris, n)); track(d_tris);
        TRY(dmalloc(&d_scene, 6)); track(d_scene);
        TRYCUDA(cudaMemcpyAsync(d_prefix, prefix.data(), prefix.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        int init[6] = {0x7fffffff, 0x7fffffff, 0x7fffffff, (int)0x80000000, (int)0x80000000, (int)0x80000000};
        TRYCUDA(cudaMemcpyAsync(d_scene, init, sizeof(init), cudaMemcpyHostToDevice, st));
        const int B = 256;
        const uint32_t G = (n + B - 1) / B;
        extract_triangles_kernel<<<G, B, 0, st>>>(ctx->d_vertices, ctx->d_indices, ctx->d_primitives, d_prefix, ctx->n_primitives, n, d_tris, d_scene);
        TRYCUDA(cudaGetLastError()); ctx->launches++;

        TRY(dmalloc(&d_keys, n)); track(d_keys);
        TRY(dmalloc(&d_keys2, n)); track(d_keys2);
        TRY(dmalloc(&d_vals, n)); track(d_vals);
        TRY(dmalloc(&d_vals2, n)); track(d_vals2);
        morton_kernel<<<G, B, 0, st>>>(d_tris, n, d_scene, d_keys, d_vals, getenv("VHR_MORTON_MODE") ? atoi(getenv("VHR_MORTON_MODE")) : 1);
        TRYCUDA(cudaGetLastError()); ctx->launches++;
        size_t tmp_bytes = 0;
        TRYCUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_keys, d_keys2, d_vals, d_vals2, (int)n, 0, 63, st));
        TRYCUDA(cudaMalloc(&d_tmp, std::max<size_t>(tmp_bytes, 16))); track(d_tmp);
        TRYCUDA(cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, d_keys, d_keys2, d_vals, d_vals2, (int)n, 0, 63, st));
        ctx->launches += 8;   // CUB radix sort passes (upsweep/scan/downsweep per digit; approximate)

        const uint32_t n_inner = n > 1 ? n - 1 : 0;
        TRY(dmalloc(&t.child_l, n_inner)); track(t.child_l);
        TRY(dmalloc(&t.child_r, n_inner)); track(t.child_r);
        TRY(dmalloc(&t.parent, 2 * (size_t)n)); track(t.parent);
        TRY(dmalloc(&t.range_first, n_inner)); track(t.range_first);
        TRY(dmalloc(&t.bmin, 2 * (size_t)n)); track(t.bmin);
        TRY(dmalloc(&t.bmax, 2 * (size_t)n)); track(t.bmax);
        TRY(dmalloc(&t.cluster, n_inner)); track(t.cluster);
        TRY(dmalloc(&t.visit, n_inner)); track(t.visit);
        TRY(dmalloc(&t.lcount, 2 * (size_t)n)); track(t.lcount);
        TRYCUDA(cudaMemsetAsync(t.visit, 0, std::max<size_t>(n_inner, 1) * sizeof(int), st));
        TRYCUDA(cudaMemsetAsync(t.cluster, 0, std::max<size_t>(n_inner, 1), st));
        // 1 (default) PLOC, 0 radix tree (Karras). Measured at 1080p, shadow + AO / reflection pass (profiles/traces/r01i_trace.log): 260 k triangles
        // 0.637 -> 0.596 / 0.888 -> 0.795 ms, 1 M 0.716 -> 0.681 / 1.058 -> 0.978 ms, 3 M 0.782 -> 0.756 / 1.25 -> 1.09 ms; build 10 -> 14.5 ms at 3 M.
        // (builder: parameter)
        if (n_inner && builder == 1) {
            const int radius = getenv("VHR_PLOC_RADIUS") ? atoi(getenv("VHR_PLOC_RADIUS")) : 8;      // 4 / 8 / 16 / 32 measured: 8 is the fastest to trace
            PlocBuf buf[2];
            uint32_t *d_nn = nullptr, *d_valid = nullptr, *d_pos = nullptr;
            for (int k = 0; k < 2; ++k) {
                TRY(dmalloc(&buf[k].id, n)); track(buf[k].id);
                TRY(dmalloc(&buf[k].mn, n)); track(buf[k].mn);
                TRY(dmalloc(&buf[k].mx, n)); track(buf[k].mx);
            }
            TRY(dmalloc(&d_nn, n)); track(d_nn);
            TRY(dmalloc(&d_valid, (size_t)n + 1)); track(d_valid);
            TRY(dmalloc(&d_pos, (size_t)n + 1)); track(d_pos);
            size_t scan_bytes = 0;
            TRYCUDA(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, d_valid, d_pos, (int)n + 1, st));
            void *d_scan = nullptr;
            TRYCUDA(cudaMalloc(&d_scan, std::max<size_t>(scan_bytes, 16))); track(d_scan);
            ploc_init_kernel<<<G, B, 0, st>>>(d_tris, d_vals2, n, buf[0]);
            TRYCUDA(cudaGetLastError()); ctx->launches++;
            uint32_t N = n;
            int cur = 0;
            for (int round = 0; N > 1; ++round) {
                if (round > kPlocMaxRounds) { *retry_radix = true; rc = fail(VHR_ERR_INVALID, "PLOC needs more than %d rounds (one merge per round: geometry strung out with growing gaps)", kPlocMaxRounds); goto done; }
                const uint32_t g = (N + 255) / 256;
                if (radius >= 32) ploc_nn_kernel<32><<<g, 256, 0, st>>>(buf[cur], N, d_nn);
                else if (radius >= 16) ploc_nn_kernel<16><<<g, 256, 0, st>>>(buf[cur], N, d_nn);
                else if (radius >= 8) ploc_nn_kernel<8><<<g, 256, 0, st>>>(buf[cur], N, d_nn);
                else ploc_nn_kernel<4><<<g, 256, 0, st>>>(buf[cur], N, d_nn);
                ploc_flag_kernel<<<(N + 256) / 256, 256, 0, st>>>(N, d_nn, d_valid);
                TRYCUDA(cub::DeviceScan::ExclusiveSum(d_scan, scan_bytes, d_valid, d_pos, (int)N + 1, st));
                ploc_merge_kernel<<<g, 256, 0, st>>>(buf[cur], buf[cur ^ 1], N, d_nn, d_valid, d_pos, t);
                TRYCUDA(cudaGetLastError()); ctx->launches += 4;
                uint32_t next = 0;
                TRYCUDA(cudaMemcpyAsync(&next, d_pos + N, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
                TRYCUDA(cudaStreamSynchronize(st));
                if (next >= N || next == 0) { rc = fail(VHR_ERR_CUDA, "PLOC round %d merged nothing", round); goto done; }
                N = next;
                cur ^= 1;
            }
            const uint32_t none = 0xffffffffu;
            TRYCUDA(cudaMemcpyAsync(t.parent, &none, sizeof(uint32_t), cudaMemcpyHostToDevice, st));      // the last merge got id 0: the root
            TRYCUDA(cudaStreamSynchronize(st));
        } else if (n_inner) {
            karras_kernel<<<(n_inner + B - 1) / B, B, 0, st>>>(d_keys2, t);
            TRYCUDA(cudaGetLastError()); ctx->launches++;
        }
        const float tri_cost = getenv("VHR_BVH_CT") ? (float)atof(getenv("VHR_BVH_CT")) : 1.0f;
        // triangles per leaf slot: kMaxLeafTris (3) by default; VHR_MAX_LEAF_TRIS=1..4 for study (8 slots x 4 = the 32-bit triangle mask of a node)
        const uint32_t max_leaf = (uint32_t)std::min(4, std::max(1, getenv("VHR_MAX_LEAF_TRIS") ? atoi(getenv("VHR_MAX_LEAF_TRIS")) : kMaxLeafTris));
        bvh.stats.max_leaf_size = max_leaf;
        refit_kernel<<<G, B, 0, st>>>(d_tris, d_vals2, t, tri_cost, max_leaf);
        TRYCUDA(cudaGetLastError()); ctx->launches++;
        // VHR_COLLAPSE: 1 (default) cost-optimal collapse (collapse_dp_kernel), 0 the greedy one; VHR_DP_NODE_COST = cost of visiting one
        // 8-wide node in units of one triangle test per unit area (0.5 / 1 / 2 measured: 1). Measured at 1080p (profiles/traces/r01v_trace.log):
        // slots in use 69.5 -> 89.1 %, SAH 31.8 -> 30.5, shadow + AO pass 0.721 -> 0.702 ms at 3 M triangles and 0.581 -> 0.562 ms at 260 k,
        // reflection pass 0.890 -> 0.876 / 0.691 -> 0.646 ms, primary rays 0.554 -> 0.552 / 0.428 -> 0.406 ms; build + 1 ms.
        const int collapse = getenv("VHR_COLLAPSE") ? atoi(getenv("VHR_COLLAPSE")) : 1;
        DpTables dp = {nullptr, nullptr};
        if (collapse == 1 && n_inner) {
            TRY(dmalloc(&dp.cost, (size_t)n_inner * 7)); track(dp.cost);
            TRY(dmalloc(&dp.dec, (size_t)n_inner * 8)); track(dp.dec);
            TRYCUDA(cudaMemsetAsync(t.visit, 0, (size_t)n_inner * sizeof(int), st));
            collapse_dp_kernel<<<G, B, 0, st>>>(t, dp, getenv("VHR_DP_NODE_COST") ? (float)atof(getenv("VHR_DP_NODE_COST")) : 1.0f, tri_cost, max_leaf);
            TRYCUDA(cudaGetLastError()); ctx->launches++;
        }

        // widen
        TRY(dmalloc(&d_wide, (size_t)n)); track(d_wide);
        TRY(dmalloc(&d_tris_out, (size_t)n));   // kept on success
        TRY(dmalloc(&d_q[0], (size_t)n)); track(d_q[0]);
        TRY(dmalloc(&d_q[1], (size_t)n)); track(d_q[1]);
        TRY(dmalloc(&d_counters, 4)); track(d_counters);
        TRY(dmalloc(&d_sah, 1)); track(d_sah);
        uint32_t counters[4] = {1, 0, 0, 0};
        TRYCUDA(cudaMemcpyAsync(d_counters, counters, sizeof(counters), cudaMemcpyHostToDevice, st));
        TRYCUDA(cudaMemsetAsync(d_sah, 0, sizeof(float), st));
        WidenArgs a;
        a.t = t; a.tris = d_tris; a.order = d_vals2; a.wide = d_wide; a.tris_out = d_tris_out;
        a.counters = d_counters; a.sah = d_sah;
        a.child_sort = getenv("VHR_CHILD_SORT") ? atoi(getenv("VHR_CHILD_SORT")) : 2;
        a.dp_dec = dp.dec;
        uint8_t root_cluster = 0;
        if (n_inner) {
            TRYCUDA(cudaMemcpyAsync(&root_cluster, t.cluster, 1, cudaMemcpyDeviceToHost, st));
            TRYCUDA(cudaStreamSynchronize(st));
        }
        if (n_inner == 0 || root_cluster) {
            a.queue_in = d_q[0]; a.queue_out = d_q[1]; a.n_in = 0;
            bvh.stats.wide_depth = 1;
            widen_single_leaf_kernel<<<1, 32, 0, st>>>(a, 0u);
            TRYCUDA(cudaGetLastError()); ctx->launches++;
        } else {
            uint2 root_item = make_uint2(0u, 0u);
            TRYCUDA(cudaMemcpyAsync(d_q[0], &root_item, sizeof(uint2), cudaMemcpyHostToDevice, st));
            uint32_t n_in = 1;
            int cur = 0;
            for (int level = 0; n_in > 0 && level < 256; ++level) {
                bvh.stats.wide_depth = (uint32_t)level + 1;
                TRYCUDA(cudaMemsetAsync(d_counters + 2, 0, sizeof(uint32_t), st));
                a.queue_in = d_q[cur]; a.queue_out = d_q[cur ^ 1]; a.n_in = n_in;
                if (a.dp_dec) widen_dp_kernel<<<(n_in + 127) / 128, 128, 0, st>>>(a);
                else widen_kernel<<<(n_in + 127) / 128, 128, 0, st>>>(a);
                TRYCUDA(cudaGetLastError()); ctx->launches++;
                TRYCUDA(cudaMemcpyAsync(&n_in, d_counters + 2, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
                TRYCUDA(cudaStreamSynchronize(st));
                cur ^= 1;
            }
            if (n_in) { *retry_radix = builder == 1; rc = fail(VHR_ERR_CUDA, "BVH widening did not terminate"); goto done; }
            if ((int)bvh.stats.wide_depth > kStackSize) {
                *retry_radix = builder == 1;
                rc = fail(VHR_ERR_INVALID, "BVH is %u levels deep, the traversal stack holds %d (degenerate geometry?)", bvh.stats.wide_depth, kStackSize);
                goto done;
            }
        }
        TRYCUDA(cudaMemcpyAsync(counters, d_counters, sizeof(counters), cudaMemcpyDeviceToHost, st));
        float sah_num = 0.0f;
        TRYCUDA(cudaMemcpyAsync(&sah_num, d_sah, sizeof(float), cudaMemcpyDeviceToHost, st));
        int scene[6];
        TRYCUDA(cudaMemcpyAsync(scene, d_scene, sizeof(scene), cudaMemcpyDeviceToHost, st));
        TRYCUDA(cudaStreamSynchronize(st));
        if (counters[1] != n) { rc = fail(VHR_ERR_CUDA, "BVH build emitted %u of %u triangles", counters[1], n); goto done; }
        bvh.n_wide = counters[0];
        bvh.n_nodes2 = 2 * n - 1;
        // shrink the node array to its final size
        WideNode *final_nodes = nullptr;
        TRYCUDA(cudaMalloc((void **)&final_nodes, (size_t)bvh.n_wide * sizeof(WideNode)));
        TRYCUDA(cudaMemcpyAsync(final_nodes, d_wide, (size_t)bvh.n_wide * sizeof(WideNode), cudaMemcpyDeviceToDevice, st));
        bvh.wide_nodes = final_nodes;
        bvh.tri_verts = reinterpret_cast<float4 *>(d_tris_out);
        d_tris_out = nullptr;
        for (int k = 0; k < 3; ++k) {
            bvh.stats.scene_min[k] = ordered_to_float(scene[k]);
            bvh.stats.scene_max[k] = ordered_to_float(scene[3 + k]);
        }
        float ex = bvh.stats.scene_max[0] - bvh.stats.scene_min[0], ey = bvh.stats.scene_max[1] - bvh.stats.scene_min[1],
              ez = bvh.stats.scene_max[2] - bvh.stats.scene_min[2];
        float root_area = ex * ey + ey * ez + ez * ex;
        bvh.stats.sah_cost = root_area > 0.0f ? sah_num / root_area : 0.0f;
        bvh.stats.n_bvh2_nodes = bvh.n_nodes2;
        bvh.stats.n_wide_nodes = bvh.n_wide;
        bvh.stats.n_used_slots = counters[3];
        TRYCUDA(cudaEventRecord(ev1, st));
        TRYCUDA(cudaEventSynchronize(ev1));
        float ms = 0.0f;
        cudaEventElapsedTime(&ms, ev0, ev1);
        bvh.stats.build_ms = ms;
    }
done:
    cudaStreamSynchronize(st);
    for (void *p : scratch) cudaFree(p);
    if (d_tris_out) cudaFree(d_tris_out);
    cudaEventDestroy(ev0);
    cudaEventDestroy(ev1);
    if (rc) free_bvh(ctx);
    return rc;
#undef TRY
#undef TRYCUDA
}

// PLOC by default (VHR_BVH_BUILDER=0: radix tree). Agglomerative clustering degenerates on geometry strung out along a line with
// steadily growing gaps (every cluster's nearest neighbour is on the same side: one merge per round, a tree as deep as it is long);
// the radix tree's depth is bounded by the key length, so such a scene is rebuilt with it instead of being refused.
int build_bvh(vhr_context *ctx) {
    const int builder = getenv("VHR_BVH_BUILDER") ? atoi(getenv("VHR_BVH_BUILDER")) : 1;
    bool retry_radix = false;
    int rc = build_bvh_with(ctx, builder, &retry_radix);
    if (rc != VHR_OK && retry_radix) {
        retry_radix = false;
        rc = build_bvh_with(ctx, 0, &retry_radix);
    }
    return rc;
}

}  // namespace vhr
