// treelet.cuh — treelet optimisation of the bottom-level hierarchy (Karras & Aila 2013, treelets of 7 leaves).
//
// B200-native replacement of FL/ClearBuffers.hlsl + FL/FindTreelets.hlsl:31-89 + FL/TreeletReorder.hlsl:40-345 as
// driven by FL/TreeletReorder.cpp:38-109 ("FL/" = externals/D3D12RaytracingFallback/src/).  Same result as the CPU
// restatement used by the parity tests, bit for bit: every internal node whose subtree holds at least
// MinTrianglesPerTreelet triangles is optimised once, children before parents, with the reference's cost arithmetic.
// What differs is HOW:
//   * one WARP per base treelet (the reference runs 32-thread groups whose thread 0 forms and re-forms the treelet
//     serially): the five greedy expansions are a warp arg-max, the 127 subset areas and the subset DP are spread
//     over the lanes, the six rewritten nodes are stored by six lanes;
//   * everything a step needs (7 leaf boxes, 128 costs, 128 partitions) lives in 1.3 KB of shared memory per warp;
//   * node boxes are read from the seven leaf boxes already on chip instead of being re-fetched per subset.
// Cross-warp hand-over is the reference's: the second child to arrive at a parent (atomic triangle counter) continues.
// Nodes written by other warps are read with ld.cg (L2), after the counter's fence.
#pragma once
#include <cfloat>

#include "common.cuh"

namespace treelet {

constexpr uint32_t kFull = 7;                   // FullTreeletSize (FL/TreeletReorderBindings.h:34)
constexpr uint32_t kSubsets = 1u << kFull;
constexpr uint32_t kCollapseBit = 0x80000000u;  // HierarchyNode::IsCollapseChildren (FL/RayTracingHlslCompat.h:57)
constexpr int kWarps = 4;                       // independent warps per block

// the 120 subsets of {0..6} with 2..7 members, grouped by size; kSizeBegin[s - 2] = first subset of size s
__constant__ uint8_t kSubsetBySize[120] = {
    3,  5,  6,  9,  10, 12, 17, 18, 20,  24,  33,  34,  36,  40,  48,  65,  66,  68,  72,  80,  96,  7,   11,  13,
    14, 19, 21, 22, 25, 26, 28, 35, 37,  38,  41,  42,  44,  49,  50,  52,  56,  67,  69,  70,  73,  74,  76,  81,
    82, 84, 88, 97, 98, 100, 104, 112, 15, 23,  27,  29,  30,  39,  43,  45,  46,  51,  53,  54,  57,  58,  60,  71,
    75, 77, 78, 83, 85, 86, 89, 90, 92,  99,  101, 102, 105, 106, 108, 113, 114, 116, 120, 31,  47,  55,  59,  61,
    62, 79, 87, 91, 93, 94, 103, 107, 109, 110, 115, 117, 118, 121, 122, 124, 63,  95,  111, 119, 123, 125, 126, 127};
__constant__ uint8_t kSizeBegin[7] = {0, 21, 56, 91, 112, 119, 120};

struct A6 {
    float mn[3], mx[3];
};
// node boxes: 6 floats per node (the reference's AABB struct), 8-byte aligned
__device__ __forceinline__ A6 ld_aabb(const float *aabbs, uint32_t i) {
    const float2 *p = reinterpret_cast<const float2 *>(aabbs + 6 * size_t(i));
    const float2 a = __ldcg(p), b = __ldcg(p + 1), c = __ldcg(p + 2);
    return A6{{a.x, a.y, b.x}, {b.y, c.x, c.y}};
}
__device__ __forceinline__ void st_aabb(float *aabbs, uint32_t i, const A6 &a) {
    float2 *p = reinterpret_cast<float2 *>(aabbs + 6 * size_t(i));
    __stcg(p, make_float2(a.mn[0], a.mn[1]));
    __stcg(p + 1, make_float2(a.mn[2], a.mx[0]));
    __stcg(p + 2, make_float2(a.mx[1], a.mx[2]));
}
__device__ __forceinline__ A6 combine(const A6 &a, const A6 &b) {  // CombineAABB
    A6 r;
#pragma unroll
    for (int k = 0; k < 3; ++k) r.mn[k] = fminf(a.mn[k], b.mn[k]), r.mx[k] = fmaxf(a.mx[k], b.mx[k]);
    return r;
}
// ComputeBoxSurfaceArea (FL/TreeletReorderBindings.h:91-95), left to right, unfused
__device__ __forceinline__ float area(const A6 &a) {
    const float dx = sub_(a.mx[0], a.mn[0]), dy = sub_(a.mx[1], a.mn[1]), dz = sub_(a.mx[2], a.mn[2]);
    return mul_(2.0f, add_(add_(mul_(dx, dy), mul_(dx, dz)), mul_(dy, dz)));
}

// FL/FindTreelets.hlsl:31-89 (+ ClearBuffers: num_tris and base[0] are zeroed by the caller).  One thread per leaf
// climbs, storing node boxes; the first node with >= min_tris triangles on its path is appended to base[1..].
__global__ void __launch_bounds__(256) k_find_treelets(uint32_t n, const rt_hierarchy_node *hier, const rt_packed_tri *tris,
                                                       uint32_t *num_tris, float *aabbs, uint32_t *base, uint32_t min_tris) {
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= n) return;
    const uint32_t nInternal = n - 1;
    uint32_t node = nInternal + slot;
    A6 a;
    {
        // ComputeLeafAABB: BoundingBoxToAABB(GetBoxDataFromTriangle(...)) — the box round trip is part of the value
        const float4 *q = reinterpret_cast<const float4 *>(tris + slot);
        const float4 q0 = q[0], q1 = q[1], q2 = q[2];
        const float v[9] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x};
        if (__float_as_uint(q2.w) & RT_PACKED_PROCEDURAL) {
            // GetProceduralPrimitiveAABB: the AABB as given, no box round trip (FindTreelets.hlsl:25-28)
#pragma unroll
            for (int k = 0; k < 3; ++k) a.mn[k] = v[k], a.mx[k] = v[3 + k];
        } else {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                float mn = fminf(fminf(v[k], v[3 + k]), v[6 + k]);
                const float mx = fmaxf(fmaxf(v[k], v[3 + k]), v[6 + k]);
                mn = fminf(mn, sub_(mx, 0.001f));  // AABB_Min_Padding
                const float c = mul_(add_(mn, mx), 0.5f), h = sub_(mx, c);
                a.mn[k] = sub_(c, h), a.mx[k] = add_(c, h);
            }
        }
    }
    uint32_t count = 1;
    while (true) {
        st_aabb(aabbs, node, a);
        if (count >= min_tris) {
            base[1 + atomicAdd(&base[0], 1u)] = node;
            return;
        }
        __threadfence();
        const uint32_t parent = hier[node].parent & ~kCollapseBit;
        const uint32_t other = atomicAdd(&num_tris[parent], count);
        if (other == 0) return;  // the sibling subtree continues
        __threadfence();
        node = parent;
        count += other;
        a = combine(ld_aabb(aabbs, hier[node].left), ld_aabb(aabbs, hier[node].right));
    }
}

// FL/TreeletReorder.hlsl:312-345: one warp per base treelet root; optimise, then climb while this warp is the second
// child to arrive.
// `local` (may be null): the per-node "subtree lies in one fit block" flags of k_hierarchy.  Re-forming a treelet
// whose root is not local can move leaves from outside the block under one of its inner nodes, so those nodes lose the
// flag; a treelet under a local root only permutes nodes and leaves of that block, and the flags stay true.
#ifndef RT_TREELET_MINBLOCKS
#define RT_TREELET_MINBLOCKS 12  // resident blocks per SM the register budget is set for: 40 registers, 48 warps per SM (A/B at 10 M triangles, one pass: 1 -> 10.68 ms, 10 -> 10.44, 12 -> 9.97, 16 -> 11.31)
#endif
__global__ void __launch_bounds__(32 * kWarps, RT_TREELET_MINBLOCKS) k_treelet_reorder(uint32_t n, rt_hierarchy_node *hier, uint32_t *num_tris,
                                                                 float *aabbs, const uint32_t *base, uint8_t *local) {
    __shared__ float s_cost[kWarps][kSubsets];
    __shared__ float s_box[kWarps][kFull][6];
    __shared__ uint32_t s_leaf[kWarps][8], s_int[kWarps][8];
    __shared__ uint32_t s_plan[kWarps][kFull - 1][4];  // {node, subset | collapse << 8, left, right}
    __shared__ uint8_t s_part[kWarps][kSubsets];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t id = blockIdx.x * kWarps + w;
    if (id >= base[0]) return;
    const uint32_t nInternal = n - 1;
    uint32_t node = base[1 + id];
    uint32_t *hw = reinterpret_cast<uint32_t *>(hier);  // {parent, left, right} x 3 words
    const unsigned full = 0xffffffffu;

    // children of `node`, known to every lane at the top of the loop
    uint32_t cl = 0, cr = 0;
    if (lane == 0) cl = __ldcg(hw + 3 * size_t(node) + 1), cr = __ldcg(hw + 3 * size_t(node) + 2);
    cl = __shfl_sync(full, cl, 0), cr = __shfl_sync(full, cr, 0);

    while (true) {
        // lane 0 starts the loads the climb at the end of this step needs; none of them is touched by this step
        // (ReformTree leaves the root's parent link and box as they are: the box is the exact union of the same leaves)
        uint32_t up = 0, ours = 0, pl = 0, pr = 0;
        A6 rootBox{};
        if (lane == 0 && node != 0) {
            up = __ldcg(hw + 3 * size_t(node)) & ~kCollapseBit;
            ours = __ldcg(num_tris + node);
            rootBox = ld_aabb(aabbs, node);
        }
        // ---- FormTreelet (:40-83): lane i owns treelet leaf i and prefetches that node's children with its box
        uint32_t my = 0xffffffffu, myL = 0, myR = 0;
        float myArea = 0.0f;
        A6 myBox{};
        auto fetch = [&]() {
            myBox = ld_aabb(aabbs, my);
            if (my < nInternal) myL = __ldcg(hw + 3 * size_t(my) + 1), myR = __ldcg(hw + 3 * size_t(my) + 2);
            myArea = area(myBox);
        };
        if (lane == 0) my = cl, s_int[w][0] = node;
        if (lane == 1) my = cr;
        if (lane < 2) fetch();
        bool formed = true;
        for (int size = 2; size < int(kFull); ++size) {
            // "surfaceArea > largestSurfaceArea" from 0.0, scanning i upwards: the largest positive area, lowest i on ties.
            // Positive floats order like their bit patterns: one integer max-reduction and one ballot.
            const uint32_t key = (lane < size && my < nInternal && myArea > 0.0f) ? __float_as_uint(myArea) : 0u;
            const uint32_t top = __reduce_max_sync(full, key);
            if (top == 0u) {  // no splittable leaf with a positive area: leave the treelet alone (unreachable for finite input: leaf boxes are padded)
                formed = false;
                break;
            }
            const int who = __ffs(__ballot_sync(full, key == top)) - 1;
            const uint32_t pick = __shfl_sync(full, my, who);
            const uint32_t l = __shfl_sync(full, myL, who), r = __shfl_sync(full, myR, who);
            if (lane == 0) s_int[w][size - 1] = pick;
            if (lane == who) my = l;
            if (lane == size) my = r;
            if (lane == who || lane == size) fetch();
        }
        if (lane == 0 && node != 0) pl = __ldcg(hw + 3 * size_t(up) + 1), pr = __ldcg(hw + 3 * size_t(up) + 2);

        if (formed) {
            if (lane < int(kFull)) {
                s_leaf[w][lane] = my;
#pragma unroll
                for (int k = 0; k < 3; ++k) s_box[w][lane][k] = myBox.mn[k], s_box[w][lane][3 + k] = myBox.mx[k];
            }
            __syncwarp();
            // ---- FindOptimalPartitions (:85-193)
            auto subset_box = [&](uint32_t mask) {
                A6 u{{FLT_MAX, FLT_MAX, FLT_MAX}, {-FLT_MAX, -FLT_MAX, -FLT_MAX}};
#pragma unroll
                for (int i = 0; i < int(kFull); ++i)
                    if (mask & (1u << i)) {
#pragma unroll
                        for (int k = 0; k < 3; ++k) u.mn[k] = fminf(u.mn[k], s_box[w][i][k]), u.mx[k] = fmaxf(u.mx[k], s_box[w][i][3 + k]);
                    }
                return u;
            };
            {
                // 127 subset areas, four per lane: subsets lane, lane+32, lane+64, lane+96 share the members given by the
                // lane's five low bits, so that union is formed once and extended by leaves 5 and 6
                A6 lo{{FLT_MAX, FLT_MAX, FLT_MAX}, {-FLT_MAX, -FLT_MAX, -FLT_MAX}};
#pragma unroll
                for (int i = 0; i < 5; ++i)
                    if (lane & (1 << i)) {
#pragma unroll
                        for (int k = 0; k < 3; ++k) lo.mn[k] = fminf(lo.mn[k], s_box[w][i][k]), lo.mx[k] = fmaxf(lo.mx[k], s_box[w][i][3 + k]);
                    }
                A6 b5, b6;
#pragma unroll
                for (int k = 0; k < 3; ++k)
                    b5.mn[k] = s_box[w][5][k], b5.mx[k] = s_box[w][5][3 + k], b6.mn[k] = s_box[w][6][k], b6.mx[k] = s_box[w][6][3 + k];
                const A6 u5 = combine(lo, b5);
                if (lane) s_cost[w][lane] = area(lo);  // intermediate value: raw surface area (:122)
                s_cost[w][lane + 32] = area(u5);
                s_cost[w][lane + 64] = area(combine(lo, b6));
                s_cost[w][lane + 96] = area(combine(u5, b6));
            }
            __syncwarp();
            const float rootArea = s_cost[w][kSubsets - 1];  // the root box is the union of the seven leaves
            __syncwarp();
            if (lane < int(kFull)) s_cost[w][1u << lane] = div_(mul_(1.2f, myArea), rootArea);  // CalculateCost (:24-28)
            __syncwarp();
            // Subset DP.  The reference's partition loop (:153-164) visits p = deposit(i, delta) for i = 1 .. 2^(size-1)-1 in
            // increasing i and keeps the first minimum, so a range of i can be given to each lane and the lanes merged
            // with (cost, i) ordered lexicographically.  Sizes 2-5: one lane per subset; size 6: four lanes per subset
            // (7 x 31 partitions); size 7: all lanes on the one subset (63 partitions).
            const float *cost = s_cost[w];
            auto finish = [&](uint32_t mask, int size, float lowest, uint32_t bestPart) {
                const float raw = cost[mask];
                const float asLeaf = mul_(mul_(1.0f, raw), float(size));  // COMBINE_LEAF_NODES = 1
                const float asInternal = add_(mul_(1.2f, raw), lowest);
                s_cost[w][mask] = fminf(asInternal, asLeaf);
                s_part[w][mask] = uint8_t(bestPart | (asLeaf < asInternal ? 0x80u : 0u));
            };
            for (int size = 2; size <= 5; ++size) {
                const int begin = kSizeBegin[size - 2], end = kSizeBegin[size - 1];
                for (int j = begin + lane; j < end; j += 32) {
                    const uint32_t mask = kSubsetBySize[j];
                    float lowest = FLT_MAX;
                    uint32_t bestPart = 0;
                    const uint32_t delta = (mask - 1) & mask;
                    uint32_t p = (0u - delta) & mask;
                    do {
                        const float c = add_(cost[p], cost[mask ^ p]);
                        if (c < lowest) lowest = c, bestPart = p;
                        p = (p - delta) & mask;
                    } while (p != 0);
                    finish(mask, size, lowest, bestPart);
                }
                __syncwarp();
            }
            auto deposit = [](uint32_t i, uint32_t delta) {  // bits of i into the set positions of delta, lowest first
                uint32_t r = 0;
                while (delta) {
                    const uint32_t low = delta & (0u - delta);
                    if (i & 1u) r |= low;
                    i >>= 1, delta ^= low;
                }
                return r;
            };
            auto merge = [&](float &lowest, uint32_t &first, uint32_t &bestPart, int o) {
                const float ol = __shfl_xor_sync(full, lowest, o);
                const uint32_t of = __shfl_xor_sync(full, first, o), op = __shfl_xor_sync(full, bestPart, o);
                if (ol < lowest || (ol == lowest && of < first)) lowest = ol, first = of, bestPart = op;
            };
            {  // size 6
                const uint32_t mask = kSubsetBySize[kSizeBegin[4] + min(lane >> 2, 6)];
                const uint32_t delta = (mask - 1) & mask;
                const uint32_t i0 = 1 + 8 * (lane & 3), i1 = min(i0 + 8, 32u);
                float lowest = FLT_MAX;
                uint32_t bestPart = 0, first = 0xffffffffu;
                uint32_t p = deposit(i0, delta);
                if (lane < 28)
                    for (uint32_t i = i0; i < i1; ++i) {
                        const float c = add_(cost[p], cost[mask ^ p]);
                        if (c < lowest) lowest = c, bestPart = p, first = i;
                        p = (p - delta) & mask;
                    }
                merge(lowest, first, bestPart, 1);
                merge(lowest, first, bestPart, 2);
                if (lane < 28 && (lane & 3) == 0) finish(mask, 6, lowest, bestPart);
            }
            __syncwarp();
            {  // size 7: delta = 0b1111110, deposit(i) = i << 1
                const uint32_t mask = kSubsets - 1, delta = mask - 1;
                float lowest = FLT_MAX;
                uint32_t bestPart = 0, first = 0xffffffffu;
#pragma unroll
                for (uint32_t i = 1 + 2 * lane; i < 3 + 2 * lane; ++i)
                    if (i < 64) {
                        const uint32_t p = (i << 1) & delta;
                        const float c = add_(cost[p], cost[mask ^ p]);
                        if (c < lowest) lowest = c, bestPart = p, first = i;
                    }
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) merge(lowest, first, bestPart, o);
                if (lane == 0) finish(mask, 7, lowest, bestPart);
            }
            __syncwarp();
            // ---- ReformTree (:195-266): lane 0 walks the partition, six lanes store the six nodes
            if (lane == 0) {
                uint32_t stackMask[kFull], stackNode[kFull];
                uint32_t allocated = 1, sp = 1, k = 0;
                stackMask[0] = kSubsets - 1, stackNode[0] = s_int[w][0];
                while (sp > 0) {
                    --sp;
                    const uint32_t curMask = stackMask[sp], curNode = stackNode[sp];
                    uint32_t lm = s_part[w][curMask];
                    const uint32_t collapse = lm >> 7;
                    lm &= kSubsets - 1;
                    uint32_t ln, rn;
                    if (__popc(lm) > 1) {
                        ln = s_int[w][allocated++];
                        stackMask[sp] = lm, stackNode[sp] = ln, ++sp;
                    } else {
                        ln = s_leaf[w][__ffs(lm) - 1];
                    }
                    const uint32_t rm = curMask ^ lm;
                    if (__popc(rm) > 1) {
                        rn = s_int[w][allocated++];
                        stackMask[sp] = rm, stackNode[sp] = rn, ++sp;
                    } else {
                        rn = s_leaf[w][__ffs(rm) - 1];
                    }
                    s_plan[w][k][0] = curNode, s_plan[w][k][1] = curMask | (collapse << 8), s_plan[w][k][2] = ln, s_plan[w][k][3] = rn;
                    ++k;
                }
            }
            __syncwarp();
            if (lane < int(kFull) - 1) {
                const uint32_t nd = s_plan[w][lane][0], mk = s_plan[w][lane][1], l = s_plan[w][lane][2], r = s_plan[w][lane][3];
                const uint32_t up = nd | ((mk >> 8) ? kCollapseBit : 0u);
                __stcg(hw + 3 * size_t(nd) + 1, l);
                __stcg(hw + 3 * size_t(nd) + 2, r);
                __stcg(hw + 3 * size_t(l), up);
                __stcg(hw + 3 * size_t(r), up);
                st_aabb(aabbs, nd, subset_box(mk & (kSubsets - 1)));
                if (local != nullptr && nd != node && __ldcg(local + node) == 0) __stcg(local + nd, uint8_t(0));
            }
        }

        // ---- TraverseToParent (:268-309)
        __threadfence();
        __syncwarp();
        uint32_t next = 0xffffffffu;
        if (lane == 0 && node != 0) {
            const uint32_t other = atomicAdd(&num_tris[up], ours);
            if (other != 0) {  // second to arrive: both subtrees are final
                __threadfence();
                st_aabb(aabbs, up, combine(rootBox, ld_aabb(aabbs, pl == node ? pr : pl)));
                __threadfence();
                next = up;
            }
        }
        next = __shfl_sync(full, next, 0);
        if (next == 0xffffffffu) return;
        node = next;
        cl = __shfl_sync(full, pl, 0), cr = __shfl_sync(full, pr, 0);
    }
}

}  // namespace treelet
