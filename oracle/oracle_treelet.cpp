// oracle_treelet.cpp — CPU restatement of the Fallback Layer's treelet optimisation pass
// (Karras & Aila 2013, n = 7), which GpuBvh2Builder runs on every bottom-level hierarchy between
// BuildBVHSplits and ComputeAABBs (FL/GpuBVH2Builder.cpp:312-326, ENABLE_TREELET_REORDERING = 1,
// FL/RayTracingHlslCompat.h:23).  TEST INFRASTRUCTURE (see oracle.h).  "FL/" =
// /root/reference/externals/D3D12RaytracingFallback/src/.
//
//   FL/TreeletReorder.cpp:38-109   pass count: PREFER_FAST_BUILD 0, PREFER_FAST_TRACE 3, otherwise 1;
//                                  MinTrianglesPerTreelet = 7, doubled every pass; a pass is skipped
//                                  when it exceeds the element count
//   FL/ClearBuffers.hlsl           counters and base-treelet list reset per pass
//   FL/FindTreelets.hlsl:31-89     bottom-up AABBs + triangle counts; the first node on every leaf-to-root
//                                  path whose subtree holds >= MinTrianglesPerTreelet triangles is a base
//                                  treelet root
//   FL/TreeletReorder.hlsl:40-345  per base root: FormTreelet -> FindOptimalPartitions -> ReformTree, then
//                                  climb (second child to arrive continues), i.e. EVERY internal node whose
//                                  subtree holds >= MinTrianglesPerTreelet triangles is optimised once,
//                                  children before parents
//
// The GPU pass is deterministic although its scheduling is not: a node is optimised only after both child
// subtrees are final, and the triangle count that admits it is the invariant leaf count of its subtree.
// The sequential order used here is the reverse pre-order of the hierarchy as it stands before the pass.
//
// Arithmetic: every expression is evaluated left to right in IEEE fp32 without contraction (the HLSL is not
// `precise`; pinned unfused, like the rest of the oracle).  The reference's cost model mixes normalised leaf
// costs with raw surface areas (TreeletReorder.hlsl:126-133 vs :180-183); it is restated literally.
// One deviation: FormTreelet's "largest surface area > 0.0" scan (:57-71) falls through to node 0 when no
// internal treelet leaf has a positive finite area — impossible for finite input because every leaf box is
// padded by AABB_Min_Padding; here such a treelet is left unchanged instead of corrupting the tree.
// FindTreelets.hlsl:74 reads ParentIndex without GetActualParentIndex, which is out of bounds from the
// second PREFER_FAST_TRACE pass on (the collapse bit is set by then); the climb here always masks the bit.
#include <cfloat>

#include "oracle_internal.h"

namespace orc {

namespace {

constexpr uint32_t kFull = 7;                    // FullTreeletSize (FL/TreeletReorderBindings.h:34)
constexpr uint32_t kSubsets = 1u << kFull;       // NumTreeletSplitPermutations
constexpr uint32_t kCollapseBit = 0x80000000u;   // HierarchyNode::IsCollapseChildren (FL/RayTracingHlslCompat.h:57)
constexpr float kCostBox = 1.2f, kCostTri = 1.0f;  // TreeletReorder.hlsl:21-22

inline Aabb combine(Aabb a, Aabb b) { return Aabb{vmin(a.mn, b.mn), vmax(a.mx, b.mx)}; }
// ComputeBoxSurfaceArea: FL/TreeletReorderBindings.h:91-95
inline float surface_area(Aabb a) {
    f3 d = a.mx - a.mn;
    return 2.0f * ((d.x * d.y + d.x * d.z) + d.y * d.z);
}

struct Pass {
    uint32_t n, nInternal;
    rt_hierarchy_node *hier;
    std::vector<Aabb> aabb;

    bool is_leaf(uint32_t node) const { return node >= nInternal; }

    void optimise(uint32_t root) {
        uint32_t leaves[kFull], internals[kFull - 1];
        // FormTreelet: TreeletReorder.hlsl:40-83
        internals[0] = root;
        leaves[0] = hier[root].left;
        leaves[1] = hier[root].right;
        for (uint32_t size = 2; size < kFull; ++size) {
            float largest = 0.0f;
            int pick = -1;
            for (uint32_t i = 0; i < size; ++i) {
                if (is_leaf(leaves[i])) continue;
                float sa = surface_area(aabb[leaves[i]]);
                if (sa > largest) largest = sa, pick = int(i);
            }
            if (pick < 0) return;  // see header: unreachable for finite input
            const uint32_t node = leaves[pick];
            internals[size - 1] = node;
            leaves[pick] = hier[node].left;
            leaves[size] = hier[node].right;
        }

        // FindOptimalPartitions: TreeletReorder.hlsl:85-193
        float cost[kSubsets];
        uint32_t part[kSubsets];
        cost[0] = 0.0f, part[0] = 0;
        for (uint32_t mask = 1; mask < kSubsets; ++mask) {
            Aabb a{mk(FLT_MAX, FLT_MAX, FLT_MAX), mk(-FLT_MAX, -FLT_MAX, -FLT_MAX)};
            for (uint32_t i = 0; i < kFull; ++i)
                if (mask & (1u << i)) a = combine(a, aabb[leaves[i]]);
            cost[mask] = surface_area(a);  // intermediate value (:122)
            part[mask] = 0;
        }
        const float rootArea = surface_area(aabb[root]);
        for (uint32_t i = 0; i < kFull; ++i) cost[1u << i] = kCostBox * surface_area(aabb[leaves[i]]) / rootArea;  // CalculateCost
        for (uint32_t size = 2; size <= kFull; ++size) {
            for (uint32_t mask = 1; mask < kSubsets; ++mask) {
                if (uint32_t(__builtin_popcount(mask)) != size) continue;
                float lowest = FLT_MAX;
                uint32_t best = 0;
                const uint32_t delta = (mask - 1) & mask;
                uint32_t p = (0u - delta) & mask;
                do {
                    const float c = cost[p] + cost[mask ^ p];
                    if (c < lowest) lowest = c, best = p;
                    p = (p - delta) & mask;
                } while (p != 0);
                // COMBINE_LEAF_NODES = 1 (FL/RayTracingHelper.hlsli:28)
                const float asLeaf = kCostTri * cost[mask] * float(size);
                const float asInternal = kCostBox * cost[mask] + lowest;
                cost[mask] = fminf(asInternal, asLeaf);
                part[mask] = best;
                if (asLeaf < asInternal) part[mask] |= 1u << kFull;  // bCollapseChildren flag
            }
        }

        // ReformTree: TreeletReorder.hlsl:195-266
        struct Entry {
            uint32_t mask, node;
        };
        Entry stack[kFull];
        uint32_t allocated = 1, sp = 1;
        stack[0] = Entry{kSubsets - 1, internals[0]};
        while (sp > 0) {
            const Entry cur = stack[--sp];
            Entry l, r;
            l.mask = part[cur.mask];
            const bool collapse = (l.mask & (1u << kFull)) != 0;
            l.mask &= kSubsets - 1;
            if (__builtin_popcount(l.mask) > 1) {
                l.node = internals[allocated++];
                stack[sp++] = l;
            } else {
                l.node = leaves[__builtin_ctz(l.mask)];
            }
            r.mask = cur.mask ^ l.mask;
            if (__builtin_popcount(r.mask) > 1) {
                r.node = internals[allocated++];
                stack[sp++] = r;
            } else {
                r.node = leaves[__builtin_ctz(r.mask)];
            }
            hier[cur.node].left = l.node;
            hier[cur.node].right = r.node;
            hier[l.node].parent = cur.node | (collapse ? kCollapseBit : 0u);
            hier[r.node].parent = cur.node | (collapse ? kCollapseBit : 0u);
        }
        for (int j = int(kFull) - 2; j >= 0; --j) {
            const uint32_t node = internals[j];
            aabb[node] = combine(aabb[hier[node].left], aabb[hier[node].right]);
        }
    }

    void run(const rt_primitive *sorted_prims, uint32_t minTris) {
        const uint32_t total = 2 * n - 1;
        aabb.resize(total);
        std::vector<uint32_t> count(total, 0), order, st;
        order.reserve(total);
        st.push_back(0);
        while (!st.empty()) {  // pre-order of the hierarchy as it stands before this pass
            const uint32_t v = st.back();
            st.pop_back();
            order.push_back(v);
            if (!is_leaf(v)) st.push_back(hier[v].left), st.push_back(hier[v].right);
        }
        for (size_t i = order.size(); i-- > 0;) {
            const uint32_t v = order[i];
            if (is_leaf(v)) {
                // ComputeLeafAABB (FindTreelets.hlsl:16-29): BoundingBoxToAABB(GetBoxDataFromTriangle(...))
                const float *p = sorted_prims[v - nInternal].v;
                f3 v0 = mk(p[0], p[1], p[2]), v1 = mk(p[3], p[4], p[5]), v2 = mk(p[6], p[7], p[8]);
                if (sorted_prims[v - nInternal].type == RT_PRIMITIVE_TYPE_PROCEDURAL) {
                    aabb[v] = Aabb{v0, v1};  // GetProceduralPrimitiveAABB, no box round trip (FindTreelets.hlsl:25-28)
                } else {
                    Aabb a{vmin(vmin(v0, v1), v2), vmax(vmax(v0, v1), v2)};
                    a.mn = vmin(a.mn, a.mx - mk(0.001f, 0.001f, 0.001f));  // AABB_Min_Padding
                    aabb[v] = box_to_aabb(aabb_to_box(a));
                }
                count[v] = 1;
            } else {
                const uint32_t l = hier[v].left, r = hier[v].right;
                aabb[v] = combine(aabb[l], aabb[r]);
                count[v] = count[l] + count[r];
                if (count[v] >= minTris) optimise(v);
            }
        }
    }
};

}  // namespace

uint32_t treelet_pass_count(uint32_t build_flags) {
    if (build_flags & RT_BUILD_FLAG_PREFER_FAST_BUILD) return 0;
    if (build_flags & RT_BUILD_FLAG_PREFER_FAST_TRACE) return 3;
    return 1;
}

void treelet_optimise(uint32_t n, rt_hierarchy_node *hier, const rt_primitive *sorted_prims, uint32_t build_flags) {
    if (n == 0) return;
    uint32_t minTris = kFull;
    const uint32_t passes = treelet_pass_count(build_flags);
    for (uint32_t i = 0; i < passes; ++i) {
        if (minTris > n) break;
        Pass p{n, n - 1, hier, {}};
        p.run(sorted_prims, minTris);
        minTris *= 2;
    }
}

}  // namespace orc

extern "C" void orc_treelet_optimise(uint32_t n, rt_hierarchy_node *hier, const rt_primitive *sorted_prims, uint32_t build_flags) {
    orc::treelet_optimise(n, hier, sorted_prims, build_flags);
}
