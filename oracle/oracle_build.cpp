// oracle_build.cpp — CPU restatement of the Fallback Layer's LBVH build passes.
// TEST INFRASTRUCTURE (see oracle.h).  Paths below are relative to
// /root/reference/externals/D3D12RaytracingFallback/src/ ("FL/").
#include <algorithm>
#include <cfloat>
#include <numeric>

#include "oracle_internal.h"

namespace orc {

// ---------------------------------------------------------------- load triangles
// FL/BottomLevelLoadTriangles.hlsli:15-126 (index fetch for R32 / R16 / no IB, optional 3x4
// transform) and FL/LoadPrimitivesBindings.h:72-79 (metadata).
static void load_triangles(const rt_geometry_desc *geoms, uint32_t n_geoms, std::vector<rt_primitive> &prims,
                           std::vector<rt_primitive_meta> &meta) {
    for (uint32_t g = 0; g < n_geoms; ++g) {
        const rt_geometry_desc &d = geoms[g];
        const uint8_t *vb = static_cast<const uint8_t *>(d.vertex_buffer);
        if (d.type == RT_GEOMETRY_TYPE_PROCEDURAL_AABBS) {
            // FL/LoadProceduralGeometry.hlsl:16-42, FL/LoadPrimitivesPass.cpp:122-152; pinned by UT:2616-2664
            for (uint32_t t = 0; t < d.vertex_count; ++t) {
                rt_primitive p;  // CreateProceduralGeometryPrimitive: NullPrimitive() + type + {min, max}
                p.type = RT_PRIMITIVE_TYPE_PROCEDURAL;
                std::memcpy(p.v, vb + size_t(t) * d.vertex_stride_bytes, 24);
                p.v[6] = p.v[7] = p.v[8] = 0.0f;
                prims.push_back(p);
                rt_primitive_meta m;
                m.geometryContributionToHitGroupIndex = g;
                m.primitiveIndex = t;
                m.geometryFlags = d.flags;
                meta.push_back(m);
            }
            continue;
        }
        uint32_t ntri = (d.index_format == 0 ? d.vertex_count : d.index_count) / 3;
        for (uint32_t t = 0; t < ntri; ++t) {
            uint32_t idx[3];
            for (int k = 0; k < 3; ++k) {
                if (d.index_format == 32)
                    idx[k] = static_cast<const uint32_t *>(d.index_buffer)[3 * t + k];
                else if (d.index_format == 16)
                    idx[k] = static_cast<const uint16_t *>(d.index_buffer)[3 * t + k];
                else
                    idx[k] = 3 * t + k;
            }
            rt_primitive p;
            p.type = 1;  // TRIANGLE_TYPE
            for (int k = 0; k < 3; ++k) {
                float v[3];
                std::memcpy(v, vb + size_t(idx[k]) * d.vertex_stride_bytes, 12);
                f3 q = mk(v[0], v[1], v[2]);
                if (d.transform3x4) q = xform_point(d.transform3x4, q);  // TransformVertex
                p.v[3 * k + 0] = q.x;
                p.v[3 * k + 1] = q.y;
                p.v[3 * k + 2] = q.z;
            }
            prims.push_back(p);
            rt_primitive_meta m;
            m.geometryContributionToHitGroupIndex = g;  // LoadPrimitivesPass.cpp:106
            m.primitiveIndex = t;                       // local index within its geometry
            m.geometryFlags = d.flags;
            meta.push_back(m);
        }
    }
}

// ---------------------------------------------------------------- Morton codes
// FL/CalculateMortonCodes.hlsli:73-118 (non-scaled variant; SCALED_MORTON_CODES is never defined).
static uint32_t morton_from_unit(f3 unit) {
    const uint32_t numBits = 10;
    const float maxCoord = 1024.0f;  // pow(2, numBits)
    f3 adj = mk(fminf(fmaxf(unit.x * maxCoord, 0.0f), maxCoord - 1), fminf(fmaxf(unit.y * maxCoord, 0.0f), maxCoord - 1),
                fminf(fmaxf(unit.z * maxCoord, 0.0f), maxCoord - 1));
    uint32_t coords[3] = {uint32_t(adj.y), uint32_t(adj.x), uint32_t(adj.z)};  // axis order (y, x, z)
    uint32_t code = 0;
    for (uint32_t b = 0; b < numBits; ++b)
        for (uint32_t a = 0; a < 3; ++a)
            if (coords[a] & (1u << b)) code |= 1u << (b * 3 + a);
    return code;
}

static uint32_t morton_from_centroid(f3 c, const float aabb[6]) {
    const float epsilon = 0.00001f;
    f3 mn = mk(aabb[0], aabb[1], aabb[2]), mx = mk(aabb[3], aabb[4], aabb[5]);
    f3 dim = vmax(mx - mn, mk(epsilon, epsilon, epsilon));
    f3 unit = (c - mn) / dim;
    return morton_from_unit(unit);
}

// ---------------------------------------------------------------- Karras hierarchy
// FL/BuildBVHSplits.hlsli:18-131
struct Karras {
    const uint32_t *codes;
    int64_t n;
    static int clz(uint32_t x) { return x ? __builtin_clz(x) : 32; }  // 31 - firstbithigh(x)
    int lcp(int64_t a, int64_t b) const {
        // the HLSL takes uints: a negative int wraps to >= 2^31 >= NumberOfElements
        if (a < 0 || b < 0 || a >= n || b >= n) return -1;
        uint32_t ca = codes[a], cb = codes[b];
        if (ca != cb) return clz(ca ^ cb);
        return clz(uint32_t(a) ^ uint32_t(b)) + 31;
    }
    void range(int64_t idx, int64_t &first, int64_t &last) const {
        int d = lcp(idx, idx + 1) - lcp(idx, idx - 1);
        d = std::min(std::max(d, -1), 1);
        int minPrefix = lcp(idx, idx - d);
        int64_t maxLength = 2;
        while (lcp(idx, idx + maxLength * d) > minPrefix) maxLength *= 4;
        int64_t length = 0;
        for (int64_t t = maxLength / 2; t > 0; t /= 2)
            if (lcp(idx, idx + (length + t) * d) > minPrefix) length += t;
        int64_t j = idx + length * d;
        first = std::min(idx, j);
        last = std::max(idx, j);
    }
    int64_t split(int64_t first, int64_t last) const {
        int commonPrefix = lcp(first, last);
        int64_t s = first, step = last - first;
        do {
            step = (step + 1) >> 1;
            int64_t ns = s + step;
            if (ns < last && lcp(first, ns) > commonPrefix) s = ns;
        } while (step > 1);
        return s;
    }
};

static void build_hierarchy(const uint32_t *codes, uint32_t n, rt_hierarchy_node *nodes) {
    if (n == 0) return;
    for (uint32_t i = 0; i < 2 * n - 1; ++i) nodes[i] = rt_hierarchy_node{0, 0, 0};
    Karras k{codes, int64_t(n)};
    const uint32_t leafOffset = n - 1;
    for (uint32_t idx = 0; idx + 1 < n; ++idx) {
        int64_t first, last;
        k.range(idx, first, last);
        int64_t s = k.split(first, last);
        uint32_t a = (s == first) ? leafOffset + uint32_t(s) : uint32_t(s);
        uint32_t b = (s + 1 == last) ? leafOffset + uint32_t(s) + 1 : uint32_t(s) + 1;
        nodes[idx].left = a;
        nodes[idx].right = b;
        nodes[a].parent = idx;
        nodes[b].parent = idx;
    }
}

// ---------------------------------------------------------------- bottom-up AABB fit
// FL/ComputeAABBs.hlsli:69-175.  The GPU pass lets the second child to arrive process the parent
// and order the children "smaller triangle count on the left"; on EQUAL counts the outcome depends
// on arrival order (:296-298).  Pinned here: swap iff count(right) < count(left), i.e. no swap on ties.
template <class LeafBox>
static void fit_boxes(uint32_t n, const rt_hierarchy_node *hier, rt_aabb_node *nodes, LeafBox leaf_box) {
    if (n == 0) return;
    const uint32_t nInternal = n - 1, total = 2 * n - 1;
    std::vector<uint32_t> count(total, 0);
    // iterative post-order
    std::vector<uint32_t> order;
    order.reserve(total);
    std::vector<uint32_t> st;
    st.push_back(0);
    while (!st.empty()) {
        uint32_t v = st.back();
        st.pop_back();
        order.push_back(v);
        if (v < nInternal) {
            st.push_back(hier[v].left);
            st.push_back(hier[v].right);
        }
    }
    for (size_t i = order.size(); i-- > 0;) {
        uint32_t v = order[i];
        if (v >= nInternal) {  // leaf
            uint32_t slot = v - nInternal;
            Box b = leaf_box(slot);
            nodes[v].center[0] = b.center.x, nodes[v].center[1] = b.center.y, nodes[v].center[2] = b.center.z;
            nodes[v].halfDim[0] = b.half.x, nodes[v].halfDim[1] = b.half.y, nodes[v].halfDim[2] = b.half.z;
            nodes[v].flags = slot | RT_NODE_LEAF_FLAG;
            nodes[v].right = 1;
            count[v] = 1;
        } else {
            uint32_t l = hier[v].left, r = hier[v].right;
            if (count[r] < count[l]) std::swap(l, r);
            Box bl{mk(nodes[l].center[0], nodes[l].center[1], nodes[l].center[2]),
                   mk(nodes[l].halfDim[0], nodes[l].halfDim[1], nodes[l].halfDim[2])};
            Box br{mk(nodes[r].center[0], nodes[r].center[1], nodes[r].center[2]),
                   mk(nodes[r].halfDim[0], nodes[r].halfDim[1], nodes[r].halfDim[2])};
            // GetBoxFromChildBoxes: FL/RayTracingHelper.hlsli:297-307
            Aabb a;
            a.mn = vmin(bl.center - bl.half, br.center - br.half);
            a.mx = vmax(bl.center + bl.half, br.center + br.half);
            Box b = aabb_to_box(a);
            nodes[v].center[0] = b.center.x, nodes[v].center[1] = b.center.y, nodes[v].center[2] = b.center.z;
            nodes[v].halfDim[0] = b.half.x, nodes[v].halfDim[1] = b.half.y, nodes[v].halfDim[2] = b.half.z;
            nodes[v].flags = l & 0x00ffffffu;
            nodes[v].right = r;
            count[v] = count[l] + count[r];
        }
    }
}

// PERFORM_UPDATE (FL/ComputeAABBs.hlsli:38-67): the topology is read back from the nodes already in the result buffer
// (children = {flags & 0xffffff, right}), only the boxes are re-fitted.  The subtree sizes are those of the original
// build and the children are already ordered, so the "smaller on the left" rule (:142-148) changes nothing except on
// ties, which are arrival-order dependent in the reference and pinned to "no swap" here exactly as in the full build.
template <class LeafBox>
static void refit_boxes(uint32_t n, rt_aabb_node *nodes, LeafBox leaf_box) {
    if (n == 0) return;
    const uint32_t nInternal = n - 1;
    std::vector<uint32_t> order, st;
    order.reserve(2 * size_t(n) - 1);
    st.push_back(0);
    while (!st.empty()) {
        uint32_t v = st.back();
        st.pop_back();
        order.push_back(v);
        if (v < nInternal) {
            st.push_back(nodes[v].flags & 0x00ffffffu);
            st.push_back(nodes[v].right);
        }
    }
    for (size_t i = order.size(); i-- > 0;) {
        const uint32_t v = order[i];
        Box b;
        if (v >= nInternal) {
            b = leaf_box(v - nInternal);
        } else {
            const uint32_t l = nodes[v].flags & 0x00ffffffu, r = nodes[v].right;
            Box bl{mk(nodes[l].center[0], nodes[l].center[1], nodes[l].center[2]),
                   mk(nodes[l].halfDim[0], nodes[l].halfDim[1], nodes[l].halfDim[2])};
            Box br{mk(nodes[r].center[0], nodes[r].center[1], nodes[r].center[2]),
                   mk(nodes[r].halfDim[0], nodes[r].halfDim[1], nodes[r].halfDim[2])};
            Aabb a;
            a.mn = vmin(bl.center - bl.half, br.center - br.half);
            a.mx = vmax(bl.center + bl.half, br.center + br.half);
            b = aabb_to_box(a);
        }
        nodes[v].center[0] = b.center.x, nodes[v].center[1] = b.center.y, nodes[v].center[2] = b.center.z;
        nodes[v].halfDim[0] = b.half.x, nodes[v].halfDim[1] = b.half.y, nodes[v].halfDim[2] = b.half.z;
    }
}

// What an ALLOW_UPDATE build leaves behind the blob (FL/GpuBVH2Builder.cpp:444-451): sort cache = load order ->
// sorted slot (FL/RearrangeTriangles.hlsl:25-28), parents = parent of every node (FL/ComputeAABBs.hlsli:160-164;
// the root's entry is never written there, pinned to 0).
static void make_update_cache(uint32_t n, const uint32_t *perm, const rt_aabb_node *nodes, std::vector<uint32_t> &sort_cache,
                              std::vector<uint32_t> &parents) {
    sort_cache.assign(n, 0);
    parents.assign(n ? 2 * size_t(n) - 1 : 0, 0);
    for (uint32_t i = 0; i < n; ++i) sort_cache[perm[i]] = i;
    for (uint32_t v = 0; v + 1 < n; ++v) {
        parents[nodes[v].flags & 0x00ffffffu] = v;
        parents[nodes[v].right] = v;
    }
}

static Box triangle_leaf_box(const rt_primitive &p) {
    const float *v = p.v;
    if (p.type == RT_PRIMITIVE_TYPE_PROCEDURAL) {
        // ComputeLeafAABB, procedural branch (FL/BottomLevelComputeAABBs.hlsl:32-40): no padding
        Aabb a{mk(v[0], v[1], v[2]), mk(v[3], v[4], v[5])};
        return aabb_to_box(a);
    }
    // GetBoxDataFromTriangle: FL/RayTracingHelper.hlsli:273-285
    f3 v0 = mk(v[0], v[1], v[2]), v1 = mk(v[3], v[4], v[5]), v2 = mk(v[6], v[7], v[8]);
    Aabb a;
    a.mn = vmin(vmin(v0, v1), v2);
    a.mx = vmax(vmax(v0, v1), v2);
    const float pad = 0.001f;  // AABB_Min_Padding
    a.mn = vmin(a.mn, a.mx - mk(pad, pad, pad));
    return aabb_to_box(a);
}

// FL/RayTracingHelper.hlsli:309-338.  Terms multiplied by the literal 0.0f in the reference
// vanish and the "* 1.0f" factors are exact, so only the surviving products are written,
// in the reference's order.
void invert_affine(const float t[12], float out[12]) {
#define T(r, c) t[(r)*4 + (c)]
    float det = T(0, 0) * T(1, 1) * T(2, 2) - T(0, 0) * T(2, 1) * T(1, 2) - T(1, 0) * T(0, 1) * T(2, 2) +
                T(1, 0) * T(2, 1) * T(0, 2) + T(2, 0) * T(0, 1) * T(1, 2) - T(2, 0) * T(1, 1) * T(0, 2);
    float invDet = 1.0f / det;
    float i00 = invDet * (T(1, 1) * T(2, 2) + T(2, 1) * (0.0f - T(1, 2)));
    float i10 = invDet * (T(1, 2) * T(2, 0) + T(2, 2) * (0.0f - T(1, 0)));
    float i20 = invDet * (T(1, 0) * T(2, 1) - T(2, 0) * T(1, 1));
    float i01 = invDet * (T(2, 1) * T(0, 2) + T(0, 1) * (0.0f - T(2, 2)));
    float i11 = invDet * (T(2, 2) * T(0, 0) + T(0, 2) * (0.0f - T(2, 0)));
    float i21 = invDet * (T(2, 0) * T(0, 1) - T(0, 0) * T(2, 1));
    float i02 = invDet * (T(0, 1) * T(1, 2) + T(1, 1) * (0.0f - T(0, 2)));
    float i12 = invDet * (T(0, 2) * T(1, 0) + T(1, 2) * (0.0f - T(0, 0)));
    float i22 = invDet * (T(0, 0) * T(1, 1) - T(1, 0) * T(0, 1));
    float i03 = invDet * (T(0, 1) * (T(2, 2) * T(1, 3) - T(1, 2) * T(2, 3)) + T(1, 1) * (T(0, 2) * T(2, 3) - T(2, 2) * T(0, 3)) +
                          T(2, 1) * (T(1, 2) * T(0, 3) - T(0, 2) * T(1, 3)));
    float i13 = invDet * (T(0, 2) * (T(2, 0) * T(1, 3) - T(1, 0) * T(2, 3)) + T(1, 2) * (T(0, 0) * T(2, 3) - T(2, 0) * T(0, 3)) +
                          T(2, 2) * (T(1, 0) * T(0, 3) - T(0, 0) * T(1, 3)));
    float i23 = invDet * (T(0, 3) * (T(2, 0) * T(1, 1) - T(1, 0) * T(2, 1)) + T(1, 3) * (T(0, 0) * T(2, 1) - T(2, 0) * T(0, 1)) +
                          T(2, 3) * (T(1, 0) * T(0, 1) - T(0, 0) * T(1, 1)));
#undef T
    out[0] = i00, out[1] = i01, out[2] = i02, out[3] = i03;
    out[4] = i10, out[5] = i11, out[6] = i12, out[7] = i13;
    out[8] = i20, out[9] = i21, out[10] = i22, out[11] = i23;
}

// FL/RayTracingHelper.hlsli:340-366
Aabb transform_aabb(Aabb box, const float m[12]) {
    f3 c[8] = {mk(box.mn.x, box.mn.y, box.mn.z), mk(box.mn.x, box.mn.y, box.mx.z), mk(box.mn.x, box.mx.y, box.mx.z),
               mk(box.mn.x, box.mx.y, box.mn.z), mk(box.mx.x, box.mn.y, box.mn.z), mk(box.mx.x, box.mx.y, box.mn.z),
               mk(box.mx.x, box.mn.y, box.mx.z), mk(box.mx.x, box.mx.y, box.mx.z)};
    Aabb out{mk(FLT_MAX, FLT_MAX, FLT_MAX), mk(-FLT_MAX, -FLT_MAX, -FLT_MAX)};
    for (int i = 0; i < 8; ++i) {
        f3 v = xform_point(m, c[i]);
        out.mn = vmin(out.mn, v);
        out.mx = vmax(out.mx, v);
    }
    return out;
}

}  // namespace orc

using namespace orc;

extern "C" {

void orc_scene_aabb(const rt_primitive *prims, uint32_t n, float out[6]) {
    f3 mn = mk(FLT_MAX, FLT_MAX, FLT_MAX), mx = mk(-FLT_MAX, -FLT_MAX, -FLT_MAX);
    for (uint32_t i = 0; i < n; ++i)
        // triangles: all three vertices; procedural: min and max (CalculateSceneAABBFromPrimitives.hlsl:26-37)
        for (int k = 0; k < (prims[i].type == RT_PRIMITIVE_TYPE_PROCEDURAL ? 2 : 3); ++k) {
            f3 v = mk(prims[i].v[3 * k], prims[i].v[3 * k + 1], prims[i].v[3 * k + 2]);
            mn = vmin(mn, v);
            mx = vmax(mx, v);
        }
    out[0] = mn.x, out[1] = mn.y, out[2] = mn.z, out[3] = mx.x, out[4] = mx.y, out[5] = mx.z;
}

void orc_morton_codes(const rt_primitive *prims, uint32_t n, const float aabb[6], uint32_t *codes) {
    for (uint32_t i = 0; i < n; ++i) {
        const float *v = prims[i].v;
        // GetCentroid: (tri.v0 + tri.v1 + tri.v2) / 3.0  (CalculateMortonCodesForPrimitives.hlsl:22-25);
        // procedural: (aabb.min + aabb.max) / 2.0 (:26-30)
        f3 c = prims[i].type == RT_PRIMITIVE_TYPE_PROCEDURAL
                   ? (mk(v[0], v[1], v[2]) + mk(v[3], v[4], v[5])) / 2.0f
                   : ((mk(v[0], v[1], v[2]) + mk(v[3], v[4], v[5])) + mk(v[6], v[7], v[8])) / 3.0f;
        codes[i] = morton_from_centroid(c, aabb);
    }
}

uint32_t orc_morton_code_from_centroid(const float c[3], const float aabb[6]) {
    return morton_from_centroid(mk(c[0], c[1], c[2]), aabb);
}

void orc_sort_pairs(const uint32_t *codes, uint32_t n, uint32_t *sorted_codes, uint32_t *perm) {
    std::vector<uint32_t> idx(n);
    std::iota(idx.begin(), idx.end(), 0u);
    std::stable_sort(idx.begin(), idx.end(), [&](uint32_t a, uint32_t b) { return codes[a] < codes[b]; });
    for (uint32_t i = 0; i < n; ++i) {
        perm[i] = idx[i];
        sorted_codes[i] = codes[idx[i]];
    }
}

void orc_build_hierarchy(const uint32_t *sorted_codes, uint32_t n, rt_hierarchy_node *nodes) {
    build_hierarchy(sorted_codes, n, nodes);
}

orc_blas *orc_blas_build(const rt_geometry_desc *geoms, uint32_t n_geoms, uint32_t build_flags) {
    orc_blas *b = new orc_blas();
    load_triangles(geoms, n_geoms, b->prims, b->meta);
    const uint32_t n = b->n = uint32_t(b->prims.size());
    if (n == 0) return b;
    orc_scene_aabb(b->prims.data(), n, b->aabb);
    b->morton.resize(n);
    b->sorted_morton.resize(n);
    b->perm.resize(n);
    orc_morton_codes(b->prims.data(), n, b->aabb, b->morton.data());
    orc_sort_pairs(b->morton.data(), n, b->sorted_morton.data(), b->perm.data());
    b->hier.resize(2 * size_t(n) - 1);
    build_hierarchy(b->sorted_morton.data(), n, b->hier.data());

    // result blob: [BVHOffsets][2n-1 AABBNode][n Primitive][n PrimitiveMetaData]
    // (FL/BottomLevelPrepareForComputeAABBs.hlsl:40-51, RayTracingHlslCompat.h:464-468)
    rt_bvh_offsets off;
    off.offsetToBoxes = 16;
    off.offsetToVertices = 16 + 32 * (2 * n - 1);
    off.offsetToPrimitiveMetaData = off.offsetToVertices + 40 * n;
    off.totalSize = off.offsetToPrimitiveMetaData + 12 * n;
    b->blob.assign(off.totalSize, 0);
    std::memcpy(b->blob.data(), &off, 16);
    rt_aabb_node *nodes = reinterpret_cast<rt_aabb_node *>(b->blob.data() + 16);
    rt_primitive *sp = reinterpret_cast<rt_primitive *>(b->blob.data() + off.offsetToVertices);
    rt_primitive_meta *sm = reinterpret_cast<rt_primitive_meta *>(b->blob.data() + off.offsetToPrimitiveMetaData);
    for (uint32_t i = 0; i < n; ++i) {  // FL/RearrangeTriangles.hlsl:18-36
        sp[i] = b->prims[b->perm[i]];
        sm[i] = b->meta[b->perm[i]];
    }
    // FL/GpuBVH2Builder.cpp:312-326: the treelet pass works on the hierarchy and the sorted triangles
    treelet_optimise(n, b->hier.data(), sp, build_flags);
    fit_boxes(n, b->hier.data(), nodes, [&](uint32_t slot) { return triangle_leaf_box(sp[slot]); });
    for (uint32_t slot = 0; slot < n; ++slot)  // flags.x = primitiveIndex | IsLeafFlag | IsProceduralGeometryFlag
        if (sp[slot].type == RT_PRIMITIVE_TYPE_PROCEDURAL) nodes[n - 1 + slot].flags |= RT_NODE_PROCEDURAL_FLAG;  // BottomLevelComputeAABBs.hlsl:34
    make_update_cache(n, b->perm.data(), nodes, b->sort_cache, b->parents);
    return b;
}

// PERFORM_UPDATE on a bottom-level structure (FL/GpuBVH2Builder.cpp:152-204): the load pass writes every triangle and
// its metadata straight to its cached sorted slot (FL/BottomLevelLoadTriangles.hlsli:107-112 via the sort cache), the
// hierarchy passes are skipped, ComputeAABBs re-fits on the stored topology.
int orc_blas_update(orc_blas *b, const rt_geometry_desc *geoms, uint32_t n_geoms) {
    std::vector<rt_primitive> prims;
    std::vector<rt_primitive_meta> meta;
    load_triangles(geoms, n_geoms, prims, meta);
    if (prims.size() != b->n) return -1;
    b->prims = prims;
    b->meta = meta;
    const rt_bvh_offsets *off = reinterpret_cast<const rt_bvh_offsets *>(b->blob.data());
    rt_aabb_node *nodes = reinterpret_cast<rt_aabb_node *>(b->blob.data() + 16);
    rt_primitive *sp = reinterpret_cast<rt_primitive *>(b->blob.data() + off->offsetToVertices);
    rt_primitive_meta *sm = reinterpret_cast<rt_primitive_meta *>(b->blob.data() + off->offsetToPrimitiveMetaData);
    for (uint32_t i = 0; i < b->n; ++i) {
        sp[b->sort_cache[i]] = prims[i];
        sm[b->sort_cache[i]] = meta[i];
    }
    refit_boxes(b->n, nodes, [&](uint32_t slot) { return triangle_leaf_box(sp[slot]); });
    return 0;
}
const uint32_t *orc_blas_sort_cache(const orc_blas *b) { return b->sort_cache.data(); }
const uint32_t *orc_blas_parents(const orc_blas *b) { return b->parents.data(); }

void orc_blas_free(orc_blas *b) { delete b; }
uint32_t orc_blas_num_prims(const orc_blas *b) { return b->n; }
const rt_primitive *orc_blas_unsorted_prims(const orc_blas *b) { return b->prims.data(); }
const float *orc_blas_scene_aabb(const orc_blas *b) { return b->aabb; }
const uint32_t *orc_blas_morton(const orc_blas *b) { return b->morton.data(); }
const uint32_t *orc_blas_sorted_morton(const orc_blas *b) { return b->sorted_morton.data(); }
const uint32_t *orc_blas_perm(const orc_blas *b) { return b->perm.data(); }
const rt_hierarchy_node *orc_blas_hierarchy(const orc_blas *b) { return b->hier.data(); }
const uint8_t *orc_blas_blob(const orc_blas *b, uint64_t *bytes) {
    if (bytes) *bytes = b->blob.size();
    return b->blob.data();
}

// FL/TopLevelLoadAABBs.hlsli:58-100: world box of the BLAS root box + BVHMetadata of every instance, load order.
static void load_instances(const rt_instance_desc *inst, uint32_t n, std::vector<orc::Box> &leaf, std::vector<rt_bvh_metadata> &md) {
    for (uint32_t i = 0; i < n; ++i) {
        const orc_blas *b = reinterpret_cast<const orc_blas *>(uintptr_t(inst[i].blas));
        const rt_aabb_node &root = b->nodes()[0];
        Aabb box = box_to_aabb(Box{mk(root.center[0], root.center[1], root.center[2]), mk(root.halfDim[0], root.halfDim[1], root.halfDim[2])});
        leaf[i] = aabb_to_box(transform_aabb(box, inst[i].transform));
        md[i].instanceDesc = inst[i];
        invert_affine(inst[i].transform, md[i].instanceDesc.transform);  // ObjectToWorld -> WorldToObject
        std::memcpy(md[i].objectToWorld, inst[i].transform, 48);
        md[i].instanceIndex = i;
    }
}

// TLAS: FL/TopLevelLoadAABBs.hlsli:58-100 -> CalculateSceneAABBFromBVHs.hlsl -> CalculateMortonCodesForAABBs.hlsl
// -> sort -> RearrangeBVHs.hlsl -> BuildBVHSplits -> TopLevelComputeAABBs.hlsl (no treelet pass).
orc_tlas *orc_tlas_build(const rt_instance_desc *inst, uint32_t n, uint32_t /*build_flags*/) {
    orc_tlas *t = new orc_tlas();
    t->n = n;
    // header: FL/TopLevelPrepareForComputeAABBs.hlsl:25-52
    const uint32_t nNodes = n == 0 ? 1 : 2 * n - 1;
    rt_bvh_offsets off;
    off.offsetToBoxes = 16;
    off.offsetToVertices = 16 + 32 * nNodes;  // offsetToLeafNodeMetaData
    off.offsetToPrimitiveMetaData = 0;
    off.totalSize = off.offsetToVertices + 116 * n;
    t->blob.assign(off.totalSize, 0);
    std::memcpy(t->blob.data(), &off, 16);
    if (n == 0) return t;  // empty AS: node 0 = zero box with zero flags

    std::vector<Box> leaf(n);
    std::vector<rt_bvh_metadata> md(n);
    load_instances(inst, n, leaf, md);
    // scene AABB from leaf boxes (re-derived from center/halfDim): CalculateSceneAABBFromBVHs.hlsl:16-40
    f3 mn = mk(FLT_MAX, FLT_MAX, FLT_MAX), mx = mk(-FLT_MAX, -FLT_MAX, -FLT_MAX);
    for (uint32_t i = 0; i < n; ++i) {
        Aabb a = box_to_aabb(leaf[i]);
        mn = vmin(a.mn, mn);
        mx = vmax(a.mx, mx);
    }
    float aabb[6] = {mn.x, mn.y, mn.z, mx.x, mx.y, mx.z};
    t->morton.resize(n);
    t->sorted_morton.resize(n);
    t->perm.resize(n);
    for (uint32_t i = 0; i < n; ++i) t->morton[i] = morton_from_centroid(leaf[i].center, aabb);
    orc_sort_pairs(t->morton.data(), n, t->sorted_morton.data(), t->perm.data());
    t->hier.resize(2 * size_t(n) - 1);
    build_hierarchy(t->sorted_morton.data(), n, t->hier.data());

    rt_aabb_node *nodes = reinterpret_cast<rt_aabb_node *>(t->blob.data() + 16);
    rt_bvh_metadata *smd = reinterpret_cast<rt_bvh_metadata *>(t->blob.data() + off.offsetToVertices);
    for (uint32_t i = 0; i < n; ++i) smd[i] = md[t->perm[i]];
    fit_boxes(n, t->hier.data(), nodes, [&](uint32_t slot) { return leaf[t->perm[slot]]; });
    make_update_cache(n, t->perm.data(), nodes, t->sort_cache, t->parents);
    return t;
}

// PERFORM_UPDATE on a top-level structure: instances are re-loaded to their cached sorted slots (the instance count
// and order of the descs must be those of the original build), then the boxes are re-fitted.
int orc_tlas_update(orc_tlas *t, const rt_instance_desc *inst, uint32_t n) {
    if (n != t->n) return -1;
    if (n == 0) return 0;
    std::vector<Box> leaf(n);
    std::vector<rt_bvh_metadata> md(n);
    load_instances(inst, n, leaf, md);
    const rt_bvh_offsets *off = reinterpret_cast<const rt_bvh_offsets *>(t->blob.data());
    rt_aabb_node *nodes = reinterpret_cast<rt_aabb_node *>(t->blob.data() + 16);
    rt_bvh_metadata *smd = reinterpret_cast<rt_bvh_metadata *>(t->blob.data() + off->offsetToVertices);
    for (uint32_t i = 0; i < n; ++i) smd[t->sort_cache[i]] = md[i];
    refit_boxes(n, nodes, [&](uint32_t slot) { return leaf[t->perm[slot]]; });
    return 0;
}
const uint32_t *orc_tlas_sort_cache(const orc_tlas *t) { return t->sort_cache.data(); }
const uint32_t *orc_tlas_parents(const orc_tlas *t) { return t->parents.data(); }

void orc_tlas_free(orc_tlas *t) { delete t; }
const uint8_t *orc_tlas_blob(const orc_tlas *t, uint64_t *bytes) {
    if (bytes) *bytes = t->blob.size();
    return t->blob.data();
}
const uint32_t *orc_tlas_sorted_morton(const orc_tlas *t) { return t->sorted_morton.data(); }
const uint32_t *orc_tlas_perm(const orc_tlas *t) { return t->perm.data(); }

}  // extern "C"
