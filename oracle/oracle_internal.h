// oracle_internal.h — shared internals of the CPU oracle (TEST INFRASTRUCTURE, see oracle.h).
#pragma once
#include "oracle.h"

#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace orc {

// IEEE fp32 -> fp16 (round to nearest even) -> fp32: what a store to an R16G16B16A16_FLOAT target followed by a load returns.
inline float half_round(float f) {
    uint32_t x;
    std::memcpy(&x, &f, 4);
    const uint32_t sign = x & 0x80000000u;
    const uint32_t e = (x >> 23) & 0xFFu;
    uint32_t man = x & 0x7FFFFFu;
    if (e == 0xFFu) return f;  // Inf / NaN survive
    const int32_t exp = int32_t(e) - 127 + 15;
    uint16_t h;
    if (exp >= 31) h = 0x7C00u;
    else if (exp <= 0) {
        if (exp < -10) h = 0;
        else {
            man |= 0x800000u;
            const uint32_t shift = uint32_t(14 - exp);
            uint32_t v = man >> shift;
            const uint32_t rem = man & ((1u << shift) - 1u), halfway = 1u << (shift - 1);
            if (rem > halfway || (rem == halfway && (v & 1u))) ++v;
            h = uint16_t(v);
        }
    } else {
        uint32_t v = (uint32_t(exp) << 10) | (man >> 13);
        const uint32_t rem = man & 0x1FFFu;
        if (rem > 0x1000u || (rem == 0x1000u && (v & 1u))) ++v;
        h = uint16_t(v);
    }
    // back to fp32
    const uint32_t he = (h >> 10) & 0x1Fu, hm = h & 0x3FFu;
    uint32_t bits;
    if (he == 0) {
        if (hm == 0) bits = sign;
        else {
            float v = float(hm) * 5.9604644775390625e-08f;  // 2^-24
            std::memcpy(&bits, &v, 4);
            bits |= sign;
        }
    } else if (he == 31) bits = sign | 0x7F800000u;
    else bits = sign | ((he - 15 + 127) << 23) | (hm << 13);
    float r;
    std::memcpy(&r, &bits, 4);
    return r;
}
float store_value(float v);  // oracle_shade.cpp: rounds through fp16 when half render targets are emulated


struct f3 {
    float x, y, z;
    float &operator[](int i) { return (&x)[i]; }
    float operator[](int i) const { return (&x)[i]; }
};
static inline f3 mk(float x, float y, float z) { return f3{x, y, z}; }
static inline f3 operator+(f3 a, f3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline f3 operator-(f3 a, f3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline f3 operator*(f3 a, f3 b) { return mk(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline f3 operator*(f3 a, float s) { return mk(a.x * s, a.y * s, a.z * s); }
static inline f3 operator*(float s, f3 a) { return mk(s * a.x, s * a.y, s * a.z); }
static inline f3 operator/(f3 a, float s) { return mk(a.x / s, a.y / s, a.z / s); }
static inline f3 operator/(f3 a, f3 b) { return mk(a.x / b.x, a.y / b.y, a.z / b.z); }
static inline f3 operator-(f3 a) { return mk(-a.x, -a.y, -a.z); }
// HLSL min/max return the non-NaN operand; fminf/fmaxf have the same rule.
static inline f3 vmin(f3 a, f3 b) { return mk(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z)); }
static inline f3 vmax(f3 a, f3 b) { return mk(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z)); }
static inline f3 vabs(f3 a) { return mk(fabsf(a.x), fabsf(a.y), fabsf(a.z)); }
// dot(): left-to-right mul/add, never fused (the library is built with -ffp-contract=off).
static inline float dot(f3 a, f3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
static inline f3 cross(f3 a, f3 b) {
    return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
static inline float length(f3 a) { return sqrtf(dot(a, a)); }
// HLSL normalize(v) = v * rsqrt(dot(v,v)); restated with IEEE sqrt + divide (DESIGN.md "Float semantics").
static inline f3 normalize(f3 a) {
    float inv = 1.0f / sqrtf(dot(a, a));
    return a * inv;
}
static inline float saturate(float x) { return fminf(fmaxf(x, 0.0f), 1.0f); }

// 3x4 affine transform applied to a point (w=1) or a vector (w=0): mul(M, float4(v,w)).
// Pinned evaluation order: ((m0*x + m1*y) + m2*z) + m3*w.
static inline f3 xform_point(const float m[12], f3 v) {
    return mk(((m[0] * v.x + m[1] * v.y) + m[2] * v.z) + m[3], ((m[4] * v.x + m[5] * v.y) + m[6] * v.z) + m[7],
              ((m[8] * v.x + m[9] * v.y) + m[10] * v.z) + m[11]);
}
static inline f3 xform_vector(const float m[12], f3 v) {
    return mk((m[0] * v.x + m[1] * v.y) + m[2] * v.z, (m[4] * v.x + m[5] * v.y) + m[6] * v.z,
              (m[8] * v.x + m[9] * v.y) + m[10] * v.z);
}

struct Box {
    f3 center, half;
};
struct Aabb {
    f3 mn, mx;
};

// FL/RayTracingHelper.hlsli:251-265
static inline Box aabb_to_box(Aabb a) {
    Box b;
    b.center = (a.mn + a.mx) * 0.5f;
    b.half = a.mx - b.center;
    return b;
}
static inline Aabb box_to_aabb(Box b) { return Aabb{b.center - b.half, b.center + b.half}; }

void invert_affine(const float t[12], float out[12]);
// FL/TreeletReorder.cpp:38-109 (oracle_treelet.cpp)
uint32_t treelet_pass_count(uint32_t build_flags);
void treelet_optimise(uint32_t n, rt_hierarchy_node *hier, const rt_primitive *sorted_prims, uint32_t build_flags);
Aabb transform_aabb(Aabb box, const float m[12]);

}  // namespace orc

struct orc_blas {
    uint32_t n = 0;
    std::vector<rt_primitive> prims;        // load order
    std::vector<rt_primitive_meta> meta;    // load order
    float aabb[6] = {0, 0, 0, 0, 0, 0};
    std::vector<uint32_t> morton, sorted_morton, perm;
    std::vector<rt_hierarchy_node> hier;
    std::vector<uint8_t> blob;
    std::vector<uint32_t> sort_cache, parents;  // what ALLOW_UPDATE appends (FL/GpuBVH2Builder.cpp:444-451)
    // convenience views into blob
    const rt_aabb_node *nodes() const { return reinterpret_cast<const rt_aabb_node *>(blob.data() + 16); }
    const rt_primitive *sorted_prims() const {
        return reinterpret_cast<const rt_primitive *>(blob.data() + reinterpret_cast<const rt_bvh_offsets *>(blob.data())->offsetToVertices);
    }
    const rt_primitive_meta *sorted_meta() const {
        return reinterpret_cast<const rt_primitive_meta *>(blob.data() + reinterpret_cast<const rt_bvh_offsets *>(blob.data())->offsetToPrimitiveMetaData);
    }
};

struct orc_tlas {
    uint32_t n = 0;
    std::vector<uint32_t> morton, sorted_morton, perm;
    std::vector<rt_hierarchy_node> hier;
    std::vector<uint8_t> blob;
    std::vector<uint32_t> sort_cache, parents;
    const rt_aabb_node *nodes() const { return reinterpret_cast<const rt_aabb_node *>(blob.data() + 16); }
    const rt_bvh_metadata *metadata() const {
        return reinterpret_cast<const rt_bvh_metadata *>(blob.data() + reinterpret_cast<const rt_bvh_offsets *>(blob.data())->offsetToVertices);
    }
};

namespace orc {

struct HitInfo {
    bool hit = false;
    float t = 0;
    float bary[2] = {0, 0};
    uint32_t primitiveIndex = RT_NO_HIT, instanceIndex = 0, geometryIndex = 0, instanceId = 0, leafSlot = 0;
    uint32_t hitGroupContribution = 0;  // instanceContribution + rayContribution + geom*multiplier
    uint32_t hitKind = 0;               // HitKind(): 0xFE for triangles, the intersection program's value otherwise
};

struct TraceCounters {
    uint64_t internal = 0, leaf = 0, inst = 0, max_stack = 0;
};

// Fallback_TraceRay without the shader call-out: FL/TraverseShader.hlsli:21-73.
HitInfo trace_ray(const orc_tlas *t, f3 origin, float tmin, f3 dir, float tmax, uint32_t rayFlags, uint32_t mask,
                  uint32_t rayContribution, uint32_t geomMultiplier, TraceCounters *ctr,
                  const rt_hit_group_programs *programs = nullptr, uint32_t n_programs = 0);

}  // namespace orc
