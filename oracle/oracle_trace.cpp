// oracle_trace.cpp — CPU restatement of the Fallback Layer's two-level BVH2 traversal.
// TEST INFRASTRUCTURE (see oracle.h).  Follows FL/TraverseFunction.hlsli:173-282,438-460,520-799
// and FL/TraverseShader.hlsli:21-73 ("FL/" = externals/D3D12RaytracingFallback/src/).
#include <atomic>
#include <thread>

#include "oracle_internal.h"

namespace orc {

struct RayData {  // FL/TraverseFunction.hlsli:429-460
    f3 invDir, originTimesInvDir, shear;
    int kx, ky, kz;
};

static RayData get_ray_data(f3 o, f3 d) {
    RayData r;
    r.invDir = mk(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);  // rcp()
    r.originTimesInvDir = o * r.invDir;
    f3 a = vabs(d);
    int z = (a.x > a.y && a.x > a.z) ? 0 : (a.y > a.z ? 1 : 2);  // GetIndexOfBiggestChannel
    r.kx = (z + 1) % 3;
    r.ky = (z + 2) % 3;
    r.kz = z;
    if (d[r.kz] < 0.0f) std::swap(r.kx, r.ky);
    r.shear = mk(d[r.kx] / d[r.kz], d[r.ky] / d[r.kz], 1.0f / d[r.kz]);
    return r;
}

// RayBoxTest: FL/TraverseFunction.hlsli:173-191
static bool ray_box(float &resultT, float closestT, const RayData &rd, const rt_aabb_node &n) {
    f3 c = mk(n.center[0], n.center[1], n.center[2]), h = mk(n.halfDim[0], n.halfDim[1], n.halfDim[2]);
    // The HLSL expressions are not `precise`, so the shader compiler is free to contract them into
    // mads; they are pinned here as fused multiply-adds (what both DXC->driver and nvcc emit) so that
    // the CUDA traversal can be compared box test for box test.
    f3 ai = vabs(rd.invDir);
    f3 rel = mk(fmaf(c.x, rd.invDir.x, -rd.originTimesInvDir.x), fmaf(c.y, rd.invDir.y, -rd.originTimesInvDir.y),
                fmaf(c.z, rd.invDir.z, -rd.originTimesInvDir.z));
    f3 maxL = mk(fmaf(h.x, ai.x, rel.x), fmaf(h.y, ai.y, rel.y), fmaf(h.z, ai.z, rel.z));
    f3 minL = mk(fmaf(-h.x, ai.x, rel.x), fmaf(-h.y, ai.y, rel.y), fmaf(-h.z, ai.z, rel.z));
    float minT = fmaxf(fmaxf(minL.x, minL.y), minL.z);
    float maxT = fminf(fminf(maxL.x, maxL.y), maxL.z);
    resultT = fmaxf(minT, 0.0f);
    return fmaxf(minT, 0.0f) < fminf(maxT, closestT);
}

// RayTriangleIntersect: FL/TraverseFunction.hlsli:200-282 (Woop/Benthin/Wald 2013).
// Returns true and updates hitT/bary only on acceptance; U,V,W are "precise" (unfused).
static bool ray_triangle(float &hitT, float bary[2], uint32_t instanceFlags, uint32_t rayFlags, f3 o, const RayData &rd,
                         f3 v0, f3 v1, f3 v2) {
    bool useCulling = !(instanceFlags & RT_INSTANCE_FLAG_TRIANGLE_CULL_DISABLE);
    bool flipFaces = (instanceFlags & RT_INSTANCE_FLAG_TRIANGLE_FRONT_COUNTERCLOCKWISE) != 0;
    uint32_t backFlag = flipFaces ? RT_RAY_FLAG_CULL_FRONT_FACING_TRIANGLES : RT_RAY_FLAG_CULL_BACK_FACING_TRIANGLES;
    uint32_t frontFlag = flipFaces ? RT_RAY_FLAG_CULL_BACK_FACING_TRIANGLES : RT_RAY_FLAG_CULL_FRONT_FACING_TRIANGLES;
    bool cullBack = useCulling && (rayFlags & backFlag);
    bool cullFront = useCulling && (rayFlags & frontFlag);

    f3 a0 = v0 - o, b0 = v1 - o, c0 = v2 - o;
    float Ax = a0[rd.kx], Ay = a0[rd.ky], Az = a0[rd.kz];
    float Bx = b0[rd.kx], By = b0[rd.ky], Bz = b0[rd.kz];
    float Cx = c0[rd.kx], Cy = c0[rd.ky], Cz = c0[rd.kz];
    Ax = Ax - rd.shear.x * Az, Ay = Ay - rd.shear.y * Az;
    Bx = Bx - rd.shear.x * Bz, By = By - rd.shear.y * Bz;
    Cx = Cx - rd.shear.x * Cz, Cy = Cy - rd.shear.y * Cz;
    float U = Cx * By - Cy * Bx;
    float V = Ax * Cy - Ay * Cx;
    float W = Bx * Ay - By * Ax;
    float det = (U + V) + W;
    if (cullFront) {
        if (U > 0.0f || V > 0.0f || W > 0.0f) return false;
    } else if (cullBack) {
        if (U < 0.0f || V < 0.0f || W < 0.0f) return false;
    } else {
        if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return false;
    }
    if (det == 0.0f) return false;
    Az = rd.shear.z * Az;
    Bz = rd.shear.z * Bz;
    Cz = rd.shear.z * Cz;
    const float T = (U * Az + V * Bz) + W * Cz;
    if (cullFront) {
        if (T > 0.0f || T < hitT * det) return false;
    } else if (cullBack) {
        if (T < 0.0f || T > hitT * det) return false;
    } else {
        float s = fabsf(T);
        if ((T > 0.0f) != (det > 0.0f)) s = -s;
        if (s < 0.0f || s > hitT * fabsf(det)) return false;
    }
    const float rcpDet = 1.0f / det;
    bary[0] = V * rcpDet;
    bary[1] = W * rcpDet;
    hitT = T * rcpDet;
    return true;
}

// ---------------------------------------------------------------- compiled-in any-hit / intersection programs
// The reference links application HLSL here (Fallback_CallIndirect(stateId)); the programs below are OURS (the
// application has only the no-op ShadowAnyHit and no intersection shader), so they are "parity unpinned" — what IS
// restated from the reference is the machinery around them: InvokeAnyHit / IgnoreHit / AcceptHitAndEndSearch
// (FL/TraverseFunction.hlsli:102-117), Fallback_ReportHit (:136-158) and the leaf handling (:635-735).
enum { kEndSearch = -1, kIgnore = 0, kAccept = 1 };  // :11-13

static int run_any_hit(uint32_t program, float ax, float ay) {
    // InvokeAnyHit: Fallback_SetAnyHitResult(ACCEPT); call; return Fallback_AnyHitResult()
    switch (program) {
        case RT_ANYHIT_IGNORE: return kIgnore;
        case RT_ANYHIT_END_SEARCH: return kEndSearch;
        case RT_ANYHIT_CUTOUT: return ((int(floorf(8.0f * ax)) + int(floorf(8.0f * ay))) & 1) ? kIgnore : kAccept;
        default: return kAccept;  // RT_ANYHIT_ACCEPT: no-op body
    }
}

struct ReportCtx {  // the Fallback_* registers ReportHit touches
    float tmin;
    float *tCurrent;
    uint32_t rayFlags, anyHit;
    int anyHitResult;  // Fallback_AnyHitResult()
    bool committed;
    float t, ax, ay;
    uint32_t kind;
};

// Fallback_ReportHit: FL/TraverseFunction.hlsli:136-158.  geomOpaque is the literal `true` and the instance flags the
// literal 0 there ("TODO" in the reference), so an any-hit shader runs for a procedural hit only under
// RAY_FLAG_FORCE_NON_OPAQUE.
static int report_hit(ReportCtx &c, float tHit, uint32_t hitKind, float ax, float ay) {
    if (tHit < c.tmin || *c.tCurrent <= tHit) return 0;
    int ret = kAccept;
    bool opaque = true;
    if (c.rayFlags & RT_RAY_FLAG_FORCE_OPAQUE) opaque = true;
    else if (c.rayFlags & RT_RAY_FLAG_FORCE_NON_OPAQUE) opaque = false;
    if (c.anyHit > 0 && !opaque) ret = c.anyHitResult = run_any_hit(c.anyHit, ax, ay);
    if (ret != kIgnore) {
        *c.tCurrent = tHit;  // Fallback_CommitHit
        c.committed = true;
        c.t = tHit, c.ax = ax, c.ay = ay, c.kind = hitKind;
        if (c.rayFlags & RT_RAY_FLAG_ACCEPT_FIRST_HIT_AND_END_SEARCH) ret = kEndSearch;
    }
    return ret;
}

static inline float dot3(f3 a, f3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }

// RT_INTERSECTION_BOX: slab test of the object-space ray against the primitive's own AABB.
static void intersect_box(ReportCtx &c, f3 o, f3 d, f3 mn, f3 mx) {
    f3 t0 = (mn - o) / d, t1 = (mx - o) / d;
    float tNear = fmaxf(fmaxf(fminf(t0.x, t1.x), fminf(t0.y, t1.y)), fminf(t0.z, t1.z));
    float tFar = fminf(fminf(fmaxf(t0.x, t1.x), fmaxf(t0.y, t1.y)), fmaxf(t0.z, t1.z));
    if (!(tNear <= tFar)) return;
    if (report_hit(c, tNear, RT_HIT_KIND_BOX_ENTER, 0.0f, 0.0f) == 0) report_hit(c, tFar, RT_HIT_KIND_BOX_EXIT, 0.0f, 0.0f);
}

// RT_INTERSECTION_SPHERE: the sphere inscribed in the AABB; attributes = (normal.x, normal.y).
static void intersect_sphere(ReportCtx &c, f3 o, f3 d, f3 mn, f3 mx) {
    f3 ctr = (mn + mx) * 0.5f, h = mx - ctr;
    float r = fminf(fminf(h.x, h.y), h.z);
    if (!(r > 0.0f)) return;
    f3 oc = o - ctr;
    float a = dot3(d, d), b = dot3(oc, d), cc = dot3(oc, oc) - r * r;
    float disc = b * b - a * cc;
    if (!(disc >= 0.0f) || !(a > 0.0f)) return;
    float s = sqrtf(disc);
    float tA = (-b - s) / a, tB = (-b + s) / a;
    f3 nA = (oc + d * tA) / r, nB = (oc + d * tB) / r;
    if (report_hit(c, tA, RT_HIT_KIND_SPHERE_ENTER, nA.x, nA.y) == 0) report_hit(c, tB, RT_HIT_KIND_SPHERE_EXIT, nB.x, nB.y);
}

HitInfo trace_ray(const orc_tlas *tl, f3 origin, float tmin, f3 dir, float tmax, uint32_t rayFlags, uint32_t mask,
                  uint32_t rayContribution, uint32_t geomMultiplier, TraceCounters *ctr, const rt_hit_group_programs *programs,
                  uint32_t n_programs) {
    HitInfo hit;
    float tCurrent = tmax;  // Fallback_TraceRayBegin: RayTCurrent() = TMax
    if (tl->n == 0) return hit;

    // The reference stack holds TRAVERSAL_MAX_STACK_DEPTH = 32 entries and overflows silently
    // (FL/TraverseFunction.hlsli:15-16).  The oracle uses a growable stack and records the depth.
    std::vector<uint32_t> stack;
    stack.reserve(64);
    const rt_aabb_node *tnodes = tl->nodes();
    const rt_bvh_metadata *tmeta = tl->metadata();

    RayData world = get_ray_data(origin, dir);
    RayData cur = world;
    f3 curOrigin = origin, curDir = dir;
    bool bottom = false;
    uint32_t nodesToProcess[2] = {0, 0};
    const rt_aabb_node *nodes = tnodes;
    const orc_blas *blas = nullptr;
    uint32_t instanceIndex = 0, instanceFlags = 0, instanceOffset = 0, instanceId = 0;
    bool endSearch = false;

    float unusedT;
    if (ray_box(unusedT, tCurrent, world, tnodes[0])) {
        stack.push_back(0);
        nodesToProcess[0]++;
    }
    while (nodesToProcess[0] != 0) {
        do {
            uint32_t nodeIndex = stack.back();
            stack.pop_back();
            nodesToProcess[bottom]--;
            const rt_aabb_node &node = nodes[nodeIndex];
            if (node.flags & RT_NODE_LEAF_FLAG) {
                uint32_t leafIndex = node.flags & ~(RT_NODE_LEAF_FLAG | RT_NODE_PROCEDURAL_FLAG);
                if (!bottom) {
                    if (ctr) ctr->inst++;
                    const rt_bvh_metadata &md = tmeta[leafIndex];
                    instanceIndex = md.instanceIndex;
                    instanceOffset = RT_INSTANCE_HIT_GROUP(md.instanceDesc);
                    instanceId = RT_INSTANCE_ID(md.instanceDesc);
                    if (RT_INSTANCE_MASK(md.instanceDesc) & mask) {
                        bottom = true;
                        stack.push_back(0);
                        blas = reinterpret_cast<const orc_blas *>(uintptr_t(md.instanceDesc.blas));
                        nodes = blas->nodes();
                        instanceFlags = RT_INSTANCE_FLAGS(md.instanceDesc);
                        curOrigin = xform_point(md.instanceDesc.transform, origin);
                        f3 objDir = xform_vector(md.instanceDesc.transform, dir);
                        curDir = objDir;
                        cur = get_ray_data(curOrigin, objDir);
                        nodesToProcess[1] = 1;
                    }
                } else {
                    if (ctr) ctr->leaf++;
                    const rt_primitive_meta &pm = blas->sorted_meta()[leafIndex];
                    bool geomOpaque = (pm.geometryFlags & RT_GEOMETRY_FLAG_OPAQUE) != 0;
                    bool opaque = geomOpaque;  // IsOpaque(): FL/TraverseFunction.hlsli:119-134
                    if (instanceFlags & RT_INSTANCE_FLAG_FORCE_OPAQUE) opaque = true;
                    else if (instanceFlags & RT_INSTANCE_FLAG_FORCE_NON_OPAQUE) opaque = false;
                    if (rayFlags & RT_RAY_FLAG_FORCE_OPAQUE) opaque = true;
                    else if (rayFlags & RT_RAY_FLAG_FORCE_NON_OPAQUE) opaque = false;
                    bool culled = (opaque && (rayFlags & RT_RAY_FLAG_CULL_OPAQUE)) || (!opaque && (rayFlags & RT_RAY_FLAG_CULL_NON_OPAQUE));
                    const uint32_t record = rayContribution + pm.geometryContributionToHitGroupIndex * geomMultiplier + instanceOffset;
                    const rt_hit_group_programs none = {RT_ANYHIT_NONE, RT_INTERSECTION_NONE};
                    const rt_hit_group_programs &prog = (programs && record < n_programs) ? programs[record] : none;
                    auto commit = [&](float t, float a0, float a1, uint32_t kind) {
                        tCurrent = t;
                        hit.hit = true;
                        hit.t = t;
                        hit.bary[0] = a0, hit.bary[1] = a1;
                        hit.primitiveIndex = pm.primitiveIndex;
                        hit.geometryIndex = pm.geometryContributionToHitGroupIndex;
                        hit.instanceIndex = instanceIndex;
                        hit.instanceId = instanceId;
                        hit.leafSlot = leafIndex;
                        hit.hitGroupContribution = record;
                        hit.hitKind = kind;
                    };
                    const float *v = blas->sorted_prims()[leafIndex].v;
                    if (!culled && (node.flags & RT_NODE_PROCEDURAL_FLAG)) {
                        // :656-671 — the intersection shader of the hit group runs with the object-space ray; it
                        // calls ReportHit; the search ends only if an any-hit shader said AcceptHitAndEndSearch
                        // (the END_SEARCH that ReportHit returns under ACCEPT_FIRST_HIT is not looked at, :670).
                        ReportCtx rc{tmin, &tCurrent, rayFlags, prog.any_hit, kAccept, false, 0, 0, 0, 0};
                        if (prog.intersection == RT_INTERSECTION_BOX)
                            intersect_box(rc, curOrigin, curDir, mk(v[0], v[1], v[2]), mk(v[3], v[4], v[5]));
                        else if (prog.intersection == RT_INTERSECTION_SPHERE)
                            intersect_sphere(rc, curOrigin, curDir, mk(v[0], v[1], v[2]), mk(v[3], v[4], v[5]));
                        if (rc.committed) commit(rc.t, rc.ax, rc.ay, rc.kind);
                        endSearch = rc.anyHitResult == kEndSearch;
                    } else if (!culled) {
                        float t0 = tCurrent, bary[2];
                        bool ok = ray_triangle(t0, bary, instanceFlags, rayFlags, curOrigin, cur, mk(v[0], v[1], v[2]),
                                               mk(v[3], v[4], v[5]), mk(v[6], v[7], v[8]));
                        // TestLeafNodeIntersections :385
                        if (ok && t0 < tCurrent && t0 > tmin) {
                            // :699-722.  Opaque: commit.  Non-opaque: the hit group's any-hit shader (if any) decides;
                            // the application registers only the no-op ShadowAnyHit and all its geometry is OPAQUE
                            // (libs/DXRFramework/Helpers/BottomLevelASGenerator.cpp:109-110).  Literal quirk kept:
                            // under ACCEPT_FIRST_HIT an IGNOREd candidate still ends the search (:721).
                            int ret = kAccept;
                            if (!opaque && prog.any_hit) ret = run_any_hit(prog.any_hit, bary[0], bary[1]);
                            if (ret != kIgnore) commit(t0, bary[0], bary[1], RT_HIT_KIND_TRIANGLE_FRONT_FACE);
                            endSearch = ret == kEndSearch || (rayFlags & RT_RAY_FLAG_ACCEPT_FIRST_HIT_AND_END_SEARCH);
                        }
                    }
                    if (endSearch) {
                        nodesToProcess[1] = 0;
                        nodesToProcess[0] = 0;
                    }
                }
            } else {
                if (ctr) ctr->internal++;
                uint32_t l = node.flags & 0x00ffffffu, r = node.right;
                float lt, rt;
                bool lh = ray_box(lt, tCurrent, cur, nodes[l]);
                bool rh = ray_box(rt, tCurrent, cur, nodes[r]);
                if (lh && rh) {
                    bool rightFirst = rt < lt;  // ties: left first (:776-778)
                    // StackPush2(selector, A=left, B=right): store0 = selector ? A : B; popped last-in first
                    stack.push_back(rightFirst ? l : r);
                    stack.push_back(rightFirst ? r : l);
                    nodesToProcess[bottom] += 2;
                } else if (lh || rh) {
                    stack.push_back(rh ? r : l);
                    nodesToProcess[bottom] += 1;
                }
            }
            if (ctr && stack.size() > ctr->max_stack) ctr->max_stack = stack.size();
        } while (nodesToProcess[bottom] != 0);
        bottom = false;
        cur = world;
        curOrigin = origin;
        curDir = dir;
        nodes = tnodes;
    }
    return hit;
}

}  // namespace orc

using namespace orc;

template <class F>
static void parallel_for(uint64_t n, int threads, F f) {
    if (threads <= 1 || n < 2) {
        f(0, n, 0);
        return;
    }
    std::vector<std::thread> pool;
    std::atomic<uint64_t> next{0};
    const uint64_t chunk = std::max<uint64_t>(1, std::min<uint64_t>(4096, n / (uint64_t(threads) * 8) + 1));
    for (int t = 0; t < threads; ++t)
        pool.emplace_back([&, t] {
            for (;;) {
                uint64_t b = next.fetch_add(chunk);
                if (b >= n) break;
                f(b, std::min(n, b + chunk), t);
            }
        });
    for (auto &th : pool) th.join();
}

extern "C" void orc_trace_hit_groups(const orc_tlas *t, const rt_ray *rays, uint64_t n, uint32_t ray_flags, uint32_t instance_mask,
                                     uint32_t ray_contribution, uint32_t geometry_multiplier,
                                     const rt_hit_group_programs *programs, uint32_t n_programs, rt_hit *hits, int threads) {
    parallel_for(n, threads, [&](uint64_t b, uint64_t e, int) {
        for (uint64_t i = b; i < e; ++i) {
            const rt_ray &r = rays[i];
            HitInfo h = trace_ray(t, mk(r.origin[0], r.origin[1], r.origin[2]), r.tmin, mk(r.direction[0], r.direction[1], r.direction[2]),
                                  r.tmax, ray_flags, instance_mask, ray_contribution, geometry_multiplier, nullptr, programs, n_programs);
            rt_hit &o = hits[i];
            o.t = h.hit ? h.t : r.tmax;
            o.bary[0] = h.bary[0], o.bary[1] = h.bary[1];
            o.primitive_index = h.hit ? h.primitiveIndex : RT_NO_HIT;
            o.instance_index = h.instanceIndex;
            o.geometry_index = h.geometryIndex;
            o.instance_id = h.instanceId;
            o.leaf_slot = h.leafSlot | (h.hit ? h.hitKind << 24 : 0u);  // HitKind() in bits 31:24
        }
    });
}

extern "C" void orc_trace(const orc_tlas *t, const rt_ray *rays, uint64_t n, uint32_t ray_flags, uint32_t instance_mask,
                          rt_hit *hits, rt_trace_stats *stats, int threads) {
    int nt = threads <= 1 ? 1 : threads;
    std::vector<TraceCounters> ctrs(nt);
    parallel_for(n, threads, [&](uint64_t b, uint64_t e, int tid) {
        for (uint64_t i = b; i < e; ++i) {
            const rt_ray &r = rays[i];
            HitInfo h = trace_ray(t, mk(r.origin[0], r.origin[1], r.origin[2]), r.tmin, mk(r.direction[0], r.direction[1], r.direction[2]),
                                  r.tmax, ray_flags, instance_mask, 0, 0, &ctrs[tid]);
            rt_hit &o = hits[i];
            o.t = h.hit ? h.t : r.tmax;
            o.bary[0] = h.bary[0], o.bary[1] = h.bary[1];
            o.primitive_index = h.hit ? h.primitiveIndex : RT_NO_HIT;
            o.instance_index = h.instanceIndex;
            o.geometry_index = h.geometryIndex;
            o.instance_id = h.instanceId;
            o.leaf_slot = h.leafSlot;
        }
    });
    if (stats) {
        stats->rays += n;
        for (auto &c : ctrs) {
            stats->internal_visits += c.internal;
            stats->leaf_visits += c.leaf;
            stats->instance_visits += c.inst;
            if (c.max_stack > stats->max_stack) stats->max_stack = c.max_stack;
        }
    }
}
