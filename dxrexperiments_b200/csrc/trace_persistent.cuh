// trace_persistent.cuh — persistent-warp queue traversal with dynamic ray replacement.
//
// The plain one-thread-one-ray loop (trace.cuh) runs incoherent rays at ~5 of 32 lanes active: rays in a warp
// finish after wildly different numbers of node visits and the warp waits for its longest ray
// (profiles/r1_ncu_trace_baseline.md).  This kernel keeps the same traversal semantics and ORDER (so hits,
// ties and visit counts are those of the reference restatement) but schedules the work differently:
//   * persistent warps pull rays from a global counter; when at least kFetchThreshold lanes are idle the warp
//     refills exactly those lanes (Aila & Laine's "dynamic fetch", warp-synchronous via __ballot_sync);
//   * internal-node steps and leaf steps are separate warp-wide phases: a lane that reaches a leaf parks until
//     kLeafThreshold lanes hold one (or nobody can advance), so the triangle test never runs for 1-2 lanes;
//   * node fetches are four 16-byte loads, triangle fetches three.
#pragma once
#include "trace.cuh"

constexpr int kFetchThreshold = 8;   // refill when >= this many lanes are idle
constexpr int kLeafThreshold = 8;    // run a leaf phase when >= this many lanes hold a leaf

template <bool ANY>
struct QueueSink {  // where results of the ray queue go
    float4 *hitA;
    uint32_t *hitRec;
    uint8_t *vis;
    __device__ __forceinline__ void miss_inactive(uint32_t i) const {
        if (ANY) vis[i] = 1;
        else hitA[i] = make_float4(0, 0, 0, __uint_as_float(RT_NO_HIT)), hitRec[i] = 0xffffffffu;
    }
};

template <bool ANY>
__global__ void __launch_bounds__(128) k_trace_persistent(const void *tlas, const rt_ray *rays, const uint32_t *count, uint32_t mult,
                                                          QueueSink<ANY> sink, uint32_t *status, uint32_t *nextRay) {
    const uint32_t n = count[0] * mult;
    const TraceAccel A = resolve_tlas(tlas);
    const int lane = threadIdx.x & 31;
    const unsigned ltMask = (1u << lane) - 1u;
    constexpr unsigned FULL = 0xffffffffu;
    constexpr uint32_t rayFlags = ANY ? (RT_RAY_FLAG_ACCEPT_FIRST_HIT_AND_END_SEARCH | RT_RAY_FLAG_SKIP_CLOSEST_HIT_SHADER) : 0u;
    constexpr uint32_t rayContribution = ANY ? 1u : 0u;  // shadow rays use hit group 1 (S/RaytracingCommon.hlsli:94)

    uint32_t stack[RT_STACK_SIZE];
    // per-lane ray state
    bool alive = false;
    uint32_t rayIdx = 0, ref = RT_SENTINEL;
    int sp = 0, blasBase = -1;
    bool bottom = false;
    float wox = 0, woy = 0, woz = 0, wdx = 0, wdy = 0, wdz = 1, tmin = 0, tCur = 0;
    RayPre cur;
    cur.ox = cur.oy = cur.oz = cur.ix = cur.iy = cur.iz = cur.oix = cur.oiy = cur.oiz = cur.sx = cur.sy = cur.sz = 0;
    cur.kx = 0, cur.ky = 1, cur.kz = 2;
    const rt_wide_node *nodes = A.wide;
    const rt_packed_tri *tris = nullptr;
    uint32_t instFlags = 0, instOffset = 0;
    int cull = 0;
    float hu = 0, hv = 0;
    uint32_t hprim = RT_NO_HIT, hrec = 0;
    bool noMore = (A.count == 0 && n == 0);

    auto finish = [&]() {  // write the result of a finished ray
        if (ANY) sink.vis[rayIdx] = (hprim != RT_NO_HIT) ? 0 : 1;
        else {
            sink.hitA[rayIdx] = make_float4(tCur, hu, hv, __uint_as_float(hprim));
            sink.hitRec[rayIdx] = hrec;
        }
        alive = false;
    };

    for (;;) {
        // ------------------------------------------------------------------ fetch
        const unsigned aliveMask = __ballot_sync(FULL, alive);
        const int idle = 32 - __popc(aliveMask);
        if (!noMore && (idle >= kFetchThreshold)) {
            const unsigned dead = ~aliveMask;
            const int leader = __ffs(dead) - 1;
            uint32_t base = 0;
            if (lane == leader) base = atomicAdd(nextRay, uint32_t(idle));
            base = __shfl_sync(FULL, base, leader);
            noMore = base + uint32_t(idle) >= n;
            if (!alive) {
                rayIdx = base + __popc(dead & ltMask);
                if (rayIdx < n) {
                    const float4 *rp = reinterpret_cast<const float4 *>(rays + rayIdx);
                    const float4 a = __ldcs(rp), b = __ldcs(rp + 1);
                    if (b.w < 0.0f || A.count == 0) {
                        sink.miss_inactive(rayIdx);
                    } else {
                        wox = a.x, woy = a.y, woz = a.z, tmin = a.w, wdx = b.x, wdy = b.y, wdz = b.z, tCur = b.w;
                        hprim = RT_NO_HIT, hu = hv = 0.0f, hrec = 0;
                        cur = make_ray_pre(wox, woy, woz, wdx, wdy, wdz);
                        nodes = A.wide;
                        bottom = false, blasBase = -1, sp = 0;
                        float tU;
                        alive = true;
                        if (ray_box(tU, tCur, cur, A.root_c[0], A.root_c[1], A.root_c[2], A.root_h[0], A.root_h[1], A.root_h[2]))
                            ref = A.root_ref;
                        else
                            finish();
                    }
                }
            }
            continue;  // re-evaluate occupancy (lanes that drew inactive rays refill again)
        }
        if (aliveMask == 0) {
            if (noMore) break;
            continue;
        }
        // ------------------------------------------------------------------ phase selection
        const bool atLeaf = alive && (ref & RT_NODE_LEAF_FLAG);
        const unsigned leafMask = __ballot_sync(FULL, atLeaf);
        const unsigned intMask = aliveMask & ~leafMask;
        if (intMask == 0 || __popc(leafMask) >= kLeafThreshold) {
            // ---------------------------------------------------------------- leaf phase
            if (atLeaf) {
                const uint32_t slot = ref & 0x00ffffffu;
                ref = RT_SENTINEL;
                if (!bottom) {
                    // TLAS leaf: TraverseFunction.hlsli:598-634
                    const uint4 *ip = reinterpret_cast<const uint4 *>(A.inst + slot);
                    const uint4 m3 = __ldg(ip + 3);
                    if ((m3.x >> 24) & 0xFFu) {  // InstanceInclusionMask is 0xFF for every ray of the pipelines
                        const float4 r0 = __ldg(reinterpret_cast<const float4 *>(ip));
                        const float4 r1 = __ldg(reinterpret_cast<const float4 *>(ip + 1));
                        const float4 r2 = __ldg(reinterpret_cast<const float4 *>(ip + 2));
                        const uint4 m4 = __ldg(ip + 4);
                        instFlags = m3.y >> 24;
                        instOffset = m3.y & 0x00ffffffu;
                        {
                            const bool useCulling = !(instFlags & RT_INSTANCE_FLAG_TRIANGLE_CULL_DISABLE);
                            const bool flip = (instFlags & RT_INSTANCE_FLAG_TRIANGLE_FRONT_COUNTERCLOCKWISE) != 0;
                            const uint32_t backFlag = flip ? RT_RAY_FLAG_CULL_FRONT_FACING_TRIANGLES : RT_RAY_FLAG_CULL_BACK_FACING_TRIANGLES;
                            const uint32_t frontFlag = flip ? RT_RAY_FLAG_CULL_BACK_FACING_TRIANGLES : RT_RAY_FLAG_CULL_FRONT_FACING_TRIANGLES;
                            cull = (useCulling && (rayFlags & frontFlag)) ? 2 : ((useCulling && (rayFlags & backFlag)) ? 1 : 0);
                        }
                        const float m[12] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w, r2.x, r2.y, r2.z, r2.w};
                        const f3 o2 = xform_point(m, mk3(wox, woy, woz));
                        const f3 d2 = xform_vector(m, mk3(wdx, wdy, wdz));
                        cur = make_ray_pre(o2.x, o2.y, o2.z, d2.x, d2.y, d2.z);
                        nodes = reinterpret_cast<const rt_wide_node *>(uintptr_t(uint64_t(m4.x) | (uint64_t(m4.y) << 32)));
                        tris = reinterpret_cast<const rt_packed_tri *>(uintptr_t(uint64_t(m4.z) | (uint64_t(m4.w) << 32)));
                        bottom = true;
                        blasBase = sp;
                        ref = m3.w;
                    }
                } else {
                    // BLAS leaf: TraverseFunction.hlsli:635-735
                    const float4 *tp = reinterpret_cast<const float4 *>(tris + slot);
                    const float4 p0 = __ldg(tp), p1 = __ldg(tp + 1), p2 = __ldg(tp + 2);
                    // every geometry of the pipelines is hit by these rays: their flags carry no FORCE_* / CULL_(NON_)OPAQUE bits
                    float t0 = tCur, bu, bv;
                    if (ray_triangle(t0, bu, bv, cull, cur, p0, p1, p2.x) && t0 < tCur && t0 > tmin) {
                        tCur = t0, hu = bu, hv = bv;
                        hprim = __float_as_uint(p2.y);
                        hrec = rayContribution + instOffset;  // geometry multiplier is 0 in both shader libraries
                        if (ANY) {
                            finish();
                        }
                    }
                }
            }
        } else {
            // ---------------------------------------------------------------- internal phase
            if (alive && !atLeaf) {
                const float4 *np = reinterpret_cast<const float4 *>(nodes + ref);
                const float4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3);
                float lt, rt;
                const bool lh = ray_box(lt, tCur, cur, n0.x, n0.y, n0.z, n1.x, n1.y, n1.z);
                const bool rh = ray_box(rt, tCur, cur, n2.x, n2.y, n2.z, n3.x, n3.y, n3.z);
                const uint32_t l = __float_as_uint(n0.w), r = __float_as_uint(n1.w);
                if (lh && rh) {
                    const bool rightFirst = rt < lt;
                    if (sp < RT_STACK_SIZE) stack[sp++] = rightFirst ? l : r;
                    else atomicOr(status, 1u);
                    ref = rightFirst ? r : l;
                } else if (lh || rh) {
                    ref = rh ? r : l;
                } else {
                    ref = RT_SENTINEL;
                }
            }
        }
        // ------------------------------------------------------------------ pop
        if (alive && ref == RT_SENTINEL) {
            if (bottom && sp == blasBase) {  // leaving the BLAS: back to the world-space ray
                bottom = false;
                cur = make_ray_pre(wox, woy, woz, wdx, wdy, wdz);
                nodes = A.wide;
                blasBase = -1;
            }
            if (sp == 0) finish();
            else ref = stack[--sp];
        }
    }
}
