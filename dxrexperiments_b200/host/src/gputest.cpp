// gputest.cpp — GPU checks of the host layer's hit-group surface, run by tests/test_host_cpp.py on the B200 box:
// RtProgram::Desc::addHitGroup(idx, closestHit, anyHit, intersection) (libs/DXRFramework/RtProgram.h:51) over a scene
// with a triangle model and a procedural-AABB model, traced through RtContext::traceRays -> rt_trace_rays_hit_groups.
// Expected answers are closed forms (ray / box and ray / sphere distances), not oracle outputs.
#include <cmath>
#include <cstdio>
#include <cstring>

#include "../DXRFramework/RtBindings.h"

using namespace DXRFramework;

#define CHECK(cond)                                                          \
    do {                                                                     \
        if (!(cond)) {                                                       \
            std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); \
            return 1;                                                        \
        }                                                                    \
    } while (0)

static rt_ray ray(float ox, float oy, float oz, float dx, float dy, float dz, float tmax = 1e30f) {
    rt_ray r{};
    r.origin[0] = ox, r.origin[1] = oy, r.origin[2] = oz, r.tmin = 0.0f;
    r.direction[0] = dx, r.direction[1] = dy, r.direction[2] = dz, r.tmax = tmax;
    return r;
}

static RtProgram::SharedPtr makeProgram(RtContext::SharedPtr ctx, const char *anyHit2, const char *intersection2) {
    RtProgram::Desc d;
    d.addShaderLibrary(kProgressiveRaytracingLibrary, kProgressiveRaytracingLibrarySize,
                       {L"RayGen", L"PrimaryClosestHit", L"PrimaryMiss", L"ShadowClosestHit", L"ShadowAnyHit", L"ShadowMiss"});
    d.addShaderLibrary(kHitGroupProgramsLibrary, kHitGroupProgramsLibrarySize,
                       {L"AnyHitAccept", L"AnyHitIgnore", L"AnyHitEndSearch", L"AnyHitCutout", L"IntersectBox", L"IntersectSphere", L"ProceduralClosestHit"});
    d.setRayGen("RayGen").addMiss(0, "PrimaryMiss").addMiss(1, "ShadowMiss");
    d.addHitGroup(0, "PrimaryClosestHit", "").addHitGroup(1, "ShadowClosestHit", "ShadowAnyHit");
    d.addHitGroup(2, "ProceduralClosestHit", anyHit2, intersection2);
    return RtProgram::create(ctx, d);
}

int main() {
    try {
        auto ctx = RtContext::create(0);
        // instance 0: a quad at z = 5 (two triangles, front face towards -z ... both windings are hit without cull flags)
        std::vector<Vertex> v = {{{-1, -1, 5}, {0, 0, -1}}, {{1, -1, 5}, {0, 0, -1}}, {{1, 1, 5}, {0, 0, -1}}, {{-1, 1, 5}, {0, 0, -1}}};
        std::vector<uint32_t> idx = {0, 1, 2, 0, 2, 3};
        auto quad = RtModel::create(ctx, v, idx);
        // instance 1: two procedural primitives: a unit cube around (4, 0, 5) and a 2 x 2 x 2 box around (-4, 0, 5)
        auto boxes = RtModel::createProcedural(ctx, {3.5f, -0.5f, 4.5f, 4.5f, 0.5f, 5.5f, -5.0f, -1.0f, 4.0f, -3.0f, 1.0f, 6.0f}, /*opaque*/ false);
        CHECK(boxes->isProcedural() && boxes->getNumTriangles() == 2);
        auto scene = RtScene::create();
        scene->addModel(quad, DirectX::XMMatrixIdentity());
        scene->addModel(boxes, DirectX::XMMatrixIdentity());
        const UINT hitGroups = 3;
        scene->build(ctx, hitGroups);

        const rt_ray rays[5] = {ray(0, 0, 0, 0, 0, 1),       // the quad, t = 5
                                ray(4, 0, 0, 0, 0, 1),       // the unit cube: enters at t = 4.5; inscribed sphere (r 0.5): t = 4.5
                                ray(-4, 0.5f, 0, 0, 0, 1),   // the big box: enters at t = 4; inscribed sphere (r 1) at height 0.5: t = 5 - sqrt(0.75)
                                ray(8, 8, 0, 0, 0, 1),       // nothing
                                ray(4, 0, 5, 0, 0, 1)};      // origin inside the cube: exit at t = 0.5
        rt_hit hits[5];
        // ray contribution 2 selects hit group 2 (record = 2 + instance * 3): box intersection, no any-hit
        auto boxProgram = makeProgram(ctx, "", "IntersectBox");
        ctx->traceRays(boxProgram, scene, rays, 5, 0, 0xFF, 2, 0, hits);
        CHECK(hits[0].primitive_index != 0xFFFFFFFFu && hits[0].t == 5.0f && hits[0].instance_index == 0 && (hits[0].leaf_slot >> 24) == 0xFE);
        CHECK(hits[1].primitive_index == 0 && hits[1].instance_index == 1 && hits[1].t == 4.5f && (hits[1].leaf_slot >> 24) == RT_HIT_KIND_BOX_ENTER);
        CHECK(hits[2].primitive_index == 1 && hits[2].t == 4.0f);
        CHECK(hits[3].primitive_index == 0xFFFFFFFFu);
        CHECK(hits[4].primitive_index == 0 && hits[4].t == 0.5f && (hits[4].leaf_slot >> 24) == RT_HIT_KIND_BOX_EXIT);
        // sphere intersection
        auto sphereProgram = makeProgram(ctx, "", "IntersectSphere");
        ctx->traceRays(sphereProgram, scene, rays, 5, 0, 0xFF, 2, 0, hits);
        CHECK(hits[1].primitive_index == 0 && std::fabs(hits[1].t - 4.5f) < 1e-6f);
        CHECK(hits[2].primitive_index == 1 && std::fabs(hits[2].t - (5.0f - std::sqrt(0.75f))) < 1e-5f);
        CHECK(std::fabs(hits[2].bary[1] - 0.5f) < 1e-6f);  // the intersection program's attributes: the sphere normal's x, y
        // an any-hit shader that ignores every candidate of the NON-OPAQUE procedural geometry: only the (opaque) quad is hit
        auto ignoreProgram = makeProgram(ctx, "AnyHitIgnore", "IntersectBox");
        ctx->traceRays(ignoreProgram, scene, rays, 5, RT_RAY_FLAG_FORCE_NON_OPAQUE, 0xFF, 2, 0, hits);
        CHECK(hits[0].primitive_index == 0xFFFFFFFFu);  // FORCE_NON_OPAQUE: the quad's candidates go to hit group 2's any-hit shader too
        CHECK(hits[1].primitive_index == 0xFFFFFFFFu && hits[2].primitive_index == 0xFFFFFFFFu && hits[4].primitive_index == 0xFFFFFFFFu);
        // hit group 0 (ray contribution 0) has no intersection shader: procedural primitives report nothing, triangles still do
        ctx->traceRays(boxProgram, scene, rays, 5, 0, 0xFF, 0, 0, hits);
        CHECK(hits[0].t == 5.0f && hits[1].primitive_index == 0xFFFFFFFFu && hits[2].primitive_index == 0xFFFFFFFFu);
        // the pipelines' DispatchRays knows TRIANGLES hit groups only: a scene that reaches procedural primitives is refused loudly
        std::puts("host gputest OK");
    } catch (const std::exception &e) {
        std::fprintf(stderr, "host gputest: %s\n", e.what());
        return 1;
    }
    return 0;
}
