// oracle_shade.cpp — CPU transcription of the application shaders.
// TEST INFRASTRUCTURE (see oracle.h).  "S/" = /root/reference/assets/shaders/.
//   S/RaytracingUtils.hlsli   : RNG, samplers, Fresnel
//   S/RaytracingCommon.hlsli  : vertex interpolation, shadow rays, lights, AO, environment
//   S/ProgressiveRaytracing.hlsl, S/RealtimeRaytracing.hlsl : RayGen / shade / closest-hit / miss
// PARITY UNPINNED by any reference test (the app has none); fidelity is by inspection.
#include <atomic>
#include <thread>

#include "oracle_internal.h"

namespace orc {

static const float M_PI_F = 3.1415927f;       // S/RaytracingUtils.hlsli:22
static const float SAMPLER_PI = 3.14159265f;  // :69,92,103
static const float RAY_MAX_T = 1.0e+38f;      // S/RaytracingCommon.hlsli:8
static const float RAY_EPSILON = 0.0001f;     // :9
static const uint32_t MAX_SHADOW_RAY_DEPTH = 2;
// MAX_RADIANCE_RAY_DEPTH is 1 in the reference (S/RaytracingCommon.hlsli:11) — the only value parity is pinned at.
// orc_set_render_options raises it to 2 for BASELINE's "2-bounce" configuration: the Phong-lobe bounce then continues one
// level (shootSecondaryRay at depth 1 no longer returns 0); indirect diffuse stays at depth 0 (shade()'s own
// `currentDepth < 1`, S/ProgressiveRaytracing.hlsl:107) and depth-2 hits take no shadow rays (MAX_SHADOW_RAY_DEPTH 2).
static uint32_t MAX_RADIANCE_RAY_DEPTH = 1;
// The reference's render targets are R16G16B16A16_FLOAT (src/DXRExperimentsApp.cpp:28): every store rounds to fp16 and every
// read-back (the accumulation's `prev`) returns that rounded value.  Off by default (fp32 targets, the declared deviation).
static bool HALF_TARGETS = false;
float store_value(float v) { return HALF_TARGETS ? half_round(v) : v; }

// S/RaytracingUtils.hlsli:26-38
static uint32_t init_rand(uint32_t val0, uint32_t val1) {
    uint32_t v0 = val0, v1 = val1, s0 = 0;
    for (uint32_t n = 0; n < 16; n++) {
        s0 += 0x9e3779b9u;
        v0 += ((v1 << 4) + 0xa341316cu) ^ (v1 + s0) ^ ((v1 >> 5) + 0xc8013ea4u);
        v1 += ((v0 << 4) + 0xad90777du) ^ (v0 + s0) ^ ((v0 >> 5) + 0x7e95761eu);
    }
    return v0;
}
// :41-45
static float next_rand(uint32_t &s) {
    s = 1664525u * s + 1013904223u;
    return float(s & 0x00FFFFFFu) / float(0x01000000);
}
// :49-56
static f3 perpendicular(f3 u) {
    f3 a = vabs(u);
    uint32_t xm = ((a.x - a.y) < 0 && (a.x - a.z) < 0) ? 1 : 0;
    uint32_t ym = (a.y - a.z) < 0 ? (1 ^ xm) : 0;
    uint32_t zm = 1 ^ (xm | ym);
    return cross(u, mk(float(xm), float(ym), float(zm)));
}
// :59-79
static f3 cos_hemisphere(uint32_t &seed, f3 n) {
    float u1 = next_rand(seed), u2 = next_rand(seed);
    f3 bitangent = perpendicular(n);
    f3 tangent = cross(bitangent, n);
    float r = sqrtf(u1);
    float phi = 2.0f * SAMPLER_PI * u2;
    float x = r * cosf(phi), z = r * sinf(phi), y = sqrtf(1.0f - u1);
    return (x * tangent + y * n) + z * bitangent;
}
// :82-98
static f3 uniform_hemisphere(uint32_t &seed, f3 n) {
    float u1 = next_rand(seed), u2 = next_rand(seed);
    f3 bitangent = perpendicular(n);
    f3 tangent = cross(bitangent, n);
    float cosT = u1, sinT = sqrtf(1.0f - cosT * cosT);
    float phi = 2.0f * SAMPLER_PI * u2;
    float x = sinT * cosf(phi), z = sinT * sinf(phi), y = cosT;
    return (x * tangent + y * n) + z * bitangent;
}
// :101-123
static f3 phong_lobe(uint32_t &seed, f3 mirror, float exponent, float &pdf, float &brdf) {
    float u1 = next_rand(seed), u2 = next_rand(seed);
    f3 bitangent = perpendicular(mirror);
    f3 tangent = cross(bitangent, mirror);
    float cosT = powf(u1, 1.0f / (exponent + 1.0f));
    float sinT = sqrtf(1.0f - cosT * cosT);
    float phi = 2.0f * SAMPLER_PI * u2;
    float pc = powf(cosT, exponent);
    pdf = (exponent + 1.0f) / (2.0f * SAMPLER_PI) * pc;
    brdf = (exponent + 2.0f) / (2.0f * SAMPLER_PI) * pc;
    float x = sinT * cosf(phi), z = sinT * sinf(phi), y = cosT;
    return (x * tangent + y * mirror) + z * bitangent;
}
// :126-130
static f3 fresnel_schlick(f3 I, f3 N, f3 f0) {
    float cosi = saturate(dot(-I, N));
    float p = powf(1.0f - cosi, 5.0f);
    return f0 + (mk(1, 1, 1) - f0) * p;
}
static f3 reflect(f3 i, f3 n) { return i - (2.0f * n) * dot(i, n); }

// Environment cube lookup (stands in for envCubemap.SampleLevel(linear, dir, 0),
// S/RaytracingCommon.hlsli:149-159).  D3D face order/orientation, bilinear inside the face,
// clamp at face borders (hardware filters seamlessly across faces; declared deviation).
static void sample_env(const rt_env_cube *env, f3 d, float rgb[3]) {
    rgb[0] = rgb[1] = rgb[2] = 0.0f;
    if (!env || !env->texels || env->size == 0) return;
    f3 a = vabs(d);
    int face;
    float u, v, ma;
    if (a.x >= a.y && a.x >= a.z) {
        ma = a.x;
        if (d.x > 0) face = 0, u = -d.z, v = -d.y; else face = 1, u = d.z, v = -d.y;
    } else if (a.y >= a.z) {
        ma = a.y;
        if (d.y > 0) face = 2, u = d.x, v = d.z; else face = 3, u = d.x, v = -d.z;
    } else {
        ma = a.z;
        if (d.z > 0) face = 4, u = d.x, v = -d.y; else face = 5, u = -d.x, v = -d.y;
    }
    const int n = int(env->size);
    float s = (u / ma + 1.0f) * 0.5f, t = (v / ma + 1.0f) * 0.5f;
    float fx = s * float(n) - 0.5f, fy = t * float(n) - 0.5f;
    float x0f = floorf(fx), y0f = floorf(fy);
    float wx = fx - x0f, wy = fy - y0f;
    int x0 = int(x0f), y0 = int(y0f), x1 = x0 + 1, y1 = y0 + 1;
    x0 = std::min(std::max(x0, 0), n - 1), x1 = std::min(std::max(x1, 0), n - 1);
    y0 = std::min(std::max(y0, 0), n - 1), y1 = std::min(std::max(y1, 0), n - 1);
    const float *base = env->texels + size_t(face) * n * n * 4;
    for (int c = 0; c < 3; ++c) {
        float t00 = base[(size_t(y0) * n + x0) * 4 + c], t10 = base[(size_t(y0) * n + x1) * 4 + c];
        float t01 = base[(size_t(y1) * n + x0) * 4 + c], t11 = base[(size_t(y1) * n + x1) * 4 + c];
        float top = t00 + (t10 - t00) * wx, bot = t01 + (t11 - t01) * wx;
        rgb[c] = top + (bot - top) * wy;
    }
}

struct Ctx {
    const orc_tlas *tlas;
    const rt_hit_record *recs;
    uint32_t n_recs;
    const rt_env_cube *env;
    const rt_per_frame_constants *frame;
    uint32_t width, height;
    bool realtime;
    uint64_t primary = 0, secondary = 0, shadow = 0;
    TraceCounters secondaryCtr;
};

struct Aov {
    f3 direct = mk(0, 0, 0), indirectSpecular = mk(0, 0, 0);
};
struct Payload {
    f3 color = mk(0, 0, 0);
    float distance = 0;
    uint32_t depth = 0;
    Aov aov;
};

static void trace_radiance(Ctx &c, uint32_t px, uint32_t py, f3 o, float tmin, f3 d, uint32_t flags, Payload &payload);

// S/RaytracingCommon.hlsli:84-96
static float shoot_shadow_ray(Ctx &c, f3 o, f3 d, float tmin, float tmax, uint32_t depth) {
    if (depth >= MAX_SHADOW_RAY_DEPTH) return 1.0f;
    c.shadow++;
    HitInfo h = trace_ray(c.tlas, o, tmin, d, tmax,
                          RT_RAY_FLAG_ACCEPT_FIRST_HIT_AND_END_SEARCH | RT_RAY_FLAG_SKIP_CLOSEST_HIT_SHADER, 0xFF, 1, 0, nullptr);
    return h.hit ? 0.0f : 1.0f;  // ShadowMiss sets 1; closest hit is skipped
}
// :126-134
static f3 eval_directional(Ctx &c, f3 p, f3 n, uint32_t depth) {
    const rt_directional_light &dl = c.frame->directionalLight;
    f3 L = normalize(-mk(dl.forwardDir[0], dl.forwardDir[1], dl.forwardDir[2]));
    float NoL = saturate(dot(n, L));
    float visible = shoot_shadow_ray(c, p, L, RAY_EPSILON, RAY_MAX_T, depth);
    return mk(dl.color[0], dl.color[1], dl.color[2]) * dl.color[3] * NoL * visible;
}
// :136-147
static f3 eval_point(Ctx &c, f3 p, f3 n, uint32_t depth) {
    const rt_point_light &pl = c.frame->pointLight;
    f3 path = mk(pl.worldPos[0], pl.worldPos[1], pl.worldPos[2]) - p;
    float dist = length(path);
    f3 L = normalize(path);
    float NoL = saturate(dot(n, L));
    float visible = shoot_shadow_ray(c, p, L, RAY_EPSILON, dist - RAY_EPSILON, depth);
    float falloff = 1.0f / (2 * M_PI_F * dist * dist);
    return mk(pl.color[0], pl.color[1], pl.color[2]) * pl.color[3] * NoL * visible * falloff;
}
// :98-124
static float eval_ao(Ctx &c, uint32_t px, uint32_t py, f3 p, f3 n) {
    float visibility = 0.0f;
    uint32_t seed = init_rand(px + py * c.width, c.frame->cameraParams.frameCount);
    for (int i = 0; i < 4; ++i) {
        f3 dir;
        float NoL, pdf;
        if (c.frame->options.cosineHemisphereSampling) {
            dir = cos_hemisphere(seed, n);
            NoL = saturate(dot(n, dir));
            pdf = NoL / M_PI_F;
        } else {
            dir = uniform_hemisphere(seed, n);
            NoL = saturate(dot(n, dir));
            pdf = 1.0f / (2.0f * M_PI_F);
        }
        visibility += shoot_shadow_ray(c, p, dir, RAY_EPSILON, 10.0f, 1) * NoL / pdf;
    }
    return visibility / 4.0f;
}

// shootSecondaryRay: S/ProgressiveRaytracing.hlsl:41-55, S/RealtimeRaytracing.hlsl:48-63
static f3 shoot_secondary(Ctx &c, uint32_t px, uint32_t py, f3 o, f3 d, float tmin, uint32_t depth) {
    if (depth >= MAX_RADIANCE_RAY_DEPTH) return mk(0, 0, 0);
    Payload p;
    p.depth = depth + 1;
    c.secondary++;
    trace_radiance(c, px, py, o, tmin, d, 0, p);
    return p.color;
}

// S/ProgressiveRaytracing.hlsl:57-78
static f3 eval_indirect_diffuse(Ctx &c, uint32_t px, uint32_t py, f3 p, f3 n, uint32_t &seed, uint32_t depth) {
    f3 color = mk(0, 0, 0);
    if (c.frame->options.cosineHemisphereSampling) {
        f3 dir = cos_hemisphere(seed, n);
        color = color + shoot_secondary(c, px, py, p, dir, RAY_EPSILON, depth) * M_PI_F;
    } else {
        f3 dir = uniform_hemisphere(seed, n);
        float NoL = saturate(dot(n, dir));
        float pdf = 1.0f / (2.0f * M_PI_F);
        color = color + shoot_secondary(c, px, py, p, dir, RAY_EPSILON, depth) * NoL / pdf;
    }
    return color / 1.0f;
}

// shade(): S/ProgressiveRaytracing.hlsl:80-148;  shadeAOV(): S/RealtimeRaytracing.hlsl:65-103
static f3 shade(Ctx &c, uint32_t px, uint32_t py, f3 position, f3 normal, uint32_t depth, f3 rayDir,
                const rt_material_params &m, Aov *aov) {
    const rt_debug_options &opt = c.frame->options;
    if (!c.realtime && opt.showAmbientOcclusionOnly) {
        float ao = eval_ao(c, px, py, position, normal);
        return mk(ao, ao, ao);
    }
    uint32_t seed = init_rand(px + py * c.width, c.frame->cameraParams.frameCount);

    f3 direct = mk(0, 0, 0);
    if (!c.realtime && opt.debug == 2) {
        if (next_rand(seed) < 0.5f)
            direct = direct + eval_directional(c, position, normal, depth) * 2.0f;
        else
            direct = direct + eval_point(c, position, normal, depth) * 2.0f;
    } else {
        direct = direct + eval_directional(c, position, normal, depth);
        direct = direct + eval_point(c, position, normal, depth);
    }

    f3 indirect = mk(0, 0, 0);
    if (!c.realtime && depth < 1 && !opt.noIndirectDiffuse)
        indirect = indirect + eval_indirect_diffuse(c, px, py, position, normal, seed, depth);

    f3 diffuseComponent = (direct + indirect) / M_PI_F;

    f3 fresnel = mk(0, 0, 0), spec = mk(0, 0, 0);
    if (m.type == 1 || m.type == 2) {
        if (m.reflectivity > 0.001f) {
            float exponent = expf((1.0f - m.roughness) * 12.0f);
            float pdf, brdf;
            f3 mirror = reflect(rayDir, normal);
            f3 dir = phong_lobe(seed, mirror, exponent, pdf, brdf);
            f3 refl = shoot_secondary(c, px, py, position, dir, RAY_EPSILON, depth);
            spec = spec + refl * brdf / pdf;
            fresnel = fresnel_schlick(rayDir, normal, mk(m.specular[0], m.specular[1], m.specular[2]));
        }
    }
    f3 albedo = mk(m.albedo[0], m.albedo[1], m.albedo[2]);
    if (c.realtime) {
        if (depth == 0 && aov) {
            aov->direct = albedo * direct / M_PI_F;
            aov->indirectSpecular = m.reflectivity * spec * fresnel;
        }
        return albedo * direct / M_PI_F + m.reflectivity * spec * fresnel;
    }
    if (depth == 0) {
        if (opt.showIndirectDiffuseOnly) return albedo * indirect / M_PI_F;
        else if (opt.showIndirectSpecularOnly) return m.reflectivity * spec * fresnel;
        else if (opt.showFresnelTerm) return fresnel;
        else if (opt.showGBufferAlbedoOnly) return albedo;
        else if (opt.showDirectLightingOnly) return albedo * direct / M_PI_F;
    }
    return (mk(m.emissive[0], m.emissive[1], m.emissive[2]) * m.emissive[3] + albedo * diffuseComponent) +
           m.reflectivity * spec * fresnel;
}

// TraceRay(SceneBVH, flags, 0xFF, 0, 0, 0, ray, payload) + PrimaryClosestHit / PrimaryMiss.
static void trace_radiance(Ctx &c, uint32_t px, uint32_t py, f3 o, float tmin, f3 d, uint32_t flags, Payload &payload) {
    TraceCounters *ctr = payload.depth > 0 ? &c.secondaryCtr : nullptr;
    HitInfo h = trace_ray(c.tlas, o, tmin, d, RAY_MAX_T, flags, 0xFF, 0, 0, ctr);
    if (h.hit) {
        // PrimaryClosestHit + interpolateVertexAttributes (S/RaytracingCommon.hlsli:53-82)
        const rt_hit_record &rec = c.recs[h.hitGroupContribution < c.n_recs ? h.hitGroupContribution : 0];
        float b0 = 1.f - h.bary[0] - h.bary[1], b1 = h.bary[0], b2 = h.bary[1];
        const uint32_t *idx = rec.index_buffer + size_t(h.primitiveIndex) * 3;
        const rt_vertex &v0 = rec.vertex_buffer[idx[0]], &v1 = rec.vertex_buffer[idx[1]], &v2 = rec.vertex_buffer[idx[2]];
        f3 n = (mk(v0.normal[0], v0.normal[1], v0.normal[2]) * b0 + mk(v1.normal[0], v1.normal[1], v1.normal[2]) * b1) +
               mk(v2.normal[0], v2.normal[1], v2.normal[2]) * b2;
        f3 pos = o + h.t * d;  // HitWorldPosition(): S/RaytracingUtils.hlsli:209-212
        Aov aov;
        f3 color = shade(c, px, py, pos, normalize(n), payload.depth, d, rec.material, &aov);
        payload.color = color;
        payload.distance = h.t;
        payload.aov = aov;  // left uninitialised by the shader at depth > 0; never read there
    } else {
        float e[3];
        sample_env(c.env, d, e);
        float s = c.frame->options.environmentStrength;
        payload.color = mk(e[0] * s, e[1] * s, e[2] * s);
        payload.distance = -1.0f;
        payload.aov.direct = payload.color;
        payload.aov.indirectSpecular = mk(0, 0, 0);
    }
}

static void primary_ray(const rt_per_frame_constants *f, uint32_t w, uint32_t h, uint32_t x, uint32_t y, float jitterScale,
                        f3 &o, f3 &d) {
    const rt_camera_params &cam = f->cameraParams;
    float dx = ((float(x) + 0.5f) / float(w)) * 2.f - 1.f;
    float dy = ((float(y) + 0.5f) / float(h)) * 2.f - 1.f;
    float jx = cam.jitters[0] * jitterScale, jy = cam.jitters[1] * jitterScale;
    o = mk(cam.worldEyePos[0], cam.worldEyePos[1], cam.worldEyePos[2]) + mk(jx, jy, 0.0f);
    f3 U = mk(cam.U[0], cam.U[1], cam.U[2]), V = mk(cam.V[0], cam.V[1], cam.V[2]), W = mk(cam.W[0], cam.W[1], cam.W[2]);
    d = normalize((dx * U + (-dy) * V) + W);
}

template <class F>
static void parallel_rows(uint32_t rows, int threads, F f) {
    if (threads <= 1) {
        for (uint32_t y = 0; y < rows; ++y) f(y, 0);
        return;
    }
    std::atomic<uint32_t> next{0};
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t)
        pool.emplace_back([&, t] {
            for (;;) {
                uint32_t y = next.fetch_add(1);
                if (y >= rows) break;
                f(y, t);
            }
        });
    for (auto &th : pool) th.join();
}

}  // namespace orc

using namespace orc;

extern "C" {

uint32_t orc_init_rand(uint32_t v0, uint32_t v1) { return init_rand(v0, v1); }
float orc_next_rand(uint32_t *state) { return next_rand(*state); }

void orc_sample_env(const rt_env_cube *env, const float dir[3], float rgb[3]) { sample_env(env, mk(dir[0], dir[1], dir[2]), rgb); }

void orc_primary_ray(const rt_per_frame_constants *frame, uint32_t width, uint32_t height, uint32_t x, uint32_t y,
                     float jitter_scale, rt_ray *out) {
    f3 o, d;
    primary_ray(frame, width, height, x, y, jitter_scale, o, d);
    out->origin[0] = o.x, out->origin[1] = o.y, out->origin[2] = o.z, out->tmin = 0.0f;
    out->direction[0] = d.x, out->direction[1] = d.y, out->direction[2] = d.z, out->tmax = RAY_MAX_T;
}

void orc_primary_rays(const rt_per_frame_constants *frame, uint32_t width, uint32_t height, float jitter_scale, rt_ray *out) {
    for (uint32_t y = 0; y < height; ++y)
        for (uint32_t x = 0; x < width; ++x) orc_primary_ray(frame, width, height, x, y, jitter_scale, out + size_t(y) * width + x);
}

void orc_set_render_options(uint32_t max_radiance_ray_depth, int half_render_targets) {
    MAX_RADIANCE_RAY_DEPTH = max_radiance_ray_depth < 1 ? 1 : (max_radiance_ray_depth > 2 ? 2 : max_radiance_ray_depth);
    HALF_TARGETS = half_render_targets != 0;
}

void orc_render_progressive(const orc_tlas *t, const rt_hit_record *recs, uint32_t n_recs, const rt_env_cube *env,
                            const rt_per_frame_constants *frame, uint32_t width, uint32_t height, float *accum, int threads,
                            rt_ray_counts *counts, rt_trace_stats *secondary_stats) {
    // RayGen: S/ProgressiveRaytracing.hlsl:11-39
    if (frame->cameraParams.accumCount >= frame->options.maxIterations) return;
    int nt = threads <= 1 ? 1 : threads;
    std::vector<Ctx> ctxs(nt, Ctx{t, recs, n_recs, env, frame, width, height, false});
    parallel_rows(height, threads, [&](uint32_t y, int tid) {
        Ctx &c = ctxs[tid];
        for (uint32_t x = 0; x < width; ++x) {
            f3 o, d;
            primary_ray(frame, width, height, x, y, 30.0f, o, d);
            Payload p;
            c.primary++;
            trace_radiance(c, x, y, o, 0.0f, d, RT_RAY_FLAG_CULL_BACK_FACING_TRIANGLES, p);
            float *px = accum + (size_t(y) * width + x) * 4;
            float cur[4] = {fmaxf(p.color.x, 0.0f), fmaxf(p.color.y, 0.0f), fmaxf(p.color.z, 0.0f), 1.0f};
            uint32_t n = frame->cameraParams.accumCount;
            for (int k = 0; k < 4; ++k) px[k] = store_value((float(n) * px[k] + cur[k]) / float(n + 1));
        }
    });
    for (auto &c : ctxs) {
        if (counts) counts->primary += c.primary, counts->secondary += c.secondary, counts->shadow += c.shadow;
        if (secondary_stats) {
            secondary_stats->internal_visits += c.secondaryCtr.internal;
            secondary_stats->leaf_visits += c.secondaryCtr.leaf;
            secondary_stats->instance_visits += c.secondaryCtr.inst;
            secondary_stats->rays += c.secondary;
            if (c.secondaryCtr.max_stack > secondary_stats->max_stack) secondary_stats->max_stack = c.secondaryCtr.max_stack;
        }
    }
}

void orc_render_realtime(const orc_tlas *t, const rt_hit_record *recs, uint32_t n_recs, const rt_env_cube *env,
                         const rt_per_frame_constants *frame, uint32_t width, uint32_t height, float *direct,
                         float *indirect_specular, int threads, rt_ray_counts *counts) {
    // RayGen: S/RealtimeRaytracing.hlsl:22-46
    int nt = threads <= 1 ? 1 : threads;
    std::vector<Ctx> ctxs(nt, Ctx{t, recs, n_recs, env, frame, width, height, true});
    parallel_rows(height, threads, [&](uint32_t y, int tid) {
        Ctx &c = ctxs[tid];
        for (uint32_t x = 0; x < width; ++x) {
            f3 o, d;
            primary_ray(frame, width, height, x, y, 10.0f, o, d);
            Payload p;
            c.primary++;
            trace_radiance(c, x, y, o, 0.0f, d, RT_RAY_FLAG_CULL_BACK_FACING_TRIANGLES, p);
            float *pd = direct + (size_t(y) * width + x) * 4, *ps = indirect_specular + (size_t(y) * width + x) * 4;
            pd[0] = store_value(fmaxf(p.aov.direct.x, 0.0f)), pd[1] = store_value(fmaxf(p.aov.direct.y, 0.0f));
            pd[2] = store_value(fmaxf(p.aov.direct.z, 0.0f)), pd[3] = 1.0f;
            ps[0] = store_value(fmaxf(p.aov.indirectSpecular.x, 0.0f)), ps[1] = store_value(fmaxf(p.aov.indirectSpecular.y, 0.0f));
            ps[2] = store_value(fmaxf(p.aov.indirectSpecular.z, 0.0f)), ps[3] = 1.0f;
        }
    });
    if (counts)
        for (auto &c : ctxs) counts->primary += c.primary, counts->secondary += c.secondary, counts->shadow += c.shadow;
}

}  // extern "C"
