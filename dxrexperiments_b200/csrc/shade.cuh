// shade.cuh — device restatement of the application's shader library (RNG, samplers, lights, BRDF,
// environment).  Follows /root/reference/assets/shaders/RaytracingUtils.hlsli:26-130 and
// RaytracingCommon.hlsli:53-159 ("S/" below).  The library is compiled with -fmad=false, so every
// expression is evaluated as written, operation by operation, like the CPU restatement; only the
// transcendental functions (sinf, cosf, powf, expf) differ from glibc by their own few-ulp error.
#pragma once
#include "common.cuh"

#define RT_M_PI 3.1415927f          // S/RaytracingUtils.hlsli:22
#define RT_SAMPLER_PI 3.14159265f   // :69,92,103
#define RT_RAY_MAX_T 1.0e+38f       // S/RaytracingCommon.hlsli:8
#define RT_RAY_EPSILON 0.0001f      // :9

__device__ __forceinline__ f3 operator+(f3 a, f3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ f3 operator-(f3 a, f3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ f3 operator*(f3 a, f3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ f3 operator*(f3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ f3 operator*(float s, f3 a) { return mk3(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ f3 operator/(f3 a, float s) { return mk3(a.x / s, a.y / s, a.z / s); }
__device__ __forceinline__ f3 operator-(f3 a) { return mk3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ float dot3(f3 a, f3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
__device__ __forceinline__ f3 cross3(f3 a, f3 b) { return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
__device__ __forceinline__ float length3(f3 a) { return sqrtf(dot3(a, a)); }
__device__ __forceinline__ f3 normalize3(f3 a) {
    float inv = 1.0f / sqrtf(dot3(a, a));
    return a * inv;
}
__device__ __forceinline__ float saturatef(float x) { return fminf(fmaxf(x, 0.0f), 1.0f); }

// initRand: TEA, 16 rounds (S/RaytracingUtils.hlsli:26-38)
__device__ __forceinline__ uint32_t init_rand(uint32_t val0, uint32_t val1) {
    uint32_t v0 = val0, v1 = val1, s0 = 0;
#pragma unroll
    for (int n = 0; n < 16; n++) {
        s0 += 0x9e3779b9u;
        v0 += ((v1 << 4) + 0xa341316cu) ^ (v1 + s0) ^ ((v1 >> 5) + 0xc8013ea4u);
        v1 += ((v0 << 4) + 0xad90777du) ^ (v0 + s0) ^ ((v0 >> 5) + 0x7e95761eu);
    }
    return v0;
}
// nextRand: 32-bit LCG, 24-bit mantissa (:41-45)
__device__ __forceinline__ float next_rand(uint32_t &s) {
    s = 1664525u * s + 1013904223u;
    return float(s & 0x00FFFFFFu) / float(0x01000000);
}
// getPerpendicularVector (:49-56)
__device__ __forceinline__ f3 perpendicular(f3 u) {
    float ax = fabsf(u.x), ay = fabsf(u.y), az = fabsf(u.z);
    uint32_t xm = ((ax - ay) < 0 && (ax - az) < 0) ? 1 : 0;
    uint32_t ym = (ay - az) < 0 ? (1 ^ xm) : 0;
    uint32_t zm = 1 ^ (xm | ym);
    return cross3(u, mk3(float(xm), float(ym), float(zm)));
}
// getCosHemisphereSample (:59-79)
__device__ __forceinline__ f3 cos_hemisphere(uint32_t &seed, f3 n) {
    float u1 = next_rand(seed), u2 = next_rand(seed);
    f3 bitangent = perpendicular(n);
    f3 tangent = cross3(bitangent, n);
    float r = sqrtf(u1);
    float phi = 2.0f * RT_SAMPLER_PI * u2;
    float x = r * cosf(phi), z = r * sinf(phi), y = sqrtf(1.0f - u1);
    return (x * tangent + y * n) + z * bitangent;
}
// getUniformHemisphereSample (:82-98)
__device__ __forceinline__ f3 uniform_hemisphere(uint32_t &seed, f3 n) {
    float u1 = next_rand(seed), u2 = next_rand(seed);
    f3 bitangent = perpendicular(n);
    f3 tangent = cross3(bitangent, n);
    float cosT = u1, sinT = sqrtf(1.0f - cosT * cosT);
    float phi = 2.0f * RT_SAMPLER_PI * u2;
    float x = sinT * cosf(phi), z = sinT * sinf(phi), y = cosT;
    return (x * tangent + y * n) + z * bitangent;
}
// samplePhongLobe (:101-123)
__device__ __forceinline__ f3 phong_lobe(uint32_t &seed, f3 mirror, float exponent, float &pdf, float &brdf) {
    float u1 = next_rand(seed), u2 = next_rand(seed);
    f3 bitangent = perpendicular(mirror);
    f3 tangent = cross3(bitangent, mirror);
    float cosT = powf(u1, 1.0f / (exponent + 1.0f));
    float sinT = sqrtf(1.0f - cosT * cosT);
    float phi = 2.0f * RT_SAMPLER_PI * u2;
    float pc = powf(cosT, exponent);
    pdf = (exponent + 1.0f) / (2.0f * RT_SAMPLER_PI) * pc;
    brdf = (exponent + 2.0f) / (2.0f * RT_SAMPLER_PI) * pc;
    float x = sinT * cosf(phi), z = sinT * sinf(phi), y = cosT;
    return (x * tangent + y * mirror) + z * bitangent;
}
// FresnelReflectanceSchlick (:126-130)
__device__ __forceinline__ f3 fresnel_schlick(f3 I, f3 N, f3 f0) {
    float cosi = saturatef(dot3(-I, N));
    float p = powf(1.0f - cosi, 5.0f);
    return f0 + (mk3(1, 1, 1) - f0) * p;
}
__device__ __forceinline__ f3 reflect3(f3 i, f3 n) { return i - (2.0f * n) * dot3(i, n); }

// sampleEnvironment's cube lookup (S/RaytracingCommon.hlsli:149-159): D3D face order, bilinear inside
// the face, clamped at face borders (same definition as the CPU restatement).
__device__ __forceinline__ f3 sample_env(const float *texels, uint32_t size, f3 d) {
    if (texels == nullptr || size == 0) return mk3(0, 0, 0);
    float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
    int face;
    float u, v, ma;
    if (ax >= ay && ax >= az) {
        ma = ax;
        if (d.x > 0) face = 0, u = -d.z, v = -d.y; else face = 1, u = d.z, v = -d.y;
    } else if (ay >= az) {
        ma = ay;
        if (d.y > 0) face = 2, u = d.x, v = d.z; else face = 3, u = d.x, v = -d.z;
    } else {
        ma = az;
        if (d.z > 0) face = 4, u = d.x, v = -d.y; else face = 5, u = -d.x, v = -d.y;
    }
    const int n = int(size);
    float s = (u / ma + 1.0f) * 0.5f, t = (v / ma + 1.0f) * 0.5f;
    float fx = s * float(n) - 0.5f, fy = t * float(n) - 0.5f;
    float x0f = floorf(fx), y0f = floorf(fy);
    float wx = fx - x0f, wy = fy - y0f;
    int x0 = int(x0f), y0 = int(y0f), x1 = x0 + 1, y1 = y0 + 1;
    x0 = min(max(x0, 0), n - 1), x1 = min(max(x1, 0), n - 1);
    y0 = min(max(y0, 0), n - 1), y1 = min(max(y1, 0), n - 1);
    const float4 *base = reinterpret_cast<const float4 *>(texels) + size_t(face) * n * n;
    float4 t00 = __ldg(base + size_t(y0) * n + x0), t10 = __ldg(base + size_t(y0) * n + x1);
    float4 t01 = __ldg(base + size_t(y1) * n + x0), t11 = __ldg(base + size_t(y1) * n + x1);
    f3 top = mk3(t00.x + (t10.x - t00.x) * wx, t00.y + (t10.y - t00.y) * wx, t00.z + (t10.z - t00.z) * wx);
    f3 bot = mk3(t01.x + (t11.x - t01.x) * wx, t01.y + (t11.y - t01.y) * wx, t01.z + (t11.z - t01.z) * wx);
    return mk3(top.x + (bot.x - top.x) * wy, top.y + (bot.y - top.y) * wy, top.z + (bot.z - top.z) * wy);
}

// interpolateVertexAttributes (S/RaytracingCommon.hlsli:53-82): normal only (the position is unused).
__device__ __forceinline__ f3 interpolate_normal(const rt_hit_record_dev &rec, uint32_t prim, float bu, float bv) {
    float b0 = 1.f - bu - bv, b1 = bu, b2 = bv;
    const uint32_t *idx = rec.ib + size_t(prim) * 3;
    const uint32_t i0 = __ldg(idx), i1 = __ldg(idx + 1), i2 = __ldg(idx + 2);
    const float *v0 = rec.vb + size_t(i0) * 6 + 3, *v1 = rec.vb + size_t(i1) * 6 + 3, *v2 = rec.vb + size_t(i2) * 6 + 3;
    f3 n0 = mk3(__ldg(v0), __ldg(v0 + 1), __ldg(v0 + 2)), n1 = mk3(__ldg(v1), __ldg(v1 + 1), __ldg(v1 + 2)),
       n2 = mk3(__ldg(v2), __ldg(v2 + 1), __ldg(v2 + 2));
    return (n0 * b0 + n1 * b1) + n2 * b2;
}

// RayGen's camera ray (S/ProgressiveRaytracing.hlsl:17-31, S/RealtimeRaytracing.hlsl:25-40).
__device__ __forceinline__ void primary_ray(const rt_per_frame_constants &f, uint32_t w, uint32_t h, uint32_t x, uint32_t y,
                                            float jitterScale, f3 &o, f3 &d) {
    const rt_camera_params &cam = f.cameraParams;
    float dx = ((float(x) + 0.5f) / float(w)) * 2.f - 1.f;
    float dy = ((float(y) + 0.5f) / float(h)) * 2.f - 1.f;
    float jx = cam.jitters[0] * jitterScale, jy = cam.jitters[1] * jitterScale;
    o = mk3(cam.worldEyePos[0], cam.worldEyePos[1], cam.worldEyePos[2]) + mk3(jx, jy, 0.0f);
    f3 U = mk3(cam.U[0], cam.U[1], cam.U[2]), V = mk3(cam.V[0], cam.V[1], cam.V[2]), W = mk3(cam.W[0], cam.W[1], cam.W[2]);
    d = normalize3((dx * U + (-dy) * V) + W);
}
