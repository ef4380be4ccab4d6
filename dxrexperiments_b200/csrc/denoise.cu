// denoise.cu — shared-memory-tiled separable joint-bilateral denoiser + compositor for sm_100a.
//
// Replaces DenoiseCompositor::dispatch's two compute passes (src/DenoiseCompositor.cpp:109-148;
// assets/shaders/BilateralFilter.hlsli:50-118, DenoiseCommon.hlsli:46-77).  Same arithmetic per pixel
// (weights LUT, L1 range weight, color/weight, composite + exposure + Reinhard + gamma), different tiling:
// the reference uses 64x1 / 1x64 line groups (poor 2-D locality for the vertical pass); here both passes
// use 32x8-pixel tiles whose rows are 128-byte coalesced, with the +-k halo staged in shared memory once
// per tile, so every texel is fetched from L2/HBM ~(1 + 2k/tile) times instead of 2k+1 times.
// Out-of-image texels read as 0 for both the input and the joint image (D3D out-of-bounds load).
#include "common.cuh"

namespace {

constexpr int MAX_EXTENT = 20, KERNEL_TAPS = 6;
constexpr int TX = 32, TY = 8;  // pass H: 32 wide x 8 rows;  pass V: 32 wide x 8 rows with a vertical halo

struct DenoiseArgs {
    const float4 *joint, *input;
    float4 *out;
    int w, h, k;
    rt_denoiser_params prm;
    float wts[2 * MAX_EXTENT + 1];
};

__device__ __forceinline__ float4 fetch(const float4 *img, int x, int y, int w, int h) {
    if (x < 0 || y < 0 || x >= w || y >= h) return make_float4(0, 0, 0, 0);
    return __ldg(img + size_t(y) * w + x);
}

__device__ __forceinline__ float range_weight(float4 s, float4 c) {
    float dist = ((fabsf(s.x - c.x) + fabsf(s.y - c.y)) + fabsf(s.z - c.z)) * 10.0f;
    return 1.0f - fminf(fmaxf(dist, 0.0f), 1.0f);
}

// PASS 0 = horizontal (DenoiseCompositorH.hlsl), PASS 1 = vertical + composite (DenoiseCompositorV.hlsl).
template <int PASS>
__global__ void __launch_bounds__(TX * TY) k_denoise(const __grid_constant__ DenoiseArgs A) {
    constexpr int HX = PASS == 0 ? MAX_EXTENT : 0, HY = PASS == 1 ? MAX_EXTENT : 0;
    constexpr int SW = TX + 2 * HX, SH = TY + 2 * HY;
    __shared__ float4 sIn[SH][SW];
    __shared__ float4 sJoint[SH][SW];
    const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
    const int bx = blockIdx.x * TX, by = blockIdx.y * TY;
    const int k = A.k;
    // stage the tile and the +-k halo actually needed
    const int hx = PASS == 0 ? k : 0, hy = PASS == 1 ? k : 0;
    const int cw = TX + 2 * hx, ch = TY + 2 * hy;
    for (int i = threadIdx.x; i < cw * ch; i += TX * TY) {
        const int cx = i % cw, cy = i / cw;
        const int gx = bx + cx - hx, gy = by + cy - hy;
        sIn[cy + (HY - hy)][cx + (HX - hx)] = fetch(A.input, gx, gy, A.w, A.h);
        sJoint[cy + (HY - hy)][cx + (HX - hx)] = fetch(A.joint, gx, gy, A.w, A.h);
    }
    __syncthreads();
    const int x = bx + tx, y = by + ty;
    if (x >= A.w || y >= A.h) return;
    float cr, cg, cb;
    if (A.prm.debugVisualize == 2) {
        const float4 s = sIn[ty + HY][tx + HX];
        cr = s.x, cg = s.y, cb = s.z;
    } else {
        // filterKernel: BilateralFilter.hlsli:75-118
        const float4 cj = sJoint[ty + HY][tx + HX];
        float r = 0.0f, g = 0.0f, b = 0.0f, weight = 0.0f;
        for (int i = -k; i <= k; ++i) {
            const int sx = tx + HX + (PASS == 0 ? i : 0), sy = ty + HY + (PASS == 1 ? i : 0);
            const float4 s = sIn[sy][sx], sj = sJoint[sy][sx];
            const float bw = A.wts[i + MAX_EXTENT] * range_weight(sj, cj);
            r += s.x * bw, g += s.y * bw, b += s.z * bw;
            weight += bw;
        }
        cr = r / weight, cg = g / weight, cb = b / weight;
    }
    if (PASS == 1) {  // DenoiseCommon.hlsli:56-74
        const float4 d = sJoint[ty + HY][tx + HX];
        if (A.prm.debugVisualize == 0) cr += d.x, cg += d.y, cb += d.z;
        else if (A.prm.debugVisualize == 3) cr = d.x, cg = d.y, cb = d.z;
        cr *= A.prm.exposure, cg *= A.prm.exposure, cb *= A.prm.exposure;
        if (A.prm.tonemap) {  // reinhardToneMap :33-38
            const float lum = (cr * 0.299f + cg * 0.587f) + cb * 0.114f;
            const float reinhard = lum / (lum + 1);
            const float s = reinhard / lum;
            cr = fmaxf(cr * s, 0.0f), cg = fmaxf(cg * s, 0.0f), cb = fmaxf(cb * s, 0.0f);
        }
        if (A.prm.gammaCorrect) {
            const float e = 1.0f / A.prm.gamma;
            cr = fminf(fmaxf(powf(cr, e), 0.0f), 1.0f), cg = fminf(fmaxf(powf(cg, e), 0.0f), 1.0f), cb = fminf(fmaxf(powf(cb, e), 0.0f), 1.0f);
        }
    }
    A.out[size_t(y) * A.w + x] = make_float4(cr, cg, cb, 1.0f);
}

}  // namespace

extern "C" int rt_denoise(rt_context *ctx, const float *direct, const float *indirect_specular, float *tmp, float *out, uint32_t width,
                          uint32_t height, const rt_denoiser_params *prm) {
    RT_REQUIRE(ctx && direct && indirect_specular && tmp && out && prm, "null argument");
    RT_REQUIRE(width > 0 && height > 0, "empty image");
    RT_REQUIRE(((uintptr_t(direct) | uintptr_t(indirect_specular) | uintptr_t(tmp) | uintptr_t(out)) & 15) == 0, "images must be 16-byte aligned");
    RT_CUDA(cudaSetDevice(ctx->device));
    DenoiseArgs A{};
    A.joint = reinterpret_cast<const float4 *>(direct);
    A.w = int(width), A.h = int(height);
    // maxKernelSize beyond MAX_EXTENT reads outside the reference's LDS tile (undefined); clamp.
    A.k = prm->maxKernelSize < 0 ? 0 : (prm->maxKernelSize > MAX_EXTENT ? MAX_EXTENT : prm->maxKernelSize);
    A.prm = *prm;
    // per-group weight table: BilateralFilter.hlsli:80-90
    for (int i = -MAX_EXTENT; i <= MAX_EXTENT; ++i) {
        int ai = i < 0 ? -i : i;
        float radius = float(prm->maxKernelSize);
        float den = 0.001f + (radius * 0.8f < 0 ? -(radius * 0.8f) : radius * 0.8f);
        int idx = int(float(ai * (KERNEL_TAPS - 1)) / den);
        idx = idx < 0 ? 0 : (idx > KERNEL_TAPS ? KERNEL_TAPS : idx);
        A.wts[i + MAX_EXTENT] = idx < 2 ? 1.0f : (idx < 3 ? 0.9f : (idx < 4 ? 0.75f : (idx < 5 ? 0.6f : (idx < 6 ? 0.5f : 0.0f))));
    }
    dim3 grid(rt_div_up(width, TX), rt_div_up(height, TY));
    A.input = reinterpret_cast<const float4 *>(indirect_specular);
    A.out = reinterpret_cast<float4 *>(tmp);
    k_denoise<0><<<grid, TX * TY, 0, ctx->stream>>>(A);
    A.input = reinterpret_cast<const float4 *>(tmp);
    A.out = reinterpret_cast<float4 *>(out);
    k_denoise<1><<<grid, TX * TY, 0, ctx->stream>>>(A);
    ctx->launches += 2;
    RT_LAUNCH_CHECK();
    return RT_OK;
}
