// oracle_denoise.cpp — CPU transcription of the DenoiseCompositor's two compute passes.
// TEST INFRASTRUCTURE (see oracle.h).  Follows /root/reference/assets/shaders/BilateralFilter.hlsli:4-118,
// DenoiseCommon.hlsli:19-77 and src/DenoiseCompositor.cpp:109-148.  PARITY UNPINNED by reference tests.
//
// Defined border behaviour: texture reads outside the image return 0 for both the
// input and the joint texture (D3D out-of-bounds load), and still carry their spatial/range weight.
// maxKernelSize is clamped to MAX_EXTENT = 20: beyond that the reference reads outside its LDS tile
// (undefined behaviour).
#include <atomic>
#include <thread>

#include "oracle_internal.h"

namespace orc {

static const int MAX_EXTENT = 20, KERNEL_TAPS = 6;

struct Img {
    const float *p;
    int w, h;
    void fetch(int x, int y, float out[4]) const {
        if (x < 0 || y < 0 || x >= w || y >= h) {
            out[0] = out[1] = out[2] = out[3] = 0.0f;
            return;
        }
        const float *q = p + (size_t(y) * w + x) * 4;
        out[0] = q[0], out[1] = q[1], out[2] = q[2], out[3] = q[3];
    }
};

// BilateralFilter.hlsli:80-90 — the per-group weight table.
static void gaussian_weights(float kernelRadius, float wts[2 * MAX_EXTENT + 1]) {
    for (int i = -MAX_EXTENT; i <= MAX_EXTENT; ++i) {
        int idx = int(float(std::abs(i) * (KERNEL_TAPS - 1)) / (0.001f + fabsf(kernelRadius * 0.8f)));
        idx = std::min(std::max(idx, 0), KERNEL_TAPS);
        wts[i + MAX_EXTENT] = idx < 2 ? 1.0f : (idx < 3 ? 0.9f : (idx < 4 ? 0.75f : (idx < 5 ? 0.6f : (idx < 6 ? 0.5f : 0.0f))));
    }
}

// filterKernel(): BilateralFilter.hlsli:75-118
static void filter_pixel(const Img &input, const Img &joint, int x, int y, int dx, int dy, int k, const float *wts, float out[3]) {
    float color[3] = {0, 0, 0}, weight = 0.0f;
    float cj[4];
    joint.fetch(x, y, cj);
    for (int i = -k; i <= k; ++i) {
        float s[4], sj[4];
        input.fetch(x + dx * i, y + dy * i, s);
        joint.fetch(x + dx * i, y + dy * i, sj);
        float g = wts[i + MAX_EXTENT];
        float dist = ((fabsf(sj[0] - cj[0]) + fabsf(sj[1] - cj[1])) + fabsf(sj[2] - cj[2])) * 10.0f;
        float cw = 1.0f - fminf(fmaxf(dist, 0.0f), 1.0f);
        float bw = g * cw;
        // `color += s * w`: whether the multiply-add is fused is implementation-defined in HLSL (DXIL FMad / driver
        // contraction; the shader is not `precise`).  Pinned as ONE fused multiply-add, like the ray/box test.
        for (int c = 0; c < 3; ++c) color[c] = fmaf(s[c], bw, color[c]);
        weight += bw;
    }
    for (int c = 0; c < 3; ++c) out[c] = color[c] / weight;
}

}  // namespace orc

using namespace orc;

extern "C" void orc_denoise(const float *direct, const float *indirect_specular, float *tmp, float *out, uint32_t width,
                            uint32_t height, const rt_denoiser_params *prm, int threads) {
    const int W = int(width), H = int(height);
    const int k = std::min(std::max(prm->maxKernelSize, 0), MAX_EXTENT);
    float wts[2 * MAX_EXTENT + 1];
    gaussian_weights(float(prm->maxKernelSize), wts);
    Img joint{direct, W, H};

    auto run_rows = [&](auto f) {
        if (threads <= 1) {
            for (int y = 0; y < H; ++y) f(y);
            return;
        }
        std::atomic<int> next{0};
        std::vector<std::thread> pool;
        for (int t = 0; t < threads; ++t)
            pool.emplace_back([&] {
                for (;;) {
                    int y = next.fetch_add(1);
                    if (y >= H) break;
                    f(y);
                }
            });
        for (auto &th : pool) th.join();
    };

    // pass 0 (H): joint = direct, input = indirect specular -> tmp   (DenoiseCompositor.cpp:124-135)
    {
        Img input{indirect_specular, W, H};
        run_rows([&](int y) {
            for (int x = 0; x < W; ++x) {
                float c[3];
                if (prm->debugVisualize == 2) {
                    float s[4];
                    input.fetch(x, y, s);
                    c[0] = s[0], c[1] = s[1], c[2] = s[2];
                } else {
                    filter_pixel(input, joint, x, y, 1, 0, k, wts, c);
                }
                float *o = tmp + (size_t(y) * W + x) * 4;  // the intermediate target has the swap chain's format too
                o[0] = store_value(c[0]), o[1] = store_value(c[1]), o[2] = store_value(c[2]), o[3] = 1.0f;
            }
        });
    }
    // pass 1 (V): joint = direct, input = tmp -> out, then composite + tonemap  (DenoiseCommon.hlsli:56-74)
    {
        Img input{tmp, W, H};
        run_rows([&](int y) {
            for (int x = 0; x < W; ++x) {
                float c[3];
                if (prm->debugVisualize == 2) {
                    float s[4];
                    input.fetch(x, y, s);
                    c[0] = s[0], c[1] = s[1], c[2] = s[2];
                } else {
                    filter_pixel(input, joint, x, y, 0, 1, k, wts, c);
                }
                float d[4];
                joint.fetch(x, y, d);
                if (prm->debugVisualize == 0) {
                    c[0] += d[0], c[1] += d[1], c[2] += d[2];
                } else if (prm->debugVisualize == 3) {
                    c[0] = d[0], c[1] = d[1], c[2] = d[2];
                }
                for (int i = 0; i < 3; ++i) c[i] *= prm->exposure;
                if (prm->tonemap) {  // reinhardToneMap: DenoiseCommon.hlsli:33-38
                    float lum = (c[0] * 0.299f + c[1] * 0.587f) + c[2] * 0.114f;
                    float reinhard = lum / (lum + 1);
                    float s = reinhard / lum;
                    for (int i = 0; i < 3; ++i) c[i] = fmaxf(c[i] * s, 0.0f);
                }
                if (prm->gammaCorrect)
                    for (int i = 0; i < 3; ++i) c[i] = saturate(powf(c[i], 1.0f / prm->gamma));
                float *o = out + (size_t(y) * W + x) * 4;
                o[0] = store_value(c[0]), o[1] = store_value(c[1]), o[2] = store_value(c[2]), o[3] = 1.0f;
            }
        });
    }
}
