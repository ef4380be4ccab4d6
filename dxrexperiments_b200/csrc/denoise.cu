// denoise.cu — shared-memory-tiled separable joint-bilateral denoiser + compositor for sm_100a.
//
// Replaces DenoiseCompositor::dispatch's two compute passes (src/DenoiseCompositor.cpp:109-148;
// assets/shaders/BilateralFilter.hlsli:50-118, DenoiseCommon.hlsli:46-77).  Same arithmetic per pixel, in the same
// order (weights LUT, L1 range weight, taps accumulated from -k to +k, color/weight, composite + exposure + Reinhard
// + gamma); what differs is how the taps are fetched.  The reference's 64x1 / 1x64 line groups read 2k+1 LDS texels
// per output; a first version here did the same from a 2-D tile and was bound by shared-memory bandwidth
// (50 LDS.128 per pixel and pass: 138 us per 1080p frame, 22 % of the HBM roofline).  Now:
//   * every thread produces R = 8 consecutive outputs ALONG the filter axis and walks the R + 2k texels they need
//     once, keeping the 8 centre joints and 8 accumulators in registers: (R + 2k) / R = 4 texel fetches per output
//     instead of 25;
//   * the tile is stored as six float planes (rgb of input and joint; alpha is never read), so lanes touch
//     consecutive banks — pass V directly (lanes along x), pass H through a +1-per-8 column padding that spreads the
//     lanes' 8-pixel strides over all 32 banks;
//   * pass H tiles are 256 x 8 pixels (halo overhead 2k/256), pass V tiles 32 x 64 (2k/64); rows are staged with
//     coalesced 16-byte loads and pass H writes its results back through shared memory so stores are coalesced too.
// Out-of-image texels read as 0 for both the input and the joint image (D3D out-of-bounds load).
#include <cuda_fp16.h>

#include <type_traits>

#include "common.cuh"

namespace {

constexpr int MAX_EXTENT = 20, KERNEL_TAPS = 6;
constexpr int R = 8;                  // outputs per thread along the filter axis
constexpr int kThreadsDn = 256;       // 8 warps
constexpr int H_TW = 32 * R, H_TH = 8;                        // pass H tile: one warp per row, 8 pixels per lane
constexpr int H_PITCH = ((H_TW + 2 * MAX_EXTENT + 8) * 9 + 7) / 8 + 1;  // padded columns: c -> c + (c >> 3); +8: the last unrolled block may read (never use) up to 7 texels past the halo
constexpr int V_TW = 32, V_TH = 8 * R;                        // pass V tile: lanes along x, one 8-row strip per warp
constexpr int V_ROWS = V_TH + 2 * MAX_EXTENT + 8;  // +8: as H_PITCH
constexpr int H_PLANE = H_TH * H_PITCH, V_PLANE = V_ROWS * V_TW;
constexpr size_t H_SMEM = (6 * H_PLANE + 64) * sizeof(float), V_SMEM = (6 * V_PLANE + 64) * sizeof(float);

struct DenoiseArgs {
    const float4 *joint, *input;
    float4 *out;
    int w, h, k;
    int half;  // stores round through fp16 (the reference's intermediate and output targets are R16G16B16A16_FLOAT)
    rt_denoiser_params prm;
    float wts[2 * MAX_EXTENT + 1];
};

__device__ __forceinline__ float4 fetch(const float4 *img, int x, int y, int w, int h) {
    if (x < 0 || y < 0 || x >= w || y >= h) return make_float4(0, 0, 0, 0);
    return __ldg(img + size_t(y) * w + x);
}

[[maybe_unused]] __device__ __forceinline__ float range_weight(float sx, float sy, float sz, float cx, float cy, float cz) {
    float dist = ((fabsf(sx - cx) + fabsf(sy - cy)) + fabsf(sz - cz)) * 10.0f;
    return 1.0f - fminf(fmaxf(dist, 0.0f), 1.0f);
}

__device__ __forceinline__ int hpad(int c) { return c + (c >> 3); }

// Packed fp32 pairs (sm_100 FADD2 / FMUL2 / FFMA2: two IEEE round-to-nearest operations per issue slot, each lane's result
// bit-identical to the scalar instruction).  The kernel is issue-bound (ncu: 66 % of the issue slots, DRAM 16 %), so the
// filter's (texel, output) work is done for two neighbouring outputs per instruction wherever the operation has a packed
// form; |x| and saturate only exist as scalar operand modifiers and stay scalar.
#ifndef RT_DENOISE_PACKED
#define RT_DENOISE_PACKED 1
#endif
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk2(f32x2 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

// PASS 0 = horizontal (DenoiseCompositorH.hlsl), PASS 1 = vertical + composite (DenoiseCompositorV.hlsl).
template <int PASS>
__global__ void __launch_bounds__(kThreadsDn, 3) k_denoise(const __grid_constant__ DenoiseArgs A) {
    extern __shared__ float smem[];
    // pass V sizes its planes for the kernel radius in use (k = 12: 74 KB, three blocks per SM; k = 20: 86 KB, two)
    const int PLANE = PASS == 0 ? H_PLANE : (V_TH + 2 * A.k + 8) * V_TW;
    float *sW = smem;            // 41 tap weights
    float *pl = smem + 64;       // planes: in.r in.g in.b joint.r joint.g joint.b
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int k = A.k;
    if (threadIdx.x < 64) {  // sW[RT_DENOISE_PACKED + jj] = weight of tap i = jj - k; zero outside the kernel
        const int i = int(threadIdx.x) - RT_DENOISE_PACKED - k;
        sW[threadIdx.x] = (i >= -k && i <= k) ? A.wts[i + MAX_EXTENT] : 0.0f;
    }
    // ---- stage the tile and the +-k halo along the filter axis
    int x0, y0;  // first output pixel of this thread
    if (PASS == 0) {
        const int bx = blockIdx.x * H_TW, by = blockIdx.y * H_TH;
        const int cols = H_TW + 2 * k;
        for (int row = warp; row < H_TH; row += kThreadsDn / 32) {
            const int gy = by + row;
            for (int c = lane; c < cols; c += 32) {
                const int gx = bx + c - k;
                const float4 a = fetch(A.input, gx, gy, A.w, A.h), j = fetch(A.joint, gx, gy, A.w, A.h);
                float *q = pl + row * H_PITCH + hpad(c);
                q[0] = a.x, q[PLANE] = a.y, q[2 * PLANE] = a.z, q[3 * PLANE] = j.x, q[4 * PLANE] = j.y, q[5 * PLANE] = j.z;
            }
        }
        x0 = bx + R * lane, y0 = by + warp;
    } else {
        const int bx = blockIdx.x * V_TW, by = blockIdx.y * V_TH;
        const int rows = V_TH + 2 * k;
        const int gx = bx + lane;
        for (int row = warp; row < rows; row += kThreadsDn / 32) {
            const int gy = by + row - k;
            const float4 a = fetch(A.input, gx, gy, A.w, A.h), j = fetch(A.joint, gx, gy, A.w, A.h);
            float *q = pl + row * V_TW + lane;
            q[0] = a.x, q[PLANE] = a.y, q[2 * PLANE] = a.z, q[3 * PLANE] = j.x, q[4 * PLANE] = j.y, q[5 * PLANE] = j.z;
        }
        x0 = bx + lane, y0 = by + R * warp;
    }
    __syncthreads();
    // texel at filter-axis offset j from this thread's first output (j in [-k, R-1+k])
    const float *base = PASS == 0 ? pl + warp * H_PITCH : pl + lane;
    auto at = [&](int j) -> const float * {
        return PASS == 0 ? base + hpad(R * lane + j + k) : base + (R * warp + j + k) * V_TW;
    };
    float cjx[R], cjy[R], cjz[R], ar[R], ag[R], ab[R], aw[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const float *q = at(r);
        cjx[r] = q[3 * PLANE], cjy[r] = q[4 * PLANE], cjz[r] = q[5 * PLANE];
        ar[r] = ag[r] = ab[r] = aw[r] = 0.0f;
    }
    if (A.prm.debugVisualize == 2) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const float *q = at(r);
            ar[r] = q[0], ag[r] = q[PLANE], ab[r] = q[2 * PLANE], aw[r] = 1.0f;
        }
    } else {
        // filterKernel (BilateralFilter.hlsli:75-118) for R outputs at once.  Texel jj (0-based from the first halo
        // texel) is tap i = jj - k - r of output r, valid iff 0 <= jj - r <= 2k; every output still receives its taps in
        // the order i = -k .. k.  The tap weight w[jj - r] lives in ring slot (jj - r) & 7: a slot is written once, when
        // its texel is fetched (r = 0), and read by output r exactly r steps later, so with the jj loop unrolled by 8
        // all slot indices are compile-time constants — one weight fetch per texel instead of one per tap.
#if RT_DENOISE_PACKED
        // Outputs r and r + 1 share every instruction that has a packed form.  A tap past an output's last one needs no
        // test: its weight sW[jj - r] is 0 there, and x + 0 * s is x (s is finite), so such a half-pair is value-neutral.
        f32x2 cx2[R / 2], cy2[R / 2], cz2[R / 2], ar2[R / 2], ag2[R / 2], ab2[R / 2], aw2[R / 2];
#pragma unroll
        for (int h = 0; h < R / 2; ++h) {
            cx2[h] = pk2(cjx[2 * h], cjx[2 * h + 1]), cy2[h] = pk2(cjy[2 * h], cjy[2 * h + 1]), cz2[h] = pk2(cjz[2 * h], cjz[2 * h + 1]);
            ar2[h] = ag2[h] = ab2[h] = aw2[h] = pk2(0.0f, 0.0f);
        }
        // ring2[p] = {w[jj], w[jj - 1]}: the weights outputs r and r + 1 apply to texel jj + r, fetched as a pair when texel jj
        // is (the table is stored shifted by one, sW[0] = 0, so that w[-1] reads as 0)
        f32x2 ring2[8];
#pragma unroll
        for (int p = 0; p < 8; ++p) ring2[p] = pk2(0.0f, 0.0f);  // "before the first tap" reads as weight 0
        const f32x2 one2 = pk2(1.0f, 1.0f);
        auto block = [&](const int b, auto firstTag, auto checkTag) {
            constexpr bool FIRST = decltype(firstTag)::value, CHECK = decltype(checkTag)::value;
#pragma unroll
            for (int p = 0; p < 8; ++p) {
                const int jj = 8 * b + p;
                const float *q = at(jj - k);
                const float sx = q[0], sy = q[PLANE], sz = q[2 * PLANE], jx = q[3 * PLANE], jy = q[4 * PLANE], jz = q[5 * PLANE];
                ring2[p] = pk2(sW[jj + 1], sW[jj]);
                const f32x2 sx2 = pk2(sx, sx), sy2 = pk2(sy, sy), sz2 = pk2(sz, sz), jx2 = pk2(jx, jx), jy2 = pk2(jy, jy), jz2 = pk2(jz, jz);
#pragma unroll
                for (int h = 0; h < R / 2; ++h) {
                    const int r = 2 * h;
                    if (FIRST && p < r) continue;              // jj - r < 0: before the first tap of both outputs
                    if (CHECK && jj - r - 1 > 2 * k) continue;  // past the last tap of both
                    float dxl, dxh, dyl, dyh, dzl, dzh;
                    upk2(sub2(jx2, cx2[h]), dxl, dxh), upk2(sub2(jy2, cy2[h]), dyl, dyh), upk2(sub2(jz2, cz2[h]), dzl, dzh);
                    const float ml = __saturatef(((fabsf(dxl) + fabsf(dyl)) + fabsf(dzl)) * 10.0f);
                    const float mh = __saturatef(((fabsf(dxh) + fabsf(dyh)) + fabsf(dzh)) * 10.0f);
                    const f32x2 bw = mul2(ring2[(p - r) & 7], sub2(one2, pk2(ml, mh)));
                    ar2[h] = fma2(sx2, bw, ar2[h]), ag2[h] = fma2(sy2, bw, ag2[h]), ab2[h] = fma2(sz2, bw, ab2[h]);
                    aw2[h] = add2(aw2[h], bw);
                }
            }
        };
        const int nblocks = (2 * k + R + 7) / 8, nfull = (2 * k + 1) / 8;  // blocks b < nfull have jj <= 2k for every p
        block(0, std::true_type{}, std::integral_constant<bool, true>{});
        int b = 1;
        for (; b < nfull; ++b) block(b, std::false_type{}, std::false_type{});
        for (; b < nblocks; ++b) block(b, std::false_type{}, std::true_type{});
#pragma unroll
        for (int h = 0; h < R / 2; ++h) {
            upk2(ar2[h], ar[2 * h], ar[2 * h + 1]), upk2(ag2[h], ag[2 * h], ag[2 * h + 1]);
            upk2(ab2[h], ab[2 * h], ab[2 * h + 1]), upk2(aw2[h], aw[2 * h], aw[2 * h + 1]);
        }
#else
        float ring[8];
        auto block = [&](const int b, auto firstTag, auto checkTag) {
            constexpr bool FIRST = decltype(firstTag)::value, CHECK = decltype(checkTag)::value;
#pragma unroll
            for (int p = 0; p < 8; ++p) {
                const int jj = 8 * b + p;
                const float *q = at(jj - k);
                const float sx = q[0], sy = q[PLANE], sz = q[2 * PLANE], jx = q[3 * PLANE], jy = q[4 * PLANE], jz = q[5 * PLANE];
                ring[p] = sW[jj];
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    if (FIRST && p < r) continue;          // jj - r < 0: before the first tap of output r
                    if (CHECK && jj - r > 2 * k) continue;  // past its last tap
                    const float bw = ring[(p - r) & 7] * range_weight(jx, jy, jz, cjx[r], cjy[r], cjz[r]);
                    ar[r] = __fmaf_rn(sx, bw, ar[r]), ag[r] = __fmaf_rn(sy, bw, ag[r]), ab[r] = __fmaf_rn(sz, bw, ab[r]);
                    aw[r] += bw;
                }
            }
        };
        const int nblocks = (2 * k + R + 7) / 8, nfull = (2 * k + 1) / 8;  // blocks b < nfull have jj <= 2k for every p
        block(0, std::true_type{}, std::integral_constant<bool, true>{});
        int b = 1;
        for (; b < nfull; ++b) block(b, std::false_type{}, std::false_type{});
        for (; b < nblocks; ++b) block(b, std::false_type{}, std::true_type{});
#endif
    }
    if (PASS == 0) {
        // results go back through shared memory (the input planes are dead) so that the global stores are coalesced
        __syncthreads();
        if (A.prm.debugVisualize != 2) {
#pragma unroll
            for (int r = 0; r < R; ++r) ar[r] = ar[r] / aw[r], ag[r] = ag[r] / aw[r], ab[r] = ab[r] / aw[r];
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            float *q = pl + warp * H_PITCH + hpad(R * lane + r);
            q[0] = ar[r], q[PLANE] = ag[r], q[2 * PLANE] = ab[r];
        }
        __syncwarp();
        if (y0 < A.h) {
            const int bx = blockIdx.x * H_TW;
            for (int c = lane; c < H_TW; c += 32) {
                const int x = bx + c;
                if (x < A.w) {
                    const float *q = pl + warp * H_PITCH + hpad(c);
                    float4 v = make_float4(q[0], q[PLANE], q[2 * PLANE], 1.0f);
                    if (A.half) v.x = __half2float(__float2half_rn(v.x)), v.y = __half2float(__float2half_rn(v.y)), v.z = __half2float(__float2half_rn(v.z));
                    A.out[size_t(y0) * A.w + x] = v;
                }
            }
        }
        return;
    }
    // PASS 1: divide, composite (DenoiseCommon.hlsli:56-74) and store; a warp's 32 lanes write 512 contiguous bytes
    if (x0 >= A.w) return;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int y = y0 + r;
        if (y >= A.h) break;
        float cr = ar[r], cg = ag[r], cb = ab[r];
        if (A.prm.debugVisualize != 2) cr = cr / aw[r], cg = cg / aw[r], cb = cb / aw[r];
        if (A.prm.debugVisualize == 0) cr += cjx[r], cg += cjy[r], cb += cjz[r];
        else if (A.prm.debugVisualize == 3) cr = cjx[r], cg = cjy[r], cb = cjz[r];
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
        if (A.half) cr = __half2float(__float2half_rn(cr)), cg = __half2float(__float2half_rn(cg)), cb = __half2float(__float2half_rn(cb));
        A.out[size_t(y) * A.w + x0] = make_float4(cr, cg, cb, 1.0f);
    }
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
    A.half = int(ctx->render_options.half_render_targets);
    // per-group weight table: BilateralFilter.hlsli:80-90
    for (int i = -MAX_EXTENT; i <= MAX_EXTENT; ++i) {
        int ai = i < 0 ? -i : i;
        float radius = float(prm->maxKernelSize);
        float den = 0.001f + (radius * 0.8f < 0 ? -(radius * 0.8f) : radius * 0.8f);
        int idx = int(float(ai * (KERNEL_TAPS - 1)) / den);
        idx = idx < 0 ? 0 : (idx > KERNEL_TAPS ? KERNEL_TAPS : idx);
        A.wts[i + MAX_EXTENT] = idx < 2 ? 1.0f : (idx < 3 ? 0.9f : (idx < 4 ? 0.75f : (idx < 5 ? 0.6f : (idx < 6 ? 0.5f : 0.0f))));
    }
    // the opt-in is a per-device attribute of the function: set it on every call (a process may hold contexts on several GPUs)
    RT_CUDA(cudaFuncSetAttribute(k_denoise<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(H_SMEM)));
    RT_CUDA(cudaFuncSetAttribute(k_denoise<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(V_SMEM)));
    A.input = reinterpret_cast<const float4 *>(indirect_specular);
    A.out = reinterpret_cast<float4 *>(tmp);
    k_denoise<0><<<dim3(rt_div_up(width, H_TW), rt_div_up(height, H_TH)), kThreadsDn, H_SMEM, ctx->stream>>>(A);
    A.input = reinterpret_cast<const float4 *>(tmp);
    A.out = reinterpret_cast<float4 *>(out);
    const size_t v_smem = (6 * size_t(V_TH + 2 * A.k + 8) * V_TW + 64) * sizeof(float);
    k_denoise<1><<<dim3(rt_div_up(width, V_TW), rt_div_up(height, V_TH)), kThreadsDn, v_smem, ctx->stream>>>(A);
    ctx->launches += 2;
    RT_LAUNCH_CHECK();
    return RT_OK;
}
