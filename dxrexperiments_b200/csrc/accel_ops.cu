// accel_ops.cu — operations on finished acceleration structures: clone / compact copy, post-build info, host query.
//
// Replaces GpuBvh2Builder::CopyRaytracingAccelerationStructure (FL/GpuBVH2Builder.cpp:330-347 -> FL/GpuBvh2Copy.hlsl:17-27,
// a 4-byte-per-thread strided copy of BVHOffsets.totalSize bytes) and EmitRaytracingAccelerationStructurePostbuildInfo
// (FL/GpuBVH2Builder.cpp:459-470 -> FL/GetBVHCompactedSize.hlsl:22-63, which reports BVHOffsets.totalSize).
// Here a result buffer is [reference blob][rt_ext_header][traversal section][update caches], so the byte count comes from
// the ext header; the copy moves 16 bytes per thread per step and sizes its grid to the SM count.
//
// A bottom-level result buffer holds offsets only, so its bytes are position independent: cloning, downloading and
// uploading it elsewhere (another GPU, a file) yields a usable BLAS — that is the serialised form.  A top-level buffer
// holds the device addresses of its BLASes (as the reference's holds WRAPPED_GPU_POINTERs) and is only valid while
// they stay where they are; after relocating BLASes the TLAS is rebuilt (microseconds).
#include <algorithm>

#include "common.cuh"

namespace {

__device__ __forceinline__ const rt_ext_header *ext_of(const uint8_t *as) {
    const rt_bvh_offsets *o = reinterpret_cast<const rt_bvh_offsets *>(as);
    return reinterpret_cast<const rt_ext_header *>(as + align_up(o->totalSize, 64));
}

// The byte count is read on the device from the source's own header (as GpuBvh2Copy.hlsl:20 does); the host has
// validated the header before the launch.
__global__ void __launch_bounds__(256) k_as_copy(uint8_t *dst, const uint8_t *src, int compact) {
    const rt_ext_header *e = ext_of(src);
    if (e->magic != RT_EXT_MAGIC) return;
    const uint64_t bytes = compact ? e->compacted_bytes : e->total_bytes;  // both are multiples of 4; buffers are 64-B aligned
    const uint64_t n16 = bytes / 16;
    const uint4 *s = reinterpret_cast<const uint4 *>(src);
    uint4 *d = reinterpret_cast<uint4 *>(dst);
    const uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
    for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n16; i += stride) d[i] = __ldcs(s + i);
    if (blockIdx.x == 0 && threadIdx.x < ((bytes - n16 * 16) >> 2)) {
        const uint64_t w = n16 * 4 + threadIdx.x;
        reinterpret_cast<uint32_t *>(dst)[w] = reinterpret_cast<const uint32_t *>(src)[w];
    }
}

// COMPACT drops the update caches: the copy can be traced but no longer updated.
__global__ void k_as_patch_compacted(uint8_t *dst, const uint8_t *src) {
    const rt_ext_header *se = ext_of(src);
    if (se->magic != RT_EXT_MAGIC) return;
    const rt_bvh_offsets *o = reinterpret_cast<const rt_bvh_offsets *>(src);
    rt_ext_header *e = reinterpret_cast<rt_ext_header *>(dst + align_up(o->totalSize, 64));
    e->off_sort_cache = 0;
    e->off_parents = 0;
    e->build_flags = se->build_flags & ~uint32_t(RT_BUILD_FLAG_ALLOW_UPDATE);
    e->total_bytes = se->compacted_bytes;
    e->compacted_bytes = se->compacted_bytes;
}

struct SrcBatch {
    const uint8_t *p[30];  // the reference binds at most 30 BVHs per dispatch too (GetBVHCompactedSize.hlsl:28-59)
};
__global__ void k_as_sizes(SrcBatch b, uint32_t n, uint64_t *out) {
    const uint32_t i = threadIdx.x;
    if (i >= n) return;
    const rt_ext_header *e = ext_of(b.p[i]);
    out[i] = e->magic == RT_EXT_MAGIC ? e->compacted_bytes : 0ull;
}

int read_header(rt_context *ctx, const void *as, rt_bvh_offsets *off, rt_ext_header *e) {
    RT_CUDA(cudaSetDevice(ctx->device));
    RT_CUDA(cudaMemcpyAsync(off, as, sizeof(*off), cudaMemcpyDeviceToHost, ctx->stream));
    RT_CUDA(cudaStreamSynchronize(ctx->stream));
    RT_REQUIRE(off->offsetToBoxes == 16 && off->totalSize >= 48, "not an acceleration structure (BVHOffsets)");
    RT_CUDA(cudaMemcpyAsync(e, static_cast<const uint8_t *>(as) + align_up(off->totalSize, 64), sizeof(*e), cudaMemcpyDeviceToHost,
                            ctx->stream));
    RT_CUDA(cudaStreamSynchronize(ctx->stream));
    RT_REQUIRE(e->magic == RT_EXT_MAGIC, "not an rt_core acceleration structure (ext header magic)");
    return RT_OK;
}

}  // namespace

extern "C" {

int rt_as_get_info(rt_context *ctx, const void *as, rt_as_info *info) {
    RT_REQUIRE(ctx && as && info, "null argument");
    rt_bvh_offsets off{};
    rt_ext_header e{};
    int rc = read_header(ctx, as, &off, &e);
    if (rc) return rc;
    info->count = e.count;
    info->top_level = e.top_level;
    info->build_flags = e.build_flags;
    info->has_procedural = e.has_procedural;
    info->blob_bytes = off.totalSize;
    info->total_bytes = e.total_bytes;
    info->compacted_bytes = e.compacted_bytes;
    return RT_OK;
}

int rt_as_copy(rt_context *ctx, void *dst, uint64_t dst_bytes, const void *src, int mode) {
    RT_REQUIRE(ctx && dst && src, "null argument");
    // FL/GpuBVH2Builder.cpp:337-346: only CLONE and COMPACT are supported, anything else is E_INVALIDARG.
    RT_REQUIRE(mode == RT_COPY_MODE_CLONE || mode == RT_COPY_MODE_COMPACT, "copy mode must be CLONE or COMPACT");
    RT_REQUIRE((uintptr_t(dst) & 63) == 0 && (uintptr_t(src) & 63) == 0, "buffers must be 64-byte aligned");
    RT_REQUIRE(dst != src, "source and destination are the same buffer");
    rt_bvh_offsets off{};
    rt_ext_header e{};
    int rc = read_header(ctx, src, &off, &e);
    if (rc) return rc;
    const uint64_t need = mode == RT_COPY_MODE_COMPACT ? e.compacted_bytes : e.total_bytes;
    if (dst_bytes < need) {
        rt_set_error("buffer too small: destination %llu < %llu", (unsigned long long)dst_bytes, (unsigned long long)need);
        return RT_ERR_TOO_SMALL;
    }
    const uint8_t *s = static_cast<const uint8_t *>(src);
    uint8_t *d = static_cast<uint8_t *>(dst);
    RT_REQUIRE(d + need <= s || s + need <= d, "source and destination overlap");
    const int grid = int(std::min<uint64_t>(uint64_t(ctx->num_sms) * 8, std::max<uint64_t>(1, need / (16 * 256))));
    k_as_copy<<<grid, 256, 0, ctx->stream>>>(d, s, mode == RT_COPY_MODE_COMPACT);
    ctx->launches++;
    if (mode == RT_COPY_MODE_COMPACT && e.compacted_bytes != e.total_bytes) {
        k_as_patch_compacted<<<1, 1, 0, ctx->stream>>>(d, s);
        ctx->launches++;
    }
    RT_LAUNCH_CHECK();
    return RT_OK;
}

int rt_as_emit_postbuild_info(rt_context *ctx, uint64_t *dst_sizes, uint32_t n, const void *const *sources) {
    RT_REQUIRE(ctx && (n == 0 || (dst_sizes && sources)), "null argument");
    RT_CUDA(cudaSetDevice(ctx->device));
    for (uint32_t base = 0; base < n; base += 30) {
        SrcBatch b{};
        const uint32_t m = std::min(30u, n - base);
        for (uint32_t i = 0; i < m; ++i) {
            RT_REQUIRE(sources[base + i] != nullptr, "null acceleration structure");
            b.p[i] = static_cast<const uint8_t *>(sources[base + i]);
        }
        k_as_sizes<<<1, 32, 0, ctx->stream>>>(b, m, dst_sizes + base);
        ctx->launches++;
    }
    RT_LAUNCH_CHECK();
    return RT_OK;
}

}  // extern "C"
