// RtContext.h — headless counterpart of libs/DXRFramework/RtContext.h:15-60.
//
// The reference's RtContext wraps an ID3D12RaytracingFallbackDevice + command list + one descriptor heap and
// exposes raytrace().  Here it wraps an rt_context (one CUDA device + stream).  "Descriptor handles" become
// plain device addresses, so createBufferSRVHandle & co. reduce to returning the buffer's address.
#pragma once
#include "RtPrefix.h"

namespace DXRFramework {

class RtBindings;
class RtState;
class RtProgram;
class RtScene;

// A device allocation (the ComPtr<ID3D12Resource> of the reference); freed with the last reference.
class RtBuffer {
public:
    using SharedPtr = std::shared_ptr<RtBuffer>;
    ~RtBuffer();
    void *ptr() const { return mPtr; }
    uint64_t size() const { return mBytes; }
    uint64_t gpuHandle() const { return reinterpret_cast<uint64_t>(mPtr); }
    void upload(const void *host, uint64_t bytes, uint64_t offset = 0);
    void download(void *host, uint64_t bytes, uint64_t offset = 0) const;
    void clear();

private:
    friend class RtContext;
    RtBuffer(rt_context *ctx, uint64_t bytes);
    rt_context *mCtx;
    void *mPtr = nullptr;
    uint64_t mBytes = 0;
};

class RtContext : public std::enable_shared_from_this<RtContext> {
public:
    using SharedPtr = std::shared_ptr<RtContext>;

    // Reference: create(ID3D12Device*, ID3D12GraphicsCommandList*, bool forceComputeFallback)  (RtContext.cpp:12-29).
    // There is exactly one backend here (CUDA, sm_100a); a missing device throws.
    static SharedPtr create(int deviceOrdinal = 0);
    ~RtContext();

    rt_context *getNative() const { return mCtx; }
    bool isUsingNativeDxr() const { return false; }

    // RtContext::raytrace (RtContext.cpp:192-222): `depth` is passed through and ignored by the core, as the
    // compute Fallback Layer ignores it.
    void raytrace(std::shared_ptr<RtBindings> bindings, std::shared_ptr<RtState> state, uint32_t width, uint32_t height,
                  uint32_t depth);

    // Strip-interleaved dispatch: the rows of every `groups`-th strip of `stripRows` image rows, starting with strip
    // `group` (screen-tile sharding of one frame across GPUs; rt_dispatch_rays_interleaved).
    void raytraceStrips(std::shared_ptr<RtBindings> bindings, std::shared_ptr<RtState> state, uint32_t width, uint32_t height,
                        uint32_t stripRows, uint32_t groups, uint32_t group);

    // A pixel rectangle [x0, x1) x [y0, y1) of the frame (rt_dispatch_rays_region): the row bands of a realtime frame
    // sharded across GPUs render the rows their DenoiseCompositor pass needs.
    void raytraceRegion(std::shared_ptr<RtBindings> bindings, std::shared_ptr<RtState> state, uint32_t width, uint32_t height,
                        uint32_t x0, uint32_t y0, uint32_t x1, uint32_t y1);

    // TraceRay() with the program's hit groups honoured in full — any-hit and intersection shaders, procedural
    // primitives, every ray flag (FL/TraverseFunction.hlsli:520-799) — for `n` caller-supplied rays (host arrays).
    // Record of a candidate = rayContribution + geometryIndex * geometryMultiplier + InstanceContributionToHitGroupIndex
    // (= instance * hitGroupCount, RtScene.cpp:29); its hit group = record % hitGroupCount.
    void traceRays(std::shared_ptr<RtProgram> program, std::shared_ptr<RtScene> scene, const rt_ray *rays, uint64_t n, uint32_t rayFlags,
                   uint32_t instanceMask, uint32_t rayContribution, uint32_t geometryMultiplier, rt_hit *hits);

    // Multi-GPU accumulation (one process per GPU): joins the NCCL communicator of `worldSize` ranks.  `idFile` is the
    // side channel for rank 0's unique id (rank 0 writes it, the others wait for it).  reduceAccumulation sums
    // weight_r * buffer_r over the ranks onto `root` (rt_accum_reduce); `recv` may be null on the other ranks.
    void joinCommunicator(int worldSize, int rank, const std::string &idFile);
    void reduceAccumulation(const RtBuffer::SharedPtr &send, const RtBuffer::SharedPtr &recv, uint64_t floats, float weight, int root = 0);
    int worldSize() const { return mWorld; }
    int rank() const { return mRank; }

    // CreateBuffer / AllocateUploadBuffer (Helpers/DirectXRaytracingHelper.h)
    RtBuffer::SharedPtr createBuffer(uint64_t bytes);
    RtBuffer::SharedPtr createBuffer(const void *initialData, uint64_t bytes);

    // In the reference these allocate descriptors and return GPU handles; a handle here is the device address.
    uint64_t createBufferSRVHandle(const RtBuffer::SharedPtr &b) const { return b ? b->gpuHandle() : 0; }
    uint64_t createBufferUAVHandle(const RtBuffer::SharedPtr &b) const { return b ? b->gpuHandle() : 0; }

    void waitForGpu();              // DeviceResources::WaitForGpu
    void checkDeviceStatus();       // throws if a traversal overflowed its stack
    void insertUAVBarrier(const RtBuffer::SharedPtr &) {}  // stream order already serialises dependent kernels
    uint64_t launchCount() const;

private:
    explicit RtContext(int deviceOrdinal);
    rt_context *mCtx = nullptr;
    rt_comm *mComm = nullptr;
    int mWorld = 1, mRank = 0;
};

}  // namespace DXRFramework
