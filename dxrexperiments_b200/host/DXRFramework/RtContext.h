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
};

}  // namespace DXRFramework
