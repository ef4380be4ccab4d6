// DenoiseCompositor.h — headless counterpart of include/DenoiseCompositor.h:6-58: two compute passes (separable
// joint-bilateral filter H then V + composite/tonemap), src/DenoiseCompositor.cpp:109-148.
#pragma once
#include "../DXRFramework/RtContext.h"
#include "RaytracingPipeline.h"

class DenoiseCompositor {
public:
    using SharedPtr = std::shared_ptr<DenoiseCompositor>;
    static SharedPtr create(DXRFramework::RtContext::SharedPtr context) { return SharedPtr(new DenoiseCompositor(context)); }
    ~DenoiseCompositor() = default;

    void userInterface() {}

    struct InputComponents {
        uint64_t directLightingSrv;    // device address of an RGBA fp32 image (0: use the mock resources)
        uint64_t indirectSpecularSrv;
    };
    void dispatch(InputComponents inputs, UINT frameIndex, UINT width, UINT height);
    // one rank's row band of a frame sharded across GPUs: filters rows [row0, row1), keeps rows [core0, core1) (Pipelines.cpp)
    void dispatchBand(InputComponents inputs, UINT width, UINT height, UINT row0, UINT row1, UINT core0, UINT core1);

    // loadMockResources: the reference loads assets/textures/{DirectLighting,IndirectSpecular}.PNG
    // (src/DenoiseCompositor.cpp:52-70); here the mock inputs are set with setMockResources().
    void loadResources(UINT frameCount, bool loadMockResources);
    void setMockResources(DXRFramework::RtBuffer::SharedPtr direct, DXRFramework::RtBuffer::SharedPtr indirectSpecular);
    void createOutputResource(DXGI_FORMAT format, UINT width, UINT height);

    DXRFramework::RtBuffer::SharedPtr getOutputResource() { return mOutputResource[1]; }
    uint64_t getOutputUavHandle() { return mOutputResource[1] ? mOutputResource[1]->gpuHandle() : 0; }

    bool mActive = true;
    rt_denoiser_params mConstantBuffer;  // exposure 1, gamma 2.2, tonemap on, gamma off, maxKernelSize 12, debugVisualize 0

private:
    explicit DenoiseCompositor(DXRFramework::RtContext::SharedPtr context);
    DXRFramework::RtContext::SharedPtr mRtContext;
    DXRFramework::RtBuffer::SharedPtr mOutputResource[2];
    DXRFramework::RtBuffer::SharedPtr mMock[2];
};
