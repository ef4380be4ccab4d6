// RealtimeRaytracingPipeline.h — headless counterpart of include/RealtimeRaytracingPipeline.h: 1 spp, two AOV
// outputs (direct lighting u0, indirect specular u1; assets/shaders/RealtimeRaytracing.hlsl:3-4,44-45) that the
// DenoiseCompositor consumes.  accumCount is always 0 (src/RealtimeRaytracingPipeline.cpp:182).
#pragma once
#include "ProgressiveRaytracingPipeline.h"

class RealtimeRaytracingPipeline : public RaytracingPipelineBase {
public:
    using SharedPtr = std::shared_ptr<RealtimeRaytracingPipeline>;
    static SharedPtr create(DXRFramework::RtContext::SharedPtr context) { return SharedPtr(new RealtimeRaytracingPipeline(context)); }

    void update(float elapsedTime, UINT elapsedFrames, UINT prevFrameIndex, UINT frameIndex, UINT width, UINT height) override;
    int getNumOutputs() override { return 2; }
    const char *getName() override { return "Realtime Ray Tracing Pipeline"; }

private:
    explicit RealtimeRaytracingPipeline(DXRFramework::RtContext::SharedPtr context);
};
