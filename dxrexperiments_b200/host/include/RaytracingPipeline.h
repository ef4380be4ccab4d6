// RaytracingPipeline.h — the operator interface that stays (include/RaytracingPipeline.h:8-39 of the reference),
// minus D3D12 types: the command list / command queue arguments are dropped (work is stream ordered inside the
// RtContext), DXGI_FORMAT is accepted and ignored (outputs are RGBA fp32 — see DESIGN.md), and output
// resources are RtBuffers instead of ID3D12Resource + descriptor handles.
#pragma once
#include <memory>
#include <random>

#include "../DXRFramework/RtBindings.h"
#include "../DXRFramework/RtContext.h"
#include "../DXRFramework/RtScene.h"
#include "Camera.h"

typedef rt_material_params MaterialParams;          // assets/shaders/RaytracingHlslCompat.h:87-96
typedef rt_per_frame_constants PerFrameConstants;   // :79-85
typedef rt_debug_options DebugOptions;              // :64-77
typedef rt_camera_params CameraParams;              // :41-50
enum DXGI_FORMAT { DXGI_FORMAT_R16G16B16A16_FLOAT = 10, DXGI_FORMAT_R32G32B32A32_FLOAT = 2 };

class RaytracingPipeline {
public:
    using SharedPtr = std::shared_ptr<RaytracingPipeline>;
    virtual ~RaytracingPipeline() {}

    virtual void userInterface() = 0;  // ImGui in the reference; here: no-op (options are plain members)
    virtual void update(float elapsedTime, UINT elapsedFrames, UINT prevFrameIndex, UINT frameIndex, UINT width, UINT height) = 0;
    virtual void render(UINT frameIndex, UINT width, UINT height) = 0;

    virtual void loadResources(UINT frameCount) = 0;
    virtual void createOutputResource(DXGI_FORMAT format, UINT width, UINT height) = 0;
    virtual void buildAccelerationStructures() = 0;

    struct Material {
        MaterialParams params;
    };

    virtual void addMaterial(Material material) = 0;
    virtual void setCamera(std::shared_ptr<Math::Camera> camera) = 0;
    virtual void setScene(DXRFramework::RtScene::SharedPtr scene) = 0;

    virtual int getNumOutputs() = 0;
    virtual DXRFramework::RtBuffer::SharedPtr getOutputResource(UINT id) = 0;
    virtual uint64_t getOutputUavHandle(UINT id) = 0;
    virtual uint64_t getOutputSrvHandle(UINT id) = 0;

    virtual bool *isActive() = 0;
    virtual const char *getName() = 0;

    // Headless additions: the reference seeds its jitter RNG from the wall clock
    // (src/ProgressiveRaytracingPipeline.cpp:86-88); a fixed seed makes frames reproducible.
    virtual void setJitterSeed(uint32_t seed) = 0;
    virtual const PerFrameConstants &getFrameConstants() const = 0;
};
