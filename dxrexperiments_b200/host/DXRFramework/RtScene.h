// RtScene.h — headless counterpart of libs/DXRFramework/RtScene.h:9-45: a list of (model, transform) instances
// and the TLAS built over them (RtScene.cpp:18-52).
#pragma once
#include "RtModel.h"

namespace DXRFramework {

class RtScene {
public:
    using SharedPtr = std::shared_ptr<RtScene>;
    static SharedPtr create() { return SharedPtr(new RtScene()); }
    ~RtScene() = default;

    void addModel(RtModel::SharedPtr model, DirectX::XMMATRIX transform) { mInstances.push_back({model, transform}); }
    RtModel::SharedPtr getModel(UINT index) const { return mInstances.at(index).model; }
    UINT getNumInstances() const { return static_cast<UINT>(mInstances.size()); }

    RtBuffer::SharedPtr getTlasResource() const { return mTlasBuffer; }
    const void *getTlasWrappedPtr() const { return mTlasBuffer ? mTlasBuffer->ptr() : nullptr; }

    // Builds every model's BLAS, then the TLAS with InstanceID = i, InstanceContributionToHitGroupIndex =
    // i * hitGroupCount, mask 0xFF, flags NONE (RtScene.cpp:27-30, Helpers/TopLevelASGenerator.cpp:343-362).
    void build(RtContext::SharedPtr context, UINT hitGroupCount);

private:
    RtScene() = default;
    struct Node {
        RtModel::SharedPtr model;
        DirectX::XMMATRIX transform;
    };
    std::vector<Node> mInstances;
    RtBuffer::SharedPtr mTlasBuffer;
};

}  // namespace DXRFramework
