// RtBindings.h — headless counterparts of libs/DXRFramework/RtParams.h:9-52, RtBindings.h:10-67 and
// RtState.h:10-43: the shader binding table.
//
// The pipelines fill per-record root arguments exactly as in the reference (appendHeapRanges /
// append32BitConstants), RtBindings lays the records out as [raygen][miss x M][hit(rayType) per instance]
// (RtBindings.cpp:131-164) and apply() hands every record to the core: a hit record's bytes are
// {VB handle, IB handle, 16 dwords MaterialParams}, a miss record's {env 2D handle, env cube handle}.
#pragma once
#include "RtProgram.h"
#include "RtScene.h"

namespace DXRFramework {

// Descriptor of an environment cube (what createTextureSRVHandle(resource, cubemap=true) stands for).
struct RtTexture {
    using SharedPtr = std::shared_ptr<RtTexture>;
    RtBuffer::SharedPtr texels;  // 6 x size x size RGBA fp32
    UINT size = 0;
    bool cubemap = false;
};

// createTextureSRVHandle: the "descriptor" of a texture is the address of its RtTexture (kept alive by the pipeline);
// RtBindings::apply resolves it when it serialises a miss record.  0 = null descriptor.
inline UINT64 createTextureSRVHandle(const RtTexture::SharedPtr &t) { return reinterpret_cast<UINT64>(t.get()); }

class RtParams {
public:
    using SharedPtr = std::shared_ptr<RtParams>;
    static SharedPtr create(UINT initialOffset = 0) { return SharedPtr(new RtParams(initialOffset)); }

    void allocateStorage(UINT sizeInBytes) { mData.assign(sizeInBytes, 0); }
    void appendHeapRanges(UINT64 gpuHandle);
    void appendDescriptor(UINT64 descriptorHandle) { appendHeapRanges(descriptorHandle); }
    void append32BitConstants(const void *constants, UINT num32BitConstants);
    // Copies the arguments written since the last apply into `record` and rewinds (RtParams.cpp:17-27).
    UINT applyRootParams(uint8_t *record);

private:
    explicit RtParams(UINT initialOffset) : mRootOffset(initialOffset), mInitialOffset(initialOffset) {}
    void write(const void *src, UINT size, UINT alignment);
    std::vector<uint8_t> mData;
    UINT mRootOffset, mInitialOffset;
};

class RtState {
public:
    using SharedPtr = std::shared_ptr<RtState>;
    static SharedPtr create(RtContext::SharedPtr) { return SharedPtr(new RtState()); }
    void setProgram(RtProgram::SharedPtr pProg) { mProgram = pProg; }
    RtProgram::SharedPtr getProgram() const { return mProgram; }
    void setMaxTraceRecursionDepth(uint32_t maxDepth) { mMaxTraceRecursionDepth = maxDepth; }
    uint32_t getMaxTraceRecursionDepth() const { return mMaxTraceRecursionDepth; }
    void setMaxPayloadSize(uint32_t maxSize) { mMaxPayloadSize = maxSize; }
    uint32_t getMaxPayloadSize() const { return mMaxPayloadSize; }
    void setMaxAttributeSize(uint32_t maxSize) { mMaxAttributeSize = maxSize; }
    uint32_t getMaxAttributeSize() const { return mMaxAttributeSize; }

private:
    RtState() = default;
    uint32_t mMaxTraceRecursionDepth = 1, mMaxPayloadSize = 20, mMaxAttributeSize = 8;
    RtProgram::SharedPtr mProgram;
};

class RtBindings {
public:
    using SharedPtr = std::shared_ptr<RtBindings>;
    static SharedPtr create(RtContext::SharedPtr context, RtProgram::SharedPtr program, RtScene::SharedPtr scene) {
        return SharedPtr(new RtBindings(context, program, scene));
    }

    // Serialises every record into the table and binds it in the core (RtBindings.cpp:100-129).
    void apply(RtContext::SharedPtr context, RtState::SharedPtr state);

    uint32_t getRecordSize() const { return mRecordSize; }
    uint32_t getRayGenRecordIndex() const { return 0; }
    uint32_t getFirstMissRecordIndex() const { return 1; }
    uint32_t getFirstHitRecordIndex() const { return mFirstHitVarEntry; }
    uint32_t getHitProgramsCount() const { return mHitProgCount; }
    uint32_t getMissProgramsCount() const { return mMissProgCount; }
    const std::vector<uint8_t> &getShaderTableData() const { return mShaderTableData; }

    const RtParams::SharedPtr &getHitVars(uint32_t rayID, uint32_t meshID) { return mHitParams.at(rayID).at(meshID); }
    const RtParams::SharedPtr &getRayGenVars() { return mRayGenParams; }
    const RtParams::SharedPtr &getMissVars(uint32_t rayID) { return mMissParams.at(rayID); }
    const RtProgram::SharedPtr &getProgram() { return mProgram; }

private:
    RtBindings(RtContext::SharedPtr context, RtProgram::SharedPtr program, RtScene::SharedPtr scene);
    uint8_t *recordPtr(uint32_t index) { return mShaderTableData.data() + size_t(index) * mRecordSize; }

    RtProgram::SharedPtr mProgram;
    RtScene::SharedPtr mScene;
    std::vector<uint8_t> mShaderTableData;
    uint32_t mMissProgCount = 0, mHitProgCount = 0, mFirstHitVarEntry = 0, mRecordSize = 0;
    static const uint32_t kProgramIdentifierSize = 32;  // D3D12_SHADER_IDENTIFIER_SIZE_IN_BYTES
    RtParams::SharedPtr mRayGenParams;
    std::vector<std::vector<RtParams::SharedPtr>> mHitParams;
    std::vector<RtParams::SharedPtr> mMissParams;
};

}  // namespace DXRFramework
