// RtProgram.h — headless counterpart of libs/DXRFramework/RtProgram.h:18-128 and RtShader.h.
//
// In the reference a program is a DXIL library + entry-point names + root signatures, linked at run time by the
// Fallback Layer into one uber compute shader.  Here the two shader libraries of the application are compiled
// into librt_core.so, so `addShaderLibrary` takes a library TOKEN (kProgressiveRaytracingLibrary /
// kRealtimeRaytracingLibrary, standing in for g_pProgressiveRaytracing / g_pRealtimeRaytracing bytecode) and
// the entry-point names are validated against the fixed export table of that library.  Root-signature
// configurators are accepted and recorded: they define the byte layout of the shader-table records that
// RtBindings parses (two 8-byte descriptor handles + 16 root constants for hit groups; two handles for miss).
#pragma once
#include "RtContext.h"

namespace DXRFramework {

extern const uint8_t kProgressiveRaytracingLibrary[];  // "rt_core:ProgressiveRaytracing"
extern const UINT kProgressiveRaytracingLibrarySize;
extern const uint8_t kRealtimeRaytracingLibrary[];      // "rt_core:RealtimeRaytracing"
extern const UINT kRealtimeRaytracingLibrarySize;
// The compiled-in any-hit and intersection programs of rt_trace_rays_hit_groups (include/rt_types.h RT_ANYHIT_*,
// RT_INTERSECTION_*), exported as AnyHitAccept, AnyHitIgnore, AnyHitEndSearch, AnyHitCutout, IntersectBox,
// IntersectSphere and the closest-hit stand-in ProceduralClosestHit.  Added NEXT TO a pipeline library, it lets
// addHitGroup(idx, closestHit, anyHit, intersection) (libs/DXRFramework/RtProgram.h:51) name them.
extern const uint8_t kHitGroupProgramsLibrary[];        // "rt_core:HitGroupPrograms"
extern const UINT kHitGroupProgramsLibrarySize;

enum class RootParameterType { SRV, UAV, CBV, Constants32Bit, DescriptorTable };

// Stand-in for nv_helpers_dx12::RootSignatureGenerator: records parameters, computes the argument layout.
class RootSignatureGenerator {
public:
    struct Parameter {
        RootParameterType type;
        UINT shaderRegister, registerSpace, numConstants;
    };
    void AddRootParameter(RootParameterType type, UINT shaderRegister = 0, UINT registerSpace = 0, UINT numRootConstants = 1) {
        mParams.push_back({type, shaderRegister, registerSpace, type == RootParameterType::Constants32Bit ? numRootConstants : 0});
    }
    void AddHeapRangesParameter(UINT shaderRegister, UINT registerSpace) {
        mParams.push_back({RootParameterType::DescriptorTable, shaderRegister, registerSpace, 0});
    }
    const std::vector<Parameter> &parameters() const { return mParams; }
    UINT argumentBytes() const;  // DXR packing: 8-byte handles aligned to 8, constants to 4

private:
    std::vector<Parameter> mParams;
};

class RtShader {
public:
    using SharedPtr = std::shared_ptr<RtShader>;
    enum class Type { RayGeneration, Miss, ClosestHit, AnyHit, Intersection };
    RtShader(Type type, std::string entryPoint) : mType(type), mEntryPoint(std::move(entryPoint)) {}
    const std::string &getEntryPoint() const { return mEntryPoint; }
    Type getType() const { return mType; }

private:
    Type mType;
    std::string mEntryPoint;
};

class RtProgram {
public:
    using SharedPtr = std::shared_ptr<RtProgram>;

    class Desc {
    public:
        Desc() = default;
        Desc &addShaderLibrary(const uint8_t *bytecode, UINT bytecodeSize, const std::vector<std::wstring> &symbolExports);
        Desc &setRayGen(const std::string &raygen);
        Desc &addMiss(uint32_t missIndex, const std::string &miss);
        Desc &addHitGroup(uint32_t hitIndex, const std::string &closestHit, const std::string &anyHit, const std::string &intersection = "");

        using RootSignatureConfigurator = std::function<void(RootSignatureGenerator &config)>;
        Desc &configureGlobalRootSignature(RootSignatureConfigurator configure);
        Desc &configureRayGenRootSignature(RootSignatureConfigurator configure);
        Desc &configureHitGroupRootSignature(RootSignatureConfigurator configure);
        Desc &configureMissRootSignature(RootSignatureConfigurator configure);

    private:
        friend class RtProgram;
        struct HitProgramEntry {
            std::string intersection, anyHit, closestHit;
        };
        int mLibrary = -1;  // rt_program_kind, -1 = none added
        std::vector<std::string> mExports;
        std::string mRayGen;
        std::vector<std::string> mMiss;
        std::vector<HitProgramEntry> mHit;
        RootSignatureGenerator mGlobalRootSignatureConfig, mRayGenRootSignatureConfig, mHitGroupRootSignatureConfig, mMissRootSignatureConfig;
    };

    struct HitGroup {
        RtShader::SharedPtr mClosestHit, mAnyHit, mIntersection;
        std::string mExportName;
    };

    static SharedPtr create(RtContext::SharedPtr context, const Desc &desc, uint32_t maxPayloadSize = 64, uint32_t maxAttributesSize = 8);
    ~RtProgram();

    RtShader::SharedPtr getRayGenProgram() const { return mRayGenProgram; }
    uint32_t getHitProgramCount() const { return (uint32_t)mHitPrograms.size(); }
    HitGroup getHitProgram(uint32_t rayIndex) const { return mHitPrograms.at(rayIndex); }
    uint32_t getMissProgramCount() const { return (uint32_t)mMissPrograms.size(); }
    RtShader::SharedPtr getMissProgram(uint32_t rayIndex) const { return mMissPrograms.at(rayIndex); }

    // The any-hit / intersection program ids of hit group `rayIndex` (RT_ANYHIT_* / RT_INTERSECTION_*): what
    // GetAnyHitAndIntersectionStateId resolves from the hit-group record in the reference (FL/TraverseFunction.hlsli:651-660).
    rt_hit_group_programs getHitGroupPrograms(uint32_t rayIndex) const;

    rt_program *getNative() const { return mProgram; }
    rt_program_kind getKind() const { return mKind; }
    UINT getHitGroupArgumentBytes() const { return mDesc.mHitGroupRootSignatureConfig.argumentBytes(); }
    UINT getMissArgumentBytes() const { return mDesc.mMissRootSignatureConfig.argumentBytes(); }

private:
    RtProgram(RtContext::SharedPtr context, const Desc &desc);
    Desc mDesc;
    RtContext::SharedPtr mContext;
    rt_program *mProgram = nullptr;
    rt_program_kind mKind = RT_PROGRAM_PROGRESSIVE;
    RtShader::SharedPtr mRayGenProgram;
    std::vector<HitGroup> mHitPrograms;
    std::vector<RtShader::SharedPtr> mMissPrograms;
};

}  // namespace DXRFramework
