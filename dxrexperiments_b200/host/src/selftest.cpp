// selftest.cpp — CPU-only checks of the host layer's own logic (no CUDA calls): OBJ parsing, PFM round trip,
// DDS cube reading, half conversion, shader-table argument packing and program-description validation.
// Run by tests/test_host_cpp.py; exits non-zero on the first failure.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iterator>
#include <string>
#include <vector>

#include "../DXRFramework/RtBindings.h"
#include "../include/Camera.h"
#include "../include/ImageIO.h"

using namespace DXRFramework;

#define CHECK(cond)                                                          \
    do {                                                                     \
        if (!(cond)) {                                                       \
            std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); \
            return 1;                                                        \
        }                                                                    \
    } while (0)

static int testObj(const std::string &dir) {
    const std::string path = dir + "/quad.obj";
    {
        std::ofstream f(path);
        f << "# a quad (polygon face, negative indices, no normals) and a triangle with normals\n"
             "v 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\n"
             "f 1 2 3 4\n"
             "v 0 0 1\nv 1 0 1\nv 0 1 1\nvn 0 0 1\n"
             "f -3//1 -2//1 -1//1\n";
    }
    std::vector<Vertex> v;
    std::vector<uint32_t> idx;
    CHECK(RtModel::loadObj(path, v, idx));
    CHECK(idx.size() == 9);  // fan-triangulated quad (2) + triangle (1)
    CHECK(v.size() == 7);
    CHECK(idx[0] == 0 && idx[1] == 1 && idx[2] == 2 && idx[3] == 0 && idx[4] == 2 && idx[5] == 3);
    CHECK(v[idx[6]].position.z == 1.0f && v[idx[6]].normal.z == 1.0f);
    CHECK(std::fabs(v[0].normal.z - 1.0f) < 1e-6f);  // generated smooth normal of the quad
    CHECK(!RtModel::loadObj(dir + "/does_not_exist.obj", v, idx));
    return 0;
}

static int testPfmAndHalf(const std::string &dir) {
    const uint32_t w = 5, h = 3;
    std::vector<float> img(w * h * 4);
    for (size_t i = 0; i < img.size(); ++i) img[i] = float(i) * 0.25f - 3.0f;
    CHECK(ImageIO::writePFM(dir + "/t.pfm", img.data(), w, h));
    std::vector<float> back;
    uint32_t bw = 0, bh = 0;
    CHECK(ImageIO::readPFM(dir + "/t.pfm", back, bw, bh));
    CHECK(bw == w && bh == h);
    for (uint32_t p = 0; p < w * h; ++p)
        for (int c = 0; c < 3; ++c) CHECK(back[4 * p + c] == img[4 * p + c]);
    CHECK(ImageIO::halfToFloat(0x3C00) == 1.0f && ImageIO::halfToFloat(0xC000) == -2.0f && ImageIO::halfToFloat(0x0000) == 0.0f);
    CHECK(ImageIO::halfToFloat(0x0001) == 5.9604644775390625e-08f && ImageIO::halfToFloat(0x7BFF) == 65504.0f);
    return 0;
}

static int testDds(const std::string &dir) {
    // a 2x2 R16G16B16A16_FLOAT cube with 2 mips, face f filled with the value f
    const std::string path = dir + "/cube.dds";
    {
        std::ofstream f(path, std::ios::binary);
        uint32_t hdr[32] = {};
        hdr[0] = 0x20534444u, hdr[1] = 124, hdr[3] = 2, hdr[4] = 2, hdr[7] = 2, hdr[21] = 0x30315844u;
        f.write(reinterpret_cast<const char *>(hdr), 128);
        const uint32_t dx10[5] = {10, 3, 0x4, 1, 0};
        f.write(reinterpret_cast<const char *>(dx10), 20);
        const uint16_t halves[6] = {0x0000, 0x3C00, 0x4000, 0x4200, 0x4400, 0x4500};  // 0,1,2,3,4,5
        for (int face = 0; face < 6; ++face) {
            for (int i = 0; i < 2 * 2 * 4 + 4; ++i) f.write(reinterpret_cast<const char *>(&halves[face]), 2);  // mip0 + 1x1 mip
        }
    }
    std::vector<float> texels;
    uint32_t size = 0;
    CHECK(ImageIO::readDDSCube(path, texels, size));
    CHECK(size == 2 && texels.size() == 6 * 2 * 2 * 4);
    for (int face = 0; face < 6; ++face)
        for (int i = 0; i < 16; ++i) CHECK(texels[face * 16 + i] == float(face));
    std::vector<float> sky;
    ImageIO::proceduralSkyCube(8, sky);
    CHECK(sky.size() == 6 * 8 * 8 * 4 && sky[3] == 1.0f && sky[1] > 0.3f);
    return 0;
}

static int testBindingsLayout() {
    RootSignatureGenerator hit;
    hit.AddHeapRangesParameter(0, 1);
    hit.AddHeapRangesParameter(1, 1);
    hit.AddRootParameter(RootParameterType::Constants32Bit, 0, 1, 16);
    CHECK(hit.argumentBytes() == 80);  // 2 x 8-byte handles + 16 dwords (the reference reserves maxRootSigSize = 80)
    auto p = RtParams::create(32);
    p->allocateStorage(80);
    rt_material_params m{};
    m.reflectivity = 0.7f;
    m.type = 1;
    p->appendHeapRanges(0x1122334455667788ull);
    p->appendHeapRanges(0x99ull);
    p->append32BitConstants(&m, 16);
    uint8_t rec[80] = {};
    CHECK(p->applyRootParams(rec) == 80);
    uint64_t a, b;
    rt_material_params back;
    std::memcpy(&a, rec, 8), std::memcpy(&b, rec + 8, 8), std::memcpy(&back, rec + 16, 64);
    CHECK(a == 0x1122334455667788ull && b == 0x99ull && back.reflectivity == 0.7f && back.type == 1);
    CHECK(p->applyRootParams(rec) == 0);  // rewound after apply
    bool threw = false;
    try {
        p->appendHeapRanges(1), p->appendHeapRanges(2), p->append32BitConstants(&m, 16), p->append32BitConstants(&m, 1);
    } catch (const std::logic_error &) {
        threw = true;
    }
    CHECK(threw);  // out-of-bounds root arguments are an error, not a silent overwrite
    return 0;
}

static int testProgramDesc() {
    RtProgram::Desc d;
    d.addShaderLibrary(kProgressiveRaytracingLibrary, kProgressiveRaytracingLibrarySize, {L"RayGen", L"PrimaryClosestHit", L"PrimaryMiss"});
    d.setRayGen("RayGen");
    bool threw = false;
    try {
        d.addMiss(1, "ShadowMiss");  // not exported above -> unknown shader identifier (RtBindings.cpp:77-79 behaviour)
    } catch (const std::logic_error &) {
        threw = true;
    }
    CHECK(threw);
    threw = false;
    try {
        const uint8_t junk[] = "DXBC....";
        RtProgram::Desc().addShaderLibrary(junk, sizeof(junk), {});
    } catch (const std::logic_error &) {
        threw = true;
    }
    CHECK(threw);
    Math::Camera cam;
    cam.SetEyeAtUp({8, 10, 30}, {0, 1.5f, 0}, {0, 1, 0});
    CHECK(std::fabs(Math::Length(cam.GetForwardVec()) - 1.0f) < 1e-6f);
    CHECK(std::fabs(Math::Dot(cam.GetForwardVec(), cam.GetUpVec())) < 1e-6f && std::fabs(Math::Dot(cam.GetRightVec(), cam.GetUpVec())) < 1e-6f);
    return 0;
}

static int testExr(const std::string &dir) {
    const uint32_t w = 7, h = 4;
    std::vector<float> img(w * h * 4);
    for (size_t i = 0; i < img.size(); ++i) img[i] = float(i) * 0.37f - 5.0f;
    img[5] = 1e-8f, img[9] = 65520.0f, img[13] = 3.14159274f;
    CHECK(ImageIO::writeEXR(dir + "/t.exr", img.data(), w, h, false));
    std::vector<float> back;
    uint32_t bw = 0, bh = 0;
    CHECK(ImageIO::readEXR(dir + "/t.exr", back, bw, bh));
    CHECK(bw == w && bh == h && back.size() == img.size());
    CHECK(std::memcmp(back.data(), img.data(), img.size() * 4) == 0);  // fp32 files keep the accumulation buffer exactly
    // file anatomy: magic, version 2, attribute block ends before the offset table of h entries, then h chunks of 8 + w*16 bytes
    std::ifstream f(dir + "/t.exr", std::ios::binary);
    std::vector<char> raw((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    uint32_t magic;
    std::memcpy(&magic, raw.data(), 4);
    CHECK(magic == 20000630u && raw[4] == 2 && std::string(raw.data() + 8) == "channels");
    CHECK(ImageIO::writeEXR(dir + "/h.exr", img.data(), w, h, true));
    CHECK(ImageIO::readEXR(dir + "/h.exr", back, bw, bh));
    for (size_t i = 0; i < img.size(); ++i) CHECK(back[i] == ImageIO::halfToFloat(ImageIO::floatToHalf(img[i])));
    // round-to-nearest-even half conversion: exact values, ties, overflow, subnormals
    CHECK(ImageIO::floatToHalf(1.0f) == 0x3C00 && ImageIO::floatToHalf(-2.0f) == 0xC000 && ImageIO::floatToHalf(65504.0f) == 0x7BFF);
    CHECK(ImageIO::floatToHalf(65520.0f) == 0x7C00 && ImageIO::floatToHalf(1e-8f) == 0x0000 && ImageIO::floatToHalf(5.9604644775390625e-08f) == 0x0001);
    CHECK(ImageIO::floatToHalf(1.0f + 1.0f / 2048.0f) == 0x3C00 && ImageIO::floatToHalf(1.0f + 3.0f / 2048.0f) == 0x3C02);  // ties to even
    for (uint32_t hbits = 0; hbits < 0x7C00; hbits += 7) CHECK(ImageIO::floatToHalf(ImageIO::halfToFloat(uint16_t(hbits))) == hbits);
    CHECK(!ImageIO::readEXR(dir + "/t.pfm", back, bw, bh));
    return 0;
}

static int testHitGroupDesc() {
    // RtProgram::Desc::addHitGroup(idx, closestHit, anyHit, intersection) with the compiled-in hit-group programs
    RtProgram::Desc d;
    d.addShaderLibrary(kProgressiveRaytracingLibrary, kProgressiveRaytracingLibrarySize,
                       {L"RayGen", L"PrimaryClosestHit", L"PrimaryMiss", L"ShadowClosestHit", L"ShadowAnyHit", L"ShadowMiss"});
    d.addShaderLibrary(kHitGroupProgramsLibrary, kHitGroupProgramsLibrarySize, {L"AnyHitIgnore", L"AnyHitCutout", L"IntersectSphere", L"ProceduralClosestHit"});
    d.setRayGen("RayGen").addMiss(0, "PrimaryMiss").addMiss(1, "ShadowMiss");
    d.addHitGroup(0, "PrimaryClosestHit", "").addHitGroup(1, "ShadowClosestHit", "ShadowAnyHit");
    d.addHitGroup(2, "ProceduralClosestHit", "AnyHitCutout", "IntersectSphere");
    bool threw = false;
    try {
        d.addHitGroup(3, "ProceduralClosestHit", "", "IntersectBox");  // not among the exports listed above
    } catch (const std::logic_error &) {
        threw = true;
    }
    CHECK(threw);
    threw = false;
    try {
        d.addHitGroup(3, "ProceduralClosestHit", "IntersectSphere", "");  // an intersection shader is not an any-hit shader
    } catch (const std::logic_error &) {
        threw = true;
    }
    CHECK(threw);
    threw = false;
    try {
        RtProgram::Desc().addShaderLibrary(kHitGroupProgramsLibrary, kHitGroupProgramsLibrarySize, {L"RayGen"});
    } catch (const std::logic_error &) {
        threw = true;
    }
    CHECK(threw);
    return 0;
}

int main(int argc, char **argv) {
    const std::string dir = argc > 1 ? argv[1] : "/tmp";
    if (testObj(dir) || testPfmAndHalf(dir) || testDds(dir) || testBindingsLayout() || testProgramDesc() || testExr(dir) || testHitGroupDesc()) return 1;
    std::puts("host selftest OK");
    return 0;
}
