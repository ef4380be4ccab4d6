// ImageIO.h — image files of the headless host: PFM output (the swap chain of the reference becomes a file) and
// a minimal DDS cube-map reader standing in for DirectXTK12's CreateDDSTextureFromFile
// (src/ProgressiveRaytracingPipeline.cpp:114-115) for R16G16B16A16_FLOAT / R32G32B32A32_FLOAT cubes.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace ImageIO {

// Writes the RGB channels of a width x height RGBA fp32 image as a little-endian colour PFM ("PF", bottom row first).
bool writePFM(const std::string &path, const float *rgba, uint32_t width, uint32_t height);
// Reads a colour PFM into RGBA (alpha = 1).
bool readPFM(const std::string &path, std::vector<float> &rgba, uint32_t &width, uint32_t &height);
// Reads mip 0 of the six faces of a DDS cube map (DX10 header, DXGI_FORMAT 10 = R16G16B16A16_FLOAT or 2 =
// R32G32B32A32_FLOAT) into 6 x size x size RGBA fp32, face order +X,-X,+Y,-Y,+Z,-Z.
bool readDDSCube(const std::string &path, std::vector<float> &texels, uint32_t &size);
// The procedural sky used when no DDS is given (same function as dxrexperiments_b200/scenes.py:sky_cube).
void proceduralSkyCube(uint32_t size, std::vector<float> &texels);
float halfToFloat(uint16_t h);

}  // namespace ImageIO
