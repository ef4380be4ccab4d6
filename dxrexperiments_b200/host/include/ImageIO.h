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
// Writes a width x height RGBA fp32 image as an OpenEXR 2 file: single part, scan lines, NO_COMPRESSION, four FLOAT
// channels A, B, G, R, increasing-Y line order.  The reference's render targets are R16G16B16A16_FLOAT
// (src/DXRExperimentsApp.cpp:28); with half = true the channels are stored as HALF (round to nearest even), which
// is that format's precision, with half = false the fp32 accumulation buffer is kept exactly.
bool writeEXR(const std::string &path, const float *rgba, uint32_t width, uint32_t height, bool half = false);
// Reads back what writeEXR wrote (uncompressed scan-line RGBA FLOAT / HALF files).
bool readEXR(const std::string &path, std::vector<float> &rgba, uint32_t &width, uint32_t &height);
uint16_t floatToHalf(float f);
// Reads mip 0 of the six faces of a DDS cube map (DX10 header, DXGI_FORMAT 10 = R16G16B16A16_FLOAT or 2 =
// R32G32B32A32_FLOAT) into 6 x size x size RGBA fp32, face order +X,-X,+Y,-Y,+Z,-Z.
bool readDDSCube(const std::string &path, std::vector<float> &texels, uint32_t &size);
// The procedural sky used when no DDS is given (same function as dxrexperiments_b200/scenes.py:sky_cube).
void proceduralSkyCube(uint32_t size, std::vector<float> &texels);
float halfToFloat(uint16_t h);

}  // namespace ImageIO
