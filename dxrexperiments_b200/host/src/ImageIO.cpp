// ImageIO.cpp — PFM writer/reader, DDS cube reader, procedural sky (see ImageIO.h).
#include "../include/ImageIO.h"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>

namespace ImageIO {

bool writePFM(const std::string &path, const float *rgba, uint32_t width, uint32_t height) {
    FILE *f = std::fopen(path.c_str(), "wb");
    if (!f) return false;
    std::fprintf(f, "PF\n%u %u\n-1.0\n", width, height);
    std::vector<float> row(size_t(width) * 3);
    for (uint32_t y = 0; y < height; ++y) {
        const float *src = rgba + size_t(height - 1 - y) * width * 4;  // PFM stores the bottom row first
        for (uint32_t x = 0; x < width; ++x) {
            row[3 * x + 0] = src[4 * x + 0];
            row[3 * x + 1] = src[4 * x + 1];
            row[3 * x + 2] = src[4 * x + 2];
        }
        std::fwrite(row.data(), sizeof(float), row.size(), f);
    }
    std::fclose(f);
    return true;
}

bool readPFM(const std::string &path, std::vector<float> &rgba, uint32_t &width, uint32_t &height) {
    FILE *f = std::fopen(path.c_str(), "rb");
    if (!f) return false;
    char tag[3] = {0, 0, 0};
    float scale = 0;
    if (std::fscanf(f, "%2s %u %u %f", tag, &width, &height, &scale) != 4 || std::strcmp(tag, "PF") != 0 || scale >= 0) {
        std::fclose(f);
        return false;
    }
    std::fgetc(f);  // the single whitespace after the header
    rgba.assign(size_t(width) * height * 4, 1.0f);
    std::vector<float> row(size_t(width) * 3);
    for (uint32_t y = 0; y < height; ++y) {
        if (std::fread(row.data(), sizeof(float), row.size(), f) != row.size()) {
            std::fclose(f);
            return false;
        }
        float *dst = rgba.data() + size_t(height - 1 - y) * width * 4;
        for (uint32_t x = 0; x < width; ++x) {
            dst[4 * x + 0] = row[3 * x + 0];
            dst[4 * x + 1] = row[3 * x + 1];
            dst[4 * x + 2] = row[3 * x + 2];
        }
    }
    std::fclose(f);
    return true;
}

float halfToFloat(uint16_t h) {
    const uint32_t sign = uint32_t(h & 0x8000u) << 16;
    uint32_t exp = (h >> 10) & 0x1Fu, man = h & 0x3FFu, bits;
    if (exp == 0) {
        if (man == 0) bits = sign;
        else {  // subnormal half -> normal float
            exp = 127 - 15 + 1;
            while (!(man & 0x400u)) {
                man <<= 1;
                --exp;
            }
            bits = sign | (exp << 23) | ((man & 0x3FFu) << 13);
        }
    } else if (exp == 31) {
        bits = sign | 0x7F800000u | (man << 13);
    } else {
        bits = sign | ((exp + 127 - 15) << 23) | (man << 13);
    }
    float out;
    std::memcpy(&out, &bits, 4);
    return out;
}

bool readDDSCube(const std::string &path, std::vector<float> &texels, uint32_t &size) {
    std::ifstream in(path, std::ios::binary);
    if (!in) return false;
    uint32_t hdr[32];  // magic + DDS_HEADER (124 bytes)
    in.read(reinterpret_cast<char *>(hdr), 128);
    if (!in || hdr[0] != 0x20534444u /* "DDS " */ || hdr[1] != 124) return false;
    const uint32_t height = hdr[3], width = hdr[4], mipCount = hdr[7] ? hdr[7] : 1;
    const uint32_t fourCC = hdr[21];
    if (fourCC != 0x30315844u /* "DX10" */ || width != height) return false;
    uint32_t dx10[5];  // dxgiFormat, resourceDimension, miscFlag, arraySize, miscFlags2
    in.read(reinterpret_cast<char *>(dx10), 20);
    if (!in) return false;
    const uint32_t format = dx10[0];
    const bool isCube = (dx10[2] & 0x4u) != 0;  // DDS_RESOURCE_MISC_TEXTURECUBE
    if (!isCube || (format != 10 && format != 2)) return false;
    const uint32_t bpp = format == 10 ? 8 : 16;
    size = width;
    texels.assign(size_t(6) * size * size * 4, 0.0f);
    for (int face = 0; face < 6; ++face) {
        std::vector<uint8_t> mip0(size_t(size) * size * bpp);
        in.read(reinterpret_cast<char *>(mip0.data()), mip0.size());
        if (!in) return false;
        float *dst = texels.data() + size_t(face) * size * size * 4;
        if (format == 10) {
            const uint16_t *h = reinterpret_cast<const uint16_t *>(mip0.data());
            for (size_t i = 0; i < size_t(size) * size * 4; ++i) dst[i] = halfToFloat(h[i]);
        } else {
            std::memcpy(dst, mip0.data(), mip0.size());
        }
        uint64_t skip = 0;  // the remaining mips of this face
        for (uint32_t m = 1; m < mipCount; ++m) {
            const uint64_t s = (size >> m) ? (size >> m) : 1;
            skip += s * s * bpp;
        }
        in.seekg(skip, std::ios::cur);
    }
    return true;
}

void proceduralSkyCube(uint32_t size, std::vector<float> &texels) {
    texels.assign(size_t(6) * size * size * 4, 1.0f);
    double sun[3] = {0.4, 0.7, 0.6};
    const double sl = std::sqrt(sun[0] * sun[0] + sun[1] * sun[1] + sun[2] * sun[2]);
    for (double &c : sun) c /= sl;
    for (uint32_t face = 0; face < 6; ++face)
        for (uint32_t row = 0; row < size; ++row)
            for (uint32_t col = 0; col < size; ++col) {
                const double u = (col + 0.5) / size * 2 - 1, v = (row + 0.5) / size * 2 - 1;
                double d[3];
                switch (face) {
                    case 0: d[0] = 1, d[1] = -v, d[2] = -u; break;
                    case 1: d[0] = -1, d[1] = -v, d[2] = u; break;
                    case 2: d[0] = u, d[1] = 1, d[2] = v; break;
                    case 3: d[0] = u, d[1] = -1, d[2] = -v; break;
                    case 4: d[0] = u, d[1] = -v, d[2] = 1; break;
                    default: d[0] = -u, d[1] = -v, d[2] = -1; break;
                }
                const double l = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
                for (double &c : d) c /= l;
                const double up = d[1] * 0.5 + 0.5;
                double dotSun = d[0] * sun[0] + d[1] * sun[1] + d[2] * sun[2];
                dotSun = dotSun < 0 ? 0 : (dotSun > 1 ? 1 : dotSun);
                const double lobe = std::pow(dotSun, 64.0);
                float *t = texels.data() + ((size_t(face) * size + row) * size + col) * 4;
                t[0] = float(0.25 + 0.35 * up + lobe * 6.0);
                t[1] = float(0.3 + 0.45 * up + lobe * 5.0);
                t[2] = float(0.35 + 0.65 * up + lobe * 3.5);
                t[3] = 1.0f;
            }
}

}  // namespace ImageIO
