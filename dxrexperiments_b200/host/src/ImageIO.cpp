// ImageIO.cpp — PFM writer/reader, DDS cube reader, procedural sky (see ImageIO.h).
#include "../include/ImageIO.h"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iterator>
#include <string>
#include <vector>

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

uint16_t floatToHalf(float f) {  // round to nearest even; overflow -> inf, NaN kept
    uint32_t x;
    std::memcpy(&x, &f, 4);
    const uint32_t sign = (x >> 16) & 0x8000u;
    const int32_t exp = int32_t((x >> 23) & 0xFFu) - 127 + 15;
    uint32_t man = x & 0x7FFFFFu;
    if (((x >> 23) & 0xFFu) == 0xFFu) return uint16_t(sign | 0x7C00u | (man ? 0x200u : 0u));
    if (exp >= 31) return uint16_t(sign | 0x7C00u);
    if (exp <= 0) {
        if (exp < -10) return uint16_t(sign);
        man |= 0x800000u;
        const uint32_t shift = uint32_t(14 - exp);
        uint32_t h = man >> shift;
        const uint32_t rem = man & ((1u << shift) - 1u), halfway = 1u << (shift - 1);
        if (rem > halfway || (rem == halfway && (h & 1u))) ++h;
        return uint16_t(sign | h);
    }
    uint32_t h = (uint32_t(exp) << 10) | (man >> 13);
    const uint32_t rem = man & 0x1FFFu;
    if (rem > 0x1000u || (rem == 0x1000u && (h & 1u))) ++h;  // may carry into the exponent: still the right value
    return uint16_t(sign | h);
}

namespace {
void putAttr(std::vector<uint8_t> &o, const char *name, const char *type, const void *data, uint32_t size) {
    o.insert(o.end(), name, name + std::strlen(name) + 1);
    o.insert(o.end(), type, type + std::strlen(type) + 1);
    const uint8_t *s = reinterpret_cast<const uint8_t *>(&size);
    o.insert(o.end(), s, s + 4);
    const uint8_t *d = static_cast<const uint8_t *>(data);
    o.insert(o.end(), d, d + size);
}
}  // namespace

bool writeEXR(const std::string &path, const float *rgba, uint32_t width, uint32_t height, bool half) {
    if (width == 0 || height == 0) return false;
    std::vector<uint8_t> hdr;
    const uint32_t magic = 20000630u, version = 2u;
    hdr.insert(hdr.end(), reinterpret_cast<const uint8_t *>(&magic), reinterpret_cast<const uint8_t *>(&magic) + 4);
    hdr.insert(hdr.end(), reinterpret_cast<const uint8_t *>(&version), reinterpret_cast<const uint8_t *>(&version) + 4);
    std::vector<uint8_t> ch;  // chlist: name, pixel type (1 HALF / 2 FLOAT), pLinear + 3 reserved, x / y sampling
    for (const char *name : {"A", "B", "G", "R"}) {
        ch.push_back(uint8_t(name[0])), ch.push_back(0);
        const int32_t rec[4] = {half ? 1 : 2, 0, 1, 1};
        ch.insert(ch.end(), reinterpret_cast<const uint8_t *>(rec), reinterpret_cast<const uint8_t *>(rec) + 16);
    }
    ch.push_back(0);
    putAttr(hdr, "channels", "chlist", ch.data(), uint32_t(ch.size()));
    const uint8_t noCompression = 0, increasingY = 0;
    putAttr(hdr, "compression", "compression", &noCompression, 1);
    const int32_t window[4] = {0, 0, int32_t(width) - 1, int32_t(height) - 1};
    putAttr(hdr, "dataWindow", "box2i", window, 16);
    putAttr(hdr, "displayWindow", "box2i", window, 16);
    putAttr(hdr, "lineOrder", "lineOrder", &increasingY, 1);
    const float one = 1.0f, centre[2] = {0.0f, 0.0f};
    putAttr(hdr, "pixelAspectRatio", "float", &one, 4);
    putAttr(hdr, "screenWindowCenter", "v2f", centre, 8);
    putAttr(hdr, "screenWindowWidth", "float", &one, 4);
    hdr.push_back(0);
    const uint64_t sample = half ? 2 : 4, lineBytes = uint64_t(width) * 4 * sample, chunk = 8 + lineBytes;
    FILE *f = std::fopen(path.c_str(), "wb");
    if (!f) return false;
    std::fwrite(hdr.data(), 1, hdr.size(), f);
    const uint64_t first = hdr.size() + 8ull * height;
    for (uint32_t y = 0; y < height; ++y) {
        const uint64_t off = first + chunk * y;
        std::fwrite(&off, 8, 1, f);
    }
    std::vector<uint8_t> line(lineBytes);
    const int order[4] = {3, 2, 1, 0};  // A, B, G, R planes of the scan line
    for (uint32_t y = 0; y < height; ++y) {
        const float *src = rgba + size_t(y) * width * 4;
        for (int c = 0; c < 4; ++c)
            for (uint32_t x = 0; x < width; ++x) {
                const float v = src[4 * x + order[c]];
                if (half) {
                    const uint16_t hv = floatToHalf(v);
                    std::memcpy(line.data() + (size_t(c) * width + x) * 2, &hv, 2);
                } else {
                    std::memcpy(line.data() + (size_t(c) * width + x) * 4, &v, 4);
                }
            }
        const int32_t yy = int32_t(y), size = int32_t(lineBytes);
        std::fwrite(&yy, 4, 1, f);
        std::fwrite(&size, 4, 1, f);
        std::fwrite(line.data(), 1, line.size(), f);
    }
    const bool ok = std::ferror(f) == 0;
    std::fclose(f);
    return ok;
}

bool readEXR(const std::string &path, std::vector<float> &rgba, uint32_t &width, uint32_t &height) {
    std::ifstream in(path, std::ios::binary);
    if (!in) return false;
    std::vector<uint8_t> d((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
    if (d.size() < 8) return false;
    uint32_t magic, version;
    std::memcpy(&magic, d.data(), 4), std::memcpy(&version, d.data() + 4, 4);
    if (magic != 20000630u || (version & 0xFFu) != 2u || (version & 0x1E00u)) return false;  // single-part scan-line files only
    size_t p = 8;
    int pixelType = -1, nChannels = 0;
    int32_t win[4] = {0, 0, -1, -1};
    uint8_t compression = 255;
    std::string names;
    while (p < d.size() && d[p] != 0) {
        const std::string name(reinterpret_cast<const char *>(d.data() + p));
        p += name.size() + 1;
        const std::string type(reinterpret_cast<const char *>(d.data() + p));
        p += type.size() + 1;
        uint32_t size;
        std::memcpy(&size, d.data() + p, 4);
        p += 4;
        if (p + size > d.size()) return false;
        if (name == "channels") {
            size_t q = p;
            while (d[q] != 0) {
                const std::string cn(reinterpret_cast<const char *>(d.data() + q));
                q += cn.size() + 1;
                int32_t pt;
                std::memcpy(&pt, d.data() + q, 4);
                q += 16;
                if (pixelType >= 0 && pt != pixelType) return false;
                pixelType = pt;
                names += cn;
                ++nChannels;
            }
        } else if (name == "dataWindow") std::memcpy(win, d.data() + p, 16);
        else if (name == "compression") compression = d[p];
        p += size;
    }
    ++p;
    if (compression != 0 || names != "ABGR" || nChannels != 4 || (pixelType != 1 && pixelType != 2)) return false;
    width = uint32_t(win[2] - win[0] + 1), height = uint32_t(win[3] - win[1] + 1);
    const size_t sample = pixelType == 1 ? 2 : 4;
    rgba.assign(size_t(width) * height * 4, 0.0f);
    for (uint32_t y = 0; y < height; ++y) {
        uint64_t off;
        std::memcpy(&off, d.data() + p + 8ull * y, 8);
        if (off + 8 + size_t(width) * 4 * sample > d.size()) return false;
        int32_t yy;
        std::memcpy(&yy, d.data() + off, 4);
        const uint8_t *line = d.data() + off + 8;
        const int order[4] = {3, 2, 1, 0};
        for (int c = 0; c < 4; ++c)
            for (uint32_t x = 0; x < width; ++x) {
                float v;
                if (pixelType == 1) {
                    uint16_t hv;
                    std::memcpy(&hv, line + (size_t(c) * width + x) * 2, 2);
                    v = halfToFloat(hv);
                } else {
                    std::memcpy(&v, line + (size_t(c) * width + x) * 4, 4);
                }
                rgba[(size_t(yy - win[1]) * width + x) * 4 + order[c]] = v;
            }
    }
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
