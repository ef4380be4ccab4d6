// RtPrefix.h — common includes and the small math vocabulary of the headless host layer.
//
// The reference's RtPrefix.h pulls in d3d12.h / DirectXMath / the Fallback Layer COM headers
// (libs/DXRFramework/RtPrefix.h:7-15).  Here the only dependency is the C ABI of librt_core.so
// (include/rt_core.h); DirectXMath's XMFLOAT* / XMMATRIX are replaced by plain structs with the same
// member names so that pipeline code reads like the reference's.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../include/rt_core.h"

typedef uint32_t UINT;
typedef uint64_t UINT64;

namespace DirectX {
struct XMFLOAT2 { float x, y; };
struct XMFLOAT3 { float x, y, z; };
struct XMFLOAT4 { float x, y, z, w; };
// Row-major 4x4, row-vector convention (v' = v * M) like DirectXMath.
struct XMMATRIX {
    float m[4][4];
};
inline XMMATRIX XMMatrixIdentity() {
    XMMATRIX r{};
    for (int i = 0; i < 4; ++i) r.m[i][i] = 1.0f;
    return r;
}
inline XMMATRIX XMMatrixTranspose(const XMMATRIX &a) {
    XMMATRIX r{};
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) r.m[i][j] = a.m[j][i];
    return r;
}
inline XMMATRIX XMMatrixTranslation(float x, float y, float z) {
    XMMATRIX r = XMMatrixIdentity();
    r.m[3][0] = x, r.m[3][1] = y, r.m[3][2] = z;
    return r;
}
inline XMMATRIX XMMatrixScaling(float x, float y, float z) {
    XMMATRIX r = XMMatrixIdentity();
    r.m[0][0] = x, r.m[1][1] = y, r.m[2][2] = z;
    return r;
}
inline XMMATRIX XMMatrixRotationY(float a) {
    XMMATRIX r = XMMatrixIdentity();
    const float c = cosf(a), s = sinf(a);
    r.m[0][0] = c, r.m[0][2] = -s, r.m[2][0] = s, r.m[2][2] = c;
    return r;
}
inline XMMATRIX XMMatrixMultiply(const XMMATRIX &a, const XMMATRIX &b) {
    XMMATRIX r{};
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            float s = 0.0f;
            for (int k = 0; k < 4; ++k) s += a.m[i][k] * b.m[k][j];
            r.m[i][j] = s;
        }
    return r;
}
inline XMFLOAT4 XMVector4Transform(const XMFLOAT4 &v, const XMMATRIX &M) {
    return XMFLOAT4{v.x * M.m[0][0] + v.y * M.m[1][0] + v.z * M.m[2][0] + v.w * M.m[3][0],
                    v.x * M.m[0][1] + v.y * M.m[1][1] + v.z * M.m[2][1] + v.w * M.m[3][1],
                    v.x * M.m[0][2] + v.y * M.m[1][2] + v.z * M.m[2][2] + v.w * M.m[3][2],
                    v.x * M.m[0][3] + v.y * M.m[1][3] + v.z * M.m[2][3] + v.w * M.m[3][3]};
}
}  // namespace DirectX

namespace DXRFramework {

// rt_core status -> exception, the analogue of ThrowIfFailed (libs/DXRFramework/Helpers/DirectXRaytracingHelper.h).
inline void ThrowIfFailed(int status, const char *what = "rt_core") {
    if (status != RT_OK) throw std::runtime_error(std::string(what) + ": " + rt_last_error());
}
inline void ThrowIfFalse(bool cond, const char *what) {
    if (!cond) throw std::logic_error(what);
}

}  // namespace DXRFramework
