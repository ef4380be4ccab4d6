// Camera.h — the subset of MiniEngine's Math::Camera the pipelines use (libs/MiniEngine/Camera.h:88-156,
// Camera.cpp:19-36): eye/at/up, vertical FOV (default pi/4), aspect ratio, and a view-projection matrix that
// only serves to detect camera motion (hasCameraMoved, src/ProgressiveRaytracingPipeline.cpp:170-175).
#pragma once
#include <cmath>

#include "../DXRFramework/RtPrefix.h"

namespace Math {

struct Vector3 {
    float x, y, z;
};
inline Vector3 operator-(Vector3 a, Vector3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline Vector3 operator*(Vector3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline float Dot(Vector3 a, Vector3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline Vector3 Cross(Vector3 a, Vector3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline float Length(Vector3 a) { return sqrtf(Dot(a, a)); }
inline Vector3 Normalize(Vector3 a) { return a * (1.0f / Length(a)); }

class Camera {
public:
    Camera() { SetEyeAtUp({0, 0, 0}, {0, 0, -1}, {0, 1, 0}); }
    void SetEyeAtUp(Vector3 eye, Vector3 at, Vector3 up) {
        mPosition = eye;
        // BaseCamera::SetLookDirection (Camera.cpp:19-36)
        Vector3 forward = at - eye;
        float lenSq = Dot(forward, forward);
        forward = lenSq < 0.000001f ? Vector3{0, 0, -1} : forward * (1.0f / sqrtf(lenSq));
        Vector3 right = Cross(forward, up);
        float rl = Dot(right, right);
        right = rl < 0.000001f ? Vector3{1, 0, 0} : right * (1.0f / sqrtf(rl));
        mForward = forward;
        mRight = right;
        mUp = Cross(right, forward);
    }
    void SetAspectRatio(float heightOverWidthInverse) { mAspect = heightOverWidthInverse; }  // width / height, as the app passes it
    void SetFOV(float verticalFovRadians) { mFov = verticalFovRadians; }
    void SetZRange(float n, float f) { mNear = n, mFar = f; }

    Vector3 GetPosition() const { return mPosition; }
    Vector3 GetForwardVec() const { return mForward; }
    Vector3 GetUpVec() const { return mUp; }
    Vector3 GetRightVec() const { return mRight; }
    float GetFOV() const { return mFov; }
    float GetAspectRatio() const { return mAspect; }

    // 16 floats that change whenever the view changes (stand-in for GetViewProjMatrix()).
    void GetViewProjSignature(float out[16]) const {
        const float v[16] = {mPosition.x, mPosition.y, mPosition.z, mFov, mForward.x, mForward.y, mForward.z, mAspect,
                             mUp.x, mUp.y, mUp.z, mNear, mRight.x, mRight.y, mRight.z, mFar};
        for (int i = 0; i < 16; ++i) out[i] = v[i];
    }

private:
    Vector3 mPosition{0, 0, 0}, mForward{0, 0, -1}, mUp{0, 1, 0}, mRight{1, 0, 0};
    float mFov = 3.14159265358979f / 4.0f, mAspect = 16.0f / 9.0f, mNear = 1.0f, mFar = 1000.0f;
};

}  // namespace Math
