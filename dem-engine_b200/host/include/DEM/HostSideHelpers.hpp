// DEM/HostSideHelpers.hpp -- the float3/float4 helpers reference demo scripts rely on
// (counterpart of src/kernel/CUDAMathHelpers.cuh + src/DEM/HostSideHelpers.hpp of the reference, host side only).
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <numeric>
#include <stdexcept>
#include <string>
#include <vector>

#include <vector_functions.h>
#include <vector_types.h>

inline float3 make_float3(float s) { return make_float3(s, s, s); }
inline float3 operator+(float3 a, float3 b) { return make_float3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline float3 operator-(float3 a, float3 b) { return make_float3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline float3 operator-(float3 a) { return make_float3(-a.x, -a.y, -a.z); }
inline float3 operator*(float3 a, float s) { return make_float3(a.x * s, a.y * s, a.z * s); }
inline float3 operator*(float s, float3 a) { return make_float3(a.x * s, a.y * s, a.z * s); }
inline float3 operator*(float3 a, float3 b) { return make_float3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline float3 operator/(float3 a, float s) { return make_float3(a.x / s, a.y / s, a.z / s); }
inline void operator+=(float3& a, float3 b) { a.x += b.x; a.y += b.y; a.z += b.z; }
inline void operator-=(float3& a, float3 b) { a.x -= b.x; a.y -= b.y; a.z -= b.z; }
inline void operator*=(float3& a, float s) { a.x *= s; a.y *= s; a.z *= s; }
inline void operator*=(float3& a, double s) { a.x = (float)(a.x * s); a.y = (float)(a.y * s); a.z = (float)(a.z * s); }
inline float dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float3 cross(float3 a, float3 b) { return make_float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
inline float length(float3 v) { return sqrtf(dot(v, v)); }
inline float3 normalize(float3 v) { return v * (1.0f / sqrtf(dot(v, v))); }

namespace deme {
/// Rotate a vector by a unit quaternion given as float4 (x,y,z,w) -- applyOriQToVector3 of the reference
inline float3 Rotate(float3 v, float4 q) {
    const float w = q.w, x = q.x, y = q.y, z = q.z;
    float3 r;
    r.x = (2.0f * (w * w + x * x) - 1.0f) * v.x + (2.0f * (x * y - w * z)) * v.y + (2.0f * (x * z + w * y)) * v.z;
    r.y = (2.0f * (x * y + w * z)) * v.x + (2.0f * (w * w + y * y) - 1.0f) * v.y + (2.0f * (y * z - w * x)) * v.z;
    r.z = (2.0f * (x * z - w * y)) * v.x + (2.0f * (y * z + w * x)) * v.y + (2.0f * (w * w + z * z) - 1.0f) * v.z;
    return r;
}
/// QuatFromAxisAngle (src/DEM/HostSideHelpers.hpp:321-328)
inline float4 QuatFromAxisAngle(const float3& axis, const float& theta) {
    float4 Q;
    Q.x = axis.x * sinf(theta / 2);
    Q.y = axis.y * sinf(theta / 2);
    Q.z = axis.z * sinf(theta / 2);
    Q.w = cosf(theta / 2);
    return Q;
}
/// applyOriQToVector3 (src/kernel/DEMHelperKernels.cuh:161-173): rotate (X, Y, Z) in place by the quaternion (w, x, y, z)
template <typename T1, typename T2>
inline void applyOriQToVector3(T1& X, T1& Y, T1& Z, const T2& Qw, const T2& Qx, const T2& Qy, const T2& Qz) {
    const T1 oldX = X, oldY = Y, oldZ = Z;
    X = ((T2)2.0 * (Qw * Qw + Qx * Qx) - (T2)1.0) * oldX + ((T2)2.0 * (Qx * Qy - Qw * Qz)) * oldY + ((T2)2.0 * (Qx * Qz + Qw * Qy)) * oldZ;
    Y = ((T2)2.0 * (Qx * Qy + Qw * Qz)) * oldX + ((T2)2.0 * (Qw * Qw + Qy * Qy) - (T2)1.0) * oldY + ((T2)2.0 * (Qy * Qz - Qw * Qx)) * oldZ;
    Z = ((T2)2.0 * (Qx * Qz - Qw * Qy)) * oldX + ((T2)2.0 * (Qy * Qz + Qw * Qx)) * oldY + ((T2)2.0 * (Qw * Qw + Qz * Qz) - (T2)1.0) * oldZ;
}
/// Rotate pos by rot_Q, then translate by vec (src/DEM/HostSideHelpers.hpp:563-569)
template <typename T1, typename T2, typename T3>
inline void applyFrameTransformLocalToGlobal(T1& pos, const T2& vec, const T3& rot_Q) {
    applyOriQToVector3(pos.x, pos.y, pos.z, rot_Q.w, rot_Q.x, rot_Q.y, rot_Q.z);
    pos.x += vec.x;
    pos.y += vec.y;
    pos.z += vec.z;
}
/// Translate by -vec, then rotate by the inverse of rot_Q (src/DEM/HostSideHelpers.hpp:608-614)
template <typename T1, typename T2, typename T3>
inline void applyFrameTransformGlobalToLocal(T1& pos, const T2& vec, const T3& rot_Q) {
    pos.x -= vec.x;
    pos.y -= vec.y;
    pos.z -= vec.z;
    applyOriQToVector3(pos.x, pos.y, pos.z, rot_Q.w, -rot_Q.x, -rot_Q.y, -rot_Q.z);
}
/// Decimal rendering with n digits after the point (src/DEM/HostSideHelpers.hpp:637-648)
inline std::string to_string_with_precision(const double a_value, const unsigned int n = 17) {
    std::string out(330 + n, '\0');
    const int len = std::snprintf(&out[0], out.size(), "%.*f", (int)n, a_value);
    if (len < 0 || (size_t)len >= out.size()) throw std::runtime_error("to_string_with_precision: value cannot be rendered");
    out.resize((size_t)len);
    return out;
}
}  // namespace deme
