// DEM/HostSideHelpers.hpp -- the float3/float4 helpers reference demo scripts rely on
// (counterpart of src/kernel/CUDAMathHelpers.cuh + src/DEM/HostSideHelpers.hpp of the reference, host side only).
#pragma once
#include <algorithm>
#include <cassert>
#include <cctype>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <numeric>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <unordered_map>
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
/// Quaternion product Q1 * Q2, float4 = (x, y, z, w) (src/DEM/HostSideHelpers.hpp:331-338)
inline float4 hostHamiltonProduct(const float4& p, const float4& q) {
    const float3 pv = make_float3(p.x, p.y, p.z), qv = make_float3(q.x, q.y, q.z);
    const float3 v = p.w * qv + q.w * pv + cross(pv, qv);
    return make_float4(v.x, v.y, v.z, p.w * q.w - dot(pv, qv));
}
/// The orientation `quat` turned further by theta about a unit axis given in the global frame (:341-346)
inline float4 RotateQuat(const float4& quat, const float3& axis, const float& theta) {
    return hostHamiltonProduct(QuatFromAxisAngle(axis, theta), quat);
}
/// Rodrigues' rotation of a vector by theta about a unit axis (:349-353); the trigonometry is done in double like there
inline float3 Rodrigues(const float3& vec, const float3& axis, const float& theta) {
    const double c = std::cos((double)theta), s = std::sin((double)theta);
    const float3 side = cross(axis, vec);
    const float along = dot(axis, vec);
    return make_float3((float)(vec.x * c + side.x * s + axis.x * along * (1. - c)),
                       (float)(vec.y * c + side.y * s + axis.y * along * (1. - c)),
                       (float)(vec.z * c + side.z * s + axis.z * along * (1. - c)));
}
/// Small generic helpers demo scripts call (src/DEM/HostSideHelpers.hpp:69-135, 252-260, 355-361, 553-583)
template <typename T1>
inline int sign_func(const T1& val) {
    return (val > T1(0) ? 1 : 0) - (val < T1(0) ? 1 : 0);
}
template <typename T1>
inline T1 vector_sum(const std::vector<T1>& vect) {
    T1 total = T1(0);
    for (const T1& v : vect) total += v;
    return total;
}
template <typename T>
inline bool isBetween(const T& x, const T& L, const T& U) {
    return !(x < L) && !(x > U);
}
inline bool isBetween(const float3& coord, const float3& L, const float3& U) {
    return isBetween(coord.x, L.x, U.x) && isBetween(coord.y, L.y, U.y) && isBetween(coord.z, L.z, U.z);
}
inline std::string str_to_upper(const std::string& input) {
    std::string out = input;
    for (char& ch : out) ch = (char)std::toupper((unsigned char)ch);
    return out;
}
/// the elements of vec whose flag is not set
template <typename T1>
inline std::vector<T1> hostRemoveElem(const std::vector<T1>& vec, const std::vector<bool>& flags) {
    std::vector<T1> kept;
    for (size_t i = 0; i < vec.size(); i++)
        if (!flags.at(i)) kept.push_back(vec[i]);
    return kept;
}
template <typename T1>
inline std::vector<T1> hostSort(std::vector<T1> input) {
    std::sort(input.begin(), input.end());
    return input;
}
template <typename T1>
inline bool check_exist(const std::set<T1>& the_set, const T1& key) {
    return the_set.count(key) != 0;
}
template <typename T1>
inline bool check_exist(const std::vector<T1>& vec, const T1& key) {
    for (const T1& v : vec)
        if (v == key) return true;
    return false;
}
template <typename T1, typename T2>
inline bool check_exist(const std::unordered_map<T1, T2>& map, const T1& key) {
    return map.count(key) != 0;
}
inline std::vector<std::string> parse_string_line(const std::string& in_str, const char separator = ',') {
    std::vector<std::string> fields(1);
    for (const char ch : in_str) {
        if (ch == separator) fields.emplace_back();
        else fields.back().push_back(ch);
    }
    return fields;
}
/// {x, y, z} / {x, y, z, w} as std::vector, the form the Python-facing getters of the reference return
inline std::vector<float> Real3ToVec(const float3& v) { return {v.x, v.y, v.z}; }
inline std::vector<float> Real4ToVec(const float4& q) { return {q.x, q.y, q.z, q.w}; }
inline std::vector<std::vector<float>> Real3VectorToVecOfVec(const std::vector<float3>& in) {
    std::vector<std::vector<float>> out;
    out.reserve(in.size());
    for (const float3& v : in) out.push_back(Real3ToVec(v));
    return out;
}
inline std::vector<std::vector<float>> Real4VectorToVecOfVec(const std::vector<float4>& in) {
    std::vector<std::vector<float>> out;
    out.reserve(in.size());
    for (const float4& q : in) out.push_back(Real4ToVec(q));
    return out;
}
/// Vector forms of the two frame transforms (positions and translations {x, y, z}, rotation {x, y, z, w}; :585-633)
inline std::vector<double> FrameTransformLocalToGlobal(const std::vector<double>& pos, const std::vector<double>& vec,
                                                       const std::vector<double>& rot_Q) {
    double3 p = make_double3(pos.at(0), pos.at(1), pos.at(2));
    const double3 shift = make_double3(vec.at(0), vec.at(1), vec.at(2));
    const double4 q = make_double4(rot_Q.at(0), rot_Q.at(1), rot_Q.at(2), rot_Q.at(3));
    applyFrameTransformLocalToGlobal(p, shift, q);
    return {p.x, p.y, p.z};
}
inline std::vector<double> FrameTransformGlobalToLocal(const std::vector<double>& pos, const std::vector<double>& vec,
                                                       const std::vector<double>& rot_Q) {
    double3 p = make_double3(pos.at(0), pos.at(1), pos.at(2));
    const double3 shift = make_double3(vec.at(0), vec.at(1), vec.at(2));
    const double4 q = make_double4(rot_Q.at(0), rot_Q.at(1), rot_Q.at(2), rot_Q.at(3));
    applyFrameTransformGlobalToLocal(p, shift, q);
    return {p.x, p.y, p.z};
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
