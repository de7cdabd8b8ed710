// vec3 -- three packed floats, the value type of every raylib class
// (reference: core/vec3.h:10-229).  The exact operation order of a few helpers is
// part of the numeric contract with the GPU kernels and the oracle:
//   Normalize():      k = 1/len, then three multiplies
//   v / f:            three true divisions         v /= f: multiply by 1/f
//   cross().y:        -(x1*z2 - z1*x2)
#pragma once

#include "core/int_types.h"
#include "core/assertion.h"
#include <algorithm>

#ifndef _USE_MATH_DEFINES
#define _USE_MATH_DEFINES
#endif
#include <math.h>

struct vec3
{
	float x, y, z;

	vec3() : x(0.0f), y(0.0f), z(0.0f) {}
	vec3(float s) : x(s), y(s), z(s) {}
	vec3(float inX, float inY, float inZ) : x(inX), y(inY), z(inZ) {}

	const vec3& operator+() const { return *this; }
	vec3 operator-() const { return vec3(-x, -y, -z); }

	vec3& operator+=(const vec3& r) { x += r.x; y += r.y; z += r.z; return *this; }
	vec3& operator-=(const vec3& r) { x -= r.x; y -= r.y; z -= r.z; return *this; }
	vec3& operator*=(const vec3& r) { x *= r.x; y *= r.y; z *= r.z; return *this; }
	vec3& operator/=(const vec3& r) { x /= r.x; y /= r.y; z /= r.z; return *this; }
	vec3& operator+=(const float s) { x += s; y += s; z += s; return *this; }
	vec3& operator-=(const float s) { x -= s; y -= s; z -= s; return *this; }
	vec3& operator*=(const float s) { x *= s; y *= s; z *= s; return *this; }
	vec3& operator/=(const float s) { const float k = 1.0f / s; x *= k; y *= k; z *= k; return *this; }

	float operator[](int32 axis) const
	{
		if (axis == 0) return x;
		if (axis == 1) return y;
		if (axis == 2) return z;
		CHECK_NO_ENTRY();
		return NAN;
	}

	float LengthSquared() const { return x * x + y * y + z * z; }
	float Length() const { return sqrtf(x * x + y * y + z * z); }
	void Normalize() { const float k = 1.0f / Length(); x *= k; y *= k; z *= k; }
};

inline bool operator==(const vec3& a, const vec3& b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
inline bool operator!=(const vec3& a, const vec3& b) { return a.x != b.x || a.y != b.y || a.z != b.z; }

inline vec3 operator+(const vec3& a, const vec3& b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline vec3 operator-(const vec3& a, const vec3& b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline vec3 operator*(const vec3& a, const vec3& b) { return vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline vec3 operator/(const vec3& a, const vec3& b) { return vec3(a.x / b.x, a.y / b.y, a.z / b.z); }

inline vec3 operator+(const vec3& a, float s) { return vec3(a.x + s, a.y + s, a.z + s); }
inline vec3 operator+(float s, const vec3& a) { return vec3(a.x + s, a.y + s, a.z + s); }
inline vec3 operator-(const vec3& a, float s) { return vec3(a.x - s, a.y - s, a.z - s); }
inline vec3 operator-(float s, const vec3& a) { return vec3(s - a.x, s - a.y, s - a.z); }
inline vec3 operator*(const vec3& a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
inline vec3 operator*(float s, const vec3& a) { return vec3(a.x * s, a.y * s, a.z * s); }
inline vec3 operator/(const vec3& a, float s) { return vec3(a.x / s, a.y / s, a.z / s); }
inline vec3 operator/(float s, const vec3& a) { return vec3(s / a.x, s / a.y, s / a.z); }

inline float dot(const vec3& a, const vec3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float absDot(const vec3& a, const vec3& b) { return std::abs(a.x * b.x + a.y * b.y + a.z * b.z); }
inline vec3 cross(const vec3& a, const vec3& b)
{
	return vec3(a.y * b.z - a.z * b.y, -(a.x * b.z - a.z * b.x), a.x * b.y - a.y * b.x);
}
inline vec3 normalize(const vec3& v) { vec3 r = v; r.Normalize(); return r; }
inline vec3 mix(const vec3& a, const vec3& b, float t) { return (1.0f - t) * a + t * b; }
inline vec3 reflect(const vec3& v, const vec3& n) { return v - 2.0f * dot(v, n) * n; }
inline bool refract(const vec3& v, const vec3& n, float niOverNt, vec3& outRefracted)
{
	const vec3 unit = normalize(v);
	const float dt = dot(unit, n);
	const float disc = 1.0f - niOverNt * niOverNt * (1.0f - dt * dt);
	if (disc > 0.0f)
	{
		outRefracted = niOverNt * (unit - n * dt) - n * sqrtf(disc);
		return true;
	}
	return false;
}

inline vec3 abs(const vec3& v) { return vec3(std::abs(v.x), std::abs(v.y), std::abs(v.z)); }
inline vec3 min(const vec3& a, const vec3& b) { return vec3(std::min(a.x, b.x), std::min(a.y, b.y), std::min(a.z, b.z)); }
inline vec3 max(const vec3& a, const vec3& b) { return vec3(std::max(a.x, b.x), std::max(a.y, b.y), std::max(a.z, b.z)); }
inline vec3 pow(const vec3& v, float p) { return vec3(powf(v.x, p), powf(v.y, p), powf(v.z, p)); }
inline vec3 saturate(const vec3& v) { return max(vec3(0.0f), min(vec3(1.0f), v)); }
inline bool isnan(const vec3& v) { return std::isnan(v.x) || std::isnan(v.y) || std::isnan(v.z); }
