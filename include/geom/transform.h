// Yaw/pitch/roll rotation and TRS transform applied to mesh vertices on the host
// (reference: geom/transform.h:8-61, geom/transform.cc:47-94).  Rotator::rotate is
// also what the miss shader applies to the sky lookup direction (renderer.cc:166-168);
// the device copy of that matrix is built from this very function.
#pragma once

#include "raylib_types.h"
#include "core/int_types.h"
#include "core/vec3.h"
#include <vector>

struct Rotator
{
	static Rotator directionToYawPitch(const vec3& dir);
	vec3 toDirection() const;
	RAYLIB_API vec3 rotate(const vec3& position) const;

	Rotator() : yaw(0.0f), pitch(0.0f), roll(0.0f) {}
	Rotator(float inYaw, float inPitch, float inRoll) : yaw(inYaw), pitch(inPitch), roll(inRoll) {}

	float yaw;    // degrees
	float pitch;
	float roll;
};

class Transform
{
public:
	Transform() { Init(vec3(0.0f), Rotator(), vec3(1.0f)); }

	RAYLIB_API void Init(const vec3& inLocation, const Rotator& inRotation, const vec3& inScale);
	void SetLocation(const vec3& inLocation) { Init(inLocation, rotation, scale); }
	void SetScale(const vec3& inScale) { Init(location, rotation, inScale); }

	RAYLIB_API void TransformVectors(std::vector<vec3>& inoutVectors) const;
	void TransformVectors(const std::vector<vec3>& inVectors, std::vector<vec3>& outVectors) const;

private:
	vec3 location;
	Rotator rotation;
	vec3 scale;
};
