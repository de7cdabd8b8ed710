// Thin-lens camera with a shutter interval (reference: render/camera.h:8-102).
// The host keeps the user parameters and the derived frame; rays are generated on
// the device by the raygen kernel from a copy of the derived block.
#pragma once

#include "core/random.h"
#include "geom/hit.h"
#include <math.h>

class Camera
{
public:
	// Unlike the reference (whose default constructor leaves every member
	// uninitialised, camera.h:14-21) a default camera here is well defined.
	Camera() { Set(vec3(0.0f), vec3(0.0f, 0.0f, -1.0f), 60.0f, 16.0f / 9.0f, 0.0f, 1.0f, 0.0f, 0.0f); }

	Camera(const vec3& inLocation, const vec3& inLookAt,
		float inFovY_degrees, float inAspectWH,
		float inAperture, float inFocalDistance,
		float inBeginTime, float inEndTime)
	{
		Set(inLocation, inLookAt, inFovY_degrees, inAspectWH, inAperture, inFocalDistance, inBeginTime, inEndTime);
	}

	// Recompute the derived frame; call after touching any public field.
	void UpdateInternal()
	{
		lensRadius = aperture * 0.5f;
		timePeriod = endTime - beginTime;

		w = normalize(origin - lookAt);
		vec3 up(0.0f, 1.0f, 0.0f);
		if (dot(w, up) >= 0.9f) up = vec3(1.0f, 0.0f, 0.0f);
		u = normalize(cross(up, w));
		v = cross(w, u);

		const float theta = fovY_degrees * float(3.1415926535897932385) / 180.0f;
		const float halfH = tan(theta / 2.0f);
		const float halfW = aspectWH * halfH;

		top_left = origin - (halfW * focalDistance * u) - (halfH * focalDistance * v) - (focalDistance * w);
		horizontal = 2.0f * halfW * focalDistance * u;
		vertical = 2.0f * halfH * focalDistance * v;
	}

	vec3 origin;
	vec3 lookAt;
	float fovY_degrees;
	float aspectWH;
	float aperture;
	float focalDistance;
	float beginTime, endTime;

private:
	friend struct RtSceneFlattener;
	void Set(const vec3& loc, const vec3& at, float fov, float aspect, float ap, float focal, float t0, float t1)
	{
		origin = loc; lookAt = at; fovY_degrees = fov; aspectWH = aspect;
		aperture = ap; focalDistance = focal; beginTime = t0; endTime = t1;
		UpdateInternal();
	}

	float lensRadius;
	float timePeriod;
	vec3 top_left;
	vec3 horizontal;
	vec3 vertical;
	vec3 u, v, w;
};
