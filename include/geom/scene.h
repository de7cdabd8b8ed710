// Scene = element list + top-level BVH + distant lighting
// (reference: geom/scene.h:9-44, geom/scene.cc:6-31).  On top of the reference
// fields the B200 library keeps a side table (not in this class, so the layout is
// unchanged) that maps a finalized Scene to its flattened GPU copy.
#pragma once

#include "raylib_types.h"
#include "geom/hit.h"

class Scene
{
public:
	Scene();
	~Scene();

	void AddSceneElement(Hitable* hitable);

	void SetSkyPanorama(ImageHandle skyImage) { skyPanorama = skyImage; }
	void SetSunIlluminance(const vec3& illuminance) { sunIlluminance = illuminance; }
	void SetSunDirection(const vec3& direction) { sunDirection = normalize(direction); }

	BVHNode* Finalize();

	ImageHandle GetSkyPanorama() const { return skyPanorama; }
	void GetSun(vec3& outIlluminance, vec3& outDirection) const
	{
		outIlluminance = sunIlluminance;
		outDirection = sunDirection;
	}
	const BVHNode* GetAccelStruct() const { return accelStruct; }

private:
	friend struct RtSceneFlattener;

	HitableList hitableList;
	BVHNode* accelStruct = nullptr;
	ImageHandle skyPanorama = 0;
	vec3 sunIlluminance;
	vec3 sunDirection;
	bool bFinalized = false;
};
