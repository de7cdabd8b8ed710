// Scene = element list + top-level BVH + distant lighting
// (reference: geom/scene.h:9-44, geom/scene.cc:6-31).  On top of the reference
// fields the B200 library keeps a side table (not in this class, so the layout is
// unchanged) that maps a finalized Scene to its flattened GPU copy.
#pragma once
#include "geom/hit.h"
#include "raylib_types.h"

class Scene
{
	friend struct RtSceneFlattener;

	// layout as in the reference: the elements, the BVH Finalize() builds over them, then the distant lighting
	// (equirectangular sky image handle, sun illuminance + normalised direction) and the "no more elements" flag
	HitableList hitableList;
	BVHNode*    accelStruct = nullptr;
	ImageHandle skyPanorama = 0;
	vec3        sunIlluminance, sunDirection;
	bool        bFinalized = false;

public:
	Scene(); ~Scene();

	// elements are owned by the caller; none may be added once Finalize() has built the top-level BVH
	void     AddSceneElement(Hitable* hitable);
	BVHNode* Finalize();
	const BVHNode* GetAccelStruct() const { return accelStruct; }

	void SetSunDirection(const vec3& direction)     { sunDirection = normalize(direction); }
	void SetSunIlluminance(const vec3& illuminance) { sunIlluminance = illuminance; }
	void SetSkyPanorama(ImageHandle skyImage)       { skyPanorama = skyImage; }
	void GetSun(vec3& outIlluminance, vec3& outDirection) const { outIlluminance = sunIlluminance; outDirection = sunDirection; }
	ImageHandle GetSkyPanorama() const { return skyPanorama; }
};
