// Host-side random helpers kept for API parity with the reference
// (core/random.h:68-73).  Clients use Random()/RandomInUnitSphere() while building
// scenes (src/main.cc:927-940).  Unlike the reference's random_device-seeded
// tables, these draw from the counter stream of include/rt_rng.h, so a scene
// generated twice is identical; RaylibB200_SeedHostRandom() re-keys the stream
// of the calling thread.  Rendering itself never calls these: the GPU path keys
// its own stream per (pixel, sample).
#pragma once

#include "raylib_types.h"
#include "core/int_types.h"
#include "core/vec3.h"

// Stateful view of one counter stream (same surface as the reference's RNG class).
class RNG
{
public:
	RAYLIB_API explicit RNG(uint32 nSamples);
	void Seek(int32 ix) { counter = (uint32)ix; }
	RAYLIB_API float Peek();
private:
	uint64 key;
	uint32 counter;
};

RAYLIB_API float Random();
RAYLIB_API vec3 RandomInUnitSphere();          // uniform ON the unit sphere, as in the reference
RAYLIB_API vec3 RandomInHemisphere(const vec3& axis);
RAYLIB_API vec3 RandomInUnitDisk();
RAYLIB_API vec3 RandomInCosineHemisphere();

extern "C" RAYLIB_API void RaylibB200_SeedHostRandom(uint64_t key);
