// oracle/shadow/core/random.h -- TEST INFRASTRUCTURE (not product code).
// Put AHEAD of /root/reference/raylib on the include path so that every
// reference translation unit that says #include "core/random.h" gets this
// deterministic stand-in for raylib/core/random.h:13-73.  Same API surface
// (class RNG{RNG(uint32); Seek; Peek}, the five Random* prototypes); the
// reference's own raylib/core/random.cc is compiled against it unchanged.
// Every RNG instance reads one thread_local counter stream whose definition
// lives in include/rt_rng.h and is shared with the CUDA renderer.
#pragma once

#include "raylib_types.h"
#include "core/int_types.h"
#include "core/vec3.h"
#include "../../../include/rt_rng.h"

#include <random>
#include <vector>
#include <algorithm>

struct OracleRngCtx { uint64_t key; uint32_t ctr; uint64_t draws; };
extern thread_local OracleRngCtx g_oracleRng;

class RNG
{
public:
	RNG(uint32) {}
	inline void Seek(int32) {}
	inline float Peek()
	{
		g_oracleRng.draws++;
		return rt_uniform(g_oracleRng.key, ++g_oracleRng.ctr);
	}
};

RAYLIB_API float Random();
RAYLIB_API vec3 RandomInUnitSphere();
vec3 RandomInHemisphere(const vec3& axis);
vec3 RandomInUnitDisk();
vec3 RandomInCosineHemisphere();
