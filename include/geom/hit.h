// Ray, bounding box, hit record and the Hitable object model that clients use to
// describe a scene (reference: geom/ray.h:5-23, geom/aabb.h:14-67, geom/hit.h:16-89,
// geom/bvh.h:6-23).  Class names, virtual-function order and field order follow the
// reference so that client code -- and client binaries -- built against either
// header set agree on layout (sizes are pinned by static_asserts in
// csrc/host/abi_layout_check.cc).
//
// In this library the objects are a *description*: Raylib_FinalizeScene walks the
// graph and flattens it into GPU records.  Ray queries run on the device only;
// the virtual Hit() entry points exist for ABI compatibility and report an error
// if a client calls them on the host.
#pragma once

#include "raylib_types.h"
#include "core/vec3.h"
#include "core/assertion.h"

#include <limits>
#include <vector>

#define FLOAT_MIN std::numeric_limits<float>::min()
#define FLOAT_MAX std::numeric_limits<float>::max()

class Material;

class ray
{
public:
	ray() : t(0.0f) {}
	ray(const vec3& origin, const vec3& direction, float worldTime) : o(origin), d(direction), t(worldTime) {}
	vec3 at(float s) const { return o + s * d; }

	vec3 o;
	vec3 d;
	float t;   // world time the ray was generated at (motion blur), not a ray parameter
};

class AABB
{
public:
	AABB() {}
	AABB(const vec3& inMin, const vec3& inMax) : minBounds(inMin), maxBounds(inMax) {}

	vec3 minBounds;
	vec3 maxBounds;
};

inline AABB operator+(const AABB& a, const AABB& b)
{
	return AABB(min(a.minBounds, b.minBounds), max(a.maxBounds, b.maxBounds));
}

struct HitResult
{
	float t;
	vec3  p;
	vec3  n;
	float paramU;
	float paramV;
	Material* material;

	RAYLIB_API void BuildOrthonormalBasis();
	RAYLIB_API vec3 LocalToWorld(const vec3& localDirection) const;
	RAYLIB_API vec3 WorldToLocal(const vec3& worldDirection) const;
private:
	vec3 tangent;
	vec3 bitangent;
};

class Hitable
{
public:
	virtual ~Hitable() = default;
	RAYLIB_API virtual bool Hit(const ray& r, float t_min, float t_max, HitResult& outResult) const = 0;
	virtual bool BoundingBox(float t0, float t1, AABB& outBox) const = 0;
};

class HitableList : public Hitable
{
public:
	HitableList() {}
	HitableList(std::vector<Hitable*> inList) : hitables(inList) {}

	RAYLIB_API virtual bool Hit(const ray& r, float t_min, float t_max, HitResult& outResult) const;

	// Exactly the reference's behaviour (geom/hit.h:62-86): the union box of the members is computed into a local and
	// NEVER written to the out-parameter, so a caller that passed a default-constructed AABB -- every caller does --
	// keeps the zero box [0,0,0]..[0,0,0] (AABB() and vec3() zero-initialise, geom/aabb.h:9, core/vec3.h:16).  A raw
	// list added as a scene element is therefore sorted and bounded as a point at the origin by the BVH build; its
	// members are still reached whenever the box of the BVHNode holding the list passes.  Kept as is: the flattener and
	// the GPU traversal reproduce what the reference renders for such a scene.
	virtual bool BoundingBox(float t0, float t1, AABB& outBox) const override
	{
		if (hitables.empty()) return false;
		AABB acc;
		if (!hitables[0]->BoundingBox(t0, t1, acc)) return false;
		for (size_t i = 1; i < hitables.size(); ++i)
		{
			AABB next;
			if (!hitables[i]->BoundingBox(t0, t1, next)) return false;
			acc = acc + next;
		}
		(void)outBox;
		return true;
	}

	std::vector<Hitable*> hitables;
};

// Binary bounding-volume hierarchy node.  The constructor reproduces the
// reference build (geom/bvh.cc:10-80): random split axis, sort by box minimum,
// median split, 1-2 primitives per leaf; the caller's array is reordered in place.
class BVHNode : public Hitable
{
public:
	RAYLIB_API BVHNode(HitableList* list, float t0, float t1);

	RAYLIB_API virtual bool Hit(const ray& r, float tMin, float tMax, HitResult& outResult) const override;
	RAYLIB_API virtual bool BoundingBox(float t0, float t1, AABB& outBox) const override;

private:
	BVHNode(Hitable** list, int32 n, float t0, float t1);
	friend struct RtBvhBuilder;

public:
	Hitable* left = nullptr;
	Hitable* right = nullptr;
	AABB box;
};
