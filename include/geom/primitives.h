// Scene primitives: Sphere, Cube, Triangle, StaticMesh
// (reference: geom/sphere.h:8-24, geom/cube.h:8-43, geom/triangle.h:8-65,
//  geom/static_mesh.h:10-43).  Field order is ABI.
#pragma once

#include "geom/hit.h"
#include "geom/transform.h"

class BVHNode;

class Sphere : public Hitable
{
public:
	Sphere(const vec3& inCenter, float inRadius, Material* inMaterial)
		: center(inCenter), radius(inRadius), material(inMaterial) {}

	RAYLIB_API virtual bool Hit(const ray& r, float t_min, float t_max, HitResult& result) const;
	RAYLIB_API virtual bool BoundingBox(float t0, float t1, AABB& outBox) const override;

	vec3 center;
	float radius;
	Material* material;
};

// Axis-aligned box that starts translating with `velocity` at `timeStartMove`.
class Cube : public Hitable
{
public:
	static Cube FromMinMaxBounds(const vec3& lo, const vec3& hi, float inTimeStartMove, vec3 inVelocity, Material* inMaterial)
	{
		return Cube(lo, hi, inTimeStartMove, inVelocity, inMaterial);
	}
	static Cube FromOriginAndExtent(const vec3& origin, const vec3& extent, float inTimeStartMove, vec3 inVelocity, Material* inMaterial)
	{
		return Cube(origin - extent, origin + extent, inTimeStartMove, inVelocity, inMaterial);
	}

	Cube(const vec3& lo, const vec3& hi, float inTimeStartMove, vec3 inVelocity, Material* inMaterial)
		: minBounds(lo), maxBounds(hi), timeStartMove(inTimeStartMove), velocity(inVelocity), material(inMaterial) {}
	Cube() : Cube(vec3(0.0f), vec3(0.0f), 0.0f, vec3(0.0f), nullptr) {}

	RAYLIB_API virtual bool Hit(const ray& r, float t_min, float t_max, HitResult& outResult) const;
	RAYLIB_API virtual bool BoundingBox(float t0, float t1, AABB& outBox) const override;

	vec3 minBounds;
	vec3 maxBounds;
	float timeStartMove;
	vec3 velocity;
	Material* material;
};

class Triangle : public Hitable
{
public:
	RAYLIB_API Triangle(
		const vec3& inV0, const vec3& inV1, const vec3& inV2,
		const vec3& inN0, const vec3& inN1, const vec3& inN2,
		Material* inMaterial);

	RAYLIB_API virtual bool Hit(const ray& r, float t_min, float t_max, HitResult& outResult) const;
	RAYLIB_API virtual bool BoundingBox(float t0, float t1, AABB& outBox) const override;

	// Per-vertex surface parameterisation (texture coordinates).
	void SetParameterization(float inS0, float inT0, float inS1, float inT1, float inS2, float inT2)
	{
		s0 = inS0; t0 = inT0; s1 = inS1; t1 = inT1; s2 = inS2; t2 = inT2;
	}
	void GetParameterization(float& outS0, float& outT0, float& outS1, float& outT1, float& outS2, float& outT2) const
	{
		outS0 = s0; outT0 = t0; outS1 = s1; outT1 = t1; outS2 = s2; outT2 = t2;
	}

	RAYLIB_API void GetVertices(vec3& outV0, vec3& outV1, vec3& outV2) const;
	RAYLIB_API void SetVertices(const vec3& inV0, const vec3& inV1, const vec3& inV2);
	RAYLIB_API void GetNormals(vec3& outN0, vec3& outN1, vec3& outN2) const;
	RAYLIB_API void SetNormals(const vec3& inN0, const vec3& inN1, const vec3& inN2);

private:
	friend struct RtSceneFlattener;
	void RefreshDerived();   // face normal + bounds from the vertices

	vec3 v0, v1, v2;
	vec3 n;                  // unit face normal, derived
	vec3 n0, n1, n2;         // shading normals, barycentrically blended
	AABB bounds;
	float s0, t0, s1, t1, s2, t2;
	Material* material;
};

class StaticMesh : public Hitable
{
public:
	RAYLIB_API StaticMesh() {}
	~StaticMesh();

	RAYLIB_API void AddTriangle(const Triangle& triangle);
	RAYLIB_API void SetBounds(const AABB& inBounds);
	RAYLIB_API void CalculateBounds();
	RAYLIB_API void ApplyTransform(const Transform& transform);
	RAYLIB_API void Finalize();              // locks the mesh and builds its BVH

	RAYLIB_API virtual bool Hit(const ray& r, float t_min, float t_max, HitResult& outResult) const override;
	RAYLIB_API virtual bool BoundingBox(float t0, float t1, AABB& outBox) const override;

private:
	friend struct RtSceneFlattener;

	std::vector<Triangle> triangles;
	AABB bounds;
	bool boundsValid = false;
	BVHNode* bvh = nullptr;
	bool bLocked = false;
};
