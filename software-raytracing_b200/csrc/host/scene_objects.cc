// scene_objects.cc -- host side of the geometry classes: parameter storage, bounding boxes and
// the BVH *build*.  Ray queries are answered on the device (csrc/device); the virtual Hit()
// entry points below exist for ABI compatibility only and report an error when called.
//
// The BVH build reproduces the reference tree exactly (raylib/geom/bvh.cc:10-80): the split
// axis of the node with pre-order index i is int(U(i+1) * 3) where U is the counter stream keyed
// by the build key (the reference draws Random() once per constructor, in pre-order), elements
// are std::sort-ed by their box minimum on that axis, and the list is halved at n/2.  Sorting
// records {pointer, box} instead of pointers makes the same sequence of comparisons (introsort
// is comparison-driven), hence the same permutation, without two virtual calls per compare;
// independent subtrees are built on separate threads because the axis stream is counter-based.
#include "compact_mesh.h"
#include "geom/hit.h"
#include "geom/primitives.h"
#include "geom/transform.h"
#include "geom/scene.h"
#include "render/material.h"
#include "rt_rng.h"
#include "host_internal.h"

#include <algorithm>
#include <future>
#include <mutex>
#include <unordered_map>

// ---------------------------------------------------------------------------------------------
// HitResult (geom/hit.cc:6-29)

void HitResult::BuildOrthonormalBasis()
{
	vec3 T = (std::abs(n.x) > 0.9f) ? vec3(0.0f, 1.0f, 0.0f) : vec3(1.0f, 0.0f, 0.0f);
	bitangent = normalize(cross(T, n));
	tangent = normalize(cross(n, bitangent));
}

vec3 HitResult::LocalToWorld(const vec3& v) const
{
	return vec3(
		dot(vec3(tangent.x, bitangent.x, n.x), v),
		dot(vec3(tangent.y, bitangent.y, n.y), v),
		dot(vec3(tangent.z, bitangent.z, n.z), v));
}

vec3 HitResult::WorldToLocal(const vec3& v) const
{
	return vec3(dot(v, tangent), dot(v, bitangent), dot(v, n));
}

bool HitableList::Hit(const ray&, float, float, HitResult&) const { return RtHostQueryUnsupported("HitableList::Hit"); }

// ---------------------------------------------------------------------------------------------
// BVH build

struct RtBvhBuilder
{
	struct Item { Hitable* object; AABB box; };

	uint64_t key;
	float t0, t1;
	std::unordered_map<int32, int64_t> nodeCountMemo;
	std::vector<BVHNode*> allocated;     // inner nodes created by this builder (owned by the root)

	// number of BVHNode constructors the reference runs for a list of n elements
	int64_t NodesIn(int32 n)
	{
		if (n <= 2) return 1;
		auto it = nodeCountMemo.find(n);
		if (it != nodeCountMemo.end()) return it->second;
		const int64_t c = 1 + NodesIn(n / 2) + NodesIn(n - n / 2);
		nodeCountMemo[n] = c;
		return c;
	}

	void SortByAxis(Item* items, int32 n, int64_t preorderIndex) const
	{
		const int32 axis = int32(rt_uniform(key, (uint32_t)(preorderIndex + 1)) * 3);
		if (axis == 0) std::sort(items, items + n, [](const Item& l, const Item& r) { return l.box.minBounds.x < r.box.minBounds.x; });
		else if (axis == 1) std::sort(items, items + n, [](const Item& l, const Item& r) { return l.box.minBounds.y < r.box.minBounds.y; });
		else std::sort(items, items + n, [](const Item& l, const Item& r) { return l.box.minBounds.z < r.box.minBounds.z; });
	}

	// Fills `node` (already allocated) from items[0..n).
	void Fill(BVHNode* node, Item* items, int32 n, int64_t preorderIndex, int parallelDepth)
	{
		SortByAxis(items, n, preorderIndex);
		AABB leftBox, rightBox;
		if (n == 1)
		{
			node->left = node->right = items[0].object;
			leftBox = rightBox = items[0].box;
		}
		else if (n == 2)
		{
			node->left = items[0].object; node->right = items[1].object;
			leftBox = items[0].box; rightBox = items[1].box;
		}
		else
		{
			const int32 nl = n / 2, nr = n - n / 2;
			BVHNode* l = AllocateNode();
			BVHNode* r = AllocateNode();
			allocated.push_back(l);
			allocated.push_back(r);
			const int64_t leftIndex = preorderIndex + 1, rightIndex = preorderIndex + 1 + NodesIn(nl);
			if (parallelDepth > 0 && n >= 65536)
			{
				auto task = std::async(std::launch::async, [=]() {
					RtBvhBuilder sub{ key, t0, t1, {}, {} };
					sub.Fill(l, items, nl, leftIndex, parallelDepth - 1);
					return std::move(sub.allocated);
				});
				Fill(r, items + nl, nr, rightIndex, parallelDepth - 1);
				std::vector<BVHNode*> fromLeft = task.get();
				allocated.insert(allocated.end(), fromLeft.begin(), fromLeft.end());
			}
			else
			{
				Fill(l, items, nl, leftIndex, 0);
				Fill(r, items + nl, nr, rightIndex, 0);
			}
			node->left = l; node->right = r;
			leftBox = l->box; rightBox = r->box;
		}
		node->box = leftBox + rightBox;
	}

	static BVHNode* AllocateNode();
};

// Inner nodes belong to the root that built them.  BVHNode has no room for an ownership flag
// (its layout is ABI), so ownership lives in a side table keyed by the root.
namespace
{
	std::mutex g_ownedMutex;
	std::unordered_map<const BVHNode*, std::vector<BVHNode*>> g_ownedInnerNodes;
}

static void RegisterOwnedNodes(const BVHNode* root, std::vector<BVHNode*>&& nodes)
{
	std::lock_guard<std::mutex> lock(g_ownedMutex);
	g_ownedInnerNodes[root] = std::move(nodes);
}

// Deletes the root and every inner node it built; never touches client-owned elements.
static void DestroyBvhTree(BVHNode* root)
{
	if (!root) return;
	std::vector<BVHNode*> nodes;
	{
		std::lock_guard<std::mutex> lock(g_ownedMutex);
		auto it = g_ownedInnerNodes.find(root);
		if (it != g_ownedInnerNodes.end()) { nodes = std::move(it->second); g_ownedInnerNodes.erase(it); }
	}
	for (BVHNode* n : nodes) delete n;
	delete root;
}

// A BVHNode can only be created through its (building) constructors; inner nodes are carved
// from raw storage and filled by the builder.
BVHNode* RtBvhBuilder::AllocateNode()
{
	Hitable* one[1] = { nullptr };
	return new BVHNode(one, 0, 0.0f, 0.0f);
}

BVHNode::BVHNode(HitableList* list, float t0, float t1)
	: BVHNode(list->hitables.data(), (int32)list->hitables.size(), t0, t1)
{
}

BVHNode::BVHNode(Hitable** list, int32 n, float t0, float t1)
{
	if (n <= 0) return;    // builder-internal empty shell (also what an empty scene gets)

	std::vector<RtBvhBuilder::Item> items((size_t)n);
	for (int32 i = 0; i < n; ++i)
	{
		items[i].object = list[i];
		if (!list[i]->BoundingBox(t0, t1, items[i].box)) { CHECK_NO_ENTRY(); }
	}
	RtBvhBuilder builder{ RtGetBvhBuildKey(), t0, t1, {}, {} };
	builder.Fill(this, items.data(), n, 0, 4);
	RegisterOwnedNodes(this, std::move(builder.allocated));
	// the reference sorts the caller's array in place (bvh.cc:46-54); keep that visible side effect
	for (int32 i = 0; i < n; ++i) list[i] = items[i].object;
}

bool BVHNode::Hit(const ray&, float, float, HitResult&) const { return RtHostQueryUnsupported("BVHNode::Hit"); }

bool BVHNode::BoundingBox(float, float, AABB& outBox) const
{
	outBox = box;
	return true;
}

// ---------------------------------------------------------------------------------------------
// Sphere / Cube (geom/sphere.cc:47-52, geom/cube.cc:45-52)

bool Sphere::Hit(const ray&, float, float, HitResult&) const { return RtHostQueryUnsupported("Sphere::Hit"); }

bool Sphere::BoundingBox(float, float, AABB& outBox) const
{
	const vec3 extent(radius, radius, radius);
	outBox = AABB(center - extent, center + extent);
	return true;
}

bool Cube::Hit(const ray&, float, float, HitResult&) const { return RtHostQueryUnsupported("Cube::Hit"); }

bool Cube::BoundingBox(float t0, float t1, AABB& outBox) const
{
	const vec3 shift0 = velocity * std::max(0.0f, t0 - timeStartMove);
	const vec3 shift1 = velocity * std::max(0.0f, t1 - timeStartMove);
	outBox = AABB(minBounds + shift0, maxBounds + shift0) + AABB(minBounds + shift1, maxBounds + shift1);
	return true;
}

// ---------------------------------------------------------------------------------------------
// Triangle (geom/triangle.cc:4-16, :60-94)

Triangle::Triangle(const vec3& inV0, const vec3& inV1, const vec3& inV2,
                   const vec3& inN0, const vec3& inN1, const vec3& inN2, Material* inMaterial)
	: v0(inV0), v1(inV1), v2(inV2), n0(inN0), n1(inN1), n2(inN2)
	, s0(0.0f), t0(0.0f), s1(0.0f), t1(0.0f), s2(0.0f), t2(0.0f), material(inMaterial)
{
	RefreshDerived();
}

void Triangle::RefreshDerived()
{
	n = cross(v1 - v0, v2 - v0);
	n.Normalize();
	bounds = AABB(min(min(v0, v1), v2), max(max(v0, v1), v2));
}

bool Triangle::Hit(const ray&, float, float, HitResult&) const { return RtHostQueryUnsupported("Triangle::Hit"); }

bool Triangle::BoundingBox(float, float, AABB& outBox) const
{
	outBox = bounds;
	return true;
}

void Triangle::GetVertices(vec3& a, vec3& b, vec3& c) const { a = v0; b = v1; c = v2; }
void Triangle::SetVertices(const vec3& a, const vec3& b, const vec3& c) { v0 = a; v1 = b; v2 = c; RefreshDerived(); }
void Triangle::GetNormals(vec3& a, vec3& b, vec3& c) const { a = n0; b = n1; c = n2; }
void Triangle::SetNormals(const vec3& a, const vec3& b, const vec3& c) { n0 = a; n1 = b; n2 = c; }

// ---------------------------------------------------------------------------------------------
// StaticMesh (geom/static_mesh.cc:6-134)

StaticMesh::~StaticMesh()
{
	DestroyBvhTree(bvh);
	RtDropCompactMesh(this);
}

// A mesh imported by Raylib_LoadOBJModel keeps its faces as arrays (compact_mesh.h).  Code that wants to ADD a triangle
// needs the reference's representation back: the Triangle objects are built from the arrays and the mesh is an ordinary
// one from then on.
void RtMaterializeMesh(StaticMesh* mesh, std::vector<Triangle>& triangles)
{
	RtCompactMesh* compact = RtFindCompactMesh(mesh);
	if (!compact || compact->finalized) return;
	triangles.reserve(compact->NumTriangles());
	for (size_t t = 0; t < compact->NumTriangles(); ++t)
	{
		Triangle tri(compact->positions[3 * t], compact->positions[3 * t + 1], compact->positions[3 * t + 2],
			compact->normals[3 * t], compact->normals[3 * t + 1], compact->normals[3 * t + 2], compact->materials[t]);
		const float* st = &compact->texcoords[6 * t];
		tri.SetParameterization(st[0], st[1], st[2], st[3], st[4], st[5]);
		triangles.push_back(tri);
	}
	RtDropCompactMesh(mesh);
}

void StaticMesh::AddTriangle(const Triangle& triangle)
{
	CHECK(!bLocked);
	if (bLocked) return;
	RtMaterializeMesh(this, triangles);
	triangles.push_back(triangle);
}

void StaticMesh::SetBounds(const AABB& inBounds)
{
	CHECK(!bLocked);
	if (!bLocked) { bounds = inBounds; boundsValid = true; }
}

void StaticMesh::CalculateBounds()
{
	CHECK(!bLocked);
	if (bLocked) return;
	if (RtCompactMesh* compact = RtFindCompactMesh(this))
	{
		RtCompactCalculateBounds(*compact);
		bounds = compact->bounds; boundsValid = true;
		return;
	}
	vec3 lo(FLOAT_MAX, FLOAT_MAX, FLOAT_MAX), hi(-FLOAT_MAX, -FLOAT_MAX, -FLOAT_MAX);
	for (const Triangle& tri : triangles)
	{
		vec3 a, b, c;
		tri.GetVertices(a, b, c);
		lo = min(min(min(lo, a), b), c);
		hi = max(max(max(hi, a), b), c);
	}
	bounds = AABB(lo, hi);
	boundsValid = true;
}

void StaticMesh::ApplyTransform(const Transform& transform)
{
	CHECK(!bLocked);
	if (bLocked) return;
	if (RtCompactMesh* compact = RtFindCompactMesh(this))
	{
		RtCompactApplyTransform(*compact, transform);
		boundsValid = false;
		return;
	}
	Transform rotationOnly = transform;
	rotationOnly.SetLocation(vec3(0.0f));
	rotationOnly.SetScale(vec3(1.0f));
	std::vector<vec3> positions(3), normals(3);
	for (Triangle& tri : triangles)
	{
		tri.GetVertices(positions[0], positions[1], positions[2]);
		tri.GetNormals(normals[0], normals[1], normals[2]);
		transform.TransformVectors(positions);
		rotationOnly.TransformVectors(normals);
		tri.SetVertices(positions[0], positions[1], positions[2]);
		tri.SetNormals(normals[0], normals[1], normals[2]);
	}
	boundsValid = false;
}

void StaticMesh::Finalize()
{
	if (bLocked) return;
	if (RtCompactMesh* compact = RtFindCompactMesh(this))
	{
		// OBJ fast path: the mesh's share of the flattened scene straight from the face arrays -- no Triangle objects, no
		// BVHNode objects (bvh stays null; the flattener takes the fragment, compact_mesh.h)
		RtCompactFinalize(*compact);
		bounds = compact->bounds; boundsValid = true;
		bLocked = true;
		return;
	}
	CalculateBounds();
	std::vector<Hitable*> pointers(triangles.size());
	for (size_t i = 0; i < triangles.size(); ++i) pointers[i] = &triangles[i];
	HitableList list(pointers);
	bvh = new BVHNode(&list, 0.0f, 0.0f);
	bLocked = true;
}

bool StaticMesh::Hit(const ray&, float, float, HitResult&) const { return RtHostQueryUnsupported("StaticMesh::Hit"); }

bool StaticMesh::BoundingBox(float, float, AABB& outBox) const
{
	outBox = bounds;
	return boundsValid;
}

// ---------------------------------------------------------------------------------------------
// Rotator / Transform (geom/transform.cc:19-94)

static const float kPi = float(3.1415926535897932385);
static float ToRadians(float degrees) { return degrees * kPi / 180.0f; }
static float ToDegrees(float radians) { return radians * 180.0f / kPi; }

Rotator Rotator::directionToYawPitch(const vec3& dir)
{
	const float mag = dir.Length();
	if (mag < 0.0001f) return Rotator(0.0f, 0.0f, 0.0f);
	const float yawRad = atan2f(-dir.z, dir.x);
	const float pitchRad = asinf(dir.y / mag);
	// roll of a y-up frame seen from this direction is zero
	const float rollRad = asinf(0.0f * sinf(yawRad) + 0.0f * -cosf(yawRad));
	return Rotator(ToDegrees(yawRad), ToDegrees(pitchRad), ToDegrees(rollRad));
}

vec3 Rotator::toDirection() const
{
	const float theta = ToRadians(yaw), phi = ToRadians(pitch);
	const float cosPhi = cosf(phi);
	return vec3(sinf(theta) * cosPhi, sinf(phi), cosf(theta) * cosPhi);
}

vec3 Rotator::rotate(const vec3& position) const
{
	const float ch = cosf(ToRadians(yaw)),   sh = sinf(ToRadians(yaw));
	const float cp = cosf(ToRadians(pitch)), sp = sinf(ToRadians(pitch));
	const float cb = cosf(ToRadians(roll)),  sb = sinf(ToRadians(roll));
	const vec3 row0(ch * cb + sh * sp * sb, sb * cp, -sh * cb + ch * sp * sb);
	const vec3 row1(-ch * sb + sh * sp * cb, cb * cp, sb * sh + ch * sp * cb);
	const vec3 row2(sh * cp, -sp, ch * cp);
	return vec3(dot(row0, position), dot(row1, position), dot(row2, position));
}

void Transform::Init(const vec3& inLocation, const Rotator& inRotation, const vec3& inScale)
{
	location = inLocation;
	rotation = inRotation;
	scale = inScale;
}

void Transform::TransformVectors(std::vector<vec3>& vectors) const
{
	for (vec3& v : vectors) v = (rotation.rotate(v) * scale) + location;
}

void Transform::TransformVectors(const std::vector<vec3>& in, std::vector<vec3>& out) const
{
	out.resize(in.size());
	for (size_t i = 0; i < in.size(); ++i) out[i] = (rotation.rotate(in[i]) * scale) + location;
}

// ---------------------------------------------------------------------------------------------
// Scene (geom/scene.cc:6-31)

Scene::Scene()
{
	sunIlluminance = vec3(0.0f);
	sunDirection = normalize(vec3(0.0f, -1.0f, -0.5f));
}

Scene::~Scene()
{
	RtForgetScene(this);
	DestroyBvhTree(accelStruct);
}

void Scene::AddSceneElement(Hitable* hitable)
{
	if (!bFinalized) hitableList.hitables.push_back(hitable);
}

BVHNode* Scene::Finalize()
{
	if (!bFinalized)
	{
		bFinalized = true;
		accelStruct = new BVHNode(&hitableList, 0.0f, 0.0f);
	}
	return accelStruct;
}
