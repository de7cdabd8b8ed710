// compact_mesh.h -- OBJ fast path (reference: loader/obj_loader.cc:128-245, geom/static_mesh.cc:80-95).
//
// The reference's importer turns every face into a 152-byte Triangle object, and StaticMesh::Finalize then builds a
// pointer tree of 48-byte BVHNode objects over them -- all of which this library would only walk once to flatten.  A mesh
// that comes from Raylib_LoadOBJModel keeps its faces as plain arrays instead (a "compact" mesh: a side table keyed by the
// StaticMesh, whose triangle list stays empty), and Finalize() produces the mesh's share of the flattened scene directly:
// the reference topology is built over indices (same split-axis stream, same std::sort comparisons, hence the same
// in-order leaf ranks, gates and node records as the object path -- tests compare the two bit for bit), no Triangle and
// no BVHNode is ever allocated.  Anything that needs the objects -- StaticMesh::AddTriangle from client code -- turns
// the mesh back into an ordinary one first (RtMaterializeMesh).
#pragma once
#include "bvh_sah.h"
#include "geom/primitives.h"
#include "geom/transform.h"
#include <memory>
#include <vector>

struct RtCompactMesh
{
	// per triangle, as the importer produced them
	std::vector<vec3> positions;      // 3 per triangle
	std::vector<vec3> normals;        // 3 per triangle (face normal already substituted where a corner had none)
	std::vector<float> texcoords;     // 6 per triangle: s0 t0 s1 t1 s2 t2
	std::vector<Material*> materials; // 1 per triangle
	AABB bounds;
	bool boundsValid = false;

	// after Finalize(): the mesh's fragment of the flattened scene, indices local to the mesh
	bool finalized = false;
	RtArray<RtTriHot>  triHot;        // gate / rank words local, material word = index into distinctMaterials
	RtArray<RtTriCold> triCold;
	RtArray<float>     gateBoxes;     // 8 per gate
	RtArray<uint32_t>  triGate;
	RtArray<RtNode>    refNodes;      // local references
	RtLeafGroups       groups;        // one per triangle, tight box
	std::vector<const Material*> distinctMaterials;     // in the order of the triangle list
	uint32_t topRef = 0;
	float topLo[3], topHi[3];
	uint32_t maxNodeDepth = 0;

	size_t NumTriangles() const { return materials.size(); }
};

// Side table.  All functions are thread-safe with respect to each other for DIFFERENT meshes.
RtCompactMesh* RtFindCompactMesh(const StaticMesh* mesh);
RtCompactMesh* RtAttachCompactMesh(const StaticMesh* mesh);      // creates the entry
void RtDropCompactMesh(const StaticMesh* mesh);

// The operations StaticMesh forwards to while the mesh is compact.
void RtCompactCalculateBounds(RtCompactMesh& m);
void RtCompactApplyTransform(RtCompactMesh& m, const Transform& transform);
void RtCompactFinalize(RtCompactMesh& m);
