// OBJModel / OBJLoader -- Wavefront OBJ + MTL import into StaticMeshes and raylib materials
// (reference: raylib/loader/obj_loader.h:16-75, raylib/loader/obj_loader.cc:45-400).
//
// The reference delegates parsing to tinyobjloader v2.0.0rc10, which is neither vendored in the
// reference repository nor available here; csrc/host/obj_loader.cc carries its own parser for the
// statements the reference consumes (v / vt / vn / f / o / g / usemtl / mtllib; newmtl, Kd, Ks, Ke, Tf,
// Ns, Ni, illum, Pr, Pm, map_Kd, map_Pr, map_Pm, map_Ke, norm, map_bump / bump) and then follows the
// reference's conversion rules literally.  Host-side only; not part of the GPU hot path.
#pragma once

#include "raylib_types.h"
#include "core/noncopyable.h"
#include "core/vec3.h"

#include <map>
#include <memory>
#include <string>
#include <vector>

class Material;
class StaticMesh;
class Hitable;
class Image2D;

// CAUTION (as in the reference): finalize the meshes with StaticMesh::Finalize() or
// OBJModel::FinalizeAllMeshes() before adding the model to a scene.
struct OBJModel
{
	OBJModel()
		: rootObject(nullptr)
		, localMinBound(vec3(0.0f, 0.0f, 0.0f))
		, localMaxBound(vec3(0.0f, 0.0f, 0.0f))
	{
	}

	RAYLIB_API void FinalizeAllMeshes();

	Hitable* rootObject;
	std::vector<StaticMesh*> staticMeshes;

	// Invalid after transforms have been applied to the meshes.
	vec3 localMinBound;
	vec3 localMaxBound;
};

class OBJLoader : public Noncopyable
{
public:
	static void Initialize();
	static void Destroy();

	RAYLIB_API static bool LoadModelFromFile(const char* filepath, OBJModel* outModel);

public:
	explicit OBJLoader();
	~OBJLoader();

	bool LoadFromFile(const char* filepath, OBJModel& outModel);

private:
	std::map<std::string, std::shared_ptr<Image2D>> imageDB;
	std::vector<Material*> materials;
};
