// OBJModel / OBJLoader -- Wavefront OBJ + MTL import into StaticMeshes and raylib materials
// (reference: raylib/loader/obj_loader.h:16-75, raylib/loader/obj_loader.cc:45-400).
//
// The reference delegates parsing to tinyobjloader v2.0.0rc10, which is neither vendored in the
// reference repository nor available here; csrc/host/obj_loader.cc carries its own parser for the
// statements the reference consumes (v / vt / vn / f / o / g / usemtl / mtllib; newmtl, Kd, Ks, Ke, Tf,
// Ns, Ni, illum, Pr, Pm, map_Kd, map_Pr, map_Pm, map_Ke, norm, map_bump / bump) and then follows the
// reference's conversion rules literally.  Host-side only; not part of the GPU hot path.
#pragma once
#include <map>
#include <memory>
#include <string>
#include <vector>
#include "raylib_types.h"
#include "core/noncopyable.h"
#include "core/vec3.h"

class Material; class StaticMesh; class Hitable; class Image2D;

// One loaded model: a root element for Raylib_AddOBJModelToScene plus the meshes behind it.  The meshes have to be
// finalized (StaticMesh::Finalize on each, or FinalizeAllMeshes) before the model goes into a scene; the local bounds
// describe the file as loaded and go stale once a transform has been applied.
struct OBJModel
{
	Hitable* rootObject = nullptr;
	std::vector<StaticMesh*> staticMeshes;
	vec3 localMinBound = vec3(0.0f, 0.0f, 0.0f);
	vec3 localMaxBound = vec3(0.0f, 0.0f, 0.0f);

	RAYLIB_API void FinalizeAllMeshes();
};

class OBJLoader : public Noncopyable
{
	std::map<std::string, std::shared_ptr<Image2D>> imageDB;      // textures by path, shared between materials
	std::vector<Material*> materials;                              // owned: one per .mtl entry

public:
	explicit OBJLoader(); ~OBJLoader();
	bool LoadFromFile(const char* filepath, OBJModel& outModel);

	// process-wide set-up / tear-down (Raylib_Initialize / Raylib_Terminate) and the one-call form used by the C API
	static void Initialize(); static void Destroy();
	RAYLIB_API static bool LoadModelFromFile(const char* filepath, OBJModel* outModel);
};
