// flatten.h -- host flattener: client object graph -> arrays of include/rt_scene_format.h.
#pragma once
#include "rt_scene_format.h"
#include "rt_array.h"
#include <memory>
#include <string>
#include <utility>
#include <vector>

class Scene;
class Camera;

struct RtFlatScene
{
	RtArray<RtNode>     nodes;       // binary SAH tree over the leaf groups (host only: equivalence tests)
	RtArray<RtNode4>    wideNodes;   // nodes collapsed to 4-wide records, exact boxes (host only)
	RtArray<RtNodeQ4>   quantNodes;  // wideNodes quantized to 64 bytes: what the device walks
	RtArray<RtNode>         refNodes;    // reference topology
	RtArray<RtTriHot>       triHot;
	RtArray<RtTriCold>      triCold;
	RtArray<uint32_t>       triRank;
	RtArray<uint32_t>       triGate;
	RtArray<float>          gateBoxes;   // 8 floats per gate
	std::vector<RtSphere>   spheres;
	std::vector<uint32_t>   sphereMaterial;
	std::vector<uint32_t>   sphereRank;
	std::vector<uint32_t>   sphereGate;
	std::vector<RtCube>     cubes;
	std::vector<uint32_t>   cubeRank;
	std::vector<uint32_t>   cubeGate;
	std::vector<RtMaterial> materials;
	std::vector<RtTexture>  textures;
	std::vector<float>      texels;      // RGBA float4 per texel
	RtSceneDesc desc;                    // pointers into the vectors above

	uint64_t HostBytes() const;
};

// Walks scene->GetAccelStruct() (reference object model: geom/bvh.h, geom/static_mesh.h, ...) and
// fills `out`.  Returns false and sets `error` when the graph holds something the device path
// cannot express (an unfinalized mesh, a raw HitableList element, a user-defined Hitable/Material).
bool RtFlattenScene(const Scene* scene, RtFlatScene& out, std::string& error);

void RtFlattenCamera(const Camera* camera, RtCamera& out);

// Points flat.desc at flat's own vectors (after the vectors were filled, moved or read from disk).
void RtBindFlatScene(RtFlatScene& flat);

// On-disk form of a flattened scene (skip the object graph, the reference BVH build and the SAH /
// collapse / quantize pipeline on re-loads of a large scene; reference path: loader/obj_loader.cc:128-245 +
// geom/static_mesh.cc:80-95 every time).  Layout: 64-byte header {magic "RTFLAT01", record sizes, payload checksum},
// the scalar part of RtSceneDesc, then every array as {uint64 count, raw records}.  The host-only arrays (binary SAH
// tree, exact 4-wide nodes) are not stored: a scene loaded from disk renders and answers ray queries, the CPU
// equivalence tests run on freshly flattened scenes.
bool RtSaveFlatScene(const RtFlatScene& flat, const char* path, std::string& error);
bool RtLoadFlatScene(const char* path, RtFlatScene& out, std::string& error);
