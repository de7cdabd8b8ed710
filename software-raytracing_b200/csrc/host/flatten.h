// flatten.h -- host flattener: client object graph -> arrays of include/rt_scene_format.h.
#pragma once
#include "rt_scene_format.h"
#include <string>
#include <vector>

class Scene;
class Camera;

struct RtFlatScene
{
	std::vector<RtNode>     nodes;       // binary SAH tree over the leaf groups (host only: equivalence tests)
	std::vector<RtNode4>    wideNodes;   // nodes collapsed to 4-wide records, exact boxes (host only)
	std::vector<RtNodeQ4>   quantNodes;  // wideNodes quantized to 64 bytes: what the device walks
	std::vector<RtNode>     refNodes;    // reference topology
	std::vector<RtTriHot>   triHot;
	std::vector<RtTriCold>  triCold;
	std::vector<uint32_t>   triRank;
	std::vector<uint32_t>   triGate;
	std::vector<float>      gateBoxes;   // 8 floats per gate
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
