// oracle/ref_shims.cc -- TEST INFRASTRUCTURE (not product code).
// Link-time stand-ins for the three pieces of the reference that cannot be
// built in this container:
//   * raylib/loader/dll_loader.cc  (#error off Windows, :49-51)  -> null FreeImage pointers
//   * raylib/loader/obj_loader.cc  (needs tinyobjloader)         -> OBJ loading reports failure
//   * the block-scope declaration `void LogMain();` inside namespace Logger
//     (raylib/core/logger.cc:30) names Logger::LogMain under g++, while the
//     definition (:76) is global.
#include "loader/dll_loader.h"
#include "loader/obj_loader.h"
#include "geom/static_mesh.h"
#include <atomic>

namespace FreeImage
{
	PFN_Initialise         Initialise         = nullptr;
	PFN_DeInitialise       DeInitialise       = nullptr;
	PFN_GetFIFFromFilename GetFIFFromFilename = nullptr;
	PFN_Load               Load               = nullptr;
	PFN_ConvertToRGBAF     ConvertToRGBAF     = nullptr;
	PFN_Unload             Unload             = nullptr;
	PFN_GetBits            GetBits            = nullptr;
	PFN_GetWidth           GetWidth           = nullptr;
	PFN_GetHeight          GetHeight          = nullptr;
	PFN_GetPitch           GetPitch           = nullptr;
	PFN_ConvertTo32Bits    ConvertTo32Bits    = nullptr;
	PFN_ConvertFromRawBits ConvertFromRawBits = nullptr;
	PFN_Save               Save               = nullptr;
	bool LoadDLL() { return true; }
}

void LogMain();
namespace Logger { void LogMain() { ::LogMain(); } }

void OBJLoader::Initialize() {}
void OBJLoader::Destroy() {}
bool OBJLoader::LoadModelFromFile(const char*, OBJModel*) { return false; }
void OBJModel::FinalizeAllMeshes()
{
	for (StaticMesh* mesh : staticMeshes) mesh->Finalize();
}

static std::atomic<long> g_debugbreaks{0};
extern "C" void oracle_on_debugbreak(void) { g_debugbreaks++; }
extern "C" long oracle_debugbreak_count(void) { return g_debugbreaks.load(); }
