// host_internal.h -- declarations shared by the host translation units of libraylib (not installed).
#pragma once
#include <stdint.h>
#include <stddef.h>

class Scene;
class Camera;
class Image2D;
struct RendererSettings;

// Reports (once per call site name) that a CPU-side query was attempted.  Ray queries and
// scattering only exist on the device in this library; there is deliberately no CPU fallback.
bool RtHostQueryUnsupported(const char* what);

// Stream key for BVH split axes (reference draws them from Random(), geom/bvh.cc:43).
uint64_t RtGetBvhBuildKey();
void RtSetBvhBuildKey(uint64_t key);

// Scene -> device bookkeeping (gpu_state.cc)
void RtForgetScene(const Scene* scene);

// true for Image2D objects created (and destroyed) by this library: Raylib_CreateImage / Raylib_LoadImage (raylib_api.cc).
// Only their storage is page-locked for read-backs -- the library sees it go away.
bool RtIsLibraryImage(const Image2D* image);
