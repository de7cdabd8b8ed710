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
