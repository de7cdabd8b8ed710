// gpu_state.h -- per-process GPU bookkeeping of libraylib: the device in use, its render context
// (path-state arenas) and the uploaded copy of every finalized scene.
#pragma once
#include "rt_device_abi.h"
#include "raylib_b200.h"
#include <memory>
#include <string>

struct RtFlatScene;
class Scene;
class Camera;
class Image2D;
struct RendererSettings;

namespace RtGpu
{
	int  DeviceCount();
	int  CurrentDevice();
	bool SetDevice(int device);
	// Devices a Raylib_Render frame is spread over: the first `count` visible devices (0 = all).  Returns how many.
	int  SetDevices(int count);
	int  ActiveDeviceCount();
	// Tuning knobs (RtTuning): read from the environment once per process, again on ReloadTuning().
	void ReloadTuning();
	const RtTuning& Tuning();
	// Drops the page-lock of an Image2D's storage before the storage goes away (Raylib_DestroyImage, Image2D::Reallocate).
	void ForgetHostImage(const Image2D* image);

	void SetFrameSeed(uint64_t seed);
	uint64_t FrameSeed();
	void SetCollectStats(bool enable);
	void SetTimeStages(bool enable);
	void SetSamplesPerPass(uint32_t samples);
	void SetPipes(uint32_t pipes);
	void SetFusedPass(uint32_t mode);

	void SetLastError(const std::string& message);
	const char* LastError();
	void SetLastStats(const RaylibB200Stats& stats);
	bool GetLastStats(RaylibB200Stats* out);

	// Flatten + upload (once per scene and device). nullptr on failure (LastError says why).
	const RtDeviceScene* AcquireScene(const Scene* scene, uint64_t* outCounts8 = nullptr);
	RtRenderContext* AcquireContext();
	// A scene whose flattened form was read from disk (RaylibB200_LoadFlattenedScene): AcquireScene uploads it as is.
	void AdoptPrebuilt(const Scene* scene, std::shared_ptr<RtFlatScene> flat);
	std::shared_ptr<RtFlatScene> Prebuilt(const Scene* scene);
	void ReleaseAll();

	// Shared implementation of every render entry point.  Exactly one of hostImage / deviceImage /
	// deviceShard is non-null.
	bool Render(const RendererSettings* settings, const Scene* scene, const Camera* camera,
	            Image2D* hostImage, void* deviceImage, void* deviceShard,
	            uint32_t shardRank, uint32_t shardCount, uint32_t renderModeOverride, void* stream, bool pinHostImage = false);

	// Denoiser inputs in one primary-hit pass (reference: the Albedo and MicrosurfaceNormal debug renders of
	// src/main.cc:464-476, two extra Raylib_Render calls).  Host images are resized to the viewport.
	bool RenderAux(const RendererSettings* settings, const Scene* scene, const Camera* camera, Image2D* albedoImage, Image2D* normalImage);

	// Image2D::PostProcess on the GPU: device-resident RGBA float4 image in place, optional packed ARGB8 copy.
	bool PostProcessDevice(void* deviceImage, uint32_t width, uint32_t height, void* deviceOutArgb8, float* outMaxWhite, void* stream);
	// Same for a host Image2D (H2D, kernels, D2H); used for large frames.
	bool PostProcessHostImage(Image2D* image);
}
