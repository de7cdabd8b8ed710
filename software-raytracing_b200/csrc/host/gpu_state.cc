// gpu_state.cc -- see gpu_state.h.  Also implements Renderer (reference: raylib/render/renderer.cc:273-356):
// where the reference builds 8x8 work cells, spins up a thread pool and polls it, this library hands the
// frame to the CUDA wavefront path tracer and copies the finished pixels back.
#include "gpu_state.h"
#include "flatten.h"
#include "host_internal.h"
#include "geom/scene.h"
#include "render/camera.h"
#include "render/image.h"
#include "render/renderer.h"
#include "core/logger.h"
#include "core/assertion.h"
#include "rt_rng.h"

#include <chrono>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>

namespace
{
	struct SceneEntry
	{
		RtDeviceScene* device = nullptr;
		uint64_t counts[8] = { 0 };
	};

	std::recursive_mutex g_mutex;
	int g_device = -1;
	uint64_t g_frameSeed = RT_RNG_DEFAULT_FRAME_SEED;
	bool g_collectStats = false;
	bool g_timeStages = false;
	uint32_t g_samplesPerPass = 0;
	uint32_t g_pipes = 0;
	std::map<int, RtRenderContext*> g_contexts;                      // per device
	std::map<std::pair<const Scene*, int>, SceneEntry> g_scenes;     // (scene, device)
	std::map<const Scene*, std::shared_ptr<RtFlatScene>> g_prebuilt; // scenes that came from RaylibB200_LoadFlattenedScene
	struct Scratch { void* ptr = nullptr; uint64_t bytes = 0; };
	std::map<std::pair<int, int>, Scratch> g_scratch;                // (device, slot)

	thread_local std::string t_lastError;
	thread_local RaylibB200Stats t_lastStats;
	thread_local bool t_haveStats = false;

	int DefaultDevice()
	{
		const char* names[] = { "RAYLIB_B200_DEVICE", "LOCAL_RANK" };
		for (const char* name : names)
		{
			const char* v = getenv(name);
			if (v && *v) return atoi(v);
		}
		return 0;
	}

	void* ScratchBuffer(int device, int slot, uint64_t bytes)
	{
		Scratch& s = g_scratch[{device, slot}];
		if (s.bytes >= bytes && s.ptr) return s.ptr;
		if (s.ptr) rt_device_free(device, s.ptr);
		s.ptr = nullptr; s.bytes = 0;
		if (rt_device_alloc(device, bytes, &s.ptr) != 0) return nullptr;
		s.bytes = bytes;
		return s.ptr;
	}
}

namespace RtGpu
{
	int DeviceCount() { return rt_device_count(); }

	int CurrentDevice()
	{
		std::lock_guard<std::recursive_mutex> lock(g_mutex);
		if (g_device < 0)
		{
			const int n = DeviceCount();
			g_device = n > 0 ? DefaultDevice() % n : 0;
		}
		return g_device;
	}

	bool SetDevice(int device)
	{
		std::lock_guard<std::recursive_mutex> lock(g_mutex);
		if (device < 0 || device >= DeviceCount()) { SetLastError("RaylibB200_SetDevice: no such CUDA device"); return false; }
		g_device = device;
		return true;
	}

	void SetFrameSeed(uint64_t seed) { g_frameSeed = seed; }
	uint64_t FrameSeed() { return g_frameSeed; }
	void SetCollectStats(bool enable) { g_collectStats = enable; }
	void SetTimeStages(bool enable) { g_timeStages = enable; }
	void SetSamplesPerPass(uint32_t samples) { g_samplesPerPass = samples; }
	void SetPipes(uint32_t pipes) { g_pipes = pipes; }

	void SetLastError(const std::string& message)
	{
		t_lastError = message;
		if (!message.empty()) fprintf(stderr, "raylib-b200: %s\n", message.c_str());
	}
	const char* LastError() { return t_lastError.c_str(); }
	void SetLastStats(const RaylibB200Stats& stats) { t_lastStats = stats; t_haveStats = true; }
	bool GetLastStats(RaylibB200Stats* out)
	{
		if (!t_haveStats || !out) return false;
		*out = t_lastStats;
		return true;
	}

	RtRenderContext* AcquireContext()
	{
		std::lock_guard<std::recursive_mutex> lock(g_mutex);
		if (DeviceCount() <= 0)
		{
			SetLastError("no CUDA device is available; this library renders on the GPU only (no CPU path)");
			return nullptr;
		}
		const int device = CurrentDevice();
		auto it = g_contexts.find(device);
		if (it != g_contexts.end()) return it->second;
		RtRenderContext* ctx = nullptr;
		if (rt_context_create(device, &ctx) != 0) { SetLastError(std::string("rt_context_create: ") + rt_last_error()); return nullptr; }
		g_contexts[device] = ctx;
		return ctx;
	}

	const RtDeviceScene* AcquireScene(const Scene* scene, uint64_t* outCounts8)
	{
		std::lock_guard<std::recursive_mutex> lock(g_mutex);
		if (!scene) { SetLastError("null scene handle"); return nullptr; }
		if (DeviceCount() <= 0)
		{
			SetLastError("no CUDA device is available; this library renders on the GPU only (no CPU path)");
			return nullptr;
		}
		const int device = CurrentDevice();
		auto key = std::make_pair(scene, device);
		auto it = g_scenes.find(key);
		if (it == g_scenes.end())
		{
			RtFlatScene fresh;
			std::string why;
			const auto t0 = std::chrono::steady_clock::now();
			const auto pre = g_prebuilt.find(scene);
			if (pre == g_prebuilt.end() && !RtFlattenScene(scene, fresh, why)) { SetLastError("cannot flatten scene: " + why); return nullptr; }
			const RtFlatScene& flat = pre != g_prebuilt.end() ? *pre->second : fresh;
			const auto t1 = std::chrono::steady_clock::now();
			SceneEntry entry;
			if (rt_scene_upload(device, &flat.desc, &entry.device) != 0)
			{
				SetLastError(std::string("rt_scene_upload: ") + rt_last_error());
				return nullptr;
			}
			const auto t2 = std::chrono::steady_clock::now();
			entry.counts[0] = flat.nodes.empty() ? flat.quantNodes.size() : flat.nodes.size(); entry.counts[1] = flat.triHot.size(); entry.counts[2] = flat.spheres.size();
			entry.counts[3] = flat.cubes.size(); entry.counts[4] = flat.materials.size(); entry.counts[5] = flat.textures.size();
			entry.counts[6] = flat.desc.maxStackDepth; entry.counts[7] = flat.desc.numLeaves;
			LOG("[STAT] scene -> GPU %d: %llu nodes, %llu triangles, %llu spheres, %.1f MB, flatten %.1f ms, upload %.1f ms",
				device, (unsigned long long)entry.counts[0], (unsigned long long)entry.counts[1], (unsigned long long)entry.counts[2],
				(double)flat.HostBytes() / 1.0e6,
				std::chrono::duration<double, std::milli>(t1 - t0).count(), std::chrono::duration<double, std::milli>(t2 - t1).count());
			it = g_scenes.emplace(key, entry).first;
		}
		if (outCounts8) memcpy(outCounts8, it->second.counts, sizeof(it->second.counts));
		return it->second.device;
	}

	void ReleaseAll()
	{
		std::lock_guard<std::recursive_mutex> lock(g_mutex);
		for (auto& kv : g_scenes) rt_scene_free(kv.second.device);
		g_scenes.clear();
		g_prebuilt.clear();
		for (auto& kv : g_scratch) rt_device_free(kv.first.first, kv.second.ptr);
		g_scratch.clear();
		for (auto& kv : g_contexts) rt_context_destroy(kv.second);
		g_contexts.clear();
	}

	bool Render(const RendererSettings* settings, const Scene* scene, const Camera* camera,
	            Image2D* hostImage, void* deviceImage, void* deviceShard,
	            uint32_t shardRank, uint32_t shardCount, uint32_t renderModeOverride, void* stream)
	{
		std::lock_guard<std::recursive_mutex> lock(g_mutex);
		const auto wall0 = std::chrono::steady_clock::now();
		t_lastError.clear();
		if (!settings || !scene || !camera) { SetLastError("Raylib_Render: null settings/scene/camera"); return false; }
		if (settings->viewportWidth == 0 || settings->viewportHeight == 0) { SetLastError("Raylib_Render: empty viewport"); return false; }
		if (settings->renderMode >= RAYLIB_RENDERMODE_MAX && renderModeOverride == 0) { SetLastError("Raylib_Render: invalid renderMode"); return false; }

		const RtDeviceScene* deviceScene = AcquireScene(scene);
		if (!deviceScene) return false;
		RtRenderContext* ctx = AcquireContext();
		if (!ctx) return false;
		const int device = CurrentDevice();

		RtCamera cam;
		RtFlattenCamera(camera, cam);

		RtRenderParams params;
		memset(&params, 0, sizeof(params));
		params.width = settings->viewportWidth;
		params.height = settings->viewportHeight;
		params.samplesPerPixel = settings->samplesPerPixel;
		params.maxPathLength = settings->maxPathLength;
		params.rayTMin = settings->rayTMin;
		params.renderMode = renderModeOverride ? renderModeOverride : settings->renderMode;
		params.frameSeed = g_frameSeed;
		// hostImage: always the whole frame; deviceShard / deviceImage: the tiles of (shardRank, shardCount)
		params.shardRank = hostImage ? 0 : shardRank;
		params.shardCount = hostImage ? 1 : (shardCount ? shardCount : 1);
		if (params.shardRank >= params.shardCount) { SetLastError("Raylib_Render: shardRank >= shardCount"); return false; }
		params.samplesPerPass = g_samplesPerPass;
		params.pipes = g_pipes;
		params.collectStats = g_collectStats ? 1u : 0u;
		params.timeStages = g_timeStages ? 1u : 0u;

		// Final pixels go straight into the row-major frame (this GPU's memory, or another rank's frame mapped over NVLink:
		// the last accumulate IS the gather) unless the caller asked for a tile-major shard buffer.
		const uint64_t imageBytes = (uint64_t)params.width * params.height * 16ull;
		void* imageBuffer = nullptr;
		if (!deviceShard)
		{
			imageBuffer = deviceImage ? deviceImage : ScratchBuffer(device, 1, imageBytes);
			if (!imageBuffer) { SetLastError(std::string("device allocation failed: ") + rt_last_error()); return false; }
			params.imageOut = imageBuffer;
		}

		RtRenderStats rs;
		if (rt_render_shard(ctx, deviceScene, &cam, &params, deviceShard, stream, &rs) != 0)
		{
			SetLastError(std::string("rt_render_shard: ") + rt_last_error());
			return false;
		}

		uint64_t d2h = 0;
		if (hostImage)
		{
			// Pixel is four packed floats, same as the device float4 image
			if (rt_copy_to_host(device, hostImage->MutablePixels(), imageBuffer, imageBytes, stream) != 0)
			{
				SetLastError(std::string("device -> host copy failed: ") + rt_last_error());
				return false;
			}
			d2h = imageBytes;
		}

		RaylibB200Stats st;
		memset(&st, 0, sizeof(st));
		st.rayQueries = rs.rayQueries; st.pixelSamples = rs.pixelSamples;
		st.boxTests = rs.boxTests; st.triTests = rs.triTests; st.sphereTests = rs.sphereTests; st.nodeVisits = rs.nodeVisits;
		st.refBoxTests = rs.refBoxTests; st.refTriTests = rs.refTriTests; st.refSphereTests = rs.refSphereTests; st.statRays = rs.statRays;
		st.deviceMs = rs.deviceMs;
		st.extendMs = rs.extendMs; st.extendLaunches = rs.extendLaunches;
		st.nodeIters = rs.nodeIters; st.nodeStep = rs.nodeStep; st.nodeAlive = rs.nodeAlive; st.leafIters = rs.leafIters; st.leafBusy = rs.leafBusy;
		st.gateTests = rs.gateTests; st.cubeTests = rs.cubeTests;
		st.totalMs = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - wall0).count();
		st.h2dBytes = sizeof(RtCamera) + sizeof(RtRenderParams);
		st.d2hBytes = d2h;
		st.kernelLaunches = rs.kernelLaunches;
		st.passes = rs.passes;
		st.device = (uint32_t)device;
		SetLastStats(st);
		return true;
	}
}

namespace RtGpu
{
	bool RenderAux(const RendererSettings* settings, const Scene* scene, const Camera* camera, Image2D* albedoImage, Image2D* normalImage)
	{
		std::lock_guard<std::recursive_mutex> lock(g_mutex);
		const auto wall0 = std::chrono::steady_clock::now();
		t_lastError.clear();
		if (!settings || !scene || !camera || !albedoImage || !normalImage) { SetLastError("RaylibB200_RenderAux: null argument"); return false; }
		if (settings->viewportWidth == 0 || settings->viewportHeight == 0) { SetLastError("RaylibB200_RenderAux: empty viewport"); return false; }
		const RtDeviceScene* deviceScene = AcquireScene(scene);
		if (!deviceScene) return false;
		RtRenderContext* ctx = AcquireContext();
		if (!ctx) return false;
		const int device = CurrentDevice();
		const uint32_t W = settings->viewportWidth, H = settings->viewportHeight;
		for (Image2D* img : { albedoImage, normalImage })
			if (img->GetWidth() != W || img->GetHeight() != H) img->Reallocate(W, H);

		RtCamera cam;
		RtFlattenCamera(camera, cam);
		const uint64_t shardBytes = (uint64_t)rt_shard_tile_capacity(W, H, 1) * RT_TILE_PIXELS * 16ull;
		const uint64_t imageBytes = (uint64_t)W * H * 16ull;
		void* shardA = ScratchBuffer(device, 0, shardBytes);
		void* shardB = ScratchBuffer(device, 2, shardBytes);
		void* image = ScratchBuffer(device, 1, imageBytes);
		if (!shardA || !shardB || !image) { SetLastError(std::string("device allocation failed: ") + rt_last_error()); return false; }

		RtRenderParams params;
		memset(&params, 0, sizeof(params));
		params.width = W; params.height = H;
		params.samplesPerPixel = 1; params.maxPathLength = settings->maxPathLength;
		params.rayTMin = settings->rayTMin;
		params.renderMode = RT_RENDERMODE_AUX;
		params.frameSeed = g_frameSeed;
		params.shardRank = 0; params.shardCount = 1;
		params.auxShardOut = shardB;
		RtRenderStats rs;
		if (rt_render_shard(ctx, deviceScene, &cam, &params, shardA, nullptr, &rs) != 0)
		{
			SetLastError(std::string("rt_render_shard: ") + rt_last_error());
			return false;
		}
		Image2D* targets[2] = { albedoImage, normalImage };
		void* shards[2] = { shardA, shardB };
		for (int i = 0; i < 2; ++i)
		{
			if (rt_assemble(device, shards[i], 1, W, H, image, nullptr) != 0 ||
			    rt_copy_to_host(device, targets[i]->MutablePixels(), image, imageBytes, nullptr) != 0)
			{
				SetLastError(std::string("aux readback failed: ") + rt_last_error());
				return false;
			}
		}
		RaylibB200Stats st;
		memset(&st, 0, sizeof(st));
		st.rayQueries = rs.rayQueries; st.pixelSamples = rs.pixelSamples; st.deviceMs = rs.deviceMs;
		st.totalMs = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - wall0).count();
		st.h2dBytes = sizeof(RtCamera) + sizeof(RtRenderParams); st.d2hBytes = 2 * imageBytes;
		st.kernelLaunches = rs.kernelLaunches + 2; st.device = (uint32_t)device;
		SetLastStats(st);
		return true;
	}

	bool PostProcessDevice(void* deviceImage, uint32_t width, uint32_t height, void* deviceOutArgb8, float* outMaxWhite, void* stream)
	{
		std::lock_guard<std::recursive_mutex> lock(g_mutex);
		t_lastError.clear();
		if (DeviceCount() <= 0) { SetLastError("no CUDA device is available; this library renders on the GPU only (no CPU path)"); return false; }
		if (rt_postprocess(CurrentDevice(), deviceImage, width, height, (uint32_t*)deviceOutArgb8, outMaxWhite, stream) != 0)
		{
			SetLastError(std::string("rt_postprocess: ") + rt_last_error());
			return false;
		}
		return rt_stream_sync(CurrentDevice(), stream) == 0;
	}

	bool PostProcessHostImage(Image2D* hostImage)
	{
		std::lock_guard<std::recursive_mutex> lock(g_mutex);
		t_lastError.clear();
		if (!hostImage || hostImage->GetWidth() == 0 || hostImage->GetHeight() == 0) { SetLastError("RaylibB200_PostProcessGPU: empty image"); return false; }
		if (DeviceCount() <= 0) { SetLastError("no CUDA device is available; this library renders on the GPU only (no CPU path)"); return false; }
		const int device = CurrentDevice();
		const uint64_t bytes = (uint64_t)hostImage->GetWidth() * hostImage->GetHeight() * 16ull;
		void* image = ScratchBuffer(device, 1, bytes);
		if (!image) { SetLastError(std::string("device allocation failed: ") + rt_last_error()); return false; }
		float maxWhite = 1.0f;
		if (rt_copy_to_device(device, image, hostImage->MutablePixels(), bytes, nullptr) != 0 ||
		    rt_postprocess(device, image, hostImage->GetWidth(), hostImage->GetHeight(), nullptr, &maxWhite, nullptr) != 0 ||
		    rt_copy_to_host(device, hostImage->MutablePixels(), image, bytes, nullptr) != 0)
		{
			SetLastError(std::string("GPU post-process failed: ") + rt_last_error());
			return false;
		}
		LOG("Max white luminance: %f", maxWhite);
		return true;
	}
}

namespace RtGpu
{
	void AdoptPrebuilt(const Scene* scene, std::shared_ptr<RtFlatScene> flat)
	{
		std::lock_guard<std::recursive_mutex> lock(g_mutex);
		g_prebuilt[scene] = std::move(flat);
	}
	std::shared_ptr<RtFlatScene> Prebuilt(const Scene* scene)
	{
		std::lock_guard<std::recursive_mutex> lock(g_mutex);
		auto it = g_prebuilt.find(scene);
		return it == g_prebuilt.end() ? nullptr : it->second;
	}
}

void RtForgetScene(const Scene* scene)
{
	std::lock_guard<std::recursive_mutex> lock(g_mutex);
	g_prebuilt.erase(scene);
	for (auto it = g_scenes.begin(); it != g_scenes.end();)
	{
		if (it->first.first == scene) { rt_scene_free(it->second.device); it = g_scenes.erase(it); }
		else ++it;
	}
}

// ---------------------------------------------------------------------------------------------
// Renderer

bool Renderer::IsDenoiserSupported() { return false; }    // OIDN is a Windows-only prebuilt in the reference (renderer.cc:28-33)

void Renderer::RenderScene(const RendererSettings* settings, const Scene* world, const Camera* camera, Image2D* outImage)
{
	CHECK(settings != nullptr && world != nullptr && camera != nullptr && outImage != nullptr);
	if (!settings || !world || !camera || !outImage) return;
	CHECK(world->GetAccelStruct() != nullptr);

	if (settings->viewportWidth != outImage->GetWidth() || settings->viewportHeight != outImage->GetHeight())
		outImage->Reallocate(settings->viewportWidth, settings->viewportHeight);

	if (!RtGpu::Render(settings, world, camera, outImage, nullptr, nullptr, 0, 1, 0, nullptr))
	{
		LOG("Raylib_Render FAILED: %s", RtGpu::LastError());
		return;
	}
	RaylibB200Stats st;
	if (RtGpu::GetLastStats(&st))
	{
		LOG("[STAT] GPU render: %.3f ms device, %.3f ms total, %.2f Mrays/s, %.2f M pixel-samples/s",
			st.deviceMs, st.totalMs,
			st.deviceMs > 0.0 ? (double)st.rayQueries / st.deviceMs * 1.0e-3 : 0.0,
			st.deviceMs > 0.0 ? (double)st.pixelSamples / st.deviceMs * 1.0e-3 : 0.0);
	}
}

bool Renderer::DenoiseScene(Image2D* mainImage, bool, Image2D*, Image2D*, Image2D* outDenoisedImage)
{
	CHECKF(mainImage != nullptr, "mainImage should not be null");
	CHECKF(outDenoisedImage != nullptr, "outDenoisedImage should not be null");
	return false;
}
