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

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <condition_variable>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

namespace
{
	struct SceneEntry
	{
		RtDeviceScene* device = nullptr;
		bool referenceTree = false;      // uploaded with the reference topology (statistics frames only)
		uint64_t counts[8] = { 0 };
		// distant lighting the upload was made with: the sky panorama lives in the uploaded texture arrays, so a scene whose
		// sky changed is flattened again; the sun is passed per frame (RtRenderParams.lightingOverride)
		ImageHandle skyHandle = 0;
		uint64_t skySignature = 0;
	};

	std::recursive_mutex g_mutex;
	std::vector<int> g_devices;      // devices a Raylib_Render frame is spread over; [0] is the primary (frame + device-only calls)
	bool g_devicesResolved = false;
	uint64_t g_frameSeed = RT_RNG_DEFAULT_FRAME_SEED;
	bool g_collectStats = false;
	bool g_timeStages = false;
	uint32_t g_samplesPerPass = 0;
	uint32_t g_pipes = 0;
	uint32_t g_fusedPass = 0;       // RaylibB200_SetFusedPass: 0 = the environment's / automatic
	RtTuning g_tuning;
	bool g_tuningLoaded = false;
	uint64_t g_multiDeviceMinSamples = 2ull << 20;   // frames with fewer pixel-samples stay on the primary device
	std::map<int, RtRenderContext*> g_contexts;                      // per device
	std::map<std::pair<const Scene*, int>, SceneEntry> g_scenes;     // (scene, device)
	std::map<const Scene*, std::shared_ptr<RtFlatScene>> g_prebuilt; // scenes that came from RaylibB200_LoadFlattenedScene
	struct Scratch { void* ptr = nullptr; uint64_t bytes = 0; };
	std::map<std::pair<int, int>, Scratch> g_scratch;                // (device, slot)
	std::map<std::pair<int, int>, bool> g_peerOk;                    // (device, peer) -> device can store into peer's memory
	std::map<const Image2D*, std::pair<void*, uint64_t>> g_pinnedImages;   // library-created images whose storage is page-locked

	// One persistent host thread per additional device: it keeps its CUDA thread state (a fresh thread pays ~10 ms for its
	// first runtime call on a device, every frame) and issues that device's launches while the caller's thread drives the
	// primary device.
	class DeviceWorker
	{
	public:
		DeviceWorker() : thread([this]() { Loop(); }) {}
		~DeviceWorker()
		{
			{ std::lock_guard<std::mutex> lock(mutex); quit = true; }
			wake.notify_all();
			thread.join();
		}
		void Submit(std::function<void()> job)
		{
			{ std::lock_guard<std::mutex> lock(mutex); task = std::move(job); busy = true; }
			wake.notify_all();
		}
		void Wait()
		{
			std::unique_lock<std::mutex> lock(mutex);
			done.wait(lock, [this]() { return !busy; });
		}
	private:
		void Loop()
		{
			for (;;)
			{
				std::function<void()> job;
				{
					std::unique_lock<std::mutex> lock(mutex);
					wake.wait(lock, [this]() { return quit || (busy && task); });
					if (quit) return;
					job = std::move(task);
					task = nullptr;
				}
				job();
				{ std::lock_guard<std::mutex> lock(mutex); busy = false; }
				done.notify_all();
			}
		}
		std::mutex mutex;
		std::condition_variable wake, done;
		std::function<void()> task;
		bool busy = false, quit = false;
		std::thread thread;
	};
	std::map<int, std::unique_ptr<DeviceWorker>> g_workers;         // per device

	thread_local std::string t_lastError;
	thread_local RaylibB200Stats t_lastStats;
	thread_local bool t_haveStats = false;

	const char* Env(const char* name) { const char* v = getenv(name); return (v && *v) ? v : nullptr; }

	// The knobs of DESIGN.md section 9, read from the environment once per process (and again on RaylibB200_ReloadTuning).
	void LoadTuning()
	{
		memset(&g_tuning, 0, sizeof(g_tuning));
		auto u32 = [](const char* name, uint32_t lo, uint32_t hi) -> uint32_t
		{
			const char* v = Env(name);
			return v ? (uint32_t)std::max<long>(lo, std::min<long>(hi, atol(v))) : 0u;
		};
		g_tuning.refillThreshold = u32("RAYLIB_B200_REFILL", 1, 32);
		g_tuning.walkThreshold = u32("RAYLIB_B200_WALK", 1, 32);
		g_tuning.traversalCtas = u32("RAYLIB_B200_TRAVERSAL_CTAS", 1, 16);
		g_tuning.pathsM = u32("RAYLIB_B200_PATHS_M", 1, 4096);
		g_tuning.binOriginBits = Env("RAYLIB_B200_BIN_OBITS") ? (int32_t)u32("RAYLIB_B200_BIN_OBITS", 0, 18) : -1;
		g_tuning.binDirBits = Env("RAYLIB_B200_BIN_DBITS") ? (int32_t)u32("RAYLIB_B200_BIN_DBITS", 0, 4) : -1;
		g_tuning.pipes = u32("RAYLIB_B200_PIPES", 1, 4);
		g_tuning.extendRing = u32("RAYLIB_B200_RING", 0, 1);
		g_tuning.dumpBounces = Env("RAYLIB_B200_DUMP_BOUNCES") ? 1u : 0u;
		g_tuning.dumpTimeline = Env("RAYLIB_B200_DUMP_TIMELINE") ? 1u : 0u;
		g_tuning.fusedPass = u32("RAYLIB_B200_FUSED", 0, 2);
		g_tuning.fusedPathsK = u32("RAYLIB_B200_FUSED_PATHS_K", 1, 1u << 20);
		g_tuning.pooledTraversal = u32("RAYLIB_B200_POOL", 0, 1);
		g_tuning.poolNodeThreshold = u32("RAYLIB_B200_POOL_NODE", 1, 32);
		g_tuning.poolRefill = u32("RAYLIB_B200_POOL_REFILL", 1, 64);
		g_tuning.poolCtas = u32("RAYLIB_B200_POOL_CTAS", 1, 16);
		if (const char* v = Env("RAYLIB_B200_MULTI_MIN_SAMPLES")) g_multiDeviceMinSamples = (uint64_t)std::max<long long>(0, atoll(v));
		g_tuningLoaded = true;
	}

	// Which devices a frame is spread over.  An explicit single device -- RaylibB200_SetDevice, $RAYLIB_B200_DEVICE, or
	// $LOCAL_RANK of a one-process-per-GPU launch -- pins the process to it; otherwise $RAYLIB_B200_DEVICES = all | k |
	// a,b,c; otherwise every visible device, like the reference's Raylib_Render uses every core (renderer.cc:286).
	void ResolveDevices()
	{
		if (g_devicesResolved) return;
		g_devicesResolved = true;
		g_devices.clear();
		const int n = rt_device_count();
		if (n <= 0) return;
		for (const char* name : { "RAYLIB_B200_DEVICE", "LOCAL_RANK" })
			if (const char* v = Env(name))
			{
				const int d = atoi(v);
				if (d < 0 || d >= n)
					fprintf(stderr, "raylib-b200: %s=%s names no CUDA device (%d visible); using device %d\n", name, v, n, ((d % n) + n) % n);
				g_devices.push_back(((d % n) + n) % n);
				return;
			}
		if (const char* v = Env("RAYLIB_B200_DEVICES"))
		{
			if (strchr(v, ','))
			{
				for (const char* c = v; *c;)
				{
					const int d = atoi(c);
					if (d >= 0 && d < n && std::find(g_devices.begin(), g_devices.end(), d) == g_devices.end()) g_devices.push_back(d);
					c = strchr(c, ',');
					if (!c) break;
					++c;
				}
			}
			else if (strcmp(v, "all") != 0)
			{
				const int k = std::max(1, std::min(n, atoi(v)));
				for (int d = 0; d < k; ++d) g_devices.push_back(d);
			}
		}
		if (g_devices.empty()) for (int d = 0; d < n; ++d) g_devices.push_back(d);
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

	// Cheap fingerprint of the sky panorama: the reference reads the image live on every miss (renderer.cc:170-180), so a
	// client may swap or repaint it between frames; the uploaded copy is refreshed when this changes.
	uint64_t SkySignature(ImageHandle sky)
	{
		const Image2D* img = (const Image2D*)sky;
		if (!img) return 0;
		const uint64_t w = img->GetWidth(), h = img->GetHeight(), n = w * h;
		uint64_t sig = 0x9E3779B97F4A7C15ull ^ (w << 32) ^ h;
		const std::vector<Pixel>& px = img->GetPixelArray();
		if (n == 0 || px.size() < n) return sig;
		const uint64_t step = std::max<uint64_t>(1, n / 4096);
		for (uint64_t i = 0; i < n; i += step)
		{
			uint32_t bits[4];
			memcpy(bits, &px[i], 16);
			for (uint32_t b : bits) { sig ^= b; sig *= 0x100000001B3ull; }
		}
		return sig;
	}

	RtRenderContext* ContextOn(int device)
	{
		auto it = g_contexts.find(device);
		if (it != g_contexts.end()) return it->second;
		RtRenderContext* ctx = nullptr;
		if (rt_context_create(device, &ctx) != 0) { RtGpu::SetLastError(std::string("rt_context_create: ") + rt_last_error()); return nullptr; }
		g_contexts[device] = ctx;
		return ctx;
	}

	bool PeerOk(int device, int peer)
	{
		if (device == peer) return true;
		auto it = g_peerOk.find({ device, peer });
		if (it != g_peerOk.end()) return it->second;
		const bool ok = rt_peer_enable(device, peer) == 0;
		g_peerOk[{ device, peer }] = ok;
		return ok;
	}

	// The scene's copy on `device`: flattened + uploaded the first time any device needs it, cloned device-to-device
	// (NVLink peer copy) for every further device -- the host flattens ONCE whatever the number of GPUs.
	const SceneEntry* SceneOn(const Scene* scene, int device)
	{
		auto key = std::make_pair(scene, device);
		auto it = g_scenes.find(key);
		const auto pre = g_prebuilt.find(scene);
		if (it != g_scenes.end() && pre == g_prebuilt.end())
		{
			// lighting that lives in the upload: a changed sky invalidates every device's copy
			const ImageHandle sky = scene->GetSkyPanorama();
			if (sky != it->second.skyHandle || SkySignature(sky) != it->second.skySignature)
			{
				const bool prebuiltScene = false;
				(void)prebuiltScene;
				for (auto e = g_scenes.begin(); e != g_scenes.end();)
				{
					if (e->first.first == scene) { rt_scene_free(e->second.device); e = g_scenes.erase(e); }
					else ++e;
				}
				it = g_scenes.end();
			}
		}
		// a statistics frame walks the reference topology, which plain uploads leave on the host: upload again with it
		if (it != g_scenes.end() && g_collectStats && !it->second.referenceTree)
		{
			for (auto e = g_scenes.begin(); e != g_scenes.end();)
			{
				if (e->first.first == scene) { rt_scene_free(e->second.device); e = g_scenes.erase(e); }
				else ++e;
			}
			it = g_scenes.end();
		}
		if (it != g_scenes.end()) return &it->second;

		// another device already holds it: clone over the fabric
		for (auto& kv : g_scenes)
		{
			if (kv.first.first != scene) continue;
			SceneEntry entry = kv.second;
			entry.device = nullptr;
			const auto t0 = std::chrono::steady_clock::now();
			PeerOk(device, kv.first.second);      // lets cudaMemcpyPeer go straight over NVLink; a staged copy works without it
			if (rt_scene_clone(kv.second.device, device, &entry.device) != 0)
			{
				RtGpu::SetLastError(std::string("rt_scene_clone: ") + rt_last_error());
				return nullptr;
			}
			LOG("[STAT] scene -> GPU %d: cloned from GPU %d in %.1f ms", device, kv.first.second,
				std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
			return &g_scenes.emplace(key, entry).first->second;
		}

		RtFlatScene fresh;
		std::string why;
		const auto t0 = std::chrono::steady_clock::now();
		if (pre == g_prebuilt.end() && !RtFlattenScene(scene, fresh, why)) { RtGpu::SetLastError("cannot flatten scene: " + why); return nullptr; }
		const RtFlatScene& flat = pre != g_prebuilt.end() ? *pre->second : fresh;
		const auto t1 = std::chrono::steady_clock::now();
		SceneEntry entry;
		entry.referenceTree = g_collectStats;
		if (rt_scene_upload(device, &flat.desc, entry.referenceTree ? RT_UPLOAD_REFERENCE_TREE : 0u, &entry.device) != 0)
		{
			RtGpu::SetLastError(std::string("rt_scene_upload: ") + rt_last_error());
			return nullptr;
		}
		const auto t2 = std::chrono::steady_clock::now();
		entry.counts[0] = flat.nodes.empty() ? flat.quantNodes.size() : flat.nodes.size(); entry.counts[1] = flat.triHot.size(); entry.counts[2] = flat.spheres.size();
		entry.counts[3] = flat.cubes.size(); entry.counts[4] = flat.materials.size(); entry.counts[5] = flat.textures.size();
		entry.counts[6] = flat.desc.maxStackDepth; entry.counts[7] = flat.desc.numLeaves;
		if (pre == g_prebuilt.end()) { entry.skyHandle = scene->GetSkyPanorama(); entry.skySignature = SkySignature(entry.skyHandle); }
		LOG("[STAT] scene -> GPU %d: %llu nodes, %llu triangles, %llu spheres, %.1f MB, flatten %.1f ms, upload %.1f ms",
			device, (unsigned long long)entry.counts[0], (unsigned long long)entry.counts[1], (unsigned long long)entry.counts[2],
			(double)flat.HostBytes() / 1.0e6,
			std::chrono::duration<double, std::milli>(t1 - t0).count(), std::chrono::duration<double, std::milli>(t2 - t1).count());
		return &g_scenes.emplace(key, entry).first->second;
	}

	void AddStats(RaylibB200Stats& st, const RtRenderStats& rs)
	{
		st.rayQueries += rs.rayQueries; st.pixelSamples += rs.pixelSamples;
		st.boxTests += rs.boxTests; st.triTests += rs.triTests; st.sphereTests += rs.sphereTests; st.nodeVisits += rs.nodeVisits;
		st.refBoxTests += rs.refBoxTests; st.refTriTests += rs.refTriTests; st.refSphereTests += rs.refSphereTests; st.statRays += rs.statRays;
		st.deviceMs = std::max(st.deviceMs, rs.deviceMs);      // the devices run side by side
		st.extendMs += rs.extendMs; st.extendLaunches += rs.extendLaunches;
		st.nodeIters += rs.nodeIters; st.nodeStep += rs.nodeStep; st.nodeAlive += rs.nodeAlive; st.leafIters += rs.leafIters; st.leafBusy += rs.leafBusy;
		st.gateTests += rs.gateTests; st.cubeTests += rs.cubeTests;
		st.kernelLaunches += rs.kernelLaunches;
		st.passes = std::max(st.passes, rs.passes);
	}
}

namespace RtGpu
{
	int DeviceCount() { return rt_device_count(); }

	int CurrentDevice()
	{
		std::lock_guard<std::recursive_mutex> lock(g_mutex);
		ResolveDevices();
		return g_devices.empty() ? 0 : g_devices[0];
	}

	bool SetDevice(int device)
	{
		std::lock_guard<std::recursive_mutex> lock(g_mutex);
		if (device < 0 || device >= DeviceCount()) { SetLastError("RaylibB200_SetDevice: no such CUDA device"); return false; }
		g_devices.assign(1, device);
		g_devicesResolved = true;
		return true;
	}

	int SetDevices(int count)
	{
		std::lock_guard<std::recursive_mutex> lock(g_mutex);
		const int n = DeviceCount();
		if (n <= 0) { SetLastError("RaylibB200_SetDevices: no CUDA device is available"); return 0; }
		const int k = count <= 0 ? n : std::min(count, n);
		g_devices.clear();
		for (int d = 0; d < k; ++d) g_devices.push_back(d);
		g_devicesResolved = true;
		return k;
	}

	int ActiveDeviceCount()
	{
		std::lock_guard<std::recursive_mutex> lock(g_mutex);
		ResolveDevices();
		return (int)g_devices.size();
	}

	void ReloadTuning() { std::lock_guard<std::recursive_mutex> lock(g_mutex); LoadTuning(); }
	const RtTuning& Tuning()
	{
		std::lock_guard<std::recursive_mutex> lock(g_mutex);
		if (!g_tuningLoaded) LoadTuning();
		return g_tuning;
	}

	void SetFrameSeed(uint64_t seed) { g_frameSeed = seed; }
	uint64_t FrameSeed() { return g_frameSeed; }
	void SetCollectStats(bool enable) { g_collectStats = enable; }
	void SetTimeStages(bool enable) { g_timeStages = enable; }
	void SetSamplesPerPass(uint32_t samples) { g_samplesPerPass = samples; }
	void SetPipes(uint32_t pipes) { g_pipes = pipes; }
	void SetFusedPass(uint32_t mode) { g_fusedPass = std::min(mode, 2u); }

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
		return ContextOn(CurrentDevice());
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
		const SceneEntry* entry = SceneOn(scene, CurrentDevice());
		if (!entry) return nullptr;
		if (outCounts8) memcpy(outCounts8, entry->counts, sizeof(entry->counts));
		return entry->device;
	}

	void ForgetHostImage(const Image2D* image)
	{
		std::lock_guard<std::recursive_mutex> lock(g_mutex);
		auto it = g_pinnedImages.find(image);
		if (it == g_pinnedImages.end()) return;
		if (it->second.first) rt_host_unregister(it->second.first);
		g_pinnedImages.erase(it);
	}

	void ReleaseAll()
	{
		std::lock_guard<std::recursive_mutex> lock(g_mutex);
		for (auto& kv : g_pinnedImages) if (kv.second.first) rt_host_unregister(kv.second.first);
		g_pinnedImages.clear();
		for (auto& kv : g_scenes) rt_scene_free(kv.second.device);
		g_scenes.clear();
		g_prebuilt.clear();
		for (auto& kv : g_scratch) rt_device_free(kv.first.first, kv.second.ptr);
		g_scratch.clear();
		g_workers.clear();      // joins the per-device threads
		for (auto& kv : g_contexts) rt_context_destroy(kv.second);
		g_contexts.clear();
		g_peerOk.clear();
	}

	// Page-locks the storage of a library-created Image2D once, so that the read-back of every later frame is one
	// full-speed DMA instead of a staged pageable copy (a third of the time of millisecond frames otherwise).
	static void PinHostImage(Image2D* image, uint64_t bytes)
	{
		void* ptr = image->MutablePixels();
		auto it = g_pinnedImages.find(image);
		if (it != g_pinnedImages.end())
		{
			if ((it->second.first == ptr || it->second.first == nullptr) && it->second.second == bytes) return;
			if (it->second.first) rt_host_unregister(it->second.first);
			g_pinnedImages.erase(it);
		}
		if (rt_host_register(ptr, bytes) == 0) g_pinnedImages[image] = { ptr, bytes };
		else
		{
			// remember the refusal (an entry with no registration) so that the attempt is not repeated on every frame
			g_pinnedImages[image] = { nullptr, bytes };
			fprintf(stderr, "raylib-b200: could not page-lock a %llu-byte image (%s); read-backs into it use pageable copies\n", (unsigned long long)bytes, rt_last_error());
		}
	}

	bool Render(const RendererSettings* settings, const Scene* scene, const Camera* camera,
	            Image2D* hostImage, void* deviceImage, void* deviceShard,
	            uint32_t shardRank, uint32_t shardCount, uint32_t renderModeOverride, void* stream, bool pinHostImage)
	{
		std::lock_guard<std::recursive_mutex> lock(g_mutex);
		const auto wall0 = std::chrono::steady_clock::now();
		t_lastError.clear();
		if (!settings || !scene || !camera) { SetLastError("Raylib_Render: null settings/scene/camera"); return false; }
		if (settings->viewportWidth == 0 || settings->viewportHeight == 0) { SetLastError("Raylib_Render: empty viewport"); return false; }
		if (settings->renderMode >= RAYLIB_RENDERMODE_MAX && renderModeOverride == 0) { SetLastError("Raylib_Render: invalid renderMode"); return false; }
		if (DeviceCount() <= 0)
		{
			SetLastError("no CUDA device is available; this library renders on the GPU only (no CPU path)");
			return false;
		}
		ResolveDevices();
		const int primary = CurrentDevice();

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
		params.samplesPerPass = g_samplesPerPass;
		params.pipes = g_pipes;
		params.collectStats = g_collectStats ? 1u : 0u;
		params.timeStages = g_timeStages ? 1u : 0u;
		params.tuning = Tuning();
		if (g_fusedPass) params.tuning.fusedPass = g_fusedPass;
		if (g_prebuilt.find(scene) == g_prebuilt.end())
		{
			// the sun as the Scene object holds it NOW (the reference reads it on every miss, renderer.cc:160-191)
			vec3 illuminance, direction;
			scene->GetSun(illuminance, direction);
			params.lightingOverride = 1;
			params.sunIlluminance[0] = illuminance.x; params.sunIlluminance[1] = illuminance.y; params.sunIlluminance[2] = illuminance.z;
			params.sunDirection[0] = direction.x; params.sunDirection[1] = direction.y; params.sunDirection[2] = direction.z;
		}

		// ---- which devices render this frame -------------------------------------------------------------
		// A whole frame (read back into a host image, or left in a row-major device image on the primary device) is spread over the
		// active devices by interleaved 16x16 tiles; shards, debug views and statistics frames stay on the primary device.
		const uint64_t pixelSamples = (uint64_t)params.width * params.height * (uint64_t)std::max(1, params.samplesPerPixel);
		std::vector<int> devices(1, primary);
		const bool wholeFrame = hostImage || (deviceImage && !deviceShard && shardCount <= 1);
		if (wholeFrame && params.renderMode == 0u && !g_collectStats && !g_timeStages && g_devices.size() > 1 &&
		    pixelSamples >= g_multiDeviceMinSamples)
			devices = g_devices;
		const uint32_t n = (uint32_t)devices.size();

		// hostImage: always the whole frame; deviceShard / deviceImage: the tiles of (shardRank, shardCount)
		if (hostImage) { shardRank = 0; shardCount = 1; }
		if (shardCount == 0) shardCount = 1;
		if (shardRank >= shardCount) { SetLastError("Raylib_Render: shardRank >= shardCount"); return false; }

		struct PerDevice
		{
			int device = 0;
			RtRenderContext* ctx = nullptr;
			const RtDeviceScene* scene = nullptr;
			void* shard = nullptr;         // tile-major shard buffer (gather fall-back only)
			RtRenderParams params;
			RtRenderStats stats;
			int rc = 0;
			std::string error;
		};
		std::vector<PerDevice> work(n);
		for (uint32_t i = 0; i < n; ++i)
		{
			PerDevice& w = work[i];
			w.device = devices[i];
			const SceneEntry* entry = SceneOn(scene, w.device);
			if (!entry) return false;
			w.scene = entry->device;
			w.ctx = ContextOn(w.device);
			if (!w.ctx) return false;
		}

		// Final pixels go straight into the row-major frame on the primary device -- its own kernels store locally, the
		// other devices store through peer mappings over NVLink: the last accumulate IS the gather -- unless the caller
		// asked for a tile-major shard buffer, or some device cannot address the primary's memory (then: shard buffers +
		// peer copies + one de-interleave pass on the primary).
		const uint64_t imageBytes = (uint64_t)params.width * params.height * 16ull;
		void* imageBuffer = nullptr;
		bool direct = true;
		for (uint32_t i = 1; i < n; ++i) direct = direct && PeerOk(work[i].device, primary);
		if (!deviceShard)
		{
			imageBuffer = deviceImage ? deviceImage : ScratchBuffer(primary, 1, imageBytes);
			if (!imageBuffer) { SetLastError(std::string("device allocation failed: ") + rt_last_error()); return false; }
		}
		const uint64_t shardBytes = (uint64_t)rt_shard_tile_capacity(params.width, params.height, n) * RT_TILE_PIXELS * 16ull;
		for (uint32_t i = 0; i < n; ++i)
		{
			PerDevice& w = work[i];
			w.params = params;
			w.params.shardRank = n > 1 ? i : shardRank;
			w.params.shardCount = n > 1 ? n : shardCount;
			if (deviceShard) w.shard = deviceShard;
			else if (direct) w.params.imageOut = imageBuffer;
			else
			{
				w.shard = ScratchBuffer(w.device, 0, shardBytes);
				if (!w.shard) { SetLastError(std::string("device allocation failed: ") + rt_last_error()); return false; }
			}
		}

		// one host thread per device issues that device's launches (about ten thousand per 4K frame) and waits for it;
		// they share nothing but the read-only camera block
		auto renderOn = [&cam](PerDevice& w, void* onStream)
		{
			w.rc = rt_render_shard(w.ctx, w.scene, &cam, &w.params, w.shard, onStream, &w.stats);
			if (w.rc != 0) w.error = rt_last_error();
		};
		{
			std::vector<DeviceWorker*> helpers;
			for (uint32_t i = 1; i < n; ++i)
			{
				std::unique_ptr<DeviceWorker>& worker = g_workers[work[i].device];
				if (!worker) worker.reset(new DeviceWorker);
				PerDevice* w = &work[i];
				worker->Submit([renderOn, w]() { renderOn(*w, nullptr); });
				helpers.push_back(worker.get());
			}
			renderOn(work[0], n > 1 ? nullptr : stream);
			for (DeviceWorker* h : helpers) h->Wait();
		}
		for (const PerDevice& w : work)
			if (w.rc != 0) { SetLastError("rt_render_shard (GPU " + std::to_string(w.device) + "): " + w.error); return false; }

		if (n > 1 && !direct)
		{
			void* gathered = ScratchBuffer(primary, 3, shardBytes * n);
			if (!gathered) { SetLastError(std::string("device allocation failed: ") + rt_last_error()); return false; }
			for (uint32_t i = 0; i < n; ++i)
				if (rt_copy_peer(primary, (char*)gathered + shardBytes * i, work[i].device, work[i].shard, shardBytes) != 0)
				{
					SetLastError(std::string("peer copy of a shard failed: ") + rt_last_error());
					return false;
				}
			if (rt_assemble(primary, gathered, n, params.width, params.height, imageBuffer, nullptr) != 0 || rt_stream_sync(primary, nullptr) != 0)
			{
				SetLastError(std::string("rt_assemble: ") + rt_last_error());
				return false;
			}
		}

		uint64_t d2h = 0;
		if (hostImage)
		{
			if (pinHostImage) PinHostImage(hostImage, imageBytes);
			// Pixel is four packed floats, same as the device float4 image
			if (rt_copy_to_host(primary, hostImage->MutablePixels(), imageBuffer, imageBytes, n > 1 ? nullptr : stream) != 0)
			{
				SetLastError(std::string("device -> host copy failed: ") + rt_last_error());
				return false;
			}
			d2h = imageBytes;
		}

		RaylibB200Stats st;
		memset(&st, 0, sizeof(st));
		for (const PerDevice& w : work) AddStats(st, w.stats);
		st.totalMs = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - wall0).count();
		st.h2dBytes = (sizeof(RtCamera) + sizeof(RtRenderParams)) * n;
		st.d2hBytes = d2h;
		st.device = (uint32_t)primary;
		st.devicesUsed = n;
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
	CHECK(world->GetAccelStruct() != nullptr || RtGpu::Prebuilt(world) != nullptr);      // scenes read from the flattened-scene cache have no object graph

	if (settings->viewportWidth != outImage->GetWidth() || settings->viewportHeight != outImage->GetHeight())
		outImage->Reallocate(settings->viewportWidth, settings->viewportHeight);

	if (!RtGpu::Render(settings, world, camera, outImage, nullptr, nullptr, 0, 1, 0, nullptr, RtIsLibraryImage(outImage)))
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
