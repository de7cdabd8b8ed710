// rt_device_abi.h -- thin C ABI between the host C++ of libraylib and its CUDA
// translation units (csrc/device/*.cu).  Plain pointers and sizes only.
//
// What each call replaces in the reference:
//   rt_scene_upload   nothing (the reference renders from the heap objects); it is the
//                     device-side half of Raylib_FinalizeScene (raylib/raylib.cc:212-215)
//   rt_render_shard   Renderer::RenderScene + ThreadPool + GenerateCell + TraceScene +
//                     every Hit/Scatter below them (render/renderer.cc:62-356)
//   rt_assemble       the implicit "all threads write one Image2D" of renderer.cc:248,
//                     needed here because GPUs own interleaved tiles
//   rt_trace_closest  BVHNode::Hit on caller-supplied rays (geom/bvh.cc:82-107); used by the
//                     parity tests and by the primary-visibility debug export
#pragma once
#include <stdint.h>
#include "rt_scene_format.h"

#ifdef __cplusplus
extern "C" {
#endif

// Exported from libraylib_b200.so (the rest of the device code has hidden visibility): a host written in another
// language can drive the kernels through this header alone.
#if defined(__GNUC__)
#define RT_DEVICE_API __attribute__((visibility("default")))
#else
#define RT_DEVICE_API
#endif

typedef struct RtDeviceScene RtDeviceScene;     // opaque, one per (scene, device)
typedef struct RtRenderContext RtRenderContext; // opaque: path-state arenas + queues for one device

#define RT_TILE_W 16
#define RT_TILE_H 16
#define RT_TILE_PIXELS (RT_TILE_W * RT_TILE_H)

// Tuning / development knobs, read from the environment ONCE per process by the host (RtGpu::Tuning, gpu_state.cc) and
// again on RaylibB200_ReloadTuning(); 0 = the built-in default of every field.
typedef struct RtTuning
{
	uint32_t refillThreshold;     // RAYLIB_B200_REFILL: a warp refills its idle lanes below this many live rays (default 20)
	uint32_t walkThreshold;       // RAYLIB_B200_WALK: node phase yields to the leaf phase below this many steppable lanes (20)
	uint32_t traversalCtas;       // RAYLIB_B200_TRAVERSAL_CTAS: CTAs per SM of k_extend / k_shadow while two pipes are active (7)
	uint32_t pathsM;              // RAYLIB_B200_PATHS_M: Mi paths in flight over all pipes (64)
	int32_t  binOriginBits;       // RAYLIB_B200_BIN_OBITS: -1 = automatic (12, 0 below 65 536 primitives)
	int32_t  binDirBits;          // RAYLIB_B200_BIN_DBITS: -1 = automatic (2)
	uint32_t pipes;               // RAYLIB_B200_PIPES (0 = default 2)
	uint32_t extendRing;          // RAYLIB_B200_RING
	uint32_t dumpBounces, dumpTimeline;   // RAYLIB_B200_DUMP_BOUNCES / _DUMP_TIMELINE (with timeStages)
	uint32_t fusedPass;           // RAYLIB_B200_FUSED: 0 = automatic (frames of at most fusedPathsK Ki paths), 1 = never, 2 = always:
	                              //   the whole pass as ONE cooperative launch with grid-wide barriers between the stages (k_pass_fused)
	uint32_t fusedPathsK;         // RAYLIB_B200_FUSED_PATHS_K: the automatic limit in Ki paths (width x height x samples; default 1024)
	uint32_t pooledTraversal;     // RAYLIB_B200_POOL: 1 = k_extend_pool (a warp regroups its rays at every step), 0 = k_extend
	uint32_t poolNodeThreshold;   // RAYLIB_B200_POOL_NODE: the pooled kernel takes a node step while this many rays can step (default 24)
	uint32_t poolRefill;          // RAYLIB_B200_POOL_REFILL: ... refills once this many of a warp's ray slots are free (default 24)
	uint32_t poolCtas;            // RAYLIB_B200_POOL_CTAS: CTAs per SM of the pooled kernel (0 = what fits)
} RtTuning;

typedef struct RtRenderParams
{
	uint32_t width, height;
	int32_t  samplesPerPixel;
	int32_t  maxPathLength;
	float    rayTMin;
	uint32_t renderMode;          // ERenderMode
	uint64_t frameSeed;
	uint32_t shardRank;           // this device renders tiles t with t % shardCount == shardRank
	uint32_t shardCount;
	uint32_t samplesPerPass;      // 0 = choose automatically
	uint32_t collectStats;        // 1 = count box/triangle/sphere tests (slower; for the roofline figures)
	uint32_t timeStages;          // 1 = bracket every k_extend launch with CUDA events (stats->extendMs)
	uint32_t pipes;               // passes in flight at once, each on its own stream and arena (0 = default 2, max 4)
	void*    auxShardOut;         // renderMode RT_RENDERMODE_AUX only: second shard buffer (microsurface normals)
	void*    imageOut;            // optional: row-major W x H float4 frame (this or a peer GPU's memory); final pixels go there
	                              //   directly and deviceShardOut may be NULL
	// distant lighting read LIVE from the Scene object at every render (the reference reads Scene::GetSun on every miss,
	// renderer.cc:160-191, so a client may change it between frames); lightingOverride = 0 keeps the uploaded values
	uint32_t lightingOverride;
	float    sunIlluminance[3];
	float    sunDirection[3];
	RtTuning tuning;
} RtRenderParams;

// Internal render modes beyond ERenderMode (raylib_types.h: 0..6)
#define RT_RENDERMODE_PRIMARY_EXPORT 100u   // pixel = (t, leaf-rank bits, bu, bv) of the primary hit
#define RT_RENDERMODE_AUX            101u   // denoiser inputs in ONE primary-hit pass: out = Albedo view, auxShardOut = MicrosurfaceNormal view

typedef struct RtRenderStats
{
	uint64_t rayQueries;          // scene-level queries: camera + scattered + sun-shadow + debug second rays
	uint64_t pixelSamples;        // pixels * spp rendered by this shard
	uint64_t boxTests, triTests, sphereTests;   // only when collectStats
	uint64_t nodeVisits;          // RtNode records fetched (collectStats)
	uint64_t refBoxTests, refTriTests, refSphereTests;   // what the reference's exhaustive traversal does for the
	uint64_t statRays;                                   //   statRays closest-hit rays of k_extend (collectStats)
	double   deviceMs;            // CUDA-event time, first launch -> shard buffer complete
	double   extendMs;            // sum of the k_extend launch durations (timeStages)
	uint32_t extendLaunches;
	uint32_t kernelLaunches;
	uint32_t passes;
	uint32_t tilesRendered;
	uint32_t pad;
	// collectStats: lane-iterations of k_extend's traversal loop (every lane of a warp counts each iteration it sits through)
	uint64_t nodeIters, nodeStep, nodeAlive;   // node phase: all lanes / lanes that stepped / lanes that owned a ray
	uint64_t leafIters, leafBusy;              // leaf phase: all lanes / lanes that tested a leaf
	uint64_t gateTests, cubeTests;             // collectStats: exact gate-box tests of accepted hits (32 B each), cube tests
} RtRenderStats;

// ---- device management -------------------------------------------------------
RT_DEVICE_API int  rt_device_count(void);                       // 0 when no CUDA device is usable
RT_DEVICE_API const char* rt_last_error(void);                  // thread-local message of the last failing call

// ---- scene --------------------------------------------------------------------
// flags: RT_UPLOAD_REFERENCE_TREE also uploads the reference topology (refNodes, 64 B per reference BVHNode: 640 MB for 10 M
// triangles), which only the statistics build walks (reference-work counters of the roofline accounting); without it those
// counters stay zero.  Per-triangle rank / gate words travel inside the hot records; the separate host arrays are not uploaded.
#define RT_UPLOAD_REFERENCE_TREE 1u
RT_DEVICE_API int  rt_scene_upload(int device, const RtSceneDesc* desc, uint32_t flags, RtDeviceScene** outScene);   // 0 = ok
RT_DEVICE_API int  rt_scene_has_reference_tree(const RtDeviceScene* scene);
// Copy of an uploaded scene on another device, made with device-to-device peer copies (NVLink) -- no second pass through host memory.
RT_DEVICE_API int  rt_scene_clone(const RtDeviceScene* scene, int device, RtDeviceScene** outScene);
RT_DEVICE_API void rt_scene_free(RtDeviceScene* scene);
RT_DEVICE_API uint64_t rt_scene_device_bytes(const RtDeviceScene* scene);

// ---- rendering ----------------------------------------------------------------
// Number of tiles (RT_TILE_W x RT_TILE_H pixels) a shard owns; every shard's buffer is padded to
// rt_shard_tile_capacity so gathers move equal-sized slabs.
RT_DEVICE_API uint32_t rt_shard_tile_capacity(uint32_t width, uint32_t height, uint32_t shardCount);

RT_DEVICE_API int  rt_context_create(int device, RtRenderContext** outCtx);
RT_DEVICE_API void rt_context_destroy(RtRenderContext* ctx);

// Renders this shard's tiles into `deviceShardOut` (device memory, rt_shard_tile_capacity * RT_TILE_PIXELS
// float4 RGBA pixels, tile-major).  `stream` is a cudaStream_t (0 = default stream).  Asynchronous with
// respect to the host unless stats != NULL (then it synchronises the stream before returning).
RT_DEVICE_API int  rt_render_shard(RtRenderContext* ctx, const RtDeviceScene* scene, const RtCamera* camera,
                     const RtRenderParams* params, void* deviceShardOut, void* stream, RtRenderStats* stats);

// De-interleaves `shardCount` gathered shard buffers (concatenated, rank-major) into a row-major W x H
// float4 image on the device.
RT_DEVICE_API int  rt_assemble(int device, const void* deviceShards, uint32_t shardCount, uint32_t width, uint32_t height,
                 void* deviceImageOut, void* stream);

// Closest hit for caller-provided rays (host arrays; 8 floats per ray: o.xyz, time, d.xyz, unused).
// outRank = global in-order leaf rank or -1, outT = hit distance or 0.
RT_DEVICE_API int  rt_trace_closest(RtRenderContext* ctx, const RtDeviceScene* scene, const float* hostRays, int64_t numRays,
                      float tMin, int32_t* hostOutRank, float* hostOutT, RtRenderStats* stats);

// Image2D::PostProcess (raylib/render/image.cc:44-103) on a device-resident W x H RGBA float4 image, in place:
// max-luminance reduction, extended Reinhard on luminance, clamp to white, gamma 2.2.  If outArgb8 is not NULL it
// also receives Pixel::ToUint32 of every pixel (raylib/render/image.h:57-64) -- a quarter of the bytes to read back.
RT_DEVICE_API int  rt_postprocess(int device, void* deviceImage, uint32_t width, uint32_t height, uint32_t* deviceOutArgb8,
                    float* hostOutMaxWhite, void* stream);

// CUDA IPC handles (64 bytes) for the one-process-per-GPU launch: rank 0 exports its frame, the others map it and
// render their tiles straight into it over NVLink (RtRenderParams.imageOut).
RT_DEVICE_API int  rt_ipc_export(int device, void* devicePtr, unsigned char* outHandle64);
RT_DEVICE_API int  rt_ipc_open(int device, const unsigned char* handle64, void** outPtr);
RT_DEVICE_API int  rt_ipc_close(int device, void* ptr);

// Several devices in ONE process (Raylib_Render over all visible GPUs): peer access so that device `device` can store
// into memory of `peer` (the frame on device 0), and a peer copy for boxes without peer access.
RT_DEVICE_API int  rt_peer_enable(int device, int peer);       // 0 = `device` may now address memory of `peer`
RT_DEVICE_API int  rt_copy_peer(int dstDevice, void* dst, int srcDevice, const void* src, uint64_t bytes);
// Page-locks caller memory (the Image2D storage a frame is read back into) for full-speed D2H copies.
RT_DEVICE_API int  rt_host_register(void* ptr, uint64_t bytes);
RT_DEVICE_API int  rt_host_unregister(void* ptr);
RT_DEVICE_API int  rt_device_free_bytes(int device, uint64_t* outFree, uint64_t* outTotal);

// The device's transcendentals (include/rt_libm.h) on caller-supplied arguments, for the parity tests against the host C
// library: fn 0..9 = sinf cosf tanf asinf acosf atanf expf logf powf(x, y) atan2f(x, y); hostY may be NULL for 0..7.
RT_DEVICE_API int  rt_libm_eval(int device, int fn, const float* hostX, const float* hostY, float* hostOut, uint64_t n);

// Plain device-memory helpers so that host C++ never includes cuda_runtime.h.
RT_DEVICE_API int  rt_device_alloc(int device, uint64_t bytes, void** outPtr);
RT_DEVICE_API void rt_device_free(int device, void* ptr);
RT_DEVICE_API int  rt_copy_to_host(int device, void* hostDst, const void* deviceSrc, uint64_t bytes, void* stream);
RT_DEVICE_API int  rt_copy_to_device(int device, void* deviceDst, const void* hostSrc, uint64_t bytes, void* stream);
RT_DEVICE_API int  rt_stream_sync(int device, void* stream);

#ifdef __cplusplus
}
#endif
