// oracle/oracle_driver.cc -- TEST INFRASTRUCTURE (not product code).
//
// Links against the UNMODIFIED reference sources (compiled from /root/reference
// by oracle/Makefile) and drives them deterministically:
//   * oracle_render        owns the pixel loop of raylib/render/renderer.cc:229-270
//                          and calls the reference's own Camera::GetCameraRay
//                          (raylib/render/camera.h:44-53), TraceScene (:114-208)
//                          and TraceSceneDebugMode (:62-111), re-keying the
//                          shadowed RNG per (pixel, sample).
//   * oracle_primary_hits  walks the reference's BVHNode/StaticMesh objects
//                          (logic of raylib/geom/bvh.cc:82-107 and
//                          raylib/geom/static_mesh.cc:97-109) with the reference's
//                          own AABB::Hit and leaf Hit to label every primary hit
//                          with its in-order leaf rank, and counts the box /
//                          triangle / sphere tests the reference performs.
//   * oracle_native_render times Renderer::RenderScene (renderer.cc:273-356), the
//                          reference's own thread-pool pixel loop.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference leg may load the library this file is built into.
//
// Compiled with -fno-access-control (StaticMesh::bvh/bounds and
// Scene::accelStruct are private in the reference headers).
#include "raylib.h"
#include "geom/scene.h"
#include "geom/bvh.h"
#include "geom/static_mesh.h"
#include "geom/sphere.h"
#include "geom/triangle.h"
#include "geom/cube.h"
#include "render/camera.h"
#include "render/image.h"
#include "render/material.h"
#include "render/renderer.h"
#include "core/random.h"

#include <atomic>
#include <chrono>
#include <thread>
#include <unordered_map>
#include <mutex>

thread_local OracleRngCtx g_oracleRng = { RT_RNG_DEFAULT_BVH_KEY, 0, 0 };

// Signatures of the reference's external-linkage functions in renderer.cc.
struct RayPayload { int32 maxRecursion; float rayTMin; };          // renderer.cc:48-51
vec3 TraceScene(const ray& cameraRay, const Scene* world, const RayPayload& settings);  // :202-208
vec3 TraceSceneDebugMode(const ray& pathRay, const Scene* world, const RayPayload& settings, ERenderMode debugMode); // :62

// ---------------------------------------------------------------------------
// Ray counting: a pass-through root that counts GetAccelStruct()->Hit() calls.

static thread_local uint64_t t_rayQueries = 0;

class CountingRoot : public BVHNode
{
public:
	CountingRoot(HitableList* one, const BVHNode* inReal) : BVHNode(one, 0.0f, 0.0f), real(inReal) {}
	bool Hit(const ray& r, float tMin, float tMax, HitResult& out) const override
	{
		++t_rayQueries;
		return real->Hit(r, tMin, tMax, out);
	}
	const BVHNode* real;
};

struct ScopedCountingRoot
{
	ScopedCountingRoot(Scene* inScene) : scene(inScene)
	{
		OracleRngCtx saved = g_oracleRng;
		real = scene->accelStruct;
		list.hitables.push_back(real);
		counting = new CountingRoot(&list, real);
		scene->accelStruct = counting;
		g_oracleRng = saved;
	}
	~ScopedCountingRoot()
	{
		scene->accelStruct = real;
		counting->left = counting->right = nullptr;
		delete counting;
	}
	Scene* scene;
	BVHNode* real;
	CountingRoot* counting;
	HitableList list;
};

// ---------------------------------------------------------------------------

struct OracleRenderStats
{
	uint64_t rayQueries;   // scene-level Hit() calls: camera + scattered + sun-shadow + debug second rays
	uint64_t rngDraws;
	double   seconds;
	int32_t  threads;
	int32_t  debugbreaks;
};

extern "C" long oracle_debugbreak_count(void);

template<typename RowFn>
static void ParallelRows(int32 y0, int32 y1, int nthreads, RowFn fn)
{
	if (nthreads <= 0) nthreads = (int)std::max(1u, std::thread::hardware_concurrency());
	std::atomic<int32> next(y0);
	std::vector<std::thread> pool;
	for (int i = 0; i < nthreads; ++i)
	{
		pool.emplace_back([&, i]() {
			for (;;) {
				int32 y = next.fetch_add(1);
				if (y >= y1) break;
				fn(y, i);
			}
		});
	}
	for (auto& th : pool) th.join();
}

extern "C" {

void oracle_rng_reset(uint64_t key) { g_oracleRng.key = key; g_oracleRng.ctr = 0; }

// Called by scenes/scenes.cc right before every top-level BVH build
// (StaticMesh::Finalize, Raylib_FinalizeScene) so the random split axes
// (raylib/geom/bvh.cc:43) are reproducible.
void scene_hook_before_bvh_build(void) { oracle_rng_reset(RT_RNG_DEFAULT_BVH_KEY); }

int32_t oracle_hardware_threads(void) { return (int32_t)std::max(1u, std::thread::hardware_concurrency()); }

// Deterministic restatement of the GenerateCell pixel loop around the reference's own
// camera + TraceScene.  The crop [x0,x1) x [y0,y1) limits the work (pixels outside are untouched).
void oracle_render_region(
	const RendererSettings* settings, SceneHandle sceneH, CameraHandle cameraH, ImageHandle imageH,
	uint64_t frameSeed, int32_t nthreads,
	int32_t x0, int32_t y0, int32_t x1, int32_t y1,
	OracleRenderStats* outStats)
{
	Scene* scene = (Scene*)sceneH;
	const Camera* camera = (const Camera*)cameraH;
	Image2D* image = (Image2D*)imageH;
	if (settings->viewportWidth != image->GetWidth() || settings->viewportHeight != image->GetHeight())
		image->Reallocate(settings->viewportWidth, settings->viewportHeight);

	const int32 W = (int32)image->GetWidth(), H = (int32)image->GetHeight();
	const float imageWidth = (float)W, imageHeight = (float)H;
	x0 = std::max(0, x0); y0 = std::max(0, y0); x1 = std::min(W, x1); y1 = std::min(H, y1);

	RayPayload rt{ settings->maxPathLength, settings->rayTMin };
	const int32 SPP = std::max(1, settings->samplesPerPixel);
	const bool pathTrace = settings->renderMode == RAYLIB_RENDERMODE_Default;

	ScopedCountingRoot counting(scene);
	std::atomic<uint64_t> rays(0), draws(0);
	long breaks0 = oracle_debugbreak_count();
	auto tStart = std::chrono::steady_clock::now();
	int usedThreads = nthreads > 0 ? nthreads : oracle_hardware_threads();

	ParallelRows(y0, y1, usedThreads, [&](int32 y, int) {
		t_rayQueries = 0; g_oracleRng.draws = 0;
		for (int32 x = x0; x < x1; ++x)
		{
			const uint32_t pixel = (uint32_t)(y * W + x);
			if (pathTrace)
			{
				vec3 accum(0.0f, 0.0f, 0.0f);
				for (int32 s = 0; s < SPP; ++s)
				{
					g_oracleRng.key = rt_sample_key(frameSeed, pixel, (uint32_t)s);
					g_oracleRng.ctr = 0;
					RNG randomsAA(0);
					float u = (float)x / imageWidth;
					float v = (float)y / imageHeight;
					if (s != 0) {
						u += (randomsAA.Peek() - 0.5f) * 2.0f / imageWidth;
						v += (randomsAA.Peek() - 0.5f) * 2.0f / imageHeight;
					}
					ray cameraRay = camera->GetCameraRay(u, v);
					vec3 Li = TraceScene(cameraRay, scene, rt);
					accum += Li;
				}
				accum /= (float)SPP;
				image->SetPixel(x, y, Pixel(accum.x, accum.y, accum.z));
			}
			else
			{
				g_oracleRng.key = rt_sample_key(frameSeed, pixel, 0u);
				g_oracleRng.ctr = 0;
				float u = (float)x / imageWidth;
				float v = (float)y / imageHeight;
				ray cameraRay = camera->GetCameraRay(u, v);
				vec3 dbg = TraceSceneDebugMode(cameraRay, scene, rt, (ERenderMode)settings->renderMode);
				image->SetPixel(x, y, Pixel(dbg.x, dbg.y, dbg.z));
			}
		}
		rays += t_rayQueries; draws += g_oracleRng.draws;
	});

	double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - tStart).count();
	if (outStats)
	{
		outStats->rayQueries = rays.load();
		outStats->rngDraws = draws.load();
		outStats->seconds = sec;
		outStats->threads = usedThreads;
		outStats->debugbreaks = (int32_t)(oracle_debugbreak_count() - breaks0);
	}
}

void oracle_render(
	const RendererSettings* settings, SceneHandle sceneH, CameraHandle cameraH, ImageHandle imageH,
	uint64_t frameSeed, int32_t nthreads, OracleRenderStats* outStats)
{
	oracle_render_region(settings, sceneH, cameraH, imageH, frameSeed, nthreads,
		0, 0, (int32_t)settings->viewportWidth, (int32_t)settings->viewportHeight, outStats);
}

// The reference's own thread-pool renderer (non-deterministic work split, same RNG shim).
double oracle_native_render(const RendererSettings* settings, SceneHandle sceneH, CameraHandle cameraH, ImageHandle imageH,
	OracleRenderStats* outStats)
{
	Scene* scene = (Scene*)sceneH;
	ScopedCountingRoot counting(scene);
	// worker threads are detached pool threads; their thread_local counters are not collectable,
	// so ray counts for this mode come from oracle_render on the same workload.
	auto tStart = std::chrono::steady_clock::now();
	Renderer renderer;
	renderer.RenderScene(settings, scene, (const Camera*)cameraH, (Image2D*)imageH);
	double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - tStart).count();
	if (outStats) { outStats->rayQueries = 0; outStats->rngDraws = 0; outStats->seconds = sec;
		outStats->threads = oracle_hardware_threads(); outStats->debugbreaks = 0; }
	return sec;
}

} // extern "C"

// ---------------------------------------------------------------------------
// Primary-hit labelling walk.

enum WalkKind : int32_t { WK_NODE = 0, WK_MESH = 1, WK_TRI = 2, WK_SPHERE = 3, WK_CUBE = 4, WK_OTHER = 5, WK_LIST = 6 };

struct WalkNode
{
	const Hitable* obj;
	int32_t kind;
	int32_t left, right;   // WalkNode indices (NODE), child (MESH: left = mesh bvh)
	int32_t rank;          // in-order leaf rank for primitives
	std::vector<int32_t> members;   // WK_LIST: WalkNode index of every member, in the list's own order
};

struct WalkTree
{
	std::vector<WalkNode> nodes;
	int32_t numLeaves = 0;
	int32_t maxDepth = 0;

	int32_t Build(const Hitable* h, int depth)
	{
		maxDepth = std::max(maxDepth, depth);
		int32_t me = (int32_t)nodes.size();
		nodes.push_back(WalkNode{ h, WK_OTHER, -1, -1, -1, {} });
		if (const BVHNode* n = dynamic_cast<const BVHNode*>(h))
		{
			nodes[me].kind = WK_NODE;
			int32_t l = Build(n->left, depth + 1);
			int32_t r = (n->left == n->right) ? -1 : Build(n->right, depth + 1);
			nodes[me].left = l; nodes[me].right = r;
		}
		else if (const StaticMesh* m = dynamic_cast<const StaticMesh*>(h))
		{
			nodes[me].kind = WK_MESH;
			int32_t c = Build(m->bvh, depth + 1);
			nodes[me].left = c;
		}
		else if (const HitableList* list = dynamic_cast<const HitableList*>(h))
		{
			// a raw list as a scene element: its members take consecutive ranks, spheres in reverse list order first, then
			// the other members in list order -- the labelling under which "minimum t, ties to the highest rank" is what the
			// reference's scan with a shrinking upper bound selects (strict range test of spheres, inclusive of triangles
			// and cubes; geom/hit.cc:34-50).  The product's flattener numbers list members the same way.
			nodes[me].kind = WK_LIST;
			std::vector<int32_t> ids(list->hitables.size(), -1);
			for (size_t i = 0; i < list->hitables.size(); ++i) { int32_t c = Build(list->hitables[i], depth + 1); ids[i] = c; }
			// Build() numbered the members in list order; renumber inside the same range
			int32_t base = numLeaves - (int32_t)ids.size();
			int32_t next = base;
			for (size_t i = ids.size(); i-- > 0;) if (nodes[ids[i]].kind == WK_SPHERE) nodes[ids[i]].rank = next++;
			for (size_t i = 0; i < ids.size(); ++i) if (nodes[ids[i]].kind != WK_SPHERE) nodes[ids[i]].rank = next++;
			nodes[me].members = ids;
		}
		else
		{
			if (dynamic_cast<const Triangle*>(h)) nodes[me].kind = WK_TRI;
			else if (dynamic_cast<const Sphere*>(h)) nodes[me].kind = WK_SPHERE;
			else if (dynamic_cast<const Cube*>(h)) nodes[me].kind = WK_CUBE;
			nodes[me].rank = numLeaves++;
		}
		return me;
	}
};

struct WalkCounts { uint64_t box, tri, sphere, other, rays; };

struct WalkHit { bool hit; float t; int32_t rank; };

static WalkHit Walk(const WalkTree& tree, int32_t ix, const ray& r, float tMin, float tMax, WalkCounts& c)
{
	const WalkNode& wn = tree.nodes[ix];
	switch (wn.kind)
	{
	case WK_NODE: {
		const BVHNode* n = static_cast<const BVHNode*>(wn.obj);
		c.box++;
		if (!n->box.Hit(r, tMin, tMax)) return WalkHit{ false, 0.0f, -1 };
		WalkHit a = Walk(tree, wn.left, r, tMin, tMax, c);
		WalkHit b = (wn.right < 0) ? WalkHit{ false, 0.0f, -1 } : Walk(tree, wn.right, r, tMin, tMax, c);
		if (a.hit && b.hit) return (a.t < b.t) ? a : b;
		if (a.hit) return a;
		return b;
	}
	case WK_MESH: {
		const StaticMesh* m = static_cast<const StaticMesh*>(wn.obj);
		c.box++;
		if (!m->bounds.Hit(r, tMin, tMax)) return WalkHit{ false, 0.0f, -1 };
		return Walk(tree, wn.left, r, tMin, tMax, c);
	}
	case WK_LIST: {
		// HitableList::Hit (geom/hit.cc:34-50): every member is asked with the closest t so far as its upper bound
		WalkHit best{ false, 0.0f, -1 };
		float closest = tMax;
		for (int32_t member : wn.members)
		{
			WalkHit h = Walk(tree, member, r, tMin, closest, c);
			if (h.hit) { best = h; closest = h.t; }
		}
		return best;
	}
	default: {
		if (wn.kind == WK_TRI) c.tri++; else if (wn.kind == WK_SPHERE) c.sphere++; else c.other++;
		HitResult hr;
		bool h = wn.obj->Hit(r, tMin, tMax, hr);
		return WalkHit{ h, h ? hr.t : 0.0f, h ? wn.rank : -1 };
	}
	}
}

static std::mutex g_walkMutex;
static std::unordered_map<const Scene*, WalkTree*> g_walkTrees;

static const WalkTree& GetWalkTree(const Scene* scene)
{
	std::lock_guard<std::mutex> lock(g_walkMutex);
	auto it = g_walkTrees.find(scene);
	if (it != g_walkTrees.end()) return *it->second;
	WalkTree* t = new WalkTree;
	t->Build(scene->accelStruct, 1);
	g_walkTrees[scene] = t;
	return *t;
}

extern "C" {

struct OraclePrimaryStats
{
	uint64_t boxTests, triTests, sphereTests, otherTests, rays;
	uint64_t walkVsHitMismatches;   // walk result differs from root->Hit (must be 0)
	int32_t  numLeaves, maxDepth, numNodes, pad;
	double   seconds;
};

// One unjittered camera ray per pixel (sample 0 of the frame's stream), as in
// renderer.cc:256-260; labels each with (leaf rank | -1, t | 0).
// `rayDump` (nullable): 8 floats per pixel = o.xyz, time, d.xyz, 0.
void oracle_primary_hits(
	const RendererSettings* settings, SceneHandle sceneH, CameraHandle cameraH,
	uint64_t frameSeed, int32_t nthreads,
	int32_t* outRank, float* outT, float* rayDump, OraclePrimaryStats* outStats)
{
	const Scene* scene = (const Scene*)sceneH;
	const Camera* camera = (const Camera*)cameraH;
	const WalkTree& tree = GetWalkTree(scene);
	const int32 W = (int32)settings->viewportWidth, H = (int32)settings->viewportHeight;
	const float imageWidth = (float)W, imageHeight = (float)H;
	const float tMin = settings->rayTMin;

	std::mutex m;
	WalkCounts total{ 0, 0, 0, 0, 0 };
	uint64_t mismatches = 0;
	auto tStart = std::chrono::steady_clock::now();
	ParallelRows(0, H, nthreads, [&](int32 y, int) {
		WalkCounts c{ 0, 0, 0, 0, 0 };
		uint64_t mm = 0;
		for (int32 x = 0; x < W; ++x)
		{
			const uint32_t pixel = (uint32_t)(y * W + x);
			g_oracleRng.key = rt_sample_key(frameSeed, pixel, 0u);
			g_oracleRng.ctr = 0;
			ray cameraRay = camera->GetCameraRay((float)x / imageWidth, (float)y / imageHeight);
			WalkHit wh = Walk(tree, 0, cameraRay, tMin, FLOAT_MAX, c);
			c.rays++;
			HitResult hr;
			bool h = scene->accelStruct->Hit(cameraRay, tMin, FLOAT_MAX, hr);
			if (h != wh.hit || (h && hr.t != wh.t)) mm++;
			outRank[pixel] = wh.hit ? wh.rank : -1;
			outT[pixel] = wh.hit ? wh.t : 0.0f;
			if (rayDump)
			{
				float* q = rayDump + 8 * (size_t)pixel;
				q[0] = cameraRay.o.x; q[1] = cameraRay.o.y; q[2] = cameraRay.o.z; q[3] = cameraRay.t;
				q[4] = cameraRay.d.x; q[5] = cameraRay.d.y; q[6] = cameraRay.d.z; q[7] = 0.0f;
			}
		}
		std::lock_guard<std::mutex> lock(m);
		total.box += c.box; total.tri += c.tri; total.sphere += c.sphere; total.other += c.other; total.rays += c.rays;
		mismatches += mm;
	});
	if (outStats)
	{
		outStats->boxTests = total.box; outStats->triTests = total.tri; outStats->sphereTests = total.sphere;
		outStats->otherTests = total.other; outStats->rays = total.rays;
		outStats->walkVsHitMismatches = mismatches;
		outStats->numLeaves = tree.numLeaves; outStats->maxDepth = tree.maxDepth; outStats->numNodes = (int32_t)tree.nodes.size();
		outStats->pad = 0;
		outStats->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - tStart).count();
	}
}

// Closest hit of arbitrary rays (8 floats each: o.xyz, time, d.xyz, unused) through the reference tree.
void oracle_trace_rays(SceneHandle sceneH, const float* rays, int64_t numRays, float tMin, int32_t nthreads,
	int32_t* outRank, float* outT, OraclePrimaryStats* outStats)
{
	const Scene* scene = (const Scene*)sceneH;
	const WalkTree& tree = GetWalkTree(scene);
	std::mutex m;
	WalkCounts total{ 0, 0, 0, 0, 0 };
	const int64_t chunk = 4096;
	const int32 numChunks = (int32)((numRays + chunk - 1) / chunk);
	ParallelRows(0, numChunks, nthreads, [&](int32 ci, int) {
		WalkCounts c{ 0, 0, 0, 0, 0 };
		for (int64_t i = ci * chunk; i < std::min(numRays, (ci + 1) * chunk); ++i)
		{
			const float* q = rays + 8 * i;
			ray r(vec3(q[0], q[1], q[2]), vec3(q[4], q[5], q[6]), q[3]);
			WalkHit wh = Walk(tree, 0, r, tMin, FLOAT_MAX, c);
			c.rays++;
			outRank[i] = wh.hit ? wh.rank : -1;
			outT[i] = wh.hit ? wh.t : 0.0f;
		}
		std::lock_guard<std::mutex> lock(m);
		total.box += c.box; total.tri += c.tri; total.sphere += c.sphere; total.other += c.other; total.rays += c.rays;
	});
	if (outStats)
	{
		*outStats = OraclePrimaryStats{};
		outStats->boxTests = total.box; outStats->triTests = total.tri; outStats->sphereTests = total.sphere;
		outStats->otherTests = total.other; outStats->rays = total.rays;
		outStats->numLeaves = tree.numLeaves; outStats->maxDepth = tree.maxDepth; outStats->numNodes = (int32_t)tree.nodes.size();
	}
}

void oracle_forget_scene(SceneHandle sceneH)
{
	std::lock_guard<std::mutex> lock(g_walkMutex);
	auto it = g_walkTrees.find((const Scene*)sceneH);
	if (it != g_walkTrees.end()) { delete it->second; g_walkTrees.erase(it); }
}

void scene_hook_on_destroy(SceneHandle sceneH) { oracle_forget_scene(sceneH); }

// Raw RGBA access to a reference Image2D, so that a test can hand the SAME HDR frame to the reference's own
// Raylib_PostProcess (render/image.cc:44-103) and to the product's GPU post-process, and compare all four channels.
void oracle_image_set_rgba(ImageHandle imageH, uint32_t width, uint32_t height, const float* rgba)
{
	Image2D* image = (Image2D*)imageH;
	if (image->GetWidth() != width || image->GetHeight() != height) image->Reallocate(width, height);
	for (uint32_t y = 0; y < height; ++y)
		for (uint32_t x = 0; x < width; ++x)
		{
			const float* p = rgba + 4 * ((size_t)y * width + x);
			image->SetPixel((int32)x, (int32)y, Pixel(p[0], p[1], p[2], p[3]));
		}
}

void oracle_image_get_rgba(ImageHandle imageH, float* rgba)
{
	const Image2D* image = (const Image2D*)imageH;
	const uint32_t width = image->GetWidth(), height = image->GetHeight();
	for (uint32_t y = 0; y < height; ++y)
		for (uint32_t x = 0; x < width; ++x)
		{
			const Pixel px = image->GetPixel((int32)x, (int32)y);
			float* p = rgba + 4 * ((size_t)y * width + x);
			p[0] = px.r; p[1] = px.g; p[2] = px.b; p[3] = px.a;
		}
}

} // extern "C"
