// trav_sim.cc -- development tool (host only, no GPU): how many nodes / boxes / triangles a near-first, pruning
// traversal touches per ray when the binary SAH tree of a scene is collapsed to 2-, 4- or 8-wide nodes, with the
// children visited in entry-distance order (what k_extend does) or in a fixed per-octant order (what an 8-wide
// node without a sorting network would do).  Used to decide what to build next; it is not part of the product and
// does not use the oracle.  Boxes are the exact ones (no quantization), triangles are tested with a plain
// Moller-Trumbore (statistics only).
//
//   tools/build_trav_sim.sh && tools/trav_sim <config id> [size]
#include "raylib.h"
#include "raylib_b200.h"
#include "rt_scene_format.h"
#include "bvh_sah.h"          // RtSahResult (csrc/host): the record array tools/bvh_reinsert.cc rewrites

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <cstring>
#include <thread>
#include <vector>

struct DemoSceneInfo
{
	SceneHandle  scene;
	CameraHandle camera;
	RendererSettings settings;
	uint64_t numTriangles, numSpheres, numMeshes;
	float cameraPos[3], cameraLookAt[3];
	float fovY, aperture, focalDistance, shutterBegin, shutterEnd;
};
extern "C" int32_t demo_scene_create(int32_t config, int32_t sizeParam, DemoSceneInfo* out);
extern "C" void demo_scene_destroy(SceneHandle scene, CameraHandle camera);

void RtReinsertSahTree(RtSahResult& tree, int iterations, double batchFraction, unsigned threads);      // tools/bvh_reinsert.cc

struct V3 { float x, y, z; };
static V3 operator+(V3 a, V3 b) { return { a.x + b.x, a.y + b.y, a.z + b.z }; }
static V3 operator-(V3 a, V3 b) { return { a.x - b.x, a.y - b.y, a.z - b.z }; }
static V3 operator*(float s, V3 a) { return { s * a.x, s * a.y, s * a.z }; }
static float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static V3 cross(V3 a, V3 b) { return { a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x }; }
static V3 normalize(V3 a) { const float k = 1.0f / std::sqrt(dot(a, a)); return k * a; }

struct Ray { V3 o, d, inv; };
static Ray makeRay(V3 o, V3 d) { return { o, d, { 1.0f / d.x, 1.0f / d.y, 1.0f / d.z } }; }

struct WideNode { int n; float lo[8][3], hi[8][3]; uint32_t ref[8]; };

struct Tree
{
	int width;
	std::vector<WideNode> nodes;
	uint32_t root;
};

static double area(const float* lo, const float* hi)
{
	const double dx = std::min((double)hi[0] - lo[0], 1e18), dy = std::min((double)hi[1] - lo[1], 1e18), dz = std::min((double)hi[2] - lo[2], 1e18);
	return (dx < 0 || dy < 0 || dz < 0) ? 0.0 : dx * dy + dy * dz + dz * dx;
}

// greedy collapse: expand the inner child with the largest surface area until `width` slots are used
static uint32_t collapse(const RtNode* bin, uint32_t binIndex, int width, std::vector<WideNode>& out)
{
	WideNode w; w.n = 2;
	const RtNode& b = bin[binIndex];
	memcpy(w.lo[0], b.lmin, 12); memcpy(w.hi[0], b.lmax, 12); w.ref[0] = b.lref;
	memcpy(w.lo[1], b.rmin, 12); memcpy(w.hi[1], b.rmax, 12); w.ref[1] = b.rref;
	while (w.n < width)
	{
		int best = -1; double bestA = -1.0;
		for (int i = 0; i < w.n; ++i)
			if (RT_REF_KIND(w.ref[i]) == RT_REF_NODE) { const double a = area(w.lo[i], w.hi[i]); if (a > bestA) { bestA = a; best = i; } }
		if (best < 0) break;
		const RtNode& c = bin[RT_REF_INDEX(w.ref[best])];
		memcpy(w.lo[best], c.lmin, 12); memcpy(w.hi[best], c.lmax, 12); w.ref[best] = c.lref;
		memcpy(w.lo[w.n], c.rmin, 12); memcpy(w.hi[w.n], c.rmax, 12); w.ref[w.n] = c.rref; w.n++;
	}
	const uint32_t index = (uint32_t)out.size();
	out.push_back(w);
	for (int i = 0; i < w.n; ++i)
		if (RT_REF_KIND(w.ref[i]) == RT_REF_NODE)
		{
			const uint32_t child = collapse(bin, RT_REF_INDEX(w.ref[i]), width, out);
			out[index].ref[i] = RT_MAKE_REF(RT_REF_NODE, child);
		}
	return index;
}

static bool slab(const float* lo, const float* hi, const Ray& r, float tMin, float tMax, float& entry)
{
	float t0 = tMin, t1 = tMax;
	const float o[3] = { r.o.x, r.o.y, r.o.z }, inv[3] = { r.inv.x, r.inv.y, r.inv.z };
	for (int a = 0; a < 3; ++a)
	{
		float ta = (lo[a] - o[a]) * inv[a], tb = (hi[a] - o[a]) * inv[a];
		if (inv[a] < 0.0f) std::swap(ta, tb);
		if (ta > t0) t0 = ta;
		if (tb < t1) t1 = tb;
	}
	entry = t0;
	return !(t1 < t0);
}

static bool triangle(const RtTriHot& t, const Ray& r, float tMin, float tMax, float& outT)
{
	const V3 v0 = { t.q[0], t.q[1], t.q[2] }, e1 = { t.q[6], t.q[7], t.q[8] }, e2 = { t.q[9], t.q[10], t.q[11] };
	const V3 p = cross(r.d, e2);
	const float det = dot(e1, p);
	if (std::fabs(det) < 1e-20f) return false;
	const float inv = 1.0f / det;
	const V3 s = r.o - v0;
	const float u = dot(s, p) * inv;
	if (u < 0.0f || u > 1.0f) return false;
	const V3 q = cross(s, e1);
	const float v = dot(r.d, q) * inv;
	if (v < 0.0f || u + v > 1.0f) return false;
	const float tt = dot(e2, q) * inv;
	if (tt < tMin || tt > tMax) return false;
	outT = tt;
	return true;
}

struct Counts { double nodes = 0, boxes = 0, tris = 0, pushes = 0, rays = 0; };

enum Order { ORDER_SORTED = 0, ORDER_OCTANT = 1, ORDER_AXIS = 2, ORDER_XOR = 3, ORDER_COUNT = 4 };
// popCull: 0 = every stacked child is visited when popped (inner children travel as one group entry without entry
// distances, as in compressed wide BVHs; leaves keep their entry distance), 1 = entries beyond the best hit are dropped
static int g_popCull = 1;   // 2 = nothing is dropped at pop, leaves included

// closest hit; returns t (FLT_MAX = miss) and the hit triangle index
static float trace(const Tree& tree, const RtSceneDesc* S, const Ray& r, float tMin, Order order, Counts& c, uint32_t& hitTri)
{
	struct Entry { uint32_t ref; float t; };
	Entry stack[256]; int sp = 0;
	float best = FLT_MAX; hitTri = 0xFFFFFFFFu;
	uint32_t cur = tree.root;
	c.rays += 1;
	for (;;)
	{
		if (RT_REF_KIND(cur) == RT_REF_NODE)
		{
			const WideNode& w = tree.nodes[RT_REF_INDEX(cur)];
			c.nodes += 1; c.boxes += w.n;
			Entry hits[8]; int nh = 0;
			for (int i = 0; i < w.n; ++i)
			{
				float e;
				if (slab(w.lo[i], w.hi[i], r, tMin, best, e)) hits[nh++] = { w.ref[i], e };
			}
			if (order == ORDER_SORTED) std::sort(hits, hits + nh, [](const Entry& a, const Entry& b) { return a.t < b.t; });
			else if (order == ORDER_AXIS)
			{
				// children ordered along ONE axis per node (the axis over which their centres spread most), reversed for
				// rays going the other way: a node needs 2 bits and the ray one sign test
				int ax = 0; float spread = -1.0f;
				for (int a = 0; a < 3; ++a)
				{
					float mn = FLT_MAX, mx = -FLT_MAX;
					for (int i = 0; i < w.n; ++i) { const float cc = std::min(std::max(w.lo[i][a] + w.hi[i][a], -1e18f), 1e18f); mn = std::min(mn, cc); mx = std::max(mx, cc); }
					if (mx - mn > spread) { spread = mx - mn; ax = a; }
				}
				const float sg = (ax == 0 ? r.d.x : ax == 1 ? r.d.y : r.d.z) < 0 ? -1.0f : 1.0f;
				auto key = [&](uint32_t ref) {
					for (int i = 0; i < w.n; ++i) if (w.ref[i] == ref) return sg * (w.lo[i][ax] + w.hi[i][ax]);
					return 0.0f; };
				std::sort(hits, hits + nh, [&](const Entry& a, const Entry& b) { return key(a.ref) < key(b.ref); });
			}
			else if (order == ORDER_XOR)
			{
				// slots hold the children sorted along the (+,+,+) diagonal; a ray visits slot (position ^ x), x in 0..3 chosen per node and
				// octant as the one with the fewest inversions against that octant's diagonal order (2 bits per octant in the node)
				const float sg[3] = { r.d.x < 0 ? -1.0f : 1.0f, r.d.y < 0 ? -1.0f : 1.0f, r.d.z < 0 ? -1.0f : 1.0f };
				int slotOf[8]; float k0[8], ko[8];
				for (int i = 0; i < w.n; ++i)
				{
					slotOf[i] = i;
					float c3[3]; for (int a = 0; a < 3; ++a) c3[a] = std::min(std::max(w.lo[i][a] + w.hi[i][a], -1e18f), 1e18f);
					k0[i] = c3[0] + c3[1] + c3[2]; ko[i] = sg[0] * c3[0] + sg[1] * c3[1] + sg[2] * c3[2];
				}
				std::sort(slotOf, slotOf + w.n, [&](int a, int b) { return k0[a] < k0[b]; });     // slotOf[s] = child in slot s
				const int W2 = tree.width;
				int bestX = 0, bestInv = 1 << 30;
				for (int x = 0; x < W2; ++x)
				{
					int inv = 0;
					for (int p = 0; p < W2; ++p) for (int q = p + 1; q < W2; ++q)
					{
						const int sa = p ^ x, sb = q ^ x;
						if (sa >= w.n || sb >= w.n) continue;
						if (ko[slotOf[sa]] > ko[slotOf[sb]]) inv++;
					}
					if (inv < bestInv) { bestInv = inv; bestX = x; }
				}
				auto key = [&](uint32_t ref) {
					for (int sl = 0; sl < w.n; ++sl) if (w.ref[slotOf[sl]] == ref) return (float)(sl ^ bestX);
					return 0.0f; };
				std::sort(hits, hits + nh, [&](const Entry& a, const Entry& b) { return key(a.ref) < key(b.ref); });
			}
			else
			{
				// fixed order per ray octant: children by the position of their box centre along the octant's diagonal
				const float sx = r.d.x < 0 ? -1.0f : 1.0f, sy = r.d.y < 0 ? -1.0f : 1.0f, sz = r.d.z < 0 ? -1.0f : 1.0f;
				auto key = [&](uint32_t ref) {
					for (int i = 0; i < w.n; ++i) if (w.ref[i] == ref)
						return sx * (w.lo[i][0] + w.hi[i][0]) + sy * (w.lo[i][1] + w.hi[i][1]) + sz * (w.lo[i][2] + w.hi[i][2]);
					return 0.0f; };
				std::sort(hits, hits + nh, [&](const Entry& a, const Entry& b) { return key(a.ref) < key(b.ref); });
			}
			for (int i = nh - 1; i >= 1; --i) { stack[sp++] = hits[i]; c.pushes += 1; }
			if (nh) { cur = hits[0].ref; continue; }
		}
		else
		{
			const uint32_t kind = RT_REF_KIND(cur), first = RT_REF_INDEX(cur);
			const int count = kind == RT_REF_TRI2 ? 2 : 1;
			if (kind == RT_REF_TRI || kind == RT_REF_TRI2)
				for (int i = 0; i < count; ++i)
				{
					float t; c.tris += 1;
					if (triangle(S->triHot[first + i], r, tMin, best, t)) { best = t; hitTri = first + i; }
				}
		}
		// pop, dropping entries that start beyond the best hit
		for (;;)
		{
			if (sp == 0) return best;
			const Entry e = stack[--sp];
			if (e.t <= best || (g_popCull == 0 && RT_REF_KIND(e.ref) == RT_REF_NODE) || g_popCull == 2) { cur = e.ref; break; }
		}
	}
}

static uint32_t g_rng = 12345u;
static float rnd() { g_rng = g_rng * 1664525u + 1013904223u; return (float)(g_rng >> 8) / 16777216.0f; }

int main(int argc, char** argv)
{
	const int cfg = argc > 1 ? atoi(argv[1]) : 4, size = argc > 2 ? atoi(argv[2]) : 0;
	Raylib_Initialize();
	DemoSceneInfo info;
	if (!demo_scene_create(cfg, size, &info)) { fprintf(stderr, "cannot create demo scene %d\n", cfg); return 1; }
	const RtSceneDesc* S = RaylibB200_FlattenForInspection(info.scene);
	if (!S || RT_REF_KIND(S->rootRef) != RT_REF_NODE) { fprintf(stderr, "flatten failed: %s\n", RaylibB200_GetLastError()); return 1; }
	RtCamera cam;
	RaylibB200_CameraBlock(info.camera, &cam);
	const float tMin = info.settings.rayTMin;

	// rays: a 192 x 108 grid of camera rays, then two generations of uniform-hemisphere bounces from their hits
	std::vector<Ray> primary;
	const int W = getenv("TRAV_SIM_WARP") ? 480 : 192, H = getenv("TRAV_SIM_WARP") ? 270 : 108;      // more rays for the warp model
	for (int y = 0; y < H; ++y)
		for (int x = 0; x < W; ++x)
		{
			const float u = (x + 0.5f) / W, v = (y + 0.5f) / H;
			const V3 o = { cam.origin[0], cam.origin[1], cam.origin[2] };
			const V3 p = { cam.topLeft[0] + u * cam.horizontal[0] + (1.0f - v) * cam.vertical[0],
			               cam.topLeft[1] + u * cam.horizontal[1] + (1.0f - v) * cam.vertical[1],
			               cam.topLeft[2] + u * cam.horizontal[2] + (1.0f - v) * cam.vertical[2] };
			primary.push_back(makeRay(o, normalize(p - o)));
		}

	// TRAV_SIM_REINSERT=<iterations> [TRAV_SIM_REINSERT_FRACTION=<share of the nodes per iteration>]: the statistics below are
	// taken on a copy of the binary tree after insertion-based optimisation (tools/bvh_reinsert.cc)
	RtSahResult optimised;
	const RtNode* binaryNodes = S->nodes;
	uint32_t binaryRoot = S->rootRef;
	if (const char* it = getenv("TRAV_SIM_REINSERT"))
	{
		optimised.nodes.resize(S->numNodes);
		memcpy(optimised.nodes.data(), S->nodes, sizeof(RtNode) * (size_t)S->numNodes);
		memcpy(optimised.rootMin, S->rootMin, 12); memcpy(optimised.rootMax, S->rootMax, 12);
		optimised.rootRef = S->rootRef; optimised.maxDepth = 0; optimised.cost = 0.0;
		const char* fr = getenv("TRAV_SIM_REINSERT_FRACTION");
		const auto t0 = std::chrono::steady_clock::now();
		RtReinsertSahTree(optimised, atoi(it), fr ? atof(fr) : 1.0, std::max(1u, std::thread::hardware_concurrency()));
		printf("re-insertion: %d iterations in %.1f s, cost %.2f, depth %u\n", atoi(it),
		       std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(), optimised.cost, optimised.maxDepth);
		binaryNodes = optimised.nodes.data(); binaryRoot = optimised.rootRef;
	}

	Tree trees[3];
	const int widths[3] = { 2, 4, 8 };
	for (int i = 0; i < 3; ++i)
	{
		trees[i].width = widths[i];
		trees[i].nodes.reserve(S->numNodes);
		trees[i].root = RT_MAKE_REF(RT_REF_NODE, collapse(binaryNodes, RT_REF_INDEX(binaryRoot), widths[i], trees[i].nodes));
	}

	// generate the bounce rays once, with the 4-wide tree
	std::vector<Ray> gen[3];
	gen[0] = primary;
	for (int g = 0; g < 2; ++g)
	{
		Counts dummy;
		for (const Ray& r : gen[g])
		{
			uint32_t tri;
			const float t = trace(trees[1], S, r, tMin, ORDER_SORTED, dummy, tri);
			if (t == FLT_MAX) continue;
			const RtTriHot& T = S->triHot[tri];
			V3 n = { T.q[3], T.q[4], T.q[5] };
			if (dot(n, r.d) > 0.0f) n = -1.0f * n;
			V3 d;
			do { d = { 2.0f * rnd() - 1.0f, 2.0f * rnd() - 1.0f, 2.0f * rnd() - 1.0f }; } while (dot(d, d) > 1.0f || dot(d, d) < 1e-4f);
			d = normalize(d);
			if (dot(d, n) < 0.0f) d = -1.0f * d;
			gen[g + 1].push_back(makeRay(r.o + t * r.d + 1e-3f * n, d));
		}
	}

	// SAH cost of the binary tree: sum of the surface areas of all inner nodes over the root's (expected nodes visited by a
	// random line through the root box, no occlusion)
	double sah = 0.0;
	{
		const double rootA = area(S->rootMin, S->rootMax);
		for (uint32_t i = 0; i < S->numNodes; ++i)
		{
			const RtNode& n = binaryNodes[i];
			float lo[3], hi[3];
			for (int a = 0; a < 3; ++a) { lo[a] = std::min(n.lmin[a], n.rmin[a]); hi[a] = std::max(n.lmax[a], n.rmax[a]); }
			sah += area(lo, hi) / rootA;
		}
	}
	printf("config %d: %u triangles, binary SAH nodes %u, SAH cost (sum of node areas / root area) %.2f\n", cfg, S->numTris, S->numNodes, sah);
	printf("%-8s %-8s %-9s | %10s %10s %10s %10s | wide nodes\n", "rays", "width", "order", "nodes/ray", "boxes/ray", "tris/ray", "pushes/ray");
	const char* genName[3] = { "camera", "bounce1", "bounce2" };
	for (int g = 0; g < 3; ++g)
		for (int i = 0; i < 3; ++i)
			for (int o = 0; o < ORDER_COUNT; ++o)
				for (int cull = 2; cull >= 0; --cull)
				{
					if ((widths[i] == 2 && o != 0) || (g < 1)) continue;
					g_popCull = cull;
					Counts c;
					for (const Ray& r : gen[g]) { uint32_t tri; trace(trees[i], S, r, tMin, (Order)o, c, tri); }
					static const char* orderName[ORDER_COUNT] = { "sorted", "octant", "axis", "xor" };
					printf("%-8s %-8d %-9s %-7s | %10.2f %10.2f %10.2f %10.2f | %zu\n", genName[g], widths[i], orderName[o], cull == 1 ? "cull" : cull == 2 ? "nocull+" : "nocull",
					       c.nodes / c.rays, c.boxes / c.rays, c.tris / c.rays, c.pushes / c.rays, trees[i].nodes.size());
				}
	// The product's own 4-wide tree: exact child boxes (wideNodes) against the decoded 7-bit boxes with the device's
	// ray-space arithmetic and slack (rt_traverse.cuh trav_step) -- what the quantization costs in node visits.
	// TRAV_SIM_FREE_SCALE=1 re-quantizes every node with a scale that is not restricted to powers of two.
	if (S->wideNodes && S->quantNodes && RT_REF_KIND(S->wideRootRef) == RT_REF_NODE)
	{
		struct QBox { float lo[4][3], hi[4][3]; };
		const bool freeScale = getenv("TRAV_SIM_FREE_SCALE") != nullptr;
		// TRAV_SIM_FREE_SCALE=2: all 256 byte values, m = as_float(0x3F000000 | byte << 16) in [0.5, 2): 128 steps of S/256 below
		// m = 1 and 128 steps of S/128 above (the device decode as it is), S free
		const bool wideGrid = freeScale && atoi(getenv("TRAV_SIM_FREE_SCALE")) == 2;
		std::vector<QBox> decoded(S->numWideNodes);
		for (uint32_t i = 0; i < S->numWideNodes; ++i)
		{
			const RtNodeQ4& q = S->quantNodes[i];
			const RtNode4& w = S->wideNodes[i];
			const float scale[3] = { q.scaleX, q.scaleY, q.scaleZ };
			const float* wlo[3] = { w.lox, w.loy, w.loz }; const float* whi[3] = { w.hix, w.hiy, w.hiz };
			for (int a = 0; a < 3; ++a)
			{
				float mn = FLT_MAX, mx = -FLT_MAX;
				for (int k = 0; k < 4; ++k) if (w.ref[k] != RT_REF_ABSENT) { mn = std::min(mn, std::max(wlo[a][k], -1e18f)); mx = std::max(mx, std::min(whi[a][k], 1e18f)); }
				for (int k = 0; k < 4; ++k)
				{
					if (!freeScale)
					{
						decoded[i].lo[k][a] = rt_q4_plane((q.qlo[a] >> (8 * k)) & 0xFFu, scale[a], q.base[a]);
						decoded[i].hi[k][a] = rt_q4_plane((q.qhi[a] >> (8 * k)) & 0xFFu, scale[a], q.base[a]);
					}
					else
					{
						if (wideGrid)
						{
							const double S = std::max((double)mx - mn, 1e-30) * 1.0001 / 1.4921875, base = mn - 0.5 * S;      // m in [0.5, 255/128]
							auto plane = [&](int b) { return base + (b < 128 ? 0.5 + b / 256.0 : 1.0 + (b - 128) / 128.0) * S; };
							int bl = 0, bh = 255;
							const double l = std::max(wlo[a][k], -1e18f), h = std::min(whi[a][k], 1e18f);
							while (bl < 255 && plane(bl + 1) <= l) ++bl;
							while (bh > 0 && plane(bh - 1) >= h) --bh;
							decoded[i].lo[k][a] = (float)plane(bl); decoded[i].hi[k][a] = (float)plane(bh);
							if (decoded[i].lo[k][a] > l) decoded[i].lo[k][a] = std::nextafter(decoded[i].lo[k][a], -FLT_MAX);
							if (decoded[i].hi[k][a] < h) decoded[i].hi[k][a] = std::nextafter(decoded[i].hi[k][a], FLT_MAX);
							continue;
						}
						// 127 steps over exactly [mn, mx] (+ a hair), planes rounded outwards
						const double step = std::max((double)mx - mn, 1e-30) * 1.0001 / 127.0;
						decoded[i].lo[k][a] = (float)(mn + std::floor((std::max(wlo[a][k], -1e18f) - (double)mn) / step) * step);
						decoded[i].hi[k][a] = (float)(mn + std::ceil((std::min(whi[a][k], 1e18f) - (double)mn) / step) * step);
					}
				}
			}
		}
		for (int mode = 0; mode < 2; ++mode)
			for (int g = 1; g < 3; ++g)
			{
				double nodes = 0, tris = 0;
				for (const Ray& r : gen[g])
				{
					struct Entry { uint32_t ref; float t; };
					Entry stack[256]; int sp = 0;
					float best = FLT_MAX;
					uint32_t cur = S->wideRootRef;
					for (;;)
					{
						if (RT_REF_KIND(cur) == RT_REF_NODE)
						{
							const uint32_t ni = RT_REF_INDEX(cur);
							const RtNode4& w = S->wideNodes[ni];
							nodes += 1;
							Entry hits[4]; int nh = 0;
							for (int k = 0; k < 4; ++k)
							{
								if (w.ref[k] == RT_REF_ABSENT) continue;
								float lo[3], hi[3], e;
								if (mode == 0) { lo[0] = w.lox[k]; lo[1] = w.loy[k]; lo[2] = w.loz[k]; hi[0] = w.hix[k]; hi[1] = w.hiy[k]; hi[2] = w.hiz[k]; }
								else { memcpy(lo, decoded[ni].lo[k], 12); memcpy(hi, decoded[ni].hi[k], 12); }
								if (slab(lo, hi, r, tMin, best, e)) hits[nh++] = { w.ref[k], e };
							}
							std::sort(hits, hits + nh, [](const Entry& a, const Entry& b) { return a.t < b.t; });
							for (int i = nh - 1; i >= 1; --i) stack[sp++] = hits[i];
							if (nh) { cur = hits[0].ref; continue; }
						}
						else
						{
							const uint32_t kind = RT_REF_KIND(cur), first = RT_REF_INDEX(cur);
							const int count = kind == RT_REF_TRI2 ? 2 : 1;
							if (kind == RT_REF_TRI || kind == RT_REF_TRI2)
								for (int i = 0; i < count; ++i) { float t; tris += 1; if (triangle(S->triHot[first + i], r, tMin, best, t)) best = t; }
						}
						bool done = false;
						for (;;)
						{
							if (sp == 0) { done = true; break; }
							const Entry e = stack[--sp];
							if (e.t <= best) { cur = e.ref; break; }
						}
						if (done) break;
					}
				}
				printf("product tree %-8s %-22s | nodes/ray %7.2f  tris/ray %6.2f\n", genName[g], mode == 0 ? "exact child boxes" : (wideGrid ? "8-bit [0.5,2), free (sim)" : freeScale ? "7-bit, free scale (sim)" : "product quantNodes"), nodes / gen[g].size(), tris / gen[g].size());
			}

		// ---- TRAV_SIM_WARP=1: a model of k_extend's warp schedule (rt_device.cu extend_stage + rt_traverse.cuh trav_run) -----
		// 32 lanes, one ray each for the ray's whole walk, idle lanes refilled from the (binned) queue when fewer than
		// `refill` lanes are busy; node phase = all lanes that can step do so together, ended when fewer than `walk` can and a
		// lane is blocked on leaves; leaf phase = every lane with a postponed leaf tests one.  A lane can postpone
		// `leafQueue` leaves before it blocks (the kernels: 1).  Counts warp iterations of both phases and the lanes in them:
		// what a scheduling change would buy in issued instructions, before anyone writes the kernel.
		if (getenv("TRAV_SIM_WARP"))
		{
			// bounced rays in the order the binning pass would hand them out: by origin cell (4 bits per axis), then direction octant
			std::vector<uint32_t> order(gen[2].size());
			std::vector<uint32_t> key(gen[2].size());
			for (size_t i = 0; i < order.size(); ++i)
			{
				order[i] = (uint32_t)i;
				const Ray& r = gen[2][i];
				const float o[3] = { r.o.x, r.o.y, r.o.z }, dd[3] = { r.d.x, r.d.y, r.d.z };
				uint32_t cell[3], k = 0;
				for (int a = 0; a < 3; ++a)
				{
					const float lo = std::max(S->rootMin[a], -1e3f), hi = std::min(S->rootMax[a], 1e3f);
					cell[a] = (uint32_t)std::min(15.0f, std::max(0.0f, (o[a] - lo) / std::max(hi - lo, 1e-20f) * 16.0f));
				}
				for (int b = 3; b >= 0; --b) for (int a = 0; a < 3; ++a) k = (k << 1) | ((cell[a] >> b) & 1u);       // Morton order
				for (int a = 0; a < 3; ++a) k = (k << 1) | (dd[a] < 0.0f ? 1u : 0u);
				key[i] = k;
			}
			std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return key[a] < key[b]; });

			struct Lane
			{
				int state = 0;                 // 0 empty, 1 active
				uint32_t ray = 0, cur = 0, leaf[4], nLeaf = 0;
				float best = FLT_MAX;
				struct E { uint32_t ref; float t; } stack[128]; int sp = 0;
			};
			const uint32_t DONE = 0xFFFFFFFEu, POP = 0xFFFFFFFDu;
			auto isLeaf = [&](uint32_t ref) { return ref != DONE && ref != POP && RT_REF_KIND(ref) != RT_REF_NODE && RT_REF_KIND(ref) != RT_REF_NONE; };
			printf("%-38s | node iters/ray  lanes | leaf iters/ray  lanes | nodes/ray tris/ray | issue estimate (209 / 150 instructions per iteration)\n", "warp model (bounce2, binned order)");
			const int variants[][4] = { { 20, 20, 1, 0 }, { 20, 16, 1, 0 }, { 20, 24, 1, 0 }, { 24, 20, 1, 0 }, { 20, 20, 2, 0 }, { 20, 20, 3, 0 }, { 20, 20, 2, 1 }, { 20, 20, 3, 1 }, { 20, 28, 2, 0 }, { 20, 28, 3, 1 } };
			for (const auto& v : variants)
			{
				const uint32_t refill = (uint32_t)v[0], walk = (uint32_t)v[1], leafQueue = (uint32_t)v[2]; const bool drain = v[3] != 0;
				double nodeIters = 0, nodeLanes = 0, leafIters = 0, leafLanes = 0, nodeVisits = 0, triTests = 0;
				size_t cursor = 0;
				// a grid of persistent warps shares the queue; warps are independent in this model, so they run one after another
				// on chunks the size a warp really sees between refills does not matter: one warp eats the whole queue
				std::vector<Lane> lanes(32);
				bool exhausted = false;
				auto canStep = [&](const Lane& L) { return L.cur != DONE && !(L.nLeaf >= leafQueue && isLeaf(L.cur)); };
				auto finished = [&](const Lane& L) { return L.cur == DONE && L.nLeaf == 0; };
				auto step = [&](Lane& L)
				{
					const Ray& r = gen[2][L.ray];
					uint32_t cur = L.cur;
					if (cur != POP && RT_REF_KIND(cur) == RT_REF_NODE)
					{
						const uint32_t ni = RT_REF_INDEX(cur);
						const RtNode4& w = S->wideNodes[ni];
						nodeVisits += 1;
						Lane::E hits[4]; int nh = 0;
						for (int k = 0; k < 4; ++k)
						{
							if (w.ref[k] == RT_REF_ABSENT) continue;
							float e;
							if (slab(decoded[ni].lo[k], decoded[ni].hi[k], r, tMin, L.best, e)) hits[nh++] = { w.ref[k], e };
						}
						std::sort(hits, hits + nh, [](const Lane::E& a, const Lane::E& b) { return a.t < b.t; });
						for (int i = nh - 1; i >= 1; --i) L.stack[L.sp++] = hits[i];
						cur = nh ? hits[0].ref : POP;
					}
					if (L.nLeaf < leafQueue && isLeaf(cur)) { L.leaf[L.nLeaf++] = cur; cur = POP; }
					for (int attempt = 0; attempt < 2 && cur == POP; ++attempt)
					{
						if (L.sp == 0) { cur = DONE; break; }
						const Lane::E e = L.stack[--L.sp];
						cur = e.t > L.best ? POP : e.ref;
					}
					L.cur = cur;
				};
				auto testLeaf = [&](Lane& L)
				{
					const Ray& r = gen[2][L.ray];
					const uint32_t leaf = L.leaf[0];
					for (uint32_t i = 1; i < L.nLeaf; ++i) L.leaf[i - 1] = L.leaf[i];
					L.nLeaf--;
					const uint32_t kind = RT_REF_KIND(leaf), first = RT_REF_INDEX(leaf);
					const int count = kind == RT_REF_TRI2 ? 2 : 1;
					if (kind == RT_REF_TRI || kind == RT_REF_TRI2)
						for (int i = 0; i < count; ++i) { float t; triTests += 1; if (triangle(S->triHot[first + i], r, tMin, L.best, t)) L.best = t; }
					if (L.nLeaf < leafQueue && isLeaf(L.cur)) { L.leaf[L.nLeaf++] = L.cur; L.cur = POP; }
				};
				for (;;)
				{
					for (Lane& L : lanes) if (L.state == 1 && finished(L)) L.state = 0;
					if (!exhausted)
						for (Lane& L : lanes)
							if (L.state == 0)
							{
								if (cursor >= order.size()) { exhausted = true; break; }
								L.ray = order[cursor++]; L.cur = S->wideRootRef; L.nLeaf = 0; L.sp = 0; L.best = FLT_MAX; L.state = 1;
							}
					uint32_t active = 0;
					for (const Lane& L : lanes) active += L.state == 1 && !finished(L) ? 1u : 0u;
					if (active == 0) { if (exhausted) break; continue; }
					const uint32_t keepGoing = exhausted ? 1u : refill;
					for (;;)
					{
						for (;;)
						{
							uint32_t nStep = 0, blocked = 0;
							for (const Lane& L : lanes) { const bool alive = L.state == 1 && !finished(L); if (alive && canStep(L)) nStep++; else if (alive) blocked++; }
							if (nStep == 0) break;
							if (nStep < walk && blocked) break;
							nodeIters += 1; nodeLanes += nStep;
							for (Lane& L : lanes) if (L.state == 1 && !finished(L) && canStep(L)) step(L);
						}
						do
						{
							uint32_t busy = 0;
							for (const Lane& L : lanes) busy += (L.state == 1 && L.nLeaf) ? 1u : 0u;
							leafIters += 1; leafLanes += busy;
							for (Lane& L : lanes) if (L.state == 1 && L.nLeaf) testLeaf(L);
							if (!drain) break;
							busy = 0;
							for (const Lane& L : lanes) busy += (L.state == 1 && L.nLeaf) ? 1u : 0u;
							if (busy == 0) break;
						} while (true);
						uint32_t alive = 0;
						for (const Lane& L : lanes) alive += (L.state == 1 && !finished(L)) ? 1u : 0u;
						if (alive < keepGoing) break;
					}
				}
				const double n = (double)order.size();
				char name[96];
				snprintf(name, sizeof(name), "refill %d walk %d leaf queue %d%s", v[0], v[1], v[2], drain ? " drained" : "");
				printf("%-38s | %14.4f %6.2f | %14.4f %6.2f | %9.2f %8.2f | %8.1f\n", name, nodeIters / n, nodeLanes / std::max(1.0, nodeIters),
				       leafIters / n, leafLanes / std::max(1.0, leafIters), nodeVisits / n, triTests / n, (nodeIters * 209.0 + leafIters * 150.0) / n);
			}
		}
	}
	demo_scene_destroy(info.scene, info.camera);
	Raylib_Terminate();
	return 0;
}
