// bvh_sah.cc -- binned surface-area-heuristic build (64 bins per axis, one or all three axes per split), top levels in
// parallel.  Node records are laid out in depth-first pre-order: a subtree over m groups owns a contiguous
// block of m-1 records, which makes the layout deterministic and independent of the thread schedule.
#include "bvh_sah.h"

#include <algorithm>
#include <atomic>
#include <cfloat>
#include <cstring>
#include <future>
#include <thread>
#include <limits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstdlib>

#ifndef RT_SAH_ROTATION_PASSES
#define RT_SAH_ROTATION_PASSES 1
#endif

namespace
{
	const int kBins = 64;      // per axis; 64 instead of 16 bins: scatter scene -6 % node visits per ray
	const int kCoarseBins = 16;             // ... for subtrees below kFineBinsFrom items
	const uint32_t kFineBinsFrom = 128;

	struct Box
	{
		float lo[3], hi[3];
		void Reset()
		{
			for (int a = 0; a < 3; ++a) { lo[a] = std::numeric_limits<float>::infinity(); hi[a] = -std::numeric_limits<float>::infinity(); }
		}
		void Grow(const float* l, const float* h)
		{
			for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], l[a]); hi[a] = std::max(hi[a], h[a]); }
		}
		void GrowPoint(const float* p) { Grow(p, p); }
		double HalfArea() const
		{
			const double dx = (double)hi[0] - lo[0], dy = (double)hi[1] - lo[1], dz = (double)hi[2] - lo[2];
			if (!(dx >= 0.0) || !(dy >= 0.0) || !(dz >= 0.0)) return 0.0;
			// boxes can be unbounded (a bare primitive under the scene root has no gate): keep the cost finite
			const double cx = std::min(dx, 1.0e18), cy = std::min(dy, 1.0e18), cz = std::min(dz, 1.0e18);
			return cx * cy + cy * cz + cz * cx;
		}
	};

	inline float Centroid(const RtLeafGroup& g, int a)
	{
		// finite even for huge boxes
		return 0.5f * std::max(-1.0e18f, g.lo[a]) + 0.5f * std::min(1.0e18f, g.hi[a]);
	}

	// Nodes with at least this many items run their passes (centroid bounds, binning, partition) on several threads: the
	// top levels of a 10 M-item build are otherwise as long as the rest of the tree (one thread touches every item at
	// level 0, two at level 1, ...).  Results do not depend on the thread count: min/max/+ merges and a STABLE partition.
	uint32_t ParallelNodeFrom()
	{
		static const uint32_t value = []() {
			const char* v = getenv("RAYLIB_B200_SAH_PARALLEL_FROM");       // tests lower it to cover the threaded passes on small scenes
			return v ? (uint32_t)std::max(2, atoi(v)) : (1u << 19); }();
		return value;
	}

	unsigned ChunkThreads(uint32_t count, unsigned threads)
	{
		const uint32_t grain = ParallelNodeFrom() < (1u << 19) ? 64u : 65536u;       // items per thread at least
		return std::max(1u, std::min<unsigned>(threads, count / grain + 1u));
	}
	template<typename Fn> void ParallelChunks(uint32_t count, unsigned threads, Fn fn)
	{
		threads = ChunkThreads(count, threads);
		std::vector<std::thread> pool;
		for (unsigned t = 1; t < threads; ++t)
			pool.emplace_back([=]() { fn(t, (uint32_t)((uint64_t)count * t / threads), (uint32_t)((uint64_t)count * (t + 1) / threads)); });
		fn(0u, 0u, (uint32_t)((uint64_t)count / threads));
		for (std::thread& th : pool) th.join();
	}

	struct Builder
	{
		RtLeafGroup* groups;
		RtNode* nodes;
		bool searchAllAxes = true;
		uint32_t maxDepth = 0;

		// Builds the subtree over groups[first, first+count) into nodes[nodeBase, nodeBase+count-1).
		// Returns the child descriptor (ref + exact union box) for the parent.
		struct Child { uint32_t ref; Box box; uint32_t depth; };

		Child Build(uint32_t first, uint32_t count, uint32_t nodeBase, int parallelDepth)
		{
			Child me;
			if (count == 1)
			{
				const RtLeafGroup& g = groups[first];
				me.ref = g.ref;
				memcpy(me.box.lo, g.lo, 12); memcpy(me.box.hi, g.hi, 12);
				me.depth = 0;
				return me;
			}

			const bool big = count >= ParallelNodeFrom();      // the chunked code path (stable partition) whatever the core count,
			                                                   // so that the tree does not depend on the machine
			const unsigned workers = big ? std::max(1u, std::thread::hardware_concurrency() >> (5 - std::min(5, parallelDepth))) : 1u;
			Box centroidBounds; centroidBounds.Reset();
			if (big)
			{
				std::vector<Box> part(workers);
				for (Box& b : part) b.Reset();
				ParallelChunks(count, workers, [&](unsigned t, uint32_t b, uint32_t e) {
					Box acc; acc.Reset();
					for (uint32_t i = first + b; i < first + e; ++i)
					{
						const float c[3] = { Centroid(groups[i], 0), Centroid(groups[i], 1), Centroid(groups[i], 2) };
						acc.GrowPoint(c);
					}
					part[t] = acc; });
				for (const Box& b : part) if (b.lo[0] <= b.hi[0]) centroidBounds.Grow(b.lo, b.hi);
			}
			else
			for (uint32_t i = first; i < first + count; ++i)
			{
				const float c[3] = { Centroid(groups[i], 0), Centroid(groups[i], 1), Centroid(groups[i], 2) };
				centroidBounds.GrowPoint(c);
			}
			// longest centroid axis: the fallback split and the only candidate when a single axis is searched
			int axis = 0;
			float extent = centroidBounds.hi[0] - centroidBounds.lo[0];
			for (int a = 1; a < 3; ++a)
			{
				const float e = centroidBounds.hi[a] - centroidBounds.lo[a];
				if (e > extent) { extent = e; axis = a; }
			}

			uint32_t mid = first + count / 2;
			bool split = false;
			int binsUsed = kBins;
			if (extent > 0.0f && count > 2)
			{
				// binned SAH over every axis with a non-degenerate centroid extent: the best (axis, bin boundary) wins.
				// One pass over the items fills the bins of all candidate axes; small subtrees (most of the nodes) use
				// 16 bins -- resetting 3 x 64 boxes per node would dominate the build.
				const int nb = count >= kFineBinsFrom ? kBins : (count >= 32 ? kCoarseBins : (count >= 8 ? 8 : 4));     // no more bins than a handful of items can fill: the sweeps over empty bins were half of the build
				binsUsed = nb;
				double bestCost = DBL_MAX; int bestSplit = -1, bestAxis = axis;
				const int numAxes = searchAllAxes ? 3 : 1;
				Box binBox[3][kBins]; uint32_t binCount[3][kBins];
				int ax[3]; float lo[3], scale[3]; bool use[3];
				for (int pass = 0; pass < numAxes; ++pass)
				{
					ax[pass] = (axis + pass) % 3;        // longest axis first: it keeps exact ties (regular grids)
					const float ext = centroidBounds.hi[ax[pass]] - centroidBounds.lo[ax[pass]];
					use[pass] = ext > 0.0f;
					lo[pass] = centroidBounds.lo[ax[pass]];
					scale[pass] = use[pass] ? (float)nb / ext : 0.0f;
					if (use[pass]) for (int b = 0; b < nb; ++b) { binBox[pass][b].Reset(); binCount[pass][b] = 0; }
				}
				auto binRange = [&](uint32_t b0, uint32_t e0, Box (*bb)[kBins], uint32_t (*bc)[kBins]) {
					for (uint32_t i = first + b0; i < first + e0; ++i)
					{
						const RtLeafGroup& g = groups[i];
						for (int pass = 0; pass < numAxes; ++pass)
						{
							if (!use[pass]) continue;
							const int b = std::min(nb - 1, std::max(0, (int)((Centroid(g, ax[pass]) - lo[pass]) * scale[pass])));
							bb[pass][b].Grow(g.lo, g.hi);
							bc[pass][b]++;
						}
					} };
				if (big)
				{
					struct Bins { Box box[3][kBins]; uint32_t count[3][kBins]; };
					std::vector<Bins> part(workers);
					ParallelChunks(count, workers, [&](unsigned t, uint32_t b, uint32_t e) {
						Bins& mine = part[t];
						for (int pass = 0; pass < 3; ++pass) for (int k = 0; k < kBins; ++k) { mine.box[pass][k].Reset(); mine.count[pass][k] = 0; }
						binRange(b, e, mine.box, mine.count); });
					for (const Bins& pb : part)
						for (int pass = 0; pass < numAxes; ++pass)
						{
							if (!use[pass]) continue;
							for (int k = 0; k < nb; ++k)
								if (pb.count[pass][k]) { binBox[pass][k].Grow(pb.box[pass][k].lo, pb.box[pass][k].hi); binCount[pass][k] += pb.count[pass][k]; }
						}
				}
				else binRange(0, count, binBox, binCount);
				for (int pass = 0; pass < numAxes; ++pass)
				{
					if (!use[pass]) continue;
					double rightArea[kBins]; uint32_t rightCount[kBins];
					Box acc; acc.Reset(); uint32_t n = 0;
					for (int b = nb - 1; b > 0; --b)
					{
						if (binCount[pass][b]) acc.Grow(binBox[pass][b].lo, binBox[pass][b].hi);
						n += binCount[pass][b];
						rightArea[b] = acc.HalfArea(); rightCount[b] = n;
					}
					acc.Reset(); n = 0;
					for (int b = 0; b < nb - 1; ++b)
					{
						if (binCount[pass][b]) acc.Grow(binBox[pass][b].lo, binBox[pass][b].hi);
						n += binCount[pass][b];
						if (n == 0 || rightCount[b + 1] == 0) continue;
						const double cost = acc.HalfArea() * (double)n + rightArea[b + 1] * (double)rightCount[b + 1];
						if (cost < bestCost) { bestCost = cost; bestSplit = b; bestAxis = ax[pass]; }
					}
				}
				if (bestSplit >= 0)
				{
					const float lo = centroidBounds.lo[bestAxis];
					const float scale = (float)binsUsed / (centroidBounds.hi[bestAxis] - centroidBounds.lo[bestAxis]);
					auto goesLeft = [&](const RtLeafGroup& g) {
						return std::min(binsUsed - 1, std::max(0, (int)((Centroid(g, bestAxis) - lo) * scale))) <= bestSplit; };
					if (big)
					{
						// stable partition through a scratch copy: per-chunk counts, prefix, scatter
						std::vector<uint32_t> lefts(workers + 1, 0);
						unsigned used = 0;
						ParallelChunks(count, workers, [&](unsigned t, uint32_t b, uint32_t e) {
							uint32_t n = 0;
							for (uint32_t i = first + b; i < first + e; ++i) n += goesLeft(groups[i]) ? 1u : 0u;
							lefts[t + 1] = n; });
						used = ChunkThreads(count, workers);
						for (unsigned t = 0; t < used; ++t) lefts[t + 1] += lefts[t];
						const uint32_t totalLeft = lefts[used];
						std::vector<RtLeafGroup> scratch(groups + first, groups + first + count);
						ParallelChunks(count, workers, [&](unsigned t, uint32_t b, uint32_t e) {
							uint32_t l = first + lefts[t], r = first + totalLeft + (b - lefts[t]);
							for (uint32_t i = b; i < e; ++i)
							{
								if (goesLeft(scratch[i])) groups[l++] = scratch[i];
								else groups[r++] = scratch[i];
							} });
						mid = first + totalLeft;
					}
					else
					{
						RtLeafGroup* m = std::partition(groups + first, groups + first + count, goesLeft);
						mid = (uint32_t)(m - groups);
					}
					split = mid > first && mid < first + count;
				}
			}
			if (!split)
			{
				// coincident centroids or two items: split by position along the axis, ties by order
				mid = first + count / 2;
				if (extent > 0.0f)
					std::nth_element(groups + first, groups + mid, groups + first + count,
						[&](const RtLeafGroup& a, const RtLeafGroup& b) { return Centroid(a, axis) < Centroid(b, axis); });
			}

			const uint32_t nl = mid - first, nr = count - nl;
			// pre-order blocks: this node, then the left subtree's nl-1 records, then the right subtree's
			const uint32_t leftBase = nodeBase + 1, rightBase = nodeBase + 1 + (nl - 1);
			Child l, r;
			if (parallelDepth > 0 && count >= 32768)
			{
				auto task = std::async(std::launch::async, [=]() {
					Builder sub{ groups, nodes, searchAllAxes };
					Child c = sub.Build(first, nl, leftBase, parallelDepth - 1);
					return c;
				});
				r = Build(mid, nr, rightBase, parallelDepth - 1);
				l = task.get();
			}
			else
			{
				l = Build(first, nl, leftBase, 0);
				r = Build(mid, nr, rightBase, 0);
			}

			RtNode& rec = nodes[nodeBase];
			memcpy(rec.lmin, l.box.lo, 12); memcpy(rec.lmax, l.box.hi, 12); rec.lref = l.ref; rec.lRefBoxTests = 0;
			memcpy(rec.rmin, r.box.lo, 12); memcpy(rec.rmax, r.box.hi, 12); rec.rref = r.ref; rec.rRefBoxTests = 0;
			me.ref = RT_MAKE_REF(RT_REF_NODE, nodeBase);
			me.box = l.box;
			me.box.Grow(r.box.lo, r.box.hi);      // exact union (min/max of floats): keeps the inclusion monotone
			me.depth = 1 + std::max(l.depth, r.depth);
			return me;
		}
	};
}

void RtBuildSahTree(RtLeafGroups& groups, RtSahResult& out, bool allAxes)
{
	out.nodes.clear();
	out.maxDepth = 0;
	out.cost = 0.0;
	const uint32_t n = (uint32_t)groups.size();
	if (n == 0)
	{
		out.rootRef = RT_MAKE_REF(RT_REF_NONE, RT_REF_INDEX_MASK);
		for (int a = 0; a < 3; ++a) { out.rootMin[a] = 0.0f; out.rootMax[a] = 0.0f; }
		return;
	}
	out.nodes.resize(n - 1);
	Builder builder{ groups.data(), out.nodes.data(), allAxes };
	const Builder::Child root = builder.Build(0, n, 0, 5);
	out.rootRef = root.ref;
	memcpy(out.rootMin, root.box.lo, 12); memcpy(out.rootMax, root.box.hi, 12);
	out.maxDepth = root.depth;
	const double rootArea = root.box.HalfArea();
	double sum = 0.0;
	for (const RtNode& node : out.nodes)
	{
		Box u; u.Reset();
		u.Grow(node.lmin, node.lmax); u.Grow(node.rmin, node.rmax);
		sum += u.HalfArea();
	}
	out.cost = rootArea > 0.0 ? sum / rootArea : 0.0;
}

// Tree rotations (after Kensler 2008): at every inner node, swapping one child with a grandchild under the other child
// changes only that other child's box; the swap that shrinks it most is applied.  Passes run over the records in
// reverse order (children before parents in the builder's pre-order layout).  Boxes stay exact unions of the leaf
// boxes below them, which is all the equivalence argument of bvh_sah.h needs.
namespace
{
	struct Side { float lo[3], hi[3]; uint32_t ref; };
	inline void GetSide(const RtNode& n, int right, Side& s)
	{
		if (right) { memcpy(s.lo, n.rmin, 12); memcpy(s.hi, n.rmax, 12); s.ref = n.rref; }
		else       { memcpy(s.lo, n.lmin, 12); memcpy(s.hi, n.lmax, 12); s.ref = n.lref; }
	}
	inline void SetSide(RtNode& n, int right, const Side& s)
	{
		if (right) { memcpy(n.rmin, s.lo, 12); memcpy(n.rmax, s.hi, 12); n.rref = s.ref; }
		else       { memcpy(n.lmin, s.lo, 12); memcpy(n.lmax, s.hi, 12); n.lref = s.ref; }
	}
	inline double UnionArea(const Side& a, const Side& b, Side* out = nullptr)
	{
		Box u; u.Reset(); u.Grow(a.lo, a.hi); u.Grow(b.lo, b.hi);
		if (out) { memcpy(out->lo, u.lo, 12); memcpy(out->hi, u.hi, 12); }
		return u.HalfArea();
	}
	uint32_t DepthOf(const RtArray<RtNode>& nodes, uint32_t ref)
	{
		// iterative: rotations can make the tree deeper than the host stack likes
		if (RT_REF_KIND(ref) != RT_REF_NODE) return 0;
		std::vector<std::pair<uint32_t, uint32_t>> stack{ { RT_REF_INDEX(ref), 1u } };
		uint32_t deepest = 0;
		while (!stack.empty())
		{
			const auto [i, d] = stack.back(); stack.pop_back();
			deepest = std::max(deepest, d);
			if (RT_REF_KIND(nodes[i].lref) == RT_REF_NODE) stack.push_back({ RT_REF_INDEX(nodes[i].lref), d + 1 });
			if (RT_REF_KIND(nodes[i].rref) == RT_REF_NODE) stack.push_back({ RT_REF_INDEX(nodes[i].rref), d + 1 });
		}
		return deepest;
	}
}

void RtRotateSahTree(RtSahResult& tree, int passes)
{
	RtArray<RtNode>& nodes = tree.nodes;
	if (nodes.size() < 2 || RT_REF_KIND(tree.rootRef) != RT_REF_NODE) return;
	for (int pass = 0; pass < passes; ++pass)
	{
		size_t applied = 0;
		for (size_t i = nodes.size(); i-- > 0;)
		{
			RtNode& p = nodes[i];
			double bestGain = 0.0; int bestX = -1, bestG = -1;
			for (int x = 0; x < 2; ++x)          // x: the child whose box changes (must be inner); the other child s is swapped down
			{
				Side X, S;
				GetSide(p, x, X); GetSide(p, 1 - x, S);
				if (RT_REF_KIND(X.ref) != RT_REF_NODE) continue;
				const RtNode& xn = nodes[RT_REF_INDEX(X.ref)];
				Side g[2];
				GetSide(xn, 0, g[0]); GetSide(xn, 1, g[1]);
				const double before = UnionArea(g[0], g[1]);
				for (int k = 0; k < 2; ++k)      // grandchild k goes up, s takes its place next to grandchild 1-k
				{
					const double gain = before - UnionArea(S, g[1 - k]);
					if (gain > bestGain) { bestGain = gain; bestX = x; bestG = k; }
				}
			}
			if (bestX < 0) continue;
			Side X, S, g[2], merged;
			GetSide(p, bestX, X); GetSide(p, 1 - bestX, S);
			RtNode& xn = nodes[RT_REF_INDEX(X.ref)];
			GetSide(xn, 0, g[0]); GetSide(xn, 1, g[1]);
			SetSide(xn, bestG, S);                                   // s moves under x
			UnionArea(S, g[1 - bestG], &merged);
			merged.ref = X.ref;
			SetSide(p, bestX, merged);                               // x's box is the new union
			SetSide(p, 1 - bestX, g[bestG]);                         // the grandchild moves up
			applied++;
		}
		if (applied == 0) break;
	}
	tree.maxDepth = DepthOf(nodes, tree.rootRef);
	Box rootBox; rootBox.Reset(); rootBox.Grow(tree.rootMin, tree.rootMax);
	const double rootArea = rootBox.HalfArea();
	double sum = 0.0;
	for (const RtNode& node : nodes) { Box u; u.Reset(); u.Grow(node.lmin, node.lmax); u.Grow(node.rmin, node.rmax); sum += u.HalfArea(); }
	tree.cost = rootArea > 0.0 ? sum / rootArea : 0.0;
}

void RtBuildBestSahTree(RtLeafGroups& groups, RtSahResult& out)
{
	const char* axes = getenv("RAYLIB_B200_SAH_AXES");
	if (axes && (atoi(axes) == 1 || atoi(axes) == 3)) { RtBuildSahTree(groups, out, atoi(axes) == 3); return; }
	if (groups.size() < 2) { RtBuildSahTree(groups, out, true); return; }
	const char* rot = getenv("RAYLIB_B200_SAH_ROTATIONS");
	const int passes = rot ? std::max(0, atoi(rot)) : RT_SAH_ROTATION_PASSES;
	RtLeafGroups copy(groups);
	RtSahResult other;
	auto task = std::async(std::launch::async, [&]() { RtBuildSahTree(copy, other, false); if (passes) RtRotateSahTree(other, passes); });
	RtBuildSahTree(groups, out, true);
	if (passes) RtRotateSahTree(out, passes);
	task.get();
	if (getenv("RAYLIB_B200_VERBOSE")) fprintf(stderr, "raylib-b200: SAH tree cost: all axes %.2f, longest axis %.2f\n", out.cost, other.cost);
	if (other.cost < out.cost)
	{
		out.nodes.swap(other.nodes);
		memcpy(out.rootMin, other.rootMin, 12); memcpy(out.rootMax, other.rootMax, 12);
		out.rootRef = other.rootRef; out.maxDepth = other.maxDepth; out.cost = other.cost;
	}
}

// ---------------------------------------------------------------------------------------------
// two-level build

namespace
{
	double AreaSum(const RtNode* nodes, size_t count)
	{
		double sum = 0.0;
		for (size_t i = 0; i < count; ++i) { Box u; u.Reset(); u.Grow(nodes[i].lmin, nodes[i].lmax); u.Grow(nodes[i].rmin, nodes[i].rmax); sum += u.HalfArea(); }
		return sum;
	}

	// One subtree over `items` (permuted in place) as a self-contained RtSahResult whose node indices start at 0: both
	// split policies, one rotation pass each, the tree with the smaller area sum kept (same rule as RtBuildBestSahTree).
	void BuildBestLocal(RtLeafGroups& items, RtSahResult& best, int rotationPasses, int forcedAxes)
	{
		if (forcedAxes == 1 || forcedAxes == 3 || items.size() < 3)
		{
			RtBuildSahTree(items, best, forcedAxes != 1);
			if (rotationPasses) RtRotateSahTree(best, rotationPasses);
			return;
		}
		RtLeafGroups copy(items);
		RtSahResult other;
		RtBuildSahTree(items, best, true);
		RtBuildSahTree(copy, other, false);
		if (rotationPasses) { RtRotateSahTree(best, rotationPasses); RtRotateSahTree(other, rotationPasses); }
		// the two trees span the same root box: their normalised costs compare directly
		if (other.cost < best.cost)
		{
			best.nodes.swap(other.nodes);
			memcpy(best.rootMin, other.rootMin, 12); memcpy(best.rootMax, other.rootMax, 12);
			best.rootRef = other.rootRef; best.maxDepth = other.maxDepth; best.cost = other.cost;
		}
	}
}

void RtBuildTwoLevelSahTree(RtLeafGroups& groups, const std::vector<std::pair<uint32_t, uint32_t>>& meshRanges, RtSahResult& out)
{
	const uint32_t n = (uint32_t)groups.size();
	if (n < 2 || meshRanges.empty()) { RtBuildBestSahTree(groups, out); return; }
	const char* axesEnv = getenv("RAYLIB_B200_SAH_AXES");
	const int forcedAxes = axesEnv ? atoi(axesEnv) : 0;
	const char* rot = getenv("RAYLIB_B200_SAH_ROTATIONS");
	const int passes = rot ? std::max(0, atoi(rot)) : RT_SAH_ROTATION_PASSES;

	// top-level items: one per mesh (filled in once its subtree exists) + every group outside the ranges
	const size_t numMeshes = meshRanges.size();
	RtLeafGroups top(numMeshes);
	{
		uint32_t cursor = 0;
		for (const auto& range : meshRanges)
		{
			for (uint32_t i = cursor; i < range.first; ++i) top.push_back(groups[i]);
			cursor = range.second;
		}
		for (uint32_t i = cursor; i < n; ++i) top.push_back(groups[i]);
	}
	// record layout: the top tree first (parents before children everywhere), then one block per mesh
	const uint32_t topNodes = (uint32_t)top.size() - 1u;
	std::vector<uint32_t> blockBase(numMeshes);
	{
		uint32_t base = topNodes;
		for (size_t m = 0; m < numMeshes; ++m) { blockBase[m] = base; base += (meshRanges[m].second - meshRanges[m].first) - 1u; }
		out.nodes.resize(base);      // == n - 1; not initialised (RtArray): every record is written below
	}
	std::vector<uint32_t> meshDepth(numMeshes, 0);
	std::vector<double> meshArea(numMeshes, 0.0);

	std::atomic<size_t> next(0);
	auto worker = [&]()
	{
		RtLeafGroups items;
		RtSahResult local;
		for (;;)
		{
			const size_t m = next.fetch_add(1);
			if (m >= numMeshes) break;
			const uint32_t first = meshRanges[m].first, count = meshRanges[m].second - meshRanges[m].first;
			items.assign(groups.begin() + first, groups.begin() + first + count);
			BuildBestLocal(items, local, passes, forcedAxes);
			// relocate the block: inner references move by the block base, leaf references stay
			const uint32_t base = blockBase[m];
			for (uint32_t i = 0; i + 1 < count; ++i)
			{
				RtNode rec = local.nodes[i];
				if (RT_REF_KIND(rec.lref) == RT_REF_NODE) rec.lref = RT_MAKE_REF(RT_REF_NODE, RT_REF_INDEX(rec.lref) + base);
				if (RT_REF_KIND(rec.rref) == RT_REF_NODE) rec.rref = RT_MAKE_REF(RT_REF_NODE, RT_REF_INDEX(rec.rref) + base);
				out.nodes[base + i] = rec;
			}
			RtLeafGroup& root = top[m];
			memcpy(root.lo, local.rootMin, 12); memcpy(root.hi, local.rootMax, 12);
			root.ref = RT_REF_KIND(local.rootRef) == RT_REF_NODE ? RT_MAKE_REF(RT_REF_NODE, RT_REF_INDEX(local.rootRef) + base) : local.rootRef;
			meshDepth[m] = local.maxDepth;
			meshArea[m] = AreaSum(local.nodes.data(), local.nodes.size());
		}
	};
	{
		const unsigned threads = std::max(1u, std::min<unsigned>(std::thread::hardware_concurrency(), (unsigned)numMeshes));
		std::vector<std::thread> pool;
		for (unsigned t = 1; t < threads; ++t) pool.emplace_back(worker);
		worker();
		for (std::thread& th : pool) th.join();
	}

	// the top tree: its leaves are mesh roots (references to records beyond the top block) and loose groups
	RtSahResult topTree;
	{
		RtLeafGroups items(top);
		// rotations look INTO inner children: restrict them to the top block by building it stand-alone with the mesh roots
		// disguised as leaves (kind RT_REF_NONE + index of the mesh), and putting the real references back afterwards
		for (size_t m = 0; m < numMeshes; ++m) items[m].ref = RT_MAKE_REF(RT_REF_NONE, (uint32_t)m);
		BuildBestLocal(items, topTree, passes, forcedAxes);
	}
	uint32_t deepest = 0;
	{
		// put the real mesh-root references back, and find the deepest chain of inner nodes through the two levels
		// (a walk from the root: rotations do not keep the records in pre-order)
		for (size_t i = 0; i < topTree.nodes.size(); ++i) out.nodes[i] = topTree.nodes[i];
		auto resolve = [&](uint32_t& ref, uint32_t depth)
		{
			if (RT_REF_KIND(ref) == RT_REF_NONE && RT_REF_INDEX(ref) < numMeshes)
			{
				deepest = std::max(deepest, depth + meshDepth[RT_REF_INDEX(ref)]);
				ref = top[RT_REF_INDEX(ref)].ref;
				return false;
			}
			if (RT_REF_KIND(ref) == RT_REF_NODE) return true;
			deepest = std::max(deepest, depth);
			return false;
		};
		std::vector<std::pair<uint32_t, uint32_t>> stack;
		if (RT_REF_KIND(topTree.rootRef) == RT_REF_NODE) stack.push_back({ RT_REF_INDEX(topTree.rootRef), 1u });
		while (!stack.empty())
		{
			const auto [i, d] = stack.back(); stack.pop_back();
			RtNode& rec = out.nodes[i];
			const uint32_t l = rec.lref, r = rec.rref;
			if (resolve(rec.lref, d)) stack.push_back({ RT_REF_INDEX(l), d + 1 });
			if (resolve(rec.rref, d)) stack.push_back({ RT_REF_INDEX(r), d + 1 });
		}
		if (topTree.nodes.empty()) deepest = meshDepth[0];
	}
	uint32_t rootRef = topTree.rootRef;
	if (RT_REF_KIND(rootRef) == RT_REF_NONE && RT_REF_INDEX(rootRef) < numMeshes) rootRef = top[RT_REF_INDEX(rootRef)].ref;
	out.rootRef = rootRef;
	memcpy(out.rootMin, topTree.rootMin, 12); memcpy(out.rootMax, topTree.rootMax, 12);
	out.maxDepth = deepest;
	Box rootBox; rootBox.Reset(); rootBox.Grow(out.rootMin, out.rootMax);
	const double rootArea = rootBox.HalfArea();
	double sum = AreaSum(topTree.nodes.data(), topTree.nodes.size());
	for (double a : meshArea) sum += a;
	out.cost = rootArea > 0.0 ? sum / rootArea : 0.0;
	if (getenv("RAYLIB_B200_VERBOSE")) fprintf(stderr, "raylib-b200: two-level SAH tree: %zu meshes, %zu top items, cost %.2f, depth %u\n", numMeshes, top.size(), out.cost, out.maxDepth);
}

// ---------------------------------------------------------------------------------------------
// binary -> 4-wide

namespace
{
	struct Slot { uint32_t ref; float lo[3], hi[3]; };

	inline double SlotArea(const Slot& s)
	{
		const double dx = (double)s.hi[0] - s.lo[0], dy = (double)s.hi[1] - s.lo[1], dz = (double)s.hi[2] - s.lo[2];
		if (!(dx >= 0.0) || !(dy >= 0.0) || !(dz >= 0.0)) return 0.0;
		const double cx = std::min(dx, 1.0e18), cy = std::min(dy, 1.0e18), cz = std::min(dz, 1.0e18);
		return cx * cy + cy * cz + cz * cx;
	}

	struct Collapser
	{
		const RtNode* bin;
		RtArray<RtNode4>& out;
		uint32_t maxStack = 0, maxDepth = 0;

		static void Children(const RtNode& n, Slot& l, Slot& r)
		{
			l.ref = n.lref; memcpy(l.lo, n.lmin, 12); memcpy(l.hi, n.lmax, 12);
			r.ref = n.rref; memcpy(r.lo, n.rmin, 12); memcpy(r.hi, n.rmax, 12);
		}

		// ---- optimal collapse (dynamic program over the binary tree, after Ylitie et al. 2017, section 3.1) ----------
		// Leaves are atomic here, so the only cost that depends on the collapse is the sum of the surface areas of the wide
		// nodes (the chance that a node is visited).  cost[k-1][n] = cheapest way to hand the subtree of binary node n to
		// its parent as AT MOST k slots; a leaf costs nothing in any number of slots.
		const std::vector<float>* cost = nullptr;       // [3] arrays, shared read-only between the workers
		bool optimal = false;
		// parallel collapse: subtrees rooted `splitDepth` wide levels down are emitted by worker threads into their own
		// arrays and appended afterwards
		struct Deferred { uint32_t binIndex, stacked, depth, parentNode, parentSlot; };
		std::vector<Deferred>* deferred = nullptr;
		uint32_t splitDepth = 0;

		static double NodeArea(const RtNode& n)
		{
			Slot u;
			for (int a = 0; a < 3; ++a) { u.lo[a] = std::min(n.lmin[a], n.rmin[a]); u.hi[a] = std::max(n.lmax[a], n.rmax[a]); }
			return SlotArea(u);
		}
		double C(uint32_t ref, int k) const { return RT_REF_KIND(ref) == RT_REF_NODE ? (double)cost[k - 1][RT_REF_INDEX(ref)] : 0.0; }
		double D(const RtNode& n, int j, int* outLeft = nullptr) const       // best split of j slots between the two children
		{
			double best = DBL_MAX; int bestK = 1;
			for (int k = 1; k < j; ++k)
			{
				const double c = C(n.lref, k) + C(n.rref, j - k);
				if (c < best) { best = c; bestK = k; }
			}
			if (outLeft) *outLeft = bestK;
			return best;
		}
		void SolveCosts(std::vector<float>* storage, size_t numNodes, uint32_t rootIndex, double rootArea)
		{
			for (int k = 0; k < 3; ++k) storage[k].assign(numNodes, 0.0f);
			cost = storage;
			// children before parents: explicit post-order walk (tree rotations break the builder's pre-order index order).
			// Areas are normalised by the root's so that floats hold the sums.
			const double scale = rootArea > 0.0 ? 1.0 / rootArea : 1.0;
			std::vector<std::pair<uint32_t, bool>> stack{ { rootIndex, false } };
			while (!stack.empty())
			{
				const auto [i, expanded] = stack.back(); stack.pop_back();
				const RtNode& n = bin[i];
				if (!expanded)
				{
					stack.push_back({ i, true });
					if (RT_REF_KIND(n.lref) == RT_REF_NODE) stack.push_back({ RT_REF_INDEX(n.lref), false });
					if (RT_REF_KIND(n.rref) == RT_REF_NODE) stack.push_back({ RT_REF_INDEX(n.rref), false });
					continue;
				}
				const double c1 = NodeArea(n) * scale + D(n, 4);
				const double c2 = std::min(D(n, 2), c1);
				const double c3 = std::min(D(n, 3), c2);
				storage[0][i] = (float)c1; storage[1][i] = (float)c2; storage[2][i] = (float)c3;
			}
			optimal = true;
		}
		// The slots the subtree under `ref` contributes when it may use at most j of them
		void Distribute(uint32_t ref, const float* lo, const float* hi, int j, Slot* slots, uint32_t& n) const
		{
			if (RT_REF_KIND(ref) == RT_REF_NODE && j > 1)
			{
				const uint32_t i = RT_REF_INDEX(ref);
				if (j <= 3 && !(cost[j - 1][i] < cost[j - 2][i])) { Distribute(ref, lo, hi, j - 1, slots, n); return; }
				int k = 1;
				D(bin[i], j, &k);
				Distribute(bin[i].lref, bin[i].lmin, bin[i].lmax, k, slots, n);
				Distribute(bin[i].rref, bin[i].rmin, bin[i].rmax, j - k, slots, n);
				return;
			}
			Slot& s = slots[n++];
			s.ref = ref; memcpy(s.lo, lo, 12); memcpy(s.hi, hi, 12);
		}

		// `stacked`: entries already on the stack when a walk arrives at this node
		uint32_t Emit(uint32_t binIndex, uint32_t stacked, uint32_t depth)
		{
			Slot slots[4];
			uint32_t n = 2;
			Children(bin[binIndex], slots[0], slots[1]);
			if (optimal)
			{
				int k = 1;
				D(bin[binIndex], 4, &k);
				n = 0;
				Distribute(bin[binIndex].lref, bin[binIndex].lmin, bin[binIndex].lmax, k, slots, n);
				Distribute(bin[binIndex].rref, bin[binIndex].rmin, bin[binIndex].rmax, 4 - k, slots, n);
			}
			while (!optimal && n < 4)
			{
				int best = -1; double bestArea = -1.0;
				for (uint32_t i = 0; i < n; ++i)
				{
					if (RT_REF_KIND(slots[i].ref) != RT_REF_NODE) continue;
					const double a = SlotArea(slots[i]);
					if (a > bestArea) { bestArea = a; best = (int)i; }
				}
				if (best < 0) break;
				Slot l, r;
				Children(bin[RT_REF_INDEX(slots[best].ref)], l, r);
				slots[best] = l;
				slots[n++] = r;
			}
			const uint32_t index = (uint32_t)out.size();
			out.push_back(RtNode4());
			maxDepth = std::max(maxDepth, depth + 1);
			maxStack = std::max(maxStack, stacked + (n - 1));
			uint32_t refs[4];
			for (uint32_t i = 0; i < n; ++i)
			{
				refs[i] = slots[i].ref;
				// a walk that descends into one child holds at most the other n-1 siblings on its stack
				if (RT_REF_KIND(slots[i].ref) == RT_REF_NODE)
				{
					if (deferred && depth + 1 == splitDepth)
					{
						deferred->push_back({ RT_REF_INDEX(slots[i].ref), stacked + (n - 1), depth + 1, index, i });
						refs[i] = RT_MAKE_REF(RT_REF_NODE, 0);          // patched when the subtree is appended
					}
					else refs[i] = RT_MAKE_REF(RT_REF_NODE, Emit(RT_REF_INDEX(slots[i].ref), stacked + (n - 1), depth + 1));
				}
			}
			RtNode4& rec = out[index];
			const float inf = std::numeric_limits<float>::infinity();
			for (uint32_t i = 0; i < 4; ++i)
			{
				if (i < n)
				{
					rec.lox[i] = slots[i].lo[0]; rec.loy[i] = slots[i].lo[1]; rec.loz[i] = slots[i].lo[2];
					rec.hix[i] = slots[i].hi[0]; rec.hiy[i] = slots[i].hi[1]; rec.hiz[i] = slots[i].hi[2];
					rec.ref[i] = refs[i];
				}
				else
				{
					rec.lox[i] = rec.loy[i] = rec.loz[i] = inf;
					rec.hix[i] = rec.hiy[i] = rec.hiz[i] = -inf;
					rec.ref[i] = RT_REF_ABSENT;
				}
				rec.pad[i] = 0;
			}
			return index;
		}
	};
}

void RtCollapseToWide(const RtSahResult& binary, RtWideResult& out)
{
	out.nodes.clear();
	out.maxStack = 0; out.maxDepth = 0;
	out.rootRef = binary.rootRef;
	if (RT_REF_KIND(binary.rootRef) != RT_REF_NODE) return;       // empty scene or a single leaf
	out.nodes.reserve(binary.nodes.size() / 2 + 1);
	Collapser c{ binary.nodes.data(), out.nodes };
	// Default: expand the child with the largest surface area first.  RAYLIB_B200_COLLAPSE=dp selects the optimal collapse:
	// 19 % fewer (fuller) wide nodes, but measured neutral on the GPU (node visits per ray 19.8 -> 20.0, frame time +-0.5 %).
	std::vector<float> costStorage[3];
	const char* mode = getenv("RAYLIB_B200_COLLAPSE");
	if (mode && strcmp(mode, "dp") == 0)
	{
		Slot root; memcpy(root.lo, binary.rootMin, 12); memcpy(root.hi, binary.rootMax, 12);
		c.SolveCosts(costStorage, binary.nodes.size(), RT_REF_INDEX(binary.rootRef), SlotArea(root));
	}
	// Large trees: the top five wide levels are emitted here, the (up to 1024) subtrees below them by worker threads into
	// their own arrays, appended in a fixed order -- the layout does not depend on the thread schedule.
	std::vector<Collapser::Deferred> deferred;
	const unsigned threads = std::max(1u, std::thread::hardware_concurrency());
	const char* from = getenv("RAYLIB_B200_COLLAPSE_PARALLEL_FROM");      // tests lower it to cover the parallel path on small scenes
	const size_t parallelFrom = from ? (size_t)std::max(1, atoi(from)) : 262144;
	if (binary.nodes.size() >= parallelFrom && threads > 1) { c.deferred = &deferred; c.splitDepth = from ? 3 : 5; }
	// the recursion is as deep as the wide tree (<= binary depth), fine for the host stack
	out.rootRef = RT_MAKE_REF(RT_REF_NODE, c.Emit(RT_REF_INDEX(binary.rootRef), 0, 0));
	out.maxStack = c.maxStack;
	out.maxDepth = c.maxDepth;
	if (deferred.empty()) return;

	struct Sub { RtArray<RtNode4> nodes; uint32_t maxStack = 0, maxDepth = 0; };
	std::vector<Sub> subs(deferred.size());
	std::atomic<size_t> next{ 0 };
	auto worker = [&]() {
		for (size_t t = next.fetch_add(1); t < deferred.size(); t = next.fetch_add(1))
		{
			Collapser w{ binary.nodes.data(), subs[t].nodes };
			w.cost = c.cost; w.optimal = c.optimal;
			w.Emit(deferred[t].binIndex, deferred[t].stacked, deferred[t].depth);
			subs[t].maxStack = w.maxStack; subs[t].maxDepth = w.maxDepth;
		}
	};
	std::vector<std::thread> pool;
	for (unsigned i = 1; i < std::min<unsigned>(threads, (unsigned)deferred.size()); ++i) pool.emplace_back(worker);
	worker();
	for (std::thread& th : pool) th.join();
	// append: bases by prefix sum, then every worker copies (and re-bases) its share of the subtrees
	std::vector<uint32_t> base(deferred.size());
	uint32_t total = (uint32_t)out.nodes.size();
	for (size_t t = 0; t < deferred.size(); ++t) { base[t] = total; total += (uint32_t)subs[t].nodes.size(); }
	out.nodes.resize(total);
	next = 0;
	auto appender = [&]() {
		for (size_t t = next.fetch_add(1); t < deferred.size(); t = next.fetch_add(1))
		{
			RtNode4* dst = out.nodes.data() + base[t];
			for (size_t k = 0; k < subs[t].nodes.size(); ++k)
			{
				RtNode4 n = subs[t].nodes[k];
				for (int i = 0; i < 4; ++i)
					if (n.ref[i] != RT_REF_ABSENT && RT_REF_KIND(n.ref[i]) == RT_REF_NODE) n.ref[i] = RT_MAKE_REF(RT_REF_NODE, RT_REF_INDEX(n.ref[i]) + base[t]);
				dst[k] = n;
			}
			RtArray<RtNode4>().swap(subs[t].nodes);
		}
	};
	pool.clear();
	for (unsigned i = 1; i < std::min<unsigned>(threads, (unsigned)deferred.size()); ++i) pool.emplace_back(appender);
	appender();
	for (std::thread& th : pool) th.join();
	for (size_t t = 0; t < deferred.size(); ++t)
	{
		out.nodes[deferred[t].parentNode].ref[deferred[t].parentSlot] = RT_MAKE_REF(RT_REF_NODE, base[t]);
		out.maxStack = std::max(out.maxStack, subs[t].maxStack);
		out.maxDepth = std::max(out.maxDepth, subs[t].maxDepth);
	}
}

// ---------------------------------------------------------------------------------------------
// 4-wide -> quantized

namespace
{
	inline float ClampCoord(float v, float whenNaN)
	{
		if (std::isnan(v)) return whenNaN;
		return std::min(RT_Q4_COORD_LIMIT, std::max(-RT_Q4_COORD_LIMIT, v));
	}

	// The 256-value grid (default).  The device decode m = as_float(0x3F000000 | byte << 16) is monotone over ALL byte
	// values: bytes 0..127 give m = 0.5 + byte/256 (steps of S/256), bytes 128..255 give m = 1 + (byte-128)/128 (steps of
	// S/128) -- 1.4921875 S from the first plane to the last.  With a scale that is free (not a power of two) the grid is
	// laid exactly over the node's extent: the finest planes are extent/382 apart and the coarsest extent/191, against
	// extent/120 ... extent/60 for 127 steps of a power-of-two scale.  tools/trav_sim: what quantization costs in node
	// visits falls from 2.9-3.7 % to 0.9-1.1 % (profiles/r02_quantization_cost.jsonl).  Same decode on the device, same
	// guarantee: every byte is chosen by evaluating rt_q4_plane until the decoded box contains the exact one.
	void QuantizeAxisWide(const float* loIn, const float* hiIn, const bool* use, float& outBase, float& outScale, uint32_t& outLoWord, uint32_t& outHiWord)
	{
		float lo[4], hi[4];
		float minLo = std::numeric_limits<float>::infinity(), maxHi = -std::numeric_limits<float>::infinity();
		for (int k = 0; k < 4; ++k)
		{
			lo[k] = ClampCoord(loIn[k], -RT_Q4_COORD_LIMIT); hi[k] = ClampCoord(hiIn[k], RT_Q4_COORD_LIMIT);
			if (use[k]) { minLo = std::min(minLo, lo[k]); maxHi = std::max(maxHi, hi[k]); }
		}
		outLoWord = 0xFFFFFFFFu; outHiWord = 0x00000000u;      // unused slots: lo = top of the grid, hi = bottom (inverted)
		outBase = 0.0f; outScale = 1.0f;
		if (!(minLo <= maxHi)) return;
		const float extent = std::max(maxHi - minLo, 1.0e-30f);
		const float span = 255.0f / 128.0f - 0.5f;                       // m(0xFF) - m(0x00)
		float S = std::max(extent / span * 1.0005f, 1.0e-30f), base = 0.0f;
		for (int attempt = 0; attempt < 200; ++attempt)
		{
			// first plane = fl(base + S/2) must not lie above the smallest lo plane, the last not below the largest hi plane
			const float bump = S * (1.0f / 8388608.0f) + std::fabs(minLo) * (1.0f / 8388608.0f);
			int guard = 0;
			base = minLo - 0.5f * S;
			while (rt_q4_plane(0x00u, S, base) > minLo && guard < 64) base = (minLo - 0.5f * S) - bump * (float)(++guard);
			if (rt_q4_plane(0xFFu, S, base) >= maxHi && guard < 64) break;
			S *= attempt < 8 ? 1.002f : 1.5f;                               // too short after rounding: stretch it
		}
		auto guess = [&](float plane) {
			const float m = (plane - base) / S;
			const float b = m < 1.0f ? (m - 0.5f) * 256.0f : 128.0f + (m - 1.0f) * 128.0f;
			return std::min(255, std::max(0, (int)std::floor(b))); };
		uint32_t loWord = 0, hiWord = 0;
		for (int k = 0; k < 4; ++k)
		{
			uint32_t ql = 255u, qh = 0u;                             // inverted box for unused slots
			if (use[k])
			{
				int q = guess(lo[k]);
				while (q > 0 && rt_q4_plane((uint32_t)q, S, base) > lo[k]) --q;
				while (q < 255 && rt_q4_plane((uint32_t)(q + 1), S, base) <= lo[k]) ++q;
				ql = (uint32_t)q;
				q = std::min(255, guess(hi[k]) + 1);
				while (q < 255 && rt_q4_plane((uint32_t)q, S, base) < hi[k]) ++q;
				while (q > 0 && rt_q4_plane((uint32_t)(q - 1), S, base) >= hi[k]) --q;
				qh = (uint32_t)q;
			}
			loWord |= ql << (8 * k);
			hiWord |= qh << (8 * k);
		}
		outBase = base; outScale = S; outLoWord = loWord; outHiWord = hiWord;
	}

	// RAYLIB_B200_Q4_GRID=7: the round-1 grid (127 steps of a power-of-two scale), kept for A/B runs
	bool UseWideGrid()
	{
		static const bool wide = []() { const char* v = getenv("RAYLIB_B200_Q4_GRID"); return !(v && atoi(v) == 7); }();
		return wide;
	}

	// One axis of one node on the 7-bit grid: base, power-of-two scale and the 8 plane bytes (0x80 | q).
	void QuantizeAxis(const float* loIn, const float* hiIn, const bool* use, float& outBase, float& outScale, uint32_t& outLoWord, uint32_t& outHiWord)
	{
		float lo[4], hi[4];
		float minLo = std::numeric_limits<float>::infinity(), maxHi = -std::numeric_limits<float>::infinity();
		for (int k = 0; k < 4; ++k)
		{
			lo[k] = ClampCoord(loIn[k], -RT_Q4_COORD_LIMIT); hi[k] = ClampCoord(hiIn[k], RT_Q4_COORD_LIMIT);
			if (use[k]) { minLo = std::min(minLo, lo[k]); maxHi = std::max(maxHi, hi[k]); }
		}
		outLoWord = 0xFFFFFFFFu; outHiWord = 0x80808080u;      // unused slots: lo = top of the grid, hi = bottom (inverted)
		outBase = 0.0f; outScale = 1.0f;
		if (!(minLo <= maxHi)) return;
		// S = 2^e with room for the whole extent on 127 steps
		int e = 0;
		const float extent = std::max(maxHi - minLo, 1.0e-30f);
		std::frexp(extent * (128.0f / 120.0f), &e);                 // extent * 1.07 < 2^e
		e = std::min(126, std::max(-100, e));
		float S = std::ldexp(1.0f, e), base = 0.0f;
		for (;;)
		{
			// bottom of the grid = fl(base + S) must not lie above the smallest lo plane
			const float bump = S * (1.0f / 8388608.0f) + std::fabs(minLo) * (1.0f / 8388608.0f);
			int guard = 0;
			base = minLo - S;
			while (rt_q4_plane(0x80u, S, base) > minLo && guard < 64) base = (minLo - S) - bump * (float)(++guard);
			if (rt_q4_plane(0xFFu, S, base) >= maxHi && guard < 64) break;
			if (e >= 126) break;
			S = std::ldexp(1.0f, ++e);                                // grid too short after rounding: double it
		}
		const float step = S / 128.0f;
		const float origin = rt_q4_plane(0x80u, S, base);
		uint32_t loWord = 0, hiWord = 0;
		for (int k = 0; k < 4; ++k)
		{
			uint32_t ql = 127u, qh = 0u;                             // inverted box for unused slots
			if (use[k])
			{
				int q = (int)std::floor((lo[k] - origin) / step);
				q = std::min(127, std::max(0, q));
				while (q > 0 && rt_q4_plane(0x80u | (uint32_t)q, S, base) > lo[k]) --q;
				while (q < 127 && rt_q4_plane(0x80u | (uint32_t)(q + 1), S, base) <= lo[k]) ++q;
				ql = (uint32_t)q;
				q = (int)std::ceil((hi[k] - origin) / step);
				q = std::min(127, std::max(0, q));
				while (q < 127 && rt_q4_plane(0x80u | (uint32_t)q, S, base) < hi[k]) ++q;
				while (q > 0 && rt_q4_plane(0x80u | (uint32_t)(q - 1), S, base) >= hi[k]) --q;
				qh = (uint32_t)q;
			}
			loWord |= (0x80u | ql) << (8 * k);
			hiWord |= (0x80u | qh) << (8 * k);
		}
		outBase = base; outScale = S; outLoWord = loWord; outHiWord = hiWord;
	}
}

void RtQuantizeWide(const RtArray<RtNode4>& wide, RtArray<RtNodeQ4>& out)
{
	out.resize(wide.size());
	const bool wideGrid = UseWideGrid();
	auto range = [&](size_t first, size_t last) {
		for (size_t i = first; i < last; ++i)
		{
			const RtNode4& n = wide[i];
			RtNodeQ4& q = out[i];
			memset(&q, 0, sizeof(q));
			bool use[4];
			for (int k = 0; k < 4; ++k) { use[k] = n.ref[k] != RT_REF_ABSENT; q.ref[k] = n.ref[k]; }
			auto axis = wideGrid ? QuantizeAxisWide : QuantizeAxis;
			axis(n.lox, n.hix, use, q.base[0], q.scaleX, q.qlo[0], q.qhi[0]);
			axis(n.loy, n.hiy, use, q.base[1], q.scaleY, q.qlo[1], q.qhi[1]);
			axis(n.loz, n.hiz, use, q.base[2], q.scaleZ, q.qlo[2], q.qhi[2]);
		}
	};
	// nodes are independent: split the array across the host cores
	const size_t workers = wide.size() < 65536 ? 1 : std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
	if (workers == 1) { range(0, wide.size()); return; }
	std::vector<std::future<void>> tasks;
	const size_t chunk = (wide.size() + workers - 1) / workers;
	for (size_t w = 0; w < workers; ++w)
	{
		const size_t first = w * chunk, last = std::min(wide.size(), first + chunk);
		if (first < last) tasks.push_back(std::async(std::launch::async, range, first, last));
	}
	for (auto& t : tasks) t.get();
}
