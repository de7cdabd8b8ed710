// bvh_reinsert.cc -- development tool, linked into tools/trav_sim only (TRAV_SIM_REINSERT=<iterations>); NOT part of the
// product: measured on the three large scenes it lowers the tree cost by 0.2-1.8 % (node visits per ray by the same) for
// 3x-50x the build time (profiles/r02_reinsertion_experiment.jsonl), so the product keeps binned SAH + rotations.
//
// Insertion-based optimisation of a finished binary SAH tree (after Bittner, Hapala, Havran 2013,
// "Fast insertion-based optimization of bounding volume hierarchies"): a subtree is cut out of the tree and put back
// where the sum of the surface areas of all inner nodes grows least.  The search for that place climbs from the
// subtree's parent towards the root and descends into the sibling of every node on the way, with the area already
// saved (the parent record disappears, the boxes on the path shrink) as the budget that bounds the descent.
//
// What stays true afterwards is all that bvh_sah.h's equivalence argument needs: every inner box is the exact union
// (float min / max, order-independent) of the leaf boxes below it, every leaf group is referenced exactly once, and the
// n - 1 records are laid out parents-before-children (depth-first pre-order).  Which primitive a ray reports does not
// depend on the topology (minimum t, ties to the highest reference rank), only the number of nodes visited does.
//
// Determinism: candidates are searched on a frozen tree (in parallel for large trees: read-only), then applied in one
// fixed order (largest gain first, ties by node index) by one thread; the result does not depend on the thread count.
#include "bvh_sah.h"

#include <algorithm>
#include <atomic>
#include <cstring>
#include <limits>
#include <thread>

namespace
{
	const uint32_t kNone = 0xFFFFFFFFu;

	struct WNode
	{
		float lo[3], hi[3];
		uint32_t parent, left, right;      // leaves: left == kNone
		uint32_t ref;                      // leaves: the leaf reference of the record slot they came from
	};

	inline double HalfArea(const float* lo, const float* hi)
	{
		const double dx = (double)hi[0] - lo[0], dy = (double)hi[1] - lo[1], dz = (double)hi[2] - lo[2];
		if (!(dx >= 0.0) || !(dy >= 0.0) || !(dz >= 0.0)) return 0.0;
		const double cx = std::min(dx, 1.0e18), cy = std::min(dy, 1.0e18), cz = std::min(dz, 1.0e18);      // unbounded gates: finite cost (as bvh_sah.cc)
		return cx * cy + cy * cz + cz * cx;
	}
	inline double HalfArea(const WNode& n) { return HalfArea(n.lo, n.hi); }
	inline double UnionHalfArea(const WNode& a, const float* lo, const float* hi)
	{
		float l[3], h[3];
		for (int k = 0; k < 3; ++k) { l[k] = std::min(a.lo[k], lo[k]); h[k] = std::max(a.hi[k], hi[k]); }
		return HalfArea(l, h);
	}

	struct Move { uint32_t from, to; double gain; };

	struct Optimizer
	{
		std::vector<WNode> T;
		uint32_t root = kNone;
		uint32_t numInner = 0;

		uint32_t Sibling(uint32_t n) const { const WNode& p = T[T[n].parent]; return p.left == n ? p.right : p.left; }

		void Load(const RtSahResult& tree)
		{
			numInner = (uint32_t)tree.nodes.size();
			T.resize((size_t)numInner * 2 + 1);
			uint32_t nextLeaf = numInner;
			for (uint32_t i = 0; i < numInner; ++i)
			{
				const RtNode& rec = tree.nodes[i];
				WNode& w = T[i];
				for (int k = 0; k < 3; ++k) { w.lo[k] = std::min(rec.lmin[k], rec.rmin[k]); w.hi[k] = std::max(rec.lmax[k], rec.rmax[k]); }
				w.ref = 0;
				auto side = [&](uint32_t ref, const float* lo, const float* hi) -> uint32_t
				{
					if (RT_REF_KIND(ref) == RT_REF_NODE && RT_REF_INDEX(ref) < numInner) { T[RT_REF_INDEX(ref)].parent = i; return RT_REF_INDEX(ref); }
					WNode& leaf = T[nextLeaf];
					memcpy(leaf.lo, lo, 12); memcpy(leaf.hi, hi, 12);
					leaf.parent = i; leaf.left = kNone; leaf.right = kNone; leaf.ref = ref;
					return nextLeaf++;
				};
				w.left = side(rec.lref, rec.lmin, rec.lmax);
				w.right = side(rec.rref, rec.rmin, rec.rmax);
			}
			root = RT_REF_INDEX(tree.rootRef);
			T[root].parent = kNone;
		}

		// The best place for `node`: the gain is the decrease of the tree's area sum.
		Move Find(uint32_t node, std::vector<std::pair<double, uint32_t>>& stack) const
		{
			Move best{ node, kNone, 0.0 };
			const uint32_t parent = T[node].parent;
			const WNode& N = T[node];
			const double nodeArea = HalfArea(N);
			double budget = HalfArea(T[parent]);              // the parent record goes away with the cut
			uint32_t sibling = Sibling(node), pivot = parent;
			float pivotLo[3], pivotHi[3];                        // what is left below the pivot once `node` is gone
			memcpy(pivotLo, T[sibling].lo, 12); memcpy(pivotHi, T[sibling].hi, 12);
			for (;;)
			{
				stack.clear();
				stack.push_back({ budget, sibling });
				while (!stack.empty())
				{
					const std::pair<double, uint32_t> top = stack.back(); stack.pop_back();
					if (top.first - nodeArea <= best.gain) continue;          // the new parent is at least as large as `node`
					const WNode& D = T[top.second];
					const double merged = UnionHalfArea(D, N.lo, N.hi);
					const double gain = top.first - merged;
					if (gain > best.gain) { best.to = top.second; best.gain = gain; }
					if (D.left != kNone)
					{
						const double below = gain + HalfArea(D);                // D itself grows to `merged` when the subtree goes below it
						stack.push_back({ below, D.left });
						stack.push_back({ below, D.right });
					}
				}
				if (pivot != parent)
				{
					for (int k = 0; k < 3; ++k) { pivotLo[k] = std::min(pivotLo[k], T[sibling].lo[k]); pivotHi[k] = std::max(pivotHi[k], T[sibling].hi[k]); }
					budget += HalfArea(T[pivot]) - HalfArea(pivotLo, pivotHi);
				}
				const uint32_t up = T[pivot].parent;
				if (up == kNone) break;
				sibling = Sibling(pivot);
				pivot = up;
			}
			if (best.to == Sibling(node)) best.to = kNone;
			return best;
		}

		void Refit(uint32_t n)
		{
			while (n != kNone)
			{
				WNode& w = T[n];
				const WNode& l = T[w.left]; const WNode& r = T[w.right];
				float lo[3], hi[3];
				for (int k = 0; k < 3; ++k) { lo[k] = std::min(l.lo[k], r.lo[k]); hi[k] = std::max(l.hi[k], r.hi[k]); }
				if (memcmp(lo, w.lo, 12) == 0 && memcmp(hi, w.hi, 12) == 0) break;
				memcpy(w.lo, lo, 12); memcpy(w.hi, hi, 12);
				n = w.parent;
			}
		}

		void Apply(const Move& m)
		{
			const uint32_t node = m.from, to = m.to;
			const uint32_t parent = T[node].parent, sibling = Sibling(node), grand = T[parent].parent;
			// cut: the sibling takes the parent's place
			T[sibling].parent = grand;
			if (grand == kNone) root = sibling;
			else
			{
				if (T[grand].left == parent) T[grand].left = sibling; else T[grand].right = sibling;
				Refit(grand);
			}
			// paste: the freed parent record becomes the parent of `to` and `node`
			const uint32_t above = T[to].parent;
			WNode& p = T[parent];
			p.parent = above; p.left = to; p.right = node;
			T[to].parent = parent; T[node].parent = parent;
			for (int k = 0; k < 3; ++k) { p.lo[k] = std::min(T[to].lo[k], T[node].lo[k]); p.hi[k] = std::max(T[to].hi[k], T[node].hi[k]); }
			if (above == kNone) root = parent;
			else
			{
				if (T[above].left == to) T[above].left = parent; else T[above].right = parent;
				Refit(above);
			}
		}

		// true when `to` lies in the subtree of `node` or is its parent (possible only when moves of one batch interact)
		bool Invalid(const Move& m) const
		{
			if (m.to == kNone || m.to == T[m.from].parent || m.to == Sibling(m.from)) return true;
			for (uint32_t n = m.to; n != kNone; n = T[n].parent) if (n == m.from) return true;
			return false;
		}

		void Store(RtSahResult& tree) const
		{
			// pre-order: a node's record, its left subtree, its right subtree
			std::vector<uint32_t> index(numInner, 0);
			{
				uint32_t next = 0;
				std::vector<uint32_t> stack{ root };
				while (!stack.empty())
				{
					const uint32_t n = stack.back(); stack.pop_back();
					index[n] = next++;
					if (T[T[n].right].left != kNone) stack.push_back(T[n].right);
					if (T[T[n].left].left != kNone) stack.push_back(T[n].left);
				}
			}
			uint32_t deepest = 0;
			std::vector<std::pair<uint32_t, uint32_t>> stack{ { root, 1u } };
			while (!stack.empty())
			{
				const auto [n, depth] = stack.back(); stack.pop_back();
				deepest = std::max(deepest, depth);
				RtNode& rec = tree.nodes[index[n]];
				const WNode& l = T[T[n].left]; const WNode& r = T[T[n].right];
				memcpy(rec.lmin, l.lo, 12); memcpy(rec.lmax, l.hi, 12); rec.lRefBoxTests = 0;
				memcpy(rec.rmin, r.lo, 12); memcpy(rec.rmax, r.hi, 12); rec.rRefBoxTests = 0;
				if (l.left != kNone) { rec.lref = RT_MAKE_REF(RT_REF_NODE, index[T[n].left]); stack.push_back({ T[n].left, depth + 1 }); } else rec.lref = l.ref;
				if (r.left != kNone) { rec.rref = RT_MAKE_REF(RT_REF_NODE, index[T[n].right]); stack.push_back({ T[n].right, depth + 1 }); } else rec.rref = r.ref;
			}
			tree.rootRef = RT_MAKE_REF(RT_REF_NODE, index[root]);
			tree.maxDepth = deepest;
			const double rootArea = HalfArea(T[root]);
			double sum = 0.0;
			for (uint32_t i = 0; i < numInner; ++i) sum += HalfArea(T[i]);
			tree.cost = rootArea > 0.0 ? sum / rootArea : 0.0;
		}
	};
}

// Per iteration the `batchFraction` largest subtrees are each offered the place where the area sum grows least;
// non-interacting moves are applied.  `threads` only parallelises the (read-only) searches.
void RtReinsertSahTree(RtSahResult& tree, int iterations, double batchFraction, unsigned threads)
{
	if (iterations <= 0 || tree.nodes.size() < 3 || RT_REF_KIND(tree.rootRef) != RT_REF_NODE) return;
	Optimizer opt;
	opt.Load(tree);
	const uint32_t total = (uint32_t)opt.T.size();
	std::vector<uint32_t> order(total);
	std::vector<double> areas(total);
	std::vector<Move> moves;
	std::vector<uint8_t> touched(total);
	std::vector<std::pair<double, uint32_t>> scratch;
	for (int it = 0; it < iterations; ++it)
	{
		// candidates: the largest subtrees (inner nodes and leaves alike) that are not the root
		uint32_t count = 0;
		for (uint32_t n = 0; n < total; ++n) if (n != opt.root) { order[count++] = n; areas[n] = HalfArea(opt.T[n]); }
		const uint32_t batch = std::max(1u, std::min(count, (uint32_t)((double)count * batchFraction)));
		auto larger = [&](uint32_t a, uint32_t b) { return areas[a] != areas[b] ? areas[a] > areas[b] : a < b; };
		if (batch < count) std::nth_element(order.begin(), order.begin() + batch, order.begin() + count, larger);
		std::sort(order.begin(), order.begin() + batch, larger);
		moves.assign(batch, Move{ kNone, kNone, 0.0 });
		const unsigned workers = batch >= 4096 ? std::max(1u, threads) : 1u;
		if (workers == 1) for (uint32_t i = 0; i < batch; ++i) moves[i] = opt.Find(order[i], scratch);
		else
		{
			std::atomic<uint32_t> next{ 0 };
			auto work = [&]() {
				std::vector<std::pair<double, uint32_t>> mine;
				for (;;)
				{
					const uint32_t b = next.fetch_add(256);
					if (b >= batch) break;
					for (uint32_t i = b; i < std::min(batch, b + 256); ++i) moves[i] = opt.Find(order[i], mine);
				} };
			std::vector<std::thread> pool;
			for (unsigned t = 1; t < workers; ++t) pool.emplace_back(work);
			work();
			for (std::thread& th : pool) th.join();
		}
		std::sort(moves.begin(), moves.end(), [](const Move& a, const Move& b) { return a.gain != b.gain ? a.gain > b.gain : a.from < b.from; });
		std::fill(touched.begin(), touched.end(), 0);
		size_t applied = 0;
		for (const Move& m : moves)
		{
			if (m.to == kNone || !(m.gain > 0.0)) break;
			if (opt.Invalid(m)) continue;
			// moves that share a node with an earlier move of this batch wait for the next iteration: their gain was
			// computed on a tree that no longer exists around them
			const uint32_t near[5] = { m.from, m.to, opt.T[m.from].parent, opt.Sibling(m.from), opt.T[m.to].parent };
			bool clash = false;
			for (uint32_t n : near) if (n != kNone && touched[n]) clash = true;
			if (clash) continue;
			for (uint32_t n : near) if (n != kNone) touched[n] = 1;
			opt.Apply(m);
			applied++;
		}
		if (applied == 0) break;
	}
	opt.Store(tree);
}
