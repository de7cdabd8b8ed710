// compact_mesh.cc -- see compact_mesh.h.
#include "compact_mesh.h"
#include "host_internal.h"
#include "rt_rng.h"
#include "render/material.h"

#include <algorithm>
#include <cstring>
#include <mutex>
#include <unordered_map>

namespace
{
	std::mutex g_mutex;
	std::unordered_map<const StaticMesh*, std::unique_ptr<RtCompactMesh>> g_compact;

	void Store3(float* dst, const vec3& v) { dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; }
}

RtCompactMesh* RtFindCompactMesh(const StaticMesh* mesh)
{
	std::lock_guard<std::mutex> lock(g_mutex);
	auto it = g_compact.find(mesh);
	return it == g_compact.end() ? nullptr : it->second.get();
}

RtCompactMesh* RtAttachCompactMesh(const StaticMesh* mesh)
{
	std::lock_guard<std::mutex> lock(g_mutex);
	std::unique_ptr<RtCompactMesh>& slot = g_compact[mesh];
	slot.reset(new RtCompactMesh);
	return slot.get();
}

void RtDropCompactMesh(const StaticMesh* mesh)
{
	std::lock_guard<std::mutex> lock(g_mutex);
	g_compact.erase(mesh);
}

// StaticMesh::CalculateBounds (geom/static_mesh.cc:40-57) over the position array
void RtCompactCalculateBounds(RtCompactMesh& m)
{
	vec3 lo(FLOAT_MAX, FLOAT_MAX, FLOAT_MAX), hi(-FLOAT_MAX, -FLOAT_MAX, -FLOAT_MAX);
	for (size_t t = 0; t < m.NumTriangles(); ++t)
	{
		const vec3& a = m.positions[3 * t], & b = m.positions[3 * t + 1], & c = m.positions[3 * t + 2];
		lo = min(min(min(lo, a), b), c);
		hi = max(max(max(hi, a), b), c);
	}
	m.bounds = AABB(lo, hi);
	m.boundsValid = true;
}

// StaticMesh::ApplyTransform (geom/static_mesh.cc:59-78): positions through the full transform, normals through its rotation
void RtCompactApplyTransform(RtCompactMesh& m, const Transform& transform)
{
	Transform rotationOnly = transform;
	rotationOnly.SetLocation(vec3(0.0f));
	rotationOnly.SetScale(vec3(1.0f));
	std::vector<vec3> positions(3), normals(3);
	for (size_t t = 0; t < m.NumTriangles(); ++t)
	{
		for (int k = 0; k < 3; ++k) { positions[k] = m.positions[3 * t + k]; normals[k] = m.normals[3 * t + k]; }
		transform.TransformVectors(positions);
		rotationOnly.TransformVectors(normals);
		for (int k = 0; k < 3; ++k) { m.positions[3 * t + k] = positions[k]; m.normals[3 * t + k] = normals[k]; }
	}
	m.boundsValid = false;
}

namespace
{
	// The reference build (geom/bvh.cc:10-80) over triangle INDICES: the same recursion as RtBvhBuilder::Fill in
	// scene_objects.cc -- split axis from the stream keyed by the node's pre-order position, std::sort by the box minimum on
	// that axis (same comparator results as sorting the object pointers, so the same permutation), halves -- but instead of
	// allocating BVHNodes it emits the flattened records in the order RtSceneFlattener::Emit would reach them.
	struct Item { uint32_t tri; AABB box; };

	struct FragmentBuilder
	{
		RtCompactMesh& m;
		uint64_t key;
		std::unordered_map<int32, int64_t> nodeCountMemo;
		std::vector<uint32_t> materialLocal;      // per triangle: index into distinctMaterials
		uint32_t nextTri = 0, nextGate = 0;

		struct Child { uint32_t ref; AABB box; };

		int64_t NodesIn(int32 n)
		{
			if (n <= 2) return 1;
			auto it = nodeCountMemo.find(n);
			if (it != nodeCountMemo.end()) return it->second;
			const int64_t c = 1 + NodesIn(n / 2) + NodesIn(n - n / 2);
			nodeCountMemo[n] = c;
			return c;
		}

		void SortByAxis(Item* items, int32 n, int64_t preorderIndex) const
		{
			const int32 axis = int32(rt_uniform(key, (uint32_t)(preorderIndex + 1)) * 3);
			if (axis == 0) std::sort(items, items + n, [](const Item& l, const Item& r) { return l.box.minBounds.x < r.box.minBounds.x; });
			else if (axis == 1) std::sort(items, items + n, [](const Item& l, const Item& r) { return l.box.minBounds.y < r.box.minBounds.y; });
			else std::sort(items, items + n, [](const Item& l, const Item& r) { return l.box.minBounds.z < r.box.minBounds.z; });
		}

		uint32_t EmitTriangle(const Item& item)
		{
			const uint32_t t = item.tri, index = nextTri++;
			const vec3 v0 = m.positions[3 * t], v1 = m.positions[3 * t + 1], v2 = m.positions[3 * t + 2];
			// Triangle::RefreshDerived (geom/triangle.cc:4-16): unit face normal
			vec3 n = cross(v1 - v0, v2 - v0);
			n.Normalize();
			const vec3 e1 = v1 - v0, e2 = v2 - v0;
			RtTriHot hot;
			hot.q[0] = v0.x; hot.q[1] = v0.y; hot.q[2] = v0.z;
			hot.q[3] = n.x;  hot.q[4] = n.y;  hot.q[5] = n.z;
			hot.q[6] = e1.x; hot.q[7] = e1.y; hot.q[8] = e1.z;
			hot.q[9] = e2.x; hot.q[10] = e2.y; hot.q[11] = e2.z;
			const uint32_t words[4] = { RT_NO_GATE, materialLocal[t], index, 0u };      // rank = local in-order index; type patched at placement
			memcpy(&hot.q[RT_TRI_GATE], words, 16);
			RtTriCold cold;
			Store3(cold.n0, m.normals[3 * t]); Store3(cold.n1, m.normals[3 * t + 1]); Store3(cold.n2, m.normals[3 * t + 2]);
			memcpy(cold.st, &m.texcoords[6 * t], 24);
			cold.material = materialLocal[t];
			m.triHot[index] = hot; m.triCold[index] = cold; m.triGate[index] = RT_NO_GATE;
			RtLeafGroup group;
			Store3(group.lo, item.box.minBounds); Store3(group.hi, item.box.maxBounds);
			group.ref = RT_MAKE_REF(RT_REF_TRI, index);
			m.groups[index] = group;
			return index;
		}

		Child Build(Item* items, int32 n, int64_t preorderIndex, uint32_t nodeDepth)
		{
			SortByAxis(items, n, preorderIndex);
			Child me;
			if (n <= 2)
			{
				// a leaf BVHNode over one or two triangles: no record of its own, one gate for its members
				const uint32_t first = EmitTriangle(items[0]);
				if (n == 2) EmitTriangle(items[1]);
				me.box = n == 2 ? items[0].box + items[1].box : items[0].box + items[0].box;
				me.ref = RT_MAKE_REF(n == 2 ? RT_REF_TRI2 : RT_REF_TRI, first);
				const uint32_t gate = nextGate++;
				const float g8[8] = { me.box.minBounds.x, me.box.minBounds.y, me.box.minBounds.z, 0.0f, me.box.maxBounds.x, me.box.maxBounds.y, me.box.maxBounds.z, 0.0f };
				memcpy(m.gateBoxes.data() + (size_t)gate * 8, g8, sizeof(g8));
				for (int32 k = 0; k < n; ++k)
				{
					m.triGate[first + k] = gate;
					memcpy(&m.triHot[first + k].q[RT_TRI_GATE], &gate, 4);
				}
				return me;
			}
			const uint32_t index = (uint32_t)m.refNodes.size();
			m.refNodes.push_back(RtNode());
			if (nodeDepth + 1 > m.maxNodeDepth) m.maxNodeDepth = nodeDepth + 1;
			const int32 nl = n / 2, nr = n - n / 2;
			const Child l = Build(items, nl, preorderIndex + 1, nodeDepth + 1);
			const Child r = Build(items + nl, nr, preorderIndex + 1 + NodesIn(nl), nodeDepth + 1);
			RtNode& rec = m.refNodes[index];
			Store3(rec.lmin, l.box.minBounds); Store3(rec.lmax, l.box.maxBounds); rec.lref = l.ref; rec.lRefBoxTests = 1;
			Store3(rec.rmin, r.box.minBounds); Store3(rec.rmax, r.box.maxBounds); rec.rref = r.ref; rec.rRefBoxTests = 1;
			me.box = l.box + r.box;
			me.ref = RT_MAKE_REF(RT_REF_NODE, index);
			return me;
		}
	};
}

// StaticMesh::Finalize (geom/static_mesh.cc:80-95) for a compact mesh: bounds, then the reference topology as records.
void RtCompactFinalize(RtCompactMesh& m)
{
	if (m.finalized) return;
	RtCompactCalculateBounds(m);
	const uint32_t n = (uint32_t)m.NumTriangles();
	m.triHot.resize(n); m.triCold.resize(n); m.triGate.resize(n); m.groups.resize(n);
	m.gateBoxes.resize((size_t)n * 8);      // at most one gate per triangle; trimmed below
	m.refNodes.clear(); m.refNodes.reserve(n / 2 + 2);
	m.maxNodeDepth = 0;
	FragmentBuilder b{ m, RtGetBvhBuildKey(), {}, {}, 0, 0 };
	// materials in the order of the triangle list (the flattener registers a mesh's materials in that order)
	b.materialLocal.resize(n);
	{
		std::unordered_map<const Material*, uint32_t> seen;
		const Material* last = nullptr; uint32_t lastIndex = 0;
		for (uint32_t t = 0; t < n; ++t)
		{
			const Material* mat = m.materials[t];
			if (mat != last)
			{
				auto it = seen.find(mat);
				if (it == seen.end()) { it = seen.emplace(mat, (uint32_t)m.distinctMaterials.size()).first; m.distinctMaterials.push_back(mat); }
				last = mat; lastIndex = it->second;
			}
			b.materialLocal[t] = lastIndex;
		}
	}
	if (n == 0) { m.finalized = true; return; }
	std::vector<Item> items(n);
	for (uint32_t t = 0; t < n; ++t)
	{
		const vec3& a = m.positions[3 * t], & bb = m.positions[3 * t + 1], & c = m.positions[3 * t + 2];
		items[t].tri = t;
		items[t].box = AABB(min(min(a, bb), c), max(max(a, bb), c));       // Triangle::RefreshDerived
	}
	const FragmentBuilder::Child top = b.Build(items.data(), (int32)n, 0, 0);
	m.gateBoxes.resize((size_t)b.nextGate * 8);
	m.topRef = top.ref;
	Store3(m.topLo, top.box.minBounds); Store3(m.topHi, top.box.maxBounds);
	m.finalized = true;
	// the face arrays are not needed any more once the records exist
	std::vector<vec3>().swap(m.positions);
	std::vector<vec3>().swap(m.normals);
	std::vector<float>().swap(m.texcoords);
}
