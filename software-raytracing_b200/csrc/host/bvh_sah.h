// bvh_sah.h -- the device traversal tree: a binned-SAH binary BVH built over the reference tree's LEAF
// GROUPS instead of re-using the reference's random-axis median-split topology.
//
// Why this is exact (DESIGN.md "Traversal tree"): a primitive is a candidate in the reference iff every
// BVHNode / StaticMesh box on its root path passes AABB::Hit (geom/bvh.cc:84).  Boxes on that path are
// exact unions of the boxes below them, and the slab arithmetic of geom/aabb.h:41-53 is monotone under box
// inclusion (fl(a-o) and multiplication by the same 1/d are monotone, NaN terms are ignored by the
// comparisons), so "the innermost box passes" already implies "all ancestors pass".  The candidate set is
// therefore { primitives whose GATE box passes }, where the gate is the box of the BVHNode that holds the
// primitive directly -- independent of the topology above it.  Any tree whose inner boxes are exact unions
// of the gates below (again monotone) finds exactly the same candidates; the winner is then chosen by the
// reference rule (minimum t, ties to the highest in-order rank).
#pragma once
#include "rt_scene_format.h"
#include "rt_array.h"
#include <vector>

struct RtLeafGroup
{
	float    lo[3], hi[3];   // gate box: box of the reference BVHNode that owns the primitive(s)
	uint32_t ref;            // RT_REF_TRI / TRI2 / SPHERE / SPHERE2 / CUBE / CUBE2
};

struct RtSahResult
{
	RtArray<RtNode> nodes;
	float    rootMin[3], rootMax[3];
	uint32_t rootRef;
	uint32_t maxDepth;       // deepest chain of inner nodes
	double   cost;           // sum of the surface areas of all inner nodes over the root's: the number of nodes a random
	                         // line through the scene box is expected to cross.  Tracks the measured node visits per ray
	                         // across builders and scenes (tools/trav_sim.cc), unlike the greedy per-split estimate.
};

// Groups are reordered in place (only their order changes).  `allAxes`: every split evaluates the binned SAH on all
// three axes instead of the longest centroid axis only.
typedef RtArray<RtLeafGroup> RtLeafGroups;      // sized by the flattener, filled by its worker threads
void RtBuildSahTree(RtLeafGroups& groups, RtSahResult& out, bool allAxes);

// Builds both variants (concurrently) and keeps the tree with the lower total cost: searching all axes wins on most
// scenes (scatter -6 % node visits per ray, grid -4 %) but the greedy choice loses badly on some (a room with large
// wall triangles: +14 % visits, total cost 69 vs 62) -- the total cost tells which.  RAYLIB_B200_SAH_AXES=1|3 forces one.
void RtBuildBestSahTree(RtLeafGroups& groups, RtSahResult& out);

// Two-level build for scenes made of many meshes (an OBJ with thousands of shapes, an instance scatter): every range
// of `groups` in `meshRanges` ([begin, end) pairs, disjoint, ascending -- the triangles of one StaticMesh) gets its own
// subtree, built, rotated and chosen between the two split policies independently of all others (one task per mesh: the
// build scales with the core count and works on cache-sized arrays), then one small tree is built over the mesh roots
// and the loose groups.  Same node records, same invariants (boxes are exact unions, pre-order: parents before
// children) as the one-level build; the tree differs only where meshes interpenetrate.
void RtBuildTwoLevelSahTree(RtLeafGroups& groups, const std::vector<std::pair<uint32_t, uint32_t>>& meshRanges, RtSahResult& out);

// Tree rotations on a finished binary tree (child <-> grandchild swaps that shrink a node's box); updates cost and depth.
void RtRotateSahTree(RtSahResult& tree, int passes);

// Collapses the binary tree into 4-wide nodes: starting from a node's two children, the inner child with the
// largest surface area is replaced by its own two children until four slots are used.  Child boxes are copied
// from the binary records, so they stay exact unions of the leaf boxes below them (the monotonicity argument
// above carries over unchanged).  Records are laid out in depth-first pre-order.
struct RtWideResult
{
	RtArray<RtNode4> nodes;
	uint32_t rootRef;        // RT_REF_NODE index, or the single leaf reference
	uint32_t maxStack;       // upper bound of simultaneously stacked entries during a near-first walk
	uint32_t maxDepth;
};
void RtCollapseToWide(const RtSahResult& binary, RtWideResult& out);

// Quantizes every wide node to the 64-byte RtNodeQ4 (same indices).  Conservative by construction: each stored
// byte is chosen by evaluating the device's own decode (rt_q4_plane) until the decoded lo plane is <= the exact
// one and the decoded hi plane is >= the exact one.
void RtQuantizeWide(const RtArray<RtNode4>& wide, RtArray<RtNodeQ4>& out);
