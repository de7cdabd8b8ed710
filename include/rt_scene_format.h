// rt_scene_format.h -- the flattened, GPU-resident form of a raylib scene.
//
// Raylib_FinalizeScene leaves the client's object graph (BVHNode / StaticMesh /
// Triangle / Sphere / Cube / Material objects scattered over the heap; reference
// geom/bvh.h:19-22, geom/triangle.h:47-63, geom/static_mesh.h:33-42) and the host
// flattener (csrc/host/flatten.cc) re-expresses it as the arrays below.  Plain C
// so the host, the CUDA kernels and the CPU oracle restatement read one definition.
//
// Reference semantics the layout preserves exactly
//   * every BVHNode / StaticMesh box the reference tests (geom/bvh.cc:84,
//     geom/static_mesh.cc:101) is a child box of some RtNode (or the root box);
//     primitives hanging directly under a BVHNode are NOT box-tested by the
//     reference and get an infinite box here;
//   * primitives are emitted in in-order (left-to-right) leaf order, so "index
//     order" within one primitive array == reference tie-break order
//     (geom/bvh.cc:90-93: on equal t the RIGHT child wins); the *Rank arrays give
//     the global in-order leaf rank across primitive kinds.
#pragma once
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

// ---- child reference: 4-bit kind | 28-bit index -----------------------------
#define RT_REF_INDEX_BITS 28u
#define RT_REF_INDEX_MASK 0x0FFFFFFFu
enum RtRefKind
{
	RT_REF_NODE     = 0,   // index into nodes[]
	RT_REF_TRI      = 1,   // one triangle:  tris[index]
	RT_REF_TRI2     = 2,   // a collapsed 2-triangle leaf BVHNode: tris[index], tris[index+1]
	RT_REF_SPHERE   = 3,
	RT_REF_SPHERE2  = 4,
	RT_REF_CUBE     = 5,
	RT_REF_CUBE2    = 6,
	RT_REF_NONE     = 15   // absent child (single-primitive BVHNode keeps left==right, bvh.cc:57-61)
};
#define RT_MAKE_REF(kind, index) (((uint32_t)(kind) << RT_REF_INDEX_BITS) | ((uint32_t)(index) & RT_REF_INDEX_MASK))
#define RT_REF_KIND(ref)  ((uint32_t)(ref) >> RT_REF_INDEX_BITS)
#define RT_REF_INDEX(ref) ((uint32_t)(ref) & RT_REF_INDEX_MASK)

// ---- inner node: both children's boxes + refs, 64 B, 64-B aligned ------------
// Read as four float4: {lmin,lref} {lmax,rref} {rmin,lRefBoxTests} {rmax,rRefBoxTests}
// xRefBoxTests = how many box tests the REFERENCE performs on its way into that child when the box
// passes: 0 for a bare primitive, 1 for a BVHNode, 2 for a StaticMesh (its bounds, then the
// identical root box of its BVH).  Only the statistics build reads it (roofline accounting).
typedef struct RtNode
{
	float    lmin[3]; uint32_t lref;
	float    lmax[3]; uint32_t rref;
	float    rmin[3]; uint32_t lRefBoxTests;
	float    rmax[3]; uint32_t rRefBoxTests;
} RtNode;

// ---- 4-wide inner node of the device traversal tree: 128 B, 128-B aligned ------------------------
// Child boxes in structure-of-arrays form so one 256-bit load brings the same plane of all four children:
// {lox[4] loy[4]} {loz[4] hix[4]} {hiy[4] hiz[4]} {ref[4] pad[4]}.  An absent child has ref = RT_REF_ABSENT and an
// inverted box (lo = +inf, hi = -inf) that no ray passes.  Built by collapsing the binary SAH tree
// (csrc/host/bvh_sah.cc: RtCollapseToWide); the boxes are the exact unions the binary tree holds.
#define RT_REF_ABSENT 0xFFFFFFFFu
typedef struct RtNode4
{
	float    lox[4], loy[4], loz[4];
	float    hix[4], hiy[4], hiz[4];
	uint32_t ref[4];
	uint32_t pad[4];
} RtNode4;

// ---- quantized 4-wide node: 64 B, 64-B aligned -- what the kernels fetch (two 256-bit loads) -----------
// The L1TEX data pipe charges one cycle per lane per load instruction for per-lane gathers
// (tools/microbench/l1_gather.cu), so the node must come in as few loads as possible.
// Child boxes are stored on a 256-value grid local to the node: plane = base + m * S, where the stored byte goes straight
// into the top of a float's fraction (and the lowest exponent bit) with one PRMT:  m = as_float(0x3F000000 | byte << 16),
// i.e. m = 0.5 + byte/256 for bytes 0..127 and m = 1 + (byte-128)/128 for bytes 128..255 -- monotone over all byte values,
// 1.4921875 S from the first plane to the last.  S is a free per-axis scale (the grid is laid over the node's extent,
// base = lowest plane - S/2); plane = fma(m, S, base).  (RAYLIB_B200_Q4_GRID=7 selects the earlier grid: bytes 0x80 | q,
// 127 steps of a power-of-two S.  The decode is the same.)
// The host picks every byte by evaluating this very expression (rt_q4_plane below), lo planes rounded down and
// hi planes rounded up, so a decoded box always CONTAINS the exact box of RtNode4 -- inner boxes only cull, the
// exact verdict comes from the gate test (see bvh_sah.h).  Non-finite boxes are clamped to +-1e18 first.
// An absent child has ref = RT_REF_ABSENT and an inverted box (lo at the top of the grid, hi at the bottom).
typedef struct RtNodeQ4
{
	float    base[3];  float scaleX;      // word 0..3
	uint32_t qlo[3];                      // word 4..6   child k in byte k
	uint32_t qhi[3];                      // word 7..9
	uint32_t ref[4];                      // word 10..13
	float    scaleY, scaleZ;              // word 14..15
} RtNodeQ4;
#define RT_Q4_COORD_LIMIT 1.0e18f

#if defined(__CUDACC__)
#define RT_FMT_FN __host__ __device__ static __forceinline__
#else
#define RT_FMT_FN static inline
#endif
#if !defined(__CUDACC__)
#include <math.h>
// The one definition of the plane decode (host quantizer and CPU oracle; the device spells the same two
// operations with intrinsics).  `byte` is the stored byte.
RT_FMT_FN float rt_q4_plane(uint32_t byte, float scale, float base)
{
	union { uint32_t u; float f; } m; m.u = 0x3F000000u | (byte << 16);
	return fmaf(m.f, scale, base);
}
#endif

// ---- triangle, hot part: 64 B = two 256-bit loads (LDG.E.256 on sm_100a) --------
// q[0..2] = v0, q[3..5] = n (unit face normal), q[6..8] = e1 = v1-v0, q[9..11] = e2 = v2-v0 -- exactly the
// values Triangle::Hit recomputes per ray (geom/triangle.cc:22-33) -- then three integer words:
// q[12] = gate index (RT_NO_GATE: none), q[13] = material index, q[14] = in-order leaf rank, q[15] = material TYPE
// (RtMaterialType: lets k_extend pick the shade queue of a hit without two more dependent loads).
#define RT_TRI_GATE 12
#define RT_TRI_MATERIAL 13
#define RT_TRI_RANK 14
#define RT_TRI_MATTYPE 15
typedef struct RtTriHot { float q[16]; } RtTriHot;

// ---- triangle, cold part (read once per accepted hit): 64 B ------------------
typedef struct RtTriCold
{
	float    n0[3], n1[3], n2[3];   // shading normals
	float    st[6];                 // s0 t0 s1 t1 s2 t2
	uint32_t material;
} RtTriCold;

typedef struct RtSphere { float center[3]; float radius; } RtSphere;   // 16 B

typedef struct RtCube                                                   // 48 B
{
	float    minBounds[3]; float timeStartMove;
	float    maxBounds[3]; uint32_t material;
	float    velocity[3];  uint32_t pad;
} RtCube;

// ---- materials ---------------------------------------------------------------
enum RtMaterialType
{
	RT_MAT_LAMBERTIAN = 0,
	RT_MAT_METAL      = 1,
	RT_MAT_DIELECTRIC = 2,
	RT_MAT_MIRROR     = 3,
	RT_MAT_LIGHT      = 4,
	RT_MAT_MICROFACET = 5,
	RT_MAT_NUM_TYPES  = 6
};
enum { RT_TEX_ALBEDO = 0, RT_TEX_NORMAL = 1, RT_TEX_ROUGHNESS = 2, RT_TEX_METALLIC = 3, RT_TEX_EMISSIVE = 4 };

typedef struct RtMaterial                                               // 64 B
{
	uint32_t type;
	float    color[3];      // albedo | transmission filter | base colour | light intensity | albedo fallback
	float    param0;        // metal fuzziness | dielectric ref_idx | microfacet roughness fallback
	float    param1;        // microfacet metallic fallback
	float    emissive[3];   // microfacet emissive fallback
	int32_t  tex[5];        // RT_TEX_* -> textures[] index, -1 = none
	uint32_t pad[2];
} RtMaterial;

typedef struct RtTexture                                                // 32 B
{
	uint64_t texelOffset;   // in float4 texels into texels[]
	uint32_t width, height;
	uint32_t srgb;          // decode pow(x, 2.2) on all four channels after the fetch
	uint32_t pad[3];
} RtTexture;

// ---- whole scene as handed to rt_scene_upload ---------------------------------
#define RT_NO_GATE 0xFFFFFFFFu
#define RT_SCENE_FLAG_ALPHA_TEST 1u   // some material carries an albedo texture -> cut-out test during traversal

enum RtTreeKind { RT_TREE_REFERENCE = 0, RT_TREE_SAH = 1 };

typedef struct RtSceneDesc
{
	// traversal tree used by the kernels: either the reference topology itself or a SAH tree over the
	// reference's leaf groups (csrc/host/bvh_sah.h explains why both give identical results)
	const RtNode*    nodes;      uint32_t numNodes;
	const RtTriHot*  triHot;
	const RtTriCold* triCold;
	const uint32_t*  triRank;    uint32_t numTris;
	// SAH mode culls triangles by their own (slightly inflated) boxes; a triangle hit only counts if the box the
	// REFERENCE tested last before it -- its gate, the box of the BVHNode holding it -- passes too.
	const uint32_t*  triGate;          // per triangle: index into gateBoxes, RT_NO_GATE = accept without a gate test
	const float*     gateBoxes;  uint32_t numGates;   // 8 floats per gate: min.xyz, 0, max.xyz, 0
	// spheres and cubes sit in the traversal tree with their gate as their box; since the inner-node test became a
	// superset test (quantized boxes, ray-space slabs) their accepted hits need the exact gate check as well
	const uint32_t*  sphereGate;       // per sphere: index into gateBoxes (RT_NO_GATE: none)
	const uint32_t*  cubeGate;         // per cube
	const RtSphere*  spheres;
	const uint32_t*  sphereMaterial;
	const uint32_t*  sphereRank; uint32_t numSpheres;
	const RtCube*    cubes;
	const uint32_t*  cubeRank;   uint32_t numCubes;
	const RtMaterial* materials; uint32_t numMaterials;
	const RtTexture* textures;   uint32_t numTextures;
	const float*     texels;     uint64_t numTexels;      // float4 texels, RGBA

	float    rootMin[3], rootMax[3];   // box of the traversal tree's root
	uint32_t rootRef;
	uint32_t maxStackDepth;            // deepest chain of RT_REF_NODE levels of nodes[]
	uint32_t treeKind;                 // RtTreeKind of nodes[]

	// what the kernels actually walk: nodes[] collapsed to 4-wide records (same root box).  nodes[] itself stays on
	// the host (CPU equivalence tests); only wideNodes is uploaded.
	const RtNode4*   wideNodes;  uint32_t numWideNodes;     // exact boxes (host only: CPU equivalence tests)
	const RtNodeQ4*  quantNodes;                             // same topology and indices, quantized: uploaded
	uint32_t wideRootRef;              // RT_REF_NODE index into wideNodes, or a leaf reference for one-primitive scenes
	uint32_t wideMaxStack;             // most stack entries any root-to-leaf walk can hold (sizes the traversal stack)

	// the reference's own topology, flattened 1:1 (statistics build: "what would the reference traverse";
	// CPU oracle restatement).  Same array as nodes[] when treeKind == RT_TREE_REFERENCE.
	const RtNode*    refNodes;   uint32_t numRefNodes;
	float    refRootMin[3], refRootMax[3];
	uint32_t refRootRef;
	uint32_t refRootBoxTests;
	uint32_t refMaxDepth;
	uint32_t flags;
	uint32_t materialTypeMask;         // bit t set if some material has type t
	uint32_t numLeaves;                // total primitives = highest rank + 1

	int32_t  skyTexture;               // textures[] index of the equirect sky, -1 = none
	float    skyRotation[9];           // rows of Rotator(yaw=90).rotate as computed on the host (renderer.cc:166-168)
	float    sunIlluminance[3];
	float    sunDirection[3];
} RtSceneDesc;

// ---- camera block: the derived frame of render/camera.h:55-78 ------------------
typedef struct RtCamera
{
	float origin[3];     float lensRadius;
	float topLeft[3];    float beginTime;
	float horizontal[3]; float timePeriod;
	float vertical[3];   float pad0;
	float u[3];          float pad1;
	float v[3];          float pad2;
} RtCamera;

#ifdef __cplusplus
}
#endif
