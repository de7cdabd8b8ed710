// rt_math.cuh -- float3 helpers whose operation ORDER mirrors the reference's vec3
// (raylib/core/vec3.h:10-229).  This translation unit is compiled with --fmad=false,
// IEEE division and square root, so every expression below rounds exactly like the
// reference built with g++ -O2 -ffp-contract=off on x86-64.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define RT_DEV __device__ __forceinline__

// Transcendentals.  RT_EXACT_LIBM 1 (default): include/rt_libm.h, glibc's float algorithms restated operation for operation
// (double-precision polynomials, the library's own tables): the shaders consume the SAME BITS as the reference's host build,
// so the only remaining differences from the oracle are the documented epsilon ties of the traversal.  0: CUDA's libdevice
// functions (1-2 ulp away from glibc; what round 1 shipped).
#ifndef RT_EXACT_LIBM
#define RT_EXACT_LIBM 1
#endif
#if RT_EXACT_LIBM
#include "rt_libm.h"
#define rt_m_sinf rt_sinf
#define rt_m_cosf rt_cosf
#define rt_m_tanf rt_tanf
#define rt_m_asinf rt_asinf
#define rt_m_acosf rt_acosf
#define rt_m_atanf rt_atanf
#define rt_m_atan2f rt_atan2f
#define rt_m_expf rt_expf
#define rt_m_logf rt_logf
#define rt_m_powf rt_powf
#else
#define rt_m_sinf sinf
#define rt_m_cosf cosf
#define rt_m_tanf tanf
#define rt_m_asinf asinf
#define rt_m_acosf acosf
#define rt_m_atanf atanf
#define rt_m_atan2f atan2f
#define rt_m_expf expf
#define rt_m_logf logf
#define rt_m_powf powf
#endif

RT_DEV float3 v3(float x, float y, float z) { return make_float3(x, y, z); }
RT_DEV float3 v3(float s) { return make_float3(s, s, s); }
RT_DEV float3 v3(const float* p) { return make_float3(p[0], p[1], p[2]); }

RT_DEV float3 operator+(float3 a, float3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
RT_DEV float3 operator-(float3 a, float3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
RT_DEV float3 operator*(float3 a, float3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
RT_DEV float3 operator-(float3 a) { return v3(-a.x, -a.y, -a.z); }
RT_DEV float3 operator*(float3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
RT_DEV float3 operator*(float s, float3 a) { return v3(a.x * s, a.y * s, a.z * s); }
RT_DEV float3 operator/(float3 a, float s) { return v3(a.x / s, a.y / s, a.z / s); }     // true division (vec3.h:92-94)
RT_DEV float3 operator+(float3 a, float s) { return v3(a.x + s, a.y + s, a.z + s); }
RT_DEV float3 operator+(float s, float3 a) { return v3(a.x + s, a.y + s, a.z + s); }
RT_DEV float3 operator-(float s, float3 a) { return v3(s - a.x, s - a.y, s - a.z); }
RT_DEV float3 operator-(float3 a, float s) { return v3(a.x - s, a.y - s, a.z - s); }

RT_DEV float dot3(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
RT_DEV float absdot3(float3 a, float3 b) { return fabsf(a.x * b.x + a.y * b.y + a.z * b.z); }
RT_DEV float length3(float3 a) { return sqrtf(a.x * a.x + a.y * a.y + a.z * a.z); }
RT_DEV float3 cross3(float3 a, float3 b)
{
	return v3(a.y * b.z - a.z * b.y, -(a.x * b.z - a.z * b.x), a.x * b.y - a.y * b.x);
}
// vec3::Normalize: k = 1/len, three multiplies (vec3.h:49-54)
RT_DEV float3 normalize3(float3 a) { const float k = 1.0f / length3(a); return v3(a.x * k, a.y * k, a.z * k); }
// vec3::operator/=(float): multiply by the reciprocal (vec3.h:214-220)
RT_DEV float3 div_assign3(float3 a, float s) { const float k = 1.0f / s; return v3(a.x * k, a.y * k, a.z * k); }
RT_DEV float3 reflect3(float3 v, float3 n) { return v - 2.0f * dot3(v, n) * n; }
RT_DEV float3 mix3(float3 a, float3 b, float t) { return (1.0f - t) * a + t * b; }
RT_DEV float3 xyz(float4 q) { return v3(q.x, q.y, q.z); }

// 128-bit read-only loads of scene records (ld.global.nc.v4)
RT_DEV float4 ldg4(const float4* p) { return __ldg(p); }

// 256-bit read-only load (sm_100a: LDG.E.ENL2.256.CONSTANT).  Scene records are 64 B: two of these per node or
// triangle put half as many wavefronts through L1TEX as four 128-bit loads -- the traversal kernels are bound
// by that pipe, not by DRAM or issue slots (profiles/README.md).  `p` must be 32-byte aligned.
struct __align__(32) RtF8 { float4 lo, hi; };
#define RT_LDG8_ASM(hint, r, p) asm volatile("ld.global.nc" hint ".v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" \
		: "=f"(r.lo.x), "=f"(r.lo.y), "=f"(r.lo.z), "=f"(r.lo.w), "=f"(r.hi.x), "=f"(r.hi.y), "=f"(r.hi.z), "=f"(r.hi.w) : "l"(p))
RT_DEV RtF8 ldg8(const void* p) { RtF8 r; RT_LDG8_ASM("", r, p); return r; }
// The same load with an L1 policy.  The traversal stack lives in thread-local memory, i.e. in L1 next to the scene data:
// ncu shows 40 % of the stack pops missing L1 when every triangle record a ray ever tests is allocated there too.
// Inner nodes (re-used by the other lanes and warps of the SM) are kept with evict_last; triangle records (hardly ever
// re-used before eviction) are streamed.  RT_NODE_LOAD_HINT / RT_TRI_LOAD_HINT: 0 plain, 1 no_allocate, 2 evict_first, 3 evict_last.
#ifndef RT_NODE_LOAD_HINT
#define RT_NODE_LOAD_HINT 0
#endif
#ifndef RT_TRI_LOAD_HINT
#define RT_TRI_LOAD_HINT 0
#endif
template<int HINT> RT_DEV RtF8 ldg8_hint(const void* p)
{
	RtF8 r;
	if (HINT == 1) RT_LDG8_ASM(".L1::no_allocate", r, p);
	else if (HINT == 2) RT_LDG8_ASM(".L1::evict_first", r, p);
	else if (HINT == 3) RT_LDG8_ASM(".L1::evict_last", r, p);
	else RT_LDG8_ASM("", r, p);
	return r;
}
// Two IEEE fused multiply-adds in one instruction (sm_100a: FFMA2, fma.rn.f32x2): {a.x*b.x+c.x, a.y*b.y+c.y}, each half
// rounded exactly like __fmaf_rn.  Used by the conservative slab test of the inner nodes (rt_traverse.cuh).
RT_DEV float2 fma2_rn(float2 a, float2 b, float2 c)
{
	float2 d;
	asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\t"
	    "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
	    "fma.rn.f32x2 rd, ra, rb, rc;\n\t"
	    "mov.b64 {%0, %1}, rd;\n\t}"
	    : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
	return d;
}

RT_DEV RtF8 ldg8_node(const void* p) { return ldg8_hint<RT_NODE_LOAD_HINT>(p); }
RT_DEV RtF8 ldg8_tri(const void* p) { return ldg8_hint<RT_TRI_LOAD_HINT>(p); }
