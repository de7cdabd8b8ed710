// rt_libm.h -- the float transcendentals of the hot path, restated so that the GPU computes the SAME BITS as the
// reference's host build.
//
// The reference calls sinf / cosf / expf / logf / powf / ... from the C library (core/random.cc:3-50,
// render/material.cc:83-190, render/renderer.cc:166-180).  On the x86-64 hosts this project runs on that is glibc 2.39,
// whose float functions (sysdeps/ieee754/flt-32, the ARM "optimized routines" algorithms) evaluate short polynomials in
// double precision and round once; CUDA's own libdevice implementations differ from them by 1-2 ulp, which a path
// tracer amplifies into different pixels after a few specular bounces.  The functions below follow glibc's algorithms
// operation for operation -- including the fused multiply-adds of the FMA build that glibc's ifunc resolver selects on
// every CPU with AVX2+FMA (checked against the disassembly of __sinf_fma etc.) -- with the constants read from that
// library's tables (tools/libm_extract.py).  fp64 arithmetic and fma are IEEE on both sides, so equal inputs give equal
// outputs; oracle/ checks every function against the host's sinf etc. over all 2^32 arguments (tests/test_cpu_libm.py,
// oracle/libm_check.c).
//
// One header for the device (nvcc, __device__) and for the host checker (gcc -mfma).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define RT_LIBM_FN __device__ __forceinline__
#define RT_LIBM_FMA(a, b, c) __fma_rn((a), (b), (c))
#define RT_LIBM_D2F(x) __double2float_rn(x)
RT_LIBM_FN uint32_t rt_libm_bits(float f) { return __float_as_uint(f); }
RT_LIBM_FN float rt_libm_float(uint32_t u) { return __uint_as_float(u); }
RT_LIBM_FN uint64_t rt_libm_dbits(double d) { return (uint64_t)__double_as_longlong(d); }
RT_LIBM_FN double rt_libm_double(uint64_t u) { return __longlong_as_double((long long)u); }
// cvttsd2si: truncation; the arguments here are always in range
RT_LIBM_FN int32_t rt_libm_d2i(double d) { return __double2int_rz(d); }
#else
#include <string.h>
#define RT_LIBM_FN static inline
#define RT_LIBM_FMA(a, b, c) __builtin_fma((a), (b), (c))
#define RT_LIBM_D2F(x) ((float)(x))
RT_LIBM_FN uint32_t rt_libm_bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
RT_LIBM_FN float rt_libm_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
RT_LIBM_FN uint64_t rt_libm_dbits(double d) { uint64_t u; memcpy(&u, &d, 8); return u; }
RT_LIBM_FN double rt_libm_double(uint64_t u) { double d; memcpy(&d, &u, 8); return d; }
RT_LIBM_FN int32_t rt_libm_d2i(double d) { return (int32_t)d; }
#endif

// ---- sinf / cosf (glibc sysdeps/ieee754/flt-32/s_sinf.c, s_cosf.c, s_sincosf.h, s_sincosf_data.c) -------------------
// Polynomial sets: [0] for quadrants where sin(x+n*pi/2) keeps its sign pattern, [1] negated.  Members in the order the
// code uses them; sign[] is shared.
struct RtSinCosTable { double c0, c1, c2, c3, c4, s1, s2, s3; };
RT_LIBM_FN RtSinCosTable rt_sincos_table(int negated)
{
	RtSinCosTable t;
	t.c0 = 0x1p0;                    t.c1 = -0x1.ffffffd0c621cp-2;    t.c2 = 0x1.55553e1068f19p-5;
	t.c3 = -0x1.6c087e89a359dp-10;   t.c4 = 0x1.99343027bf8c3p-16;
	t.s1 = -0x1.555545995a603p-3;    t.s2 = 0x1.1107605230bc4p-7;     t.s3 = -0x1.994eb3774cf24p-13;
	if (negated) { t.c0 = -t.c0; t.c1 = -t.c1; t.c2 = -t.c2; t.c3 = -t.c3; t.c4 = -t.c4; }
	return t;
}
// x*sign already applied by the caller for the sine branch; n odd = cosine polynomial
RT_LIBM_FN float rt_sinf_poly(double x, double x2, const RtSinCosTable& p, int n)
{
	if ((n & 1) == 0)
	{
		const double x3 = x * x2;
		const double s1 = RT_LIBM_FMA(x2, p.s3, p.s2);
		const double x7 = x3 * x2;
		const double s = RT_LIBM_FMA(x3, p.s1, x);
		return RT_LIBM_D2F(RT_LIBM_FMA(s1, x7, s));
	}
	const double x4 = x2 * x2;
	const double c = RT_LIBM_FMA(x2, p.c1, p.c0);
	const double c2 = RT_LIBM_FMA(x2, p.c4, p.c3);
	const double x6 = x4 * x2;
	const double c1 = RT_LIBM_FMA(x4, p.c2, c);
	return RT_LIBM_D2F(RT_LIBM_FMA(c2, x6, c1));
}
RT_LIBM_FN double rt_sincos_sign(int i) { return (i == 1 || i == 2) ? -1.0 : 1.0; }
// |x| < 120: n = round(x * 2/pi) through a 2^24 fixed-point product, x - n*pi/2 with one fma
RT_LIBM_FN double rt_reduce_fast(double x, int* np)
{
	const double r = x * 0x1.45f306dc9c883p+23;
	const int n = (rt_libm_d2i(r) + 0x800000) >> 24;
	*np = n;
	return RT_LIBM_FMA(-(double)n, 0x1.921fb54442d18p+0, x);
}
// |x| >= 120: 192 bits of 4/pi against the 24-bit mantissa, Payne-Hanek style
RT_LIBM_FN double rt_reduce_large(uint32_t xi, int* np)
{
	const uint32_t inv_pio4[24] = {
		0xa2, 0xa2f9, 0xa2f983, 0xa2f9836e, 0xf9836e4e, 0x836e4e44, 0x6e4e4415, 0x4e441529, 0x441529fc, 0x1529fc27, 0x29fc2757, 0xfc2757d1,
		0x2757d1f5, 0x57d1f534, 0xd1f534dd, 0xf534ddc0, 0x34ddc0db, 0xddc0db62, 0xc0db6295, 0xdb629599, 0x6295993c, 0x95993c43, 0x993c4390, 0x3c439041 };
	const uint32_t* arr = &inv_pio4[(xi >> 26) & 15];
	const int shift = (xi >> 23) & 7;
	uint64_t n, res0, res1, res2;
	xi = (xi & 0xffffff) | 0x800000;
	xi <<= shift;
	res0 = (uint32_t)(xi * arr[0]);
	res1 = (uint64_t)xi * arr[4];
	res2 = (uint64_t)xi * arr[8];
	res0 = (res2 >> 32) | (res0 << 32);
	res0 += res1;
	n = (res0 + (1ULL << 61)) >> 62;
	res0 -= n << 62;
	const double x = (double)(int64_t)res0;
	*np = (int)n;
	return x * 0x1.921fb54442d18p-62;
}
RT_LIBM_FN uint32_t rt_abstop12(float x) { return (rt_libm_bits(x) >> 20) & 0x7ff; }

RT_LIBM_FN float rt_sinf(float y)
{
	double x = (double)y;
	int n;
	const uint32_t top = rt_abstop12(y);
	if (top < 0x3f4u)                       // |y| < pi/4
	{
		if (top < 0x398u) return y;         // |y| < 2^-12: sin y == y to float precision
		return rt_sinf_poly(x, x * x, rt_sincos_table(0), 0);
	}
	if (top < 0x42fu)                       // |y| < 120
	{
		x = rt_reduce_fast(x, &n);
		const double s = rt_sincos_sign(n & 3);
		return rt_sinf_poly(x * s, x * x, rt_sincos_table(n & 2), n);
	}
	if (top < 0x7f8u)
	{
		const uint32_t xi = rt_libm_bits(y);
		const int sign = (int)(xi >> 31);
		x = rt_reduce_large(xi, &n);
		const double s = rt_sincos_sign((n + sign) & 3);
		return rt_sinf_poly(x * s, x * x, rt_sincos_table((n + sign) & 2), n);
	}
	return y - y;                           // inf / nan -> nan
}

RT_LIBM_FN float rt_cosf(float y)
{
	double x = (double)y;
	int n;
	const uint32_t top = rt_abstop12(y);
	if (top < 0x3f4u)
	{
		if (top < 0x398u) return 1.0f;
		return rt_sinf_poly(x, x * x, rt_sincos_table(0), 1);
	}
	if (top < 0x42fu)
	{
		x = rt_reduce_fast(x, &n);
		const double s = rt_sincos_sign(n & 3);
		return rt_sinf_poly(x * s, x * x, rt_sincos_table(n & 2), n ^ 1);
	}
	if (top < 0x7f8u)
	{
		const uint32_t xi = rt_libm_bits(y);
		const int sign = (int)(xi >> 31);
		x = rt_reduce_large(xi, &n);
		const double s = rt_sincos_sign((n + sign) & 3);
		return rt_sinf_poly(x * s, x * x, rt_sincos_table((n + sign) & 2), n ^ 1);
	}
	return y - y;
}
