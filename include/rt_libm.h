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
#define RT_LIBM_TABLE static __device__ const      // in global memory (L1-resident): no per-call copy into local memory
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
#define RT_LIBM_TABLE static const
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
RT_LIBM_TABLE uint32_t rt_libm_inv_pio4[24] = {
	0xa2, 0xa2f9, 0xa2f983, 0xa2f9836e, 0xf9836e4e, 0x836e4e44, 0x6e4e4415, 0x4e441529, 0x441529fc, 0x1529fc27, 0x29fc2757, 0xfc2757d1,
	0x2757d1f5, 0x57d1f534, 0xd1f534dd, 0xf534ddc0, 0x34ddc0db, 0xddc0db62, 0xc0db6295, 0xdb629599, 0x6295993c, 0x95993c43, 0x993c4390, 0x3c439041 };
RT_LIBM_FN double rt_reduce_large(uint32_t xi, int* np)
{
	const uint32_t* arr = &rt_libm_inv_pio4[(xi >> 26) & 15];
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

// ---- expf (glibc sysdeps/ieee754/flt-32/e_expf.c, e_exp2f_data.c; N = 32 table entries) ------------------------------
RT_LIBM_TABLE uint64_t rt_libm_exp2f_T[32] = {
		0x3ff0000000000000ull, 0x3fefd9b0d3158574ull, 0x3fefb5586cf9890full, 0x3fef9301d0125b51ull, 0x3fef72b83c7d517bull, 0x3fef54873168b9aaull,
		0x3fef387a6e756238ull, 0x3fef1e9df51fdee1ull, 0x3fef06fe0a31b715ull, 0x3feef1a7373aa9cbull, 0x3feedea64c123422ull, 0x3feece086061892dull,
		0x3feebfdad5362a27ull, 0x3feeb42b569d4f82ull, 0x3feeab07dd485429ull, 0x3feea47eb03a5585ull, 0x3feea09e667f3bcdull, 0x3fee9f75e8ec5f74ull,
		0x3feea11473eb0187ull, 0x3feea589994cce13ull, 0x3feeace5422aa0dbull, 0x3feeb737b0cdc5e5ull, 0x3feec49182a3f090ull, 0x3feed503b23e255dull,
		0x3feee89f995ad3adull, 0x3feeff76f2fb5e47ull, 0x3fef199bdd85529cull, 0x3fef3720dcef9069ull, 0x3fef5818dcfba487ull, 0x3fef7c97337b9b5full,
		0x3fefa4afa2a490daull, 0x3fefd0765b6e4540ull };
RT_LIBM_FN uint64_t rt_exp2f_tab(uint32_t i) { return rt_libm_exp2f_T[i & 31u]; }
// +-0x1.9p-150 rounded to float: the smallest subnormal (what __math_may_uflowf's 0x1.4p-75f * 0x1.4p-75f gives)
RT_LIBM_FN float rt_libm_may_uflow(uint32_t sign) { return rt_libm_float((sign ? 0x80000000u : 0u) | 1u); }
RT_LIBM_FN float rt_libm_inf(uint32_t sign) { return rt_libm_float((sign ? 0x80000000u : 0u) | 0x7f800000u); }
RT_LIBM_FN float rt_libm_zero(uint32_t sign) { return rt_libm_float(sign ? 0x80000000u : 0u); }

RT_LIBM_FN float rt_expf(float x)
{
	const double xd = (double)x;
	const uint32_t abstop = rt_abstop12(x);
	if (abstop >= 0x42bu)                    // |x| >= 88 or nan
	{
		if (rt_libm_bits(x) == 0xff800000u) return 0.0f;
		if (abstop >= 0x7f8u) return x + x;
		if (x > 0x1.62e42ep6f) return rt_libm_inf(0);             // x > log(0x1p128)
		if (x < -0x1.9fe368p6f) return 0.0f;                       // x < log(0x1p-150)
		if (x < -0x1.9d1d9ep6f) return rt_libm_may_uflow(0);       // x < log(0x1p-149)
	}
	const double InvLn2N = 0x1.71547652b82fep+5, SHIFT = 0x1.8p+52;
	double kd = RT_LIBM_FMA(InvLn2N, xd, SHIFT);
	const uint64_t ki = rt_libm_dbits(kd);
	kd -= SHIFT;
	const double r = RT_LIBM_FMA(InvLn2N, xd, -kd);
	const double s = rt_libm_double(rt_exp2f_tab((uint32_t)ki) + (ki << 47));
	const double z = RT_LIBM_FMA(0x1.c6af84b912394p-20, r, 0x1.ebfce50fac4f3p-13);
	const double r2 = r * r;
	double y = RT_LIBM_FMA(r, 0x1.62e42ff0c52d6p-6, 1.0);
	y = RT_LIBM_FMA(z, r2, y);
	return RT_LIBM_D2F(y * s);
}

// ---- logf (glibc sysdeps/ieee754/flt-32/e_logf.c, e_logf_data.c; 16 table entries) -----------------------------------
RT_LIBM_TABLE double rt_libm_logf_T[32] = {
		0x1.661ec79f8f3bep+0, -0x1.57bf7808caadep-2, 0x1.571ed4aaf883dp+0, -0x1.2bef0a7c06ddbp-2, 0x1.49539f0f010b0p+0, -0x1.01eae7f513a67p-2,
		0x1.3c995b0b80385p+0, -0x1.b31d8a68224e9p-3, 0x1.30d190c8864a5p+0, -0x1.6574f0ac07758p-3, 0x1.25e227b0b8ea0p+0, -0x1.1aa2bc79c8100p-3,
		0x1.1bb4a4a1a343fp+0, -0x1.a4e76ce8c0e5ep-4, 0x1.12358f08ae5bap+0, -0x1.1973c5a611cccp-4, 0x1.0953f419900a7p+0, -0x1.252f438e10c1ep-5,
		0x1.0000000000000p+0, 0x0.0p+0, 0x1.e608cfd9a47acp-1, 0x1.aa5aa5df25984p-5, 0x1.ca4b31f026aa0p-1, 0x1.c5e53aa362eb4p-4,
		0x1.b2036576afce6p-1, 0x1.526e57720db08p-3, 0x1.9c2d163a1aa2dp-1, 0x1.bc2860d224770p-3, 0x1.886e6037841edp-1, 0x1.1058bc8a07ee1p-2,
		0x1.767dcf5534862p-1, 0x1.4043057b6ee09p-2 };
RT_LIBM_FN void rt_logf_tab(uint32_t i, double* invc, double* logc) { *invc = rt_libm_logf_T[2u * (i & 15u)]; *logc = rt_libm_logf_T[2u * (i & 15u) + 1u]; }
RT_LIBM_FN float rt_logf(float x)
{
	uint32_t ix = rt_libm_bits(x);
	if (ix == 0x3f800000u) return 0.0f;
	if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u)
	{
		if (ix * 2u == 0u) return rt_libm_inf(1);                  // log(+-0) = -inf
		if (ix == 0x7f800000u) return x;                          // log(inf) = inf
		if ((ix & 0x80000000u) || ix * 2u >= 0xff000000u) return (x - x) / (x - x) + x;      // negative or nan
		ix = rt_libm_bits(x * 0x1p23f);                           // subnormal: normalise
		ix -= 23u << 23;
	}
	const uint32_t tmp = ix - 0x3f330000u;
	const uint32_t i = (tmp >> 19) & 15u;
	const int32_t k = (int32_t)tmp >> 23;
	const uint32_t iz = ix - (tmp & 0xff800000u);
	double invc, logc;
	rt_logf_tab(i, &invc, &logc);
	const double z = (double)rt_libm_float(iz);
	const double r = RT_LIBM_FMA(z, invc, -1.0);
	const double y0 = RT_LIBM_FMA((double)k, 0x1.62e42fefa39efp-1, logc);
	const double r2 = r * r;
	double y = RT_LIBM_FMA(r, 0x1.5575b0be00b6ap-2, -0x1.ffffef20a4123p-2);     // A[1]*r + A[2]
	const double y0r = r + y0;
	y = RT_LIBM_FMA(r2, -0x1.00ea348b88334p-2, y);                              // A[0]*r2 + y
	return RT_LIBM_D2F(RT_LIBM_FMA(r2, y, y0r));
}

// ---- powf (glibc sysdeps/ieee754/flt-32/e_powf.c, e_powf_log2_data.c; log2 table of 16, exp2 table of 32) -----------
RT_LIBM_TABLE double rt_libm_powf_log2_T[32] = {
		0x1.661ec79f8f3bep+0, -0x1.efec65b963019p-2, 0x1.571ed4aaf883dp+0, -0x1.b0b6832d4fca4p-2, 0x1.49539f0f010b0p+0, -0x1.7418b0a1fb77bp-2,
		0x1.3c995b0b80385p+0, -0x1.39de91a6dcf7bp-2, 0x1.30d190c8864a5p+0, -0x1.01d9bf3f2b631p-2, 0x1.25e227b0b8ea0p+0, -0x1.97c1d1b3b7af0p-3,
		0x1.1bb4a4a1a343fp+0, -0x1.2f9e393af3c9fp-3, 0x1.12358f08ae5bap+0, -0x1.960cbbf788d5cp-4, 0x1.0953f419900a7p+0, -0x1.a6f9db6475fcep-5,
		0x1.0000000000000p+0, 0x0.0p+0, 0x1.e608cfd9a47acp-1, 0x1.338ca9f24f53dp-4, 0x1.ca4b31f026aa0p-1, 0x1.476a9543891bap-3,
		0x1.b2036576afce6p-1, 0x1.e840b4ac4e4d2p-3, 0x1.9c2d163a1aa2dp-1, 0x1.40645f0c6651cp-2, 0x1.886e6037841edp-1, 0x1.88e9c2c1b9ff8p-2,
		0x1.767dcf5534862p-1, 0x1.ce0a44eb17bccp-2 };
RT_LIBM_FN void rt_powf_log2_tab(uint32_t i, double* invc, double* logc) { *invc = rt_libm_powf_log2_T[2u * (i & 15u)]; *logc = rt_libm_powf_log2_T[2u * (i & 15u) + 1u]; }
// 0: y is not an integer, 1: odd integer, 2: even integer
RT_LIBM_FN int rt_powf_checkint(uint32_t iy)
{
	const int e = (int)((iy >> 23) & 0xffu);
	if (e < 0x7f) return 0;
	if (e > 0x7f + 23) return 2;
	if (iy & ((1u << (0x7f + 23 - e)) - 1u)) return 0;
	if (iy & (1u << (0x7f + 23 - e))) return 1;
	return 2;
}
RT_LIBM_FN bool rt_powf_zeroinfnan(uint32_t ix) { return 2u * ix - 1u >= 2u * 0x7f800000u - 1u; }
RT_LIBM_FN bool rt_libm_issignaling(uint32_t ix) { return ((ix ^ 0x00400000u) & 0x7fffffffu) > 0x7fc00000u; }

RT_LIBM_FN float rt_powf(float x, float y)
{
	uint32_t signBias = 0;
	uint32_t ix = rt_libm_bits(x);
	const uint32_t iy = rt_libm_bits(y);
	if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u || rt_powf_zeroinfnan(iy))
	{
		if (rt_powf_zeroinfnan(iy))
		{
			if (2u * iy == 0u) return rt_libm_issignaling(ix) ? x + y : 1.0f;
			if (ix == 0x3f800000u) return rt_libm_issignaling(iy) ? x + y : 1.0f;
			if (2u * ix > 2u * 0x7f800000u || 2u * iy > 2u * 0x7f800000u) return x + y;
			if (2u * ix == 2u * 0x3f800000u) return 1.0f;
			if ((2u * ix < 2u * 0x3f800000u) == !(iy & 0x80000000u)) return 0.0f;     // |x| < 1 && y == inf, or |x| > 1 && y == -inf
			return y * y;
		}
		if (rt_powf_zeroinfnan(ix))
		{
			float x2 = x * x;
			if ((ix & 0x80000000u) && rt_powf_checkint(iy) == 1) { x2 = -x2; signBias = 1; }
			if (2u * ix == 0u && (iy & 0x80000000u)) return rt_libm_inf(signBias);
			return (iy & 0x80000000u) ? 1.0f / x2 : x2;
		}
		if (ix & 0x80000000u)                 // finite x < 0
		{
			const int yint = rt_powf_checkint(iy);
			if (yint == 0) return (x - x) / (x - x);
			if (yint == 1) signBias = 1u << 16;
			ix &= 0x7fffffffu;
		}
		if (ix < 0x00800000u)                 // subnormal x: normalise
		{
			ix = rt_libm_bits(x * 0x1p23f);
			ix &= 0x7fffffffu;
			ix -= 23u << 23;
		}
	}
	// log2(x)
	const uint32_t tmp = ix - 0x3f330000u;
	const uint32_t i = (tmp >> 19) & 15u;
	const uint32_t top = tmp & 0xff800000u;
	const uint32_t iz = ix - top;
	const int32_t k = (int32_t)top >> 23;
	double invc, logc;
	rt_powf_log2_tab(i, &invc, &logc);
	const double z = (double)rt_libm_float(iz);
	const double r = RT_LIBM_FMA(z, invc, -1.0);
	const double y0 = logc + (double)k;
	const double A0 = 0x1.27616c9496e0bp-2, A1 = -0x1.71969a075c67ap-2, A2 = 0x1.ec70a6ca7baddp-2, A3 = -0x1.7154748bef6c8p-1, A4 = 0x1.71547652ab82bp+0;
	const double r2 = r * r;
	const double p01 = RT_LIBM_FMA(A0, r, A1);
	const double p23 = RT_LIBM_FMA(A2, r, A3);
	const double r4 = r2 * r2;
	double q = RT_LIBM_FMA(A4, r, y0);
	q = RT_LIBM_FMA(p23, r2, q);
	const double logx = RT_LIBM_FMA(p01, r4, q);
	const double ylogx = (double)y * logx;
	if (((rt_libm_dbits(ylogx) >> 47) & 0xffffu) >= (0x405f800000000000ull >> 47))       // |y * log2(x)| >= 126
	{
		if (ylogx > 0x1.fffffffd1d571p+6) return rt_libm_inf(signBias);
		if (ylogx <= -150.0) return rt_libm_zero(signBias);
		if (ylogx < -149.0) return rt_libm_may_uflow(signBias);
	}
	// exp2(ylogx)
	const double SHIFT = 0x1.8p+47;
	double kd = ylogx + SHIFT;
	const uint64_t ki = rt_libm_dbits(kd);
	kd -= SHIFT;
	const double rr = ylogx - kd;
	const double s = rt_libm_double(rt_exp2f_tab((uint32_t)ki) + ((ki + signBias) << 47));
	const double zz = RT_LIBM_FMA(0x1.c6af84b912394p-5, rr, 0x1.ebfce50fac4f3p-3);
	const double rr2 = rr * rr;
	double yy = RT_LIBM_FMA(rr, 0x1.62e42ff0c52d6p-1, 1.0);
	yy = RT_LIBM_FMA(zz, rr2, yy);
	return RT_LIBM_D2F(yy * s);
}

// ---- asinf / acosf / atanf / atan2f (glibc e_asinf.c, e_acosf.c, s_atanf.c, e_atan2f.c: float arithmetic, no fma) ------
// These translation units are compiled without contraction on both sides (gcc -ffp-contract=off, nvcc --fmad=false).
RT_LIBM_FN float rt_libm_fabsf(float x) { return rt_libm_float(rt_libm_bits(x) & 0x7fffffffu); }
#if defined(__CUDACC__)
RT_LIBM_FN float rt_libm_sqrtf(float x) { return __fsqrt_rn(x); }
RT_LIBM_FN float rt_libm_divf(float a, float b) { return __fdiv_rn(a, b); }
#else
RT_LIBM_FN float rt_libm_sqrtf(float x) { return __builtin_sqrtf(x); }
RT_LIBM_FN float rt_libm_divf(float a, float b) { return a / b; }
#endif

RT_LIBM_FN float rt_asinf(float x)
{
	const float pio2_hi = 0x1.921fb6p+0f, pio2_lo = -0x1.777a5cp-25f, pio4_hi = 0x1.921fb6p-1f;
	const float p0 = 0x1.5555c8p-3f, p1 = 0x1.3301e4p-4f, p2 = 0x1.747e4ap-5f, p3 = 0x1.8c283cp-6f, p4 = 0x1.596d28p-5f;
	const uint32_t hx = rt_libm_bits(x), ix = hx & 0x7fffffffu;
	if (ix == 0x3f800000u) return x * pio2_hi + x * pio2_lo;
	if (ix > 0x3f800000u) return rt_libm_divf(x - x, x - x);
	if (ix < 0x3f000000u)
	{
		if (ix < 0x32000000u) return x;
		const float t = x * x;
		const float w = t * (p0 + t * (p1 + t * (p2 + t * (p3 + t * p4))));
		return x + x * w;
	}
	float w = 1.0f - rt_libm_fabsf(x);
	float t = w * 0.5f;
	float p = t * (p0 + t * (p1 + t * (p2 + t * (p3 + t * p4))));
	const float s = rt_libm_sqrtf(t);
	if (ix >= 0x3f79999au) t = pio2_hi - (2.0f * (s + s * p) - pio2_lo);
	else
	{
		w = rt_libm_float(rt_libm_bits(s) & 0xfffff000u);
		const float c = rt_libm_divf(t - w * w, s + w);
		const float r = p;
		p = 2.0f * s * r - (pio2_lo - 2.0f * c);
		const float q = pio4_hi - 2.0f * w;
		t = pio4_hi - (p - q);
	}
	return (hx & 0x80000000u) ? -t : t;
}

RT_LIBM_FN float rt_acosf(float x)
{
	const float pi = 0x1.921fb4p+1f, pio2_hi = 0x1.921fb4p+0f, pio2_lo = 0x1.4442d0p-24f, twoPio2Lo = 0x1.4442d0p-23f;
	const float pS0 = 0x1.555556p-3f, pS1 = -0x1.4d6120p-2f, pS2 = 0x1.9c1550p-3f, pS3 = -0x1.48228cp-5f, pS4 = 0x1.9efe08p-11f, pS5 = 0x1.23de10p-15f;
	const float qS1 = -0x1.33a272p+1f, qS2 = 0x1.02ae5ap+1f, qS3 = -0x1.6066c2p-1f, qS4 = 0x1.3b8c5cp-4f;
	const uint32_t hx = rt_libm_bits(x), ix = hx & 0x7fffffffu;
	if (ix == 0x3f800000u) return (hx & 0x80000000u) ? twoPio2Lo + pi : 0.0f;
	if (ix > 0x3f800000u) return rt_libm_divf(x - x, x - x);
	if (ix < 0x3f000000u)
	{
		if (ix <= 0x32800000u) return pio2_lo + pio2_hi;
		const float z = x * x;
		const float p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
		const float q = 1.0f + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
		const float r = rt_libm_divf(p, q);
		return pio2_hi - (x - (pio2_lo - r * x));
	}
	if (hx & 0x80000000u)
	{
		const float z = (1.0f + x) * 0.5f;
		const float p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
		const float q = 1.0f + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
		const float s = rt_libm_sqrtf(z);
		const float r = rt_libm_divf(p, q);
		const float w = r * s - pio2_lo;
		return pi - 2.0f * (s + w);
	}
	const float z = (1.0f - x) * 0.5f;
	const float s = rt_libm_sqrtf(z);
	const float df = rt_libm_float(rt_libm_bits(s) & 0xfffff000u);
	const float c = rt_libm_divf(z - df * df, s + df);
	const float p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
	const float q = 1.0f + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
	const float r = rt_libm_divf(p, q);
	const float w = r * s + c;
	return 2.0f * (df + w);
}

RT_LIBM_FN float rt_atanf(float x)
{
	const float aT0 = 0x1.555556p-2f, aT1 = -0x1.99999ap-3f, aT2 = 0x1.24924ap-3f, aT3 = -0x1.c71c70p-4f, aT4 = 0x1.745cdcp-4f, aT5 = -0x1.3b0f2ap-4f;
	const float aT6 = 0x1.10d66ap-4f, aT7 = -0x1.dde2d6p-5f, aT8 = 0x1.97b4b2p-5f, aT9 = -0x1.2b4442p-5f, aT10 = 0x1.0ad3aep-6f;
	const uint32_t hx = rt_libm_bits(x), ix = hx & 0x7fffffffu;
	float hi = 0.0f, lo = 0.0f;
	int id;
	if (ix >= 0x4c000000u)                   // |x| >= 2^25
	{
		if (ix > 0x7f800000u) return x + x;
		return (hx & 0x80000000u) ? -0x1.921fb4p+0f - 0x1.4442d0p-24f : 0x1.4442d0p-24f + 0x1.921fb4p+0f;
	}
	if (ix < 0x3ee00000u)                    // |x| < 0.4375
	{
		if (ix < 0x31000000u) return x;
		id = -1;
	}
	else
	{
		x = rt_libm_fabsf(x);
		if (ix < 0x3f980000u)
		{
			if (ix < 0x3f300000u) { id = 0; x = rt_libm_divf(2.0f * x - 1.0f, 2.0f + x); hi = 0x1.dac670p-2f; lo = 0x1.586ed2p-28f; }
			else { id = 1; x = rt_libm_divf(x - 1.0f, x + 1.0f); hi = 0x1.921fb4p-1f; lo = 0x1.4442d0p-25f; }
		}
		else
		{
			if (ix < 0x401c0000u) { id = 2; x = rt_libm_divf(x - 1.5f, 1.0f + 1.5f * x); hi = 0x1.f730bcp-1f; lo = 0x1.281f68p-25f; }
			else { id = 3; x = rt_libm_divf(-1.0f, x); hi = 0x1.921fb4p+0f; lo = 0x1.4442d0p-24f; }
		}
	}
	const float z = x * x;
	const float w = z * z;
	const float s1 = z * (aT0 + w * (aT2 + w * (aT4 + w * (aT6 + w * (aT8 + w * aT10)))));
	const float s2 = w * (aT1 + w * (aT3 + w * (aT5 + w * (aT7 + w * aT9))));
	if (id < 0) return x - x * (s1 + s2);
	const float r = hi - ((x * (s1 + s2) - lo) - x);
	return (hx & 0x80000000u) ? -r : r;
}

RT_LIBM_FN float rt_atan2f(float y, float x)
{
	const float tiny = 0x1.4484c0p-100f, pi_o_4 = 0x1.921fb6p-1f, pi_o_2 = 0x1.921fb6p+0f, pi = 0x1.921fb6p+1f, pi_lo = -0x1.777a5cp-24f;
	const uint32_t hx = rt_libm_bits(x), hy = rt_libm_bits(y);
	const uint32_t ix = hx & 0x7fffffffu, iy = hy & 0x7fffffffu;
	if (ix > 0x7f800000u || iy > 0x7f800000u) return x + y;
	if (hx == 0x3f800000u) return rt_atanf(y);
	const uint32_t m = ((hy >> 31) & 1u) | ((hx >> 30) & 2u);
	if (iy == 0u)
	{
		if (m < 2u) return y;
		return m == 2u ? pi + tiny : -pi - tiny;
	}
	if (ix == 0u) return (hy & 0x80000000u) ? -pi_o_2 - tiny : pi_o_2 + tiny;
	if (ix == 0x7f800000u)
	{
		if (iy == 0x7f800000u)
		{
			switch (m) { case 0: return pi_o_4 + tiny; case 1: return -pi_o_4 - tiny; case 2: return 3.0f * pi_o_4 + tiny; default: return -3.0f * pi_o_4 - tiny; }
		}
		switch (m) { case 0: return 0.0f; case 1: return -0.0f; case 2: return pi + tiny; default: return -pi - tiny; }
	}
	if (iy == 0x7f800000u) return (hy & 0x80000000u) ? -pi_o_2 - tiny : pi_o_2 + tiny;
	const int32_t k = ((int32_t)iy - (int32_t)ix) >> 23;
	float z;
	if (k > 60) z = pi_o_2 + 0.5f * pi_lo;
	else if ((hx & 0x80000000u) && k < -60) z = 0.0f;
	else z = rt_atanf(rt_libm_fabsf(rt_libm_divf(y, x)));
	switch (m)
	{
	case 0: return z;
	case 1: return rt_libm_float(rt_libm_bits(z) ^ 0x80000000u);
	case 2: return pi - (z - pi_lo);
	default: return (z - pi_lo) - pi;
	}
}

// ---- tanf (glibc s_tanf.c + e_rem_pio2f.c + k_tanf.c: double-precision argument reduction WITHOUT fma, float kernel) ---
RT_LIBM_FN float rt_kernel_tanf(float x, float y, int iy)
{
	const float pio4 = 0x1.921fb4p-1f, pio4lo = 0x1.4442d0p-25f;
	const float T0 = 0x1.555556p-2f, T1 = 0x1.111112p-3f, T2 = 0x1.ba1ba2p-5f, T3 = 0x1.664f48p-6f, T4 = 0x1.226e3ep-7f, T5 = 0x1.d6d22cp-9f, T6 = 0x1.7dbc90p-10f;
	const float T7 = 0x1.344d90p-11f, T8 = 0x1.026f72p-12f, T9 = 0x1.47e88ap-14f, T10 = 0x1.2b80f4p-14f, T11 = -0x1.375cbep-16f, T12 = 0x1.b2a708p-16f;
	const uint32_t hx = rt_libm_bits(x), ix = hx & 0x7fffffffu;
	if (ix < 0x39000000u)                    // |x| < 2^-13
	{
		if ((ix | (uint32_t)(iy + 1)) == 0u) return rt_libm_divf(1.0f, rt_libm_fabsf(x));
		if (iy == 1) return x;
		return rt_libm_divf(-1.0f, x);
	}
	if (ix >= 0x3f2ca140u)                   // |x| >= 0.6744
	{
		if (hx & 0x80000000u) { x = -x; y = -y; }
		const float z0 = pio4 - x;
		const float w0 = pio4lo - y;
		x = z0 + w0; y = 0.0f;
		if (rt_libm_fabsf(x) < 0x1p-13f)
			return (float)((1 - (int)((hx >> 30) & 2u)) * iy) * (1.0f - (float)(2 * iy) * x);
	}
	const float z = x * x;
	const float w = z * z;
	float r = T1 + w * (T3 + w * (T5 + w * (T7 + w * (T9 + w * T11))));
	float v = z * (T2 + w * (T4 + w * (T6 + w * (T8 + w * (T10 + w * T12)))));
	float s = z * x;
	r = y + z * (s * (r + v) + y);
	r += T0 * s;
	const float ww = x + r;
	if (ix >= 0x3f2ca140u)
	{
		v = (float)iy;
		return (float)(1 - (int)((hx >> 30) & 2u)) * (v - 2.0f * (x - (rt_libm_divf(ww * ww, ww + v) - r)));
	}
	if (iy == 1) return ww;
	// -1 / (x + r), accurately
	const float zt = rt_libm_float(rt_libm_bits(ww) & 0xfffff000u);
	v = r - (zt - x);
	const float a = rt_libm_divf(-1.0f, ww);
	const float t = rt_libm_float(rt_libm_bits(a) & 0xfffff000u);
	s = 1.0f + t * zt;
	return t + a * (s + t * v);
}
RT_LIBM_FN float rt_tanf(float x)
{
	const uint32_t ix = rt_libm_bits(x) & 0x7fffffffu;
	if (ix <= 0x3f490fdau) return rt_kernel_tanf(x, 0.0f, 1);
	if (ix >= 0x7f800000u) return x - x;
	double dx = (double)x;
	int n;
	if (rt_abstop12(x) < 0x42fu)
	{
		// reduce_fast as compiled into this translation unit: multiply and subtract, not fused
		const double r = dx * 0x1.45f306dc9c883p+23;
		n = (rt_libm_d2i(r) + 0x800000) >> 24;
		dx = dx - (double)n * 0x1.921fb54442d18p+0;
	}
	else
	{
		const uint32_t xi = rt_libm_bits(x);
		dx = rt_reduce_large(xi, &n);
		if (xi >> 31) dx = -dx;
	}
	const float y0 = RT_LIBM_D2F(dx);
	const float y1 = RT_LIBM_D2F(dx - (double)y0);
	return rt_kernel_tanf(y0, y1, 1 - ((n & 1) << 1));
}
