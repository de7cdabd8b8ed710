// quantizer_stress.cc -- CPU test helper (tests/test_cpu_host.py::test_quantizer_stress): RtQuantizeWide
// (csrc/host/bvh_sah.cc) on 400 000 synthetic 4-wide nodes that cover what scenes throw at it -- ordinary boxes, boxes
// 1e5 and 3e7 units from the origin (planes closer than one ulp), denormal and 1e15-sized extents, flat boxes,
// unbounded (infinite) gates, absent children.  Every decoded child box must contain the exact one (clamped to
// +-RT_Q4_COORD_LIMIT) with finite, positive scales.  Prints the mean relative padding of the ordinary nodes.
// Built by the test with:  g++ -std=c++17 -O2 -pthread -Iinclude -Isoftware-raytracing_b200/csrc/host <this> .../bvh_sah.cc
#include "bvh_sah.h"
#include <cstdio>
#include <cmath>
#include <random>
#include <limits>
int main()
{
	std::mt19937 rng(1);
	std::uniform_real_distribution<float> U(0.0f, 1.0f);
	const size_t N = 400000;
	RtArray<RtNode4> wide; wide.resize(N);
	const float inf = std::numeric_limits<float>::infinity();
	for (size_t i = 0; i < N; ++i)
	{
		RtNode4& n = wide[i];
		const int mode = (int)(i % 8);
		const float centre = mode == 1 ? 1.0e5f * (U(rng) - 0.5f) : mode == 2 ? 3.0e7f : mode == 3 ? 0.0f : 10.0f * (U(rng) - 0.5f);
		const float size = mode == 4 ? 1.0e-12f : mode == 5 ? 1.0e15f : mode == 3 ? 1.0e-38f : std::pow(10.0f, 6.0f * U(rng) - 4.0f);
		const int count = 2 + (int)(U(rng) * 3.0f); 
		for (int k = 0; k < 4; ++k)
		{
			float* lo[3] = { &n.lox[k], &n.loy[k], &n.loz[k] }; float* hi[3] = { &n.hix[k], &n.hiy[k], &n.hiz[k] };
			if (k >= count) { for (int a = 0; a < 3; ++a) { *lo[a] = inf; *hi[a] = -inf; } n.ref[k] = RT_REF_ABSENT; continue; }
			n.ref[k] = RT_MAKE_REF(RT_REF_TRI, (uint32_t)i);
			for (int a = 0; a < 3; ++a)
			{
				float x = centre + size * (U(rng) - 0.5f), y = centre + size * (U(rng) - 0.5f);
				if (mode == 6 && U(rng) < 0.3f) x = y;                     // flat boxes
				if (mode == 7 && U(rng) < 0.2f) { x = -inf; }              // unbounded
				if (mode == 7 && U(rng) < 0.2f) { y = inf; }
				*lo[a] = std::min(x, y); *hi[a] = std::max(x, y);
			}
		}
	}
	RtArray<RtNodeQ4> q;
	RtQuantizeWide(wide, q);
	size_t bad = 0; double slackSum = 0; size_t slackN = 0;
	for (size_t i = 0; i < N; ++i)
	{
		const RtNode4& w = wide[i]; const RtNodeQ4& n = q[i];
		const float sc[3] = { n.scaleX, n.scaleY, n.scaleZ };
		for (int k = 0; k < 4; ++k)
		{
			if (w.ref[k] == RT_REF_ABSENT) continue;
			const float wl[3] = { w.lox[k], w.loy[k], w.loz[k] }, wh[3] = { w.hix[k], w.hiy[k], w.hiz[k] };
			for (int a = 0; a < 3; ++a)
			{
				const float lo = rt_q4_plane((n.qlo[a] >> (8 * k)) & 255u, sc[a], n.base[a]), hi = rt_q4_plane((n.qhi[a] >> (8 * k)) & 255u, sc[a], n.base[a]);
				const float el = std::max(wl[a], -RT_Q4_COORD_LIMIT), eh = std::min(wh[a], RT_Q4_COORD_LIMIT);
				if (!(lo <= el) || !(hi >= eh) || !std::isfinite(sc[a]) || !std::isfinite(n.base[a]) || !(sc[a] > 0)) { if (bad < 5) printf("bad node %zu mode %zu: lo %g <= %g, hi %g >= %g S %g base %g\n", i, i % 8, lo, el, hi, eh, sc[a], n.base[a]); bad++; }
				if (i % 8 == 0 && eh > el) { slackSum += ((el - lo) + (hi - eh)) / (eh - el + 1e-30); slackN++; }
			}
		}
	}
	printf("violations %zu of %zu nodes; mean relative padding (ordinary nodes) %.4f\n", bad, N, slackSum / slackN);
	return bad != 0;
}
