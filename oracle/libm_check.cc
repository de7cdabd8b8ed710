/* libm_check.cc -- TEST INFRASTRUCTURE.  Compares include/rt_libm.h with the host C library (the functions the compiled
 * reference calls) over ALL 2^32 float arguments of every one-argument function, and over dense grids for powf / atan2f.
 *   g++ -O2 -mfma -ffp-contract=off -pthread oracle/libm_check.cc -o oracle/lib/libm_check -lm
 *   oracle/lib/libm_check sinf [stride]        -> prints "sinf: N mismatches of M" and the first few
 * Results are compared bit for bit; NaN results are considered equal whatever their payload. */
#define _GNU_SOURCE
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../include/rt_libm.h"

typedef float (*fn1)(float);
struct Job { fn1 mine, ref; uint64_t first, last, stride; uint64_t mismatches; uint32_t firstBad[8]; int nbad; };

static int same(float a, float b)
{
	uint32_t x, y; memcpy(&x, &a, 4); memcpy(&y, &b, 4);
	if (x == y) return 1;
	return (a != a) && (b != b);
}

static void* run(void* p)
{
	struct Job* j = (struct Job*)p;
	for (uint64_t i = j->first; i < j->last; i += j->stride)
	{
		float x; uint32_t u = (uint32_t)i; memcpy(&x, &u, 4);
		if (!same(j->mine(x), j->ref(x))) { if (j->nbad < 8) j->firstBad[j->nbad++] = u; j->mismatches++; }
	}
	return NULL;
}

#define WRAP(name) static float mine_##name(float x) { return rt_##name(x); } static float ref_##name(float x) { return name(x); }
WRAP(sinf) WRAP(cosf)
WRAP(expf) WRAP(logf) WRAP(asinf) WRAP(acosf) WRAP(atanf) WRAP(tanf)

// two-argument functions: every float x against a set of fixed second arguments, then pseudo-random pairs
typedef float (*fn2)(float, float);
struct Job2 { fn2 mine, ref; uint64_t first, last, stride; const float* fixed; int nfixed; int fixedIsSecond; uint64_t mismatches; uint32_t bad[8][2]; int nbad; };
static void* run2(void* p)
{
	struct Job2* j = (struct Job2*)p;
	uint64_t rng = 0x9E3779B97F4A7C15ull * (j->first + 1);
	for (uint64_t i = j->first; i < j->last; i += j->stride)
	{
		float x; uint32_t u = (uint32_t)i; memcpy(&x, &u, 4);
		for (int k = 0; k <= j->nfixed; ++k)
		{
			float other;
			if (k < j->nfixed) other = j->fixed[k];
			else { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; uint32_t r = (uint32_t)(rng >> 16); memcpy(&other, &r, 4); }
			const float a = j->fixedIsSecond ? x : other, b = j->fixedIsSecond ? other : x;
			if (!same(j->mine(a, b), j->ref(a, b)))
			{
				if (j->nbad < 8) { memcpy(&j->bad[j->nbad][0], &a, 4); memcpy(&j->bad[j->nbad][1], &b, 4); j->nbad++; }
				j->mismatches++;
			}
		}
	}
	return NULL;
}
static float mine_powf(float x, float y) { return rt_powf(x, y); }
static float ref_powf(float x, float y) { return powf(x, y); }
static float mine_atan2f(float y, float x) { return rt_atan2f(y, x); }
static float ref_atan2f(float y, float x) { return atan2f(y, x); }

static int main2(const char* name, uint64_t stride)
{
	static const float powExponents[] = { 2.2f, 1.0f / 2.2f, 5.0f, 0.5f, 2.0f, 3.0f, -1.0f, 0.4265f, 1.3f, 0.75f, 17.0f, -2.5f };
	static const float atanSeconds[] = { 1.0f, -1.0f, 0.5f, 3.0f, -0.25f, 1e-3f, 1e10f, 0.0f };
	fn2 mine = !strcmp(name, "powf") ? mine_powf : mine_atan2f, ref = !strcmp(name, "powf") ? ref_powf : ref_atan2f;
	enum { T = 8 };
	pthread_t th[T]; static struct Job2 jobs[T];
	for (int t = 0; t < T; ++t)
	{
		memset(&jobs[t], 0, sizeof(jobs[t]));
		jobs[t].mine = mine; jobs[t].ref = ref; jobs[t].stride = stride; jobs[t].fixedIsSecond = 1;
		jobs[t].fixed = !strcmp(name, "powf") ? powExponents : atanSeconds;
		jobs[t].nfixed = !strcmp(name, "powf") ? (int)(sizeof(powExponents) / 4) : (int)(sizeof(atanSeconds) / 4);
		jobs[t].first = (1ull << 32) / T * t; jobs[t].last = (1ull << 32) / T * (t + 1);
		pthread_create(&th[t], NULL, run2, &jobs[t]);
	}
	uint64_t bad = 0, total = 0;
	for (int t = 0; t < T; ++t) { pthread_join(th[t], NULL); bad += jobs[t].mismatches; total += ((1ull << 32) / T / stride) * (jobs[t].nfixed + 1); }
	printf("%s: %llu mismatches of %llu argument pairs\n", name, (unsigned long long)bad, (unsigned long long)total);
	for (int t = 0; t < T; ++t)
		for (int k = 0; k < jobs[t].nbad && k < 2; ++k)
		{
			float a, b; memcpy(&a, &jobs[t].bad[k][0], 4); memcpy(&b, &jobs[t].bad[k][1], 4);
			printf("  (%a, %a): mine %a libm %a\n", a, b, mine(a, b), ref(a, b));
		}
	return bad ? 1 : 0;
}

int main(int argc, char** argv)
{
	const char* name = argc > 1 ? argv[1] : "sinf";
	const uint64_t stride = argc > 2 ? strtoull(argv[2], NULL, 10) : 1;
	if (!strcmp(name, "powf") || !strcmp(name, "atan2f")) return main2(name, stride);
	fn1 mine = NULL, ref = NULL;
#define PICK(n) if (!strcmp(name, #n)) { mine = mine_##n; ref = ref_##n; }
	PICK(sinf) PICK(cosf)
	PICK(expf) PICK(logf) PICK(asinf) PICK(acosf) PICK(atanf) PICK(tanf)
	if (!mine) { fprintf(stderr, "unknown function %s\n", name); return 2; }
	enum { T = 8 };
	pthread_t th[T]; struct Job jobs[T];
	for (int t = 0; t < T; ++t)
	{
		memset(&jobs[t], 0, sizeof(jobs[t]));
		jobs[t].mine = mine; jobs[t].ref = ref; jobs[t].stride = stride;
		jobs[t].first = (1ull << 32) / T * t; jobs[t].last = (1ull << 32) / T * (t + 1);
		pthread_create(&th[t], NULL, run, &jobs[t]);
	}
	uint64_t bad = 0;
	for (int t = 0; t < T; ++t) { pthread_join(th[t], NULL); bad += jobs[t].mismatches; }
	printf("%s: %llu mismatches of %llu arguments\n", name, (unsigned long long)bad, (unsigned long long)((1ull << 32) / stride));
	for (int t = 0; t < T; ++t)
		for (int k = 0; k < jobs[t].nbad && k < 3; ++k)
		{
			float x; memcpy(&x, &jobs[t].firstBad[k], 4);
			float a = mine(x), b = ref(x); uint32_t ua, ub; memcpy(&ua, &a, 4); memcpy(&ub, &b, 4);
			printf("  x=%a (0x%08x): mine %a (0x%08x) libm %a (0x%08x)\n", x, jobs[t].firstBad[k], a, ua, b, ub);
		}
	return bad ? 1 : 0;
}
