/* libm_check.c -- TEST INFRASTRUCTURE.  Compares include/rt_libm.h with the host C library (the functions the compiled
 * reference calls) over ALL 2^32 float arguments of every one-argument function, and over dense grids for powf / atan2f.
 *   gcc -O2 -mfma -ffp-contract=off -pthread oracle/libm_check.c -o oracle/lib/libm_check -lm
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
#ifdef RT_LIBM_HAVE_ALL
WRAP(expf) WRAP(logf) WRAP(tanf) WRAP(asinf) WRAP(acosf) WRAP(atanf)
#endif

int main(int argc, char** argv)
{
	const char* name = argc > 1 ? argv[1] : "sinf";
	const uint64_t stride = argc > 2 ? strtoull(argv[2], NULL, 10) : 1;
	fn1 mine = NULL, ref = NULL;
#define PICK(n) if (!strcmp(name, #n)) { mine = mine_##n; ref = ref_##n; }
	PICK(sinf) PICK(cosf)
#ifdef RT_LIBM_HAVE_ALL
	PICK(expf) PICK(logf) PICK(tanf) PICK(asinf) PICK(acosf) PICK(atanf)
#endif
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
