// rt_rng.h -- the ONE definition of the counter-based random stream.
//
// The reference draws every random number from std::random_device-seeded,
// thread_local mt19937 tables (raylib/core/random.h:13-65, raylib/core/random.cc:3-50),
// so two runs never agree.  The B200 renderer replaces that with a stateless
// stream: one 64-bit key per (frame seed, pixel, sample) and a 32-bit draw
// counter.  The oracle's include-shadowed core/random.h (oracle/shadow/) pulls
// in this very header, so CPU reference and GPU consume identical uniforms.
//
// Plain C subset: compiles as C, C++ and CUDA device code.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define RT_RNG_FN __host__ __device__ static __forceinline__
#else
#define RT_RNG_FN static inline
#endif

#define RT_RNG_GOLDEN      0x9E3779B97F4A7C15ull
#define RT_RNG_DEFAULT_FRAME_SEED 1337u
#define RT_RNG_DEFAULT_BVH_KEY    0xB7ull

// splitmix64 finaliser
RT_RNG_FN uint64_t rt_mix64(uint64_t z)
{
	z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull;
	z ^= z >> 27; z *= 0x94D049BB133111EBull;
	z ^= z >> 31;
	return z;
}

// Key of the stream that serves sample `s` of pixel `pixel` (= y*W + x, global
// image coordinates, so the stream is independent of tiling and GPU count).
RT_RNG_FN uint64_t rt_sample_key(uint64_t frameSeed, uint32_t pixel, uint32_t s)
{
	uint64_t k = rt_mix64(frameSeed + RT_RNG_GOLDEN);
	k = rt_mix64(k ^ ((uint64_t)pixel * 0xD6E8FEB86659FD93ull + 0x2545F4914F6CDD1Dull));
	k = rt_mix64(k + (uint64_t)s * RT_RNG_GOLDEN + 1ull);
	return k;
}

// n-th draw (n = 1, 2, 3, ...) of the stream `key`: top 24 bits -> [0,1).
RT_RNG_FN float rt_uniform(uint64_t key, uint32_t n)
{
	uint64_t h = rt_mix64(key + (uint64_t)n * RT_RNG_GOLDEN);
	return (float)(uint32_t)(h >> 40) * (1.0f / 16777216.0f);
}
