// core_services.cc -- logging, runtime checks and the host random helpers.
// Behavioural reference: raylib/core/logger.cc:22-96 (async log thread drained every 100 ms),
// raylib/core/assertion.cc:4-26 (print then break), raylib/core/random.cc:3-73.
#include "core/logger.h"
#include "core/assertion.h"
#include "core/random.h"
#include "rt_rng.h"
#include "host_internal.h"

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <csignal>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <deque>
#include <mutex>
#include <set>
#include <string>
#include <thread>

// ---------------------------------------------------------------------------------------------
// Logger

namespace
{
	std::mutex g_logMutex;
	std::deque<std::string> g_logQueue;
	std::atomic<bool> g_logThreadRunning{ false };
	std::atomic<bool> g_logThreadStop{ false };
	std::atomic<bool> g_logThreadExited{ true };

	void DrainLocked(std::unique_lock<std::mutex>& lock)
	{
		while (!g_logQueue.empty())
		{
			std::string line = std::move(g_logQueue.front());
			g_logQueue.pop_front();
			lock.unlock();
			printf("%s\n", line.c_str());
			lock.lock();
		}
		fflush(stdout);
	}

	void LogThreadMain()
	{
		while (!g_logThreadStop.load())
		{
			{
				std::unique_lock<std::mutex> lock(g_logMutex);
				DrainLocked(lock);
			}
			std::this_thread::sleep_for(std::chrono::milliseconds(100));
		}
		{
			std::unique_lock<std::mutex> lock(g_logMutex);
			DrainLocked(lock);
		}
		g_logThreadExited = true;
	}
}

namespace Logger
{
	void StartLogThread()
	{
		bool expected = false;
		if (!g_logThreadRunning.compare_exchange_strong(expected, true)) return;
		g_logThreadStop = false;
		g_logThreadExited = false;
		std::thread(LogThreadMain).detach();   // detached like the reference's (logger.cc:31-32): safe at process exit
	}

	void FlushLogThread()
	{
		if (!g_logThreadRunning.load())
		{
			std::unique_lock<std::mutex> lock(g_logMutex);
			DrainLocked(lock);
			return;
		}
		for (;;)
		{
			{
				std::lock_guard<std::mutex> lock(g_logMutex);
				if (g_logQueue.empty()) break;
			}
			std::this_thread::sleep_for(std::chrono::milliseconds(10));
		}
	}

	void KillAndWaitForLogThread()
	{
		if (!g_logThreadRunning.load()) return;
		g_logThreadStop = true;
		while (!g_logThreadExited.load()) std::this_thread::sleep_for(std::chrono::milliseconds(5));
		g_logThreadRunning = false;
	}
}

void LOG(const char* format, ...)
{
	char buffer[1024];
	va_list ap;
	va_start(ap, format);
	vsnprintf(buffer, sizeof(buffer), format, ap);
	va_end(ap);

	if (!g_logThreadRunning.load())
	{
		// library not initialised (or already terminated): print synchronously instead of queueing forever
		printf("%s\n", buffer);
		fflush(stdout);
		return;
	}
	std::lock_guard<std::mutex> lock(g_logMutex);
	g_logQueue.emplace_back(buffer);
}

// ---------------------------------------------------------------------------------------------
// CHECK

static void AfterFailedCheck()
{
	fflush(stdout);
	const char* trap = getenv("RAYLIB_B200_TRAP");
	if (trap && trap[0] == '1') raise(SIGTRAP);
}

extern "C" void CHECK_IMPL(int x, const char* file, int line)
{
	if (x) return;
	printf("Assertion failed !!! [FILE=%s] [LINE=%d]\n", file, line);
	AfterFailedCheck();
}

extern "C" void CHECKF_IMPL(int x, const char* msg, const char* file, int line)
{
	if (x) return;
	printf("Assertion failed !!! [MSG=%s] [FILE=%s] [LINE=%d]\n", msg, file, line);
	AfterFailedCheck();
}

bool RtHostQueryUnsupported(const char* what)
{
	static std::mutex m;
	static std::set<std::string> reported;
	std::lock_guard<std::mutex> lock(m);
	if (reported.insert(what).second)
	{
		fprintf(stderr, "raylib-b200: %s was called on the host. Ray queries and scattering run on the GPU only "
		                "(there is no CPU rendering path); the call returns false.\n", what);
	}
	return false;
}

// ---------------------------------------------------------------------------------------------
// Random helpers (host, scene construction only)

namespace
{
	struct HostStream { uint64_t key; uint32_t ctr; };
	thread_local HostStream t_hostStream = { 0x5EEDC0DEull, 0 };
	std::atomic<uint64_t> g_bvhBuildKey{ RT_RNG_DEFAULT_BVH_KEY };
}

extern "C" void RaylibB200_SeedHostRandom(uint64_t key) { t_hostStream.key = key; t_hostStream.ctr = 0; }

uint64_t RtGetBvhBuildKey() { return g_bvhBuildKey.load(); }
void RtSetBvhBuildKey(uint64_t key) { g_bvhBuildKey = key; }

RNG::RNG(uint32 nSamples) : key(rt_mix64(0xC0FFEEull + nSamples)), counter(0) {}
float RNG::Peek() { return rt_uniform(key, ++counter); }

float Random() { return rt_uniform(t_hostStream.key, ++t_hostStream.ctr); }

vec3 RandomInUnitSphere()
{
	const float u1 = Random();
	const float u2 = Random();
	const float z = 1.0f - 2.0f * u1;
	const float r = sqrtf(std::max(0.0f, 1.0f - z * z));
	const float phi = 2.0f * 3.141592f * u2;
	return vec3(r * cosf(phi), r * sinf(phi), z);
}

vec3 RandomInHemisphere(const vec3& axis)
{
	vec3 v = RandomInUnitSphere();
	if (dot(v, axis) < 0.0f) v = -v;
	return v;
}

vec3 RandomInUnitDisk()
{
	const float u1 = Random();
	const float u2 = Random();
	const float r = sqrtf(u1);
	const float theta = 2.0f * (float)M_PI * u2;
	return vec3(r * cosf(theta), r * sinf(theta), 0.0f);
}

vec3 RandomInCosineHemisphere()
{
	// concentric-disk mapping lifted to the hemisphere around +z
	float a = 2.0f * Random() - 1.0f;
	float b = 2.0f * Random() - 1.0f;
	if (a != 0.0f && b != 0.0f)
	{
		float radius, theta;
		if (std::abs(a) > std::abs(b)) { radius = a; theta = (float)M_PI_4 * (b / a); }
		else { radius = b; theta = (float)M_PI_2 - (float)M_PI_4 * (a / b); }
		a = radius * cosf(theta);
		b = radius * sinf(theta);
	}
	return vec3(a, b, sqrtf(std::max(0.0f, 1.0f - a * a - b * b)));
}
