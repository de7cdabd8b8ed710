// Scoped wall-clock timer that reports through LOG (reference: core/stat.h:8-31).
#pragma once
#include <chrono>
#include "core/logger.h"

// SCOPED_CPU_COUNTER(Name) at the top of a scope logs "[STAT] Name: <ms> ms (<s> s)" when the scope ends.
struct ScopedCycleCounter
{
	explicit ScopedCycleCounter(const char* inLabel) : label(inLabel), startTime(std::chrono::system_clock::now()) {}
	~ScopedCycleCounter()
	{
		using namespace std::chrono;
		const long long ms = (long long)duration_cast<milliseconds>(system_clock::now() - startTime).count();
		LOG("[STAT] %s: %u ms (%.3f s)", label, (unsigned)ms, 0.001f * (float)ms);
	}
private:
	const char* label; std::chrono::system_clock::time_point startTime;
};
#define SCOPED_CPU_COUNTER(custom_label) ScopedCycleCounter __scoped_cycle_counter(#custom_label);
