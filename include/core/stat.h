// Scoped wall-clock timer that reports through LOG (reference: core/stat.h:8-31).
#pragma once
#include "core/logger.h"
#include <chrono>

struct ScopedCycleCounter
{
	explicit ScopedCycleCounter(const char* inLabel)
		: label(inLabel), startTime(std::chrono::system_clock::now()) {}
	~ScopedCycleCounter()
	{
		const auto ms = std::chrono::duration_cast<std::chrono::milliseconds>(
			std::chrono::system_clock::now() - startTime).count();
		LOG("[STAT] %s: %u ms (%.3f s)", label, (unsigned)ms, (float)ms * 0.001f);
	}
private:
	const char* label;
	std::chrono::system_clock::time_point startTime;
};

#define SCOPED_CPU_COUNTER(custom_label) ScopedCycleCounter __scoped_cycle_counter(#custom_label);
