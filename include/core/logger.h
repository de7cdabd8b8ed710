// Host-side logging (reference: core/logger.h, core/logger.cc:22-96).
// LOG() is printf-style and thread-safe; messages are printed by a background
// thread once Raylib_Initialize() started it, synchronously before that.
#pragma once
#include "raylib_types.h"

namespace Logger
{
	void StartLogThread();
	RAYLIB_API void FlushLogThread();
	void KillAndWaitForLogThread();
}

RAYLIB_API void LOG(const char* format, ...);
