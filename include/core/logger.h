// Host-side logging (reference: core/logger.h, core/logger.cc:22-96).
// LOG() is printf-style and thread-safe; messages are printed by a background
// thread once Raylib_Initialize() started it, synchronously before that.
#pragma once
#include "raylib_types.h"

RAYLIB_API void LOG(const char* format, ...);

// life cycle of the printing thread: started by Raylib_Initialize, drained by Raylib_FlushLogThread (blocks until every
// queued line is out), stopped -- after a last drain -- by Raylib_Terminate
namespace Logger { void StartLogThread(); RAYLIB_API void FlushLogThread(); void KillAndWaitForLogThread(); }
