// Runtime invariant checks (reference: core/assertion.h:10-19, core/assertion.cc:4-26).
// A failed check prints file/line and raises SIGTRAP when a debugger is expected
// (RAYLIB_B200_TRAP=1); otherwise it only prints, so FFI hosts are not torn down.
#pragma once
#include "raylib_types.h"

#define STATIC_ASSERT(x) static_assert(x)
#define CHECK_NO_ENTRY() CHECK(false);
#define CHECKF(x, msg)   CHECKF_IMPL(!!(x), msg, __FILE__, __LINE__)
#define CHECK(x)         CHECK_IMPL(!!(x), __FILE__, __LINE__)

// out of line on purpose (exported, C linkage): the macros above expand to one call
extern "C" RAYLIB_API void CHECKF_IMPL(int x, const char* msg, const char* file, int line);
extern "C" RAYLIB_API void CHECK_IMPL(int x, const char* file, int line);
