// Runtime invariant checks (reference: core/assertion.h:10-19, core/assertion.cc:4-26).
// A failed check prints file/line and raises SIGTRAP when a debugger is expected
// (RAYLIB_B200_TRAP=1); otherwise it only prints, so FFI hosts are not torn down.
#pragma once
#include "raylib_types.h"

extern "C" {
	RAYLIB_API void CHECK_IMPL(int x, const char* file, int line);
	RAYLIB_API void CHECKF_IMPL(int x, const char* msg, const char* file, int line);
}

#define CHECK(x)         CHECK_IMPL(!!(x), __FILE__, __LINE__)
#define CHECKF(x, msg)   CHECKF_IMPL(!!(x), msg, __FILE__, __LINE__)
#define CHECK_NO_ENTRY() CHECK(false);
#define STATIC_ASSERT(x) static_assert(x)
