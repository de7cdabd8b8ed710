// oracle/compat.h -- TEST INFRASTRUCTURE (not product code).
// Force-included (-include) when the UNMODIFIED reference sources under
// /root/reference/raylib are compiled with g++ on Linux.  It only papers over
// MSVC-isms; it changes no behaviour of the path-tracing code.
//   __declspec            raylib/raylib_types.h:7-11
//   sprintf_s/vsprintf_s  raylib/core/assertion.cc:9,12  raylib/core/logger.cc:68
//   __debugbreak          raylib/core/assertion.cc:12,24
//   unqualified isnan     raylib/render/texture.cc:41, raylib/core/vec3.h:168
#pragma once
#include <cstdio>
#include <cstdarg>
#include <cstring>
#include <cmath>
#include <csignal>
#include <string>
#include <map>
#include <memory>
#include <vector>

#define __declspec(x)
#define sprintf_s(buf, ...) snprintf(buf, sizeof(buf), __VA_ARGS__)
#define vsprintf_s(buf, n, fmt, ap) vsnprintf(buf, n, fmt, ap)
#define _countof(a) (sizeof(a) / sizeof((a)[0]))

// CHECK failures print and then "break"; count them instead of trapping so a
// failed runtime assert inside the reference shows up in the oracle report.
extern "C" void oracle_on_debugbreak(void);
#define __debugbreak() oracle_on_debugbreak()

using std::isnan;
using std::isinf;
