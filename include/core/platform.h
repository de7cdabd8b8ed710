// PLATFORM_WINDOWS is 1 on Win32/Win64 toolchains and 0 everywhere else (reference: core/platform.h).
#pragma once
#ifdef PLATFORM_WINDOWS
  #undef PLATFORM_WINDOWS
#endif
#if defined(_WIN64) || defined(_WIN32)
  #define PLATFORM_WINDOWS 1
#else
  #define PLATFORM_WINDOWS 0
#endif
