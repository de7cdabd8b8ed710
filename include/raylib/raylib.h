// Same header reachable as "raylib/raylib.h" (how src/main.cc:12 includes it).
#pragma once
#include "../raylib.h"
