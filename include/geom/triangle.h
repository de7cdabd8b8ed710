#pragma once
#include "geom/primitives.h"
