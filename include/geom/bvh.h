#pragma once
#include "geom/hit.h"
