// Constants shared by host material code; the BRDF terms themselves
// (Beckmann D, Smith G, Schlick F -- reference render/brdf.h:14-115) are evaluated
// on the device in csrc/device/rt_shade.cuh.
#pragma once
#include "core/vec3.h"

namespace BRDF
{
	const float PI = 3.14159265359f;
}
