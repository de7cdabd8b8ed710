#pragma once
#include "render/texture.h"
