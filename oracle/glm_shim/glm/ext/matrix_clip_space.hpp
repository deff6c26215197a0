// Part of the GLM stand-in (see ../glm.hpp). The reference includes this header
// (src/vrt/rt.cpp:2) but uses nothing from it.
#pragma once
#include "../glm.hpp"
