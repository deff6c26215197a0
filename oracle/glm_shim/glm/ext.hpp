// Part of the GLM stand-in (see glm.hpp).
#pragma once
#include "glm.hpp"
