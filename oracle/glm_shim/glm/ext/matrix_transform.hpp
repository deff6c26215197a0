// Part of the GLM stand-in (see ../glm.hpp). translate/rotate/lookAt live in glm.hpp.
#pragma once
#include "../glm.hpp"
