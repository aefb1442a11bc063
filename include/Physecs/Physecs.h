// Drop-in header name of the reference (include/Physecs/Physecs.h); declarations live in detail/b200_scene.hpp.
#pragma once
#include "detail/b200_scene.hpp"
