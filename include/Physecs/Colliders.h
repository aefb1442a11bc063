// Drop-in header name of the reference (include/Physecs/Colliders.h); the declarations live in detail/b200_types.hpp.
#pragma once
#include "detail/b200_types.hpp"
