// Drop-in header name of the reference (include/Physecs/Joint.h); declarations live in detail/b200_joints.hpp.
#pragma once
#include "detail/b200_joints.hpp"
