// Drop-in header name of the reference (include/Physecs/Joints/FixedJoint.h).
#pragma once
#include "../detail/b200_joints.hpp"
