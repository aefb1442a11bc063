// Drop-in header name of the reference (include/Physecs/Joints/PrismaticJoint.h).
#pragma once
#include "../detail/b200_joints.hpp"
