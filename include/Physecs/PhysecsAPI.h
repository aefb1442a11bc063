// Drop-in header name of the reference (include/Physecs/PhysecsAPI.h).
#pragma once
#ifndef PHYSECS_API
#define PHYSECS_API __attribute__((visibility("default")))
#endif
