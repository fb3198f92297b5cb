// global_includes.h -- reference src/global_includes.h:29-38
#pragma once
#include "logger.h"
#define APP_EXIT_PAUSE 0
inline float reflection2Admitance(float coef) { return (1.f - coef) / (1.f + coef); }
inline float admitance2Reflection(float coef) { return reflection2Admitance(coef); }
