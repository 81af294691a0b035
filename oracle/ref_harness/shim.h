// libstdc++ 13 has no std::modff / std::atan2f, which the reference uses (Loader.h:86-99,
// Global.h:83,87). Force-included with -include so the reference headers stay untouched.
#pragma once
#include <cmath>
#include <math.h>
namespace std { using ::modff; using ::atan2f; }
