/* TEST INFRASTRUCTURE ONLY. Force-included when compiling the reference CPU
 * backend with g++ (the reference expects clang+libc++):
 * src/physics/xpbd.cpp:57,62 call unqualified isnan(). */
#include <cmath>
using std::isnan;
