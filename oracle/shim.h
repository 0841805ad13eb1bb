/* TEST INFRASTRUCTURE ONLY. Force-included when compiling the reference CPU
 * backend with g++ (the reference expects clang+libc++):
 * src/physics/xpbd.cpp:57,62 call unqualified isnan(),
 * include/madrona/mesh_bvh.inl:785 unqualified signbit(). */
#include <cmath>
using std::isnan;
using std::signbit;
