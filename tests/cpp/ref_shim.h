// g++ (not the reference's clang/libc++) needs the unqualified <cmath> names the reference headers use
#include <cmath>
using std::isnan;
using std::signbit;
