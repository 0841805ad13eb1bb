// Reference: include/madrona/utils.hpp (subset used by simulators).
#pragma once
#include <madrona/types.hpp>
namespace madrona {
namespace utils {
template <typename T>
MB2_HD constexpr inline T divideRoundUp(T a, T b) { return (a + (b - 1)) / b; }
template <typename T>
MB2_HD constexpr inline T roundUp(T v, T mult) { return divideRoundUp(v, mult) * mult; }
MB2_HD constexpr inline uint64_t roundUpPow2(uint64_t v, uint64_t p) { return (v + p - 1) & ~(p - 1); }
MB2_HD constexpr inline bool isPower2(uint64_t v) { return v && !(v & (v - 1)); }
MB2_HD constexpr inline uint32_t u32mulhi(uint32_t a, uint32_t b)
{
    return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
}
MB2_HD constexpr inline uint32_t int32NextPow2(uint32_t v)
{
    v--; v |= v >> 1; v |= v >> 2; v |= v >> 4; v |= v >> 8; v |= v >> 16;
    return v + 1;
}
MB2_HD constexpr inline uint32_t int32Log2(uint32_t v) { return 31u - (uint32_t)MB2_CLZ(v); }
template <typename T> MB2_HD constexpr inline T clamp(T v, T lo, T hi) { return v < lo ? lo : (v > hi ? hi : v); }
}
}
