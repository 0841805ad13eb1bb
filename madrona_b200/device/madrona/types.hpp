// Reference: include/madrona/types.hpp:13-53 (CountT is int32_t on the GPU).
#pragma once
#include <cstdint>
#include <madrona/macros.hpp>
namespace madrona {
using u64 = uint64_t;
using i64 = int64_t;
using u32 = uint32_t;
using i32 = int32_t;
using u16 = uint16_t;
using i16 = int16_t;
using u8 = uint8_t;
using i8 = int8_t;
using f32 = float;
MB2_HD inline constexpr u32 operator "" _u32(unsigned long long v) { return uint32_t(v); }
MB2_HD inline constexpr u64 operator "" _u64(unsigned long long v) { return uint64_t(v); }
MB2_HD inline constexpr i32 operator "" _i32(unsigned long long v) { return int32_t(v); }
MB2_HD inline constexpr i64 operator "" _i64(unsigned long long v) { return int64_t(v); }
#ifdef MADRONA_GPU_MODE
using CountT = int32_t;
#else
using CountT = int64_t;
#endif
template <typename T> concept EnumType = __is_enum(T);
}
