// Counter-based RNG: threefry2x32, 20 rounds.  Must be bit-exact with the
// reference (include/madrona/rand.inl:31-277; KATs tests/rand.cpp:131-141):
// pure 32-bit integer arithmetic, so it is.
#pragma once
#include <madrona/macros.hpp>
#include <madrona/types.hpp>
#include <madrona/math.hpp>
#include <madrona/utils.hpp>

namespace madrona {

struct RandKey {
    uint32_t a;
    uint32_t b;
};

namespace rand {

MB2_HD constexpr inline uint32_t mb2Rotl(uint32_t v, uint32_t d)
{
    return (v << d) | (v >> (32 - d));
}

// One application of the threefry2x32-20 block function keyed by `src` to
// the counter (idx, idx_upper).
MB2_HD constexpr inline RandKey split_i(RandKey src, uint32_t idx,
                                        uint32_t idx_upper = 0)
{
    const uint32_t rot[8] = { 13, 15, 26, 6, 17, 29, 16, 24 };
    uint32_t ks[3] = { src.a, src.b, 0x1BD11BDAu ^ src.a ^ src.b };
    uint32_t x0 = idx + ks[0];
    uint32_t x1 = idx_upper + ks[1];

    // five groups of four rounds; after group g inject ks[(g+1)%3], ks[(g+2)%3]+g+1
    for (uint32_t g = 0; g < 5; g++) {
        const uint32_t *r = (g & 1) ? rot + 4 : rot;
        for (int i = 0; i < 4; i++) {
            x0 += x1;
            x1 = mb2Rotl(x1, r[i]);
            x1 ^= x0;
        }
        x0 += ks[(g + 1) % 3];
        x1 += ks[(g + 2) % 3] + (g + 1);
    }
    return RandKey { x0, x1 };
}

MB2_HD constexpr inline RandKey initKey(uint32_t seed, uint32_t seed_upper = 0)
{
    return split_i(RandKey { seed, seed_upper }, 0);
}

MB2_HD constexpr inline uint32_t bits32(RandKey k) { return k.a ^ k.b; }
MB2_HD constexpr inline uint64_t bits64(RandKey k)
{
    return ((uint64_t)k.b << 32) | (uint64_t)k.a;
}

// [0,1) with a 24-bit mantissa (rand.inl:198-221).
MB2_HD constexpr inline float bitsToFloat01(uint32_t rand_bits)
{
    return (rand_bits >> 8) * 0x1p-24f;
}

// Lemire's unbiased bounded integer (rand.inl:110-160), rejection re-keys
// with split_i(k, 0).
MB2_HD constexpr inline int32_t sampleI32(RandKey k, int32_t a, int32_t b)
{
    uint32_t s = (uint32_t)(b - a);
    uint64_t m = (uint64_t)bits32(k) * (uint64_t)s;
    uint32_t l = (uint32_t)m;
    if (l < s) {
        uint32_t t = (0u - s) % s;
        while (l < t) {
            k = split_i(k, 0);
            m = (uint64_t)bits32(k) * (uint64_t)s;
            l = (uint32_t)m;
        }
    }
    return (int32_t)(uint32_t)(m >> 32) + a;
}

MB2_HD constexpr inline int32_t sampleI32Biased(RandKey k, int32_t a, int32_t b)
{
    // NB: the reference does not add `a` here (rand.inl:162-168).
    return utils::u32mulhi(bits32(k), (uint32_t)(b - a));
}

MB2_HD constexpr inline float sampleUniform(RandKey k) { return bitsToFloat01(bits32(k)); }

MB2_HD constexpr inline bool sampleBool(RandKey k)
{
    return (MB2_POPC(bits32(k)) & 1) == 0;
}

MB2_HD constexpr inline math::Vector2 sample2xUniform(RandKey k)
{
    return math::Vector2 { bitsToFloat01(k.a), bitsToFloat01(k.b) };
}

}

class RNG {
public:
    MB2_HD inline RNG() : k_(RandKey { 0, 0 }), count_(0) {}
    MB2_HD inline RNG(RandKey k) : k_(k), count_(0) {}
    MB2_HD inline RNG(uint32_t seed) : RNG(rand::initKey(seed)) {}

    MB2_HD inline int32_t sampleI32(int32_t a, int32_t b) { return rand::sampleI32(advance(), a, b); }
    MB2_HD inline int32_t sampleI32Biased(int32_t a, int32_t b) { return rand::sampleI32Biased(advance(), a, b); }
    MB2_HD inline float sampleUniform() { return rand::sampleUniform(advance()); }
    MB2_HD inline bool sampleBool() { return rand::sampleBool(advance()); }
    MB2_HD inline RandKey randKey() { return advance(); }

    RNG(const RNG &) = default;
    RNG(RNG &&) = default;
    RNG &operator=(const RNG &) = default;
    RNG &operator=(RNG &&) = default;

private:
    MB2_HD inline RandKey advance()
    {
        RandKey s = rand::split_i(k_, count_);
        count_ += 1;
        return s;
    }

    RandKey k_;
    uint32_t count_;
};

}
