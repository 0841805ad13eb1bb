// Reference: include/madrona/ecs.hpp:17-85, ecs.inl.
#pragma once
#include <madrona/fwd.hpp>
#include <madrona/types.hpp>
#include <cstdint>
namespace madrona {

struct Entity {
    uint32_t gen;
    int32_t id;
    MB2_HD static constexpr inline Entity none() { return Entity { 0xFFFFFFFFu, (int32_t)0xFFFFFFFF }; }
};

struct Loc {
    uint32_t archetype;
    int32_t row;
    MB2_HD inline bool valid() const { return archetype != 0xFFFFFFFFu; }
    MB2_HD static inline Loc none() { return Loc { 0xFFFFFFFFu, 0 }; }
};

struct IndexHelper {
    uint32_t prev;
    uint32_t next;
};

template <typename... ComponentTs> struct Bundle {
    using Base = Bundle<ComponentTs...>;
};

template <typename... ComponentTs> struct Archetype {
    using Base = Archetype<ComponentTs...>;
};

struct WorldID {
    int32_t idx;
};

struct ComponentID { uint32_t id; };
struct ArchetypeID { uint32_t id; };

class WorldBase {
public:
    MB2_HD inline WorldBase(Context &) {}
    WorldBase(const WorldBase &) = delete;
};

MB2_HD inline bool operator==(Entity a, Entity b) { return a.gen == b.gen && a.id == b.id; }
MB2_HD inline bool operator!=(Entity a, Entity b) { return !(a == b); }
MB2_HD inline bool operator==(Loc a, Loc b) { return a.archetype == b.archetype && a.row == b.row; }
MB2_HD inline bool operator!=(Loc a, Loc b) { return !(a == b); }

}
