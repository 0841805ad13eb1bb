// Reference: include/madrona/ecs_flags.hpp:17-38.  Flags are accepted and
// recorded; storage policy here is decided by the engine (all columns are
// device memory owned by the executor).
#pragma once
#include <madrona/ecs.hpp>
#include <madrona/span.hpp>
namespace madrona {

enum class ArchetypeFlags : uint32_t {
    None = 0,
    ImportOffsets = 1_u32 << 0,
};

enum class ComponentFlags : uint32_t {
    None = 0,
    ExportMemory = 1_u32 << 0,
    ImportMemory = 1_u32 << 1,
    CudaReserveMemory = 1_u32 << 2,
    CudaAllocMemory = 1_u32 << 3,
};

template <typename... ComponentTs>
struct ComponentMetadataSelector {
    ComponentFlags flags[sizeof...(ComponentTs) == 0 ? 1 : sizeof...(ComponentTs)];

    inline ComponentMetadataSelector() : flags {} {}
    inline ComponentMetadataSelector(ComponentFlags f)
    {
        for (unsigned i = 0; i < sizeof...(ComponentTs); i++) flags[i] = f;
    }
    template <typename... FlagTs>
    inline ComponentMetadataSelector(FlagTs... in_flags) : flags { in_flags... } {}
};

inline ArchetypeFlags operator|(ArchetypeFlags a, ArchetypeFlags b) { return ArchetypeFlags(uint32_t(a) | uint32_t(b)); }
inline ArchetypeFlags &operator|=(ArchetypeFlags &a, ArchetypeFlags b) { a = a | b; return a; }
inline ArchetypeFlags operator&(ArchetypeFlags a, ArchetypeFlags b) { return ArchetypeFlags(uint32_t(a) & uint32_t(b)); }
inline ArchetypeFlags &operator&=(ArchetypeFlags &a, ArchetypeFlags b) { a = a & b; return a; }
inline ComponentFlags operator|(ComponentFlags a, ComponentFlags b) { return ComponentFlags(uint32_t(a) | uint32_t(b)); }
inline ComponentFlags &operator|=(ComponentFlags &a, ComponentFlags b) { a = a | b; return a; }
inline ComponentFlags operator&(ComponentFlags a, ComponentFlags b) { return ComponentFlags(uint32_t(a) & uint32_t(b)); }
inline ComponentFlags &operator&=(ComponentFlags &a, ComponentFlags b) { a = a & b; return a; }

}
