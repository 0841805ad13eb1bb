// Type IDs are assigned in registration-call order (reference GPU:
// src/mw/device/include/madrona/type_tracker.hpp:17-30).  One device global
// per type in the JIT module.
#pragma once
#include <madrona/types.hpp>
namespace madrona {

template <typename T>
__device__ uint32_t mb2TypeIDStorage = 0xFFFFFFFFu;

template <typename T>
struct TypeIDHolder {
    static inline uint32_t &ref() { return mb2TypeIDStorage<T>; }
};

class TypeTracker {
public:
    static constexpr uint32_t unassignedTypeID = 0xFFFFFFFFu;

    template <typename T>
    static inline uint32_t typeID() { return mb2TypeIDStorage<T>; }

    template <typename T>
    static inline void registerType(uint32_t *next_id_ptr)
    {
        if (mb2TypeIDStorage<T> == unassignedTypeID) {
            mb2TypeIDStorage<T> = (*next_id_ptr)++;
        }
    }
};

}
