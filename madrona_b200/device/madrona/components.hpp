// Reference: include/madrona/components.hpp:9-44, src/core/base.cpp.
#pragma once
#include <madrona/math.hpp>
#include <madrona/fwd.hpp>
#include <madrona/taskgraph.hpp>
namespace madrona {
namespace base {

struct Position : math::Vector3 {
    inline Position(math::Vector3 v) : Vector3(v) {}
};

struct Rotation : math::Quat {
    inline Rotation(math::Quat q) : Quat(q) {}
};

struct Scale : math::Diag3x3 {
    inline Scale(math::Diag3x3 d) : Diag3x3(d) {}
};

struct ObjectID {
    int32_t idx;
};

struct ObjectInstance : Bundle<Position, Rotation, Scale, ObjectID> {};

inline void registerTypes(ECSRegistry &registry)
{
    registry.registerComponent<Position>();
    registry.registerComponent<Rotation>();
    registry.registerComponent<Scale>();
    registry.registerComponent<ObjectID>();
    registry.registerBundle<ObjectInstance>();
}

}
}
