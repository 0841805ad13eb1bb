// Fixture simulator "balls": a walled pen with a few loose cubes and a volley of
// spheres -- exercises the sphere pair types of the narrowphase (sphere-sphere,
// sphere-plane and, through GJK, sphere-hull) that the box-world room fixture
// never produces.  Physics only: no agents, no lidar (the reference's
// BVH::traceRay asserts on sphere primitives, src/physics/broadphase.cpp:871).
// Compiled unchanged for the reference CPU backend (oracle/harness_balls.cpp)
// and, through NVRTC, for this engine.
#pragma once

#include <madrona/taskgraph_builder.hpp>
#include <madrona/custom_context.hpp>
#include <madrona/components.hpp>
#include <madrona/physics.hpp>
#include <madrona/rand.hpp>

namespace balls {

using madrona::Entity;
using madrona::CountT;
using madrona::base::Position;
using madrona::base::Rotation;
using madrona::base::Scale;
using madrona::base::ObjectID;
using madrona::phys::Velocity;
using madrona::phys::ResponseType;
using madrona::phys::ExternalForce;
using madrona::phys::ExternalTorque;

constexpr int32_t kNumWalls = 4;
#if defined(BALLS_MANY) && BALLS_MANY == 2
// build variant (GPU only): 145 bodies per world, beyond the engine's documented
// per-world body cap -- the step must report the overflow, not corrupt memory
constexpr int32_t kNumCubes = 10;
constexpr int32_t kNumBalls = 130;
constexpr float kPen = 12.f;
#elif defined(BALLS_MANY)
// build variant: 95 bodies per world (> 64: more than one word of the engine's
// candidate-search leaf masks), a bigger pen
constexpr int32_t kNumCubes = 10;
constexpr int32_t kNumBalls = 80;
constexpr float kPen = 9.f;            // pen interior: [-kPen, kPen]^2
#else
constexpr int32_t kNumCubes = 3;
constexpr int32_t kNumBalls = 8;
constexpr float kPen = 4.f;
#endif
constexpr int32_t kMaxBodies = 1 + kNumWalls + kNumCubes + kNumBalls;

enum class ExportID : uint32_t {
    BodyPos,
    BodyRot,
    BodyVel,
    BodyEntity,
    NumExports,
};

enum class TaskGraphID : uint32_t {
    Step,
    NumTaskGraphs,
};

// indices into the ObjectManager built by sims/objects.py:balls_objects()
enum class SimObject : uint32_t {
    Cube,
    Wall,
    Agent,
    Plane,
    Ball,
    NumObjects,
};

struct Body : public madrona::Archetype<madrona::phys::RigidBody> {};

struct Config {
    madrona::phys::ObjectManager *objMgr;
};

struct WorldInit {
    uint32_t seed;
};

class Engine;

struct Sim : public madrona::WorldBase {
    static void registerTypes(madrona::ECSRegistry &registry, const Config &cfg);
    static void setupTasks(madrona::TaskGraphManager &mgr, const Config &cfg);

    Sim(Engine &ctx, const Config &cfg, const WorldInit &init);

    madrona::RNG rng;
};

class Engine : public madrona::CustomContext<Engine, Sim> {
public:
    using CustomContext::CustomContext;
};

}
