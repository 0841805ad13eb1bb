// Fixture simulator 4 ("arena", BASELINE.json configs[2]: the Hide&Seek-class
// workload).  Per world: a ground plane, four border walls, a central room made
// of eight wall segments with two door gaps, two doors latched to the walls by
// FIXED joints until an agent unlatches them,
// two static hexagonal pillars, six force-driven agents (three hiders, three
// seekers; hexagonal-prism hulls: 12 vertices, 8 faces), twelve cubes, eight long
// boxes, three ramps (wedge hulls: 6 vertices, 5 faces) and two loose hexagonal
// barrels = 49 rigid bodies.  Hiders grab / release movable objects through
// FIXED joints, seekers shove them with one-step HINGE joints (action driven,
// see grabSystem); XPBD with 4 substeps at dt = 0.04; 16-ray lidar
// and agent-to-agent visibility rays through the broadphase tree; every movable
// object, both doors and all joints are destroyed and recreated on each episode
// reset (entity churn + compaction of two tables).
// The upstream Hide&Seek sources are not available here; this
// fixture defines the workload by construction.  No transcendental functions:
// yaw angles and directions come from literal tables, so with FP contraction
// off on both sides floats compare bit for bit against the reference CPU backend.
#pragma once

#include <madrona/taskgraph_builder.hpp>
#include <madrona/custom_context.hpp>
#include <madrona/components.hpp>
#include <madrona/physics.hpp>
#include <madrona/rand.hpp>

namespace arena {

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

constexpr int32_t kNumAgents = 6;
constexpr int32_t kNumHiders = 3;
constexpr int32_t kNumCubes = 12;
constexpr int32_t kNumLongBoxes = 8;
constexpr int32_t kNumRamps = 3;
constexpr int32_t kNumBarrels = 2;
constexpr int32_t kNumDoors = 2;
constexpr int32_t kNumMovable = kNumCubes + kNumLongBoxes + kNumRamps + kNumBarrels;
constexpr int32_t kNumBorderWalls = 4;
constexpr int32_t kNumRoomWalls = 8;
constexpr int32_t kNumPillars = 2;
constexpr int32_t kNumLidar = 16;
// plane + walls + pillars + doors + movable objects (the PhysicsEntity table)
constexpr int32_t kNumStaticBodies = 1 + kNumBorderWalls + kNumRoomWalls + kNumPillars;
constexpr int32_t kNumPhysicsEntities = kNumStaticBodies + kNumDoors + kNumMovable;
constexpr int32_t kMaxBodies = 64;

enum class ExportID : uint32_t {
    Reset,
    Action,
    Reward,
    Done,
    SelfObs,
    OtherObs,
    Lidar,
    AgentPos,
    AgentRot,
    BodyCount,
    JointCount,
    BodyPos,
    BodyRot,
    BodyEntity,
    BodyVel,
    NumExports,
};

enum class TaskGraphID : uint32_t {
    Step,
    NumTaskGraphs,
};

enum class SimObject : uint32_t {
    Cube,
    LongBox,
    Ramp,
    Barrel,
    Door,
    Wall,
    Pillar,
    Agent,
    Plane,
    NumObjects,
};

enum class EntityType : uint32_t {
    None,
    Cube,
    LongBox,
    Ramp,
    Barrel,
    Door,
    Wall,
    Pillar,
    Agent,
    Plane,
};

struct WorldReset { int32_t reset; };
struct BodyCount { int32_t count; };
struct JointCount { int32_t count; };

struct Action {
    int32_t moveAmount;   // [0, 3]
    int32_t moveAngle;    // [0, 7], multiples of 45 degrees in the agent frame
    int32_t rotate;       // [0, 4], 2 = none
    int32_t grab;         // 1 = toggle: grab the movable object in reach / let go
};

struct Reward { float v; };
struct Done { int32_t v; };
struct StepsRemaining { uint32_t t; };
struct Team { int32_t isHider; };

// the joint an agent holds (1: fixed) or shoves (2: hinge, one step) an object with
struct Grip {
    Entity joint;
    int32_t holding;
    int32_t cooldown;     // steps until a seeker may shove again
};

struct SelfObs {
    float x, y, z;
    float qw, qx, qy, qz;
    float isHider;
    float holding;
    float stepsRemaining;
};

struct OtherObs {
    // per other agent: offset in the arena frame, team, visible (ray test)
    float v[kNumAgents - 1][4];
};

struct LidarSample {
    float depth;
    float type;
};

struct Lidar {
    LidarSample samples[kNumLidar];
};

struct Agent : public madrona::Archetype<
    madrona::phys::RigidBody,
    Action, Reward, Done, StepsRemaining, Team, Grip, SelfObs, OtherObs, Lidar, EntityType
> {};

struct PhysicsEntity : public madrona::Archetype<
    madrona::phys::RigidBody,
    EntityType
> {};

struct Config {
    madrona::phys::ObjectManager *objMgr;
    uint32_t episodeLen;
    uint32_t pad;
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
    uint32_t episodeLen;
    uint32_t episode;

    Entity plane;
    Entity borders[kNumBorderWalls];
    Entity roomWalls[kNumRoomWalls];
    Entity pillars[kNumPillars];
    Entity agents[kNumAgents];
    // recreated on every episode reset
    Entity movable[kNumMovable];
    Entity doors[kNumDoors];
    Entity latches[kNumDoors];
    int32_t latched[kNumDoors];
    int32_t numJoints;
};

class Engine : public madrona::CustomContext<Engine, Sim> {
public:
    using CustomContext::CustomContext;
};

}
