#include "sim.hpp"

#ifdef MADRONA_GPU_MODE
#include <madrona/mw_gpu_entry.hpp>
#endif

using namespace madrona;
using namespace madrona::math;
using namespace madrona::phys;

namespace arena {

constexpr float kDeltaT = 0.04f;
constexpr CountT kNumSubsteps = 4;
constexpr float kHalf = 12.f;          // arena: x, y in [-12, 12]
constexpr float kWallThick = 0.5f;
constexpr float kWallHeight = 2.5f;
constexpr float kRoomHalf = 4.f;       // central room: wall centre lines at +-4
constexpr float kDoorGap = 2.5f;
constexpr float kDoorLen = 2.3f;
constexpr float kDoorThick = 0.3f;
constexpr float kDoorHeight = 2.0f;
constexpr float kDoorZ = 1.05f;        // hangs 5 cm above the floor
constexpr int32_t kNumSlots = 36;      // 6 x 6 placement grid, 4 m pitch

void Sim::registerTypes(ECSRegistry &registry, const Config &)
{
    base::registerTypes(registry);
    PhysicsSystem::registerTypes(registry);

    registry.registerComponent<Action>();
    registry.registerComponent<Reward>();
    registry.registerComponent<Done>();
    registry.registerComponent<StepsRemaining>();
    registry.registerComponent<Team>();
    registry.registerComponent<Grip>();
    registry.registerComponent<SelfObs>();
    registry.registerComponent<OtherObs>();
    registry.registerComponent<Lidar>();
    registry.registerComponent<EntityType>();

    registry.registerSingleton<WorldReset>();
    registry.registerSingleton<BodyCount>();
    registry.registerSingleton<JointCount>();

    registry.registerArchetype<Agent>(
        ComponentMetadataSelector<> {}, ArchetypeFlags::None, kNumAgents);
    registry.registerArchetype<PhysicsEntity>();

    registry.exportSingleton<WorldReset>((uint32_t)ExportID::Reset);
    registry.exportColumn<Agent, Action>((uint32_t)ExportID::Action);
    registry.exportColumn<Agent, Reward>((uint32_t)ExportID::Reward);
    registry.exportColumn<Agent, Done>((uint32_t)ExportID::Done);
    registry.exportColumn<Agent, SelfObs>((uint32_t)ExportID::SelfObs);
    registry.exportColumn<Agent, OtherObs>((uint32_t)ExportID::OtherObs);
    registry.exportColumn<Agent, Lidar>((uint32_t)ExportID::Lidar);
    registry.exportColumn<Agent, Position>((uint32_t)ExportID::AgentPos);
    registry.exportColumn<Agent, Rotation>((uint32_t)ExportID::AgentRot);
    registry.exportSingleton<BodyCount>((uint32_t)ExportID::BodyCount);
    registry.exportSingleton<JointCount>((uint32_t)ExportID::JointCount);
    registry.exportColumn<PhysicsEntity, Position>((uint32_t)ExportID::BodyPos);
    registry.exportColumn<PhysicsEntity, Rotation>((uint32_t)ExportID::BodyRot);
    registry.exportColumn<PhysicsEntity, Entity>((uint32_t)ExportID::BodyEntity);
    registry.exportColumn<PhysicsEntity, Velocity>((uint32_t)ExportID::BodyVel);
}

// yaw by multiples of 45 degrees: (cos(a/2), 0, 0, sin(a/2)) as literals
static inline Quat yawQuat(int32_t eighth)
{
    constexpr float c = 0.92387953f, s = 0.38268343f, d = 0.70710678f;
    switch (eighth & 7) {
    case 0: return Quat { 1, 0, 0, 0 };
    case 1: return Quat { c, 0, 0, s };
    case 2: return Quat { d, 0, 0, d };
    case 3: return Quat { s, 0, 0, c };
    case 4: return Quat { 0, 0, 0, 1 };
    case 5: return Quat { -s, 0, 0, c };
    case 6: return Quat { -d, 0, 0, d };
    default: return Quat { -c, 0, 0, s };
    }
}

static inline void setupBody(Engine &ctx, Entity e, Vector3 pos, Quat rot,
                             Diag3x3 scale, SimObject obj, ResponseType resp,
                             EntityType type)
{
    ObjectID obj_id { (int32_t)obj };
    ctx.get<Position>(e) = pos;
    ctx.get<Rotation>(e) = rot;
    ctx.get<Scale>(e) = scale;
    ctx.get<ObjectID>(e) = obj_id;
    ctx.get<ResponseType>(e) = resp;
    ctx.get<Velocity>(e) = Velocity { Vector3::zero(), Vector3::zero() };
    ctx.get<ExternalForce>(e) = Vector3::zero();
    ctx.get<ExternalTorque>(e) = Vector3::zero();
    ctx.get<EntityType>(e) = type;
    ctx.get<broadphase::LeafID>(e) = PhysicsSystem::registerEntity(ctx, e, obj_id);
}

static inline void placeWall(Engine &ctx, Entity e, float x0, float x1,
                             float y0, float y1)
{
    Vector3 pos { 0.5f * (x0 + x1), 0.5f * (y0 + y1), 0.5f * kWallHeight };
    Diag3x3 scale { x1 - x0, y1 - y0, kWallHeight };
    setupBody(ctx, e, pos, Quat { 1, 0, 0, 0 }, scale, SimObject::Wall,
              ResponseType::Static, EntityType::Wall);
}

static inline Vector3 slotPos(int32_t slot, RNG &rng, float z)
{
    int32_t col = slot % 6, row = slot / 6;
    float x = -10.f + 4.f * (float)col + (rng.sampleUniform() - 0.5f) * 0.6f;
    float y = -10.f + 4.f * (float)row + (rng.sampleUniform() - 0.5f) * 0.6f;
    return Vector3 { x, y, z };
}

// (Re)generate the layout of this world.  Persistent entities (plane, walls,
// pillars, agents) are re-placed and re-registered with the broadphase; doors,
// hinges and movable objects are created fresh (the previous ones were
// destroyed by the caller).
static void generateWorld(Engine &ctx)
{
    Sim &sim = ctx.data();
    RNG &rng = sim.rng;

    PhysicsSystem::reset(ctx);

    setupBody(ctx, sim.plane, Vector3 { 0, 0, 0 }, Quat { 1, 0, 0, 0 },
              Diag3x3 { 1, 1, 1 }, SimObject::Plane, ResponseType::Static,
              EntityType::Plane);

    const float hw = kHalf, t = kWallThick, ht = 0.5f * kWallThick;
    placeWall(ctx, sim.borders[0], -hw - t, hw + t, -hw - t, -hw);
    placeWall(ctx, sim.borders[1], -hw - t, hw + t, hw, hw + t);
    placeWall(ctx, sim.borders[2], -hw - t, -hw, -hw, hw);
    placeWall(ctx, sim.borders[3], hw, hw + t, -hw, hw);

    const float r = kRoomHalf, g = 0.5f * kDoorGap;
    placeWall(ctx, sim.roomWalls[0], -r - ht, r + ht, r - ht, r + ht);       // +y side
    placeWall(ctx, sim.roomWalls[1], -r - ht, -r + ht, -r + ht, r - ht);     // -x side
    placeWall(ctx, sim.roomWalls[2], r - ht, r + ht, -r + ht, -g);           // +x side, below the gap
    placeWall(ctx, sim.roomWalls[3], r - ht, r + ht, g, r - ht);             // +x side, above the gap
    placeWall(ctx, sim.roomWalls[4], -r - ht, -g, -r - ht, -r + ht);         // -y side, left of the gap
    placeWall(ctx, sim.roomWalls[5], g, r + ht, -r - ht, -r + ht);           // -y side, right of the gap
    placeWall(ctx, sim.roomWalls[6], -11.5f, -6.5f, 8.f - ht, 8.f + ht);     // free-standing partitions
    placeWall(ctx, sim.roomWalls[7], 8.f - ht, 8.f + ht, -11.5f, -6.5f);

    // every object that needs floor space takes one cell of the 6 x 6 grid:
    // slot(i) = (offset + i * stride) mod 36 with stride coprime to 36
    const int32_t strides[8] = { 5, 7, 11, 13, 17, 19, 23, 25 };
    const int32_t offset = rng.sampleI32(0, kNumSlots);
    const int32_t stride = strides[rng.sampleI32(0, 8)];
    int32_t next = 0;
    auto takeSlot = [&]() {
        int32_t s = (offset + next * stride) % kNumSlots;
        next += 1;
        return s;
    };

    for (int32_t i = 0; i < kNumPillars; i++) {
        Vector3 p = slotPos(takeSlot(), rng, 1.25f);
        setupBody(ctx, sim.pillars[i], p, yawQuat(rng.sampleI32(0, 8)),
                  Diag3x3 { 1.5f, 1.5f, 2.5f }, SimObject::Pillar,
                  ResponseType::Static, EntityType::Pillar);
    }

    // doors: slabs standing in the two gaps of the central room, each latched to
    // the wall segment next to it by a FIXED joint (static body <-> dynamic body)
    // until an agent unlatches it
    {
        const float edge_y = -g + 0.05f;
        Entity door = ctx.makeEntity<PhysicsEntity>();
        sim.doors[0] = door;
        Vector3 door_pos { r, edge_y + 0.5f * kDoorLen, kDoorZ };
        setupBody(ctx, door, door_pos, Quat { 1, 0, 0, 0 },
                  Diag3x3 { kDoorThick, kDoorLen, kDoorHeight },
                  SimObject::Door, ResponseType::Dynamic, EntityType::Door);
        Vector3 wall_pos = ctx.get<Position>(sim.roomWalls[2]);
        sim.latches[0] = PhysicsSystem::makeFixedJoint(ctx, sim.roomWalls[2], door,
            Quat { 1, 0, 0, 0 }, Quat { 1, 0, 0, 0 },
            Vector3 { r, edge_y, kDoorZ } - wall_pos,
            Vector3 { 0.f, -0.5f * kDoorLen, 0.f }, 0.f);
    }
    {
        const float edge_x = -g + 0.05f;
        Entity door = ctx.makeEntity<PhysicsEntity>();
        sim.doors[1] = door;
        Vector3 door_pos { edge_x + 0.5f * kDoorLen, -r, kDoorZ };
        setupBody(ctx, door, door_pos, Quat { 1, 0, 0, 0 },
                  Diag3x3 { kDoorLen, kDoorThick, kDoorHeight },
                  SimObject::Door, ResponseType::Dynamic, EntityType::Door);
        Vector3 wall_pos = ctx.get<Position>(sim.roomWalls[4]);
        sim.latches[1] = PhysicsSystem::makeFixedJoint(ctx, sim.roomWalls[4], door,
            Quat { 1, 0, 0, 0 }, Quat { 1, 0, 0, 0 },
            Vector3 { edge_x, -r, kDoorZ } - wall_pos,
            Vector3 { -0.5f * kDoorLen, 0.f, 0.f }, 0.f);
    }
    sim.latched[0] = 1;
    sim.latched[1] = 1;
    sim.numJoints = kNumDoors;

    int32_t m = 0;
    for (int32_t i = 0; i < kNumCubes; i++) {
        // one in four starts in the air
        float z = 0.75f + (rng.sampleI32(0, 4) == 0 ? 1.5f : 0.f);
        Vector3 p = slotPos(takeSlot(), rng, z);
        Entity e = ctx.makeEntity<PhysicsEntity>();
        sim.movable[m++] = e;
        setupBody(ctx, e, p, yawQuat(rng.sampleI32(0, 8)), Diag3x3 { 1.5f, 1.5f, 1.5f },
                  SimObject::Cube, ResponseType::Dynamic, EntityType::Cube);
    }
    for (int32_t i = 0; i < kNumLongBoxes; i++) {
        Vector3 p = slotPos(takeSlot(), rng, 0.5f);
        Entity e = ctx.makeEntity<PhysicsEntity>();
        sim.movable[m++] = e;
        setupBody(ctx, e, p, yawQuat(rng.sampleI32(0, 8)), Diag3x3 { 2.4f, 0.8f, 1.0f },
                  SimObject::LongBox, ResponseType::Dynamic, EntityType::LongBox);
    }
    for (int32_t i = 0; i < kNumRamps; i++) {
        // wedge mesh: z spans [-1/3, 2/3] of its height around the centroid
        Vector3 p = slotPos(takeSlot(), rng, 0.5f);
        Entity e = ctx.makeEntity<PhysicsEntity>();
        sim.movable[m++] = e;
        setupBody(ctx, e, p, yawQuat(rng.sampleI32(0, 8)), Diag3x3 { 2.1f, 2.0f, 1.5f },
                  SimObject::Ramp, ResponseType::Dynamic, EntityType::Ramp);
    }
    for (int32_t i = 0; i < kNumBarrels; i++) {
        Vector3 p = slotPos(takeSlot(), rng, 0.6f + 1.f * (float)i);
        Entity e = ctx.makeEntity<PhysicsEntity>();
        sim.movable[m++] = e;
        setupBody(ctx, e, p, yawQuat(rng.sampleI32(0, 8)), Diag3x3 { 1.2f, 1.2f, 1.2f },
                  SimObject::Barrel, ResponseType::Dynamic, EntityType::Barrel);
    }

    for (int32_t i = 0; i < kNumAgents; i++) {
        Entity agent = sim.agents[i];
        Vector3 p = slotPos(takeSlot(), rng, 0.75f);
        setupBody(ctx, agent, p, yawQuat(rng.sampleI32(0, 8)), Diag3x3 { 1.f, 1.f, 1.5f },
                  SimObject::Agent, ResponseType::Dynamic, EntityType::Agent);
        ctx.get<StepsRemaining>(agent).t = sim.episodeLen;
        ctx.get<Grip>(agent) = Grip { Entity::none(), 0, 0 };
    }
    sim.episode += 1;
}

// 45-degree steps: literal constants, no trigonometry at run time
static inline Vector3 moveDir(int32_t angle)
{
    constexpr float d = 0.70710678f;
    switch (angle & 7) {
    case 0: return Vector3 { 0, 1, 0 };
    case 1: return Vector3 { d, d, 0 };
    case 2: return Vector3 { 1, 0, 0 };
    case 3: return Vector3 { d, -d, 0 };
    case 4: return Vector3 { 0, -1, 0 };
    case 5: return Vector3 { -d, -d, 0 };
    case 6: return Vector3 { -1, 0, 0 };
    default: return Vector3 { -d, d, 0 };
    }
}

inline void movementSystem(Engine &, Action &action, Rotation &rot,
                           ExternalForce &force, ExternalTorque &torque)
{
    constexpr float move_max = 4000.f;
    constexpr float turn_max = 320.f;
    float f = move_max * (float)action.moveAmount * (1.f / 3.f);
    Vector3 dir = moveDir(action.moveAngle);
    Quat q = rot;
    force = q.rotateVec(Vector3 { f * dir.x, f * dir.y, 0.f });
    float t_z = turn_max * ((float)action.rotate - 2.f) * 0.5f;
    torque = Vector3 { 0.f, 0.f, t_z };
}

inline void agentZeroVelSystem(Engine &, Velocity &vel, Action &)
{
    vel.linear.x = 0.f;
    vel.linear.y = 0.f;
    vel.linear.z = fminf(vel.linear.z, 0.f);
    vel.angular = Vector3::zero();
}

// One invocation per world: agents in index order, so joints are created and
// destroyed in the same order on every backend.  A grab action (1) first
// unlatches a latched door in reach; otherwise hiders hold the movable object in
// front of them with a FIXED joint (toggle), seekers shove it with a HINGE joint
// that lives for exactly one step.  (The reference's hinge constraint is
// anti-restoring -- src/physics/xpbd.cpp:686-696 feeds p2 - p1 and a1 x a2 into
// updates that apply the negative magnitude to body 1, the opposite sign of its
// Fixed branch :650-685 -- so a hinge that stays alive diverges to NaN within a
// few steps on the reference itself; alive for one step it is a finite,
// deterministic push that exercises exactly that code path.)
inline void grabSystem(Engine &ctx, JointCount &)
{
    Sim &sim = ctx.data();
    for (int32_t a = 0; a < kNumAgents; a++) {
        Entity agent = sim.agents[a];
        Grip &grip = ctx.get<Grip>(agent);
        if (grip.holding == 2) {
            ctx.destroyEntity(grip.joint);
            grip.joint = Entity::none();
            grip.holding = 0;
            sim.numJoints -= 1;
        }
        if (grip.cooldown > 0) {
            grip.cooldown -= 1;
            continue;
        }
        if (ctx.get<Action>(agent).grab != 1) {
            continue;
        }
        if (grip.holding != 0) {
            ctx.destroyEntity(grip.joint);
            grip.joint = Entity::none();
            grip.holding = 0;
            sim.numJoints -= 1;
            continue;
        }
        Vector3 agent_pos = ctx.get<Position>(agent);
        bool unlatched = false;
        for (int32_t d = 0; d < kNumDoors; d++) {
            if (sim.latched[d] != 0 &&
                    ctx.get<Position>(sim.doors[d]).distance2(agent_pos) < 9.f) {
                ctx.destroyEntity(sim.latches[d]);
                sim.latched[d] = 0;
                sim.numJoints -= 1;
                unlatched = true;
                break;
            }
        }
        if (unlatched) {
            continue;
        }
        Quat agent_rot = ctx.get<Rotation>(agent);
        Vector3 hold_point = agent_pos + agent_rot.rotateVec(Vector3 { 0.f, 1.5f, 0.f });
        for (int32_t i = 0; i < kNumMovable; i++) {
            Entity obj = sim.movable[i];
            Vector3 obj_pos = ctx.get<Position>(obj);
            if (obj_pos.distance2(hold_point) >= 4.f) {
                continue;
            }
            // seekers leave the light barrels (the last kNumBarrels movables) alone
            if (ctx.get<Team>(agent).isHider == 0 && i >= kNumMovable - kNumBarrels) {
                continue;
            }
            // anchor = the object's centre, expressed in the agent's frame
            Vector3 r1 = agent_rot.inv().rotateVec(obj_pos - agent_pos);
            if (ctx.get<Team>(agent).isHider != 0) {
                // keep the current relative pose
                Quat obj_rot = ctx.get<Rotation>(obj);
                Quat attach2 = (obj_rot.inv() * agent_rot).normalize();
                grip.joint = PhysicsSystem::makeFixedJoint(ctx, agent, obj,
                    Quat { 1, 0, 0, 0 }, attach2, r1, Vector3 { 0.f, 0.f, 0.f }, 0.f);
                grip.holding = 1;
            } else {
                // the agent's up axis in the object's frame, tilted by 0.5 mrad, and
                // 0.2 mm of anchor offset: small non-zero angular and positional
                // hinge errors to start from
                Quat obj_rot = ctx.get<Rotation>(obj);
                Vector3 axis2 = obj_rot.inv().rotateVec(agent_rot.rotateVec(math::up));
                axis2.x += 0.0005f;
                grip.joint = PhysicsSystem::makeHingeJoint(ctx, agent, obj,
                    math::up, axis2, math::right, math::right,
                    r1, Vector3 { 0.f, 0.0002f, 0.f });
                grip.holding = 2;
                grip.cooldown = 6;
            }
            sim.numJoints += 1;
            break;
        }
    }
}

// hiders are rewarded for keeping every seeker at a distance, seekers for closing in
inline void rewardSystem(Engine &ctx, Position &pos, Team &team, Reward &reward,
                         StepsRemaining &steps, Done &done)
{
    Sim &sim = ctx.data();
    Vector3 p = pos;
    float r;
    if (team.isHider != 0) {
        r = 1.f;
        for (int32_t i = kNumHiders; i < kNumAgents; i++) {
            Vector3 o = ctx.get<Position>(sim.agents[i]);
            if (o.distance2(p) < 36.f) {
                r = -1.f;
            }
        }
    } else {
        r = -1.f;
        for (int32_t i = 0; i < kNumHiders; i++) {
            Vector3 o = ctx.get<Position>(sim.agents[i]);
            if (o.distance2(p) < 36.f) {
                r = 1.f;
            }
        }
    }
    reward.v = r;
    steps.t -= 1;
    done.v = steps.t == 0 ? 1 : 0;
}

inline void resetSystem(Engine &ctx, WorldReset &reset)
{
    Sim &sim = ctx.data();
    bool should_reset = reset.reset != 0;
    for (int32_t i = 0; i < kNumAgents; i++) {
        if (ctx.get<Done>(sim.agents[i]).v != 0) {
            should_reset = true;
        }
    }
    if (should_reset) {
        reset.reset = 0;
        for (int32_t i = 0; i < kNumAgents; i++) {
            Grip &grip = ctx.get<Grip>(sim.agents[i]);
            if (grip.holding != 0) {
                ctx.destroyEntity(grip.joint);
                grip.holding = 0;
            }
        }
        for (int32_t i = 0; i < kNumDoors; i++) {
            if (sim.latched[i] != 0) {
                ctx.destroyEntity(sim.latches[i]);
            }
            ctx.destroyEntity(sim.doors[i]);
        }
        for (int32_t i = 0; i < kNumMovable; i++) {
            ctx.destroyEntity(sim.movable[i]);
        }
        generateWorld(ctx);
    }
    ctx.singleton<BodyCount>().count = kNumPhysicsEntities;
    ctx.singleton<JointCount>().count = sim.numJoints;
}

inline void selfObsSystem(Engine &, Position &pos, Rotation &rot, Team &team,
                          Grip &grip, StepsRemaining &steps, SelfObs &obs)
{
    obs.x = pos.x * (1.f / kHalf);
    obs.y = pos.y * (1.f / kHalf);
    obs.z = pos.z;
    obs.qw = rot.w;
    obs.qx = rot.x;
    obs.qy = rot.y;
    obs.qz = rot.z;
    obs.isHider = (float)team.isHider;
    obs.holding = (float)grip.holding;
    obs.stepsRemaining = (float)steps.t;
}

// offsets to the other agents + a line-of-sight ray to each of them
inline void otherObsSystem(Engine &ctx, Entity e, Position &pos, OtherObs &obs)
{
    Sim &sim = ctx.data();
    broadphase::BVH &bvh = ctx.singleton<broadphase::BVH>();
    Vector3 origin = pos;
    origin.z += 0.25f;
#ifdef MADRONA_GPU_MODE
    // GPU backend: 8 threads per agent, one other agent each (three idle)
    const int32_t first = (int32_t)(threadIdx.x % 8);
    const int32_t last = first + 1 < kNumAgents - 1 ? first + 1 : kNumAgents - 1;
#else
    const int32_t first = 0;
    const int32_t last = kNumAgents - 1;
#endif
    int32_t self_idx = 0;
    for (int32_t i = 0; i < kNumAgents; i++) {
        if (sim.agents[i] == e) {
            self_idx = i;
        }
    }
    for (int32_t k = first; k < last; k++) {
        int32_t j = k < self_idx ? k : k + 1;
        Entity other = sim.agents[j];
        Vector3 other_pos = ctx.get<Position>(other);
        Vector3 target = other_pos;
        target.z += 0.25f;
        Vector3 to = target - origin;
        float visible = 0.f;
        float len2 = to.length2();
        if (len2 > 1.f) {
            Vector3 dir = to / sqrtf(len2);
            float hit_t;
            Vector3 hit_normal;
            Entity hit = bvh.traceRay(origin + 0.8f * dir, dir, &hit_t, &hit_normal, 100.f);
            if (hit == other) {
                visible = 1.f;
            }
        } else {
            visible = 1.f;
        }
        obs.v[k][0] = (other_pos.x - pos.x) * (1.f / (2.f * kHalf));
        obs.v[k][1] = (other_pos.y - pos.y) * (1.f / (2.f * kHalf));
        obs.v[k][2] = (float)ctx.get<Team>(other).isHider;
        obs.v[k][3] = visible;
    }
}

inline void lidarSystem(Engine &ctx, Entity e, Position &pos, Rotation &rot,
                        Lidar &lidar)
{
    // unit directions at multiples of 22.5 degrees (literals)
    constexpr float c1 = 0.92387953f, s1 = 0.38268343f, d = 0.70710678f;
    const Vector3 dirs[kNumLidar] = {
        { 0, 1, 0 }, { s1, c1, 0 }, { d, d, 0 }, { c1, s1, 0 },
        { 1, 0, 0 }, { c1, -s1, 0 }, { d, -d, 0 }, { s1, -c1, 0 },
        { 0, -1, 0 }, { -s1, -c1, 0 }, { -d, -d, 0 }, { -c1, -s1, 0 },
        { -1, 0, 0 }, { -c1, s1, 0 }, { -d, d, 0 }, { -s1, c1, 0 },
    };
    broadphase::BVH &bvh = ctx.singleton<broadphase::BVH>();
    Quat q = rot;
    Vector3 origin = pos;
    origin.z += 0.25f;
#ifdef MADRONA_GPU_MODE
    const int32_t first_ray = (int32_t)(threadIdx.x % kNumLidar);
    const int32_t last_ray = first_ray + 1;
#else
    const int32_t first_ray = 0;
    const int32_t last_ray = kNumLidar;
#endif
    for (int32_t i = first_ray; i < last_ray; i++) {
        Vector3 ray_dir = q.rotateVec(dirs[i]);
        Vector3 ray_o = origin + 0.8f * ray_dir;
        float hit_t;
        Vector3 hit_normal;
        Entity hit = bvh.traceRay(ray_o, ray_dir, &hit_t, &hit_normal, 200.f);
        if (hit == Entity::none() || hit == e) {
            lidar.samples[i] = LidarSample { 0.f, 0.f };
        } else {
            EntityType type = ctx.get<EntityType>(hit);
            lidar.samples[i] = LidarSample { hit_t, (float)(uint32_t)type };
        }
    }
}

void Sim::setupTasks(TaskGraphManager &mgr, const Config &)
{
    TaskGraphBuilder &builder = mgr.init(TaskGraphID::Step);

    auto move = builder.addToGraph<ParallelForNode<Engine, movementSystem,
        Action, Rotation, ExternalForce, ExternalTorque>>({});

    auto broadphase = PhysicsSystem::setupBroadphaseTasks(builder, {move});
    auto physics = PhysicsSystem::setupPhysicsStepTasks(builder, {broadphase},
                                                        kNumSubsteps);

    auto zero_vel = builder.addToGraph<ParallelForNode<Engine, agentZeroVelSystem,
        Velocity, Action>>({physics});
    auto cleanup = PhysicsSystem::setupCleanupTasks(builder, {zero_vel});

    auto grab = builder.addToGraph<ParallelForNode<Engine, grabSystem,
        JointCount>>({cleanup});
    auto reward = builder.addToGraph<ParallelForNode<Engine, rewardSystem,
        Position, Team, Reward, StepsRemaining, Done>>({grab});
    auto reset = builder.addToGraph<ParallelForNode<Engine, resetSystem,
        WorldReset>>({reward});

    auto compact = builder.addToGraph<CompactArchetypeNode<PhysicsEntity>>({reset});
#ifdef MADRONA_GPU_MODE
    auto recycle = builder.addToGraph<RecycleEntitiesNode>({compact});
    auto post_reset = recycle;
#else
    auto post_reset = compact;
#endif
    auto post_bvh = PhysicsSystem::setupBroadphaseTasks(builder, {post_reset});

    // three independent observation systems (they only read the world): parallel
    // branches of the step graph on the GPU
    builder.addToGraph<ParallelForNode<Engine, selfObsSystem,
        Position, Rotation, Team, Grip, StepsRemaining, SelfObs>>({post_bvh});
#ifdef MADRONA_GPU_MODE
    builder.addToGraph<CustomParallelForNode<Engine, otherObsSystem, 8, 1,
        Entity, Position, OtherObs>>({post_bvh});
    builder.addToGraph<CustomParallelForNode<Engine, lidarSystem, kNumLidar, 1,
        Entity, Position, Rotation, Lidar>>({post_bvh});
#else
    builder.addToGraph<ParallelForNode<Engine, otherObsSystem,
        Entity, Position, OtherObs>>({post_bvh});
    builder.addToGraph<ParallelForNode<Engine, lidarSystem,
        Entity, Position, Rotation, Lidar>>({post_bvh});
#endif
}

Sim::Sim(Engine &ctx, const Config &cfg, const WorldInit &init)
    : WorldBase(ctx),
      rng(init.seed),
      episodeLen(cfg.episodeLen),
      episode(0),
      numJoints(0)
{
    PhysicsSystem::init(ctx, cfg.objMgr, kDeltaT, kNumSubsteps,
                        -9.8f * math::up, kMaxBodies);

    plane = ctx.makeEntity<PhysicsEntity>();
    for (int32_t i = 0; i < kNumBorderWalls; i++) {
        borders[i] = ctx.makeEntity<PhysicsEntity>();
    }
    for (int32_t i = 0; i < kNumRoomWalls; i++) {
        roomWalls[i] = ctx.makeEntity<PhysicsEntity>();
    }
    for (int32_t i = 0; i < kNumPillars; i++) {
        pillars[i] = ctx.makeEntity<PhysicsEntity>();
    }
    for (int32_t i = 0; i < kNumAgents; i++) {
        agents[i] = ctx.makeEntity<Agent>();
        ctx.get<Action>(agents[i]) = Action { 0, 0, 2, 0 };
        ctx.get<Done>(agents[i]).v = 0;
        ctx.get<Reward>(agents[i]).v = 0.f;
        ctx.get<Team>(agents[i]).isHider = i < kNumHiders ? 1 : 0;
        ctx.get<SelfObs>(agents[i]) = SelfObs {};
        ctx.get<OtherObs>(agents[i]) = OtherObs {};
        ctx.get<Lidar>(agents[i]) = Lidar {};
    }
    ctx.singleton<WorldReset>().reset = 0;
    generateWorld(ctx);
    ctx.singleton<BodyCount>().count = kNumPhysicsEntities;
    ctx.singleton<JointCount>().count = numJoints;
}

}

#ifdef MADRONA_GPU_MODE
MADRONA_BUILD_MWGPU_ENTRY(arena::Engine, arena::Sim, arena::Config, arena::WorldInit);
#endif
