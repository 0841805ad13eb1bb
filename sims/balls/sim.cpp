#include "sim.hpp"

#ifdef MADRONA_GPU_MODE
#include <madrona/mw_gpu_entry.hpp>
#endif

using namespace madrona;
using namespace madrona::math;
using namespace madrona::phys;

namespace balls {

constexpr float kDeltaT = 0.04f;
constexpr CountT kNumSubsteps = 4;
constexpr float kWallThick = 0.5f;
constexpr float kWallHeight = 8.f;     // tall: the volley stays inside the pen
constexpr float kBallScale = 1.2f;     // radius 0.6

void Sim::registerTypes(ECSRegistry &registry, const Config &)
{
    base::registerTypes(registry);
    PhysicsSystem::registerTypes(registry);

    registry.registerArchetype<Body>();

    registry.exportColumn<Body, Position>((uint32_t)ExportID::BodyPos);
    registry.exportColumn<Body, Rotation>((uint32_t)ExportID::BodyRot);
    registry.exportColumn<Body, Velocity>((uint32_t)ExportID::BodyVel);
    registry.exportColumn<Body, Entity>((uint32_t)ExportID::BodyEntity);
}

static inline Entity makeBody(Engine &ctx, Vector3 pos, Quat rot, Diag3x3 scale,
                              SimObject obj, ResponseType resp, Vector3 lin_vel)
{
    Entity e = ctx.makeEntity<Body>();
    ObjectID obj_id { (int32_t)obj };
    ctx.get<Position>(e) = pos;
    ctx.get<Rotation>(e) = rot;
    ctx.get<Scale>(e) = scale;
    ctx.get<ObjectID>(e) = obj_id;
    ctx.get<ResponseType>(e) = resp;
    ctx.get<Velocity>(e) = Velocity { lin_vel, Vector3::zero() };
    ctx.get<ExternalForce>(e) = Vector3::zero();
    ctx.get<ExternalTorque>(e) = Vector3::zero();
    ctx.get<broadphase::LeafID>(e) = PhysicsSystem::registerEntity(ctx, e, obj_id);
    return e;
}

void Sim::setupTasks(TaskGraphManager &mgr, const Config &)
{
    TaskGraphBuilder &builder = mgr.init(TaskGraphID::Step);
    auto broadphase = PhysicsSystem::setupBroadphaseTasks(builder, {});
    auto physics = PhysicsSystem::setupPhysicsStepTasks(builder, {broadphase},
                                                        kNumSubsteps);
    PhysicsSystem::setupCleanupTasks(builder, {physics});
}

Sim::Sim(Engine &ctx, const Config &cfg, const WorldInit &init)
    : WorldBase(ctx),
      rng(init.seed)
{
    PhysicsSystem::init(ctx, cfg.objMgr, kDeltaT, kNumSubsteps,
                        -9.8f * math::up, kMaxBodies);
    PhysicsSystem::reset(ctx);

    const Quat upright { 1, 0, 0, 0 };
    const Vector3 still = Vector3::zero();

    makeBody(ctx, Vector3 { 0, 0, 0 }, upright, Diag3x3 { 1, 1, 1 }, SimObject::Plane,
             ResponseType::Static, still);

    // four walls around the pen (static hulls the balls run into)
    const float reach = kPen + kWallThick;
    const float mid = kPen + 0.5f * kWallThick;
    makeBody(ctx, Vector3 { 0, -mid, 0.5f * kWallHeight }, upright,
             Diag3x3 { 2.f * reach, kWallThick, kWallHeight }, SimObject::Wall,
             ResponseType::Static, still);
    makeBody(ctx, Vector3 { 0, mid, 0.5f * kWallHeight }, upright,
             Diag3x3 { 2.f * reach, kWallThick, kWallHeight }, SimObject::Wall,
             ResponseType::Static, still);
    makeBody(ctx, Vector3 { -mid, 0, 0.5f * kWallHeight }, upright,
             Diag3x3 { kWallThick, 2.f * kPen, kWallHeight }, SimObject::Wall,
             ResponseType::Static, still);
    makeBody(ctx, Vector3 { mid, 0, 0.5f * kWallHeight }, upright,
             Diag3x3 { kWallThick, 2.f * kPen, kWallHeight }, SimObject::Wall,
             ResponseType::Static, still);

    // loose cubes resting on the floor, spread over the pen
    for (int32_t i = 0; i < kNumCubes; i++) {
        float span = 2.f * (kPen - 1.5f);
        float x = -0.5f * span + span * ((float)i + 0.5f) / (float)kNumCubes +
            (rng.sampleUniform() - 0.5f);
        float y = (rng.sampleUniform() - 0.5f) * kPen;
        makeBody(ctx, Vector3 { x, y, 0.75f }, upright, Diag3x3 { 1.5f, 1.5f, 1.5f },
                 SimObject::Cube, ResponseType::Dynamic, still);
    }

    // balls: dropped from different heights, thrown at the walls and the cubes
    for (int32_t i = 0; i < kNumBalls; i++) {
        float x = (rng.sampleUniform() - 0.5f) * 1.5f * kPen;
        float y = (rng.sampleUniform() - 0.5f) * 1.5f * kPen;
        float z = 0.6f + 0.9f * (float)(i % 4) + rng.sampleUniform();
        Vector3 vel { (rng.sampleUniform() - 0.5f) * 8.f,
                      (rng.sampleUniform() - 0.5f) * 8.f,
                      (rng.sampleUniform() - 0.5f) * 2.f };
        makeBody(ctx, Vector3 { x, y, z }, upright,
                 Diag3x3 { kBallScale, kBallScale, kBallScale }, SimObject::Ball,
                 ResponseType::Dynamic, vel);
    }
}

}

#ifdef MADRONA_GPU_MODE
MADRONA_BUILD_MWGPU_ENTRY(balls::Engine, balls::Sim, balls::Config, balls::WorldInit);
#endif
