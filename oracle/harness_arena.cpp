// TEST INFRASTRUCTURE ONLY: sims/arena (Hide&Seek-class fixture: hinge + fixed
// joints, wedge / hexagonal-prism hulls, entity churn) on the reference CPU
// backend (TaskGraphExecutor + the reference's own broadphase / narrowphase / XPBD).
#include <madrona/mw_cpu.hpp>
#include "../sims/arena/sim.hpp"
#include "harness.hpp"

using namespace arena;

int main(int argc, char **argv)
{
    oracle::Args args = oracle::parseArgs(argc, argv);
    Config cfg { (madrona::phys::ObjectManager *)oracle::loadObjectsBlob(oracle::objectsPathArg(argc, argv)),
                 (uint32_t)(args.extra[0] ? args.extra[0] : 100), 0 };
    std::vector<WorldInit> inits(args.numWorlds);
    for (int64_t i = 0; i < args.numWorlds; i++) inits[i].seed = (uint32_t)(args.extra[1] + i);

    using Exec = madrona::TaskGraphExecutor<Engine, Sim, Config, WorldInit>;
    Exec exec({
        .numWorlds = (uint32_t)args.numWorlds,
        .numExportedBuffers = (uint32_t)ExportID::NumExports,
        .numWorkers = (uint32_t)args.numWorkers,
    }, cfg, inits.data(), (madrona::CountT)TaskGraphID::NumTaskGraphs);

    const size_t W = (size_t)args.numWorlds;
    const size_t A = W * kNumAgents;
    auto total_bodies = [&exec, W]() {
        const int32_t *counts = (const int32_t *)exec.getExported((int)ExportID::BodyCount);
        size_t n = 0;
        for (size_t i = 0; i < W; i++) n += (size_t)counts[i];
        return n;
    };
    return oracle::runTrace(exec, args,
        { { (int)ExportID::Reset, 4 }, { (int)ExportID::Action, sizeof(Action) * kNumAgents } },
        { { (int)ExportID::Reward, [=] { return A * 4; } },
          { (int)ExportID::Done, [=] { return A * 4; } },
          { (int)ExportID::SelfObs, [=] { return A * sizeof(SelfObs); } },
          { (int)ExportID::OtherObs, [=] { return A * sizeof(OtherObs); } },
          { (int)ExportID::Lidar, [=] { return A * sizeof(Lidar); } },
          { (int)ExportID::AgentPos, [=] { return A * 12; } },
          { (int)ExportID::AgentRot, [=] { return A * 16; } },
          { (int)ExportID::BodyCount, [=] { return W * 4; } },
          { (int)ExportID::JointCount, [=] { return W * 4; } },
          { (int)ExportID::BodyPos, [=] { return total_bodies() * 12; } },
          { (int)ExportID::BodyRot, [=] { return total_bodies() * 16; } },
          { (int)ExportID::BodyEntity, [=] { return total_bodies() * 8; } },
          { (int)ExportID::BodyVel, [=] { return total_bodies() * 24; } } });
}
